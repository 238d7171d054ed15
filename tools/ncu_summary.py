"""Summarise an `ncu --set full` report (one kernel launch) as JSON: python tools/ncu_summary.py report.ncu-rep [note]
Reads the report with `ncu -i ... --page raw --csv` (B200_PROFILING.md recipe); keeps the numbers DESIGN.md / bench.py quote."""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_rate_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_sectors_read_from_sm",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, val = rows[0], rows[1], rows[-1]
    d = {"report": rep.split("/")[-1], "kernel": val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None}
    for i, name in enumerate(hdr):
        if name in KEYS:
            try:
                v = float(val[i])
            except ValueError:
                v = val[i]
            d[KEYS[name]] = v
            d[KEYS[name] + "_unit"] = units[i]
    def to_bytes(k):
        u = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(d.get(k + "_unit"), 1.0)
        return d.get(k, 0.0) * u
    d["dram_bytes_total"] = to_bytes("dram_bytes_read") + to_bytes("dram_bytes_write")
    if len(sys.argv) > 2:
        d["note"] = sys.argv[2]
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
