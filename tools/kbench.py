"""Row-kernel micro-benchmark: achieved GB/s of every row-kernel mode on the fine-level matrices of the
BASELINE configs.  usage: python tools/kbench.py [poisson:CELLS | elasticity:CELLS | stokes:CELLS ...] [modes=a,b] [reps=N] [k=v options]
Reports, per mode: time, ALGORITHMIC GB/s (12 B per non-zero + per-row vector traffic) and the GB/s
of the bytes the block-SELL format actually stores (8 + 4/bs^2 B per stored entry incl. padding)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsb200 as gsb
from gsb200 import synth

PER_ROW = {"spmv": 20, "spmv_dot": 28, "residual": 28, "sweep": 44, "spmv_add": 36}


def peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    specs = [a for a in sys.argv[1:] if ":" in a] or ["poisson:128"]
    opts = [a.split("=", 1) for a in sys.argv[1:] if "=" in a and ":" not in a]
    modes = ("spmv", "residual", "sweep", "spmv_dot")
    reps = 30
    for k, v in list(opts):  # tool arguments (not library options): modes=sweep,spmv reps=3
        if k == "modes":
            modes = tuple(v.split(","))
            opts.remove([k, v])
        elif k == "reps":
            reps = int(v)
            opts.remove([k, v])
    ctx = gsb.Context()
    for k, v in opts:
        ctx.set_option(k, v)
    pk = peak()
    for spec in specs:
        kind, c = spec.split(":")
        c = int(c)
        if kind == "poisson":
            lp = synth.make_level_part((c,) * 3, (1, 1, 1), 0)
            rp, col, val, b = synth.poisson_rows(lp)
            n = lp.n_own
        elif kind == "elasticity":
            rp, col, val, b, n = synth.elasticity_rows((c,) * 3)
        elif kind == "stokes":
            st = synth.stokes_cavity_host((c, c), nlevels=1)
            rp, col, val = st["A"]
            n = rp.shape[0] - 1
        else:
            raise SystemExit("unknown problem " + kind)
        A = gsb.SparseMatrix(ctx, n, n, 0, rp, col, val)
        nnz = int(rp[-1])
        fmt = A.format()
        out = {"problem": spec, "rows": n, "nnz": nnz, "format": fmt, "opts": dict(opts)}
        for mode in modes:
            ms = A.bench_rows(mode, reps)
            gbs = (12 * nnz + PER_ROW[mode] * n) / (ms * 1e-3) / 1e9
            fgbs = (fmt["bytes_per_pass"] + PER_ROW[mode] * n - 4 * n) / (ms * 1e-3) / 1e9
            out[mode] = {"us": round(ms * 1e3, 1), "GBps_algorithmic": round(gbs, 0), "frac": round(gbs / pk, 3),
                         "GBps_format_bytes": round(fgbs, 0), "frac_format": round(fgbs / pk, 3)}
        print(json.dumps(out), flush=True)
        del A


if __name__ == "__main__":
    main()
