#!/usr/bin/env python
"""bench.py -- GMG-preconditioned CG solve of the 3D Q1 Poisson problem (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells-per-gpu C]
  torchrun ... bench.py --gpus N ...      (one rank per GPU, NCCL)

A "step" is one complete solve!(x, ns, b): CGSolver preconditioned by GMGLinearSolver (V-cycle,
RichardsonSmoother(JacobiLinearSolver(),10,2/3) pre/post, LU coarse solve, mode=:preconditioner,
maxiter=1) from x0 = 0 to a relative residual of 1e-8 -- config C2 (128^3 cells, 4 levels) at
N=1; for N>1 the C3-style weak scaling series (256^3 cells per GPU; 512^3 = 133M DOFs on 8).
Set-up (matrix upload, inverse diagonals, coarse inverse) is outside the timed region, as in the
reference's drivers (tic!/toc! around solve! only, test/LinearSolvers/GMGTests.jl:127-129).

Prints ONE JSON line (contract in the task statement).  `value` = DOFs solved to 1e-8 per second
(whole job); `ms_per_step` = the solve time; `e2e` = same through the host-buffer C-ABI call
(gsb_solve_host: H2D of b and x0, solve, D2H of x inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's banner / debug output on stderr
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

# part grids (px,py,pz): split the slowest-varying directions first so that halo faces are contiguous planes
PARTS = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}
RTOL, ATOL, MAXITER = 1e-8, 1e-14, 100


def n_levels(cells_per_gpu: int, nranks: int) -> int:
    if nranks == 1 and cells_per_gpu == 128:
        return 4  # C2: 128/64/32/16
    lv, c = 1, cells_per_gpu
    while c % 2 == 0 and c // 2 >= 8:  # coarsest level keeps 8 cells per part and direction
        c //= 2
        lv += 1
    return lv


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        with open(self.path) as f:
            for line in f:
                c = [t.strip() for t in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def build_solver(gsb, dh, nlev):
    sm = gsb.Fill(gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), 10, 2.0 / 3.0), nlev - 1)
    gmg = gsb.GMGLinearSolver(dh.A, dh.P, dh.R, pre_smoothers=sm, post_smoothers=sm, coarsest_solver=gsb.LUSolver(),
                              maxiter=1, mode="preconditioner", cycle_type="v_cycle")
    solver = gsb.CGSolver(gmg, maxiter=MAXITER, atol=ATOL, rtol=RTOL)
    ns = gsb.numerical_setup(gsb.symbolic_setup(solver, dh.A[0]), dh.A[0])
    return solver, ns


def oracle_solver(hh, nlev, maxiter=MAXITER):
    """the CPU restatement of the same solver stack on the same assembled system (oracle/)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import solvers as OS
    from util import oracle_hierarchy

    mats, P, R = oracle_hierarchy(hh)
    sm = [OS.RichardsonSmoother(OS.JacobiLinearSolver(), 10, 2.0 / 3.0)] * (nlev - 1)
    gmg = OS.GMGLinearSolver(mats, P, R, pre_smoothers=sm, post_smoothers=sm, coarsest_solver=OS.LUSolver(), maxiter=1)
    s = OS.CGSolver(gmg, maxiter=maxiter, atol=ATOL, rtol=RTOL)
    ns = OS.numerical_setup(OS.symbolic_setup(s, mats[0]), mats[0])
    return OS, s, ns, mats


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path.  The Julia package cannot run on this box (no
    julia), so this times the oracle port of the same algorithm (kind "port") with all host threads."""
    if rank != 0:
        return
    import gsb200  # noqa: F401  (host-side generator only; no device call is made on this arm)
    from gsb200 import synth
    from oracle import linalg as ola

    cells = args.cells_per_gpu or 128
    nlev = n_levels(cells, 1)
    hh = synth.poisson_hierarchy_host((cells,) * 3, nlev)
    threads = ola.set_threaded(True)
    OS, s, ns, mats = oracle_solver(hh, nlev)
    n = mats[0].shape[0]
    times = []
    for it in range(args.warmup + args.steps):
        x = np.zeros(n)
        t0 = time.perf_counter()
        OS.solve_(x, ns, hh.b)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = n / (ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": "GMG-PCG solve to 1e-8 rtol (3D Poisson Q1)", "value": val, "unit": "MDOF/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "solve_time_ms": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "iterations": int(s.log.num_iters),
        "config": {"workload": f"C2: CG+GMG(4 lev, V(10,10) Jacobi-Richardson 2/3, LU coarse) 3D Poisson Q1 {cells}^3 cells, "
                               f"{n} DOFs, rtol 1e-8; CPU restatement of the reference (oracle/), {threads} OpenMP threads"},
        "cpu_baseline": {"value": val, "unit": "MDOF/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full solves ({s.log.num_iters} PCG iterations each) of the same {cells}^3 system"},
        "e2e": {"value": val, "unit": "MDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gsb200")
    ap.add_argument("--cells-per-gpu", type=int, default=int(os.environ.get("GSB_BENCH_CELLS", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scaling-base", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    assert args.gpus in PARTS, "--gpus must be 1, 2, 4 or 8"
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    import gsb200 as gsb
    from gsb200 import synth

    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ids = [gsb.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]
    ctx = gsb.Context(device=local_rank, nranks=world, rank=rank, nccl_id=nccl_id)

    def barrier():
        ctx.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(cells, steps, warmup, with_e2e=True, with_profile=True):
        parts = PARTS[world]
        nlev = n_levels(cells, world)
        ncell = tuple(cells * p for p in parts)
        t0 = time.perf_counter()
        hh = synth.poisson_hierarchy_host(ncell, nlev, parts=parts, rank=rank, lengths=tuple(float(p) for p in parts))
        t_gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        dh = synth.upload_hierarchy(ctx, hh)
        solver, ns = build_solver(gsb, dh, nlev)
        ctx.synchronize()
        t_setup = time.perf_counter() - t0
        A = dh.A[0]
        n_own = hh.levels[0].n_own
        n_glob = int(np.prod([c - 1 for c in ncell]))
        x, b = gsb.allocate_in_domain(A), gsb.allocate_in_domain(A)
        b.set(hh.b)
        # ---- device-resident timing: inputs already in HBM
        for _ in range(warmup):
            x.fill(0.0)
            gsb.solve_(x, ns, b)
        sampler = ClockSampler(local_rank)
        barrier()
        if rank == 0:
            sampler.start()
        l0 = ctx.launch_count()
        dev_ms = 0.0
        wall0 = time.perf_counter()
        for _ in range(steps):
            x.fill(0.0)
            ctx.timer_start()
            gsb.solve_(x, ns, b)
            dev_ms += ctx.timer_stop()
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - wall0)
        clocks = sampler.stop() if rank == 0 else None
        launches = ctx.launch_count() - l0
        ms = max_over_ranks(dev_ms / steps)
        iters = solver.log.num_iters
        hist = solver.log.history()
        out = dict(cells=cells, ncell=ncell, nlev=nlev, n_glob=n_glob, n_own=n_own, ms=ms, wall_ms=wall_ms / steps, iters=iters,
                   hist=hist, launches=launches // steps, clocks=clocks, t_gen=t_gen, t_setup=t_setup, hh=hh,
                   nnz=[int(a[0][-1]) for a in hh.A], rows=[lp.n_own for lp in hh.levels])
        # ---- end-to-end through the host-buffer C-ABI entry point (pinned host memory)
        if with_e2e:
            xh = torch.zeros(n_own, dtype=torch.float64).pin_memory().numpy()
            bh = torch.from_numpy(hh.b.copy()).pin_memory().numpy()
            for _ in range(2):
                xh[:] = 0.0
                gsb.solve_(xh, ns, bh)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                xh[:] = 0.0
                gsb.solve_(xh, ns, bh)
            barrier()
            out["e2e_ms"] = max_over_ranks(1e3 * (time.perf_counter() - t0) / steps)
            out["h2d"], out["d2h"] = 2 * 8 * n_own, 8 * n_own
            out["x_err"] = float(np.max(np.abs(xh - synth.exact_solution(hh.levels[0]))))
        # ---- per-kernel durations inside a real solve (CUDA events around every row-kernel launch)
        if with_profile:
            x.fill(0.0)
            ctx.profile_start()
            gsb.solve_(x, ns, b)
            out["prof"] = ctx.profile_stop()
        return out

    cells = args.cells_per_gpu or (128 if world == 1 else 256)
    res = measure(cells, args.steps, args.warmup)
    peak, peak_src = peaks()

    # roofline of the dominant kernel: the fused Jacobi-Richardson sweep on the finest level
    def algo_bytes(mode, nrows, nnz, impl=""):
        per_row = {"spmv": 20, "spmv_dot": 28, "residual": 28, "sweep": 44, "spmv_add": 36, "sweeps_pipelined": 44}[mode]
        nsweeps = int(impl.rsplit("S", 1)[1]) if mode == "sweeps_pipelined" else 1  # S sweeps per launch
        return nsweeps * (12 * nnz + per_row * nrows)

    kernels = []
    for p in res.get("prof", []):
        t = p["total_ms"] / p["count"] * 1e-3
        kernels.append({"kernel": p["mode"], "impl": p["impl"], "rows": p["nrows"], "nnz": p["nnz"],
                        "launches": p["count"], "avg_us": round(t * 1e6, 2),
                        "GBps": round(algo_bytes(p["mode"], p["nrows"], p["nnz"], p["impl"]) / t / 1e9, 1),
                        "share_of_solve": round(p["total_ms"] / res["ms"], 4)})
    kernels.sort(key=lambda k: -k["share_of_solve"])
    top = next((k for k in kernels if k["kernel"] in ("sweep", "sweeps_pipelined") and k["rows"] == res["rows"][0]), kernels[0] if kernels else None)
    roofline = None
    if top:
        pipelined = top["kernel"] == "sweeps_pipelined"
        nsw = int(top["impl"].rsplit("S", 1)[1]) if pipelined else 1
        roofline = {"bound": "hbm", "kernel": ("sell_pipe_kernel level 1 (%d fused Jacobi-Richardson sweeps per launch, matrix re-read from L2)" % nsw) if pipelined
                    else "csr_sell_kernel<sweep> level 1 (fused Jacobi-Richardson sweep, SELL-32)",
                    "achieved": top["GBps"], "peak": peak, "unit": "GB/s", "frac": round(top["GBps"] / peak, 4),
                    "frac_of_nominal_8TBs": round(top["GBps"] / 8000.0, 4), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": algo_bytes(top["kernel"], top["rows"], top["nnz"], top["impl"]),
                    "sweeps_per_launch": nsw,
                    "avg_launch_us": top["avg_us"], "traffic": None}
        tfile = os.path.join(ROOT, "profiles", "traffic_r01.json")
        if os.path.exists(tfile):
            try:
                with open(tfile) as f:
                    roofline["traffic"] = json.load(f).get("sweep_level1_dram_bytes_per_launch")
            except Exception:
                pass

    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    n_glob = res["n_glob"]
    value = n_glob / (res["ms"] * 1e-3) / 1e6
    line = {
        "metric": "GMG-PCG solve to 1e-8 rtol (3D Poisson Q1)", "value": round(value, 3), "unit": "MDOF/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(res["ms"], 4),
        "solve_time_ms": round(res["ms"], 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "iterations": res["iters"], "final_rel_residual": float(res["hist"][-1] / res["hist"][0]),
        "config": {
            "workload": (f"{'C2' if (world == 1 and cells == 128) else 'C3-style weak scaling'}: CGSolver(GMGLinearSolver V-cycle, "
                         f"{res['nlev']} levels, RichardsonSmoother(Jacobi,10,2/3) pre+post, LU coarse, maxiter=1) on 3D Poisson Q1, "
                         f"{'x'.join(str(c) for c in res['ncell'])} cells on [0,{PARTS[world][0]}]x[0,{PARTS[world][1]}]x[0,{PARTS[world][2]}] (cubic cells) = {n_glob} DOFs, "
                         f"{cells}^3 cells per GPU, rtol 1e-8, x0=0"),
            "partition": "x".join(str(p) for p in PARTS[world]), "levels_rows_rank0": res["rows"], "levels_nnz_rank0": res["nnz"],
            "l2_policy": "inputs larger than L2 (fine-level CSR matrix %.0f MB >> 126 MB L2); coarse levels are L2-resident by construction" % (12 * res["nnz"][0] / 1e6),
            "setup_s": round(res["t_setup"], 3), "host_generation_s": round(res["t_gen"], 3),
        },
        "gpu_launches": res["launches"],
        "clocks": res["clocks"],
        "e2e": {"value": round(n_glob / (res["e2e_ms"] * 1e-3) / 1e6, 3), "unit": "MDOF/s", "ms_per_step": round(res["e2e_ms"], 4),
                "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"], "max_abs_error_vs_exact": res["x_err"]},
        "roofline": roofline,
        "kernels": kernels[:12],
    }
    # CPU baseline (oracle port) on a bounded sample, rank 0, N=1 only
    if world == 1 and not args.no_cpu_baseline:
        from oracle import linalg as ola

        threads = ola.set_threaded(True)
        sample_iters = 2
        OS, s, ons, mats = oracle_solver(res["hh"], res["nlev"], maxiter=sample_iters)
        xo = np.zeros(mats[0].shape[0])
        t0 = time.perf_counter()
        OS.solve_(xo, ons, res["hh"].b)
        dt = time.perf_counter() - t0
        est = dt * res["iters"] / sample_iters
        ola.set_threaded(False)
        line["cpu_baseline"] = {"value": round(n_glob / est / 1e6, 4), "unit": "MDOF/s", "cores": threads, "kind": "port",
                                "est_solve_time_ms": round(est * 1e3, 1),
                                "sample": f"{sample_iters} PCG iterations (of {res['iters']}) of the same {cells}^3 system on the oracle port "
                                          f"with {threads} OpenMP threads, scaled by {res['iters']}/{sample_iters}; residuals after the sample "
                                          f"agree with the GPU history to {abs(s.log.history()[-1] - res['hist'][sample_iters]) / res['hist'][0]:.1e}"}
    # apples-to-apples denominator for the weak-scaling series (same per-GPU unit as N>1)
    if world == 1 and cells == 128 and not args.no_scaling_base:
        del res
        base = measure(256, max(2, args.steps // 3), 3, with_e2e=False, with_profile=False)
        line["weak_scaling_base"] = {"cells_per_gpu": 256, "dofs": base["n_glob"], "ms_per_step": round(base["ms"], 3),
                                     "value": round(base["n_glob"] / (base["ms"] * 1e-3) / 1e6, 3), "unit": "MDOF/s",
                                     "iterations": base["iters"], "levels": base["nlev"],
                                     "note": "N=1 run of the per-GPU unit the N>1 series uses (256^3 cells per GPU)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


if __name__ == "__main__":
    main()
