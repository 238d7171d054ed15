#!/usr/bin/env python
"""bench.py -- solve-phase benchmark of the BASELINE.json configurations on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c4|c5] [--impl reference] [--cells-per-gpu C]
  torchrun ... bench.py --gpus N ...      (one rank per GPU, NCCL)

A "step" is one complete solve!(x, ns, b) from x0 = 0 to a relative residual of 1e-8:
  c2 (default; the config the BASELINE metric is quoted on): CGSolver preconditioned by GMGLinearSolver (V-cycle,
     RichardsonSmoother(JacobiLinearSolver(),10,2/3) pre/post, LU coarse solve, mode=:preconditioner, maxiter=1) on
     3D Poisson Q1, 128^3 cells, 4 levels at N=1; for N>1 the C3-style weak-scaling series (256^3 cells per GPU).
  c4: FGMRESSolver(30, GMG) on 3D linear elasticity Q2 (vector-valued, 3x3-block sparsity), 64^3 cells.
  c5: GMRESSolver(30; Pr=BlockTriangularSolver([GMG velocity block, CG-Jacobi on the pressure mass])) on the 2D
      Stokes lid-driven cavity, Q2-P1disc, 256^2 cells.
Set-up (matrix upload, format conversion, inverse diagonals, coarse inverse) is outside the timed region, as in the
reference's drivers (tic!/toc! around solve! only, test/LinearSolvers/GMGTests.jl:127-129).

Prints ONE JSON line (contract in the task statement).  `value` = DOFs solved to 1e-8 per second (whole job);
`ms_per_step` = the solve time; `e2e` = same through the host-buffer C-ABI call (gsb_solve_host: H2D of b, solve,
D2H of x inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
# stdout carries exactly one JSON line: keep NCCL's banner / debug output on stderr
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

# part grids (px,py,pz): only the slowest-varying directions are split, so that halo faces are contiguous planes AND
# every mesh line (x direction) stays on one part: a split in x puts a ghost column at the end of every line, which
# breaks the diagonal alignment of one slice in four/eight (measured: 105 ms instead of ~88 ms on 8 GPUs with 2x2x2)
PARTS = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (1, 2, 4)}
RTOL, ATOL, MAXITER = 1e-8, 1e-14, 100
PER_ROW = {"spmv": 20, "spmv_dot": 28, "residual": 28, "sweep": 44, "spmv_add": 36}  # SURVEY.md 8d vector bytes per row


def host_threads() -> int:
    """threads for the CPU arms: all host cores, set explicitly (torchrun exports OMP_NUM_THREADS=1)"""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def n_levels(cells: int, min_coarse: int = 8) -> int:
    lv, c = 1, cells
    while c % 2 == 0 and c // 2 >= min_coarse:
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
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
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


# ------------------------------------------------------------------------------------------------ solver stacks
# the same constructor sequence builds the device solver (S = gsb200) and the CPU restatement (S = oracle.solvers)


def gmg_stack(S, mats, P, R, nlev, redist=None, **kw):
    sm = [S.RichardsonSmoother(S.JacobiLinearSolver(), 10, 2.0 / 3.0)] * (nlev - 1)
    if redist is not None:  # device side only: coarse levels on fewer parts
        kw["redist"] = redist
    return S.GMGLinearSolver(mats, P, R, pre_smoothers=sm, post_smoothers=sm, coarsest_solver=S.LUSolver(), **kw)


def c2_solver(S, mats, P, R, nlev, maxiter=MAXITER, redist=None):
    gmg = gmg_stack(S, mats, P, R, nlev, redist=redist, maxiter=1, mode="preconditioner", cycle_type="v_cycle")
    return S.CGSolver(gmg, maxiter=maxiter, atol=ATOL, rtol=RTOL)


def c4_solver(S, mats, P, R, nlev, maxiter=MAXITER, redist=None):
    gmg = gmg_stack(S, mats, P, R, nlev, redist=redist, maxiter=1, mode="preconditioner", cycle_type="v_cycle")
    return S.FGMRESSolver(30, gmg, maxiter=maxiter, atol=ATOL, rtol=RTOL)


def c5_precond(S, mats, P, R, nlev, Mp):
    # joss_paper/demo.jl:64-85: GMG(maxiter=4, mode=:solver) on the velocity block, CG(Jacobi; maxiter=20, rtol=1e-6)
    # on the pressure mass matrix (BiformBlock), upper block-triangular
    gmg = gmg_stack(S, mats, P, R, nlev, maxiter=4, mode="solver")
    cgp = S.CGSolver(S.JacobiLinearSolver(), maxiter=20, atol=1e-14, rtol=1e-6)
    return S.BlockTriangularSolver([gmg, cgp], half="upper", diag_mats=[None, Mp])


def oracle_mats(hh_A, hh_P, hh_R, ns_own):
    from oracle import linalg as ola
    from util import host_to_scipy

    mats = [ola.CSR(host_to_scipy(a, n)) for a, n in zip(hh_A, ns_own)]
    P = [ola.CSR(host_to_scipy(p, ns_own[l + 1])) for l, p in enumerate(hh_P)]
    R = [ola.CSR(host_to_scipy(r, ns_own[l])) for l, r in enumerate(hh_R)]
    return mats, P, R


class Problem:
    """host data of one configuration + builders of the device and the oracle solver"""

    def __init__(self, config, cells, parts=(1, 1, 1), rank=0, serial_of=None, agglomerate_rows=0):
        """serial_of=(px,py,pz): the SERIAL system with the global mesh of that part grid (the oracle side of the
        multi-GPU parity check)"""
        from gsb200 import synth

        self.config, self.cells, self.parts, self.rank = config, cells, parts, rank
        t0 = time.perf_counter()
        if config == "c2":
            grid = serial_of or parts
            nranks = int(np.prod(grid))
            self.nlev = 4 if (nranks == 1 and cells == 128) else n_levels(cells)
            self.ncell = tuple(cells * p for p in grid)
            # levels with at most `agglomerate_rows` rows per part live on rank 0 only (the reference's np_per_level):
            # no halo exchange below that level, one redistribution on the way down and one on the way up
            ppl = None
            if agglomerate_rows > 0 and int(np.prod(parts)) > 1:
                ppl = [tuple(parts) if (cells >> l) ** 3 > agglomerate_rows else (1, 1, 1) for l in range(self.nlev)]
                ppl[0] = tuple(parts)
                for l in range(1, self.nlev):  # monotone: once on one part, always on one part
                    if ppl[l - 1] == (1, 1, 1):
                        ppl[l] = (1, 1, 1)
            self.parts_per_level = ppl
            self.hh = synth.poisson_hierarchy_host(self.ncell, self.nlev, parts=parts, rank=rank, lengths=tuple(float(p) for p in grid),
                                                   parts_per_level=ppl)
            self.n_own = self.hh.levels[0].n_own
            self.n_glob = int(np.prod([c - 1 for c in self.ncell]))
            self.b = self.hh.b
            self.name = "C2" if (nranks == 1 and cells == 128) else "C3-style weak scaling"
            self.metric = "GMG-PCG solve to 1e-8 rtol (3D Poisson Q1)"
            self.what = (f"CGSolver(GMGLinearSolver V-cycle, {self.nlev} levels, RichardsonSmoother(Jacobi,10,2/3) pre+post, LU coarse, "
                         f"maxiter=1) on 3D Poisson Q1, {'x'.join(str(c) for c in self.ncell)} cells on [0,{parts[0]}]x[0,{parts[1]}]x[0,{parts[2]}] "
                         f"(cubic cells) = {self.n_glob} DOFs, {cells}^3 cells per GPU")
        elif config == "c4":
            assert int(np.prod(parts)) == 1, "c4 runs on one GPU"
            self.nlev = n_levels(cells, 4)
            self.ncell = (cells,) * 3
            self.hh = synth.elasticity_hierarchy_host(self.ncell, self.nlev)
            self.n_own = self.n_glob = self.hh.levels[0].n_own
            self.b = self.hh.b
            self.name = "C4"
            self.metric = "FGMRES(30)+GMG solve to 1e-8 rtol (3D linear elasticity Q2)"
            self.what = (f"FGMRESSolver(30, GMGLinearSolver V-cycle, {self.nlev} levels, RichardsonSmoother(Jacobi,10,2/3) pre+post, LU coarse, "
                         f"maxiter=1) on 3D linear elasticity Q2 (vector-valued, node-major 3x3 blocks), {cells}^3 cells, lambda=mu=1 "
                         f"(deviation from the reference's lambda=100 test value, SURVEY 8d), clamped x=0 face, body force (0,0,-1) = {self.n_glob} DOFs")
        elif config == "c5":
            assert int(np.prod(parts)) == 1, "c5 runs on one GPU"
            self.nlev = n_levels(cells, 8)
            self.ncell = (cells, cells)
            self.st = synth.stokes_cavity_host(self.ncell, nlevels=self.nlev)
            self.n_u, self.n_p = self.st["n_u"], self.st["n_p"]
            self.n_own = self.n_glob = self.n_u + self.n_p
            self.b = np.concatenate([self.st["fu"], self.st["fp"]])
            self.name = "C5"
            self.metric = "GMRES(30)+BlockTriangular[GMG,CG-Jacobi] solve to 1e-8 rtol (2D Stokes Q2-P1disc)"
            self.what = (f"GMRESSolver(30; Pr=BlockTriangularSolver([GMGLinearSolver(velocity block, {self.nlev} levels, maxiter=4, mode=:solver), "
                         f"CGSolver(Jacobi; maxiter=20, rtol=1e-6) on the pressure mass matrix])) on the 2D lid-driven cavity, Q2-P1disc, "
                         f"{cells}^2 cells = {self.n_u} velocity + {self.n_p} pressure DOFs (no zero-mean pressure constraint)")
        else:
            raise SystemExit(f"unknown --config {config}")
        self.t_gen = time.perf_counter() - t0

    # ---- device side
    def build_device(self, gsb, ctx):
        from gsb200 import synth

        t0 = time.perf_counter()
        if self.config in ("c2", "c4"):
            dh = synth.upload_hierarchy(ctx, self.hh)
            self.dh = dh
            solver = (c2_solver if self.config == "c2" else c4_solver)(gsb, dh.A, dh.P, dh.R, self.nlev, redist=dh.redist)
            A = dh.A[0]
            self.fine = dh.A[0]
            self.level_rows = [lp.n_own for lp in self.hh.levels]
            self.level_nnz = [int(a[0][-1]) for a in self.hh.A]
        else:
            st = self.st
            mk = lambda t, nc: gsb.SparseMatrix(ctx, t[0].shape[0] - 1, nc, 0, *t)
            ns_own = [m[0].shape[0] - 1 for m in st["mats"]]
            mats = [mk(m, n) for m, n in zip(st["mats"], ns_own)]
            Pm = [mk(p, ns_own[l + 1]) for l, p in enumerate(st["P"])]
            Rm = [mk(r, ns_own[l]) for l, r in enumerate(st["R"])]
            B, Bt, Mp = mk(st["B"], self.n_u), mk(st["Bt"], self.n_p), mk(st["Mp"], self.n_p)
            A = gsb.BlockSparseMatrix([[mats[0], Bt], [B, None]])
            self.keep = (mats, Pm, Rm, B, Bt, Mp)
            solver = gsb.GMRESSolver(30, Pr=c5_precond(gsb, mats, Pm, Rm, self.nlev, Mp), maxiter=MAXITER, atol=ATOL, rtol=RTOL)
            self.fine = mats[0]
            self.level_rows = ns_own
            self.level_nnz = [int(m[0][-1]) for m in st["mats"]]
        ns = gsb.numerical_setup(gsb.symbolic_setup(solver, A), A)
        ctx.synchronize()
        self.t_setup = time.perf_counter() - t0
        if self.config == "c5":
            x, b = gsb.Vector(ctx, A.n_own_cols), gsb.Vector(ctx, A.n_rows)
        else:
            x, b = gsb.allocate_in_domain(A), gsb.allocate_in_domain(A)
        b.set(self.b)
        return solver, ns, x, b

    # ---- oracle side (serial problems only)
    def build_oracle(self, maxiter=MAXITER):
        from oracle import linalg as ola
        from oracle import solvers as OS
        from util import host_to_scipy

        if self.config in ("c2", "c4"):
            ns_own = [lp.n_own for lp in self.hh.levels]
            mats, P, R = oracle_mats(self.hh.A, self.hh.P, self.hh.R, ns_own)
            s = (c2_solver if self.config == "c2" else c4_solver)(OS, mats, P, R, self.nlev, maxiter=maxiter)
            A = mats[0]
        else:
            st = self.st
            ns_own = [m[0].shape[0] - 1 for m in st["mats"]]
            mats, P, R = oracle_mats(st["mats"], st["P"], st["R"], ns_own)
            B, Bt = ola.CSR(host_to_scipy(st["B"], self.n_u)), ola.CSR(host_to_scipy(st["Bt"], self.n_p))
            Mp = ola.CSR(host_to_scipy(st["Mp"], self.n_p))
            A = OS.BlockMatrix([[mats[0], Bt], [B, None]])
            s = OS.GMRESSolver(30, Pr=OS.BlockPrecondAdapter(c5_precond(OS, mats, P, R, self.nlev, Mp)), maxiter=maxiter, atol=ATOL, rtol=RTOL)
        ns = OS.numerical_setup(OS.symbolic_setup(s, A), A)
        return OS, s, ns, A


def oracle_sample(prob, sample_iters, threads):
    """bounded CPU sample: `sample_iters` outer Krylov iterations of the same system on the oracle port"""
    from oracle import linalg as ola

    ola.set_threaded(True, threads)
    OS, s, ns, A = prob.build_oracle(maxiter=sample_iters)
    x = np.zeros(A.shape[1])
    t0 = time.perf_counter()
    OS.solve_(x, ns, prob.b)
    dt = time.perf_counter() - t0
    ola.set_threaded(False)
    return dt, s.log.history(), int(s.log.num_iters)


SAMPLE_ITERS = {"c2": 2, "c4": 2, "c5": 3}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path.  The Julia package cannot run on this box (no julia), so this
    times the oracle port of the same algorithm (kind "port") with ALL host threads (set explicitly).  Each step is a
    bounded sample (a few outer Krylov iterations of the same system) scaled to the full iteration count.  Under
    torchrun only rank 0 works; at N>1 it times the per-GPU unit of the weak-scaling series (throughput in MDOF/s is
    what the driver compares; one host serves all N GPUs)."""
    if rank != 0:
        return
    import gsb200  # noqa: F401  (host-side generator only; no device call is made on this arm)

    threads = host_threads()
    cfg = args.config
    cells = args.cells_per_gpu or {"c2": 128 if args.gpus == 1 else 256, "c4": 64, "c5": 256}[cfg]
    prob = Problem(cfg, cells)
    # full iteration count of this system: known from the residual history only by solving; the bounded sample
    # runs SAMPLE_ITERS iterations, the full count is taken from one complete solve when it is cheap (c2 at 128^3)
    # or from the committed GPU/oracle parity record otherwise
    full_iters = {"c2": 4, "c4": None, "c5": None}[cfg]
    rec = os.path.join(ROOT, "profiles", "iterations_r02.json")
    if os.path.exists(rec):
        try:
            full_iters = json.load(open(rec)).get(f"{cfg}:{cells}", full_iters)
        except Exception:
            pass
    if full_iters is None or (cfg == "c2" and cells not in (128, 256)):
        full_iters = None
        if prob.n_glob <= 300000:  # small system: one complete solve gives the count
            _, _, full_iters = oracle_sample(prob, MAXITER, threads)
    k = SAMPLE_ITERS[cfg]
    times, hist, it = [], None, 0
    for i in range(args.warmup + args.steps):
        dt, hist, it = oracle_sample(prob, k, threads)
        if i >= args.warmup:
            times.append(dt)
    scale = (full_iters / it) if full_iters else 1.0
    ms = 1e3 * float(np.mean(times)) * scale
    val = prob.n_glob / (ms * 1e-3) / 1e6
    sample = (f"{args.steps} samples of {it} outer iterations of the {prob.name} system ({prob.n_glob} DOFs), "
              + (f"scaled by {full_iters}/{it} to the full solve" if full_iters else "NOT scaled (full iteration count unknown): value is per-sample"))
    line = {
        "impl": "reference", "metric": prob.metric, "value": val, "unit": "MDOF/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "solve_time_ms": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "iterations": full_iters or it,
        "config": {"workload": f"{prob.name}: {prob.what}, rtol 1e-8, x0=0; CPU restatement of the reference (oracle/), {threads} OpenMP threads"
                               + ("" if args.gpus == 1 else f"; the per-GPU unit of the {args.gpus}-GPU weak-scaling series (one host, MDOF/s comparable)")},
        "cpu_baseline": {"value": val, "unit": "MDOF/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "MDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gsb200")
    ap.add_argument("--config", default=os.environ.get("GSB_BENCH_CONFIG", "c2"), choices=["c2", "c4", "c5"])
    ap.add_argument("--cells-per-gpu", type=int, default=int(os.environ.get("GSB_BENCH_CELLS", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scaling-base", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--agglomerate-rows", type=int, default=int(os.environ.get("GSB_BENCH_AGGLOMERATE_ROWS", "40000")),
                    help="N>1: GMG levels with at most this many rows per GPU live on rank 0 only (0 = every level on every GPU)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    assert args.gpus in PARTS, "--gpus must be 1, 2, 4 or 8"
    assert args.config == "c2" or world == 1, "c4 / c5 are single-GPU configurations"
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
    threads = host_threads()

    def barrier(c=None):
        (c or ctx).synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(config, cells, steps, warmup, c=None, parts=None, with_e2e=True, with_profile=True, distributed=True):
        c = c or ctx
        parts = parts or (PARTS[world] if distributed else (1, 1, 1))
        prob = Problem(config, cells, parts=parts, rank=rank if distributed else 0,
                       agglomerate_rows=args.agglomerate_rows if distributed else 0)
        solver, ns, x, b = prob.build_device(gsb, c)
        sync = barrier if distributed else (lambda: c.synchronize())
        mx = max_over_ranks if distributed else (lambda v: v)
        # ---- device-resident timing: inputs already in HBM
        hist_first = None
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()  # nvidia-smi needs ~50 ms to deliver its first sample: started before the warm-up solves (same load)
        for _ in range(warmup):
            x.fill(0.0)
            gsb.solve_(x, ns, b)
            if hist_first is None:
                # the parity record uses the FIRST solve after set-up: the block solvers warm-start their inner
                # iterative solvers from the previous solve's caches (reference quirk, BlockTriangularSolvers.jl:
                # 188-242: the y caches are zeroed at set-up only), so later solves follow a different history
                hist_first = solver.log.history()
        sync()
        l0 = c.launch_count()
        dev_ms, iters_seen = 0.0, []
        wall0 = time.perf_counter()
        for _ in range(steps):
            x.fill(0.0)
            c.timer_start()
            gsb.solve_(x, ns, b)
            dev_ms += c.timer_stop()
            iters_seen.append(solver.log.num_iters)
        sync()
        wall_ms = 1e3 * (time.perf_counter() - wall0)
        clocks = sampler.stop() if rank == 0 else None
        launches = c.launch_count() - l0
        out = dict(prob=prob, ms=mx(dev_ms / steps), wall_ms=wall_ms / steps, iters=solver.log.num_iters, iters_seen=iters_seen,
                   hist=solver.log.history(), hist_first=hist_first, flag=solver.log.flag, launches=launches // steps, clocks=clocks)
        # ---- end-to-end through the host-buffer C-ABI entry point (pinned host memory, x0 = 0 not uploaded)
        if with_e2e:
            n = prob.n_own
            xh = torch.zeros(n, dtype=torch.float64).pin_memory().numpy()
            bh = torch.from_numpy(prob.b.copy()).pin_memory().numpy()
            for _ in range(2):
                gsb.solve_(xh, ns, bh, zero_initial_guess=True)
            sync()
            t0 = time.perf_counter()
            for _ in range(steps):
                gsb.solve_(xh, ns, bh, zero_initial_guess=True)
            sync()
            out["e2e_ms"] = mx(1e3 * (time.perf_counter() - t0) / steps)
            out["h2d"], out["d2h"] = 8 * n, 8 * n
            out["xh"] = xh
            if config == "c2":
                out["x_err"] = float(np.max(np.abs(xh - synth.exact_solution(prob.hh.levels[0]))))
        # ---- per-kernel durations inside a real solve (CUDA events around every row-kernel launch)
        if with_profile:
            x.fill(0.0)
            c.profile_start()
            gsb.solve_(x, ns, b)
            out["prof"] = c.profile_stop()
            out["fine_format"] = prob.fine.format()
        return out

    cfg = args.config
    cells = args.cells_per_gpu or {"c2": 128 if world == 1 else 256, "c4": 64, "c5": 256}[cfg]
    peak, peak_src = peaks()

    # ---- multi-GPU parity, before the timed region: a 64^3-cells-per-GPU problem solved by all ranks together,
    #      iteration count and residual history against the serial CPU oracle of the same global system (rank 0)
    parity = None
    if world > 1 and not args.no_parity:
        pcells = 64
        pp = Problem("c2", pcells, parts=PARTS[world], rank=rank, agglomerate_rows=args.agglomerate_rows)
        psolver, pns, px, pb = pp.build_device(gsb, ctx)
        gsb.solve_(px, pns, pb)
        xerr = max_over_ranks(float(np.max(np.abs(px.get() - synth.exact_solution(pp.hh.levels[0])))))
        if rank == 0:
            from util import rel_hist_diff

            ps = Problem("c2", pcells, serial_of=PARTS[world])  # the serial system with the same global mesh
            dt, ohist, oit = oracle_sample(ps, MAXITER, threads)
            parity = {"problem": f"{'x'.join(str(c_) for c_ in pp.ncell)} cells ({pcells}^3 per GPU), {pp.nlev} levels",
                      "gpu_iterations": int(psolver.log.num_iters), "oracle_iterations": oit,
                      "rel_residual_history_max_diff": rel_hist_diff(psolver.log.history(), ohist),
                      "max_abs_error_vs_exact": xerr, "oracle": f"serial CPU restatement (oracle/), {threads} threads",
                      "ok": bool(psolver.log.num_iters == oit and rel_hist_diff(psolver.log.history(), ohist) < 1e-10)}
        del psolver, pns, px, pb, pp
        barrier()

    # ---- same-unit denominator of the weak-scaling series: rank 0 alone solves the per-GPU unit on its GPU
    scaling_base = None
    if cfg == "c2" and not args.no_scaling_base and (world > 1 or cells == 128):
        ucells = cells if world > 1 else 256
        if rank == 0:
            c1 = gsb.Context(device=local_rank, nranks=1, rank=0) if world > 1 else ctx
            base = measure("c2", ucells, max(2, args.steps // 3), 3, c=c1, with_e2e=False, with_profile=False, distributed=False)
            scaling_base = {"cells_per_gpu": ucells, "dofs": base["prob"].n_glob, "ms_per_step": round(base["ms"], 3),
                            "value": round(base["prob"].n_glob / (base["ms"] * 1e-3) / 1e6, 3), "unit": "MDOF/s",
                            "iterations": base["iters"], "levels": base["prob"].nlev,
                            "note": "N=1 run of the per-GPU unit of the N>1 series, same box, same run"}
            del base
            if world > 1:
                c1.close()
        if world > 1:
            dist.barrier()

    res = measure(cfg, cells, args.steps, args.warmup)
    prob = res["prob"]

    # roofline of the dominant kernel: the fused Jacobi-Richardson sweep on the finest level
    fmt = res.get("fine_format") or {}

    def algo_bytes(mode, nrows, nnz):
        return 12 * nnz + PER_ROW[mode] * nrows

    kernels = []
    for p in res.get("prof", []):
        t = p["total_ms"] / p["count"] * 1e-3
        kernels.append({"kernel": p["mode"], "impl": p["impl"], "rows": p["nrows"], "nnz": p["nnz"],
                        "launches": p["count"], "avg_us": round(t * 1e6, 2),
                        "GBps": round(algo_bytes(p["mode"], p["nrows"], p["nnz"]) / t / 1e9, 1),
                        "share_of_solve": round(p["total_ms"] / res["ms"], 4)})
    kernels.sort(key=lambda k: -k["share_of_solve"])
    top = next((k for k in kernels if k["kernel"] == "sweep" and k["rows"] == prob.level_rows[0]), kernels[0] if kernels else None)
    roofline = None
    if top:
        fbytes = fmt.get("bytes_per_pass", 0) + (PER_ROW[top["kernel"]] - 4) * top["rows"]
        roofline = {"bound": "hbm", "kernel": f"sell_kernel<{top['kernel']}> level 1 ({top['impl']}: fused Jacobi-Richardson sweep, block-SELL-32 "
                                              f"{fmt.get('block_size', 1)}x{fmt.get('block_size', 1)} blocks{', rows sorted in windows of 256' if fmt.get('sorted') else ''})",
                    "achieved": top["GBps"], "peak": peak, "unit": "GB/s", "frac": round(top["GBps"] / peak, 4),
                    "frac_of_nominal_8TBs": round(top["GBps"] / 8000.0, 4), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": algo_bytes(top["kernel"], top["rows"], top["nnz"]),
                    "algorithmic_bytes_model": "SURVEY 8d: 12 B per non-zero (fp64 value + int32 column) + 44 B per row",
                    "format_bytes_per_launch": fbytes,
                    "format_GBps": round(fbytes / (top["avg_us"] * 1e-6) / 1e9, 1), "frac_format_bytes": round(fbytes / (top["avg_us"] * 1e-6) / 1e9 / peak, 4),
                    "avg_launch_us": top["avg_us"], "traffic": None,
                    "note": "achieved / frac use the ALGORITHMIC bytes of SURVEY 8d (CSR: 12 B per non-zero) as the measurement contract asks; the "
                            "block-SELL format streams fewer bytes (format_bytes_per_launch, confirmed by the ncu DRAM `traffic`), so frac can exceed 1: "
                            "the HBM utilisation of the kernel is frac_format_bytes"}
        tfile = os.path.join(ROOT, "profiles", "traffic_r02.json")
        if os.path.exists(tfile):
            try:
                with open(tfile) as f:
                    roofline["traffic"] = json.load(f).get(f"{cfg}:{cells}:sweep_level1_dram_bytes_per_launch")
            except Exception:
                pass

    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    n_glob = prob.n_glob
    value = n_glob / (res["ms"] * 1e-3) / 1e6
    line = {
        "metric": prob.metric, "value": round(value, 3), "unit": "MDOF/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(res["ms"], 4),
        "solve_time_ms": round(res["ms"], 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "iterations": res["iters"], "iterations_per_step": res["iters_seen"], "final_rel_residual": float(res["hist"][-1] / res["hist"][0]),
        "converged_flag": res["flag"],
        "config": {
            "workload": f"{prob.name}: {prob.what}, rtol 1e-8, x0=0",
            "partition": "x".join(str(p) for p in PARTS[world]), "levels_rows_rank0": prob.level_rows, "levels_nnz_rank0": prob.level_nnz,
            "parts_per_level": ["x".join(str(q) for q in pl) for pl in (getattr(prob, "parts_per_level", None) or [])] or None,
            "fine_matrix_format": fmt,
            "l2_policy": "inputs larger than L2 (fine-level matrix %.0f MB as stored vs 126 MB L2); coarse levels are L2-resident by construction" % (fmt.get("bytes_per_pass", 0) / 1e6),
            "setup_s": round(prob.t_setup, 3), "host_generation_s": round(prob.t_gen, 3),
        },
        "gpu_launches": res["launches"],
        "clocks": res["clocks"],
        "e2e": {"value": round(n_glob / (res["e2e_ms"] * 1e-3) / 1e6, 3), "unit": "MDOF/s", "ms_per_step": round(res["e2e_ms"], 4),
                "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"],
                "note": "gsb_solve_host_zero_guess: pinned host buffers, H2D of b, solve, D2H of x inside the timed region (x0 = 0 is not uploaded)"},
        "roofline": roofline,
        "kernels": kernels[:12],
    }
    if "x_err" in res:
        line["e2e"]["max_abs_error_vs_exact"] = res["x_err"]
    # CPU baseline (oracle port) on a bounded sample, rank 0, N=1 only; doubles as the parity record of this line
    if world == 1 and not args.no_cpu_baseline:
        from util import rel_hist_diff

        k = SAMPLE_ITERS[cfg]
        dt, ohist, oit = oracle_sample(prob, k, threads)
        est = dt * res["iters"] / max(oit, 1)
        d = rel_hist_diff(res["hist_first"][: oit + 1], ohist)
        line["cpu_baseline"] = {"value": round(n_glob / est / 1e6, 4), "unit": "MDOF/s", "cores": threads, "kind": "port",
                                "est_solve_time_ms": round(est * 1e3, 1),
                                "sample": f"{oit} outer iterations (of {res['iters']}) of the same {prob.name} system on the oracle port with "
                                          f"{threads} OpenMP threads, scaled by {res['iters']}/{oit}"}
        line["parity"] = {"against": "CPU restatement (oracle/) on the same assembled system; first solve after set-up on both sides", "iterations_compared": oit,
                          "rel_residual_history_max_diff": d, "ok": bool(d < 1e-10)}
    if world > 1:
        # halo volume of one consistent! on the finest level (rank 0) against the NVLink roofline (900 GB/s per direction)
        lp0 = prob.hh.levels[0]
        snd = int(lp0.snd_ptrs[-1]) * 8
        line["halo"] = {"level1_send_bytes_per_exchange_rank0": snd, "neighbours_rank0": int(len(lp0.nbr_snd)),
                        "nvlink_time_at_900GBps_us": round(snd / 900e9 * 1e6, 2),
                        "note": "an exchange costs 20-30 us end to end (DESIGN.md section 6): latency-bound, a few percent of the NVLink roofline"}
    if parity is not None:
        line["parity"] = parity
    if scaling_base is not None:
        line["weak_scaling_base"] = scaling_base
        if world > 1:
            line["weak_scaling_efficiency_same_unit"] = round(scaling_base["ms_per_step"] / res["ms"], 4)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


if __name__ == "__main__":
    main()
