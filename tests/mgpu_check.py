"""Multi-GPU parity check, launched with torchrun (one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_check.py
Checks, through the C ABI over NCCL:
  1. consistent!: ghosts receive the owners' values (integer-exact index maps)
  2. distributed SpMV == oracle CSR mul on the local block (own-first column order) bit-exactly, and
     == the serial product up to rounding
  3. distributed GMG-PCG (V-cycle) iteration count and residual history == serial oracle solve (1e-10)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PARTS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def main():
    import torch
    import torch.distributed as dist

    import gsb200 as gsb
    from gsb200 import synth
    from oracle import linalg as ola
    from oracle import solvers as OS
    from util import oracle_hierarchy, rel_hist_diff

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ids = [gsb.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx = gsb.Context(device=local, nranks=world, rank=rank, nccl_id=ids[0])
    parts = PARTS[world]
    cells = int(os.environ.get("MGPU_CELLS", "16"))
    nlev = 3
    ncell = tuple(cells * p for p in parts)
    lengths = tuple(float(p) for p in parts)  # cubic cells: the domain grows with the part grid
    hh = synth.poisson_hierarchy_host(ncell, nlev, parts=parts, rank=rank, lengths=lengths)
    dh = synth.upload_hierarchy(ctx, hh)
    lp = hh.levels[0]
    gid = synth.lexicographic_ids(lp)
    N = int(np.prod([c - 1 for c in ncell]))
    xg = np.sin(np.arange(N, dtype=np.float64))

    # 1. consistent!
    A = dh.A[0]
    x = gsb.allocate_in_domain(A)
    x.set(xg[gid[: lp.n_own]])
    gsb.consistent_(x, dh.plans[0])
    assert np.array_equal(x.get_local(), xg[gid]), "ghost values differ from owners'"

    # 2. SpMV (the "auto" runs use the own/ghost split with the halo exchange overlapped on a second stream)
    ctx.set_option("overlap", "1")
    ctx.set_option("overlap_min_rows", os.environ.get("MGPU_OVERLAP_MIN_ROWS", "1"))
    for kern in ("auto", "sell", "stream", "vector"):
        ctx.set_option("spmv", kern)
        y = gsb.allocate_in_range(A)
        x.set(xg[gid[: lp.n_own]])
        gsb.mul_(y, A, x)
        Ao = ola.CSR(synth.to_scipy(*hh.A[0], lp.n_own + lp.n_ghost))
        yo = np.zeros(lp.n_own)
        ola.mul(yo, Ao, xg[gid])
        assert np.array_equal(y.get(), yo), f"distributed SpMV ({kern}) not bit-identical to the own-first CSR order"
        if kern == "auto":  # 5-arg form and repeated exchanges through the split path
            y.set(yo)
            gsb.mul_(y, A, x, -0.5, 1.0)
            y2 = yo.copy()
            ola.mul5(y2, Ao, xg[gid], -0.5, 1.0)
            assert np.array_equal(y.get(), y2)
    ctx.set_option("spmv", "auto")
    d = gsb.dot(x, x)
    assert abs(d - float(xg @ xg)) <= 1e-13 * float(xg @ xg)

    # 3. GMG-PCG vs the serial oracle: (a) overlapped two-stream halo exchange, (b) default path
    #    (peer-memory exchange inside a replayed CUDA graph); both must give the same history
    sm = gsb.Fill(gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), 10, 2.0 / 3.0), nlev - 1)
    gmg = gsb.GMGLinearSolver(dh.A, dh.P, dh.R, pre_smoothers=sm, post_smoothers=sm, maxiter=1)
    s = gsb.CGSolver(gmg, maxiter=30, atol=1e-14, rtol=1e-8)
    ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
    xs, b = gsb.allocate_in_domain(A), gsb.allocate_in_domain(A)
    b.set(hh.b)
    gsb.solve_(xs, ns, b)
    hist_overlap = s.log.history()
    ctx.set_option("overlap", "0")
    for _ in range(2):  # second solve replays the captured graph
        xs.fill(0.0)
        gsb.solve_(xs, ns, b)
        assert np.array_equal(s.log.history(), hist_overlap), "graph / overlap paths disagree"
    err = float(np.max(np.abs(xs.get() - synth.exact_solution(lp))))
    if rank == 0:
        hs = synth.poisson_hierarchy_host(ncell, nlev, lengths=lengths)
        mats, P, R = oracle_hierarchy(hs)
        smo = [OS.RichardsonSmoother(OS.JacobiLinearSolver(), 10, 2.0 / 3.0)] * (nlev - 1)
        go = OS.GMGLinearSolver(mats, P, R, pre_smoothers=smo, post_smoothers=smo, maxiter=1)
        so = OS.CGSolver(go, maxiter=30, atol=1e-14, rtol=1e-8)
        xo = np.zeros(mats[0].shape[0])
        OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, mats[0]), mats[0]), hs.b)
        dd = rel_hist_diff(s.log.history(), so.log.history())
        assert s.log.num_iters == so.log.num_iters, (s.log.num_iters, so.log.num_iters)
        assert dd < 1e-10, dd
        print(f"mgpu_check ok: world={world} cells={ncell} iters={s.log.num_iters} hist_diff={dd:.2e} max_err={err:.2e}", flush=True)
    assert err < 1e-7
    # 4. the NCCL send/recv exchange (fallback when peer memory cannot be mapped) gives the same bits
    ctx.set_option("p2p", "0")
    plan2 = gsb.ExchangePlan(ctx, lp.n_own, lp.n_ghost, lp.nbr_snd, lp.snd_ptrs, lp.snd_ids, lp.nbr_rcv, lp.rcv_ptrs, lp.rcv_ids)
    A2 = gsb.SparseMatrix(ctx, lp.n_own, lp.n_own, lp.n_ghost, *hh.A[0], plan=plan2)
    x2, y2 = gsb.allocate_in_domain(A2), gsb.allocate_in_range(A2)
    x2.set(xg[gid[: lp.n_own]])
    gsb.mul_(y2, A2, x2)
    assert np.array_equal(y2.get(), yo), "NCCL halo path differs"
    ctx.set_option("p2p", "1")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
