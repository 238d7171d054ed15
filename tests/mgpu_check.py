"""Multi-GPU parity check, launched with torchrun (one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_check.py
Checks, through the C ABI over NCCL:
  1. consistent!: ghosts receive the owners' values (integer-exact index maps)
  2. distributed SpMV == oracle CSR mul on the local block (own-first column order) bit-exactly, and
     == the serial product up to rounding
  3. distributed GMG-PCG (V-cycle) iteration count and residual history == serial oracle solve (1e-10)
  4. the NCCL send/recv halo path gives the same bits as the peer-memory path
  5. levels on fewer parts (np_per_level): redistribute! round trip is exact, and GMG-PCG with the two coarsest levels
     agglomerated on rank 0 reproduces the serial oracle's history (1e-10)
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
    for kern in ("auto", "vector"):  # block-SELL row kernel (interior / boundary split) and the CSR fallback kernel
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

    # 5. levels on fewer parts: 4 levels, the two coarsest on rank 0 only
    nlev2 = 4
    one = (1,) * len(parts)
    ppl = [parts, parts, one, one]
    hh2 = synth.poisson_hierarchy_host(ncell, nlev2, parts=parts, rank=rank, lengths=lengths, parts_per_level=ppl)
    dh2 = synth.upload_hierarchy(ctx, hh2)
    assert dh2.redist is not None and (hh2.levels[2].n_own > 0) == (rank == 0)
    #    redistribute there and back: own values of the level-3 coarse space in the partition of level 2's parts
    cr = hh2.coarse_red[1]
    gcr = synth.lexicographic_ids(cr)
    N3 = int(np.prod([c // 4 - 1 for c in ncell]))
    zg = np.cos(np.arange(N3, dtype=np.float64))
    src = gsb.Vector(ctx, cr.n_own)
    src.set(zg[gcr[: cr.n_own]])
    agg = gsb.Vector(ctx, hh2.levels[2].n_own)
    gsb.redistribute_(agg, dh2.to_coarse[1], src)
    if rank == 0:
        assert np.array_equal(agg.get(), zg), "redistribute to the agglomerated level is not the serial vector"
    back = gsb.Vector(ctx, cr.n_own)
    gsb.redistribute_(back, dh2.to_fine[1], agg)
    assert np.array_equal(back.get(), zg[gcr[: cr.n_own]]), "redistribute round trip"
    sm2 = gsb.Fill(gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), 10, 2.0 / 3.0), nlev2 - 1)
    gmg2 = gsb.GMGLinearSolver(dh2.A, dh2.P, dh2.R, pre_smoothers=sm2, post_smoothers=sm2, maxiter=1, redist=dh2.redist)
    s2 = gsb.CGSolver(gmg2, maxiter=30, atol=1e-14, rtol=1e-8)
    A2 = dh2.A[0]
    ns2 = gsb.numerical_setup(gsb.symbolic_setup(s2, A2), A2)
    xs2, b2 = gsb.allocate_in_domain(A2), gsb.allocate_in_domain(A2)
    b2.set(hh2.b)
    hist2 = None
    for _ in range(3):  # eager, captured, replayed
        xs2.fill(0.0)
        gsb.solve_(xs2, ns2, b2)
        if hist2 is None:
            hist2 = s2.log.history()
        assert np.array_equal(s2.log.history(), hist2), "graph replay changes the history of the agglomerated GMG"
    err2 = float(np.max(np.abs(xs2.get() - synth.exact_solution(hh2.levels[0]))))
    if rank == 0:
        hs2 = synth.poisson_hierarchy_host(ncell, nlev2, lengths=lengths)
        mats2, P2, R2 = oracle_hierarchy(hs2)
        smo2 = [OS.RichardsonSmoother(OS.JacobiLinearSolver(), 10, 2.0 / 3.0)] * (nlev2 - 1)
        go2 = OS.GMGLinearSolver(mats2, P2, R2, pre_smoothers=smo2, post_smoothers=smo2, maxiter=1)
        so2 = OS.CGSolver(go2, maxiter=30, atol=1e-14, rtol=1e-8)
        xo2 = np.zeros(mats2[0].shape[0])
        OS.solve_(xo2, OS.numerical_setup(OS.symbolic_setup(so2, mats2[0]), mats2[0]), hs2.b)
        dd2 = rel_hist_diff(hist2, so2.log.history())
        assert s2.log.num_iters == so2.log.num_iters, (s2.log.num_iters, so2.log.num_iters)
        assert dd2 < 1e-10, dd2
        print(f"mgpu_check agglomerated ok: world={world} levels on parts {ppl} iters={s2.log.num_iters} hist_diff={dd2:.2e} "
              f"max_err={err2:.2e}", flush=True)
    assert err2 < 1e-7
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
