"""CPU tests (no GPU) of the host-side logic: C-ABI surface, PartitionedArrays-style partition /
exchange plans (integer-exact), and the N>1 path with world_size-2 gloo processes."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import gsb200
from gsb200 import synth
from oracle import fem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gsb200.h")).read()
    declared = set(re.findall(r"\b(gsb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 45
    L = gsb200._lib.lib()  # (imports torch first so that one libnccl.so.2 serves both)
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    # the ctypes table binds exactly the declared surface
    assert set(gsb200._lib.SIGNATURES) == declared


def test_julia_shim_binds_only_declared_symbols():
    """every `ccall((:gsb_xxx, libgsb), ...)` of the Julia shim names an entry point of include/gsb200.h"""
    hdr = open(os.path.join(ROOT, "include", "gsb200.h")).read()
    declared = set(re.findall(r"\b(gsb_[a-z0-9_]+)\s*\(", hdr))
    jl = open(os.path.join(ROOT, "gridapsolvers.jl_b200", "julia", "GridapSolversB200.jl")).read()
    used = set(re.findall(r":(gsb_[a-z0-9_]+)", jl))
    assert len(used) >= 15
    assert used <= declared, sorted(used - declared)


def test_null_handles_are_errors_not_crashes():
    """every entry point rejects a NULL first handle with GSB_EINVAL and a message (no device needed)"""
    L = gsb200._lib.lib()
    n = ctypes.c_int64()
    d = ctypes.c_double()
    assert L.gsb_vec_fill(None, 1.0) == 1
    assert b"NULL handle" in L.gsb_last_error(None)
    assert L.gsb_spmv(None, None, None, 1.0, 0.0) == 1
    assert L.gsb_solve(None, None, None) == 1
    assert L.gsb_mat_info(None, ctypes.byref(n), None, None, None) == 1
    assert L.gsb_dot(None, None, ctypes.byref(d)) == 1
    assert L.gsb_solver_log(None, None, None, 0, None) == 1
    assert L.gsb_mat_destroy(None) == 0 and L.gsb_solver_destroy(None) == 0  # destroying NULL is a no-op


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gsb200.GSBError) as e:
        gsb200.Context()
    assert "no CUDA device" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gridapsolvers.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


@pytest.mark.parametrize("nc,parts", [((8, 8), (2, 2)), ((16, 8, 8), (2, 1, 1)), ((8, 8, 8), (2, 2, 2)), ((16, 16, 8), (4, 2, 1))])
def test_partition_reassembles_the_serial_system(nc, parts):
    """own rows of all parts, mapped through (own-first local id -> lexicographic id), give exactly
    the serial matrix / rhs; own dofs of part p get consecutive global ids after those of part p-1"""
    sysm = fem.poisson(nc)
    nranks = int(np.prod(parts))
    N = sysm.A.shape[0]
    Aglob = sp.lil_matrix((N, N))
    bglob = np.zeros(N)
    seen = np.zeros(N, dtype=int)
    off = 0
    for r in range(nranks):
        lp = synth.make_level_part(nc, parts, r)
        assert lp.own_offset_global() == off
        off += lp.n_own
        rp, col, val, b = synth.poisson_rows(lp)
        gid = synth.lexicographic_ids(lp)
        Al = synth.to_scipy(rp, col, val, lp.n_own + lp.n_ghost).tocoo()
        Aglob[gid[Al.row], gid[Al.col]] = Al.data
        bglob[gid[: lp.n_own]] = b
        seen[gid[: lp.n_own]] += 1
        # columns ascending within rows, own before ghost
        for i in range(lp.n_own):
            c = col[rp[i]:rp[i + 1]]
            assert (np.diff(c) > 0).all()
        # ghosts sorted by (owner, owner-local id); receive lists tile the ghost tail
        key = lp.ghost_owner.astype(np.int64) * (1 << 40) + lp.ghost_owner_lid
        assert (np.diff(key) > 0).all()
        assert lp.rcv_ptrs[-1] == lp.n_ghost and (lp.rcv_ids == lp.n_own + np.arange(lp.n_ghost)).all()
    assert off == N and (seen == 1).all()
    assert abs(Aglob.tocsr() - sysm.A).max() < 1e-15
    assert np.abs(bglob - sysm.b).max() < 1e-15


@pytest.mark.parametrize("nc,parts", [((16, 8, 8), (2, 1, 1)), ((8, 8, 8), (2, 2, 2))])
def test_exchange_plans_are_mutually_consistent(nc, parts):
    """what p sends to q is exactly what q expects from p, in the same order (global ids match)"""
    nranks = int(np.prod(parts))
    lps = [synth.make_level_part(nc, parts, r) for r in range(nranks)]
    gids = [synth.lexicographic_ids(lp) for lp in lps]
    for p, lp in enumerate(lps):
        assert lp.snd_ids.max(initial=-1) < lp.n_own
        for k, q in enumerate(lp.nbr_snd):
            sent = gids[p][lp.snd_ids[lp.snd_ptrs[k]:lp.snd_ptrs[k + 1]]]
            lq = lps[q]
            kk = list(lq.nbr_rcv).index(p)
            expected = gids[q][lq.rcv_ids[lq.rcv_ptrs[kk]:lq.rcv_ptrs[kk + 1]]]
            assert np.array_equal(sent, expected)
        assert sorted(lp.nbr_snd) == sorted(lp.nbr_rcv)  # symmetric stencil => symmetric neighbours


def test_transfer_operators_distributed_match_serial():
    nc, parts = (16, 16, 8), (2, 2, 1)
    H = fem.poisson_hierarchy(nc, 2)
    Pg = sp.lil_matrix(H.P[0].shape)
    Rg = sp.lil_matrix(H.R[0].shape)
    for r in range(4):
        hh = synth.poisson_hierarchy_host(nc, 2, parts=parts, rank=r)
        f, c = hh.levels
        gf, gc = synth.lexicographic_ids(f), synth.lexicographic_ids(c)
        Pl = synth.to_scipy(*hh.P[0], c.n_own + c.n_ghost).tocoo()
        Pg[gf[Pl.row], gc[Pl.col]] = Pl.data
        Rl = synth.to_scipy(*hh.R[0], f.n_own + f.n_ghost).tocoo()
        Rg[gc[Rl.row], gf[Rl.col]] = Rl.data
    assert abs(Pg.tocsr() - H.P[0]).max() == 0.0
    assert abs(Rg.tocsr() - H.R[0]).max() == 0.0


@pytest.mark.parametrize("nc,src,dst", [((16, 8, 8), (2, 1, 1), (1, 1, 1)), ((8, 8, 8), (2, 2, 2), (1, 1, 1)),
                                        ((8, 8, 8), (2, 2, 2), (2, 1, 1)), ((8, 8, 8), (1, 1, 1), (2, 2, 1))])
def test_redistribution_lists_move_every_own_value_once(nc, src, dst):
    """RedistributionOperator data for Cartesian partitions (GridTransferOperators.jl:447-532): what p sends to q is what
    q expects from p, in the same order; applying the lists to the distributed pieces of a global vector yields the
    pieces of the same vector in the destination partition; ranks beyond the destination grid end up empty"""
    nranks = max(int(np.prod(src)), int(np.prod(dst)))
    L = [synth.redistribution_lists(nc, src, dst, r) for r in range(nranks)]
    sp_ = [synth.level_part_or_empty(nc, src, r) for r in range(nranks)]
    dp_ = [synth.level_part_or_empty(nc, dst, r) for r in range(nranks)]
    N = int(np.prod([n - 1 for n in nc]))
    xg = np.cos(np.arange(N, dtype=np.float64))
    gs = [synth.lexicographic_ids(lp)[: lp.n_own] if lp.n_own else np.zeros(0, dtype=np.int64) for lp in sp_]
    gd = [synth.lexicographic_ids(lp)[: lp.n_own] if lp.n_own else np.zeros(0, dtype=np.int64) for lp in dp_]
    out = [np.full(lp.n_own, np.nan) for lp in dp_]
    for p_, (ns, nd, nbs, sptr, sids, nbr, rptr, rids) in enumerate(L):
        assert ns == sp_[p_].n_own and nd == dp_[p_].n_own
        assert sptr[-1] == ns and rptr[-1] == nd  # every own value leaves once, every own slot is filled once
        assert np.array_equal(np.sort(sids), np.arange(ns)) and np.array_equal(np.sort(rids), np.arange(nd))
        for k, q_ in enumerate(nbs):
            sent = sids[sptr[k]:sptr[k + 1]]
            nq = L[q_]
            kk = list(nq[5]).index(p_)
            slots = nq[7][nq[6][kk]:nq[6][kk + 1]]
            assert np.array_equal(gs[p_][sent], gd[q_][slots])  # same global dofs, same order
            out[q_][slots] = xg[gs[p_][sent]]
    for r in range(nranks):
        assert np.array_equal(out[r], xg[gd[r]])
        if r >= int(np.prod(dst)):
            assert dp_[r].n_own == 0


def test_hierarchy_with_levels_on_fewer_parts_matches_serial_transfers():
    """np_per_level semantics: the two finest levels on 2x2x1 parts, the coarse ones on one part.  Across the boundary
    the restriction acts in the partition of the finer level and is followed by the redistribution (and the prolongation
    preceded by the reverse one): together they reproduce the serial R x and P x"""
    nc, parts, nlev = (16, 16, 8), (2, 2, 1), 3
    ppl = [parts, parts, (1, 1, 1)]
    H = fem.poisson_hierarchy(nc, nlev)
    hhs = [synth.poisson_hierarchy_host(nc, nlev, parts=parts, rank=r, parts_per_level=ppl) for r in range(4)]
    for r, hh in enumerate(hhs):
        assert hh.coarse_red[0] is None and hh.coarse_red[1] is not None
        assert (hh.levels[2].n_own > 0) == (r == 0)  # the coarsest level lives on rank 0 only
        if r == 0:  # ... where it is the serial matrix in the serial ordering
            A2 = synth.to_scipy(*hh.A[2], hh.levels[2].n_own)
            assert abs(A2 - H.mats[2]).max() < 1e-14
    # restriction from level 2 (distributed) to level 3 (rank 0): y = redistribute(R_local x_local)
    N1, N2 = H.mats[1].shape[0], H.mats[2].shape[0]
    x1 = np.sin(1.0 + np.arange(N1, dtype=np.float64))
    y_ref = H.R[1] @ x1
    y0 = np.full(N2, np.nan)
    z2 = np.cos(np.arange(N2, dtype=np.float64))
    p_ref = H.P[1] @ z2
    p_out = np.full(N1, np.nan)
    for r, hh in enumerate(hhs):
        f, cr = hh.levels[1], hh.coarse_red[1]
        gf, gc = synth.lexicographic_ids(f), synth.lexicographic_ids(cr)
        yl = synth.to_scipy(*hh.R[1], f.n_own + f.n_ghost) @ x1[gf]           # own rows of the redistributed coarse space
        ns, nd, nbs, sptr, sids, nbr, rptr, rids = hh.to_coarse[1]
        assert list(nbs) == [0] and ns == cr.n_own                            # everything goes to rank 0
        k0 = list(hhs[0].to_coarse[1][5]).index(r)
        slots = hhs[0].to_coarse[1][7][hhs[0].to_coarse[1][6][k0]:hhs[0].to_coarse[1][6][k0 + 1]]
        y0[slots] = yl[sids]
        # prolongation: rank 0 scatters z2 (its own, serial order) to the redistributed layout, ghosts filled, P applied
        zl = z2[gc]                                                            # what redistribute + consistent! deliver
        pl = synth.to_scipy(*hh.P[1], cr.n_own + cr.n_ghost) @ zl
        p_out[gf[: f.n_own]] = pl
        ns2, nd2, nbs2, sptr2, sids2, nbr2, rptr2, rids2 = hh.to_fine[1]
        assert nd2 == cr.n_own and (list(nbr2) == [0] if cr.n_own else True)
    assert np.abs(y0 - y_ref).max() <= 1e-14 * np.abs(y_ref).max()
    assert np.abs(p_out - p_ref).max() <= 1e-14 * np.abs(p_ref).max()


def _gloo_worker(rank, world, nc, parts, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lp = synth.make_level_part(nc, parts, rank)
        rp, col, val, b = synth.poisson_rows(lp)
        gid = synth.lexicographic_ids(lp)
        N = int(np.prod([n - 1 for n in nc]))
        xg = np.sin(np.arange(N, dtype=np.float64))
        x = np.zeros(lp.n_own + lp.n_ghost)
        x[: lp.n_own] = xg[gid[: lp.n_own]]
        # consistent!(x): the same pack / send / recv / unpack sequence libgsb200 runs over NCCL
        reqs, bufs = [], []
        for k, qn in enumerate(lp.nbr_rcv):
            buf = torch.zeros(int(lp.rcv_ptrs[k + 1] - lp.rcv_ptrs[k]), dtype=torch.float64)
            bufs.append(buf)
            reqs.append(dist.irecv(buf, src=int(qn)))
        for k, qn in enumerate(lp.nbr_snd):
            s = torch.from_numpy(x[lp.snd_ids[lp.snd_ptrs[k]:lp.snd_ptrs[k + 1]]].copy())
            reqs.append(dist.isend(s, dst=int(qn)))
        for r in reqs:
            r.wait()
        for k, buf in enumerate(bufs):
            x[lp.rcv_ids[lp.rcv_ptrs[k]:lp.rcv_ptrs[k + 1]]] = buf.numpy()
        assert np.array_equal(x, xg[gid])  # ghosts hold the owners' values
        y = synth.to_scipy(rp, col, val, lp.n_own + lp.n_ghost) @ x
        # dot over own values + allreduce == serial dot (PartitionedArrays dot semantics)
        t = torch.tensor([float(np.dot(x[: lp.n_own], y))], dtype=torch.float64)
        dist.all_reduce(t)
        q.put((rank, gid[: lp.n_own], y, float(t.item())))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_halo_spmv_and_dot():
    import torch.multiprocessing as mp

    nc, parts = (16, 8, 8), (2, 1, 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, nc, parts, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sysm = fem.poisson(nc)
    N = sysm.A.shape[0]
    xg = np.sin(np.arange(N, dtype=np.float64))
    yg = sysm.A @ xg
    y = np.zeros(N)
    for rank, gid, yl, d in res:
        y[gid] = yl
        assert abs(d - float(xg @ yg)) <= 1e-12 * abs(float(xg @ yg)) + 1e-12
    assert np.abs(y - yg).max() < 1e-13


def _gloo_pcg_worker(rank, world, nc, parts, port, q):
    """Jacobi-preconditioned CG over a 2-part PartitionedArrays-style distribution, numpy local kernels,
    gloo for consistent! and the dot all-reduces: the same statement sequence as CGSolvers.jl:73-120."""
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lp = synth.make_level_part(nc, parts, rank)
        rp, col, val, b = synth.poisson_rows(lp)
        A = synth.to_scipy(rp, col, val, lp.n_own + lp.n_ghost)
        n = lp.n_own

        def consistent(v):
            reqs, bufs = [], []
            for k, qn in enumerate(lp.nbr_rcv):
                buf = torch.zeros(int(lp.rcv_ptrs[k + 1] - lp.rcv_ptrs[k]), dtype=torch.float64)
                bufs.append(buf)
                reqs.append(dist.irecv(buf, src=int(qn)))
            for k, qn in enumerate(lp.nbr_snd):
                reqs.append(dist.isend(torch.from_numpy(v[lp.snd_ids[lp.snd_ptrs[k]:lp.snd_ptrs[k + 1]]].copy()), dst=int(qn)))
            for r in reqs:
                r.wait()
            for k, buf in enumerate(bufs):
                v[lp.rcv_ids[lp.rcv_ptrs[k]:lp.rcv_ptrs[k + 1]]] = buf.numpy()

        def pdot(a, c):
            t = torch.tensor([float(np.dot(a[:n], c[:n]))], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t.item())

        def mul(v):  # mul!(w,A,v): halo of v, then the local rows
            consistent(v)
            return A @ v

        inv_diag = 1.0 / A.diagonal()[:n]  # own-own block diagonal (JacobiLinearSolvers.jl:29-34)
        nl = lp.n_own + lp.n_ghost
        x, p, z, r, w = (np.zeros(nl) for _ in range(5))
        r[:n] = b - mul(x)
        gamma, hist = 1.0, [np.sqrt(pdot(r, r))]
        for _ in range(200):
            z[:n] = inv_diag * r[:n]
            beta = gamma
            gamma = pdot(z, r)
            beta = gamma / beta
            p[:n] = z[:n] + beta * p[:n]
            w[:n] = mul(p)
            alpha = gamma / pdot(p, w)
            x[:n] += alpha * p[:n]
            r[:n] -= alpha * w[:n]
            hist.append(np.sqrt(pdot(r, r)))
            if hist[-1] / hist[0] < 1e-8:
                break
        q.put((rank, synth.lexicographic_ids(lp)[:n], x[:n].copy(), hist))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_distributed_pcg_matches_serial_oracle():
    import torch.multiprocessing as mp

    from oracle import linalg as ola
    from oracle import solvers as OS

    nc, parts = (8, 8, 16), (1, 1, 2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_pcg_worker, args=(r, 2, nc, parts, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sysm = fem.poisson(nc)
    Ao = ola.CSR(sysm.A)
    s = OS.CGSolver(OS.JacobiLinearSolver(), maxiter=200, atol=0.0, rtol=1e-8)
    xo = np.zeros(Ao.shape[0])
    OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(s, Ao), Ao), sysm.b)
    x = np.zeros(Ao.shape[0])
    for rank, gid, xl, hist in res:
        x[gid] = xl
        assert len(hist) - 1 == s.log.num_iters
        assert np.max(np.abs(np.array(hist) - s.log.history())) <= 1e-10 * hist[0]
    assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)


def test_host_mirror_plumbing():
    """HierarchicalArray / with_level / tolerances: host-only pieces of the reference interface"""
    # levels on fewer parts (np_per_level = [4, 4, 1]): rank 2 does not hold level 3 (HierarchicalArrays.jl:96-149)
    ranks = [None, range(4), [0]]
    h = gsb200.HierarchicalArray(["A1", "A2", "A3"], ranks, rank=2)
    assert gsb200.num_levels(h) == 3 and h[1] == "A2" and h[2] is None
    assert gsb200.with_level(lambda a: a + "!", h, 1) == "A1!"
    assert gsb200.with_level(lambda a: a, h, 3, default="skip") == "skip"  # rank not in the level
    assert gsb200.with_level(lambda a: a, h, 3) is None
    m = h.map(lambda a, b: a + b, gsb200.HierarchicalArray(["x", "y", "z"], ranks, rank=2))
    assert m.array == ["A1x", "A2y", None] and m.ranks == ranks
    h0 = gsb200.HierarchicalArray(["A1", "A2", "A3"], ranks, rank=0)
    assert gsb200.with_level(lambda a: a, h0, 3, default="skip") == "A3"
    assert gsb200.with_level(lambda a: a * 2, [1, 2, 3], 2) == 4  # plain arrays: every rank holds every level
    s = gsb200.CGSolver(maxiter=7)
    assert s.log.residuals.shape[0] == 8 and gsb200.get_solver_tolerances(s).maxiter == 7
    gsb200.set_solver_tolerances_(s, maxiter=12, rtol=1e-9)
    assert s.log.tols.maxiter == 12 and s.log.tols.rtol == 1e-9 and s.log.residuals.shape[0] == 13
    # Fill shares ONE smoother object; pre is post => shared caches (GMGLinearSolvers.jl:52,190-194)
    sm = gsb200.Fill(gsb200.RichardsonSmoother(gsb200.JacobiLinearSolver(), 10, 2.0 / 3.0), 2)
    assert sm[0] is sm[1]
    g = gsb200.GMGLinearSolver(gsb200.HierarchicalArray([1, 2, 3]), [1, 2], [1, 2], pre_smoothers=sm, post_smoothers=sm)
    assert g.pre_smoothers is g.post_smoothers and g.log.tols.maxiter == 100 and g.log.tols.rtol == 1e-8
    with pytest.raises(AssertionError):
        gsb200.GMGLinearSolver([1, 2, 3], [1], [1, 2])


def test_explicit_transfer_by_coloured_probing():
    """setup_transfer_operators hands the GMG objects that only support mul! (GridTransferOperators.jl:350-401); the
    shim materialises them by coloured probing (julia/GridapSolversB200.jl explicit_transfer, mirrored in api.py).
    Probing the oracle's prolongation / restriction through mul only recovers the matrices exactly: 2^d colours on the
    coarse node grid for P, 3^d on the fine grid for R = P'; an invalid colouring is detected"""
    nc = (8, 8, 8)
    H = fem.poisson_hierarchy(nc, 2)
    P, R = H.P[0], H.R[0]
    nf, ncs = tuple(n - 1 for n in nc), tuple(n // 2 - 1 for n in nc)  # free nodes per direction, lexicographic (x fastest)

    def coords(j, dims):
        out = []
        for n in dims:
            out.append(j % n)
            j //= n
        return out

    colour_c = lambda j: sum((c % 2) << k for k, c in enumerate(coords(j, ncs)))
    colour_f = lambda j: sum((c % 3) * 3 ** k for k, c in enumerate(coords(j, nf)))
    mulP = lambda y, x: y.__setitem__(slice(None), P @ x)
    mulR = lambda y, x: y.__setitem__(slice(None), R @ x)
    Pp = gsb200.explicit_transfer(mulP, P.shape[1], P.shape[0], colour_c, 8)
    Rp = gsb200.explicit_transfer(mulR, R.shape[1], R.shape[0], colour_f, 27)
    assert abs(Pp - P).max() == 0.0 and Pp.nnz == P.nnz
    assert abs(Rp - R).max() == 0.0 and abs(Rp - Pp.T).max() <= 1e-15
    with pytest.raises(ValueError):
        gsb200.explicit_transfer(mulP, P.shape[1], P.shape[0], lambda j: 0, 1)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (CPU oracle port) prints one JSON line with the contract's keys"""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cells-per-gpu", "16"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None


def _sell_plan(A, n_ghost=0, blocks=1, sort=-1):
    A = sp.csr_matrix(A)
    A.sort_indices()
    rp, col = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    out = np.zeros(12, dtype=np.int64)
    nsl = -(-A.shape[0] // 32)
    prow, plen = np.full(nsl * 32, -9, dtype=np.int32), np.zeros(nsl * 32, dtype=np.int32)
    pmask = np.zeros(nsl * 32, dtype=np.int32)
    words = np.zeros(max(1, int(A.nnz) + 64 * nsl), dtype=np.int32)  # generous: one word per (slice, k) pair
    rc = gsb200._lib.lib().gsb_diag_sell_plan(A.shape[0], A.shape[1] - n_ghost, n_ghost, rp.ctypes.data, col.ctypes.data, blocks, sort,
                                              out.ctypes.data, prow.ctypes.data, plen.ctypes.data, pmask.ctypes.data, words.ctypes.data)
    assert rc == 0
    keys = ("ok", "bs", "sorted", "n_brows", "n_slices", "blocks", "sum_blocks", "bnd_slices", "explicit_lines", "pairs",
            "aligned_slices", "triple_slices")
    d = dict(zip(keys, (int(v) for v in out)))
    d["col_words"] = words[: d["pairs"]].copy()
    d["col"] = col
    ns = d["n_slices"] * 32
    d["pos_mask"] = pmask[:ns]
    return d, prow[:ns], plen[:ns], rp


def _check_col_words(d, prow, plen, rp):
    """the layout reproduces the CSR matrix: the valid slots of a position (mask bits, or k < length in slices wider
    than 32) are as many as its blocks; the q-th valid slot of an AFFINE (slice, k) pair is the q-th block column of
    the row (word + lane); explicit lines are numbered 0.. in (slice, k) order.  Returns the affine fraction."""
    bs, col, words, pmask = d["bs"], d["col"], d["col_words"], d["pos_mask"]
    assert d["pairs"] * 32 == d["blocks"]
    nsl = d["n_slices"]
    # slice widths from the words array cannot be read back directly: recompute them from the masks / lengths
    width = np.zeros(nsl, dtype=np.int64)
    # a slice is wide (> 32 slots) iff some length exceeds 32; else its width is the highest mask bit + 1
    for sl in range(nsl):
        ln = plen[sl * 32:(sl + 1) * 32]
        if ln.max(initial=0) > 32:
            width[sl] = ln.max()
        else:
            m = int(np.bitwise_or.reduce(pmask[sl * 32:(sl + 1) * 32].astype(np.int64) & 0xFFFFFFFF))
            width[sl] = m.bit_length()
    assert width.sum() == d["pairs"], (width.sum(), d["pairs"])
    off = np.concatenate([[0], np.cumsum(width)])
    nexp = 0
    for sl in range(nsl):
        wide = width[sl] > 32
        q = [0] * 32
        for k in range(width[sl]):
            w = int(words[off[sl] + k])
            if w < 0:
                assert ~w == nexp
                nexp += 1
            for l in range(32):
                pos = sl * 32 + l
                b, ln = int(prow[pos]), int(plen[pos])
                if b < 0:
                    continue
                on = (k < int(pmask[pos])) if wide else bool((int(pmask[pos]) >> k) & 1)
                if on:
                    if w >= 0:
                        assert col[rp[b * bs] + q[l] * bs] // bs == w + l
                    q[l] += 1
        for l in range(32):
            if prow[sl * 32 + l] >= 0:
                assert q[l] == plen[sl * 32 + l]
    assert nexp == d["explicit_lines"]
    return 1.0 - nexp / max(1, d["pairs"])


def test_sell_plan_scalar_q1_is_unsorted_and_tight():
    """Q1 Poisson: rows of one mesh line have the same length => no sorting, < 3 % padding, block size 1"""
    d, prow, plen, rp = _sell_plan(fem.poisson((24, 20, 12)).A)
    assert d["ok"] and d["bs"] == 1 and not d["sorted"]
    assert d["blocks"] <= 1.06 * d["sum_blocks"]
    n = d["n_brows"]
    assert np.array_equal(prow[:n], np.arange(n)) and (prow[n:] == -1).all()
    assert np.array_equal(plen[:n], np.diff(rp))
    # mesh-ordered stencil: slices are diagonal-aligned, (almost) every column word is affine -- also in the slices
    # that cross mesh-line ends (the first slices keep explicit lines for the slots that reach before column 0)
    assert d["aligned_slices"] >= 0.95 * d["n_slices"]
    assert _check_col_words(d, prow, plen, rp) > 0.9
    # the x-neighbours (c-1, c, c+1) of the 27-point stencil: slots in runs of three consecutive columns
    assert d["triple_slices"] >= 0.9 * d["n_slices"]
    w = d["col_words"]
    assert d["pairs"] % 3 == 0 or d["triple_slices"] < d["n_slices"]
    # without alignment the padding is a little smaller and a good part of the words is explicit
    assert d["blocks"] <= 1.06 * d["sum_blocks"]


def test_sell_plan_q2_elasticity_detects_3x3_blocks_and_sorts_rows():
    """Q2 vector-valued elasticity (C4): node-major dofs => aligned 3x3 blocks; 125/75/45/27-node stencils alternate
    along a mesh line => rows are sorted by length inside windows of 256 block rows, which brings the padding of
    the 32-row slices from ~25 % down to a few %; the permutation is a bijection that stays inside its window"""
    sysm = fem.elasticity((6, 6, 6), order=2)
    A = sysm.A
    d, prow, plen, rp = _sell_plan(A)
    assert d["ok"] and d["bs"] == 3 and d["sorted"]
    nb = A.shape[0] // 3
    assert d["n_brows"] == nb and d["sum_blocks"] * 9 == A.nnz
    assert d["blocks"] <= 1.12 * d["sum_blocks"]
    d0, _, _, _ = _sell_plan(A, sort=0)
    assert d0["bs"] == 3 and not d0["sorted"] and d0["blocks"] > 1.15 * d0["sum_blocks"]
    real = prow[prow >= 0]
    assert np.array_equal(np.sort(real), np.arange(nb))
    pos = np.flatnonzero(prow >= 0)
    assert (pos // 256 == real // 256).all()
    assert np.array_equal(plen[pos], np.diff(rp)[real * 3] // 3)
    for w in range(0, len(prow), 256):  # descending inside every window
        assert (np.diff(plen[w:w + 256]) <= 0).all()
    # block detection can be switched off: same matrix as scalar rows
    d1, _, _, _ = _sell_plan(A, blocks=0)
    assert d1["bs"] == 1 and d1["sum_blocks"] == A.nnz
    _check_col_words(d, prow, plen, rp)


def test_sell_plan_rejects_false_block_structure():
    """a matrix whose size is a multiple of 3 but whose sparsity is not made of aligned 3x3 blocks keeps block size 1;
    2-component vector problems (Stokes velocity block) get 2x2 blocks"""
    d, _, _, _ = _sell_plan(fem.poisson((10, 10)).A)  # 81 rows
    assert d["bs"] == 1
    st = fem.stokes_cavity((8, 8))
    d, _, _, _ = _sell_plan(st["A"])
    assert d["ok"] and d["bs"] == 2
    # one entry removed from one row of a block row breaks the structure
    A = fem.elasticity((2, 2, 2), order=2).A.tolil()
    A[4, A.rows[4][0]] = 0
    A = sp.csr_matrix(A)
    A.eliminate_zeros()
    d, _, _, _ = _sell_plan(A)
    assert d["bs"] == 1


def test_sell_plan_boundary_slices_follow_ghost_columns():
    lp = synth.make_level_part((16, 8, 8), (2, 1, 1), 0)
    rp, col, val, b = synth.poisson_rows(lp)
    A = synth.to_scipy(rp, col, val, lp.n_own + lp.n_ghost)
    d, prow, plen, _ = _sell_plan(A, n_ghost=lp.n_ghost)
    has_ghost = np.array([col[rp[i + 1] - 1] >= lp.n_own for i in range(lp.n_own)])
    expect = {int(i) // 32 for i in np.flatnonzero(has_ghost)} if not d["sorted"] else None
    assert d["ok"] and d["bnd_slices"] > 0
    if expect is not None:
        assert d["bnd_slices"] == len(expect)


def test_sell_plan_prolongation_short_rows():
    """prolongations have 1/2/4/8 entries per row: sorting keeps the padding small"""
    H = fem.poisson_hierarchy((16, 16, 16), 2)
    d, _, _, _ = _sell_plan(H.P[0])
    assert d["ok"] and d["bs"] == 1
    assert d["blocks"] <= 1.6 * d["sum_blocks"]
