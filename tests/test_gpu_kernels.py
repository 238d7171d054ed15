"""GPU parity tests of the device primitives, through the C ABI, against the CPU oracle.

Bar: the row kernels (SpMV, residual, fused Jacobi-Richardson sweep, transfers) accumulate each row
in ascending column order without FMA contraction, exactly like the oracle (and Julia's
SparseArrays mul!) => BIT-EXACT with one lane per row.  Reductions (dot/norm) use a different
summation tree => relative 1e-14.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import fem
from oracle import linalg as ola
from oracle import solvers as OS
from util import dev_matrix, dev_vec

pytestmark = pytest.mark.gpu


def _rand(n, seed):
    return np.random.default_rng(seed).standard_normal(n)


KERNELS = ["vector", "sell", "sell_sorted"]


def _select(ctx, kernel):
    """row-kernel family for the matrices created next: CSR fallback kernel, block-SELL-32 in natural row order
    (sorting forced off), block-SELL-32 with the rows sorted by length inside windows of 256 (forced on)"""
    ctx.set_option("spmv", "vector" if kernel == "vector" else "auto")
    ctx.set_option("sell_sort", "1" if kernel == "sell_sorted" else ("0" if kernel == "sell" else "auto"))


def _reset(ctx):
    for k, v in (("spmv", "auto"), ("sell_sort", "auto"), ("block", "1"), ("keep_csr_max_nnz", "16777216"), ("fuse_smoother", "1")):
        ctx.set_option(k, v)


def _format(gsb, A):
    import ctypes

    kind, bs, srt = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    ent, byt = ctypes.c_int64(), ctypes.c_int64()
    gsb._lib.check(gsb._lib.lib().gsb_mat_format(A.h, ctypes.byref(kind), ctypes.byref(bs), ctypes.byref(srt), ctypes.byref(ent),
                                                 ctypes.byref(byt)))
    return dict(kind=kind.value, bs=bs.value, sorted=bool(srt.value), entries=ent.value, bytes=byt.value)


@pytest.mark.parametrize("nc", [(8, 8), (33, 17), (8, 8, 8), (20, 24, 16)])
@pytest.mark.parametrize("kernel", KERNELS)
def test_spmv_bit_exact(gsb, ctx, nc, kernel):
    _select(ctx, kernel)
    try:
        sysm = fem.poisson(nc)
        A = dev_matrix(gsb, ctx, sysm.A)
        Ao = ola.CSR(sysm.A)
        x = _rand(sysm.A.shape[1], 1)
        y0 = _rand(sysm.A.shape[0], 2)
        xd, yd = dev_vec(gsb, A, x), dev_vec(gsb, A, y0, domain=False)
        gsb.mul_(yd, A, xd)
        yo = np.zeros_like(y0)
        ola.mul(yo, Ao, x)
        assert np.array_equal(yd.get(), yo)
        for alpha, beta in [(-0.75, 1.0), (2.5, -0.5), (1.0, 1.0)]:
            yd.set(y0)
            gsb.mul_(yd, A, xd, alpha, beta)
            yo = y0.copy()
            ola.mul5(yo, Ao, x, alpha, beta)
            assert np.array_equal(yd.get(), yo), (alpha, beta)
        f = _format(gsb, A)
        assert f["kind"] == 1 and (kernel == "vector" or f["sorted"] == (kernel == "sell_sorted"))
    finally:
        _reset(ctx)


@pytest.mark.parametrize("kernel", ["sell", "sell_sorted"])
@pytest.mark.parametrize("nc", [(64, 64, 64), (96, 80, 72), (700, 600)])
def test_sell_many_slices_bit_exact(gsb, ctx, nc, kernel):
    """sizes with thousands of slices / sorting windows, matrices whose CSR arrays are released after the device-side
    conversion: SpMV, residual and fused sweeps stay bit-identical to the oracle, run after run"""
    from gsb200 import synth
    from util import host_to_scipy

    _select(ctx, kernel)
    ctx.set_option("keep_csr_max_nnz", "1000")
    try:
        hh = synth.poisson_hierarchy_host(nc, 1)
        n = hh.levels[0].n_own
        As = host_to_scipy(hh.A[0], n)
        A, Ao = dev_matrix(gsb, ctx, As), ola.CSR(As)
        x = _rand(n, 21)
        xd, yd = dev_vec(gsb, A, x), dev_vec(gsb, A, domain=False)
        yo = np.zeros(n)
        ola.mul(yo, Ao, x)
        for _ in range(3):
            gsb.mul_(yd, A, xd)
            assert np.array_equal(yd.get(), yo)
        s = gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), 4, 2.0 / 3.0)
        ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
        x0, r0 = _rand(n, 22), _rand(n, 23)
        xd, rd = dev_vec(gsb, A, x0), dev_vec(gsb, A, r0)
        gsb.solve_(xd, ns, rd)
        so = OS.RichardsonSmoother(OS.JacobiLinearSolver(), 4, 2.0 / 3.0)
        xo, ro = x0.copy(), r0.copy()
        OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, Ao), Ao), ro)
        assert np.array_equal(xd.get(), xo) and np.array_equal(rd.get(), ro)
        # numerical_setup! on a matrix without CSR mirror: values and inverse diagonal are refreshed
        A.update_values(3.0 * As.data)
        gsb.numerical_setup_(ns, A)
        gsb.mul_(yd, A, dev_vec(gsb, A, x))
        ola.mul(yo, ola.CSR(3.0 * As), x)
        assert np.array_equal(yd.get(), yo)
        xd, rd = dev_vec(gsb, A, x0), dev_vec(gsb, A, r0)
        gsb.solve_(xd, ns, rd)
        Ao3 = ola.CSR(3.0 * As)
        xo, ro = x0.copy(), r0.copy()
        OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, Ao3), Ao3), ro)
        assert np.array_equal(xd.get(), xo) and np.array_equal(rd.get(), ro)
    finally:
        _reset(ctx)


@pytest.mark.parametrize("problem", ["elasticity_q2_3d", "stokes_velocity_q2_2d", "poisson_q2_3d"])
@pytest.mark.parametrize("block", [1, 0])
def test_block_sell_bit_exact(gsb, ctx, problem, block):
    """long-row matrices (Q2 elements: 27..125-node stencils): 3x3 / 2x2 DOF blocks with one lane per block row and
    rows sorted by length keep the sequential ascending-column order of every row => bit-identical to the oracle's
    CSR loop for SpMV (3- and 5-argument), residual, fused Jacobi-Richardson sweeps and the transfer kernels"""
    ctx.set_option("block", block)
    try:
        if problem == "elasticity_q2_3d":
            H = fem.elasticity_hierarchy((6, 6, 6), 2, order=2)
            bs = 3
        elif problem == "stokes_velocity_q2_2d":
            st = fem.stokes_cavity((24, 24), nlevels=2)
            H = fem.Hierarchy([fem.AffineSystem(A=m, b=None, M=None, xstar=None, free=None, grid=None) for m in st["mats"]], st["P"], st["R"])
            bs = 2
        else:
            H = fem.poisson_hierarchy((8, 8, 8), 2, order=2)
            bs = 1
        As = H.mats[0]
        A, Ao = dev_matrix(gsb, ctx, As), ola.CSR(As)
        f = _format(gsb, A)
        assert f["kind"] == 1 and f["bs"] == (bs if block else 1) and f["sorted"]
        assert f["entries"] <= 1.25 * As.nnz
        n = As.shape[0]
        x, y0 = _rand(n, 41), _rand(n, 42)
        xd, yd = dev_vec(gsb, A, x), dev_vec(gsb, A, y0, domain=False)
        gsb.mul_(yd, A, xd)
        yo = np.zeros(n)
        ola.mul(yo, Ao, x)
        assert np.array_equal(yd.get(), yo)
        for alpha, beta in [(-0.75, 1.0), (2.5, -0.5)]:
            yd.set(y0)
            gsb.mul_(yd, A, xd, alpha, beta)
            yo = y0.copy()
            ola.mul5(yo, Ao, x, alpha, beta)
            assert np.array_equal(yd.get(), yo), (alpha, beta)
        for niter in (1, 6):
            s = gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), niter, 2.0 / 3.0)
            ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
            x0, r0 = _rand(n, 43), _rand(n, 44)
            xs, rs = dev_vec(gsb, A, x0), dev_vec(gsb, A, r0)
            gsb.solve_(xs, ns, rs)
            so = OS.RichardsonSmoother(OS.JacobiLinearSolver(), niter, 2.0 / 3.0)
            xo, ro = x0.copy(), r0.copy()
            OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, Ao), Ao), ro)
            assert np.array_equal(xs.get(), xo) and np.array_equal(rs.get(), ro), niter
        # transfers of the same hierarchy (prolongation rows of 1..27 entries, restriction rows up to 125)
        for T in (H.P[0], H.R[0]):
            Td, To = dev_matrix(gsb, ctx, T), ola.CSR(T)
            xt = _rand(T.shape[1], 45)
            xtd, ytd = dev_vec(gsb, Td, xt), dev_vec(gsb, Td, domain=False)
            gsb.mul_(ytd, Td, xtd)
            yt = np.zeros(T.shape[0])
            ola.mul(yt, To, xt)
            assert np.array_equal(ytd.get(), yt)
    finally:
        _reset(ctx)


def test_spmv_csc_int64_one_based_upload(gsb, ctx):
    """Julia's default local matrix: SparseMatrixCSC{Float64,Int64}, 1-based."""
    sysm = fem.poisson((9, 7))
    Acsc = sp.csc_matrix(sysm.A)
    A = gsb.SparseMatrix(ctx, Acsc.shape[0], Acsc.shape[1], 0, Acsc.indptr.astype(np.int64) + 1,
                         Acsc.indices.astype(np.int64) + 1, Acsc.data, fmt="csc", index_base=1)
    x = _rand(Acsc.shape[1], 3)
    xd, yd = dev_vec(gsb, A, x), dev_vec(gsb, A, domain=False)
    gsb.mul_(yd, A, xd)
    yo = np.zeros(Acsc.shape[0])
    ola.mul(yo, ola.CSR(sysm.A), x)
    assert np.array_equal(yd.get(), yo)
    # numerical_setup!-style value update in the caller's (CSC) order
    A.update_values(2.0 * Acsc.data)
    gsb.mul_(yd, A, xd)
    ola.mul(yo, ola.CSR(2.0 * sysm.A), x)
    assert np.array_equal(yd.get(), yo)


def test_spmv_unsorted_rows_and_empty_rows(gsb, ctx):
    rng = np.random.default_rng(0)
    n = 300
    A = sp.random(n, n, density=0.03, random_state=5, format="csr")
    A = A.tolil()
    A[7, :] = 0  # empty row
    A = sp.csr_matrix(A)
    A.eliminate_zeros()
    A.sort_indices()
    # shuffle entries inside rows
    ptr, idx, val = A.indptr.copy(), A.indices.copy(), A.data.copy()
    for i in range(n):
        p = rng.permutation(ptr[i + 1] - ptr[i]) + ptr[i]
        idx[ptr[i]:ptr[i + 1]], val[ptr[i]:ptr[i + 1]] = idx[p], val[p]
    Ad = gsb.SparseMatrix(ctx, n, n, 0, ptr, idx, val)
    x = _rand(n, 4)
    xd, yd = dev_vec(gsb, Ad, x), dev_vec(gsb, Ad, domain=False)
    gsb.mul_(yd, Ad, xd)
    yo = np.zeros(n)
    ola.mul(yo, ola.CSR(A), x)
    assert np.array_equal(yd.get(), yo)
    assert yd.get()[7] == 0.0


@pytest.mark.parametrize("G_rows", [60, 200])  # average row length; the CSR fallback kernel uses 4 / 16 lanes per row
@pytest.mark.parametrize("kernel", ["vector", "sell_auto"])
def test_spmv_long_rows(gsb, ctx, G_rows, kernel):
    _select(ctx, kernel)
    try:
        n = 4000
        A = sp.random(n, n, density=G_rows / n, random_state=11, format="csr")
        A.sort_indices()
        Ad = dev_matrix(gsb, ctx, A)
        x = _rand(n, 6)
        xd, yd = dev_vec(gsb, Ad, x), dev_vec(gsb, Ad, domain=False)
        gsb.mul_(yd, Ad, xd)
        yo = np.zeros(n)
        ola.mul(yo, ola.CSR(A), x)
        if kernel == "sell_auto":
            # random row lengths: sorted block-SELL, one lane per row => still the sequential order, bit for bit
            assert _format(gsb, Ad)["kind"] == 1 and _format(gsb, Ad)["sorted"]
            assert np.array_equal(yd.get(), yo)
        else:
            # several lanes per row => different summation tree: tolerance, not bit equality
            assert np.allclose(yd.get(), yo, rtol=0, atol=1e-13 * np.abs(A).dot(np.abs(x)).max())
    finally:
        _reset(ctx)


def test_blas1(gsb, ctx):
    n = 100003
    A = dev_matrix(gsb, ctx, sp.identity(n, format="csr"))
    a, b = _rand(n, 7), _rand(n, 8)
    ad, bd, zd = dev_vec(gsb, A, a), dev_vec(gsb, A, b), dev_vec(gsb, A)
    assert abs(gsb.dot(ad, bd) - np.dot(a, b)) <= 1e-13 * np.dot(np.abs(a), np.abs(b))
    assert abs(gsb.norm(ad) - np.linalg.norm(a)) <= 1e-14 * np.linalg.norm(a)
    gsb.axpby_(zd, 1.0, ad, -0.3, bd)
    assert np.array_equal(zd.get(), a + (-0.3) * b)  # z = x - s*y, two roundings
    gsb.axpby_(zd, 2.0, ad, 1.0, bd)
    assert np.array_equal(zd.get(), 2.0 * a + b)
    # reductions are deterministic run to run
    assert gsb.dot(ad, bd) == gsb.dot(ad, bd)
    zd.fill(2.5)
    assert np.array_equal(zd.get(), np.full(n, 2.5))
    zd.fill(0.0)
    assert not zd.get().any()


@pytest.mark.parametrize("nc", [(16, 16), (12, 12, 12)])
@pytest.mark.parametrize("kernel", KERNELS)
def test_richardson_jacobi_bit_exact(gsb, ctx, nc, kernel):
    """fused Jacobi-Richardson sweeps == reference statement sequence (RichardsonSmoothers.jl:84-98)"""
    _select(ctx, kernel)
    try:
        sysm = fem.poisson(nc)
        A = dev_matrix(gsb, ctx, sysm.A)
        Ao = ola.CSR(sysm.A)
        for niter, omega in [(1, 1.0), (5, 2.0 / 3.0), (10, 2.0 / 3.0)]:
            x0, r0 = _rand(Ao.shape[0], 9), _rand(Ao.shape[0], 10)
            s = gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), niter, omega)
            ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
            xd, rd = dev_vec(gsb, A, x0), dev_vec(gsb, A, r0)
            gsb.solve_(xd, ns, rd)
            so = OS.RichardsonSmoother(OS.JacobiLinearSolver(), niter, omega)
            nso = OS.numerical_setup(OS.symbolic_setup(so, Ao), Ao)
            xo, ro = x0.copy(), r0.copy()
            OS.solve_(xo, nso, ro)
            assert np.array_equal(xd.get(), xo), (niter, omega)
            assert np.array_equal(rd.get(), ro), (niter, omega)
            # unfused device path (generic inner solver sequence) gives the same bits
            ctx.set_option("fuse_smoother", "0")
            xd2, rd2 = dev_vec(gsb, A, x0), dev_vec(gsb, A, r0)
            gsb.solve_(xd2, ns, rd2)
            ctx.set_option("fuse_smoother", "1")
            assert np.array_equal(xd2.get(), xo) and np.array_equal(rd2.get(), ro)
    finally:
        _reset(ctx)


def test_linear_solver_from_smoother(gsb, ctx):
    sysm = fem.poisson((10, 10))
    A, Ao = dev_matrix(gsb, ctx, sysm.A), ola.CSR(sysm.A)
    b = _rand(Ao.shape[0], 12)
    s = gsb.LinearSolverFromSmoother(gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), 5, 2.0 / 3.0))
    ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
    xd, bd = dev_vec(gsb, A, _rand(Ao.shape[0], 13)), dev_vec(gsb, A, b)
    gsb.solve_(xd, ns, bd)
    so = OS.LinearSolverFromSmoother(OS.RichardsonSmoother(OS.JacobiLinearSolver(), 5, 2.0 / 3.0))
    xo = _rand(Ao.shape[0], 13)
    OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, Ao), Ao), b)
    assert np.array_equal(xd.get(), xo)
    assert np.array_equal(bd.get(), b)  # b must not be mutated


def test_transfers_bit_exact(gsb, ctx):
    H = fem.poisson_hierarchy((16, 16, 16), 2)
    P, R = dev_matrix(gsb, ctx, H.P[0]), dev_matrix(gsb, ctx, H.R[0])
    xc, xf = _rand(H.P[0].shape[1], 14), _rand(H.P[0].shape[0], 15)
    xcd, yfd = dev_vec(gsb, P, xc), dev_vec(gsb, P, domain=False)
    gsb.mul_(yfd, P, xcd)
    yo = np.zeros(H.P[0].shape[0])
    ola.mul(yo, ola.CSR(H.P[0]), xc)
    assert np.array_equal(yfd.get(), yo)
    xfd, ycd = dev_vec(gsb, R, xf), dev_vec(gsb, R, domain=False)
    gsb.mul_(ycd, R, xfd)
    yo = np.zeros(H.R[0].shape[0])
    ola.mul(yo, ola.CSR(H.R[0]), xf)
    assert np.array_equal(ycd.get(), yo)


@pytest.mark.parametrize("nc", [(6, 6), (8, 8, 8), (16, 16, 16)])
def test_dense_coarse_solver(gsb, ctx, nc):
    sysm = fem.poisson(nc)
    A = dev_matrix(gsb, ctx, sysm.A)
    ns = gsb.numerical_setup(gsb.symbolic_setup(gsb.LUSolver(), A), A)
    b = _rand(sysm.A.shape[0], 16)
    xd, bd = dev_vec(gsb, A), dev_vec(gsb, A, b)
    gsb.solve_(xd, ns, bd)
    x = xd.get()
    xo = np.zeros_like(b)
    OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(OS.LUSolver(), ola.CSR(sysm.A)), ola.CSR(sysm.A)), b)
    assert np.linalg.norm(x - xo) <= 1e-12 * np.linalg.norm(xo)
    assert np.linalg.norm(sysm.A @ x - b) <= 1e-12 * np.linalg.norm(b)


def test_dense_coarse_solver_nonsymmetric_needs_pivoting(gsb, ctx):
    n = 64
    rng = np.random.default_rng(3)
    M = rng.standard_normal((n, n))
    M[0, 0] = 0.0  # zero leading pivot
    A = dev_matrix(gsb, ctx, sp.csr_matrix(M))
    ns = gsb.numerical_setup(gsb.symbolic_setup(gsb.LUSolver(), A), A)
    b = _rand(n, 17)
    xd, bd = dev_vec(gsb, A), dev_vec(gsb, A, b)
    gsb.solve_(xd, ns, bd)
    assert np.linalg.norm(M @ xd.get() - b) <= 1e-10 * np.linalg.norm(b)


def test_errors_are_reported_not_crashes(gsb, ctx):
    sysm = fem.poisson((4, 4))
    A = dev_matrix(gsb, ctx, sysm.A)
    x = dev_vec(gsb, A)
    with pytest.raises(gsb.GSBError):
        gsb.mul_(x, A, x)  # aliasing
    short = gsb.Vector(ctx, 3)
    with pytest.raises(gsb.GSBError):
        gsb.mul_(x, A, short)
    gmg = gsb.GMGLinearSolver([A, A], [A], [A])
    ns = gsb.numerical_setup(gsb.symbolic_setup(gmg, A), A)
    with pytest.raises(gsb.GSBError):  # GMGLinearSolvers.jl:249-258
        gsb.numerical_setup_(ns, A)
