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


@pytest.mark.parametrize("nc", [(8, 8), (33, 17), (8, 8, 8), (20, 24, 16)])
@pytest.mark.parametrize("kernel", ["vector", "stream", "sell"])
def test_spmv_bit_exact(gsb, ctx, nc, kernel):
    ctx.set_option("spmv", kernel)
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
    finally:
        ctx.set_option("spmv", "auto")


@pytest.mark.parametrize("stream_kernel", ["ws", "v1"])
@pytest.mark.parametrize("nc", [(64, 64, 64), (96, 80, 72), (700, 600)])
def test_stream_ring_wraparound_bit_exact(gsb, ctx, nc, stream_kernel):
    """sizes at which every persistent CTA wraps its shared-memory ring several times and the
    consumer warps drift apart: SpMV, residual and fused sweeps stay bit-identical to the oracle"""
    from gsb200 import synth
    from util import host_to_scipy

    ctx.set_option("spmv", "stream")
    ctx.set_option("stream_kernel", stream_kernel)
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
    finally:
        ctx.set_option("spmv", "auto")
        ctx.set_option("stream_kernel", "ws")


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


@pytest.mark.parametrize("G_rows", [60, 200])  # average row length selects 4 / 16 lanes per row
@pytest.mark.parametrize("kernel", ["vector", "stream", "sell"])
def test_spmv_long_rows(gsb, ctx, G_rows, kernel):
    ctx.set_option("spmv", kernel)
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
        # several lanes per row => different summation tree: tolerance, not bit equality
        assert np.allclose(yd.get(), yo, rtol=0, atol=1e-13 * np.abs(A).dot(np.abs(x)).max())
    finally:
        ctx.set_option("spmv", "auto")


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
@pytest.mark.parametrize("kernel", ["vector", "stream", "sell"])
def test_richardson_jacobi_bit_exact(gsb, ctx, nc, kernel):
    """fused Jacobi-Richardson sweeps == reference statement sequence (RichardsonSmoothers.jl:84-98)"""
    ctx.set_option("spmv", kernel)
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
        ctx.set_option("spmv", "auto")
        ctx.set_option("fuse_smoother", "1")


@pytest.mark.parametrize("nc,force", [((20, 24, 16), True), ((33, 17), True), ((96, 96, 80), False)])
@pytest.mark.parametrize("stages", [2, 3, 5, 10])
def test_richardson_l2_pipelined_sweeps_bit_exact(gsb, ctx, nc, force, stages):
    """S sweeps fused in one launch (inter-CTA dependency pipeline through L2) == the reference sequence, bit for bit"""
    from gsb200 import synth
    from util import host_to_scipy

    ctx.set_option("pipe_stages", stages)
    ctx.set_option("pipe_min_chunks", 1 if force else 2368)
    try:
        hh = synth.poisson_hierarchy_host(nc, 1)
        n = hh.levels[0].n_own
        As = host_to_scipy(hh.A[0], n)
        A, Ao = dev_matrix(gsb, ctx, As), ola.CSR(As)
        for niter in (10, 3):
            s = gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), niter, 2.0 / 3.0)
            ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
            x0, r0 = _rand(n, 31), _rand(n, 32)
            xd, rd = dev_vec(gsb, A, x0), dev_vec(gsb, A, r0)
            l0 = ctx.launch_count()
            gsb.solve_(xd, ns, rd)
            assert ctx.launch_count() - l0 == 1 + -(-niter // stages)  # prologue + ceil(niter/S) launches
            so = OS.RichardsonSmoother(OS.JacobiLinearSolver(), niter, 2.0 / 3.0)
            xo, ro = x0.copy(), r0.copy()
            OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, Ao), Ao), ro)
            assert np.array_equal(xd.get(), xo) and np.array_equal(rd.get(), ro)
            # run to run reproducible
            xd2, rd2 = dev_vec(gsb, A, x0), dev_vec(gsb, A, r0)
            gsb.solve_(xd2, ns, rd2)
            assert np.array_equal(xd2.get(), xo) and np.array_equal(rd2.get(), ro)
    finally:
        ctx.set_option("pipe_stages", 1)
        ctx.set_option("pipe_min_chunks", 2368)


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
