"""GPU parity tests of the solver layer (through the C ABI) against the CPU oracle, plus the
reference's own known-answer assertions (KrylovTests.jl:21-25 E < 1e-6, SmoothersTests.jl:44 E < 1e-8).

Tolerance (BASELINE.json north_star): identical iteration counts (+-1) and a relative residual
history within 1e-10 of the CPU solve on the same assembled system.
"""
import numpy as np
import pytest

import gsb200
from gsb200 import synth
from oracle import fem
from oracle import linalg as ola
from oracle import solvers as OS
from util import dev_matrix, dev_vec, oracle_hierarchy, rel_hist_diff

pytestmark = pytest.mark.gpu

HIST_TOL = 1e-10


def _run_pair(gsb, ctx, sysm, make_gpu, make_ref, x0=None):
    A, Ao = dev_matrix(gsb, ctx, sysm.A), ola.CSR(sysm.A)
    sg, so = make_gpu(gsb), make_ref(OS)
    ns = gsb.numerical_setup(gsb.symbolic_setup(sg, A), A)
    xd = dev_vec(gsb, A, x0)
    bd = dev_vec(gsb, A, sysm.b)
    gsb.solve_(xd, ns, bd)
    nso = OS.numerical_setup(OS.symbolic_setup(so, Ao), Ao)
    xo = np.zeros(Ao.shape[1]) if x0 is None else x0.copy()
    OS.solve_(xo, nso, sysm.b)
    return sg, so, xd.get(), xo


KRYLOV_CASES = {
    # KrylovTests.jl:66-93
    "gmres40_PrPl": lambda S: S.GMRESSolver(40, Pr=S.JacobiLinearSolver(), Pl=S.JacobiLinearSolver(), rtol=1e-8),
    "gmres10": lambda S: S.GMRESSolver(10, rtol=1e-8),
    "gmres10_restart": lambda S: S.GMRESSolver(10, restart=True, rtol=1e-8),
    "fgmres10": lambda S: S.FGMRESSolver(10, S.JacobiLinearSolver(), rtol=1e-8),
    "fgmres10_restart": lambda S: S.FGMRESSolver(10, S.JacobiLinearSolver(), restart=True, rtol=1e-8),
    "cg": lambda S: S.CGSolver(rtol=1e-8),
    "pcg": lambda S: S.CGSolver(S.JacobiLinearSolver(), rtol=1e-8),
    "fpcg": lambda S: S.CGSolver(S.JacobiLinearSolver(), flexible=True, rtol=1e-8),
    "minres": lambda S: S.MINRESSolver(Pl=S.JacobiLinearSolver(), rtol=1e-8),
    # SmoothersTests.jl:46-56
    "cg_richardson": lambda S: S.CGSolver(S.LinearSolverFromSmoother(S.RichardsonSmoother(S.JacobiLinearSolver(), 5, 2.0 / 3.0)), rtol=1e-8),
    "gmres_identity": lambda S: S.GMRESSolver(20, Pl=S.IdentitySolver(), rtol=1e-8),
}


@pytest.mark.parametrize("nc", [(8, 8), (8, 8, 8)])
@pytest.mark.parametrize("case", sorted(KRYLOV_CASES))
def test_krylov_known_answer_and_parity(gsb, ctx, nc, case):
    sysm = fem.poisson(nc)
    mk = KRYLOV_CASES[case]
    sg, so, xg, xo = _run_pair(gsb, ctx, sysm, mk, mk)
    # reference known-answer assertion
    E = fem.l2_error_sq(sysm, xg)
    assert E < (1e-8 if case == "cg_richardson" else 1e-6)
    # parity with the oracle
    assert abs(sg.log.num_iters - so.log.num_iters) <= 1
    assert sg.log.flag == so.log.flag
    assert rel_hist_diff(sg.log.history(), so.log.history()) < HIST_TOL
    assert np.linalg.norm(xg - xo) <= 1e-9 * np.linalg.norm(xo)


def test_log_semantics(gsb, ctx):
    """ConvergenceLog: residuals has maxiter+1 slots, maxiter stops the loop (flag 2), atol wins at it 0"""
    sysm = fem.poisson((8, 8))
    A = dev_matrix(gsb, ctx, sysm.A)
    s = gsb.CGSolver(maxiter=3, rtol=1e-30, atol=0.0)
    ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
    gsb.solve_(dev_vec(gsb, A), ns, dev_vec(gsb, A, sysm.b))
    assert s.log.num_iters == 3 and s.log.flag == gsb200.api.SOLVER_DIVERGED_MAXITER
    assert s.log.residuals.shape[0] == 4
    s = gsb.CGSolver(atol=1e300)
    ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
    gsb.solve_(dev_vec(gsb, A), ns, dev_vec(gsb, A, sysm.b))
    assert s.log.num_iters == 0 and s.log.flag == gsb200.api.SOLVER_CONVERGED_ATOL


def test_gmres_basis_growth(gsb, ctx):
    """restart=false: the basis grows by m_add when full (GMRESSolvers.jl:154-157)"""
    sysm = fem.poisson((8, 8))
    mk = lambda S: S.GMRESSolver(3, m_add=2, rtol=1e-8)
    sg, so, xg, xo = _run_pair(gsb, ctx, sysm, mk, mk)
    assert sg.log.num_iters == so.log.num_iters
    assert rel_hist_diff(sg.log.history(), so.log.history()) < HIST_TOL


def test_nonzero_initial_guess_and_host_buffers(gsb, ctx):
    sysm = fem.poisson((8, 8, 8))
    x0 = np.random.default_rng(5).standard_normal(sysm.A.shape[0])
    mk = lambda S: S.CGSolver(S.JacobiLinearSolver(), rtol=1e-8)
    sg, so, xg, xo = _run_pair(gsb, ctx, sysm, mk, mk, x0=x0)
    assert rel_hist_diff(sg.log.history(), so.log.history()) < HIST_TOL
    # host-buffer entry point (e2e path): same answer as the device-vector call
    A = dev_matrix(gsb, ctx, sysm.A)
    s = mk(gsb)
    ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
    xh = x0.copy()
    b = sysm.b.copy()
    gsb.solve_(xh, ns, b)
    assert np.array_equal(xh, xg)
    assert np.array_equal(b, sysm.b)


def _gmg_pair(gsb, ctx, nc, nlev, cycle, outer):
    hh = synth.poisson_hierarchy_host(nc, nlev)
    dh = synth.upload_hierarchy(ctx, hh)
    mats, P, R = oracle_hierarchy(hh)
    sm_g = gsb.Fill(gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), 10, 2.0 / 3.0), nlev - 1)  # GMGTests.jl:52
    sm_o = [OS.RichardsonSmoother(OS.JacobiLinearSolver(), 10, 2.0 / 3.0)] * (nlev - 1)
    gmg_g = gsb.GMGLinearSolver(dh.A, dh.P, dh.R, pre_smoothers=sm_g, post_smoothers=sm_g, coarsest_solver=gsb.LUSolver(),
                                maxiter=1, mode="preconditioner", cycle_type=cycle)  # GMGTests.jl:109-117
    gmg_o = OS.GMGLinearSolver(mats, P, R, pre_smoothers=sm_o, post_smoothers=sm_o, coarsest_solver=OS.LUSolver(),
                               maxiter=1, mode="preconditioner", cycle_type=cycle)
    sg, so = outer(gsb, gmg_g), outer(OS, gmg_o)
    ns = gsb.numerical_setup(gsb.symbolic_setup(sg, dh.A[0]), dh.A[0])
    xd, bd = dev_vec(gsb, dh.A[0]), dev_vec(gsb, dh.A[0], hh.b)
    gsb.solve_(xd, ns, bd)
    nso = OS.numerical_setup(OS.symbolic_setup(so, mats[0]), mats[0])
    xo = np.zeros(mats[0].shape[0])
    OS.solve_(xo, nso, hh.b)
    return hh, sg, so, gmg_g, gmg_o, xd.get(), xo


@pytest.mark.parametrize("nc,nlev", [((16, 16), 3), ((32, 32), 4), ((16, 16, 16), 3), ((32, 32, 32), 4)])
def test_gmg_pcg_v_cycle_parity(gsb, ctx, nc, nlev):
    outer = lambda S, gmg: S.CGSolver(gmg, maxiter=20, atol=1e-14, rtol=1e-8)
    hh, sg, so, gmg_g, gmg_o, xg, xo = _gmg_pair(gsb, ctx, nc, nlev, "v_cycle", outer)
    assert sg.log.num_iters == so.log.num_iters
    assert rel_hist_diff(sg.log.history(), so.log.history()) < HIST_TOL
    # the preconditioner's own log (2 norms per application, quirk App. C.1) agrees too
    assert gmg_g.log.num_iters == gmg_o.log.num_iters == 1
    assert rel_hist_diff(gmg_g.log.history(), gmg_o.log.history()) < 1e-8
    assert np.linalg.norm(xg - xo) <= 1e-9 * np.linalg.norm(xo)
    assert np.linalg.norm(xg - synth.exact_solution(hh.levels[0])) <= 1e-6


@pytest.mark.parametrize("cycle", ["w_cycle", "f_cycle"])
def test_gmg_fgmres_w_f_cycles(gsb, ctx, cycle):
    outer = lambda S, gmg: S.FGMRESSolver(5, gmg, maxiter=20, atol=1e-14, rtol=1e-8)  # GMGTests.jl:121-122
    hh, sg, so, *_ = _gmg_pair(gsb, ctx, (16, 16, 16), 3, cycle, outer)
    assert sg.log.num_iters == so.log.num_iters
    assert rel_hist_diff(sg.log.history(), so.log.history()) < HIST_TOL


def test_gmg_with_iterative_coarsest_solver_is_not_graph_captured(gsb, ctx):
    """any LinearSolver is legal as coarsest solver (GMGLinearSolvers.jl:48-58).  A Krylov coarse solver reads scalars
    back on the host every iteration, so the preconditioner application must NOT be captured into a CUDA graph
    (capturable() guard): three consecutive solves on the same vector pair -- the pattern that triggers the capture --
    keep returning the oracle's history"""
    nc, nlev = (16, 16, 16), 3
    hh = synth.poisson_hierarchy_host(nc, nlev)
    dh = synth.upload_hierarchy(ctx, hh)
    mats, P, R = oracle_hierarchy(hh)
    mk = lambda S, m, p_, r_, sm: S.CGSolver(S.GMGLinearSolver(m, p_, r_, pre_smoothers=sm, post_smoothers=sm, maxiter=1,
                                                               coarsest_solver=S.CGSolver(S.JacobiLinearSolver(), maxiter=200, atol=1e-14, rtol=1e-12)),
                                             maxiter=20, atol=1e-14, rtol=1e-8)
    sg = mk(gsb, dh.A, dh.P, dh.R, gsb.Fill(gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), 10, 2.0 / 3.0), nlev - 1))
    so = mk(OS, mats, P, R, [OS.RichardsonSmoother(OS.JacobiLinearSolver(), 10, 2.0 / 3.0)] * (nlev - 1))
    xo = np.zeros(mats[0].shape[0])
    OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, mats[0]), mats[0]), hh.b)
    ns = gsb.numerical_setup(gsb.symbolic_setup(sg, dh.A[0]), dh.A[0])
    xd, bd = dev_vec(gsb, dh.A[0]), dev_vec(gsb, dh.A[0], hh.b)
    for _ in range(3):
        xd.fill(0.0)
        gsb.solve_(xd, ns, bd)
        assert sg.log.num_iters == so.log.num_iters
        assert rel_hist_diff(sg.log.history(), so.log.history()) < 1e-8  # inner CG to 1e-12: reduction-order noise amplified
    assert np.linalg.norm(xd.get() - xo) <= 1e-7 * np.linalg.norm(xo)


def test_gmg_solver_mode(gsb, ctx):
    hh = synth.poisson_hierarchy_host((16, 16), 3)
    dh = synth.upload_hierarchy(ctx, hh)
    mats, P, R = oracle_hierarchy(hh)
    g = gsb.GMGLinearSolver(dh.A, dh.P, dh.R, mode="solver", maxiter=20, rtol=1e-8)  # default smoothers, omega=1
    o = OS.GMGLinearSolver(mats, P, R, mode="solver", maxiter=20, rtol=1e-8)
    ns = gsb.numerical_setup(gsb.symbolic_setup(g, dh.A[0]), dh.A[0])
    xd = dev_vec(gsb, dh.A[0])
    gsb.solve_(xd, ns, dev_vec(gsb, dh.A[0], hh.b))
    xo = np.zeros(mats[0].shape[0])
    OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(o, mats[0]), mats[0]), hh.b)
    assert g.log.num_iters == o.log.num_iters
    assert rel_hist_diff(g.log.history(), o.log.history()) < HIST_TOL


def test_c2_full_size_properties(gsb, ctx):
    """BASELINE config C2 at full size (128^3, 4 levels): size-independent properties --
    convergence to rtol 1e-8, true residual agrees with the recurrence, discrete solution equals
    the nodal interpolant of u = x+y, SpMV linearity."""
    hh = synth.poisson_hierarchy_host((128, 128, 128), 4)
    dh = synth.upload_hierarchy(ctx, hh)
    sm = gsb.Fill(gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), 10, 2.0 / 3.0), 3)
    gmg = gsb.GMGLinearSolver(dh.A, dh.P, dh.R, pre_smoothers=sm, post_smoothers=sm, maxiter=1)
    s = gsb.CGSolver(gmg, maxiter=50, atol=1e-14, rtol=1e-8)
    A = dh.A[0]
    ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
    xd, bd = dev_vec(gsb, A), dev_vec(gsb, A, hh.b)
    gsb.solve_(xd, ns, bd)
    assert s.log.flag == gsb200.api.SOLVER_CONVERGED_RTOL and s.log.num_iters <= 12
    h = s.log.history()
    assert h[-1] / h[0] < 1e-8
    rd = dev_vec(gsb, A, domain=False)
    gsb.mul_(rd, A, xd)
    true_res = np.linalg.norm(hh.b - rd.get())
    assert abs(true_res - h[-1]) <= 1e-6 * h[0] * 1e-2
    assert np.max(np.abs(xd.get() - synth.exact_solution(hh.levels[0]))) < 1e-7
    # the oracle on the SAME assembled system at full size (threaded C kernels: ~3 s): iteration count and
    # relative residual history within the north-star tolerance
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    ola.set_threaded(True)
    try:
        mats, P, R = oracle_hierarchy(hh)
        smo = [OS.RichardsonSmoother(OS.JacobiLinearSolver(), 10, 2.0 / 3.0)] * 3
        go = OS.GMGLinearSolver(mats, P, R, pre_smoothers=smo, post_smoothers=smo, maxiter=1)
        so = OS.CGSolver(go, maxiter=50, atol=1e-14, rtol=1e-8)
        xo = np.zeros(mats[0].shape[0])
        OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, mats[0]), mats[0]), hh.b)
    finally:
        ola.set_threaded(False)
    assert s.log.num_iters == so.log.num_iters
    assert rel_hist_diff(h, so.log.history()) < HIST_TOL
    assert np.linalg.norm(xd.get() - xo) <= 1e-9 * np.linalg.norm(xo)
    # linearity of the fine-level SpMV: A(2x) == 2 A(x) exactly (power-of-two scaling)
    x2 = dev_vec(gsb, A, 2.0 * xd.get())
    r2 = dev_vec(gsb, A, domain=False)
    gsb.mul_(r2, A, x2)
    assert np.array_equal(r2.get(), 2.0 * rd.get())


def test_numerical_setup_update_refreshes_value_dependent_data(gsb, ctx):
    """numerical_setup!(ns,A) after the matrix values changed (CGSolvers.jl:57-63, JacobiLinearSolvers.jl:25-41)"""
    sysm = fem.poisson((8, 8, 8))
    A = dev_matrix(gsb, ctx, sysm.A)
    s = gsb.CGSolver(gsb.JacobiLinearSolver(), rtol=1e-10)
    ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
    xd = dev_vec(gsb, A)
    gsb.solve_(xd, ns, dev_vec(gsb, A, sysm.b))
    x1 = xd.get()
    A.update_values(3.0 * sysm.A.data)  # same sparsity, new values
    gsb.numerical_setup_(ns, A)
    xd.fill(0.0)
    gsb.solve_(xd, ns, dev_vec(gsb, A, sysm.b))
    assert np.linalg.norm(3.0 * xd.get() - x1) <= 1e-8 * np.linalg.norm(x1)
    so = OS.CGSolver(OS.JacobiLinearSolver(), rtol=1e-10)
    Ao = ola.CSR(3.0 * sysm.A)
    xo = np.zeros(Ao.shape[0])
    OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, Ao), Ao), sysm.b)
    assert s.log.num_iters == so.log.num_iters and rel_hist_diff(s.log.history(), so.log.history()) < HIST_TOL


def test_richardson_linear_solver_parity(gsb, ctx):
    sysm = fem.poisson((8, 8))
    mk = lambda S: S.RichardsonLinearSolver(0.5, 1000, Pl=S.JacobiLinearSolver(), rtol=1e-8, atol=1e-14)
    sg, so, xg, xo = _run_pair(gsb, ctx, sysm, mk, mk)
    assert sg.log.num_iters == so.log.num_iters and sg.log.flag == so.log.flag
    assert rel_hist_diff(sg.log.history(), so.log.history()) < HIST_TOL
    assert fem.l2_error_sq(sysm, xg) < 1e-8
    mk = lambda S: S.RichardsonLinearSolver(0.2, 25, rtol=1e-8, atol=1e-14)  # no preconditioner, stops at maxiter
    sg, so, xg, xo = _run_pair(gsb, ctx, sysm, mk, mk)
    assert sg.log.num_iters == so.log.num_iters == 25
    assert rel_hist_diff(sg.log.history(), so.log.history()) < HIST_TOL


def test_lanczos_diagnostic_matches_oracle(gsb, ctx):
    """PCG-Jacobi with LanczosDiagnostic (KrylovTests.jl:96-137): condition-number estimate from CG's alpha/beta"""
    sysm = fem.poisson((20, 20))
    A, Ao = dev_matrix(gsb, ctx, sysm.A), ola.CSR(sysm.A)
    dg, do = gsb.LanczosDiagnostic(400), OS.LanczosDiagnostic(400)
    sg = gsb.CGSolver(gsb.JacobiLinearSolver(), maxiter=400, rtol=1e-12, diagnostic=dg)
    so = OS.CGSolver(OS.JacobiLinearSolver(), maxiter=400, rtol=1e-12, diagnostic=do)
    b = np.random.default_rng(0).standard_normal(Ao.shape[0])
    ns = gsb.numerical_setup(gsb.symbolic_setup(sg, A), A)
    gsb.solve_(dev_vec(gsb, A), ns, dev_vec(gsb, A, b))
    xo = np.zeros(Ao.shape[0])
    OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, Ao), Ao), b)
    assert dg.k == sg.log.num_iters and abs(dg.k - do.k) <= 1
    n = min(dg.k, do.k, 30)
    assert np.allclose(dg.delta[:n], do.delta[:n], rtol=1e-9) and np.allclose(dg.gamma[:n], do.gamma[:n], rtol=1e-9)
    Dm12 = 1.0 / np.sqrt(sysm.A.diagonal())
    lam = np.linalg.eigvalsh((sysm.A.toarray() * Dm12).T * Dm12)
    # (the reference's off-diagonal uses alpha_k rather than alpha_{k-1}, CGSolvers.jl:133, so its estimate is
    #  only roughly the condition number of the preconditioned operator: 92 vs 80.9 here)
    assert 0.5 * lam.max() / lam.min() < dg.estimate() < 1.5 * lam.max() / lam.min()
    assert abs(dg.estimate() - do.estimate()) <= 1e-6 * do.estimate()


def test_schur_complement_solver(gsb, ctx):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    st = fem.stokes_cavity((6, 6))
    Au, B, Bt = st["A"], st["B"], st["Bt"]
    D = -1e-2 * st["Mp"]
    Sm = (D - B @ spla.spsolve(Au.tocsc(), Bt.tocsc())).tocsr()
    Ag, Sg, Bg, Cg = (dev_matrix(gsb, ctx, m) for m in (Au, Sm, Bt, B))
    A_ns = gsb.numerical_setup(gsb.symbolic_setup(gsb.LUSolver(), Ag), Ag)
    S_ns = gsb.numerical_setup(gsb.symbolic_setup(gsb.LUSolver(), Sg), Sg)
    sc = gsb.SchurComplementSolver(A_ns, Bg, Cg, S_ns)
    ns = gsb.numerical_setup(gsb.symbolic_setup(sc, None), None)
    n = Au.shape[0] + B.shape[0]
    y = np.random.default_rng(2).standard_normal(n)
    xd, yd = gsb.Vector(ctx, n), gsb.Vector(ctx, n)
    yd.set(y)
    gsb.solve_(xd, ns, yd)
    M = sp.bmat([[Au, Bt], [B, D]], format="csr")
    assert np.linalg.norm(M @ xd.get() - y) <= 1e-9 * np.linalg.norm(y)


def test_redistribution_plan_single_part_and_gmg_with_redistributed_level(gsb, ctx):
    """gsb_redist_create / gsb_vec_redistribute on one part (a permutation of own values) and GMG built through
    gsb_gmg_create_redist with identity plans on one level boundary: the redistributed V-cycle must reproduce the plain
    one bit for bit (the multi-rank case runs in tests/mgpu_check.py)"""
    from gsb200 import synth

    n = 1000
    perm = np.random.default_rng(5).permutation(n).astype(np.int64)
    plan = gsb.RedistributionPlan(ctx, n, n, [0], [0, n], np.arange(n), [0], [0, n], perm)
    x = np.random.default_rng(6).standard_normal(n)
    src, dst = gsb.Vector(ctx, n), gsb.Vector(ctx, n)
    src.set(x)
    gsb.redistribute_(dst, plan, src)
    out = np.empty(n)
    out[perm] = x
    assert np.array_equal(dst.get(), out)
    with pytest.raises(gsb.GSBError):
        gsb.redistribute_(gsb.Vector(ctx, n + 1), plan, src)
    nlev = 3
    hh = synth.poisson_hierarchy_host((16, 16, 16), nlev)
    dh = synth.upload_hierarchy(ctx, hh)
    hists = []
    for use_redist in (False, True):
        redist = None
        if use_redist:
            n2 = hh.levels[2].n_own
            ident = lambda: gsb.RedistributionPlan(ctx, n2, n2, [0], [0, n2], np.arange(n2), [0], [0, n2], np.arange(n2))
            redist = ([None, ident()], [None, ident()])
        sm = gsb.Fill(gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), 5, 2.0 / 3.0), nlev - 1)
        gmg = gsb.GMGLinearSolver(dh.A, dh.P, dh.R, pre_smoothers=sm, post_smoothers=sm, maxiter=1, redist=redist)
        s = gsb.CGSolver(gmg, maxiter=30, atol=1e-14, rtol=1e-8)
        ns = gsb.numerical_setup(gsb.symbolic_setup(s, dh.A[0]), dh.A[0])
        xs, b = gsb.allocate_in_domain(dh.A[0]), gsb.allocate_in_domain(dh.A[0])
        b.set(hh.b)
        for _ in range(2):
            xs.fill(0.0)
            gsb.solve_(xs, ns, b)
        hists.append(s.log.history())
    assert np.array_equal(hists[0], hists[1])


def test_solve_affine_operator_entry_points(gsb, ctx):
    """solve!(x, ls, op::AffineOperator, cache[, newmatrix]) -- SolverInterfaces/GridapExtras.jl:33-58: first call sets up
    and returns the cache (ns, y); later calls reuse it, newmatrix=True refreshes the numerical set-up after the values
    of the matrix changed; x is only touched through own-value copies"""
    sysm = fem.poisson((24, 24))
    A = dev_matrix(gsb, ctx, sysm.A)
    b = dev_vec(gsb, A, sysm.b)
    op = gsb.AffineOperator(A, b)
    ls = gsb.CGSolver(gsb.JacobiLinearSolver(), maxiter=500, atol=1e-14, rtol=1e-10)
    x = dev_vec(gsb, A)
    cache = gsb.solve_affine_(x, ls, op)
    so = OS.CGSolver(OS.JacobiLinearSolver(), maxiter=500, atol=1e-14, rtol=1e-10)
    Ao = ola.CSR(sysm.A)
    xo = np.zeros(sysm.A.shape[0])
    OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, Ao), Ao), sysm.b)
    assert ls.log.num_iters == so.log.num_iters and rel_hist_diff(ls.log.history(), so.log.history()) < HIST_TOL
    assert np.linalg.norm(x.get() - xo) <= 1e-9 * np.linalg.norm(xo)
    # same cache, new matrix values (2 A): the solution halves, the set-up objects are reused
    A.update_values(2.0 * sysm.A.data)  # same sparsity, new values
    x.fill(0.0)
    cache2 = gsb.solve_affine_(x, ls, op, cache, newmatrix=True)
    assert cache2 is cache
    assert np.linalg.norm(x.get() - 0.5 * xo) <= 1e-8 * np.linalg.norm(xo)
