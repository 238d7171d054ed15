"""CPU tests (no GPU): the oracle against (1) every known-answer assertion the reference's own
tests hold for the solve phase and (2) the committed oracle-generated fixtures.

Reference known answers (no golden vectors exist in the reference tree):
  test/LinearSolvers/KrylovTests.jl:14-26,66-93   L2 error^2 < 1e-6 for every Krylov variant
  test/LinearSolvers/SmoothersTests.jl:12-44      CG + Richardson(Jacobi,5,2/3): L2 error^2 < 1e-8
"""
import json
import os

import numpy as np
import pytest

from oracle import fem
from oracle import linalg as ola
from oracle import solvers as S

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_histories.json")))
P = S.JacobiLinearSolver

CASES = {
    "gmres40_PrPl": lambda: S.GMRESSolver(40, Pr=P(), Pl=P(), rtol=1e-8),
    "gmres10": lambda: S.GMRESSolver(10, rtol=1e-8),
    "gmres10_restart": lambda: S.GMRESSolver(10, restart=True, rtol=1e-8),
    "fgmres10": lambda: S.FGMRESSolver(10, P(), rtol=1e-8),
    "fgmres10_restart": lambda: S.FGMRESSolver(10, P(), restart=True, rtol=1e-8),
    "cg": lambda: S.CGSolver(rtol=1e-8),
    "pcg": lambda: S.CGSolver(P(), rtol=1e-8),
    "fpcg": lambda: S.CGSolver(P(), flexible=True, rtol=1e-8),
    "minres": lambda: S.MINRESSolver(Pl=P(), rtol=1e-8),
    "cg_richardson": lambda: S.CGSolver(S.LinearSolverFromSmoother(S.RichardsonSmoother(P(), 5, 2.0 / 3.0)), rtol=1e-8),
}


@pytest.mark.parametrize("nc", [(8, 8), (8, 8, 8)])
@pytest.mark.parametrize("case", sorted(CASES))
def test_reference_known_answers_and_fixture(nc, case):
    sysm = fem.poisson(nc)
    A = ola.CSR(sysm.A)
    s = CASES[case]()
    x = S.allocate_in_domain(A)
    S.solve_(x, S.numerical_setup(S.symbolic_setup(s, A), A), sysm.b)
    E = fem.l2_error_sq(sysm, x)
    assert E < (1e-8 if case == "cg_richardson" else 1e-6)
    g = GOLD["krylov"]["%s/%s" % ("x".join(map(str, nc)), case)]
    assert s.log.num_iters == g["num_iters"] and s.log.flag == g["flag"]
    assert np.allclose(s.log.history(), g["residuals"], rtol=1e-9, atol=1e-14 * g["residuals"][0])


def test_stencil_values_match_survey_appendix_d():
    """Q1 Laplacian stencils: 2D centre 8/3, neighbours -1/3; 3D centre 8h/3, faces 0, edges -h/6, corners -h/12"""
    A = fem.poisson((6, 6)).A.toarray()
    i = 2 * 5 + 2  # interior node (3,3)
    assert abs(A[i, i] - 8.0 / 3.0) < 1e-14 and abs(A[i, i + 1] + 1.0 / 3.0) < 1e-14 and abs(A[i, i + 6] + 1.0 / 3.0) < 1e-14
    n = 6
    h = 1.0 / n
    sysm = fem.poisson((n, n, n))
    A3 = sysm.A.tocsr()
    m = n - 1
    c = 2 + m * (2 + m * 2)
    row = A3[c].toarray().ravel()
    assert abs(row[c] - 8 * h / 3) < 1e-15
    assert abs(row[c + 1]) < 1e-16 and abs(row[c + m]) < 1e-16 and abs(row[c + m * m]) < 1e-16  # faces: structural zeros
    assert abs(row[c + 1 + m] + h / 6) < 1e-15 and abs(row[c + 1 + m + m * m] + h / 12) < 1e-15
    assert A3[c].nnz == 27  # the numerically-zero face couplings stay stored (sparsity parity)


def test_sizes_match_survey_section_8():
    for n, N, nnz in [(16, 3375, 79507), (32, 29791, 753571)]:
        A = fem.poisson((n, n, n)).A
        assert A.shape[0] == N and A.nnz == nnz
    A = fem.poisson((64, 64)).A
    assert A.shape[0] == 63 * 63 and A.nnz == (3 * 63 - 2) ** 2


@pytest.mark.parametrize("key", sorted(GOLD["gmg"]))
def test_gmg_fixture(key):
    ncs, nlev, cyc = key.split("/")
    nc, nlev = tuple(int(v) for v in ncs.split("x")), int(nlev)
    if np.prod(nc) > 20000:
        pytest.skip("kept for the GPU parity tests")
    H = fem.poisson_hierarchy(nc, nlev)
    mats = [ola.CSR(m) for m in H.mats]
    sm = [S.RichardsonSmoother(P(), 10, 2.0 / 3.0)] * (nlev - 1)
    gmg = S.GMGLinearSolver(mats, [ola.CSR(p) for p in H.P], [ola.CSR(r) for r in H.R], pre_smoothers=sm,
                            post_smoothers=sm, maxiter=1, cycle_type=cyc)
    s = S.CGSolver(gmg, maxiter=20, atol=1e-14, rtol=1e-8) if cyc == "v_cycle" else S.FGMRESSolver(5, gmg, maxiter=20, atol=1e-14, rtol=1e-8)
    x = S.allocate_in_domain(mats[0])
    S.solve_(x, S.numerical_setup(S.symbolic_setup(s, mats[0]), mats[0]), H.systems[0].b)
    g = GOLD["gmg"][key]
    assert s.log.num_iters == g["num_iters"]
    assert np.allclose(s.log.history(), g["residuals"], rtol=1e-8, atol=1e-13 * g["residuals"][0])
    assert fem.l2_error_sq(H.systems[0], x) < 1e-6


def test_restriction_is_transpose_and_galerkin_consistency():
    """R = P^T (dual projection, GridTransferOperators.jl:206-208,536-561) and P reproduces
    coarse FE functions: P * (nodal values of x+y with zero boundary contribution) is exact in the interior"""
    H = fem.poisson_hierarchy((8, 8, 8), 2)
    assert abs(H.R[0] - H.P[0].T).max() == 0.0
    P0 = H.P[0]
    assert set(np.unique(P0.data)) <= {1.0, 0.5, 0.25, 0.125}
    # partition of unity away from the Dirichlet boundary
    rs = np.asarray(P0.sum(axis=1)).ravel()
    fine = H.systems[0]
    mi = fine.grid.node_multi_index()[fine.free]
    inner = ((mi >= 2) & (mi <= 6)).all(axis=1)
    assert np.allclose(rs[inner], 1.0)


def test_convergence_log_semantics():
    """ConvergenceLogs.jl:101-150 / SolverTolerances.jl:97-128"""
    tols = S.SolverTolerances(maxiter=3, atol=1e-12, rtol=1e-6)
    log = S.ConvergenceLog("t", tols)
    assert log.residuals.shape[0] == 4
    assert S.init_(log, 1.0) is False
    assert S.update_(log, 0.5) is False and log.num_iters == 1
    assert S.update_(log, 0.5e-6) is True  # r/r0 < rtol
    assert S.finalize_(log, 0.5e-6) == S.SOLVER_CONVERGED_RTOL
    assert S.init_(log, 1e-13) is True  # atol at iteration 0
    assert S.finalize_(log, 1e-13) == S.SOLVER_CONVERGED_ATOL
    S.init_(log, 1.0)
    for _ in range(3):
        done = S.update_(log, 0.9)
    assert done and S.finalize_(log, 0.9) == S.SOLVER_DIVERGED_MAXITER
    # strict inequalities: r/r0 == rtol does not converge
    S.init_(log, 1.0)
    assert S.update_(log, 1e-6) is False


def test_givens_matches_dlartg_conventions():
    for f, g in [(3.0, 4.0), (-3.0, 4.0), (-5.0, 1.0), (0.0, 2.0), (2.0, 0.0), (1e-3, -7.0)]:
        c, s, r = S.givens_algorithm(f, g)
        assert abs(c * f + s * g - r) < 1e-14 * max(1, abs(r)) and abs(-s * f + c * g) < 1e-14 * max(1, abs(r))
        assert abs(c * c + s * s - 1) < 1e-14
        if abs(f) > abs(g):
            assert c > 0


def test_block_triangular_and_stokes_oracle():
    """C5-style stack on the oracle: GMRES + BlockTriangularSolver[LU velocity, CG-Jacobi(pressure mass)]
    (test/Applications/Stokes.jl:92-110 asserts ||Ax-b|| < 1e-7 for FGMRES+BlockTriangular)"""
    st = fem.stokes_cavity((8, 8))
    A, B, Bt, Mp = (ola.CSR(st[k]) for k in ("A", "B", "Bt", "Mp"))
    M = S.BlockMatrix([[A, Bt], [B, None]])
    bt = S.BlockTriangularSolver([S.LUSolver(), S.CGSolver(P(), rtol=1e-10)], coeffs=[[1.0, 1.0], [0.0, 1.0]], half="upper",
                                 diag_mats=[None, ola.CSR(-1.0 * st["Mp"])])
    s = S.FGMRESSolver(20, S.BlockPrecondAdapter(bt), maxiter=200, atol=1e-14, rtol=1e-10)
    b = np.concatenate([st["fu"], st["fp"]])
    x = np.zeros(M.shape[1])
    S.solve_(x, S.numerical_setup(S.symbolic_setup(s, M), M), b)
    assert np.linalg.norm(M.to_scipy() @ x - b) < 1e-7
    assert s.log.flag in (S.SOLVER_CONVERGED_RTOL, S.SOLVER_CONVERGED_ATOL)


def test_richardson_linear_solver_known_answer():
    """test/LinearSolvers/RichardsonLinearTests.jl: Richardson iteration with a Jacobi left preconditioner converges
    to the discrete solution of the Poisson problem (L2 error^2 < 1e-8 there with rtol 1e-8)."""
    sysm = fem.poisson((8, 8))
    A = ola.CSR(sysm.A)
    s = S.RichardsonLinearSolver(0.5, 1000, Pl=P(), rtol=1e-8, atol=1e-14)
    x = S.allocate_in_domain(A)
    S.solve_(x, S.numerical_setup(S.symbolic_setup(s, A), A), sysm.b)
    assert s.log.flag == S.SOLVER_CONVERGED_RTOL
    assert fem.l2_error_sq(sysm, x) < 1e-8
    assert np.linalg.norm(sysm.A @ x - sysm.b) <= 1.01e-8 * np.linalg.norm(sysm.b)


def test_schur_complement_solver_is_exact_block_inverse():
    st = fem.stokes_cavity((6, 6))
    Au, B, Bt = st["A"], st["B"], st["Bt"]
    import scipy.sparse.linalg as spla
    import scipy.sparse as sp

    nu, npr = Au.shape[0], B.shape[0]
    D = -1e-2 * st["Mp"]  # a stabilised (regular) 2x2 block system [[A, Bt], [B, D]]
    Sm = (D - B @ spla.spsolve(Au.tocsc(), Bt.tocsc())).tocsr()
    A_ns = S.numerical_setup(S.symbolic_setup(S.LUSolver(), ola.CSR(Au)), ola.CSR(Au))
    S_ns = S.numerical_setup(S.symbolic_setup(S.LUSolver(), ola.CSR(Sm)), ola.CSR(Sm))
    sc = S.SchurComplementSolver(A_ns, ola.CSR(Bt), ola.CSR(B), S_ns)
    ns = sc._numerical_setup(None)
    rng = np.random.default_rng(1)
    y = [rng.standard_normal(nu), rng.standard_normal(npr)]
    x = [np.zeros(nu), np.zeros(npr)]
    ns.solve(x, y)
    M = sp.bmat([[Au, Bt], [B, D]], format="csr")
    assert np.linalg.norm(M @ np.concatenate(x) - np.concatenate(y)) <= 1e-9 * np.linalg.norm(np.concatenate(y))


def test_block_diagonal_solver_known_answer():
    """test/BlockSolvers/BlockDiagonalSolversTests.jl:45,55,66: ||x - x_direct|| < 1e-8 when GMRES is right-
    preconditioned by a BlockDiagonalSolver of exact diagonal-block solves on a block-diagonal-dominant system"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    s1, s2 = fem.poisson((6, 6)), fem.poisson((5, 7))
    n1, n2 = s1.A.shape[0], s2.A.shape[0]
    rng = np.random.default_rng(4)
    C12 = sp.random(n1, n2, density=0.05, random_state=1, format="csr") * 0.05
    C21 = sp.random(n2, n1, density=0.05, random_state=2, format="csr") * 0.05
    M = S.BlockMatrix([[ola.CSR(s1.A), ola.CSR(C12)], [ola.CSR(C21), ola.CSR(s2.A)]])
    bd = S.BlockDiagonalSolver([S.LUSolver(), S.LUSolver()])
    solver = S.GMRESSolver(10, Pr=S.BlockPrecondAdapter(bd), maxiter=50, atol=1e-14, rtol=1e-12)
    b = rng.standard_normal(n1 + n2)
    x = np.zeros(n1 + n2)
    S.solve_(x, S.numerical_setup(S.symbolic_setup(solver, M), M), b)
    xd = spla.spsolve(M.to_scipy().tocsc(), b)
    assert np.linalg.norm(x - xd) < 1e-8


def test_oracle_krylov_histories_against_independent_implementations():
    """The oracle restates GridapSolvers' OWN variants of the Krylov methods; this cross-check ties its residual
    histories to independent implementations of the published algorithms (scipy): for a consistent Jacobi
    preconditioner the PCG iterates are unique, so the true residuals of scipy's iterates must equal the oracle's
    recursively updated residual norms (CGSolvers.jl:85,111); the multigrid V-cycle as a stationary iteration
    (mode=:solver) must contract the error like the two-grid operator I - M A it defines (checked through the exact
    solve of the same system)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    sysm = fem.poisson((24, 20))
    A, b = sp.csr_matrix(sysm.A), sysm.b
    Ao = ola.CSR(A)
    s = S.CGSolver(S.JacobiLinearSolver(), maxiter=60, atol=1e-30, rtol=1e-12)
    x = np.zeros(A.shape[0])
    S.solve_(x, S.numerical_setup(S.symbolic_setup(s, Ao), Ao), b)
    hist = s.log.history()
    true_res = [np.linalg.norm(b)]
    Minv = sp.diags(1.0 / A.diagonal())
    spla.cg(A, b, x0=np.zeros_like(b), M=Minv, rtol=1e-14, atol=0.0, maxiter=len(hist) - 1,
            callback=lambda xk: true_res.append(np.linalg.norm(b - A @ xk)))
    n = min(len(hist), len(true_res))
    assert n >= 20
    assert np.max(np.abs(np.array(true_res[:n]) - hist[:n])) <= 1e-9 * hist[0]
    # GMG V-cycle in solver mode converges to the direct solution of the same system at a mesh-independent rate
    H = fem.poisson_hierarchy((16, 16), 3)
    mats = [ola.CSR(m) for m in H.mats]
    g = S.GMGLinearSolver(mats, [ola.CSR(p) for p in H.P], [ola.CSR(r) for r in H.R], mode="solver", maxiter=30, rtol=1e-10)
    xg = np.zeros(H.mats[0].shape[0])
    S.solve_(xg, S.numerical_setup(S.symbolic_setup(g, mats[0]), mats[0]), H.systems[0].b)
    xd = spla.spsolve(sp.csc_matrix(H.mats[0]), H.systems[0].b)
    assert np.linalg.norm(xg - xd) <= 1e-8 * np.linalg.norm(xd)
    h = g.log.history()
    rates = h[2:] / h[1:-1]
    assert rates.max() < 0.35  # textbook V(10,10)-cycle contraction for the 2D Laplacian with damped Jacobi
