"""GPU parity tests of the remaining BASELINE configs (scaled to sizes the oracle finishes in seconds):
  C1  CGSolver + JacobiLinearSolver, 2D Poisson Q1 256x256 (full size)
  C4  FGMRESSolver(30) + GMG on 3D linear elasticity Q2 (vector-valued, 3x3 block sparsity)
  C5  GMRES + BlockTriangularSolver[GMG velocity block, CG-Jacobi pressure mass] on 2D Stokes Q2-P1disc
"""
import numpy as np
import pytest

import gsb200
from gsb200 import synth
from oracle import fem
from oracle import linalg as ola
from oracle import solvers as OS
from util import dev_matrix, dev_vec, rel_hist_diff

pytestmark = pytest.mark.gpu


def test_c1_cg_jacobi_poisson2d_256(gsb, ctx):
    """C1 at full size.  Several hundred CG iterations amplify summation-order differences (loss of
    orthogonality), so the bar is: iteration count +-1, the first 50 residuals to 1e-10 relative, the
    whole history to 1e-6 relative, and the reference's own L2-error assertion."""
    hh = synth.poisson_hierarchy_host((256, 256), 1)
    n = hh.levels[0].n_own
    assert n == 65025 and int(hh.A[0][0][-1]) == 582169
    As = synth.to_scipy(*hh.A[0], n)
    A, Ao = dev_matrix(gsb, ctx, As), ola.CSR(As)
    mk = lambda S: S.CGSolver(S.JacobiLinearSolver(), maxiter=1000, rtol=1e-8)
    sg, so = mk(gsb), mk(OS)
    ns = gsb.numerical_setup(gsb.symbolic_setup(sg, A), A)
    xd = dev_vec(gsb, A)
    gsb.solve_(xd, ns, dev_vec(gsb, A, hh.b))
    xo = np.zeros(n)
    OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, Ao), Ao), hh.b)
    assert abs(sg.log.num_iters - so.log.num_iters) <= 1
    hg, ho = sg.log.history(), so.log.history()
    assert rel_hist_diff(hg[:50], ho[:50]) < 1e-10
    assert rel_hist_diff(hg, ho) < 1e-6
    M = synth.to_scipy(*synth.mass_rows(hh.levels[0]), n)
    e = xd.get() - synth.exact_solution(hh.levels[0])
    assert float(e @ (M @ e)) < 1e-6  # KrylovTests.jl:21-25


def _gmg_solvers(S, mats, P, R, nlev, outer):
    sm = [S.RichardsonSmoother(S.JacobiLinearSolver(), 10, 2.0 / 3.0)] * (nlev - 1)
    gmg = S.GMGLinearSolver(mats, P, R, pre_smoothers=sm, post_smoothers=sm, coarsest_solver=S.LUSolver(), maxiter=1)
    return outer(S, gmg), gmg


def test_c4_fgmres_gmg_elasticity_q2(gsb, ctx):
    """FGMRESSolver(30, GMG) on Q2 vector-valued elasticity (lambda = mu = 1, clamped x=0 face);
    long rows (81..375 nnz) exercise the multi-lane row kernels => tolerance parity, not bit equality."""
    nlev = 2
    H = fem.elasticity_hierarchy((4, 4, 4), nlev, order=2)
    mats_o = [ola.CSR(m) for m in H.mats]
    P_o, R_o = [ola.CSR(p) for p in H.P], [ola.CSR(r) for r in H.R]
    mats_g = [dev_matrix(gsb, ctx, m) for m in H.mats]
    P_g, R_g = [dev_matrix(gsb, ctx, p) for p in H.P], [dev_matrix(gsb, ctx, r) for r in H.R]
    outer = lambda S, gmg: S.FGMRESSolver(30, gmg, maxiter=200, atol=1e-14, rtol=1e-8)
    sg, _ = _gmg_solvers(gsb, mats_g, P_g, R_g, nlev, outer)
    so, _ = _gmg_solvers(OS, mats_o, P_o, R_o, nlev, outer)
    b = H.systems[0].b
    A = mats_g[0]
    ns = gsb.numerical_setup(gsb.symbolic_setup(sg, A), A)
    xd = dev_vec(gsb, A)
    gsb.solve_(xd, ns, dev_vec(gsb, A, b))
    xo = np.zeros(b.shape[0])
    OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, mats_o[0]), mats_o[0]), b)
    assert so.log.flag == OS.SOLVER_CONVERGED_RTOL and sg.log.flag == so.log.flag
    assert abs(sg.log.num_iters - so.log.num_iters) <= 1
    assert rel_hist_diff(sg.log.history(), so.log.history()) < 1e-10
    assert np.linalg.norm(H.mats[0] @ xd.get() - b) <= 1e-7 * np.linalg.norm(b)
    # the vector-valued matrix has the 3x3-block sparsity of node-major dofs
    assert H.mats[0].shape[0] % 3 == 0 and H.mats[0][0].nnz % 3 == 0


def _bench_parity(gsb, ctx, config, cells):
    """device solve of a bench.py configuration vs the threaded CPU oracle on the same generated system"""
    import bench

    prob = bench.Problem(config, cells)
    solver, ns, x, b = prob.build_device(gsb, ctx)
    gsb.solve_(x, ns, b)
    hist_first = solver.log.history()
    iters_first = solver.log.num_iters
    _, ohist, oit = bench.oracle_sample(prob, bench.MAXITER, bench.host_threads())
    assert solver.log.flag == OS.SOLVER_CONVERGED_RTOL
    assert abs(iters_first - oit) <= 1, (iters_first, oit)
    n = min(iters_first, oit) + 1
    return prob, solver, ns, x, b, rel_hist_diff(hist_first[:n], ohist[:n]), iters_first


def test_c4_fgmres_gmg_elasticity_q2_32cubed(gsb, ctx):
    """C4 at 32^3 Q2 cells (823 k dofs, 155 M non-zeros, 4 levels): FGMRES(30)+GMG against the threaded oracle --
    identical iteration count and relative residual history to 1e-10; the fine matrix must be stored as sorted
    3x3 block-SELL (the long-row format) and its CSR arrays released"""
    prob, solver, ns, x, b, d, iters = _bench_parity(gsb, ctx, "c4", 32)
    f = prob.fine.format()
    assert f["kind"] == "bsell32" and f["block_size"] == 3 and f["sorted"]
    assert f["stored_entries"] <= 1.10 * prob.level_nnz[0]
    assert f["bytes_per_pass"] < 0.78 * 12 * prob.level_nnz[0]  # 8.44 B per non-zero + padding vs CSR's 12
    assert d < 1e-10, d
    # true residual of the device solution on the generated system
    from util import host_to_scipy

    A = host_to_scipy(prob.hh.A[0], prob.n_own)
    assert np.linalg.norm(A @ x.get() - prob.b) <= 2e-8 * np.linalg.norm(prob.b)


def test_c5_gmres_block_triangular_stokes_256(gsb, ctx):
    """C5 at full size (256^2 Q2-P1disc cells, 522 k velocity + 197 k pressure dofs, 6 velocity levels): GMRES(30)
    right-preconditioned by the block-triangular solver of joss_paper/demo.jl against the threaded oracle"""
    prob, solver, ns, x, b, d, iters = _bench_parity(gsb, ctx, "c5", 256)
    assert prob.fine.format()["block_size"] == 2
    assert d < 1e-8, d  # inner CG / GMG iterations amplify reduction-order differences: 1e-8 on the relative history


def _stokes_pair(gsb, ctx, outer):
    nlev = 2
    st = fem.stokes_cavity((16, 16), nlevels=nlev)
    Mpneg = -1.0 * st["Mp"]
    # ---- oracle
    A_o, B_o, Bt_o = ola.CSR(st["A"]), ola.CSR(st["B"]), ola.CSR(st["Bt"])
    M_o = OS.BlockMatrix([[A_o, Bt_o], [B_o, None]])
    mats_o = [ola.CSR(m) for m in st["mats"]]
    sm_o = [OS.RichardsonSmoother(OS.JacobiLinearSolver(), 10, 2.0 / 3.0)] * (nlev - 1)
    gmg_o = OS.GMGLinearSolver(mats_o, [ola.CSR(p) for p in st["P"]], [ola.CSR(r) for r in st["R"]], pre_smoothers=sm_o,
                               post_smoothers=sm_o, maxiter=3, mode="solver", rtol=1e-10)
    bt_o = OS.BlockTriangularSolver([gmg_o, OS.CGSolver(OS.JacobiLinearSolver(), rtol=1e-10)], coeffs=[[1.0, 1.0], [0.0, 1.0]],
                                    half="upper", diag_mats=[None, ola.CSR(Mpneg)])
    so = outer(OS, OS.BlockPrecondAdapter(bt_o))
    b = np.concatenate([st["fu"], st["fp"]])
    xo = np.zeros(M_o.shape[1])
    OS.solve_(xo, OS.numerical_setup(OS.symbolic_setup(so, M_o), M_o), b)
    # ---- device
    A_g, B_g, Bt_g, Mp_g = (dev_matrix(gsb, ctx, m) for m in (st["A"], st["B"], st["Bt"], Mpneg))
    M_g = gsb.BlockSparseMatrix([[A_g, Bt_g], [B_g, None]])
    mats_g = [A_g] + [dev_matrix(gsb, ctx, m) for m in st["mats"][1:]]
    sm_g = gsb.Fill(gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), 10, 2.0 / 3.0), nlev - 1)
    gmg_g = gsb.GMGLinearSolver(mats_g, [dev_matrix(gsb, ctx, p) for p in st["P"]], [dev_matrix(gsb, ctx, r) for r in st["R"]],
                                pre_smoothers=sm_g, post_smoothers=sm_g, maxiter=3, mode="solver", rtol=1e-10)
    bt_g = gsb.BlockTriangularSolver([gmg_g, gsb.CGSolver(gsb.JacobiLinearSolver(), rtol=1e-10)], coeffs=[[1.0, 1.0], [0.0, 1.0]],
                                     half="upper", diag_mats=[None, Mp_g])
    sg = outer(gsb, bt_g)
    ns = gsb.numerical_setup(gsb.symbolic_setup(sg, M_g), M_g)
    xd, bd = gsb.Vector(ctx, M_g.n_own_cols), gsb.Vector(ctx, M_g.n_rows)
    bd.set(b)
    gsb.solve_(xd, ns, bd)
    return M_o, b, sg, so, xd.get(), xo


def test_c5_gmres_block_triangular_stokes(gsb, ctx):
    """C5: GMRES(30; Pr = BlockTriangularSolver[GMG(velocity, mode=:solver, maxiter=3), CG-Jacobi(-pressure
    mass)]) on the 2D lid-driven cavity Q2-P1disc (joss_paper/demo.jl:20-91, stokes_gmg.jl:41-63).
    The inner iterative block solvers are warm-started (quirk App. C.6: the `y` caches are zeroed at
    set-up only), which the device path must reproduce for the histories to agree."""
    outer = lambda S, Pr: S.GMRESSolver(30, Pr=Pr, maxiter=300, atol=1e-14, rtol=1e-10)
    M_o, b, sg, so, xg, xo = _stokes_pair(gsb, ctx, outer)
    assert abs(sg.log.num_iters - so.log.num_iters) <= 1
    n = min(sg.log.num_iters, so.log.num_iters) + 1
    assert rel_hist_diff(sg.log.history()[:n], so.log.history()[:n]) < 1e-8
    assert np.linalg.norm(xg - xo) <= 1e-6 * np.linalg.norm(xo)


def test_c5_fgmres_block_triangular_stokes_known_answer(gsb, ctx):
    """test/Applications/Stokes.jl:92-110: FGMRES + BlockTriangular => ||Ax - b|| < 1e-7"""
    outer = lambda S, Pr: S.FGMRESSolver(20, Pr, maxiter=300, atol=1e-14, rtol=1e-10)
    M_o, b, sg, so, xg, xo = _stokes_pair(gsb, ctx, outer)
    assert np.linalg.norm(M_o.to_scipy() @ xg - b) < 1e-7
    assert np.linalg.norm(M_o.to_scipy() @ xo - b) < 1e-7
    assert abs(sg.log.num_iters - so.log.num_iters) <= 1
    n = min(sg.log.num_iters, so.log.num_iters) + 1
    assert rel_hist_diff(sg.log.history()[:n], so.log.history()[:n]) < 1e-8


def test_block_diagonal_solver(gsb, ctx):
    """BlockDiagonalSolversTests.jl:45 : ||x - x_ref|| < 1e-8 with exact diagonal-block solves"""
    s1, s2 = fem.poisson((6, 6)), fem.poisson((5, 7))
    A1, A2 = dev_matrix(gsb, ctx, s1.A), dev_matrix(gsb, ctx, s2.A)
    M = gsb.BlockSparseMatrix([[A1, None], [None, A2]])
    bd = gsb.BlockDiagonalSolver([gsb.LUSolver(), gsb.LUSolver()])
    ns = gsb.numerical_setup(gsb.symbolic_setup(bd, M), M)
    b = np.concatenate([s1.b, s2.b])
    xd, bv = gsb.Vector(ctx, M.n_own_cols), gsb.Vector(ctx, M.n_rows)
    bv.set(b)
    gsb.solve_(xd, ns, bv)
    import scipy.sparse.linalg as spla

    xr = np.concatenate([spla.spsolve(s1.A.tocsc(), s1.b), spla.spsolve(s2.A.tocsc(), s2.b)])
    assert np.linalg.norm(xd.get() - xr) < 1e-8
    # block SpMV of the assembled block matrix
    yd = gsb.Vector(ctx, M.n_rows)
    gsb.mul_(yd, M, xd)
    assert np.linalg.norm(yd.get() - b) <= 1e-10 * np.linalg.norm(b)
