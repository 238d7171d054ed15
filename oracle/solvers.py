"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Statement-by-statement CPU restatement of the reference's solve phase
(gridap/GridapSolvers.jl v0.7.1, all paths relative to /root/reference/src):

  SolverInterfaces/SolverTolerances.jl:40-49,97-128     -> SolverTolerances, finished_flag
  SolverInterfaces/ConvergenceLogs.jl:42-150            -> ConvergenceLog, init_/update_/finalize_
  LinearSolvers/JacobiLinearSolvers.jl:20-56            -> JacobiLinearSolver
  LinearSolvers/RichardsonSmoothers.jl:84-105           -> RichardsonSmoother
  LinearSolvers/LinearSolverFromSmoothers.jl:44-50      -> LinearSolverFromSmoother
  LinearSolvers/IdentityLinearSolvers.jl:23-26          -> IdentitySolver
  LinearSolvers/GMGLinearSolvers.jl:48-69,183-210,451-645 -> GMGLinearSolver (V/W/F cycles)
  LinearSolvers/Krylov/KrylovUtils.jl:17-54             -> krylov_mul_, krylov_residual_
  LinearSolvers/Krylov/CGSolvers.jl:73-120              -> CGSolver
  LinearSolvers/Krylov/GMRESSolvers.jl:132-210          -> GMRESSolver
  LinearSolvers/Krylov/FGMRESSolvers.jl:130-199         -> FGMRESSolver
  LinearSolvers/Krylov/MINRESSolvers.jl:75-149          -> MINRESSolver
  BlockSolvers/BlockTriangularSolvers.jl:188-242        -> BlockTriangularSolver
  BlockSolvers/BlockDiagonalSolvers.jl:165-177          -> BlockDiagonalSolver
  LinearSolvers/RichardsonLinearSolvers.jl:79-106       -> RichardsonLinearSolver
  LinearSolvers/SchurComplementSolvers.jl:55-74         -> SchurComplementSolver
  Gridap.Algebra.LUSolver (dep, UMFPACK)                -> LUSolver (SuperLU: exact sparse direct solve)

Julia's `f!` is spelled `f_` here.  Vectors are numpy fp64 arrays; matrices are oracle.linalg.CSR.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import linalg as la
from .linalg import CSR

# ------------------------------------------------------------------ tolerances / logs

SOLVER_CONVERGED_ATOL = 0
SOLVER_CONVERGED_RTOL = 1
SOLVER_DIVERGED_MAXITER = 2
SOLVER_DIVERGED_BREAKDOWN = 3


class SolverTolerances:
    """SolverTolerances.jl:40-49 (dtol stored, never consulted :117-128)."""

    def __init__(self, maxiter=1000, atol=np.finfo(np.float64).eps, rtol=1e-5, dtol=math.inf):
        self.maxiter, self.atol, self.rtol, self.dtol = int(maxiter), float(atol), float(rtol), float(dtol)


def converged(tols, niter, e_a, e_r):  # SolverTolerances.jl:126-128
    return bool((e_r < tols.rtol) or (e_a < tols.atol))


def finished(tols, niter, e_a, e_r):  # SolverTolerances.jl:117-119
    return bool((niter >= tols.maxiter) or converged(tols, niter, e_a, e_r))


def finished_flag(tols, niter, e_a, e_r):  # SolverTolerances.jl:97-110
    if e_r < tols.rtol:
        return SOLVER_CONVERGED_RTOL
    elif e_a < tols.atol:
        return SOLVER_CONVERGED_ATOL
    elif niter >= tols.maxiter:
        return SOLVER_DIVERGED_MAXITER
    return SOLVER_DIVERGED_BREAKDOWN


class ConvergenceLog:
    """ConvergenceLogs.jl:42-60: residuals has fixed length maxiter+1."""

    def __init__(self, name, tols, verbose=0, depth=0):
        self.name, self.tols = name, tols
        self.num_iters = 0
        self.residuals = np.zeros(tols.maxiter + 1)
        self.verbose, self.depth = int(verbose), depth

    def history(self):
        return self.residuals[: self.num_iters + 1].copy()


def reset_(log):  # ConvergenceLogs.jl:90-94
    log.num_iters = 0
    log.residuals[:] = 0.0
    return log


def init_(log, r0):  # ConvergenceLogs.jl:101-112
    reset_(log)
    log.residuals[0] = r0
    if log.verbose > 1:
        print(" " * (2 + 2 * log.depth) + "> Iteration %3i - Residuals: %.2e,   %.2e " % (0, r0, 1))
    return finished(log.tols, log.num_iters, r0, 1.0)


def update_(log, r):  # ConvergenceLogs.jl:119-129
    log.num_iters += 1
    log.residuals[log.num_iters] = r
    r_rel = r / log.residuals[0]
    if log.verbose > 1:
        print(" " * (2 + 2 * log.depth) + "> Iteration %3i - Residuals: %.2e,   %.2e " % (log.num_iters, r, r_rel))
    return finished(log.tols, log.num_iters, r, r_rel)


def finalize_(log, r):  # ConvergenceLogs.jl:136-150
    r_rel = r / log.residuals[0]
    flag = finished_flag(log.tols, log.num_iters, r, r_rel)
    log.flag = flag
    return flag


# ------------------------------------------------------------------ generic API


def allocate_in_domain(A):
    return np.zeros(A.shape[1])


def allocate_in_range(A):
    return np.zeros(A.shape[0])


def symbolic_setup(solver, A):
    return solver.symbolic_setup(A)


def numerical_setup(ss, A):
    return ss.numerical_setup(A)


def numerical_setup_(ns, A):
    return ns.update(A)


def solve_(x, ns, b):
    return ns.solve(x, b)


class _SS:
    def __init__(self, solver):
        self.solver = solver

    def numerical_setup(self, A):
        return self.solver._numerical_setup(A)


class _Solver:
    def symbolic_setup(self, A):
        return _SS(self)


def _setup(solver, A):
    return None if solver is None else numerical_setup(symbolic_setup(solver, A), A)


# ------------------------------------------------------------------ simple solvers


class IdentitySolver(_Solver):  # IdentityLinearSolvers.jl:2-26
    def _numerical_setup(self, A):
        return IdentityNS()


class IdentityNS:
    def solve(self, x, b):
        x[:] = b
        return x

    def update(self, A):
        return self


class JacobiLinearSolver(_Solver):  # JacobiLinearSolvers.jl:6
    def _numerical_setup(self, A):
        return JacobiNS(1.0 / A.diag())  # :20-23


class JacobiNS:
    def __init__(self, inv_diag):
        self.inv_diag = inv_diag

    def update(self, A):  # :25-27
        self.inv_diag[:] = 1.0 / A.diag()
        return self

    def solve(self, x, b):  # :43-47   x .= inv_diag .* b
        la.emul(x, self.inv_diag, b)
        return x


class LUSolver(_Solver):
    """Gridap.Algebra.LUSolver (dep): sparse direct solve; SuperLU stands in for UMFPACK."""

    def _numerical_setup(self, A):
        return LUNS(A)


class LUNS:
    def __init__(self, A):
        self.update(A)

    def update(self, A):
        self.lu = spla.splu(sp.csc_matrix(A.to_scipy()))
        return self

    def solve(self, x, b):
        x[:] = self.lu.solve(b)
        return x


class RichardsonSmoother(_Solver):  # RichardsonSmoothers.jl:20-38
    def __init__(self, M, niter=1, omega=1.0):
        self.M, self.niter, self.omega = M, int(niter), float(omega)

    def _numerical_setup(self, A):  # :58-63
        return RichardsonSmootherNS(self, A, allocate_in_range(A), allocate_in_domain(A), _setup(self.M, A))


class RichardsonSmootherNS:
    def __init__(self, smoother, A, Adx, dx, Mns):
        self.smoother, self.A, self.Adx, self.dx, self.Mns = smoother, A, Adx, dx, Mns

    def update(self, A):  # :72-76
        self.Mns.update(A)
        self.A = A
        return self

    def solve(self, x, r):  # :84-98 -- mutates x AND r
        Adx, dx, Mns = self.Adx, self.dx, self.Mns
        niter, w = self.smoother.niter, self.smoother.omega
        it = 1
        dx[:] = 0.0
        while it <= niter:
            Mns.solve(dx, r)  # solve!(dx,Mns,r)
            la.scale(dx, w, dx)  # dx .= w .* dx
            la.add(x, x, dx)  # x  .= x .+ dx
            la.mul(Adx, self.A, dx)  # mul!(Adx, A, dx)
            la.sub(r, r, Adx)  # r  .= r .- Adx
            it += 1
        return x


class LinearSolverFromSmoother(_Solver):  # LinearSolverFromSmoothers.jl:1-50
    def __init__(self, smoother):
        self.smoother = smoother

    def _numerical_setup(self, A):
        return LinearSolverFromSmootherNS(_setup(self.smoother, A), allocate_in_domain(A))


class LinearSolverFromSmootherNS:
    def __init__(self, ns, r):
        self.smoother_ns, self.r = ns, r

    def update(self, A):
        self.smoother_ns.update(A)
        return self

    def solve(self, x, b):  # :44-50
        x[:] = 0.0
        self.r[:] = b
        self.smoother_ns.solve(x, self.r)
        return x


# ------------------------------------------------------------------ transfer operators


class MatrixTransfer:
    """Stand-in for DistributedGridTransferOperator: `mul!(y,op,x)` with an explicit sparse matrix
    (legal: GMG only needs mul!, GMGLinearSolvers.jl:484,491).  Prolongation == P
    (GridTransferOperators.jl:391-401), mode=:residual restriction == P^T (:206-208,:536-561)."""

    def __init__(self, M):
        self.M = M if isinstance(M, CSR) else CSR(M)

    def mul(self, y, x):
        return la.mul(y, self.M, x)


# ------------------------------------------------------------------ GMG


class GMGLinearSolver(_Solver):  # GMGLinearSolvers.jl:48-69
    def __init__(self, smatrices, interp, restrict, pre_smoothers=None, post_smoothers=None,
                 coarsest_solver=None, mode="preconditioner", cycle_type="v_cycle",
                 maxiter=100, atol=1.0e-14, rtol=1.0e-08, verbose=0):
        n = len(smatrices)
        if pre_smoothers is None:
            pre_smoothers = [RichardsonSmoother(JacobiLinearSolver(), 10)] * (n - 1)  # Fill(...) :52
        if post_smoothers is None:
            post_smoothers = pre_smoothers
        assert n - 1 == len(interp) == len(restrict) == len(pre_smoothers) == len(post_smoothers)
        assert mode in ("preconditioner", "solver") and cycle_type in ("v_cycle", "w_cycle", "f_cycle")
        self.smatrices = list(smatrices)
        self.interp = [t if hasattr(t, "mul") else MatrixTransfer(t) for t in interp]
        self.restrict = [t if hasattr(t, "mul") else MatrixTransfer(t) for t in restrict]
        self.pre_smoothers, self.post_smoothers = pre_smoothers, post_smoothers
        self.coarsest_solver = coarsest_solver if coarsest_solver is not None else LUSolver()
        self.mode, self.cycle_type = mode, cycle_type
        self.log = ConvergenceLog("GMG", SolverTolerances(maxiter=maxiter, atol=atol, rtol=rtol), verbose=verbose)

    def _numerical_setup(self, mat):  # :183-210
        s = self
        s.smatrices[0] = mat  # gmg_compute_matrices :336-340
        sm = s.smatrices
        nlev = len(sm)
        finest = allocate_in_domain(sm[0])  # :391-396
        work = []  # :451-466
        for lev in range(nlev - 1):
            dxh, Adxh = allocate_in_domain(sm[lev]), allocate_in_range(sm[lev])
            rH, dxH = allocate_in_domain(sm[lev + 1]), allocate_in_domain(sm[lev + 1])
            work.append((dxh, Adxh, dxH, rH))
        pre = [_setup(sm_, A) for sm_, A in zip(s.pre_smoothers, sm[: nlev - 1])]  # :398-408
        if s.pre_smoothers is s.post_smoothers:  # :190-194
            post = pre
        else:
            post = [_setup(sm_, A) for sm_, A in zip(s.post_smoothers, sm[: nlev - 1])]
        coarse = _setup(s.coarsest_solver, sm[nlev - 1])  # :423-434
        return GMGNS(s, sm, finest, pre, post, coarse, work)


class GMGNS:
    def __init__(self, solver, smatrices, finest, pre, post, coarse, work):
        self.solver, self.smatrices, self.finest_level_cache = solver, smatrices, finest
        self.pre, self.post, self.coarse, self.work = pre, post, coarse, work

    def update(self, A):  # :249-258: logs an error, does not throw
        raise NotImplementedError("GMGLinearSolverFromMatrices does not support updates")

    def _cycle(self, kind, lev, xh, rh):
        ns, s = self, self.solver
        nlev = len(self.smatrices)
        if lev == nlev - 1:  # coarsest :472-474
            ns.coarse.solve(xh, rh)
            return
        Ah = ns.smatrices[lev]
        restrict, interp = s.restrict[lev], s.interp[lev]
        dxh, Adxh, dxH, rH = ns.work[lev]
        ns.pre[lev].solve(xh, rh)  # :481
        restrict.mul(rH, rh)  # :484
        dxH[:] = 0.0  # :487
        first = {"v_cycle": "v_cycle", "w_cycle": "w_cycle", "f_cycle": "f_cycle"}[kind]
        self._cycle(first, lev + 1, dxH, rH)  # :488 / :524 / :578
        interp.mul(dxh, dxH)  # :491
        la.add(xh, xh, dxh)  # :494
        la.mul(Adxh, Ah, dxh)  # :495
        la.sub(rh, rh, Adxh)  # :496
        if kind != "v_cycle":  # W: :533-551, F: :587-605
            ns.post[lev].solve(xh, rh)  # re-smooth
            restrict.mul(rH, rh)
            dxH[:] = 0.0
            self._cycle("w_cycle" if kind == "w_cycle" else "v_cycle", lev + 1, dxH, rH)
            interp.mul(dxh, dxH)
            la.add(xh, xh, dxh)
            la.mul(Adxh, Ah, dxh)
            la.sub(rh, rh, Adxh)
        ns.post[lev].solve(xh, rh)  # :499

    def solve(self, x, b):  # :612-645
        s = self.solver
        log = s.log
        rh = self.finest_level_cache
        if s.mode == "preconditioner":
            x[:] = 0.0
            rh[:] = b
        else:
            la.mul(rh, self.smatrices[0], x)
            la.sub(rh, b, rh)
        res = la.norm(rh)
        done = init_(log, res)
        while not done:
            self._cycle(s.cycle_type, 0, x, rh)
            res = la.norm(rh)
            done = update_(log, res)
        finalize_(log, res)
        return x


# ------------------------------------------------------------------ Krylov utils


def krylov_mul_(y, A, x, Pr, Pl, wr, wl):  # KrylovUtils.jl:17-32
    if Pr is not None and Pl is not None:
        Pr.solve(wr, x)
        la.mul(wl, A, wr)
        Pl.solve(y, wl)
    elif Pr is not None:
        Pr.solve(wr, x)
        la.mul(y, A, wr)
    elif Pl is not None:
        la.mul(wl, A, x)
        Pl.solve(y, wl)
    else:
        la.mul(y, A, x)


def krylov_residual_(r, x, A, b, Pl, w):  # KrylovUtils.jl:46-54
    if Pl is not None:
        la.mul(w, A, x)
        la.sub(w, b, w)
        Pl.solve(r, w)
    else:
        la.mul(r, A, x)
        la.sub(r, b, r)


def givens_algorithm(f: float, g: float):
    """LinearAlgebra.givensAlgorithm(f,g) for reals (Julia stdlib givens.jl, a port of LAPACK 3.x
    dlartg), normal-range branch; the under/overflow rescaling loops are not needed for the
    magnitudes met here and are asserted away."""
    if g == 0.0:
        return 1.0, 0.0, f
    if f == 0.0:
        return 0.0, 1.0, g
    scale = max(abs(f), abs(g))
    assert 1e-140 < scale < 1e140, "givens rescaling branch not restated"
    r = math.sqrt(f * f + g * g)
    cs, sn = f / r, g / r
    if abs(f) > abs(g) and cs < 0:
        cs, sn, r = -cs, -sn, -r
    return cs, sn, r


# ------------------------------------------------------------------ CG


class LanczosDiagnostic:  # KrylovUtils.jl:58-90
    def __init__(self, max_iters):
        self.k = 0
        self.delta = np.zeros(max_iters)
        self.gamma = np.zeros(max_iters)

    def reset(self):
        self.k = 0
        self.delta[:] = 0
        self.gamma[:] = 0

    def update(self, d, g):
        self.delta[self.k] = d
        self.gamma[self.k] = g
        self.k += 1

    def estimate(self):
        k = self.k
        if k < 2:
            return 1.0
        import scipy.linalg as sl

        lam = sl.eigvalsh_tridiagonal(self.delta[:k], self.gamma[1:k])
        return abs(lam.max() / lam.min())


class CGSolver(_Solver):  # CGSolvers.jl:10-23
    def __init__(self, Pl=None, maxiter=1000, atol=1e-12, rtol=1.0e-6, diagnostic=None, flexible=False,
                 verbose=0, name="CG"):
        self.Pl, self.flexible, self.diag = Pl, flexible, diagnostic
        self.log = ConvergenceLog(name, SolverTolerances(maxiter=maxiter, atol=atol, rtol=rtol), verbose=verbose)

    def _numerical_setup(self, A):  # :50-55
        caches = tuple(allocate_in_domain(A) for _ in range(4))
        return CGNS(self, A, _setup(self.Pl, A), caches)


class CGNS:
    def __init__(self, solver, A, Pl_ns, caches):
        self.solver, self.mat, self.Pl_ns, self.caches = solver, A, Pl_ns, caches

    def update(self, A):  # :57-63
        if self.Pl_ns is not None:
            self.Pl_ns.update(A)
        self.mat = A
        return self

    def solve(self, x, b):  # :73-120
        solver, A, Pl = self.solver, self.mat, self.Pl_ns
        flexible, log = solver.flexible, solver.log
        w, p, z, r = self.caches
        la.mul(w, A, x)
        la.sub(r, b, w)  # :79
        p[:] = 0.0
        z[:] = 0.0
        gamma = 1.0
        alpha_last = 1.0
        res = la.norm(r)
        done = init_(log, res)
        if solver.diag is not None:
            solver.diag.reset()
        while not done:
            if Pl is None:
                z[:] = r
                beta = gamma
                gamma = la.dot(r, r)
                beta = gamma / beta
            elif not flexible:
                Pl.solve(z, r)
                beta = gamma
                gamma = la.dot(z, r)
                beta = gamma / beta
            else:
                delta = la.dot(z, r)
                Pl.solve(z, r)
                beta = gamma
                gamma = la.dot(z, r)
                beta = (gamma - delta) / beta
            la.axpy(p, z, beta, p)  # p .= z .+ beta .* p
            la.mul(w, A, p)
            alpha = gamma / la.dot(p, w)
            la.axpy(x, x, alpha, p)  # x .+= alpha .* p
            la.axmy(r, r, alpha, w)  # r .-= alpha .* w
            res = la.norm(r)
            done = update_(log, res)
            if solver.diag is not None:  # :122-138
                d = solver.diag
                if d.k == 0:
                    d.update(1.0 / alpha, 0.0)
                else:
                    d.update((1.0 / alpha) + (beta / alpha_last), math.sqrt(beta) / alpha)
            alpha_last = alpha
        finalize_(log, res)
        return x


# ------------------------------------------------------------------ GMRES / FGMRES


class GMRESSolver(_Solver):  # GMRESSolvers.jl:16-29
    flexible_basis = False

    def __init__(self, m, Pr=None, Pl=None, restart=False, m_add=1, maxiter=100, atol=1e-12, rtol=1.0e-6,
                 verbose=0, name="GMRES"):
        self.m, self.restart, self.m_add, self.Pr, self.Pl = int(m), restart, int(m_add), Pr, Pl
        self.log = ConvergenceLog(name, SolverTolerances(maxiter=maxiter, atol=atol, rtol=rtol), verbose=verbose)

    def _numerical_setup(self, A):  # :57-69,94-100
        m = self.m
        V = [allocate_in_domain(A) for _ in range(m + 1)]
        Z = [allocate_in_domain(A) for _ in range(m)] if self.flexible_basis else None
        zr = allocate_in_domain(A) if (self.Pr is not None and not self.flexible_basis) else None
        zl = allocate_in_domain(A)
        H, g, c, s = np.zeros((m + 1, m)), np.zeros(m + 1), np.zeros(m), np.zeros(m)
        return GMRESNS(self, A, _setup(self.Pr, A), _setup(self.Pl, A), [V, Z, zr, zl, H, g, c, s])


class FGMRESSolver(GMRESSolver):  # FGMRESSolvers.jl:17-30
    flexible_basis = True

    def __init__(self, m, Pr, Pl=None, restart=False, m_add=1, maxiter=100, atol=1e-12, rtol=1.0e-6,
                 verbose=0, name="FGMRES"):
        super().__init__(m, Pr=Pr, Pl=Pl, restart=restart, m_add=m_add, maxiter=maxiter, atol=atol, rtol=rtol,
                         verbose=verbose, name=name)


class GMRESNS:
    def __init__(self, solver, A, Pr_ns, Pl_ns, caches):
        self.solver, self.mat, self.Pr_ns, self.Pl_ns, self.caches = solver, A, Pr_ns, Pl_ns, caches

    def update(self, A):
        if self.Pr_ns is not None:
            self.Pr_ns.update(A)
        if self.Pl_ns is not None:
            self.Pl_ns.update(A)
        self.mat = A
        return self

    def _expand(self):  # GMRESSolvers.jl:76-92 / FGMRESSolvers.jl:77-95
        V, Z, zr, zl, H, g, c, s = self.caches
        m = len(V) - 1
        m_new = m + self.solver.m_add
        for _ in range(self.solver.m_add):
            V.append(allocate_in_domain(self.mat))
            if Z is not None:
                Z.append(allocate_in_domain(self.mat))
        Hn = np.zeros((m_new + 1, m_new)); Hn[: m + 1, :m] = H
        gn = np.zeros(m_new + 1); gn[: m + 1] = g
        cn = np.zeros(m_new); cn[:m] = c
        sn = np.zeros(m_new); sn[:m] = s
        self.caches = [V, Z, zr, zl, Hn, gn, cn, sn]

    def solve(self, x, b):  # GMRESSolvers.jl:132-210 ; FGMRESSolvers.jl:130-199
        solver, A, Pl, Pr = self.solver, self.mat, self.Pl_ns, self.Pr_ns
        flex = solver.flexible_basis
        V, Z, zr, zl, H, g, c, s = self.caches
        m = len(V) - 1
        log = solver.log
        V[0][:] = 0.0
        if zr is not None:
            zr[:] = 0.0
        zl[:] = 0.0
        krylov_residual_(V[0], x, A, b, Pl, zl)
        beta = la.norm(V[0])
        done = init_(log, beta)
        while not done:
            j = 1
            np.divide(V[0], beta, out=V[0])  # V[1] ./= beta
            H[:] = 0.0
            g[:] = 0.0
            g[0] = beta
            while (not done) and not (solver.restart and j > solver.m):
                if j > m:
                    self._expand()
                    V, Z, zr, zl, H, g, c, s = self.caches
                    m = len(V) - 1
                V[j][:] = 0.0
                if flex:
                    Z[j - 1][:] = 0.0  # FGMRESSolvers.jl:158
                    krylov_mul_(V[j], A, V[j - 1], Pr, Pl, Z[j - 1], zl)
                else:
                    krylov_mul_(V[j], A, V[j - 1], Pr, Pl, zr, zl)  # zr NOT re-zeroed :161
                for i in range(j):  # modified Gram-Schmidt :162-165
                    H[i, j - 1] = la.dot(V[j], V[i])
                    la.axmy(V[j], V[j], H[i, j - 1], V[i])
                H[j, j - 1] = la.norm(V[j])
                np.divide(V[j], H[j, j - 1], out=V[j])
                for i in range(j - 1):  # :170-174
                    gam = c[i] * H[i, j - 1] + s[i] * H[i + 1, j - 1]
                    H[i + 1, j - 1] = -s[i] * H[i, j - 1] + c[i] * H[i + 1, j - 1]
                    H[i, j - 1] = gam
                c[j - 1], s[j - 1], _ = givens_algorithm(H[j - 1, j - 1], H[j, j - 1])  # :177
                H[j - 1, j - 1] = c[j - 1] * H[j - 1, j - 1] + s[j - 1] * H[j, j - 1]
                H[j, j - 1] = 0.0
                g[j] = -s[j - 1] * g[j - 1]
                g[j - 1] = c[j - 1] * g[j - 1]
                beta = abs(g[j])
                j += 1
                done = update_(log, beta)
            j = j - 1
            for i in range(j - 1, -1, -1):  # :188-190
                g[i] = (g[i] - np.dot(H[i, i + 1: j], g[i + 1: j])) / H[i, i]
            if flex:
                for i in range(j):
                    la.axpy(x, x, g[i], Z[i])
            elif Pr is None:
                for i in range(j):
                    la.axpy(x, x, g[i], V[i])
            else:
                zl[:] = 0.0
                for i in range(j):
                    la.axpy(zl, zl, g[i], V[i])
                Pr.solve(zr, zl)
                la.add(x, x, zr)
            krylov_residual_(V[0], x, A, b, Pl, zl)
        finalize_(log, beta)
        return x


# ------------------------------------------------------------------ MINRES


class MINRESSolver(_Solver):  # MINRESSolvers.jl:11-20
    def __init__(self, Pl=None, maxiter=1000, atol=1e-12, rtol=1.0e-6, verbose=0, name="MINRES"):
        self.Pl = Pl
        self.log = ConvergenceLog(name, SolverTolerances(maxiter=maxiter, atol=atol, rtol=rtol), verbose=verbose)

    def _numerical_setup(self, A):  # :39-44
        caches = [[allocate_in_domain(A) for _ in range(3)] for _ in range(3)]
        return MINRESNS(self, A, _setup(self.Pl, A), caches)


class MINRESNS:
    def __init__(self, solver, A, Pl_ns, caches):
        self.solver, self.A, self.Pl_ns, self.caches = solver, A, Pl_ns, caches

    def update(self, A):
        if self.Pl_ns is not None:
            self.Pl_ns.update(A)
        self.A = A
        return self

    def solve(self, x, b):  # :75-149
        solver, A, Pl = self.solver, self.A, self.Pl_ns
        Vs, Ws, Zs = self.caches
        log = solver.log
        Vnew, V, Vold = Vs
        Wnew, W, Wold = Ws
        Znew, Z, Zold = Zs
        W[:] = 0.0; Wold[:] = 0.0; Vold[:] = 0.0; Zold[:] = 0.0
        la.mul(Vnew, A, x)
        la.sub(Vnew, b, Vnew)
        Znew[:] = 0.0
        if Pl is not None:
            Pl.solve(Znew, Vnew)
        else:
            Znew[:] = Vnew
        beta_r = la.norm(Znew)
        beta_p = la.dot(Znew, Vnew)
        assert beta_p > 0.0  # @check :97
        gnew, gam, gold = 0.0, math.sqrt(beta_p), 1.0
        cnew, c, cold = 0.0, 1.0, 1.0
        snew, s, sold = 0.0, 0.0, 0.0
        np.divide(Vnew, gam, out=V)
        np.divide(Znew, gam, out=Z)
        eta = gam
        done = init_(log, beta_r)
        while not done:
            la.mul(Vnew, A, Z)
            if Pl is not None:
                Pl.solve(Znew, Vnew)
            else:
                Znew[:] = Vnew
            delta = la.dot(Vnew, Z)
            # Vnew .= Vnew .- delta .* V .- gam .* Vold  (left-to-right)
            la.axmy(Vnew, Vnew, delta, V); la.axmy(Vnew, Vnew, gam, Vold)
            la.axmy(Znew, Znew, delta, Z); la.axmy(Znew, Znew, gam, Zold)
            beta_p = la.dot(Znew, Vnew)
            gnew = math.sqrt(beta_p)
            np.divide(Vnew, gnew, out=Vnew)
            np.divide(Znew, gnew, out=Znew)
            a0 = c * delta - cold * s * gam
            cnew, snew, a1 = givens_algorithm(a0, gnew)
            a2 = s * delta + cold * c * gam
            a3 = sold * gam
            # Wnew .= (Z .- a2 .* W .- a3 .* Wold) ./ a1
            la.axmy(Wnew, Z, a2, W); la.axmy(Wnew, Wnew, a3, Wold)
            np.divide(Wnew, a1, out=Wnew)
            la.axpy(x, x, cnew * eta, Wnew)
            eta = -snew * eta
            beta_r = abs(snew) * beta_r
            # swap3(xnew,x,xold) = xold, xnew, x
            Vnew, V, Vold = Vold, Vnew, V
            Wnew, W, Wold = Wold, Wnew, W
            Znew, Z, Zold = Zold, Znew, Z
            gnew, gam, gold = gold, gnew, gam
            cnew, c, cold = cold, cnew, c
            snew, s, sold = sold, snew, s
            done = update_(log, beta_r)
        finalize_(log, beta_r)
        return x


# ------------------------------------------------------------------ block solvers (C5)


class BlockTriangularSolver(_Solver):
    """BlockTriangularSolvers.jl:26-58,135-143,188-242 with LinearSystemBlock()s: `mats[i][j]`
    are the blocks of the system matrix; diagonal solvers act on mats[i][i] unless `diag_mats`
    overrides them (MatrixBlock / BiformBlock, BlockSolverInterfaces.jl:162,262)."""

    def __init__(self, solvers, coeffs=None, half="upper", diag_mats=None):
        self.solvers, self.half, self.diag_mats = solvers, half, diag_mats
        n = len(solvers)
        self.coeffs = np.ones((n, n)) if coeffs is None else np.asarray(coeffs, dtype=float)

    def _numerical_setup(self, mats):
        n = len(self.solvers)
        dm = [(self.diag_mats[i] if (self.diag_mats and self.diag_mats[i] is not None) else mats[i][i]) for i in range(n)]
        block_ns = [_setup(self.solvers[i], dm[i]) for i in range(n)]
        w = [np.zeros(dm[i].shape[0]) for i in range(n)]
        y = [np.zeros(dm[i].shape[0]) for i in range(n)]  # zeroed only at setup :139
        return BlockTriangularNS(self, mats, block_ns, w, y)


class BlockTriangularNS:
    def __init__(self, solver, mats, block_ns, w, y):
        self.solver, self.mats, self.block_ns, self.w, self.y = solver, mats, block_ns, w, y

    def solve(self, x, b):  # x, b: lists of block vectors
        NB = len(self.block_ns)
        c = self.solver.coeffs
        order = range(NB) if self.solver.half == "lower" else range(NB - 1, -1, -1)
        for iB in order:
            wi = self.w[iB]
            wi[:] = b[iB]
            js = range(iB) if self.solver.half == "lower" else range(iB + 1, NB)
            for jB in js:
                cij = c[iB, jB]
                if abs(cij) > np.spacing(abs(cij)):  # abs(cij) > eps(cij)
                    la.mul5(wi, self.mats[iB][jB], x[jB], -cij, 1.0)
            self.block_ns[iB].solve(self.y[iB], wi)
            x[iB][:] = self.y[iB]
        return x


class BlockDiagonalSolver(_Solver):  # BlockDiagonalSolvers.jl:22-45,165-177
    def __init__(self, solvers, diag_mats=None):
        self.solvers, self.diag_mats = solvers, diag_mats

    def _numerical_setup(self, mats):
        n = len(self.solvers)
        dm = [(self.diag_mats[i] if (self.diag_mats and self.diag_mats[i] is not None) else mats[i][i]) for i in range(n)]
        return BlockDiagonalNS([_setup(self.solvers[i], dm[i]) for i in range(n)],
                               [np.zeros(dm[i].shape[0]) for i in range(n)])


class BlockDiagonalNS:
    def __init__(self, block_ns, y):
        self.block_ns, self.y = block_ns, y

    def solve(self, x, b):
        for iB, bns in enumerate(self.block_ns):
            bns.solve(self.y[iB], b[iB])
            x[iB][:] = self.y[iB]
        return x


class BlockMatrix:
    """Minimal BlockArrays.BlockMatrix stand-in so Krylov solvers can run on block systems:
    vectors are single concatenated arrays; `mul!` loops over blocks."""

    def __init__(self, blocks):
        self.blocks = blocks
        self.rsizes = [next(b for b in row if b is not None).shape[0] for row in blocks]
        ncol = len(blocks[0])
        self.csizes = [next(blocks[i][j] for i in range(len(blocks)) if blocks[i][j] is not None).shape[1] for j in range(ncol)]
        self.shape = (sum(self.rsizes), sum(self.csizes))
        full = sp.bmat([[None if b is None else b.to_scipy() for b in row] for row in blocks], format="csr")
        self.full = CSR(full)
        # so that la.mul works on it directly
        self.rowptr, self.col, self.val, self.sp = self.full.rowptr, self.full.col, self.full.val, self.full.sp

    def diag(self):
        return self.full.diag()

    def to_scipy(self):
        return self.full.sp

    def split(self, v, sizes=None):
        sizes = sizes or self.csizes
        offs = np.concatenate([[0], np.cumsum(sizes)])
        return [v[offs[i]: offs[i + 1]] for i in range(len(sizes))]

    def __getitem__(self, i):
        return self.blocks[i]

    def __len__(self):
        return len(self.blocks)


class BlockPrecondAdapter(_Solver):
    """Lets a block solver be used as Pr/Pl of a Krylov solver on concatenated vectors."""

    def __init__(self, block_solver):
        self.block_solver = block_solver

    def _numerical_setup(self, A):
        ns = self.block_solver._numerical_setup(A)
        return _BlockAdapterNS(ns, A)


class _BlockAdapterNS:
    def __init__(self, ns, A):
        self.ns, self.A = ns, A

    def solve(self, x, b):
        self.ns.solve(self.A.split(x), self.A.split(b, self.A.rsizes))
        return x


# ------------------------------------------------------------------ "next" rows


class RichardsonLinearSolver(_Solver):
    """RichardsonLinearSolvers.jl:12-23 (scalar relaxation parameter)."""

    def __init__(self, omega, maxiter, Pl=None, rtol=1e-10, atol=1e-6, verbose=0, name="RichardsonLinearSolver"):
        self.omega, self.Pl = float(omega), Pl
        self.log = ConvergenceLog(name, SolverTolerances(maxiter=maxiter, atol=atol, rtol=rtol), verbose=verbose)

    def _numerical_setup(self, A):
        return RichardsonLinearNS(self, A, _setup(self.Pl, A), allocate_in_domain(A), allocate_in_domain(A))


class RichardsonLinearNS:
    def __init__(self, solver, A, Pl_ns, z, r):
        self.solver, self.A, self.Pl_ns, self.z, self.r = solver, A, Pl_ns, z, r

    def update(self, A):
        if self.Pl_ns is not None:
            self.Pl_ns.update(A)
        self.A = A
        return self

    def solve(self, x, b):  # :79-106
        s, A, Pl, z, r, log = self.solver, self.A, self.Pl_ns, self.z, self.r, self.solver.log
        r[:] = b
        la.mul5(r, A, x, -1.0, 1.0)  # mul!(r, A, x, -1, 1)
        done = init_(log, la.norm(r))
        while not done:
            if Pl is not None:
                Pl.solve(z, r)
                la.axpy(x, x, s.omega, z)  # x .+= w .* z
            else:
                la.axpy(x, x, s.omega, r)
            r[:] = b
            la.mul5(r, A, x, -1.0, 1.0)
            done = update_(log, la.norm(r))
        finalize_(log, la.norm(r))
        return x


class SchurComplementSolver(_Solver):
    """SchurComplementSolvers.jl:8-24: A, S are NumericalSetups, B, C matrices; acts on [u; p]."""

    def __init__(self, A_ns, B, C, S_ns):
        self.A, self.B, self.C, self.S = A_ns, B, C, S_ns

    def _numerical_setup(self, mat):
        return SchurComplementNS(self, allocate_in_domain(self.C), allocate_in_domain(self.C), allocate_in_domain(self.B))


class SchurComplementNS:
    def __init__(self, solver, du, bu, bp):
        self.s, self.du, self.bu, self.bp = solver, du, bu, bp

    def solve(self, x, y):  # x, y: [u, p] lists of block vectors   :55-74
        s = self.s
        x_u, x_p = x
        y_u, y_p = y
        s.A.solve(x_u, y_u)
        self.bp[:] = y_p
        la.mul5(self.bp, s.C, x_u, -1.0, 1.0)
        s.S.solve(x_p, self.bp)
        la.mul(self.bu, s.B, x_p)
        s.A.solve(self.du, self.bu)
        la.sub(x_u, x_u, self.du)
        return x
