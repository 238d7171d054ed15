"""Host-side mirror of the reference's solver interface, on top of the C ABI (include/gsb200.h).

The reference is Julia; its toolchain is absent here, so this Python layer plays the role of the
Julia shim (gridapsolvers.jl_b200/julia/GridapSolversB200.jl binds the same entry points with
ccall): same names, argument meaning and error behaviour as the reference's
`Gridap.Algebra.LinearSolver` API -- `symbolic_setup / numerical_setup / numerical_setup! /
solve!` (Julia's `f!` is spelled `f_`), solver constructors with the reference's keyword
defaults, and the public `solver.log :: ConvergenceLog` filled after every solve.

Reference (paths relative to /root/reference/src):
  CGSolver               LinearSolvers/Krylov/CGSolvers.jl:10-23
  GMRESSolver            LinearSolvers/Krylov/GMRESSolvers.jl:16-29
  FGMRESSolver           LinearSolvers/Krylov/FGMRESSolvers.jl:17-30
  MINRESSolver           LinearSolvers/Krylov/MINRESSolvers.jl:11-20
  JacobiLinearSolver     LinearSolvers/JacobiLinearSolvers.jl:6
  RichardsonSmoother     LinearSolvers/RichardsonSmoothers.jl:20-38
  LinearSolverFromSmoother LinearSolvers/LinearSolverFromSmoothers.jl:1-3
  IdentitySolver         LinearSolvers/IdentityLinearSolvers.jl:2
  GMGLinearSolver        LinearSolvers/GMGLinearSolvers.jl:48-69
  BlockTriangularSolver  BlockSolvers/BlockTriangularSolvers.jl:26-58
  BlockDiagonalSolver    BlockSolvers/BlockDiagonalSolvers.jl:22-45
  SolverTolerances / ConvergenceLog  SolverInterfaces/{SolverTolerances,ConvergenceLogs}.jl
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _lib
from ._lib import c_p, check

GSB_FMT_CSR, GSB_FMT_CSC = 0, 1
SOLVER_CONVERGED_ATOL, SOLVER_CONVERGED_RTOL, SOLVER_DIVERGED_MAXITER, SOLVER_DIVERGED_BREAKDOWN = 0, 1, 2, 3
_MODES = {"preconditioner": 0, "solver": 1}
_CYCLES = {"v_cycle": 0, "w_cycle": 1, "f_cycle": 2}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(c_p)


# --------------------------------------------------------------------------- context


class Context:
    """One part of a PartitionedArrays distribution == one process == one B200."""

    _default = None

    def __init__(self, device: int = 0, nranks: int = 1, rank: int = 0, nccl_id: bytes | None = None):
        L = _lib.lib()
        h = c_p()
        idbuf = ctypes.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        check(L.gsb_init(device, nranks, rank, idbuf, ctypes.byref(h)))
        self.h, self.nranks, self.rank, self.device = h, nranks, rank, device

    @classmethod
    def default(cls) -> "Context":
        if cls._default is None:
            cls._default = Context()
        return cls._default

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        check(_lib.lib().gsb_nccl_unique_id(buf))
        return buf.raw

    def set_option(self, key: str, value) -> None:
        check(_lib.lib().gsb_set_option(self.h, key.encode(), str(value).encode()))

    def synchronize(self):
        check(_lib.lib().gsb_synchronize(self.h))

    def timer_start(self):
        check(_lib.lib().gsb_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = ctypes.c_float()
        check(_lib.lib().gsb_timer_stop(self.h, ctypes.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        n = ctypes.c_int64()
        check(_lib.lib().gsb_launch_count(self.h, ctypes.byref(n)))
        return int(n.value)

    def profile_start(self):
        check(_lib.lib().gsb_profile_start(self.h))

    def profile_stop(self, cap: int = 256):
        """-> list of dicts(mode, stream_kernel, nrows, nnz, count, total_ms) for the row kernels"""
        n = ctypes.c_int()
        mode, st, cnt = (np.zeros(cap, dtype=np.int32) for _ in range(3))
        nrows, nnz = (np.zeros(cap, dtype=np.int64) for _ in range(2))
        ms = np.zeros(cap)
        check(_lib.lib().gsb_profile_stop(self.h, cap, ctypes.byref(n), _ptr(mode), _ptr(st), _ptr(nrows), _ptr(nnz),
                                          _ptr(cnt), _ptr(ms)))
        names = {0: "spmv", 1: "residual", 2: "sweep", 3: "spmv_dot", 4: "spmv_add"}

        def impl(k):  # 0 CSR fallback kernel; 2 + 10*block_size (+100 sorted rows) block-SELL-32
            if k == 0:
                return "csr_vector"
            bs = (k % 100 - 2) // 10
            return "bsell32_%dx%d%s" % (bs, bs, "_sorted" if k >= 100 else "")

        return [dict(mode=names[int(mode[i])], impl=impl(int(st[i])), nrows=int(nrows[i]), nnz=int(nnz[i]),
                     count=int(cnt[i]), total_ms=float(ms[i])) for i in range(n.value)]

    def host_register(self, array: np.ndarray):
        """page-lock a caller-owned numpy buffer (pinned-speed gsb_solve_host / set / get)"""
        check(_lib.lib().gsb_host_register(self.h, _ptr(array), array.nbytes))

    def host_unregister(self, array: np.ndarray):
        check(_lib.lib().gsb_host_unregister(self.h, _ptr(array)))

    def close(self):
        if self.h:
            _lib.lib().gsb_finalize(self.h)
            self.h = None


# --------------------------------------------------------------------------- PartitionedArrays mirrors


class ExchangePlan:
    """consistent!/assemble! cache of an index partition (neighbours + local id lists)."""

    def __init__(self, ctx, n_own, n_ghost, nbr_snd, snd_ptrs, snd_ids, nbr_rcv, rcv_ptrs, rcv_ids, index_base=0):
        self.ctx = ctx
        a32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
        a64 = lambda v: np.ascontiguousarray(v, dtype=np.int64)
        self._keep = (a32(nbr_snd), a64(snd_ptrs), a64(snd_ids), a32(nbr_rcv), a64(rcv_ptrs), a64(rcv_ids))
        k = self._keep
        h = c_p()
        check(_lib.lib().gsb_plan_create(ctx.h, n_own, n_ghost, len(k[0]), _ptr(k[0]), _ptr(k[1]), _ptr(k[2]),
                                         len(k[3]), _ptr(k[3]), _ptr(k[4]), _ptr(k[5]), index_base, ctypes.byref(h)))
        self.h, self.n_own, self.n_ghost = h, n_own, n_ghost

    def __del__(self):
        try:
            if self.h:
                _lib.lib().gsb_plan_destroy(self.h)
        except Exception:
            pass


class RedistributionPlan:
    """Moves own values between two row partitions of the same global index space (MultilevelTools'
    RedistributionOperator / redistribute_free_values!, GridTransferOperators.jl:447-532): one direction per plan."""

    def __init__(self, ctx, n_src_own, n_dst_own, nbr_snd, snd_ptrs, snd_ids, nbr_rcv, rcv_ptrs, rcv_ids, index_base=0):
        self.ctx = ctx
        a32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
        a64 = lambda v: np.ascontiguousarray(v, dtype=np.int64)
        self._keep = (a32(nbr_snd), a64(snd_ptrs), a64(snd_ids), a32(nbr_rcv), a64(rcv_ptrs), a64(rcv_ids))
        k = self._keep
        h = c_p()
        check(_lib.lib().gsb_redist_create(ctx.h, n_src_own, n_dst_own, len(k[0]), _ptr(k[0]), _ptr(k[1]), _ptr(k[2]),
                                           len(k[3]), _ptr(k[3]), _ptr(k[4]), _ptr(k[5]), index_base, ctypes.byref(h)))
        self.h, self.n_src_own, self.n_dst_own = h, n_src_own, n_dst_own

    def __del__(self):
        try:
            if self.h:
                _lib.lib().gsb_plan_destroy(self.h)
        except Exception:
            pass


def redistribute_(dst, plan: "RedistributionPlan", src):
    """dst (own values, destination layout) <- src (own values, source layout)"""
    check(_lib.lib().gsb_vec_redistribute(plan.h, src.h, dst.h))
    return dst


class SparseMatrix:
    """Device mirror of the local block of a PSparseMatrix (own rows x own+ghost columns)."""

    def __init__(self, ctx, n_rows, n_own_cols, n_ghost_cols, ptr, idx, vals, fmt="csr", index_base=0, plan=None):
        ptr = np.ascontiguousarray(ptr)
        idx = np.ascontiguousarray(idx)
        if ptr.dtype not in (np.int32, np.int64):
            ptr = ptr.astype(np.int64)
        if idx.dtype == np.int32 and ptr.dtype == np.int64 and int(ptr[-1]) < 2**31 - 8:
            ptr = ptr.astype(np.int32)  # int32 ids are uploaded as they are (no widened copy of the id array)
        idx = idx.astype(ptr.dtype, copy=False)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        h = c_p()
        check(_lib.lib().gsb_mat_create(ctx.h, n_rows, n_own_cols, n_ghost_cols, GSB_FMT_CSC if fmt == "csc" else GSB_FMT_CSR,
                                        index_base, ptr.dtype.itemsize, _ptr(ptr), _ptr(idx), _ptr(vals),
                                        plan.h if plan is not None else None, ctypes.byref(h)))
        self.ctx, self.h, self.plan = ctx, h, plan
        self.n_rows, self.n_own_cols, self.n_ghost_cols = n_rows, n_own_cols, n_ghost_cols
        self.nnz = int(ptr[-1] - ptr[0])
        self.shape = (n_rows, n_own_cols)

    @classmethod
    def from_scipy(cls, A, ctx=None, n_ghost_cols=0, plan=None):
        """Serial (or local-block) matrix from a scipy CSR/CSC matrix."""
        import scipy.sparse as sp

        ctx = ctx or Context.default()
        if sp.isspmatrix_csc(A):
            return cls(ctx, A.shape[0], A.shape[1] - n_ghost_cols, n_ghost_cols, A.indptr, A.indices, A.data, fmt="csc", plan=plan)
        A = sp.csr_matrix(A)
        return cls(ctx, A.shape[0], A.shape[1] - n_ghost_cols, n_ghost_cols, A.indptr, A.indices, A.data, fmt="csr", plan=plan)

    def bench_rows(self, mode: str, reps: int = 20) -> float:
        """average ms of one row-kernel launch (diagnostics / roofline numbers)"""
        ms = ctypes.c_float()
        k = {"spmv": 0, "residual": 1, "sweep": 2, "spmv_dot": 3, "spmv_add": 4}[mode]
        check(_lib.lib().gsb_bench_rows(self.h, k, reps, ctypes.byref(ms)))
        return float(ms.value)

    def format(self) -> dict:
        """device storage streamed by the row kernels (gsb_mat_format)"""
        kind, bs, srt = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        ent, byt = ctypes.c_int64(), ctypes.c_int64()
        check(_lib.lib().gsb_mat_format(self.h, ctypes.byref(kind), ctypes.byref(bs), ctypes.byref(srt), ctypes.byref(ent), ctypes.byref(byt)))
        return dict(kind="bsell32" if kind.value == 1 else "csr", block_size=bs.value, sorted=bool(srt.value),
                    stored_entries=ent.value, bytes_per_pass=byt.value)

    def update_values(self, vals):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        assert vals.shape[0] == self.nnz
        check(_lib.lib().gsb_mat_update_values(self.h, _ptr(vals)))

    def __del__(self):
        try:
            if self.h:
                _lib.lib().gsb_mat_destroy(self.h)
        except Exception:
            pass


class BlockSparseMatrix:
    """BlockMatrix of SparseMatrix blocks acting on concatenated vectors (C5)."""

    def __init__(self, blocks, ctx=None):
        self.blocks = blocks
        nb = len(blocks)
        first = next(b for row in blocks for b in row if b is not None)
        self.ctx = ctx or first.ctx
        arr = (c_p * (nb * nb))(*[(b.h if b is not None else None) for row in blocks for b in row])
        h = c_p()
        check(_lib.lib().gsb_block_mat_create(self.ctx.h, nb, arr, ctypes.byref(h)))
        self.h, self.plan = h, None
        nr, nc, ng, nnz = (ctypes.c_int64() for _ in range(4))
        check(_lib.lib().gsb_mat_info(h, ctypes.byref(nr), ctypes.byref(nc), ctypes.byref(ng), ctypes.byref(nnz)))
        self.n_rows, self.n_own_cols, self.n_ghost_cols, self.nnz = nr.value, nc.value, 0, nnz.value
        self.shape = (self.n_rows, self.n_own_cols)

    def __getitem__(self, i):
        return self.blocks[i]


class Vector:
    """Device mirror of the local part of a PVector: own values first, ghost values after."""

    def __init__(self, ctx, n_own, n_ghost=0, h=None):
        if h is None:
            h = c_p()
            check(_lib.lib().gsb_vec_create(ctx.h, n_own, n_ghost, ctypes.byref(h)))
        self.ctx, self.h, self.n_own, self.n_ghost = ctx, h, n_own, n_ghost

    def set(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        check(_lib.lib().gsb_vec_set(self.h, _ptr(v), v.shape[0]))
        return self

    def get(self) -> np.ndarray:
        out = np.empty(self.n_own)
        check(_lib.lib().gsb_vec_get(self.h, _ptr(out), self.n_own))
        return out

    def get_local(self) -> np.ndarray:
        out = np.empty(self.n_own + self.n_ghost)
        check(_lib.lib().gsb_vec_get_local(self.h, _ptr(out), out.shape[0]))
        return out

    def fill(self, value):
        check(_lib.lib().gsb_vec_fill(self.h, float(value)))
        return self

    def __del__(self):
        try:
            if self.h:
                _lib.lib().gsb_vec_destroy(self.h)
        except Exception:
            pass


def allocate_in_domain(A) -> Vector:
    h = c_p()
    check(_lib.lib().gsb_vec_create_domain(A.h, ctypes.byref(h)))
    return Vector(A.ctx, A.n_own_cols, A.n_ghost_cols, h)


def allocate_in_range(A) -> Vector:
    h = c_p()
    check(_lib.lib().gsb_vec_create_range(A.h, ctypes.byref(h)))
    return Vector(A.ctx, A.n_rows, 0, h)


def mul_(y: Vector, A, x: Vector, alpha: float = 1.0, beta: float = 0.0) -> Vector:
    """LinearAlgebra.mul!(y,A,x[,alpha,beta])"""
    check(_lib.lib().gsb_spmv(A.h, x.h, y.h, float(alpha), float(beta)))
    return y


def dot(a: Vector, b: Vector) -> float:
    out = ctypes.c_double()
    check(_lib.lib().gsb_dot(a.h, b.h, ctypes.byref(out)))
    return out.value


def norm(a: Vector) -> float:
    out = ctypes.c_double()
    check(_lib.lib().gsb_norm2(a.h, ctypes.byref(out)))
    return out.value


def axpby_(z: Vector, alpha, x: Vector, beta, y: Vector) -> Vector:
    """z .= alpha .* x .+ beta .* y"""
    check(_lib.lib().gsb_axpby(z.h, float(alpha), x.h, float(beta), y.h))
    return z


def copy_(dst: Vector, src: Vector) -> Vector:
    check(_lib.lib().gsb_vec_copy(dst.h, src.h))
    return dst


def consistent_(v: Vector, plan: ExchangePlan) -> Vector:
    check(_lib.lib().gsb_vec_consistent(v.h, plan.h if plan is not None else None))
    return v


def assemble_(v: Vector, plan: ExchangePlan) -> Vector:
    """assemble!(v) |> wait: ghost contributions added to their owners, ghosts zeroed"""
    check(_lib.lib().gsb_vec_assemble(v.h, plan.h if plan is not None else None))
    return v


# --------------------------------------------------------------------------- tolerances / logs


class SolverTolerances:
    """SolverInterfaces/SolverTolerances.jl:40-49"""

    def __init__(self, maxiter=1000, atol=np.finfo(np.float64).eps, rtol=1e-5, dtol=math.inf):
        self.maxiter, self.atol, self.rtol, self.dtol = int(maxiter), float(atol), float(rtol), float(dtol)


class ConvergenceLog:
    """SolverInterfaces/ConvergenceLogs.jl:42-60; filled from the device solve after solve!."""

    def __init__(self, name, tols, verbose=0, depth=0):
        self.name, self.tols = name, tols
        self.num_iters = 0
        self.residuals = np.zeros(tols.maxiter + 1)
        self.verbose, self.depth = int(verbose), depth
        self.flag = None

    def history(self):
        return self.residuals[: self.num_iters + 1].copy()

    def _fill(self, handle):
        n, flag = ctypes.c_int(), ctypes.c_int()
        self.residuals[:] = 0.0
        check(_lib.lib().gsb_solver_log(handle, ctypes.byref(n), _ptr(self.residuals), self.residuals.shape[0], ctypes.byref(flag)))
        self.num_iters, self.flag = n.value, flag.value
        if self.verbose > 1:  # ConvergenceLogs.jl:103-109,121-126 (printed after the device solve returns)
            t = " " * (2 + 2 * self.depth)
            print(" " * (2 * self.depth) + (("-" * 15) + f" Starting {self.name} solver ").ljust(55, "-"))
            r0 = self.residuals[0]
            for k in range(self.num_iters + 1):
                r = self.residuals[k]
                print(t + "> Iteration %3i - Residuals: %.2e,   %.2e " % (k, r, (r / r0) if (k and r0) else 1))
        if self.verbose > 0:  # ConvergenceLogs.jl:139-148
            t = " " * (2 * self.depth)
            r = self.residuals[self.num_iters]
            print(f"{t}Solver {self.name} finished with reason {self.flag}")
            print(t + "Iterations: %3i - Residuals: %.2e,   %.2e " % (self.num_iters, r, r / self.residuals[0] if self.residuals[0] else float('nan')))


def get_solver_tolerances(solver) -> SolverTolerances:
    """SolverTolerances.jl:51-56"""
    return solver.log.tols


def set_solver_tolerances_(solver, maxiter=1000, atol=np.finfo(np.float64).eps, rtol=1e-5, dtol=math.inf):
    """set_solver_tolerances!(s; maxiter, atol, rtol, dtol) -- SolverTolerances.jl:58-84 (takes effect at the next
    numerical_setup, which is when the device-side NumericalSetup copies the tolerances)."""
    t = solver.log.tols
    t.maxiter, t.atol, t.rtol, t.dtol = int(maxiter), float(atol), float(rtol), float(dtol)
    solver.log.residuals = np.zeros(t.maxiter + 1)
    return t


class HierarchicalArray:
    """MultilevelTools/HierarchicalArrays.jl:13-22: one entry per level plus the ranks taking part in it.
    `ranks[l]` is the collection of ranks that hold level l (None: every rank); on a rank outside it the entry is
    `None` -- the reference's `nothing` -- and `map` / `with_level` skip it (HierarchicalArrays.jl:96-149).  The device
    side of a level a rank does not hold is a zero-row matrix (synth.level_part_or_empty, gsb_gmg_create_redist)."""

    def __init__(self, array, ranks=None, rank=0):
        self.array = list(array)
        self.ranks = list(ranks) if ranks is not None else [None] * len(self.array)
        self.rank = rank
        assert len(self.ranks) == len(self.array)
        for l in range(len(self.array)):
            if not self.i_am_in(l):
                self.array[l] = None

    def i_am_in(self, l) -> bool:  # GridapDistributed.i_am_in(ranks[l]) of this process
        r = self.ranks[l]
        return r is None or self.rank in r

    def __len__(self):
        return len(self.array)

    def __getitem__(self, i):
        return self.array[i]

    def __setitem__(self, i, v):
        self.array[i] = v

    def __iter__(self):
        return iter(self.array)

    def map(self, f, *others):  # Base.map(f, args::HierarchicalArray...), HierarchicalArrays.jl:96-120
        for o in others:
            assert isinstance(o, HierarchicalArray) and o.ranks == self.ranks, "matching_level_parts"
        out = [f(self.array[l], *[o.array[l] for o in others]) if self.i_am_in(l) else None for l in range(len(self))]
        return HierarchicalArray(out, self.ranks, self.rank)


def num_levels(a) -> int:  # HierarchicalArrays.jl:71
    return len(a)


def with_level(f, a, lev, default=None):
    """with_level(f, a, lev; default=nothing), HierarchicalArrays.jl:139-149 (1-based level like the reference): `default`
    when this rank does not belong to the level's ranks; plain sequences always run `f`"""
    if isinstance(a, HierarchicalArray):
        if lev < 1 or lev > len(a) or not a.i_am_in(lev - 1):
            return default
        return f(a[lev - 1])
    return f(a[lev - 1])


# --------------------------------------------------------------------------- setup protocol


class SymbolicSetup:
    def __init__(self, solver):
        self.solver = solver


class NumericalSetup:
    """Owns the device-side NumericalSetup handle (and keeps its children alive)."""

    def __init__(self, solver, handle, children=(), mat=None):
        self.solver, self.h, self.children, self.mat = solver, handle, list(children), mat

    def _logs(self):
        if hasattr(self.solver, "log") and self.solver.log is not None:
            self.solver.log._fill(self.h)
        for c in self.children:
            if c is not None:
                c._logs()

    def __del__(self):
        try:
            if self.h:
                _lib.lib().gsb_solver_destroy(self.h)
        except Exception:
            pass


def symbolic_setup(solver, A, x=None) -> SymbolicSetup:
    return SymbolicSetup(solver)


def numerical_setup(ss: SymbolicSetup, A, x=None) -> NumericalSetup:
    return ss.solver._numerical_setup(A)


def numerical_setup_(ns: NumericalSetup, A, x=None) -> NumericalSetup:
    """numerical_setup!(ns,A[,x]): refresh value-dependent data after A's values changed."""
    check(_lib.lib().gsb_solver_update(ns.h, A.h))
    ns.mat = A
    return ns


def solve_(x, ns: NumericalSetup, b, zero_initial_guess: bool = False):
    """solve!(x,ns,b).  x, b: device `Vector`s, or host numpy arrays of own values (the e2e path:
    H2D of b and x, solve, D2H of x inside one C call; zero_initial_guess skips the upload of x)."""
    L = _lib.lib()
    if isinstance(x, np.ndarray):
        assert x.dtype == np.float64 and b.dtype == np.float64 and x.flags.c_contiguous and b.flags.c_contiguous
        fn = L.gsb_solve_host_zero_guess if zero_initial_guess else L.gsb_solve_host
        check(fn(ns.h, _ptr(x), _ptr(b), x.shape[0]))
    else:
        check(L.gsb_solve(ns.h, x.h, b.h))
    ns._logs()
    if getattr(ns, "_after_solve", None) is not None:
        ns._after_solve(ns)
    return x


def ldiv_(x, ns, b):
    return solve_(x, ns, b)


class AffineOperator:
    """Gridap.Algebra.AffineOperator: the pair (matrix, vector) a linear FE operator hands to the solver."""

    def __init__(self, matrix, vector):
        self.matrix, self.vector = matrix, vector


def solve_affine_(x, ls, op: AffineOperator, cache=None, newmatrix: bool = False, plan=None):
    """solve!(x::PVector, ls::LinearSolver, op::AffineOperator, cache[, newmatrix]) -- SolverInterfaces/GridapExtras.jl:33-58.
    `x` may carry another ghost layout than the columns of op.matrix (an FE-space vector): the solve runs on a vector `y`
    allocated in the domain of the matrix, own values are copied in and out, and the ghosts of `x` are made consistent
    afterwards (`plan`: the exchange plan of x's layout; None on one part).  First call (cache None): symbolic +
    numerical set-up, returns the cache (ns, y); later calls reuse it and refresh the set-up when newmatrix."""
    A, b = op.matrix, op.vector
    if cache is None:
        ns = numerical_setup(symbolic_setup(ls, A), A)
        y = allocate_in_domain(A)
        cache = (ns, y)
    else:
        ns, y = cache
        if newmatrix:
            numerical_setup_(ns, A)
    copy_(y, x)
    solve_(y, ns, b)
    copy_(x, y)
    if plan is not None:
        consistent_(x, plan)
    return cache


def explicit_transfer(mul, n_in: int, n_out: int, colour, ncolours: int):
    """Materialise a linear transfer operator given only as `mul(y, x)` (y = op x on own values) as a scipy CSR matrix by
    coloured probing -- the algorithm of `explicit_transfer` in julia/GridapSolversB200.jl (the reference's GMG accepts
    any object with mul! as interp/restrict, GMGLinearSolvers.jl:484,491; setup_transfer_operators returns such
    objects, GridTransferOperators.jl:350-401).  `colour(j)` in 0..ncolours-1 must differ for columns whose images
    overlap (Q1, factor-2 refinement: 2^d colours on the coarse node grid; 3^d for the restriction).  Two probes per
    colour: x = 1 on the colour's columns gives the values, x = column id gives the column of every non-zero row."""
    import scipy.sparse as sp

    cols_of = [[] for _ in range(ncolours)]
    for j in range(n_in):
        cols_of[colour(j)].append(j)
    I, J, V = [], [], []
    x, y, y2 = np.zeros(n_in), np.zeros(n_out), np.zeros(n_out)
    for cols in cols_of:
        if not cols:
            continue
        cols = np.asarray(cols)
        x[:] = 0.0
        x[cols] = 1.0
        mul(y, x)
        x[cols] = cols + 1.0  # 1-based ids: column 0 stays distinguishable from "no contribution"
        mul(y2, x)
        rows = np.flatnonzero(y)
        q = y2[rows] / y[rows]
        if np.any(np.abs(q - np.rint(q)) > 1e-6):
            raise ValueError("explicit_transfer: two columns of one colour overlap (invalid colouring)")
        I.append(rows)
        J.append(np.rint(q).astype(np.int64) - 1)
        V.append(y[rows].copy())
    if not I:
        return sp.csr_matrix((n_out, n_in))
    return sp.csr_matrix((np.concatenate(V), (np.concatenate(I), np.concatenate(J))), shape=(n_out, n_in))


def _child(solver, A):
    return None if solver is None else numerical_setup(symbolic_setup(solver, A), A)


def _h(ns):
    return None if ns is None else ns.h


# --------------------------------------------------------------------------- solvers


class LinearSolver:
    log = None


class IdentitySolver(LinearSolver):
    def _numerical_setup(self, A):
        h = c_p()
        check(_lib.lib().gsb_identity_create(A.ctx.h, ctypes.byref(h)))
        return NumericalSetup(self, h)


class JacobiLinearSolver(LinearSolver):
    def _numerical_setup(self, A):
        h = c_p()
        check(_lib.lib().gsb_jacobi_create(A.h, ctypes.byref(h)))
        return NumericalSetup(self, h, mat=A)


class LUSolver(LinearSolver):
    """Gridap.Algebra.LUSolver stand-in: dense fp64 inverse on the device (coarsest GMG level)."""

    def _numerical_setup(self, A):
        h = c_p()
        check(_lib.lib().gsb_dense_lu_create(A.h, ctypes.byref(h)))
        return NumericalSetup(self, h, mat=A)


class RichardsonSmoother(LinearSolver):
    def __init__(self, M, niter: int = 1, omega: float = 1.0):
        self.M, self.niter, self.omega = M, int(niter), float(omega)

    def _numerical_setup(self, A):
        Mns = _child(self.M, A)
        h = c_p()
        check(_lib.lib().gsb_richardson_create(A.h, Mns.h, self.niter, self.omega, ctypes.byref(h)))
        return NumericalSetup(self, h, [Mns], mat=A)


class LinearSolverFromSmoother(LinearSolver):
    def __init__(self, smoother):
        self.smoother = smoother

    def _numerical_setup(self, A):
        sns = _child(self.smoother, A)
        h = c_p()
        check(_lib.lib().gsb_from_smoother_create(A.h, sns.h, ctypes.byref(h)))
        return NumericalSetup(self, h, [sns], mat=A)


def Fill(value, n):
    """FillArrays.Fill(value,n): n references to the SAME solver object (GMGLinearSolvers.jl:52)."""
    return [value] * n


class GMGLinearSolver(LinearSolver):
    """GMGLinearSolver(matrices, prolongations, restrictions; ...) -- GMGLinearSolvers.jl:48-69.
    Transfer operators are explicit sparse matrices (P, and R = P^T for mode=:residual)."""

    def __init__(self, smatrices, interp, restrict, pre_smoothers=None, post_smoothers=None, coarsest_solver=None,
                 mode="preconditioner", cycle_type="v_cycle", maxiter=100, atol=1.0e-14, rtol=1.0e-08, verbose=False,
                 redist=None):
        """redist = (to_coarse, to_fine): per level boundary a RedistributionPlan pair (or None) for hierarchies whose
        coarse levels live on fewer parts (the transfer operators then carry redist = Val{true} in the reference,
        GridTransferOperators.jl:391-401,536-561)"""
        n = len(smatrices)
        self.redist = redist
        if pre_smoothers is None:
            pre_smoothers = Fill(RichardsonSmoother(JacobiLinearSolver(), 10), n - 1)
        if post_smoothers is None:
            post_smoothers = pre_smoothers
        assert n - 1 == len(interp) == len(restrict) == len(pre_smoothers) == len(post_smoothers)  # @check :59
        same_smoothers = pre_smoothers is post_smoothers
        assert mode in _MODES and cycle_type in _CYCLES  # @check :60-61
        self.smatrices, self.interp, self.restrict = list(smatrices), list(interp), list(restrict)
        self.pre_smoothers = list(pre_smoothers)
        self.post_smoothers = self.pre_smoothers if same_smoothers else list(post_smoothers)
        self.coarsest_solver = coarsest_solver if coarsest_solver is not None else LUSolver()
        self.mode, self.cycle_type = mode, cycle_type
        self.log = ConvergenceLog("GMG", SolverTolerances(maxiter=maxiter, atol=atol, rtol=rtol), verbose=verbose)

    def _numerical_setup(self, mat):  # GMGLinearSolvers.jl:183-210
        self.smatrices[0] = mat  # :336-340
        sm = self.smatrices
        n = len(sm)
        pre = [_child(s, A) for s, A in zip(self.pre_smoothers, sm[: n - 1])]
        post = pre if self.pre_smoothers is self.post_smoothers else [_child(s, A) for s, A in zip(self.post_smoothers, sm[: n - 1])]
        coarse = _child(self.coarsest_solver, sm[n - 1])
        arr = lambda objs: (c_p * max(1, len(objs)))(*[o.h for o in objs])
        h = c_p()
        if self.redist is None:
            check(_lib.lib().gsb_gmg_create(mat.ctx.h, n, arr(sm), arr(self.interp), arr(self.restrict), arr(pre), arr(post),
                                            coarse.h, _MODES[self.mode], _CYCLES[self.cycle_type], self.log.tols.maxiter,
                                            self.log.tols.atol, self.log.tols.rtol, ctypes.byref(h)))
        else:
            parr = lambda plans: (c_p * max(1, len(plans)))(*[(p.h if p is not None else None) for p in plans])
            tc, tf = self.redist
            assert len(tc) == len(tf) == n - 1
            check(_lib.lib().gsb_gmg_create_redist(mat.ctx.h, n, arr(sm), arr(self.interp), arr(self.restrict), arr(pre), arr(post),
                                                   coarse.h, _MODES[self.mode], _CYCLES[self.cycle_type], self.log.tols.maxiter,
                                                   self.log.tols.atol, self.log.tols.rtol, parr(tc), parr(tf), ctypes.byref(h)))
        kids = pre + ([] if post is pre else post) + [coarse]
        return NumericalSetup(self, h, kids, mat=mat)


class LanczosDiagnostic:
    """Krylov/KrylovUtils.jl:58-90: (delta, gamma) of the Lanczos tridiagonal recorded from CG's alpha/beta
    (CGSolvers.jl:122-138); `estimate()` = condition-number estimate |lambda_max / lambda_min|."""

    def __init__(self, max_iters: int):
        self.k = 0
        self.delta = np.zeros(max_iters)
        self.gamma = np.zeros(max_iters)

    def reset(self):
        self.k = 0
        self.delta[:] = 0
        self.gamma[:] = 0

    def _update_from(self, alpha, beta):
        self.reset()
        a_last = 1.0
        for a, b in zip(alpha, beta):
            if self.k == 0:
                d, g = 1.0 / a, 0.0
            else:
                d, g = (1.0 / a) + (b / a_last), math.sqrt(b) / a
            self.delta[self.k], self.gamma[self.k] = d, g
            self.k += 1
            a_last = a

    def estimate(self) -> float:
        k = self.k
        if k < 2:
            return 1.0
        T = np.diag(self.delta[:k]) + np.diag(self.gamma[1:k], 1) + np.diag(self.gamma[1:k], -1)
        lam = np.linalg.eigvalsh(T)
        return float(abs(lam.max() / lam.min()))


class CGSolver(LinearSolver):
    def __init__(self, Pl=None, maxiter=1000, atol=1e-12, rtol=1.0e-6, diagnostic=None, flexible=False, verbose=0, name="CG"):
        self.Pl, self.flexible, self.diag = Pl, bool(flexible), diagnostic
        self.log = ConvergenceLog(name, SolverTolerances(maxiter=maxiter, atol=atol, rtol=rtol), verbose=verbose)

    def _numerical_setup(self, A):
        Pl = _child(self.Pl, A)
        t = self.log.tols
        h = c_p()
        check(_lib.lib().gsb_cg_create(A.h, _h(Pl), int(self.flexible), t.maxiter, t.atol, t.rtol, ctypes.byref(h)))
        ns = NumericalSetup(self, h, [Pl], mat=A)
        if self.diag is not None:
            check(_lib.lib().gsb_cg_record_coefficients(h, 1))
            ns._after_solve = self._fill_diagnostic
        return ns

    def _fill_diagnostic(self, ns):
        cap = self.log.tols.maxiter
        alpha, beta, n = np.zeros(cap), np.zeros(cap), ctypes.c_int64()
        check(_lib.lib().gsb_cg_coefficients(ns.h, _ptr(alpha), _ptr(beta), cap, ctypes.byref(n)))
        self.diag._update_from(alpha[: n.value], beta[: n.value])


class RichardsonLinearSolver(LinearSolver):
    """RichardsonLinearSolver(omega, maxiter; Pl, rtol, atol) -- RichardsonLinearSolvers.jl:12-23 (scalar omega)"""

    def __init__(self, omega, maxiter, Pl=None, rtol=1e-10, atol=1e-6, verbose=False, name="RichardsonLinearSolver"):
        self.omega, self.Pl = float(omega), Pl
        self.log = ConvergenceLog(name, SolverTolerances(maxiter=maxiter, atol=atol, rtol=rtol), verbose=verbose)

    def _numerical_setup(self, A):
        Pl = _child(self.Pl, A)
        t = self.log.tols
        h = c_p()
        check(_lib.lib().gsb_richardson_linear_create(A.h, _h(Pl), self.omega, t.maxiter, t.atol, t.rtol, ctypes.byref(h)))
        return NumericalSetup(self, h, [Pl], mat=A)


class SchurComplementSolver(LinearSolver):
    """SchurComplementSolver(A_ns, B, C, S_ns) -- SchurComplementSolvers.jl:8-24; A_ns, S_ns are NumericalSetups"""

    def __init__(self, A_ns, B, C, S_ns):
        self.A, self.B, self.C, self.S = A_ns, B, C, S_ns

    def _numerical_setup(self, mat):
        h = c_p()
        check(_lib.lib().gsb_schur_complement_create(self.B.ctx.h, self.A.h, self.B.h, self.C.h, self.S.h, ctypes.byref(h)))
        return NumericalSetup(self, h, [self.A, self.S], mat=mat)


class GMRESSolver(LinearSolver):
    _create = "gsb_gmres_create"

    def __init__(self, m, Pr=None, Pl=None, restart=False, m_add=1, maxiter=100, atol=1e-12, rtol=1.0e-6,
                 verbose=False, name="GMRES"):
        self.m, self.restart, self.m_add, self.Pr, self.Pl = int(m), bool(restart), int(m_add), Pr, Pl
        self.log = ConvergenceLog(name, SolverTolerances(maxiter=maxiter, atol=atol, rtol=rtol), verbose=verbose)

    def _numerical_setup(self, A):
        Pr, Pl = _child(self.Pr, A), _child(self.Pl, A)
        t = self.log.tols
        h = c_p()
        check(getattr(_lib.lib(), self._create)(A.h, _h(Pr), _h(Pl), self.m, int(self.restart), self.m_add, t.maxiter,
                                                t.atol, t.rtol, ctypes.byref(h)))
        return NumericalSetup(self, h, [Pr, Pl], mat=A)


class FGMRESSolver(GMRESSolver):
    _create = "gsb_fgmres_create"

    def __init__(self, m, Pr, Pl=None, restart=False, m_add=1, maxiter=100, atol=1e-12, rtol=1.0e-6,
                 verbose=False, name="FGMRES"):
        super().__init__(m, Pr=Pr, Pl=Pl, restart=restart, m_add=m_add, maxiter=maxiter, atol=atol, rtol=rtol,
                         verbose=verbose, name=name)


class MINRESSolver(LinearSolver):
    def __init__(self, Pl=None, maxiter=1000, atol=1e-12, rtol=1.0e-6, verbose=False, name="MINRES"):
        self.Pl = Pl
        self.log = ConvergenceLog(name, SolverTolerances(maxiter=maxiter, atol=atol, rtol=rtol), verbose=verbose)

    def _numerical_setup(self, A):
        Pl = _child(self.Pl, A)
        t = self.log.tols
        h = c_p()
        check(_lib.lib().gsb_minres_create(A.h, _h(Pl), t.maxiter, t.atol, t.rtol, ctypes.byref(h)))
        return NumericalSetup(self, h, [Pl], mat=A)


class _BlockSolver(LinearSolver):
    diagonal = False

    def __init__(self, solvers, coeffs=None, half="upper", diag_mats=None):
        self.solvers, self.half, self.diag_mats = list(solvers), half, diag_mats
        nb = len(solvers)
        self.coeffs = np.ones((nb, nb)) if coeffs is None else np.ascontiguousarray(coeffs, dtype=np.float64)

    def _numerical_setup(self, A):
        """A: BlockSparseMatrix.  Diagonal solvers are set up on A[i][i] unless diag_mats[i] overrides
        it (MatrixBlock / BiformBlock semantics, BlockSolverInterfaces.jl:162,262)."""
        nb = len(self.solvers)
        dm = [(self.diag_mats[i] if (self.diag_mats and self.diag_mats[i] is not None) else A[i][i]) for i in range(nb)]
        kids = [_child(self.solvers[i], dm[i]) for i in range(nb)]
        blocks = (c_p * (nb * nb))(*[(A[i][j].h if A[i][j] is not None else None) for i in range(nb) for j in range(nb)])
        sol = (c_p * nb)(*[k.h for k in kids])
        h = c_p()
        check(_lib.lib().gsb_block_solver_create(A.ctx.h, nb, blocks, sol, _ptr(self.coeffs), 1 if self.half == "lower" else 0,
                                                 int(self.diagonal), ctypes.byref(h)))
        return NumericalSetup(self, h, kids, mat=A)


class BlockTriangularSolver(_BlockSolver):
    pass


class BlockDiagonalSolver(_BlockSolver):
    diagonal = True

    def __init__(self, solvers, diag_mats=None):
        super().__init__(solvers, None, "upper", diag_mats)
