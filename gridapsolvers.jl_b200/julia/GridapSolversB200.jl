# GridapSolversB200.jl -- Julia shim over libgsb200.so (include/gsb200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build container has no `julia`.  It is written against the
# same C ABI the Python mirror (gridapsolvers.jl_b200/api.py) exercises on every GPU test, and against
# the reference's interfaces as cited.  PartitionedArrays accessor names follow PartitionedArrays 0.3.x
# as used by the reference at src/SolverInterfaces/PAExtras.jl:9-168.
#
# Usage (drop-in):
#     using GridapSolvers, GridapSolversB200
#     gmg    = GMGLinearSolver(mats, Ps, Rs; pre_smoothers=sm, post_smoothers=sm, maxiter=1)   # explicit sparse P / R
#     solver = B200Solver(CGSolver(gmg; rtol=1e-8); ctx=B200Context(comm))
#     ns = numerical_setup(symbolic_setup(solver, A), A)
#     solve!(x, ns, b)            # x, b :: PVector (or Vector); solver.solver.log is filled as usual
module GridapSolversB200

using LinearAlgebra, SparseArrays
using Gridap, Gridap.Algebra
using PartitionedArrays
using GridapSolvers
using GridapSolvers.LinearSolvers
using GridapSolvers.SolverInterfaces
using GridapSolvers.BlockSolvers
using BlockArrays
import MPI

const libgsb = get(ENV, "GSB200_LIB", "libgsb200.so")

check(code::Cint) = code == 0 ? nothing :
  error("libgsb200: ", unsafe_string(ccall((:gsb_last_error, libgsb), Cstring, (Ptr{Cvoid},), C_NULL)))

# ------------------------------------------------------------------ context (one part == one rank == one GPU)
mutable struct B200Context
  h::Ptr{Cvoid}
  comm
end

function B200Context(comm=MPI.COMM_SELF; device=nothing)
  rank, nranks = MPI.Comm_rank(comm), MPI.Comm_size(comm)
  id = zeros(UInt8, 128)
  if nranks > 1
    rank == 0 && check(ccall((:gsb_nccl_unique_id, libgsb), Cint, (Ptr{UInt8},), id))
    MPI.Bcast!(id, 0, comm)
  end
  dev = isnothing(device) ? rank % max(1, parse(Int, get(ENV, "GSB200_GPUS_PER_NODE", "8"))) : device
  h = Ref{Ptr{Cvoid}}()
  check(ccall((:gsb_init, libgsb), Cint, (Cint, Cint, Cint, Ptr{UInt8}, Ref{Ptr{Cvoid}}), dev, nranks, rank, id, h))
  ctx = B200Context(h[], comm)
  finalizer(c -> ccall((:gsb_finalize, libgsb), Cint, (Ptr{Cvoid},), c.h), ctx)  # cf. joss_paper/scalability/src/stokes_gmg.jl:73-81
  return ctx
end

# ------------------------------------------------------------------ PSparseMatrix / PVector mirrors
mutable struct B200Matrix
  h::Ptr{Cvoid}
  plan::Ptr{Cvoid}
  own_to_local::Vector{Int}      # to move values between Julia's local numbering and own-first numbering
  ghost_to_local::Vector{Int}
  ctx::B200Context
  own_rows::Vector{Int}          # local ids of the own rows (PSparseMatrix parts; empty for serial matrices)
  l2new::Vector{Int}             # Julia local column numbering -> own-first numbering
end
B200Matrix(h, plan, o2l, g2l, ctx) = B200Matrix(h, plan, o2l, g2l, ctx, Int[], Int[])

"values of one part in the order the device matrix was created with (same sparsity => same order)"
function _own_first_values(Al::SparseMatrixCSC, own_rows::Vector{Int}, l2new::Vector{Int}, n_cols::Int)
  I, J, V = findnz(Al[own_rows, :])
  return sparse(I, l2new[J], V, length(own_rows), n_cols).nzval
end

"Serial matrix: SparseMatrixCSC{Float64,Int64}, 1-based, passed untouched (fmt = CSC)."
function B200Matrix(ctx::B200Context, A::SparseMatrixCSC{Float64,Int64})
  h = Ref{Ptr{Cvoid}}()
  check(ccall((:gsb_mat_create, libgsb), Cint,
    (Ptr{Cvoid}, Int64, Int64, Int64, Cint, Cint, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
    ctx.h, size(A,1), size(A,2), 0, 1, 1, 8, A.colptr, A.rowval, A.nzval, C_NULL, h))
  return B200Matrix(h[], C_NULL, collect(1:size(A,2)), Int[], ctx)
end

"""
One part of a PSparseMatrix.  Rows = own rows; columns are renumbered own-first (own columns in
own order, then ghost columns in ghost order) so that the device accumulation order equals
PartitionedArrays' `c_own = A_oo*b_own; c_own += A_oh*b_ghost` (SURVEY.md App. B).  The exchange plan
is the column partition's assembly cache reversed for `consistent!` (neighbours + local id lists),
read exactly like src/SolverInterfaces/PAExtras.jl:15-60 reads it.
"""
function B200Matrix(ctx::B200Context, A::PSparseMatrix)
  mats  = partition(A)
  rows  = partition(axes(A,1))
  cols  = partition(axes(A,2))
  # consistent! direction (owner -> ghost) = the assembly caches with snd/rcv swapped, exactly as the reference
  # reads them at src/SolverInterfaces/PAExtras.jl:84-86 ("Reversed caches")
  nbors_rcv, nbors_snd = assembly_neighbors(cols)
  lids_rcv, lids_snd   = assembly_local_indices(cols, nbors_rcv, nbors_snd)
  map(mats, rows, cols, nbors_snd, nbors_rcv, lids_snd, lids_rcv) do Al, ri, ci, nb_snd, nb_rcv, l_snd, l_rcv
    o2l, g2l = own_to_local(ci), ghost_to_local(ci)
    n_own, n_ghost = length(o2l), length(g2l)
    l2new = zeros(Int, n_own + n_ghost)           # Julia local numbering -> own-first numbering
    l2new[o2l] .= 1:n_own
    l2new[g2l] .= n_own .+ (1:n_ghost)
    Aoo = Al[own_to_local(ri), :]                 # own rows only
    I, J, V = findnz(Aoo)
    B = sparse(I, l2new[J], V, length(own_to_local(ri)), n_own + n_ghost)   # CSC, own-first columns
    snd_ptrs, rcv_ptrs = Int64.(l_snd.ptrs), Int64.(l_rcv.ptrs)
    snd_ids = Int64.(l2new[l_snd.data])           # own entries to pack, per send neighbour
    rcv_ids = Int64.(l2new[l_rcv.data])           # ghost entries to fill, per receive neighbour
    plan = Ref{Ptr{Cvoid}}()
    check(ccall((:gsb_plan_create, libgsb), Cint,
      (Ptr{Cvoid}, Int64, Int64, Cint, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Cint, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Cint, Ref{Ptr{Cvoid}}),
      ctx.h, n_own, n_ghost, length(nb_snd), Int32.(nb_snd .- 1), snd_ptrs, snd_ids,
      length(nb_rcv), Int32.(nb_rcv .- 1), rcv_ptrs, rcv_ids, 1, plan))
    h = Ref{Ptr{Cvoid}}()
    check(ccall((:gsb_mat_create, libgsb), Cint,
      (Ptr{Cvoid}, Int64, Int64, Int64, Cint, Cint, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
      ctx.h, size(B,1), n_own, n_ghost, 1, 1, 8, B.colptr, B.rowval, B.nzval, plan[], h))
    B200Matrix(h[], plan[], collect(o2l), collect(g2l), ctx, collect(own_to_local(ri)), l2new)
  end |> PartitionedArrays.getany   # one part per process under with_mpi
end

mutable struct B200Vector
  h::Ptr{Cvoid}
  n_own::Int
end
function allocate_like_domain(A::B200Matrix)
  h = Ref{Ptr{Cvoid}}(); check(ccall((:gsb_vec_create_domain, libgsb), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), A.h, h))
  n = Ref{Int64}(); check(ccall((:gsb_vec_size, libgsb), Cint, (Ptr{Cvoid}, Ref{Int64}, Ptr{Int64}), h[], n, C_NULL))
  B200Vector(h[], n[])
end
set_own!(v::B200Vector, x::AbstractVector{Float64}) = check(ccall((:gsb_vec_set, libgsb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), v.h, x, length(x)))
get_own!(x::AbstractVector{Float64}, v::B200Vector) = check(ccall((:gsb_vec_get, libgsb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), v.h, x, length(x)))

# ------------------------------------------------------------------ solver tree -> NumericalSetup handles
# Each method is numerical_setup(symbolic_setup(s,A),A) of the corresponding reference solver.
_h(x) = isnothing(x) ? C_NULL : x
function _create(f::Symbol, argt, args...)
  h = Ref{Ptr{Cvoid}}()
  check(ccall((f, libgsb), Cint, (argt..., Ref{Ptr{Cvoid}}), args..., h))
  return h[]
end

device_ns(::Nothing, A, reg) = nothing
device_ns(s::LinearSolvers.JacobiLinearSolver, A::B200Matrix, reg) = _create(:gsb_jacobi_create, (Ptr{Cvoid},), A.h)
device_ns(s::LinearSolvers.IdentitySolver, A::B200Matrix, reg) = _create(:gsb_identity_create, (Ptr{Cvoid},), A.ctx.h)
device_ns(s::Gridap.Algebra.LUSolver, A::B200Matrix, reg) = _create(:gsb_dense_lu_create, (Ptr{Cvoid},), A.h)
function device_ns(s::LinearSolvers.RichardsonSmoother, A::B200Matrix, reg)
  M = device_ns(s.M, A, reg)
  _create(:gsb_richardson_create, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cdouble), A.h, M, s.niter, s.ω)
end
function device_ns(s::LinearSolvers.LinearSolverFromSmoother, A::B200Matrix, reg)
  sm = device_ns(s.smoother, A, reg)
  _create(:gsb_from_smoother_create, (Ptr{Cvoid}, Ptr{Cvoid}), A.h, sm)
end
function device_ns(s::LinearSolvers.CGSolver, A::B200Matrix, reg)
  Pl = device_ns(s.Pl, A, reg); t = s.log.tols
  h = _create(:gsb_cg_create, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cdouble, Cdouble), A.h, _h(Pl), s.flexible, t.maxiter, t.atol, t.rtol)
  push!(reg, (s.log, h)); h
end
function device_ns(s::LinearSolvers.GMRESSolver, A::B200Matrix, reg)
  Pr = device_ns(s.Pr, A, reg); Pl = device_ns(s.Pl, A, reg); t = s.log.tols
  h = _create(:gsb_gmres_create, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cdouble, Cdouble),
              A.h, _h(Pr), _h(Pl), s.m, s.restart, s.m_add, t.maxiter, t.atol, t.rtol)
  push!(reg, (s.log, h)); h
end
function device_ns(s::LinearSolvers.FGMRESSolver, A::B200Matrix, reg)
  Pr = device_ns(s.Pr, A, reg); Pl = device_ns(s.Pl, A, reg); t = s.log.tols
  h = _create(:gsb_fgmres_create, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cdouble, Cdouble),
              A.h, _h(Pr), _h(Pl), s.m, s.restart, s.m_add, t.maxiter, t.atol, t.rtol)
  push!(reg, (s.log, h)); h
end
function device_ns(s::LinearSolvers.MINRESSolver, A::B200Matrix, reg)
  Pl = device_ns(s.Pl, A, reg); t = s.log.tols
  h = _create(:gsb_minres_create, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cdouble, Cdouble), A.h, _h(Pl), t.maxiter, t.atol, t.rtol)
  push!(reg, (s.log, h)); h
end
function device_ns(s::LinearSolvers.RichardsonLinearSolver, A::B200Matrix, reg)
  s.ω isa Float64 || error("RichardsonLinearSolver: only a scalar relaxation parameter is mirrored")
  Pl = device_ns(s.Pl, A, reg); t = s.log.tols
  h = _create(:gsb_richardson_linear_create, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Cdouble, Cdouble), A.h, _h(Pl), s.ω, t.maxiter, t.atol, t.rtol)
  push!(reg, (s.log, h)); h
end
"""
BlockTriangularSolver with `LinearSystemBlock`s (src/BlockSolvers/BlockTriangularSolvers.jl:26-58): `blocks` is the
matrix of B200Matrix blocks of the system (nothing = zero block); diagonal solvers act on blocks[i,i].
"""
function device_block_triangular(ctx::B200Context, blocks::Matrix, solvers::Vector, coeffs::Matrix{Float64}, half::Symbol, reg)
  nb = length(solvers)
  hs = Ptr{Cvoid}[device_ns(solvers[i], blocks[i,i], reg) for i in 1:nb]
  hb = Ptr{Cvoid}[isnothing(blocks[i,j]) ? C_NULL : blocks[i,j].h for i in 1:nb for j in 1:nb]   # row-major
  cf = Float64[coeffs[i,j] for i in 1:nb for j in 1:nb]
  _create(:gsb_block_solver_create, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Float64}, Cint, Cint),
          ctx.h, nb, hb, hs, cf, half == :lower ? 1 : 0, 0)
end
"""
GMGLinearSolverFromMatrices (src/LinearSolvers/GMGLinearSolvers.jl:8-18).  `interp[l]` / `restrict[l]`
must be explicit sparse matrices (legal for the reference: GMG only calls mul! on them, :484,491);
`explicit_transfer(op)` below extracts them from a DistributedGridTransferOperator by coloured probing.
"""
function device_ns(s::LinearSolvers.GMGLinearSolverFromMatrices, A::B200Matrix, reg)
  nlev = length(s.smatrices)
  mats = [l == 1 ? A : B200Matrix(A.ctx, s.smatrices[l]) for l in 1:nlev]     # smatrices[1] = A, :336-340
  Ps   = [B200Matrix(A.ctx, s.interp[l])   for l in 1:nlev-1]
  Rs   = [B200Matrix(A.ctx, s.restrict[l]) for l in 1:nlev-1]
  pre  = [device_ns(s.pre_smoothers[l], mats[l], reg) for l in 1:nlev-1]
  post = s.pre_smoothers === s.post_smoothers ? pre : [device_ns(s.post_smoothers[l], mats[l], reg) for l in 1:nlev-1]  # :190-194
  coarse = device_ns(s.coarsest_solver, mats[nlev], reg)
  t = s.log.tols
  mode  = s.mode == :preconditioner ? 0 : 1
  cycle = Dict(:v_cycle => 0, :w_cycle => 1, :f_cycle => 2)[s.cycle_type]
  hp(v) = Ptr{Cvoid}[m isa B200Matrix ? m.h : m for m in v]
  h = _create(:gsb_gmg_create,
    (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Cint, Cint, Cint, Cdouble, Cdouble),
    A.ctx.h, nlev, hp(mats), hp(Ps), hp(Rs), hp(pre), hp(post), coarse, mode, cycle, t.maxiter, t.atol, t.rtol)
  push!(reg, (s.log, h)); push!(reg, (:keepalive, (mats, Ps, Rs)))
  return h
end

# ------------------------------------------------------------------ the drop-in LinearSolver
struct B200Solver{S} <: Gridap.Algebra.LinearSolver
  solver::S            # any reference solver tree made of the types handled above
  ctx::B200Context
end
B200Solver(s; ctx=B200Context()) = B200Solver(s, ctx)

struct B200SymbolicSetup{S} <: Gridap.Algebra.SymbolicSetup
  solver::B200Solver{S}
end
mutable struct B200NumericalSetup{S} <: Gridap.Algebra.NumericalSetup
  solver::B200Solver{S}
  A::B200Matrix
  h::Ptr{Cvoid}
  logs::Vector{Any}     # (ConvergenceLog, handle) pairs to fill after each solve
  xd::B200Vector
  bd::B200Vector
end

Gridap.Algebra.symbolic_setup(s::B200Solver, A::AbstractMatrix) = B200SymbolicSetup(s)
function Gridap.Algebra.numerical_setup(ss::B200SymbolicSetup, A::AbstractMatrix)
  Ad  = B200Matrix(ss.solver.ctx, A)
  reg = Any[]
  h   = device_ns(ss.solver.solver, Ad, reg)
  return B200NumericalSetup(ss.solver, Ad, h, reg, allocate_like_domain(Ad), allocate_like_domain(Ad))
end
"numerical_setup!(ns,A): same sparsity, new values (CGSolvers.jl:57-63)."
function Gridap.Algebra.numerical_setup!(ns::B200NumericalSetup, A::SparseMatrixCSC)
  check(ccall((:gsb_mat_update_values, libgsb), Cint, (Ptr{Cvoid}, Ptr{Float64}), ns.A.h, A.nzval))
  check(ccall((:gsb_solver_update, libgsb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ns.h, ns.A.h))
  return ns
end
"numerical_setup!(ns,A) for a PSparseMatrix with the sparsity ns was built with: the own rows of every part are
re-extracted in the creation order (own-first columns, CSC) and uploaded as values only."
function Gridap.Algebra.numerical_setup!(ns::B200NumericalSetup, A::PSparseMatrix)
  map(partition(A)) do Al
    vals = _own_first_values(Al, ns.A.own_rows, ns.A.l2new, length(ns.A.own_to_local) + length(ns.A.ghost_to_local))
    check(ccall((:gsb_mat_update_values, libgsb), Cint, (Ptr{Cvoid}, Ptr{Float64}), ns.A.h, vals))
  end
  check(ccall((:gsb_solver_update, libgsb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ns.h, ns.A.h))
  return ns
end
Gridap.Algebra.numerical_setup!(ns::B200NumericalSetup, A::AbstractMatrix, x::AbstractVector) = numerical_setup!(ns, A)  # GridapExtras.jl:11-13

function _fill_logs!(ns::B200NumericalSetup)
  for (log, h) in ns.logs
    log === :keepalive && continue
    n, flag = Ref{Cint}(), Ref{Cint}()
    fill!(log.residuals, 0.0)
    check(ccall((:gsb_solver_log, libgsb), Cint, (Ptr{Cvoid}, Ref{Cint}, Ptr{Float64}, Int64, Ref{Cint}),
                h, n, log.residuals, length(log.residuals), flag))
    log.num_iters = n[]
  end
end

"solve!(x,ns,b) for serial vectors: host buffers straight through gsb_solve_host (H2D, solve, D2H)."
function Gridap.Algebra.solve!(x::Vector{Float64}, ns::B200NumericalSetup, b::Vector{Float64})
  check(ccall((:gsb_solve_host, libgsb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), ns.h, x, b, length(x)))
  _fill_logs!(ns)
  return x
end
"solve!(x,ns,b) for PVectors: own values in, own values out; ghosts of x are left to the caller's consistent! (GridapExtras.jl:42)."
function Gridap.Algebra.solve!(x::PVector, ns::B200NumericalSetup, b::PVector)
  map(own_values(x), own_values(b)) do xo, bo
    set_own!(ns.xd, collect(xo)); set_own!(ns.bd, collect(bo))
    check(ccall((:gsb_solve, libgsb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ns.h, ns.xd.h, ns.bd.h))
    tmp = collect(xo); get_own!(tmp, ns.xd); xo .= tmp
  end
  _fill_logs!(ns)
  return x
end
LinearAlgebra.ldiv!(x, ns::B200NumericalSetup, b) = solve!(x, ns, b)

# The AffineOperator entry points of src/SolverInterfaces/GridapExtras.jl:33-58, spelled out for B200Solver (the
# reference's generic methods for ::LinearSolver would dispatch here as well): `x` may carry the FE-space ghost
# layout, the solve runs on a vector with the layout of the matrix' columns, ghosts of x are made consistent.
function Gridap.Algebra.solve!(x::PVector, ls::B200Solver, op::Gridap.Algebra.AffineOperator, cache::Nothing)
  A, b = op.matrix, op.vector
  ns = numerical_setup(symbolic_setup(ls, A), A)
  y = allocate_in_domain(A)
  copy!(y, x)
  solve!(y, ns, b)
  copy!(x, y)
  consistent!(x) |> wait
  return ns, y
end
function Gridap.Algebra.solve!(x::PVector, ls::B200Solver, op::Gridap.Algebra.AffineOperator, cache, newmatrix::Bool)
  A, b = op.matrix, op.vector
  ns, y = cache
  newmatrix && numerical_setup!(ns, A)
  copy!(y, x)
  solve!(y, ns, b)
  copy!(x, y)
  consistent!(x) |> wait
  return cache
end

# ------------------------------------------------------------------ block systems (BlockSolvers/)
"Serial block matrix (BlockArrays.BlockMatrix of SparseMatrixCSC): one device matrix per non-zero block + the block
matrix that acts on concatenated vectors (usable as the A of the Krylov solvers)."
struct B200BlockMatrix
  h::Ptr{Cvoid}
  blocks::Matrix{Union{Nothing,B200Matrix}}
  ctx::B200Context
end
function B200BlockMatrix(ctx::B200Context, A::BlockArrays.AbstractBlockMatrix)
  nb = blocksize(A, 1)
  @assert nb == blocksize(A, 2)
  bl = Matrix{Union{Nothing,B200Matrix}}(nothing, nb, nb)
  for i in 1:nb, j in 1:nb
    Aij = A[Block(i, j)]
    nnz(Aij) > 0 && (bl[i, j] = B200Matrix(ctx, SparseMatrixCSC{Float64,Int64}(Aij)))
  end
  hb = Ptr{Cvoid}[isnothing(bl[i, j]) ? C_NULL : bl[i, j].h for i in 1:nb for j in 1:nb]   # row-major
  h = _create(:gsb_block_mat_create, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}), ctx.h, nb, hb)
  return B200BlockMatrix(h, bl, ctx)
end
"BlockTriangularSolver / BlockDiagonalSolver whose blocks are LinearSystemBlocks (taken from the system matrix;
BlockTriangularSolvers.jl:26-58, BlockDiagonalSolvers.jl:22-60).  Matrix/Biform blocks: build the block with the
reference on the host, upload it with B200Matrix and call device_block_triangular directly."
function device_ns(s::BlockSolvers.BlockTriangularSolver{Val{H}}, A::B200BlockMatrix, reg) where H
  all(b -> b isa BlockSolvers.LinearSystemBlock, s.blocks) || error("B200: only LinearSystemBlock blocks are dispatched automatically")
  device_block_triangular(A.ctx, A.blocks, collect(s.solvers), Matrix{Float64}(s.coeffs), H, reg)
end
function device_ns(s::BlockSolvers.BlockDiagonalSolver, A::B200BlockMatrix, reg)
  all(b -> b isa BlockSolvers.LinearSystemBlock, s.blocks) || error("B200: only LinearSystemBlock blocks are dispatched automatically")
  nb = length(s.solvers)
  hs = Ptr{Cvoid}[device_ns(s.solvers[i], A.blocks[i, i], reg) for i in 1:nb]
  hb = Ptr{Cvoid}[(i == j && !isnothing(A.blocks[i, j])) ? A.blocks[i, j].h : C_NULL for i in 1:nb for j in 1:nb]
  _create(:gsb_block_solver_create, (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Float64}, Cint, Cint),
          A.ctx.h, nb, hb, hs, C_NULL, 0, 1)
end
# Krylov solvers on a block system: the operator is the block matrix, preconditioners see the blocks
for (T, f) in ((:GMRESSolver, :gsb_gmres_create), (:FGMRESSolver, :gsb_fgmres_create))
  @eval function device_ns(s::LinearSolvers.$T, A::B200BlockMatrix, reg)
    Pr = device_ns(s.Pr, A, reg); Pl = device_ns(s.Pl, A, reg); t = s.log.tols
    h = _create($(QuoteNode(f)), (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cdouble, Cdouble),
                A.h, _h(Pr), _h(Pl), s.m, s.restart, s.m_add, t.maxiter, t.atol, t.rtol)
    push!(reg, (s.log, h)); h
  end
end
function Gridap.Algebra.numerical_setup(ss::B200SymbolicSetup, A::BlockArrays.AbstractBlockMatrix)
  Ad  = B200BlockMatrix(ss.solver.ctx, A)
  reg = Any[(:keepalive, Ad)]
  h   = device_ns(ss.solver.solver, Ad, reg)
  n   = size(A, 1)
  mk() = (v = Ref{Ptr{Cvoid}}(); check(ccall((:gsb_vec_create, libgsb), Cint, (Ptr{Cvoid}, Int64, Int64, Ref{Ptr{Cvoid}}), Ad.ctx.h, n, 0, v)); B200Vector(v[], n))
  A11 = first(b for b in Ad.blocks if !isnothing(b))
  return B200NumericalSetup(ss.solver, A11, h, reg, mk(), mk())
end
"solve!(x,ns,b) for serial block vectors: the blocks are stored contiguously (BlockArrays' BlockedVector / mortar of Vectors)"
function Gridap.Algebra.solve!(x::BlockArrays.AbstractBlockVector, ns::B200NumericalSetup, b::BlockArrays.AbstractBlockVector)
  xf, bf = collect(x), collect(b)
  check(ccall((:gsb_solve_host, libgsb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), ns.h, xf, bf, length(xf)))
  copyto!(x, xf)
  _fill_logs!(ns)
  return x
end

"""
Materialise a transfer operator as a sparse matrix by probing `mul!(y,op,x)` with coloured 0/1 vectors
(SURVEY.md section 7, "Drop-in transfer operators"; precedent for assembling an interpolation operator as a
matrix: ext/GridapPETScExt/PETScUtils.jl:59-80).  `colour(j)` must give different colours to columns whose
images overlap (Q1, factor-2 refinement: 2^d colours on the coarse node grid).
"""
function explicit_transfer(op, n_in::Int, n_out::Int, colour::Function, ncolours::Int)
  I, J, V = Int[], Int[], Float64[]
  x, y = zeros(n_in), zeros(n_out)
  for c in 1:ncolours
    cols = findall(j -> colour(j) == c, 1:n_in)
    fill!(x, 0.0); x[cols] .= 1.0
    mul!(y, op, x)
    # attribute each nonzero of y to the unique column of this colour whose support contains the row:
    # done by a second probe with x[cols] .= cols (row value / first probe value = column id)
    fill!(x, 0.0); x[cols] .= Float64.(cols); y2 = similar(y); mul!(y2, op, x)
    for i in findall(!iszero, y)
      push!(I, i); push!(J, round(Int, y2[i] / y[i])); push!(V, y[i])
    end
  end
  return sparse(I, J, V, n_out, n_in)
end

end # module
