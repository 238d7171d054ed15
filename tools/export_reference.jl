# tools/export_reference.jl -- dump what pins this repository's oracle and CUDA path to the REAL GridapSolvers.jl.
#
# NOT EXECUTED IN THIS REPOSITORY'S CONTAINER (no `julia`).  Run it once where GridapSolvers.jl is installed:
#
#     julia --project=<env with GridapSolvers, GridapDistributed, PartitionedArrays, NPZ> tools/export_reference.jl tests/golden
#
# and commit the tests/golden/reference_*.npz files it writes.  tests/test_reference_pinning.py picks every such file
# up (it skips loudly while there is none), feeds the reference's OWN assembled matrices, transfer matrices and
# right-hand side to oracle/ and to libgsb200.so, and asserts
#   * iteration counts equal (+-1) and relative residual histories within 1e-10 of `solver.log.residuals`,
#   * the exported prolongations / restrictions satisfy what the device path assumes (R == P' for mode=:residual),
#   * for the distributed cases: the own/ghost index maps of every part (own-first renumbering, ghost owners) are the
#     ones gridapsolvers.jl_b200/julia/GridapSolversB200.jl builds from the same PRange.
# Cases (SURVEY.md section 8c, BASELINE.json configs): C1 (CG + Jacobi, 2D Poisson Q1), C2-small (CG + GMG V-cycle,
# 3D Poisson Q1, Jacobi-Richardson smoothers, LU coarse solve; the driver of test/LinearSolvers/GMGTests.jl:107-145),
# C2-small on 2x2 parts (DebugArray), C5-small (GMRES + BlockTriangularSolver on 2D Stokes, test/Applications/Stokes.jl).
#
# File layout (arrays are 1-based exactly as Julia holds them; the Python side converts):
#   kind            "cg_jacobi" | "gmg_pcg"
#   nlev, nparts
#   A<l>_colptr / A<l>_rowval / A<l>_nzval / A<l>_shape         level matrices (serial assembled, CSC)
#   P<l>_* , R<l>_*                                             explicit transfer matrices (columns probed with unit vectors)
#   b, x, residuals, num_iters, rtol, atol, maxiter, niter_smooth, omega
#   part<p>_own_to_global_<l>, part<p>_ghost_to_global_<l>, part<p>_ghost_owner_<l>     (distributed cases)
using LinearAlgebra, SparseArrays, FillArrays
using Gridap, Gridap.Algebra, Gridap.ReferenceFEs, Gridap.Geometry
using PartitionedArrays, GridapDistributed
using GridapSolvers, GridapSolvers.LinearSolvers, GridapSolvers.MultilevelTools
using NPZ

outdir = length(ARGS) >= 1 ? ARGS[1] : "tests/golden"
mkpath(outdir)

"serial CSC copy of a (P)SparseMatrix in GLOBAL numbering"
function global_csc(A::PSparseMatrix)
  I, J, V = Int[], Int[], Float64[]
  map(partition(A), partition(axes(A,1)), partition(axes(A,2))) do Al, rows, cols
    r2g, c2g = local_to_global(rows), local_to_global(cols)
    own = own_to_local(rows)
    i, j, v = findnz(Al[own, :])
    append!(I, r2g[own[i]]); append!(J, c2g[j]); append!(V, v)
  end
  sparse(I, J, V, size(A,1), size(A,2))
end
global_csc(A::SparseMatrixCSC) = A

function put_csc!(d, name, A::SparseMatrixCSC)
  d[name*"_colptr"] = Int64.(A.colptr); d[name*"_rowval"] = Int64.(A.rowval)
  d[name*"_nzval"] = Float64.(A.nzval); d[name*"_shape"] = Int64[size(A,1), size(A,2)]
end

"global (own values gathered) copy of a PVector"
function global_vec(v::PVector)
  out = zeros(length(v))
  map(own_values(v), partition(axes(v,1))) do vo, ids
    out[own_to_global(ids)] .= vo
  end
  out
end

"explicit matrix of a transfer operator: one mul! per column (small problems only)"
function probe(op, x::PVector, y::PVector)
  n_in, n_out = length(x), length(y)
  I, J, V = Int[], Int[], Float64[]
  for j in 1:n_in
    fill!(x, 0.0)
    map(own_values(x), partition(axes(x,1))) do xo, ids
      k = findfirst(==(j), own_to_global(ids)); isnothing(k) || (xo[k] = 1.0)
    end
    consistent!(x) |> wait
    mul!(y, op, x)
    col = global_vec(y)
    for i in findall(!iszero, col)
      push!(I, i); push!(J, j); push!(V, col[i])
    end
  end
  sparse(I, J, V, n_out, n_in)
end

function export_gmg_poisson(name, np, ncells, nlev; niter=10, ω=2.0/3.0, rtol=1e-8, atol=1e-14, maxiter=30)
  D = length(ncells)
  with_debug() do distribute
    parts = distribute(LinearIndices((prod(np),)))
    domain = Tuple(vcat([[0.0, 1.0] for _ in 1:D]...))
    mh = CartesianModelHierarchy(parts, fill(np, nlev), domain, ncells .÷ 2^(nlev-1))   # coarsest mesh refined nlev-1 times
    u(x) = x[1] + x[2]
    f(x) = -Δ(u)(x)
    biform(u,v,dΩ) = ∫(∇(v)⋅∇(u))dΩ
    liform(v,dΩ)   = ∫(v*f)dΩ
    qdegree = 3
    reffe   = ReferenceFE(lagrangian, Float64, 1)
    tests   = TestFESpace(mh, reffe, dirichlet_tags="boundary")
    trials  = TrialFESpace(tests, u)
    restrictions, prolongations = setup_transfer_operators(tests, qdegree; mode=:residual, solver=CGSolver(JacobiLinearSolver(); rtol=1e-14))
    smoothers = HierarchicalArray(Fill(RichardsonSmoother(JacobiLinearSolver(), niter, ω), nlev-1), view(get_level_parts(mh), 1:nlev-1))
    smatrices, A, b = compute_hierarchy_matrices(trials, tests, biform, liform, qdegree)
    gmg = GMGLinearSolver(smatrices, prolongations, restrictions; pre_smoothers=smoothers, post_smoothers=smoothers,
                          coarsest_solver=LUSolver(), maxiter=1, mode=:preconditioner, cycle_type=:v_cycle)
    solver = CGSolver(gmg; maxiter=maxiter, atol=atol, rtol=rtol)
    ns = numerical_setup(symbolic_setup(solver, A), A)
    x = pfill(0.0, partition(axes(A,2)))
    solve!(x, ns, b)
    d = Dict{String,Any}("kind" => "gmg_pcg", "nlev" => nlev, "nparts" => prod(np), "rtol" => rtol, "atol" => atol,
                         "maxiter" => maxiter, "niter_smooth" => niter, "omega" => ω, "ncells" => Int64.(collect(ncells)))
    for l in 1:nlev
      Al = smatrices[l]
      put_csc!(d, "A$l", global_csc(Al))
      map(partition(axes(Al,2))) do ids   # index maps of every part (1 entry per part under DebugArray)
        nothing
      end
      for (p, ids) in enumerate(partition(axes(Al,2)).items)
        d["part$(p)_own_to_global_$l"]   = Int64.(own_to_global(ids))
        d["part$(p)_ghost_to_global_$l"] = Int64.(ghost_to_global(ids))
        d["part$(p)_ghost_owner_$l"]     = Int64.(ghost_to_owner(ids))
      end
      if l < nlev
        xH = pfill(0.0, partition(axes(smatrices[l+1],2))); xh = pfill(0.0, partition(axes(Al,2)))
        put_csc!(d, "P$l", probe(prolongations[l], xH, xh))
        put_csc!(d, "R$l", probe(restrictions[l], xh, xH))
      end
    end
    d["b"] = global_vec(b); d["x"] = global_vec(x)
    d["num_iters"] = solver.log.num_iters
    d["residuals"] = Float64.(solver.log.residuals[1:solver.log.num_iters+1])
    npzwrite(joinpath(outdir, "reference_$(name).npz"), d)
    println("wrote reference_$(name).npz : ", solver.log.num_iters, " iterations")
  end
end

function export_cg_jacobi(name, ncells; rtol=1e-8, atol=1e-14, maxiter=1000)
  model = CartesianDiscreteModel((0,1,0,1), ncells)
  u(x) = x[1] + x[2]
  f(x) = -Δ(u)(x)
  V = TestFESpace(model, ReferenceFE(lagrangian, Float64, 1), dirichlet_tags="boundary")
  U = TrialFESpace(V, u)
  Ω = Triangulation(model); dΩ = Measure(Ω, 3)
  op = AffineFEOperator((u,v) -> ∫(∇(v)⋅∇(u))dΩ, v -> ∫(v*f)dΩ, U, V)
  A, b = get_matrix(op), get_vector(op)
  solver = CGSolver(JacobiLinearSolver(); maxiter=maxiter, atol=atol, rtol=rtol)
  ns = numerical_setup(symbolic_setup(solver, A), A)
  x = zeros(size(A,2))
  solve!(x, ns, b)
  d = Dict{String,Any}("kind" => "cg_jacobi", "nlev" => 1, "nparts" => 1, "rtol" => rtol, "atol" => atol, "maxiter" => maxiter,
                       "ncells" => Int64.(collect(ncells)), "b" => b, "x" => x, "num_iters" => solver.log.num_iters,
                       "residuals" => Float64.(solver.log.residuals[1:solver.log.num_iters+1]))
  put_csc!(d, "A1", A)
  npzwrite(joinpath(outdir, "reference_$(name).npz"), d)
  println("wrote reference_$(name).npz : ", solver.log.num_iters, " iterations")
end

export_cg_jacobi("c1_cg_jacobi_poisson2d_64", (64, 64))
export_gmg_poisson("c2_gmg_pcg_poisson3d_16_serial", (1,1,1), (16,16,16), 3)
export_gmg_poisson("c2_gmg_pcg_poisson3d_16_parts2x2x1", (2,2,1), (16,16,16), 3)
