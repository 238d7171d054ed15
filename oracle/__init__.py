"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy + a small C library) of the solve phase of gridap/GridapSolvers.jl
v0.7.1: Krylov solvers, GMG V/W/F cycles, Richardson/Jacobi smoothers, transfer operators,
convergence logs.  Every function cites the reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package, and only as the CHECKER / the timed CPU baseline -- never on the
product path.  The product (`gridapsolvers.jl_b200/`) has no CPU fallback and never imports it.

PARITY STATUS ("how the oracle is pinned"):
  * The reference is pure Julia; there is no `julia` in this container and no stored golden
    vector, matrix or residual history anywhere in the reference tree.  The
    reference therefore cannot be executed or compiled here (`oracle/_ref` does not exist).
  * The arithmetic lives in un-vendored dependencies: PartitionedArrays 0.3.x, SparseArrays /
    SparseMatricesCSR 0.6.7, LinearAlgebra (BLAS dot/nrm2, `givensAlgorithm` = LAPACK dlartg),
    Gridap 0.19 `LUSolver` (UMFPACK).  Their published algorithms are restated here.
  * The oracle IS pinned against every known-answer test the reference holds for this path:
    KrylovTests.jl:14-26,66-93 (L2 error^2 < 1e-6 for every Krylov variant, 8^d Poisson),
    SmoothersTests.jl:12-44 (CG + Richardson(Jacobi,5,2/3): < 1e-8) -- tests/test_oracle_*.py.
  * GMG-PCG iteration counts / residual histories are NOT asserted by any reference test
    (GMGTests.jl:139-142,413) => "parity unpinned" for those; they are pinned only oracle <-> CUDA.
"""
