/*
 * gsb200.h -- C ABI of libgsb200.so: a B200-native (sm_100a) solve phase for GridapSolvers.jl.
 *
 * This is the boundary a Julia shim binds with `ccall` (INTEGRATION.md shows the binding).  It
 * holds device-resident mirrors of PartitionedArrays' PSparseMatrix / PVector (one part per
 * process / GPU) and NumericalSetup mirrors of the reference's LinearSolver objects.  Each
 * entry point cites the reference interface (paths relative to /root/reference/src) it replaces.
 *
 * Conventions
 *   - every call returns 0 on success, a GSB_E* code otherwise; gsb_last_error() gives the text.
 *     Non-convergence is NOT an error (ConvergenceLogs.jl:136-150): it is reported by the flag.
 *   - handles are opaque pointers; the caller owns them and destroys them (precedent for foreign
 *     NumericalSetups owning C objects: ext/GridapPETScExt/PETScCaches.jl:2-47).
 *   - all floating point data is fp64; index arrays are int32 or int64 (index_bytes), 0- or
 *     1-based (index_base) -- Julia passes its Int64 1-based arrays untouched.
 *   - vectors are laid out own values first, ghost values after (PartitionedArrays own-first
 *     local numbering).  Reductions only ever see own values.
 *   - one context per process, one GPU per context, calls from one host thread; gsb_solve,
 *     gsb_vec_get, gsb_dot and friends block until the result is on the host.
 *   - there is NO CPU fallback: every numerical entry point needs a CUDA device.
 */
#ifndef GSB200_H
#define GSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gsb_ctx_s *gsb_ctx_t;
typedef struct gsb_plan_s *gsb_plan_t;
typedef struct gsb_mat_s *gsb_mat_t;
typedef struct gsb_vec_s *gsb_vec_t;
typedef struct gsb_solver_s *gsb_solver_t;

enum { GSB_OK = 0, GSB_EINVAL = 1, GSB_ECUDA = 2, GSB_ENCCL = 3, GSB_ENOMEM = 4, GSB_EUNSUPPORTED = 5 };
enum { GSB_FMT_CSR = 0, GSB_FMT_CSC = 1 };
/* SolverConvergenceFlag, SolverInterfaces/SolverTolerances.jl:11-16 */
enum { GSB_CONVERGED_ATOL = 0, GSB_CONVERGED_RTOL = 1, GSB_DIVERGED_MAXITER = 2, GSB_DIVERGED_BREAKDOWN = 3 };
/* GMGLinearSolvers.jl:56-57 (mode, cycle_type) */
enum { GSB_GMG_PRECONDITIONER = 0, GSB_GMG_SOLVER = 1 };
enum { GSB_V_CYCLE = 0, GSB_W_CYCLE = 1, GSB_F_CYCLE = 2 };
enum { GSB_UPPER = 0, GSB_LOWER = 1 };

#define GSB_NCCL_ID_BYTES 128

/* ---------------------------------------------------------------- library / context */
int gsb_version(void);
/* text of the last error raised on this thread (ctx may be NULL) */
const char *gsb_last_error(gsb_ctx_t ctx);
/* rank 0 creates the NCCL unique id and distributes the 128 bytes by its own means
 * (MPI.bcast in Julia, torch.distributed in the Python harness). */
int gsb_nccl_unique_id(void *out_id_128_bytes);
/* one part of a PartitionedArrays distribution == one rank == one GPU.  nranks==1: id may be NULL.
 * Replaces: with_mpi/distribute + MPI communicator (test/LinearSolvers/mpi/runtests.jl:5-20). */
int gsb_init(int device, int nranks, int rank, const void *nccl_id, gsb_ctx_t *out);
int gsb_finalize(gsb_ctx_t ctx);
int gsb_synchronize(gsb_ctx_t ctx);
/* CUDA-event timer on the context's compute stream (what bench.py times kernels with) */
int gsb_timer_start(gsb_ctx_t ctx);
int gsb_timer_stop(gsb_ctx_t ctx, float *ms);
/* number of kernels this library has launched on this context so far */
int gsb_launch_count(gsb_ctx_t ctx, int64_t *out);
/* per-launch CUDA-event timing of the row kernels (SpMV / residual / sweep / transfer), aggregated
 * by (mode, kernel kind, rows, nnz); mode = 0 spmv, 1 residual, 2 fused sweep, 3 spmv+dot, 4 spmv+add;
 * kernel kind = 0 CSR fallback kernel, 2 + 10*block_size (+100 when the rows are sorted) block-SELL-32 */
int gsb_profile_start(gsb_ctx_t ctx);
int gsb_profile_stop(gsb_ctx_t ctx, int cap, int *n_out, int *mode, int *stream_kernel, int64_t *nrows, int64_t *nnz,
                     int *count, double *total_ms);
/* diagnostics (pure host): the block-SELL-32 plan the library builds for a CSR matrix (int32, 0-based, ascending
 * columns): out[12] = {ok, block size, sorted, block rows, slices, stored blocks incl. padding, blocks without
 * padding, boundary slices, explicit column-id lines, (slice,k) pairs, diagonal-aligned slices, slices made of runs of three consecutive columns}; pos_row / pos_len
 * / pos_mask (n_slices*32 ints or NULL): block row (-1 = padding lane), length in blocks and slot-validity word (slice
 * width <= 32: bit k set <=> slot k holds a block; else the length) of every (slice, lane) position; col_words (one
 * int per (slice,k) pair = stored blocks / 32, or NULL): >= 0 the affine base block column (lane l uses base + l),
 * < 0 the complement of the explicit id line.  sort_mode: -1 auto, 0 never, 1 always */
int gsb_diag_sell_plan(int64_t n_rows, int64_t n_own_cols, int64_t n_ghost_cols, const int *rowptr, const int *col,
                       int detect_blocks, int sort_mode, int64_t *out, int *pos_row, int *pos_len, int *pos_mask, int *col_words);
/* diagnostics: average duration of `reps` back-to-back launches of one row-kernel mode on scratch vectors */
int gsb_bench_rows(gsb_mat_t A, int mode, int reps, float *avg_ms);
/* runtime knobs (also read from the environment variable GSB_OPTIONS="k=v,k=v" at gsb_init); for tests/tuning:
 *   spmv=auto|vector   row-kernel family (vector = CSR fallback kernel, only for matrices that keep their CSR arrays)
 *   sell=1|0, block=1|0, sell_sort=auto|0|1   block-SELL-32 storage: on/off, DOF-block detection, row sorting (at matrix creation)
 *   keep_csr=0|1, keep_csr_max_nnz=N          keep the CSR column ids / values on the device next to block-SELL
 *   graph=1|0      CUDA-graph replay of a maxiter=1 GMG      gmg_defer_log=1|0     device-resident GMG log norms
 *   p2p=1|0        NVLink peer-memory halo (at plan creation) overlap=0|1           two-stream halo overlap
 *   fuse_smoother=1|0  fused Jacobi-Richardson sweeps         fuse_mgs=1|0          fused modified Gram-Schmidt steps */
int gsb_set_option(gsb_ctx_t ctx, const char *key, const char *value);

/* ---------------------------------------------------------------- exchange plan
 * The neighbour / local-id lists of a PartitionedArrays index partition's assembly cache, in
 * the direction of consistent! (owner -> ghost).  snd_local_ids index OWN entries to pack for
 * each send neighbour, rcv_local_ids index local (>= n_own) ghost entries to fill per receive
 * neighbour.  Replaces: PartitionedArrays consistent!/assemble! caches as read at
 * SolverInterfaces/PAExtras.jl:9-110; call sites GridapExtras.jl:42,56.
 * COLLECTIVE when nranks > 1: every rank must create its plans in the same order (the CUDA-IPC handles of the
 * peer-memory receive buffers are all-gathered here; if any rank cannot map a peer, all ranks fall back to
 * ncclSend/ncclRecv for this plan).  Option "p2p=0" (gsb_set_option / GSB_OPTIONS) forces the NCCL path. */
int gsb_plan_create(gsb_ctx_t ctx, int64_t n_own, int64_t n_ghost, int n_nbr_snd, const int32_t *nbr_snd,
                    const int64_t *snd_ptrs, const int64_t *snd_local_ids, int n_nbr_rcv,
                    const int32_t *nbr_rcv, const int64_t *rcv_ptrs, const int64_t *rcv_local_ids,
                    int index_base, gsb_plan_t *out);
int gsb_plan_destroy(gsb_plan_t plan);
/* Redistribution plan: moves the OWN values of a vector between two row partitions of the same global index space
 * (the same plan type and transport as the exchange plan; destroy with gsb_plan_destroy).  snd_local_ids index own
 * entries of the SOURCE layout (n_src_own of them on this rank) to send to each neighbour, rcv_local_ids own entries
 * of the DESTINATION layout (n_dst_own) to fill; a rank may list itself as a neighbour.  Replaces: MultilevelTools
 * RedistributionOperator / redistribute_free_values! (MultilevelTools/GridTransferOperators.jl:2-157,447-532,
 * RedistributionOperators.jl) and the level-on-fewer-parts semantics of HierarchicalArrays.jl:96-149 (a rank that
 * does not hold a level passes matrices / vectors with zero own rows).  COLLECTIVE like gsb_plan_create.  The two
 * directions of a repartition are two plans; callers alternate them (restrict ... prolongate), which is what keeps
 * the double-buffered peer-memory transport safe for the one-directional neighbour graph of a redistribution. */
int gsb_redist_create(gsb_ctx_t ctx, int64_t n_src_own, int64_t n_dst_own, int n_nbr_snd, const int32_t *nbr_snd,
                      const int64_t *snd_ptrs, const int64_t *snd_local_ids, int n_nbr_rcv,
                      const int32_t *nbr_rcv, const int64_t *rcv_ptrs, const int64_t *rcv_local_ids,
                      int index_base, gsb_plan_t *out);
/* dst(own, destination layout) <- src(own, source layout) through a redistribution plan */
int gsb_vec_redistribute(gsb_plan_t plan, gsb_vec_t src, gsb_vec_t dst);

/* ---------------------------------------------------------------- PSparseMatrix mirror
 * Local block of partition(A): n_rows own rows x (n_own_cols + n_ghost_cols) local columns in
 * own-first numbering.  fmt CSC (Julia's default SparseMatrixCSC{Float64,Int64}) or CSR
 * (SparseMatricesCSR); stored on device as CSR fp64 / int32 with ascending columns per row, so
 * per-row accumulation order equals the reference's.  plan may be NULL when
 * n_ghost_cols == 0.  Replaces: the matrix argument of symbolic_setup/numerical_setup
 * (LinearSolvers/Krylov/CGSolvers.jl:31-55). */
int gsb_mat_create(gsb_ctx_t ctx, int64_t n_rows, int64_t n_own_cols, int64_t n_ghost_cols, int fmt,
                   int index_base, int index_bytes, const void *ptr, const void *idx, const double *vals,
                   gsb_plan_t plan, gsb_mat_t *out);
/* same sparsity, new values, in the order given at creation -- numerical_setup!(ns,A) support */
int gsb_mat_update_values(gsb_mat_t A, const double *vals);
int gsb_mat_info(gsb_mat_t A, int64_t *n_rows, int64_t *n_own_cols, int64_t *n_ghost_cols, int64_t *nnz);
/* device storage the row kernels stream: kind 0 = CSR, 1 = block-SELL-32 (block_size x block_size DOF blocks, rows
 * optionally sorted by length inside windows of 256); stored_entries includes padding; bytes_per_pass = bytes
 * of matrix data one SpMV-type kernel reads (values + ids + per-row metadata).  Any out pointer may be NULL. */
int gsb_mat_format(gsb_mat_t A, int *kind, int *block_size, int *sorted, int64_t *stored_entries, int64_t *bytes_per_pass);
int gsb_mat_destroy(gsb_mat_t A);

/* ---------------------------------------------------------------- PVector mirror */
int gsb_vec_create(gsb_ctx_t ctx, int64_t n_own, int64_t n_ghost, gsb_vec_t *out);
/* allocate_in_domain(A) / allocate_in_range(A) (RichardsonSmoothers.jl:59-60), zero-filled */
int gsb_vec_create_domain(gsb_mat_t A, gsb_vec_t *out);
int gsb_vec_create_range(gsb_mat_t A, gsb_vec_t *out);
int gsb_vec_destroy(gsb_vec_t v);
int gsb_vec_size(gsb_vec_t v, int64_t *n_own, int64_t *n_ghost);
/* own values <-> host (own_values(v) .= host / host .= own_values(v)) */
int gsb_vec_set(gsb_vec_t v, const double *host, int64_t n);
int gsb_vec_get(gsb_vec_t v, double *host, int64_t n);
/* whole local vector incl. ghosts (tests) */
int gsb_vec_get_local(gsb_vec_t v, double *host, int64_t n);
int gsb_vec_fill(gsb_vec_t v, double value);  /* fill!(v,value) */
int gsb_vec_copy(gsb_vec_t dst, gsb_vec_t src); /* copy!(dst,src): own values */
/* consistent!(v) |> wait : owner -> ghost halo update through the plan */
int gsb_vec_consistent(gsb_vec_t v, gsb_plan_t plan);
/* assemble!(v) |> wait : ghost -> owner accumulation through the same plan reversed (contributions added
 * neighbour after neighbour), ghost entries zeroed afterwards.  Call sites in the reference:
 * LinearSolvers/SchwarzLinearSolvers.jl:44-49, MultilevelTools/GridTransferOperators.jl:425,544 */
int gsb_vec_assemble(gsb_vec_t v, gsb_plan_t plan);
/* page-lock / release a caller-owned host buffer (e.g. the storage of a Julia Vector{Float64}) so that
 * gsb_solve_host / gsb_vec_set / gsb_vec_get run at pinned-memory speed */
int gsb_host_register(gsb_ctx_t ctx, void *ptr, int64_t bytes);
int gsb_host_unregister(gsb_ctx_t ctx, void *ptr);

/* ---------------------------------------------------------------- array primitives */
/* mul!(y,A,x,alpha,beta): y = beta*y + A*(alpha*x), halo of x included (3-arg mul! = alpha 1, beta 0) */
int gsb_spmv(gsb_mat_t A, gsb_vec_t x, gsb_vec_t y, double alpha, double beta);
int gsb_dot(gsb_vec_t a, gsb_vec_t b, double *out);   /* dot(a,b): own values + allreduce */
int gsb_norm2(gsb_vec_t a, double *out);              /* norm(a) */
int gsb_axpby(gsb_vec_t z, double alpha, gsb_vec_t x, double beta, gsb_vec_t y); /* z .= alpha.*x .+ beta.*y */

/* ---------------------------------------------------------------- NumericalSetup mirrors
 * Each *_create is numerical_setup(symbolic_setup(solver,A),A) of the reference solver; the
 * returned handle is the NumericalSetup.  Nested solvers are passed as handles and stay owned by
 * the caller (they must outlive the parent). */
/* IdentitySolver, IdentityLinearSolvers.jl:2-26 */
int gsb_identity_create(gsb_ctx_t ctx, gsb_solver_t *out);
/* JacobiLinearSolver, JacobiLinearSolvers.jl:20-56 (inv_diag of the own-own block; own values only) */
int gsb_jacobi_create(gsb_mat_t A, gsb_solver_t *out);
/* RichardsonSmoother(M,niter,omega), RichardsonSmoothers.jl:20-98 -- solve! mutates x AND r */
int gsb_richardson_create(gsb_mat_t A, gsb_solver_t M, int niter, double omega, gsb_solver_t *out);
/* LinearSolverFromSmoother, LinearSolverFromSmoothers.jl:1-50 */
int gsb_from_smoother_create(gsb_mat_t A, gsb_solver_t smoother, gsb_solver_t *out);
/* Gridap.Algebra.LUSolver stand-in for the GMG coarsest level: dense fp64 inverse computed and
 * applied on the device (GMGLinearSolvers.jl:423-434,472-474).  Gathers over all ranks. */
int gsb_dense_lu_create(gsb_mat_t A, gsb_solver_t *out);
/* GMGLinearSolver(matrices,prolongations,restrictions;...), GMGLinearSolvers.jl:48-69,183-210.
 * interp[l]/restrict[l] are explicit sparse transfer matrices (any object with mul! is legal for
 * the reference, GMGLinearSolvers.jl:484,491); pre/post have nlev-1 entries and may alias. */
int gsb_gmg_create(gsb_ctx_t ctx, int nlev, const gsb_mat_t *mats, const gsb_mat_t *interp,
                   const gsb_mat_t *restrict_, const gsb_solver_t *pre, const gsb_solver_t *post,
                   gsb_solver_t coarsest, int mode, int cycle_type, int maxiter, double atol, double rtol,
                   gsb_solver_t *out);
/* GMG whose coarse levels live on fewer parts (ModelHierarchy np_per_level; GridTransferOperators.jl:391-401,536-561 with
 * redist = Val{true}): to_coarse[l] / to_fine[l] (nlev-1 entries, NULL = level l+2 lives on the parts of level l+1) are
 * redistribution plans between the row layout of restrict[l] / the column layout of interp[l] (coarse space in the
 * partition of the finer level's parts) and the layout of mats[l+1].  Ranks that do not hold a level pass empty
 * (zero-row) matrices for it; every rank still makes every call (the collectives are global). */
int gsb_gmg_create_redist(gsb_ctx_t ctx, int nlev, const gsb_mat_t *mats, const gsb_mat_t *interp,
                          const gsb_mat_t *restrict_, const gsb_solver_t *pre, const gsb_solver_t *post,
                          gsb_solver_t coarsest, int mode, int cycle_type, int maxiter, double atol, double rtol,
                          const gsb_plan_t *to_coarse, const gsb_plan_t *to_fine, gsb_solver_t *out);
/* CGSolver(Pl;maxiter,atol,rtol,flexible), Krylov/CGSolvers.jl:10-120.  Pl may be NULL. */
int gsb_cg_create(gsb_mat_t A, gsb_solver_t Pl, int flexible, int maxiter, double atol, double rtol,
                  gsb_solver_t *out);
/* GMRESSolver(m;Pr,Pl,restart,m_add,...), Krylov/GMRESSolvers.jl:16-210 */
int gsb_gmres_create(gsb_mat_t A, gsb_solver_t Pr, gsb_solver_t Pl, int m, int restart, int m_add,
                     int maxiter, double atol, double rtol, gsb_solver_t *out);
/* FGMRESSolver(m,Pr;Pl,restart,m_add,...), Krylov/FGMRESSolvers.jl:17-199 */
int gsb_fgmres_create(gsb_mat_t A, gsb_solver_t Pr, gsb_solver_t Pl, int m, int restart, int m_add,
                      int maxiter, double atol, double rtol, gsb_solver_t *out);
/* MINRESSolver(;Pl,...), Krylov/MINRESSolvers.jl:11-149 */
int gsb_minres_create(gsb_mat_t A, gsb_solver_t Pl, int maxiter, double atol, double rtol, gsb_solver_t *out);
/* BlockTriangularSolver / BlockDiagonalSolver on an nb x nb block system stored as one
 * concatenated vector per side (BlockSolvers/BlockTriangularSolvers.jl:188-242,
 * BlockDiagonalSolvers.jl:165-177).  blocks is row-major nb*nb (NULL = zero block); coeffs
 * row-major nb*nb or NULL (all ones); half GSB_UPPER/GSB_LOWER; for the diagonal solver pass
 * diagonal = 1.  Work caches y are zeroed only at creation (BlockTriangularSolvers.jl:139). */
int gsb_block_solver_create(gsb_ctx_t ctx, int nb, const gsb_mat_t *blocks, const gsb_solver_t *solvers,
                            const double *coeffs, int half, int diagonal, gsb_solver_t *out);
/* a block matrix acting on concatenated vectors, usable as the A of the Krylov solvers */
int gsb_block_mat_create(gsb_ctx_t ctx, int nb, const gsb_mat_t *blocks, gsb_mat_t *out);

/* RichardsonLinearSolver(omega,maxiter;Pl,rtol,atol), LinearSolvers/RichardsonLinearSolvers.jl:12-106 */
int gsb_richardson_linear_create(gsb_mat_t A, gsb_solver_t Pl, double omega, int maxiter, double atol, double rtol,
                                 gsb_solver_t *out);
/* SchurComplementSolver(A_ns,B,C,S_ns) on a 2-block concatenated vector [u;p],
 * LinearSolvers/SchurComplementSolvers.jl:8-74 */
int gsb_schur_complement_create(gsb_ctx_t ctx, gsb_solver_t A_ns, gsb_mat_t B, gsb_mat_t C, gsb_solver_t S_ns,
                                gsb_solver_t *out);
/* LanczosDiagnostic support (Krylov/KrylovUtils.jl:58-90, CGSolvers.jl:122-138): record alpha_k, beta_k of a
 * CG NumericalSetup during solve! and read them back (n = number recorded, at most cap copied) */
int gsb_cg_record_coefficients(gsb_solver_t cg_ns, int enable);
int gsb_cg_coefficients(gsb_solver_t cg_ns, double *alpha, double *beta, int64_t cap, int64_t *n);
/* numerical_setup!(ns,A): refresh value-dependent data (inv_diag, dense inverse) after
 * gsb_mat_update_values; GMG-from-matrices does not support it (GMGLinearSolvers.jl:249-258). */
int gsb_solver_update(gsb_solver_t ns, gsb_mat_t A);
/* solve!(x,ns,b) on device vectors */
int gsb_solve(gsb_solver_t ns, gsb_vec_t x, gsb_vec_t b);
/* solve!(x,ns,b) with HOST own-value buffers: H2D of b and x0, solve, D2H of x (the e2e path) */
int gsb_solve_host(gsb_solver_t ns, double *x_host, const double *b_host, int64_t n);
/* same, the caller vouching for x0 = 0 (the usual `x = allocate_in_domain(A); fill!(x,0)` of the reference's
 * drivers, test/LinearSolvers/GMGTests.jl:124-128): the initial guess is not uploaded */
int gsb_solve_host_zero_guess(gsb_solver_t ns, double *x_host, const double *b_host, int64_t n);
/* ConvergenceLog read-back (ConvergenceLogs.jl:42-49): num_iters, residuals[0..num_iters], flag */
int gsb_solver_log(gsb_solver_t ns, int *num_iters, double *residuals, int64_t cap, int *flag);
int gsb_solver_destroy(gsb_solver_t ns);

#ifdef __cplusplus
}
#endif
#endif /* GSB200_H */
