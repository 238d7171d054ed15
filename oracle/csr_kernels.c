/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product
 * path (gridapsolvers.jl_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * CPU restatement of the array primitives the reference's solve phase delegates to its
 * (un-vendored) dependencies SparseArrays / SparseMatricesCSR / LinearAlgebra / PartitionedArrays.
 * PARITY STATUS: the reference is pure Julia and cannot run in this container (no julia, no
 * depot); these kernels restate the *published* algorithms of those dependencies:
 *
 *  - SparseArrays.mul!(C, A::SparseMatrixCSC, B, alpha, beta) (Julia stdlib, spmatmul):
 *        C .*= beta (or fill 0); for each column j ascending: axj = B[j]*alpha;
 *        for each stored row i of column j: C[i] += nzval * axj
 *    => per output row the products are accumulated in ASCENDING COLUMN order, starting from
 *       beta*C[i], each product rounded before the add (Julia never contracts a*b+c to an FMA
 *       unless muladd/@fastmath is written).  A CSR row loop with ascending columns and
 *       -ffp-contract=off reproduces exactly that rounding sequence.
 *    Call sites in the reference: CGSolvers.jl:79,104; RichardsonSmoothers.jl:94;
 *    GMGLinearSolvers.jl:495,623; KrylovUtils.jl:19-52; BlockTriangularSolvers.jl:202,230.
 *  - PartitionedArrays.mul!(c, A::PSparseMatrix, b): c_own = A_oo*b_own, then
 *    c_own += A_oh*b_ghost == the same row loop when the local matrix uses
 *    own-first column numbering.
 *
 * Compile with -ffp-contract=off (the Makefile does); OpenMP only parallelises across rows /
 * elements, it never changes the per-row rounding sequence.
 */
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* y = A x   (3-arg mul!, beta == false => y zeroed first) */
void oracle_csr_spmv(int64_t nrows, const int64_t *rowptr, const int32_t *col, const double *val,
                     const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nrows; ++i) {
    double s = 0.0;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      double p = val[k] * x[col[k]];
      s = s + p;
    }
    y[i] = s;
  }
}

/* y = beta*y + A*(alpha*x)   (5-arg mul!, e.g. mul!(wi,A_ij,xj,-cij,1.0) BlockTriangularSolvers.jl:230) */
void oracle_csr_spmv5(int64_t nrows, const int64_t *rowptr, const int32_t *col, const double *val,
                      const double *x, double *y, double alpha, double beta) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nrows; ++i) {
    double s = (beta == 1.0) ? y[i] : (beta == 0.0 ? 0.0 : beta * y[i]);
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      double ax = x[col[k]] * alpha;
      double p = val[k] * ax;
      s = s + p;
    }
    y[i] = s;
  }
}

/* z = a .* b */
void oracle_ew_mul(int64_t n, const double *a, const double *b, double *z) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) z[i] = a[i] * b[i];
}
/* z = s .* a */
void oracle_ew_scale(int64_t n, double s, const double *a, double *z) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) z[i] = s * a[i];
}
/* z = a .+ b */
void oracle_ew_add(int64_t n, const double *a, const double *b, double *z) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) z[i] = a[i] + b[i];
}
/* z = a .- b */
void oracle_ew_sub(int64_t n, const double *a, const double *b, double *z) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) z[i] = a[i] - b[i];
}
/* z = a .+ s .* b   (two roundings: t = s*b; z = a + t) */
void oracle_ew_axpy(int64_t n, const double *a, double s, const double *b, double *z) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double t = s * b[i];
    z[i] = a[i] + t;
  }
}
/* z = a .- s .* b */
void oracle_ew_axmy(int64_t n, const double *a, double s, const double *b, double *z) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double t = s * b[i];
    z[i] = a[i] - t;
  }
}

/* threaded reductions: only used by the CPU *timing* baseline (summation order depends on the
 * thread count).  The checker uses numpy's BLAS dot, like Julia's LinearAlgebra.dot does. */
double oracle_dot(int64_t n, const double *a, const double *b) {
  double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
