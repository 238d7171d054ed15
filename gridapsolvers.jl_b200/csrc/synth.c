/*
 * synth.c -- host-side generator of the synthetic benchmark systems (libgsb200_synth.so).
 *
 * NOT part of the solve path: it plays the role of the Gridap/GridapDistributed assembly the
 * Julia host performs before calling the solvers (SURVEY.md 8d "Concrete synthetic inputs"):
 * Q1 Poisson on a uniform Cartesian mesh of [0,L1]x..x[0,Ld] (L = 1 by default), Dirichlet on the whole boundary,
 * manufactured u = x + y (test/LinearSolvers/KrylovTests.jl:11-12,46-61, GMGTests.jl:204-215),
 * and the factor-2 nodal prolongation / its transpose between nested levels
 * (src/MultilevelTools/GridTransferOperators.jl:391-401,536-561).  Each rank generates only its
 * own rows in PartitionedArrays-style own-first local numbering: `ext_lid` maps every node of the
 * rank's extended box (own box + one layer) to a local id (>=0), -2 = Dirichlet node.
 * The values are the assembled Q1 stencils; tests/ check them against the
 * oracle's genuine element-by-element assembly.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXD 3

static inline int64_t ext_index(int d, const int64_t *elo, const int64_t *ehi, const int64_t *g) {
  int64_t idx = 0, stride = 1;
  for (int k = 0; k < d; ++k) {
    idx += (g[k] - elo[k]) * stride;
    stride *= (ehi[k] - elo[k]);
  }
  return idx;
}

static void sort_row(int n, int32_t *c, double *v) {
  for (int i = 1; i < n; ++i) {
    int32_t ci = c[i];
    double vi = v[i];
    int j = i - 1;
    while (j >= 0 && c[j] > ci) {
      c[j + 1] = c[j];
      v[j + 1] = v[j];
      --j;
    }
    c[j + 1] = ci;
    v[j + 1] = vi;
  }
}

/* rows of the Q1 Laplacian for the own box [olo,ohi) (global node coords), and the Dirichlet
 * lift b_i = -sum_{j Dirichlet} A_ij (x_j + y_j).  rowptr has n_own+1 entries.
 * pass 0: fill rowptr counts (rowptr[i+1] = nnz of row i, caller prefix-sums); pass 1: fill. */
void synth_poisson_rows(int d, const int64_t *ncell, const double *lengths, const int64_t *elo, const int64_t *ehi, const int32_t *ext_lid,
                        const int64_t *olo, const int64_t *ohi, int pass, int64_t *rowptr, int32_t *col, double *val,
                        double *b) {
  double h[MAXD], kd[MAXD][3], md[MAXD][3];
  for (int k = 0; k < d; ++k) {
    h[k] = lengths[k] / (double)ncell[k];
    kd[k][0] = kd[k][2] = -1.0 / h[k];
    kd[k][1] = 2.0 / h[k];
    md[k][0] = md[k][2] = h[k] / 6.0;
    md[k][1] = 2.0 * h[k] / 3.0;
  }
  int64_t on[MAXD] = {1, 1, 1};
  for (int k = 0; k < d; ++k) on[k] = ohi[k] - olo[k];
  const int64_t nown = on[0] * on[1] * on[2];
  const int nst = d == 2 ? 9 : (d == 3 ? 27 : 3);
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < nown; ++r) {
    int64_t g[MAXD];
    g[0] = olo[0] + r % on[0];
    if (d > 1) g[1] = olo[1] + (r / on[0]) % on[1];
    if (d > 2) g[2] = olo[2] + r / (on[0] * on[1]);
    int32_t c[27];
    double v[27];
    int cnt = 0;
    double bi = 0.0;
    for (int s = 0; s < nst; ++s) {
      int a[MAXD];
      a[0] = s % 3 - 1;
      a[1] = d > 1 ? (s / 3) % 3 - 1 : 0;
      a[2] = d > 2 ? s / 9 - 1 : 0;
      int64_t q[MAXD];
      for (int k = 0; k < d; ++k) q[k] = g[k] + a[k];
      double sv = 0.0;
      for (int t = 0; t < d; ++t) {
        double p = 1.0;
        for (int k = 0; k < d; ++k) p *= (k == t) ? kd[k][a[k] + 1] : md[k][a[k] + 1];
        sv += p;
      }
      const int32_t lid = ext_lid[ext_index(d, elo, ehi, q)];
      if (lid >= 0) {
        c[cnt] = lid;
        v[cnt] = sv;
        cnt++;
      } else if (lid == -2) {
        const double gx = (double)q[0] * h[0], gy = d > 1 ? (double)q[1] * h[1] : 0.0;
        bi -= sv * (gx + gy);
      }
    }
    if (pass == 0) {
      rowptr[r + 1] = cnt;
    } else {
      sort_row(cnt, c, v);
      const int64_t e0 = rowptr[r];
      for (int i = 0; i < cnt; ++i) {
        col[e0 + i] = c[i];
        val[e0 + i] = v[i];
      }
      if (b) b[r] = bi;
    }
  }
}

/* mass-matrix rows (for L2 errors in the known-answer tests); same calling convention */
void synth_mass_rows(int d, const int64_t *ncell, const double *lengths, const int64_t *elo, const int64_t *ehi, const int32_t *ext_lid,
                     const int64_t *olo, const int64_t *ohi, int pass, int64_t *rowptr, int32_t *col, double *val) {
  double h[MAXD], md[MAXD][3];
  for (int k = 0; k < d; ++k) {
    h[k] = lengths[k] / (double)ncell[k];
    md[k][0] = md[k][2] = h[k] / 6.0;
    md[k][1] = 2.0 * h[k] / 3.0;
  }
  int64_t on[MAXD] = {1, 1, 1};
  for (int k = 0; k < d; ++k) on[k] = ohi[k] - olo[k];
  const int64_t nown = on[0] * on[1] * on[2];
  const int nst = d == 2 ? 9 : (d == 3 ? 27 : 3);
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < nown; ++r) {
    int64_t g[MAXD];
    g[0] = olo[0] + r % on[0];
    if (d > 1) g[1] = olo[1] + (r / on[0]) % on[1];
    if (d > 2) g[2] = olo[2] + r / (on[0] * on[1]);
    int32_t c[27];
    double v[27];
    int cnt = 0;
    for (int s = 0; s < nst; ++s) {
      int a[MAXD];
      a[0] = s % 3 - 1;
      a[1] = d > 1 ? (s / 3) % 3 - 1 : 0;
      a[2] = d > 2 ? s / 9 - 1 : 0;
      int64_t q[MAXD];
      for (int k = 0; k < d; ++k) q[k] = g[k] + a[k];
      double p = 1.0;
      for (int k = 0; k < d; ++k) p *= md[k][a[k] + 1];
      const int32_t lid = ext_lid[ext_index(d, elo, ehi, q)];
      if (lid >= 0) { c[cnt] = lid; v[cnt] = p; cnt++; }
    }
    if (pass == 0) rowptr[r + 1] = cnt;
    else {
      sort_row(cnt, c, v);
      const int64_t e0 = rowptr[r];
      for (int i = 0; i < cnt; ++i) { col[e0 + i] = c[i]; val[e0 + i] = v[i]; }
    }
  }
}

/* prolongation rows: fine own box [olo,ohi) (fine node coords) x coarse local ids (coarse ext box) */
void synth_prolong_rows(int d, const int64_t *olo, const int64_t *ohi, const int64_t *celo, const int64_t *cehi,
                        const int32_t *c_ext_lid, int pass, int64_t *rowptr, int32_t *col, double *val) {
  int64_t on[MAXD] = {1, 1, 1};
  for (int k = 0; k < d; ++k) on[k] = ohi[k] - olo[k];
  const int64_t nown = on[0] * on[1] * on[2];
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < nown; ++r) {
    int64_t g[MAXD] = {0, 0, 0};
    g[0] = olo[0] + r % on[0];
    if (d > 1) g[1] = olo[1] + (r / on[0]) % on[1];
    if (d > 2) g[2] = olo[2] + r / (on[0] * on[1]);
    int64_t ci[MAXD][2];
    double cw[MAXD][2];
    int cn[MAXD] = {1, 1, 1};
    for (int k = 0; k < d; ++k) {
      if (g[k] % 2 == 0) { cn[k] = 1; ci[k][0] = g[k] / 2; cw[k][0] = 1.0; }
      else { cn[k] = 2; ci[k][0] = (g[k] - 1) / 2; ci[k][1] = (g[k] + 1) / 2; cw[k][0] = cw[k][1] = 0.5; }
    }
    int32_t c[8];
    double v[8];
    int cnt = 0;
    for (int iz = 0; iz < (d > 2 ? cn[2] : 1); ++iz)
      for (int iy = 0; iy < (d > 1 ? cn[1] : 1); ++iy)
        for (int ix = 0; ix < cn[0]; ++ix) {
          int64_t q[MAXD];
          double w = cw[0][ix];
          q[0] = ci[0][ix];
          if (d > 1) { q[1] = ci[1][iy]; w *= cw[1][iy]; }
          if (d > 2) { q[2] = ci[2][iz]; w *= cw[2][iz]; }
          const int32_t lid = c_ext_lid[ext_index(d, celo, cehi, q)];
          if (lid >= 0) { c[cnt] = lid; v[cnt] = w; cnt++; }
        }
    if (pass == 0) rowptr[r + 1] = cnt;
    else {
      sort_row(cnt, c, v);
      const int64_t e0 = rowptr[r];
      for (int i = 0; i < cnt; ++i) { col[e0 + i] = c[i]; val[e0 + i] = v[i]; }
    }
  }
}

/* restriction rows (= P^T): coarse own box [olo,ohi) (coarse node coords) x fine local ids (fine ext box) */
void synth_restrict_rows(int d, const int64_t *olo, const int64_t *ohi, const int64_t *felo, const int64_t *fehi,
                         const int32_t *f_ext_lid, int pass, int64_t *rowptr, int32_t *col, double *val) {
  int64_t on[MAXD] = {1, 1, 1};
  for (int k = 0; k < d; ++k) on[k] = ohi[k] - olo[k];
  const int64_t nown = on[0] * on[1] * on[2];
  const int nst = d == 2 ? 9 : (d == 3 ? 27 : 3);
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < nown; ++r) {
    int64_t g[MAXD] = {0, 0, 0};
    g[0] = olo[0] + r % on[0];
    if (d > 1) g[1] = olo[1] + (r / on[0]) % on[1];
    if (d > 2) g[2] = olo[2] + r / (on[0] * on[1]);
    int32_t c[27];
    double v[27];
    int cnt = 0;
    for (int s = 0; s < nst; ++s) {
      int a[MAXD];
      a[0] = s % 3 - 1;
      a[1] = d > 1 ? (s / 3) % 3 - 1 : 0;
      a[2] = d > 2 ? s / 9 - 1 : 0;
      int64_t q[MAXD];
      double w = 1.0;
      for (int k = 0; k < d; ++k) { q[k] = 2 * g[k] + a[k]; w *= a[k] == 0 ? 1.0 : 0.5; }
      const int32_t lid = f_ext_lid[ext_index(d, felo, fehi, q)];
      if (lid >= 0) { c[cnt] = lid; v[cnt] = w; cnt++; }
    }
    if (pass == 0) rowptr[r + 1] = cnt;
    else {
      sort_row(cnt, c, v);
      const int64_t e0 = rowptr[r];
      for (int i = 0; i < cnt; ++i) { col[e0 + i] = c[i]; val[e0 + i] = v[i]; }
    }
  }
}
