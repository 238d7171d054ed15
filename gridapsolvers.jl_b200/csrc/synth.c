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

/* -------------------------------------------------------------------------------------------
 * Generic tensor-product Lagrange generator (Q_p, p = order, ncomp components per node, node-major
 * dofs) on a uniform Cartesian mesh: rows of the assembled matrix for the free nodes, computed row by
 * row from ONE element matrix Ke (all cells are congruent) -- the full-size C4 system (64^3 Q2 cells,
 * 1.24e9 non-zeros) is generated in parallel without any COO intermediate.  Plays the role of the
 * Gridap assembly of test/Applications/Elasticity.jl:31-37 (C4) and of the velocity block of
 * joss_paper/demo.jl:20-91 (C5); tests/ check it against the oracle's element-by-element assembly.
 *
 *   Ke        : (nb*ncomp)^2 row-major, local dof = a*ncomp + c, local node a lexicographic (x fastest)
 *   Fe        : nb*ncomp element load vector or NULL
 *   node_free : per node of the (order*ncell+1)^d node grid (x fastest): index among the free nodes or -1
 *   free_nodes: node id of every free node (ascending)
 *   ud        : Dirichlet values per (node, comp) or NULL; b (n_free*ncomp) receives Fe sums - K_ij ud_j
 * pass 0: rowptr[r+1] = entries of row r (caller prefix-sums); pass 1: fill col/val (ascending columns) and b. */
void synth_fe_rows(int d, int order, int ncomp, const int64_t *ncell, const double *Ke, const double *Fe,
                   const int32_t *node_free, int64_t n_free_nodes, const int64_t *free_nodes, const double *ud, int pass,
                   int64_t *rowptr, int32_t *col, double *val, double *b) {
  int64_t nn[MAXD] = {1, 1, 1};
  for (int k = 0; k < d; ++k) nn[k] = (int64_t)order * ncell[k] + 1;
  const int p1 = order + 1, w = 2 * order + 1;
  int nb = 1, nslot = 1;
  for (int k = 0; k < d; ++k) { nb *= p1; nslot *= w; }
  const int nv = nb * ncomp;
#pragma omp parallel
  {
    double *S = (double *)malloc(sizeof(double) * (size_t)ncomp * nslot * ncomp);
    unsigned char *touched = (unsigned char *)malloc((size_t)nslot);
    int64_t *slot_node = (int64_t *)malloc(sizeof(int64_t) * (size_t)nslot);
#pragma omp for schedule(static)
    for (int64_t fn = 0; fn < n_free_nodes; ++fn) {
      const int64_t nid = free_nodes[fn];
      int64_t g[MAXD] = {0, 0, 0};
      g[0] = nid % nn[0];
      if (d > 1) g[1] = (nid / nn[0]) % nn[1];
      if (d > 2) g[2] = nid / (nn[0] * nn[1]);
      memset(touched, 0, (size_t)nslot);
      if (pass) memset(S, 0, sizeof(double) * (size_t)ncomp * nslot * ncomp);
      double fsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      /* cells containing the node, per direction: (cell index, local node index) */
      int64_t cc[MAXD][2];
      int ca[MAXD][2], cn[MAXD] = {1, 1, 1};
      for (int k = 0; k < d; ++k) {
        if (g[k] % order == 0) {
          const int64_t c1 = g[k] / order;
          cn[k] = 0;
          if (c1 - 1 >= 0) { cc[k][cn[k]] = c1 - 1; ca[k][cn[k]] = order; cn[k]++; }
          if (c1 < ncell[k]) { cc[k][cn[k]] = c1; ca[k][cn[k]] = 0; cn[k]++; }
        } else {
          cn[k] = 1; cc[k][0] = g[k] / order; ca[k][0] = (int)(g[k] % order);
        }
      }
      for (int iz = 0; iz < (d > 2 ? cn[2] : 1); ++iz)
        for (int iy = 0; iy < (d > 1 ? cn[1] : 1); ++iy)
          for (int ix = 0; ix < cn[0]; ++ix) {
            const int sel[MAXD] = {ix, iy, iz};
            int a = 0, stride = 1;
            for (int k = 0; k < d; ++k) { a += ca[k][sel[k]] * stride; stride *= p1; }
            if (pass && Fe)
              for (int c = 0; c < ncomp; ++c) fsum[c] += Fe[a * ncomp + c];
            for (int bn = 0; bn < nb; ++bn) {
              int rem = bn, slot = 0, sstride = 1;
              for (int k = 0; k < d; ++k) {
                const int bk = rem % p1;
                rem /= p1;
                const int64_t o = cc[k][sel[k]] * order + bk - g[k];  /* in [-order, order] */
                slot += (int)(o + order) * sstride;
                sstride *= w;
              }
              touched[slot] = 1;
              if (pass)
                for (int c = 0; c < ncomp; ++c)
                  for (int c2 = 0; c2 < ncomp; ++c2)
                    S[((size_t)c * nslot + slot) * ncomp + c2] += Ke[(size_t)(a * ncomp + c) * nv + bn * ncomp + c2];
            }
          }
      /* neighbour node of every touched slot */
      int cnt = 0;
      for (int slot = 0; slot < nslot; ++slot) {
        if (!touched[slot]) continue;
        int rem = slot;
        int64_t q = 0, qstride = 1;
        for (int k = 0; k < d; ++k) {
          const int64_t o = rem % w - order;
          rem /= w;
          q += (g[k] + o) * qstride;
          qstride *= nn[k];
        }
        slot_node[slot] = q;
        if (node_free[q] >= 0) cnt += ncomp;
      }
      for (int c = 0; c < ncomp; ++c) {
        const int64_t row = fn * ncomp + c;
        if (pass == 0) { rowptr[row + 1] = cnt; continue; }
        int64_t e = rowptr[row];
        double bi = fsum[c];
        for (int slot = 0; slot < nslot; ++slot) {
          if (!touched[slot]) continue;
          const int64_t q = slot_node[slot];
          const int32_t fq = node_free[q];
          const double *Sr = S + ((size_t)c * nslot + slot) * ncomp;
          if (fq >= 0) {
            for (int c2 = 0; c2 < ncomp; ++c2) { col[e] = fq * ncomp + c2; val[e] = Sr[c2]; ++e; }
          } else if (ud) {
            for (int c2 = 0; c2 < ncomp; ++c2) bi -= Sr[c2] * ud[q * ncomp + c2];
          }
        }
        if (b) b[row] = bi;
      }
    }
    free(S); free(touched); free(slot_node);
  }
}
