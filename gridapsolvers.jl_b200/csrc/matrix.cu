// matrix.cu -- device mirror of the local block of a PSparseMatrix: upload (CSR or CSC, int32/int64, 0/1-based),
// host-side analysis of the sparsity (DOF block size, row-length sorting, interior/boundary slices) and
// conversion to the block-SELL-32 layout the row kernels stream (the conversion itself runs on the device,
// core.cu sell_build).  Pure host code + CUDA runtime calls; the kernels live in kernels.cuh / core.cu.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "api_macros.h"
#include "ops.h"

using namespace gsb;

int64_t gsb_mat_s::format_bytes() const {
  if (nb > 0) {
    int64_t t = 0;
    for (gsb_mat_t b : blocks)
      if (b) t += b->format_bytes();
    return t;
  }
  if (sell_ok)  // values + column words (one per 32 blocks) + explicit id lines + per-position length (+ permutation) + slice offsets
    return sell_blocks * ((int64_t)bs * bs * 8) + (sell_blocks / 32) * 4 + sell_explicit * 128 + n_slices * 32 * (sorted ? 8 : 4) +
           (n_slices + 1) * 4;
  return nnz * 12 + (n_rows + 1) * 4;
}

namespace {

int64_t rd_idx(const void *p, int bytes, int64_t i) {
  return bytes == 8 ? ((const int64_t *)p)[i] : (int64_t)((const int32_t *)p)[i];
}

// true when rows [BS*b0, BS*b1) are made of aligned BS x BS blocks: the BS rows of a block row have the same
// column list, and that list is a sequence of runs {BS*m, .., BS*m + BS-1}
bool blocks_ok(int BS, const int *rowptr, const int *col, int64_t b0, int64_t b1) {
  for (int64_t b = b0; b < b1; ++b) {
    const int e0 = rowptr[b * BS];
    const int L = rowptr[b * BS + 1] - e0;
    if (L % BS) return false;
    for (int i = 1; i < BS; ++i)
      if (rowptr[b * BS + i + 1] - rowptr[b * BS + i] != L) return false;
    for (int q = 0; q < L; q += BS) {
      const int c0 = col[e0 + q];
      if (c0 % BS) return false;
      for (int j = 1; j < BS; ++j)
        if (col[e0 + q + j] != c0 + j) return false;
    }
    for (int i = 1; i < BS; ++i)
      if (L && std::memcmp(col + rowptr[b * BS + i], col + e0, sizeof(int) * (size_t)L) != 0) return false;
  }
  return true;
}

int detect_block_size(int64_t n_rows, int64_t n_cols, const int *rowptr, const int *col) {
  for (int BS : {3, 2}) {
    if (n_rows == 0 || n_rows % BS || n_cols % BS) continue;
    const int64_t nb = n_rows / BS;
    // cheap probes first (most scalar matrices fail on the first block row)
    const int64_t probe = std::min<int64_t>(nb, 64);
    if (!blocks_ok(BS, rowptr, col, 0, probe)) continue;
    if (!blocks_ok(BS, rowptr, col, nb / 2, std::min(nb, nb / 2 + probe))) continue;
    int ok = 1;
    const int64_t chunk = 4096;
    const int64_t nchunks = (nb + chunk - 1) / chunk;
#pragma omp parallel for schedule(dynamic, 4) reduction(&& : ok)
    for (int64_t c = 0; c < nchunks; ++c)
      if (ok) ok = ok && blocks_ok(BS, rowptr, col, c * chunk, std::min(nb, (c + 1) * chunk));
    if (ok) return BS;
  }
  return 1;
}

constexpr int64_t SELL_WINDOW = 256;  // sigma: block rows are sorted by length inside windows of this many

// host-side plan of the block-SELL layout of a CSR matrix (pure host code: also reachable without a device
// through gsb_diag_sell_plan, for the CPU test-suite)
struct SellPlan {
  bool ok = false;
  int bs = 1;
  bool sorted = false;
  int64_t n_brows = 0, n_slices = 0, blocks = 0, sum_blocks = 0, n_explicit = 0;
  int64_t aligned_slices = 0;
  std::vector<int> perm, blen, off;      // perm empty when !sorted; blen by position (blocks per block row), padded with 0
  std::vector<int> lmask;                // per position: validity word of its slots (bit mask, or length when width > 32)
  std::vector<int> kbase;                // per (slice, k): affine base block column, or ~(explicit line)
  std::vector<int> kind;                 // per slice: 1 = aligned slice made of runs of three consecutive columns
  int64_t triple_slices = 0;
  std::vector<int> int_slices, bnd_slices;
};

SellPlan plan_sell(int64_t n_rows, int64_t n_own_cols, int64_t n_ghost_cols, const int *rowptr, const int *col, bool detect_blocks,
                   const std::string &sort_opt, bool affine = true, bool dia = true, bool triples = true) {
  SellPlan P;
  if (n_rows == 0) return P;
  const int64_t n_cols = n_own_cols + n_ghost_cols;
  const int64_t nnz = rowptr[n_rows];
  const int BS = detect_blocks ? detect_block_size(n_rows, n_cols, rowptr, col) : 1;
  const int64_t nb = n_rows / BS;
  const int64_t nsl = (nb + 31) / 32;
  const int64_t npos = nsl * 32;
  std::vector<int> blen((size_t)npos, 0);
  int64_t sum = 0;
#pragma omp parallel for schedule(static) reduction(+ : sum)
  for (int64_t b = 0; b < nb; ++b) {
    blen[(size_t)b] = (rowptr[b * BS + 1] - rowptr[b * BS]) / BS;
    sum += blen[(size_t)b];
  }
  auto padded_total = [&](const std::vector<int> &len) {
    int64_t tot = 0;
#pragma omp parallel for schedule(static) reduction(+ : tot)
    for (int64_t sl = 0; sl < nsl; ++sl) {
      int w = 0;
      for (int l = 0; l < 32; ++l) w = std::max(w, len[(size_t)(sl * 32 + l)]);
      tot += w;
    }
    return tot * 32;
  };
  const int64_t tot_unsorted = padded_total(blen);
  bool sorted = false;
  std::vector<int> perm;
  if (sort_opt == "1" || (sort_opt == "auto" && (double)tot_unsorted > 1.03 * (double)sum + 1024.0)) {
    perm.assign((size_t)npos, -1);
    std::vector<int> slen((size_t)npos, 0);
    const int64_t nwin = (nb + SELL_WINDOW - 1) / SELL_WINDOW;
#pragma omp parallel for schedule(static)
    for (int64_t w = 0; w < nwin; ++w) {
      const int64_t lo = w * SELL_WINDOW, hi = std::min(nb, lo + SELL_WINDOW);
      int idx[SELL_WINDOW];
      for (int64_t i = lo; i < hi; ++i) idx[i - lo] = (int)i;
      std::stable_sort(idx, idx + (hi - lo), [&](int a, int b) { return blen[(size_t)a] > blen[(size_t)b]; });
      for (int64_t i = lo; i < hi; ++i) {
        perm[(size_t)i] = idx[i - lo];
        slen[(size_t)i] = blen[(size_t)idx[i - lo]];
      }
    }
    const int64_t tot_sorted = padded_total(slen);
    if (sort_opt == "1" || (double)tot_sorted < 0.97 * (double)tot_unsorted) {
      sorted = true;
      blen.swap(slen);
    } else {
      perm.clear();
    }
  }
  // ---- per-slice layout.  PACKED slice: lane stores its blocks k = 0..len-1, width = longest block row.
  // DIAGONAL-ALIGNED slice (unsorted matrices): the k-slots of the slice are the distinct diagonal offsets
  // d = (block column) - (slice*32 + lane) of its lanes in ascending order; a lane fills the slots whose offset
  // it has, in ascending order = its CSR order, the other slots are masked out.  Every (slice, k) pair of an
  // aligned slice is affine (block column = slice*32 + d_k + lane): a mesh-ordered stencil matrix stores one
  // word per 32 blocks instead of 32 ids, including the slices that cross mesh-line ends.  A slice is aligned
  // when it has at most 32 distinct offsets and aligning costs at most 25 % more slots than packing.
  std::vector<int> width((size_t)nsl, 0);
  std::vector<unsigned char> aligned((size_t)nsl, 0);
  std::vector<std::vector<int>> offs(dia && !sorted ? (size_t)nsl : 0);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t sl = 0; sl < nsl; ++sl) {
    int w = 0;
    for (int l = 0; l < 32; ++l) w = std::max(w, blen[(size_t)(sl * 32 + l)]);
    width[(size_t)sl] = w;
    if (!dia || sorted || w == 0) continue;
    int64_t d[32 * 48];
    int nd = 0;
    bool too_many = false;
    for (int l = 0; l < 32 && !too_many; ++l) {
      const int64_t b = sl * 32 + l;
      if (b >= nb) break;
      const int e0 = rowptr[b * BS], L = blen[(size_t)b];
      if (L > 32) { too_many = true; break; }
      for (int q = 0; q < L; ++q) d[nd++] = (int64_t)(col[(size_t)e0 + (size_t)q * BS] / BS) - b;
    }
    if (too_many) continue;
    std::sort(d, d + nd);
    nd = (int)(std::unique(d, d + nd) - d);
    if (nd > 32 || nd * 4 > w * 5) continue;
    aligned[(size_t)sl] = 1;
    width[(size_t)sl] = nd;
    offs[(size_t)sl].assign(d, d + nd);
  }
  int64_t tot = 0;
  for (int64_t sl = 0; sl < nsl; ++sl) tot += width[(size_t)sl];
  tot *= 32;
  P.bs = BS; P.sorted = sorted; P.n_brows = nb; P.n_slices = nsl; P.blocks = tot; P.sum_blocks = sum;
  // padding budget: 25 %; very short rows (prolongations: 1/2/4/8 entries) may pad up to 3x -- a padded
  // coalesced slice still beats the row-pointer-chasing CSR kernel there
  const double avg = (double)nnz / (double)n_rows;
  const double pad_ok = (avg <= 8.0) ? 3.0 : 1.25;
  if (tot / 32 >= INT32_MAX || (double)tot > pad_ok * (double)std::max<int64_t>(sum, 1) + 4096.0) return P;
  P.off.assign((size_t)nsl + 1, 0);
  for (int64_t sl = 0; sl < nsl; ++sl) P.off[(size_t)sl + 1] = P.off[(size_t)sl] + width[(size_t)sl];
  // column words + per-position validity word (lmask): width <= 32: bit k set <=> the lane has a block in slot
  // k; width > 32 (packed only): the number of blocks of the lane
  {
    P.kbase.assign((size_t)P.off[(size_t)nsl], -1);
    P.lmask.assign((size_t)npos, 0);
#pragma omp parallel for schedule(static)
    for (int64_t sl = 0; sl < nsl; ++sl) {
      const int w = width[(size_t)sl];
      int *kb = P.kbase.data() + P.off[(size_t)sl];
      if (aligned[(size_t)sl]) {
        const std::vector<int> &d = offs[(size_t)sl];
        for (int k = 0; k < w; ++k) kb[k] = (int)(sl * 32 + d[(size_t)k]);  // may be < 0: see the fix-up below
        for (int l = 0; l < 32; ++l) {
          const int64_t b = sl * 32 + l;
          if (b >= nb) break;
          const int e0 = rowptr[b * BS], L = blen[(size_t)b];
          unsigned m = 0;
          int k = 0;
          for (int q = 0; q < L; ++q) {
            const int dq = (int)((int64_t)(col[(size_t)e0 + (size_t)q * BS] / BS) - b);
            while (d[(size_t)k] != dq) ++k;
            m |= 1u << k;
          }
          P.lmask[(size_t)b] = (int)m;
        }
        continue;
      }
      int e0[32], ln[32];
      for (int l = 0; l < 32; ++l) {
        const int64_t pos = sl * 32 + l;
        const int64_t b = sorted ? perm[(size_t)pos] : (pos < nb ? pos : -1);
        ln[l] = b < 0 ? 0 : blen[(size_t)pos];
        e0[l] = b < 0 ? 0 : rowptr[b * BS];
        P.lmask[(size_t)pos] = w <= 32 ? (int)(ln[l] >= 32 ? 0xffffffffu : ((1u << ln[l]) - 1u)) : ln[l];
      }
      for (int k = 0; k < w && affine; ++k) {
        bool any = false, ok = true;
        int64_t base = 0;
        for (int l = 0; l < 32 && ok; ++l) {
          if (k >= ln[l]) continue;
          const int64_t c = col[(size_t)e0[l] + (size_t)k * BS] / BS;
          if (!any) { base = c - l; any = true; }
          else if (c - l != base) ok = false;
        }
        if (any && ok && base >= 0) kb[k] = (int)base;
      }
    }
    // a negative word means "explicit line": affine bases of aligned slices that are negative (first slices:
    // slot offsets reach before column 0 for the lanes that do not use them) are stored biased instead
    int64_t nexp = 0;
    for (int64_t sl = 0; sl < nsl; ++sl) {
      int *kb = P.kbase.data() + P.off[(size_t)sl];
      const int w = width[(size_t)sl];
      if (aligned[(size_t)sl]) {
        bool neg = false;
        for (int k = 0; k < w; ++k) neg = neg || kb[k] < 0;
        if (!neg) continue;
        // rare (only the first few slices): fall back to explicit lines for the negative slots
        for (int k = 0; k < w; ++k)
          if (kb[k] < 0) kb[k] = (int)~nexp++;
        continue;
      }
      for (int k = 0; k < w; ++k)
        if (kb[k] < 0) kb[k] = (int)~nexp++;
    }
    P.n_explicit = nexp;
    P.aligned_slices = 0;
    for (int64_t sl = 0; sl < nsl; ++sl) P.aligned_slices += aligned[(size_t)sl];
    // aligned slices whose slots come in runs of three consecutive columns (stencil x-neighbours) and have no
    // explicit line: the kernel gathers every run once and shuffles (kernels.cuh, sell_kernel)
    P.kind.assign((size_t)nsl, 0);
    if (BS == 1 && triples) {
      int64_t ntri = 0;
#pragma omp parallel for schedule(static) reduction(+ : ntri)
      for (int64_t sl = 0; sl < nsl; ++sl) {
        if (!aligned[(size_t)sl]) continue;
        const int w = width[(size_t)sl];
        const int *kb = P.kbase.data() + P.off[(size_t)sl];
        bool ok = w > 0 && w % 3 == 0;
        for (int k = 0; k < w && ok; k += 3) ok = kb[k] >= 0 && kb[k + 1] == kb[k] + 1 && kb[k + 2] == kb[k] + 2;
        if (ok) { P.kind[(size_t)sl] = 1; ntri += 1; }
      }
      P.triple_slices = ntri;
    }
  }
  // interior / boundary slice lists (only for matrices with ghost columns): a slice is "boundary" when one
  // of its block rows references a ghost column (columns ascend: the last one decides)
  if (n_ghost_cols > 0) {
    std::vector<unsigned char> bnd((size_t)nsl, 0);
#pragma omp parallel for schedule(static)
    for (int64_t sl = 0; sl < nsl; ++sl) {
      for (int l = 0; l < 32 && !bnd[(size_t)sl]; ++l) {
        const int64_t pos = sl * 32 + l;
        const int64_t b = sorted ? perm[(size_t)pos] : (pos < nb ? pos : -1);
        if (b < 0) continue;
        const int e0 = rowptr[b * BS], e1 = rowptr[b * BS + 1];
        if (e1 > e0 && col[(size_t)e1 - 1] >= n_own_cols) bnd[(size_t)sl] = 1;
      }
    }
    for (int64_t sl = 0; sl < nsl; ++sl) (bnd[(size_t)sl] ? P.bnd_slices : P.int_slices).push_back((int)sl);
  }
  P.perm.swap(perm);
  P.blen.swap(blen);
  P.ok = true;
  return P;
}

// fills A's block-SELL description and builds the arrays on the device
void build_sell(gsb_mat_s *A, const int *rowptr, const int *col) {
  gsb_ctx_t ctx = A->ctx;
  A->sell_ok = false;
  A->split_ok = false;
  if (A->n_rows == 0 || ctx->opt("sell", "1") != "1") return;
  SellPlan P = plan_sell(A->n_rows, A->n_own_cols, A->n_ghost_cols, rowptr, col, ctx->opt("block", "1") == "1",
                         ctx->opt("sell_sort", "auto"), ctx->opt("sell_affine", "1") == "1", ctx->opt("sell_align", "1") == "1", ctx->opt("sell_triples", "1") == "1");
  if (!P.ok) return;
  A->bs = P.bs;
  A->sorted = P.sorted;
  A->n_brows = P.n_brows;
  A->n_slices = P.n_slices;
  A->sell_blocks = P.blocks;
  A->sell_explicit = P.n_explicit;
  A->sell_aligned = P.aligned_slices;
  auto up = [&](DevBuf<int> &d, const std::vector<int> &h) {
    d.alloc(std::max<size_t>(1, h.size()));
    if (!h.empty()) GSB_CUDA(cudaMemcpyAsync(d.p, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
  };
  up(A->sell_lmask, P.lmask);
  up(A->sell_off, P.off);
  if (P.sorted) up(A->sell_perm, P.perm);
  up(A->sell_kbase, P.kbase);
  up(A->sell_kind, P.kind);
  A->sell_bcol.alloc((size_t)std::max<int64_t>(P.n_explicit * 32, 1));
  A->sell_val.alloc((size_t)std::max<int64_t>(P.blocks, 1) * P.bs * P.bs);
  sell_fill(A, /*values_only=*/false, A->val.p);
  A->sell_ok = true;
  if (A->n_ghost_cols > 0) {
    A->n_int_slices = (int64_t)P.int_slices.size();
    A->n_bnd_slices = (int64_t)P.bnd_slices.size();
    up(A->int_slices, P.int_slices);
    up(A->bnd_slices, P.bnd_slices);
    A->split_ok = true;
  }
  GSB_CUDA(cudaStreamSynchronize(ctx->stream));  // the plan's host arrays go out of scope here
}

// rowptr/col/val: CSR with int32 0-based ascending columns (host; may be the caller's own buffers)
void finish_matrix(gsb_mat_s *A, const int *rowptr, const int *col, const double *val) {
  gsb_ctx_t ctx = A->ctx;
  A->nnz = rowptr[A->n_rows];
  const size_t nnz_alloc = (size_t)std::max<int64_t>(A->nnz, 1);
  A->rowptr.alloc((size_t)A->n_rows + 1);
  A->col.alloc(nnz_alloc);
  A->val.alloc(nnz_alloc);
  // uploads go through the context's (non-blocking) stream: a plain cudaMemcpy from pageable memory returns once
  // the data is staged, its DMA is ordered in the legacy stream only and would race with the kernels below
  GSB_CUDA(cudaMemcpyAsync(A->rowptr.p, rowptr, sizeof(int) * ((size_t)A->n_rows + 1), cudaMemcpyHostToDevice, ctx->stream));
  if (A->nnz) {
    GSB_CUDA(cudaMemcpyAsync(A->col.p, col, sizeof(int) * (size_t)A->nnz, cudaMemcpyHostToDevice, ctx->stream));
    GSB_CUDA(cudaMemcpyAsync(A->val.p, val, sizeof(double) * (size_t)A->nnz, cudaMemcpyHostToDevice, ctx->stream));
  }
  A->csr_kept = true;
  int mx = 0;
#pragma omp parallel for schedule(static) reduction(max : mx)
  for (int64_t i = 0; i < A->n_rows; ++i) mx = std::max(mx, rowptr[i + 1] - rowptr[i]);
  A->max_row_nnz = mx;
  const double avg = A->n_rows ? (double)A->nnz / (double)A->n_rows : 0.0;
  A->G = avg <= 48.0 ? 1 : (avg <= 160.0 ? 4 : 16);
  A->diag.alloc((size_t)std::max<int64_t>(1, A->n_rows));
  csr_diag(A, A->val.p);
  build_sell(A, rowptr, col);
  // large matrices stream the block-SELL arrays only: release the CSR column ids / values
  const int64_t keep_max = std::stoll(ctx->opt("keep_csr_max_nnz", "16777216"));
  GSB_CUDA(cudaStreamSynchronize(ctx->stream));  // the caller's / the converter's host arrays may go away now
  if (A->sell_ok && A->nnz > keep_max && ctx->opt("keep_csr", "0") != "1") {
    A->col.release();
    A->val.release();
    A->csr_kept = false;
  }
}

}  // namespace

extern "C" {

int gsb_mat_create(gsb_ctx_t ctx, int64_t n_rows, int64_t n_own_cols, int64_t n_ghost_cols, int fmt, int index_base,
                   int index_bytes, const void *ptr, const void *idx, const double *vals, gsb_plan_t plan,
                   gsb_mat_t *out) {
  GSB_NULLCHK(ctx)
  API_BEGIN
  GSB_CHECK(index_bytes == 4 || index_bytes == 8, "mat: index_bytes must be 4 or 8");
  GSB_CHECK(fmt == GSB_FMT_CSR || fmt == GSB_FMT_CSC, "mat: unknown format");
  GSB_CHECK(n_ghost_cols == 0 || ctx->nranks == 1 || plan != nullptr, "mat: ghost columns need an exchange plan");
  const int64_t n_cols = n_own_cols + n_ghost_cols;
  GSB_CHECK(n_rows >= 0 && n_cols >= 0 && n_rows < INT32_MAX && n_cols < INT32_MAX, "mat: local dimensions exceed int32");
  std::unique_ptr<gsb_mat_s> A(new gsb_mat_s());
  A->ctx = ctx; A->n_rows = n_rows; A->n_own_cols = n_own_cols; A->n_ghost_cols = n_ghost_cols; A->plan = plan;
  const int64_t nptr = (fmt == GSB_FMT_CSR ? n_rows : n_cols);
  const int64_t p0 = rd_idx(ptr, index_bytes, 0);
  const int64_t nnz = rd_idx(ptr, index_bytes, nptr) - p0;
  GSB_CHECK(nnz >= 0 && nnz < INT32_MAX - 8, "mat: local nnz exceeds int32");
  if (fmt == GSB_FMT_CSR) {
    // validate (range, ascending columns) straight on the caller's arrays
    int bad_range = 0, unsorted = 0;
#pragma omp parallel for schedule(static) reduction(|| : bad_range, unsorted)
    for (int64_t i = 0; i < n_rows; ++i) {
      const int64_t a = rd_idx(ptr, index_bytes, i) - p0, b = rd_idx(ptr, index_bytes, i + 1) - p0;
      if (a > b || b > nnz) { bad_range = 1; continue; }
      int64_t prev = -1;
      for (int64_t e = a; e < b; ++e) {
        const int64_t c = rd_idx(idx, index_bytes, e) - index_base;
        if (c < 0 || c >= n_cols) bad_range = 1;
        if (c <= prev) unsorted = 1;
        prev = c;
      }
    }
    GSB_CHECK(!bad_range, "mat: column index out of range");
    if (!unsorted && index_bytes == 4 && index_base == 0 && p0 == 0) {
      // zero-copy: the caller's int32 0-based sorted CSR is uploaded as it is
      finish_matrix(A.get(), (const int *)ptr, (const int *)idx, vals);
    } else {
      std::vector<int> rowptr((size_t)n_rows + 1), col((size_t)nnz);
      std::vector<double> val;
      const double *vsrc = vals;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i <= n_rows; ++i) rowptr[(size_t)i] = (int)(rd_idx(ptr, index_bytes, i) - p0);
#pragma omp parallel for schedule(static)
      for (int64_t e = 0; e < nnz; ++e) col[(size_t)e] = (int)(rd_idx(idx, index_bytes, e) - index_base);
      if (unsorted) {
        A->perm.resize((size_t)nnz);
        val.resize((size_t)nnz);
        std::vector<int> c2((size_t)nnz);
        int dup = 0;
#pragma omp parallel for schedule(dynamic, 1024)
        for (int64_t i = 0; i < n_rows; ++i) {
          const int a = rowptr[(size_t)i], b = rowptr[(size_t)i + 1];
          for (int e = a; e < b; ++e) A->perm[(size_t)e] = e;
          std::sort(A->perm.begin() + a, A->perm.begin() + b, [&](int64_t x, int64_t y) { return col[(size_t)x] < col[(size_t)y]; });
          for (int e = a; e < b; ++e) {
            c2[(size_t)e] = col[(size_t)A->perm[(size_t)e]];
            val[(size_t)e] = vals[A->perm[(size_t)e]];
          }
          for (int e = a + 1; e < b; ++e)
            if (c2[(size_t)e] == c2[(size_t)e - 1]) {
#pragma omp atomic write
              dup = 1;
            }
        }
        GSB_CHECK(!dup, "mat: duplicate column index in a row");
        col.swap(c2);
        vsrc = val.data();
      }
      finish_matrix(A.get(), rowptr.data(), col.data(), vsrc);
    }
  } else {  // CSC -> CSR; visiting columns in ascending order leaves every row sorted
    std::vector<int> rowptr((size_t)n_rows + 1, 0), col((size_t)nnz);
    std::vector<double> val((size_t)nnz);
    A->perm.resize((size_t)nnz);
    for (int64_t e = 0; e < nnz; ++e) {
      int64_t r = rd_idx(idx, index_bytes, e) - index_base;
      GSB_CHECK(r >= 0 && r < n_rows, "mat: row index out of range");
      rowptr[(size_t)r + 1]++;
    }
    for (int64_t i = 0; i < n_rows; ++i) rowptr[(size_t)i + 1] += rowptr[(size_t)i];
    std::vector<int> fillp(rowptr.begin(), rowptr.end() - 1);
    for (int64_t j = 0; j < n_cols; ++j) {
      const int64_t a = rd_idx(ptr, index_bytes, j) - p0, b = rd_idx(ptr, index_bytes, j + 1) - p0;
      for (int64_t e = a; e < b; ++e) {
        const int64_t r = rd_idx(idx, index_bytes, e) - index_base;
        const int pos = fillp[(size_t)r]++;
        col[(size_t)pos] = (int)j;
        val[(size_t)pos] = vals[e];
        A->perm[(size_t)pos] = e;
      }
    }
    finish_matrix(A.get(), rowptr.data(), col.data(), val.data());
  }
  *out = A.release();
  API_END(ctx)
}

int gsb_mat_update_values(gsb_mat_t A, const double *vals) {
  GSB_NULLCHK(A)
  API_BEGIN
  GSB_CHECK(A->nb == 0, "update_values: block matrix");
  gsb_ctx_t ctx = A->ctx;
  const double *src = vals;
  std::vector<double> v;
  if (!A->perm.empty()) {
    v.resize((size_t)A->nnz);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < A->nnz; ++e) v[(size_t)e] = vals[A->perm[(size_t)e]];
    src = v.data();
  }
  // CSR-ordered values on the device: in place when the CSR mirror is kept, else a temporary
  DevBuf<double> tmp;
  double *dval = A->val.p;
  if (!A->csr_kept) {
    tmp.alloc((size_t)std::max<int64_t>(A->nnz, 1));
    dval = tmp.p;
  }
  if (A->nnz) GSB_CUDA(cudaMemcpyAsync(dval, src, sizeof(double) * (size_t)A->nnz, cudaMemcpyHostToDevice, ctx->stream));
  refresh_diag(A, dval);  // through the positions of the diagonal entries recorded at creation
  if (A->sell_ok) sell_fill(A, /*values_only=*/true, dval);
  GSB_CUDA(cudaStreamSynchronize(ctx->stream));
  API_END(A->ctx)
}

int gsb_mat_info(gsb_mat_t A, int64_t *n_rows, int64_t *n_own_cols, int64_t *n_ghost_cols, int64_t *nnz) {
  GSB_NULLCHK(A)
  if (n_rows) *n_rows = A->n_rows;
  if (n_own_cols) *n_own_cols = A->n_own_cols;
  if (n_ghost_cols) *n_ghost_cols = A->n_ghost_cols;
  if (nnz) *nnz = A->nnz;
  return GSB_OK;
}

// storage the row kernels stream: kind 0 = CSR (fallback kernel), 1 = block-SELL-32
int gsb_mat_format(gsb_mat_t A, int *kind, int *block_size, int *sorted, int64_t *stored_entries, int64_t *bytes_per_pass) {
  GSB_NULLCHK(A)
  if (kind) *kind = A->sell_ok ? 1 : 0;
  if (block_size) *block_size = A->sell_ok ? A->bs : 1;
  if (sorted) *sorted = A->sell_ok && A->sorted ? 1 : 0;
  if (stored_entries) *stored_entries = A->sell_ok ? A->sell_blocks * A->bs * A->bs : A->nnz;
  if (bytes_per_pass) *bytes_per_pass = A->format_bytes();
  return GSB_OK;
}

int gsb_mat_destroy(gsb_mat_t A) {
  delete A;
  return GSB_OK;
}

// diagnostics (pure host, no device needed): the block-SELL plan the library would build for a CSR matrix
// (int32, 0-based, ascending columns).  out[12] = {ok, block size, sorted, block rows, slices, stored blocks incl.
// padding, blocks without padding, boundary slices, explicit (slice,k) id lines, (slice,k) pairs, diagonal-aligned
// slices, slices made of runs of three consecutive columns}; pos_row (n_slices*32 ints or NULL) receives the block row of every (slice, lane) position (-1 =
// padding lane), pos_len its length in blocks, pos_mask its slot-validity word, col_words (out[9] ints or NULL) the
// column word of every (slice, k) pair (>= 0: affine base)
int gsb_diag_sell_plan(int64_t n_rows, int64_t n_own_cols, int64_t n_ghost_cols, const int *rowptr, const int *col,
                       int detect_blocks, int sort_mode, int64_t *out, int *pos_row, int *pos_len, int *pos_mask, int *col_words) {
  API_BEGIN
  GSB_CHECK(rowptr && col && out && n_rows >= 0, "sell plan: bad arguments");
  SellPlan P = plan_sell(n_rows, n_own_cols, n_ghost_cols, rowptr, col, detect_blocks != 0,
                         sort_mode < 0 ? "auto" : (sort_mode ? "1" : "0"));
  out[0] = P.ok; out[1] = P.bs; out[2] = P.sorted; out[3] = P.n_brows; out[4] = P.n_slices; out[5] = P.blocks;
  out[6] = P.sum_blocks; out[7] = (int64_t)P.bnd_slices.size(); out[8] = P.n_explicit; out[9] = (int64_t)P.kbase.size();
  out[10] = P.aligned_slices; out[11] = P.triple_slices;
  if (P.ok) {
    if (col_words) std::copy(P.kbase.begin(), P.kbase.end(), col_words);
    for (int64_t pos = 0; pos < P.n_slices * 32; ++pos) {
      if (pos_row) pos_row[pos] = P.sorted ? P.perm[(size_t)pos] : (pos < P.n_brows ? (int)pos : -1);
      if (pos_len) pos_len[pos] = P.blen[(size_t)pos];
      if (pos_mask) pos_mask[pos] = P.lmask[(size_t)pos];
    }
  }
  API_END(nullptr)
}

int gsb_block_mat_create(gsb_ctx_t ctx, int nb, const gsb_mat_t *blocks, gsb_mat_t *out) {
  GSB_NULLCHK(ctx)
  API_BEGIN
  GSB_CHECK(nb >= 1, "block matrix: nb < 1");
  std::unique_ptr<gsb_mat_s> A(new gsb_mat_s());
  A->ctx = ctx; A->nb = nb;
  A->blocks.assign(blocks, blocks + (size_t)nb * nb);
  std::vector<int64_t> rs((size_t)nb, -1), cs((size_t)nb, -1);
  for (int i = 0; i < nb; ++i)
    for (int j = 0; j < nb; ++j) {
      gsb_mat_t B = A->blocks[(size_t)i * nb + j];
      if (!B) continue;
      GSB_CHECK(B->n_ghost_cols == 0, "block matrix: distributed blocks not supported");
      GSB_CHECK(rs[(size_t)i] < 0 || rs[(size_t)i] == B->n_rows, "block matrix: inconsistent block rows");
      GSB_CHECK(cs[(size_t)j] < 0 || cs[(size_t)j] == B->n_own_cols, "block matrix: inconsistent block cols");
      rs[(size_t)i] = B->n_rows; cs[(size_t)j] = B->n_own_cols;
      A->nnz += B->nnz;
    }
  A->row_off.assign((size_t)nb + 1, 0); A->col_off.assign((size_t)nb + 1, 0);
  for (int i = 0; i < nb; ++i) {
    GSB_CHECK(rs[(size_t)i] >= 0 && cs[(size_t)i] >= 0, "block matrix: empty block row/column");
    A->row_off[(size_t)i + 1] = A->row_off[(size_t)i] + rs[(size_t)i];
    A->col_off[(size_t)i + 1] = A->col_off[(size_t)i] + cs[(size_t)i];
  }
  A->n_rows = A->row_off.back(); A->n_own_cols = A->col_off.back();
  *out = A.release();
  API_END(ctx)
}

}  // extern "C"
