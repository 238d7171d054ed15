// kernels.cuh -- hand-written sm_100a kernels of the gsb200 solve phase.
//
// Every kernel here is HBM-bandwidth bound: the design rules are a matrix stream made of fully
// coalesced 256 B lines whose addresses do not depend on row pointers (block-SELL-32, one lane per block row),
// as few bytes per non-zero as the structure allows (column ids shared by DOF blocks and compressed to one word
// per 32 blocks where the columns of a slice are consecutive), coalesced gathers of the input vector (the k-th
// neighbours of 32 consecutive rows are 32 consecutive entries on a mesh-ordered matrix; runs of consecutive
// columns are gathered once and shuffled), fusion of every elementwise epilogue into the row kernel, and
// deterministic two-stage reductions.
//
// Rounding contract (DESIGN.md "parity"): a row sum is accumulated in ascending column order,
// product rounded before the add (no FMA contraction) -- the sequence Julia's SparseArrays /
// SparseMatricesCSR mul! produce (oracle/csr_kernels.c) -- so with one lane per row the SpMV,
// smoother and transfer kernels are bit-identical to the CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "types.h"

namespace gsb {

// ------------------------------------------------------------------------------------------
// device scalars: Krylov coefficients live in a small device array; kernels take references
// to them so that the host never has to read gamma/alpha/beta back (one sync per iteration,
// for the stopping test only).
__device__ __forceinline__ double load_scalar(const ScalarRef &r, const double *__restrict__ scal) {
  double s = r.v;
  if (r.num >= 0) {
    s = scal[r.num];
    if (r.sub >= 0) s = __dsub_rn(s, scal[r.sub]);
    if (r.den >= 0) s = __ddiv_rn(s, scal[r.den]);
  }
  return r.neg ? -s : s;
}

// ------------------------------------------------------------------------------------------
// deterministic block reduction + "last block finalises" grid reduction
template <int THREADS>
__device__ __forceinline__ double block_reduce_sum(double v, double *smem /* THREADS/32 doubles */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) smem[w] = v;
  __syncthreads();
  double t = 0.0;
  if (w == 0) {
    t = (l < THREADS / 32) ? smem[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in warp 0 (all lanes)
}

// every block calls this with its partial(s); the last block to arrive sums all partials in a
// fixed order and stores the totals => bitwise reproducible for a fixed grid size.
template <int THREADS, int NRED>
__device__ __forceinline__ void grid_reduce_finish(double (&v)[NRED], const ReduceOut &ro, double *smem) {
  __shared__ bool is_last;
#pragma unroll
  for (int k = 0; k < NRED; ++k) {
    double t = block_reduce_sum<THREADS>(v[k], smem);
    if (threadIdx.x == 0) ro.partials[(size_t)k * gridDim.x + blockIdx.x] = t;
  }
  __threadfence();
  if (threadIdx.x == 0) {
    unsigned int prev = atomicAdd(ro.ticket, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
#pragma unroll
    for (int k = 0; k < NRED; ++k) {
      double t = 0.0;
      const volatile double *p = ro.partials + (size_t)k * gridDim.x;
      for (unsigned int i = threadIdx.x; i < gridDim.x; i += THREADS) t += p[i];
      t = block_reduce_sum<THREADS>(t, smem);
      if (threadIdx.x == 0) ro.scal[ro.slot[k]] = t;
    }
    if (threadIdx.x == 0) *ro.ticket = 0u;
  }
}

// ------------------------------------------------------------------------------------------
// row epilogues shared by the CSR kernels
template <int MODE>
__device__ __forceinline__ double row_init(const RowArgs &a, int64_t row) {
  if (MODE == ROW_SPMV) {
    if (a.beta == 0.0) return 0.0;
    if (a.beta == 1.0) return a.y[row];
    return __dmul_rn(a.beta, a.y[row]);
  }
  return 0.0;
}

template <int MODE>
__device__ __forceinline__ void row_epilogue(const RowArgs &a, int64_t row, double s, double &acc) {
  if (MODE == ROW_SPMV) {
    a.y[row] = s;
  } else if (MODE == ROW_RESID) {
    a.out[row] = __dsub_rn(a.b[row], s);
  } else if (MODE == ROW_SWEEP) {
    const double r = __dsub_rn(a.b[row], s);
    a.out[row] = r;
    const double d = __dmul_rn(a.omega, __dmul_rn(a.invd[row], r));
    a.dxout[row] = d;
    a.xacc[row] = __dadd_rn(a.xacc[row], d);
  } else if (MODE == ROW_SPMV_DOT) {
    a.y[row] = s;
    acc = __dadd_rn(acc, __dmul_rn(a.dotv[row], s));
  } else if (MODE == ROW_SPMV_ADD) {
    a.y[row] = s;
    a.xacc[row] = __dadd_rn(a.xacc[row], s);
  }
}

// ------------------------------------------------------------------------------------------
// generic CSR kernel: G lanes per row, direct global loads.  Used for small levels (launch /
// latency bound) and for matrices whose block-SELL padding would exceed the budget (matrix.cu plan_sell).
template <int G, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) csr_vector_kernel(int64_t nrows, const int *__restrict__ rowptr,
                                                            const int *__restrict__ col,
                                                            const double *__restrict__ val, RowArgs a) {
  __shared__ double red_smem[THREADS / 32];
  const int64_t row = ((int64_t)blockIdx.x * THREADS + threadIdx.x) / G;
  const int lane = threadIdx.x % G;
  double acc = 0.0;
  double s = 0.0;
  const bool valid = row < nrows;
  if (valid) {
    const int e0 = rowptr[row], e1 = rowptr[row + 1];
    const double al = a.alpha;
    if (lane == 0) s = row_init<MODE>(a, row);
    for (int e = e0 + lane; e < e1; e += G) {
      double xv = __ldg(a.x + col[e]);
      if (MODE == ROW_SPMV) xv = __dmul_rn(xv, al);
      s = __dadd_rn(s, __dmul_rn(val[e], xv));
    }
  }
  if (G > 1) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
  }
  if (valid && lane == 0) row_epilogue<MODE>(a, row, s, acc);
  if (MODE == ROW_SPMV_DOT) {
    double v[1] = {acc};
    grid_reduce_finish<THREADS, 1>(v, a.red, red_smem);
  }
}

template <int MODE>
struct RowPre {  // epilogue operands, loaded before the gather loop so their latency is hidden
  double b, invd, xacc, dotv;
};
template <int MODE>
__device__ __forceinline__ void row_prefetch(const RowArgs &a, int64_t row, RowPre<MODE> &p) {
  if (MODE == ROW_RESID) p.b = a.b[row];
  if (MODE == ROW_SWEEP) { p.b = a.b[row]; p.invd = a.invd[row]; p.xacc = a.xacc[row]; }
  if (MODE == ROW_SPMV_DOT) p.dotv = a.dotv[row];
  if (MODE == ROW_SPMV_ADD) p.xacc = a.xacc[row];
}
template <int MODE>
__device__ __forceinline__ void row_epilogue_pre(const RowArgs &a, int64_t row, double s, const RowPre<MODE> &p, double &acc) {
  if (MODE == ROW_SPMV) {
    a.y[row] = s;
  } else if (MODE == ROW_RESID) {
    a.out[row] = __dsub_rn(p.b, s);
  } else if (MODE == ROW_SWEEP) {
    const double r = __dsub_rn(p.b, s);
    a.out[row] = r;
    const double d = __dmul_rn(a.omega, __dmul_rn(p.invd, r));
    a.dxout[row] = d;
    a.xacc[row] = __dadd_rn(p.xacc, d);
  } else if (MODE == ROW_SPMV_DOT) {
    a.y[row] = s;
    acc = __dadd_rn(acc, __dmul_rn(p.dotv, s));
  } else if (MODE == ROW_SPMV_ADD) {
    a.y[row] = s;
    a.xacc[row] = __dadd_rn(p.xacc, s);
  }
}

// ------------------------------------------------------------------------------------------
// Block-SELL-32 row kernel (the hot kernel of every configuration).
//
// Storage (built on the device from the uploaded CSR arrays, sell_fill_kernel below):
//   * rows are grouped in BLOCK ROWS of BS consecutive rows (BS = 1, 2, 3: the DOF block of a
//     node-major vector-valued FE space; BS = 1 for scalar problems and for any matrix whose sparsity is
//     not made of aligned BS x BS blocks).  One lane owns one block row, one warp owns one SLICE of 32
//     block rows, the blocks of a slice are stored column-major: block k of the 32 lanes is contiguous
//     (one 128 B line of block-column ids, BS*BS 256 B lines of values) -- every matrix load of a warp is
//     a fully coalesced transaction whose address does not depend on a row pointer.
//   * slice width = longest block row of the slice.  When consecutive rows have very different lengths
//     (Q2 elements: 125/75/45/27-node stencils alternate along a mesh line) the block rows are sorted by
//     length inside windows of 256 block rows (SELL-C-sigma with C = 32, sigma = 256) and `perm` maps a
//     (slice, lane) position to its block row; padding lanes have perm = -1.
//   * block-column ids are compressed per (slice, k): when the k-th blocks of the 32 lanes of a slice sit in
//     CONSECUTIVE block columns (base + lane: the rule on a mesh-ordered matrix -- the k-th neighbours of 32
//     consecutive rows are 32 consecutive columns) only `base` is stored (kbase[] >= 0, 4 B per 32 blocks);
//     otherwise kbase[] = ~e and the 32 ids are line e of bcol[] (explicit).  On the C2 fine level 83 % of the
//     (slice, k) pairs are affine: 8.8 B per stored non-zero instead of 12.
//   * bytes per stored non-zero: 8 + (0.125 .. 4.125)/BS^2 instead of CSR's 12.
// Arithmetic: lane-sequential, ascending column order inside every row, product rounded before the add
// (__dmul_rn/__dadd_rn) -- row i of a block row visits block k = 0,1,.. and inside a block column
// j = 0..BS-1, i.e. exactly the CSR order of that row => bit-identical to the oracle's sequential CSR
// loop for every BS and with or without the permutation.  Matrix data is streamed with L1::no_allocate
// so that L1 keeps the gathered vector.
__device__ __forceinline__ double ldg_stream_f64(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ldg_stream_s32(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ldg_nc_f64(const double *p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ldg_nc_s32(const int *p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// block column of this lane for a (slice, k) pair whose kbase word is kb (warp-uniform => no divergence)
__device__ __forceinline__ int sell_bcol_of(int kb, int lane, const int *__restrict__ bcol) {
  return kb >= 0 ? kb + lane : ldg_stream_s32(bcol + (((size_t)(~kb)) << 5) + lane);
}

struct SellArgs {
  const int *slice_list;  // optional indirection: the slices this launch processes (nullptr = all)
  int64_t n_list;         // number of slices this launch processes
  const int *perm;        // PERM only: block row of every (slice, lane) position, -1 = padding lane
  const int *lmask;       // per position: slice width <= 32: bit k set <=> the lane has a block in slot k; else its length
  const int *slice_off;   // per slice: offset of the slice in units of 32 blocks; nslices+1 entries
  const int *kbase;       // per (slice, k): >= 0 affine base block column (lane l reads base + l), < 0: ~line of bcol
  const int *bcol;        // explicit block-column ids, one 32-lane line per non-affine (slice, k)
  const double *val;      // BS*BS values per block, each (k, i, j) a 32-lane line
  int64_t n_brows;        // number of block rows
  const int *kind;        // per slice: 1 = diagonal-aligned slice whose slots come in runs of three consecutive columns
  int64_t n_cols;         // entries of the gathered vector (own + ghost columns)
};

template <int MODE, int BS, bool PERM, int THREADS, int U, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) sell_kernel(SellArgs m, RowArgs a) {
  __shared__ double red_smem[THREADS / 32];
  constexpr int BB = BS * BS;
  const int64_t widx = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;  // one warp per slice
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  if (widx < m.n_list) {  // warp-uniform
    const int64_t slice = m.slice_list ? (int64_t)m.slice_list[widx] : widx;
    const int64_t pos = (slice << 5) + lane;
    int64_t brow = pos;
    bool valid;
    if (PERM) {
      const int p = m.perm[pos];
      brow = p;
      valid = p >= 0;
    } else {
      valid = pos < m.n_brows;
    }
    const int lm = m.lmask[pos];
    const int so0 = m.slice_off[slice], so1 = m.slice_off[slice + 1];
    const int width = so1 - so0;  // slots per lane in this slice (warp-uniform)
    const bool wide = width > 32;
    auto slot_on = [&](int k) { return wide ? (k < lm) : (((unsigned)lm >> k) & 1u) != 0u; };
    RowPre<MODE> pre{};
    double s[BS];
#pragma unroll
    for (int i = 0; i < BS; ++i) s[i] = 0.0;
    if (valid) {
      if (BS == 1) row_prefetch<MODE>(a, brow, pre);
#pragma unroll
      for (int i = 0; i < BS; ++i) s[i] = row_init<MODE>(a, brow * BS + i);
    }
    const int *kp = m.kbase + so0;
    const double *vp = m.val + (((size_t)so0 * BB) << 5) + lane;
    const double al = a.alpha;
    int k = 0;
    if (BS == 1 && !PERM && m.kind[slice] == 1) {
      // Slots in runs of three consecutive columns (c, c+1, c+2: the x-neighbours of a stencil).  The three
      // gathers of a run read the 34 entries x[c .. c+33]: ONE full-width gather (lane l: x[c+l]) plus one
      // two-lane gather (x[c+32], x[c+33]); the shifted operands come from warp shuffles.  A third of the L1
      // requests of the generic loop; the values, their order and every rounding are unchanged.
      const int nt = width / 3;
      auto run3 = [&](int t, const double (&v)[3], double M, double E) {
        const double d1 = __shfl_down_sync(0xffffffffu, M, 1), d2 = __shfl_down_sync(0xffffffffu, M, 2);
        const double e0 = __shfl_sync(0xffffffffu, E, 0), e1 = __shfl_sync(0xffffffffu, E, 1);
        double xs[3];
        xs[0] = M;
        xs[1] = lane == 31 ? e0 : d1;
        xs[2] = lane == 30 ? e0 : (lane == 31 ? e1 : d2);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (slot_on(3 * t + j)) {
            const double xj = (MODE == ROW_SPMV) ? __dmul_rn(xs[j], al) : xs[j];
            s[0] = __dadd_rn(s[0], __dmul_rn(v[j], xj));
          }
        }
      };
      auto gather3 = [&](int kb, double &M, double &E) {
        const int c = kb + lane;
        M = ((unsigned)c < (unsigned)m.n_cols) ? ldg_nc_f64(a.x + c) : 0.0;
        E = (lane < 2 && (unsigned)(c + 32) < (unsigned)m.n_cols) ? ldg_nc_f64(a.x + c + 32) : 0.0;
      };
      int t = 0;
      for (; t + 3 <= nt; t += 3) {
        int kb[3];
        double vv[3][3], M[3], E[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) kb[q] = ldg_nc_s32(kp + 3 * (t + q));
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int j = 0; j < 3; ++j) vv[q][j] = ldg_stream_f64(vp + (size_t)(3 * (t + q) + j) * 32);
#pragma unroll
        for (int q = 0; q < 3; ++q) gather3(kb[q], M[q], E[q]);
#pragma unroll
        for (int q = 0; q < 3; ++q) run3(t + q, vv[q], M[q], E[q]);
      }
      for (; t < nt; ++t) {
        double vv[3], M, E;
        const int kb = ldg_nc_s32(kp + 3 * t);
#pragma unroll
        for (int j = 0; j < 3; ++j) vv[j] = ldg_stream_f64(vp + (size_t)(3 * t + j) * 32);
        gather3(kb, M, E);
        run3(t, vv, M, E);
      }
      k = width;  // nothing left for the generic loops
    }
    for (; k + U <= width; k += U) {
      int cc[U];
      double vv[U][BB], xv[U][BS];
      // burst order: all column words, then all values (contiguous bursts per warp), then the gathers
#pragma unroll
      for (int u = 0; u < U; ++u) cc[u] = ldg_nc_s32(kp + k + u);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int e = 0; e < BB; ++e) vv[u][e] = ldg_stream_f64(vp + ((size_t)(k + u) * BB + e) * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) cc[u] = sell_bcol_of(cc[u], lane, m.bcol);
      bool on[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        on[u] = slot_on(k + u);  // masked slots gather nothing (their column word may point outside x)
#pragma unroll
        for (int j = 0; j < BS; ++j) xv[u][j] = on[u] ? ldg_nc_f64(a.x + (int64_t)cc[u] * BS + j) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (on[u]) {
          double t[BS];
#pragma unroll
          for (int j = 0; j < BS; ++j) t[j] = (MODE == ROW_SPMV) ? __dmul_rn(xv[u][j], al) : xv[u][j];
#pragma unroll
          for (int i = 0; i < BS; ++i)
#pragma unroll
            for (int j = 0; j < BS; ++j) s[i] = __dadd_rn(s[i], __dmul_rn(vv[u][i * BS + j], t[j]));
        }
      }
    }
    for (; k < width; ++k) {
      const int c = sell_bcol_of(ldg_nc_s32(kp + k), lane, m.bcol);
      double vv[BB], xv[BS];
#pragma unroll
      for (int e = 0; e < BB; ++e) vv[e] = ldg_stream_f64(vp + ((size_t)k * BB + e) * 32);
      const bool on = slot_on(k);
#pragma unroll
      for (int j = 0; j < BS; ++j) xv[j] = on ? ldg_nc_f64(a.x + (int64_t)c * BS + j) : 0.0;
      if (on) {
#pragma unroll
        for (int j = 0; j < BS; ++j)
          if (MODE == ROW_SPMV) xv[j] = __dmul_rn(xv[j], al);
#pragma unroll
        for (int i = 0; i < BS; ++i)
#pragma unroll
          for (int j = 0; j < BS; ++j) s[i] = __dadd_rn(s[i], __dmul_rn(vv[i * BS + j], xv[j]));
      }
    }
    if (valid) {
      if (BS == 1) {
        row_epilogue_pre<MODE>(a, brow, s[0], pre, acc);
      } else {
#pragma unroll
        for (int i = 0; i < BS; ++i) row_epilogue<MODE>(a, brow * BS + i, s[i], acc);
      }
    }
  }
  if (MODE == ROW_SPMV_DOT) {
    double v[1] = {acc};
    grid_reduce_finish<THREADS, 1>(v, a.red, red_smem);
  }
}

// CSR -> block-SELL conversion on the device: one lane per (slice, lane) position copies its block row
// into the slice (coalesced writes; padding written as zero blocks; explicit id lines get block column 0
// in their padding lanes, affine (slice, k) pairs store no ids at all -- kbase[] comes from the host plan).
// values_only: refresh of the values after gsb_mat_update_values (same sparsity).
template <int BS>
__global__ void __launch_bounds__(256) sell_fill_kernel(int64_t npos, int64_t n_brows, const int *__restrict__ perm,
                                                        const int *__restrict__ lmask, const int *__restrict__ slice_off,
                                                        const int *__restrict__ kbase,
                                                        const int *__restrict__ rowptr, const int *__restrict__ col,
                                                        const double *__restrict__ val, int *__restrict__ bcol,
                                                        double *__restrict__ sval, int values_only) {
  constexpr int BB = BS * BS;
  const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= npos) return;
  const int64_t slice = pos >> 5;
  const int lane = (int)(pos & 31);
  const int64_t brow = perm ? (int64_t)perm[pos] : (pos < n_brows ? pos : -1);
  const int so0 = slice_off[slice], width = slice_off[slice + 1] - so0;
  int e0[BS] = {};
  if (brow >= 0) {
#pragma unroll
    for (int i = 0; i < BS; ++i) e0[i] = rowptr[brow * BS + i];
  }
  const int lm = brow >= 0 ? lmask[pos] : 0;
  const bool wide = width > 32;
  double *vp = sval + (((size_t)so0 * BB) << 5) + lane;
  int q = 0;  // next block of the block row (slots are filled in ascending column order)
  for (int k = 0; k < width; ++k) {
    const bool in = wide ? (k < lm) : (((unsigned)lm >> k) & 1u) != 0u;
    if (!values_only) {
      const int kb = kbase[so0 + k];
      if (kb < 0) bcol[(((size_t)(~kb)) << 5) + lane] = in ? col[e0[0] + q * BS] / BS : 0;
    }
#pragma unroll
    for (int i = 0; i < BS; ++i)
#pragma unroll
      for (int j = 0; j < BS; ++j) vp[((size_t)k * BB + i * BS + j) * 32] = in ? val[e0[i] + q * BS + j] : 0.0;
    if (in) ++q;
  }
}

// diag(A) of the own-own block (0 for a missing entry, like diag() of a Julia sparse matrix)
__global__ void csr_diag_kernel(int64_t nrows, const int *__restrict__ rowptr, const int *__restrict__ col,
                                const double *__restrict__ val, double *__restrict__ diag, int *__restrict__ diag_pos) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  double d = 0.0;
  int pos = -1;
  for (int e = rowptr[i]; e < rowptr[i + 1]; ++e)
    if (col[e] == i) { d = val[e]; pos = e; }
  diag[i] = d;
  diag_pos[i] = pos;
}
__global__ void refresh_diag_kernel(int64_t nrows, const int *__restrict__ diag_pos, const double *__restrict__ val,
                                    double *__restrict__ diag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) diag[i] = diag_pos[i] >= 0 ? val[diag_pos[i]] : 0.0;
}

// ------------------------------------------------------------------------------------------
// BLAS-1: z = ((a*x + b*y) + c*w) / d   evaluated left to right, each product rounded before
// the add (matches Julia's broadcast of `x .+ s .* y`, `x .- s .* y`, `(z .- a2.*w .- a3.*wo) ./ a1`).
// a == 1 is exact (1*x == x), b*y with b = -s equals -(s*y) exactly.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) ew_kernel(EwArgs g) {
  const double a = load_scalar(g.a, g.scal), b = load_scalar(g.b, g.scal), c = load_scalar(g.c, g.scal),
               d = load_scalar(g.d, g.scal);
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < g.n; i += (int64_t)gridDim.x * THREADS) {
    double v;
    if (g.mul_xy) {
      v = __dmul_rn(g.x[i], g.y[i]);
    } else {
      v = __dmul_rn(a, g.x[i]);
      if (g.has_y) v = __dadd_rn(v, __dmul_rn(b, g.y[i]));
      if (g.has_w) v = __dadd_rn(v, __dmul_rn(c, g.w[i]));
      if (g.has_d) v = __ddiv_rn(v, d);
    }
    g.z[i] = v;
  }
}

__global__ void fill_kernel(int64_t n, double value, double *__restrict__ v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = value;
}

// Jacobi-Richardson prologue: dx = omega*(invd*r) ; x += dx      (RichardsonSmoothers.jl:91-93)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) jacobi_step_kernel(int64_t n, const double *__restrict__ invd,
                                                             const double *__restrict__ r, double omega,
                                                             double *__restrict__ dx, double *__restrict__ x,
                                                             int x_is_zero) {
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
    const double d = __dmul_rn(omega, __dmul_rn(invd[i], r[i]));
    dx[i] = d;
    x[i] = x_is_zero ? __dadd_rn(0.0, d) : __dadd_rn(x[i], d);
  }
}

// dot(a,b) over own values -> scal[slot]
template <int THREADS>
__global__ void __launch_bounds__(THREADS) dot_kernel(int64_t n, const double *__restrict__ a,
                                                     const double *__restrict__ b, ReduceOut ro) {
  __shared__ double red_smem[THREADS / 32];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS)
    acc = __dadd_rn(acc, __dmul_rn(a[i], b[i]));
  double v[1] = {acc};
  grid_reduce_finish<THREADS, 1>(v, ro, red_smem);
}

// z = invd .* r ; scal[slot] = dot(z, r)        (Jacobi Pl fused with CG's gamma, CGSolvers.jl:94-95)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) jacobi_dot_kernel(int64_t n, const double *__restrict__ invd,
                                                            const double *__restrict__ r, double *__restrict__ z,
                                                            ReduceOut ro) {
  __shared__ double red_smem[THREADS / 32];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
    const double ri = r[i];
    const double zi = __dmul_rn(invd[i], ri);
    z[i] = zi;
    acc = __dadd_rn(acc, __dmul_rn(zi, ri));
  }
  double v[1] = {acc};
  grid_reduce_finish<THREADS, 1>(v, ro, red_smem);
}

// CG tail: x += alpha p ; r -= alpha w ; scal[slot] = r.r      (CGSolvers.jl:105-111)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) cg_update_kernel(int64_t n, ScalarRef alpha, const double *__restrict__ p,
                                                           const double *__restrict__ w, double *__restrict__ x,
                                                           double *__restrict__ r, ReduceOut ro) {
  __shared__ double red_smem[THREADS / 32];
  const double al = load_scalar(alpha, ro.scal);
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
    x[i] = __dadd_rn(x[i], __dmul_rn(al, p[i]));
    const double ri = __dsub_rn(r[i], __dmul_rn(al, w[i]));
    r[i] = ri;
    acc = __dadd_rn(acc, __dmul_rn(ri, ri));
  }
  double v[1] = {acc};
  grid_reduce_finish<THREADS, 1>(v, ro, red_smem);
}

// modified Gram-Schmidt, one launch per step (GMRESSolvers.jl:162-165): w -= h_prev * vprev (the axpy of
// the previous step, h_prev read from its device slot) fused with the dot of the updated w against v
// (v == nullptr: against itself -> the norm that closes the column).  Element-wise the same roundings
// and the same grid-stride summation order as the separate ew_kernel / dot_kernel launches.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) mgs_step_kernel(int64_t n, double *__restrict__ w, const double *__restrict__ vprev,
                                                           int slot_prev, const double *__restrict__ v, ReduceOut ro) {
  __shared__ double red_smem[THREADS / 32];
  const double b = vprev ? -ro.scal[slot_prev] : 0.0;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
    double wi = w[i];
    if (vprev) {
      wi = __dadd_rn(__dmul_rn(1.0, wi), __dmul_rn(b, vprev[i]));
      w[i] = wi;
    }
    acc = __dadd_rn(acc, __dmul_rn(wi, v ? v[i] : wi));
  }
  double r[1] = {acc};
  grid_reduce_finish<THREADS, 1>(r, ro, red_smem);
}

// x += sum_i g_i z_i, applied vector after vector per element (the rounding sequence of the loop of
// `x .+= g[i] .* Z[i]` broadcasts, GMRESSolvers.jl:193-196 / FGMRESSolvers.jl:191-193), one pass over x
constexpr int MAXPY_MAX = 16;
struct MultiAxpyArgs {
  double *x;
  const double *z[MAXPY_MAX];
  double g[MAXPY_MAX];
  int cnt;
  int64_t n;
};
template <int THREADS>
__global__ void __launch_bounds__(THREADS) multi_axpy_kernel(MultiAxpyArgs a) {
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * THREADS) {
    double xi = a.x[i];
    double zv[MAXPY_MAX];
#pragma unroll
    for (int q = 0; q < MAXPY_MAX; ++q)
      if (q < a.cnt) zv[q] = a.z[q][i];
#pragma unroll
    for (int q = 0; q < MAXPY_MAX; ++q)
      if (q < a.cnt) xi = __dadd_rn(__dmul_rn(1.0, xi), __dmul_rn(a.g[q], zv[q]));
    a.x[i] = xi;
  }
}

// gather / scatter for halo exchange
__global__ void pack_kernel(int64_t n, const int *__restrict__ ids, const double *__restrict__ v, double *__restrict__ buf) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) buf[i] = v[ids[i]];
}
__global__ void unpack_kernel(int64_t n, const int *__restrict__ ids, const double *__restrict__ buf, double *__restrict__ v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[ids[i]] = buf[i];
}

// ------------------------------------------------------------------------------------------
// halo exchange over NVLink peer memory (one process per GPU, buffers shared with CUDA IPC):
//   push kernel  : gathers the own entries each neighbour needs and stores them straight into that
//                  neighbour's receive buffer (peer addresses -> NVLink/NVSwitch writes); the last
//                  block to finish publishes this exchange's sequence number in every neighbour's
//                  flag word (system-scope release)
//   wait/unpack  : spins on the local flag words until every neighbour has published the sequence
//                  number (system-scope acquire), then scatters the receive buffer into the ghost tail
// Receive buffers are double-buffered by the parity of the sequence number: a sender can only run one
// exchange ahead of a receiver (it needs the receiver's data of exchange n before it can produce n+1).
struct P2PPush {
  int64_t nsnd;
  const int *snd_ids;                     // own local ids to send, grouped by neighbour
  const int *snd_nbr;                     // neighbour index of every element
  const int64_t *snd_ptrs;                // n_nbr+1 group offsets
  double *const *peer_buf0;               // per neighbour: my slot in its receive buffer, parity 0 / 1
  double *const *peer_buf1;
  unsigned long long *const *peer_flag;   // per neighbour: its flag word for me
  int n_nbr;
  unsigned long long *seq;                // device-resident exchange counter (CUDA-graph replayable)
  unsigned int *ticket;
};
__global__ void __launch_bounds__(256) p2p_push_kernel(P2PPush p, const double *__restrict__ v) {
  const unsigned long long seq = *(volatile unsigned long long *)p.seq + 1ull;  // this exchange's number
  double *const *peer_buf = (seq & 1ull) ? p.peer_buf1 : p.peer_buf0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.nsnd; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = p.snd_nbr[i];
    peer_buf[k][i - p.snd_ptrs[k]] = v[p.snd_ids[i]];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) is_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence_system();
    for (int k = threadIdx.x; k < p.n_nbr; k += blockDim.x) {
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peer_flag[k]), "l"(seq) : "memory");
    }
    if (threadIdx.x == 0) {
      *p.ticket = 0u;
      *p.seq = seq;  // every block has read the old value before arriving at the ticket
    }
  }
}

struct P2PWait {
  int n_nbr;
  const int *nbr_rank;                    // ranks I receive from
  const unsigned long long *flags;        // my flag words, indexed by sender rank
  const unsigned long long *seq;          // device counter, already advanced by the push kernel
  int64_t nrcv;
  const int *rcv_ids;                     // ghost local ids, grouped by neighbour
  const double *rcv_buf0;                 // parity 0 / 1
  const double *rcv_buf1;
};
__global__ void __launch_bounds__(256) p2p_wait_unpack_kernel(P2PWait w, double *__restrict__ v) {
  const unsigned long long seq = *w.seq;
  const double *rcv_buf = (seq & 1ull) ? w.rcv_buf1 : w.rcv_buf0;
  if (threadIdx.x < w.n_nbr) {
    const unsigned long long *f = w.flags + w.nbr_rank[threadIdx.x];
    const long long t0 = clock64();
    for (;;) {
      unsigned long long cur;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(cur) : "l"(f) : "memory");
      if (cur >= seq) break;
      if (clock64() - t0 > 20000000000LL) __trap();  // ~10 s: a peer died; fail instead of hanging
      __nanosleep(64);
    }
  }
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < w.nrcv; i += (int64_t)gridDim.x * blockDim.x)
    v[w.rcv_ids[i]] = __ldcg(rcv_buf + i);
}

// the whole exchange in ONE launch: every block pushes its share, the last block to finish publishes the sequence
// number to the neighbours, then all blocks wait for the neighbours' numbers and unpack.  The grid is never larger
// than what is co-resident (<= 2 blocks per SM), so the blocks that spin cannot starve the block that publishes.
__global__ void __launch_bounds__(256) p2p_exchange_kernel(P2PPush p, P2PWait w, const double *v, double *vdst) {
  const unsigned long long seq = *(volatile unsigned long long *)p.seq + 1ull;  // this exchange's number
  double *const *peer_buf = (seq & 1ull) ? p.peer_buf1 : p.peer_buf0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.nsnd; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = p.snd_nbr[i];
    peer_buf[k][i - p.snd_ptrs[k]] = v[p.snd_ids[i]];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) is_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence_system();
    for (int k = threadIdx.x; k < p.n_nbr; k += blockDim.x) {
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peer_flag[k]), "l"(seq) : "memory");
    }
    if (threadIdx.x == 0) {
      *p.ticket = 0u;
      *p.seq = seq;  // every block has read the old value before arriving at the ticket
    }
  }
  const double *rcv_buf = (seq & 1ull) ? w.rcv_buf1 : w.rcv_buf0;
  if (threadIdx.x < w.n_nbr) {
    const unsigned long long *f = w.flags + w.nbr_rank[threadIdx.x];
    const long long t0 = clock64();
    for (;;) {
      unsigned long long cur;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(cur) : "l"(f) : "memory");
      if (cur >= seq) break;
      if (clock64() - t0 > 20000000000LL) __trap();  // ~10 s: a peer died; fail instead of hanging
      __nanosleep(32);
    }
  }
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < w.nrcv; i += (int64_t)gridDim.x * blockDim.x)
    vdst[w.rcv_ids[i]] = __ldcg(rcv_buf + i);
}

// invd[i] = 1/A[i,i]  (JacobiLinearSolvers.jl:20-23,29-34; diagonal of the own-own block, kept per matrix)
__global__ void inv_diag_kernel(int64_t nrows, const double *__restrict__ diag, double *__restrict__ invd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) invd[i] = __ddiv_rn(1.0, diag[i]);
}

// assemble!: v[ids[i]] += buf[i] over one neighbour's segment (ids unique within a segment)
__global__ void add_at_kernel(int64_t n, const int *__restrict__ ids, const double *__restrict__ buf, double *__restrict__ v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[ids[i]] = __dadd_rn(v[ids[i]], buf[i]);
}

// ------------------------------------------------------------------------------------------
// dense coarse-level solver: Gauss-Jordan inverse with partial pivoting on the device (set-up),
// applied as a GEMV (HBM-bound: n^2*8 bytes) -- the "small coarse-level solve kept on-device".
__global__ void csr_to_dense_kernel(int64_t nrows, int64_t row_off, int64_t ld, const int *__restrict__ rowptr,
                                    const int *__restrict__ col, const int64_t *__restrict__ col_gid,
                                    const double *__restrict__ val, double *__restrict__ M) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
    const int64_t j = col_gid ? col_gid[col[e]] : col[e];
    M[(row_off + i) * ld + j] = val[e];
  }
}
__global__ void gj_set_identity_kernel(int64_t n, double *__restrict__ M) {  // right half of [A | I]
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) M[i * 2 * n + n + i] = 1.0;
}

// grid-wide barrier of a cooperative launch (all CTAs co-resident): monotone arrival counter
__device__ __forceinline__ void grid_barrier(unsigned int *ctr, unsigned int &gen) {
  __syncthreads();
  gen += 1;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const unsigned int target = gen * gridDim.x;
    for (;;) {
      unsigned int cur;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(ctr) : "memory");
      if (cur >= target) break;
    }
  }
  __syncthreads();
}

// Gauss-Jordan inverse with partial pivoting of the augmented matrix M = [A | I] (n x 2n, row-major), the
// WHOLE elimination in ONE cooperative launch: per column (a) every CTA scans its rows for the pivot
// candidate, (b) after a grid barrier every CTA reduces the candidates, one pass saves the scaled pivot row
// and the elimination factors (taken before the row swap) and swaps, (c) after a second barrier all rows are
// eliminated, third barrier.  Matrix reads use ld.global.cg: L1 is not coherent between SMs inside a launch.  Scratch: cand (2 doubles + 1 int per CTA), prow (2n), fcol (n).
struct GJArgs {
  int64_t n;
  double *M;
  double *prow, *fcol;
  double *cand_abs, *cand_val;
  int *cand_idx;
  unsigned int *barrier;  // zero at launch
};
template <int THREADS>
__global__ void __launch_bounds__(THREADS) gj_inverse_kernel(GJArgs g) {
  __shared__ double sv[THREADS], sval[THREADS];
  __shared__ int si[THREADS];
  const int64_t n = g.n, ld = 2 * g.n;
  unsigned int gen = 0;
  for (int64_t k = 0; k < n; ++k) {
    // (a) pivot candidates: argmax_{i>=k} |M[i][k]|, smallest index among equals
    {
      double best = -1.0, bval = 0.0;
      int bi = (int)k;
      for (int64_t i = k + (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
        const double v = __ldcg(g.M + i * ld + k);
        if (fabs(v) > best) { best = fabs(v); bi = (int)i; bval = v; }
      }
      sv[threadIdx.x] = best; si[threadIdx.x] = bi; sval[threadIdx.x] = bval;
      __syncthreads();
      for (int o = THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
          const int q = threadIdx.x + o;
          if (sv[q] > sv[threadIdx.x] || (sv[q] == sv[threadIdx.x] && si[q] < si[threadIdx.x])) {
            sv[threadIdx.x] = sv[q]; si[threadIdx.x] = si[q]; sval[threadIdx.x] = sval[q];
          }
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) { g.cand_abs[blockIdx.x] = sv[0]; g.cand_idx[blockIdx.x] = si[0]; g.cand_val[blockIdx.x] = sval[0]; }
    }
    grid_barrier(g.barrier, gen);
    // (b) every CTA reduces the candidates (same result everywhere)
    {
      double best = -1.0, bval = 0.0;
      int bi = (int)k;
      for (unsigned int q = threadIdx.x; q < gridDim.x; q += THREADS) {
        const double v = __ldcg(g.cand_abs + q);
        const int idx = __ldcg(g.cand_idx + q);
        if (v > best || (v == best && idx < bi)) { best = v; bi = idx; bval = __ldcg(g.cand_val + q); }
      }
      sv[threadIdx.x] = best; si[threadIdx.x] = bi; sval[threadIdx.x] = bval;
      __syncthreads();
      for (int o = THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
          const int q = threadIdx.x + o;
          if (sv[q] > sv[threadIdx.x] || (sv[q] == sv[threadIdx.x] && si[q] < si[threadIdx.x])) {
            sv[threadIdx.x] = sv[q]; si[threadIdx.x] = si[q]; sval[threadIdx.x] = sval[q];
          }
        }
        __syncthreads();
      }
    }
    const int64_t p = si[0];
    const double piv = sval[0];
    __syncthreads();
    // scaled pivot row, elimination factors (taken before the swap: row p will hold old row k), row swap
    // row p <- row k.  Only row p is written here, and each element of it by the thread that read it.
    for (int64_t t = (int64_t)blockIdx.x * THREADS + threadIdx.x; t < ld; t += (int64_t)gridDim.x * THREADS) {
      if (t < n) g.fcol[t] = (t == p) ? __ldcg(g.M + k * ld + k) : __ldcg(g.M + t * ld + k);
      const double a = __ldcg(g.M + p * ld + t);
      if (p != k) g.M[p * ld + t] = __ldcg(g.M + k * ld + t);
      g.prow[t] = a / piv;
    }
    grid_barrier(g.barrier, gen);
    // (c) eliminate: row k <- prow, row i <- row i - fcol[i] * prow
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
      double *Mi = g.M + i * ld;
      if (i == k) {
        for (int64_t j = threadIdx.x; j < ld; j += THREADS) Mi[j] = __ldcg(g.prow + j);
      } else {
        const double f = __ldcg(g.fcol + i);
        if (f != 0.0)
          for (int64_t j = k + threadIdx.x; j < ld; j += THREADS) Mi[j] = __ldcg(Mi + j) - f * __ldcg(g.prow + j);
      }
    }
    grid_barrier(g.barrier, gen);
  }
}
__global__ void gj_extract_kernel(int64_t n, int64_t row0, int64_t nrows, const double *__restrict__ M,
                                  double *__restrict__ inv) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j < n && i < nrows) inv[i * n + j] = M[(row0 + i) * 2 * n + n + j];
}
// y[i] = sum_j inv[i][j] * b[j]  : one warp per row, coalesced
template <int THREADS>
__global__ void __launch_bounds__(THREADS) dense_gemv_kernel(int64_t nrows, int64_t n, const double *__restrict__ inv,
                                                            const double *__restrict__ b, double *__restrict__ y) {
  const int64_t row = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= nrows) return;
  const double *r = inv + row * n;
  double s = 0.0;
  for (int64_t j = lane; j < n; j += 32) s += r[j] * b[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[row] = s;
}

}  // namespace gsb
