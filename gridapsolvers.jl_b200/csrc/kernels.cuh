// kernels.cuh -- hand-written sm_100a kernels of the gsb200 solve phase.
//
// Every kernel here is HBM-bandwidth bound: the design rules are perfectly
// streamed matrix traffic (TMA bulk copies of the CSR value / column arrays into a shared-memory
// ring, persistent CTAs), coalesced gathers of the input vector (one lane per row => the k-th
// neighbours of consecutive rows are consecutive in memory on mesh-ordered matrices), fusion of
// every elementwise epilogue into the row kernel, and deterministic two-stage reductions.
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
// latency bound) and for matrices whose rows do not fit the streaming kernel's ring.
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
// SELL-32 kernel: the matrix is also kept in sliced-ELLPACK form with slice height 32 (one warp =
// one slice, one lane = one row, entries of a slice stored column-major: entry k of the 32 rows
// is contiguous).  Every matrix load of a warp is then one fully coalesced 256 B (values) /
// 128 B (columns) transaction with addresses known in advance -- no row-pointer -> data
// dependence, no shared-memory staging, no barriers -- and with one lane per row the gathers of x
// are coalesced too and the row sum keeps the sequential ascending-column order (bit-exact with
// the oracle).  Matrix data is streamed with L1::no_allocate so that L1 keeps x.
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

struct SellArgs {
  const int *slice_list;  // optional indirection: the slices this launch processes (nullptr = all)
  int64_t n_list;         // number of slices this launch processes
  const int *rowptr;      // CSR row pointers (row lengths)
  const int *slice_off;   // per slice: offset of the slice in units of 32 entries; nslices+1 entries
  const int *col;         // padded, column-major per slice
  const double *val;
  int64_t nrows;
};

template <int MODE, int THREADS, int U, int MINB = 1, int STYLE = 0>
__global__ void __launch_bounds__(THREADS, MINB) csr_sell_kernel(SellArgs m, RowArgs a) {
  __shared__ double red_smem[THREADS / 32];
  const int64_t widx = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;  // one warp per slice
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  if (widx < m.n_list) {  // warp-uniform
    const int64_t slice = m.slice_list ? (int64_t)m.slice_list[widx] : widx;
    const int64_t row = (slice << 5) + lane;
    const bool valid = row < m.nrows;
    const int so0 = m.slice_off[slice], so1 = m.slice_off[slice + 1];
    const int width = so1 - so0;  // entries per row in this slice (warp-uniform)
    int len = 0;
    RowPre<MODE> pre{};
    double s = 0.0;
    if (valid) {
      len = m.rowptr[row + 1] - m.rowptr[row];
      row_prefetch<MODE>(a, row, pre);
      s = row_init<MODE>(a, row);
    }
    const size_t base = ((size_t)so0 << 5) + lane;
    const int *cp = m.col + base;
    const double *vp = m.val + base;
    const double al = a.alpha;
    int k = 0;
    for (; k + U <= width; k += U) {
      int cc[U];
      double vv[U], xv[U];
      if (STYLE == 0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          cc[u] = ldg_stream_s32(cp + (size_t)(k + u) * 32);
          vv[u] = ldg_stream_f64(vp + (size_t)(k + u) * 32);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = __ldg(a.x + cc[u]);
      } else {
        // burst order: all column ids, then all values (two contiguous bursts per warp), then the gathers
#pragma unroll
        for (int u = 0; u < U; ++u) cc[u] = ldg_stream_s32(cp + (size_t)(k + u) * 32);
#pragma unroll
        for (int u = 0; u < U; ++u) vv[u] = ldg_stream_f64(vp + (size_t)(k + u) * 32);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          double t;
          asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(t) : "l"(a.x + cc[u]));
          xv[u] = t;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k + u < len) {
          double t = xv[u];
          if (MODE == ROW_SPMV) t = __dmul_rn(t, al);
          s = __dadd_rn(s, __dmul_rn(vv[u], t));
        }
      }
    }
    for (; k < width; ++k) {
      const int c = ldg_stream_s32(cp + (size_t)k * 32);
      const double v = ldg_stream_f64(vp + (size_t)k * 32);
      double t = __ldg(a.x + c);
      if (k < len) {
        if (MODE == ROW_SPMV) t = __dmul_rn(t, al);
        s = __dadd_rn(s, __dmul_rn(v, t));
      }
    }
    if (valid) row_epilogue_pre<MODE>(a, row, s, pre, acc);
  }
  if (MODE == ROW_SPMV_DOT) {
    double v[1] = {acc};
    grid_reduce_finish<THREADS, 1>(v, a.red, red_smem);
  }
}

// ------------------------------------------------------------------------------------------
// L2-pipelined Jacobi-Richardson sweeps: S consecutive sweeps of the smoother in ONE launch, so that
// the matrix is streamed from HBM once and re-read S-1 times from the 126 MB L2.
//
// Work item = (stage j, chunk c of 256 rows).  Items are handed out in the order
//     tau = 0,1,2,... ; j = 0..S-1 ; c = tau - j*LAG
// i.e. stage j trails stage j-1 by LAG chunks.  Stage j of chunk c gathers dx produced by stage j-1
// on chunks [c-reach, c+reach] (reach = matrix bandwidth in chunks) and overwrites the dx buffer that
// stage j-1 itself gathered from, so it may start only when stage j-1 has completed every chunk up
// to c+reach: one monotone counter per stage (`prefix[j]` = number of leading chunks completed)
// carries both the RAW and the WAR dependency.  LAG > reach + (items in flight)/S makes the wait
// almost never spin, and every dependency of an item has a smaller ticket, so persistent CTAs that
// take tickets in order can never deadlock.  Arithmetic per row is exactly that of ROW_SWEEP /
// ROW_RESID (same order, same roundings): the result is bit-identical to S separate launches.
// Vectors written inside the kernel are read with ld.global.cg (L2) -- L1 is not coherent.
// matrix loads with an L2 cache-policy hint (createpolicy): evict_last while a later pipeline stage will
// re-read the line, evict_first on its final use
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double ldg_hint_f64(const double *p, uint64_t pol) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int ldg_hint_s32(const int *p, uint64_t pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}

struct PipeArgs {
  SellArgs m;
  double *r, *x;
  const double *invd;
  double *dxbuf[2];
  double omega;
  int k0;        // global index (1-based) of the sweep stage 0 performs
  int S;         // stages in this launch
  int niter;     // total sweeps of the smoother application: sweep niter only updates r
  int nchunks, lag, reach;
  unsigned int *ticket;  // work counter (zero at launch; reset by the last CTA)
  unsigned int *exited;
  int *prefix;   // S counters, zero at launch
  int *done;     // S * nchunks flags, compared against epoch
  int epoch;
  int l2_hints;  // use L2 eviction-priority hints on the matrix stream
};

template <int THREADS, int U>
__global__ void __launch_bounds__(THREADS, 4) sell_pipe_kernel(PipeArgs p) {
  __shared__ unsigned int s_t[2];
  __shared__ int s_lo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int total = (unsigned int)(p.nchunks + (p.S - 1) * p.lag) * (unsigned int)p.S;
  if (threadIdx.x == 0) s_t[0] = atomicAdd(p.ticket, 1u);
  const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
  int it = 0;
  for (;; ++it) {
    __syncthreads();
    const unsigned int t = s_t[it & 1];
    if (t >= total) break;
    // take the next ticket now: its latency overlaps this item's work
    if (threadIdx.x == 0) s_t[(it + 1) & 1] = atomicAdd(p.ticket, 1u);
    const int tau = (int)(t / (unsigned int)p.S), j = (int)(t % (unsigned int)p.S);
    const int c = tau - j * p.lag;
    if (c < 0 || c >= p.nchunks) continue;
    if (j > 0) {
      // stage j-1 must have completed every chunk below `need`: start from the published lower
      // bound and check the completion flags of the remaining chunks with the whole CTA
      const int need = min(p.nchunks, c + p.reach + 1);
      int *pf = p.prefix + (j - 1);
      const int *dn = p.done + (size_t)(j - 1) * p.nchunks;
      const long long t0 = clock64();
      for (;;) {
        // one thread reads the lower bound: every thread must scan from the SAME value, or the
        // strided scans would leave holes
        if (threadIdx.x == 0) {
          int v;
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(pf) : "memory");
          s_lo = v;
        }
        __syncthreads();
        const int lo = s_lo;
        bool ok = true;
        for (int q = lo + (int)threadIdx.x; q < need; q += THREADS) {
          int fl;
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(fl) : "l"(dn + q) : "memory");
          ok = ok && (fl == p.epoch);
        }
        if (__syncthreads_and(ok)) {
          if (threadIdx.x == 0 && need > lo) atomicMax(pf, need);
          break;
        }
        if (clock64() - t0 > 20000000000LL) __trap();
        __nanosleep(200);
      }
    }
    // a later stage re-reads this chunk's matrix rows -> ask L2 to keep them; final use -> evict first
    const uint64_t pol = (j + 1 < p.S) ? pol_keep : pol_drop;
    const int k = p.k0 + j;            // sweep index
    const bool last = (k == p.niter);  // the final sweep only updates the residual
    const double *xin = p.dxbuf[(k - 1) & 1];
    double *dxout = p.dxbuf[k & 1];
    const int64_t slice = (int64_t)c * (THREADS / 32) + warp;
    const int64_t nslices = (p.m.nrows + 31) >> 5;
    if (slice < nslices) {
      const int64_t row = (slice << 5) + lane;
      const bool valid = row < p.m.nrows;
      const int so0 = p.m.slice_off[slice], so1 = p.m.slice_off[slice + 1];
      const int width = so1 - so0;
      int len = 0;
      double rb = 0.0, idg = 0.0, xa = 0.0, s = 0.0;
      if (valid) {
        len = p.m.rowptr[row + 1] - p.m.rowptr[row];
        rb = __ldcg(p.r + row);
        if (!last) { idg = __ldg(p.invd + row); xa = __ldcg(p.x + row); }
      }
      const size_t base = ((size_t)so0 << 5) + lane;
      const int *cp = p.m.col + base;
      const double *vp = p.m.val + base;
      int kk = 0;
      for (; kk + U <= width; kk += U) {
        int cc[U];
        double vv[U], xv[U];
        if (p.l2_hints) {
#pragma unroll
          for (int u = 0; u < U; ++u) cc[u] = ldg_hint_s32(cp + (size_t)(kk + u) * 32, pol);
#pragma unroll
          for (int u = 0; u < U; ++u) vv[u] = ldg_hint_f64(vp + (size_t)(kk + u) * 32, pol);
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u) cc[u] = ldg_stream_s32(cp + (size_t)(kk + u) * 32);
#pragma unroll
          for (int u = 0; u < U; ++u) vv[u] = ldg_stream_f64(vp + (size_t)(kk + u) * 32);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = __ldcg(xin + cc[u]);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (kk + u < len) s = __dadd_rn(s, __dmul_rn(vv[u], xv[u]));
      }
      for (; kk < width; ++kk) {
        const int cidx = ldg_stream_s32(cp + (size_t)kk * 32);
        const double v = ldg_stream_f64(vp + (size_t)kk * 32);
        const double tx = __ldcg(xin + cidx);
        if (kk < len) s = __dadd_rn(s, __dmul_rn(v, tx));
      }
      if (valid) {
        const double rn = __dsub_rn(rb, s);
        p.r[row] = rn;
        if (!last) {
          const double d = __dmul_rn(p.omega, __dmul_rn(idg, rn));
          dxout[row] = d;
          p.x[row] = __dadd_rn(xa, d);
        }
      }
    }
    // publish: every thread's stores first (fence), then this chunk's completion flag
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.done + (size_t)j * p.nchunks + c), "r"(p.epoch) : "memory");
  }
  // the last CTA to leave resets the counters for the next launch
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.exited, 1u) == gridDim.x - 1) {
      *p.ticket = 0u;
      *p.exited = 0u;
      for (int j = 0; j < p.S; ++j) p.prefix[j] = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------
// SELL-32 with a staged x window (opt-in `xstage=1`; planner in xstage.h, NOT yet validated on hardware --
// round-2 work item 1).  One CTA = one chunk of THREADS rows: the contiguous column segments the chunk
// references are copied from the gathered vector into shared memory, the per-entry column id is a 16-bit
// offset into that window, and the row loop gathers from shared memory.  Same accumulation order as
// csr_sell_kernel (bit-identical results).
struct XStageArgs {
  const int *rowptr;
  const int *slice_off;
  const unsigned short *lcol;   // SELL layout, window offsets
  const double *val;            // SELL layout
  const int *chunk_seg_ptr;     // nchunks + 1
  const int *seg_start, *seg_len, *seg_off;
  int64_t nrows;
};
__device__ __forceinline__ unsigned short ldg_stream_u16(const unsigned short *p) {
  unsigned short v;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}

template <int MODE, int THREADS, int U>
__global__ void __launch_bounds__(THREADS, 4) csr_sell_xs_kernel(XStageArgs m, RowArgs a) {
  extern __shared__ double xwin[];
  __shared__ double red_smem[THREADS / 32];
  const int chunk = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // stage the window: segment after segment, coalesced
  const int sg0 = m.chunk_seg_ptr[chunk], sg1 = m.chunk_seg_ptr[chunk + 1];
  for (int sgi = sg0; sgi < sg1; ++sgi) {
    const int start = m.seg_start[sgi], len = m.seg_len[sgi], off = m.seg_off[sgi];
    for (int t = threadIdx.x; t < len; t += THREADS) xwin[off + t] = __ldg(a.x + start + t);
  }
  __syncthreads();
  double acc = 0.0;
  const int64_t slice = (int64_t)chunk * (THREADS / 32) + warp;
  const int64_t nslices = (m.nrows + 31) >> 5;
  if (slice < nslices) {
    const int64_t row = (slice << 5) + lane;
    const bool valid = row < m.nrows;
    const int so0 = m.slice_off[slice], so1 = m.slice_off[slice + 1];
    const int width = so1 - so0;
    int len = 0;
    RowPre<MODE> pre{};
    double s = 0.0;
    if (valid) {
      len = m.rowptr[row + 1] - m.rowptr[row];
      row_prefetch<MODE>(a, row, pre);
      s = row_init<MODE>(a, row);
    }
    const size_t base = ((size_t)so0 << 5) + lane;
    const unsigned short *cp = m.lcol + base;
    const double *vp = m.val + base;
    const double al = a.alpha;
    int k = 0;
    for (; k + U <= width; k += U) {
      unsigned short cc[U];
      double vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) cc[u] = ldg_stream_u16(cp + (size_t)(k + u) * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) vv[u] = ldg_stream_f64(vp + (size_t)(k + u) * 32);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k + u < len) {
          double t = xwin[cc[u]];
          if (MODE == ROW_SPMV) t = __dmul_rn(t, al);
          s = __dadd_rn(s, __dmul_rn(vv[u], t));
        }
      }
    }
    for (; k < width; ++k) {
      const unsigned short c = ldg_stream_u16(cp + (size_t)k * 32);
      const double v = ldg_stream_f64(vp + (size_t)k * 32);
      if (k < len) {
        double t = xwin[c];
        if (MODE == ROW_SPMV) t = __dmul_rn(t, al);
        s = __dadd_rn(s, __dmul_rn(v, t));
      }
    }
    if (valid) row_epilogue_pre<MODE>(a, row, s, pre, acc);
  }
  if (MODE == ROW_SPMV_DOT) {
    double v[1] = {acc};
    grid_reduce_finish<THREADS, 1>(v, a.red, red_smem);
  }
}

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA 1-D bulk copy (cp.async.bulk, SASS: UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------------------------------------
// streaming CSR kernel (the hot kernel): persistent CTAs, each owning a contiguous block of rows
// (balanced by nnz at set-up time).  One elected thread streams the CTA's slice of val[] / col[]
// through a shared-memory ring with TMA bulk copies (CHUNK non-zeros per copy, one mbarrier per
// ring slot); the other threads never touch DRAM for matrix data.  THREADS/G rows are processed
// per step, G lanes per row walking the row's segment of the ring; gathers of x go through L1/L2.
//   ring bytes = RING*(8+4); in flight per SM ~ (RING - group span) * 12 B.
template <int G, int MODE, int THREADS, int RING_LOG2, int CHUNK_LOG2>
__global__ void __launch_bounds__(THREADS) csr_stream_kernel(StreamArgs m, RowArgs a) {
  constexpr int RING = 1 << RING_LOG2;
  constexpr int CHUNK = 1 << CHUNK_LOG2;
  constexpr int NSLOT = RING / CHUNK;
  constexpr int ROWS = THREADS / G;
  constexpr int SPAN_MAX = RING / 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *sval = reinterpret_cast<double *>(smem_raw);
  int *scol = reinterpret_cast<int *>(smem_raw + (size_t)RING * 8);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)RING * 12);
  __shared__ double red_smem[THREADS / 32];

  const int R0 = m.cta_rows[blockIdx.x], R1 = m.cta_rows[blockIdx.x + 1];
  double acc = 0.0;
  if (R0 < R1) {
    const int E0 = m.rowptr[R0], E1 = m.rowptr[R1];
    const int S0 = E0 & ~3;  // 16-byte aligned start of this CTA's stream
    const int nchunks = (E1 - S0 + CHUNK - 1) >> CHUNK_LOG2;
    if (threadIdx.x == 0) {
      for (int s = 0; s < NSLOT; ++s) mbar_init(&full[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int issued = 0;  // next chunk to issue (thread 0 only)
    const int t_row = threadIdx.x / G, lane = threadIdx.x % G;
    int g0 = R0;
    // row pointers of the first group
    int rp0 = 0, rp1 = 0;
    {
      const int r = g0 + t_row;
      if (r < R1) { rp0 = m.rowptr[r]; rp1 = m.rowptr[r + 1]; }
    }
    __shared__ int s_gbase;
    while (g0 < R1) {
      if (threadIdx.x == 0) s_gbase = rp0;  // thread 0 holds the row pointer of row g0
      // barrier #1: the previous group is fully consumed (its ring slots may be refilled) and
      // the base of this group is published
      __syncthreads();
      const int gbase = s_gbase;
      const bool in_rng = (g0 + t_row < R1);
      // barrier #2: n = number of rows of this group (row-pointer prefix within SPAN_MAX)
      const int n = __syncthreads_count(in_rng && lane == 0 && (rp1 - gbase) <= SPAN_MAX);
      const int clo = (gbase - S0) >> CHUNK_LOG2;
      if (threadIdx.x == 0) {
        const int lim = min(nchunks, clo + NSLOT);
        for (; issued < lim; ++issued) {
          const int slot = issued & (NSLOT - 1);
          const int64_t start = (int64_t)S0 + ((int64_t)issued << CHUNK_LOG2);
          int64_t cnt = m.nnz_padded - start;
          if (cnt > CHUNK) cnt = CHUNK;
          mbar_expect_tx(&full[slot], (uint32_t)cnt * 12u);
          tma_bulk_g2s(sval + (size_t)slot * CHUNK, m.val + start, (uint32_t)cnt * 8u, &full[slot]);
          tma_bulk_g2s(scol + (size_t)slot * CHUNK, m.col + start, (uint32_t)cnt * 4u, &full[slot]);
        }
      }
      const bool active = (t_row < n);
      const int e0 = rp0, e1 = rp1;
      const int64_t row = g0 + t_row;
      // prefetch the next group's row pointers while this one is processed
      {
        const int r = g0 + n + t_row;
        rp0 = 0; rp1 = 0;
        if (r < R1) { rp0 = m.rowptr[r]; rp1 = m.rowptr[r + 1]; }
      }
      double s = 0.0;
      if (active) {
        if (e1 > e0) {
          const int c0 = (e0 - S0) >> CHUNK_LOG2, c1 = (e1 - 1 - S0) >> CHUNK_LOG2;
          for (int c = c0; c <= c1; ++c) mbar_wait(&full[c & (NSLOT - 1)], (uint32_t)((c / NSLOT) & 1));
        }
        if (lane == 0) s = row_init<MODE>(a, row);
        const double al = a.alpha;
#pragma unroll 4
        for (int e = e0 + lane; e < e1; e += G) {
          const int p = (e - S0) & (RING - 1);
          const int c = scol[p];
          const double v = sval[p];
          double xv = __ldg(a.x + c);
          if (MODE == ROW_SPMV) xv = __dmul_rn(xv, al);
          s = __dadd_rn(s, __dmul_rn(v, xv));
        }
      }
      if (G > 1) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
      }
      if (active && lane == 0) row_epilogue<MODE>(a, row, s, acc);
      g0 += n;
    }
  }
  if (MODE == ROW_SPMV_DOT) {
    double v[1] = {acc};
    grid_reduce_finish<THREADS, 1>(v, a.red, red_smem);
  }
}

// ------------------------------------------------------------------------------------------
// warp-specialised streaming CSR kernel (the production hot kernel).
//   warp NW          : producer -- one lane streams the CTA's val[]/col[] slice through the ring
//                      with TMA bulk copies, re-filling a slot as soon as every consumer warp has
//                      released it (empty[] mbarriers, count NW)
//   warps 0 .. NW-1  : consumers -- each owns every NW-th step of 32/G consecutive rows, waits on the
//                      full[] mbarriers of the chunks its row touches, walks the row in batches of U
//                      entries (U independent x-gathers in flight per lane), applies the fused
//                      epilogue, then releases the chunks it has moved past.
// There is no block-wide barrier in the main loop: warps drift apart by up to the ring capacity,
// so DRAM streaming, shared-memory reads and L1/L2 gathers of different warps overlap.
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int G, int MODE, int NW, int RING_LOG2, int CHUNK_LOG2, int U>
__global__ void __launch_bounds__((NW + 1) * 32, 1) csr_stream_ws_kernel(StreamArgs m, RowArgs a) {
  constexpr int THREADS = (NW + 1) * 32;
  constexpr int RING = 1 << RING_LOG2;
  constexpr int CHUNK = 1 << CHUNK_LOG2;
  constexpr int NSLOT = RING / CHUNK;
  constexpr int RW = 32 / G;  // rows per warp step
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *sval = reinterpret_cast<double *>(smem_raw);
  int *scol = reinterpret_cast<int *>(smem_raw + (size_t)RING * 8);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)RING * 12);
  uint64_t *empty = full + NSLOT;
  __shared__ double red_smem[(THREADS + 31) / 32];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R0 = m.cta_rows[blockIdx.x], R1 = m.cta_rows[blockIdx.x + 1];
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  double acc = 0.0;
  if (R0 < R1) {
    const int E0 = m.rowptr[R0], E1 = m.rowptr[R1];
    const int S0 = E0 & ~3;  // 16-byte aligned start of this CTA's stream
    const int nchunks = (E1 - S0 + CHUNK - 1) >> CHUNK_LOG2;
    if (warp == NW) {
      // ------------------------------------------------ producer
      if (lane == 0) {
        for (int c = 0; c < nchunks; ++c) {
          const int slot = c & (NSLOT - 1);
          if (c >= NSLOT) mbar_wait(&empty[slot], (uint32_t)(((c / NSLOT) - 1) & 1));
          const int64_t start = (int64_t)S0 + ((int64_t)c << CHUNK_LOG2);
          int64_t cnt = m.nnz_padded - start;
          if (cnt > CHUNK) cnt = CHUNK;
          mbar_expect_tx(&full[slot], (uint32_t)cnt * 12u);
          tma_bulk_g2s(sval + (size_t)slot * CHUNK, m.val + start, (uint32_t)cnt * 8u, &full[slot]);
          tma_bulk_g2s(scol + (size_t)slot * CHUNK, m.col + start, (uint32_t)cnt * 4u, &full[slot]);
        }
      }
    } else {
      // ------------------------------------------------ consumers
      const int sub = lane / G, gl = lane % G;
      int released = 0;  // meaningful in lane 0
      int first = R0 + warp * RW;  // first row of this warp's current step
      int row = first + sub;
      int rp0 = 0, rp1 = 0;
      if (row < R1) { rp0 = m.rowptr[row]; rp1 = m.rowptr[row + 1]; }
      while (first < R1) {
        const bool valid = row < R1;
        RowPre<MODE> pre{};
        if (valid && gl == 0) row_prefetch<MODE>(a, row, pre);
        // row pointers of this warp's next step
        const int nfirst = first + NW * RW, nrow = nfirst + sub;
        int nrp0 = 0, nrp1 = 0;
        if (nrow < R1) { nrp0 = m.rowptr[nrow]; nrp1 = m.rowptr[nrow + 1]; }
        double s = 0.0;
        if (valid) {
          const int e0 = rp0, e1 = rp1;
          if (e1 > e0) {
            const int c0 = (e0 - S0) >> CHUNK_LOG2, c1 = (e1 - 1 - S0) >> CHUNK_LOG2;
            for (int c = c0; c <= c1; ++c) mbar_wait(&full[c & (NSLOT - 1)], (uint32_t)((c / NSLOT) & 1));
          }
          if (gl == 0) s = row_init<MODE>(a, row);
          const double al = a.alpha;
          for (int k = e0 + gl; k < e1; k += U * G) {
            int cc[U];
            double vv[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int e = k + u * G;
              const int p = (e - S0) & (RING - 1);
              cc[u] = (e < e1) ? scol[p] : 0;
              vv[u] = (e < e1) ? sval[p] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = (k + u * G < e1) ? __ldg(a.x + cc[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (k + u * G < e1) {
                double t = xv[u];
                if (MODE == ROW_SPMV) t = __dmul_rn(t, al);
                s = __dadd_rn(s, __dmul_rn(vv[u], t));
              }
            }
          }
        }
        if (G > 1) {
#pragma unroll
          for (int o = G / 2; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
        }
        if (valid && gl == 0) row_epilogue_pre<MODE>(a, row, s, pre, acc);
        // release the chunks this warp has moved past
        const int e_next = __shfl_sync(0xffffffffu, (nfirst < R1) ? nrp0 : E1, 0);
        const int pass = (nfirst < R1) ? ((e_next - S0) >> CHUNK_LOG2) : nchunks;
        __syncwarp();
        // (a chunk is released only once it has landed: a warp that skips a chunk must not arrive
        //  on empty[] before the slot's previous use has been released by every warp, or its
        //  arrival would be counted in the wrong phase)
        if (lane == 0)
          for (; released < pass; ++released) {
            mbar_wait(&full[released & (NSLOT - 1)], (uint32_t)((released / NSLOT) & 1));
            mbar_arrive(&empty[released & (NSLOT - 1)]);
          }
        first = nfirst; row = nrow; rp0 = nrp0; rp1 = nrp1;
      }
      if (lane == 0)
        for (; released < nchunks; ++released) {
          mbar_wait(&full[released & (NSLOT - 1)], (uint32_t)((released / NSLOT) & 1));
          mbar_arrive(&empty[released & (NSLOT - 1)]);
        }
    }
  }
  if (MODE == ROW_SPMV_DOT) {
    double v[1] = {acc};
    grid_reduce_finish<THREADS, 1>(v, a.red, red_smem);
  }
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

// diag extraction: invd[i] = 1/A[i,i]  (JacobiLinearSolvers.jl:20-23,29-34; own-own block)
__global__ void inv_diag_kernel(int64_t nrows, const int *__restrict__ rowptr, const int *__restrict__ col,
                                const double *__restrict__ val, double *__restrict__ invd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  double d = 0.0;  // diag() of a sparse matrix returns 0 for a missing entry
  for (int e = rowptr[i]; e < rowptr[i + 1]; ++e)
    if (col[e] == i) d = val[e];
  invd[i] = __ddiv_rn(1.0, d);
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
// one block: pivot = argmax_{i>=k} |M[i][k]|
__global__ void gj_pivot_kernel(int64_t n, int64_t k, const double *__restrict__ M, int *__restrict__ piv,
                                double *__restrict__ pivval) {
  __shared__ double sv[256];
  __shared__ int si[256];
  double best = -1.0;
  int bi = (int)k;
  for (int64_t i = k + threadIdx.x; i < n; i += blockDim.x) {
    const double v = fabs(M[i * 2 * n + k]);
    if (v > best) { best = v; bi = (int)i; }
  }
  sv[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      if (sv[threadIdx.x + o] > sv[threadIdx.x] ||
          (sv[threadIdx.x + o] == sv[threadIdx.x] && si[threadIdx.x + o] < si[threadIdx.x])) {
        sv[threadIdx.x] = sv[threadIdx.x + o];
        si[threadIdx.x] = si[threadIdx.x + o];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *piv = si[0];
    *pivval = M[(int64_t)si[0] * 2 * n + k];
  }
}
// factors of the elimination step, taken BEFORE the row swap: fcol[i] = M[i][k], and the row
// that will hold old row k after the swap (row piv) gets old M[k][k]
__global__ void gj_fcol_kernel(int64_t n, int64_t k, const double *__restrict__ M, const int *__restrict__ piv,
                               double *__restrict__ fcol) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t p = *piv;
  fcol[i] = (i == p) ? M[k * 2 * n + k] : M[i * 2 * n + k];
}
// prow = row piv / pivot ; row piv <- row k (row k itself is written by the eliminate kernel)
__global__ void gj_swap_scale_kernel(int64_t n, int64_t k, double *__restrict__ M, const int *__restrict__ piv,
                                     const double *__restrict__ pivval, double *__restrict__ prow) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= 2 * n) return;
  const int64_t p = *piv;
  const double a = M[p * 2 * n + j];
  if (p != k) M[p * 2 * n + j] = M[k * 2 * n + j];
  prow[j] = a / *pivval;
}
__global__ void gj_eliminate_kernel(int64_t n, int64_t k, double *__restrict__ M, const double *__restrict__ prow,
                                    const double *__restrict__ fcol) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j >= 2 * n) return;
  if (i == k) {
    M[i * 2 * n + j] = prow[j];
  } else {
    const double f = fcol[i];
    if (f != 0.0) M[i * 2 * n + j] -= f * prow[j];
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
