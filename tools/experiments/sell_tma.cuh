// EXPERIMENT, NOT PART OF THE BUILD (round 2).  Kept for the record next to profiles/sell_tma_r02.json:
// a TMA-fed persistent variant of the block-SELL kernel.  Findings on B200 (C2 fine level, 2 048 383 rows):
//   * with the gathers of x made trivial (all column words 0) the bulk-copy value stream runs at 1.005 of the
//     measured HBM copy peak (SpMV 73.9 us for 454 MB) -- the register-fed kernel reaches 0.84 in the same test;
//   * with the real gathers every variant (2/3 stages, 16..24 warps per SM, gathers software-pipelined one chunk
//     ahead, x windows staged in shared memory with LDGSTS) lands at 94..100 us: the kernel is bound by the NUMBER
//     of L1 gather requests (T ~= 0.074 us/MB + 31 us per million warp-gathers), not by bytes in flight;
//   * the fix that paid is in the format / register kernel instead: runs of three consecutive columns are gathered
//     once and shuffled (kernels.cuh sell_kernel, `kind` slices): SpMV 78.8 us, residual 79.8 us, sweep 95.2 us.
// To try it again: copy next to kernels.cuh, include from core.cu and dispatch from launch_sell_list.
// sell_tma.cuh -- TMA-fed persistent variant of the block-SELL-32 row kernel (kernels.cuh sell_kernel).
//
// Why: with the column ids compressed to one word per 32 blocks (diagonal-aligned slices) the register-fed
// kernel is no longer bound by HBM but by the bytes it can keep in flight: every outstanding matrix load
// needs a destination register, and 64 registers x 32 warps hold ~70 KB per SM of which only part is ever
// in flight because a warp alternates between issuing loads and consuming them.  Here the matrix VALUES
// (97 % of the traffic) never pass through registers on their way in: every warp owns a small ring of
// shared-memory stages, one elected lane issues one 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) per
// chunk of KC slots of its current slice -- the chunk is one contiguous byte range of the SELL value array
// -- the column words of the chunk follow with 4-byte cp.async copies, and both complete on the stage's
// mbarrier (complete_tx bytes + cp.async.mbarrier.arrive).  Chunks are issued NST-1 ahead of the one being
// consumed, so a warp keeps (NST-1) x KC x BS^2 x 256 B in flight with no register cost, and the only
// register-fed loads left in the loop are the gathers of x (L1/L2 hits on a mesh-ordered matrix).
//
// Same arithmetic as sell_kernel: one lane per block row, slots visited in ascending order, product rounded
// before the add => bit-identical results (tests/test_gpu_kernels.py runs both kernels against the oracle).
// Warps are persistent (warp w takes slices w, w + W, w + 2W, ..): no CTA-level synchronisation in the loop.
#pragma once
#include "kernels.cuh"

namespace gsb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst_smem, const void *src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
// the mbarrier tracks this thread's cp.async copies issued so far: pending count +1 now, -1 when they have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// predicated read-only gather as ONE volatile instruction: keeps its place in the instruction stream (the
// compiler moves plain invariant loads across anything) and needs no branch
__device__ __forceinline__ void ldg_nc_f64_if(double &v, const double *p, unsigned on) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.u32 q, %2, 0;\n"
      "@q ld.global.nc.f64 %0, [%1];\n"
      "}\n"
      : "+d"(v)
      : "l"(p), "r"(on));
}

// one ring stage: values | column words | slot-validity words | row ids
template <int BS, int KC>
struct SellTmaCfg {
  static constexpr int BB = BS * BS;
  static constexpr int VAL_BYTES = KC * BB * 256;  // KC slots x BS^2 lines of 32 doubles
  static constexpr int KW_OFF = VAL_BYTES, LM_OFF = VAL_BYTES + 128, PERM_OFF = VAL_BYTES + 256;
  static constexpr int STAGE_BYTES = VAL_BYTES + 384;
};

// gathers of one chunk whose column words are all affine: predicated by the validity mask, in groups of G slots
// skipped by a warp-uniform branch
template <int BS, int KC, int G>
__device__ __forceinline__ void tma_chunk_gather(const double *__restrict__ x, unsigned lmc, int kc, int lane, uint32_t kw_addr,
                                                 double (&xv)[KC][BS]) {
#pragma unroll
  for (int g = 0; g < KC; g += G) {
    if (g < kc) {
#pragma unroll
      for (int u = g; u < g + G; ++u) {
        const int c = lds_s32(kw_addr + u * 4) + lane;  // broadcast read of the column word
        const double *xp = x + (int64_t)c * BS;
#pragma unroll
        for (int j = 0; j < BS; ++j) {
          xv[u][j] = 0.0;
          ldg_nc_f64_if(xv[u][j], xp + j, (lmc >> u) & 1u);
        }
      }
    }
  }
}
// products of one chunk out of shared memory, slot after slot (ascending column order inside every row)
template <int BS, int KC, int G, bool SCALE>
__device__ __forceinline__ void tma_chunk_multiply(const double (&xv)[KC][BS], double al, unsigned lmc, int kc, uint32_t vbase,
                                                   double (&s)[BS]) {
  constexpr int BB = BS * BS;
#pragma unroll
  for (int g = 0; g < KC; g += G) {
    if (g < kc) {
#pragma unroll
      for (int u = g; u < g + G; ++u) {
        if ((lmc >> u) & 1u) {
          double t[BS];
#pragma unroll
          for (int j = 0; j < BS; ++j) t[j] = SCALE ? __dmul_rn(xv[u][j], al) : xv[u][j];
#pragma unroll
          for (int i = 0; i < BS; ++i)
#pragma unroll
            for (int j = 0; j < BS; ++j)
              s[i] = __dadd_rn(s[i], __dmul_rn(lds_f64(vbase + (uint32_t)((u * BB + i * BS + j) * 256)), t[j]));
        }
      }
    }
  }
}

// Each warp is the producer and the consumer of its own stream of chunks (a chunk = at most KC slots of one slice),
// software-pipelined over three stages of shared memory:
//   issue(n+2)  : starts the bulk copy of the chunk's values and 4-byte cp.async copies of its column words (and,
//                 on the first chunk of a slice, of the slice's slot-validity words and row ids) -- everything the
//                 warp needs arrives in shared memory, asynchronously, on the stage's mbarrier;
//   gather(n+1) : waits for the stage (issued a whole step earlier) and issues the gathers of x for all its slots
//                 into registers (fully unrolled, predicated by the validity mask);
//   multiply(n) : the gathers issued in the previous step have landed: products out of shared memory.
// So the gathers of a chunk are in flight while the previous chunk is multiplied and the chunk after it is
// streamed from HBM: neither latency is exposed.  Chunk descriptors (slice, first slot, width) are warp-uniform
// registers; the offsets of the next slice are loaded one slice ahead.  A slice without slots still gets one
// (empty) chunk so that its rows see their epilogue.
struct TmaChunk {
  int sl, so0, width, k0;
};
template <int BS>
struct TmaLaneState {  // per-lane state that travels with a chunk from the gather step to the multiply step
  unsigned lmc;        // validity of the chunk's slots
  int brow;            // block row of the lane (-1: padding lane), set on the first chunk of a slice
  bool has_explicit;   // the chunk has explicit column-id lines: gathered in the multiply step instead
};

template <int MODE, int BS, bool PERM, int WARPS, int KC, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) sell_tma_kernel(SellArgs m, RowArgs a) {
  using Cfg = SellTmaCfg<BS, KC>;
  constexpr int BB = Cfg::BB;
  constexpr int THREADS = WARPS * 32;
  constexpr int NST = 3;
  constexpr int G = (KC % 4 == 0) ? 4 : 3;
  static_assert(KC <= 32 && KC % G == 0, "chunk = at most 32 slots, a multiple of the group size");
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ double red_smem[THREADS / 32];
  __shared__ __align__(8) uint64_t bars[WARPS * NST];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t wbuf = smem_u32(dyn_smem) + (uint32_t)warp * NST * Cfg::STAGE_BYTES;
  const uint32_t wbar = smem_u32(bars) + (uint32_t)warp * NST * 8;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) mbar_init(wbar + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, GW = (int64_t)gridDim.x * WARPS;
  const int64_t n_mine = gw < m.n_list ? (m.n_list - gw + GW - 1) / GW : 0;
  double acc = 0.0;

  // offsets of slice ordinal `nx_ord`, loaded one slice before they are needed
  int64_t nx_ord = 0;
  int nx_sl = 0, nx_o0 = 0, nx_o1 = 0;
  auto prefetch_slice = [&]() {
    if (nx_ord < n_mine) {
      const int64_t g = gw + nx_ord * GW;
      nx_sl = m.slice_list ? m.slice_list[g] : (int)g;
      nx_o0 = m.slice_off[nx_sl];
      nx_o1 = m.slice_off[nx_sl + 1];
    }
  };
  auto issue = [&](const TmaChunk &c, unsigned st) {
    const uint32_t sbase = wbuf + st * Cfg::STAGE_BYTES, bar = wbar + st * 8;
    const int kc = min(KC, c.width - c.k0);  // 0 for a slice without slots
    if (lane < kc) cp_async_4(sbase + Cfg::KW_OFF + lane * 4, m.kbase + c.so0 + c.k0 + lane);
    if (c.k0 == 0) {
      const int64_t pos = ((int64_t)c.sl << 5) + lane;
      cp_async_4(sbase + Cfg::LM_OFF + lane * 4, m.lmask + pos);
      if (PERM) cp_async_4(sbase + Cfg::PERM_OFF + lane * 4, m.perm + pos);
    }
    cp_async_mbar_arrive(bar);
    __syncwarp();
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)kc * BB * 256u;
      if (bytes) {
        mbar_expect_tx(bar, bytes);
        tma_bulk_g2s(sbase, m.val + ((m.dbg & 1) ? (size_t)0 : (((size_t)(c.so0 + c.k0) * BB) << 5)), bytes, bar);
      } else {
        mbar_arrive(bar);
      }
    }
  };
  // the chunk after c; returns false at the end of the warp's stream
  auto advance = [&](TmaChunk &c) -> bool {
    if (c.k0 + KC < c.width) {
      c.k0 += KC;
      return true;
    }
    if (nx_ord >= n_mine) return false;
    c.sl = nx_sl; c.so0 = nx_o0; c.width = nx_o1 - nx_o0; c.k0 = 0;
    nx_ord += 1;
    prefetch_slice();
    return true;
  };

  // multiply-side state of the current slice
  int64_t brow = 0;
  bool valid = false;
  RowPre<MODE> pre{};
  double s[BS];
#pragma unroll
  for (int i = 0; i < BS; ++i) s[i] = 0.0;
  const double al = a.alpha;
  int lm = 0;  // gather-side: validity word of the lane in the slice being gathered

  // gather step of chunk c (stage st, use number `use` of that stage): returns the per-lane state of the chunk
  auto gather_step = [&](const TmaChunk &c, unsigned st, unsigned use, double (&xv)[KC][BS], TmaLaneState<BS> &ls) {
    const uint32_t sbase = wbuf + st * Cfg::STAGE_BYTES;
    mbar_wait(wbar + st * 8, use & 1u);
    const int kc = min(KC, c.width - c.k0);
    if (c.k0 == 0) {
      const int64_t pos = ((int64_t)c.sl << 5) + lane;
      lm = lds_s32(sbase + Cfg::LM_OFF + lane * 4);
      ls.brow = PERM ? lds_s32(sbase + Cfg::PERM_OFF + lane * 4) : (pos < m.n_brows ? (int)pos : -1);
      if (ls.brow < 0) lm = 0;
    }
    ls.lmc = 0u;
    ls.has_explicit = false;
    if (kc > 0) {
      if (c.width > 32) {
        const int cnt = min(max(lm - c.k0, 0), kc);
        ls.lmc = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
      } else {
        ls.lmc = ((unsigned)lm >> c.k0) & (kc >= 32 ? 0xffffffffu : ((1u << kc) - 1u));
      }
      const int kw_l = lds_s32(sbase + Cfg::KW_OFF + lane * 4);
      ls.has_explicit = __any_sync(0xffffffffu, lane < kc && kw_l < 0);
      if (!ls.has_explicit) tma_chunk_gather<BS, KC, G>(a.x, ls.lmc, kc, lane, sbase + Cfg::KW_OFF, xv);
    }
  };
  // multiply step of chunk c (its gathers were issued one step earlier)
  auto multiply_step = [&](const TmaChunk &c, unsigned st, const double (&xv)[KC][BS], const TmaLaneState<BS> &ls) {
    const uint32_t sbase = wbuf + st * Cfg::STAGE_BYTES;
    const int kc = min(KC, c.width - c.k0);
    if (c.k0 == 0) {  // ---- slice prologue
      brow = ls.brow;
      valid = ls.brow >= 0;
#pragma unroll
      for (int i = 0; i < BS; ++i) s[i] = 0.0;
      if (valid) {
        if (BS == 1) row_prefetch<MODE>(a, brow, pre);
#pragma unroll
        for (int i = 0; i < BS; ++i) s[i] = row_init<MODE>(a, brow * BS + i);
      }
    }
    if (kc > 0) {
      const uint32_t vbase = sbase + lane * 8;
      if (!ls.has_explicit) {
        if (MODE == ROW_SPMV && al != 1.0)
          tma_chunk_multiply<BS, KC, G, true>(xv, al, ls.lmc, kc, vbase, s);
        else  // x * 1.0 == x exactly: the unscaled path is bit-identical
          tma_chunk_multiply<BS, KC, G, false>(xv, al, ls.lmc, kc, vbase, s);
      } else {
        // general path: explicit column-id lines (packed slices of unstructured matrices, a few boundary slices)
        const int kw_l = lds_s32(sbase + Cfg::KW_OFF + lane * 4);
#pragma unroll 1
        for (int u = 0; u < kc; ++u) {
          const int kb = __shfl_sync(0xffffffffu, kw_l, u);
          const bool on = (ls.lmc >> u) & 1u;
          int cc = kb + lane;
          if (kb < 0) cc = ldg_stream_s32(m.bcol + (((size_t)(~kb)) << 5) + lane);
          if (on) {
            double t[BS];
#pragma unroll
            for (int j = 0; j < BS; ++j) {
              t[j] = __ldg(a.x + (int64_t)cc * BS + j);
              if (MODE == ROW_SPMV) t[j] = __dmul_rn(t[j], al);
            }
#pragma unroll
            for (int i = 0; i < BS; ++i)
#pragma unroll
              for (int j = 0; j < BS; ++j)
                s[i] = __dadd_rn(s[i], __dmul_rn(lds_f64(vbase + (uint32_t)((u * BB + i * BS + j) * 256)), t[j]));
          }
        }
      }
    }
    if (c.k0 + KC >= c.width) {  // ---- slice epilogue
      if (valid) {
        if (BS == 1) {
          row_epilogue_pre<MODE>(a, brow, s[0], pre, acc);
        } else {
#pragma unroll
          for (int i = 0; i < BS; ++i) row_epilogue<MODE>(a, brow * BS + i, s[i], acc);
        }
      }
    }
  };

  // chunk j lives in stage j % 3 and is the (j / 3)-th use of it.  Life of chunk j: issued at the end of step j-3
  // (into the stage chunk j-3 has just released), gathered in step j-1, multiplied in step j.
  double xa[KC][BS], xb[KC][BS];
  TmaLaneState<BS> la{0u, -1, false}, lb{0u, -1, false};
  TmaChunk cm{0, 0, 0, 0}, c1{0, 0, 0, 0}, c2{0, 0, 0, 0}, ci{0, 0, 0, 0};  // chunks n, n+1, n+2 and the issue cursor
  prefetch_slice();
  bool h0 = advance(ci);  // chunk 0
  if (h0) issue(ci, 0u);
  cm = ci;
  bool h1 = h0 && advance(ci);  // chunk 1
  if (h1) issue(ci, 1u);
  c1 = ci;
  bool h2 = h1 && advance(ci);  // chunk 2
  if (h2) issue(ci, 2u);
  c2 = ci;
  bool hi = h2;  // the issue cursor has not run off the end of the stream
  unsigned n = 0;  // index of the chunk under cm
  if (h0) gather_step(cm, 0u, 0u, xa, la);
  // at the top of step n: chunk n (cm) has been gathered into the `cur` registers, chunks n+1 (c1) and n+2 (c2) have
  // been issued (if they exist)
  auto step = [&](double (&xcur)[KC][BS], TmaLaneState<BS> &lcur, double (&xnext)[KC][BS], TmaLaneState<BS> &lnext) {
    if (h1) gather_step(c1, (n + 1u) % 3u, (n + 1u) / 3u, xnext, lnext);
    multiply_step(cm, n % 3u, xcur, lcur);
    __syncwarp();  // every lane is done with stage n % 3: chunk n+3 may land there
    hi = hi && advance(ci);
    if (hi) issue(ci, n % 3u);
    cm = c1; h0 = h1;
    c1 = c2; h1 = h2;
    c2 = ci; h2 = hi;
    n += 1u;
  };
#pragma unroll 1
  while (h0) {
    step(xa, la, xb, lb);
    if (!h0) break;
    step(xb, lb, xa, la);
  }
  if (MODE == ROW_SPMV_DOT) {
    double v[1] = {acc};
    grid_reduce_finish<THREADS, 1>(v, a.red, red_smem);
  }
}

}  // namespace gsb
