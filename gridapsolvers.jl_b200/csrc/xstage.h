// xstage.h -- host-side planner of the "staged x window" variant of the SELL-32 row kernel
// (round-2 work item 1 of DESIGN.md section 10; the kernel is opt-in, `xstage=1`).
//
// For every chunk of CHUNK_ROWS consecutive rows the distinct columns its rows reference are covered by a
// few contiguous column segments (a 27-point stencil on a mesh-ordered matrix: 3 planes x 3-4 lines);
// the kernel copies those segments of the gathered vector into shared memory once per chunk and the
// per-entry column id becomes a 16-bit offset into that window.  Per row this replaces 216 B of L1/L2
// gather traffic by ~100 B of coalesced window loads and shrinks the column stream from 4 to 2 B/nnz,
// without touching the order in which a row's products are accumulated (bit-exactness is preserved).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace gsb {

struct XStagePlan {
  int chunk_rows = 256;
  bool ok = false;
  int max_window = 0;                 // doubles
  int64_t total_window = 0;           // sum over chunks (doubles): the x traffic of one pass
  std::vector<int> chunk_seg_ptr;     // nchunks + 1
  std::vector<int> seg_start;         // first column of the segment
  std::vector<int> seg_len;           // number of columns
  std::vector<int> seg_off;           // offset of the segment inside the chunk's window
  std::vector<uint16_t> lcol;         // per SELL entry (same layout as the SELL column array)
};

// rowptr/col: CSR with ascending columns; sell_off: per slice offset in units of 32 entries (nslices+1).
// gap: two column runs closer than `gap` are merged into one segment; cap: largest window allowed.
inline XStagePlan build_xstage(int64_t n_rows, const int *rowptr, const int *col, const int *sell_off, int chunk_rows,
                               int gap, int cap) {
  XStagePlan p;
  p.chunk_rows = chunk_rows;
  const int64_t nchunks = (n_rows + chunk_rows - 1) / chunk_rows;
  const int64_t nslices = (n_rows + 31) / 32;
  const int64_t entries = (int64_t)sell_off[nslices] * 32;
  p.lcol.assign((size_t)std::max<int64_t>(entries, 1), 0);
  p.chunk_seg_ptr.assign((size_t)nchunks + 1, 0);
  std::vector<int> cols;
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t r0 = c * chunk_rows, r1 = std::min<int64_t>(n_rows, r0 + chunk_rows);
    cols.assign(col + rowptr[r0], col + rowptr[r1]);
    std::sort(cols.begin(), cols.end());
    cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
    const size_t seg0 = p.seg_start.size();
    int off = 0;
    for (size_t i = 0; i < cols.size();) {
      size_t j = i;
      while (j + 1 < cols.size() && cols[j + 1] - cols[j] <= gap) ++j;
      const int start = cols[i], len = cols[j] - cols[i] + 1;
      p.seg_start.push_back(start);
      p.seg_len.push_back(len);
      p.seg_off.push_back(off);
      off += len;
      i = j + 1;
    }
    p.chunk_seg_ptr[(size_t)c + 1] = (int)p.seg_start.size();
    p.max_window = std::max(p.max_window, off);
    p.total_window += off;
    if (off > cap || off > 65535) return p;  // ok stays false: the caller keeps the plain kernel
    // 16-bit window offsets of every entry of the chunk's rows, in SELL layout
    const size_t nseg = p.seg_start.size() - seg0;
    for (int64_t i = r0; i < r1; ++i) {
      const size_t base = ((size_t)sell_off[i >> 5] << 5) + (size_t)(i & 31);
      for (int e = rowptr[i], k = 0; e < rowptr[i + 1]; ++e, ++k) {
        // last segment whose start is <= col[e]
        size_t lo = 0, hi = nseg;
        while (hi - lo > 1) {
          const size_t mid = (lo + hi) / 2;
          if (p.seg_start[seg0 + mid] <= col[e]) lo = mid; else hi = mid;
        }
        p.lcol[base + (size_t)k * 32] = (uint16_t)(p.seg_off[seg0 + lo] + (col[e] - p.seg_start[seg0 + lo]));
      }
    }
  }
  p.ok = true;
  return p;
}

}  // namespace gsb
