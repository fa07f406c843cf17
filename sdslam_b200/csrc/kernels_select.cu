// kernels_select.cu -- keypoint culling of ComputeKeyPoints for sm_100a.
//
// Reference: /root/reference/src/ORBextractor.cc:538-604.  Per level:
//   1. per-cell quota nfeaturesCell with iterative redistribution of the deficit of poor cells (:538-575);
//   2. per cell  KeyPointsFilter::retainBest(cell, nToRetain) + resize(nToRetain)  (:583-588);
//   3. cells concatenated row-major, then  retainBest(level, nDesired) + resize  (:601-604).
// retainBest's survivors and their ORDER are defined by libstdc++'s nth_element (see introselect.cuh); the order is
// part of the output (keypoints and descriptor rows follow it), so it is reproduced move for move.
//
// Two kernels.  gather_cells_kernel: one warp per (cell, frame) walks the cell rectangle of the FAST kernel's keypoint
// map in row-major order -- cv::FAST's emission order -- and compacts the keypoints into the cell's list with ballot /
// popc prefix sums (no atomics: the lists are bit-identical on every run).  The walk is pure memory latency, so it
// lives in its own register-light kernel that fills the SMs with warps.  select_kernel: one CTA per (level, frame)
// computes the quotas, and its warps load the lists into shared memory and trim them.
//
// nth_element is sequential as written, but its cost is all in the Hoare partition, and that parallelises exactly:
// with pivot response p, the left scan stops at elements with response <= p, the right scan at elements with
// response >= p, and the t-th left stopper L_t is swapped with the t-th right stopper R_t for as long as L_t < R_t.
// So one warp pass ranks the stoppers of the range (ballot + popc), T = #{t : L_t < R_t} pairs are swapped in
// parallel and the cut is min(L_T, R_{T-1}) -- the element moves of std::__unguarded_partition, in n/32 steps.
// Median-of-three, the depth limit with its heap-select fallback and the final insertion sort stay on one lane.
// tests/test_gpu_parity.py::test_warp_nth_element_equals_std checks the routine against the real std::nth_element.
#include <climits>

#include "introselect.cuh"
#include "kernels.cuh"

namespace sdorb {

// select_kernel comes in three shapes: 4 warps per (level, frame) CTA for batches (many CTAs: throughput), 16 warps for a handful
// of frames (the single-frame call of Frame.cc:195: 8 CTAs in all, so the cells of a level are trimmed 16 at a time), 8 warps for
// the small first passes of the host pipeline (measured: 96-frame passes 6.8 -> 5.2 ms per 4096 frames, 256-frame passes 3.9 -> 4.3)
constexpr int SEL_WARPS_BATCH = 4, SEL_WARPS_MID = 8, SEL_WARPS_FEW = 16, SEL_FEW_FRAMES = 8;
#ifndef SDORB_SEL_MID_FRAMES
#define SDORB_SEL_MID_FRAMES 128  // passes up to this many frames take 8 warps per CTA: their time is the longest CTA's, not the sum
#endif
constexpr int SEL_WORK_CAP = 1024;  // entries of per-warp shared scratch; larger lists fall back to global memory + one lane

// std::nth_element(a + first, a + nth, a + last, response >) by one warp; indices must stay below 65536.
__device__ void warp_nth_element(uint32_t* a, int first, int nth, int last, int lane, uint16_t* Lidx, uint16_t* Ridx) {
  if (first == last || nth == last) return;
  int n = last - first, lg = 0;
  while (n > 1) {
    n >>= 1;
    ++lg;
  }
  int depth = 2 * lg;
  const uint32_t lt = (1u << lane) - 1u;
  while (last - first > 3) {
    if (depth == 0) {
      if (lane == 0) {
        heap_select(a, first, nth + 1, last);
        swap_u32(a, first, nth);
      }
      __syncwarp();
      return;
    }
    --depth;
    if (lane == 0) {  // __move_median_to_first(first, first+1, mid, last-1)
      const int x = first + 1, y = first + (last - first) / 2, z = last - 1;
      if (resp_gt(a[x], a[y])) {
        if (resp_gt(a[y], a[z]))
          swap_u32(a, first, y);
        else if (resp_gt(a[x], a[z]))
          swap_u32(a, first, z);
        else
          swap_u32(a, first, x);
      } else if (resp_gt(a[x], a[z]))
        swap_u32(a, first, x);
      else if (resp_gt(a[y], a[z]))
        swap_u32(a, first, z);
      else
        swap_u32(a, first, y);
    }
    __syncwarp();
    const uint32_t p = a[first] & 0xFFu;
    // rank the stoppers: left scan (first, last) stops at response <= p, right scan [first, last) at response >= p
    int nl = 0, nr = 0;
    for (int base = first; base < last; base += 32) {
      const int i = base + lane;
      const bool valid = i < last;
      const uint32_t v = valid ? (a[i] & 0xFFu) : 0u;
      const bool is_l = valid && i > first && v <= p;
      const bool is_r = valid && v >= p;
      const uint32_t bl = __ballot_sync(0xffffffffu, is_l), br = __ballot_sync(0xffffffffu, is_r);
      if (is_l) Lidx[nl + __popc(bl & lt)] = (uint16_t)i;
      if (is_r) Ridx[nr + __popc(br & lt)] = (uint16_t)i;  // ascending; R_t = Ridx[nr - 1 - t]
      nl += __popc(bl);
      nr += __popc(br);
    }
    __syncwarp();
    const int m = min(nl, nr);
    int T = 0;  // number of swaps: L_t < R_t is monotone in t
    for (int base = 0; base < m; base += 32) {
      const int t = base + lane;
      const bool ok = t < m && Lidx[t] < Ridx[nr - 1 - t];
      const uint32_t bb = __ballot_sync(0xffffffffu, ok);
      T += __popc(bb);
      if (bb != 0xffffffffu) break;
    }
    for (int t = lane; t < T; t += 32) swap_u32(a, Lidx[t], Ridx[nr - 1 - t]);
    int cut = T < nl ? (int)Lidx[T] : INT_MAX;
    if (T > 0) cut = min(cut, (int)Ridx[nr - T]);
    __syncwarp();
    if (cut <= nth)
      first = cut;
    else
      last = cut;
  }
  if (lane == 0) {  // __insertion_sort(first, last) on at most three elements
    for (int i = first + 1; i < last; ++i) {
      const uint32_t val = a[i];
      if (resp_gt(val, a[first])) {
        for (int k = i; k > first; --k) a[k] = a[k - 1];
        a[first] = val;
      } else {
        int k = i;
        while (resp_gt(val, a[k - 1])) {
          a[k] = a[k - 1];
          --k;
        }
        a[k] = val;
      }
    }
  }
  __syncwarp();
}

struct CellRect {
  int x0, x1, y0, y1;  // detectable rectangle [x0, x1) x [y0, y1) of the cell (empty when the reference skips it)
};

__device__ __forceinline__ CellRect cell_rect(const LevelGeom& L, int ci, int cj) {
  CellRect r;
  r.x0 = SDORB_EDGE + cj * L.cell_w;
  r.x1 = (cj == L.cols - 1) ? L.max_bx : r.x0 + L.cell_w;
  r.y0 = SDORB_EDGE + ci * L.cell_h;
  r.y1 = (ci == L.rows - 1) ? L.max_by : r.y0 + L.cell_h;
  // cv::FAST needs a 7 x 7 ROI: a last column / row narrower than that yields nothing (src/ORBextractor.cc:509-532)
  r.x1 = min(r.x1, L.det_x1);
  r.y1 = min(r.y1, L.det_y1);
  if (r.x1 <= r.x0 || r.y1 <= r.y0) r.x1 = r.x0, r.y1 = r.y0;
  return r;
}

// Walks the cell rectangle of the keypoint map in row-major order and writes SDORB_ENTRY(y, x, score) of every keypoint
// to dst in that order; returns their number.  Item i of the walk is the 16-byte segment (i % nseg) of row (i / nseg);
// a warp takes 32 * U consecutive items per step with all U loads in flight before the first is consumed.
template <int U>
__device__ int scan_cell(const uint8_t* __restrict__ map, int pitch, const CellRect& r, int th, int lane, uint32_t* dst, int cap) {
  if (r.x1 <= r.x0) return 0;
  const int q0 = r.x0 >> 4, nseg = ((r.x1 - 1) >> 4) - q0 + 1;
  const int items = nseg * (r.y1 - r.y0);
  const bool small = items < 65536;
  const uint32_t magic = nseg >= 2 ? 0xFFFFFFFFu / (uint32_t)nseg + 1u : 0u;  // i / nseg == umulhi(i, magic) for i < 65536
  const uint32_t lt = (1u << lane) - 1u;
  int count = 0;
  for (int base = 0; base < items; base += 32 * U) {
    uint4 v[U];
    int xs[U], ys[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + 32 * u + lane;
      v[u] = make_uint4(0u, 0u, 0u, 0u);
      xs[u] = ys[u] = 0;
      if (i < items) {
        const int row = nseg == 1 ? i : (small ? (int)__umulhi((uint32_t)i, magic) : i / nseg);
        xs[u] = (q0 + i - row * nseg) << 4;
        ys[u] = r.y0 + row;
        v[u] = *reinterpret_cast<const uint4*>(map + (int64_t)ys[u] * pitch + xs[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + 32 * u >= items) break;  // warp-uniform
      const uint32_t wv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      // one bit per pixel of the segment: byte != 0 into bit 7 of each byte, the four bits 7 of a word gathered by a multiply
      uint32_t m16 = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t msb = (((wv[j] & 0x7f7f7f7fu) + 0x7f7f7f7fu) | wv[j]) & 0x80808080u;  // bit 7 of a byte: the byte is not 0
        m16 |= ((msb * 0x00204081u) >> 28) << (4 * j);                                        // bits 7, 15, 23, 31 -> 28 .. 31
      }
      // pixels outside [x0, x1) belong to the neighbouring cells
      const int lo = min(max(r.x0 - xs[u], 0), 16), hi = min(max(r.x1 - xs[u], 0), 16);
      m16 &= (0xFFFFu << lo) & ((1u << hi) - 1u);
      const int c = __popc(m16);
      // inside one cell strict 3x3 suppression leaves at most 8 keypoints in 16 consecutive pixels: 4 count bits
      const uint32_t b0 = __ballot_sync(0xffffffffu, c & 1), b1 = __ballot_sync(0xffffffffu, c & 2),
                     b2 = __ballot_sync(0xffffffffu, c & 4), b3 = __ballot_sync(0xffffffffu, c & 8);
      if (b0 | b1 | b2 | b3) {
        int idx = count + __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt) + 8 * __popc(b3 & lt);
        // SDORB_ENTRY(y, x + q, t + th - 1) = the segment's entry + (q << 8) + t: the fields cannot carry into each other
        const uint32_t ebase = ((uint32_t)ys[u] << 20) + ((uint32_t)xs[u] << 8) + (uint32_t)(th - 1);
        for (uint32_t m = m16; m; m &= m - 1, ++idx) {
          const int q = __ffs(m) - 1;
          const uint32_t lo8 = __byte_perm(wv[0], wv[1], q & 7), hi8 = __byte_perm(wv[2], wv[3], q & 7);
          const uint32_t t = ((q & 8) ? hi8 : lo8) & 0xFFu;
          if (idx < cap) dst[idx] = ebase + ((uint32_t)q << 8) + t;
        }
        count += __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2) + 8 * __popc(b3);
      }
    }
  }
  return count;
}

constexpr int GATHER_WARPS = 4;

__global__ void __launch_bounds__(GATHER_WARPS * 32) gather_cells_kernel(const FrameGeom* __restrict__ geom, BatchPlanes p,
                                                                         SelectBuffers buf) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int gcell = blockIdx.x * GATHER_WARPS + (threadIdx.x >> 5);  // cell index over all levels of the frame
  const int frame = blockIdx.y;
  if (gcell >= geom->cells_total) return;
  int level = 0;
  while (level + 1 < geom->nlevels && gcell >= geom->lv[level + 1].cell_base) ++level;
  // levels without cells share the cell_base of the next level: skip forward to the one that owns gcell
  while (!(geom->lv[level].cols > 0 && geom->lv[level].rows > 0)) --level;
  const LevelGeom& L = geom->lv[level];
  const int c = gcell - L.cell_base;
  int32_t* seen = buf.cell_seen + (int64_t)frame * geom->cells_total + gcell;
  if (L.list_cap_cell == 0) {
    if (lane == 0) *seen = 0;
    return;
  }
  const uint8_t* map = p.nms + L.plane_base * p.batch_cap + (int64_t)frame * L.plane_bytes;
  uint32_t* list = buf.cell_list + (int64_t)frame * geom->list_total + L.list_base + (int64_t)c * L.list_cap_cell;
  const CellRect r = cell_rect(L, c / L.cols, c % L.cols);
  int n = scan_cell<4>(map, L.pitch, r, geom->th_fast, lane, list, L.list_cap_cell);
  if (geom->octree && n > 0 && n <= L.list_cap_cell) {
    // ORB-SLAM2-style mode: the cell keeps its iniThFAST keypoints, all of them (minThFAST) only when it has none of those.
    // In-place compaction in order: a step's reads precede its writes and a write never passes its own read index.
    const uint32_t lt = (1u << lane) - 1u;
    int hi = 0;
    __syncwarp();
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const uint32_t e = i < n ? list[i] : 0u;
      const bool keep = i < n && SDORB_ENTRY_SCORE(e) >= geom->ini_th;
      const uint32_t b = __ballot_sync(0xffffffffu, keep);
      if (keep) list[hi + __popc(b & lt)] = e;
      hi += __popc(b);
    }
    if (hi > 0) n = hi;
  }
  if (lane == 0) {
    *seen = n;
    if (n > L.list_cap_cell) atomicExch(buf.error_flag, 6);  // cannot happen: the capacity bounds what strict NMS leaves
  }
}

template <int SEL_WARPS>
__global__ void __launch_bounds__(SEL_WARPS * 32) select_kernel(const FrameGeom* __restrict__ geom, SelectBuffers buf,
                                                                int max_cells, int lvl_cap) {
  constexpr int SEL_THREADS = SEL_WARPS * 32;
  extern __shared__ __align__(16) uint32_t smem[];
  pdl_enter();
  int* n_total = reinterpret_cast<int*>(smem);
  int* n_retain = n_total + max_cells;
  int* offs = n_retain + max_cells;  // max_cells + 1 entries
  uint32_t* lvl = smem + 3 * max_cells + 1;
  uint32_t* work = lvl + lvl_cap;
  uint16_t* idx_scratch = reinterpret_cast<uint16_t*>(work + SEL_WARPS * SEL_WORK_CAP);
  __shared__ int s_total;

  // level-major grid: the CTAs of level 0 (the longest lists) are dispatched first, the short top levels fill the tail
  const int level = blockIdx.y, frame = blockIdx.x;
  const LevelGeom& L = geom->lv[level];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_cells = (L.cols > 0 && L.rows > 0) ? L.cols * L.rows : 0;
  int32_t* out_count = buf.sel_count + (int64_t)frame * geom->nlevels + level;
  if (n_cells == 0 || L.list_cap_cell == 0) {
    if (tid == 0) *out_count = 0;
    return;
  }
  const int32_t* seen = buf.cell_seen + (int64_t)frame * geom->cells_total + L.cell_base;
  uint32_t* lists = buf.cell_list + (int64_t)frame * geom->list_total + L.list_base;

  // keypoints per cell, counted by gather_cells_kernel
  for (int c = tid; c < n_cells; c += SEL_THREADS) n_total[c] = min(seen[c], L.list_cap_cell);
  __syncthreads();

  if (tid == 0) {
    // quota with redistribution, literally as in the reference (including cells its loops `continue` past:
    // they stay open with zero keypoints and are closed by the first redistribution pass)
    const int quota = L.n_features_cell;
    int n_no_more = 0, n_distribute = 0;
    for (int i = 0; i < L.rows; ++i)
      for (int j = 0; j < L.cols; ++j) {
        const int c = i * L.cols + j;
        const bool skipped = (i == L.rows - 1 && L.last_skipped_y) || (j == L.cols - 1 && L.last_skipped_x);
        if (skipped) {
          n_retain[c] = 0;  // bNoMore stays false
          continue;
        }
        const int n = n_total[c];
        if (n > quota) {
          n_retain[c] = quota;
        } else {
          n_retain[c] = n;
          n_distribute += quota - n;
          n_retain[c] |= 0x40000000;  // bNoMore
          ++n_no_more;
        }
      }
    while (n_distribute > 0 && n_no_more < n_cells) {
      const int q = quota + (int)ceilf((float)n_distribute / (float)(n_cells - n_no_more));
      n_distribute = 0;
      for (int c = 0; c < n_cells; ++c) {
        if (n_retain[c] & 0x40000000) continue;
        const int n = n_total[c];
        if (n > q) {
          n_retain[c] = q;
        } else {
          n_retain[c] = n | 0x40000000;
          n_distribute += q - n;
          ++n_no_more;
        }
      }
    }
    int acc = 0;
    for (int c = 0; c < n_cells; ++c) {
      n_retain[c] &= 0x3FFFFFFF;
      offs[c] = acc;
      acc += n_retain[c];
    }
    offs[n_cells] = acc;
    s_total = acc;
    if (acc > lvl_cap) atomicExch(buf.error_flag, 6);
  }
  __syncthreads();

  // pass 2: gather each cell in emission order, retainBest, append the survivors to the level list
  uint32_t* wbuf = work + warp * SEL_WORK_CAP;
  uint16_t* Lidx = idx_scratch + warp * 2 * SEL_WORK_CAP;
  uint16_t* Ridx = Lidx + SEL_WORK_CAP;
  for (int c = warp; c < n_cells; c += SEL_WARPS) {
    const int n = n_total[c], r = n_retain[c];
    if (r == 0) continue;
    const bool in_smem = n <= SEL_WORK_CAP;
    uint32_t* list = lists + (int64_t)c * L.list_cap_cell;
    uint32_t* a = list;
    if (in_smem) {
      a = wbuf;
      const int m = n > r ? n : r;  // an untrimmed cell only needs its first r (== n) entries
      for (int i = lane; i < m; i += 32) a[i] = list[i];
      __syncwarp();
    }
    if (n > r) {
      if (in_smem) {
        warp_nth_element(a, 0, r - 1, n, lane, Lidx, Ridx);
      } else {
        if (lane == 0) nth_element_resp(a, 0, r - 1, n);
        __syncwarp();
      }
    }
    const int o = offs[c];
    for (int i = lane; i < r; i += 32)
      if (o + i < lvl_cap) lvl[o + i] = a[i];
    __syncwarp();
  }
  __syncthreads();

  // per-level retainBest
  const int total = min(s_total, lvl_cap);
  int keep = total;
  if (total > L.n_desired) {
    keep = L.n_desired;
    if (keep > 0 && warp == 0) {
      if (total <= SEL_WORK_CAP) {
        warp_nth_element(lvl, 0, keep - 1, total, lane, Lidx, Ridx);
      } else if (lane == 0) {
        nth_element_resp(lvl, 0, keep - 1, total);
      }
    }
    __syncthreads();
  }
  uint32_t* sel = buf.sel + (int64_t)frame * geom->sel_total + L.sel_base;
  for (int i = tid; i < keep; i += SEL_THREADS) sel[i] = lvl[i];
  if (tid == 0) *out_count = keep;
}

static void select_caps(const FrameGeom& g, int* max_cells, int* lvl_cap) {
  int mc = 1, lc = 64;
  for (int l = 0; l < g.nlevels; ++l) {
    const LevelGeom& L = g.lv[l];
    const int nc = (L.cols > 0 && L.rows > 0) ? L.cols * L.rows : 0;
    mc = nc > mc ? nc : mc;
    const int c = 2 * (L.n_desired > 0 ? L.n_desired : 0) + 2 * nc + 64;
    lc = c > lc ? c : lc;
  }
  *max_cells = mc;
  *lvl_cap = lc;
}

static size_t select_smem_bytes_w(const FrameGeom& g, int warps) {
  int mc, lc;
  select_caps(g, &mc, &lc);
  return sizeof(uint32_t) * ((size_t)3 * mc + 1 + lc + (size_t)warps * SEL_WORK_CAP) +
         sizeof(uint16_t) * (size_t)warps * 2 * SEL_WORK_CAP + 16;
}
size_t select_smem_bytes(const FrameGeom& g) { return select_smem_bytes_w(g, SEL_WARPS_FEW); }

void launch_select(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, const SelectBuffers& b, int nframes,
                   cudaStream_t s) {
  int mc, lc;
  select_caps(g, &mc, &lc);
  if (g.cells_total > 0)
    launch_pdl(gather_cells_kernel, dim3((g.cells_total + GATHER_WARPS - 1) / GATHER_WARPS, nframes), dim3(GATHER_WARPS * 32), 0, s, d_geom, p, b);
  if (g.octree) {
    launch_octree(d_geom, g, b, nframes, s);
    return;
  }
  if (nframes <= SEL_FEW_FRAMES && select_smem_bytes_w(g, SEL_WARPS_FEW) <= 200 * 1024)
    launch_pdl(select_kernel<SEL_WARPS_FEW>, dim3(nframes, g.nlevels), dim3(SEL_WARPS_FEW * 32), select_smem_bytes_w(g, SEL_WARPS_FEW), s,
               d_geom, b, mc, lc);
  else if (nframes <= SDORB_SEL_MID_FRAMES && select_smem_bytes_w(g, SEL_WARPS_MID) <= 200 * 1024)
    launch_pdl(select_kernel<SEL_WARPS_MID>, dim3(nframes, g.nlevels), dim3(SEL_WARPS_MID * 32), select_smem_bytes_w(g, SEL_WARPS_MID), s,
               d_geom, b, mc, lc);
  else
    launch_pdl(select_kernel<SEL_WARPS_BATCH>, dim3(nframes, g.nlevels), dim3(SEL_WARPS_BATCH * 32), select_smem_bytes_w(g, SEL_WARPS_BATCH), s,
               d_geom, b, mc, lc);
}

// ---- test hook
__global__ void __launch_bounds__(32) debug_nth_element_kernel(uint32_t* entries, int n, int nth) {
  __shared__ uint32_t a[SEL_WORK_CAP];
  __shared__ uint16_t li[SEL_WORK_CAP], ri[SEL_WORK_CAP];
  const int lane = threadIdx.x;
  if (n > SEL_WORK_CAP) {
    if (lane == 0) nth_element_resp(entries, 0, nth, n);
    return;
  }
  for (int i = lane; i < n; i += 32) a[i] = entries[i];
  __syncwarp();
  warp_nth_element(a, 0, nth, n, lane, li, ri);
  for (int i = lane; i < n; i += 32) entries[i] = a[i];
}

void launch_debug_nth_element(uint32_t* d_entries, int n, int nth, cudaStream_t s) {
  debug_nth_element_kernel<<<1, 32, 0, s>>>(d_entries, n, nth);
}

int configure_kernels() {
  cudaError_t e = cudaFuncSetAttribute(select_kernel<SEL_WARPS_BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(select_kernel<SEL_WARPS_MID>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(select_kernel<SEL_WARPS_FEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  return (int)e;
}

}  // namespace sdorb
