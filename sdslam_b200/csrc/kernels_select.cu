// kernels_select.cu -- keypoint culling of ComputeKeyPoints for sm_100a.
//
// Reference: /root/reference/src/ORBextractor.cc:538-604.  Per level:
//   1. per-cell quota nfeaturesCell with iterative redistribution of the deficit of poor cells (:538-575);
//   2. per cell  KeyPointsFilter::retainBest(cell, nToRetain) + resize(nToRetain)  (:583-588);
//   3. cells concatenated row-major, then  retainBest(level, nDesired) + resize  (:601-604).
// retainBest's survivors and their ORDER are defined by libstdc++'s nth_element (see introselect.cuh); the
// order is part of the output (keypoints and descriptor rows follow it), so it is reproduced move for move.
// nth_element is inherently sequential per list, but there are cells x levels x frames independent lists:
// one CTA handles one (level, frame); its warps take the cells, each sorting its cell list back into FAST's
// emission order (bitonic network on the unique (y,x) keys) before one lane replays the introselect.
#include "introselect.cuh"
#include "kernels.cuh"

namespace sdorb {

constexpr int SEL_THREADS = 256;
constexpr int SEL_WARPS = SEL_THREADS / 32;
constexpr int SEL_WORK_CAP = 1024;  // entries of per-warp shared scratch; larger cell lists are handled in place in global memory

__device__ __forceinline__ void cmp_swap(uint32_t* a, int i, int j) {
  const uint32_t x = a[i], y = a[j];
  if (x > y) {
    a[i] = y;
    a[j] = x;
  }
}

// Ascending bitonic sort of a[0..n) by one warp; slots >= n act as +infinity and are never touched.
__device__ void warp_sort(uint32_t* a, int n, int lane) {
  if (n < 2) return;
  int P = 2;
  while (P < n) P <<= 1;
  for (int k = 2; k <= P; k <<= 1) {
    const int half = k >> 1;
    for (int t = lane; t < (P >> 1); t += 32) {
      const int blk = t / half, off = t - blk * half;
      const int i = blk * k + off, j = blk * k + (k - 1 - off);
      if (j < n) cmp_swap(a, i, j);
    }
    __syncwarp();
    for (int s = k >> 2; s > 0; s >>= 1) {
      for (int t = lane; t < (P >> 1); t += 32) {
        const int i = ((t & ~(s - 1)) << 1) | (t & (s - 1)), j = i | s;
        if (j < n) cmp_swap(a, i, j);
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(SEL_THREADS) select_kernel(const FrameGeom* __restrict__ geom, SelectBuffers buf,
                                                             int max_cells, int lvl_cap) {
  extern __shared__ __align__(16) uint32_t smem[];
  int* n_total = reinterpret_cast<int*>(smem);
  int* n_retain = n_total + max_cells;
  int* offs = n_retain + max_cells;  // max_cells + 1 entries
  uint32_t* lvl = smem + 3 * max_cells + 1;
  uint32_t* work = lvl + lvl_cap;
  __shared__ int s_total;

  const int level = blockIdx.x, frame = blockIdx.y;
  const LevelGeom& L = geom->lv[level];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_cells = (L.cols > 0 && L.rows > 0) ? L.cols * L.rows : 0;
  int32_t* out_count = buf.sel_count + (int64_t)frame * geom->nlevels + level;
  if (n_cells == 0 || L.list_cap_cell == 0) {
    if (tid == 0) *out_count = 0;
    return;
  }
  int32_t* cnt = buf.cell_count + (int64_t)frame * geom->cells_total + L.cell_base;
  int32_t* seen = buf.cell_seen + (int64_t)frame * geom->cells_total + L.cell_base;
  uint32_t* lists = buf.cell_list + (int64_t)frame * geom->list_total + L.list_base;
  for (int c = tid; c < n_cells; c += SEL_THREADS) {
    const int n = cnt[c];
    n_total[c] = min(n, L.list_cap_cell);
    seen[c] = n;  // what FAST found (parity tests read this)
    cnt[c] = 0;   // re-arm the append counters for the next pass: the FAST stage stays a single kernel
  }
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

  // per-cell retainBest
  uint32_t* wbuf = work + warp * SEL_WORK_CAP;
  for (int c = warp; c < n_cells; c += SEL_WARPS) {
    const int n = n_total[c], r = n_retain[c];
    if (r == 0) continue;
    uint32_t* list = lists + (int64_t)c * L.list_cap_cell;
    uint32_t* a = list;
    if (n <= SEL_WORK_CAP) {
      a = wbuf;
      for (int i = lane; i < n; i += 32) a[i] = list[i];
      __syncwarp();
    }
    warp_sort(a, n, lane);
    if (n > r) {
      if (lane == 0) nth_element_resp(a, 0, r - 1, n);
      __syncwarp();
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
    if (tid == 0 && keep > 0) nth_element_resp(lvl, 0, keep - 1, total);
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

size_t select_smem_bytes(const FrameGeom& g) {
  int mc, lc;
  select_caps(g, &mc, &lc);
  return sizeof(uint32_t) * ((size_t)3 * mc + 1 + lc + (size_t)SEL_WARPS * SEL_WORK_CAP);
}

void launch_select(const FrameGeom* d_geom, const FrameGeom& g, const SelectBuffers& b, int nframes, cudaStream_t s) {
  int mc, lc;
  select_caps(g, &mc, &lc);
  select_kernel<<<dim3(g.nlevels, nframes), SEL_THREADS, select_smem_bytes(g), s>>>(d_geom, b, mc, lc);
}

int configure_kernels() {
  return (int)cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

}  // namespace sdorb
