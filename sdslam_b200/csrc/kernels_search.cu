// kernels_search.cu -- the guided matchers either side of DescriptorDistance, for sm_100a.
//
// Reference: /root/reference/src/ORBmatcher.cc
//   SearchForInitialization              :256-357
//   SearchByProjection(Frame&, Frame&)   :946-1075   (from the projection (u, v, invzc) on; the pose algebra stays with the caller)
//   ComputeThreeMaxima                   :1423-1454
// and the window query both of them run per keypoint, Frame::GetFeaturesInArea, src/Frame.cc:271-321, over the CSR grid
// that sdorb_assign_grid_batch leaves on the device.
//
// Both matchers are greedy over the query keypoints IN ORDER: an accepted match changes what later queries may take
// (vMatchedDistance / vnMatches21, mvpMapPoints of the current frame).  That order is the result, so one warp owns one
// frame pair and walks its queries sequentially; the parallelism is the batch (thousands of pairs in flight) and, inside
// a query, the candidates of the window: the grid columns of the window are contiguous index spans of the CSR grid, the
// lanes flatten them with a prefix sum and take candidates k, k+32, ... ; (distance << 16 | k) min-reduced over the warp
// keeps the reference's "first minimum in candidate order" rule, the runner-up is merged alongside.
#include "kernels.cuh"
#include "match_common.cuh"

namespace sdorb {

constexpr int GRID_COLS = 64, GRID_ROWS = 48;
constexpr int HISTO_LENGTH = 30;  // src/ORBmatcher.cc:38
constexpr int DIST_NONE = 0x7FFF;  // stands for INT_MAX in the packed (distance << 16 | candidate) words

struct KP {  // cv::KeyPoint, 28 bytes
  float x, y, size, angle, response;
  int32_t octave, class_id;
};

// Frame::GetFeaturesInArea's cell window (src/Frame.cc:275-291); false = the reference returns an empty vector
__device__ __forceinline__ bool area_cells(float x, float y, float r, const SearchGrid& g, int& x0, int& x1, int& y0, int& y1) {
  x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, g.min_x), r), g.inv_w)));
  if (x0 >= GRID_COLS) return false;
  x1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, g.min_x), r), g.inv_w)));
  if (x1 < 0) return false;
  y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, g.min_y), r), g.inv_h)));
  if (y0 >= GRID_ROWS) return false;
  y1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, g.min_y), r), g.inv_h)));
  if (y1 < 0) return false;
  return true;
}

// The window's candidates in the reference's order are, per grid column ix, the contiguous span
// indices[cell_start[ix*48 + y0] .. cell_start[ix*48 + y1 + 1]).  Fills s_begin[c] / s_prefix[c] for the columns and returns the total.
__device__ __forceinline__ int window_spans(const int32_t* __restrict__ cs, int x0, int x1, int y0, int y1, int lane, int* s_begin,
                                            int* s_prefix) {
  const int ncols = x1 - x0 + 1;
  int total = 0;
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    const int c = c0 + lane;
    int b = 0, len = 0;
    if (c < ncols) {
      b = cs[(x0 + c) * GRID_ROWS + y0];
      len = cs[(x0 + c) * GRID_ROWS + y1 + 1] - b;
    }
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (c < ncols) {
      s_begin[c] = b;
      s_prefix[c] = total + incl - len;
    }
    total += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) s_prefix[ncols] = total;
  __syncwarp();
  return total;
}

// position in `indices` of candidate k
__device__ __forceinline__ int candidate_slot(int k, int ncols, const int* s_begin, const int* s_prefix) {
  int lo = 0, hi = ncols - 1;  // largest c with s_prefix[c] <= k
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (s_prefix[mid] <= k) lo = mid;
    else hi = mid - 1;
  }
  return s_begin[lo] + (k - s_prefix[lo]);
}

// the two smallest distances of the warp's candidates; first candidate on ties
__device__ __forceinline__ void warp_best2(int& best1, int& best2) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int ob1 = __shfl_xor_sync(0xffffffffu, best1, o), ob2 = __shfl_xor_sync(0xffffffffu, best2, o);
    const int lo1 = min(best1, ob1), hi1 = max(best1, ob1);
    best2 = min(min(best2, ob2), hi1 >> 16);
    best1 = lo1;
  }
}

__device__ __forceinline__ int rotation_bin(float a1, float a2) {  // src/ORBmatcher.cc:316-322
  const float factor = 1.0f / HISTO_LENGTH;
  float rot = __fsub_rn(a1, a2);
  if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
  int bin = (int)roundf(__fmul_rn(rot, factor));
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

// ComputeThreeMaxima, src/ORBmatcher.cc:1423-1454; returns a 30-bit mask of the bins that survive
__device__ __forceinline__ uint32_t three_maxima_mask(const int* histo) {
  int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
  for (int i = 0; i < HISTO_LENGTH; i++) {
    const int s = histo[i];
    if (s > max1) {
      max3 = max2;
      max2 = max1;
      max1 = s;
      ind3 = ind2;
      ind2 = ind1;
      ind1 = i;
    } else if (s > max2) {
      max3 = max2;
      max2 = s;
      ind3 = ind2;
      ind2 = i;
    } else if (s > max3) {
      max3 = s;
      ind3 = i;
    }
  }
  if ((float)max2 < __fmul_rn(0.1f, (float)max1)) {
    ind2 = -1;
    ind3 = -1;
  } else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) {
    ind3 = -1;
  }
  uint32_t m = 0;
  if (ind1 >= 0) m |= 1u << ind1;
  if (ind2 >= 0) m |= 1u << ind2;
  if (ind3 >= 0) m |= 1u << ind3;
  return m;
}

__device__ __forceinline__ void load_desc(const uint8_t* d, uint32_t (&q)[8]) {
  const uint4 lo = reinterpret_cast<const uint4*>(d)[0], hi = reinterpret_cast<const uint4*>(d)[1];
  q[0] = lo.x; q[1] = lo.y; q[2] = lo.z; q[3] = lo.w;
  q[4] = hi.x; q[5] = hi.y; q[6] = hi.z; q[7] = hi.w;
}

// ---------------------------------------------------------------------------------------- SearchForInitialization
// shared memory per pair: vMatchedDistance[cap] (int), vnMatches21[cap] (int), rotation bin of every query (int8), spans
__global__ void __launch_bounds__(32) search_init_kernel(SearchInitArgs a) {
  extern __shared__ int s_mem[];
  const int pair = blockIdx.x, lane = threadIdx.x, cap = a.capacity;
  int* s_dist = s_mem;              // vMatchedDistance
  int* s_m21 = s_mem + cap;         // vnMatches21
  int* s_begin = s_m21 + cap;       // [64]
  int* s_prefix = s_begin + GRID_COLS;  // [65]
  int* s_histo = s_prefix + GRID_COLS + 1;  // [30] (+2 pad)
  int8_t* s_bin = reinterpret_cast<int8_t*>(s_histo + 32);  // [cap]
  const int n1 = min(a.n1[pair], cap), n2 = min(a.n2[pair], cap);
  const KP* k1 = reinterpret_cast<const KP*>(a.kps1) + (int64_t)pair * cap;
  const KP* k2 = reinterpret_cast<const KP*>(a.kps2) + (int64_t)pair * cap;
  const uint8_t* d1 = a.desc1 + (int64_t)pair * cap * 32;
  const uint8_t* d2 = a.desc2 + (int64_t)pair * cap * 32;
  const int32_t* cs = a.grid.cell_start + (int64_t)pair * (GRID_COLS * GRID_ROWS + 1);
  const int32_t* idx = a.grid.indices + (int64_t)pair * cap;
  float* prev = a.prev_matched + (int64_t)pair * cap * 2;
  int32_t* m12 = a.matches12 + (int64_t)pair * cap;
  for (int i = lane; i < n2; i += 32) {
    s_dist[i] = INT_MAX;
    s_m21[i] = -1;
  }
  for (int i = lane; i < cap; i += 32) {
    m12[i] = -1;
    if (i < n1) s_bin[i] = -1;
  }
  if (lane < 32) s_histo[lane] = 0;
  __syncwarp();
  int nmatches = 0;  // lane 0's copy counts
  const float r = (float)a.window_size;
  for (int i1 = 0; i1 < n1; ++i1) {
    const int level1 = k1[i1].octave;
    if (level1 > 0) continue;
    const float x = prev[2 * i1], y = prev[2 * i1 + 1];
    int x0, x1, y0, y1;
    if (!area_cells(x, y, r, a.grid, x0, x1, y0, y1)) continue;
    const int total = window_spans(cs, x0, x1, y0, y1, lane, s_begin, s_prefix);
    uint32_t q[8];
    load_desc(d1 + (int64_t)i1 * 32, q);
    int best1 = (DIST_NONE << 16) | 0xFFFF, best2 = DIST_NONE;
    for (int k = lane; k < total; k += 32) {
      const int i2 = idx[candidate_slot(k, x1 - x0 + 1, s_begin, s_prefix)];
      const KP kp = k2[i2];
      // bCheckLevels is true (maxLevel = level1 >= 0): octave < level1 or > level1 is skipped (level1 <= 0 here)
      if (kp.octave < level1 || kp.octave > level1) continue;
      if (!(fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r)) continue;
      uint32_t t[8];
      load_desc(d2 + (int64_t)i2 * 32, t);
      const uint4 lo = make_uint4(t[0], t[1], t[2], t[3]), hi = make_uint4(t[4], t[5], t[6], t[7]);
      const int d = hamming256(q, lo, hi);
      if (s_dist[i2] <= d) continue;
      best2 = min(best2, max(d, best1 >> 16));
      best1 = min(best1, (d << 16) | k);
    }
    warp_best2(best1, best2);
    const int bestDist = best1 >> 16;
    if (bestDist <= a.th_low) {  // also false when nothing was found (DIST_NONE)
      const float second = best2 == DIST_NONE ? (float)INT_MAX : (float)best2;
      if ((float)bestDist < __fmul_rn(second, a.nnratio)) {
        if (lane == 0) {
          const int bestIdx2 = idx[candidate_slot(best1 & 0xFFFF, x1 - x0 + 1, s_begin, s_prefix)];
          const int old = s_m21[bestIdx2];
          if (old >= 0) {
            m12[old] = -1;
            nmatches--;
          }
          m12[i1] = bestIdx2;
          s_m21[bestIdx2] = i1;
          s_dist[bestIdx2] = bestDist;
          nmatches++;
          if (a.check_orientation) {
            const int bin = rotation_bin(k1[i1].angle, k2[bestIdx2].angle);
            s_bin[i1] = (int8_t)bin;
            s_histo[bin]++;
          }
        }
      }
    }
    __syncwarp();
  }
  if (a.check_orientation) {
    const uint32_t keep = three_maxima_mask(s_histo);
    int removed = 0;
    for (int i = lane; i < n1; i += 32) {
      const int bin = s_bin[i];
      if (bin >= 0 && !((keep >> bin) & 1u) && m12[i] >= 0) {
        m12[i] = -1;
        removed++;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
    nmatches -= removed;
    __syncwarp();
  }
  for (int i = lane; i < n1; i += 32) {  // :349-352
    const int m = m12[i];
    if (m >= 0) {
      prev[2 * i] = k2[m].x;
      prev[2 * i + 1] = k2[m].y;
    }
  }
  if (lane == 0) a.nmatches[pair] = nmatches;
}

size_t search_init_smem(int capacity) {
  return sizeof(int) * ((size_t)2 * capacity + GRID_COLS + GRID_COLS + 1 + 32) + (size_t)capacity + 16;
}

// ---------------------------------------------------------------------------------------- SearchByProjection (Frame, Frame)
// shared memory per pair: occupied[cap] (uint8), push bin / push index of every last-frame keypoint, spans
__global__ void __launch_bounds__(32) search_projection_kernel(SearchProjArgs a) {
  extern __shared__ int s_mem[];
  const int pair = blockIdx.x, lane = threadIdx.x, cap = a.capacity;
  int* s_push_idx = s_mem;               // [cap] bestIdx2 pushed into rotHist by last-frame keypoint i
  int* s_begin = s_mem + cap;            // [64]
  int* s_prefix = s_begin + GRID_COLS;   // [65]
  int* s_histo = s_prefix + GRID_COLS + 1;   // [32]
  int8_t* s_bin = reinterpret_cast<int8_t*>(s_histo + 32);  // [cap]
  uint8_t* s_occ = reinterpret_cast<uint8_t*>(s_bin + cap);  // [cap]
  const int nL = min(a.n_last[pair], cap), nC = min(a.n_cur[pair], cap);
  const KP* kL = reinterpret_cast<const KP*>(a.kps_last) + (int64_t)pair * cap;
  const KP* kLun = reinterpret_cast<const KP*>(a.kps_last_un) + (int64_t)pair * cap;
  const KP* kC = reinterpret_cast<const KP*>(a.kps_cur_un) + (int64_t)pair * cap;
  const float* proj = a.proj + (int64_t)pair * cap * 3;
  const uint8_t* fl = a.flags_last + (int64_t)pair * cap;
  const uint8_t* dMP = a.desc_mp + (int64_t)pair * cap * 32;
  const uint8_t* dC = a.desc_cur + (int64_t)pair * cap * 32;
  const float* uR = a.u_right_cur + (int64_t)pair * cap;
  const int32_t* cs = a.grid.cell_start + (int64_t)pair * (GRID_COLS * GRID_ROWS + 1);
  const int32_t* idx = a.grid.indices + (int64_t)pair * cap;
  int32_t* assigned = a.assigned + (int64_t)pair * cap;
  for (int i = lane; i < cap; i += 32) {
    assigned[i] = -1;
    s_bin[i] = -1;
    s_occ[i] = i < nC ? a.occupied_cur[(int64_t)pair * cap + i] : 0;
  }
  s_histo[lane] = 0;
  __syncwarp();
  int nmatches = 0;
  for (int i = 0; i < nL; ++i) {
    const int flags = fl[i];
    if (!(flags & 1)) continue;
    const float u = proj[3 * i], v = proj[3 * i + 1], invzc = proj[3 * i + 2];
    if (invzc < 0.f) continue;
    if (u < a.min_x || u > a.max_x) continue;
    if (v < a.min_y || v > a.max_y) continue;
    const int oct = kL[i].octave;
    const float radius = __fmul_rn(a.th, a.scale_factors[min(max(oct, 0), SDORB_MAX_LEVELS - 1)]);
    int minLevel, maxLevel;
    if (a.mode == 1) {
      minLevel = oct;
      maxLevel = -1;
    } else if (a.mode == 2) {
      minLevel = 0;
      maxLevel = oct;
    } else {
      minLevel = oct - 1;
      maxLevel = oct + 1;
    }
    int x0, x1, y0, y1;
    if (!area_cells(u, v, radius, a.grid, x0, x1, y0, y1)) continue;
    const int total = window_spans(cs, x0, x1, y0, y1, lane, s_begin, s_prefix);
    const bool check_levels = minLevel > 0 || maxLevel >= 0;
    uint32_t q[8];
    load_desc(dMP + (int64_t)i * 32, q);
    const float ur = __fmaf_rn(-a.mbf, invzc, u);  // u - mbf*invzc, contracted by the reference build (see the oracle)
    int best1 = (256 << 16) | 0xFFFF;
    for (int k = lane; k < total; k += 32) {
      const int i2 = idx[candidate_slot(k, x1 - x0 + 1, s_begin, s_prefix)];
      const KP kp = kC[i2];
      if (check_levels) {
        if (kp.octave < minLevel) continue;
        if (maxLevel >= 0 && kp.octave > maxLevel) continue;
      }
      if (!(fabsf(__fsub_rn(kp.x, u)) < radius && fabsf(__fsub_rn(kp.y, v)) < radius)) continue;
      if (s_occ[i2]) continue;
      const float r2 = uR[i2];
      if (r2 > 0.f) {
        if (fabsf(__fsub_rn(ur, r2)) > radius) continue;
      }
      uint32_t t[8];
      load_desc(dC + (int64_t)i2 * 32, t);
      const uint4 lo = make_uint4(t[0], t[1], t[2], t[3]), hi = make_uint4(t[4], t[5], t[6], t[7]);
      best1 = min(best1, (hamming256(q, lo, hi) << 16) | k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best1 = min(best1, __shfl_xor_sync(0xffffffffu, best1, o));
    const int bestDist = best1 >> 16;
    if (bestDist <= a.th_high && bestDist < 256) {
      if (lane == 0) {
        const int bestIdx2 = idx[candidate_slot(best1 & 0xFFFF, x1 - x0 + 1, s_begin, s_prefix)];
        assigned[bestIdx2] = i;
        s_occ[bestIdx2] = (flags & 2) ? 1 : 0;
        nmatches++;
        if (a.check_orientation) {
          const int bin = rotation_bin(kLun[i].angle, kC[bestIdx2].angle);
          s_bin[i] = (int8_t)bin;
          s_push_idx[i] = bestIdx2;
          s_histo[bin]++;
        }
      }
    }
    __syncwarp();
  }
  if (a.check_orientation) {
    const uint32_t keep = three_maxima_mask(s_histo);
    int removed = 0;
    for (int i = lane; i < nL; i += 32) {
      const int bin = s_bin[i];
      if (bin >= 0 && !((keep >> bin) & 1u)) {
        assigned[s_push_idx[i]] = -1;  // every push of a removed bin clears its slot and is counted (:1064-1067)
        removed++;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
    nmatches -= removed;
  }
  if (lane == 0) a.nmatches[pair] = nmatches;
}

size_t search_projection_smem(int capacity) {
  return sizeof(int) * ((size_t)capacity + GRID_COLS + GRID_COLS + 1 + 32) + (size_t)2 * capacity + 16;
}

// ---------------------------------------------------------------------------------------- SearchByProjection (Frame, map points)
// src/ORBmatcher.cc:43-119, the local-map search of Tracking::SearchLocalPoints: the map points are visited in order and an
// accepted match occupies its keypoint for the later ones, so one warp owns a frame; per map point the window's candidates
// are taken 32 at a time.  Best and second-best (with their pyramid levels) are the two smallest packed words
// (distance << 16 | candidate): the sequential update rule of :87-97 is exactly that order.
__device__ __forceinline__ void warp_min2(uint32_t& m1, uint32_t& m2) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t a1 = __shfl_xor_sync(0xffffffffu, m1, o), a2 = __shfl_xor_sync(0xffffffffu, m2, o);
    const uint32_t lo = min(m1, a1), hi = max(m1, a1);
    m2 = min(hi, min(m2, a2));
    m1 = lo;
  }
}

__global__ void __launch_bounds__(32) search_points_kernel(SearchPointsArgs a) {
  extern __shared__ int s_mem[];
  const int frame = blockIdx.x, lane = threadIdx.x, cap = a.capacity, capmp = a.capacity_mp;
  int* s_begin = s_mem;                   // [64]
  int* s_prefix = s_begin + GRID_COLS;    // [65]
  uint8_t* s_occ = reinterpret_cast<uint8_t*>(s_prefix + GRID_COLS + 1 + 1);  // [cap]
  const int nMP = min(a.n_mp[frame], capmp), nF = min(a.n_frame[frame], cap);
  const float* proj = a.proj + (int64_t)frame * capmp * 3;
  const float* vcos = a.view_cos + (int64_t)frame * capmp;  // not read in the Sim3 form (may be null)
  const int32_t* lvl = a.level + (int64_t)frame * capmp;
  const uint8_t* fl = a.flags + (int64_t)frame * capmp;
  const uint8_t* dMP = a.desc_mp + (int64_t)frame * capmp * 32;
  const KP* kF = reinterpret_cast<const KP*>(a.kps) + (int64_t)frame * cap;
  const uint8_t* dF = a.desc + (int64_t)frame * cap * 32;
  const float* uR = a.u_right + (int64_t)frame * cap;
  const int32_t* cs = a.grid.cell_start + (int64_t)frame * (GRID_COLS * GRID_ROWS + 1);
  const int32_t* idx = a.grid.indices + (int64_t)frame * cap;
  int32_t* assigned = a.assigned + (int64_t)frame * cap;
  for (int i = lane; i < cap; i += 32) {
    assigned[i] = -1;
    s_occ[i] = i < nF ? a.occupied[(int64_t)frame * cap + i] : 0;
  }
  __syncwarp();
  int nmatches = 0;
  const bool factor = (double)a.th != 1.0;
  for (int i = 0; i < nMP; ++i) {
    const int flags = fl[i];
    if (!(flags & 1)) continue;
    const int level = lvl[i];
    float r = a.th;  // the Sim3 form, :213: th * mvScaleFactors[nPredictedLevel]
    if (!a.sim3_form) {
      r = (double)vcos[i] > 0.998 ? 2.5f : 4.0f;  // RadiusByViewingCos, :121-126
      if (factor) r = __fmul_rn(r, a.th);
    }
    const float radius = __fmul_rn(r, a.scale_factors[min(max(level, 0), SDORB_MAX_LEVELS - 1)]);
    const float x = proj[3 * i], y = proj[3 * i + 1], xr = proj[3 * i + 2];
    int x0, x1, y0, y1;
    if (!area_cells(x, y, radius, a.grid, x0, x1, y0, y1)) continue;
    const int total = window_spans(cs, x0, x1, y0, y1, lane, s_begin, s_prefix);
    const int minLevel = level - 1, maxLevel = level;
    const bool check_levels = minLevel > 0 || maxLevel >= 0;
    uint32_t q[8];
    load_desc(dMP + (int64_t)i * 32, q);
    uint32_t m1 = (256u << 16) | 0xFFFFu, m2 = (256u << 16) | 0xFFFFu;
    for (int k = lane; k < total; k += 32) {
      const int i2 = idx[candidate_slot(k, x1 - x0 + 1, s_begin, s_prefix)];
      const KP kp = kF[i2];
      if (check_levels) {
        if (kp.octave < minLevel) continue;
        if (maxLevel >= 0 && kp.octave > maxLevel) continue;
      }
      if (!(fabsf(__fsub_rn(kp.x, x)) < radius && fabsf(__fsub_rn(kp.y, y)) < radius)) continue;
      if (s_occ[i2]) continue;
      if (!a.sim3_form) {
        const float r2 = uR[i2];
        if (r2 > 0.f && fabsf(__fsub_rn(xr, r2)) > radius) continue;
      }
      uint32_t t[8];
      load_desc(dF + (int64_t)i2 * 32, t);
      const uint4 lo = make_uint4(t[0], t[1], t[2], t[3]), hi = make_uint4(t[4], t[5], t[6], t[7]);
      const uint32_t v = ((uint32_t)hamming256(q, lo, hi) << 16) | (uint32_t)k;
      m2 = min(m2, max(m1, v));
      m1 = min(m1, v);
    }
    warp_min2(m1, m2);
    const int bestDist = (int)(m1 >> 16), bestDist2 = (int)(m2 >> 16);
    if (bestDist <= a.th_high && bestDist < 256) {
      if (lane == 0) {
        const int bestIdx = idx[candidate_slot((int)(m1 & 0xFFFFu), x1 - x0 + 1, s_begin, s_prefix)];
        const int bestLevel = kF[bestIdx].octave;
        int bestLevel2 = -1;
        if (bestDist2 < 256) bestLevel2 = kF[idx[candidate_slot((int)(m2 & 0xFFFFu), x1 - x0 + 1, s_begin, s_prefix)]].octave;
        const bool reject = !a.sim3_form && bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(a.nnratio, (float)bestDist2);
        if (!reject) {
          assigned[bestIdx] = i;
          s_occ[bestIdx] = (a.sim3_form || (flags & 2)) ? 1 : 0;  // :246: vpMatched[bestIdx] = pMP
          nmatches++;
        }
      }
    }
    __syncwarp();
  }
  if (lane == 0) a.nmatches[frame] = nmatches;
}

size_t search_points_smem(int capacity) { return sizeof(int) * (size_t)(GRID_COLS + GRID_COLS + 2) + (size_t)capacity + 16; }

void launch_search_points(const SearchPointsArgs& a, int nframes, cudaStream_t s) {
  search_points_kernel<<<nframes, 32, search_points_smem(a.capacity), s>>>(a);
}

// ---------------------------------------------------------------------------------------- SearchByPoints
// src/ORBmatcher.cc:1209-1304 (loop detection): brute force over the keypoints of the other keyframe that have a good map point and
// are not taken yet (vbMatched2), best / second-best, TH_LOW and the ratio test, rotation histogram.  Queries in order; one
// warp per keyframe pair, the lanes split the rows of the second keyframe.
__global__ void __launch_bounds__(32) search_by_points_kernel(SearchByPointsArgs a) {
  extern __shared__ int s_mem[];
  const int pair = blockIdx.x, lane = threadIdx.x, cap = a.capacity;
  uint32_t* s_taken = reinterpret_cast<uint32_t*>(s_mem);       // bitset over the rows of KF2: no good map point, or matched
  int* s_histo = s_mem + (cap + 31) / 32;                        // [32]
  int8_t* s_bin = reinterpret_cast<int8_t*>(s_histo + 32);       // [cap]
  const int n1 = min(a.n1[pair], cap), n2 = min(a.n2[pair], cap);
  const KP* k1 = reinterpret_cast<const KP*>(a.kps1) + (int64_t)pair * cap;
  const KP* k2 = reinterpret_cast<const KP*>(a.kps2) + (int64_t)pair * cap;
  const uint8_t* d1 = a.desc1 + (int64_t)pair * cap * 32;
  const uint4* d2 = reinterpret_cast<const uint4*>(a.desc2 + (int64_t)pair * cap * 32);
  const uint8_t* v1 = a.valid1 + (int64_t)pair * cap;
  const uint8_t* v2 = a.valid2 + (int64_t)pair * cap;
  int32_t* m12 = a.matches12 + (int64_t)pair * cap;
  for (int w = lane; w < (n2 + 31) / 32; w += 32) {
    uint32_t bits = 0;
    for (int b = 0; b < 32; ++b) {
      const int j = 32 * w + b;
      if (j >= n2 || !v2[j]) bits |= 1u << b;
    }
    s_taken[w] = bits;
  }
  for (int i = lane; i < cap; i += 32) {
    m12[i] = -1;
    s_bin[i] = -1;
  }
  s_histo[lane] = 0;
  __syncwarp();
  int nmatches = 0;
  for (int i1 = 0; i1 < n1; ++i1) {
    if (!v1[i1]) continue;
    uint32_t q[8];
    load_desc(d1 + (int64_t)i1 * 32, q);
    int best1 = (256 << 16) | 0xFFFF, best2 = 256;
    for (int j = lane; j < n2; j += 32) {
      if ((s_taken[j >> 5] >> (j & 31)) & 1u) continue;
      const int d = hamming256(q, d2[2 * j], d2[2 * j + 1]);
      best2 = min(best2, max(d, best1 >> 16));
      best1 = min(best1, (d << 16) | j);
    }
    warp_best2(best1, best2);
    const int bd = best1 >> 16;
    if (bd < a.th_low && (float)bd < __fmul_rn(a.nnratio, (float)best2)) {
      if (lane == 0) {
        const int j = best1 & 0xFFFF;
        m12[i1] = j;
        s_taken[j >> 5] |= 1u << (j & 31);
        if (a.check_orientation) {
          const int bin = rotation_bin(k1[i1].angle, k2[j].angle);
          s_bin[i1] = (int8_t)bin;
          s_histo[bin]++;
        }
        nmatches++;
      }
    }
    __syncwarp();
  }
  if (a.check_orientation) {
    const uint32_t keep = three_maxima_mask(s_histo);
    int removed = 0;
    for (int i = lane; i < n1; i += 32) {
      const int bin = s_bin[i];
      if (bin >= 0 && !((keep >> bin) & 1u)) {
        m12[i] = -1;
        removed++;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
    nmatches -= removed;
  }
  if (lane == 0) a.nmatches[pair] = nmatches;
}

void launch_search_by_points(const SearchByPointsArgs& a, int npairs, cudaStream_t s) {
  const size_t smem = sizeof(int) * (size_t)((a.capacity + 31) / 32 + 32) + (size_t)a.capacity + 16;
  search_by_points_kernel<<<npairs, 32, smem, s>>>(a);
}

// ---------------------------------------------------------------------------------------- Fuse (the keypoint search)
// src/ORBmatcher.cc:535-586: every map point is independent here (the map surgery that follows stays with the caller), so a
// thread takes a map point and walks the window's grid cells in the reference's order; strict < keeps the first minimum.
// Without the reprojection gate it is the search of the Sim3 overload of Fuse (:682-708) and of both directions of SearchBySim3
// (:812-845, :890-923; th_dist = TH_HIGH).
__global__ void __launch_bounds__(128) fuse_search_kernel(FuseSearchArgs a) {
  const int frame = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x, cap = a.capacity, capmp = a.capacity_mp;
  if (i >= capmp) return;
  int32_t* out_idx = a.best_idx + (int64_t)frame * capmp;
  int bestDist = 256, bestIdx = -1;
  const int nMP = min(a.n_mp[frame], capmp);
  if (i < nMP && (a.flags[(int64_t)frame * capmp + i] & 1)) {
    const float* pr = a.proj + ((int64_t)frame * capmp + i) * 3;
    const float u = pr[0], v = pr[1], ur = pr[2];
    const int level = a.level[(int64_t)frame * capmp + i];
    const float radius = __fmul_rn(a.th, a.scale_factors[min(max(level, 0), SDORB_MAX_LEVELS - 1)]);
    int x0, x1, y0, y1;
    if (area_cells(u, v, radius, a.grid, x0, x1, y0, y1)) {
      const KP* kF = reinterpret_cast<const KP*>(a.kps) + (int64_t)frame * cap;
      const uint8_t* dF = a.desc + (int64_t)frame * cap * 32;
      const float* uR = a.u_right + (int64_t)frame * cap;
      const int32_t* cs = a.grid.cell_start + (int64_t)frame * (GRID_COLS * GRID_ROWS + 1);
      const int32_t* idx = a.grid.indices + (int64_t)frame * cap;
      uint32_t q[8];
      load_desc(a.desc_mp + ((int64_t)frame * capmp + i) * 32, q);
      for (int ix = x0; ix <= x1; ++ix) {
        const int e = cs[ix * GRID_ROWS + y1 + 1];
        for (int s = cs[ix * GRID_ROWS + y0]; s < e; ++s) {  // the cells (ix, y0..y1) are one contiguous span
          const int i2 = idx[s];
          const KP kp = kF[i2];
          if (!(fabsf(__fsub_rn(kp.x, u)) < radius && fabsf(__fsub_rn(kp.y, v)) < radius)) continue;
          if (kp.octave < level - 1 || kp.octave > level) continue;
          if (a.check_reprojection) {
            const float inv = a.inv_sigma2[min(max(kp.octave, 0), SDORB_MAX_LEVELS - 1)];
            const float ex = __fsub_rn(u, kp.x), ey = __fsub_rn(v, kp.y);
            float e2 = __fmaf_rn(ex, ex, __fmul_rn(ey, ey));
            const float r2 = uR[i2];
            double lim = 5.99;
            if (r2 >= 0.f) {
              const float er = __fsub_rn(ur, r2);
              e2 = __fmaf_rn(er, er, e2);
              lim = 7.8;
            }
            if ((double)__fmul_rn(e2, inv) > lim) continue;
          }
          uint32_t t[8];
          load_desc(dF + (int64_t)i2 * 32, t);
          const int d = hamming256(q, make_uint4(t[0], t[1], t[2], t[3]), make_uint4(t[4], t[5], t[6], t[7]));
          if (d < bestDist) {
            bestDist = d;
            bestIdx = i2;
          }
        }
      }
    }
  }
  if (a.best_dist) a.best_dist[(int64_t)frame * capmp + i] = bestDist;
  out_idx[i] = bestDist <= a.th_dist ? bestIdx : -1;
}

void launch_fuse_search(const FuseSearchArgs& a, int nframes, cudaStream_t s) {
  fuse_search_kernel<<<dim3((a.capacity_mp + 127) / 128, nframes), 128, 0, s>>>(a);
}

// The agreement check of ORBmatcher::SearchBySim3, src/ORBmatcher.cc:927-941: vpMatches12[i1] takes the map point of idx2 =
// vnMatch1[i1] when vnMatch2[idx2] == i1.  One CTA per keyframe pair; the count is an integer sum (order-independent).
__global__ void __launch_bounds__(256) sim3_agreement_kernel(const int32_t* __restrict__ match1, const int32_t* __restrict__ match2,
                                                             const int32_t* __restrict__ n1, int capacity, int32_t* __restrict__ matches12,
                                                             int32_t* __restrict__ nfound) {
  __shared__ int s_found;
  const int pair = blockIdx.x, N1 = min(n1[pair], capacity);
  const int32_t* m1 = match1 + (int64_t)pair * capacity;
  const int32_t* m2 = match2 + (int64_t)pair * capacity;
  int32_t* m12 = matches12 + (int64_t)pair * capacity;
  if (threadIdx.x == 0) s_found = 0;
  __syncthreads();
  int found = 0;
  for (int i1 = threadIdx.x; i1 < capacity; i1 += 256) {
    int out = -1;
    if (i1 < N1) {
      const int idx2 = m1[i1];
      if (idx2 >= 0 && idx2 < capacity && m2[idx2] == i1) {
        out = idx2;
        ++found;
      }
    }
    m12[i1] = out;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) found += __shfl_xor_sync(0xffffffffu, found, o);
  if ((threadIdx.x & 31) == 0 && found) atomicAdd(&s_found, found);
  __syncthreads();
  if (threadIdx.x == 0) nfound[pair] = s_found;
}

void launch_sim3_agreement(const int32_t* match1, const int32_t* match2, const int32_t* n1, int capacity, int32_t* matches12,
                           int32_t* nfound, int npairs, cudaStream_t s) {
  sim3_agreement_kernel<<<npairs, 256, 0, s>>>(match1, match2, n1, capacity, matches12, nfound);
}

// ---------------------------------------------------------------------------------------- SearchForTriangulation
// src/ORBmatcher.cc:359-462 with CheckDistEpipolarLine :128-144.  This reference never sets vbMatched2, so every keypoint of
// KF1 is independent: among the keypoints of KF2 that have no map point, pass the epipolar gate, lie within TH_LOW and (for a
// mono / mono pair) away from the epipole, the LAST one with the smallest distance wins (`dist > bestDist` keeps ties
// moving on).  Thread = keypoint of KF1, CTA = frame pair; KF2 is staged through shared memory in tiles; the rotation
// histogram of the pair is finished by the same CTA.  Arithmetic as the reference build evaluates it: a, b, c in double
// with one FMA each, num / den / the epipole distance in float with one FMA each (see the oracle).
constexpr int TRI_THREADS = 256;
constexpr int TRI_TILE = 256;

struct TriTrain {
  float x, y;
  float gate;   // smallest float >= 3.84 * mvLevelSigma2[octave]  (dsqr < gate  <=>  the double comparison of :143); -1 = has a map point
  float eplim;  // 100 * mvScaleFactors[octave] for a mono keypoint, -1 for a stereo one (the epipole test does not apply)
};

__global__ void __launch_bounds__(TRI_THREADS) search_triangulation_kernel(SearchTriArgs a) {
  __shared__ uint4 s_desc[TRI_TILE][2];
  __shared__ TriTrain s_tr[TRI_TILE];
  __shared__ int s_histo[32];
  __shared__ int s_removed;
  extern __shared__ int8_t s_bin[];  // [capacity]
  const int pair = blockIdx.x, tid = threadIdx.x, cap = a.capacity;
  const int n1 = min(a.n1[pair], cap), n2 = min(a.n2[pair], cap);
  const KP* k1 = reinterpret_cast<const KP*>(a.kps1) + (int64_t)pair * cap;
  const KP* k2 = reinterpret_cast<const KP*>(a.kps2) + (int64_t)pair * cap;
  const uint8_t* d1 = a.desc1 + (int64_t)pair * cap * 32;
  const uint8_t* d2 = a.desc2 + (int64_t)pair * cap * 32;
  const uint8_t* mp1 = a.has_mp1 + (int64_t)pair * cap;
  const uint8_t* mp2 = a.has_mp2 + (int64_t)pair * cap;
  const float* ur1 = a.u_right1 + (int64_t)pair * cap;
  const float* ur2 = a.u_right2 + (int64_t)pair * cap;
  const double* F = a.F12 + (int64_t)pair * 9;
  const float ex = a.epipole[2 * pair], ey = a.epipole[2 * pair + 1];
  int32_t* m12 = a.matches12 + (int64_t)pair * cap;
  if (tid < 32) s_histo[tid] = 0;
  if (tid == 0) s_removed = 0;
  int matched = 0;
  for (int q0 = 0; q0 < n1; q0 += TRI_THREADS) {
    const int i1 = q0 + tid;
    const bool live = i1 < n1 && !mp1[i1];
    uint32_t q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float la = 0.f, lb = 0.f, lc = 0.f, den = 0.f;
    bool stereo1 = true;
    if (live) {
      load_desc(d1 + (int64_t)i1 * 32, q);
      const double x1 = (double)k1[i1].x, y1 = (double)k1[i1].y;
      la = (float)__dadd_rn(__fma_rn(x1, F[0], __dmul_rn(y1, F[3])), F[6]);
      lb = (float)__dadd_rn(__fma_rn(x1, F[1], __dmul_rn(y1, F[4])), F[7]);
      lc = (float)__dadd_rn(__fma_rn(x1, F[2], __dmul_rn(y1, F[5])), F[8]);
      den = __fmaf_rn(la, la, __fmul_rn(lb, lb));
      stereo1 = ur1[i1] >= 0.f;
    }
    const bool query_ok = live && den != 0.f;
    // packed (distance << 16 | 0xFFFF - idx2): the minimum is the smallest distance, the largest index among equals
    uint32_t best = ((uint32_t)(a.th_low + 1) << 16);
    for (int t0 = 0; t0 < n2; t0 += TRI_TILE) {
      __syncthreads();
      const int i2 = t0 + tid;
      if (tid < TRI_TILE && i2 < n2) {
        s_desc[tid][0] = reinterpret_cast<const uint4*>(d2 + (int64_t)i2 * 32)[0];
        s_desc[tid][1] = reinterpret_cast<const uint4*>(d2 + (int64_t)i2 * 32)[1];
        const KP kp = k2[i2];
        const int oct = min(max(kp.octave, 0), SDORB_MAX_LEVELS - 1);
        TriTrain tr;
        tr.x = kp.x;
        tr.y = kp.y;
        tr.gate = mp2[i2] ? -1.f : a.gate[oct];
        tr.eplim = ur2[i2] >= 0.f ? -1.f : a.eplim[oct];
        s_tr[tid] = tr;
      }
      __syncthreads();
      if (query_ok) {
        const int m = min(TRI_TILE, n2 - t0);
        for (int j = 0; j < m; ++j) {
          const TriTrain tr = s_tr[j];
          const float num = __fadd_rn(__fmaf_rn(la, tr.x, __fmul_rn(lb, tr.y)), lc);
          const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
          if (!(dsqr < tr.gate)) continue;
          const int dist = hamming256(q, s_desc[j][0], s_desc[j][1]);
          if (!stereo1) {
            const float dx = __fsub_rn(ex, tr.x), dy = __fsub_rn(ey, tr.y);
            if (__fmaf_rn(dx, dx, __fmul_rn(dy, dy)) < tr.eplim) continue;
          }
          best = min(best, ((uint32_t)dist << 16) | (uint32_t)(0xFFFF - (t0 + j)));
        }
      }
    }
    int bin = -1;
    if (i1 < n1) {
      int m = -1;
      if (query_ok && (int)(best >> 16) <= a.th_low) {
        m = 0xFFFF - (int)(best & 0xFFFFu);
        ++matched;
        if (a.check_orientation) {
          bin = rotation_bin(k1[i1].angle, k2[m].angle);
          atomicAdd(&s_histo[bin], 1);
        }
      }
      m12[i1] = m;
      s_bin[i1] = (int8_t)bin;
    }
  }
  for (int i = n1 + tid; i < cap; i += TRI_THREADS) m12[i] = -1;
  __syncthreads();
  int removed = 0;
  if (a.check_orientation) {
    const uint32_t keep = three_maxima_mask(s_histo);
    for (int i = tid; i < n1; i += TRI_THREADS) {
      const int bin = s_bin[i];
      if (bin >= 0 && !((keep >> bin) & 1u)) {
        m12[i] = -1;
        ++removed;
      }
    }
  }
  // nmatches = matches found - matches removed, summed over the CTA
  int net = matched - removed;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) net += __shfl_xor_sync(0xffffffffu, net, o);
  if ((tid & 31) == 0) atomicAdd(&s_removed, net);
  __syncthreads();
  if (tid == 0) a.nmatches[pair] = s_removed;
}

void launch_search_triangulation(const SearchTriArgs& a, int npairs, cudaStream_t s) {
  search_triangulation_kernel<<<npairs, TRI_THREADS, (size_t)a.capacity + 16, s>>>(a);
}

int configure_search_kernels() {
  cudaError_t e = cudaFuncSetAttribute(search_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(search_projection_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(search_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  return (int)e;
}

void launch_search_init(const SearchInitArgs& a, int npairs, cudaStream_t s) {
  search_init_kernel<<<npairs, 32, search_init_smem(a.capacity), s>>>(a);
}
void launch_search_projection(const SearchProjArgs& a, int npairs, cudaStream_t s) {
  search_projection_kernel<<<npairs, 32, search_projection_smem(a.capacity), s>>>(a);
}

}  // namespace sdorb
