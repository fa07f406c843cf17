// kernels_match.cu -- ORBmatcher::DescriptorDistance batched as Hamming matching, for sm_100a.
//
// Reference: /root/reference/src/ORBmatcher.cc:1459-1473 (256-bit Hamming distance as eight 32-bit SWAR popcounts,
// == __popc) and the best / second-best scan of SearchByPoints (:1239-1265):
//     if (d < best1) { best2 = best1; best1 = d; idx = j; } else if (d < best2) best2 = d;
// with j ascending, so the FIRST minimal index wins ties.  Bound: the integer popc (XU) pipe, not memory (64 KB in,
// 16 KB out per 1000x1000 pair); ncu: XU pipe 95 % busy.  One thread owns one query row (8 registers); train rows stream through shared
// memory and are read as broadcasts.  (distance << 16 | j) packs value and index so a single integer min keeps
// the first-minimum rule; the runner-up needs only  best2 = min(best2, max(d, best1)).
#include "kernels.cuh"
#include "match_common.cuh"

namespace sdorb {

struct MatchOut {
  int32_t best_idx, best_dist, second_dist, accepted;
};

constexpr int M_THREADS = 128;
constexpr int M_TILE = 256;  // train rows staged per step (8 KB)

__device__ __forceinline__ int accept_rule(int best1, int best2, float ratio, int th_low) {
  return (best1 < th_low && (float)best1 < __fmul_rn(ratio, (float)best2)) ? 1 : 0;
}

__global__ void __launch_bounds__(M_THREADS) match_kernel(const uint8_t* __restrict__ A, const int32_t* __restrict__ nA,
                                                          int strideA, const uint8_t* __restrict__ B,
                                                          const int32_t* __restrict__ nB, int strideB, float ratio,
                                                          int th_low, MatchOut* __restrict__ out) {
  __shared__ uint4 s_b[M_TILE * 2];
  const int pair = blockIdx.x;  // the batch index lives on gridDim.x (2^31 - 1), the query block on gridDim.y
  // device-resident counts are clamped to the slab like the guided-search kernels do (the host form validates them)
  const int na = min(max(nA[pair], 0), strideA), nb = min(max(nB[pair], 0), strideB);
  const int qi = blockIdx.y * M_THREADS + threadIdx.x;
  const bool active = qi < na;
  // rows past the pair's count still get a defined result ("no candidate"), so a caller may reduce over whole slabs
  if (!active && qi < strideA) out[(int64_t)pair * strideA + qi] = MatchOut{-1, 256, 256, 0};
  if (blockIdx.y * M_THREADS >= na) return;
  uint32_t q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    const uint4* qa = reinterpret_cast<const uint4*>(A + ((int64_t)pair * strideA + qi) * 32);
    const uint4 lo = qa[0], hi = qa[1];
    q[0] = lo.x; q[1] = lo.y; q[2] = lo.z; q[3] = lo.w;
    q[4] = hi.x; q[5] = hi.y; q[6] = hi.z; q[7] = hi.w;
  }
  // (distance << 16 | index) orders candidates by distance, then by index: the smallest packed value is the first
  // minimum, and the second smallest carries the second smallest distance.  Both are tracked packed (3 min / max, no shift).
  int best1 = (256 << 16) | 0xFFFF, best2 = (256 << 16) | 0xFFFF;
  const uint4* bsrc = reinterpret_cast<const uint4*>(B + (int64_t)pair * strideB * 32);
  for (int j0 = 0; j0 < nb; j0 += M_TILE) {
    const int cnt = min(M_TILE, nb - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 2; i += M_THREADS) s_b[i] = bsrc[(int64_t)j0 * 2 + i];
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const int d = hamming256(q, s_b[2 * j], s_b[2 * j + 1]);
      const int dp = d * 65536 + (j0 + j);  // IMAD: the ALU pipe is the next bottleneck after POPC
      best2 = min(best2, max(dp, best1));
      best1 = min(best1, dp);
    }
  }
  if (active) {
    MatchOut o;
    o.best_dist = best1 >> 16;
    o.best_idx = o.best_dist < 256 ? (best1 & 0xFFFF) : -1;
    o.second_dist = best2 >> 16;
    o.accepted = accept_rule(o.best_dist, o.second_dist, ratio, th_low);
    out[(int64_t)pair * strideA + qi] = o;
  }
}

// SearchByPoints' greedy form: queries in order; a train row accepted by an earlier query is masked for later
// ones (vbMatched2, src/ORBmatcher.cc:1228, 1245, 1267).  Sequential over queries by nature: one warp per pair,
// lanes split the train rows, (distance<<16 | j) min-reduced across the warp, runner-up merged alongside.
__global__ void __launch_bounds__(32) match_greedy_kernel(const uint8_t* __restrict__ A, const int32_t* __restrict__ nA,
                                                          int strideA, const uint8_t* __restrict__ B,
                                                          const int32_t* __restrict__ nB, int strideB, float ratio,
                                                          int th_low, MatchOut* __restrict__ out) {
  extern __shared__ uint32_t s_matched[];  // bitset over train rows
  const int pair = blockIdx.x, lane = threadIdx.x;
  const int na = min(max(nA[pair], 0), strideA), nb = min(max(nB[pair], 0), strideB);  // s_matched is sized from strideB
  for (int i = lane; i < (nb + 31) / 32; i += 32) s_matched[i] = 0;
  __syncwarp();
  const uint4* bsrc = reinterpret_cast<const uint4*>(B + (int64_t)pair * strideB * 32);
  for (int qi = 0; qi < na; ++qi) {
    const uint4* qa = reinterpret_cast<const uint4*>(A + ((int64_t)pair * strideA + qi) * 32);
    const uint4 lo = qa[0], hi = qa[1];
    const uint32_t q[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    int best1 = (256 << 16) | 0xFFFF, best2 = 256;
    for (int j = lane; j < nb; j += 32) {
      if ((s_matched[j >> 5] >> (j & 31)) & 1u) continue;
      const int d = hamming256(q, bsrc[2 * j], bsrc[2 * j + 1]);
      best2 = min(best2, max(d, best1 >> 16));
      best1 = min(best1, (d << 16) | j);
    }
    // merge (best1, best2) pairs: the two smallest distances of the union, first index on ties
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int ob1 = __shfl_xor_sync(0xffffffffu, best1, o), ob2 = __shfl_xor_sync(0xffffffffu, best2, o);
      const int lo1 = min(best1, ob1), hi1 = max(best1, ob1);
      best2 = min(min(best2, ob2), hi1 >> 16);
      best1 = lo1;
    }
    const int bd = best1 >> 16;
    const int ok = accept_rule(bd, best2, ratio, th_low);
    if (lane == 0) {
      MatchOut r;
      r.best_dist = bd;
      r.best_idx = bd < 256 ? (best1 & 0xFFFF) : -1;
      r.second_dist = best2;
      r.accepted = ok;
      out[(int64_t)pair * strideA + qi] = r;
      if (ok) s_matched[(best1 & 0xFFFF) >> 5] |= 1u << (best1 & 31);
    }
    __syncwarp();
  }
  for (int qi = na + lane; qi < strideA; qi += 32) out[(int64_t)pair * strideA + qi] = MatchOut{-1, 256, 256, 0};
}

__global__ void __launch_bounds__(256) hamming_matrix_kernel(const uint8_t* __restrict__ A, int nA,
                                                             const uint8_t* __restrict__ B, int nB,
                                                             uint16_t* __restrict__ out) {
  const int j = blockIdx.y * 256 + threadIdx.x, i = blockIdx.x;  // rows of A on gridDim.x: no 65535 limit
  if (j >= nB) return;
  const uint4* qa = reinterpret_cast<const uint4*>(A + (int64_t)i * 32);
  const uint4 lo = qa[0], hi = qa[1];
  const uint32_t q[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  const uint4* b = reinterpret_cast<const uint4*>(B + (int64_t)j * 32);
  out[(int64_t)i * nB + j] = (uint16_t)hamming256(q, b[0], b[1]);
}

// ---- MapPoint::ComputeDistinctiveDescriptors (/root/reference/src/MapPoint.cc:252-275), batched over map points:
// all pairwise DescriptorDistance values of a set of N observed descriptors, per row the median = element
// (size_t)(0.5 * (N - 1)) of the sorted row, the FIRST row with the smallest median wins.  One CTA per set, one thread
// per row: the row's distances go into a 257-bin histogram in shared memory (distances are 0..256), the median is the
// bin where the running count passes the rank -- a counting sort, no matrix, any N.
constexpr int D_THREADS = 64;

__global__ void __launch_bounds__(D_THREADS) distinctive_kernel(const uint8_t* __restrict__ desc, const int32_t* __restrict__ offsets,
                                                                int32_t* __restrict__ best_idx, int32_t* __restrict__ best_median) {
  __shared__ uint16_t s_hist[D_THREADS][258];  // pitch 258 halfwords = 129 words: threads start in different banks
  __shared__ int s_best[D_THREADS / 32];
  const int set = blockIdx.x, tid = threadIdx.x;
  const int first = offsets[set], n = min(offsets[set + 1] - first, 65535);  // uint16 bins and (median << 16 | row) hold 65535 rows
  if (n <= 0) {
    if (tid == 0) {
      best_idx[set] = -1;
      if (best_median) best_median[set] = 0x7fffffff;
    }
    return;
  }
  const uint4* rows = reinterpret_cast<const uint4*>(desc + (int64_t)first * 32);
  const int rank = (int)(0.5 * (n - 1));
  int best = 0x7fffffff;  // median << 16 | row: the smallest packed value is the first row with the smallest median
  for (int i = tid; i < n; i += D_THREADS) {  // n <= 65535 rows per set
    uint16_t* hist = s_hist[tid];
    for (int b = 0; b < 257; ++b) hist[b] = 0;
    const uint4 lo = __ldg(rows + 2 * i), hi = __ldg(rows + 2 * i + 1);
    const uint32_t q[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    for (int j = 0; j < n; ++j) hist[hamming256(q, __ldg(rows + 2 * j), __ldg(rows + 2 * j + 1))] += 1;  // D[i][i] = 0 included
    int acc = 0, median = 0;
    for (int b = 0; b < 257; ++b) {
      acc += hist[b];
      if (acc > rank) {
        median = b;
        break;
      }
    }
    best = min(best, (median << 16) | i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((tid & 31) == 0) s_best[tid >> 5] = best;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < D_THREADS / 32; ++w) best = min(best, s_best[w]);
    best_idx[set] = best & 0xFFFF;
    if (best_median) best_median[set] = best >> 16;
  }
}

void launch_distinctive(const uint8_t* desc, const int32_t* offsets, int nsets, int32_t* best_idx, int32_t* best_median,
                        cudaStream_t s) {
  if (nsets <= 0) return;
  distinctive_kernel<<<nsets, D_THREADS, 0, s>>>(desc, offsets, best_idx, best_median);
}

void launch_match(const uint8_t* A, const int32_t* nA, int strideA, const uint8_t* B, const int32_t* nB, int strideB,
                  int npairs, float ratio, int th_low, void* out, cudaStream_t s) {
  if (npairs <= 0 || strideA <= 0) return;
  dim3 grid(npairs, (strideA + M_THREADS - 1) / M_THREADS);
  match_kernel<<<grid, M_THREADS, 0, s>>>(A, nA, strideA, B, nB, strideB, ratio, th_low, (MatchOut*)out);
}

void launch_match_greedy(const uint8_t* A, const int32_t* nA, int strideA, const uint8_t* B, const int32_t* nB,
                         int strideB, int npairs, float ratio, int th_low, void* out, cudaStream_t s) {
  if (npairs <= 0) return;
  const size_t smem = sizeof(uint32_t) * (size_t)((strideB + 31) / 32 + 1);
  match_greedy_kernel<<<npairs, 32, smem, s>>>(A, nA, strideA, B, nB, strideB, ratio, th_low, (MatchOut*)out);
}

void launch_hamming_matrix(const uint8_t* A, int nA, const uint8_t* B, int nB, uint16_t* out, cudaStream_t s) {
  if (nA <= 0 || nB <= 0) return;
  hamming_matrix_kernel<<<dim3(nA, (nB + 255) / 256), 256, 0, s>>>(A, nA, B, nB, out);
}

}  // namespace sdorb
