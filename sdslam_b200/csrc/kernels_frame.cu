// kernels_frame.cu -- the per-keypoint post-processing Frame does right after the extractor.
//
// Reference: /root/reference/src/Frame.cc
//   PosInGrid             :323-332  posX = round((x - mnMinX) * mfGridElementWidthInv), same in y; outside the 64 x 48 grid -> dropped
//   AssignFeaturesToGrid  :179-192  mGrid[posX][posY].push_back(i) for i ascending (FRAME_GRID_COLS 64, FRAME_GRID_ROWS 48, Frame.h:37-38)
//   ComputeStereoFromRGBD :399-417  d = imDepth.at<float>(v, u) (float indices truncate); d > 0: depth = d, uRight = xUn - mbf / d
// The grid comes back in CSR form: cell c = posX * 48 + posY owns indices[cell_start[c] .. cell_start[c+1]), ascending --
// exactly the order of the reference's push_backs.  One CTA per frame; no atomics decide any order.
#include "kernels.cuh"

namespace sdorb {

constexpr int G_THREADS = 256;
constexpr int G_COLS = 64, G_ROWS = 48, G_CELLS = G_COLS * G_ROWS;
constexpr int G_CELLS_PER_THREAD = G_CELLS / G_THREADS;  // 12
static_assert(G_CELLS % G_THREADS == 0, "cells split evenly over the threads");

__global__ void __launch_bounds__(G_THREADS) assign_grid_kernel(const float* __restrict__ kps, const int32_t* __restrict__ counts,
                                                                int capacity, float min_x, float min_y, float inv_w, float inv_h,
                                                                int32_t* __restrict__ cell_start, int32_t* __restrict__ indices) {
  extern __shared__ uint16_t s_cell[];  // cell of every keypoint of the frame (0xFFFF = outside the grid)
  __shared__ int s_count[G_CELLS];
  __shared__ int s_warp_sum[G_THREADS / 32];
  const int frame = blockIdx.x, tid = threadIdx.x;
  const int n = min(counts[frame], capacity);
  const float* k = kps + (int64_t)frame * capacity * 7;
  for (int c = tid; c < G_CELLS; c += G_THREADS) s_count[c] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += G_THREADS) {
    // round(float) = half away from zero; the products are rounded to float first, as in the reference
    const float px = roundf(__fmul_rn(__fsub_rn(k[7 * i], min_x), inv_w));
    const float py = roundf(__fmul_rn(__fsub_rn(k[7 * i + 1], min_y), inv_h));
    uint16_t cell = 0xFFFF;
    if (px >= 0.f && px < (float)G_COLS && py >= 0.f && py < (float)G_ROWS) {
      cell = (uint16_t)((int)px * G_ROWS + (int)py);
      atomicAdd(&s_count[cell], 1);  // a count, not an order
    }
    s_cell[i] = cell;
  }
  __syncthreads();
  // exclusive prefix sum over the cells: thread t owns cells [12 t, 12 t + 12)
  int local[G_CELLS_PER_THREAD], sum = 0;
#pragma unroll
  for (int j = 0; j < G_CELLS_PER_THREAD; ++j) {
    local[j] = sum;
    sum += s_count[tid * G_CELLS_PER_THREAD + j];
  }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += v;
  }
  if ((tid & 31) == 31) s_warp_sum[tid >> 5] = incl;
  __syncthreads();
  int base = incl - sum;
  for (int w = 0; w < (tid >> 5); ++w) base += s_warp_sum[w];
  int32_t* cs = cell_start + (int64_t)frame * (G_CELLS + 1);
  int cursor[G_CELLS_PER_THREAD];
#pragma unroll
  for (int j = 0; j < G_CELLS_PER_THREAD; ++j) {
    cursor[j] = base + local[j];
    cs[tid * G_CELLS_PER_THREAD + j] = cursor[j];
  }
  if (tid == G_THREADS - 1) cs[G_CELLS] = base + sum;
  // every thread walks all keypoints in index order and appends those of its own cells: ascending within each cell
  int32_t* out = indices + (int64_t)frame * capacity;
  const int c_lo = tid * G_CELLS_PER_THREAD;
  for (int i = 0; i < n; ++i) {
    const int j = (int)s_cell[i] - c_lo;  // broadcast read
    if ((unsigned)j < (unsigned)G_CELLS_PER_THREAD) {
#pragma unroll
      for (int q = 0; q < G_CELLS_PER_THREAD; ++q)
        if (q == j) out[cursor[q]++] = i;
    }
  }
}

__global__ void __launch_bounds__(256) stereo_rgbd_kernel(const float* __restrict__ kps, const float* __restrict__ kps_un,
                                                          const int32_t* __restrict__ counts, int capacity,
                                                          const float* __restrict__ depth, int width, int height,
                                                          int64_t depth_row_stride, int64_t depth_frame_stride, float mbf,
                                                          float* __restrict__ u_right, float* __restrict__ z) {
  const int frame = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= capacity) return;
  float ur = -1.f, d_out = -1.f;
  if (i < min(counts[frame], capacity)) {
    const float* kp = kps + ((int64_t)frame * capacity + i) * 7;
    const int u = (int)kp[0], v = (int)kp[1];  // Mat::at<float>(float v, float u): the indices truncate
    if (u >= 0 && u < width && v >= 0 && v < height) {
      const float d = depth[(int64_t)frame * depth_frame_stride + (int64_t)v * depth_row_stride + u];
      if (d > 0.f) {
        d_out = d;
        ur = __fsub_rn(kps_un[((int64_t)frame * capacity + i) * 7], __fdiv_rn(mbf, d));
      }
    }
  }
  u_right[(int64_t)frame * capacity + i] = ur;
  z[(int64_t)frame * capacity + i] = d_out;
}

void launch_assign_grid(const void* kps, const int32_t* counts, int nframes, int capacity, float min_x, float min_y, float inv_w,
                        float inv_h, int32_t* cell_start, int32_t* indices, cudaStream_t s) {
  if (nframes <= 0) return;
  const size_t smem = sizeof(uint16_t) * (size_t)capacity;
  assign_grid_kernel<<<nframes, G_THREADS, smem, s>>>((const float*)kps, counts, capacity, min_x, min_y, inv_w, inv_h, cell_start, indices);
}

void launch_stereo_rgbd(const void* kps, const void* kps_un, const int32_t* counts, int nframes, int capacity, const float* depth,
                        int width, int height, int64_t row_stride, int64_t frame_stride, float mbf, float* u_right, float* z,
                        cudaStream_t s) {
  if (nframes <= 0 || capacity <= 0) return;
  stereo_rgbd_kernel<<<dim3((capacity + 255) / 256, nframes), 256, 0, s>>>((const float*)kps, (const float*)kps_un, counts, capacity,
                                                                             depth, width, height, row_stride, frame_stride, mbf,
                                                                             u_right, z);
}

int configure_frame_kernels() {
  return (int)cudaFuncSetAttribute(assign_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
}

}  // namespace sdorb
