// kernels_frame.cu -- the per-keypoint post-processing Frame does right after the extractor.
//
// Reference: /root/reference/src/Frame.cc
//   PosInGrid             :323-332  posX = round((x - mnMinX) * mfGridElementWidthInv), same in y; outside the 64 x 48 grid -> dropped
//   AssignFeaturesToGrid  :179-192  mGrid[posX][posY].push_back(i) for i ascending (FRAME_GRID_COLS 64, FRAME_GRID_ROWS 48, Frame.h:37-38)
//   ComputeStereoFromRGBD :399-417  d = imDepth.at<float>(v, u) (float indices truncate); d > 0: depth = d, uRight = xUn - mbf / d
//   UndistortKeyPoints    :335-366  cv::undistortPoints(pts, K, dist, noArray(), K); ComputeImageBounds :368-397 the same on 4 corners
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

// cv::undistortPoints(pts, K, dist, noArray(), K) of OpenCV 4.13 (calib3d undistort.dispatch.cpp): double arithmetic, five
// fixed-point iterations, operation order kept with round-to-nearest intrinsics (no contraction).  Shared by the host
// (ComputeImageBounds) and the device (UndistortKeyPoints).
struct Undistorter {
  double fx, fy, cx, cy, ifx, ify, k[12];
};

__host__ __device__ inline void undistort_point(const Undistorter& U, float u_in, float v_in, float* xo, float* yo) {
#ifdef __CUDA_ARCH__
#define DMUL(a, b) __dmul_rn(a, b)
#define DADD(a, b) __dadd_rn(a, b)
#define DSUB(a, b) __dsub_rn(a, b)
#define DDIV(a, b) __ddiv_rn(a, b)
#else
#define DMUL(a, b) ((a) * (b))
#define DADD(a, b) ((a) + (b))
#define DSUB(a, b) ((a) - (b))
#define DDIV(a, b) ((a) / (b))
#endif
  const double* k = U.k;
  const double u = u_in, v = v_in;
  double x = DMUL(DSUB(u, U.cx), U.ifx), y = DMUL(DSUB(v, U.cy), U.ify);
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; ++j) {
    const double r2 = DADD(DMUL(x, x), DMUL(y, y));
    const double num = DADD(1., DMUL(DADD(DMUL(DADD(DMUL(k[7], r2), k[6]), r2), k[5]), r2));
    const double den = DADD(1., DMUL(DADD(DMUL(DADD(DMUL(k[4], r2), k[1]), r2), k[0]), r2));
    const double icdist = DDIV(num, den);
    if (icdist < 0) {
      x = DMUL(DSUB(u, U.cx), U.ifx);
      y = DMUL(DSUB(v, U.cy), U.ify);
      break;
    }
    // deltaX = 2*k2*x*y + k3*(r2 + 2*x*x) + k8*r2 + k9*r2*r2, left to right
    const double dX = DADD(DADD(DADD(DMUL(DMUL(DMUL(2., k[2]), x), y), DMUL(k[3], DADD(r2, DMUL(DMUL(2., x), x)))), DMUL(k[8], r2)),
                           DMUL(DMUL(k[9], r2), r2));
    // deltaY = k2*(r2 + 2*y*y) + 2*k3*x*y + k10*r2 + k11*r2*r2
    const double dY = DADD(DADD(DADD(DMUL(k[2], DADD(r2, DMUL(DMUL(2., y), y))), DMUL(DMUL(DMUL(2., k[3]), x), y)), DMUL(k[10], r2)),
                           DMUL(DMUL(k[11], r2), r2));
    x = DMUL(DSUB(x0, dX), icdist);
    y = DMUL(DSUB(y0, dY), icdist);
  }
  // P = K, R = I:  xx = fx*x + 0*y + cx,  ww = 1 / (0*x + 0*y + 1)
  const double xx = DADD(DADD(DMUL(U.fx, x), DMUL(0., y)), U.cx), yy = DADD(DADD(DMUL(0., x), DMUL(U.fy, y)), U.cy);
  const double ww = DDIV(1., DADD(DADD(DMUL(0., x), DMUL(0., y)), 1.));
  *xo = (float)DMUL(xx, ww);
  *yo = (float)DMUL(yy, ww);
#undef DMUL
#undef DADD
#undef DSUB
#undef DDIV
}

static Undistorter make_undistorter(const float K[4], const float* dist, int ndist) {
  Undistorter U;
  U.fx = K[0];
  U.fy = K[1];
  U.cx = K[2];
  U.cy = K[3];
  U.ifx = 1. / U.fx;
  U.ify = 1. / U.fy;
  for (int i = 0; i < 12; ++i) U.k[i] = i < ndist ? (double)dist[i] : 0.;
  return U;
}

// Frame::UndistortKeyPoints (src/Frame.cc:335-366): every field but pt is copied
__global__ void __launch_bounds__(256) undistort_kernel(const float* __restrict__ kps, const int32_t* __restrict__ counts,
                                                        int capacity, Undistorter U, int identity, float* __restrict__ out) {
  const int frame = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  if (i >= min(counts[frame], capacity)) return;
  const float* a = kps + ((int64_t)frame * capacity + i) * 7;
  float* b = out + ((int64_t)frame * capacity + i) * 7;
  float x = a[0], y = a[1];
  if (!identity) undistort_point(U, x, y, &x, &y);
  b[0] = x;
  b[1] = y;
#pragma unroll
  for (int f = 2; f < 7; ++f) b[f] = a[f];
}

void launch_undistort(const void* kps, const int32_t* counts, int nframes, int capacity, const float K[4], const float* dist,
                      int ndist, void* out, cudaStream_t s) {
  if (nframes <= 0 || capacity <= 0) return;
  const int identity = !(ndist > 0 && dist[0] != 0.0f);  // src/Frame.cc:336-339
  undistort_kernel<<<dim3((capacity + 255) / 256, nframes), 256, 0, s>>>((const float*)kps, counts, capacity,
                                                                          make_undistorter(K, dist, ndist), identity, (float*)out);
}

// Frame::ComputeImageBounds (src/Frame.cc:368-397) on the host: bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY}
void host_image_bounds(int cols, int rows, const float K[4], const float* dist, int ndist, float bounds[4]) {
  if (ndist > 0 && dist[0] != 0.0f) {
    const Undistorter U = make_undistorter(K, dist, ndist);
    const float px[4] = {0.f, (float)cols, 0.f, (float)cols}, py[4] = {0.f, 0.f, (float)rows, (float)rows};
    float mx[4], my[4];
    for (int i = 0; i < 4; ++i) undistort_point(U, px[i], py[i], &mx[i], &my[i]);
    bounds[0] = mx[0] < mx[2] ? mx[0] : mx[2];
    bounds[1] = mx[1] > mx[3] ? mx[1] : mx[3];
    bounds[2] = my[0] < my[1] ? my[0] : my[1];
    bounds[3] = my[2] > my[3] ? my[2] : my[3];
  } else {
    bounds[0] = 0.f;
    bounds[1] = (float)cols;
    bounds[2] = 0.f;
    bounds[3] = (float)rows;
  }
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
