// kernels_image.cu -- pyramid resize and Gaussian blur for sm_100a.
//
// Both reproduce OpenCV's 8-bit fixed-point arithmetic bit for bit:
//   * cv::resize INTER_LINEAR (ComputePyramid, /root/reference/src/ORBextractor.cc:690): 11-bit coefficients,
//     horizontal pass in int32, vertical pass  (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2.
//   * cv::GaussianBlur 7x7 sigma 2 BORDER_REFLECT_101 (src/ORBextractor.cc:660): 8.8 kernel
//     {18,34,48,56,48,34,18}, 16-bit horizontal pass, 32-bit vertical pass, (v + 2^15) >> 16.
// Both are HBM-bound byte kernels: one read and one write of every level pixel.
#include "kernels.cuh"

namespace sdorb {

__device__ __forceinline__ const uint8_t* level_plane(const BatchPlanes& p, const LevelGeom& L, int level, int frame,
                                                      int* pitch) {
  if (level == 0) {
    *pitch = p.img0_pitch;
    return p.img0 + (int64_t)frame * p.img0_frame_stride;
  }
  *pitch = L.pitch;
  return p.pyr + L.plane_base * p.batch_cap + (int64_t)frame * L.plane_bytes;
}

// ------------------------------------------------------------------------------------------------ resize
// One thread = 4 consecutive output pixels of one row (one 32-bit store).  Grid: (x groups, rows, frames).
__global__ void __launch_bounds__(256) resize_level_kernel(const FrameGeom* __restrict__ geom, int level, BatchPlanes p,
                                                           const ResizeTap* __restrict__ taps) {
  const LevelGeom& D = geom->lv[level];
  const LevelGeom& S = geom->lv[level - 1];
  const int x4 = (blockIdx.x * 64 + threadIdx.x) * 4;
  const int y = blockIdx.y * 4 + threadIdx.y;
  const int frame = blockIdx.z;
  if (x4 >= D.w || y >= D.h) return;
  int spitch;
  const uint8_t* src = level_plane(p, S, level - 1, frame, &spitch);
  uint8_t* dst = p.pyr + D.plane_base * p.batch_cap + (int64_t)frame * D.plane_bytes;
  const ResizeTap ty = taps[D.coef_y_base + y];
  const uint8_t* r0 = src + (int64_t)ty.s0 * spitch;
  const uint8_t* r1 = src + (int64_t)ty.s1 * spitch;
  const ResizeTap* tx = taps + D.coef_x_base;
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const ResizeTap t = tx[min(x4 + i, D.w - 1)];
    const int h0 = r0[t.s0] * t.c0 + r0[t.s1] * t.c1;
    const int h1 = r1[t.s0] * t.c0 + r1[t.s1] * t.c1;
    int v = (((ty.c0 * (h0 >> 4)) >> 16) + ((ty.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
    out |= (uint32_t)v << (8 * i);
  }
  *reinterpret_cast<uint32_t*>(dst + (int64_t)y * D.pitch + x4) = out;  // pitch is a multiple of 128: padding absorbs the tail
}

void launch_resize_level(const FrameGeom* d_geom, const FrameGeom& g, int level, const BatchPlanes& p,
                         const ResizeTap* d_taps, int nframes, cudaStream_t s) {
  const LevelGeom& D = g.lv[level];
  dim3 block(64, 4);
  dim3 grid((D.w + 255) / 256, (D.h + 3) / 4, nframes);
  resize_level_kernel<<<grid, block, 0, s>>>(d_geom, level, p, d_taps);
}

// ------------------------------------------------------------------------------------------------ blur
__device__ __forceinline__ int reflect101(int p, int len) {
  if ((unsigned)p < (unsigned)len) return p;
  if (len == 1) return 0;
  do {
    p = p < 0 ? -p : 2 * (len - 1) - p;
  } while ((unsigned)p >= (unsigned)len);
  return p;
}

constexpr int BTW = SDORB_BLUR_TW, BTH = SDORB_BLUR_TH;
constexpr int B_SRC_W = BTW + 8;   // bytes per staged source row: x0-4 .. x0+TW+3
constexpr int B_ROWS = BTH + 6;

__global__ void __launch_bounds__(256) blur_all_kernel(const FrameGeom* __restrict__ geom, BatchPlanes p) {
  __shared__ __align__(16) uint8_t s_src[B_ROWS][B_SRC_W];
  __shared__ __align__(16) uint16_t s_h[B_ROWS][BTW];
  __shared__ int s_level;
  const int tid = threadIdx.x;
  if (tid == 0) {
    int l = 0;
    while (l + 1 < geom->nlevels && (int)blockIdx.x >= geom->lv[l + 1].tile_base_blur) ++l;
    s_level = l;
  }
  __syncthreads();
  const int level = s_level;
  const LevelGeom& L = geom->lv[level];
  const int frame = blockIdx.y;
  const int t = blockIdx.x - L.tile_base_blur;
  const int x0 = (t % L.tiles_x_blur) * BTW, y0 = (t / L.tiles_x_blur) * BTH;
  const int w = L.w, h = L.h;
  int spitch;
  const uint8_t* src = level_plane(p, L, level, frame, &spitch);

  // stage rows y0-3 .. y0+TH+2 (reflected), columns x0-4 .. x0+TW+3 as 32-bit words
  for (int i = tid; i < B_ROWS * (B_SRC_W / 4); i += 256) {
    const int r = i / (B_SRC_W / 4), k = i % (B_SRC_W / 4);
    const int gy = reflect101(y0 - 3 + r, h);
    const int gx = x0 - 4 + 4 * k;
    const uint8_t* row = src + (int64_t)gy * spitch;
    uint32_t v;
    if (gx >= 0 && gx + 3 < w) {
      v = *reinterpret_cast<const uint32_t*>(row + gx);
    } else {
      v = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) v |= (uint32_t)row[reflect101(gx + b, w)] << (8 * b);
    }
    *reinterpret_cast<uint32_t*>(&s_src[r][4 * k]) = v;
  }
  __syncthreads();
  // horizontal pass: 4 pixels per thread-iteration
  for (int i = tid; i < B_ROWS * (BTW / 4); i += 256) {
    const int r = i / (BTW / 4), k = i % (BTW / 4);
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(&s_src[r][4 * k]);  // tile px 4k..4k+3 live at byte 4k+4
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
    int b[12];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      b[q] = (w0 >> (8 * q)) & 0xFF;
      b[4 + q] = (w1 >> (8 * q)) & 0xFF;
      b[8 + q] = (w2 >> (8 * q)) & 0xFF;
    }
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      o[q] = 18 * (b[q + 1] + b[q + 7]) + 34 * (b[q + 2] + b[q + 6]) + 48 * (b[q + 3] + b[q + 5]) + 56 * b[q + 4];
    *reinterpret_cast<uint2*>(&s_h[r][4 * k]) = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
  }
  __syncthreads();
  // vertical pass
  uint8_t* dst = p.blur + L.plane_base * p.batch_cap + (int64_t)frame * L.plane_bytes;
  for (int i = tid; i < BTH * (BTW / 4); i += 256) {
    const int r = i / (BTW / 4), k = i % (BTW / 4);
    const int gy = y0 + r, gx = x0 + 4 * k;
    if (gy >= h || gx >= w) continue;
    uint32_t acc[4] = {0, 0, 0, 0};
    const int kw[7] = {18, 34, 48, 56, 48, 34, 18};
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const uint2 v = *reinterpret_cast<const uint2*>(&s_h[r + j][4 * k]);
      acc[0] += kw[j] * (v.x & 0xFFFF);
      acc[1] += kw[j] * (v.x >> 16);
      acc[2] += kw[j] * (v.y & 0xFFFF);
      acc[3] += kw[j] * (v.y >> 16);
    }
    const uint32_t out = ((acc[0] + 32768u) >> 16) | (((acc[1] + 32768u) >> 16) << 8) |
                         (((acc[2] + 32768u) >> 16) << 16) | (((acc[3] + 32768u) >> 16) << 24);
    *reinterpret_cast<uint32_t*>(dst + (int64_t)gy * L.pitch + gx) = out;
  }
}

void launch_blur_all(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, int nframes, cudaStream_t s) {
  if (g.tiles_total_blur == 0) return;
  blur_all_kernel<<<dim3(g.tiles_total_blur, nframes), 256, 0, s>>>(d_geom, p);
}

}  // namespace sdorb
