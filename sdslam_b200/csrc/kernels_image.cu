// kernels_image.cu -- pyramid resize and Gaussian blur for sm_100a.
//
// Both reproduce OpenCV's 8-bit fixed-point arithmetic bit for bit:
//   * cv::resize INTER_LINEAR (ComputePyramid, /root/reference/src/ORBextractor.cc:690): 11-bit coefficients,
//     horizontal pass in int32, vertical pass  (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2.
//   * cv::GaussianBlur 7x7 sigma 2 BORDER_REFLECT_101 (src/ORBextractor.cc:660): 8.8 kernel
//     {18,34,48,56,48,34,18}, 16-bit horizontal pass, 32-bit vertical pass, (v + 2^15) >> 16.
// Both are byte kernels that read and write every level pixel once.  To get near the HBM roofline the integer work
// per pixel has to be a handful of instructions, so both use the byte dot-product unit:
//   resize : the 8 source bytes holding all taps of 4 destination pixels are funnel-shifted into place, one PRMT
//            gathers (s0, s1) of two pixels and IDP.2A multiplies them with the (c0, c1) coefficient pair;
//   blur   : the horizontal 7-tap filter of a pixel is two IDP.4A over byte quadruples cut out with PRMT, the
//            vertical one four IDP.2A over vertically paired 16-bit sums, with the rounding constant as accumulator.
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
constexpr int RZ_ROWS = 8;  // consecutive destination rows per thread (a source row feeds two destination rows)

// horizontal pass of one source row for four destination pixels
__device__ __forceinline__ void resize_hrow(const uint8_t* __restrict__ row, int base, int last_word, int shift,
                                            const ResizeGroup& G, int (&hv)[4]) {
  const uint32_t w0 = *reinterpret_cast<const uint32_t*>(row + min(base, last_word));
  const uint32_t w1 = *reinterpret_cast<const uint32_t*>(row + min(base + 4, last_word));
  const uint32_t w2 = *reinterpret_cast<const uint32_t*>(row + min(base + 8, last_word));
  const uint32_t lo = __funnelshift_r(w0, w1, shift), hi = __funnelshift_r(w1, w2, shift);
  const uint32_t p01 = __byte_perm(lo, hi, G.sel01), p23 = __byte_perm(lo, hi, G.sel23);
  hv[0] = (int)__dp2a_lo(G.coef[0], p01, 0u);
  hv[1] = (int)__dp2a_hi(G.coef[1], p01, 0u);
  hv[2] = (int)__dp2a_lo(G.coef[2], p23, 0u);
  hv[3] = (int)__dp2a_hi(G.coef[3], p23, 0u);
}

// One thread = 4 consecutive destination pixels x RZ_ROWS consecutive rows.  Block (32, 4): 128 x 32 pixels.
__global__ void __launch_bounds__(128) resize_level_kernel(const FrameGeom* __restrict__ geom, int level, BatchPlanes p,
                                                           const ResizeTap* __restrict__ taps,
                                                           const ResizeGroup* __restrict__ groups) {
  const LevelGeom& D = geom->lv[level];
  const LevelGeom& S = geom->lv[level - 1];
  const int gi = blockIdx.x * 32 + threadIdx.x;
  const int x4 = gi * 4;
  const int y0 = (blockIdx.y * 4 + threadIdx.y) * RZ_ROWS;
  const int frame = blockIdx.z;
  if (x4 >= D.w || y0 >= D.h) return;
  int spitch;
  const uint8_t* src = level_plane(p, S, level - 1, frame, &spitch);
  uint8_t* dst = p.pyr + D.plane_base * p.batch_cap + (int64_t)frame * D.plane_bytes;
  const ResizeGroup G = groups[D.group_base + gi];
  const int base = G.src_x & ~3, shift = (G.src_x & 3) * 8, last_word = (S.w - 1) & ~3;
  const ResizeTap* ty = taps + D.coef_y_base;
  int h1[4] = {0, 0, 0, 0};
  int have = -1;  // source row currently held in h1
  const int y1 = min(y0 + RZ_ROWS, D.h);
  for (int y = y0; y < y1; ++y) {
    const ResizeTap t = ty[y];
    int h0[4];
    if ((int)t.s0 == have) {
#pragma unroll
      for (int i = 0; i < 4; ++i) h0[i] = h1[i];
    } else {
      resize_hrow(src + (int64_t)t.s0 * spitch, base, last_word, shift, G, h0);
    }
    if (t.s1 != t.s0) resize_hrow(src + (int64_t)t.s1 * spitch, base, last_word, shift, G, h1);
    else {
#pragma unroll
      for (int i = 0; i < 4; ++i) h1[i] = h0[i];
    }
    have = t.s1;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int v = (((t.c0 * (h0[i] >> 4)) >> 16) + ((t.c1 * (h1[i] >> 4)) >> 16) + 2) >> 2;
      v = min(max(v, 0), 255);
      out |= (uint32_t)v << (8 * i);
    }
    *reinterpret_cast<uint32_t*>(dst + (int64_t)y * D.pitch + x4) = out;  // pitch is a multiple of 128: padding absorbs the tail
  }
}

// Generic path for scale factors whose taps do not fit the 8-byte window (scale > ~2): one thread = 4 pixels of one row.
__global__ void __launch_bounds__(256) resize_level_generic_kernel(const FrameGeom* __restrict__ geom, int level, BatchPlanes p,
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
  *reinterpret_cast<uint32_t*>(dst + (int64_t)y * D.pitch + x4) = out;
}

void launch_resize_level(const FrameGeom* d_geom, const FrameGeom& g, int level, const BatchPlanes& p,
                         const ResizeTap* d_taps, const ResizeGroup* d_groups, int nframes, cudaStream_t s) {
  const LevelGeom& D = g.lv[level];
  if (D.group_base >= 0) {
    dim3 block(32, 4);
    dim3 grid((D.w + 127) / 128, (D.h + 4 * RZ_ROWS - 1) / (4 * RZ_ROWS), nframes);
    resize_level_kernel<<<grid, block, 0, s>>>(d_geom, level, p, d_taps, d_groups);
  } else {
    dim3 block(64, 4);
    dim3 grid((D.w + 255) / 256, (D.h + 3) / 4, nframes);
    resize_level_generic_kernel<<<grid, block, 0, s>>>(d_geom, level, p, d_taps);
  }
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

constexpr int BTW = SDORB_BLUR_TW, BTH = SDORB_BLUR_TH;  // 128 x 32 output pixels per block
constexpr int B_WORDS = BTW / 4;        // output words per row (= one warp)
constexpr int B_ROWS = BTH + 6;         // source rows of a tile: y0-3 .. y0+TH+2
constexpr int B_WARPS = 4;
constexpr int B_VROWS = BTH / B_WARPS;  // output rows per thread in the vertical pass
static_assert(BTW == 128, "one warp spans a tile row");

// four pixels starting at column gx of a row, BORDER_REFLECT_101 outside [0, w).  Out of line on purpose: only the one
// or two lanes of a row that straddle the right image edge ever come here.
__device__ __noinline__ uint32_t blur_edge_word(const uint8_t* __restrict__ row, int gx, int w) {
  uint32_t v = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) v |= (uint32_t)row[reflect101(gx + b, w)] << (8 * b);
  return v;
}

__global__ void __launch_bounds__(B_WARPS * 32) blur_all_kernel(const FrameGeom* __restrict__ geom, BatchPlanes p) {
  __shared__ __align__(16) uint2 s_h[B_ROWS][B_WORDS];  // horizontal sums, four 16-bit values per entry
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int level = 0;
  while (level + 1 < geom->nlevels && (int)blockIdx.x >= geom->lv[level + 1].tile_base_blur) ++level;
  const LevelGeom& L = geom->lv[level];
  const int frame = blockIdx.y;
  const int t = blockIdx.x - L.tile_base_blur;
  const int x0 = (t % L.tiles_x_blur) * BTW, y0 = (t / L.tiles_x_blur) * BTH;
  const int w = L.w, h = L.h;
  int spitch;
  const uint8_t* src = level_plane(p, L, level, frame, &spitch);

  // ---- horizontal pass straight from global memory: warp = one source row, lane = one word of it; the neighbour words
  // come from the neighbour lanes, the two halo words from one extra load on lanes 0 and 31.
  const int gx = x0 + 4 * lane;
  const bool interior = gx + 3 < w;                         // this lane's word lies fully inside the row
  const bool edge = !interior && gx < w + 4;                // straddles the right edge (taps reach 3 px beyond it)
  const int hx = lane == 0 ? x0 - 4 : x0 + BTW;             // halo word column (lanes 0 and 31 only)
  const bool h_lane = (lane == 0 && x0 > 0) || lane == 31;  // the left halo of the first tile is mirrored from w1, w2 below
  const bool h_interior = hx + 3 < w;
  const bool h_edge = !h_interior && hx < w + 4;
  constexpr uint32_t KA = 18u | (34u << 8) | (48u << 16) | (56u << 24);  // taps -3..0
  constexpr uint32_t KB = 48u | (34u << 8) | (18u << 16);                // taps +1..+3 (+4 unused)
  constexpr int B_HROWS = (B_ROWS + B_WARPS - 1) / B_WARPS;              // source rows per warp
  const int hlast = 2 * (h - 1);
  // all loads of the warp's rows are issued before the first is used (the pass is latency-bound otherwise)
  uint32_t cw[B_HROWS], ew[B_HROWS];
  if (x0 + BTW + 4 <= w) {
    // tile (and its right halo word) entirely inside the row: plain loads, no edge tests (block-uniform branch)
    const bool halo = (lane == 0 && x0 > 0) || lane == 31;
#pragma unroll
    for (int j = 0; j < B_HROWS; ++j) {
      const int r = warp + j * B_WARPS;
      cw[j] = ew[j] = 0;
      if (r < B_ROWS) {
        int gy = abs(y0 - 3 + r);  // BORDER_REFLECT_101 of the row index, see below
        gy = max(min(gy, hlast - gy), 0);
        const uint8_t* row = src + (int64_t)gy * spitch;
        cw[j] = *reinterpret_cast<const uint32_t*>(row + gx);
        if (halo) ew[j] = *reinterpret_cast<const uint32_t*>(row + hx);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < B_HROWS; ++j) {
      const int r = warp + j * B_WARPS;
      cw[j] = ew[j] = 0;
      if (r < B_ROWS) {
        // BORDER_REFLECT_101 of the row index without branches: one reflection is enough for every level that can hold
        // a keypoint (h >= 45); the clamp only keeps degenerate levels inside their plane
        int gy = abs(y0 - 3 + r);
        gy = max(min(gy, hlast - gy), 0);
        const uint8_t* row = src + (int64_t)gy * spitch;
        if (interior) cw[j] = *reinterpret_cast<const uint32_t*>(row + gx);
        else if (edge) cw[j] = blur_edge_word(row, gx, w);
        if (h_lane) {
          if (h_interior) ew[j] = *reinterpret_cast<const uint32_t*>(row + hx);
          else if (h_edge) ew[j] = blur_edge_word(row, hx, w);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < B_HROWS; ++j) {
    const int r = warp + j * B_WARPS;
    if (r >= B_ROWS) break;  // warp-uniform
    const uint32_t w1 = cw[j];
    uint32_t w0 = __shfl_up_sync(0xffffffffu, w1, 1), w2 = __shfl_down_sync(0xffffffffu, w1, 1);
    if (lane == 0) w0 = x0 > 0 ? ew[j] : __byte_perm(w1, w2, 0x1234);  // px -4..-1 mirror px 4..1
    if (lane == 31) w2 = ew[j];
    // H[x] = sum_i K[i] * src[x + i - 3]: the seven taps of a pixel are two byte quadruples of (w0 w1) and (w1 w2)
    const uint32_t h0 = __dp4a(__byte_perm(w0, w1, 0x4321), KA, __dp4a(__byte_perm(w1, w2, 0x4321), KB, 0u));
    const uint32_t h1 = __dp4a(__byte_perm(w0, w1, 0x5432), KA, __dp4a(__byte_perm(w1, w2, 0x5432), KB, 0u));
    const uint32_t h2 = __dp4a(__byte_perm(w0, w1, 0x6543), KA, __dp4a(__byte_perm(w1, w2, 0x6543), KB, 0u));
    const uint32_t h3 = __dp4a(w1, KA, __dp4a(w2, KB, 0u));
    s_h[r][lane] = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
  }
  __syncthreads();

  // ---- vertical pass: thread = one word column x B_VROWS output rows; V[y] = sum_j K[j] * H[y + j - 3] over row pairs
  if (gx >= w) return;
  uint2 hv[B_VROWS + 6];
#pragma unroll
  for (int j = 0; j < B_VROWS + 6; ++j) hv[j] = s_h[warp * B_VROWS + j][lane];
  uint8_t* dst = p.blur + L.plane_base * p.batch_cap + (int64_t)frame * L.plane_bytes;
  constexpr uint32_t K01 = 18u | (34u << 8), K23 = 48u | (56u << 8), K45 = 48u | (34u << 8);
#pragma unroll
  for (int o = 0; o < B_VROWS; ++o) {
    const int gy = y0 + warp * B_VROWS + o;
    if (gy >= h) break;  // warp-uniform
    uint32_t acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t sel = (q & 1) ? 0x7632u : 0x5410u;
      const uint32_t a0 = q < 2 ? hv[o].x : hv[o].y, a1 = q < 2 ? hv[o + 1].x : hv[o + 1].y;
      const uint32_t a2 = q < 2 ? hv[o + 2].x : hv[o + 2].y, a3 = q < 2 ? hv[o + 3].x : hv[o + 3].y;
      const uint32_t a4 = q < 2 ? hv[o + 4].x : hv[o + 4].y, a5 = q < 2 ? hv[o + 5].x : hv[o + 5].y;
      const uint32_t a6 = q < 2 ? hv[o + 6].x : hv[o + 6].y;
      uint32_t v = __dp2a_lo(__byte_perm(a0, a1, sel), K01, 32768u);
      v = __dp2a_lo(__byte_perm(a2, a3, sel), K23, v);
      v = __dp2a_lo(__byte_perm(a4, a5, sel), K45, v);
      v = __dp2a_lo(a6, (q & 1) ? (18u << 8) : 18u, v);
      acc[q] = v;
    }
    // byte 2 of each accumulator is (v + 2^15) >> 16 (v < 2^24)
    const uint32_t out = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
    *reinterpret_cast<uint32_t*>(dst + (int64_t)gy * L.pitch + gx) = out;
  }
}

void launch_blur_all(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, int nframes, cudaStream_t s) {
  if (g.tiles_total_blur == 0) return;
  blur_all_kernel<<<dim3(g.tiles_total_blur, nframes), B_WARPS * 32, 0, s>>>(d_geom, p);
}

}  // namespace sdorb
