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
constexpr int RZ_ROWS = 16;  // consecutive destination rows per thread (a source row feeds up to two destination rows)

// predicated word load: keeps the loads of a row straight-line under loop-invariant predicates.  When the predicate is
// off the result is unspecified -- the callers only switch off words none of whose bytes can reach a result.
__device__ __forceinline__ uint32_t ldg_word_if(const uint8_t* ptr, int on) {
  uint32_t v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.global.nc.b32 %0, [%1];\n\t}" : "=r"(v) : "l"(ptr), "r"(on));
  return v;
}
// the same through the coherent path: for a kernel that reads what its own CTA wrote earlier (resize_tail_kernel)
__device__ __forceinline__ uint32_t ld_word_if(const uint8_t* ptr, int on) {
  uint32_t v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.global.b32 %0, [%1];\n\t}" : "=r"(v) : "l"(ptr), "r"(on) : "memory");
  return v;
}

// The three source words that hold every tap of four destination pixels in one source row.
struct ResizeRaw {
  uint32_t w0, w1, w2;
};
template <bool NC = true>
__device__ __forceinline__ void resize_fetch(const uint8_t* __restrict__ row, int on, int ld1, int ld2, ResizeRaw& r) {
  if (NC) {
    r.w0 = ldg_word_if(row, on);
    r.w1 = ldg_word_if(row + 4, on & ld1);
    r.w2 = ldg_word_if(row + 8, on & ld2);
  } else {
    r.w0 = ld_word_if(row, on);
    r.w1 = ld_word_if(row + 4, on & ld1);
    r.w2 = ld_word_if(row + 8, on & ld2);
  }
}
// horizontal pass of one source row for four destination pixels, already shifted: a[i] = (c0*s0 + c1*s1) >> 4
__device__ __forceinline__ void resize_hrow(const ResizeRaw& r, int shift, const ResizeGroup& G, uint32_t (&a)[4]) {
  const uint32_t lo = __funnelshift_r(r.w0, r.w1, shift), hi = __funnelshift_r(r.w1, r.w2, shift);
  const uint32_t p01 = __byte_perm(lo, hi, G.sel01), p23 = __byte_perm(lo, hi, G.sel23);
  a[0] = __dp2a_lo(G.coef[0], p01, 0u) >> 4;
  a[1] = __dp2a_hi(G.coef[1], p01, 0u) >> 4;
  a[2] = __dp2a_lo(G.coef[2], p23, 0u) >> 4;
  a[3] = __dp2a_hi(G.coef[3], p23, 0u) >> 4;
}

// One thread = 4 consecutive destination pixels x RZ_ROWS consecutive rows.  Block (32, 4): 128 x 64 pixels.
// The kernel waits on memory, not on issue slots (ncu: long-scoreboard stalls): a thread walks down its source rows and
// a row's words are needed as soon as they are asked for.  So the walk is software-pipelined: the source words of
// destination row j+1 (and the vertical taps of row j+2) are requested before row j is computed.
// Coefficients are non-negative on this path (checked when the groups are built), so every intermediate fits an
// unsigned lane and the result needs no clamp: the two products of a pixel are cut to their upper halves two pixels at
// a time (PRMT), summed with the rounding constant in 16-bit lanes and shifted as one word.
__global__ void __launch_bounds__(128) resize_level_kernel(const FrameGeom* __restrict__ geom, int level, BatchPlanes p,
                                                           const ResizeTap* __restrict__ taps,
                                                           const ResizeGroup* __restrict__ groups) {
  pdl_enter();
  const LevelGeom& D = geom->lv[level];
  const LevelGeom& S = geom->lv[level - 1];
  const int gi = blockIdx.x * 32 + threadIdx.x;
  const int x4 = gi * 4;
  const int y0 = (blockIdx.y * 4 + threadIdx.y) * RZ_ROWS;
  const int frame = blockIdx.z;
  const int dh = D.h;
  if (x4 >= D.w || y0 >= dh) return;
  int spitch;
  const uint8_t* src = level_plane(p, S, level - 1, frame, &spitch);
  const int dpitch = D.pitch;
  uint8_t* dst = p.pyr + D.plane_base * p.batch_cap + (int64_t)frame * D.plane_bytes + (int64_t)y0 * dpitch + x4;
  const ResizeGroup G = groups[D.group_base + gi];
  const int base = G.src_x & ~3, shift = (G.src_x & 3) * 8, last_word = (S.w - 1) & ~3;
  // the eight bytes from src_x on hold every tap; words past the row's last word hold none
  const int ld1 = base + 4 <= last_word, ld2 = shift != 0 && base + 8 <= last_word;
  src += base;
  const uint2* ty = reinterpret_cast<const uint2*>(taps + D.coef_y_base + y0);  // {s0 | s1 << 16, c0 | c1 << 16}
  const int jlast = min(RZ_ROWS, dh - y0) - 1;  // rows past it repeat its taps: their requests are harmless re-reads
  uint2 t = ty[0], t1 = ty[min(1, jlast)];
  uint32_t s0 = t.x & 0xFFFFu, s1 = t.x >> 16;
  int need0 = 1, need1 = s1 != s0;
  ResizeRaw r0, r1;
  resize_fetch(src + (uint64_t)(s0 * (uint32_t)spitch), need0, ld1, ld2, r0);
  resize_fetch(src + (uint64_t)(s1 * (uint32_t)spitch), need1, ld1, ld2, r1);
  uint32_t a0[4], a1[4] = {0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < RZ_ROWS; ++j) {
    if (j > jlast) break;
    // requests of row j+1: a source row is new unless it is the one row j leaves behind (s1) / equals its partner
    const uint2 tn = t1;
    t1 = ty[min(j + 2, jlast)];
    const uint32_t n0 = tn.x & 0xFFFFu, n1 = tn.x >> 16;
    const int nneed0 = n0 != s1, nneed1 = n1 != n0;
    ResizeRaw q0, q1;
    resize_fetch(src + (uint64_t)(n0 * (uint32_t)spitch), nneed0, ld1, ld2, q0);
    resize_fetch(src + (uint64_t)(n1 * (uint32_t)spitch), nneed1, ld1, ld2, q1);
    // row j
    const uint32_t c0 = t.y & 0xFFFFu, c1 = t.y >> 16;
    if (need0) {
      resize_hrow(r0, shift, G, a0);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) a0[i] = a1[i];
    }
    if (need1) {
      resize_hrow(r1, shift, G, a1);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) a1[i] = a0[i];
    }
    // v = (((c0 * a0) >> 16) + ((c1 * a1) >> 16) + 2) >> 2
    const uint32_t u01 = __byte_perm(c0 * a0[0], c0 * a0[1], 0x7632), u23 = __byte_perm(c0 * a0[2], c0 * a0[3], 0x7632);
    const uint32_t l01 = __byte_perm(c1 * a1[0], c1 * a1[1], 0x7632), l23 = __byte_perm(c1 * a1[2], c1 * a1[3], 0x7632);
    const uint32_t v01 = (u01 + l01 + 0x00020002u) >> 2, v23 = (u23 + l23 + 0x00020002u) >> 2;  // lanes <= 1022: no carry
    *reinterpret_cast<uint32_t*>(dst + (uint64_t)((uint32_t)j * (uint32_t)dpitch)) = __byte_perm(v01, v23, 0x6420);  // the pitch absorbs the tail
    t = tn;
    s0 = n0;
    s1 = n1;
    need0 = nneed0;
    need1 = nneed1;
    r0 = q0;
    r1 = q1;
  }
}

// The same arithmetic with every source word of the thread requested up front.  The row-walk kernel above waits for
// memory once per destination row (ncu source page: all its stall samples sit on the first use of a row's words, and one row
// of look-ahead does not cover the latency with the warps an SM holds).  The source rows of RZP_ROWS consecutive destination
// rows are one contiguous range of at most KMAX rows, and when the image shrinks every source row is the lower tap of at
// most one destination row (both checked on the host per level, LevelGeom::rz_span).  So this variant issues all
// 3 * KMAX independent loads of a thread before it touches any, then walks the source rows in registers: H of source row
// k is computed once, and the destination row that ends on row k -- looked up in a per-warp table in shared memory, filled
// from the vertical taps by the lanes that hold them -- is emitted from H(k-1) / H(k).
constexpr int RZP_ROWS = 8;
// One warp: 128 destination pixels (word column block bx) x RZP_ROWS rows starting at y0 of `level` of `frame`.  emit = KMAX
// int4 of this warp's shared memory.  NC: the source level was written by an earlier kernel (non-coherent loads are fine).
template <int KMAX, bool NC>
__device__ __forceinline__ void resize_pre_tile(const FrameGeom* __restrict__ geom, const int level, const BatchPlanes& p,
                                                const ResizeTap* __restrict__ taps, const ResizeGroup* __restrict__ groups, const int bx,
                                                const int y0, const int frame, int4* emit, const int lane) {
  const LevelGeom& D = geom->lv[level];
  const LevelGeom& S = geom->lv[level - 1];
  // lanes beyond the row redo the row's last group: same loads, same bytes stored to the same place
  const int gi = min(bx * 32 + lane, ((D.w + 3) >> 2) - 1);
  const int dh = D.h;
  if (y0 >= dh) return;
  const int jn = min(RZP_ROWS, dh - y0);
  // vertical taps {s0 | s1 << 16, c0 | c1 << 16} of row y0 + lane, for the lanes < jn
  const uint2 tl = reinterpret_cast<const uint2*>(taps + D.coef_y_base + y0)[min(lane, jn - 1)];
  const uint32_t row0 = __shfl_sync(0xffffffffu, tl.x, 0) & 0xFFFFu, rowl = __shfl_sync(0xffffffffu, tl.x, jn - 1) >> 16;
  __syncwarp();  // the previous tile of this warp is done with the table
  if (lane < KMAX) emit[lane] = make_int4(-1, 0, 0, 0);
  __syncwarp();
  // a row with both taps on one source row (s0 == s1, the clamped first row of a level) is always the first of its group of
  // eight (checked on the host), i.e. ends on source row k = 0, where "the row before" does not exist anyway
  if (lane < jn) emit[(tl.x >> 16) - row0] = make_int4(lane, (int)(tl.y & 0xFFFFu), (int)(tl.y >> 16), 0);
  __syncwarp();
  int spitch;
  const uint8_t* src = level_plane(p, S, level - 1, frame, &spitch);
  const int dpitch = D.pitch;
  uint8_t* dst = p.pyr + D.plane_base * p.batch_cap + (int64_t)frame * D.plane_bytes + (int64_t)y0 * dpitch + gi * 4;
  const ResizeGroup G = groups[D.group_base + gi];
  const int base = G.src_x & ~3, shift = (G.src_x & 3) * 8, last_word = (S.w - 1) & ~3;
  const int ld1 = base + 4 <= last_word, ld2 = shift != 0 && base + 8 <= last_word;
  src += base + (uint64_t)(row0 * (uint32_t)spitch);
  ResizeRaw raw[KMAX];
  {
    const uint8_t* rp = src;
#pragma unroll
    for (int k = 0; k < KMAX; ++k, rp += spitch) resize_fetch<NC>(rp, row0 + k <= rowl, ld1, ld2, raw[k]);
  }
  uint32_t ap[4] = {0, 0, 0, 0}, ac[4] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (row0 + k > rowl) break;
#pragma unroll
    for (int i = 0; i < 4; ++i) ap[i] = ac[i];
    resize_hrow(raw[k], shift, G, ac);
    const int4 e = emit[k];
    if (e.x >= 0) {
      const uint32_t c0 = (uint32_t)e.y, c1 = (uint32_t)e.z;
      const uint32_t(&au)[4] = k == 0 ? ac : ap;  // the upper tap's row: H(k-1), or H(0) itself for the clamped first row
      // v = (((c0 * a0) >> 16) + ((c1 * a1) >> 16) + 2) >> 2
      const uint32_t u01 = __byte_perm(c0 * au[0], c0 * au[1], 0x7632), u23 = __byte_perm(c0 * au[2], c0 * au[3], 0x7632);
      const uint32_t l01 = __byte_perm(c1 * ac[0], c1 * ac[1], 0x7632), l23 = __byte_perm(c1 * ac[2], c1 * ac[3], 0x7632);
      const uint32_t v01 = (u01 + l01 + 0x00020002u) >> 2, v23 = (u23 + l23 + 0x00020002u) >> 2;  // lanes <= 1022: no carry
      *reinterpret_cast<uint32_t*>(dst + (uint32_t)e.x * (uint32_t)dpitch) = __byte_perm(v01, v23, 0x6420);  // the pitch absorbs the tail
    }
  }
}

template <int KMAX>
__global__ void __launch_bounds__(128) resize_level_pre_kernel(const FrameGeom* __restrict__ geom, int level, BatchPlanes p,
                                                               const ResizeTap* __restrict__ taps,
                                                               const ResizeGroup* __restrict__ groups) {
  __shared__ int4 s_emit[4][KMAX];  // per warp and source row k: {destination row or -1, c0, c1, -}
  pdl_enter();
  resize_pre_tile<KMAX, true>(geom, level, p, taps, groups, blockIdx.x, (blockIdx.y * 4 + threadIdx.y) * RZP_ROWS, blockIdx.z,
                              s_emit[threadIdx.y], threadIdx.x);
}

// The small upper levels of the chained pyramid in ONE launch: level l needs all of level l - 1, so a CTA owns a frame and
// walks the levels first_level .. nlevels - 1 with a block barrier between them (its own global writes are visible to its own
// threads after the barrier; the loads of these levels take the coherent path).  At 640x480 levels 4-7 together are 11 % of the
// level-0 pixels: four launches of 18-38 us per 512 frames at 14-22 % of the DRAM throughput become one, and the single-frame
// call loses three kernel boundaries.
constexpr int RZT_WARPS = 16;
template <int KMAX>
__global__ void __launch_bounds__(RZT_WARPS * 32, 2) resize_tail_kernel(const FrameGeom* __restrict__ geom, int first_level, BatchPlanes p,
                                                                     const ResizeTap* __restrict__ taps,
                                                                     const ResizeGroup* __restrict__ groups) {
  __shared__ int4 s_emit[RZT_WARPS][KMAX];
  pdl_enter();
  const int frame = blockIdx.x, lane = threadIdx.x, warp = threadIdx.y;
  for (int level = first_level; level < geom->nlevels; ++level) {
    const LevelGeom& D = geom->lv[level];
    const int tiles_x = (D.w + 127) >> 7, tiles_y = (D.h + RZP_ROWS - 1) / RZP_ROWS;
    for (int t = warp; t < tiles_x * tiles_y; t += RZT_WARPS) {
      const int ty = t / tiles_x, bx = t - ty * tiles_x;
      resize_pre_tile<KMAX, false>(geom, level, p, taps, groups, bx, ty * RZP_ROWS, frame, s_emit[warp], lane);
    }
    __syncthreads();
  }
}

// Generic path for scale factors whose taps do not fit the 8-byte window (scale > ~2): one thread = 4 pixels of one row.
__global__ void __launch_bounds__(256) resize_level_generic_kernel(const FrameGeom* __restrict__ geom, int level, BatchPlanes p,
                                                                   const ResizeTap* __restrict__ taps) {
  pdl_enter();
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

// First level of the fused tail launch: the smallest l >= 2 from which every level is small (at most 96 K pixels) and runs on
// the pre-loading path; nlevels when there is no such tail of at least two levels.
int resize_tail_first_level(const FrameGeom& g) {
  int first = g.nlevels;
  for (int l = g.nlevels - 1; l >= 2; --l) {
    const LevelGeom& D = g.lv[l];
    if (!(D.group_base >= 0 && D.rz_span > 0 && D.rz_span <= 11 && D.w * D.h <= 96 * 1024)) break;
    first = l;
  }
  return g.nlevels - first >= 2 ? first : g.nlevels;
}

void launch_resize_tail(const FrameGeom* d_geom, const FrameGeom& g, int first_level, const BatchPlanes& p, const ResizeTap* d_taps,
                        const ResizeGroup* d_groups, int nframes, cudaStream_t s) {
  launch_pdl(resize_tail_kernel<11>, dim3(nframes), dim3(32, RZT_WARPS), 0, s, d_geom, first_level, p, d_taps, d_groups);
}

void launch_resize_level(const FrameGeom* d_geom, const FrameGeom& g, int level, const BatchPlanes& p,
                         const ResizeTap* d_taps, const ResizeGroup* d_groups, int nframes, cudaStream_t s) {
  const LevelGeom& D = g.lv[level];
  if (D.group_base >= 0 && D.rz_span > 0 && D.rz_span <= 11) {
    dim3 block(32, 4);
    dim3 grid((D.w + 127) / 128, (D.h + 4 * RZP_ROWS - 1) / (4 * RZP_ROWS), nframes);
    launch_pdl(resize_level_pre_kernel<11>, grid, block, 0, s, d_geom, level, p, d_taps, d_groups);
  } else if (D.group_base >= 0) {
    dim3 block(32, 4);
    dim3 grid((D.w + 127) / 128, (D.h + 4 * RZ_ROWS - 1) / (4 * RZ_ROWS), nframes);
    launch_pdl(resize_level_kernel, grid, block, 0, s, d_geom, level, p, d_taps, d_groups);
  } else {
    dim3 block(64, 4);
    dim3 grid((D.w + 255) / 256, (D.h + 3) / 4, nframes);
    launch_pdl(resize_level_generic_kernel, grid, block, 0, s, d_geom, level, p, d_taps);
  }
}

// ------------------------------------------------------------------------------------------------ blur
// One warp = one strip of 128 columns x SDORB_BLUR_TH output rows; lane = one aligned word (4 pixels) of every row.
// The warp walks down its strip: per source row three word loads (previous / own / next word), the horizontal 7-tap
// sums of its four pixels as two IDP.4A each, and -- the vertical filter being symmetric -- the output row whose seventh
// source row just arrived as  56*H3 + 48*(H2+H4) + 34*(H1+H5) + 18*(H0+H6) + 2^15  over a seven-row register window
// (the loop is unrolled by seven so that the window slots are fixed registers).  Nothing goes through shared memory and
// there is no barrier; per row and word: 3 LDG, 9 PRMT, 8 IDP.4A, 12 IADD, 16 IMAD, 1 STG plus addressing.
// BORDER_REFLECT_101: rows by index arithmetic; columns in registers -- the words that reach beyond the row's ends are
// rebuilt with PRMT from the two words next to the edge (selectors per level from the host, LevelGeom::blur_sel_*).
constexpr int BTW = SDORB_BLUR_TW, BTH = SDORB_BLUR_TH;
constexpr int B_WARPS = 4;
static_assert(BTW == 128, "one warp spans a strip row");

struct BlurTileBases {
  int nlevels;
  int bth;                         // output rows per strip: SDORB_BLUR_TH for batches, short strips for a handful of frames
  int base[SDORB_MAX_LEVELS + 1];  // first strip tile of each level; base[nlevels] = total
};

struct BlurLane {
  const uint8_t* src;  // this lane's own word of row 0
  int spitch, hlast, y_in0;
  int ld0, ld1, ld2;         // which of the three words exist inside the row
  bool edge_warp, is_last, is_pre, is_first;
  uint32_t sel_last, sel_beyond;
};

__device__ __forceinline__ void blur_load_row(const BlurLane& B, int r, uint32_t (&w)[3]) {
  int gy = abs(B.y_in0 + r);           // BORDER_REFLECT_101 of the row index: one reflection is enough for every level
  gy = max(min(gy, B.hlast - gy), 0);  // that can hold a keypoint; the clamp only keeps the address inside the plane
  const uint8_t* row = B.src + (uint64_t)((uint32_t)gy * (uint32_t)B.spitch);  // a plane is smaller than 2^31 bytes
  w[1] = ldg_word_if(row, B.ld1);
  w[0] = ldg_word_if(row - 4, B.ld0);
  w[2] = ldg_word_if(row + 4, B.ld2);
}

// horizontal sums of the lane's four pixels of one source row
template <bool EDGE>
__device__ __forceinline__ void blur_hrow(const BlurLane& B, const uint32_t (&win)[3], uint32_t (&h)[4]) {
  constexpr uint32_t KA = 18u | (34u << 8) | (48u << 16) | (56u << 24);  // taps -3..0
  constexpr uint32_t KB = 48u | (34u << 8) | (18u << 16);                // taps +1..+3
  uint32_t w0 = win[0], w1 = win[1], w2 = win[2];
  if (EDGE) {
    // right end: the last word of the row with its bytes beyond the row mirrored in, and the word after it
    const uint32_t a = B.is_last ? w0 : w1, b = B.is_last ? w1 : w2;
    const uint32_t lw = __byte_perm(a, b, B.sel_last), bw = __byte_perm(a, b, B.sel_beyond);
    if (B.is_last) {
      w1 = lw;
      w2 = bw;
    } else if (B.is_pre) {
      w2 = lw;
    }
    if (B.is_first) w0 = __byte_perm(w1, w1, 0x1233);  // px -3..-1 mirror px 3..1
  }
  // every sum carries +128: the vertical weights add up to 256, which makes it the rounding constant 2^15 of the output
  h[0] = __dp4a(__byte_perm(w0, w1, 0x4321), KA, __dp4a(__byte_perm(w1, w2, 0x4321), KB, 128u));
  h[1] = __dp4a(__byte_perm(w0, w1, 0x5432), KA, __dp4a(__byte_perm(w1, w2, 0x5432), KB, 128u));
  h[2] = __dp4a(__byte_perm(w0, w1, 0x6543), KA, __dp4a(__byte_perm(w1, w2, 0x6543), KB, 128u));
  h[3] = __dp4a(w1, KA, __dp4a(w2, KB, 128u));
}

// The vertical filter works on ROW PAIRS: the horizontal sums of source rows 2j and 2j + 1 (16 bits each, the +128 included) share
// one register per pixel, and IDP.2A multiplies such a pair with two of the seven weights at once.  Output row o needs source rows
// o .. o + 6: an even o = 2m takes the pairs m .. m+3 with the weights (k0,k1) (k2,k3) (k4,k5) (k6,0), an odd o = 2m + 1 the same
// pairs with (0,k0) (k1,k2) (k3,k4) (k5,k6) -- four IDP.2A per pixel instead of four IMAD + three IADD, and a window of four pair
// slots instead of seven rows.  S = slot of the newest pair (j); the older pairs j-3, j-2, j-1 sit in slots S+1, S+2, S+3 (mod 4).
template <int S, bool ODD>
__device__ __forceinline__ uint32_t blur_vrow(const uint32_t (&P)[4][4]) {
  constexpr uint32_t K0 = 18, K1 = 34, K2 = 48, K3 = 56;  // k = K0 K1 K2 K3 K2 K1 K0
  // weights of the four pairs, two pairs per constant (low / high half-word: __dp2a_lo / __dp2a_hi)
  constexpr uint32_t WA = ODD ? ((0u | K0 << 8) | (K1 | K2 << 8) << 16) : ((K0 | K1 << 8) | (K2 | K3 << 8) << 16);
  constexpr uint32_t WB = ODD ? ((K3 | K2 << 8) | (K1 | K0 << 8) << 16) : ((K2 | K1 << 8) | (K0 | 0u << 8) << 16);
  uint32_t acc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t v = __dp2a_lo(P[(S + 1) % 4][q], WA, 0u);
    v = __dp2a_hi(P[(S + 2) % 4][q], WA, v);
    v = __dp2a_lo(P[(S + 3) % 4][q], WB, v);
    acc[q] = __dp2a_hi(P[S][q], WB, v);  // includes 256 * 128 = 2^15
  }
  // byte 2 of each accumulator is (v + 2^15) >> 16 (v < 2^24)
  return __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
}

template <bool EDGE>
__device__ __forceinline__ void blur_walk(const BlurLane& B, uint8_t* __restrict__ dst, const int dpitch, const int nout) {
  uint32_t P[4][4];
  uint32_t ra[3], rb[3], na[3], nb[3];  // the rows of the pair in work, and of the next pair (loads run one pair ahead)
  blur_load_row(B, 0, ra);
  blur_load_row(B, 1, rb);
  // source rows R, R + 1 (already in ra / rb) become the pair in slot S; the rows R + 2, R + 3 are requested first
#define SDORB_BLUR_PAIR(S, R)                                                   \
  blur_load_row(B, (R) + 2, na);                                                \
  blur_load_row(B, (R) + 3, nb);                                                \
  {                                                                             \
    uint32_t h0[4], h1[4];                                                      \
    blur_hrow<EDGE>(B, ra, h0);                                                 \
    blur_hrow<EDGE>(B, rb, h1);                                                 \
    _Pragma("unroll") for (int q = 0; q < 4; ++q) P[S][q] = __byte_perm(h0[q], h1[q], 0x5410); \
  }                                                                             \
  ra[0] = na[0], ra[1] = na[1], ra[2] = na[2];                                  \
  rb[0] = nb[0], rb[1] = nb[1], rb[2] = nb[2];
  // source rows 0..5 of the strip fill three pair slots; from then on every pair completes two output rows
  SDORB_BLUR_PAIR(0, 0)
  SDORB_BLUR_PAIR(1, 2)
  SDORB_BLUR_PAIR(2, 4)
  // output rows o + 2K, o + 2K + 1 from the pair of source rows o + 2K + 6, o + 2K + 7 (slot (K + 3) mod 4); rows past the strip are
  // read clamped and not stored
#define SDORB_BLUR_STEP(K)                                                                                                          \
  SDORB_BLUR_PAIR((K + 3) % 4, o + 2 * K + 6)                                                                                       \
  if (o + 2 * K < nout && B.ld1)                                                                                                    \
    *reinterpret_cast<uint32_t*>(dst + (uint64_t)((uint32_t)(o + 2 * K) * (uint32_t)dpitch)) = blur_vrow<(K + 3) % 4, false>(P);     \
  if (o + 2 * K + 1 < nout && B.ld1)                                                                                                \
    *reinterpret_cast<uint32_t*>(dst + (uint64_t)((uint32_t)(o + 2 * K + 1) * (uint32_t)dpitch)) = blur_vrow<(K + 3) % 4, true>(P);
  for (int o = 0; o < nout; o += 8) {
    SDORB_BLUR_STEP(0)
    SDORB_BLUR_STEP(1)
    SDORB_BLUR_STEP(2)
    SDORB_BLUR_STEP(3)
  }
#undef SDORB_BLUR_STEP
#undef SDORB_BLUR_PAIR
}

#ifndef SDORB_BLUR_MIN_CTAS
#define SDORB_BLUR_MIN_CTAS 5
#endif
__global__ void __launch_bounds__(B_WARPS * 32, SDORB_BLUR_MIN_CTAS) blur_all_kernel(const FrameGeom* __restrict__ geom, BatchPlanes p, BlurTileBases tb) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x * B_WARPS + (threadIdx.x >> 5);
  if (tile >= tb.base[tb.nlevels]) return;
  int level = 0;
  while (tile >= tb.base[level + 1]) ++level;
  const LevelGeom& L = geom->lv[level];
  const int frame = blockIdx.y;
  const int t = tile - tb.base[level];
  const int tx = t % L.tiles_x_blur, ty = t / L.tiles_x_blur;
  const int x0 = tx * BTW, y0 = ty * tb.bth;
  const int w = L.w, h = L.h;
  const int gx = x0 + 4 * lane;
  const int lastw = (w - 1) & ~3;
  BlurLane B;
  const uint8_t* plane = level_plane(p, L, level, frame, &B.spitch);
  B.src = plane + gx;
  B.hlast = 2 * (h - 1);
  B.y_in0 = y0 - 3;
  B.ld1 = gx <= lastw;
  B.ld0 = B.ld1 && gx > 0;
  B.ld2 = gx + 4 <= lastw;
  B.is_last = gx == lastw;
  B.is_pre = gx + 4 == lastw;
  B.is_first = gx == 0;
  B.edge_warp = x0 == 0 || x0 + BTW + 4 > lastw;
  B.sel_last = L.blur_sel_last;
  B.sel_beyond = L.blur_sel_beyond;
  const int nout = min(tb.bth, h - y0);
  uint8_t* dst = p.blur + L.plane_base * p.batch_cap + (int64_t)frame * L.plane_bytes + (int64_t)y0 * L.pitch + gx;
  const int dpitch = L.pitch;

  if (B.edge_warp) blur_walk<true>(B, dst, dpitch, nout);  // warp-uniform
  else blur_walk<false>(B, dst, dpitch, nout);
}

// ---- imagePyramid for a batch (src/ORBextractor.cc:620-621: the pyramid is an output of operator()).  The scratch planes are
// level-major with padded pitches; the caller's slab is frame-major with tightly packed rows: level l of frame f starts at
// slab + f * frame_bytes + offset[l] (offsets and frame_bytes are multiples of 16, sdorb_pyramid_layout).  One thread moves 16
// destination bytes with one aligned store; its source bytes may straddle a row end, so they are fetched byte by byte (L1).
struct PackLayout {
  int nlevels, first_level;
  int64_t frame_bytes;
  int64_t offset[SDORB_MAX_LEVELS];
  int chunk_base[SDORB_MAX_LEVELS + 1];  // first 16-byte chunk of each level in the per-frame chunk numbering
};

__global__ void __launch_bounds__(256) pack_pyramid_kernel(const FrameGeom* __restrict__ geom, BatchPlanes p, PackLayout lay,
                                                           uint8_t* __restrict__ dst) {
  pdl_enter();
  const int chunk = blockIdx.x * 256 + threadIdx.x, frame = blockIdx.y;
  if (chunk >= lay.chunk_base[lay.nlevels]) return;
  int level = lay.first_level;
  while (chunk >= lay.chunk_base[level + 1]) ++level;
  const LevelGeom& L = geom->lv[level];
  int spitch;
  const uint8_t* src = level_plane(p, L, level, frame, &spitch);
  const int i0 = (chunk - lay.chunk_base[level]) * 16, total = L.w * L.h;
  int y = i0 / L.w, x = i0 - y * L.w;
  uint32_t wds[4] = {0u, 0u, 0u, 0u};
  const uint8_t* row = src + (int64_t)y * spitch;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (i0 + k < total) wds[k >> 2] |= (uint32_t)row[x] << (8 * (k & 3));
    if (++x == L.w) {
      x = 0;
      row += spitch;
    }
  }
  *reinterpret_cast<uint4*>(dst + (int64_t)frame * lay.frame_bytes + lay.offset[level] + i0) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
}

// The single-frame call returns imagePyramid the way the reference holds it: every level inside a buffer padded by 19 pixels
// of BORDER_REFLECT_101 (src/ORBextractor.cc:684-696).  This kernel lays frame 0's levels out in exactly that form, level after
// level ((w + 38) x (h + 38) bytes each, 16-byte aligned starts), so that the host side is one plain copy per level.
struct PadLayout {
  int nlevels, first_level;
  int64_t offset[SDORB_MAX_LEVELS];
  int chunk_base[SDORB_MAX_LEVELS + 1];
};

__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while ((unsigned)p >= (unsigned)len) p = p < 0 ? -p : 2 * (len - 1) - p;
  return p;
}

__global__ void __launch_bounds__(256) pack_padded_kernel(const FrameGeom* __restrict__ geom, BatchPlanes p, PadLayout lay,
                                                          uint8_t* __restrict__ dst) {
  pdl_enter();
  const int chunk = blockIdx.x * 256 + threadIdx.x;
  if (chunk >= lay.chunk_base[lay.nlevels]) return;
  int level = lay.first_level;
  while (chunk >= lay.chunk_base[level + 1]) ++level;
  const LevelGeom& L = geom->lv[level];
  int spitch;
  const uint8_t* src = level_plane(p, L, level, 0, &spitch);
  const int pw = L.w + 2 * SDORB_EDGE, total = pw * (L.h + 2 * SDORB_EDGE);
  const int i0 = (chunk - lay.chunk_base[level]) * 16;
  int Y = i0 / pw, X = i0 - Y * pw;
  uint32_t wds[4] = {0u, 0u, 0u, 0u};
  const uint8_t* row = src + (int64_t)reflect101(Y - SDORB_EDGE, L.h) * spitch;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (i0 + k < total) wds[k >> 2] |= (uint32_t)row[reflect101(X - SDORB_EDGE, L.w)] << (8 * (k & 3));
    if (++X == pw) {
      X = 0;
      ++Y;
      row = src + (int64_t)reflect101(Y - SDORB_EDGE, L.h) * spitch;
    }
  }
  *reinterpret_cast<uint4*>(dst + lay.offset[level] + i0) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
}

int64_t padded_pyramid_layout(const FrameGeom& g, int64_t* offset) {
  int64_t off = 0;
  for (int l = 0; l < g.nlevels; ++l) {
    if (offset) offset[l] = off;
    off += ((int64_t)(g.lv[l].w + 2 * SDORB_EDGE) * (g.lv[l].h + 2 * SDORB_EDGE) + 15) / 16 * 16;
  }
  return off;
}

void launch_pack_padded(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, int first_level, uint8_t* dst,
                        cudaStream_t s) {
  PadLayout lay{};
  lay.nlevels = g.nlevels;
  lay.first_level = first_level;
  padded_pyramid_layout(g, lay.offset);
  int c = 0;
  for (int l = 0; l <= SDORB_MAX_LEVELS; ++l) {
    lay.chunk_base[l] = c;
    if (l < g.nlevels && l >= first_level) c += ((g.lv[l].w + 2 * SDORB_EDGE) * (g.lv[l].h + 2 * SDORB_EDGE) + 15) / 16;
  }
  if (c == 0) return;
  pack_padded_kernel<<<(c + 255) / 256, 256, 0, s>>>(d_geom, p, lay, dst);
}

void pyramid_layout(const FrameGeom& g, int64_t* offset, int64_t* frame_bytes) {
  int64_t off = 0;
  for (int l = 0; l < g.nlevels; ++l) {
    offset[l] = off;
    off += ((int64_t)g.lv[l].w * g.lv[l].h + 15) / 16 * 16;
  }
  *frame_bytes = off;
}

void launch_pack_pyramid(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, int first_level, uint8_t* dst, int nframes,
                         cudaStream_t s) {
  PackLayout lay{};
  lay.nlevels = g.nlevels;
  lay.first_level = first_level;
  pyramid_layout(g, lay.offset, &lay.frame_bytes);
  int c = 0;
  for (int l = 0; l <= g.nlevels; ++l) {
    lay.chunk_base[l] = c;
    if (l < g.nlevels && l >= first_level) c += (g.lv[l].w * g.lv[l].h + 15) / 16;
  }
  for (int l = g.nlevels + 1; l <= SDORB_MAX_LEVELS; ++l) lay.chunk_base[l] = c;
  if (c == 0 || nframes <= 0) return;
  launch_pdl(pack_pyramid_kernel, dim3((c + 255) / 256, nframes), dim3(256), 0, s, d_geom, p, lay, dst);
}

void launch_blur_all(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, int nframes, cudaStream_t s) {
  if (g.tiles_total_blur == 0) return;
  // A strip is walked by one warp, seven halo rows on top of its output rows: 128-row strips for batches (5 % halo, plenty of
  // warps), 28-row strips when there are only a few frames (the single-frame call had 21 warps on the whole GPU otherwise).
  BlurTileBases tb;
  tb.nlevels = g.nlevels;
  tb.bth = nframes <= 8 ? 28 : BTH;
  int total = 0;
  for (int l = 0; l <= SDORB_MAX_LEVELS; ++l) {
    tb.base[l] = total;
    if (l < g.nlevels && g.lv[l].tiles_y_blur > 0) total += g.lv[l].tiles_x_blur * ((g.lv[l].h + tb.bth - 1) / tb.bth);
  }
  launch_pdl(blur_all_kernel, dim3((total + B_WARPS - 1) / B_WARPS, nframes), dim3(B_WARPS * 32), 0, s, d_geom, p, tb);
}

}  // namespace sdorb
