// kernels_fast.cu -- cell-wise cv::FAST(TYPE_9_16, nonmaxSuppression=true) for sm_100a.
//
// Reference: ComputeKeyPoints runs cv::FAST separately on every grid cell of every level
// (/root/reference/src/ORBextractor.cc:495-536).  Restated per pixel (OpenCV features2d/fast.cpp, fast_score.cpp):
//   A  = max over the 16 nine-pixel arcs of  min (v - ring),   B' = max over arcs of  min (ring - v)
//   corner  <=>  max(A,B') > th ;  score = max(A,B') - 1 ;  keypoint <=> score strictly greater than the 8
//   neighbouring scores, where neighbours outside the cell's own detectable rectangle count as 0.
// The score does not depend on the cell, so one kernel scores whole-level tiles and applies the cell rule only in the
// non-max test.  The result is a dense keypoint map per level: one byte per pixel, t = score + 1 - th for a keypoint
// and 0 elsewhere.  The selection kernel walks every cell rectangle of the map in row-major order, which IS cv::FAST's
// emission order -- no atomics, no lists to sort, no capacity to overflow, and a bit-identical result on every run.
//
// The synthetic benchmark frames are corner-dense (17 % of all pixels are FAST corners, 70 % pass the usual compass
// pre-test), so the kernel scores DENSELY and branch-free instead of compacting candidates:
//   * min(v - r) over an arc is v - max(r) over the arc, so the arc minima / maxima are taken on the ring pixels themselves,
//     two pixels at a time in the 16-bit lanes of a register;
//   * the tile is staged in shared memory as HALF-PRECISION LANES: a pixel p becomes the fp16 number 1024 + p, whose bit
//     pattern is 0x6400 | p -- one byte permute per pixel pair -- in two copies (pairs that start at even / at odd columns),
//     so every ring sample of a pixel pair (x, x + 1) is ONE aligned LDS.32 with an immediate offset: no window cutting in
//     the inner loop.  Positive halves order like their bit patterns, so the integer VIMNMX(3).U16x2 instructions work on
//     them, and integers below 2048 add and subtract exactly in fp16, so part of the network runs on the FMA pipe, which
//     the integer kernels otherwise leave idle:  (min, max)(a, b) = (a - d, b + d) with d = relu(a - b)  [HFMA2.RELU, HADD2].
//     Measured (tools/probe/score_probe.cu, profiles/r2_score_probe.log): the byte-tile / PRMT-window / all-VIMNMX form of
//     round 1 needs 359 SMSP-cycles per 128-pixel row, this form 274; an instruction with three register sources takes two
//     issue slots on this SM, which is why moving MORE of the network to the FMA pipe does not pay (same log);
//   * 64 three-input min / max equivalents per pixel pair (the arcs j-1 and j share eight pixels, see score_pair);
//   * a lane owns the pairs q and q + 32 of a 128-pixel row (conflict-free LDS), scores are kept as t = max(score + 1 - th, 0)
//     in one byte per pixel in a second tile;
//   * the 3x3 strict non-max test runs on that tile in bands: a warp walks 15 consecutive output rows, H(row) = max of a row's
//     three columns serves the rows above and below it, M(row) = max of its left and right columns the row itself (neighbour
//     windows with junk low bytes against a centre with a zero low byte, per-column / per-row cell-boundary masks); the narrow
//     tiles of the last tile column keep the per-row form (nms_pair); a tile without any score writes zeros straight away;
//   * survivors are written back as map words (coalesced 120-byte row segments); rows below the detectable area are neither
//     staged nor written (nobody reads the map there).
// A compass test on the same fp16 lanes (a 9-arc contains one of ring {0, 8} and one of ring {4, 12}) is kept only to skip
// 64-pixel half rows in which no pixel can be a corner (flat image regions).  The kernel is bound by instruction issue and
// the integer ALU pipe, not by HBM: see DESIGN.md.
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace sdorb {

constexpr int OW = SDORB_FAST_TW, OH = SDORB_FAST_TH;  // output pixels per tile: 120 x 60 (30 whole words per row)
constexpr int SWORDS = 32;                             // scored words per row (128 px: outputs + one word on each side)
constexpr int SROWS = OH + 2;                          // scored rows (outputs + 1 on each side)
constexpr int PWORDS = SWORDS + 2;                     // staged pixel words per row (scored +- 4 px)
constexpr int PROWS = SROWS + 6;                       // staged pixel rows (scored +- 3)
constexpr int HW = 2 * PWORDS;                         // words per row of one fp16-lane copy (two pixels per word)
constexpr int TWORDS = SWORDS + 2;                     // score tile pitch in words (one zero word on each side)
constexpr int NT = 128;
static_assert(OH <= 64, "row flags are two 32-bit ballots");
constexpr int SMEM_BYTES = 4 * (2 * PROWS * HW + SROWS * TWORDS + 4 + 2);  // the two pixel copies, the score tile, row flags, level + flag
constexpr int CTAS_BY_SMEM = 232448 / (SMEM_BYTES + 1024);                 // 227 KB per SM, 1 KB reserved per CTA
constexpr int CTAS_PER_SM = CTAS_BY_SMEM < 7 ? CTAS_BY_SMEM : 7;           // 7 x 128 threads x 72 registers fit the register file

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }

// ---- fp16-lane arithmetic: every lane holds an integer 0 .. 2047 as a half, so add / sub are exact
__device__ __forceinline__ uint32_t h_add(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h_sub(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h_relu_sub(uint32_t a, uint32_t b) {  // max(a - b, 0) = relu(b * -1 + a): one HFMA2.RELU
  uint32_t d;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(0xBC00BC00u), "r"(a));
  return d;
}
__device__ __forceinline__ uint32_t h_bits(float x) {  // both lanes = x as a half
  const __half2 h = __floats2half2_rn(x, x);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// One staged byte word w (pixels 4 kk .. 4 kk + 3 of a tile row) becomes two words of each copy: copy 0 word j = pixels
// (2j, 2j + 1), copy 1 word j = pixels (2j - 1, 2j), each pixel as the half 0x6400 | p.  hi_prev is what this function returned
// for the previous word of the row (its pixels 2, 3 as halves: copy 1 needs its last pixel); word 0 of copy 1 is never read.
__device__ __forceinline__ uint32_t pair23(const uint32_t w) { return prmt(w, 0x64646464u, 0x4342); }
__device__ __forceinline__ void stage_word(uint32_t (&s_h)[2][PROWS][HW], const int row, const int kk, const uint32_t w, const uint32_t hi,
                                           const uint32_t hi_prev) {
  const uint32_t c = 0x64646464u;
  *reinterpret_cast<uint2*>(&s_h[0][row][2 * kk]) = make_uint2(prmt(w, c, 0x4140), hi);
  *reinterpret_cast<uint2*>(&s_h[1][row][2 * kk]) = make_uint2(prmt(hi_prev, w, 0x3412), prmt(w, c, 0x4241));
}

// Ring column DX of the pixel pair q of a row starts at staged pixel 2q + 4 + DX: copy DX & 1, word q + col_word(DX).
__host__ __device__ constexpr int col_copy(int dx) { return dx & 1; }
__host__ __device__ constexpr int col_word(int dx) { return 2 + (dx + (dx & 1)) / 2; }
template <int DX, int DY>
__device__ __forceinline__ uint32_t ring_at(const uint32_t* __restrict__ base) {  // base = &s_h[0][scored row][q]
  return base[col_copy(DX) * PROWS * HW + (DY + 3) * HW + col_word(DX)];
}

// Compass test of a pixel pair on the four ring samples at distance 3 along the axes: non-zero iff one of the two pixels
// may be a corner, i.e. differs by more than th from one of ring {0, 8} AND from one of ring {4, 12}.
__device__ __forceinline__ uint32_t compass_pair(const uint32_t r0, const uint32_t r8, const uint32_t r4, const uint32_t r12,
                                                 const uint32_t v, const uint32_t th_h) {
  const uint32_t ev = __vmaxu2(h_relu_sub(__vmaxu2(r0, r8), v), h_relu_sub(v, __vminu2(r0, r8)));
  const uint32_t eh = __vmaxu2(h_relu_sub(__vmaxu2(r4, r12), v), h_relu_sub(v, __vminu2(r4, r12)));
  return h_relu_sub(__vminu2(ev, eh), th_h);
}

// t = max(cornerScore + 1 - th, 0) of the two pixels of a pair, as the halves 1024 + t (bits 0x6400 | t).  r[] is the ring in
// OpenCV order: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3); v the centre.
__device__ __forceinline__ uint32_t score_pair(const uint32_t (&r)[16], const uint32_t v, const uint32_t th_h, const uint32_t th_back) {
  // X = min over the 16 arcs of the arc maximum, Y = max over the arcs of the arc minimum.  The arcs starting at j-1 and
  // at j (j odd) share the eight pixels j..j+7, so  min(max(arc j-1), max(arc j)) = max(max(r[j..j+7]), min(r[j-1], r[j+8]))
  // (the grouping of OpenCV's cornerScore loop); the eight-pixel extrema are built from pair extrema (j, j+1), and the six
  // pixels j+2..j+7 serve both j and j+2.  The pair extrema (lo2, hi2) go through the FMA pipe, everything else is VIMNMX.
  uint32_t lo2[8], hi2[8], pmax[8], pmin[8];  // index q <-> j = 2q+1
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int j = 2 * q + 1;
    const uint32_t d = h_relu_sub(r[j], r[(j + 1) & 15]);
    hi2[q] = h_add(r[(j + 1) & 15], d);
    lo2[q] = h_sub(r[j], d);
    pmin[q] = __vminu2(r[j - 1], r[(j + 8) & 15]);  // joins the arc maxima
    pmax[q] = __vmaxu2(r[j - 1], r[(j + 8) & 15]);  // joins the arc minima
  }
  uint32_t wmax[8], wmin[8];
#pragma unroll
  for (int q = 0; q < 8; q += 2) {
    const uint32_t chi = __vimax3_u16x2(hi2[(q + 1) & 7], hi2[(q + 2) & 7], hi2[(q + 3) & 7]);  // pixels j+2 .. j+7
    const uint32_t clo = __vimin3_u16x2(lo2[(q + 1) & 7], lo2[(q + 2) & 7], lo2[(q + 3) & 7]);
    wmax[q] = __vimax3_u16x2(chi, hi2[q], pmin[q]);
    wmin[q] = __vimin3_u16x2(clo, lo2[q], pmax[q]);
    wmax[q + 1] = __vimax3_u16x2(chi, hi2[(q + 4) & 7], pmin[q + 1]);
    wmin[q + 1] = __vimin3_u16x2(clo, lo2[(q + 4) & 7], pmax[q + 1]);
  }
  uint32_t X = __vimin3_u16x2(wmax[0], wmax[1], wmax[2]), Y = __vimax3_u16x2(wmin[0], wmin[1], wmin[2]);
  X = __vimin3_u16x2(X, wmax[3], wmax[4]);
  Y = __vimax3_u16x2(Y, wmin[3], wmin[4]);
  X = __vimin3_u16x2(X, wmax[5], wmax[6]);
  Y = __vimax3_u16x2(Y, wmin[5], wmin[6]);
  X = __vminu2(X, wmax[7]);
  Y = __vmaxu2(Y, wmin[7]);
  const uint32_t A = h_relu_sub(v, X), B = h_relu_sub(Y, v);  // max(v - X, 0), max(Y - v, 0): non-negative halves
  return h_add(__vimax3_u16x2(A, B, th_h), th_back);          // max(A, B', th) - th + 1024
}

// The ring of pair q of a scored row in OpenCV order; base = &s_h[0][scored row][q]
__device__ __forceinline__ void load_ring(const uint32_t* __restrict__ base, uint32_t (&r)[16], const bool with_compass_samples) {
  if (with_compass_samples) {
    r[0] = ring_at<0, 3>(base);
    r[4] = ring_at<3, 0>(base);
    r[8] = ring_at<0, -3>(base);
    r[12] = ring_at<-3, 0>(base);
  }
  r[1] = ring_at<1, 3>(base);
  r[2] = ring_at<2, 2>(base);
  r[3] = ring_at<3, 1>(base);
  r[5] = ring_at<3, -1>(base);
  r[6] = ring_at<2, -2>(base);
  r[7] = ring_at<1, -3>(base);
  r[9] = ring_at<-1, -3>(base);
  r[10] = ring_at<-2, -2>(base);
  r[11] = ring_at<-3, -1>(base);
  r[13] = ring_at<-3, 1>(base);
  r[14] = ring_at<-2, 2>(base);
  r[15] = ring_at<-1, 3>(base);
}
// fp16 score lanes (0x6400 | t) -> the 16-bit score word: t of pixel 2q in the low byte, of pixel 2q + 1 in the high byte
__device__ __forceinline__ uint32_t score_word(const uint32_t t) { return prmt(t, 0u, 0x4420); }

// The score word of pair q of a scored row, or 0 when the compass test says that no lane of the warp has a corner candidate in
// this pair slot (warp-uniform decision).
__device__ __forceinline__ uint32_t score_slot_tested(const uint32_t* __restrict__ base, const bool live, bool& needed, const uint32_t th_h,
                                                      const uint32_t th_back) {
  uint32_t r[16];
  r[0] = ring_at<0, 3>(base);
  r[8] = ring_at<0, -3>(base);
  r[4] = ring_at<3, 0>(base);
  r[12] = ring_at<-3, 0>(base);
  const uint32_t v = ring_at<0, 0>(base);
  const uint32_t cand = live ? compass_pair(r[0], r[8], r[4], r[12], v, th_h) : 0u;
  needed = __any_sync(0xffffffffu, cand != 0u);
  if (!needed) return 0u;
  load_ring(base, r, false);
  return score_word(score_pair(r, v, th_h, th_back));
}

// Window of the pixel pair P of a BYTE word -- P = 0: pixels (0, 2), P = 1: pixels (1, 3) -- at horizontal offset DX, from the
// row's three words: the pixel byte sits in the HIGH byte of its 16-bit lane, the low byte is whatever lies to its left.
// u16 min / max order such lanes by the pixel byte first, and taking the high byte commutes with min / max (it is
// monotone).  Pixels two apart make the sample a plain unaligned 4-byte window of the row starting at byte 3 + P + DX of the
// 12-byte window: free when that is word aligned, and the window of (P = 0, DX) is the window of (P = 1, DX - 1).
template <int P, int DX>
__device__ __forceinline__ uint32_t window_at(const uint32_t w0, const uint32_t w1, const uint32_t w2) {
  constexpr int s = 3 + P + DX;
  static_assert(s >= 0 && s <= 8, "offset out of window");
  if (s == 0) return w0;
  if (s == 4) return w1;
  if (s == 8) return w2;
  if (s < 4) {
    constexpr uint32_t sel = s | ((s + 1) << 4) | ((s + 2) << 8) | ((s + 3) << 12);
    return prmt(w0, w1, sel);
  } else {
    constexpr uint32_t t = s >= 4 ? s - 4 : 0, sel = t | ((t + 1) << 4) | ((t + 2) << 8) | ((t + 3) << 12);
    return prmt(w1, w2, sel);
  }
}

// Strict 3x3 non-max test of the two pixels of pair P (pixels (0, 2) or (1, 3) of the word) on the score tile: returns per
// 16-bit lane a word whose LOW byte is non-zero exactly for a survivor.  T[0..2] are the score words of rows y-1, y, y+1
// (previous / own / next word); lm / rm zero the neighbours that lie in another cell.  Same lane format as the ring
// windows (window_at): the neighbours' scores sit in the high bytes with junk below them, the centre's with a zero low byte.
// With the junk of the neighbour maximum forced to 0xFF,  centre > neighbours  <=>  centre lane > neighbour lane, and then
// centre - neighbours = (difference - 1) << 8 | 1.  The nine windows of the two pairs cost nine PRMT (three per row).
template <int P>
__device__ __forceinline__ uint32_t nms_pair(const uint32_t (&T)[3][3], const uint32_t lm, const uint32_t rm) {
  const uint32_t c = prmt(T[1][1], 0u, P == 0 ? 0x2404 : 0x3414);  // score << 8 of pixels (0, 2) / (1, 3)
  const uint32_t l = window_at<P, -1>(T[1][0], T[1][1], T[1][2]), r = window_at<P, 1>(T[1][0], T[1][1], T[1][2]);
  const uint32_t u = window_at<P, 0>(T[0][0], T[0][1], T[0][2]), d = window_at<P, 0>(T[2][0], T[2][1], T[2][2]);
  const uint32_t ul = window_at<P, -1>(T[0][0], T[0][1], T[0][2]), ur = window_at<P, 1>(T[0][0], T[0][1], T[0][2]);
  const uint32_t dl = window_at<P, -1>(T[2][0], T[2][1], T[2][2]), dr = window_at<P, 1>(T[2][0], T[2][1], T[2][2]);
  // the lane masks are all-or-nothing per lane, so one AND per column of three neighbours is enough
  const uint32_t ml = __vimax3_u16x2(ul, l, dl) & lm, mr = __vimax3_u16x2(ur, r, dr) & rm;
  const uint32_t nb = __vimax3_u16x2(ml, mr, __vmaxu2(u, d)) | 0x00FF00FFu;
  return c - __vminu2(c, nb);
}

__device__ __forceinline__ int div_magic(int n, int d, uint32_t magic) {  // n / d for 0 <= n < 65536 (see geometry.cc)
  return d == 1 ? n : (int)__umulhi((uint32_t)n, magic);
}

// One tile.  NARROW = false: 32 scored words per row, one warp per row (all index arithmetic folds to constants).
// NARROW = true: the last tile column of a level, fastn_words (4, 8 or 16) words per row, several rows per warp step.
template <bool NARROW>
__device__ __forceinline__ void fast_tile(const FrameGeom* __restrict__ geom, const BatchPlanes& p, const int level, const LevelGeom& L,
                                          const int tile, uint32_t (&s_h)[2][PROWS][HW], uint32_t (&s_t)[SROWS][TWORDS],
                                          uint32_t (&s_rowflags)[2][2], uint32_t& s_any) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frame = blockIdx.y;
  const int t = tile - (NARROW ? L.tile_base_fastn : L.tile_base_fast);
  const int tx = NARROW ? L.tiles_x_fast : t % L.tiles_x_fast;  // the narrow tile follows the level's full tiles
  const int ty = NARROW ? t : t / L.tiles_x_fast;
  // A tile scores nw words per row: 32 (one warp per row), or fewer in the last tile column, where a warp then takes
  // 32 / nw rows per step so that no lane scores columns beyond the detectable area.
  const int nw = NARROW ? L.fastn_words : SWORDS;
  const int lg = NARROW ? 31 - __clz(nw) : 5, rps = 32 >> lg;  // rows per warp step
  const int k = lane & (nw - 1), sub = lane >> lg;             // this lane's scored word and its row within the step
  const int a = 12 + tx * OW;                              // first scored column (multiple of 4); outputs are words 1 .. nw-2
  const int b = SDORB_EDGE + ty * OH;                      // first output row; scored rows are [b-1, b+OH+1)
  const int w = L.w, h = L.h;
  const int th = geom->th_fast;
  int pitch;
  const uint8_t* src;
  if (level == 0) {
    pitch = p.img0_pitch;
    src = p.img0 + (int64_t)frame * p.img0_frame_stride;
  } else {
    pitch = L.pitch;
    src = p.pyr + L.plane_base * p.batch_cap + (int64_t)frame * L.plane_bytes;
  }

  // ---- stage pixels: rows b-4 .. b+OH+3, columns a-4 .. a+4*nw+3, read as aligned words (one warp per row, zero outside the
  // image) and written as fp16 lanes (stage_word).  Rows are at least 4 px inside the image at the top (b >= 19) and every plane
  // has a pitch that is a multiple of 16, so an aligned word that starts inside a row is readable; bytes beyond the row's last
  // pixel never reach a detectable pixel's ring.
  if (!NARROW) {
    static_assert(PROWS % (NT / 32) == 0 && 2 * PROWS <= 2 * NT, "staging loop shape");
    const int gx = a - 4 + 4 * lane;
    // word index of (row b - 4 + warp, column gx) in the plane (a plane is far below 2^31 bytes; pitch and gx are multiples of 4);
    // rows [0, rows_ok) of the tile exist in the image -- none for a lane whose column lies beyond the row
    const uint32_t* const plane = reinterpret_cast<const uint32_t*>(src);
    int widx = ((b - 4 + warp) * pitch + gx) >> 2;
    const int wstep = (NT / 32) * (pitch >> 2);
    // Only the staged rows that a scored row inside the detectable area reads are needed (the last tile row of a level is
    // rarely full): rows [0, rows_used); they all exist in the image (det_y1 + 3 <= h).
    const int rows_used = min(L.det_y1 - (b - 1), SROWS) + 6;
    const int rows_ok = gx < w ? rows_used : 0;
    uint32_t v[PROWS / (NT / 32)];
#pragma unroll
    for (int j = 0; j < PROWS / (NT / 32); ++j) {  // all loads of the thread in flight together
      v[j] = 0;
      if (warp + (NT / 32) * j < rows_ok) v[j] = plane[widx];
      widx += wstep;
    }
#pragma unroll
    for (int j = 0; j < PROWS / (NT / 32); ++j) {
      if (warp + (NT / 32) * j < rows_used) {  // warp-uniform
        const uint32_t hi = pair23(v[j]);
        stage_word(s_h, warp + (NT / 32) * j, lane, v[j], hi, __shfl_up_sync(0xffffffffu, hi, 1));
      }
    }
    for (int i = tid; i < 2 * PROWS; i += NT) {  // the two words to the right of the 32 (words 32, 33 of each row)
      const int r = i >> 1, kk = 32 + (i & 1);
      const int gy = b - 4 + r, hx = a - 4 + 4 * kk;
      uint32_t hv = 0, hp = 0;
      if (gy < h && hx < w) hv = *reinterpret_cast<const uint32_t*>(src + (int64_t)gy * pitch + hx);
      if (gy < h && hx - 4 < w) hp = *reinterpret_cast<const uint32_t*>(src + (int64_t)gy * pitch + hx - 4);
      stage_word(s_h, r, kk, hv, pair23(hv), pair23(hp));
    }
  } else {
    const int nwords = nw + 2, total = PROWS * nwords;
    const uint32_t inv = (65536u + nwords - 1) / nwords;  // i / nwords = (i * inv) >> 16 for i < 68 * 18 and nwords = 6, 10, 18
    for (int i0 = tid; i0 < total; i0 += 4 * NT) {
      uint32_t v[4], pv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {  // four (row, word) items of the thread in flight together
        const int i = i0 + u * NT;
        const int r = (int)(((uint32_t)i * inv) >> 16), kk = i - r * nwords;
        const int gy = b - 4 + r, gx = a - 4 + 4 * kk;
        v[u] = pv[u] = 0;
        if (i < total && gy < h) {
          const uint8_t* row = src + (int64_t)gy * pitch;
          if (gx < w) v[u] = *reinterpret_cast<const uint32_t*>(row + gx);
          if (kk > 0 && gx - 4 < w) pv[u] = *reinterpret_cast<const uint32_t*>(row + gx - 4);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * NT;
        const int r = (int)(((uint32_t)i * inv) >> 16), kk = i - r * nwords;
        if (i < total) stage_word(s_h, r, kk, v[u], pair23(v[u]), pair23(pv[u]));
      }
    }
  }
  if (tid < SROWS) {
    s_t[tid][0] = 0;
    s_t[tid][nw + 1] = 0;
  }
  if (tid == NT - 1) s_any = 0;
  if (tid < 64) {  // which vertical neighbours of output row tid lie in the same cell (none for rows below the detectable area):
    const int y = b + tid;  // one bit per row, s_rowflags[row / 32][0] for the row above, [1] for the row below
    uint32_t f = 0;
    if (tid < OH && y < L.det_y1) {
      const int ci = min(div_magic(y - SDORB_EDGE, L.cell_h, L.cell_h_magic), L.rows - 1);
      const int cy0 = SDORB_EDGE + ci * L.cell_h;
      const int cy1 = (ci == L.rows - 1) ? L.max_by : cy0 + L.cell_h;
      f = (y - 1 >= cy0 ? 1u : 0u) | (y + 1 < cy1 ? 2u : 0u);
    }
    const uint32_t up = __ballot_sync(0xffffffffu, f & 1u), down = __ballot_sync(0xffffffffu, f & 2u);
    if (lane == 0) {
      s_rowflags[warp][0] = up;
      s_rowflags[warp][1] = down;
    }
  }
  __syncthreads();

  // ---- per-thread column constants: this lane owns pixels x = xw .. xw+3 in every row it touches
  const int xw = a + 4 * k;
  const int vx1 = L.det_x1, vy1 = L.det_y1;  // detectable area is [19, det_x1) x [19, det_y1)
  uint32_t lm[2] = {0, 0}, rm[2] = {0, 0};  // 16-bit lane masks: left / right neighbour lies in the same cell
  {
    const int xs = max(xw, SDORB_EDGE);
    int cj = min(div_magic(xs - SDORB_EDGE, L.cell_w, L.cell_w_magic), L.cols - 1);
    int cx0 = SDORB_EDGE + cj * L.cell_w;
    int cx1 = (cj == L.cols - 1) ? L.max_bx : cx0 + L.cell_w;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int x = xw + q;
      if (x >= SDORB_EDGE && x < vx1) {
        if (x >= cx1) {  // stepped into the next cell
          ++cj;
          cx0 = cx1;
          cx1 = (cj == L.cols - 1) ? L.max_bx : cx0 + L.cell_w;
        }
        if (x - 1 >= cx0) lm[q & 1] |= 0xFFFFu << (16 * (q >> 1));  // pixel q is lane q >> 1 of pair q & 1
        if (x + 1 < cx1) rm[q & 1] |= 0xFFFFu << (16 * (q >> 1));
      }
    }
  }
  // ---- phase S: dense scores.  Scored row sr is image row b-1+sr and staged rows sr .. sr+6; pair q of a row is the scored
  // pixels 2q, 2q+1 = image columns a + 2q, a + 2q + 1.  A warp step covers 64 pair slots: one row of a full tile (a lane takes
  // the pairs lane and lane + 32), 64 / ppr rows of a narrow tile with ppr = 2 nw pairs per row.
  const uint32_t th_h = h_bits((float)th), th_back = h_bits((float)(1024 - th));
  const int plg = lg + 1, ppr = 2 * nw, rpw = 64 >> plg;  // pairs per row (and its log2), rows per warp step
  int sq[2], srow[2];
  uint32_t svalid[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = lane + 32 * i;
    sq[i] = idx & (ppr - 1);
    srow[i] = idx >> plg;
    const int x = a + 2 * sq[i];
    svalid[i] = (x >= SDORB_EDGE && x < vx1 ? 0x00FFu : 0u) | (x + 1 >= SDORB_EDGE && x + 1 < vx1 ? 0xFF00u : 0u);
  }
  uint16_t* const s_t16 = reinterpret_cast<uint16_t*>(&s_t[0][0]);
  const int sr_lo = max(SDORB_EDGE - (b - 1), 0), sr_hi = min(vy1 - (b - 1), SROWS);  // the scored rows inside the detectable area
  int dense_rows = 0, base = 0;
  uint32_t any_score = 0;
  // Steps with the compass test.  On corner-dense tiles the test never skips anything, so a warp of a full tile that needed
  // both of its slots in two consecutive steps stops testing for the rest of the tile (scoring a pair is always correct).
  for (; base < SROWS && (NARROW || dense_rows < 2); base += (NT / 32) * rpw) {
    bool all_needed = true;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int sr = base + warp * rpw + srow[i];
      const bool in_tile = sr < SROWS;
      const bool live = sr >= sr_lo && sr < sr_hi && svalid[i] != 0u;
      uint32_t T = 0;
      if (__any_sync(0xffffffffu, live)) {
        bool needed;
        T = score_slot_tested(&s_h[0][min(sr, SROWS - 1)][sq[i]], live, needed, th_h, th_back);
        all_needed = all_needed && needed;
        T = live ? (T & svalid[i]) : 0u;
      } else {
        all_needed = false;
      }
      any_score |= T;
      if (in_tile) s_t16[sr * (2 * TWORDS) + 2 + sq[i]] = (uint16_t)T;
    }
    dense_rows = all_needed ? dense_rows + 1 : 0;
  }
  // Dense steps of a full tile: the warp's row is the same for all lanes, and both pair slots are scored in one block.  (The
  // steps above have covered rows 0 .. 7 at least, so the rows left are at or below the first detectable row.)
  if (!NARROW) {
    for (int sr = base + warp; sr < sr_hi; sr += NT / 32) {
      const uint32_t* const row = &s_h[0][sr][lane];
      uint32_t r0[16], r1[16];
      load_ring(row, r0, true);
      load_ring(row + 32, r1, true);
      const uint32_t v0 = ring_at<0, 0>(row), v1 = ring_at<0, 0>(row + 32);
      const uint32_t T0 = score_word(score_pair(r0, v0, th_h, th_back)) & svalid[0];
      const uint32_t T1 = score_word(score_pair(r1, v1, th_h, th_back)) & svalid[1];
      any_score |= T0 | T1;
      s_t16[sr * (2 * TWORDS) + 2 + lane] = (uint16_t)T0;
      s_t16[sr * (2 * TWORDS) + 2 + 32 + lane] = (uint16_t)T1;
    }
    for (int sr = max(sr_hi, base) + warp; sr < SROWS; sr += NT / 32) {  // rows below the detectable area score 0
      s_t16[sr * (2 * TWORDS) + 2 + lane] = 0;
      s_t16[sr * (2 * TWORDS) + 2 + 32 + lane] = 0;
    }
  }
  if (__any_sync(0xffffffffu, any_score != 0u) && lane == 0) s_any = 1u;
  __syncthreads();

  // ---- phase N: cell-bounded strict non-max suppression on the score tile; the survivors' t bytes go to the map
  uint8_t* map = p.nms + L.plane_base * p.batch_cap + (int64_t)frame * L.plane_bytes;
  const bool out_lane = k >= 1 && k <= nw - 2 && xw < L.pitch;
  if (!NARROW) {
    // Full tile: a warp takes a band of OH / 4 consecutive output rows and walks down the score tile once.  The 3x3 maximum is
    // taken row-wise: H(row) = max of the row's three columns (left / right masked by the cell's column limits) serves the
    // output rows above and below it, M(row) = max of its left and right columns serves the row itself, so a score row costs
    // three window PRMT and four min / max once instead of once per output row that looks at it.
    constexpr int BAND = OH / (NT / 32);
    static_assert(OH % (NT / 32) == 0, "bands of whole rows");
    const int o0 = warp * BAND;  // first output row of the band = score row o0 + 1
    // Row flags of the band as bits 0 .. BAND-1.  (A second, flag-free copy of the loop below for bands inside one cell row made
    // the kernel 7 % slower, and rolling the loop up did not make it faster: profiles/r2_define_probe.log.)
    const uint64_t upw = ((uint64_t)s_rowflags[1][0] << 32) | s_rowflags[0][0], dnw = ((uint64_t)s_rowflags[1][1] << 32) | s_rowflags[0][1];
    const uint32_t band_mask = (1u << BAND) - 1u, upb = (uint32_t)(upw >> o0) & band_mask, dnb = (uint32_t)(dnw >> o0) & band_mask;
    uint8_t* mrow = map + (int64_t)(b + o0) * L.pitch + xw;
    const int rows_left = L.det_y1 - (b + o0);  // output rows of the band inside the detectable area (nobody reads the map below it)
    if (s_any == 0u) {  // a tile without any corner (flat image region): the map is all zero
#pragma unroll 1
      for (int j = 0; j < BAND; ++j) {
        if (j < rows_left && out_lane) *reinterpret_cast<uint32_t*>(mrow) = 0u;
        mrow += L.pitch;
      }
      return;
    }
    {
      uint32_t H[BAND + 2][2], M[BAND + 2][2], Cw[BAND + 2];
#pragma unroll
      for (int i = 0; i < BAND + 2; ++i) {
        const int sr = o0 + i;  // score rows o0 .. o0 + BAND + 1; lanes 0 / 31 see the tile's zero words as T0 / T2
        const uint32_t T0 = s_t[sr][k], T1 = s_t[sr][k + 1], T2 = s_t[sr][k + 2];
        Cw[i] = T1;
        const uint32_t l0 = window_at<0, -1>(T0, T1, T2) & lm[0], l1 = window_at<1, -1>(T0, T1, T2) & lm[1];
        const uint32_t r0 = window_at<0, 1>(T0, T1, T2) & rm[0], r1 = window_at<1, 1>(T0, T1, T2) & rm[1];
        M[i][0] = __vmaxu2(l0, r0);
        M[i][1] = __vmaxu2(l1, r1);
        H[i][0] = __vimax3_u16x2(l0, r0, window_at<0, 0>(T0, T1, T2));
        H[i][1] = __vimax3_u16x2(l1, r1, window_at<1, 0>(T0, T1, T2));
        if (i >= 2) {  // output row o0 + j: score row sr - 1, the rows above / below are entries i - 2 / i
          const int j = i - 2;
          if (j >= rows_left) break;  // warp-uniform: the rest of the band lies below the detectable area
          const uint32_t cw = Cw[i - 1];
          const bool up = ((upb >> j) & 1u) != 0, down = ((dnb >> j) & 1u) != 0;
          uint32_t res[2];
#pragma unroll
          for (int P = 0; P < 2; ++P) {
            const uint32_t nb = __vimax3_u16x2(up ? H[i - 2][P] : 0u, down ? H[i][P] : 0u, M[i - 1][P]) | 0x00FF00FFu;
            const uint32_t c = prmt(cw, 0u, P == 0 ? 0x2404 : 0x3414);  // score << 8 of pixels (0, 2) / (1, 3)
            res[P] = c - __vminu2(c, nb);                               // low byte 1 for a survivor, else the lane is 0
          }
          const uint32_t keep_bytes = prmt(res[0], res[1], 0x6240) * 0xFFu;
          if (j < rows_left && out_lane) *reinterpret_cast<uint32_t*>(mrow) = cw & keep_bytes;
          mrow += L.pitch;
        }
      }
    }
    return;
  }
  for (int base = 0; base < OH; base += (NT / 32) * rps) {
    const int orow = base + warp * rps + sub;
    const int y = b + orow;
    const bool active = orow < OH && y < h;
    const int sr = min(orow, OH - 1) + 1;
    const uint32_t cw = active ? s_t[sr][k + 1] : 0u;
    uint32_t keep_bytes = 0;
    if (__any_sync(0xffffffffu, cw != 0)) {
      const int fr = sr - 1;  // bit 0: row above is in the same cell, bit 1: row below is
      const uint32_t rowflags = ((s_rowflags[fr >> 5][0] >> (fr & 31)) & 1u) | (((s_rowflags[fr >> 5][1] >> (fr & 31)) & 1u) << 1);
      uint32_t Tn[3][3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        Tn[0][j] = (rowflags & 1u) ? s_t[sr - 1][k + j] : 0u;
        Tn[1][j] = s_t[sr][k + j];
        Tn[2][j] = (rowflags & 2u) ? s_t[sr + 1][k + j] : 0u;
      }
      // the low bytes of the four difference lanes in pixel order: 1 for a survivor, else 0; times 255 = a byte mask
      const uint32_t d4 = prmt(nms_pair<0>(Tn, lm[0], rm[0]), nms_pair<1>(Tn, lm[1], rm[1]), 0x6240);
      keep_bytes = d4 * 0xFFu;
    }
    if (active && out_lane) *reinterpret_cast<uint32_t*>(map + (int64_t)y * L.pitch + xw) = cw & keep_bytes;
  }
}

// One launch for all tiles of all levels: blocks [0, tiles_total_fast) are the full tiles, the rest the narrow tiles of the
// last tile columns (their own launch ran at 57 % ALU utilisation against 89 % for the full tiles: short, few CTAs, a tail of
// its own; behind the full tiles in the same grid they fill the SMs the last full tiles leave free).
__global__ void __launch_bounds__(NT, CTAS_PER_SM) fast_tiles_kernel(const FrameGeom* __restrict__ geom, BatchPlanes p, int n_full) {
  __shared__ __align__(16) uint32_t s_h[2][PROWS][HW];
  __shared__ __align__(16) uint32_t s_t[SROWS][TWORDS];
  __shared__ uint32_t s_rowflags[2][2];
  __shared__ int s_level;
  __shared__ uint32_t s_any;  // does the tile hold a non-zero score at all?
  pdl_enter();
  const bool narrow = (int)blockIdx.x >= n_full;  // CTA-uniform
  const int t = narrow ? (int)blockIdx.x - n_full : (int)blockIdx.x;
  if (threadIdx.x == 0) {
    int l = 0;
    while (l + 1 < geom->nlevels && t >= (narrow ? geom->lv[l + 1].tile_base_fastn : geom->lv[l + 1].tile_base_fast)) ++l;
    s_level = l;
  }
  __syncthreads();
  const int level = s_level;
  if (narrow)
    fast_tile<true>(geom, p, level, geom->lv[level], t, s_h, s_t, s_rowflags, s_any);
  else
    fast_tile<false>(geom, p, level, geom->lv[level], t, s_h, s_t, s_rowflags, s_any);
}

void launch_fast_all(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, int nframes, cudaStream_t s) {
  const int total = g.tiles_total_fast + g.tiles_total_fastn;
  if (total > 0) launch_pdl(fast_tiles_kernel, dim3(total, nframes), dim3(NT), 0, s, d_geom, p, g.tiles_total_fast);
}

}  // namespace sdorb
