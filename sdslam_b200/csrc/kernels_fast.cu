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
//   * one thread owns one 32-bit word = 4 pixels; a warp owns 128 pixels of one row; everything is read from a
//     shared-memory tile as aligned words and shuffled into place with PRMT;
//   * min(v - r) over an arc is v - max(r) over the arc, so the arc minima / maxima are taken on the ring bytes
//     themselves, two pixels at a time in 16-bit lanes (VIMNMX.U16x2, 3-input forms): 64 min/max per pixel pair;
//     the two pixels of a pair lie two apart, each in the high byte of its lane with its left neighbour as the low byte,
//     so a ring sample is a plain unaligned window of the row (23 PRMT per word for its 32 samples, see ring_at) and
//     needs no masking: u16 order is byte order up to ties, and ties do not change the high byte of a min / max;
//   * scores are kept as t = max(score + 1 - th, 0) in one byte per pixel; the 3x3 strict non-max test runs on the
//     same kind of lanes (neighbour windows with junk low bytes against a centre with a zero low byte, see nms_pair)
//     with per-column / per-row cell-boundary masks;
//   * survivors are written back as map words (coalesced 120-byte row segments).
// A cheap 4-pixel SWAR compass test (VABSDIFF4) is kept only to skip pixel pairs no lane of the warp needs (flat image
// regions).  The kernel is bound by the integer ALU pipe (min/max, PRMT), not by HBM: see DESIGN.md.
#include "kernels.cuh"

namespace sdorb {

constexpr int OW = SDORB_FAST_TW, OH = SDORB_FAST_TH;  // output pixels per tile: 120 x 60 (30 whole words per row)
constexpr int SWORDS = 32;                             // scored words per row (128 px: outputs + one word on each side)
constexpr int SROWS = OH + 2;                          // scored rows (outputs + 1 on each side)
constexpr int PWORDS = SWORDS + 2;                     // staged pixel words per row (scored +- 4 px)
constexpr int PROWS = SROWS + 6;                       // staged pixel rows (scored +- 3)
constexpr int TWORDS = SWORDS + 2;                     // score tile pitch in words (one zero word on each side)
constexpr int NT = 128;

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
// per-byte (a > th) in bit 7 of each byte; C prepared by the caller from th
__device__ __forceinline__ uint32_t gt_th(uint32_t a, uint32_t C, bool th_high) {
  const uint32_t t = (a & 0x7f7f7f7fu) + C;
  return th_high ? (t & a) : (t | a);
}

// Ring sample of the pixel pair P of a word -- P = 0: pixels (0, 2), P = 1: pixels (1, 3) -- at horizontal offset DX, from
// the row's three words: the pixel byte sits in the HIGH byte of its 16-bit lane, the low byte is whatever lies to its left.
// u16 min / max order such lanes by the pixel byte first, and taking the high byte commutes with min / max (it is
// monotone), so the arc extrema come out right in the high bytes.  Pixels two apart make the sample a plain unaligned
// 4-byte window of the row starting at byte 3 + P + DX of the 12-byte window: free when that is word aligned, and the
// window of (P = 0, DX) is the window of (P = 1, DX - 1) -- 23 PRMT per word for its 32 ring samples instead of 32.
template <int P, int DX>
__device__ __forceinline__ uint32_t ring_at(const uint32_t w0, const uint32_t w1, const uint32_t w2) {
  constexpr int s = 3 + P + DX;
  static_assert(s >= 0 && s <= 8, "offset out of window");
  if (s == 0) return w0;
  if (s == 4) return w1;
  if (s == 8) return w2;
  if (s < 4) {
    constexpr uint32_t sel = s | ((s + 1) << 4) | ((s + 2) << 8) | ((s + 3) << 12);
    return prmt(w0, w1, sel);
  } else {
    constexpr uint32_t t = s - 4, sel = t | ((t + 1) << 4) | ((t + 2) << 8) | ((t + 3) << 12);
    return prmt(w1, w2, sel);
  }
}

// t = max(cornerScore + 1 - th, 0) for the two pixels of pair P (pixels (0, 2) or (1, 3) of the word), as two 16-bit lanes.  W[dy+3][0..2] are the staged
// words of rows y-3..y+3 (previous / own / next word).
template <int P>
__device__ __forceinline__ uint32_t score_pair(const uint32_t (&W)[7][3], const uint32_t th2) {
  uint32_t r[16];
  // ring in OpenCV order: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
  r[0] = ring_at<P, 0>(W[6][0], W[6][1], W[6][2]);
  r[1] = ring_at<P, 1>(W[6][0], W[6][1], W[6][2]);
  r[2] = ring_at<P, 2>(W[5][0], W[5][1], W[5][2]);
  r[3] = ring_at<P, 3>(W[4][0], W[4][1], W[4][2]);
  r[4] = ring_at<P, 3>(W[3][0], W[3][1], W[3][2]);
  r[5] = ring_at<P, 3>(W[2][0], W[2][1], W[2][2]);
  r[6] = ring_at<P, 2>(W[1][0], W[1][1], W[1][2]);
  r[7] = ring_at<P, 1>(W[0][0], W[0][1], W[0][2]);
  r[8] = ring_at<P, 0>(W[0][0], W[0][1], W[0][2]);
  r[9] = ring_at<P, -1>(W[0][0], W[0][1], W[0][2]);
  r[10] = ring_at<P, -2>(W[1][0], W[1][1], W[1][2]);
  r[11] = ring_at<P, -3>(W[2][0], W[2][1], W[2][2]);
  r[12] = ring_at<P, -3>(W[3][0], W[3][1], W[3][2]);
  r[13] = ring_at<P, -3>(W[4][0], W[4][1], W[4][2]);
  r[14] = ring_at<P, -2>(W[5][0], W[5][1], W[5][2]);
  r[15] = ring_at<P, -1>(W[6][0], W[6][1], W[6][2]);
  // X = min over the 16 arcs of the arc maximum, Y = max over the arcs of the arc minimum.  The arcs starting at j-1 and
  // at j (j odd) share the eight pixels j..j+7, so  min(max(arc j-1), max(arc j)) = max(max(r[j..j+7]), min(r[j-1], r[j+8]))
  // (the grouping of OpenCV's cornerScore loop); the eight-pixel extrema are built from pair extrema (j, j+1), and the six
  // pixels j+2..j+7 serve both j and j+2.  64 three-input min/max per pixel pair instead of 80.
  uint32_t lo2[8], hi2[8], pmax[8], pmin[8];  // index q <-> j = 2q+1
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int j = 2 * q + 1;
    lo2[q] = __vminu2(r[j], r[(j + 1) & 15]);
    hi2[q] = __vmaxu2(r[j], r[(j + 1) & 15]);
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
  // A = v - X, B' = Y - v per lane, biased by 256 so that the lanes never borrow
  const uint32_t Xc = prmt(X, 0u, 0x4341), Yc = prmt(Y, 0u, 0x4341);
  const uint32_t Vc = prmt(W[3][1], 0u, P == 0 ? 0x4240 : 0x4341);  // pixels (0, 2) / (1, 3) of the own word
  const uint32_t A = Vc + 0x01000100u - Xc, B = Yc + 0x01000100u - Vc;
  return __vimax3_u16x2(A, B, th2) - th2;     // (max(A, B') - th) if positive, else 0;  th2 = (th + 256) per lane
}

// Strict 3x3 non-max test of the two pixels of pair P (pixels (0, 2) or (1, 3) of the word) on the score tile: returns per
// 16-bit lane a word whose LOW byte is non-zero exactly for a survivor.  T[0..2] are the score words of rows y-1, y, y+1
// (previous / own / next word); lm / rm zero the neighbours that lie in another cell.  Same lane format as the ring
// samples (ring_at): the neighbours' scores sit in the high bytes with junk below them, the centre's with a zero low byte.
// With the junk of the neighbour maximum forced to 0xFF,  centre > neighbours  <=>  centre lane > neighbour lane, and then
// centre - neighbours = (difference - 1) << 8 | 1.  The nine windows of the two pairs cost nine PRMT (three per row).
template <int P>
__device__ __forceinline__ uint32_t nms_pair(const uint32_t (&T)[3][3], const uint32_t lm, const uint32_t rm) {
  const uint32_t c = prmt(T[1][1], 0u, P == 0 ? 0x2404 : 0x3414);  // score << 8 of pixels (0, 2) / (1, 3)
  const uint32_t l = ring_at<P, -1>(T[1][0], T[1][1], T[1][2]), r = ring_at<P, 1>(T[1][0], T[1][1], T[1][2]);
  const uint32_t u = ring_at<P, 0>(T[0][0], T[0][1], T[0][2]), d = ring_at<P, 0>(T[2][0], T[2][1], T[2][2]);
  const uint32_t ul = ring_at<P, -1>(T[0][0], T[0][1], T[0][2]), ur = ring_at<P, 1>(T[0][0], T[0][1], T[0][2]);
  const uint32_t dl = ring_at<P, -1>(T[2][0], T[2][1], T[2][2]), dr = ring_at<P, 1>(T[2][0], T[2][1], T[2][2]);
  // the lane masks are all-or-nothing per lane, so one AND per column of three neighbours is enough
  const uint32_t ml = __vimax3_u16x2(ul, l, dl) & lm, mr = __vimax3_u16x2(ur, r, dr) & rm;
  const uint32_t nb = __vimax3_u16x2(ml, mr, __vmaxu2(u, d)) | 0x00FF00FFu;
  return c - __vminu2(c, nb);
}

__device__ __forceinline__ int div_magic(int n, int d, uint32_t magic) {  // n / d for 0 <= n < 65536 (see geometry.cc)
  return d == 1 ? n : (int)__umulhi((uint32_t)n, magic);
}

// One tile.  NARROW = false: 32 scored words per row, one warp per row (all index arithmetic folds to constants).
// NARROW = true: the last tile column of a level, fast_last_words (4, 8 or 16) words per row, 32 / nw rows per warp step.
template <bool NARROW>
__device__ __forceinline__ void fast_tile(const FrameGeom* __restrict__ geom, const BatchPlanes& p, const int level, const LevelGeom& L,
                                          const int tile, uint32_t (&s_pix)[PROWS][PWORDS], uint32_t (&s_t)[SROWS][TWORDS],
                                          uint32_t (&s_rowflags)[OH]) {
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

  // ---- stage pixels: rows b-4 .. b+OH+3, columns a-4 .. a+4*nw+3 as aligned words, one warp per row; zero outside the image
  // Rows are at least 4 px inside the image at the top (b >= 19) and every plane has a pitch that is a multiple of 16, so
  // an aligned word that starts inside a row is readable; bytes beyond the row's last pixel never reach a detectable pixel's ring.
  if (!NARROW) {
    static_assert(PROWS % (NT / 32) == 0 && 2 * PROWS <= 2 * NT, "staging loop shape");
    const int gx = a - 4 + 4 * lane;
    const uint8_t* colp = src + gx;
    const bool col_ok = gx < w;
    uint32_t v[PROWS / (NT / 32)];
#pragma unroll
    for (int j = 0; j < PROWS / (NT / 32); ++j) {  // all loads of the thread in flight together
      const int gy = b - 4 + warp + (NT / 32) * j;
      v[j] = 0;
      if (col_ok && gy < h) v[j] = *reinterpret_cast<const uint32_t*>(colp + (int64_t)gy * pitch);
    }
#pragma unroll
    for (int j = 0; j < PROWS / (NT / 32); ++j) s_pix[warp + (NT / 32) * j][lane] = v[j];
    for (int i = tid; i < 2 * PROWS; i += NT) {  // the two words to the right of the 32 (words 32, 33 of each row)
      const int r = i >> 1, kk = 32 + (i & 1);
      const int gy = b - 4 + r, hx = a - 4 + 4 * kk;
      uint32_t hv = 0;
      if (gy < h && hx < w) hv = *reinterpret_cast<const uint32_t*>(src + (int64_t)gy * pitch + hx);
      s_pix[r][kk] = hv;
    }
  } else
  for (int r = warp; r < PROWS; r += NT / 32) {
    const int gy = b - 4 + r;
    const uint8_t* row = src + (int64_t)gy * pitch;
    const bool row_ok = gy >= 0 && gy < h;
    for (int kk = lane; kk < nw + 2; kk += 32) {
      const int gx = a - 4 + 4 * kk;
      uint32_t v = 0;
      if (row_ok && gx < w) {
        if (gx + 4 <= w) {
          v = *reinterpret_cast<const uint32_t*>(row + gx);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (gx + q < w) v |= (uint32_t)row[gx + q] << (8 * q);
        }
      }
      s_pix[r][kk] = v;
    }
  }
  if (tid < SROWS) {
    s_t[tid][0] = 0;
    s_t[tid][nw + 1] = 0;
  }
  if (tid < OH) {  // which vertical neighbours of output row tid lie in the same cell (0 for rows below the detectable area)
    const int y = b + tid;
    uint32_t f = 0;
    if (y < L.det_y1) {
      const int ci = min(div_magic(y - SDORB_EDGE, L.cell_h, L.cell_h_magic), L.rows - 1);
      const int cy0 = SDORB_EDGE + ci * L.cell_h;
      const int cy1 = (ci == L.rows - 1) ? L.max_by : cy0 + L.cell_h;
      f = (y - 1 >= cy0 ? 1u : 0u) | (y + 1 < cy1 ? 2u : 0u);
    }
    s_rowflags[tid] = f;
  }
  __syncthreads();

  // ---- per-thread column constants: this lane owns pixels x = xw .. xw+3 in every row it touches
  const int xw = a + 4 * k;
  const int vx1 = L.det_x1, vy1 = L.det_y1;  // detectable area is [19, det_x1) x [19, det_y1)
  uint32_t valid_cols = 0;                  // byte mask: pixel may carry a score
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
        valid_cols |= 0xFFu << (8 * q);
        if (x - 1 >= cx0) lm[q & 1] |= 0xFFFFu << (16 * (q >> 1));  // pixel q is lane q >> 1 of pair q & 1
        if (x + 1 < cx1) rm[q & 1] |= 0xFFFFu << (16 * (q >> 1));
      }
    }
  }
  const bool th_high = th >= 128;
  const uint32_t C = (uint32_t)(127 - (th_high ? th - 128 : th)) * 0x01010101u;
  const uint32_t th2 = (uint32_t)(th + 256) * 0x00010001u;

  // ---- phase S: dense scores.  Scored row sr is image row b-1+sr and staged row sr+3; scored word k is staged word k+1.
  bool use_compass = true;  // warp-uniform
  int dense_rows = 0;
  for (int base = 0; base < SROWS; base += (NT / 32) * rps) {
    const int sr = base + warp * rps + sub;
    const int gy = b - 1 + sr;
    const bool in_tile = sr < SROWS;
    const bool active = in_tile && gy >= SDORB_EDGE && gy < vy1;
    uint32_t T = 0;
    if (__any_sync(0xffffffffu, active)) {
      const int srl = min(sr, SROWS - 1);
      uint32_t W[7][3];
#pragma unroll
      for (int dy = 0; dy < 7; ++dy) {
        W[dy][0] = s_pix[srl + dy][k];
        W[dy][1] = s_pix[srl + dy][k + 1];
        W[dy][2] = s_pix[srl + dy][k + 2];
      }
      // compass pre-test on 4 pixels (a 9-arc contains one of ring {0,8} and one of ring {4,12}): lets the warp skip the
      // pixel pairs that no lane needs (flat image regions).  On corner-dense tiles it never skips anything, so a warp that
      // needed both pairs in two consecutive rows stops testing for the rest of the tile (scoring a pair is always correct).
      const uint32_t live = active ? valid_cols : 0u;
      bool need0 = true, need1 = true;
      if (use_compass) {
        const uint32_t c = W[3][1];
        const uint32_t lf = prmt(W[3][0], c, 0x4321), rt = prmt(c, W[3][2], 0x6543);
        const uint32_t fv = gt_th(__vabsdiffu4(W[0][1], c), C, th_high) | gt_th(__vabsdiffu4(W[6][1], c), C, th_high);
        const uint32_t fh = gt_th(__vabsdiffu4(lf, c), C, th_high) | gt_th(__vabsdiffu4(rt, c), C, th_high);
        const uint32_t cand = fv & fh & 0x80808080u & live;
        need0 = __any_sync(0xffffffffu, cand & 0x00800080u);  // pixels 0, 2
        need1 = __any_sync(0xffffffffu, cand & 0x80008000u);  // pixels 1, 3
        dense_rows = (need0 && need1) ? dense_rows + 1 : 0;
        use_compass = dense_rows < 2;
      }
      uint32_t t0 = 0, t1 = 0;
      if (need0 && need1) {  // one block: the two pairs share nine of their ring windows
        t0 = score_pair<0>(W, th2);
        t1 = score_pair<1>(W, th2);
      } else {
        if (need0) t0 = score_pair<0>(W, th2);
        if (need1) t1 = score_pair<1>(W, th2);
      }
      T = prmt(t0, t1, 0x6240) & live;  // bytes t(0), t(1), t(2), t(3)
    }
    if (in_tile) s_t[sr][k + 1] = T;
  }
  __syncthreads();

  // ---- phase N: cell-bounded strict non-max suppression on the score tile; the survivors' t bytes go to the map
  uint8_t* map = p.nms + L.plane_base * p.batch_cap + (int64_t)frame * L.plane_bytes;
  const bool out_lane = k >= 1 && k <= nw - 2 && xw < L.pitch;
  for (int base = 0; base < OH; base += (NT / 32) * rps) {
    const int orow = base + warp * rps + sub;
    const int y = b + orow;
    const bool active = orow < OH && y < h;
    const int sr = min(orow, OH - 1) + 1;
    const uint32_t cw = active ? s_t[sr][k + 1] : 0u;
    uint32_t keep_bytes = 0;
    if (__any_sync(0xffffffffu, cw != 0)) {
      const uint32_t rowflags = s_rowflags[sr - 1];  // bit 0: row above is in the same cell, bit 1: row below is
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
__global__ void __launch_bounds__(NT, 1024 / NT) fast_tiles_kernel(const FrameGeom* __restrict__ geom, BatchPlanes p, int n_full) {
  __shared__ __align__(16) uint32_t s_pix[PROWS][PWORDS];
  __shared__ __align__(16) uint32_t s_t[SROWS][TWORDS];
  __shared__ uint32_t s_rowflags[OH];
  __shared__ int s_level;
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
    fast_tile<true>(geom, p, level, geom->lv[level], t, s_pix, s_t, s_rowflags);
  else
    fast_tile<false>(geom, p, level, geom->lv[level], t, s_pix, s_t, s_rowflags);
}

void launch_fast_all(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, int nframes, cudaStream_t s) {
  const int total = g.tiles_total_fast + g.tiles_total_fastn;
  if (total > 0) launch_pdl(fast_tiles_kernel, dim3(total, nframes), dim3(NT), 0, s, d_geom, p, g.tiles_total_fast);
}

}  // namespace sdorb
