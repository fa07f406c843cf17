// kernels_fast.cu -- cell-wise cv::FAST(TYPE_9_16, nonmaxSuppression=true) for sm_100a.
//
// Reference: ComputeKeyPoints runs cv::FAST separately on every grid cell of every level
// (/root/reference/src/ORBextractor.cc:495-536).  Restated per pixel (OpenCV features2d/fast.cpp, fast_score.cpp):
//   d[k] = v - ring[k];  A = max over the 16 nine-pixel arcs of min(d),  B = max over arcs of min(-d)
//   corner  <=>  max(A,B) > th ;  score = max(A,B) - 1 ;  keypoint <=> score strictly greater than the 8
//   neighbouring scores, where neighbours outside the cell's own detectable rectangle count as 0.
// The score does not depend on the cell, so one kernel scores whole-level tiles and applies the cell rule
// only in the non-max test.  Keypoints are appended to per-cell lists as SDORB_ENTRY(y,x,score); the order of
// appends is arbitrary, the selection kernel sorts each list back into FAST's row-major emission order.
//
// Work per tile (128x32 outputs, 256 threads), three phases separated by block barriers:
//   A  SWAR pre-test on 4 pixels per 32-bit word: VABSDIFF4 against the 4 compass ring pixels; any 9-arc
//      contains one pixel of {0,8} and one of {4,12}, so  (|d0|>th or |d8|>th) and (|d4|>th or |d12|>th)
//      is necessary.  Survivors are compacted into a shared-memory candidate list.
//   B  one thread per candidate: the exact score with 3-input min/max (VIMNMX3) over the ring.
//   C  one thread per candidate: cell-bounded strict non-max test on the shared score tile, append.
#include "kernels.cuh"

namespace sdorb {

constexpr int TW = SDORB_FAST_TW, TH = SDORB_FAST_TH;
constexpr int PW = TW + 8;   // staged pixel columns x0-4 .. x0+TW+3
constexpr int PH = TH + 8;   // staged pixel rows    y0-4 .. y0+TH+3
constexpr int SW = TW + 2;   // scored columns x0-1 .. x0+TW
constexpr int SH = TH + 2;
constexpr int SP = 132;      // score tile pitch
constexpr int NT = 256;

// per-byte (a > th) in bit 7 of each byte; C and hi prepared by the caller from th
__device__ __forceinline__ uint32_t gt_th(uint32_t a, uint32_t C, bool th_high) {
  const uint32_t t = (a & 0x7f7f7f7fu) + C;
  return th_high ? (t & a) : (t | a);
}

__device__ __forceinline__ int corner_strength(const uint8_t* c, int pitch) {
  // ring in OpenCV order: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
  const int v = c[0];
  int d[16];
  d[0] = v - c[3 * pitch];
  d[1] = v - c[3 * pitch + 1];
  d[2] = v - c[2 * pitch + 2];
  d[3] = v - c[pitch + 3];
  d[4] = v - c[3];
  d[5] = v - c[-pitch + 3];
  d[6] = v - c[-2 * pitch + 2];
  d[7] = v - c[-3 * pitch + 1];
  d[8] = v - c[-3 * pitch];
  d[9] = v - c[-3 * pitch - 1];
  d[10] = v - c[-2 * pitch - 2];
  d[11] = v - c[-pitch - 3];
  d[12] = v - c[-3];
  d[13] = v - c[pitch - 3];
  d[14] = v - c[2 * pitch - 2];
  d[15] = v - c[3 * pitch - 1];
  int lo3[16], hi3[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    lo3[k] = __vimin3_s32(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    hi3[k] = __vimax3_s32(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
  }
  int A = -256, B = 256;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    A = max(A, __vimin3_s32(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]));  // min over arc k..k+8
    B = min(B, __vimax3_s32(hi3[k], hi3[(k + 3) & 15], hi3[(k + 6) & 15]));  // max over arc k..k+8
  }
  return max(A, -B);
}

__global__ void __launch_bounds__(NT) fast_all_kernel(const FrameGeom* __restrict__ geom, BatchPlanes p, SelectBuffers buf) {
  __shared__ __align__(16) uint8_t s_pix[PH][PW];
  __shared__ __align__(16) uint8_t s_score[SH][SP];
  __shared__ uint16_t s_cand[SW * SH];
  __shared__ int s_ncand;
  __shared__ int s_level;
  const int tid = threadIdx.x;
  if (tid == 0) {
    int l = 0;
    while (l + 1 < geom->nlevels && (int)blockIdx.x >= geom->lv[l + 1].tile_base_fast) ++l;
    s_level = l;
    s_ncand = 0;
  }
  __syncthreads();
  const int level = s_level;
  const LevelGeom& L = geom->lv[level];
  const int frame = blockIdx.y;
  const int t = blockIdx.x - L.tile_base_fast;
  const int x0 = 16 + (t % L.tiles_x_fast) * TW, y0 = 16 + (t / L.tiles_x_fast) * TH;
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

  // ---- stage pixels, clear scores
  for (int i = tid; i < PH * (PW / 4); i += NT) {
    const int r = i / (PW / 4), k = i % (PW / 4);
    const int gy = y0 - 4 + r, gx = x0 - 4 + 4 * k;
    uint32_t v = 0;
    if (gy >= 0 && gy < h) {
      const uint8_t* row = src + (int64_t)gy * pitch;
      if (gx + 4 <= w) {
        v = *reinterpret_cast<const uint32_t*>(row + gx);
      } else {
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (gx + b < w) v |= (uint32_t)row[gx + b] << (8 * b);
      }
    }
    *reinterpret_cast<uint32_t*>(&s_pix[r][4 * k]) = v;
  }
  for (int i = tid; i < SH * SP / 4; i += NT) reinterpret_cast<uint32_t*>(&s_score[0][0])[i] = 0;
  __syncthreads();

  // ---- phase A: compass pre-test, 4 pixels per word.  Scored pixels: x in [x0-1, x0+TW+1) and inside
  // the level's detectable area [19, det_x1) x [19, det_y1).
  const bool th_high = th >= 128;
  const uint32_t C = (uint32_t)(127 - (th_high ? th - 128 : th)) * 0x01010101u;
  const int vx0 = max(x0 - 1, SDORB_EDGE), vx1 = min(x0 + TW + 1, L.det_x1);
  const int vy0 = max(y0 - 1, SDORB_EDGE), vy1 = min(y0 + TH + 1, L.det_y1);
  for (int i = tid; i < SH * (PW / 4); i += NT) {
    const int r = i / (PW / 4) + 3, k = i % (PW / 4);  // staged row r, word k
    const int gy = y0 - 4 + r, gx = x0 - 4 + 4 * k;
    if (gy < vy0 || gy >= vy1 || gx + 3 < vx0 || gx >= vx1) continue;
    const uint32_t* rowc = reinterpret_cast<const uint32_t*>(&s_pix[r][0]);
    const uint32_t c = rowc[k];
    const uint32_t up = reinterpret_cast<const uint32_t*>(&s_pix[r - 3][0])[k];
    const uint32_t dn = reinterpret_cast<const uint32_t*>(&s_pix[r + 3][0])[k];
    const uint32_t wl = k > 0 ? rowc[k - 1] : 0u, wr = k + 1 < PW / 4 ? rowc[k + 1] : 0u;
    const uint32_t lf = __byte_perm(wl, c, 0x4321);  // pixels x-3 .. x
    const uint32_t rt = __byte_perm(c, wr, 0x6543);  // pixels x+3 .. x+6
    const uint32_t fv = gt_th(__vabsdiffu4(up, c), C, th_high) | gt_th(__vabsdiffu4(dn, c), C, th_high);
    const uint32_t fh = gt_th(__vabsdiffu4(lf, c), C, th_high) | gt_th(__vabsdiffu4(rt, c), C, th_high);
    uint32_t m = fv & fh & 0x80808080u;
    if (m == 0) continue;
    // drop lanes outside the valid column range
    int nb = 0;
    uint16_t ids[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int x = gx + b;
      if (((m >> (8 * b + 7)) & 1u) && x >= vx0 && x < vx1) ids[nb++] = (uint16_t)((r << 8) | (4 * k + b));
    }
    if (nb) {
      const int base = atomicAdd(&s_ncand, nb);
      for (int b = 0; b < nb; ++b) s_cand[base + b] = ids[b];
    }
  }
  __syncthreads();
  const int ncand = s_ncand;

  // ---- phase B: exact score per candidate
  for (int i = tid; i < ncand; i += NT) {
    const int id = s_cand[i];
    const int r = id >> 8, cx = id & 0xFF;
    const int m = corner_strength(&s_pix[r][cx], PW);
    if (m > th) s_score[r - 3][cx - 3] = (uint8_t)(m - 1);
  }
  __syncthreads();

  // ---- phase C: cell-bounded strict non-max suppression, append to the cell list
  for (int i = tid; i < ncand; i += NT) {
    const int id = s_cand[i];
    const int r = id >> 8, cx = id & 0xFF;
    const int sy = r - 3, sx = cx - 3;  // score-tile coordinates; (1,1) is pixel (x0,y0)
    if (sy < 1 || sy > TH || sx < 1 || sx > TW) continue;
    const int s = s_score[sy][sx];
    if (s == 0) continue;
    const int x = x0 - 1 + sx, y = y0 - 1 + sy;
    int cj = (x - SDORB_EDGE) / L.cell_w, ci = (y - SDORB_EDGE) / L.cell_h;
    cj = min(cj, L.cols - 1);
    ci = min(ci, L.rows - 1);
    const int cx0 = SDORB_EDGE + cj * L.cell_w, cy0 = SDORB_EDGE + ci * L.cell_h;
    const int cx1 = (cj == L.cols - 1) ? L.max_bx : cx0 + L.cell_w;
    const int cy1 = (ci == L.rows - 1) ? L.max_by : cy0 + L.cell_h;
    const bool l_ok = x - 1 >= cx0, r_ok = x + 1 < cx1, u_ok = y - 1 >= cy0, d_ok = y + 1 < cy1;
    bool keep = true;
    keep &= !l_ok || s > s_score[sy][sx - 1];
    keep &= !r_ok || s > s_score[sy][sx + 1];
    keep &= !u_ok || s > s_score[sy - 1][sx];
    keep &= !d_ok || s > s_score[sy + 1][sx];
    keep &= !(u_ok && l_ok) || s > s_score[sy - 1][sx - 1];
    keep &= !(u_ok && r_ok) || s > s_score[sy - 1][sx + 1];
    keep &= !(d_ok && l_ok) || s > s_score[sy + 1][sx - 1];
    keep &= !(d_ok && r_ok) || s > s_score[sy + 1][sx + 1];
    if (!keep) continue;
    const int cell = ci * L.cols + cj;
    int32_t* cnt = buf.cell_count + (int64_t)frame * geom->cells_total + L.cell_base + cell;
    const int slot = atomicAdd(cnt, 1);
    if (slot < L.list_cap_cell) {
      uint32_t* list = buf.cell_list + (int64_t)frame * geom->list_total + L.list_base + (int64_t)cell * L.list_cap_cell;
      list[slot] = SDORB_ENTRY(y, x, s);
    } else {
      atomicExch(buf.error_flag, 6);
    }
  }
}

void launch_fast_all(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, const SelectBuffers& b,
                     int nframes, cudaStream_t s) {
  if (g.tiles_total_fast == 0) return;
  fast_all_kernel<<<dim3(g.tiles_total_fast, nframes), NT, 0, s>>>(d_geom, p, b);
}

}  // namespace sdorb
