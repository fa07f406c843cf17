// geometry.cc -- host-side tables of the extractor: everything that depends only on the constructor
// arguments and the image size.  Mirrors the ORBextractor constructor (/root/reference/src/ORBextractor.cc:406-457),
// the level sizes of ComputePyramid (:683), the cell grid of ComputeKeyPoints (:469-532) and the coefficient
// tables cv::resize builds for INTER_LINEAR on 8-bit data (OpenCV imgproc/resize.cpp, fixed point, 11 bits).
#include "geometry.h"

#include <algorithm>
#include <cmath>

namespace sdorb {

namespace {
// cvRound: cvtss2si / cvtsd2si, i.e. round-half-to-even in the default rounding mode.
inline int round_even(float v) { return (int)lrintf(v); }
inline int round_even(double v) { return (int)lrint(v); }
inline int floor_int(double v) {
  const int i = (int)v;
  return i - (i > v);
}
inline int ceil_int(double v) {
  const int i = (int)v;
  return i + (i < v);
}
inline int align_up(int v, int a) { return (v + a - 1) / a * a; }
inline int16_t coef11(float f) {
  const int i = round_even(f * 2048.f);
  return (int16_t)std::min(32767, std::max(-32768, i));
}
}  // namespace

void build_tables(int nfeatures, float scale_factor_f, int nlevels, Tables* t) {
  // `scaleFactor` is a double member initialised from the float argument (src/ORBextractor.h:78)
  const double scale_factor = scale_factor_f;
  t->nlevels = nlevels;
  t->scale.assign(nlevels, 1.f);
  t->sigma2.assign(nlevels, 1.f);
  for (int i = 1; i < nlevels; ++i) {
    t->scale[i] = (float)(t->scale[i - 1] * scale_factor);
    t->sigma2[i] = t->scale[i] * t->scale[i];
  }
  t->inv_scale.resize(nlevels);
  t->inv_sigma2.resize(nlevels);
  for (int i = 0; i < nlevels; ++i) {
    t->inv_scale[i] = 1.0f / t->scale[i];
    t->inv_sigma2[i] = 1.0f / t->sigma2[i];
  }
  // geometric split of nfeatures over the levels; the remainder goes to the last level (:424-434)
  t->n_per_level.resize(nlevels);
  const float factor = (float)(1.0f / scale_factor);
  float per_scale = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
  int sum = 0;
  for (int l = 0; l + 1 < nlevels; ++l) {
    t->n_per_level[l] = round_even(per_scale);
    sum += t->n_per_level[l];
    per_scale *= factor;
  }
  t->n_per_level[nlevels - 1] = std::max(nfeatures - sum, 0);

  // circular patch row ends for the intensity centroid (:442-456)
  const int R = SDORB_HALF_PATCH;
  int v, v0;
  const int vmax = floor_int(R * std::sqrt(2.f) / 2 + 1);
  const int vmin = ceil_int(R * std::sqrt(2.f) / 2);
  const double r2 = (double)R * R;
  for (v = 0; v <= vmax; ++v) t->umax[v] = round_even(std::sqrt(r2 - v * v));
  for (v = R, v0 = 0; v >= vmin; --v) {
    while (t->umax[v0] == t->umax[v0 + 1]) ++v0;
    t->umax[v] = v0;
    ++v0;
  }
}

void level_size(const Tables& t, int level, int width, int height, int* lw, int* lh) {
  const float s = t.inv_scale[level];
  *lw = round_even((float)width * s);
  *lh = round_even((float)height * s);
}

// One axis of cv::resize's INTER_LINEAR tables.  `clamp_frac` distinguishes the horizontal axis (source
// index clamped AND fraction zeroed at both ends) from the vertical one (fraction kept, rows clipped).
static void resize_axis(int src, int dst, bool horizontal, ResizeTap* out) {
  const double scale = 1. / ((double)dst / src);
  for (int d = 0; d < dst; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = floor_int(f);
    f -= s;
    ResizeTap tap;
    if (horizontal) {
      if (s < 0) {
        f = 0;
        s = 0;
      }
      if (s >= src - 1) {
        f = 0;
        s = src - 1;
      }
      tap.s0 = (uint16_t)s;
      tap.s1 = (uint16_t)std::min(s + 1, src - 1);
    } else {
      auto clip = [src](int x) { return x >= 0 ? (x < src ? x : src - 1) : 0; };
      tap.s0 = (uint16_t)clip(s);
      tap.s1 = (uint16_t)clip(s + 1);
    }
    tap.c0 = coef11(1.f - f);
    tap.c1 = coef11(f);
    out[d] = tap;
  }
}

int octree_level_slots(int n_desired, int n_ini) { return std::max(std::max(n_desired, 0) + 3, 4 * n_ini); }

int build_frame_geom(const Tables& t, int nfeatures, int th_fast, int width, int height, FrameGeom* g,
                     std::vector<ResizeTap>* taps, std::vector<ResizeGroup>* groups, int min_th_fast) {
  if (width <= 0 || height <= 0 || width > SDORB_MAX_DIM || height > SDORB_MAX_DIM) return -1;
  *g = FrameGeom{};
  g->nlevels = t.nlevels;
  g->width = width;
  g->height = height;
  g->nfeatures = nfeatures;
  g->th_fast = std::min(std::max(th_fast, 0), 255);  // cv::FAST clamps the threshold
  if (min_th_fast >= 0) {
    // ORB-SLAM2-style mode: FAST runs once with the smaller threshold -- a keypoint for
    // iniThFAST is exactly a keypoint for minThFAST whose score reaches iniThFAST -- and the cells choose afterwards
    g->octree = 1;
    g->ini_th = g->th_fast;
    g->th_fast = std::min(g->th_fast, std::min(min_th_fast, 255));
  }
  taps->clear();
  if (groups) groups->clear();
  const float ratio = (float)width / height;
  int cell_base = 0, sel_base = 0, tile_fast = 0, tile_fastn = 0, tile_blur = 0;
  int64_t list_base = 0, plane_base = 0;
  for (int l = 0; l < t.nlevels; ++l) {
    LevelGeom& L = g->lv[l];
    level_size(t, l, width, height, &L.w, &L.h);
    if (L.w <= 0 || L.h <= 0) return -4;  // cv::resize asserts on an empty destination
    L.pitch = align_up(L.w, 128);
    L.plane_bytes = (int64_t)L.pitch * L.h;
    L.plane_base = plane_base;
    plane_base += L.plane_bytes;
    L.n_desired = t.n_per_level[l];
    L.scale = t.scale[l];
    L.scaled_patch_size = (int)(31 * t.scale[l]);
    L.max_bx = L.w - SDORB_EDGE;
    L.max_by = L.h - SDORB_EDGE;
    L.cols = g->octree ? 0 : (int)std::sqrt((float)L.n_desired / (5 * ratio));
    L.rows = g->octree ? 0 : (int)(ratio * L.cols);
    L.cell_base = cell_base;
    L.sel_base = sel_base;
    L.list_base = list_base;
    sel_base += std::max(L.n_desired, 0);
    bool any_detect = false;
    if (g->octree) {
      // ComputeKeyPointsOctTree of ORB-SLAM2: cells of about 30 pixels over [16, w-16) x [16, h-16); cell (i, j) is the ROI
      // [16 + j*wCell, +wCell+6) clipped at w-16, skipped from iniX >= maxBorderX-6 / iniY >= maxBorderY-3 on; its
      // detectable rectangle starts at 19 + j*wCell like the reference's cells, so the same kernels serve both grids.
      const int min_b = SDORB_EDGE - 3, bx1 = L.w - SDORB_EDGE + 3, by1 = L.h - SDORB_EDGE + 3;
      const float wf = (float)(bx1 - min_b), hf = (float)(by1 - min_b);
      const int ncols = (int)(wf / 30.f), nrows = (int)(hf / 30.f);
      if (ncols <= 0 || nrows <= 0) return -4;  // ORB-SLAM2 divides by zero here
      L.cell_w = (int)std::ceil(wf / ncols);
      L.cell_h = (int)std::ceil(hf / nrows);
      L.n_ini = (int)std::round((float)(bx1 - min_b) / (by1 - min_b));
      if (L.n_ini < 1) return -4;  // ... and here (portrait levels)
      if (L.n_ini > 8) return -7;
      L.h_x = (float)(bx1 - min_b) / L.n_ini;
      // rows / columns whose ROI is at least 7 pixels, i.e. that can hold a keypoint: iniX <= maxBorderX - 7
      L.cols = L.rows = 0;
      for (int j = 0; j < ncols; ++j)
        if (min_b + j * L.cell_w <= bx1 - 7) L.cols = j + 1;
      for (int i = 0; i < nrows; ++i)
        if (min_b + i * L.cell_h <= by1 - 7) L.rows = i + 1;
      L.cell_w_magic = L.cell_w >= 2 ? (uint32_t)(0x100000000ull / (uint64_t)L.cell_w) + 1u : 0u;
      L.cell_h_magic = L.cell_h >= 2 ? (uint32_t)(0x100000000ull / (uint64_t)L.cell_h) + 1u : 0u;
      L.n_features_cell = 0;
      sel_base += octree_level_slots(L.n_desired, L.n_ini) - std::max(L.n_desired, 0);  // DistributeOctTree may return more than N
      if (L.cols > 0 && L.rows > 0) {
        if ((int64_t)L.cols * L.rows > SDORB_MAX_CELLS_PER_LEVEL) return -7;
        any_detect = true;
        L.det_x1 = L.max_bx;
        L.det_y1 = L.max_by;
        L.last_x0 = SDORB_EDGE + (L.cols - 1) * L.cell_w;
        L.last_y0 = SDORB_EDGE + (L.rows - 1) * L.cell_h;
        const int cw = std::min(L.cell_w, L.max_bx - SDORB_EDGE), ch = std::min(L.cell_h, L.max_by - SDORB_EDGE);
        L.list_cap_cell = ((cw + 1) / 2) * ((ch + 1) / 2);
        cell_base += L.rows * L.cols;
        list_base += (int64_t)L.list_cap_cell * L.rows * L.cols;
      } else {
        L.cols = L.rows = 0;
      }
    } else if (L.cols > 0 && L.rows > 0) {
      if ((int64_t)L.cols * L.rows > SDORB_MAX_CELLS_PER_LEVEL) return -7;
      const int W = L.max_bx - SDORB_EDGE, H = L.max_by - SDORB_EDGE;
      L.cell_w = (int)std::ceil((float)W / L.cols);
      L.cell_h = (int)std::ceil((float)H / L.rows);
      L.n_features_cell = (int)std::ceil((float)L.n_desired / (L.rows * L.cols));
      L.cell_w_magic = L.cell_w >= 2 ? (uint32_t)(0x100000000ull / (uint64_t)L.cell_w) + 1u : 0u;
      L.cell_h_magic = L.cell_h >= 2 ? (uint32_t)(0x100000000ull / (uint64_t)L.cell_h) + 1u : 0u;
      // Walk the cell ROIs exactly like the reference loop (:495-532) to validate them and to find the
      // detectable rectangle of the last row / column.
      int cap = 0;
      L.det_x1 = L.det_y1 = 0;
      for (int i = 0; i < L.rows; ++i) {
        const int y0 = SDORB_EDGE + i * L.cell_h - 3;
        int hy = L.cell_h + 6;
        if (i == L.rows - 1) {
          hy = L.max_by + 3 - y0;
          L.last_y0 = y0 + 3;
          if (hy <= 0) {
            L.last_skipped_y = 1;
            continue;
          }
        }
        for (int j = 0; j < L.cols; ++j) {
          const int x0 = SDORB_EDGE + j * L.cell_w - 3;
          int hx = L.cell_w + 6;
          if (j == L.cols - 1) {
            hx = L.max_bx + 3 - x0;
            L.last_x0 = x0 + 3;
            if (hx <= 0) {
              L.last_skipped_x = 1;
              continue;
            }
          }
          // Mat::rowRange / colRange would throw
          if (y0 < 0 || hy < 0 || y0 + hy > L.h || x0 < 0 || hx < 0 || x0 + hx > L.w) return -4;
          if (hx >= 7 && hy >= 7) {
            // detectable rectangle [x0+3, x0+hx-3) x [y0+3, y0+hy-3); one past maxBorder means the
            // reference reads its descriptor pattern outside the blurred image (undefined there).
            if (x0 + hx - 3 > L.max_bx || y0 + hy - 3 > L.max_by) return -4;
            any_detect = true;
            L.det_x1 = std::max(L.det_x1, x0 + hx - 3);
            L.det_y1 = std::max(L.det_y1, y0 + hy - 3);
            // strict 3x3 non-max suppression leaves at most one keypoint per 2x2 block
            cap = std::max(cap, ((hx - 6 + 1) / 2) * ((hy - 6 + 1) / 2));
          }
        }
      }
      if (any_detect && (L.cell_w <= 0 || L.cell_h <= 0)) return -4;  // cannot happen; keeps the kernels' divisions safe
      L.list_cap_cell = any_detect ? cap : 0;
      cell_base += L.rows * L.cols;
      list_base += (int64_t)L.list_cap_cell * L.rows * L.cols;
    }
    if (!any_detect) L.det_x1 = L.det_y1 = 0;
    // flattened tile tables
    L.tile_base_fast = tile_fast;
    L.tile_base_fastn = tile_fastn;
    if (any_detect) {
      // full tiles emit 120 columns each from column 16 on; the remainder goes to one column of narrow tiles
      const int span = L.det_x1 - 16, rest = span % SDORB_FAST_TW;
      L.tiles_x_fast = span / SDORB_FAST_TW;
      L.tiles_y_fast = (L.det_y1 - SDORB_EDGE + SDORB_FAST_TH - 1) / SDORB_FAST_TH;
      if (rest > 0) {
        L.fastn_words = 32;
        for (int nw = 4; nw < 32; nw *= 2)
          if ((nw - 2) * 4 >= rest) {
            L.fastn_words = nw;
            break;
          }
        if (L.fastn_words == 32) {  // too wide for a narrow tile: one more full tile
          L.fastn_words = 0;
          ++L.tiles_x_fast;
        }
      }
    }
    tile_fast += L.tiles_x_fast * L.tiles_y_fast;
    tile_fastn += L.fastn_words ? L.tiles_y_fast : 0;
    L.tile_base_blur = tile_blur;
    // the reference blurs only levels that hold keypoints (src/ORBextractor.cc:651-660); a level without cells gets no strips
    L.tiles_x_blur = any_detect ? (L.w + SDORB_BLUR_TW - 1) / SDORB_BLUR_TW : 0;
    L.tiles_y_blur = any_detect ? (L.h + SDORB_BLUR_TH - 1) / SDORB_BLUR_TH : 0;
    tile_blur += L.tiles_x_blur * L.tiles_y_blur;
    {
      // window = the eight bytes at columns lastw-4 .. lastw+3; column x >= w mirrors to 2(w-1) - x
      const int lastw = (L.w - 1) & ~3;
      L.blur_sel_last = L.blur_sel_beyond = 0;
      for (int b = 0; b < 4; ++b) {
        const int x = lastw + b, xb = lastw + 4 + b;
        const int i0 = (x < L.w ? x : 2 * (L.w - 1) - x) - (lastw - 4);
        const int i1 = 2 * (L.w - 1) - xb - (lastw - 4);
        L.blur_sel_last |= (uint32_t)std::min(std::max(i0, 0), 7) << (4 * b);
        L.blur_sel_beyond |= (uint32_t)std::min(std::max(i1, 0), 7) << (4 * b);
      }
    }
    // resize taps from level l-1 to l
    if (l > 0) {
      const LevelGeom& P = g->lv[l - 1];
      L.coef_x_base = (int)taps->size();
      taps->resize(taps->size() + L.w);
      resize_axis(P.w, L.w, true, taps->data() + L.coef_x_base);
      L.coef_y_base = (int)taps->size();
      taps->resize(taps->size() + L.h);
      resize_axis(P.h, L.h, false, taps->data() + L.coef_y_base);
      {
        // source rows touched by every group of 8 destination rows: one contiguous, ascending range for the prefetching kernel
        const ResizeTap* tyv = taps->data() + L.coef_y_base;
        int span = 0;
        bool ok = true;
        for (int y = 0; y < L.h && ok; ++y) {
          const ResizeTap& tp = tyv[y];
          if (tp.s1 != tp.s0 && tp.s1 != tp.s0 + 1) ok = false;
          if (y > 0 && (tp.s1 <= tyv[y - 1].s1 || tp.s0 < tyv[y - 1].s0)) ok = false;  // a source row ends at most one destination row
          if (tp.s1 == tp.s0 && y % 8 != 0) ok = false;  // both taps on one row: only as the first row of a group of eight
        }
        for (int y = 0; y < L.h && ok; y += 8) {
          const int yl = std::min(y + 8, L.h) - 1;
          span = std::max(span, (int)tyv[yl].s1 - (int)tyv[y].s0 + 1);
        }
        L.rz_span = ok ? span : 0;
      }
      // four destination pixels per group, all taps inside one 8-byte source window
      L.group_base = -1;
      if (groups) {
        const ResizeTap* tx = taps->data() + L.coef_x_base;
        const int ngroups = (L.w + 3) / 4;
        std::vector<ResizeGroup> gs(ngroups);
        bool fits = true;
        for (int y = 0; y < L.h; ++y) {  // the kernel multiplies the vertical coefficients as unsigned 16-bit values
          const ResizeTap& tp = taps->data()[L.coef_y_base + y];
          if (tp.c0 < 0 || tp.c1 < 0) fits = false;
        }
        for (int gi = 0; gi < ngroups && fits; ++gi) {
          ResizeGroup& G = gs[gi];
          G.pad_ = 0;
          G.src_x = tx[4 * gi].s0;
          uint32_t sel[2] = {0, 0};
          for (int i = 0; i < 4; ++i) {
            const ResizeTap& tp = tx[std::min(4 * gi + i, L.w - 1)];
            const int o0 = (int)tp.s0 - G.src_x, o1 = (int)tp.s1 - G.src_x;
            if (o0 < 0 || o1 < 0 || o0 > 7 || o1 > 7 || tp.c0 < 0 || tp.c1 < 0) fits = false;
            sel[i >> 1] |= ((uint32_t)(o0 & 7) | ((uint32_t)(o1 & 7) << 4)) << (8 * (i & 1));
            G.coef[i] = (uint32_t)(uint16_t)tp.c0 | ((uint32_t)(uint16_t)tp.c1 << 16);
          }
          G.sel01 = sel[0];
          G.sel23 = sel[1];
        }
        if (fits) {
          L.group_base = (int)groups->size();
          groups->insert(groups->end(), gs.begin(), gs.end());
        }
      }
    }
  }
  g->cells_total = cell_base;
  g->sel_total = sel_base;
  g->list_total = list_base;
  g->plane_total = plane_base;
  g->tiles_total_fast = tile_fast;
  g->tiles_total_fastn = tile_fastn;
  g->tiles_total_blur = tile_blur;
  return 0;
}

}  // namespace sdorb
