// sdorb_internal.h -- structures shared by the host runtime and the sm_100a kernels of libsdorb.
#pragma once
#include <stddef.h>
#include <stdint.h>

#define SDORB_MAX_LEVELS 32
#define SDORB_EDGE 19        // EDGE_THRESHOLD, /root/reference/src/ORBextractor.cc:75
#define SDORB_HALF_PATCH 15  // HALF_PATCH_SIZE, src/ORBextractor.cc:74
#define SDORB_MAX_DIM 4095   // keypoint entries pack y:12 | x:12 | score:8
#define SDORB_MAX_CELLS_PER_LEVEL 4096

// Tile shapes of the all-level launches.  A FAST tile scores 128 x 62 pixels starting at column 12 + 120*tx (a
// multiple of 4, so every staged row is word aligned) and row 18 + 60*ty, and writes the keypoint map for the inner
// 120 x 60 pixels (30 whole words per row) starting at (16 + 120*tx, 19 + 60*ty); the first detectable pixel is (19,19).
// (60 rows and four warps per tile instead of 30 and two: 3 % fewer halo rows at the same 32 warps per SM.)
// What is left of a level's width after the full tiles goes to one column of narrow tiles (16, 32 or 64 pixels scored).
#define SDORB_FAST_TW 120
#ifndef SDORB_FAST_TH
#define SDORB_FAST_TH 60
#endif
#define SDORB_BLUR_TW 128
#define SDORB_BLUR_TH 128

// Packed keypoint entry used between the FAST, selection and describe kernels.
// Ascending order of the packed word == row-major (y, then x) order, the order cv::FAST emits in.
#define SDORB_ENTRY(y, x, s) (((uint32_t)(y) << 20) | ((uint32_t)(x) << 8) | (uint32_t)(s))
#define SDORB_ENTRY_Y(e) ((int)((e) >> 20))
#define SDORB_ENTRY_X(e) ((int)(((e) >> 8) & 0xFFFu))
#define SDORB_ENTRY_SCORE(e) ((int)((e)&0xFFu))

// Geometry of one pyramid level for one input size; everything ComputePyramid / ComputeKeyPoints derive
// from the image size alone (src/ORBextractor.cc:469-488, 683).
struct LevelGeom {
  int w, h;              // level image size
  int pitch;             // bytes per row of the level plane in scratch (multiple of 128)
  int n_desired;         // mnFeaturesPerLevel[level]
  int cols, rows;        // levelCols, levelRows  (0 => the level produces nothing)
  int cell_w, cell_h;    // cellW, cellH
  uint32_t cell_w_magic, cell_h_magic;  // floor(2^32 / d) + 1: n / d == umulhi(n, magic) for n < 65536, d >= 2
  int n_features_cell;   // nfeaturesCell
  int max_bx, max_by;    // maxBorderX / maxBorderY  (= w-19, h-19)
  int last_x0, last_y0;  // first detectable x / y of the last cell column / row
  int last_skipped_x, last_skipped_y;  // 1 when the last column / row is skipped (hX<=0 / hY<=0)
  int det_x1, det_y1;    // exclusive end of the detectable area over all cells
  int cell_base;         // index of this level's first cell in the per-frame cell arrays
  int sel_base;          // index of this level's first slot in the per-frame selected-entry array
  int list_cap_cell;     // capacity (entries) of every cell list of this level
  int64_t list_base;     // offset (entries) of this level's first cell list in the per-frame list array
  int64_t plane_base;    // offset (bytes) of this level's plane array in the pyramid / blur scratch
  int64_t plane_bytes;   // pitch * h
  int tile_base_fast, tiles_x_fast, tiles_y_fast;  // flattened tile tables for the all-level launches (full FAST tiles)
  int tile_base_fastn, fastn_words;  // narrow FAST tiles of the last tile column: scored words per row (4, 8, 16; 0 = none)
  int tile_base_blur, tiles_x_blur, tiles_y_blur;  // blur strips: 128 columns x SDORB_BLUR_TH rows, one warp each (none for levels without cells)
  uint32_t blur_sel_last, blur_sel_beyond;  // PRMT selectors that rebuild, from the row's last two words, the last word with its
                                            // out-of-row bytes mirrored in (BORDER_REFLECT_101) and the word after it
  int scaled_patch_size; // (int)(31 * mvScaleFactor[level])
  float scale;           // mvScaleFactor[level]
  int coef_x_base, coef_y_base;  // offsets into the resize coefficient tables (entries), levels >= 1
  int group_base;        // first ResizeGroup of this level (levels >= 1), -1 when the 8-byte window does not fit (scale > ~2)
  int rz_span;           // largest number of source rows that 8 consecutive destination rows (groups starting at multiples of 8) touch;
                         // 0 when a row's taps are not (s, s+1) / (s, s) or two rows end on one source row -- selects the resize kernel
  int n_ini;             // ORB-SLAM2-style mode: initial quadtree nodes, round(width / height) of the bordered level
  float h_x;             // ... and their width hX
};

struct FrameGeom {
  int nlevels;
  int width, height;
  int nfeatures;
  int th_fast;           // threshold the FAST kernel runs with (ORB-SLAM2-style mode: min(iniThFAST, minThFAST))
  int octree;            // 1 = ORB-SLAM2-style mode: 30-pixel cells, ini / min threshold, DistributeOctTree
  int ini_th;            // iniThFAST of that mode
  int cells_total;      // per-frame cells over all levels
  int sel_total;        // per-frame selected-entry slots (sum of n_desired)
  int tiles_total_fast, tiles_total_fastn, tiles_total_blur;
  int64_t list_total;   // per-frame cell-list entries over all levels
  int64_t plane_total;  // per-frame... unused (planes are level-major); bytes of one frame over all levels
  LevelGeom lv[SDORB_MAX_LEVELS];
};

// One bilinear tap pair of cv::resize's fixed-point tables (imgproc resize.cpp): source indices and the
// two 11-bit coefficients, rounded independently.
struct ResizeTap {
  uint16_t s0, s1;
  int16_t c0, c1;
};
// Horizontal taps of four consecutive destination pixels, prepared for PRMT + IDP.2A: the eight source bytes starting
// at src_x hold every tap of the group; sel01 / sel23 gather (s0, s1) of pixels 0,1 / 2,3 into one word each and
// coef[i] = c0 | c1 << 16 of pixel i.
struct ResizeGroup {
  int32_t src_x;
  uint32_t sel01, sel23;
  uint32_t coef[4];
  int32_t pad_;
};
