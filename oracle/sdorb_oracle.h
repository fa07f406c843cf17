/*
 * sdorb_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference's ORB front-end
 * (SD-SLAM src/ORBextractor.cc, src/ORBmatcher.cc) and of the OpenCV 4.13 / glibc / libstdc++
 * primitives it calls.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product (libsdorb.so) never links, loads or calls it.
 *
 * Parity pin (round 2: PINNED to the reference itself).  The reference has no tests or golden vectors,
 * but its own ORBextractor.cc compiles unmodified against oracle/ref_compat (a cv:: surface whose pixel primitives are the
 * ones below) into oracle/_ref/libsdorb_ref.so (oracle/ref_build/Makefile).  tests/test_ref_parity.py and tests/tools/ref_sweep.py
 * assert  restatement == reference  byte for byte (keypoints incl. order, angles, descriptors, pyramid and its borders,
 * constructor tables, where it throws) on every fixture, the staged configurations, random sizes / parameters and 6560 bench
 * frames; the fixtures under tests/golden/ are generated from the reference (tests/golden/make_golden.py).  The OpenCV / glibc /
 * libstdc++ primitives are pinned against the only importable implementation of that arithmetic, Python cv2 4.13.0
 * (tests/test_oracle_primitives.py), and the whole operator() also against an independent Python assembly of cv2 calls
 * (tests/test_oracle_e2e.py).
 */
#ifndef SDORB_ORACLE_H
#define SDORB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Binary-compatible with cv::KeyPoint (28 bytes). */
typedef struct {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} orc_keypoint;

typedef struct {
  int nfeatures;
  float scale_factor;
  int nlevels;
  int th_fast;
} orc_params;

/* ---- OpenCV / libc primitive restatements (each validated against cv2 4.13) ---- */
void orc_resize_linear_8u(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh,
                          size_t dstep);
void orc_copy_make_border_reflect101(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep,
                                     int border);
/* cv::FAST(img, kps, th, nonmax=true); returns count (may exceed cap; only cap are written). */
int orc_fast(const uint8_t* img, int w, int h, size_t step, int th, int nonmax, orc_keypoint* out, int cap);
/* cornerScore<16>() at every pixel that is a FAST-9 corner for `th`, else 0 (border 3 px = 0). */
void orc_fast_score_map(const uint8_t* img, int w, int h, size_t step, int th, uint8_t* score, size_t score_step);
void orc_gaussian_blur_7x7_s2(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep);
float orc_fast_atan2(float y, float x);
/* KeyPointsFilter::retainBest(v, n) using the real std::nth_element / std::partition; returns new size. */
int orc_retain_best(orc_keypoint* kps, int count, int n);
/* Same selection on bare float responses carrying an index payload (for the introselect twin tests). */
int orc_retain_best_idx(const float* response, int count, int n, int32_t* order_out);
/* glibc sincosf restated in double arithmetic (what the CUDA library evaluates on device). */
void orc_sincosf_restated(float x, float* s, float* c);
/* Returns number of floats in [lo_bits, hi_bits] (as IEEE bit patterns) where restated != libm sincosf. */
uint64_t orc_sincosf_mismatches(uint32_t lo_bits, uint32_t hi_bits, int nthreads);
/* rBRIEF sample offset: (cvRound(x*b + y*a), cvRound(x*a - y*b)) with the frozen FMA contraction. */
void orc_pattern_rotate(int px, int py, float a, float b, int* drow, int* dcol);

/* ---- ORBextractor restatement ---- */
typedef struct orc_extractor orc_extractor;
orc_extractor* orc_create(const orc_params* p);
void orc_destroy(orc_extractor* e);
/* Row f1: switches the extractor to the ORB-SLAM2-style mode -- per 30-pixel cell cv::FAST with
 * iniThFAST, minThFAST where that finds nothing, then DistributeOctTree per level.  Not in /root/reference: restated from the
 * public ORB-SLAM2 algorithm, PARITY UNPINNED against ORB-SLAM2 itself (see the comment in sdorb_oracle.cc).  min_th_fast < 0
 * switches back to the reference's ComputeKeyPoints. */
void orc_set_orbslam2_mode(orc_extractor* e, int ini_th_fast, int min_th_fast);
/* ORBextractor::DistributeOctTree of ORB-SLAM2 on keypoints whose coordinates are relative to (minX, minY); returns the
 * number of result keypoints (at most cap written). */
int orc_distribute_oct_tree(const orc_keypoint* keys, int n, int minX, int maxX, int minY, int maxY, int N, orc_keypoint* out,
                            int cap);
/* tables: each array has nlevels entries; umax has 16. Any pointer may be NULL. */
void orc_get_tables(const orc_extractor* e, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                    int* n_per_level, int* umax);
/* level geometry (ComputePyramid sizes + ComputeKeyPoints grid) for an input of w x h */
typedef struct {
  int width, height;       /* level image size */
  int n_desired;           /* mnFeaturesPerLevel[level] */
  int level_cols, level_rows;
  int cell_w, cell_h;
  int n_features_cell;
  int scaled_patch_size;
} orc_level_geom;
void orc_level_geometry(const orc_extractor* e, int w, int h, orc_level_geom* out /* nlevels */);

/* Optional stage dump of one extraction.  All pointers may be NULL (then that stage is not dumped).
 * pyramid / blurred: tightly packed levels (width*height bytes each, level after level).
 * raw_*: FAST keypoints per cell before any selection, in emission order (cell-local coordinates),
 *        flattened level-major then cell row-major; raw_cell_count has one entry per cell. */
typedef struct {
  uint8_t* pyramid;
  uint8_t* blurred;
  int32_t* level_count;     /* nlevels: selected keypoints per level */
  int32_t* raw_cell_count;  /* sum over levels of level_rows*level_cols entries */
  int32_t raw_cell_cap;     /* capacity of raw_cell_count */
  orc_keypoint* raw_kps;
  int32_t raw_kps_cap;
  int32_t raw_kps_total;    /* out */
  int32_t n_to_retain_cap;
  int32_t* n_to_retain;     /* same indexing as raw_cell_count */
} orc_dump;

/* ORBextractor::operator(): returns number of keypoints (<= cap written), or <0 on geometry error
 * (the reference would throw cv::Exception for an out-of-range cell ROI).  desc: n x 32 bytes. */
int orc_extract(const orc_extractor* e, const uint8_t* img, int w, int h, size_t step, orc_keypoint* kps,
                uint8_t* desc, int cap, orc_dump* dump);
/* Frame-parallel driver for the CPU baseline: frames are w*h tightly packed; returns total keypoints. */
long orc_extract_many(const orc_extractor* e, const uint8_t* imgs, int nframes, int w, int h, int nthreads,
                      orc_keypoint* kps /* nframes*cap or NULL */, uint8_t* desc /* or NULL */,
                      int32_t* counts /* or NULL */, int cap);

/* ---- ORBmatcher restatement ---- */
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b);
typedef struct {
  int32_t best_idx;   /* -1 if no candidate */
  int32_t best_dist;  /* 256 if none */
  int32_t second_dist;
  int32_t accepted;   /* best_dist < th_low && (float)best_dist < ratio*(float)second_dist */
} orc_match;
/* For each of nA queries scan all nB train rows in ascending order with the reference update rule. */
void orc_match_best2(const uint8_t* descA, int nA, const uint8_t* descB, int nB, float ratio, int th_low,
                     orc_match* out);
/* SearchByPoints-style greedy variant: train rows already matched are skipped for later queries. */
void orc_match_greedy(const uint8_t* descA, int nA, const uint8_t* descB, int nB, float ratio, int th_low,
                      orc_match* out);
void orc_match_many(const uint8_t* descA, const int32_t* nA, const uint8_t* descB, const int32_t* nB, int npairs,
                    int strideA_rows, int strideB_rows, float ratio, int th_low, int nthreads, orc_match* out);
void orc_hamming_matrix(const uint8_t* descA, int nA, const uint8_t* descB, int nB, uint16_t* out);
/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:252-275): index of the descriptor with the least median
 * distance to the others (first on ties), -1 for an empty set; the median itself through *best_median. */
int orc_distinctive(const uint8_t* desc, int n, int* best_median);
/* sets are rows [offsets[s], offsets[s+1]) of desc */
void orc_distinctive_many(const uint8_t* desc, const int32_t* offsets, int nsets, int nthreads, int32_t* best_idx,
                          int32_t* best_median);

/* ---- Frame post-processing restatement (src/Frame.cc) ---- */
/* AssignFeaturesToGrid + PosInGrid (:179-192, :323-332): cell_start has 64*48+1 entries, indices n entries. */
void orc_assign_grid(const orc_keypoint* kps_un, int n, float mnMinX, float mnMinY, float mfGridElementWidthInv,
                     float mfGridElementHeightInv, int32_t* cell_start, int32_t* indices);
/* UndistortKeyPoints (:335-366) = cv::undistortPoints(pts, K, dist, noArray(), K) restated (OpenCV 4.13, double, 5
 * iterations); K = {fx, fy, cx, cy}, dist = {k1, k2, p1, p2[, k3]}; dist[0] == 0 copies the keypoints (:336-339). */
void orc_undistort_keypoints(const orc_keypoint* kps, int n, const float K[4], const float* dist, int ndist, orc_keypoint* out);
/* ComputeImageBounds (:368-397): bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY}. */
void orc_image_bounds(int cols, int rows, const float K[4], const float* dist, int ndist, float bounds[4]);
/* ComputeStereoFromRGBD (:399-417): depth is a tightly packed float image of the given width. */
void orc_stereo_from_rgbd(const orc_keypoint* kps, const orc_keypoint* kps_un, int n, const float* depth, int width, float mbf,
                          float* u_right, float* z);


/* ---- guided matchers (src/ORBmatcher.cc) on Frame::GetFeaturesInArea (src/Frame.cc:271-321) ---- */
/* the Frame grid as orc_assign_grid leaves it, with the four Frame members the window query reads */
typedef struct {
  const int32_t* cell_start; /* 64*48+1 */
  const int32_t* indices;
  float mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv;
} orc_frame_grid;
/* Frame::GetFeaturesInArea (:271-321): indices of the keypoints inside the window, in the reference's order (cell
 * columns outer, rows inner, push order inside a cell); returns the count (at most n written to out). */
int orc_features_in_area(const orc_keypoint* kps_un, const orc_frame_grid* grid, float x, float y, float r, int minLevel,
                         int maxLevel, int32_t* out);
/* ORBmatcher::ComputeThreeMaxima (:1423-1454) on the bin sizes. */
void orc_three_maxima(const int32_t* sizes, int L, int* ind1, int* ind2, int* ind3);
/* ORBmatcher::SearchForInitialization (:256-357).  prev_matched: n1 (x, y) pairs, updated in place (:349-352);
 * matches12: n1 entries; returns nmatches. */
int orc_search_for_initialization(const orc_keypoint* kps1_un, const uint8_t* desc1, int n1, const orc_keypoint* kps2_un,
                                  const uint8_t* desc2, int n2, const orc_frame_grid* grid2, float* prev_matched,
                                  int window_size, float nnratio, int check_orientation, int32_t* matches12);
/* ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (:946-1075) from the projection on: the caller supplies
 * per keypoint i of the last frame the float triple (u, v, invzc) of :979-987 and flags (bit 0: the keypoint has a map
 * point that is not an outlier, :968-971; bit 1: that map point has Observations() > 0), its map point's descriptor, and
 * for the current frame mvuRight and `occupied` (mvpMapPoints[i2] && Observations() > 0 on entry, :1018-1020; the
 * restatement tracks it as map points are assigned).  mode: 0 neither, 1 bForward, 2 bBackward (:962-963).  bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY}.
 * assigned[i2] = last-frame index whose map point ends up in mvpMapPoints[i2] through this call, else -1. */
int orc_search_by_projection(const orc_keypoint* kps_last, const orc_keypoint* kps_last_un, const float* proj /* n x 3 */,
                             const uint8_t* flags_last, const uint8_t* desc_mp, int n_last, const orc_keypoint* kps_cur_un,
                             const uint8_t* desc_cur, const float* u_right_cur, const uint8_t* occupied_cur, int n_cur,
                             const orc_frame_grid* grid_cur, const float* scale_factors, const float bounds[4], float th,
                             float mbf, int mode, int check_orientation, int32_t* assigned, int orb_dist /* 0 = TH_HIGH */);

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:43-119): see sdorb_oracle.cc for the
 * argument conventions; returns nmatches, assigned has n_frame entries. */
int orc_search_map_points(const float* proj, const float* view_cos, const int32_t* level, const uint8_t* flags, const uint8_t* desc_mp,
                          int n_mp, const orc_keypoint* kps_un, const uint8_t* desc, const float* u_right, const uint8_t* occupied,
                          int n_frame, const orc_frame_grid* grid, const float* scale_factors, float th, float nnratio,
                          int32_t* assigned);
/* ORBmatcher::SearchByPoints (src/ORBmatcher.cc:1209-1304): valid = the keypoint has a map point that is not bad; matches12 has n1
 * entries (index into the second keyframe or -1); returns nmatches. */
int orc_search_by_points(const orc_keypoint* kps1_un, const uint8_t* desc1, const uint8_t* valid1, int n1, const orc_keypoint* kps2_un,
                         const uint8_t* desc2, const uint8_t* valid2, int n2, float nnratio, int check_orientation, int32_t* matches12);
/* The keypoint search of ORBmatcher::Fuse (src/ORBmatcher.cc:535-586; without the reprojection gate: the Sim3 overload :682-708 and
 * either direction of SearchBySim3); see sdorb_oracle.cc. */
void orc_fuse_search(const float* proj, const int32_t* level, const uint8_t* flags, const uint8_t* desc_mp, int n_mp,
                     const orc_keypoint* kps_un, const uint8_t* desc, const float* u_right, const orc_frame_grid* grid,
                     const float* scale_factors, const float* inv_level_sigma2, float th, int check_reprojection, int th_dist,
                     int32_t* best_idx, int32_t* best_dist);
/* ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:734-944) from the projections on; returns nFound. */
int orc_search_by_sim3(const float* proj1, const int32_t* level1, const uint8_t* flags1, const uint8_t* desc_mp1, int n1,
                       const float* proj2, const int32_t* level2, const uint8_t* flags2, const uint8_t* desc_mp2, int n2,
                       const orc_keypoint* kps1_un, const uint8_t* desc1, const orc_frame_grid* grid1, const orc_keypoint* kps2_un,
                       const uint8_t* desc2, const orc_frame_grid* grid2, const float* scale_factors1, const float* scale_factors2,
                       float th, int32_t* match1, int32_t* match2, int32_t* matches12);
/* ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (src/ORBmatcher.cc:146-254) from the projection on. */
int orc_search_by_projection_sim3(const float* proj, const int32_t* level, const uint8_t* flags, const uint8_t* desc_mp, int n_mp,
                                  const orc_keypoint* kps_un, const uint8_t* desc, const uint8_t* matched_in, int n_kf,
                                  const orc_frame_grid* grid, const float* scale_factors, int th, int32_t* assigned);
/* ORBmatcher::CheckDistEpipolarLine (src/ORBmatcher.cc:128-144); F12 row-major (F12(i, j) = F12[3 * i + j]), sigma2 =
 * pKF2->mvLevelSigma2[kp2.octave]. */
int orc_check_dist_epipolar_line(float x1, float y1, float x2, float y2, const double* F12, float sigma2);
/* ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:359-462) from the epipole (ex, ey) of :366-368 on.  has_mp: the keypoint
 * already has a map point; u_right: mvuRight.  matches12 receives vMatches12 (n1 entries, -1 = none; vMatchedPairs is the
 * list of (i, matches12[i]) with matches12[i] >= 0); returns nmatches. */
int orc_search_for_triangulation(const orc_keypoint* kps1_un, const uint8_t* desc1, const uint8_t* has_mp1, const float* u_right1,
                                 int n1, const orc_keypoint* kps2_un, const uint8_t* desc2, const uint8_t* has_mp2,
                                 const float* u_right2, int n2, const double* F12, float ex, float ey, const float* scale_factors,
                                 const float* level_sigma2, int check_orientation, int32_t* matches12);

#ifdef __cplusplus
}
#endif
#endif
