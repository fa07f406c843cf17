/*
 * sdorb.h -- C ABI of libsdorb.so: the B200-native (sm_100a) ORB front-end of SD-SLAM.
 *
 * This is the drop-in boundary for ONE hot path of pasensio97/SDslam (SD-SLAM):
 *   SD_SLAM::ORBextractor::operator()          /root/reference/src/ORBextractor.cc:620-678
 *   SD_SLAM::ORBmatcher::DescriptorDistance    /root/reference/src/ORBmatcher.cc:1459-1473
 *     batched in the shape of its hottest caller, the best / second-best scan of
 *     ORBmatcher::SearchByPoints               /root/reference/src/ORBmatcher.cc:1239-1265
 * The reference has no FFI layer; its boundary is the C++ class SD_SLAM::ORBextractor
 * (src/ORBextractor.h:34-90).  include/ORBextractor.h re-creates that class surface on top of the
 * functions below, so Frame / Tracking / Initializer keep compiling unchanged.
 *
 * Conventions: plain pointers and sizes only; every function returns SDORB_OK (0) or a negative
 * SDORB_ERR_* code and never throws.  There is no CPU fallback: if CUDA fails the call fails.
 * A handle owns one CUDA stream and all scratch memory; it is not re-entrant (use one handle per
 * thread / per GPU).  Results are bit-identical to the reference on the same input bytes: keypoint
 * x, y, size, response, octave and order; angle; descriptors; Hamming distances and match indices.
 */
#ifndef SDORB_H
#define SDORB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SDORB_API __attribute__((visibility("default")))
#else
#define SDORB_API
#endif

#define SDORB_OK 0
#define SDORB_ERR_BAD_ARG (-1)   /* null pointer, non-positive size, image larger than max_width/height */
#define SDORB_ERR_CAPACITY (-2)  /* caller's output capacity is smaller than nfeatures */
#define SDORB_ERR_CUDA (-3)      /* a CUDA call failed; see sdorb_last_cuda_error() */
#define SDORB_ERR_GEOMETRY (-4)  /* a cell ROI leaves its level image: the reference throws cv::Exception
                                    (or reads out of bounds) for this image size / parameter set */
#define SDORB_ERR_NOMEM (-5)     /* device or pinned-host allocation failed */
#define SDORB_ERR_OVERFLOW (-6)  /* an internal fixed-capacity list overflowed (results invalid) */
#define SDORB_ERR_UNSUPPORTED (-7)

/* largest batch (frames / frame pairs per call) of the entry points that say so below; larger requests return
 * SDORB_ERR_BAD_ARG -- split them into several calls */
#define SDORB_MAX_GRID_BATCH 65535

/* where the image / result buffers of a batch call live */
#define SDORB_MEM_HOST 0   /* host memory: the call copies in and out and returns when results are on the host */
#define SDORB_MEM_DEVICE 1 /* device memory of the handle's GPU: the call only enqueues work on `stream` */

typedef struct sdorb_handle sdorb_handle;

/* Mirrors ORBextractor(int nfeatures, float scaleFactor, int nlevels, int thFAST)
 * (src/ORBextractor.h:38, src/ORBextractor.cc:406-407) plus the resources of the GPU handle. */
typedef struct {
  int nfeatures;
  float scale_factor;
  int nlevels;
  int th_fast;      /* the reference's single FAST threshold (thFAST / north-star iniThFAST) */
  int min_th_fast;  /* -1 = reference behaviour (SD-SLAM has one FAST threshold and no quadtree).  >= 0 selects the
                       ORB-SLAM2-style mode (the north star's iniThFAST / minThFAST /
                       DistributeOctTree): th_fast is iniThFAST, cells of about 30 pixels fall back to min_th_fast when
                       iniThFAST finds nothing in them, DistributeOctTree culls every level.  That algorithm is NOT in
                       /root/reference; it follows the public ORB-SLAM2 source, and a frame may then yield up to
                       sdorb_max_keypoints() > nfeatures keypoints. */
  int device;       /* CUDA device ordinal; -1 = the calling thread's current device */
  int max_width, max_height; /* largest image the handle will be given (<= 4095 each) */
  int max_batch;    /* frames processed per internal pass; device scratch is sized for this many */
} sdorb_params;

/* Binary-compatible with cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id. */
typedef struct {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} sdorb_keypoint;

/* Destination of one pyramid level (src/ORBextractor.cc:684-686: imagePyramid[l] is a view into a buffer
 * padded by 19 px).  `data` points at the level's pixel (0,0); the callee writes width*height bytes with
 * the given row stride and, when `border` > 0, also fills `border` pixels around it with BORDER_REFLECT_101
 * (the caller must own that margin). */
typedef struct {
  uint8_t* data;
  int width, height;
  size_t stride;
  int border;
} sdorb_pyr_view;

/* Result of the best / second-best scan for one query descriptor (src/ORBmatcher.cc:1239-1265). */
typedef struct {
  int32_t best_idx;    /* first index of the minimum distance, -1 if there was no candidate */
  int32_t best_dist;   /* 256 if none */
  int32_t second_dist; /* second smallest distance (may equal best_dist), 256 if none */
  int32_t accepted;    /* best_dist < th_low && (float)best_dist < ratio * (float)second_dist */
} sdorb_match;

/* ---- lifetime ---- */
SDORB_API int sdorb_create(const sdorb_params* params, sdorb_handle** out);
SDORB_API void sdorb_destroy(sdorb_handle* h);
SDORB_API const char* sdorb_strerror(int code);
/* Text of the last CUDA error seen by this handle ("" if none). */
SDORB_API const char* sdorb_last_cuda_error(const sdorb_handle* h);

/* ---- ORBextractor getters (src/ORBextractor.h:48-70): arrays of nlevels entries, any may be NULL ---- */
SDORB_API int sdorb_get_tables(const sdorb_handle* h, float* scale_factors, float* inv_scale_factors, float* level_sigma2,
                     float* inv_level_sigma2, int* n_features_per_level);
/* Upper bound of keypoints per frame = sum of the per-level targets (normally == nfeatures; the rounding of
 * src/ORBextractor.cc:428-434 can exceed it by a few).  Output capacities must be at least this. */
SDORB_API int sdorb_max_keypoints(const sdorb_handle* h);
/* Size of pyramid level `level` for a width x height input (src/ORBextractor.cc:683). */
SDORB_API int sdorb_level_size(const sdorb_handle* h, int width, int height, int level, int* level_width, int* level_height);

/* ---- ORBextractor::operator() for one host image (src/ORBextractor.cc:620-678, call site src/Frame.cc:195) ----
 * image: 8-bit gray, `stride` bytes per row.  keypoints / descriptors: caller-allocated, capacity >= nfeatures
 * entries / rows of 32 bytes.  pyramid: NULL or nlevels views to receive the image pyramid.  Blocks until the
 * results are in host memory.  An empty image (NULL / zero size) returns SDORB_OK with *count untouched,
 * like the reference's early return (src/ORBextractor.cc:622-623). */
SDORB_API int sdorb_extract(sdorb_handle* h, const uint8_t* image, int width, int height, size_t stride,
                  sdorb_keypoint* keypoints, uint8_t* descriptors, int capacity, int* count,
                  const sdorb_pyr_view* pyramid);

/* ---- the same operator over a batch of equally sized frames ----
 * images: frame f starts at images + f*frame_stride, rows `row_stride` bytes apart.
 * keypoints: [nframes][capacity]; descriptors: [nframes][capacity][32]; counts: [nframes]; capacity >= nfeatures.
 * mem = SDORB_MEM_DEVICE: all four buffers are device pointers on the handle's GPU, work is enqueued on
 *   `stream` (a cudaStream_t passed as void*, NULL = the handle's own stream) and the call returns without
 *   synchronising; sdorb_batch_status() reports deferred errors after the stream is synchronised.
 * mem = SDORB_MEM_HOST: buffers are host pointers (pinned memory gives full copy speed); the call stages
 *   frames through the GPU in passes of max_batch, overlapping copies with kernels, and returns when all
 *   results are on the host. */
SDORB_API int sdorb_extract_batch(sdorb_handle* h, const uint8_t* images, int nframes, int width, int height,
                        size_t row_stride, size_t frame_stride, sdorb_keypoint* keypoints, uint8_t* descriptors,
                        int32_t* counts, int capacity, int mem, void* stream);

/* ---- the batch operator with the reference's FIFTH output, imagePyramid (src/ORBextractor.cc:620-621, consumed by
 * src/ImageAlign.cc:57,93) ----
 * pyramid: NULL (then identical to sdorb_extract_batch) or a frame-major slab in the same memory space as the other buffers:
 * level l of frame f starts at pyramid + f * frame_bytes + level_offset[l] with tightly packed rows (stride = level width);
 * sdorb_pyramid_layout() gives level_offset[nlevels] and frame_bytes for an input size (all multiples of 16; a device slab
 * must be 16-byte aligned).  first_level = 0 writes every level (level 0 = a copy of the input, as the reference returns it),
 * first_level = 1 leaves the level-0 bytes of every frame untouched (the caller already holds level 0: its input). */
SDORB_API int sdorb_pyramid_layout(const sdorb_handle* h, int width, int height, size_t* level_offset, size_t* frame_bytes);

/* ---- one batch over several GPUs of one node, from C / C++ ----
 * handles[g]: one handle per GPU (sdorb_params.device), all created with the same extractor parameters.  Handle g extracts
 * the contiguous frame range sdorb_shard_range(g, ndev, nframes) = [g * nframes / ndev, (g + 1) * nframes / ndev) on its own host
 * thread, through the host pipeline of sdorb_extract_batch_pyr (host buffers; pinned memory gives full copy speed), straight
 * into the caller's slabs: the result is byte-identical to one GPU doing the whole batch.  No collective, no exchange.
 * Returns the first error in handle order; pyramid may be NULL. */
SDORB_API int sdorb_shard_range(int shard, int nshards, int nframes, int* first, int* last);
SDORB_API int sdorb_extract_batch_multi(sdorb_handle* const* handles, int ndev, const uint8_t* images, int nframes, int width,
                              int height, size_t row_stride, size_t frame_stride, sdorb_keypoint* keypoints,
                              uint8_t* descriptors, int32_t* counts, int capacity, uint8_t* pyramid, int first_level);
SDORB_API int sdorb_extract_batch_pyr(sdorb_handle* h, const uint8_t* images, int nframes, int width, int height,
                            size_t row_stride, size_t frame_stride, sdorb_keypoint* keypoints, uint8_t* descriptors,
                            int32_t* counts, int capacity, uint8_t* pyramid, int first_level, int mem, void* stream);
/* After synchronising: SDORB_OK, or the first deferred device-side error (SDORB_ERR_OVERFLOW / _CUDA). */
SDORB_API int sdorb_batch_status(sdorb_handle* h);

/* ---- ORBmatcher::DescriptorDistance, batched ----
 * For pair p, every query row i < nA[p] of descA + p*strideA_rows*32 is compared against all rows j < nB[p] of
 * descB + p*strideB_rows*32 in ascending j with the reference's update rule
 *   if (d < best1) { best2 = best1; best1 = d; idx = j; } else if (d < best2) best2 = d;
 * out: [npairs][strideA_rows] (entries i >= nA[p] are not written).  mem / stream as above. */
SDORB_API int sdorb_match_batch(sdorb_handle* h, const uint8_t* descA, const int32_t* nA, int strideA_rows,
                      const uint8_t* descB, const int32_t* nB, int strideB_rows, int npairs, float ratio,
                      int th_low, sdorb_match* out, int mem, void* stream);
/* SearchByPoints' greedy form (src/ORBmatcher.cc:1228-1270): queries are visited in order and a train row that
 * has been accepted by an earlier query is skipped by later ones (vbMatched2). */
SDORB_API int sdorb_match_greedy_batch(sdorb_handle* h, const uint8_t* descA, const int32_t* nA, int strideA_rows,
                             const uint8_t* descB, const int32_t* nB, int strideB_rows, int npairs, float ratio,
                             int th_low, sdorb_match* out, int mem, void* stream);
/* Full distance matrix out[i*nB + j] = DescriptorDistance(A_i, B_j) for one pair. */
SDORB_API int sdorb_hamming_matrix(sdorb_handle* h, const uint8_t* descA, int nA, const uint8_t* descB, int nB,
                         uint16_t* out, int mem, void* stream);

/* ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:225-284), batched over map points ----
 * Set s holds the descriptors observed for one map point: rows [offsets[s], offsets[s+1]) of desc (32 bytes each, at most
 * 65535 rows per set).  best_idx[s] = row (relative to the set) with the least median DescriptorDistance to the set,
 * the first on ties, exactly as the reference picks it (median = element (size_t)(0.5*(N-1)) of the sorted row,
 * src/MapPoint.cc:266-274); -1 for an empty set.  best_median (may be NULL) receives that median.  mem / stream as above. */
SDORB_API int sdorb_distinctive_batch(sdorb_handle* h, const uint8_t* desc, const int32_t* offsets, int nsets,
                            int32_t* best_idx, int32_t* best_median, int mem, void* stream);

/* ---- Frame post-processing right after the extractor (src/Frame.cc), batched over frames ----
 * Keypoints are [nframes][capacity] sdorb_keypoint with counts[nframes], as the extractor leaves them.
 *
 * sdorb_assign_grid_batch = Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:179-192, 323-332) on the UNDISTORTED
 * keypoints (identical to the extractor's when the camera has no distortion, src/Frame.cc:336-339): grid 64 x 48
 * (FRAME_GRID_COLS / ROWS, src/Frame.h:37-38), cell of a keypoint = (round((x - min_x) * inv_w), round((y - min_y) * inv_h)),
 * keypoints outside the grid are dropped.  Output in CSR form per frame: cell c = posX * 48 + posY owns
 * indices[cell_start[c] .. cell_start[c + 1]), ascending (the reference's push_back order); cell_start has 64 * 48 + 1
 * entries per frame, indices `capacity` entries per frame.  min_x / min_y / inv_w / inv_h are mnMinX, mnMinY,
 * mfGridElementWidthInv, mfGridElementHeightInv (src/Frame.cc:106-107, 368-397). */
#define SDORB_GRID_COLS 64
#define SDORB_GRID_ROWS 48
SDORB_API int sdorb_assign_grid_batch(sdorb_handle* h, const sdorb_keypoint* keypoints_un, const int32_t* counts, int nframes,
                            int capacity, float min_x, float min_y, float inv_w, float inv_h, int32_t* cell_start,
                            int32_t* indices, int mem, void* stream);
/* sdorb_undistort_keypoints_batch = Frame::UndistortKeyPoints (src/Frame.cc:335-366), i.e. cv::undistortPoints(pts, K, dist,
 * noArray(), K) as OpenCV 4.13 computes it, applied to pt of every keypoint (all other fields copied).  K = {fx, fy, cx, cy}
 * and dist = {k1, k2, p1, p2[, k3]} are the float values the reference passes (host pointers, 4 and ndist <= 12 entries);
 * dist[0] == 0 copies the keypoints (src/Frame.cc:336-339).  out may alias keypoints.  nframes <= SDORB_MAX_GRID_BATCH. */
SDORB_API int sdorb_undistort_keypoints_batch(sdorb_handle* h, const sdorb_keypoint* keypoints, const int32_t* counts, int nframes,
                                    int capacity, const float* K, const float* dist, int ndist, sdorb_keypoint* out, int mem,
                                    void* stream);
/* Frame::ComputeImageBounds (src/Frame.cc:368-397) on the host: bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY}; needs no GPU. */
SDORB_API int sdorb_host_image_bounds(int cols, int rows, const float* K, const float* dist, int ndist, float* bounds);
/* sdorb_stereo_from_rgbd_batch = Frame::ComputeStereoFromRGBD (src/Frame.cc:399-417): d = depth(v, u) at the truncated
 * keypoint position; d > 0: z = d and u_right = x_undistorted - mbf / d, else both -1.  depth: float32 images, row /
 * frame strides in ELEMENTS; u_right, z: [nframes][capacity] (entries beyond counts[f] are set to -1).  nframes <= SDORB_MAX_GRID_BATCH. */
SDORB_API int sdorb_stereo_from_rgbd_batch(sdorb_handle* h, const sdorb_keypoint* keypoints, const sdorb_keypoint* keypoints_un,
                                 const int32_t* counts, int nframes, int capacity, const float* depth, int width, int height,
                                 size_t depth_row_stride, size_t depth_frame_stride, float mbf, float* u_right, float* z,
                                 int mem, void* stream);

/* ---- guided matchers: the callers either side of DescriptorDistance (src/ORBmatcher.cc), batched over frame pairs ----
 * All arrays are slabs [npairs][capacity] (keypoints: sdorb_keypoint, descriptors: 32 bytes per row) with per-pair counts;
 * the grid of the searched frame is what sdorb_assign_grid_batch produced for it (cell_start [npairs][64*48+1], indices
 * [npairs][capacity]) together with the four Frame members of the window query Frame::GetFeaturesInArea
 * (src/Frame.cc:271-321): mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv.  Queries are visited in index
 * order and every accepted match changes what later queries may take, exactly as in the reference; the results are
 * identical to running the reference function pair by pair.  capacity <= 16384.  mem / stream as above. */
typedef struct {
  const int32_t* cell_start;
  const int32_t* indices;
  float min_x, min_y, inv_w, inv_h;
} sdorb_frame_grid;

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:256-357), the
 * matcher constructed as ORBmatcher(nnratio, check_orientation) (:40).  prev_matched [npairs][capacity][2] (x, y) is
 * updated in place (:349-352); matches12 [npairs][capacity] receives vnMatches12 (-1 = none; entries >= n1 are -1);
 * nmatches [npairs] the return value. */
SDORB_API int sdorb_search_for_initialization_batch(sdorb_handle* h, const sdorb_keypoint* kps1_un, const uint8_t* desc1,
                                                    const int32_t* n1, const sdorb_keypoint* kps2_un, const uint8_t* desc2,
                                                    const int32_t* n2, const sdorb_frame_grid* grid2, int npairs, int capacity,
                                                    float* prev_matched, int window_size, float nnratio, int check_orientation,
                                                    int32_t* matches12, int32_t* nmatches, int mem, void* stream);

/* ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (src/ORBmatcher.cc:946-1075) from the projection
 * on.  The pose algebra of :955-987 stays with the caller (Eigen, double): per last-frame keypoint i it passes
 * proj[i] = (u, v, invzc) as the floats of :983-987, flags_last[i] (bit 0: mvpMapPoints[i] is set and not an outlier,
 * :966-971; bit 1: that map point has Observations() > 0) and desc_mp[i] = pMP->GetDescriptor(); mode = 1 for bForward,
 * 2 for bBackward, 0 for neither (:962-963).  Current frame: undistorted keypoints, descriptors, u_right (mvuRight),
 * occupied[i2] != 0 where mvpMapPoints[i2] is set with Observations() > 0 on entry (:1018-1020), its grid, bounds =
 * {mnMinX, mnMaxX, mnMinY, mnMaxY}; scale_factors = mvScaleFactors (nlevels <= 32 host floats).
 * assigned [npairs][capacity]: for every current-frame keypoint the last-frame index whose map point the call leaves in
 * CurrentFrame.mvpMapPoints, -1 where the call sets none (or clears it in the rotation check); nmatches [npairs].
 * The overload SearchByProjection(CurrentFrame, KeyFrame* pKF, th, bMono) (src/ORBmatcher.cc:1077-1207) is the same text from the
 * projection on with pKF in the place of LastFrame (kps_last = pKF->mvKeys, kps_last_un = pKF->mvKeysUn, flags bit 0 = the map
 * point exists and !isBad(), :1101-1103): it is served by this entry unchanged. */
typedef struct {
  const sdorb_keypoint* kps_last;     /* LastFrame.mvKeys (octave) */
  const sdorb_keypoint* kps_last_un;  /* LastFrame.mvKeysUn (angle) */
  const float* proj;                  /* [npairs][capacity][3] */
  const uint8_t* flags_last;          /* [npairs][capacity] */
  const uint8_t* desc_mp;             /* [npairs][capacity][32] */
  const int32_t* n_last;
  const sdorb_keypoint* kps_cur_un;
  const uint8_t* desc_cur;
  const float* u_right_cur;           /* [npairs][capacity] */
  const uint8_t* occupied_cur;        /* [npairs][capacity] */
  const int32_t* n_cur;
  sdorb_frame_grid grid_cur;
  const float* scale_factors;         /* host pointer, nlevels entries */
  int nlevels;
  float bounds[4];
  float th, mbf;
  int mode, check_orientation;
  int orb_dist;                       /* distance threshold; 0 = TH_HIGH (100).  With it the same entry point serves
                                         SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1306-1421):
                                         kps_last.octave = nPredictedLevel, mode 0, flags bit 1 set, u_right_cur = -1,
                                         occupied_cur = (mvpMapPoints[i2] != NULL), proj[2] = any value >= 0 */
} sdorb_projection_search;
SDORB_API int sdorb_search_by_projection_batch(sdorb_handle* h, const sdorb_projection_search* q, int npairs, int capacity,
                                               int32_t* assigned, int32_t* nmatches, int mem, void* stream);

/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th) (src/ORBmatcher.cc:43-119, with
 * RadiusByViewingCos :121-126): the local-map search of Tracking::SearchLocalPoints, batched over frames.  Map-point slabs are
 * [nframes][capacity_mp]: proj = (mTrackProjX, mTrackProjY, mTrackProjXR), view_cos = mTrackViewCos, level = mnTrackScaleLevel,
 * flags (bit 0: mbTrackInView && !isBad(); bit 1: Observations() > 0), desc_mp = GetDescriptor().  Frame slabs are
 * [nframes][capacity]: undistorted keypoints, descriptors, mvuRight, occupied (mvpMapPoints[idx] set with Observations() > 0 on
 * entry), and the frame's grid.  The map points are visited in order and every accepted match occupies its keypoint for the
 * later ones, exactly as in the reference.  assigned [nframes][capacity]: index of the map point the call leaves in
 * F.mvpMapPoints[idx], -1 where it sets none; nmatches [nframes].  The matcher is the one constructed as ORBmatcher(nnratio). */
typedef struct {
  const float* proj;        /* [nframes][capacity_mp][3] */
  const float* view_cos;    /* [nframes][capacity_mp] */
  const int32_t* level;     /* [nframes][capacity_mp] */
  const uint8_t* flags;     /* [nframes][capacity_mp] */
  const uint8_t* desc_mp;   /* [nframes][capacity_mp][32] */
  const int32_t* n_mp;      /* [nframes] */
  const sdorb_keypoint* kps_un;
  const uint8_t* desc;
  const float* u_right;
  const uint8_t* occupied;
  const int32_t* n_frame;
  sdorb_frame_grid grid;
  const float* scale_factors; /* host pointer, nlevels entries */
  int nlevels;
  float th, nnratio;
  /* 1: the loop-closing overload ORBmatcher::SearchByProjection(KeyFrame* pKF, Scw, vpPoints, vpMatched, th)
   * (src/ORBmatcher.cc:146-254) from the projection on (the Sim3 algebra :153-209 stays with the caller): radius
   * th * mvScaleFactors[level] (th = the int of :146 as a float), candidates from KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:529-565)
   * with the level window :233-236, single best <= TH_LOW, no ratio test, no mvuRight test; occupied = vpMatched[idx] != NULL on
   * entry and every match occupies its keypoint (:246); flags bit 0 = !isBad() && not in spAlreadyFound && passed :179-209.
   * assigned[idx] = index of the point written to vpMatched[idx].  view_cos, u_right and nnratio are not read (may be NULL). */
  int sim3_form;
} sdorb_map_point_search;
SDORB_API int sdorb_search_map_points_batch(sdorb_handle* h, const sdorb_map_point_search* q, int nframes, int capacity_mp, int capacity,
                                            int32_t* assigned, int32_t* nmatches, int mem, void* stream);

/* ORBmatcher::SearchByPoints(currentKF, pKF, matches) (src/ORBmatcher.cc:1209-1304; loop detection, src/LoopClosing.cc:255), batched
 * over keyframe pairs.  valid1 / valid2 [npairs][capacity]: the keypoint has a map point that is not bad (:1231-1236, :1246-1251).
 * matches12 [npairs][capacity]: index into the second keyframe whose map point the call puts into matches[idx1], -1 = NULL;
 * nmatches [npairs].  The matcher is the one constructed as ORBmatcher(nnratio, check_orientation). */
SDORB_API int sdorb_search_by_points_batch(sdorb_handle* h, const sdorb_keypoint* kps1_un, const uint8_t* desc1, const uint8_t* valid1,
                                           const int32_t* n1, const sdorb_keypoint* kps2_un, const uint8_t* desc2,
                                           const uint8_t* valid2, const int32_t* n2, int npairs, int capacity, float nnratio,
                                           int check_orientation, int32_t* matches12, int32_t* nmatches, int mem, void* stream);

/* The keypoint search of ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:535-586), batched over
 * keyframes: for every map point that passed the checks of :489-531 (flags bit 0; the caller does the pose algebra) the most
 * similar keypoint inside the radius th * mvScaleFactors[level] whose level is level-1 or level and whose reprojection error
 * passes the chi-square test (:560-580); proj = (u, v, ur), level = PredictScale(...).  best_idx [nframes][capacity_mp] = that
 * keypoint if its distance is <= th_dist (TH_LOW = 50 in the reference), else -1; best_dist the distance (256 = no candidate).
 * Replacing / adding the observation (:588-606) stays with the caller, which applies the results in map-point order and evaluates
 * isBad() / IsInKeyFrame(pKF) (:497) again there: those two are the only checks an earlier iteration's surgery can change.  scale_factors / inv_level_sigma2: host pointers, nlevels
 * entries.  check_reprojection = 0 is the search of the Sim3 overload Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (:682-708, also
 * TH_LOW): no chi-square test; u_right, inv_level_sigma2 and proj[2] are then not read (the pointers may be NULL). */
typedef struct {
  const float* proj;       /* [nframes][capacity_mp][3] */
  const int32_t* level;    /* [nframes][capacity_mp] */
  const uint8_t* flags;    /* [nframes][capacity_mp] */
  const uint8_t* desc_mp;  /* [nframes][capacity_mp][32] */
  const int32_t* n_mp;     /* [nframes] */
  const sdorb_keypoint* kps_un; /* [nframes][capacity]  nframes <= SDORB_MAX_GRID_BATCH. */
  const uint8_t* desc;
  const float* u_right;
  sdorb_frame_grid grid;
  const float* scale_factors;
  const float* inv_level_sigma2;
  int nlevels;
  float th;
  int check_reprojection;
  int th_dist;
} sdorb_fuse_search;
SDORB_API int sdorb_fuse_search_batch(sdorb_handle* h, const sdorb_fuse_search* q, int nframes, int capacity_mp, int capacity,
                                      int32_t* best_idx, int32_t* best_dist, int mem, void* stream);

/* ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) (src/ORBmatcher.cc:734-944) from the projections on, batched
 * over keyframe pairs.  q12: the map points of KF1 (entry i1 belongs to keypoint i1: capacity_mp == capacity, n_mp = N1; flags bit
 * 0 = pMP && !vbAlreadyMatched1[i1] && !isBad() && passed :786-808; desc_mp = pMP->GetDescriptor()) projected into KF2, whose
 * keypoints / descriptors / grid / scale factors the struct carries; q21 the reverse direction (:848-925).  Both searches run without
 * the reprojection gate (check_reprojection must be 0) against TH_HIGH (th_dist is ignored).  match1 / match2 [npairs][capacity] =
 * vnMatch1 / vnMatch2; matches12[i1] = idx2 whose map point the agreement check (:927-941) puts into vpMatches12[i1], else -1;
 * nfound [npairs] the return values.  npairs <= SDORB_MAX_GRID_BATCH. */
SDORB_API int sdorb_search_by_sim3_batch(sdorb_handle* h, const sdorb_fuse_search* q12, const sdorb_fuse_search* q21, int npairs,
                                         int capacity, int32_t* match1, int32_t* match2, int32_t* matches12, int32_t* nfound, int mem,
                                         void* stream);

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs) (src/ORBmatcher.cc:359-462, with CheckDistEpipolarLine
 * :128-144) from the epipole on: the pose algebra of :361-368 stays with the caller, which passes per pair F12 (row-major
 * doubles, F12(i, j) = F12[3 * i + j]) and the epipole (ex, ey) as the floats of :367-368.  has_mp1 / has_mp2 [npairs][capacity]:
 * the keypoint already has a map point (GetMapPointMatches()[idx] != NULL); u_right: mvuRight; scale_factors / level_sigma2:
 * pKF2's mvScaleFactors / mvLevelSigma2 (host pointers, nlevels <= 32 entries).  matches12 [npairs][capacity] receives vMatches12
 * (-1 = none; vMatchedPairs is the list of (i, matches12[i]) with matches12[i] >= 0), nmatches [npairs] the return value. */
typedef struct {
  const sdorb_keypoint* kps1_un;
  const uint8_t* desc1;
  const uint8_t* has_mp1;
  const float* u_right1;
  const int32_t* n1;
  const sdorb_keypoint* kps2_un;
  const uint8_t* desc2;
  const uint8_t* has_mp2;
  const float* u_right2;
  const int32_t* n2;
  const double* F12;     /* [npairs][9] */
  const float* epipole;  /* [npairs][2] */
  const float* scale_factors;
  const float* level_sigma2;
  int nlevels;
  int check_orientation;
} sdorb_triangulation_search;
SDORB_API int sdorb_search_for_triangulation_batch(sdorb_handle* h, const sdorb_triangulation_search* q, int npairs, int capacity,
                                                   int32_t* matches12, int32_t* nmatches, int mem, void* stream);

/* ---- host helpers (pure CPU table arithmetic, usable without a CUDA device) ---- */
/* BORDER_REFLECT_101 margin around a level (src/ORBextractor.cc:692-696), used by the C++ shim. */
SDORB_API void sdorb_fill_border_reflect101(uint8_t* level_origin, int width, int height, size_t stride, int border);
/* The constructor tables of src/ORBextractor.cc:406-457 without creating a handle: four float arrays and
 * mnFeaturesPerLevel of nlevels entries each, umax of 16 entries.  Any pointer may be NULL. */
SDORB_API int sdorb_host_tables(int nfeatures, float scale_factor, int nlevels, float* scale_factors,
                      float* inv_scale_factors, float* level_sigma2, float* inv_level_sigma2,
                      int* n_features_per_level, int* umax);
/* Level size (src/ORBextractor.cc:683) and cell grid (src/ORBextractor.cc:469-488) of every level for a
 * width x height input.  Returns SDORB_ERR_GEOMETRY where the reference would throw (see above). */
typedef struct {
  int width, height;   /* level image size */
  int n_desired;       /* mnFeaturesPerLevel[level] */
  int level_cols, level_rows;
  int cell_w, cell_h;
  int n_features_cell;
  int scaled_patch_size;
} sdorb_level_geom;
SDORB_API int sdorb_host_level_geometry(int nfeatures, float scale_factor, int nlevels, int th_fast, int width,
                              int height, sdorb_level_geom* out /* nlevels entries */);

/* ---- instrumentation (bench.py / tests) ---- */
#define SDORB_STAGE_PYRAMID 0
#define SDORB_STAGE_FAST 1
#define SDORB_STAGE_SELECT 2
#define SDORB_STAGE_BLUR 3
#define SDORB_STAGE_DESCRIBE 4
#define SDORB_STAGE_MATCH 5
#define SDORB_NUM_STAGES 6
/* When enabled, every stage of the device path is bracketed by CUDA events on the launching stream. */
SDORB_API int sdorb_set_profiling(sdorb_handle* h, int enabled);
/* Synchronises, then returns accumulated milliseconds and kernel-launch counts per stage since the last reset. */
SDORB_API int sdorb_get_stage_times(sdorb_handle* h, double* ms, int64_t* launches, int reset);
/* Total kernels this handle has launched since creation. */
SDORB_API int64_t sdorb_kernel_launches(const sdorb_handle* h);
/* Copies an intermediate of the most recent pass to the host (parity tests): what = SDORB_DBG_*;
 * returns bytes written or a negative error. */
#define SDORB_DBG_PYRAMID_LEVEL 0 /* width*height bytes of frame `frame`, level `level`, rows packed */
#define SDORB_DBG_BLURRED_LEVEL 1
#define SDORB_DBG_CELL_COUNTS 2   /* int32 per cell of the level (row-major cells): FAST keypoints after NMS */
#define SDORB_DBG_LEVEL_SELECTED 3 /* uint32 (y<<20 | x<<8 | score) per selected keypoint of the level, in order */
SDORB_API int64_t sdorb_debug_read(sdorb_handle* h, int what, int frame, int level, void* dst, size_t capacity);
/* Runs the device implementation of std::nth_element(e, e + nth, e + n, response >) -- the ordering core of
 * KeyPointsFilter::retainBest (src/ORBextractor.cc:586, 602) -- on n packed entries (low 8 bits = response) in host
 * memory, in place (parity tests against the real libstdc++ algorithm). */
SDORB_API int sdorb_debug_nth_element(sdorb_handle* h, uint32_t* entries, int n, int nth);
/* Guarded run (SDORB_GUARD=1 in the environment when the library makes its first allocation): every device buffer has a 256 KB
 * guard band on either side and a poisoned payload.  Returns the number of guard bytes any kernel has overwritten so far (0 = all
 * bands intact; the first damaged buffer is named by sdorb_last_cuda_error), -1000 when the run is not guarded. */
SDORB_API int64_t sdorb_debug_guard_check(sdorb_handle* h);
/* Measured peak of an execution pipe on the handle's GPU (a saturating micro-benchmark, kernels_probe.cu), for the roofline
 * figures of bench.py: pipe 0 = POPC (bounds the Hamming matcher), 1 = the integer ALU pipe on VIMNMX3.U16x2 (bounds FAST),
 * 2 = PRMT.  Rates in warp-instructions per second (whole GPU) and per SM clock per SM. */
SDORB_API int sdorb_debug_pipe_probe(sdorb_handle* h, int pipe, double* warp_instr_per_s, double* warp_instr_per_clk_per_sm);

#ifdef __cplusplus
}
#endif
#endif /* SDORB_H */
