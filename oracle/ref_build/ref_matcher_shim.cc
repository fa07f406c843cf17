// oracle/ref_build/ref_matcher_shim.cc -- C interface around the UNMODIFIED reference object graph, for the parity tests of the
// matcher / Frame rows.
//
// TEST INFRASTRUCTURE ONLY.  This file is ours; what it drives is the reference's own text, compiled where it lies:
//   /root/reference/src/{ORBmatcher,Frame,KeyFrame,MapPoint,Map}.{h,cc}
// against oracle/ref_compat (cv:: surface, a sliver of Eigen, a stand-in Converter.h) by oracle/ref_build/Makefile.  The shim
// builds Frame / KeyFrame / MapPoint objects from flat arrays -- the same arrays the oracle functions and the C ABI take -- calls
// the reference method, and flattens the result.  Where a reference routine does pose algebra before it searches (the
// SearchByProjection / Fuse family), the objects are posed so that the algebra is EXACT (identity rotation, zero translation,
// fx = fy = 1, cx = cy = 0, unit depth: the projection of (u, v, 1) is (u, v) with invz = 1 in one rounding-free step), so the
// comparison starts at the same (u, v, invz) the oracle and the CUDA entry points are given and nothing depends on how the Eigen
// stand-in orders its double arithmetic.
//
// `#define private public` below is applied to the reference HEADERS as seen by THIS translation unit only (the reference's own
// translation units are compiled untouched): Frame keeps UndistortKeyPoints / ComputeImageBounds / AssignFeaturesToGrid private
// and MapPoint / KeyFrame keep the members a test has to set protected.  Access specifiers do not change the class layout.
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include <opencv2/opencv.hpp>
#include <Eigen/Dense>

#define private public
#define protected public
#include "Map.h"
#include "MapPoint.h"
#include "KeyFrame.h"
#include "Frame.h"
#include "ORBmatcher.h"
#undef private
#undef protected

using namespace SD_SLAM;

namespace {
std::mutex g_lock;  // Frame keeps its bounds / grid scale / intrinsics in statics: one call at a time

struct GridParams {
  float min_x, max_x, min_y, max_y, inv_w, inv_h;
};

void set_frame_statics(const GridParams& g, float fx, float fy, float cx, float cy) {
  Frame::mnMinX = g.min_x;
  Frame::mnMaxX = g.max_x;
  Frame::mnMinY = g.min_y;
  Frame::mnMaxY = g.max_y;
  Frame::mfGridElementWidthInv = g.inv_w;
  Frame::mfGridElementHeightInv = g.inv_h;
  Frame::fx = fx;
  Frame::fy = fy;
  Frame::cx = cx;
  Frame::cy = cy;
  Frame::invfx = 1.0f / fx;
  Frame::invfy = 1.0f / fy;
  Frame::mbInitialComputations = false;
}

std::vector<cv::KeyPoint> to_kps(const orc_keypoint* k, int n) {
  std::vector<cv::KeyPoint> v((size_t)n);
  static_assert(sizeof(cv::KeyPoint) == sizeof(orc_keypoint), "cv::KeyPoint layout");
  if (n > 0) memcpy(static_cast<void*>(&v[0]), k, sizeof(orc_keypoint) * (size_t)n);
  return v;
}

cv::Mat to_desc(const uint8_t* d, int n) {
  cv::Mat m(n, 32, CV_8UC1);
  if (n > 0) memcpy(m.data, d, (size_t)n * 32);
  return m;
}

// A Frame as the reference's constructors leave it (src/Frame.cc:68-123), from flat arrays instead of an image.
void fill_frame(Frame& F, const orc_keypoint* kps, const orc_keypoint* kps_un, const uint8_t* desc, int n, const float* scale_factors,
                int nlevels, const float* u_right, float mbf) {
  F.N = n;
  F.mvKeys = to_kps(kps ? kps : kps_un, n);
  F.mvKeysUn = to_kps(kps_un, n);
  F.mDescriptors = to_desc(desc, n);
  F.mvuRight.assign((size_t)n, -1.0f);
  if (u_right)
    for (int i = 0; i < n; ++i) F.mvuRight[(size_t)i] = u_right[i];
  F.mvDepth.assign((size_t)n, -1.0f);
  F.mvpMapPoints.assign((size_t)n, static_cast<MapPoint*>(NULL));
  F.mvbOutlier.assign((size_t)n, false);
  F.mnScaleLevels = nlevels;
  F.mfScaleFactor = nlevels > 1 ? scale_factors[1] : 1.2f;
  F.mfLogScaleFactor = std::log(F.mfScaleFactor);
  F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
  F.mvInvScaleFactors.resize((size_t)nlevels);
  F.mvLevelSigma2.resize((size_t)nlevels);
  F.mvInvLevelSigma2.resize((size_t)nlevels);
  for (int l = 0; l < nlevels; ++l) {
    F.mvInvScaleFactors[(size_t)l] = 1.0f / scale_factors[l];
    F.mvLevelSigma2[(size_t)l] = scale_factors[l] * scale_factors[l];
    F.mvInvLevelSigma2[(size_t)l] = 1.0f / F.mvLevelSigma2[(size_t)l];
  }
  F.mbf = mbf;
  F.mb = mbf / Frame::fx;
  F.mThDepth = 0;
  F.mnId = Frame::nNextId++;
  F.mpReferenceKF = NULL;
  F.mpORBextractorLeft = NULL;
  F.mK = Eigen::Matrix3d::Identity();
  F.mDistCoef = cv::Mat(4, 1, CV_32F);
  for (int i = 0; i < 4; ++i) F.mDistCoef.at<float>(i) = 0.f;
  F.SetPose(Eigen::Matrix4d::Identity());
  F.AssignFeaturesToGrid();  // src/Frame.cc:179-192 (private in the reference)
}

const float kDefaultScales[8] = {1.f, 1.2f, 1.44f, 1.728f, 2.0736f, 2.48832f, 2.985984f, 3.5831808f};

// owns every MapPoint / KeyFrame a call creates (the reference manages them through Map and raw pointers)
struct Arena {
  Map map;
  std::vector<MapPoint*> points;
  std::vector<KeyFrame*> keyframes;
  ~Arena() {
    for (size_t i = 0; i < points.size(); ++i) delete points[i];
    for (size_t i = 0; i < keyframes.size(); ++i) delete keyframes[i];
  }
  KeyFrame* keyframe(Frame& F) {
    KeyFrame* kf = new KeyFrame(F, &map);  // src/KeyFrame.cc:36-76
    keyframes.push_back(kf);
    return kf;
  }
  MapPoint* point(const Eigen::Vector3d& pos, KeyFrame* ref) {
    MapPoint* p = new MapPoint(pos, ref, &map);  // src/MapPoint.cc:38-49
    points.push_back(p);
    return p;
  }
  // a one-keypoint keyframe to hang observations on (MapPoint::AddObservation reads pKF->mvuRight[idx])
  KeyFrame* observer(const float* scale_factors, int nlevels) {
    orc_keypoint kp0;
    memset(&kp0, 0, sizeof(kp0));
    std::vector<uint8_t> d0(32, 0);
    Frame F;
    fill_frame(F, NULL, &kp0, &d0[0], 1, scale_factors, nlevels, NULL, 0);
    return keyframe(F);
  }
};
}  // namespace

extern "C" {

// ---- Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:179-192, 323-332) in the oracle's CSR form: cell x * 48 + y
void ref_assign_grid(const orc_keypoint* kps_un, int n, float min_x, float min_y, float inv_w, float inv_h, int32_t* cell_start,
                     int32_t* indices) {
  std::lock_guard<std::mutex> lk(g_lock);
  GridParams g = {min_x, 0, min_y, 0, inv_w, inv_h};
  set_frame_statics(g, 1, 1, 0, 0);
  Frame F;
  fill_frame(F, NULL, kps_un, NULL, 0, kDefaultScales, 8, NULL, 0);  // empty first: the grid below is filled by the real method
  F.N = n;
  F.mvKeysUn = to_kps(kps_un, n);
  for (int i = 0; i < FRAME_GRID_COLS; ++i)
    for (int j = 0; j < FRAME_GRID_ROWS; ++j) F.mGrid[i][j].clear();
  F.AssignFeaturesToGrid();
  int acc = 0;
  for (int x = 0; x < FRAME_GRID_COLS; ++x)
    for (int y = 0; y < FRAME_GRID_ROWS; ++y) {
      cell_start[x * FRAME_GRID_ROWS + y] = acc;
      for (size_t k = 0; k < F.mGrid[x][y].size(); ++k) indices[acc++] = (int32_t)F.mGrid[x][y][k];
    }
  cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = acc;
}

// ---- Frame::GetFeaturesInArea (src/Frame.cc:271-321); returns the count, at most n indices written
int ref_features_in_area(const orc_keypoint* kps_un, int n, float min_x, float min_y, float inv_w, float inv_h, float x, float y, float r,
                         int min_level, int max_level, int32_t* out) {
  std::lock_guard<std::mutex> lk(g_lock);
  GridParams g = {min_x, 0, min_y, 0, inv_w, inv_h};
  set_frame_statics(g, 1, 1, 0, 0);
  Frame F;
  std::vector<uint8_t> d((size_t)n * 32 + 32, 0);
  fill_frame(F, NULL, kps_un, &d[0], n, kDefaultScales, 8, NULL, 0);
  const std::vector<size_t> v = F.GetFeaturesInArea(x, y, r, min_level, max_level);
  for (size_t i = 0; i < v.size() && (int)i < n; ++i) out[i] = (int32_t)v[i];
  return (int)v.size();
}

// ---- Frame::UndistortKeyPoints (src/Frame.cc:335-366); K = {fx, fy, cx, cy}
void ref_undistort_keypoints(const orc_keypoint* kps, int n, const float K[4], const float* dist, int ndist, orc_keypoint* out) {
  std::lock_guard<std::mutex> lk(g_lock);
  Frame F;
  F.N = n;
  F.mvKeys = to_kps(kps, n);
  F.mK = Eigen::Matrix3d::Identity();
  F.mK(0, 0) = K[0];
  F.mK(1, 1) = K[1];
  F.mK(0, 2) = K[2];
  F.mK(1, 2) = K[3];
  F.mDistCoef = cv::Mat(ndist > 4 ? ndist : 4, 1, CV_32F);
  for (int i = 0; i < F.mDistCoef.rows; ++i) F.mDistCoef.at<float>(i) = i < ndist ? dist[i] : 0.f;
  F.UndistortKeyPoints();
  for (int i = 0; i < n && i < (int)F.mvKeysUn.size(); ++i) memcpy(&out[i], &F.mvKeysUn[(size_t)i], sizeof(orc_keypoint));
}

// ---- Frame::ComputeImageBounds (src/Frame.cc:368-397): bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY}
void ref_image_bounds(int cols, int rows, const float K[4], const float* dist, int ndist, float bounds[4]) {
  std::lock_guard<std::mutex> lk(g_lock);
  Frame F;
  F.mK = Eigen::Matrix3d::Identity();
  F.mK(0, 0) = K[0];
  F.mK(1, 1) = K[1];
  F.mK(0, 2) = K[2];
  F.mK(1, 2) = K[3];
  F.mDistCoef = cv::Mat(ndist > 4 ? ndist : 4, 1, CV_32F);
  for (int i = 0; i < F.mDistCoef.rows; ++i) F.mDistCoef.at<float>(i) = i < ndist ? dist[i] : 0.f;
  uint8_t px = 0;
  cv::Mat img(rows, cols, CV_8UC1, &px, (size_t)cols);  // only .cols / .rows are read
  F.ComputeImageBounds(img);
  bounds[0] = Frame::mnMinX;
  bounds[1] = Frame::mnMaxX;
  bounds[2] = Frame::mnMinY;
  bounds[3] = Frame::mnMaxY;
}

// ---- Frame::ComputeStereoFromRGBD (src/Frame.cc:399-417); depth: tightly packed float image of the given size
void ref_stereo_from_rgbd(const orc_keypoint* kps, const orc_keypoint* kps_un, int n, const float* depth, int width, int height, float mbf,
                          float* u_right, float* z) {
  std::lock_guard<std::mutex> lk(g_lock);
  Frame F;
  F.N = n;
  F.mvKeys = to_kps(kps, n);
  F.mvKeysUn = to_kps(kps_un, n);
  F.mbf = mbf;
  cv::Mat d(height, width, CV_32F, const_cast<float*>(depth), (size_t)width * sizeof(float));
  F.ComputeStereoFromRGBD(d);
  for (int i = 0; i < n; ++i) {
    u_right[i] = F.mvuRight[(size_t)i];
    z[i] = F.mvDepth[(size_t)i];
  }
}

// ---- ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1423-1454) on bins of the given sizes
void ref_three_maxima(const int32_t* sizes, int L, int* ind1, int* ind2, int* ind3) {
  std::vector<std::vector<int> > histo((size_t)L);
  for (int i = 0; i < L; ++i) histo[(size_t)i].assign((size_t)sizes[i], 0);
  ORBmatcher m(0.6f, true);
  *ind1 = *ind2 = *ind3 = -1;
  m.ComputeThreeMaxima(L ? &histo[0] : NULL, L, *ind1, *ind2, *ind3);
}

// ---- ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:256-357); prev_matched: n1 (x, y) pairs, updated in place
int ref_search_for_initialization(const orc_keypoint* kps1_un, const uint8_t* desc1, int n1, const orc_keypoint* kps2_un,
                                  const uint8_t* desc2, int n2, float min_x, float min_y, float inv_w, float inv_h, float* prev_matched,
                                  int window_size, float nnratio, int check_orientation, int32_t* matches12) {
  std::lock_guard<std::mutex> lk(g_lock);
  GridParams g = {min_x, 0, min_y, 0, inv_w, inv_h};
  set_frame_statics(g, 1, 1, 0, 0);
  Frame F1, F2;
  fill_frame(F1, NULL, kps1_un, desc1, n1, kDefaultScales, 8, NULL, 0);
  fill_frame(F2, NULL, kps2_un, desc2, n2, kDefaultScales, 8, NULL, 0);
  std::vector<cv::Point2f> prev((size_t)n1);
  for (int i = 0; i < n1; ++i) prev[(size_t)i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
  std::vector<int> m12;
  ORBmatcher matcher(nnratio, check_orientation != 0);
  const int nm = matcher.SearchForInitialization(F1, F2, prev, m12, window_size);
  for (int i = 0; i < n1; ++i) {
    matches12[i] = m12[(size_t)i];
    prev_matched[2 * i] = prev[(size_t)i].x;
    prev_matched[2 * i + 1] = prev[(size_t)i].y;
  }
  return nm;
}

// ---- ORBmatcher::SearchByPoints (src/ORBmatcher.cc:1209-1304), the loop a12 batches: valid = the keypoint has a good map point;
// matches12[i] = index in the second keyframe of the map point matched to keypoint i of the first, or -1
int ref_search_by_points(const orc_keypoint* kps1_un, const uint8_t* desc1, const uint8_t* valid1, int n1, const orc_keypoint* kps2_un,
                         const uint8_t* desc2, const uint8_t* valid2, int n2, float nnratio, int check_orientation, int32_t* matches12) {
  std::lock_guard<std::mutex> lk(g_lock);
  GridParams g = {0, 640, 0, 480, 64.f / 640.f, 48.f / 480.f};
  set_frame_statics(g, 1, 1, 0, 0);
  Arena A;
  Frame F1, F2;
  fill_frame(F1, NULL, kps1_un, desc1, n1, kDefaultScales, 8, NULL, 0);
  fill_frame(F2, NULL, kps2_un, desc2, n2, kDefaultScales, 8, NULL, 0);
  KeyFrame* k1 = A.keyframe(F1);
  KeyFrame* k2 = A.keyframe(F2);
  std::map<MapPoint*, int> index2;
  for (int i = 0; i < n1; ++i)
    if (valid1[i]) k1->AddMapPoint(A.point(Eigen::Vector3d(0, 0, 1), k1), (size_t)i);
  for (int i = 0; i < n2; ++i)
    if (valid2[i]) {
      MapPoint* p = A.point(Eigen::Vector3d(0, 0, 1), k2);
      k2->AddMapPoint(p, (size_t)i);
      index2[p] = i;
    }
  std::vector<MapPoint*> matches;
  ORBmatcher matcher(nnratio, check_orientation != 0);
  const int nm = matcher.SearchByPoints(k1, k2, matches);
  for (int i = 0; i < n1; ++i) matches12[i] = (i < (int)matches.size() && matches[(size_t)i]) ? index2[matches[(size_t)i]] : -1;
  return nm;
}

// ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:225-284): the map point is observed by n keyframes, keyframe k
// holding descriptor k as the descriptor of its keypoint 0.  Returns the index of the descriptor the map point ends up with
// (found by comparing bytes: first equal row), -1 for n == 0.
int ref_distinctive(const uint8_t* desc, int n) {
  std::lock_guard<std::mutex> lk(g_lock);
  GridParams g = {0, 640, 0, 480, 64.f / 640.f, 48.f / 480.f};
  set_frame_statics(g, 1, 1, 0, 0);
  Arena A;
  if (n <= 0) return -1;
  orc_keypoint kp;
  memset(&kp, 0, sizeof(kp));
  kp.x = kp.y = 10.f;
  kp.class_id = -1;
  std::vector<KeyFrame*> kfs;
  for (int k = 0; k < n; ++k) {
    Frame F;
    fill_frame(F, NULL, &kp, desc + (size_t)k * 32, 1, kDefaultScales, 8, NULL, 0);
    kfs.push_back(A.keyframe(F));
  }
  MapPoint* p = A.point(Eigen::Vector3d(0, 0, 1), kfs[0]);
  for (int k = 0; k < n; ++k) p->AddObservation(kfs[(size_t)k], 0);
  p->ComputeDistinctiveDescriptors();
  const cv::Mat d = p->GetDescriptor();
  if (d.empty()) return -1;
  for (int k = 0; k < n; ++k)
    if (memcmp(d.ptr(0), desc + (size_t)k * 32, 32) == 0) return k;
  return -2;
}

// ---- ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (src/ORBmatcher.cc:946-1075).
// Both frames sit at the identity pose with fx = fy = 1, cx = cy = 0 and the map point of last-frame keypoint i at
// (proj[3i], proj[3i+1], 1): its projection is exactly (u, v) = (proj[3i], proj[3i+1]) with invzc = 1 (the oracle is given
// the same triples).  flags_last bit 0: the keypoint has a map point that is not an outlier, bit 1: that map point has
// Observations() > 0 (so a keypoint of the current frame that receives it counts as taken for later queries); desc_mp: that
// map point's descriptor; occupied_cur[i2]: the current frame's keypoint already has a map point with observations.
// mode 0 / 1 / 2 = neither / bForward / bBackward (:962-963): tlc = Rlw * twc + tlw is just the last frame's translation here,
// set to (0, 0, +-(mb + 1)) -- the last frame's pose enters nothing else.
// assigned[i2] = index of the last-frame keypoint whose map point ends up in CurrentFrame.mvpMapPoints[i2] through this call.
int ref_search_by_projection(const orc_keypoint* kps_last, const orc_keypoint* kps_last_un, const float* proj, const uint8_t* flags_last,
                             const uint8_t* desc_mp, int n_last, const orc_keypoint* kps_cur_un, const uint8_t* desc_cur,
                             const float* u_right_cur, const uint8_t* occupied_cur, int n_cur, const float* scale_factors, int nlevels,
                             const float bounds[4], float inv_w, float inv_h, float th, float mbf, int mode, int check_orientation,
                             int32_t* assigned, int keyframe_overload) {
  std::lock_guard<std::mutex> lk(g_lock);
  GridParams g = {bounds[0], bounds[1], bounds[2], bounds[3], inv_w, inv_h};
  set_frame_statics(g, 1, 1, 0, 0);
  Arena A;
  Frame Last, Cur;
  std::vector<uint8_t> dl((size_t)n_last * 32 + 32, 0);
  fill_frame(Last, kps_last, kps_last_un, &dl[0], n_last, scale_factors, nlevels, NULL, mbf);
  fill_frame(Cur, NULL, kps_cur_un, desc_cur, n_cur, scale_factors, nlevels, u_right_cur, mbf);
  KeyFrame* ref = A.observer(scale_factors, nlevels);
  if (mode != 0) {  // tlc(2) = tz: bForward needs tz > mb, bBackward -tz > mb  (mb = mbf / fx = mbf here)
    Eigen::Matrix4d T = Eigen::Matrix4d::Identity();
    T(2, 3) = (mode == 1 ? 1.0 : -1.0) * ((double)Cur.mb + 1.0);
    Last.SetPose(T);
  }
  std::map<MapPoint*, int> owner;
  for (int i = 0; i < n_last; ++i) {
    if (!(flags_last[i] & 1)) {  // no map point, or an outlier association (:968-971): half of them get the latter
      if (i & 1) {
        Last.mvbOutlier[(size_t)i] = true;
        Last.mvpMapPoints[(size_t)i] = A.point(Eigen::Vector3d(proj[3 * i], proj[3 * i + 1], 1.0), ref);
      }
      continue;
    }
    MapPoint* p = A.point(Eigen::Vector3d(proj[3 * i], proj[3 * i + 1], 1.0), ref);
    p->mDescriptor = to_desc(desc_mp + (size_t)i * 32, 1);
    if (flags_last[i] & 2) p->AddObservation(ref, 0);
    Last.mvpMapPoints[(size_t)i] = p;
    owner[p] = i;
  }
  std::vector<MapPoint*> pre((size_t)n_cur, static_cast<MapPoint*>(NULL));
  for (int i = 0; i < n_cur; ++i)
    if (occupied_cur[i]) {
      MapPoint* p = A.point(Eigen::Vector3d(0, 0, 1), ref);
      p->AddObservation(ref, 0);  // Observations() > 0
      Cur.mvpMapPoints[(size_t)i] = p;
      pre[(size_t)i] = p;
    }
  ORBmatcher matcher(0.9f, check_orientation != 0);
  int nm;
  if (keyframe_overload) {
    // SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, th, bMono) (:1077-1207): the same text with pKF for LastFrame and
    // pMP->isBad() for mvbOutlier[i] -- the outlier associations of the frame become bad map points of the keyframe
    for (int i = 0; i < n_last; ++i)
      if (Last.mvbOutlier[(size_t)i] && Last.mvpMapPoints[(size_t)i]) Last.mvpMapPoints[(size_t)i]->mbBad = true;
    KeyFrame* kl = A.keyframe(Last);
    nm = matcher.SearchByProjection(Cur, kl, th, false);
  } else {
    nm = matcher.SearchByProjection(Cur, Last, th, false);
  }
  for (int i = 0; i < n_cur; ++i) {
    MapPoint* p = Cur.mvpMapPoints[(size_t)i];
    assigned[i] = (p && p != pre[(size_t)i] && owner.count(p)) ? owner[p] : -1;
  }
  return nm;
}

// ---- ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th) (src/ORBmatcher.cc:43-119), the local-map
// search of Tracking::SearchLocalPoints.  Map point i carries what Frame::isInFrustum leaves on it: mbTrackInView (flags bit 0),
// mTrackProjX / Y / XR = proj[3i .. 3i+2], mnTrackScaleLevel = level[i], mTrackViewCos = view_cos[i]; points without bit 0 are
// alternately "not in view" and "bad"; flags bit 1: Observations() > 0 (a keypoint that receives the point counts as taken).
// occupied[k]: the frame's keypoint k already has a map point with observations.  assigned[k] = map point index or -1.
int ref_search_map_points(const float* proj, const float* view_cos, const int32_t* level, const uint8_t* flags, const uint8_t* desc_mp,
                          int n_mp, const orc_keypoint* kps_un, const uint8_t* desc, const float* u_right, const uint8_t* occupied,
                          int n_frame, const float* scale_factors, int nlevels, const float bounds[4], float inv_w, float inv_h, float th,
                          float nnratio, int32_t* assigned) {
  std::lock_guard<std::mutex> lk(g_lock);
  GridParams g = {bounds[0], bounds[1], bounds[2], bounds[3], inv_w, inv_h};
  set_frame_statics(g, 1, 1, 0, 0);
  Arena A;
  Frame F;
  fill_frame(F, NULL, kps_un, desc, n_frame, scale_factors, nlevels, u_right, 40.f);
  KeyFrame* ref = A.observer(scale_factors, nlevels);
  std::vector<MapPoint*> pre((size_t)n_frame, static_cast<MapPoint*>(NULL));
  for (int k = 0; k < n_frame; ++k)
    if (occupied[k]) {
      MapPoint* p = A.point(Eigen::Vector3d(0, 0, 1), ref);
      p->AddObservation(ref, 0);
      F.mvpMapPoints[(size_t)k] = p;
      pre[(size_t)k] = p;
    }
  std::vector<MapPoint*> pts;
  std::map<MapPoint*, int> index;
  for (int i = 0; i < n_mp; ++i) {
    MapPoint* p = A.point(Eigen::Vector3d(0, 0, 1), ref);
    p->mDescriptor = to_desc(desc_mp + (size_t)i * 32, 1);
    p->mbTrackInView = (flags[i] & 1) != 0 || (i & 1);
    p->mbBad = !(flags[i] & 1) && (i & 1);
    if (flags[i] & 2) p->AddObservation(ref, 0);
    p->mTrackProjX = proj[3 * i];
    p->mTrackProjY = proj[3 * i + 1];
    p->mTrackProjXR = proj[3 * i + 2];
    p->mnTrackScaleLevel = level[i];
    p->mTrackViewCos = view_cos[i];
    pts.push_back(p);
    index[p] = i;
  }
  ORBmatcher matcher(nnratio, true);
  const int nm = matcher.SearchByProjection(F, pts, th);
  for (int k = 0; k < n_frame; ++k) {
    MapPoint* p = F.mvpMapPoints[(size_t)k];
    assigned[k] = (p && p != pre[(size_t)k] && index.count(p)) ? index[p] : -1;
  }
  return nm;
}

// ---- ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:359-462) with CheckDistEpipolarLine (:128-144).  KF2 sits at the
// identity pose, KF1's camera centre at (ex, ey, 1): the epipole of :366-368 is exactly (ex, ey).  F12 row-major.
// has_mp: the keypoint already has a map point.  matches12[i] = index in KF2 or -1 (vMatchedPairs flattened); returns nmatches.
int ref_search_for_triangulation(const orc_keypoint* kps1_un, const uint8_t* desc1, const uint8_t* has_mp1, const float* u_right1, int n1,
                                 const orc_keypoint* kps2_un, const uint8_t* desc2, const uint8_t* has_mp2, const float* u_right2, int n2,
                                 const double* F12, float ex, float ey, const float* scale_factors, int nlevels, int check_orientation,
                                 int32_t* matches12) {
  std::lock_guard<std::mutex> lk(g_lock);
  GridParams g = {0, 640, 0, 480, 64.f / 640.f, 48.f / 480.f};
  set_frame_statics(g, 1, 1, 0, 0);
  Arena A;
  Frame F1, F2;
  fill_frame(F1, NULL, kps1_un, desc1, n1, scale_factors, nlevels, u_right1, 40.f);
  fill_frame(F2, NULL, kps2_un, desc2, n2, scale_factors, nlevels, u_right2, 40.f);
  Eigen::Matrix4d T1 = Eigen::Matrix4d::Identity();
  T1(0, 3) = -(double)ex;
  T1(1, 3) = -(double)ey;
  T1(2, 3) = -1.0;
  F1.SetPose(T1);
  KeyFrame* k1 = A.keyframe(F1);
  KeyFrame* k2 = A.keyframe(F2);
  for (int i = 0; i < n1; ++i)
    if (has_mp1[i]) k1->AddMapPoint(A.point(Eigen::Vector3d(0, 0, 1), k1), (size_t)i);
  for (int i = 0; i < n2; ++i)
    if (has_mp2[i]) k2->AddMapPoint(A.point(Eigen::Vector3d(0, 0, 1), k2), (size_t)i);
  Eigen::Matrix3d F;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) F(r, c) = F12[3 * r + c];
  std::vector<std::pair<size_t, size_t> > pairs;
  ORBmatcher matcher(0.6f, check_orientation != 0);
  const int nm = matcher.SearchForTriangulation(k1, k2, F, pairs);
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  for (size_t k = 0; k < pairs.size(); ++k) matches12[pairs[k].first] = (int32_t)pairs[k].second;
  return nm;
}

namespace {
// A map point that ORBmatcher::Fuse / SearchBySim3 project to exactly (u, v) in a keyframe at the identity pose with fx = fy = 1,
// cx = cy = 0, at depth z (a power of two: u * z, 1 / z and (u * z) * (1 / z) are all exact), whose PredictScale() is `level`,
// which passes the distance-invariance and viewing-angle checks, and which carries the given descriptor.
MapPoint* posed_point(Arena& A, KeyFrame* ref, float u, float v, double z, int level, const uint8_t* desc, float log_scale_factor) {
  const Eigen::Vector3d pos((double)u * z, (double)v * z, z);
  MapPoint* p = A.point(pos, ref);
  p->mDescriptor = to_desc(desc, 1);
  const float dist = (float)pos.norm();
  // ratio = mfMaxDistance / dist = scaleFactor^(level - 0.5): ceil(log(ratio) / log(scaleFactor)) = level (0 for level 0)
  p->mfMaxDistance = dist * std::exp(((float)level - 0.5f) * log_scale_factor);
  p->mfMinDistance = 0.f;
  if (p->mfMaxDistance * 1.2f < dist) p->mfMaxDistance = dist;  // level 0: keep dist3D <= 1.2 * mfMaxDistance
  p->mNormalVector = pos / pos.norm();
  return p;
}
}  // namespace

// ---- the keypoint search of ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th) (src/ORBmatcher.cc:477-615).  Map point i projects to
// (proj[3i], proj[3i+1]) at depth 1 / invz[i] (powers of two), so ur = u - bf * invz of :520; flags bit 0 clear: the point is
// bad (skipped).  The keyframe has no map points of its own, so every accepted match goes through AddObservation / AddMapPoint,
// or -- when an earlier point took the keypoint -- through Replace(): best_idx[i] is read back from either.  min_x / min_y are
// the grid origin (KeyFrame keeps them as int), the image is unbounded to the right / below.
void ref_fuse_search(const float* proj, const float* invz, const int32_t* level, const uint8_t* flags, const uint8_t* desc_mp, int n_mp,
                     const orc_keypoint* kps_un, const uint8_t* desc, const float* u_right, int n_kf, float min_x, float min_y, float inv_w,
                     float inv_h, const float* scale_factors, int nlevels, float th, float bf, int32_t* best_idx) {
  std::lock_guard<std::mutex> lk(g_lock);
  GridParams g = {min_x, 1e6f, min_y, 1e6f, inv_w, inv_h};
  set_frame_statics(g, 1, 1, 0, 0);
  Arena A;
  Frame F;
  fill_frame(F, NULL, kps_un, desc, n_kf, scale_factors, nlevels, u_right, bf);
  KeyFrame* kf = A.keyframe(F);
  KeyFrame* obs = A.observer(scale_factors, nlevels);
  std::vector<MapPoint*> pts((size_t)n_mp);
  for (int i = 0; i < n_mp; ++i) {
    MapPoint* p = posed_point(A, obs, proj[3 * i], proj[3 * i + 1], 1.0 / (double)invz[i], level[i], desc_mp + (size_t)i * 32, kf->mfLogScaleFactor);
    if (!(flags[i] & 1)) p->mbBad = true;
    pts[(size_t)i] = p;
  }
  ORBmatcher matcher(0.6f, true);
  matcher.Fuse(kf, pts, th);
  for (int i = 0; i < n_mp; ++i) {
    MapPoint* p = pts[(size_t)i];
    int idx = p->GetIndexInKeyFrame(kf);
    if (idx < 0 && p->GetReplaced()) idx = p->GetReplaced()->GetIndexInKeyFrame(kf);  // the keypoint was taken: this point was replaced by its owner
    best_idx[i] = idx;
  }
}

// ---- ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:734-944) with s12 = 1, R12 = I, t12 = 0 and both keyframes at the identity pose:
// the map point of keypoint i1 of KF1 (flags1 bit 0) lies at (proj1[3 i1], proj1[3 i1 + 1], 1) and so projects there in KF2, and
// vice versa.  matches12[i1] = index in KF2 of the map point the call stores in vpMatches12[i1], else -1; returns nFound.
int ref_search_by_sim3(const float* proj1, const int32_t* level1, const uint8_t* flags1, const uint8_t* desc_mp1, int n1, const float* proj2,
                       const int32_t* level2, const uint8_t* flags2, const uint8_t* desc_mp2, int n2, const orc_keypoint* kps1_un,
                       const uint8_t* desc1, const orc_keypoint* kps2_un, const uint8_t* desc2, float min_x, float min_y, float inv_w,
                       float inv_h, const float* scale_factors, int nlevels, float th, int32_t* matches12) {
  std::lock_guard<std::mutex> lk(g_lock);
  GridParams g = {min_x, 1e6f, min_y, 1e6f, inv_w, inv_h};
  set_frame_statics(g, 1, 1, 0, 0);
  Arena A;
  Frame F1, F2;
  fill_frame(F1, NULL, kps1_un, desc1, n1, scale_factors, nlevels, NULL, 40.f);
  fill_frame(F2, NULL, kps2_un, desc2, n2, scale_factors, nlevels, NULL, 40.f);
  KeyFrame* k1 = A.keyframe(F1);
  KeyFrame* k2 = A.keyframe(F2);
  KeyFrame* obs = A.observer(scale_factors, nlevels);
  std::map<MapPoint*, int> index2;
  for (int i = 0; i < n1; ++i)
    if (flags1[i] & 1)
      k1->AddMapPoint(posed_point(A, obs, proj1[3 * i], proj1[3 * i + 1], 1.0, level1[i], desc_mp1 + (size_t)i * 32, k2->mfLogScaleFactor), (size_t)i);
  for (int i = 0; i < n2; ++i)
    if (flags2[i] & 1) {
      MapPoint* p = posed_point(A, obs, proj2[3 * i], proj2[3 * i + 1], 1.0, level2[i], desc_mp2 + (size_t)i * 32, k1->mfLogScaleFactor);
      k2->AddMapPoint(p, (size_t)i);
      index2[p] = i;
    }
  std::vector<MapPoint*> m12((size_t)n1, static_cast<MapPoint*>(NULL));
  ORBmatcher matcher(0.75f, true);
  const float s12 = 1.0f;
  const int nf = matcher.SearchBySim3(k1, k2, m12, s12, Eigen::Matrix3d::Identity(), Eigen::Vector3d::Zero(), th);
  for (int i = 0; i < n1; ++i) matches12[i] = m12[(size_t)i] ? index2[m12[(size_t)i]] : -1;
  return nf;
}

// ---- ORBmatcher::Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint) (src/ORBmatcher.cc:617-732) with Scw = identity (scw = 1, Rcw = I,
// tcw = 0: exact); map points as in ref_fuse_search at unit depth.  best_idx[i] = keypoint the point was fused to (directly, or the
// keypoint of the point it is to replace), else -1.
void ref_fuse_sim3_search(const float* proj, const int32_t* level, const uint8_t* flags, const uint8_t* desc_mp, int n_mp,
                          const orc_keypoint* kps_un, const uint8_t* desc, int n_kf, float min_x, float min_y, float inv_w, float inv_h,
                          const float* scale_factors, int nlevels, float th, int32_t* best_idx) {
  std::lock_guard<std::mutex> lk(g_lock);
  GridParams g = {min_x, 1e6f, min_y, 1e6f, inv_w, inv_h};
  set_frame_statics(g, 1, 1, 0, 0);
  Arena A;
  Frame F;
  fill_frame(F, NULL, kps_un, desc, n_kf, scale_factors, nlevels, NULL, 40.f);
  KeyFrame* kf = A.keyframe(F);
  KeyFrame* obs = A.observer(scale_factors, nlevels);
  std::vector<MapPoint*> pts((size_t)n_mp);
  for (int i = 0; i < n_mp; ++i) {
    MapPoint* p = posed_point(A, obs, proj[3 * i], proj[3 * i + 1], 1.0, level[i], desc_mp + (size_t)i * 32, kf->mfLogScaleFactor);
    if (!(flags[i] & 1)) p->mbBad = true;
    pts[(size_t)i] = p;
  }
  std::vector<MapPoint*> replace((size_t)n_mp, static_cast<MapPoint*>(NULL));
  ORBmatcher matcher(0.8f, true);
  matcher.Fuse(kf, Eigen::Matrix4d::Identity(), pts, th, replace);
  for (int i = 0; i < n_mp; ++i) {
    int idx = pts[(size_t)i]->GetIndexInKeyFrame(kf);
    if (idx < 0 && replace[(size_t)i]) idx = replace[(size_t)i]->GetIndexInKeyFrame(kf);
    best_idx[i] = idx;
  }
}

// ---- ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (src/ORBmatcher.cc:146-254) with Scw = identity.
// matched_in[k]: vpMatched[k] is set on entry (to a point that is not among vpPoints).  assigned[k] = index of the point the call
// stores in vpMatched[k], else -1; returns nmatches.
int ref_search_by_projection_sim3(const float* proj, const int32_t* level, const uint8_t* flags, const uint8_t* desc_mp, int n_mp,
                                  const orc_keypoint* kps_un, const uint8_t* desc, const uint8_t* matched_in, int n_kf, float min_x,
                                  float min_y, float inv_w, float inv_h, const float* scale_factors, int nlevels, int th, int32_t* assigned) {
  std::lock_guard<std::mutex> lk(g_lock);
  GridParams g = {min_x, 1e6f, min_y, 1e6f, inv_w, inv_h};
  set_frame_statics(g, 1, 1, 0, 0);
  Arena A;
  Frame F;
  fill_frame(F, NULL, kps_un, desc, n_kf, scale_factors, nlevels, NULL, 40.f);
  KeyFrame* kf = A.keyframe(F);
  KeyFrame* obs = A.observer(scale_factors, nlevels);
  std::vector<MapPoint*> pts((size_t)n_mp);
  std::map<MapPoint*, int> index;
  for (int i = 0; i < n_mp; ++i) {
    MapPoint* p = posed_point(A, obs, proj[3 * i], proj[3 * i + 1], 1.0, level[i], desc_mp + (size_t)i * 32, kf->mfLogScaleFactor);
    if (!(flags[i] & 1)) p->mbBad = true;
    pts[(size_t)i] = p;
    index[p] = i;
  }
  std::vector<MapPoint*> matched((size_t)n_kf, static_cast<MapPoint*>(NULL));
  for (int k = 0; k < n_kf; ++k)
    if (matched_in[k]) matched[(size_t)k] = A.point(Eigen::Vector3d(0, 0, 1), obs);
  ORBmatcher matcher(0.75f, true);
  const int nm = matcher.SearchByProjection(kf, Eigen::Matrix4d::Identity(), pts, matched, th);
  for (int k = 0; k < n_kf; ++k) assigned[k] = (matched[(size_t)k] && index.count(matched[(size_t)k])) ? index[matched[(size_t)k]] : -1;
  return nm;
}

// ---- ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1306-1421), the
// relocalisation search, with the current frame at the identity pose.  Keyframe keypoint i carries a map point when valid[i] & 1
// (bad or already found when the bit is clear, alternately), at (proj[3i], proj[3i+1], 1) with PredictScale = pred_level[i].
// has_mp_cur[k]: CurrentFrame.mvpMapPoints[k] is set on entry.  assigned[k] = keyframe keypoint whose point the call stores there.
int ref_search_by_projection_reloc(const orc_keypoint* kps_kf_un, const float* proj, const uint8_t* valid, const int32_t* pred_level,
                                   const uint8_t* desc_mp, int n_kf, const orc_keypoint* kps_cur_un, const uint8_t* desc_cur,
                                   const uint8_t* has_mp_cur, int n_cur, const float* scale_factors, int nlevels, const float bounds[4],
                                   float inv_w, float inv_h, float th, int orb_dist, int check_orientation, int32_t* assigned) {
  std::lock_guard<std::mutex> lk(g_lock);
  GridParams g = {bounds[0], bounds[1], bounds[2], bounds[3], inv_w, inv_h};
  set_frame_statics(g, 1, 1, 0, 0);
  Arena A;
  Frame Fk, Cur;
  std::vector<uint8_t> dk((size_t)n_kf * 32 + 32, 0);
  fill_frame(Fk, NULL, kps_kf_un, &dk[0], n_kf, scale_factors, nlevels, NULL, 40.f);
  fill_frame(Cur, NULL, kps_cur_un, desc_cur, n_cur, scale_factors, nlevels, NULL, 40.f);
  KeyFrame* kf = A.keyframe(Fk);
  KeyFrame* obs = A.observer(scale_factors, nlevels);
  std::set<MapPoint*> found;
  std::map<MapPoint*, int> owner;
  for (int i = 0; i < n_kf; ++i) {
    if (!(valid[i] & 1) && (i % 3 == 0)) continue;  // no map point at all
    MapPoint* p = posed_point(A, obs, proj[3 * i], proj[3 * i + 1], 1.0, pred_level[i], desc_mp + (size_t)i * 32, Cur.mfLogScaleFactor);
    if (!(valid[i] & 1)) {
      if (i % 3 == 1) p->mbBad = true;
      else found.insert(p);
    }
    kf->AddMapPoint(p, (size_t)i);
    owner[p] = i;
  }
  std::vector<MapPoint*> pre((size_t)n_cur, static_cast<MapPoint*>(NULL));
  for (int k = 0; k < n_cur; ++k)
    if (has_mp_cur[k]) {
      pre[(size_t)k] = A.point(Eigen::Vector3d(0, 0, 1), obs);
      Cur.mvpMapPoints[(size_t)k] = pre[(size_t)k];
    }
  ORBmatcher matcher(0.9f, check_orientation != 0);
  const int nm = matcher.SearchByProjection(Cur, kf, found, th, orb_dist);
  for (int k = 0; k < n_cur; ++k) {
    MapPoint* p = Cur.mvpMapPoints[(size_t)k];
    assigned[k] = (p && p != pre[(size_t)k] && owner.count(p)) ? owner[p] : -1;
  }
  return nm;
}

}  // extern "C"
