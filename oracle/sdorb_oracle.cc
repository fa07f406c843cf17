/*
 * sdorb_oracle.cc -- CPU oracle for the SD-SLAM ORB front-end.
 *
 * TEST INFRASTRUCTURE ONLY (see sdorb_oracle.h).  Nothing here is shipped or measured as the
 * product; libsdorb.so never links this file.
 *
 * What it restates, and where the behaviour comes from (paths relative to /root/reference):
 *   - ORBextractor ctor tables            src/ORBextractor.cc:406-457
 *   - ComputePyramid                      src/ORBextractor.cc:680-700
 *   - ComputeKeyPoints (grid, FAST, quota, retainBest)   src/ORBextractor.cc:466-610
 *   - IC_Angle / computeOrientation       src/ORBextractor.cc:78-102, 459-464
 *   - computeOrbDescriptor                src/ORBextractor.cc:105-143
 *   - operator()                          src/ORBextractor.cc:620-678
 *   - DescriptorDistance                  src/ORBmatcher.cc:1459-1473
 *   - best / second-best update rule      src/ORBmatcher.cc:1239-1265
 * The pixel arithmetic lives in OpenCV (not vendored by the reference); the restatements below follow
 * the published OpenCV 4.13 algorithms (imgproc resize.cpp fixed-point INTER_LINEAR, features2d
 * fast.cpp / fast_score.cpp, smooth fixed-point Gaussian, core mathfuncs fastAtan2,
 * keypoint.cpp retainBest) and are pinned bit-for-bit against Python cv2 4.13.0 in tests/.
 * Also restated (the callers either side of the path, SURVEY.md section 8 f): every search routine of src/ORBmatcher.cc
 * (:43-119, :146-254, :256-357, :359-462, :535-586, :682-708, :734-944, :946-1207, :1209-1304, :1306-1421, :1423-1454),
 * Frame::GetFeaturesInArea / AssignFeaturesToGrid / UndistortKeyPoints / ComputeStereoFromRGBD (src/Frame.cc), MapPoint::
 * ComputeDistinctiveDescriptors (src/MapPoint.cc:225-284) and the ORB-SLAM2-style extractor mode.  The reference holds no tests
 * or vectors for any of them.  PARITY PIN: every function here that restates reference text is compared byte for byte with that
 * text itself -- src/{ORBextractor,ORBmatcher,Frame,KeyFrame,MapPoint,Map}.cc compiled unmodified into oracle/_ref by
 * oracle/ref_build/Makefile on the cv:: / Eigen surface of oracle/ref_compat (tests/test_ref_parity.py, tests/test_ref_matchers.py,
 * tests/tools/ref_sweep.py) -- and additionally against separately written Python restatements (tests/search_cases.py).  The
 * ORB-SLAM2-style mode alone is PARITY UNPINNED against ORB-SLAM2 itself (its source is not in this image).
 *
 * Build: g++ -O2 -std=c++17 -ffp-contract=off (see Makefile).  -ffp-contract=off matters: every
 * fused multiply-add the reference's build performs is written as an explicit fmaf() here.
 */
#include "sdorb_oracle.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

/* ------------------------------------------------------------------ rounding helpers */
/* cvRound(float|double) = SSE cvtss2si / cvtsd2si = round-half-to-even in the default FP mode. */
inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }
inline int cv_floor(double v) {
  int i = (int)v;
  return i - (i > v);
}
inline int cv_ceil(double v) {
  int i = (int)v;
  return i + (i < v);
}
inline short sat_short(float v) {
  int i = cv_round(v);
  return (short)(i < SHRT_MIN ? SHRT_MIN : i > SHRT_MAX ? SHRT_MAX : i);
}

/* ------------------------------------------------------------------ cv::resize INTER_LINEAR, 8UC1 */
/* OpenCV imgproc/resize.cpp: resizeGeneric_ with HResizeLinear<uchar,int,short,2048> and
 * VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>.  (The "scale exactly 2 => INTER_AREA fast"
 * shortcut of cv::resize gives (a+b+c+d+2)>>2, which is what these formulas produce for that case too.) */
void resize_linear_8u(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep) {
  if (sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) return;
  const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
  std::vector<int> xofs(dw), xofs1(dw);
  std::vector<short> alpha(2 * dw);
  for (int dx = 0; dx < dw; dx++) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = cv_floor(fx);
    fx -= sx;
    if (sx < 0) {
      fx = 0;
      sx = 0;
    }
    if (sx >= sw - 1) {
      fx = 0;
      sx = sw - 1;
    }
    xofs[dx] = sx;
    xofs1[dx] = std::min(sx + 1, sw - 1);
    alpha[2 * dx] = sat_short((1.f - fx) * 2048.f);
    alpha[2 * dx + 1] = sat_short(fx * 2048.f);
  }
  std::vector<int> row0(dw), row1(dw);
  int cached0 = -1, cached1 = -1;
  auto hresize = [&](int sy, std::vector<int>& out) {
    const uint8_t* S = src + (size_t)sy * sstep;
    for (int dx = 0; dx < dw; dx++) out[dx] = S[xofs[dx]] * alpha[2 * dx] + S[xofs1[dx]] * alpha[2 * dx + 1];
  };
  auto clip = [](int x, int a, int b) { return x >= a ? (x < b ? x : b - 1) : a; };
  for (int dy = 0; dy < dh; dy++) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = cv_floor(fy);
    fy -= sy;
    const int b0 = sat_short((1.f - fy) * 2048.f), b1 = sat_short(fy * 2048.f);
    const int sy0 = clip(sy, 0, sh), sy1 = clip(sy + 1, 0, sh);
    if (sy0 == cached1) {
      std::swap(row0, row1);
      std::swap(cached0, cached1);
    }
    if (sy0 != cached0) {
      hresize(sy0, row0);
      cached0 = sy0;
    }
    if (sy1 == cached0) {
      row1 = row0;
      cached1 = sy1;
    } else if (sy1 != cached1) {
      hresize(sy1, row1);
      cached1 = sy1;
    }
    uint8_t* D = dst + (size_t)dy * dstep;
    for (int x = 0; x < dw; x++) {
      int v = (((b0 * (row0[x] >> 4)) >> 16) + ((b1 * (row1[x] >> 4)) >> 16) + 2) >> 2;
      D[x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
  }
}

/* cv::borderInterpolate(p, len, BORDER_REFLECT_101) */
inline int reflect101(int p, int len) {
  if ((unsigned)p < (unsigned)len) return p;
  if (len == 1) return 0;
  do {
    if (p < 0)
      p = -p;
    else
      p = 2 * (len - 1) - p;
  } while ((unsigned)p >= (unsigned)len);
  return p;
}

void copy_make_border_reflect101(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep, int b) {
  for (int y = -b; y < h + b; y++) {
    const uint8_t* S = src + (size_t)reflect101(y, h) * sstep;
    uint8_t* D = dst + (size_t)(y + b) * dstep;
    for (int x = -b; x < w + b; x++) D[x + b] = S[reflect101(x, w)];
  }
}

/* ------------------------------------------------------------------ cv::FAST TYPE_9_16 */
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* features2d/fast_score.cpp cornerScore<16>: largest threshold for which the pixel stays a corner. */
inline int corner_score16(const uint8_t* p, const int* ofs, int threshold) {
  int d[25];
  const int v = p[0];
  for (int k = 0; k < 25; k++) d[k] = v - p[ofs[k & 15]];
  int a0 = threshold;
  for (int k = 0; k < 16; k += 2) {
    int a = std::min(d[k + 1], d[k + 2]);
    a = std::min(a, d[k + 3]);
    if (a <= a0) continue;
    a = std::min(a, d[k + 4]);
    a = std::min(a, d[k + 5]);
    a = std::min(a, d[k + 6]);
    a = std::min(a, d[k + 7]);
    a = std::min(a, d[k + 8]);
    a0 = std::max(a0, std::min(a, d[k]));
    a0 = std::max(a0, std::min(a, d[k + 9]));
  }
  int b0 = -a0;
  for (int k = 0; k < 16; k += 2) {
    int b = std::max(d[k + 1], d[k + 2]);
    b = std::max(b, d[k + 3]);
    b = std::max(b, d[k + 4]);
    b = std::max(b, d[k + 5]);
    if (b >= b0) continue;
    b = std::max(b, d[k + 6]);
    b = std::max(b, d[k + 7]);
    b = std::max(b, d[k + 8]);
    b0 = std::min(b0, std::max(b, d[k]));
    b0 = std::min(b0, std::max(b, d[k + 9]));
  }
  return -b0 - 1;
}

/* FAST-9 segment test of features2d/fast.cpp (FAST_t<16>): 9 contiguous ring pixels all darker than
 * v-th or all brighter than v+th. */
inline bool is_corner16(const uint8_t* p, const int* ofs, int threshold) {
  const int v = p[0];
  const int lo = v - threshold, hi = v + threshold;
  int cd = 0, cb = 0;
  for (int k = 0; k < 25; k++) {
    const int x = p[ofs[k & 15]];
    if (x < lo) {
      if (++cd > 8) return true;
    } else
      cd = 0;
    if (x > hi) {
      if (++cb > 8) return true;
    } else
      cb = 0;
  }
  return false;
}

void fast_score_map(const uint8_t* img, int w, int h, size_t step, int th, uint8_t* score, size_t sstep) {
  th = std::min(std::max(th, 0), 255);
  for (int y = 0; y < h; y++) memset(score + (size_t)y * sstep, 0, (size_t)std::max(w, 0));
  int ofs[16];
  for (int k = 0; k < 16; k++) ofs[k] = kRingDy[k] * (int)step + kRingDx[k];
  /* fast.cpp threshold_tab: 1 = darker than v-th, 2 = brighter than v+th; the antipodal-pair pretest
   * rejects most pixels before the full segment test (same decisions, fewer loads). */
  uint8_t tab[512];
  for (int i = -255; i <= 255; i++) tab[i + 255] = (uint8_t)(i < -th ? 1 : i > th ? 2 : 0);
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      const uint8_t* p = img + (size_t)y * step + x;
      const uint8_t* t = tab - p[0] + 255;
      int d = t[p[ofs[0]]] | t[p[ofs[8]]];
      if (d == 0) continue;
      d &= t[p[ofs[2]]] | t[p[ofs[10]]];
      d &= t[p[ofs[4]]] | t[p[ofs[12]]];
      d &= t[p[ofs[6]]] | t[p[ofs[14]]];
      if (d == 0) continue;
      d &= t[p[ofs[1]]] | t[p[ofs[9]]];
      d &= t[p[ofs[3]]] | t[p[ofs[11]]];
      d &= t[p[ofs[5]]] | t[p[ofs[13]]];
      d &= t[p[ofs[7]]] | t[p[ofs[15]]];
      if (d == 0) continue;
      if (is_corner16(p, ofs, th)) score[(size_t)y * sstep + x] = (uint8_t)corner_score16(p, ofs, th);
    }
}

void fast_detect(const uint8_t* img, int w, int h, size_t step, int th, bool nonmax, std::vector<orc_keypoint>& out) {
  out.clear();
  if (w < 7 || h < 7) return;
  std::vector<uint8_t> score((size_t)w * h);
  fast_score_map(img, w, h, step, th, score.data(), (size_t)w);
  /* corner flags must be tracked separately from the score: a corner whose score is 0 (only possible
   * for th == 0) is still a corner, and never survives the strict '>' non-max test. */
  th = std::min(std::max(th, 0), 255);
  int ofs[16];
  for (int k = 0; k < 16; k++) ofs[k] = kRingDy[k] * (int)step + kRingDx[k];
  for (int y = 3; y < h - 3; y++) {
    const uint8_t* pr = &score[(size_t)(y - 1) * w];
    const uint8_t* cr = &score[(size_t)y * w];
    const uint8_t* nr = &score[(size_t)(y + 1) * w];
    for (int x = 3; x < w - 3; x++) {
      const int s = cr[x];
      bool keep;
      if (s == 0) {
        if (nonmax || !is_corner16(img + (size_t)y * step + x, ofs, th)) continue;
        keep = true;
      } else {
        keep = !nonmax || (s > cr[x + 1] && s > cr[x - 1] && s > pr[x - 1] && s > pr[x] && s > pr[x + 1] &&
                           s > nr[x - 1] && s > nr[x] && s > nr[x + 1]);
      }
      /* without non-max suppression OpenCV never computes the score: response stays 0 */
      if (keep) out.push_back(orc_keypoint{(float)x, (float)y, 7.f, -1.f, nonmax ? (float)s : 0.f, 0, -1});
    }
  }
}

/* ------------------------------------------------------------------ KeyPointsFilter::retainBest */
struct ResponseGreater {
  bool operator()(const orc_keypoint& a, const orc_keypoint& b) const { return a.response > b.response; }
};
void retain_best(std::vector<orc_keypoint>& v, int n) {
  if (n >= 0 && v.size() > (size_t)n) {
    if (n == 0) {
      v.clear();
      return;
    }
    std::nth_element(v.begin(), v.begin() + n - 1, v.end(), ResponseGreater());
    const float ambiguous = v[n - 1].response;
    auto new_end = std::partition(v.begin() + n, v.end(), [ambiguous](const orc_keypoint& k) { return k.response >= ambiguous; });
    v.resize(new_end - v.begin());
  }
}

/* ------------------------------------------------------------------ cv::fastAtan2 (scalar, degrees) */
float fast_atan2(float y, float x) {
  static const float scale = (float)(180 / 3.1415926535897932384626433832795);
  static const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                     p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

/* ------------------------------------------------------------------ cv::GaussianBlur 7x7 sigma 2, 8U */
/* OpenCV 4.x 8U path: fixed-point separable filter, 8.8 kernel {18,34,48,56,48,34,18}/256,
 * horizontal pass kept in 16 bits, vertical pass in 32 bits, one rounding at the end. */
void gaussian_blur_7x7_s2(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep) {
  static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
  if (w <= 0 || h <= 0) return;
  std::vector<uint32_t> hbuf((size_t)w * h);
  for (int y = 0; y < h; y++) {
    const uint8_t* S = src + (size_t)y * sstep;
    for (int x = 0; x < w; x++) {
      uint32_t acc = 0;
      for (int i = 0; i < 7; i++) acc += k[i] * S[reflect101(x + i - 3, w)];
      hbuf[(size_t)y * w + x] = acc;
    }
  }
  for (int y = 0; y < h; y++) {
    uint8_t* D = dst + (size_t)y * dstep;
    const uint32_t* r[7];
    for (int j = 0; j < 7; j++) r[j] = &hbuf[(size_t)reflect101(y + j - 3, h) * w];
    for (int x = 0; x < w; x++) {
      uint32_t acc = 0;
      for (int j = 0; j < 7; j++) acc += k[j] * r[j][x];
      D[x] = (uint8_t)((acc + 32768u) >> 16);
    }
  }
}

/* ------------------------------------------------------------------ glibc sincosf, restated */
/* glibc 2.39 sysdeps/ieee754/flt-32/s_sincosf.c (the ARM optimized-routines algorithm): double
 * polynomial after a fast range reduction by pi/2.  Only the |x| < 120 paths are needed: the argument
 * is angle*pi/180 with angle in [0, 360].  Written with plain double mul/add (no contraction). */
struct SinCosTab {
  double sign[4];
  double hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3;
};
const SinCosTab kSinCos[2] = {
    {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2,
     0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16, -0x1.555545995a603p-3,
     0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
    {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2,
     -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3,
     0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};

inline uint32_t abstop12(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  return (u >> 20) & 0x7ff;
}

inline void sincosf_poly(double x, double x2, const SinCosTab* p, int n, float* sinp, float* cosp) {
  const double x4 = x2 * x2;
  const double x3 = x2 * x;
  const double c2 = p->c3 + x2 * p->c4;
  const double s1 = p->s2 + x2 * p->s3;
  float* tmp = (n & 1) ? cosp : sinp;
  cosp = (n & 1) ? sinp : cosp;
  sinp = tmp;
  const double c1 = p->c0 + x2 * p->c1;
  const double x5 = x3 * x2;
  const double x6 = x4 * x2;
  const double s = x + x3 * p->s1;
  const double c = c1 + x4 * p->c2;
  *sinp = (float)(s + x5 * s1);
  *cosp = (float)(c + x6 * c2);
}

void sincosf_restated(float y, float* sinp, float* cosp) {
  double x = y;
  const SinCosTab* p = &kSinCos[0];
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) { /* |y| < pi/4 */
    const double x2 = x * x;
    if (abstop12(y) < abstop12(0x1p-12f)) {
      *sinp = y;
      *cosp = 1.0f;
      return;
    }
    sincosf_poly(x, x2, p, 0, sinp, cosp);
  } else if (abstop12(y) < abstop12(120.0f)) {
    const double r = x * p->hpi_inv;
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = x - n * p->hpi;
    const double s = p->sign[n & 3];
    if (n & 2) p = &kSinCos[1];
    sincosf_poly(x * s, x * x, p, n, sinp, cosp);
  } else {
    sincosf(y, sinp, cosp); /* outside the range the extractor can produce */
  }
}

/* ------------------------------------------------------------------ the extractor */
const int PATCH_SIZE = 31;
const int HALF_PATCH_SIZE = 15;
const int EDGE_THRESHOLD = 19;

const int8_t kPattern[1024] = {
#include "orb_pattern.inc"
};

}  // namespace

struct orc_extractor {
  int nfeatures;
  double scaleFactor; /* the member is a double initialised from a float (src/ORBextractor.h:78) */
  int nlevels;
  int thFAST;
  int iniThFAST = -1, minThFAST = -1; /* >= 0: the ORB-SLAM2-style mode (row f1), see compute_keypoints_octree */
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
  std::vector<int> mnFeaturesPerLevel, umax;
};

namespace {

/* src/ORBextractor.cc:406-457 */
void build_tables(orc_extractor& e) {
  const int nlevels = e.nlevels;
  e.mvScaleFactor.resize(nlevels);
  e.mvLevelSigma2.resize(nlevels);
  e.mvScaleFactor[0] = 1.0f;
  e.mvLevelSigma2[0] = 1.0f;
  for (int i = 1; i < nlevels; i++) {
    e.mvScaleFactor[i] = (float)(e.mvScaleFactor[i - 1] * e.scaleFactor);
    e.mvLevelSigma2[i] = e.mvScaleFactor[i] * e.mvScaleFactor[i];
  }
  e.mvInvScaleFactor.resize(nlevels);
  e.mvInvLevelSigma2.resize(nlevels);
  for (int i = 0; i < nlevels; i++) {
    e.mvInvScaleFactor[i] = 1.0f / e.mvScaleFactor[i];
    e.mvInvLevelSigma2[i] = 1.0f / e.mvLevelSigma2[i];
  }
  e.mnFeaturesPerLevel.resize(nlevels);
  float factor = (float)(1.0f / e.scaleFactor);
  float nDesiredFeaturesPerScale = e.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
  int sumFeatures = 0;
  for (int level = 0; level < nlevels - 1; level++) {
    e.mnFeaturesPerLevel[level] = cv_round(nDesiredFeaturesPerScale);
    sumFeatures += e.mnFeaturesPerLevel[level];
    nDesiredFeaturesPerScale *= factor;
  }
  e.mnFeaturesPerLevel[nlevels - 1] = std::max(e.nfeatures - sumFeatures, 0);

  e.umax.resize(HALF_PATCH_SIZE + 1);
  int v, v0, vmax = cv_floor(HALF_PATCH_SIZE * sqrt(2.f) / 2 + 1);
  int vmin = cv_ceil(HALF_PATCH_SIZE * sqrt(2.f) / 2);
  const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
  for (v = 0; v <= vmax; ++v) e.umax[v] = cv_round(sqrt(hp2 - v * v));
  for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
    while (e.umax[v0] == e.umax[v0 + 1]) ++v0;
    e.umax[v] = v0;
    ++v0;
  }
}

struct Image {
  std::vector<uint8_t> buf; /* padded by EDGE_THRESHOLD on every side, like the reference's temp Mat */
  int w = 0, h = 0;
  size_t step = 0;
  uint8_t* inner() { return buf.data() + EDGE_THRESHOLD * step + EDGE_THRESHOLD; }
  const uint8_t* inner() const { return buf.data() + EDGE_THRESHOLD * step + EDGE_THRESHOLD; }
};

void level_size(const orc_extractor& e, int level, int w0, int h0, int* w, int* h) {
  const float scale = e.mvInvScaleFactor[level];
  *w = cv_round((float)w0 * scale);
  *h = cv_round((float)h0 * scale);
}

/* src/ORBextractor.cc:680-700 */
/* Returns false where cv::resize would throw (a level of zero width or height). */
bool compute_pyramid(const orc_extractor& e, const uint8_t* img, int w0, int h0, size_t step0, std::vector<Image>& pyr) {
  pyr.resize(e.nlevels);
  for (int level = 0; level < e.nlevels; ++level) {
    Image& L = pyr[level];
    level_size(e, level, w0, h0, &L.w, &L.h);
    L.step = (size_t)std::max(L.w, 0) + 2 * EDGE_THRESHOLD;
    L.buf.assign(L.step * ((size_t)std::max(L.h, 0) + 2 * EDGE_THRESHOLD), 0);
    if (L.w <= 0 || L.h <= 0) return false;
    if (level != 0) {
      const Image& P = pyr[level - 1];
      resize_linear_8u(P.inner(), P.w, P.h, P.step, L.inner(), L.w, L.h, L.step);
      copy_make_border_reflect101(L.inner(), L.w, L.h, L.step, L.buf.data(), L.step, EDGE_THRESHOLD);
    } else {
      copy_make_border_reflect101(img, w0, h0, step0, L.buf.data(), L.step, EDGE_THRESHOLD);
    }
  }
  return true;
}

/* src/ORBextractor.cc:78-102 */
float ic_angle(const uint8_t* center, int step, const std::vector<int>& u_max) {
  int m_01 = 0, m_10 = 0;
  for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
  for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
    int v_sum = 0;
    const int d = u_max[v];
    for (int u = -d; u <= d; ++u) {
      const int val_plus = center[u + v * step], val_minus = center[u - v * step];
      v_sum += (val_plus - val_minus);
      m_10 += u * (val_plus + val_minus);
    }
    m_01 += v * v_sum;
  }
  return fast_atan2((float)m_01, (float)m_10);
}

/* The reference evaluates  cvRound(px*b + py*a)  and  cvRound(px*a - py*b)  in float
 * (src/ORBextractor.cc:115-117).  Built with its own flags (-O3 -march=native, CMakeLists.txt:39-40)
 * on an FMA host g++ contracts these to fma(px, b, py*a) and fma(px, a, -(py*b)); that form is frozen
 * here (tests/test_oracle_primitives.py compiles the verbatim expression with those flags and compares). */
inline void pattern_rotate(int px, int py, float a, float b, int* drow, int* dcol) {
  const float fx = (float)px, fy = (float)py;
  *drow = cv_round(fmaf(fx, b, fy * a));
  *dcol = cv_round(fmaf(fx, a, -(fy * b)));
}

const float kFactorPI = (float)(3.1415926535897932384626433832795 / 180.f);

/* src/ORBextractor.cc:105-143 */
void orb_descriptor(const orc_keypoint& kpt, const uint8_t* img, int step, uint8_t* desc) {
  const float angle = kpt.angle * kFactorPI;
  float a, b;
  sincosf(angle, &b, &a); /* a = cos, b = sin: g++ merges the reference's cos()/sin() pair into sincosf */
  const uint8_t* center = img + (ptrdiff_t)cv_round(kpt.y) * step + cv_round(kpt.x);
  const int8_t* pat = kPattern;
  for (int i = 0; i < 32; ++i, pat += 32) {
    int val = 0;
    for (int bit = 0; bit < 8; ++bit) {
      int r0, c0, r1, c1;
      pattern_rotate(pat[4 * bit + 0], pat[4 * bit + 1], a, b, &r0, &c0);
      pattern_rotate(pat[4 * bit + 2], pat[4 * bit + 3], a, b, &r1, &c1);
      const int t0 = center[r0 * step + c0], t1 = center[r1 * step + c1];
      val |= (t0 < t1) << bit;
    }
    desc[i] = (uint8_t)val;
  }
}

struct LevelGrid {
  int nDesired, levelCols, levelRows, W, H, cellW, cellH, nCells, nfeaturesCell, scaledPatchSize;
};

/* src/ORBextractor.cc:469-488, 580.  levelCols==0 makes the reference divide by zero and convert inf to
 * int (undefined, but harmless: its loops then run zero times); such a level simply yields nothing. */
LevelGrid level_grid(const orc_extractor& e, int level, int w0, int h0, int lw, int lh) {
  LevelGrid g{};
  const float imageRatio = (float)w0 / h0;
  g.nDesired = e.mnFeaturesPerLevel[level];
  g.levelCols = (int)std::sqrt((float)g.nDesired / (5 * imageRatio));
  g.levelRows = (int)(imageRatio * g.levelCols);
  g.W = (lw - EDGE_THRESHOLD) - EDGE_THRESHOLD;
  g.H = (lh - EDGE_THRESHOLD) - EDGE_THRESHOLD;
  g.nCells = g.levelRows * g.levelCols;
  if (g.levelCols > 0 && g.levelRows > 0) {
    g.cellW = (int)std::ceil((float)g.W / g.levelCols);
    g.cellH = (int)std::ceil((float)g.H / g.levelRows);
    g.nfeaturesCell = (int)std::ceil((float)g.nDesired / g.nCells);
  }
  g.scaledPatchSize = (int)(PATCH_SIZE * e.mvScaleFactor[level]);
  return g;
}

/* src/ORBextractor.cc:466-610.  Returns false when a cell ROI leaves the level image (cv::Exception). */
bool compute_keypoints(const orc_extractor& e, const std::vector<Image>& pyr, int w0, int h0,
                       std::vector<std::vector<orc_keypoint>>& all, orc_dump* dump, int* raw_cell_pos, int* raw_kp_pos) {
  all.assign(e.nlevels, {});
  for (int level = 0; level < e.nlevels; ++level) {
    const Image& L = pyr[level];
    const LevelGrid g = level_grid(e, level, w0, h0, L.w, L.h);
    const int levelRows = g.levelRows, levelCols = g.levelCols;
    if (levelRows <= 0 || levelCols <= 0) continue;
    const int nCells = g.nCells, nfeaturesCell = g.nfeaturesCell, cellW = g.cellW, cellH = g.cellH;
    const int minBorderX = EDGE_THRESHOLD, minBorderY = EDGE_THRESHOLD;
    const int maxBorderX = L.w - EDGE_THRESHOLD, maxBorderY = L.h - EDGE_THRESHOLD;

    std::vector<std::vector<orc_keypoint>> cellKeyPoints((size_t)nCells);
    std::vector<int> nToRetain(nCells, 0), nTotal(nCells, 0);
    std::vector<char> bNoMore(nCells, 0);
    std::vector<int> iniXCol(levelCols), iniYRow(levelRows);
    int nNoMore = 0, nToDistribute = 0;

    float hY = (float)(cellH + 6);
    for (int i = 0; i < levelRows; i++) {
      const float iniY = (float)(minBorderY + i * cellH - 3);
      iniYRow[i] = (int)iniY;
      if (i == levelRows - 1) {
        hY = maxBorderY + 3 - iniY;
        if (hY <= 0) continue;
      }
      float hX = (float)(cellW + 6);
      for (int j = 0; j < levelCols; j++) {
        float iniX;
        if (i == 0) {
          iniX = (float)(minBorderX + j * cellW - 3);
          iniXCol[j] = (int)iniX;
        } else {
          iniX = (float)iniXCol[j];
        }
        if (j == levelCols - 1) {
          hX = maxBorderX + 3 - iniX;
          if (hX <= 0) continue;
        }
        const int y0 = (int)iniY, y1 = (int)(iniY + hY), x0 = (int)iniX, x1 = (int)(iniX + hX);
        /* Mat::rowRange / colRange assertions */
        if (!(0 <= y0 && y0 <= y1 && y1 <= L.h && 0 <= x0 && x0 <= x1 && x1 <= L.w)) return false;
        /* A non-empty cell whose detectable rectangle passes maxBorder (possible only when
         * (levelRows-1)*cellH > H, i.e. a level a few pixels larger than the border) yields keypoints
         * whose rotated pattern is read outside the blurred level image: undefined behaviour in the
         * reference, so no result is defined.  Reported as a geometry error as well. */
        if (x1 - x0 >= 7 && y1 - y0 >= 7 && (x1 - 3 > maxBorderX || y1 - 3 > maxBorderY)) return false;
        std::vector<orc_keypoint>& cell = cellKeyPoints[(size_t)i * levelCols + j];
        fast_detect(L.inner() + (size_t)y0 * L.step + x0, x1 - x0, y1 - y0, L.step, e.thFAST, true, cell);
        const int nKeys = (int)cell.size();
        nTotal[i * levelCols + j] = nKeys;
        if (nKeys > nfeaturesCell) {
          nToRetain[i * levelCols + j] = nfeaturesCell;
          bNoMore[i * levelCols + j] = 0;
        } else {
          nToRetain[i * levelCols + j] = nKeys;
          nToDistribute += nfeaturesCell - nKeys;
          bNoMore[i * levelCols + j] = 1;
          nNoMore++;
        }
      }
    }

    const int cell_base = *raw_cell_pos;
    *raw_cell_pos += nCells;
    if (dump) {
      for (int c = 0; c < nCells; c++) {
        if (dump->raw_cell_count && cell_base + c < dump->raw_cell_cap) dump->raw_cell_count[cell_base + c] = nTotal[c];
        for (const orc_keypoint& k : cellKeyPoints[c]) {
          if (dump->raw_kps && *raw_kp_pos < dump->raw_kps_cap) dump->raw_kps[*raw_kp_pos] = k;
          ++*raw_kp_pos;
        }
      }
    }

    while (nToDistribute > 0 && nNoMore < nCells) {
      const int nNewFeaturesCell = nfeaturesCell + (int)std::ceil((float)nToDistribute / (nCells - nNoMore));
      nToDistribute = 0;
      for (int c = 0; c < nCells; c++) {
        if (!bNoMore[c]) {
          if (nTotal[c] > nNewFeaturesCell) {
            nToRetain[c] = nNewFeaturesCell;
            bNoMore[c] = 0;
          } else {
            nToRetain[c] = nTotal[c];
            nToDistribute += nNewFeaturesCell - nTotal[c];
            bNoMore[c] = 1;
            nNoMore++;
          }
        }
      }
    }
    if (dump && dump->n_to_retain)
      for (int c = 0; c < nCells; c++)
        if (cell_base + c < dump->n_to_retain_cap) dump->n_to_retain[cell_base + c] = nToRetain[c];

    std::vector<orc_keypoint>& keypoints = all[level];
    keypoints.reserve((size_t)std::max(g.nDesired, 0) * 2);
    for (int i = 0; i < levelRows; i++)
      for (int j = 0; j < levelCols; j++) {
        std::vector<orc_keypoint>& keysCell = cellKeyPoints[(size_t)i * levelCols + j];
        retain_best(keysCell, nToRetain[i * levelCols + j]);
        if ((int)keysCell.size() > nToRetain[i * levelCols + j]) keysCell.resize(nToRetain[i * levelCols + j]);
        for (orc_keypoint& k : keysCell) {
          k.x += iniXCol[j];
          k.y += iniYRow[i];
          k.octave = level;
          k.size = (float)g.scaledPatchSize;
          keypoints.push_back(k);
        }
      }
    if ((int)keypoints.size() > g.nDesired) {
      retain_best(keypoints, g.nDesired);
      keypoints.resize(g.nDesired);
    }
  }
  for (int level = 0; level < e.nlevels; ++level) {
    const Image& L = pyr[level];
    for (orc_keypoint& k : all[level])
      k.angle = ic_angle(L.inner() + (ptrdiff_t)cv_round(k.y) * L.step + cv_round(k.x), (int)L.step, e.umax);
  }
  return true;
}

/* ------------------------------------------------------------------ ORB-SLAM2-style mode
 * iniThFAST / minThFAST fallback on 30-pixel cells + DistributeOctTree.  NOT part of /root/reference:
 * this restates the public ORB-SLAM2 algorithm (raulmur/ORB_SLAM2, src/ORBextractor.cc: ExtractorNode::DivideNode,
 * ORBextractor::DistributeOctTree, ORBextractor::ComputeKeyPointsOctTree) from its published form; the source is not in
 * this image and there is no network, so this mode is PARITY UNPINNED against ORB-SLAM2 itself (the FAST / blur /
 * descriptor primitives it shares with the reference mode are pinned as above, and a second, independently written Python
 * restatement in tests/cv2_pipeline.py must agree).
 * One point of ORB-SLAM2 is not a function of its inputs: DistributeOctTree sorts (size, ExtractorNode*) pairs, so nodes of
 * equal size are ordered by the heap addresses of std::list nodes.  Frozen here as the bump-allocator model: a node created
 * later has the higher address. */
struct ExtractorNode {
  int ULx, ULy, URx, BRy; /* UL = (ULx, ULy), UR = (URx, ULy), BL = (ULx, BRy), BR = (URx, BRy) */
  std::vector<orc_keypoint> vKeys;
  bool bNoMore = false;
  int serial = 0; /* creation order; stands in for the node's address in the (size, pointer) sort */
  void DivideNode(ExtractorNode& n1, ExtractorNode& n2, ExtractorNode& n3, ExtractorNode& n4) const {
    const int halfX = (int)std::ceil((float)(URx - ULx) / 2);
    const int halfY = (int)std::ceil((float)(BRy - ULy) / 2);
    n1.ULx = ULx, n1.ULy = ULy, n1.URx = ULx + halfX, n1.BRy = ULy + halfY;
    n2.ULx = n1.URx, n2.ULy = ULy, n2.URx = URx, n2.BRy = ULy + halfY;
    n3.ULx = ULx, n3.ULy = n1.BRy, n3.URx = n1.URx, n3.BRy = BRy;
    n4.ULx = n3.URx, n4.ULy = n3.ULy, n4.URx = URx, n4.BRy = BRy;
    for (const orc_keypoint& kp : vKeys) {
      if (kp.x < n1.URx) {
        if (kp.y < n1.BRy)
          n1.vKeys.push_back(kp);
        else
          n3.vKeys.push_back(kp);
      } else if (kp.y < n1.BRy)
        n2.vKeys.push_back(kp);
      else
        n4.vKeys.push_back(kp);
    }
    n1.bNoMore = n1.vKeys.size() == 1;
    n2.bNoMore = n2.vKeys.size() == 1;
    n3.bNoMore = n3.vKeys.size() == 1;
    n4.bNoMore = n4.vKeys.size() == 1;
  }
};
}  // namespace
#include <list>
namespace {
typedef std::list<ExtractorNode> NodeList;

std::vector<orc_keypoint> distribute_oct_tree(const std::vector<orc_keypoint>& vToDistributeKeys, int minX, int maxX, int minY,
                                              int maxY, int N) {
  const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
  const float hX = (float)(maxX - minX) / nIni;
  NodeList lNodes;
  std::vector<ExtractorNode*> vpIniNodes((size_t)nIni);
  int serial = 0;
  for (int i = 0; i < nIni; i++) {
    ExtractorNode ni;
    ni.ULx = (int)(hX * (float)i);
    ni.URx = (int)(hX * (float)(i + 1));
    ni.ULy = 0;
    ni.BRy = maxY - minY;
    ni.serial = serial++;
    lNodes.push_back(ni);
    vpIniNodes[i] = &lNodes.back();
  }
  for (const orc_keypoint& kp : vToDistributeKeys) vpIniNodes[(size_t)(kp.x / hX)]->vKeys.push_back(kp);
  for (NodeList::iterator lit = lNodes.begin(); lit != lNodes.end();) {
    if (lit->vKeys.size() == 1) {
      lit->bNoMore = true;
      ++lit;
    } else if (lit->vKeys.empty())
      lit = lNodes.erase(lit);
    else
      ++lit;
  }
  bool bFinish = false;
  typedef std::pair<int, std::pair<int, NodeList::iterator>> SizeAndNode; /* (size, (serial, node)) */
  std::vector<SizeAndNode> vSizeAndPointerToNode;
  auto add_child = [&](ExtractorNode& n, int* nToExpand) {
    if (n.vKeys.size() > 0) {
      n.serial = serial++;
      lNodes.push_front(n);
      if (n.vKeys.size() > 1) {
        if (nToExpand) ++*nToExpand;
        vSizeAndPointerToNode.push_back(std::make_pair((int)n.vKeys.size(), std::make_pair(n.serial, lNodes.begin())));
      }
    }
  };
  while (!bFinish) {
    int prevSize = (int)lNodes.size();
    NodeList::iterator lit = lNodes.begin();
    int nToExpand = 0;
    vSizeAndPointerToNode.clear();
    while (lit != lNodes.end()) {
      if (lit->bNoMore) {
        ++lit;
        continue;
      }
      ExtractorNode n1, n2, n3, n4;
      lit->DivideNode(n1, n2, n3, n4);
      add_child(n1, &nToExpand);
      add_child(n2, &nToExpand);
      add_child(n3, &nToExpand);
      add_child(n4, &nToExpand);
      lit = lNodes.erase(lit);
    }
    if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) {
      bFinish = true;
    } else if ((int)lNodes.size() + nToExpand * 3 > N) {
      while (!bFinish) {
        prevSize = (int)lNodes.size();
        std::vector<SizeAndNode> vPrev = vSizeAndPointerToNode;
        vSizeAndPointerToNode.clear();
        std::sort(vPrev.begin(), vPrev.end(), [](const SizeAndNode& a, const SizeAndNode& b) {
          return a.first < b.first || (a.first == b.first && a.second.first < b.second.first);
        });
        for (int j = (int)vPrev.size() - 1; j >= 0; j--) {
          ExtractorNode n1, n2, n3, n4;
          vPrev[j].second.second->DivideNode(n1, n2, n3, n4);
          add_child(n1, nullptr);
          add_child(n2, nullptr);
          add_child(n3, nullptr);
          add_child(n4, nullptr);
          lNodes.erase(vPrev[j].second.second);
          if ((int)lNodes.size() >= N) break;
        }
        if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) bFinish = true;
      }
    }
  }
  std::vector<orc_keypoint> vResultKeys;
  for (const ExtractorNode& n : lNodes) {
    const orc_keypoint* pKP = &n.vKeys[0];
    float maxResponse = pKP->response;
    for (size_t k = 1; k < n.vKeys.size(); k++)
      if (n.vKeys[k].response > maxResponse) {
        pKP = &n.vKeys[k];
        maxResponse = n.vKeys[k].response;
      }
    vResultKeys.push_back(*pKP);
  }
  return vResultKeys;
}

/* ORBextractor::ComputeKeyPointsOctTree.  Returns false where ORB-SLAM2 itself is undefined (a level too small for one
 * 30-pixel cell, or nIni == 0: divisions by zero). */
bool compute_keypoints_octree(const orc_extractor& e, const std::vector<Image>& pyr, std::vector<std::vector<orc_keypoint>>& all,
                              orc_dump* dump, int* raw_kp_pos) {
  all.assign(e.nlevels, {});
  const float W = 30;
  for (int level = 0; level < e.nlevels; ++level) {
    const Image& L = pyr[level];
    const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
    const int maxBorderX = L.w - EDGE_THRESHOLD + 3, maxBorderY = L.h - EDGE_THRESHOLD + 3;
    const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) return false;
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    if ((int)std::round((float)(maxBorderX - minBorderX) / (maxBorderY - minBorderY)) < 1) return false;
    std::vector<orc_keypoint> vToDistributeKeys;
    for (int i = 0; i < nRows; i++) {
      const float iniY = (float)(minBorderY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBorderY - 3) continue;
      if (maxY > maxBorderY) maxY = (float)maxBorderY;
      for (int j = 0; j < nCols; j++) {
        const float iniX = (float)(minBorderX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBorderX - 6) continue;
        if (maxX > maxBorderX) maxX = (float)maxBorderX;
        const int y0 = (int)iniY, y1 = (int)maxY, x0 = (int)iniX, x1 = (int)maxX;
        std::vector<orc_keypoint> vKeysCell;
        fast_detect(L.inner() + (size_t)y0 * L.step + x0, x1 - x0, y1 - y0, L.step, e.iniThFAST, true, vKeysCell);
        if (vKeysCell.empty())
          fast_detect(L.inner() + (size_t)y0 * L.step + x0, x1 - x0, y1 - y0, L.step, e.minThFAST, true, vKeysCell);
        for (orc_keypoint& k : vKeysCell) {
          k.x += j * wCell;
          k.y += i * hCell;
          vToDistributeKeys.push_back(k);
        }
      }
    }
    if (dump)
      for (const orc_keypoint& k : vToDistributeKeys) {
        if (dump->raw_kps && *raw_kp_pos < dump->raw_kps_cap) {
          dump->raw_kps[*raw_kp_pos] = k;
          dump->raw_kps[*raw_kp_pos].octave = level;
        }
        ++*raw_kp_pos;
      }
    std::vector<orc_keypoint>& keypoints = all[level];
    keypoints = distribute_oct_tree(vToDistributeKeys, minBorderX, maxBorderX, minBorderY, maxBorderY, e.mnFeaturesPerLevel[level]);
    const int scaledPatchSize = (int)(PATCH_SIZE * e.mvScaleFactor[level]);
    for (orc_keypoint& k : keypoints) {
      k.x += minBorderX;
      k.y += minBorderY;
      k.octave = level;
      k.size = (float)scaledPatchSize;
    }
  }
  for (int level = 0; level < e.nlevels; ++level) {
    const Image& L = pyr[level];
    for (orc_keypoint& k : all[level])
      k.angle = ic_angle(L.inner() + (ptrdiff_t)cv_round(k.y) * L.step + cv_round(k.x), (int)L.step, e.umax);
  }
  return true;
}

/* src/ORBextractor.cc:620-678 */
int extract(const orc_extractor& e, const uint8_t* img, int w, int h, size_t step, orc_keypoint* kps, uint8_t* desc,
            int cap, orc_dump* dump) {
  if (!img || w <= 0 || h <= 0) return 0;
  std::vector<Image> pyr;
  if (!compute_pyramid(e, img, w, h, step, pyr)) return -1;
  if (dump && dump->pyramid) {
    uint8_t* o = dump->pyramid;
    for (const Image& L : pyr)
      for (int y = 0; y < L.h; y++, o += L.w) memcpy(o, L.inner() + (size_t)y * L.step, (size_t)L.w);
  }
  std::vector<std::vector<orc_keypoint>> all;
  int raw_cell_pos = 0, raw_kp_pos = 0;
  if (e.minThFAST >= 0) {
    if (!compute_keypoints_octree(e, pyr, all, dump, &raw_kp_pos)) return -1;
  } else if (!compute_keypoints(e, pyr, w, h, all, dump, &raw_cell_pos, &raw_kp_pos))
    return -1;
  if (dump) dump->raw_kps_total = raw_kp_pos;

  int n = 0;
  uint8_t* bo = dump ? dump->blurred : nullptr;
  for (int level = 0; level < e.nlevels; ++level) {
    std::vector<orc_keypoint>& keypoints = all[level];
    const Image& L = pyr[level];
    if (dump && dump->level_count) dump->level_count[level] = (int)keypoints.size();
    std::vector<uint8_t> work;
    if (!keypoints.empty() || bo) {
      work.resize((size_t)std::max(L.w, 0) * std::max(L.h, 0));
      if (L.w > 0 && L.h > 0) gaussian_blur_7x7_s2(L.inner(), L.w, L.h, L.step, work.data(), (size_t)L.w);
      if (bo) {
        memcpy(bo, work.data(), work.size());
        bo += work.size();
      }
    }
    if (keypoints.empty()) continue;
    const float scale = e.mvScaleFactor[level];
    for (orc_keypoint& k : keypoints) {
      uint8_t d[32];
      orb_descriptor(k, work.data(), L.w, d);
      if (level != 0) {
        k.x *= scale;
        k.y *= scale;
      }
      if (n < cap) {
        if (kps) kps[n] = k;
        if (desc) memcpy(desc + (size_t)n * 32, d, 32);
      }
      ++n;
    }
  }
  return n;
}

inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
  /* src/ORBmatcher.cc:1459-1473, verbatim arithmetic */
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t wa, wb;
    memcpy(&wa, a + 4 * i, 4);
    memcpy(&wb, b + 4 * i, 4);
    unsigned int v = wa ^ wb;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

void match_best2(const uint8_t* A, int nA, const uint8_t* B, int nB, float ratio, int th_low, orc_match* out,
                 bool greedy) {
  std::vector<char> matched((size_t)std::max(nB, 0), 0);
  for (int i = 0; i < nA; i++) {
    int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
    for (int j = 0; j < nB; j++) {
      if (greedy && matched[j]) continue;
      const int dist = descriptor_distance(A + (size_t)i * 32, B + (size_t)j * 32);
      if (dist < bestDist1) {
        bestDist2 = bestDist1;
        bestDist1 = dist;
        bestIdx2 = j;
      } else if (dist < bestDist2) {
        bestDist2 = dist;
      }
    }
    int ok = 0;
    if (bestDist1 < th_low && (float)bestDist1 < ratio * (float)bestDist2) {
      ok = 1;
      if (greedy) matched[bestIdx2] = 1;
    }
    out[i] = orc_match{bestIdx2, bestDist1, bestDist2, ok};
  }
}

template <class F>
void parallel_for(int n, int nthreads, F f) {
  nthreads = std::max(1, std::min(nthreads, n));
  if (nthreads == 1) {
    for (int i = 0; i < n; i++) f(i);
    return;
  }
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back([&] {
      for (int i; (i = next.fetch_add(1)) < n;) f(i);
    });
  for (auto& t : th) t.join();
}

}  // namespace

/* ================================================================== C interface */
extern "C" {

void orc_resize_linear_8u(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep) {
  resize_linear_8u(src, sw, sh, sstep, dst, dw, dh, dstep);
}
void orc_copy_make_border_reflect101(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep, int border) {
  copy_make_border_reflect101(src, w, h, sstep, dst, dstep, border);
}
int orc_fast(const uint8_t* img, int w, int h, size_t step, int th, int nonmax, orc_keypoint* out, int cap) {
  std::vector<orc_keypoint> v;
  fast_detect(img, w, h, step, th, nonmax != 0, v);
  for (int i = 0; i < (int)v.size() && i < cap; i++) out[i] = v[i];
  return (int)v.size();
}
void orc_fast_score_map(const uint8_t* img, int w, int h, size_t step, int th, uint8_t* score, size_t score_step) {
  fast_score_map(img, w, h, step, th, score, score_step);
}
void orc_gaussian_blur_7x7_s2(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep) {
  gaussian_blur_7x7_s2(src, w, h, sstep, dst, dstep);
}
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
int orc_retain_best(orc_keypoint* kps, int count, int n) {
  std::vector<orc_keypoint> v(kps, kps + count);
  retain_best(v, n);
  std::copy(v.begin(), v.end(), kps);
  return (int)v.size();
}
int orc_retain_best_idx(const float* response, int count, int n, int32_t* order_out) {
  std::vector<orc_keypoint> v((size_t)count);
  for (int i = 0; i < count; i++) v[i] = orc_keypoint{0, 0, 0, 0, response[i], 0, i};
  retain_best(v, n);
  for (size_t i = 0; i < v.size(); i++) order_out[i] = v[i].class_id;
  return (int)v.size();
}
void orc_sincosf_restated(float x, float* s, float* c) { sincosf_restated(x, s, c); }
uint64_t orc_sincosf_mismatches(uint32_t lo_bits, uint32_t hi_bits, int nthreads) {
  nthreads = std::max(1, nthreads);
  std::atomic<uint64_t> bad(0);
  const uint64_t total = (uint64_t)hi_bits - lo_bits + 1;
  const int chunks = nthreads * 16;
  parallel_for(chunks, nthreads, [&](int c) {
    const uint64_t b = lo_bits + total * c / chunks, e = lo_bits + total * (c + 1) / chunks;
    uint64_t local = 0;
    for (uint64_t u = b; u < e; u++) {
      const uint32_t bits = (uint32_t)u;
      float x, s0, c0, s1, c1;
      memcpy(&x, &bits, 4);
      sincosf(x, &s0, &c0);
      sincosf_restated(x, &s1, &c1);
      local += (memcmp(&s0, &s1, 4) != 0) || (memcmp(&c0, &c1, 4) != 0);
    }
    bad += local;
  });
  return bad.load();
}
void orc_pattern_rotate(int px, int py, float a, float b, int* drow, int* dcol) { pattern_rotate(px, py, a, b, drow, dcol); }

orc_extractor* orc_create(const orc_params* p) {
  if (!p || p->nlevels <= 0 || p->nfeatures < 0) return nullptr;
  orc_extractor* e = new orc_extractor;
  e->nfeatures = p->nfeatures;
  e->scaleFactor = p->scale_factor;
  e->nlevels = p->nlevels;
  e->thFAST = p->th_fast;
  build_tables(*e);
  return e;
}
void orc_destroy(orc_extractor* e) { delete e; }
void orc_set_orbslam2_mode(orc_extractor* e, int ini_th_fast, int min_th_fast) {
  e->iniThFAST = ini_th_fast;
  e->minThFAST = min_th_fast;
}
int orc_distribute_oct_tree(const orc_keypoint* keys, int n, int minX, int maxX, int minY, int maxY, int N, orc_keypoint* out, int cap) {
  const std::vector<orc_keypoint> r = distribute_oct_tree(std::vector<orc_keypoint>(keys, keys + n), minX, maxX, minY, maxY, N);
  for (size_t i = 0; i < r.size() && (int)i < cap; i++) out[i] = r[i];
  return (int)r.size();
}
void orc_get_tables(const orc_extractor* e, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                    int* n_per_level, int* umax) {
  for (int i = 0; i < e->nlevels; i++) {
    if (scale) scale[i] = e->mvScaleFactor[i];
    if (inv_scale) inv_scale[i] = e->mvInvScaleFactor[i];
    if (sigma2) sigma2[i] = e->mvLevelSigma2[i];
    if (inv_sigma2) inv_sigma2[i] = e->mvInvLevelSigma2[i];
    if (n_per_level) n_per_level[i] = e->mnFeaturesPerLevel[i];
  }
  if (umax)
    for (int i = 0; i <= HALF_PATCH_SIZE; i++) umax[i] = e->umax[i];
}
void orc_level_geometry(const orc_extractor* e, int w, int h, orc_level_geom* out) {
  for (int l = 0; l < e->nlevels; l++) {
    int lw, lh;
    level_size(*e, l, w, h, &lw, &lh);
    const LevelGrid g = level_grid(*e, l, w, h, lw, lh);
    out[l] = orc_level_geom{lw, lh, g.nDesired, g.levelCols, g.levelRows, g.cellW, g.cellH, g.nfeaturesCell, g.scaledPatchSize};
  }
}
int orc_extract(const orc_extractor* e, const uint8_t* img, int w, int h, size_t step, orc_keypoint* kps, uint8_t* desc,
                int cap, orc_dump* dump) {
  return extract(*e, img, w, h, step, kps, desc, cap, dump);
}
long orc_extract_many(const orc_extractor* e, const uint8_t* imgs, int nframes, int w, int h, int nthreads,
                      orc_keypoint* kps, uint8_t* desc, int32_t* counts, int cap) {
  std::atomic<long> total(0);
  parallel_for(nframes, nthreads, [&](int f) {
    std::vector<orc_keypoint> k((size_t)cap);
    std::vector<uint8_t> d((size_t)cap * 32);
    const int n = extract(*e, imgs + (size_t)f * w * h, w, h, (size_t)w, k.data(), d.data(), cap, nullptr);
    const int m = std::max(0, std::min(n, cap));
    if (kps) std::copy(k.begin(), k.begin() + m, kps + (size_t)f * cap);
    if (desc) memcpy(desc + (size_t)f * cap * 32, d.data(), (size_t)m * 32);
    if (counts) counts[f] = n;
    total += std::max(n, 0);
  });
  return total.load();
}

int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }
void orc_match_best2(const uint8_t* A, int nA, const uint8_t* B, int nB, float ratio, int th_low, orc_match* out) {
  match_best2(A, nA, B, nB, ratio, th_low, out, false);
}
void orc_match_greedy(const uint8_t* A, int nA, const uint8_t* B, int nB, float ratio, int th_low, orc_match* out) {
  match_best2(A, nA, B, nB, ratio, th_low, out, true);
}
void orc_match_many(const uint8_t* A, const int32_t* nA, const uint8_t* B, const int32_t* nB, int npairs, int strideA_rows,
                    int strideB_rows, float ratio, int th_low, int nthreads, orc_match* out) {
  parallel_for(npairs, nthreads, [&](int p) {
    match_best2(A + (size_t)p * strideA_rows * 32, nA[p], B + (size_t)p * strideB_rows * 32, nB[p], ratio, th_low,
                out + (size_t)p * strideA_rows, false);
  });
}
// ---------------------------------------------------------------------------------------------- guided matchers
// Frame::GetFeaturesInArea, /root/reference/src/Frame.cc:271-321.
static std::vector<int> features_in_area(const orc_keypoint* kps, const orc_frame_grid& G, float x, float y, float r, int minLevel,
                                         int maxLevel) {
  const int COLS = 64, ROWS = 48;
  std::vector<int> vIndices;
  const int nMinCellX = std::max(0, (int)std::floor((x - G.mnMinX - r) * G.mfGridElementWidthInv));
  if (nMinCellX >= COLS) return vIndices;
  const int nMaxCellX = std::min(COLS - 1, (int)std::ceil((x - G.mnMinX + r) * G.mfGridElementWidthInv));
  if (nMaxCellX < 0) return vIndices;
  const int nMinCellY = std::max(0, (int)std::floor((y - G.mnMinY - r) * G.mfGridElementHeightInv));
  if (nMinCellY >= ROWS) return vIndices;
  const int nMaxCellY = std::min(ROWS - 1, (int)std::ceil((y - G.mnMinY + r) * G.mfGridElementHeightInv));
  if (nMaxCellY < 0) return vIndices;
  const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
    for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
      const int c = ix * ROWS + iy;
      for (int j = G.cell_start[c]; j < G.cell_start[c + 1]; j++) {
        const orc_keypoint& kpUn = kps[G.indices[j]];
        if (bCheckLevels) {
          if (kpUn.octave < minLevel) continue;
          if (maxLevel >= 0)
            if (kpUn.octave > maxLevel) continue;
        }
        const float distx = kpUn.x - x;
        const float disty = kpUn.y - y;
        if (std::fabs(distx) < r && std::fabs(disty) < r) vIndices.push_back(G.indices[j]);
      }
    }
  }
  return vIndices;
}
int orc_features_in_area(const orc_keypoint* kps_un, const orc_frame_grid* grid, float x, float y, float r, int minLevel,
                         int maxLevel, int32_t* out) {
  const std::vector<int> v = features_in_area(kps_un, *grid, x, y, r, minLevel, maxLevel);
  std::copy(v.begin(), v.end(), out);
  return (int)v.size();
}
// ORBmatcher::ComputeThreeMaxima, src/ORBmatcher.cc:1423-1454.
void orc_three_maxima(const int32_t* sizes, int L, int* ind1, int* ind2, int* ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = sizes[i];
    if (s > max1) {
      max3 = max2;
      max2 = max1;
      max1 = s;
      *ind3 = *ind2;
      *ind2 = *ind1;
      *ind1 = i;
    } else if (s > max2) {
      max3 = max2;
      max2 = s;
      *ind3 = *ind2;
      *ind2 = i;
    } else if (s > max3) {
      max3 = s;
      *ind3 = i;
    }
  }
  if (max2 < 0.1f * (float)max1) {
    *ind2 = -1;
    *ind3 = -1;
  } else if (max3 < 0.1f * (float)max1) {
    *ind3 = -1;
  }
}
static const int kHistoLength = 30, kThHigh = 100, kThLow = 50;  // src/ORBmatcher.cc:36-38
static int rotation_bin(float a1, float a2) {  // src/ORBmatcher.cc:316-322, 1042-1048
  const float factor = 1.0f / kHistoLength;
  float rot = a1 - a2;
  if (rot < 0.0) rot += 360.0f;
  int bin = (int)std::round(rot * factor);
  if (bin == kHistoLength) bin = 0;
  return bin;
}
// ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:256-357.
int orc_search_for_initialization(const orc_keypoint* k1, const uint8_t* d1s, int n1, const orc_keypoint* k2, const uint8_t* d2s,
                                  int n2, const orc_frame_grid* grid2, float* prev, int windowSize, float mfNNratio,
                                  int mbCheckOrientation, int32_t* vnMatches12) {
  int nmatches = 0;
  std::fill(vnMatches12, vnMatches12 + n1, -1);
  std::vector<std::vector<int>> rotHist(kHistoLength);
  std::vector<int> vMatchedDistance((size_t)n2, INT_MAX), vnMatches21((size_t)n2, -1);
  for (int i1 = 0; i1 < n1; i1++) {
    const int level1 = k1[i1].octave;
    if (level1 > 0) continue;
    const std::vector<int> vIndices2 = features_in_area(k2, *grid2, prev[2 * i1], prev[2 * i1 + 1], (float)windowSize, level1, level1);
    if (vIndices2.empty()) continue;
    const uint8_t* d1 = d1s + (size_t)i1 * 32;
    int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
    for (int i2 : vIndices2) {
      const int dist = descriptor_distance(d1, d2s + (size_t)i2 * 32);
      if (vMatchedDistance[i2] <= dist) continue;
      if (dist < bestDist) {
        bestDist2 = bestDist;
        bestDist = dist;
        bestIdx2 = i2;
      } else if (dist < bestDist2) {
        bestDist2 = dist;
      }
    }
    if (bestDist <= kThLow) {
      if (bestDist < (float)bestDist2 * mfNNratio) {
        if (vnMatches21[bestIdx2] >= 0) {
          vnMatches12[vnMatches21[bestIdx2]] = -1;
          nmatches--;
        }
        vnMatches12[i1] = bestIdx2;
        vnMatches21[bestIdx2] = i1;
        vMatchedDistance[bestIdx2] = bestDist;
        nmatches++;
        if (mbCheckOrientation) rotHist[rotation_bin(k1[i1].angle, k2[bestIdx2].angle)].push_back(i1);
      }
    }
  }
  if (mbCheckOrientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    int32_t sizes[kHistoLength];
    for (int i = 0; i < kHistoLength; i++) sizes[i] = (int)rotHist[i].size();
    orc_three_maxima(sizes, kHistoLength, &ind1, &ind2, &ind3);
    for (int i = 0; i < kHistoLength; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i]) {
        if (vnMatches12[idx1] >= 0) {
          vnMatches12[idx1] = -1;
          nmatches--;
        }
      }
    }
  }
  for (int i1 = 0; i1 < n1; i1++)
    if (vnMatches12[i1] >= 0) {
      prev[2 * i1] = k2[vnMatches12[i1]].x;
      prev[2 * i1 + 1] = k2[vnMatches12[i1]].y;
    }
  return nmatches;
}
// ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono), src/ORBmatcher.cc:946-1075, from
// the projection (u, v, invzc) on.  `ur = u - mbf*invzc` (:1024) is a multiply feeding a subtraction: GCC contracts it
// into one fused multiply-add under the reference's -O3 -march=native (CMakeLists.txt:39-40), frozen here with fmaf.
int orc_search_by_projection(const orc_keypoint* kL, const orc_keypoint* kLun, const float* proj, const uint8_t* flagsL,
                             const uint8_t* descMP, int nL, const orc_keypoint* kC, const uint8_t* descC, const float* uRightC,
                             const uint8_t* occupied_in, int nC, const orc_frame_grid* gridC, const float* mvScaleFactors,
                             const float bounds[4], float th, float mbf, int mode, int mbCheckOrientation, int32_t* assigned,
                             int orb_dist) {
  const int thHigh = orb_dist > 0 ? orb_dist : kThHigh; /* the KeyFrame overload (:1306-1421) passes its own ORBdist */
  int nmatches = 0;
  std::fill(assigned, assigned + nC, -1);
  std::vector<uint8_t> occupiedC(occupied_in, occupied_in + nC);
  std::vector<std::vector<int>> rotHist(kHistoLength);
  const bool bForward = mode == 1, bBackward = mode == 2;
  for (int i = 0; i < nL; i++) {
    if (!(flagsL[i] & 1)) continue;  // pMP && !mvbOutlier[i]
    const float u = proj[3 * i], v = proj[3 * i + 1], invzc = proj[3 * i + 2];
    if (invzc < 0) continue;
    if (u < bounds[0] || u > bounds[1]) continue;
    if (v < bounds[2] || v > bounds[3]) continue;
    const int nLastOctave = kL[i].octave;
    const float radius = th * mvScaleFactors[nLastOctave];
    std::vector<int> vIndices2;
    if (bForward)
      vIndices2 = features_in_area(kC, *gridC, u, v, radius, nLastOctave, -1);
    else if (bBackward)
      vIndices2 = features_in_area(kC, *gridC, u, v, radius, 0, nLastOctave);
    else
      vIndices2 = features_in_area(kC, *gridC, u, v, radius, nLastOctave - 1, nLastOctave + 1);
    if (vIndices2.empty()) continue;
    const uint8_t* dMP = descMP + (size_t)i * 32;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : vIndices2) {
      if (occupiedC[i2]) continue;  // mvpMapPoints[i2] && Observations() > 0
      if (uRightC[i2] > 0) {
        const float ur = fmaf(-mbf, invzc, u);
        const float er = std::fabs(ur - uRightC[i2]);
        if (er > radius) continue;
      }
      const int dist = descriptor_distance(dMP, descC + (size_t)i2 * 32);
      if (dist < bestDist) {
        bestDist = dist;
        bestIdx2 = i2;
      }
    }
    if (bestDist <= thHigh) {
      assigned[bestIdx2] = i;                        // CurrentFrame.mvpMapPoints[bestIdx2] = pMP
      occupiedC[bestIdx2] = (flagsL[i] & 2) ? 1 : 0;  // what a later candidate test (:1018-1020) sees for it
      nmatches++;
      if (mbCheckOrientation) rotHist[rotation_bin(kLun[i].angle, kC[bestIdx2].angle)].push_back(bestIdx2);
    }
  }
  if (mbCheckOrientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    int32_t sizes[kHistoLength];
    for (int i = 0; i < kHistoLength; i++) sizes[i] = (int)rotHist[i].size();
    orc_three_maxima(sizes, kHistoLength, &ind1, &ind2, &ind3);
    for (int i = 0; i < kHistoLength; i++) {
      if (i != ind1 && i != ind2 && i != ind3) {
        for (int idx : rotHist[i]) {
          assigned[idx] = -1;  // mvpMapPoints[idx] = NULL
          nmatches--;
        }
      }
    }
  }
  return nmatches;
}
// ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th), src/ORBmatcher.cc:43-119 (the
// local-map search of Tracking::SearchLocalPoints), with RadiusByViewingCos :121-126.  Per map point the caller supplies
// (mTrackProjX, mTrackProjY, mTrackProjXR), mTrackViewCos, mnTrackScaleLevel, flags (bit 0: mbTrackInView && !isBad(), bit 1:
// Observations() > 0) and GetDescriptor(); for the frame mvuRight and `occupied` (mvpMapPoints[idx] && Observations() > 0 on
// entry; tracked as map points are assigned).  assigned[idx] = index of the map point the call leaves in F.mvpMapPoints[idx], else -1.
int orc_search_map_points(const float* proj /* n x 3 */, const float* view_cos, const int32_t* level, const uint8_t* flags,
                          const uint8_t* descMP, int nMP, const orc_keypoint* kF, const uint8_t* descF, const float* uRight,
                          const uint8_t* occupied_in, int nF, const orc_frame_grid* grid, const float* mvScaleFactors, float th,
                          float mfNNratio, int32_t* assigned) {
  int nmatches = 0;
  std::fill(assigned, assigned + nF, -1);
  std::vector<uint8_t> occupied(occupied_in, occupied_in + nF);
  const bool bFactor = th != 1.0;
  for (int iMP = 0; iMP < nMP; iMP++) {
    if (!(flags[iMP] & 1)) continue;
    const int nPredictedLevel = level[iMP];
    float r = view_cos[iMP] > 0.998 ? 2.5f : 4.0f;  // RadiusByViewingCos
    if (bFactor) r *= th;
    const std::vector<int> vIndices = features_in_area(kF, *grid, proj[3 * iMP], proj[3 * iMP + 1],
                                                       r * mvScaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel);
    if (vIndices.empty()) continue;
    const uint8_t* MPdescriptor = descMP + (size_t)iMP * 32;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int idx : vIndices) {
      if (occupied[idx]) continue;
      if (uRight[idx] > 0) {
        const float er = std::fabs(proj[3 * iMP + 2] - uRight[idx]);
        if (er > r * mvScaleFactors[nPredictedLevel]) continue;
      }
      const int dist = descriptor_distance(MPdescriptor, descF + (size_t)idx * 32);
      if (dist < bestDist) {
        bestDist2 = bestDist;
        bestDist = dist;
        bestLevel2 = bestLevel;
        bestLevel = kF[idx].octave;
        bestIdx = idx;
      } else if (dist < bestDist2) {
        bestLevel2 = kF[idx].octave;
        bestDist2 = dist;
      }
    }
    if (bestDist <= kThHigh) {
      if (bestLevel == bestLevel2 && bestDist > mfNNratio * bestDist2) continue;
      assigned[bestIdx] = iMP;
      occupied[bestIdx] = (flags[iMP] & 2) ? 1 : 0;
      nmatches++;
    }
  }
  return nmatches;
}
// ORBmatcher::SearchByPoints(currentKF, pKF, matches), src/ORBmatcher.cc:1209-1304 (loop detection, LoopClosing.cc:255).
// valid1 / valid2: the keypoint has a map point that is not bad.  matches12[idx1] = idx2 whose map point ends up in matches[idx1].
int orc_search_by_points(const orc_keypoint* k1, const uint8_t* d1s, const uint8_t* valid1, int n1, const orc_keypoint* k2,
                         const uint8_t* d2s, const uint8_t* valid2, int n2, float mfNNratio, int mbCheckOrientation, int32_t* matches12) {
  int nmatches = 0;
  std::vector<std::vector<int>> rotHist(kHistoLength);
  std::fill(matches12, matches12 + n1, -1);
  std::vector<char> vbMatched2((size_t)n2, 0);
  for (int idx1 = 0; idx1 < n1; idx1++) {
    if (!valid1[idx1]) continue;
    const uint8_t* d1 = d1s + (size_t)idx1 * 32;
    int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
    for (int idx2 = 0; idx2 < n2; idx2++) {
      if (!valid2[idx2] || vbMatched2[idx2]) continue;
      const int dist = descriptor_distance(d1, d2s + (size_t)idx2 * 32);
      if (dist < bestDist1) {
        bestDist2 = bestDist1;
        bestDist1 = dist;
        bestIdx2 = idx2;
      } else if (dist < bestDist2) {
        bestDist2 = dist;
      }
    }
    if (bestDist1 < kThLow) {
      if ((float)bestDist1 < mfNNratio * (float)bestDist2) {
        matches12[idx1] = bestIdx2;
        vbMatched2[bestIdx2] = 1;
        if (mbCheckOrientation) rotHist[rotation_bin(k1[idx1].angle, k2[bestIdx2].angle)].push_back(idx1);
        nmatches++;
      }
    }
  }
  if (mbCheckOrientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    int32_t sizes[kHistoLength];
    for (int i = 0; i < kHistoLength; i++) sizes[i] = (int)rotHist[i].size();
    orc_three_maxima(sizes, kHistoLength, &ind1, &ind2, &ind3);
    for (int i = 0; i < kHistoLength; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i]) {
        matches12[idx1] = -1;
        nmatches--;
      }
    }
  }
  return nmatches;
}
// The keypoint search inside ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th), src/ORBmatcher.cc:535-586: per map point
// that passed the checks of :489-531 (flags bit 0) the most similar keypoint within the radius, subject to the level window and
// the chi-square test of the reprojection error; best_idx = -1 where none reaches thDist.  The map surgery of :588-606 stays with the
// caller.  e2 as the reference build contracts it: fma(er, er, fma(ex, ex, ey*ey)) (checked against the compiled expression).
// bCheckReprojection = 0, thDist = TH_LOW is the search of the Sim3 overload Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)
// (:682-708; its bestDist starts at INT_MAX, which only differs for "no candidate": reported as 256 here too);
// bCheckReprojection = 0, thDist = TH_HIGH is either direction of SearchBySim3 (:812-845, :890-923).
void orc_fuse_search(const float* proj /* n x 3: u, v, ur */, const int32_t* level, const uint8_t* flags, const uint8_t* descMP, int nMP,
                     const orc_keypoint* kKF, const uint8_t* descKF, const float* uRight, const orc_frame_grid* grid,
                     const float* mvScaleFactors, const float* mvInvLevelSigma2, float th, int bCheckReprojection, int thDist,
                     int32_t* best_idx, int32_t* best_dist) {
  for (int i = 0; i < nMP; i++) {
    best_idx[i] = -1;
    best_dist[i] = 256;
    if (!(flags[i] & 1)) continue;
    const float u = proj[3 * i], v = proj[3 * i + 1], ur = proj[3 * i + 2];
    const int nPredictedLevel = level[i];
    const float radius = th * mvScaleFactors[nPredictedLevel];
    const std::vector<int> vIndices = features_in_area(kKF, *grid, u, v, radius, -1, -1);  // KeyFrame::GetFeaturesInArea: no level filter
    int bestDist = INT_MAX, bestIdx = -1;
    for (int idx : vIndices) {
      const orc_keypoint& kp = kKF[idx];
      const int kpLevel = kp.octave;
      if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
      if (bCheckReprojection) {
        if (uRight[idx] >= 0) {
          const float ex = u - kp.x, ey = v - kp.y, er = ur - uRight[idx];
          const float e2 = fmaf(er, er, fmaf(ex, ex, ey * ey));
          if (e2 * mvInvLevelSigma2[kpLevel] > 7.8) continue;
        } else {
          const float ex = u - kp.x, ey = v - kp.y;
          const float e2 = fmaf(ex, ex, ey * ey);
          if (e2 * mvInvLevelSigma2[kpLevel] > 5.99) continue;
        }
      }
      const int dist = descriptor_distance(descMP + (size_t)i * 32, descKF + (size_t)idx * 32);
      if (dist < bestDist) {
        bestDist = dist;
        bestIdx = idx;
      }
    }
    if (bestIdx >= 0) best_dist[i] = bestDist;
    if (bestDist <= thDist) best_idx[i] = bestIdx;
  }
}
// ORBmatcher::SearchBySim3, src/ORBmatcher.cc:734-944, from the projections on: map point i1 of KF1 (flags1 bit 0: it exists, is
// not bad, is not already matched and passed :786-808) is searched among the keypoints of KF2 and vice versa (both TH_HIGH, no
// reprojection gate), then the agreement check :927-941.  vnMatch1 / vnMatch2 as in the reference; matches12[i1] = idx2 of the
// map point the call puts into vpMatches12[i1], else -1; returns nFound.
int orc_search_by_sim3(const float* proj1, const int32_t* level1, const uint8_t* flags1, const uint8_t* descMP1, int N1,
                       const float* proj2, const int32_t* level2, const uint8_t* flags2, const uint8_t* descMP2, int N2,
                       const orc_keypoint* k1, const uint8_t* d1, const orc_frame_grid* grid1, const orc_keypoint* k2, const uint8_t* d2,
                       const orc_frame_grid* grid2, const float* mvScaleFactors1, const float* mvScaleFactors2, float th,
                       int32_t* vnMatch1, int32_t* vnMatch2, int32_t* matches12) {
  std::vector<int32_t> dist((size_t)std::max(N1, N2) + 1);
  orc_fuse_search(proj1, level1, flags1, descMP1, N1, k2, d2, nullptr, grid2, mvScaleFactors2, nullptr, th, 0, kThHigh, vnMatch1, dist.data());
  orc_fuse_search(proj2, level2, flags2, descMP2, N2, k1, d1, nullptr, grid1, mvScaleFactors1, nullptr, th, 0, kThHigh, vnMatch2, dist.data());
  int nFound = 0;
  for (int i1 = 0; i1 < N1; i1++) {
    matches12[i1] = -1;
    const int idx2 = vnMatch1[i1];
    if (idx2 >= 0) {
      const int idx1 = vnMatch2[idx2];
      if (idx1 == i1) {
        matches12[i1] = idx2;
        nFound++;
      }
    }
  }
  return nFound;
}
// ORBmatcher::SearchByProjection(KeyFrame* pKF, Scw, vpPoints, vpMatched, th), src/ORBmatcher.cc:146-254 (loop closing), from the
// projection on: flags bit 0 = the point is not bad, not in spAlreadyFound and passed :179-209.  vpMatched_in[idx] != 0: the
// keypoint holds a map point on entry; every match made occupies its keypoint for the later points (:246).  assigned[idx] = index
// of the point the call writes into vpMatched[idx], else -1; returns nmatches.
int orc_search_by_projection_sim3(const float* proj /* n x 3, [2] unused */, const int32_t* level, const uint8_t* flags, const uint8_t* descMP,
                                  int nMP, const orc_keypoint* kKF, const uint8_t* descKF, const uint8_t* vpMatched_in, int nKF,
                                  const orc_frame_grid* grid, const float* mvScaleFactors, int th, int32_t* assigned) {
  int nmatches = 0;
  std::fill(assigned, assigned + nKF, -1);
  std::vector<uint8_t> vpMatched(vpMatched_in, vpMatched_in + nKF);
  for (int iMP = 0; iMP < nMP; iMP++) {
    if (!(flags[iMP] & 1)) continue;
    const int nPredictedLevel = level[iMP];
    const float radius = th * mvScaleFactors[nPredictedLevel];
    const std::vector<int> vIndices = features_in_area(kKF, *grid, proj[3 * iMP], proj[3 * iMP + 1], radius, -1, -1);
    if (vIndices.empty()) continue;
    int bestDist = 256, bestIdx = -1;
    for (int idx : vIndices) {
      if (vpMatched[idx]) continue;
      const int kpLevel = kKF[idx].octave;
      if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
      const int dist = descriptor_distance(descMP + (size_t)iMP * 32, descKF + (size_t)idx * 32);
      if (dist < bestDist) {
        bestDist = dist;
        bestIdx = idx;
      }
    }
    if (bestDist <= kThLow) {
      vpMatched[bestIdx] = 1;
      assigned[bestIdx] = iMP;
      nmatches++;
    }
  }
  return nmatches;
}
// ORBmatcher::CheckDistEpipolarLine, src/ORBmatcher.cc:128-144.  F12 is an Eigen::Matrix3d, so a, b, c are evaluated in double
// and rounded to float; the reference's -O3 -march=native contracts  p*q + r*s  into  fma(p, q, r*s)  (first product fused,
// checked against the compiled verbatim expression in tests/test_oracle_search.py), frozen here with explicit fma / fmaf.
static bool check_dist_epipolar_line(float x1, float y1, float x2, float y2, const double* F12 /* row-major */, float sigma2) {
  const float a = (float)(fma((double)x1, F12[0 * 3 + 0], (double)y1 * F12[1 * 3 + 0]) + F12[2 * 3 + 0]);
  const float b = (float)(fma((double)x1, F12[0 * 3 + 1], (double)y1 * F12[1 * 3 + 1]) + F12[2 * 3 + 1]);
  const float c = (float)(fma((double)x1, F12[0 * 3 + 2], (double)y1 * F12[1 * 3 + 2]) + F12[2 * 3 + 2]);
  const float num = fmaf(a, x2, b * y2) + c;
  const float den = fmaf(a, a, b * b);
  if (den == 0) return false;
  const float dsqr = num * num / den;
  return dsqr < 3.84 * sigma2;
}
int orc_check_dist_epipolar_line(float x1, float y1, float x2, float y2, const double* F12, float sigma2) {
  return check_dist_epipolar_line(x1, y1, x2, y2, F12, sigma2) ? 1 : 0;
}
// ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:359-462, from the epipole (ex, ey) on (:361-368 is pose algebra that stays
// with the caller).  has_mp1 / has_mp2: vpMapPoints[idx] != NULL; u_right: mvuRight.  Note that this reference never sets
// vbMatched2 (:373, :405): several keypoints of KF1 may take the same keypoint of KF2, exactly as restated.
int orc_search_for_triangulation(const orc_keypoint* k1, const uint8_t* d1s, const uint8_t* has_mp1, const float* u_right1, int n1,
                                 const orc_keypoint* k2, const uint8_t* d2s, const uint8_t* has_mp2, const float* u_right2, int n2,
                                 const double* F12, float ex, float ey, const float* mvScaleFactors, const float* mvLevelSigma2,
                                 int mbCheckOrientation, int32_t* vMatches12) {
  int nmatches = 0;
  std::vector<char> vbMatched2((size_t)n2, 0);
  std::fill(vMatches12, vMatches12 + n1, -1);
  std::vector<std::vector<int>> rotHist(kHistoLength);
  for (int idx1 = 0; idx1 < n1; idx1++) {
    if (has_mp1[idx1]) continue;
    const bool bStereo1 = u_right1[idx1] >= 0;
    const orc_keypoint& kp1 = k1[idx1];
    const uint8_t* d1 = d1s + (size_t)idx1 * 32;
    int bestDist = kThLow, bestIdx2 = -1;
    for (int idx2 = 0; idx2 < n2; idx2++) {
      if (vbMatched2[idx2] || has_mp2[idx2]) continue;
      const bool bStereo2 = u_right2[idx2] >= 0;
      const orc_keypoint& kp2 = k2[idx2];
      if (!check_dist_epipolar_line(kp1.x, kp1.y, kp2.x, kp2.y, F12, mvLevelSigma2[kp2.octave])) continue;
      const int dist = descriptor_distance(d1, d2s + (size_t)idx2 * 32);
      if (dist > kThLow || dist > bestDist) continue;
      if (!bStereo1 && !bStereo2) {
        const float distex = ex - kp2.x;
        const float distey = ey - kp2.y;
        if (fmaf(distex, distex, distey * distey) < 100 * mvScaleFactors[kp2.octave]) continue;
      }
      bestIdx2 = idx2;
      bestDist = dist;
    }
    if (bestIdx2 >= 0) {
      vMatches12[idx1] = bestIdx2;
      nmatches++;
      if (mbCheckOrientation) rotHist[rotation_bin(kp1.angle, k2[bestIdx2].angle)].push_back(idx1);
    }
  }
  if (mbCheckOrientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    int32_t sizes[kHistoLength];
    for (int i = 0; i < kHistoLength; i++) sizes[i] = (int)rotHist[i].size();
    orc_three_maxima(sizes, kHistoLength, &ind1, &ind2, &ind3);
    for (int i = 0; i < kHistoLength; i++) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i]) {
        vMatches12[idx1] = -1;
        nmatches--;
      }
    }
  }
  return nmatches;
}
// MapPoint::ComputeDistinctiveDescriptors, /root/reference/src/MapPoint.cc:252-275: all pairwise distances of the N
// observed descriptors (float matrix), per row std::sort and the element at index 0.5*(N-1) as median, the FIRST row
// with the smallest median wins.
int orc_distinctive(const uint8_t* desc, int n, int* best_median) {
  if (n <= 0) {
    if (best_median) *best_median = INT_MAX;
    return -1;
  }
  std::vector<float> D((size_t)n * n);
  for (int i = 0; i < n; i++) {
    D[(size_t)i * n + i] = 0;
    for (int j = i + 1; j < n; j++) {
      const int dij = descriptor_distance(desc + (size_t)i * 32, desc + (size_t)j * 32);
      D[(size_t)i * n + j] = (float)dij;
      D[(size_t)j * n + i] = (float)dij;
    }
  }
  int BestMedian = INT_MAX, BestIdx = 0;
  for (int i = 0; i < n; i++) {
    std::vector<int> vDists(D.begin() + (size_t)i * n, D.begin() + (size_t)(i + 1) * n);
    std::sort(vDists.begin(), vDists.end());
    const int median = vDists[(size_t)(0.5 * (n - 1))];
    if (median < BestMedian) {
      BestMedian = median;
      BestIdx = i;
    }
  }
  if (best_median) *best_median = BestMedian;
  return BestIdx;
}
void orc_distinctive_many(const uint8_t* desc, const int32_t* offsets, int nsets, int nthreads, int32_t* best_idx,
                          int32_t* best_median) {
  parallel_for(nsets, nthreads, [&](int s) {
    int med = 0;
    best_idx[s] = orc_distinctive(desc + (size_t)offsets[s] * 32, offsets[s + 1] - offsets[s], &med);
    if (best_median) best_median[s] = med;
  });
}

// Frame::AssignFeaturesToGrid + PosInGrid, /root/reference/src/Frame.cc:179-192, 323-332 (FRAME_GRID_COLS 64, FRAME_GRID_ROWS 48,
// src/Frame.h:37-38).  mGrid[x][y] is returned in CSR form: cell x * 48 + y owns indices[cell_start[c] .. cell_start[c+1]).
void orc_assign_grid(const orc_keypoint* kps_un, int n, float mnMinX, float mnMinY, float mfGridElementWidthInv,
                     float mfGridElementHeightInv, int32_t* cell_start, int32_t* indices) {
  const int COLS = 64, ROWS = 48;
  std::vector<std::vector<int>> mGrid((size_t)COLS * ROWS);
  for (int i = 0; i < n; i++) {
    const orc_keypoint& kp = kps_un[i];
    const int posX = (int)std::round((kp.x - mnMinX) * mfGridElementWidthInv);
    const int posY = (int)std::round((kp.y - mnMinY) * mfGridElementHeightInv);
    if (posX < 0 || posX >= COLS || posY < 0 || posY >= ROWS) continue;
    mGrid[(size_t)posX * ROWS + posY].push_back(i);
  }
  int acc = 0;
  for (int c = 0; c < COLS * ROWS; c++) {
    cell_start[c] = acc;
    for (int i : mGrid[c]) indices[acc++] = i;
  }
  cell_start[COLS * ROWS] = acc;
}
// cv::undistortPoints(src, dst, K, dist, noArray(), K) as OpenCV 4.13 computes it (calib3d undistort.dispatch.cpp,
// cvUndistortPointsInternal: double arithmetic, 5 fixed-point iterations, no tilt, R = identity, P = K), which is what
// Frame::UndistortKeyPoints (/root/reference/src/Frame.cc:335-366) and ComputeImageBounds (:368-397) call.  K and dist are
// the CV_32F values the reference passes (Converter::toCvMat, src/Converter.cc:53-60).  k = (k1, k2, p1, p2[, k3]).
static void undistort_point(float u_in, float v_in, const float K[4], const float* dist, int ndist, float* xo, float* yo) {
  const double fx = K[0], fy = K[1], cx = K[2], cy = K[3], ifx = 1. / fx, ify = 1. / fy;
  double k[14] = {0};
  for (int i = 0; i < ndist && i < 14; i++) k[i] = dist[i];
  const double u = u_in, v = v_in;
  double x = (u - cx) * ifx, y = (v - cy) * ify;
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; j++) {
    const double r2 = x * x + y * y;
    const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
    if (icdist < 0) {  // the distortion model folded over: OpenCV falls back to the undistorted guess
      x = (u - cx) * ifx;
      y = (v - cy) * ify;
      break;
    }
    const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
    const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
    x = (x0 - deltaX) * icdist;
    y = (y0 - deltaY) * icdist;
  }
  const double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
  *xo = (float)(xx * ww);
  *yo = (float)(yy * ww);
}
// Frame::UndistortKeyPoints, src/Frame.cc:335-366
void orc_undistort_keypoints(const orc_keypoint* kps, int n, const float K[4], const float* dist, int ndist, orc_keypoint* out) {
  for (int i = 0; i < n; i++) {
    out[i] = kps[i];
    if (ndist > 0 && dist[0] != 0.0f) undistort_point(kps[i].x, kps[i].y, K, dist, ndist, &out[i].x, &out[i].y);
  }
}
// Frame::ComputeImageBounds, src/Frame.cc:368-397: bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY}
void orc_image_bounds(int cols, int rows, const float K[4], const float* dist, int ndist, float bounds[4]) {
  if (ndist > 0 && dist[0] != 0.0f) {
    const float cx[4] = {0.f, (float)cols, 0.f, (float)cols}, cy[4] = {0.f, 0.f, (float)rows, (float)rows};
    float mx[4], my[4];
    for (int i = 0; i < 4; i++) undistort_point(cx[i], cy[i], K, dist, ndist, &mx[i], &my[i]);
    bounds[0] = std::min(mx[0], mx[2]);
    bounds[1] = std::max(mx[1], mx[3]);
    bounds[2] = std::min(my[0], my[1]);
    bounds[3] = std::max(my[2], my[3]);
  } else {
    bounds[0] = 0.0f;
    bounds[1] = (float)cols;
    bounds[2] = 0.0f;
    bounds[3] = (float)rows;
  }
}

// Frame::ComputeStereoFromRGBD, src/Frame.cc:399-417
void orc_stereo_from_rgbd(const orc_keypoint* kps, const orc_keypoint* kps_un, int n, const float* depth, int width, float mbf,
                          float* u_right, float* z) {
  for (int i = 0; i < n; i++) {
    u_right[i] = -1;
    z[i] = -1;
    const float v = kps[i].y, u = kps[i].x;
    const float d = depth[(size_t)(int)v * width + (int)u];  // Mat::at<float>(int, int) with float arguments
    if (d > 0) {
      z[i] = d;
      u_right[i] = kps_un[i].x - mbf / d;
    }
  }
}

void orc_hamming_matrix(const uint8_t* A, int nA, const uint8_t* B, int nB, uint16_t* out) {
  for (int i = 0; i < nA; i++)
    for (int j = 0; j < nB; j++) out[(size_t)i * nB + j] = (uint16_t)descriptor_distance(A + (size_t)i * 32, B + (size_t)j * 32);
}

}  // extern "C"
