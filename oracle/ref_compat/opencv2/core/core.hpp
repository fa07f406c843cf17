/*
 * oracle/ref_compat -- a minimal OpenCV-shaped `cv::` surface in front of the UNMODIFIED reference sources.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as oracle/sdorb_oracle.h).  OpenCV's C++ headers and libraries are not in this
 * image, so /root/reference/src/ORBextractor.cc cannot be built against the real thing.  This directory supplies exactly
 * the names that file uses (SURVEY.md section 8c lists them) so that oracle/ref_build/Makefile can compile the reference's
 * own text, read where it lies under /root/reference, into oracle/_ref/libsdorb_ref.so:
 *
 *   - the container types (Mat with reference-counted storage, ROI views, MatExpr assignment semantics, KeyPoint, Point_,
 *     Size, Rect, Range, InputArray / OutputArray) follow OpenCV 4.x's observable behaviour for the calls the reference makes;
 *   - the pixel arithmetic (cv::resize INTER_LINEAR 8U, cv::copyMakeBorder, cv::FAST 9/16 + NMS, cv::GaussianBlur 8U,
 *     cv::KeyPointsFilter::retainBest, cv::fastAtan2) forwards to the oracle's primitives, each of which is pinned bit for
 *     bit against the real cv2 4.13.0 (tests/test_oracle_primitives.py).
 *
 * What this buys: the CONTROL FLOW and FLOAT EXPRESSIONS of the path (cell grid, ROI arithmetic with its float -> int
 * conversions, quota redistribution, retain order, IC_Angle, the rBRIEF index math with whatever FMA contraction g++
 * applies under the reference's own flags, keypoint scaling, pyramid chaining) are the reference's own compiled text, not a
 * restatement.
 */
#ifndef SDORB_REF_COMPAT_CORE_HPP
#define SDORB_REF_COMPAT_CORE_HPP

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <iterator>
#include <list>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../../sdorb_oracle.h"

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_PI 3.1415926535897932384626433832795
#define CV_CN_SHIFT 3
#define CV_DEPTH_MAX (1 << CV_CN_SHIFT)
#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH_MASK (CV_DEPTH_MAX - 1)
#define CV_MAT_DEPTH(flags) ((flags) & CV_MAT_DEPTH_MASK)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAT_CN(flags) ((((flags) >> CV_CN_SHIFT) & 63) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

/* OpenCV: SSE2 cvtsd2si / cvtss2si, i.e. round to nearest even in the default rounding mode == lrint / lrintf */
static inline int cvRound(double value) { return (int)lrint(value); }
static inline int cvRound(float value) { return (int)lrintf(value); }
static inline int cvRound(int value) { return value; }
static inline int cvFloor(double value) { int i = (int)value; return i - (i > value); }
static inline int cvFloor(float value) { int i = (int)value; return i - (i > value); }
static inline int cvFloor(int value) { return value; }
static inline int cvCeil(double value) { int i = (int)value; return i + (i < value); }
static inline int cvCeil(float value) { int i = (int)value; return i + (i < value); }
static inline int cvCeil(int value) { return value; }

namespace cv {

class Exception : public std::exception {
 public:
  explicit Exception(const std::string& m) : msg(m) {}
  ~Exception() throw() {}
  const char* what() const throw() { return msg.c_str(); }
  std::string msg;
};
#define SDORB_CV_ASSERT(expr) do { if (!(expr)) throw cv::Exception("OpenCV assertion failed: " #expr); } while (0)
#define CV_Assert(expr) SDORB_CV_ASSERT(expr)

template <typename T> static inline T saturate_cast(float v) { return T(v); }
template <typename T> static inline T saturate_cast(double v) { return T(v); }
template <typename T> static inline T saturate_cast(int v) { return T(v); }
template <> inline int saturate_cast<int>(float v) { return cvRound(v); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }

template <typename T> class Point_ {
 public:
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> Point_(const Point_<U>& p) : x(saturate_cast<T>(p.x)), y(saturate_cast<T>(p.y)) {}
  T x, y;
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
/* OpenCV core/types.hpp: a.x = saturate_cast<T>(a.x * b) */
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, int b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, double b) { a.x = saturate_cast<T>(a.x * b); a.y = saturate_cast<T>(a.y * b); return a; }

template <typename T> class Size_ {
 public:
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
  T width, height;
};
typedef Size_<int> Size;

template <typename T> class Rect_ {
 public:
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
  T x, y, width, height;
};
typedef Rect_<int> Rect;

class Range {
 public:
  Range() : start(0), end(0) {}
  Range(int s, int e) : start(s), end(e) {}
  static Range all() { return Range(INT_MIN, INT_MAX); }
  bool operator==(const Range& r) const { return start == r.start && end == r.end; }
  int start, end;
};

/* Binary layout of cv::KeyPoint (28 bytes) == orc_keypoint. */
class KeyPoint {
 public:
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
};

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4,
       BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };

class Mat;
/* the part of MatExpr the reference uses: Mat::zeros(...) assigned to an existing Mat (see Mat::operator=(const MatExpr&)) */
struct MatExpr {
  int rows, cols, type;
};

/* Dense 2-D matrices of 8U / 32F / 64F elements with 1..4 channels (the path itself asserts CV_8UC1, src/ORBextractor.cc:626;
 * src/Frame.cc also keeps the distortion coefficients and the point list of cv::undistortPoints in CV_32F matrices). */
static inline size_t sdorb_elem_size(int type) {
  const int depth = CV_MAT_DEPTH(type);
  SDORB_CV_ASSERT(depth == CV_8U || depth == CV_16U || depth == CV_32F || depth == CV_64F);
  return (size_t)(depth == CV_8U ? 1 : depth == CV_16U ? 2 : depth == CV_32F ? 4 : 8) * (size_t)CV_MAT_CN(type);
}

class Mat {
 public:
  Mat() : rows(0), cols(0), data(0), step(0), datastart(0), whole_rows(0), whole_cols(0), refcount(0), type_(CV_8UC1) {}
  Mat(int r, int c, int type) : rows(0), cols(0), data(0), step(0), datastart(0), whole_rows(0), whole_cols(0), refcount(0), type_(CV_8UC1) {
    create(r, c, type);
  }
  Mat(Size sz, int type) : rows(0), cols(0), data(0), step(0), datastart(0), whole_rows(0), whole_cols(0), refcount(0), type_(CV_8UC1) {
    create(sz.height, sz.width, type);
  }
  /* user-allocated data: no ownership (cv::Mat(rows, cols, type, void* data, size_t step)) */
  Mat(int r, int c, int type, void* d, size_t st = 0)
      : rows(r), cols(c), data((uchar*)d), step(st ? st : (size_t)c * sdorb_elem_size(type)), datastart((uchar*)d), whole_rows(r),
        whole_cols(c), refcount(0), type_(type) {}
  Mat(const Mat& m)
      : rows(m.rows), cols(m.cols), data(m.data), step(m.step), datastart(m.datastart), whole_rows(m.whole_rows),
        whole_cols(m.whole_cols), refcount(m.refcount), type_(m.type_) {
    if (refcount) ++*refcount;
  }
  /* cv::Mat(const Mat&, const Range& rowRange, const Range& colRange) */
  Mat(const Mat& m, const Range& rr, const Range& cr)
      : rows(m.rows), cols(m.cols), data(m.data), step(m.step), datastart(m.datastart), whole_rows(m.whole_rows),
        whole_cols(m.whole_cols), refcount(m.refcount), type_(m.type_) {
    if (refcount) ++*refcount;
    try {
      if (!(rr == Range::all()) && !(rr == Range(0, rows))) {
        SDORB_CV_ASSERT(0 <= rr.start && rr.start <= rr.end && rr.end <= m.rows);
        rows = rr.end - rr.start;
        data += step * (size_t)rr.start;
      }
      if (!(cr == Range::all()) && !(cr == Range(0, cols))) {
        SDORB_CV_ASSERT(0 <= cr.start && cr.start <= cr.end && cr.end <= m.cols);
        cols = cr.end - cr.start;
        data += (size_t)cr.start * elemSize();
      }
    } catch (...) {
      release();
      throw;
    }
  }
  /* cv::Mat(const Mat&, const Rect&) */
  Mat(const Mat& m, const Rect& roi)
      : rows(roi.height), cols(roi.width), data(m.data + (size_t)roi.y * m.step + (size_t)roi.x * m.elemSize()), step(m.step),
        datastart(m.datastart), whole_rows(m.whole_rows), whole_cols(m.whole_cols), refcount(m.refcount), type_(m.type_) {
    if (refcount) ++*refcount;
    try {
      SDORB_CV_ASSERT(0 <= roi.x && 0 <= roi.width && roi.x + roi.width <= m.cols && 0 <= roi.y && 0 <= roi.height &&
                      roi.y + roi.height <= m.rows);
    } catch (...) {
      release();
      throw;
    }
  }
  ~Mat() { release(); }
  Mat& operator=(const Mat& m) {
    if (this != &m) {
      if (m.refcount) ++*m.refcount;
      release();
      rows = m.rows; cols = m.cols; data = m.data; step = m.step; datastart = m.datastart;
      whole_rows = m.whole_rows; whole_cols = m.whole_cols; refcount = m.refcount; type_ = m.type_;
    }
    return *this;
  }
  /* OpenCV: MatOp_Initializer::assign -> m.create(size, type) (keeps the buffer when size and type already match, also for a
   * view into another matrix) followed by m = Scalar(0): the zeros are written THROUGH an existing header. */
  Mat& operator=(const MatExpr& e) {
    create(e.rows, e.cols, e.type);
    for (int y = 0; y < rows; ++y) memset(data + (size_t)y * step, 0, (size_t)cols * elemSize());
    return *this;
  }
  static MatExpr zeros(int rows, int cols, int type) { MatExpr e = {rows, cols, type}; return e; }

  void create(int r, int c, int type) {
    SDORB_CV_ASSERT(r >= 0 && c >= 0);
    const size_t esz = sdorb_elem_size(type);
    if (data && rows == r && cols == c && type_ == type) return;
    if (!data && r == 0 && c == 0 && rows == 0 && cols == 0) { type_ = type; return; }
    release();
    rows = r; cols = c; type_ = type; step = (size_t)c * esz;
    whole_rows = r; whole_cols = c;
    const size_t bytes = (size_t)r * (size_t)c * esz;
    uchar* block = (uchar*)malloc(sizeof(long) * 2 + (bytes ? bytes : 1));
    if (!block) throw Exception("out of memory");
    refcount = (long*)block;
    *refcount = 1;
    data = datastart = block + sizeof(long) * 2;
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  void release() {
    if (refcount && --*refcount == 0) free(refcount);
    refcount = 0;
    data = datastart = 0;
    rows = cols = 0;
    step = 0;
    whole_rows = whole_cols = 0;
  }
  Mat clone() const {
    Mat m;
    if (rows > 0 && cols > 0) {
      m.create(rows, cols, type_);
      for (int y = 0; y < rows; ++y) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
    } else {
      m.type_ = type_;
    }
    return m;
  }
  void copyTo(Mat& dst) const {
    dst.create(rows, cols, type_);
    for (int y = 0; y < rows; ++y) memmove(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols * elemSize());
  }
  Mat row(int y) const { return Mat(*this, Range(y, y + 1), Range::all()); }
  Mat col(int x) const { return Mat(*this, Range::all(), Range(x, x + 1)); }
  Mat rowRange(int startrow, int endrow) const { return Mat(*this, Range(startrow, endrow), Range::all()); }
  Mat colRange(int startcol, int endcol) const { return Mat(*this, Range::all(), Range(startcol, endcol)); }
  Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
  Mat operator()(Range rr, Range cr) const { return Mat(*this, rr, cr); }
  /* cv::Mat::reshape(cn): same data, another channel count (continuous matrices only, rows kept) */
  Mat reshape(int cn, int new_rows = 0) const {
    SDORB_CV_ASSERT(new_rows == 0 && isContinuous() && cn >= 1 && cn <= 4);
    const int total_scalars = cols * channels();
    SDORB_CV_ASSERT(total_scalars % cn == 0);
    Mat m(*this);
    m.cols = total_scalars / cn;
    m.type_ = CV_MAKETYPE(depth(), cn);
    m.whole_rows = m.rows;
    m.whole_cols = m.cols;
    return m;
  }
  /* size of the whole matrix this header views and the view's offset in it */
  void locateROI(Size& wholeSize, Point& ofs) const {
    const ptrdiff_t delta = data - datastart;
    if (delta == 0 || step == 0) {
      ofs.x = ofs.y = 0;
    } else {
      ofs.y = (int)(delta / (ptrdiff_t)step);
      ofs.x = (int)((delta - (ptrdiff_t)step * ofs.y) / (ptrdiff_t)elemSize());
    }
    wholeSize.height = whole_rows;
    wholeSize.width = whole_cols;
  }
  bool isSubmatrix() const { return rows != whole_rows || cols != whole_cols; }
  int type() const { return type_; }
  int depth() const { return CV_MAT_DEPTH(type_); }
  int channels() const { return CV_MAT_CN(type_); }
  size_t elemSize() const { return sdorb_elem_size(type_); }
  size_t elemSize1() const { return elemSize() / (size_t)channels(); }
  size_t step1(int = 0) const { return step / elemSize1(); }
  bool empty() const { return data == 0 || rows == 0 || cols == 0; }
  size_t total() const { return (size_t)rows * (size_t)cols; }
  Size size() const { return Size(cols, rows); }
  bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
  template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
  template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }
  /* single index: element i of a single-row or single-column matrix (Mat::at(int i0)) */
  template <typename T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + (size_t)i * step); }
  template <typename T> const T& at(int i) const { return rows == 1 ? ((const T*)data)[i] : *(const T*)(data + (size_t)i * step); }

  int rows, cols;
  uchar* data;
  size_t step;
  /* bookkeeping */
  uchar* datastart;
  int whole_rows, whole_cols;
  long* refcount;

 private:
  int type_;
};

/* InputArray / OutputArray: proxies around a Mat, as the reference uses them (empty / getMat / create / release). */
class _InputArray {
 public:
  _InputArray() : m(0) {}
  _InputArray(const Mat& mat) : m(const_cast<Mat*>(&mat)) {}
  bool empty() const { return !m || m->empty(); }
  Mat getMat(int = -1) const { return m ? *m : Mat(); }
  Mat* m;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray() {}
  _OutputArray(Mat& mat) : _InputArray(mat) {}
  void create(int rows, int cols, int type) const { SDORB_CV_ASSERT(m != 0); m->create(rows, cols, type); }
  void create(Size sz, int type) const { create(sz.height, sz.width, type); }
  void release() const { if (m) m->release(); }
  bool needed() const { return m != 0; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
static inline InputArray noArray() { static _InputArray none; return none; }

static inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

/* cv::borderInterpolate for BORDER_REFLECT_101 */
static inline int sdorb_reflect101(int p, int len) {
  if ((unsigned)p < (unsigned)len) return p;
  if (len == 1) return 0;
  do {
    if (p < 0) p = -p; else p = 2 * (len - 1) - p;
  } while ((unsigned)p >= (unsigned)len);
  return p;
}

/* cv::copyMakeBorder (imgproc/src/... copyMakeBorder_8u): without BORDER_ISOLATED a source that is a view first grows into
 * the pixels that really surround it in its parent (adjustROI), only the rest is synthesised; the inner block is copied
 * unless it already lies in place. */
static inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType) {
  SDORB_CV_ASSERT(top >= 0 && bottom >= 0 && left >= 0 && right >= 0);
  Mat src = _src.getMat();
  SDORB_CV_ASSERT(src.type() == CV_8UC1);  /* one byte per pixel below */
  if (src.isSubmatrix() && (borderType & BORDER_ISOLATED) == 0) {
    Size wholeSize;
    Point ofs;
    src.locateROI(wholeSize, ofs);
    const int dtop = std::min(ofs.y, top), dbottom = std::min(wholeSize.height - src.rows - ofs.y, bottom);
    const int dleft = std::min(ofs.x, left), dright = std::min(wholeSize.width - src.cols - ofs.x, right);
    src.data -= (size_t)dtop * src.step + (size_t)dleft;
    src.rows += dtop + dbottom;
    src.cols += dleft + dright;
    top -= dtop; left -= dleft; bottom -= dbottom; right -= dright;
  }
  borderType &= ~BORDER_ISOLATED;
  SDORB_CV_ASSERT(borderType == BORDER_REFLECT_101);
  _dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
  Mat dst = _dst.getMat();
  if (top == 0 && left == 0 && bottom == 0 && right == 0) {
    if (src.data != dst.data || src.step != dst.step) src.copyTo(dst);
    return;
  }
  uchar* inner = dst.data + (size_t)top * dst.step + (size_t)left;
  if (inner != src.data || dst.step != src.step)
    for (int y = 0; y < src.rows; ++y) memmove(inner + (size_t)y * dst.step, src.data + (size_t)y * src.step, (size_t)src.cols);
  /* left / right of every inner row, then whole rows above and below (rows already carry their side borders) */
  for (int y = 0; y < src.rows; ++y) {
    uchar* row = inner + (size_t)y * dst.step;
    for (int x = 1; x <= left; ++x) row[-x] = row[sdorb_reflect101(-x, src.cols)];
    for (int x = 0; x < right; ++x) row[src.cols + x] = row[sdorb_reflect101(src.cols + x, src.cols)];
  }
  for (int y = 1; y <= top; ++y)
    memcpy(dst.data + (size_t)(top - y) * dst.step, dst.data + (size_t)(top + sdorb_reflect101(-y, src.rows)) * dst.step, (size_t)dst.cols);
  for (int y = 0; y < bottom; ++y)
    memcpy(dst.data + (size_t)(top + src.rows + y) * dst.step,
           dst.data + (size_t)(top + sdorb_reflect101(src.rows + y, src.rows)) * dst.step, (size_t)dst.cols);
}

}  // namespace cv

#endif
