// Minimal stand-in for the parts of <opencv2/core/core.hpp> that include/ORBextractor.h uses.  TEST INFRASTRUCTURE:
// OpenCV's C++ headers are not installed in this image, so the shim is compiled and run against this mock
// (tests/test_cpp_shim.py).  Semantics follow OpenCV: Mat is a ref-counted view (data / step / rows / cols), ROI views
// share the parent's buffer, OutputArray::create reallocates only on a shape change.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <vector>

#define CV_8U 0
#define CV_16U 2
#define CV_8UC1 0
#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception(#expr); } while (0)

namespace cv {
struct Exception : std::runtime_error {
  explicit Exception(const char* m) : std::runtime_error(m) {}
};
struct Rect {
  int x, y, width, height;
  Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};
struct Point2f {
  float x, y;
};
struct Point {
  int x, y;
};
struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt{0, 0}, size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
      : pt{x, y}, size(s), angle(a), response(r), octave(o), class_id(c) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint is 28 bytes");

class Mat {
 public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  size_t step = 0;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ext, size_t st) : rows(r), cols(c), data((unsigned char*)ext), step(st), type_(type) {}
  void create(int r, int c, int type) {
    if (r == rows && c == cols && type == type_ && data) return;
    const size_t es = type == CV_16U ? 2 : 1;
    buf_ = std::shared_ptr<unsigned char>(new unsigned char[(size_t)r * c * es + 1], std::default_delete<unsigned char[]>());
    rows = r;
    cols = c;
    type_ = type;
    step = (size_t)c * es;
    data = buf_.get();
  }
  void release() { *this = Mat(); }
  Mat operator()(const Rect& r) const {
    if (r.x < 0 || r.y < 0 || r.x + r.width > cols || r.y + r.height > rows) throw Exception("ROI outside the matrix");
    Mat m = *this;
    m.rows = r.height;
    m.cols = r.width;
    m.data = data + (size_t)r.y * step + (size_t)r.x * (type_ == CV_16U ? 2 : 1);
    return m;
  }
  int type() const { return type_; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return step == (size_t)cols * (type_ == CV_16U ? 2 : 1) || rows <= 1; }
  unsigned char* ptr(int r = 0) { return data + (size_t)r * step; }
  const unsigned char* ptr(int r = 0) const { return data + (size_t)r * step; }

 private:
  std::shared_ptr<unsigned char> buf_;
  int type_ = CV_8UC1;
};

class _InputArray {
 public:
  _InputArray(const Mat& m) : m_(&m) {}
  bool empty() const { return m_->empty(); }
  Mat getMat() const { return *m_; }

 private:
  const Mat* m_;
};
class _OutputArray {
 public:
  _OutputArray(Mat& m) : m_(&m) {}
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  void release() const { m_->release(); }
  Mat getMat() const { return *m_; }

 private:
  Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
}  // namespace cv
