/* oracle/ref_compat: cv::resize (INTER_LINEAR, 8UC1) and cv::GaussianBlur (7x7, sigma 2, 8UC1) -- see core/core.hpp. */
#ifndef SDORB_REF_COMPAT_IMGPROC_HPP
#define SDORB_REF_COMPAT_IMGPROC_HPP
#include "../core/core.hpp"

namespace cv {

/* cv::resize: dst.create(dsize) keeps a destination view of the right size in place (src/ORBextractor.cc:690 resizes into
 * the ROI of the padded level buffer); 8-bit INTER_LINEAR is the 11-bit fixed-point kernel of OpenCV 4.x. */
static inline void resize(InputArray _src, OutputArray _dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR) {
  Mat src = _src.getMat();
  SDORB_CV_ASSERT(!src.empty() && src.type() == CV_8UC1 && interpolation == INTER_LINEAR);
  if (dsize.width == 0 || dsize.height == 0) {
    SDORB_CV_ASSERT(fx > 0 && fy > 0);
    dsize = Size(saturate_cast<int>(src.cols * fx), saturate_cast<int>(src.rows * fy));
  }
  SDORB_CV_ASSERT(dsize.width > 0 && dsize.height > 0);
  _dst.create(dsize.height, dsize.width, src.type());
  Mat dst = _dst.getMat();
  SDORB_CV_ASSERT(dst.data != src.data);
  if (dsize.width == src.cols && dsize.height == src.rows) {  /* OpenCV: plain copy */
    src.copyTo(dst);
    return;
  }
  orc_resize_linear_8u(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

/* cv::GaussianBlur on 8U with a 7x7 kernel, sigma 2: the fixed-point separable filter of OpenCV 4.x; in-place calls
 * (src/ORBextractor.cc:660) read from a copy, as OpenCV's filter engine buffers its source rows. */
static inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY = 0,
                                int borderType = BORDER_DEFAULT) {
  Mat src = _src.getMat();
  SDORB_CV_ASSERT(src.type() == CV_8UC1);
  if (sigmaY <= 0) sigmaY = sigmaX;
  SDORB_CV_ASSERT(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2);
  SDORB_CV_ASSERT((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
  SDORB_CV_ASSERT(!src.isSubmatrix() || (borderType & BORDER_ISOLATED));  /* a view would read its real surroundings */
  _dst.create(src.rows, src.cols, src.type());
  Mat dst = _dst.getMat();
  if (src.empty()) return;
  Mat in = (dst.data == src.data) ? src.clone() : src;
  orc_gaussian_blur_7x7_s2(in.data, in.cols, in.rows, in.step, dst.data, dst.step);
}

/* cv::undistortPoints(src, dst, K, dist, R = noArray(), P): src / dst are N x 1 CV_32FC2 point lists (src/Frame.cc:353-356, 384-387
 * pass K as P and no R).  Forwards to the oracle's restatement of OpenCV 4.13's iteration, which tests/test_oracle_primitives.py
 * pins against the real cv2.undistortPoints. */
static inline void undistortPoints(InputArray _src, OutputArray _dst, InputArray _K, InputArray _dist, InputArray _R, InputArray _P) {
  Mat src = _src.getMat(), K = _K.getMat(), dist = _dist.getMat(), P = _P.getMat();
  SDORB_CV_ASSERT(_R.empty() && src.type() == CV_32FC2 && src.isContinuous() && (src.cols == 1 || src.rows == 1));
  SDORB_CV_ASSERT(K.type() == CV_32F && K.rows == 3 && K.cols == 3 && P.type() == CV_32F && P.rows == 3 && P.cols == 3);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) SDORB_CV_ASSERT(K.at<float>(i, j) == P.at<float>(i, j));
  SDORB_CV_ASSERT(dist.type() == CV_32F && (dist.rows == 1 || dist.cols == 1));
  const int n = (int)src.total(), nd = (int)dist.total();
  const float k4[4] = {K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2)};
  std::vector<float> d((size_t)nd);
  for (int i = 0; i < nd; ++i) d[(size_t)i] = dist.at<float>(i);
  std::vector<orc_keypoint> in((size_t)n), out((size_t)n);
  const float* sp = src.ptr<float>();
  for (int i = 0; i < n; ++i) {
    in[(size_t)i].x = sp[2 * i];
    in[(size_t)i].y = sp[2 * i + 1];
  }
  if (n > 0) orc_undistort_keypoints(&in[0], n, k4, nd ? &d[0] : 0, nd, &out[0]);
  _dst.create(src.rows, src.cols, src.type());
  Mat dst = _dst.getMat();
  float* dp = dst.ptr<float>();
  for (int i = 0; i < n; ++i) {
    dp[2 * i] = out[(size_t)i].x;
    dp[2 * i + 1] = out[(size_t)i].y;
  }
}

/* cv::undistort: declared for src/Frame.cc:433-436 (Frame::Undistort, an image warp outside this path); never executed here. */
static inline void undistort(InputArray, OutputArray, InputArray, InputArray) { throw Exception("cv::undistort is outside the parity library"); }

}  // namespace cv
#endif
