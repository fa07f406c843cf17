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

}  // namespace cv
#endif
