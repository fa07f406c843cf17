/* oracle/ref_compat/Converter.h -- stands in for /root/reference/src/Converter.h when the reference's Frame.cc is compiled for
 * the parity library (oracle/ref_build/Makefile): the real header pulls in g2o (Eigen template code this image cannot build),
 * and Frame.cc only needs Converter::toCvMat(Eigen::Matrix3d) = a 3x3 CV_32F copy of the matrix (src/Converter.cc).
 * TEST INFRASTRUCTURE ONLY. */
#ifndef SD_SLAM_CONVERTER_H
#define SD_SLAM_CONVERTER_H
#include <opencv2/core/core.hpp>
#include <Eigen/Dense>

namespace SD_SLAM {
class Converter {
 public:
  static cv::Mat toCvMat(const Eigen::Matrix3d& m) {
    cv::Mat out(3, 3, CV_32F);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) out.at<float>(i, j) = (float)m(i, j);
    return out;
  }
};
}  // namespace SD_SLAM
#endif
