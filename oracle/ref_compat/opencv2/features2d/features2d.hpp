/* oracle/ref_compat: cv::FAST and cv::KeyPointsFilter::retainBest -- see core/core.hpp. */
#ifndef SDORB_REF_COMPAT_FEATURES2D_HPP
#define SDORB_REF_COMPAT_FEATURES2D_HPP
#include "../core/core.hpp"

namespace cv {

/* cv::FAST(image, keypoints, threshold, nonmaxSuppression) = FAST-9/16 (FastFeatureDetector::TYPE_9_16): keypoints in
 * row-major order with size 7, angle -1, response = corner score, coordinates relative to the (view's) origin. */
static inline void FAST(InputArray _img, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
  Mat img = _img.getMat();
  keypoints.clear();
  if (img.empty()) return;
  SDORB_CV_ASSERT(img.type() == CV_8UC1);
  static_assert(sizeof(KeyPoint) == sizeof(orc_keypoint), "cv::KeyPoint layout");
  keypoints.resize(1024);
  int n = orc_fast(img.data, img.cols, img.rows, img.step, threshold, nonmaxSuppression ? 1 : 0, reinterpret_cast<orc_keypoint*>(&keypoints[0]),
                   (int)keypoints.size());
  if (n > (int)keypoints.size()) {  /* more corners than the first guess: run again with room for all of them */
    keypoints.resize((size_t)n);
    n = orc_fast(img.data, img.cols, img.rows, img.step, threshold, nonmaxSuppression ? 1 : 0, reinterpret_cast<orc_keypoint*>(&keypoints[0]), n);
  }
  keypoints.resize((size_t)(n > 0 ? n : 0));
}

class KeyPointsFilter {
 public:
  /* features2d/src/keypoint.cpp: nth_element(begin, begin + n - 1, end, response >), then keep everything that ties with
   * the n-th response (partition of the tail) -- on the real libstdc++ algorithms, inside the oracle library. */
  static void retainBest(std::vector<KeyPoint>& keypoints, int npoints) {
    if (npoints >= 0 && keypoints.size() > (size_t)npoints) {
      if (npoints == 0) {
        keypoints.clear();
        return;
      }
      const int n = orc_retain_best(reinterpret_cast<orc_keypoint*>(&keypoints[0]), (int)keypoints.size(), npoints);
      keypoints.resize((size_t)n);
    }
  }
};

}  // namespace cv
#endif
