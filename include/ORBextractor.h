/**
 * ORBextractor.h -- drop-in replacement of SD-SLAM's src/ORBextractor.h + src/ORBextractor.cc on top of libsdorb.so.
 *
 * Same namespace, class name, constructor, operator() and getters as the reference
 * (/root/reference/src/ORBextractor.h:34-90), so Frame (src/Frame.cc:78-84, 132-138, 195) and Tracking
 * (src/Tracking.cc:98,101) compile and behave unchanged: put this header in place of the reference's, drop
 * src/ORBextractor.cc from the build and link libsdorb.so.  All pixel work happens on the GPU behind
 * the C ABI of include/sdorb.h; this header only converts between cv:: types and plain buffers.
 *
 * Results are bit-identical to the reference built against OpenCV 4.13 (keypoints incl. order, angles, descriptors,
 * pyramid).  Differences a caller can observe:
 *   - a CUDA failure throws std::runtime_error (the reference cannot fail that way); there is no CPU fallback;
 *   - image sizes for which the reference itself throws cv::Exception (a cell ROI outside the level image) throw
 *     std::runtime_error("... geometry ...") instead.
 * The north-star surface ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) / operator()(image, mask,
 * keypoints, descriptors) is provided as overloads.  SD-SLAM itself has a single FAST threshold and no quadtree
 * (src/ORBextractor.cc:536): the 5-argument constructor selects the ORB-SLAM2-style mode of the
 * library (30-pixel cells with the iniThFAST / minThFAST fallback, DistributeOctTree), whose
 * results follow the public ORB-SLAM2 algorithm, not anything in /root/reference; operator() may then return a few
 * keypoints more than nfeatures, as ORB-SLAM2 does.
 */
#ifndef SD_SLAM_ORBEXTRACTOR_H
#define SD_SLAM_ORBEXTRACTOR_H

#include <cstring>
#include <list>
#include <stdexcept>
#include <string>
#include <vector>

#include <opencv2/core/core.hpp>

#include "sdorb.h"

namespace SD_SLAM {

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };  // src/ORBextractor.h:36

  // Resources of the GPU handle; not part of the reference surface, defaulted so that reference call sites compile.
  struct Options {
    int max_width, max_height;  // largest image this extractor will be given (scratch is sized for the size actually seen)
    int device;                 // CUDA device ordinal, -1 = current
    Options() : max_width(4095), max_height(4095), device(-1) {}
  };

  // src/ORBextractor.h:38, src/ORBextractor.cc:406-457
  ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _thFAST, const Options& opt = Options())
      : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), thFAST(_thFAST), handle_(NULL) {
    Init(-1, opt);
  }
  // north-star form: (nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) -- the ORB-SLAM2-style mode
  ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int iniThFAST, int minThFAST, const Options& opt = Options())
      : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), thFAST(iniThFAST), handle_(NULL) {
    Init(minThFAST < 0 ? 0 : minThFAST, opt);
  }

 protected:
  void Init(int min_th_fast, const Options& opt) {
    const int _nfeatures = nfeatures, _nlevels = nlevels, _thFAST = thFAST;
    const float _scaleFactor = scaleFactor;
    const int max_width = opt.max_width, max_height = opt.max_height, device = opt.device;
    sdorb_params p;
    p.nfeatures = _nfeatures;
    p.scale_factor = _scaleFactor;
    p.nlevels = _nlevels;
    p.th_fast = _thFAST;
    p.min_th_fast = min_th_fast;
    p.device = device;
    p.max_width = max_width;
    p.max_height = max_height;
    p.max_batch = 1;  // Frame hands over one image at a time (src/Frame.cc:195)
    Check(sdorb_create(&p, &handle_), "sdorb_create");
    mvScaleFactor.resize(nlevels);
    mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    Check(sdorb_get_tables(handle_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                           mvInvLevelSigma2.data(), mnFeaturesPerLevel.data()),
          "sdorb_get_tables");
    capacity_ = sdorb_max_keypoints(handle_);
    if (capacity_ < 1) capacity_ = 1;
    kps_.resize(capacity_);
    desc_.resize((size_t)capacity_ * 32);
  }

 public:
  ~ORBextractor() { sdorb_destroy(handle_); }
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // src/ORBextractor.h:45-46, src/ORBextractor.cc:620-678.  Mask is ignored, as in the reference.
  void operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& _keypoints,
                  cv::OutputArray _descriptors, std::vector<cv::Mat>& imagePyramid) {
    Run(_image, _keypoints, _descriptors, &imagePyramid);
  }
  // north-star form without the pyramid out-parameter
  void operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& _keypoints,
                  cv::OutputArray _descriptors) {
    Run(_image, _keypoints, _descriptors, NULL);
  }

  int inline GetLevels() { return nlevels; }
  float inline GetScaleFactor() { return scaleFactor; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  sdorb_handle* handle() { return handle_; }  // for the batched entry points of sdorb.h

 protected:
  static const int EDGE_THRESHOLD = 19;  // src/ORBextractor.cc:75

  void Check(int rc, const char* what) {
    if (rc == SDORB_OK) return;
    std::string msg = std::string("ORBextractor (libsdorb): ") + what + ": " + sdorb_strerror(rc);
    if (rc == SDORB_ERR_CUDA && handle_) msg += std::string(" [") + sdorb_last_cuda_error(handle_) + "]";
    throw std::runtime_error(msg);
  }

  void Run(cv::InputArray _image, std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors,
           std::vector<cv::Mat>* imagePyramid) {
    if (_image.empty()) return;  // src/ORBextractor.cc:622-623: outputs untouched
    cv::Mat image = _image.getMat();
    CV_Assert(image.type() == CV_8UC1);  // src/ORBextractor.cc:626

    // imagePyramid[l] is a view, at (19,19), of a buffer padded by 19 px of BORDER_REFLECT_101 (src/ORBextractor.cc:684-696)
    std::vector<sdorb_pyr_view> views;
    if (imagePyramid) {
      imagePyramid->resize(nlevels);
      views.resize(nlevels);
      for (int level = 0; level < nlevels; ++level) {
        int w = 0, h = 0;
        Check(sdorb_level_size(handle_, image.cols, image.rows, level, &w, &h), "sdorb_level_size");
        if (w <= 0 || h <= 0) throw std::runtime_error("ORBextractor (libsdorb): empty pyramid level (geometry)");
        cv::Mat temp(h + EDGE_THRESHOLD * 2, w + EDGE_THRESHOLD * 2, CV_8UC1);
        (*imagePyramid)[level] = temp(cv::Rect(EDGE_THRESHOLD, EDGE_THRESHOLD, w, h));
        cv::Mat& lvl = (*imagePyramid)[level];
        views[level].data = lvl.data;
        views[level].width = w;
        views[level].height = h;
        views[level].stride = (size_t)lvl.step;
        views[level].border = EDGE_THRESHOLD;
      }
    }
    int n = 0;
    Check(sdorb_extract(handle_, image.data, image.cols, image.rows, (size_t)image.step, kps_.data(), desc_.data(), capacity_,
                        &n, imagePyramid ? views.data() : NULL),
          "sdorb_extract");
    _keypoints.clear();
    if (n == 0) {
      _descriptors.release();  // src/ORBextractor.cc:640-641
      return;
    }
    _descriptors.create(n, 32, CV_8U);  // src/ORBextractor.cc:643
    cv::Mat descriptors = _descriptors.getMat();
    _keypoints.reserve(n);
    for (int i = 0; i < n; ++i) {
      const sdorb_keypoint& k = kps_[i];
      _keypoints.push_back(cv::KeyPoint(k.x, k.y, k.size, k.angle, k.response, k.octave, k.class_id));
      std::memcpy(descriptors.ptr(i), &desc_[(size_t)i * 32], 32);
    }
  }

  // members of the reference class kept by name (src/ORBextractor.h:72-89)
  int nfeatures;
  double scaleFactor;
  int nlevels;
  int thFAST;
  std::vector<int> mnFeaturesPerLevel;
  std::vector<float> mvScaleFactor;
  std::vector<float> mvInvScaleFactor;
  std::vector<float> mvLevelSigma2;
  std::vector<float> mvInvLevelSigma2;

 private:
  sdorb_handle* handle_;
  int capacity_;
  std::vector<sdorb_keypoint> kps_;
  std::vector<unsigned char> desc_;
};

/**
 * Batched ORBmatcher::DescriptorDistance (src/ORBmatcher.h:43, src/ORBmatcher.cc:1459-1473) in the shape of its hottest
 * caller, the best / second-best scan of ORBmatcher::SearchByPoints (src/ORBmatcher.cc:1239-1265).  The pointer-chasing
 * wrappers of ORBmatcher (MapPoint / KeyFrame graph walks) stay on the host; they hand descriptor matrices to this class
 * instead of looping over DescriptorDistance.
 */
class ORBdistance {
 public:
  static const int TH_HIGH = 100, TH_LOW = 50;  // src/ORBmatcher.cc:36-37

  explicit ORBdistance(ORBextractor& owner) : handle_(owner.handle()) {}

  // DescriptorDistance(a, b) for single 1x32 rows (parity checks; one pair per call is latency, not throughput)
  int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    uint16_t d = 0;
    Check(sdorb_hamming_matrix(handle_, a.data, 1, b.data, 1, &d, SDORB_MEM_HOST, NULL));
    return d;
  }
  // all distances: out is nA x nB, CV_16U
  void DistanceMatrix(const cv::Mat& A, const cv::Mat& B, cv::Mat& out) {
    RequireRows(A);
    RequireRows(B);
    out.create(A.rows, B.rows, CV_16U);
    if (A.rows && B.rows)
      Check(sdorb_hamming_matrix(handle_, A.data, A.rows, B.data, B.rows, reinterpret_cast<uint16_t*>(out.data), SDORB_MEM_HOST, NULL));
  }
  // MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:252-275): row of D (N x 32) with the least median distance
  // to the other rows, first on ties; -1 for an empty matrix
  int Distinctive(const cv::Mat& D, int* median = NULL) {
    RequireRows(D);
    const int32_t offsets[2] = {0, D.rows};
    int32_t best = -1, med = 0;
    const uint8_t dummy[32] = {0};
    Check(sdorb_distinctive_batch(handle_, D.rows ? D.data : dummy, offsets, 1, &best, &med, SDORB_MEM_HOST, NULL));
    if (median) *median = med;
    return best;
  }
  // per row of A: first index of the smallest distance in B, that distance, the second smallest, and the acceptance
  // best < th_low && best < ratio * second.  greedy = SearchByPoints' vbMatched2 rule (src/ORBmatcher.cc:1228-1270).
  void BestTwo(const cv::Mat& A, const cv::Mat& B, float ratio, int th_low, std::vector<sdorb_match>& out, bool greedy = false) {
    RequireRows(A);
    RequireRows(B);
    out.resize(A.rows);
    if (!A.rows) return;
    const int32_t nA = A.rows, nB = B.rows;
    const uint8_t dummy[32] = {0};
    const uint8_t* pb = nB ? B.data : dummy;
    const int strideB = nB ? nB : 1;
    Check(greedy ? sdorb_match_greedy_batch(handle_, A.data, &nA, nA, pb, &nB, strideB, 1, ratio, th_low, out.data(), SDORB_MEM_HOST, NULL)
                 : sdorb_match_batch(handle_, A.data, &nA, nA, pb, &nB, strideB, 1, ratio, th_low, out.data(), SDORB_MEM_HOST, NULL));
  }

 private:
  static void RequireRows(const cv::Mat& m) {
    if (!m.empty() && (m.type() != CV_8UC1 || m.cols != 32 || !m.isContinuous()))
      throw std::runtime_error("ORBdistance: descriptors must be a continuous N x 32 CV_8U matrix");
  }
  static void Check(int rc) {
    if (rc != SDORB_OK) throw std::runtime_error(std::string("ORBdistance (libsdorb): ") + sdorb_strerror(rc));
  }
  sdorb_handle* handle_;
};

}  // namespace SD_SLAM

#endif  // SD_SLAM_ORBEXTRACTOR_H
