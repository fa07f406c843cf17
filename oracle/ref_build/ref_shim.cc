// oracle/ref_build/ref_shim.cc -- C interface around the UNMODIFIED reference classes, for tests and the CPU baseline.
//
// TEST INFRASTRUCTURE ONLY.  This file is ours; the code it drives is the reference's own text:
//   /root/reference/src/ORBextractor.{h,cc}   (compiled where it lies, never copied into this repository)
//   /root/reference/src/ORBmatcher.cc (the whole translation unit; DescriptorDistance :1459-1473 is called from here, the
//   search routines from ref_matcher_shim.cc)
// against the cv:: surface of oracle/ref_compat.  Built into oracle/_ref/libsdorb_ref.so by oracle/ref_build/Makefile.
#include <stdint.h>

#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "ORBextractor.h"  // the reference's headers (links in oracle/_ref/gen/src, see the Makefile)
#include "ORBmatcher.h"

namespace {
// reaches the protected tables of the reference class (src/ORBextractor.h:72-89) without touching its text
class Probe : public SD_SLAM::ORBextractor {
 public:
  Probe(int n, float sf, int nl, int th) : SD_SLAM::ORBextractor(n, sf, nl, th) {}
  const std::vector<int>& perLevel() const { return mnFeaturesPerLevel; }
  const std::vector<int>& umaxTable() const { return umax; }
  const std::vector<cv::Point>& patternTable() const { return pattern; }
};

int run_one(Probe* e, const uint8_t* img, int w, int h, size_t step, orc_keypoint* kps, uint8_t* desc, int cap, uint8_t* pyr_tight,
            uint8_t* pyr_padded) {
  try {
    cv::Mat image(h, w, CV_8UC1, const_cast<uint8_t*>(img), step);
    std::vector<cv::KeyPoint> keypoints;
    cv::Mat descriptors;
    std::vector<cv::Mat> pyramid;
    (*e)(image, cv::Mat(), keypoints, descriptors, pyramid);  // src/ORBextractor.cc:620, called as src/Frame.cc:195 does
    const int n = (int)keypoints.size();
    if (n > 0 && (descriptors.rows != n || descriptors.cols != 32)) return -100;
    if (n == 0 && !descriptors.empty()) return -101;
    for (int i = 0; i < n && i < cap; ++i) {
      if (kps) memcpy(&kps[i], &keypoints[i], sizeof(orc_keypoint));
      if (desc) memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
    }
    for (size_t l = 0; l < pyramid.size(); ++l) {
      const cv::Mat& L = pyramid[l];
      if (pyr_tight) {
        for (int y = 0; y < L.rows; ++y) memcpy(pyr_tight + (size_t)y * L.cols, L.ptr(y), (size_t)L.cols);
        pyr_tight += (size_t)L.rows * L.cols;
      }
      if (pyr_padded) {  // the whole buffer imagePyramid[l] is a view of (19 px of border all around, src/ORBextractor.cc:684-686)
        cv::Size whole;
        cv::Point ofs;
        L.locateROI(whole, ofs);
        if (ofs.x != 19 || ofs.y != 19 || whole.width != L.cols + 38 || whole.height != L.rows + 38) return -102;
        const uint8_t* base = L.data - (size_t)19 * L.step - 19;
        for (int y = 0; y < whole.height; ++y) memcpy(pyr_padded + (size_t)y * whole.width, base + (size_t)y * L.step, (size_t)whole.width);
        pyr_padded += (size_t)whole.height * whole.width;
      }
    }
    return n;
  } catch (const cv::Exception&) {
    return -4;  // the reference throws cv::Exception here (a cell ROI outside its level image)
  }
}
}  // namespace

extern "C" {

void* ref_create(int nfeatures, float scale_factor, int nlevels, int th_fast) { return new Probe(nfeatures, scale_factor, nlevels, th_fast); }
void ref_destroy(void* e) { delete static_cast<Probe*>(e); }

// getters of src/ORBextractor.h:48-70 + the protected tables; arrays of nlevels (umax: 16, pattern: 1024 ints) or NULL
void ref_get_tables(void* ev, int* nlevels, float* scale_factor, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                    int* n_per_level, int* umax, int* pattern) {
  Probe* e = static_cast<Probe*>(ev);
  const int nl = e->GetLevels();
  if (nlevels) *nlevels = nl;
  if (scale_factor) *scale_factor = e->GetScaleFactor();
  const std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(),
                           d = e->GetInverseScaleSigmaSquares();
  for (int i = 0; i < nl; ++i) {
    if (scale) scale[i] = a[i];
    if (inv_scale) inv_scale[i] = b[i];
    if (sigma2) sigma2[i] = c[i];
    if (inv_sigma2) inv_sigma2[i] = d[i];
    if (n_per_level) n_per_level[i] = e->perLevel()[i];
  }
  if (umax)
    for (size_t i = 0; i < e->umaxTable().size() && i < 16; ++i) umax[i] = e->umaxTable()[i];
  if (pattern)
    for (size_t i = 0; i < e->patternTable().size() && i < 512; ++i) {
      pattern[2 * i] = e->patternTable()[i].x;
      pattern[2 * i + 1] = e->patternTable()[i].y;
    }
}

// ORBextractor::operator() on one frame.  Returns the keypoint count (at most cap written), -4 where the reference throws
// cv::Exception.  pyr_tight: levels packed one after another (w*h each); pyr_padded: the padded buffers ((w+38)*(h+38) each).
int ref_extract(void* e, const uint8_t* img, int w, int h, size_t step, orc_keypoint* kps, uint8_t* desc, int cap, uint8_t* pyr_tight,
                uint8_t* pyr_padded) {
  return run_one(static_cast<Probe*>(e), img, w, h, step, kps, desc, cap, pyr_tight, pyr_padded);
}

// Frame-parallel driver (operator() mutates no members): frames tightly packed; kps / desc / counts may be NULL.
long ref_extract_many(void* ev, const uint8_t* imgs, int nframes, int w, int h, int nthreads, orc_keypoint* kps, uint8_t* desc,
                      int32_t* counts, int cap) {
  Probe* e = static_cast<Probe*>(ev);
  std::atomic<int> next(0);
  std::atomic<long> total(0);
  auto work = [&]() {
    std::vector<orc_keypoint> k((size_t)cap);
    std::vector<uint8_t> d((size_t)cap * 32);
    for (;;) {
      const int f = next.fetch_add(1);
      if (f >= nframes) break;
      const int n = run_one(e, imgs + (size_t)f * w * h, w, h, (size_t)w, kps ? kps + (size_t)f * cap : k.data(),
                            desc ? desc + (size_t)f * cap * 32 : d.data(), cap, nullptr, nullptr);
      if (counts) counts[f] = n;
      if (n > 0) total += n;
    }
  };
  if (nthreads <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  return total.load();
}

// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1459-1473) on two 32-byte rows
int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
  cv::Mat ma(1, 32, CV_8UC1, const_cast<uint8_t*>(a)), mb(1, 32, CV_8UC1, const_cast<uint8_t*>(b));
  return SD_SLAM::ORBmatcher::DescriptorDistance(ma, mb);
}
void ref_hamming_matrix(const uint8_t* A, int nA, const uint8_t* B, int nB, uint16_t* out) {
  for (int i = 0; i < nA; ++i)
    for (int j = 0; j < nB; ++j) out[(size_t)i * nB + j] = (uint16_t)ref_descriptor_distance(A + (size_t)i * 32, B + (size_t)j * 32);
}

}  // extern "C"
