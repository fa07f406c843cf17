// Exercises include/ORBextractor.h (the C++ drop-in for SD_SLAM::ORBextractor) the way Frame / Tracking do
// (/root/reference/src/Tracking.cc:98, src/Frame.cc:78-84,195) and dumps the results for tests/test_cpp_shim.py.
//   shim_driver <in.raw> <w> <h> <nfeatures> <scale> <nlevels> <thFAST> <out.bin> [minThFAST]
// With minThFAST the north-star 5-argument constructor (iniThFAST, minThFAST: the ORB-SLAM2-style mode) is used.
// out.bin: int32 n | n*28 B keypoints | n*32 B descriptors | int32 nlevels | per level: int32 w, h, step-padded (w+38)*(h+38) bytes
//          | nlevels*4 floats of the getters | int32 distance(d0,d1) | 4*int32 best-two of row 0 against all rows
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ORBextractor.h"

int main(int argc, char** argv) {
  if (argc != 9 && argc != 10) return 2;
  const int w = atoi(argv[2]), h = atoi(argv[3]);
  std::vector<unsigned char> raw((size_t)w * h);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(raw.data(), 1, raw.size(), f) != raw.size()) return 3;
  fclose(f);
  try {
    SD_SLAM::ORBextractor* mpORBextractorLeft =
        argc == 10 ? new SD_SLAM::ORBextractor(atoi(argv[4]), (float)atof(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[9]))
                   : new SD_SLAM::ORBextractor(atoi(argv[4]), (float)atof(argv[5]), atoi(argv[6]), atoi(argv[7]));
    cv::Mat im(h, w, CV_8UC1, raw.data(), (size_t)w);
    std::vector<cv::KeyPoint> mvKeys;
    cv::Mat mDescriptors;
    std::vector<cv::Mat> mvImagePyramid;
    (*mpORBextractorLeft)(cv::Mat(), cv::Mat(), mvKeys, mDescriptors, mvImagePyramid);  // empty image: nothing happens
    if (!mvKeys.empty() || !mvImagePyramid.empty()) return 4;
    (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors, mvImagePyramid);  // src/Frame.cc:195
    std::vector<cv::KeyPoint> k2;
    cv::Mat d2;
    (*mpORBextractorLeft)(im, cv::Mat(), k2, d2);  // north-star 4-argument form
    if (k2.size() != mvKeys.size()) return 5;
    FILE* o = fopen(argv[8], "wb");
    int n = (int)mvKeys.size();
    fwrite(&n, 4, 1, o);
    fwrite(mvKeys.data(), sizeof(cv::KeyPoint), n, o);
    for (int i = 0; i < n; ++i) fwrite(mDescriptors.ptr(i), 1, 32, o);
    int nl = mpORBextractorLeft->GetLevels();
    fwrite(&nl, 4, 1, o);
    for (int l = 0; l < nl; ++l) {
      const cv::Mat& m = mvImagePyramid[l];
      fwrite(&m.cols, 4, 1, o);
      fwrite(&m.rows, 4, 1, o);
      const unsigned char* base = m.data - 19 * m.step - 19;  // the padded parent buffer
      for (int y = 0; y < m.rows + 38; ++y) fwrite(base + (size_t)y * m.step, 1, m.cols + 38, o);
    }
    std::vector<float> a = mpORBextractorLeft->GetScaleFactors(), b = mpORBextractorLeft->GetInverseScaleFactors(),
                       c = mpORBextractorLeft->GetScaleSigmaSquares(), d = mpORBextractorLeft->GetInverseScaleSigmaSquares();
    fwrite(a.data(), 4, nl, o);
    fwrite(b.data(), 4, nl, o);
    fwrite(c.data(), 4, nl, o);
    fwrite(d.data(), 4, nl, o);
    SD_SLAM::ORBdistance dist(*mpORBextractorLeft);
    int dd = n >= 2 ? dist.DescriptorDistance(mDescriptors(cv::Rect(0, 0, 32, 1)), mDescriptors(cv::Rect(0, 1, 32, 1))) : -1;
    fwrite(&dd, 4, 1, o);
    std::vector<sdorb_match> best;
    if (n) dist.BestTwo(mDescriptors(cv::Rect(0, 0, 32, 1)), mDescriptors, 0.75f, SD_SLAM::ORBdistance::TH_LOW, best);
    sdorb_match m0 = n ? best[0] : sdorb_match{-1, 256, 256, 0};
    fwrite(&m0, sizeof(m0), 1, o);
    fclose(o);
    delete mpORBextractorLeft;
  } catch (const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 10;
  }
  return 0;
}
