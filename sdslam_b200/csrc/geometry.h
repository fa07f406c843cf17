// geometry.h -- host-side tables of the extractor (see geometry.cc).
#pragma once
#include <vector>

#include "sdorb_internal.h"

namespace sdorb {

struct Tables {
  int nlevels = 0;
  std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
  std::vector<int> n_per_level;
  int umax[SDORB_HALF_PATCH + 1] = {0};
};

void build_tables(int nfeatures, float scale_factor, int nlevels, Tables* t);
void level_size(const Tables& t, int level, int width, int height, int* lw, int* lh);
// Returns 0, or the negated SDORB_ERR_* magnitude (-1 bad arg, -4 geometry, -7 unsupported).
int build_frame_geom(const Tables& t, int nfeatures, int th_fast, int width, int height, FrameGeom* g,
                     std::vector<ResizeTap>* taps, std::vector<ResizeGroup>* groups = nullptr, int min_th_fast = -1);
// ORB-SLAM2-style mode: keypoint slots of one level (DistributeOctTree stops at >= N nodes, so up to N + 2, or the
// 4 children of each of the n_ini initial nodes).
int octree_level_slots(int n_desired, int n_ini);

}  // namespace sdorb
