// ASan + UBSan run of the CPU oracle (SURVEY.md section 4, row "Sanitizers"): compiled together with oracle/sdorb_oracle.cc by
// tests/test_oracle_primitives.py::test_oracle_under_asan_ubsan with -fsanitize=address,undefined -fno-sanitize-recover=all.
// Exercises the extractor in both modes (incl. the reference defaults whose last level has a negative cell height, a zero-corner
// image and a tiny image), the primitives, and every matcher on random keyframes.  Exit code 0 = no report.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../oracle/sdorb_oracle.h"

static uint32_t rng_state = 12345u;
static uint32_t rnd() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return rng_state >> 8;
}
static float frand(float lo, float hi) { return lo + (hi - lo) * (float)(rnd() & 0xFFFF) / 65535.0f; }

static std::vector<uint8_t> image(int w, int h, int kind) {
  std::vector<uint8_t> im((size_t)w * h);
  std::vector<uint8_t> coarse((size_t)(w / 8 + 2) * (h / 8 + 2));
  for (auto& c : coarse) c = (uint8_t)(rnd() & 255);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int v = kind == 1 ? 128 : coarse[(size_t)(y / 8) * (w / 8 + 2) + x / 8] + (int)(rnd() % 41) - 20;
      im[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
  return im;
}

static long extract(const orc_params& p, int min_th, int w, int h, int kind, std::vector<orc_keypoint>* kps_out = nullptr,
                    std::vector<uint8_t>* desc_out = nullptr) {
  orc_extractor* e = orc_create(&p);
  if (min_th >= 0) orc_set_orbslam2_mode(e, p.th_fast, min_th);
  const int cap = 4 * p.nfeatures + 64;
  std::vector<orc_keypoint> kps(cap);
  std::vector<uint8_t> desc((size_t)cap * 32);
  const std::vector<uint8_t> im = image(w, h, kind);
  const int n = orc_extract(e, im.data(), w, h, (size_t)w, kps.data(), desc.data(), cap, nullptr);
  orc_destroy(e);
  if (n > 0 && kps_out) {
    kps.resize(n);
    desc.resize((size_t)n * 32);
    *kps_out = kps;
    *desc_out = desc;
  }
  return n;
}

int main() {
  long total = 0;
  const orc_params c1 = {1000, 1.2f, 8, 20}, c0 = {1000, 2.0f, 5, 20}, small = {50, 1.2f, 8, 20};
  std::vector<orc_keypoint> k1, k2;
  std::vector<uint8_t> d1, d2;
  total += extract(c1, -1, 640, 480, 0, &k1, &d1);
  total += extract(c1, -1, 333, 257, 0, &k2, &d2);
  total += extract(c0, -1, 640, 480, 0);   // level 4 is 40 x 30: negative cell height
  total += extract(c1, -1, 320, 240, 1);   // no corners
  total += extract(small, -1, 97, 83, 0);
  total += extract(c1, 7, 640, 480, 0);    // ORB-SLAM2-style mode
  total += extract(c1, 7, 752, 480, 1);
  if (k1.empty() || k2.empty()) return 2;
  const int n1 = (int)k1.size(), n2 = (int)k2.size();

  // Hamming / best-two / greedy / distinctive
  std::vector<orc_match> m(n1);
  orc_match_best2(d1.data(), n1, d2.data(), n2, 0.75f, 50, m.data());
  orc_match_greedy(d1.data(), n1, d2.data(), n2, 0.75f, 50, m.data());
  orc_match_best2(d1.data(), n1, d2.data(), 0, 0.75f, 50, m.data());
  std::vector<uint16_t> hm((size_t)100 * 90);
  orc_hamming_matrix(d1.data(), 100, d2.data(), 90, hm.data());
  int med = 0;
  total += orc_distinctive(d1.data(), 17, &med) + orc_distinctive(d1.data(), 0, &med);

  // grids and every search routine, on the extractor's own keypoints
  const float gp[4] = {0.f, 0.f, 64.f / 640.f, 48.f / 480.f};
  std::vector<int32_t> cs1(64 * 48 + 1), ix1(n1), cs2(64 * 48 + 1), ix2(n2);
  orc_assign_grid(k1.data(), n1, gp[0], gp[1], gp[2], gp[3], cs1.data(), ix1.data());
  orc_assign_grid(k2.data(), n2, gp[0], gp[1], gp[2], gp[3], cs2.data(), ix2.data());
  const orc_frame_grid g1 = {cs1.data(), ix1.data(), gp[0], gp[1], gp[2], gp[3]}, g2 = {cs2.data(), ix2.data(), gp[0], gp[1], gp[2], gp[3]};
  float sf[8], inv[8], sig[8];
  sf[0] = 1.f;
  for (int l = 1; l < 8; l++) sf[l] = sf[l - 1] * 1.2f;
  for (int l = 0; l < 8; l++) sig[l] = sf[l] * sf[l], inv[l] = 1.f / sig[l];
  std::vector<int32_t> area(n2);
  for (int q = 0; q < 200; q++)
    total += orc_features_in_area(k2.data(), &g2, frand(-100, 800), frand(-100, 600), frand(0.5f, 400), (int)(rnd() % 4) - 1, (int)(rnd() % 9) - 1,
                                  area.data());
  std::vector<float> prev((size_t)2 * n1), proj((size_t)3 * n1), vcos(n1), ur2(n2), ur1(n1);
  std::vector<int32_t> level(n1), m12(n1), a2(n2), a1(n1), bi(n1), bd(n1), m1(n1), m2(n2);
  std::vector<uint8_t> flags(n1), occ2(n2), flags2(n2), v1(n1), v2(n2);
  std::vector<float> proj2((size_t)3 * n2);
  std::vector<int32_t> level2(n2);
  for (int i = 0; i < n1; i++) {
    prev[2 * i] = k1[i].x, prev[2 * i + 1] = k1[i].y;
    proj[3 * i] = k1[i].x * 0.52f + frand(-3, 3), proj[3 * i + 1] = k1[i].y * 0.53f + frand(-3, 3), proj[3 * i + 2] = frand(0.001f, 1.f);
    vcos[i] = frand(0.5f, 1.f), level[i] = (int)(rnd() % 8), flags[i] = (uint8_t)(rnd() & 3), ur1[i] = (rnd() & 1) ? k1[i].x - 5.f : -1.f;
    v1[i] = (uint8_t)(rnd() & 1);
  }
  for (int i = 0; i < n2; i++) {
    ur2[i] = (rnd() & 1) ? k2[i].x - 5.f : -1.f, occ2[i] = (uint8_t)((rnd() & 7) == 0), flags2[i] = (uint8_t)(rnd() & 3);
    proj2[3 * i] = k2[i].x * 1.9f + frand(-3, 3), proj2[3 * i + 1] = k2[i].y * 1.85f + frand(-3, 3), proj2[3 * i + 2] = 0.f;
    level2[i] = (int)(rnd() % 8), v2[i] = (uint8_t)(rnd() & 1);
  }
  const float bounds[4] = {0.f, 640.f, 0.f, 480.f};
  total += orc_search_for_initialization(k1.data(), d1.data(), n1, k2.data(), d2.data(), n2, &g2, prev.data(), 100, 0.9f, 1, m12.data());
  for (int mode = 0; mode < 3; mode++)
    total += orc_search_by_projection(k1.data(), k1.data(), proj.data(), flags.data(), d1.data(), n1, k2.data(), d2.data(), ur2.data(), occ2.data(),
                                      n2, &g2, sf, bounds, 15.f, 40.f, mode, 1, a2.data(), mode == 2 ? 64 : 0);
  total += orc_search_map_points(proj.data(), vcos.data(), level.data(), flags.data(), d1.data(), n1, k2.data(), d2.data(), ur2.data(), occ2.data(),
                                 n2, &g2, sf, 3.f, 0.8f, a2.data());
  total += orc_search_by_points(k1.data(), d1.data(), v1.data(), n1, k2.data(), d2.data(), v2.data(), n2, 0.75f, 1, m12.data());
  orc_fuse_search(proj.data(), level.data(), flags.data(), d1.data(), n1, k2.data(), d2.data(), ur2.data(), &g2, sf, inv, 3.f, 1, 50, bi.data(),
                  bd.data());
  orc_fuse_search(proj.data(), level.data(), flags.data(), d1.data(), n1, k2.data(), d2.data(), nullptr, &g2, sf, nullptr, 4.f, 0, 50, bi.data(),
                  bd.data());
  total += orc_search_by_sim3(proj.data(), level.data(), flags.data(), d1.data(), n1, proj2.data(), level2.data(), flags2.data(), d2.data(), n2,
                              k1.data(), d1.data(), &g1, k2.data(), d2.data(), &g2, sf, sf, 7.5f, m1.data(), m2.data(), m12.data());
  total += orc_search_by_projection_sim3(proj.data(), level.data(), flags.data(), d1.data(), n1, k2.data(), d2.data(), occ2.data(), n2, &g2, sf, 10,
                                         a2.data());
  const double F12[9] = {0, -1e-3, 0.2, 1e-3, 0, -0.3, -0.2, 0.3, 0};
  const int nt1 = n1 < 300 ? n1 : 300, nt2 = n2 < 300 ? n2 : 300;
  total += orc_search_for_triangulation(k1.data(), d1.data(), v1.data(), ur1.data(), nt1, k2.data(), d2.data(), v2.data(), ur2.data(), nt2, F12, 320.f,
                                        240.f, sf, sig, 1, m12.data());
  // empty sides
  total += orc_search_by_points(k1.data(), d1.data(), v1.data(), 0, k2.data(), d2.data(), v2.data(), n2, 0.75f, 1, m12.data());
  total += orc_search_by_sim3(proj.data(), level.data(), flags.data(), d1.data(), 0, proj2.data(), level2.data(), flags2.data(), d2.data(), n2,
                              k1.data(), d1.data(), &g1, k2.data(), d2.data(), &g2, sf, sf, 7.5f, m1.data(), m2.data(), m12.data());
  // Frame post-processing
  const float K[4] = {517.3f, 516.5f, 318.6f, 255.3f}, dist[5] = {0.2624f, -0.9531f, -0.0054f, 0.0026f, 1.1633f};
  std::vector<orc_keypoint> kun(n1);
  orc_undistort_keypoints(k1.data(), n1, K, dist, 5, kun.data());
  float b4[4];
  orc_image_bounds(640, 480, K, dist, 5, b4);
  std::vector<float> depth((size_t)640 * 480, 2.5f), ur(n1), z(n1);
  orc_stereo_from_rgbd(k1.data(), kun.data(), n1, depth.data(), 640, 40.f, ur.data(), z.data());
  std::printf("oracle sanitize run ok: %ld\n", total);
  return 0;
}
