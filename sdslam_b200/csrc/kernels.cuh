// kernels.cuh -- launchers of the sm_100a kernels of libsdorb (one family per step of
// ORBextractor::operator(), /root/reference/src/ORBextractor.cc:620-678, plus the batched DescriptorDistance).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "sdorb_internal.h"

namespace sdorb {

// ---- programmatic dependent launch (sm_90+): the kernels of a pass form one dependency chain on one stream.  Each is launched
// with cudaLaunchAttributeProgrammaticStreamSerialization and begins with pdl_wait() -- griddepcontrol.wait blocks until the
// preceding grid has completed and flushed its writes, so no kernel touches memory earlier than it would have without the
// attribute -- followed by pdl_trigger() (griddepcontrol.launch_dependents): once every CTA of a grid has started, the CTAs of
// the next grid may take the SM slots its last wave leaves free and sit at their own pdl_wait().  What is saved is the launch
// latency and CTA ramp-up at each of the 12 kernel boundaries of a pass (it matters for small passes and the single-frame call).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_wait();
  pdl_trigger();
}

// Set by the host runtime per pass (thread-local: one handle per thread).  Measured on B200 (profiles/r2_pdl_probe.log): with 512-
// and 2048-frame passes the attribute COSTS 3-4 % -- the kernels run tens to hundreds of waves, nothing is gained at the
// boundaries and every CTA pays for its griddepcontrol.wait -- and the single-frame call, whose kernels are replayed from CUDA
// graphs, does not change.  So it is off unless SDORB_PDL_MAX_FRAMES asks for it; the evidence stays with the switch.
extern thread_local bool g_pdl_enabled;

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_pdl_enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// Where the planes of a batch live.  Level 0 is the caller's image (or a staged copy); levels >= 1 and all
// blurred levels are level-major scratch: level l of frame f starts at base + lv[l].plane_base * batch_cap + f * lv[l].plane_bytes.
struct BatchPlanes {
  const uint8_t* img0;       // level 0, frame 0
  int64_t img0_frame_stride; // bytes between frames of level 0
  int img0_pitch;            // bytes between rows of level 0
  uint8_t* pyr;              // scratch: unblurred levels (slot 0 = staged level 0 when used)
  uint8_t* blur;             // scratch: blurred levels
  uint8_t* nms;              // scratch: keypoint map (one byte per pixel: FAST score + 1 - th after non-max suppression, else 0)
  int batch_cap;             // frames the scratch was sized for
};

struct SelectBuffers {
  int32_t* cell_seen;   // [batch][cells_total]  FAST keypoints per cell of the last pass (written by the select kernel)
  uint32_t* cell_list;  // [batch][list_total]   overflow scratch for cells too large for shared memory
  uint32_t* sel;        // [batch][sel_total]  selected entries, level-major, in output order
  int32_t* sel_count;   // [batch][nlevels]
  int32_t* error_flag;  // device int: set to SDORB_ERR_OVERFLOW magnitude when a fixed-capacity list overflows
  uint32_t* okeys;      // [batch][list_total]   ORB-SLAM2-style mode: vToDistributeKeys of every level, contiguous
  uint16_t* onode;      // [batch][list_total]   ... and the list position of each keypoint's quadtree node
};

// ComputePyramid: level l from level l-1 for every frame (cv::resize INTER_LINEAR fixed point).
void launch_resize_level(const FrameGeom* d_geom, const FrameGeom& g, int level, const BatchPlanes& p,
                         const ResizeTap* d_taps, const ResizeGroup* d_groups, int nframes, cudaStream_t s);
// The small upper levels (first_level .. nlevels - 1, resize_tail_first_level) of every frame in one launch.
int resize_tail_first_level(const FrameGeom& g);
void launch_resize_tail(const FrameGeom* d_geom, const FrameGeom& g, int first_level, const BatchPlanes& p, const ResizeTap* d_taps,
                        const ResizeGroup* d_groups, int nframes, cudaStream_t s);
// imagePyramid of a batch into the caller's frame-major slab (levels first_level .. nlevels-1); layout: pyramid_layout().
void pyramid_layout(const FrameGeom& g, int64_t* offset, int64_t* frame_bytes);
void launch_pack_pyramid(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, int first_level, uint8_t* dst, int nframes,
                         cudaStream_t s);
// Frame 0's pyramid in the reference's padded form (19 px of BORDER_REFLECT_101 around every level), level after level.
int64_t padded_pyramid_layout(const FrameGeom& g, int64_t* offset);
void launch_pack_padded(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, int first_level, uint8_t* dst,
                        cudaStream_t s);
// cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101), 8-bit fixed point, all levels of all frames in one launch.
void launch_blur_all(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, int nframes, cudaStream_t s);
// cv::FAST(cell, thFAST, nonmax=true) for every cell of every level of every frame, one launch; writes the keypoint
// map BatchPlanes::nms.
void launch_fast_all(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, int nframes, cudaStream_t s);
// Quota redistribution + retainBest per cell + retainBest per level.
void launch_select(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, const SelectBuffers& b, int nframes,
                   cudaStream_t s);
// ORB-SLAM2-style mode (FrameGeom::octree): gather_cells_kernel with the ini / min threshold choice per cell, then
// DistributeOctTree per (level, frame) (kernels_octree.cu).
void launch_octree(const FrameGeom* d_geom, const FrameGeom& g, const SelectBuffers& b, int nframes, cudaStream_t s);
int configure_octree_kernel();
// Test hook: std::nth_element(a, a + nth, a + n, response >) as the selection kernel performs it (one warp).
void launch_debug_nth_element(uint32_t* d_entries, int n, int nth, cudaStream_t s);
// IC_Angle + rBRIEF descriptor + output assembly (coordinate scaling, cv::KeyPoint layout).
void launch_describe(const FrameGeom* d_geom, const FrameGeom& g, const BatchPlanes& p, const SelectBuffers& b,
                     const int* d_umax, void* kps_out, uint8_t* desc_out, int32_t* counts_out, int capacity,
                     int nframes, cudaStream_t s);
// Batched DescriptorDistance with best / second-best tracking.
void launch_match(const uint8_t* A, const int32_t* nA, int strideA, const uint8_t* B, const int32_t* nB, int strideB,
                  int npairs, float ratio, int th_low, void* out, cudaStream_t s);
void launch_match_greedy(const uint8_t* A, const int32_t* nA, int strideA, const uint8_t* B, const int32_t* nB,
                         int strideB, int npairs, float ratio, int th_low, void* out, cudaStream_t s);
void launch_hamming_matrix(const uint8_t* A, int nA, const uint8_t* B, int nB, uint16_t* out, cudaStream_t s);
// MapPoint::ComputeDistinctiveDescriptors for nsets descriptor sets (rows [offsets[s], offsets[s+1]) of desc).
void launch_distinctive(const uint8_t* desc, const int32_t* offsets, int nsets, int32_t* best_idx, int32_t* best_median,
                        cudaStream_t s);

// Frame::AssignFeaturesToGrid / PosInGrid and Frame::ComputeStereoFromRGBD (kernels_frame.cu); device pointers.
void launch_assign_grid(const void* kps, const int32_t* counts, int nframes, int capacity, float min_x, float min_y, float inv_w,
                        float inv_h, int32_t* cell_start, int32_t* indices, cudaStream_t s);
void launch_stereo_rgbd(const void* kps, const void* kps_un, const int32_t* counts, int nframes, int capacity, const float* depth,
                        int width, int height, int64_t row_stride, int64_t frame_stride, float mbf, float* u_right, float* z,
                        cudaStream_t s);
// Frame::UndistortKeyPoints (device) and Frame::ComputeImageBounds (host); K = {fx, fy, cx, cy}, dist = {k1, k2, p1, p2[, k3]}.
void launch_undistort(const void* kps, const int32_t* counts, int nframes, int capacity, const float K[4], const float* dist,
                      int ndist, void* out, cudaStream_t s);
void host_image_bounds(int cols, int rows, const float K[4], const float* dist, int ndist, float bounds[4]);
int configure_frame_kernels();

// Guided matchers (kernels_search.cu); all pointers are device pointers, arrays are [npairs][capacity] slabs.
struct SearchGrid {
  const int32_t* cell_start;  // [npairs][64*48+1]  CSR grid of the searched frame (launch_assign_grid)
  const int32_t* indices;     // [npairs][capacity]
  float min_x, min_y, inv_w, inv_h;  // mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv
};
struct SearchInitArgs {  // ORBmatcher::SearchForInitialization
  const void* kps1;  const uint8_t* desc1;  const int32_t* n1;   // F1: mvKeysUn, mDescriptors, N
  const void* kps2;  const uint8_t* desc2;  const int32_t* n2;   // F2
  SearchGrid grid;       // of F2
  float* prev_matched;   // [npairs][capacity][2], in / out
  int32_t* matches12;    // [npairs][capacity]
  int32_t* nmatches;     // [npairs]
  int capacity, window_size, th_low, check_orientation;
  float nnratio;
};
struct SearchProjArgs {  // ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) from the projection on
  const void* kps_last;  const void* kps_last_un;  const float* proj;  const uint8_t* flags_last;  const uint8_t* desc_mp;
  const int32_t* n_last;
  const void* kps_cur_un;  const uint8_t* desc_cur;  const float* u_right_cur;  const uint8_t* occupied_cur;  const int32_t* n_cur;
  SearchGrid grid;      // of the current frame
  int32_t* assigned;    // [npairs][capacity]
  int32_t* nmatches;    // [npairs]
  int capacity, mode, th_high, check_orientation;
  float th, mbf, min_x, max_x, min_y, max_y;
  float scale_factors[SDORB_MAX_LEVELS];
};
struct SearchTriArgs {  // ORBmatcher::SearchForTriangulation from the epipole on
  const void* kps1;  const uint8_t* desc1;  const uint8_t* has_mp1;  const float* u_right1;  const int32_t* n1;  // KF1
  const void* kps2;  const uint8_t* desc2;  const uint8_t* has_mp2;  const float* u_right2;  const int32_t* n2;  // KF2
  const double* F12;     // [npairs][9] row-major
  const float* epipole;  // [npairs][2]  (ex, ey)
  int32_t* matches12;    // [npairs][capacity]  vMatches12
  int32_t* nmatches;     // [npairs]
  int capacity, th_low, check_orientation;
  float gate[SDORB_MAX_LEVELS];   // smallest float >= 3.84 * mvLevelSigma2[level]
  float eplim[SDORB_MAX_LEVELS];  // 100 * mvScaleFactors[level]
};
struct SearchPointsArgs {  // ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)
  const float* proj;  const float* view_cos;  const int32_t* level;  const uint8_t* flags;  const uint8_t* desc_mp;  const int32_t* n_mp;
  const void* kps;  const uint8_t* desc;  const float* u_right;  const uint8_t* occupied;  const int32_t* n_frame;
  SearchGrid grid;
  int32_t* assigned;  // [nframes][capacity]
  int32_t* nmatches;  // [nframes]
  int capacity, capacity_mp, th_high;
  int sim3_form;  // 1: SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th), src/ORBmatcher.cc:146-254
  float th, nnratio;
  float scale_factors[SDORB_MAX_LEVELS];
};
void launch_search_points(const SearchPointsArgs& a, int nframes, cudaStream_t s);
struct SearchByPointsArgs {  // ORBmatcher::SearchByPoints
  const void* kps1;  const uint8_t* desc1;  const uint8_t* valid1;  const int32_t* n1;
  const void* kps2;  const uint8_t* desc2;  const uint8_t* valid2;  const int32_t* n2;
  int32_t* matches12;  // [npairs][capacity]
  int32_t* nmatches;   // [npairs]
  int capacity, th_low, check_orientation;
  float nnratio;
};
void launch_search_by_points(const SearchByPointsArgs& a, int npairs, cudaStream_t s);
struct FuseSearchArgs {  // the keypoint search of ORBmatcher::Fuse
  const float* proj;  const int32_t* level;  const uint8_t* flags;  const uint8_t* desc_mp;  const int32_t* n_mp;
  const void* kps;  const uint8_t* desc;  const float* u_right;
  SearchGrid grid;
  int32_t* best_idx;   // [nframes][capacity_mp]
  int32_t* best_dist;  // [nframes][capacity_mp], may be null
  int capacity, capacity_mp, th_dist, check_reprojection;
  float th;
  float scale_factors[SDORB_MAX_LEVELS], inv_sigma2[SDORB_MAX_LEVELS];
};
void launch_fuse_search(const FuseSearchArgs& a, int nframes, cudaStream_t s);
void launch_sim3_agreement(const int32_t* match1, const int32_t* match2, const int32_t* n1, int capacity, int32_t* matches12,
                           int32_t* nfound, int npairs, cudaStream_t s);
void launch_search_triangulation(const SearchTriArgs& a, int npairs, cudaStream_t s);
void launch_search_init(const SearchInitArgs& a, int npairs, cudaStream_t s);
void launch_search_projection(const SearchProjArgs& a, int npairs, cudaStream_t s);
size_t search_init_smem(int capacity);
size_t search_projection_smem(int capacity);
int configure_search_kernels();

// Pipe micro-benchmarks (kernels_probe.cu): 0 = POPC, 1 = VIMNMX3.U16x2, 2 = PRMT; warp-instructions per second and per clk per SM.
int run_pipe_probe(int pipe, cudaStream_t s, double* rate_per_s, double* per_clk_sm);

size_t select_smem_bytes(const FrameGeom& g);
int configure_kernels();  // one-time cudaFuncSetAttribute calls; returns cudaError_t as int

}  // namespace sdorb
