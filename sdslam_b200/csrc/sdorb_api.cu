// sdorb_api.cu -- the C ABI of libsdorb.so (include/sdorb.h): handle, scratch memory, stream pipelines.
//
// A handle owns a compute stream, two copy streams and all device scratch.  The device-memory entry points only
// enqueue kernels; the host-memory entry points stage frames through the GPU in passes of max_batch frames with
// H2D copy / kernels / D2H copy of consecutive passes overlapped on three streams.  There is no CPU path: any
// CUDA failure is returned as SDORB_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sdorb.h"
#include "geometry.h"
#include "kernels.cuh"

using namespace sdorb;

namespace sdorb {
thread_local bool g_pdl_enabled = false;
}


// ---- guarded device allocations: the debug build that stands in for compute-sanitizer (closed on this pool).  With SDORB_GUARD=1
// in the environment every device buffer of the library is allocated with a 256 KB guard band on each side filled with 0xA5 and its
// payload poisoned with 0xCD (nothing may rely on cudaMalloc handing out zeros); sdorb_debug_guard_check() verifies every band of
// every live buffer.  An out-of-bounds WRITE of any kernel lands in a band and is reported with the buffer's name; an
// out-of-bounds or uninitialised READ brings 0xA5 / 0xCD bytes into the results, where the bit-exact parity tests see it.
// tests/test_gpu_parity.py::test_guarded_allocations_stay_intact runs the parity suite that way.
namespace {
struct GuardRec {
  void* base;
  size_t bytes;
  const char* what;
};
constexpr size_t kGuard = 256 * 1024;
std::mutex g_guard_mu;
std::map<void*, GuardRec> g_guard_live;  // user pointer -> record
int g_guard_on = -1;

bool guard_on() {
  if (g_guard_on < 0) {
    const char* e = getenv("SDORB_GUARD");
    g_guard_on = (e && e[0] != '0') ? 1 : 0;
  }
  return g_guard_on == 1;
}

template <class T>
cudaError_t sd_malloc_named(T** p, size_t bytes, const char* what) {
  if (!guard_on()) return cudaMalloc(p, bytes);
  uint8_t* base = nullptr;
  cudaError_t e = cudaMalloc(&base, bytes + 2 * kGuard);
  if (e != cudaSuccess) return e;
  cudaMemset(base, 0xA5, kGuard);
  cudaMemset(base + kGuard, 0xCD, bytes);
  cudaMemset(base + kGuard + bytes, 0xA5, kGuard);
  cudaDeviceSynchronize();  // the fills run on the legacy stream, the library's streams do not wait for it
  *p = reinterpret_cast<T*>(base + kGuard);
  std::lock_guard<std::mutex> lk(g_guard_mu);
  g_guard_live[base + kGuard] = GuardRec{base, bytes, what};
  return cudaSuccess;
}
#define sd_malloc(p, bytes) sd_malloc_named(p, bytes, #p)

int64_t g_guard_bad_freed = 0;   // damage found in buffers that have been freed since the last check
std::string g_guard_freed_msg;

// bytes of the two bands of one buffer that no longer hold 0xA5; describes the first damage in *msg when it is empty
int64_t guard_scan(const GuardRec& r, std::string* msg) {
  std::vector<uint8_t> band(kGuard);
  int64_t bad = 0;
  for (int side = 0; side < 2; ++side) {
    const uint8_t* src = side == 0 ? (const uint8_t*)r.base : (const uint8_t*)r.base + kGuard + r.bytes;
    if (cudaMemcpy(band.data(), src, kGuard, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    int64_t n = 0, first = -1;
    for (size_t i = 0; i < kGuard; ++i)
      if (band[i] != 0xA5) {
        if (first < 0) first = (int64_t)i;
        ++n;
      }
    if (n) {
      bad += n;
      if (msg && msg->empty()) {
        char text[256];
        snprintf(text, sizeof text, "guard band %s %s (%zu bytes): %lld bytes overwritten, first at offset %lld", side ? "after" : "before",
                 r.what, r.bytes, (long long)n, (long long)(side ? first : first - (int64_t)kGuard));
        *msg = text;
      }
    }
  }
  return bad;
}

cudaError_t sd_free(void* p) {
  if (!p || !guard_on()) return cudaFree(p);
  void* base = p;
  {
    std::lock_guard<std::mutex> lk(g_guard_mu);
    auto it = g_guard_live.find(p);
    if (it != g_guard_live.end()) {
      cudaDeviceSynchronize();
      const int64_t bad = guard_scan(it->second, &g_guard_freed_msg);  // a buffer is checked once more before it goes away
      if (bad > 0) g_guard_bad_freed += bad;
      base = it->second.base;
      g_guard_live.erase(it);
    }
  }
  return cudaFree(base);
}
}  // namespace

struct StageEvent {
  int stage;
  cudaEvent_t a, b;
};

struct sdorb_handle {
  sdorb_params prm{};
  Tables tables;
  int device = 0;
  cudaStream_t s_compute = nullptr, s_in = nullptr, s_out = nullptr, s_aux = nullptr;
  static constexpr int kMaxInSlots = 4;
  cudaEvent_t ev_in[kMaxInSlots]{}, ev_in_free[kMaxInSlots]{}, ev_compute[2]{}, ev_out[2]{};
  // input staging slots of the host pipeline (SDORB_PIPE_SLOTS, 2..4): with more than two, uploads run further ahead of the kernels
  int in_slots = 3;  // measured on B200: 2 -> 3 slots +2.1 % end to end (138.4 k -> 141.3 k frames/s), a fourth adds nothing
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;  // the blur runs beside FAST + selection on s_aux
  bool pipe_taper = false;   // host pipeline: shrink the last passes (SDORB_PIPE_TAPER=1; measured: -1.5 %)
  bool pipe_trace = false;    // SDORB_PIPE_TRACE=1: timed events around every upload / pass / download of a host batch, printed to stderr
  int pipe_min = 0;           // first pass of the host pipeline in frames (0: max_batch / 8; SDORB_PIPE_MIN)
  int pipe_growth_pct = 112; // ... and grow the first ones by this factor (SDORB_PIPE_GROWTH, percent)
  // host pipeline, two compute lanes: odd passes run on a twin handle (own scratch arena, own stream) so that the tail of pass p
  // (partial last waves of its 13 launches, the small upper pyramid levels) is filled by the start of pass p+1.  SDORB_PIPE_DUAL
  bool pipe_dual = false;
  int pipe_const = 0;        // > 0: passes of this many frames after the first (SDORB_PIPE_CONST) instead of the geometric ramp
  sdorb_handle* twin = nullptr;
  bool is_twin = false;
  int overlap = 0;  // SDORB_OVERLAP: 1 = blur on a second stream beside FAST (measured: +0.5 % at best, both are issue-bound),
                    // 2 = beside gather + select (latency-bound, half of the issue slots idle)
  // geometry of the current image size
  int gw = 0, gh = 0;
  FrameGeom geom{};
  FrameGeom* d_geom = nullptr;
  ResizeTap* d_taps = nullptr;
  ResizeGroup* d_groups = nullptr;
  int* d_umax = nullptr;
  // scratch for max_batch frames of the current geometry
  uint8_t *d_pyr = nullptr, *d_blur = nullptr, *d_nms = nullptr;
  uint8_t* d_stage_in[kMaxInSlots] = {nullptr, nullptr, nullptr, nullptr};
  int32_t *d_cell_seen = nullptr, *d_sel_count = nullptr, *d_error = nullptr;
  uint32_t *d_cell_list = nullptr, *d_sel = nullptr;
  uint32_t* d_okeys = nullptr;  // ORB-SLAM2-style mode only
  uint16_t* d_onode = nullptr;
  // imagePyramid staging of the host pipeline (2 slots of max_batch frames each, allocated when a caller asks for the pyramid)
  uint8_t* d_pyr_out[2] = {nullptr, nullptr};
  cudaEvent_t ev_pyr_pack[2]{};
  // output staging for the host path (2 slots)
  sdorb_keypoint* d_kps[2] = {nullptr, nullptr};
  uint8_t* d_desc[2] = {nullptr, nullptr};
  int32_t* d_counts[2] = {nullptr, nullptr};
  int out_cap = 0;
  // matcher temporaries (host path)
  void* d_match_buf = nullptr;
  size_t match_buf_bytes = 0;
  // pinned staging for the pyramid read-back of the single-frame entry point
  uint8_t* h_pyr = nullptr;
  size_t h_pyr_bytes = 0;
  uint8_t* d_pyr_pad = nullptr;  // frame 0's levels in the reference's padded form (launch_pack_padded), mirrored in h_pyr
  size_t d_pyr_pad_bytes = 0;
  // single-frame entry point (sdorb_extract, the call of Frame.cc:195): pinned result block, the event that says "the pyramid
  // levels are in h_pyr", and the whole call captured once per geometry as a CUDA graph ([0] without, [1] with the pyramid)
  uint8_t* h_res = nullptr;
  size_t h_res_bytes = 0;
  uint8_t* h_in = nullptr;  // pinned image staging: a strided / pitched upload becomes one linear copy
  size_t h_in_bytes = 0;
  cudaEvent_t ev_pyr_host = nullptr;
  struct SingleGraph {
    cudaGraphExec_t exec = nullptr;
    int capacity = 0;
    int64_t launches = 0, stage_launches[SDORB_NUM_STAGES] = {0};
  } sg[2];
  bool sg_tight = false;  // the captured kernels read level 0 with pitch == width (a contiguous caller image, uploaded linearly)
  bool use_graph = true;  // SDORB_GRAPH=0 replays the same enqueue sequence on the streams instead
  // SDORB_PYRAMID_TAIL=1: the small upper pyramid levels in one launch (resize_tail_kernel).  Off by default: measured on B200
  // (profiles/r2_pyramid_tail_probe.log) the pyramid stage gets SLOWER, 2.47 -> 2.59 ms per 4096 frames in 512-frame passes and
  // 0.169 -> 0.185 ms for the single-frame call -- one CTA per frame walking four levels behind block barriers loses more than
  // the three saved kernel boundaries give.
  bool fuse_pyramid_tail = false;
  // passes of at most this many frames use programmatic dependent launch (SDORB_PDL_MAX_FRAMES).  Default 0 = never: measured on
  // B200 (tools/pdl_probe.sh, profiles/r2_pdl_probe.log) it costs 3-4 % on 512 / 2048-frame passes and changes the single-frame
  // call by less than its run-to-run spread (inside the captured graphs the kernel boundaries are already short).
  int pdl_max_frames = 0;
  // bookkeeping
  int64_t launches = 0;
  int64_t stage_launches[SDORB_NUM_STAGES] = {0};
  double stage_ms[SDORB_NUM_STAGES] = {0};
  bool profiling = false;
  std::vector<StageEvent> pending;
  std::vector<cudaEvent_t> event_pool;
  int deferred_error = 0;
  std::string cuda_error;
};

namespace {

#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      h->cuda_error = std::string(#call) + ": " + cudaGetErrorString(e_);              \
      return e_ == cudaErrorMemoryAllocation ? SDORB_ERR_NOMEM : SDORB_ERR_CUDA;       \
    }                                                                                  \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

template <class T>
void dfree(T*& p) {
  if (p) sd_free(p);
  p = nullptr;
}

void free_single_graphs(sdorb_handle* h) {
  for (auto& g : h->sg) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g = sdorb_handle::SingleGraph{};
  }
}

void free_geometry_scratch(sdorb_handle* h) {
  free_single_graphs(h);
  dfree(h->d_geom);
  dfree(h->d_taps);
  dfree(h->d_groups);
  dfree(h->d_pyr);
  dfree(h->d_blur);
  dfree(h->d_nms);
  for (auto& p : h->d_stage_in) dfree(p);
  dfree(h->d_pyr_out[0]);
  dfree(h->d_pyr_out[1]);
  dfree(h->d_pyr_pad);
  h->d_pyr_pad_bytes = 0;
  dfree(h->d_cell_seen);
  dfree(h->d_cell_list);
  dfree(h->d_okeys);
  dfree(h->d_onode);
  dfree(h->d_sel);
  dfree(h->d_sel_count);
  for (int i = 0; i < 2; ++i) {
    dfree(h->d_kps[i]);
    dfree(h->d_desc[i]);
    dfree(h->d_counts[i]);
  }
  h->gw = h->gh = 0;
  h->out_cap = 0;
}

int geom_err(int e) {
  switch (e) {
    case 0: return SDORB_OK;
    case -4: return SDORB_ERR_GEOMETRY;
    case -7: return SDORB_ERR_UNSUPPORTED;
    default: return SDORB_ERR_BAD_ARG;
  }
}

// (Re)build geometry and scratch for a width x height input.
int ensure_geometry(sdorb_handle* h, int width, int height) {
  if (width == h->gw && height == h->gh) return SDORB_OK;
  if (width <= 0 || height <= 0 || width > h->prm.max_width || height > h->prm.max_height) return SDORB_ERR_BAD_ARG;
  FrameGeom g;
  std::vector<ResizeTap> taps;
  std::vector<ResizeGroup> groups;
  const int ge = build_frame_geom(h->tables, h->prm.nfeatures, h->prm.th_fast, width, height, &g, &taps, &groups, h->prm.min_th_fast);
  if (ge) return geom_err(ge);
  CU(cudaStreamSynchronize(h->s_compute));
  free_geometry_scratch(h);
  const size_t B = (size_t)h->prm.max_batch;
  CU(sd_malloc(&h->d_geom, sizeof(FrameGeom)));
  CU(cudaMemcpy(h->d_geom, &g, sizeof(FrameGeom), cudaMemcpyHostToDevice));
  CU(sd_malloc(&h->d_taps, sizeof(ResizeTap) * std::max<size_t>(taps.size(), 1)));
  if (!taps.empty()) CU(cudaMemcpy(h->d_taps, taps.data(), sizeof(ResizeTap) * taps.size(), cudaMemcpyHostToDevice));
  CU(sd_malloc(&h->d_groups, sizeof(ResizeGroup) * std::max<size_t>(groups.size(), 1)));
  if (!groups.empty())
    CU(cudaMemcpy(h->d_groups, groups.data(), sizeof(ResizeGroup) * groups.size(), cudaMemcpyHostToDevice));
  CU(sd_malloc(&h->d_pyr, (size_t)g.plane_total * B + 256));
  CU(sd_malloc(&h->d_blur, (size_t)g.plane_total * B + 256));
  CU(sd_malloc(&h->d_nms, (size_t)g.plane_total * B + 256));
  const size_t cells_bytes = sizeof(int32_t) * std::max<size_t>((size_t)g.cells_total * B, 1);
  CU(sd_malloc(&h->d_cell_seen, cells_bytes));
  CU(cudaMemset(h->d_cell_seen, 0, cells_bytes));
  CU(sd_malloc(&h->d_cell_list, sizeof(uint32_t) * std::max<size_t>((size_t)g.list_total * B, 1)));
  if (g.octree) {
    CU(sd_malloc(&h->d_okeys, sizeof(uint32_t) * std::max<size_t>((size_t)g.list_total * B, 1)));
    CU(sd_malloc(&h->d_onode, sizeof(uint16_t) * std::max<size_t>((size_t)g.list_total * B, 1)));
  }
  CU(sd_malloc(&h->d_sel, sizeof(uint32_t) * std::max<size_t>((size_t)g.sel_total * B, 1)));
  CU(sd_malloc(&h->d_sel_count, sizeof(int32_t) * (size_t)g.nlevels * B));
  // the uploads and fills above ran on the legacy stream; the handle's streams are non-blocking and would not wait for them
  CU(cudaStreamSynchronize(cudaStreamLegacy));
  h->geom = g;
  h->gw = width;
  h->gh = height;
  return SDORB_OK;
}

int ensure_host_staging(sdorb_handle* h, int capacity) {
  const size_t B = (size_t)h->prm.max_batch;
  if (!h->d_stage_in[0]) {
    for (int i = 0; i < h->in_slots; ++i) CU(sd_malloc(&h->d_stage_in[i], (size_t)h->geom.lv[0].plane_bytes * B + 256));
  }
  if (h->out_cap != capacity) {
    free_single_graphs(h);
    for (int i = 0; i < 2; ++i) {
      dfree(h->d_kps[i]);
      dfree(h->d_desc[i]);
      dfree(h->d_counts[i]);
      CU(sd_malloc(&h->d_kps[i], sizeof(sdorb_keypoint) * (size_t)capacity * B));
      CU(sd_malloc(&h->d_desc[i], (size_t)32 * capacity * B));
      CU(sd_malloc(&h->d_counts[i], sizeof(int32_t) * B));
    }
    h->out_cap = capacity;
  }
  return SDORB_OK;
}

cudaEvent_t take_event(sdorb_handle* h) {
  if (!h->event_pool.empty()) {
    cudaEvent_t e = h->event_pool.back();
    h->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

struct StageScope {
  sdorb_handle* h;
  cudaStream_t s;
  int stage;
  cudaEvent_t a = nullptr;
  StageScope(sdorb_handle* h_, cudaStream_t s_, int stage_) : h(h_), s(s_), stage(stage_) {
    if (h->profiling) {
      a = take_event(h);
      cudaEventRecord(a, s);
    }
  }
  void launched(int n = 1) {
    h->launches += n;
    h->stage_launches[stage] += n;
  }
  ~StageScope() {
    if (h->profiling) {
      cudaEvent_t b = take_event(h);
      cudaEventRecord(b, s);
      h->pending.push_back(StageEvent{stage, a, b});
    }
  }
};

// Enqueue ORBextractor::operator() for `n` frames (n <= max_batch) whose level 0 is described by `planes`.
enum { PASS_PYRAMID = 1, PASS_REST = 2, PASS_ALL = 3 };  // the pyramid is an output of its own (imagePyramid): it can be read back while the rest runs
struct PyrOut {           // where enqueue_pass leaves imagePyramid for the n frames of the pass (nullptr: nowhere)
  uint8_t* dst = nullptr;  // frame-major slab (sdorb_pyramid_layout), frame 0 of the pass
  int first_level = 1;
  cudaEvent_t done = nullptr;  // recorded on the stream once the slab is written (before FAST starts)
};
int enqueue_pass(sdorb_handle* h, BatchPlanes planes, int n, sdorb_keypoint* d_kps, uint8_t* d_desc, int32_t* d_counts,
                 int capacity, cudaStream_t s, int parts = PASS_ALL, const PyrOut* pyr = nullptr) {
  const FrameGeom& g = h->geom;
  g_pdl_enabled = n <= h->pdl_max_frames;
  planes.pyr = h->d_pyr;
  planes.blur = h->d_blur;
  planes.nms = h->d_nms;
  planes.batch_cap = h->prm.max_batch;
  SelectBuffers sb{h->d_cell_seen, h->d_cell_list, h->d_sel, h->d_sel_count, h->d_error, h->d_okeys, h->d_onode};
  if (parts & PASS_PYRAMID) {
    StageScope st(h, s, SDORB_STAGE_PYRAMID);
    const int tail = h->fuse_pyramid_tail ? resize_tail_first_level(g) : g.nlevels;
    for (int l = 1; l < tail; ++l) {
      launch_resize_level(h->d_geom, g, l, planes, h->d_taps, h->d_groups, n, s);
      st.launched();
    }
    if (tail < g.nlevels) {  // the small upper levels in one launch
      launch_resize_tail(h->d_geom, g, tail, planes, h->d_taps, h->d_groups, n, s);
      st.launched();
    }
  }
  if ((parts & PASS_PYRAMID) && pyr && pyr->dst) {
    StageScope st(h, s, SDORB_STAGE_PYRAMID);
    launch_pack_pyramid(h->d_geom, g, planes, pyr->first_level, pyr->dst, n, s);
    st.launched();
    if (pyr->done) CU(cudaEventRecord(pyr->done, s));
  }
  if (!(parts & PASS_REST)) {
    CU(cudaGetLastError());
    return SDORB_OK;
  }
  // The blur only needs the pyramid, FAST + selection only the pyramid too: the blur (byte dot products, FMA pipe) runs
  // on a second stream beside FAST (min / max, ALU pipe) and joins before the descriptors.
  const bool fork = h->overlap != 0;
  auto fork_blur = [&]() -> int {
    CU(cudaEventRecord(h->ev_fork, s));
    CU(cudaStreamWaitEvent(h->s_aux, h->ev_fork, 0));
    StageScope st(h, h->s_aux, SDORB_STAGE_BLUR);
    launch_blur_all(h->d_geom, g, planes, n, h->s_aux);
    st.launched();
    return SDORB_OK;
  };
  if (h->overlap == 1) {
    const int rc = fork_blur();
    if (rc) return rc;
  }
  {
    StageScope st(h, s, SDORB_STAGE_FAST);
    if (g.tiles_total_fast + g.tiles_total_fastn > 0) {
      launch_fast_all(h->d_geom, g, planes, n, s);
      st.launched();  // full tiles and the narrow last-column tiles in one grid
    }
  }
  if (h->overlap == 2) {
    const int rc = fork_blur();
    if (rc) return rc;
  }
  if (h->overlap == 3) CU(cudaEventRecord(h->ev_fork, s));  // the blur is launched behind gather + select but waits for FAST only
  {
    StageScope st(h, s, SDORB_STAGE_SELECT);
    launch_select(h->d_geom, g, planes, sb, n, s);
    st.launched(g.cells_total > 0 ? 2 : 1);  // gather_cells_kernel + select_kernel
  }
  if (h->overlap == 3) {
    CU(cudaStreamWaitEvent(h->s_aux, h->ev_fork, 0));
    StageScope st(h, h->s_aux, SDORB_STAGE_BLUR);
    launch_blur_all(h->d_geom, g, planes, n, h->s_aux);
    st.launched();
  }
  if (fork) {
    CU(cudaEventRecord(h->ev_join, h->s_aux));
    CU(cudaStreamWaitEvent(s, h->ev_join, 0));
  } else {
    StageScope st(h, s, SDORB_STAGE_BLUR);
    launch_blur_all(h->d_geom, g, planes, n, s);
    st.launched();
  }
  {
    StageScope st(h, s, SDORB_STAGE_DESCRIBE);
    launch_describe(h->d_geom, g, planes, sb, h->d_umax, d_kps, d_desc, d_counts, capacity, n, s);
    st.launched();
  }
  CU(cudaGetLastError());
  return SDORB_OK;
}

int check_deferred(sdorb_handle* h) {
  int flag = 0;
  CU(cudaMemcpy(&flag, h->d_error, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) {
    h->deferred_error = -flag;
    int zero = 0;
    CU(cudaMemcpy(h->d_error, &zero, sizeof(int), cudaMemcpyHostToDevice));
  }
  return SDORB_OK;
}

bool aligned16(const void* p, size_t a, size_t b) { return ((uintptr_t)p % 16 == 0) && (a % 16 == 0) && (b % 16 == 0); }

int host_pipeline(sdorb_handle* h, const uint8_t* images, int nframes, int width, int height, size_t row_stride,
                  size_t frame_stride, sdorb_keypoint* keypoints, uint8_t* descriptors, int32_t* counts, int capacity,
                  uint8_t* pyramid, int first_level);

}  // namespace

extern "C" {

const char* sdorb_strerror(int code) {
  switch (code) {
    case SDORB_OK: return "ok";
    case SDORB_ERR_BAD_ARG: return "bad argument";
    case SDORB_ERR_CAPACITY: return "output capacity too small";
    case SDORB_ERR_CUDA: return "CUDA error";
    case SDORB_ERR_GEOMETRY: return "cell ROI outside the level image (the reference throws / reads out of bounds)";
    case SDORB_ERR_NOMEM: return "out of device or pinned memory";
    case SDORB_ERR_OVERFLOW: return "internal list overflow";
    case SDORB_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

const char* sdorb_last_cuda_error(const sdorb_handle* h) { return h ? h->cuda_error.c_str() : ""; }

int sdorb_create(const sdorb_params* params, sdorb_handle** out) {
  if (!params || !out) return SDORB_ERR_BAD_ARG;
  *out = nullptr;
  if (params->nfeatures < 0 || params->nlevels <= 0 || params->nlevels > SDORB_MAX_LEVELS || params->max_batch <= 0 ||
      params->max_batch > 65535 || params->max_width <= 0 || params->max_height <= 0 ||
      params->max_width > SDORB_MAX_DIM || params->max_height > SDORB_MAX_DIM || !(params->scale_factor > 0.f))
    return SDORB_ERR_BAD_ARG;
  sdorb_handle* h = new (std::nothrow) sdorb_handle;
  if (!h) return SDORB_ERR_NOMEM;
  h->prm = *params;
  int dev = params->device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) {
    delete h;
    return SDORB_ERR_CUDA;
  }
  h->device = dev;
  build_tables(params->nfeatures, params->scale_factor, params->nlevels, &h->tables);
  // the describe kernel takes the height of disc column u as umax[|u|]: the table must be symmetric (it is, by construction)
  for (int au = 0; au <= SDORB_HALF_PATCH; ++au) {
    int vmax = 0;
    for (int v = 1; v <= SDORB_HALF_PATCH; ++v)
      if (au <= h->tables.umax[v]) vmax = v;
    if (vmax != h->tables.umax[au]) {
      delete h;
      return SDORB_ERR_UNSUPPORTED;
    }
  }
  auto fail = [&](int code) {
    sdorb_destroy(h);
    return code;
  };
  DeviceGuard guard(dev);
  if (cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking) != cudaSuccess) return fail(SDORB_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&h->s_aux, cudaStreamNonBlocking) != cudaSuccess) return fail(SDORB_ERR_CUDA);
  if (cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess) return fail(SDORB_ERR_CUDA);
  if (cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) return fail(SDORB_ERR_CUDA);
  if (cudaEventCreateWithFlags(&h->ev_pyr_host, cudaEventDisableTiming) != cudaSuccess) return fail(SDORB_ERR_CUDA);
  if (const char* e = getenv("SDORB_GRAPH")) h->use_graph = e[0] != '0';
  if (const char* e = getenv("SDORB_PYRAMID_TAIL")) h->fuse_pyramid_tail = e[0] != '0';
  if (const char* e = getenv("SDORB_PDL_MAX_FRAMES")) h->pdl_max_frames = std::max(atoi(e), 0);
  if (const char* e = getenv("SDORB_OVERLAP")) h->overlap = atoi(e);
  if (const char* e = getenv("SDORB_PIPE_TAPER")) h->pipe_taper = e[0] != '0';
  if (const char* e = getenv("SDORB_PIPE_GROWTH")) h->pipe_growth_pct = std::min(std::max(atoi(e), 101), 1000);
  if (const char* e = getenv("SDORB_PIPE_DUAL")) h->pipe_dual = e[0] != '0';
  if (const char* e = getenv("SDORB_PIPE_CONST")) h->pipe_const = std::max(atoi(e), 0);
  if (const char* e = getenv("SDORB_PIPE_MIN")) h->pipe_min = std::max(atoi(e), 0);
  if (const char* e = getenv("SDORB_PIPE_TRACE")) h->pipe_trace = e[0] != '0';
  if (cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking) != cudaSuccess) return fail(SDORB_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking) != cudaSuccess) return fail(SDORB_ERR_CUDA);
  if (const char* e = getenv("SDORB_PIPE_SLOTS")) h->in_slots = std::min(std::max(atoi(e), 2), (int)sdorb_handle::kMaxInSlots);
  for (int i = 0; i < sdorb_handle::kMaxInSlots; ++i) {
    if (cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming) != cudaSuccess) return fail(SDORB_ERR_CUDA);
    if (cudaEventCreateWithFlags(&h->ev_in_free[i], cudaEventDisableTiming) != cudaSuccess) return fail(SDORB_ERR_CUDA);
  }
  for (int i = 0; i < 2; ++i) {
    if (cudaEventCreateWithFlags(&h->ev_compute[i], cudaEventDisableTiming) != cudaSuccess) return fail(SDORB_ERR_CUDA);
    if (cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming) != cudaSuccess) return fail(SDORB_ERR_CUDA);
    if (cudaEventCreateWithFlags(&h->ev_pyr_pack[i], cudaEventDisableTiming) != cudaSuccess) return fail(SDORB_ERR_CUDA);
  }
  if (sd_malloc(&h->d_umax, sizeof(int) * 16) != cudaSuccess) return fail(SDORB_ERR_NOMEM);
  if (cudaMemcpy(h->d_umax, h->tables.umax, sizeof(int) * 16, cudaMemcpyHostToDevice) != cudaSuccess) return fail(SDORB_ERR_CUDA);
  if (sd_malloc(&h->d_error, sizeof(int)) != cudaSuccess) return fail(SDORB_ERR_NOMEM);
  if (cudaMemset(h->d_error, 0, sizeof(int)) != cudaSuccess) return fail(SDORB_ERR_CUDA);
  if (cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess) return fail(SDORB_ERR_CUDA);  // the fills above, before any stream of the handle runs
  if (configure_kernels() != 0 || configure_frame_kernels() != 0 || configure_search_kernels() != 0 || configure_octree_kernel() != 0)
    return fail(SDORB_ERR_CUDA);
  *out = h;
  return SDORB_OK;
}

void sdorb_destroy(sdorb_handle* h) {
  if (!h) return;
  if (h->twin) sdorb_destroy(h->twin);
  h->twin = nullptr;
  {
    DeviceGuard guard(h->device);
    if (h->s_compute) cudaStreamSynchronize(h->s_compute);
    if (h->s_aux) cudaStreamSynchronize(h->s_aux);
    if (h->s_in) cudaStreamSynchronize(h->s_in);
    if (h->s_out) cudaStreamSynchronize(h->s_out);
    free_geometry_scratch(h);
    dfree(h->d_umax);
    dfree(h->d_error);
    if (h->d_match_buf) sd_free(h->d_match_buf);
    if (h->h_pyr) cudaFreeHost(h->h_pyr);
    if (h->h_res) cudaFreeHost(h->h_res);
    if (h->h_in) cudaFreeHost(h->h_in);
    if (h->ev_pyr_host) cudaEventDestroy(h->ev_pyr_host);
    for (auto& p : h->pending) {
      cudaEventDestroy(p.a);
      cudaEventDestroy(p.b);
    }
    for (auto e : h->event_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
      if (h->ev_compute[i]) cudaEventDestroy(h->ev_compute[i]);
      if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
      if (h->ev_pyr_pack[i]) cudaEventDestroy(h->ev_pyr_pack[i]);
    }
    for (int i = 0; i < sdorb_handle::kMaxInSlots; ++i) {
      if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
      if (h->ev_in_free[i]) cudaEventDestroy(h->ev_in_free[i]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->s_aux) cudaStreamDestroy(h->s_aux);
    if (h->s_compute) cudaStreamDestroy(h->s_compute);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
  }
  delete h;
}

int sdorb_get_tables(const sdorb_handle* h, float* sf, float* isf, float* s2, float* is2, int* npl) {
  if (!h) return SDORB_ERR_BAD_ARG;
  for (int i = 0; i < h->tables.nlevels; ++i) {
    if (sf) sf[i] = h->tables.scale[i];
    if (isf) isf[i] = h->tables.inv_scale[i];
    if (s2) s2[i] = h->tables.sigma2[i];
    if (is2) is2[i] = h->tables.inv_sigma2[i];
    if (npl) npl[i] = h->tables.n_per_level[i];
  }
  return SDORB_OK;
}

int sdorb_max_keypoints(const sdorb_handle* h) {
  if (!h) return SDORB_ERR_BAD_ARG;
  int s = 0;
  // ORB-SLAM2-style mode: DistributeOctTree returns up to N + 2 keypoints per level, or the 4 children of each of its
  // (at most 8) initial nodes
  for (int v : h->tables.n_per_level) s += h->prm.min_th_fast >= 0 ? octree_level_slots(v, 8) : std::max(v, 0);
  return s;
}

int sdorb_level_size(const sdorb_handle* h, int width, int height, int level, int* lw, int* lh) {
  if (!h || level < 0 || level >= h->tables.nlevels || !lw || !lh) return SDORB_ERR_BAD_ARG;
  level_size(h->tables, level, width, height, lw, lh);
  return SDORB_OK;
}

int sdorb_batch_status(sdorb_handle* h) {
  if (!h) return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  const int rc = check_deferred(h);
  if (rc) return rc;
  const int e = h->deferred_error;
  h->deferred_error = 0;
  return e;
}

int sdorb_shard_range(int shard, int nshards, int nframes, int* first, int* last) {
  if (shard < 0 || nshards <= 0 || shard >= nshards || nframes < 0 || !first || !last) return SDORB_ERR_BAD_ARG;
  *first = (int)((int64_t)shard * nframes / nshards);
  *last = (int)((int64_t)(shard + 1) * nframes / nshards);
  return SDORB_OK;
}

// Frames are independent (operator() keeps no state, src/ORBextractor.cc:620-678): handle g of ndev -- one per GPU -- extracts the
// contiguous range sdorb_shard_range(g, ndev, nframes) on its own host thread through the host pipeline, straight into the
// caller's slabs.  No exchange between the GPUs; the "gather" is that all threads write disjoint ranges of one host result.
int sdorb_extract_batch_multi(sdorb_handle* const* handles, int ndev, const uint8_t* images, int nframes, int width, int height,
                              size_t row_stride, size_t frame_stride, sdorb_keypoint* keypoints, uint8_t* descriptors,
                              int32_t* counts, int capacity, uint8_t* pyramid, int first_level) {
  if (!handles || ndev <= 0 || nframes < 0) return SDORB_ERR_BAD_ARG;
  for (int g = 0; g < ndev; ++g) {
    if (!handles[g]) return SDORB_ERR_BAD_ARG;
    for (int k = 0; k < g; ++k)
      if (handles[k] == handles[g]) return SDORB_ERR_BAD_ARG;  // a handle is not re-entrant
  }
  if (nframes == 0) return SDORB_OK;
  size_t pyr_frame = 0;
  if (pyramid) {
    const int rc = sdorb_pyramid_layout(handles[0], width, height, nullptr, &pyr_frame);
    if (rc) return rc;
  }
  std::vector<int> rcs((size_t)ndev, SDORB_OK);
  auto work = [&](int g) {
    int lo = 0, hi = 0;
    sdorb_shard_range(g, ndev, nframes, &lo, &hi);
    if (hi <= lo) return;
    rcs[(size_t)g] = sdorb_extract_batch_pyr(handles[g], images + (size_t)lo * frame_stride, hi - lo, width, height, row_stride, frame_stride,
                                             keypoints ? keypoints + (size_t)lo * capacity : nullptr,
                                             descriptors ? descriptors + (size_t)lo * capacity * 32 : nullptr, counts ? counts + lo : nullptr,
                                             capacity, pyramid ? pyramid + (size_t)lo * pyr_frame : nullptr, first_level, SDORB_MEM_HOST,
                                             nullptr);
  };
  std::vector<std::thread> pool;
  for (int g = 1; g < ndev; ++g) pool.emplace_back(work, g);
  work(0);  // the calling thread drives the first GPU
  for (auto& t : pool) t.join();
  for (int g = 0; g < ndev; ++g)
    if (rcs[(size_t)g]) return rcs[(size_t)g];
  return SDORB_OK;
}

int sdorb_pyramid_layout(const sdorb_handle* h, int width, int height, size_t* level_offset, size_t* frame_bytes) {
  if (!h || width <= 0 || height <= 0) return SDORB_ERR_BAD_ARG;
  size_t off = 0;
  for (int l = 0; l < h->tables.nlevels; ++l) {
    int lw = 0, lh = 0;
    const int rc = sdorb_level_size(h, width, height, l, &lw, &lh);
    if (rc) return rc;
    if (level_offset) level_offset[l] = off;
    off += ((size_t)std::max(lw, 0) * (size_t)std::max(lh, 0) + 15) / 16 * 16;
  }
  if (frame_bytes) *frame_bytes = off;
  return SDORB_OK;
}

int sdorb_extract_batch(sdorb_handle* h, const uint8_t* images, int nframes, int width, int height, size_t row_stride,
                        size_t frame_stride, sdorb_keypoint* keypoints, uint8_t* descriptors, int32_t* counts,
                        int capacity, int mem, void* stream) {
  return sdorb_extract_batch_pyr(h, images, nframes, width, height, row_stride, frame_stride, keypoints, descriptors, counts, capacity,
                                 nullptr, 1, mem, stream);
}

int sdorb_extract_batch_pyr(sdorb_handle* h, const uint8_t* images, int nframes, int width, int height, size_t row_stride,
                            size_t frame_stride, sdorb_keypoint* keypoints, uint8_t* descriptors, int32_t* counts,
                            int capacity, uint8_t* pyramid, int first_level, int mem, void* stream) {
  if (!h || nframes < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (pyramid && (first_level < 0 || first_level > 1 || (mem == SDORB_MEM_DEVICE && (uintptr_t)pyramid % 16))) return SDORB_ERR_BAD_ARG;
  if (nframes == 0) return SDORB_OK;
  if (!images || !keypoints || !descriptors || !counts || width <= 0 || height <= 0 || row_stride < (size_t)width ||
      (nframes > 1 && frame_stride < row_stride * (size_t)(height - 1) + (size_t)width))
    return SDORB_ERR_BAD_ARG;
  if (capacity < sdorb_max_keypoints(h)) return SDORB_ERR_CAPACITY;
  DeviceGuard guard(h->device);
  int rc = ensure_geometry(h, width, height);
  if (rc) return rc;
  const int B = h->prm.max_batch;
  const LevelGeom& L0 = h->geom.lv[0];

  if (mem == SDORB_MEM_DEVICE) {
    cudaStream_t s = stream ? (cudaStream_t)stream : h->s_compute;
    const bool direct = aligned16(images, row_stride, frame_stride);
    for (int f0 = 0; f0 < nframes; f0 += B) {
      const int n = std::min(B, nframes - f0);
      BatchPlanes pl{};
      if (direct) {
        pl.img0 = images + (size_t)f0 * frame_stride;
        pl.img0_frame_stride = (int64_t)frame_stride;
        pl.img0_pitch = (int)row_stride;
      } else {
        // unaligned caller layout: repack level 0 into the pitch-aligned scratch plane
        for (int f = 0; f < n; ++f)
          CU(cudaMemcpy2DAsync(h->d_pyr + (size_t)f * L0.plane_bytes, L0.pitch, images + (size_t)(f0 + f) * frame_stride,
                               row_stride, width, height, cudaMemcpyDeviceToDevice, s));
        pl.img0 = h->d_pyr;
        pl.img0_frame_stride = L0.plane_bytes;
        pl.img0_pitch = L0.pitch;
      }
      PyrOut po;
      if (pyramid) {
        int64_t off[SDORB_MAX_LEVELS], fb = 0;
        pyramid_layout(h->geom, off, &fb);
        po.dst = pyramid + (size_t)f0 * (size_t)fb;
        po.first_level = first_level;
      }
      rc = enqueue_pass(h, pl, n, keypoints + (size_t)f0 * capacity, descriptors + (size_t)f0 * capacity * 32, counts + f0,
                        capacity, s, PASS_ALL, &po);
      if (rc) return rc;
    }
    return SDORB_OK;
  }

  // host path: three-stream pipeline over passes of up to max_batch frames.  Only the first upload and the last
  // download are exposed, so the passes start at max_batch / 8.  A pass can only start once it is uploaded completely, and on
  // this platform PCIe delivers 640x480 frames about as fast as the kernels consume them (5.65 vs 5.5 us per frame; 5.6 vs 7 us
  // before the kernels of round 2), so the passes grow slowly: by 1.12 per pass (measured best of 1.05 ... 1.25 and of constant
  // pass sizes, profiles/r2_e2e_schedule_probe.log: 153.9 k frames/s against 144.5 k at 1.25).  Doubling stalled the kernels for
  // 2 ms per 4096-frame call; tapering the last passes costs more in small-pass efficiency than the shorter last download saves.
  rc = ensure_host_staging(h, capacity);
  if (rc) return rc;
  // Any failure after the first enqueue leaves copies into / out of the caller's buffers in flight on three (four) streams:
  // drain them before the error is returned, so that the caller may free or reuse its buffers right away.
  rc = host_pipeline(h, images, nframes, width, height, row_stride, frame_stride, keypoints, descriptors, counts, capacity, pyramid,
                     first_level);
  if (rc) {
    cudaStreamSynchronize(h->s_in);
    cudaStreamSynchronize(h->s_compute);
    cudaStreamSynchronize(h->s_out);
    if (h->twin) cudaStreamSynchronize(h->twin->s_compute);
    cudaGetLastError();
  }
  return rc;
}

}  // extern "C"

namespace {
int host_pipeline(sdorb_handle* h, const uint8_t* images, int nframes, int width, int height, size_t row_stride,
                  size_t frame_stride, sdorb_keypoint* keypoints, uint8_t* descriptors, int32_t* counts, int capacity,
                  uint8_t* pyramid, int first_level) {
  const int B = h->prm.max_batch;
  const LevelGeom& L0 = h->geom.lv[0];
  int rc = SDORB_OK;
  int64_t pyr_off[SDORB_MAX_LEVELS], pyr_frame = 0;
  pyramid_layout(h->geom, pyr_off, &pyr_frame);
  if (pyramid && !h->d_pyr_out[0])
    for (int i = 0; i < 2; ++i) CU(sd_malloc(&h->d_pyr_out[i], (size_t)pyr_frame * B + 256));
  // opt-in timeline (debugging aid): six timed events per pass, all relative to the start of the first upload
  struct PassTrace {
    int n;
    cudaEvent_t e[6];  // upload begin / end, kernels begin / end, download begin / end
  };
  std::vector<PassTrace> trace;
  auto mark = [&](int which, cudaStream_t st) {
    if (!h->pipe_trace) return;
    cudaEventCreate(&trace.back().e[which]);
    cudaEventRecord(trace.back().e[which], st);
  };
  int pass = 0;
  const int n_min = h->pipe_min > 0 ? std::min(h->pipe_min, B) : std::max(B / 8, 1);
  int ramp = n_min;
  // two compute lanes (see pipe_dual): staging slot 1 belongs to the twin's scratch arena and stream
  const bool dual = h->pipe_dual && !h->is_twin && nframes > n_min;
  if (dual) {
    if (!h->twin) {
      sdorb_params p2 = h->prm;
      p2.device = h->device;
      rc = sdorb_create(&p2, &h->twin);
      if (rc) return rc;
      h->twin->is_twin = true;
    }
    h->twin->profiling = h->profiling;
    rc = ensure_geometry(h->twin, width, height);
    if (rc) return rc;
  }
  for (int f0 = 0, n = 0; f0 < nframes; f0 += n, ++pass) {
    const int left = nframes - f0;
    n = h->pipe_taper ? std::min(std::min(ramp, B), std::max((left + 1) / 2, std::min(left, n_min))) : std::min(std::min(ramp, B), left);
    ramp = std::min(std::max(ramp + 1, (int)((int64_t)ramp * h->pipe_growth_pct / 100)), B);
    if (h->pipe_const > 0) ramp = std::min(h->pipe_const, B);
    const int slot = pass & 1;
    sdorb_handle* hc = (dual && slot) ? h->twin : h;  // the lane: scratch arena + compute stream
    cudaStream_t cs = hc->s_compute;
    const int si = pass % h->in_slots;  // input staging slot: free again once the pass that last used it has run its kernels
    if (pass >= h->in_slots) CU(cudaStreamWaitEvent(h->s_in, h->ev_in_free[si], 0));
    if (h->pipe_trace) trace.push_back(PassTrace{n, {}});
    mark(0, h->s_in);
    const bool tight = frame_stride == row_stride * (size_t)height && row_stride == (size_t)width && width % 16 == 0;
    if (tight) {
      // contiguous host frames whose rows stay 16-byte aligned: one linear copy, and the kernels read level 0 with the image
      // width as its pitch (a 2-D copy is issued row by row: measured 6.5 GB/s at 752x480 against 54 GB/s linear)
      CU(cudaMemcpyAsync(h->d_stage_in[si], images + (size_t)f0 * frame_stride, frame_stride * (size_t)n, cudaMemcpyHostToDevice,
                         h->s_in));
    } else if (frame_stride == row_stride * (size_t)height) {
      CU(cudaMemcpy2DAsync(h->d_stage_in[si], L0.pitch, images + (size_t)f0 * frame_stride, row_stride, width,
                           (size_t)height * n, cudaMemcpyHostToDevice, h->s_in));
    } else {
      for (int f = 0; f < n; ++f)
        CU(cudaMemcpy2DAsync(h->d_stage_in[si] + (size_t)f * L0.plane_bytes, L0.pitch,
                             images + (size_t)(f0 + f) * frame_stride, row_stride, width, height, cudaMemcpyHostToDevice,
                             h->s_in));
    }
    mark(1, h->s_in);
    CU(cudaEventRecord(h->ev_in[si], h->s_in));
    CU(cudaStreamWaitEvent(cs, h->ev_in[si], 0));
    if (pass >= 2) CU(cudaStreamWaitEvent(cs, h->ev_out[slot], 0));
    mark(2, cs);
    BatchPlanes pl{};
    pl.img0 = h->d_stage_in[si];
    pl.img0_frame_stride = tight ? (int64_t)frame_stride : L0.plane_bytes;
    pl.img0_pitch = tight ? width : L0.pitch;
    PyrOut po;
    po.dst = pyramid ? h->d_pyr_out[slot] : nullptr;
    po.first_level = first_level;
    po.done = h->ev_pyr_pack[slot];
    rc = enqueue_pass(hc, pl, n, h->d_kps[slot], h->d_desc[slot], h->d_counts[slot], capacity, cs, PASS_ALL, &po);
    if (rc) {
      if (hc != h) h->cuda_error = hc->cuda_error;
      return rc;
    }
    if (pyramid) {
      // the pyramid slab of the pass is complete before FAST starts: its (large) copy to the host runs beside the other kernels
      CU(cudaStreamWaitEvent(h->s_out, h->ev_pyr_pack[slot], 0));
      if (first_level == 0) {
        CU(cudaMemcpyAsync(pyramid + (size_t)f0 * pyr_frame, h->d_pyr_out[slot], (size_t)pyr_frame * n, cudaMemcpyDeviceToHost, h->s_out));
      } else {  // level 0 is the caller's own input: skip its bytes of every frame
        const size_t skip = (size_t)pyr_off[first_level];
        CU(cudaMemcpy2DAsync(pyramid + (size_t)f0 * pyr_frame + skip, (size_t)pyr_frame, h->d_pyr_out[slot] + skip, (size_t)pyr_frame,
                             (size_t)pyr_frame - skip, (size_t)n, cudaMemcpyDeviceToHost, h->s_out));
      }
    }
    mark(3, cs);
    CU(cudaEventRecord(h->ev_compute[slot], cs));
    CU(cudaEventRecord(h->ev_in_free[si], cs));
    CU(cudaStreamWaitEvent(h->s_out, h->ev_compute[slot], 0));
    mark(4, h->s_out);
    CU(cudaMemcpyAsync(keypoints + (size_t)f0 * capacity, h->d_kps[slot], sizeof(sdorb_keypoint) * (size_t)capacity * n,
                       cudaMemcpyDeviceToHost, h->s_out));
    CU(cudaMemcpyAsync(descriptors + (size_t)f0 * capacity * 32, h->d_desc[slot], (size_t)32 * capacity * n,
                       cudaMemcpyDeviceToHost, h->s_out));
    CU(cudaMemcpyAsync(counts + f0, h->d_counts[slot], sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->s_out));
    mark(5, h->s_out);
    CU(cudaEventRecord(h->ev_out[slot], h->s_out));
  }
  CU(cudaStreamSynchronize(h->s_out));
  CU(cudaStreamSynchronize(h->s_compute));
  if (h->pipe_trace && !trace.empty()) {
    fprintf(stderr, "sdorb host pipeline: %d frames in %zu passes; ms since the first upload began\n", nframes, trace.size());
    fprintf(stderr, "  pass frames   upload          kernels         download\n");
    for (size_t i = 0; i < trace.size(); ++i) {
      float t[6] = {0, 0, 0, 0, 0, 0};
      for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&t[k], trace[0].e[0], trace[i].e[k]);
      fprintf(stderr, "  %4zu %6d   %6.2f-%6.2f   %6.2f-%6.2f   %6.2f-%6.2f\n", i, trace[i].n, t[0], t[1], t[2], t[3], t[4], t[5]);
    }
    for (auto& p : trace)
      for (int k = 0; k < 6; ++k) cudaEventDestroy(p.e[k]);
  }
  if (dual) {
    CU(cudaStreamSynchronize(h->twin->s_compute));
    h->launches += h->twin->launches;
    for (int i = 0; i < SDORB_NUM_STAGES; ++i) h->stage_launches[i] += h->twin->stage_launches[i], h->twin->stage_launches[i] = 0;
    h->twin->launches = 0;
    h->pending.insert(h->pending.end(), h->twin->pending.begin(), h->twin->pending.end());  // stage events of the second lane
    h->twin->pending.clear();
    rc = check_deferred(h->twin);
    if (rc) return rc;
    if (h->twin->deferred_error && !h->deferred_error) h->deferred_error = h->twin->deferred_error;
    h->twin->deferred_error = 0;
  }
  rc = check_deferred(h);
  if (rc) return rc;
  if (h->deferred_error) {
    const int e = h->deferred_error;
    h->deferred_error = 0;
    return e;
  }
  return SDORB_OK;
}

}  // namespace

extern "C" {

void sdorb_fill_border_reflect101(uint8_t* origin, int width, int height, size_t stride, int border) {
  if (!origin || width <= 0 || height <= 0 || border <= 0) return;
  if (border < width && border < height) {  // the usual case: one reflection, no index arithmetic per pixel
    for (int y = 0; y < height; ++y) {
      uint8_t* row = origin + (ptrdiff_t)y * (ptrdiff_t)stride;
      uint8_t* re = row + width - 1;
      for (int x = 1; x <= border; ++x) {
        row[-x] = row[x];
        re[x] = re[-x];
      }
    }
    const size_t n = (size_t)width + 2 * (size_t)border;
    for (int y = 1; y <= border; ++y) {
      memcpy(origin - (ptrdiff_t)y * (ptrdiff_t)stride - border, origin + (ptrdiff_t)y * (ptrdiff_t)stride - border, n);
      memcpy(origin + (ptrdiff_t)(height - 1 + y) * (ptrdiff_t)stride - border,
             origin + (ptrdiff_t)(height - 1 - y) * (ptrdiff_t)stride - border, n);
    }
    return;
  }
  auto refl = [](int p, int len) {
    if (len == 1) return 0;
    while ((unsigned)p >= (unsigned)len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
  };
  for (int y = 0; y < height; ++y) {
    uint8_t* row = origin + (ptrdiff_t)y * (ptrdiff_t)stride;
    for (int x = 1; x <= border; ++x) {
      row[-x] = row[refl(-x, width)];
      row[width - 1 + x] = row[refl(width - 1 + x, width)];
    }
  }
  for (int y = 1; y <= border; ++y) {
    memcpy(origin - (ptrdiff_t)y * (ptrdiff_t)stride - border, origin + (ptrdiff_t)refl(-y, height) * (ptrdiff_t)stride - border,
           (size_t)width + 2 * (size_t)border);
    memcpy(origin + (ptrdiff_t)(height - 1 + y) * (ptrdiff_t)stride - border,
           origin + (ptrdiff_t)refl(height - 1 + y, height) * (ptrdiff_t)stride - border, (size_t)width + 2 * (size_t)border);
  }
}

int sdorb_host_tables(int nfeatures, float scale_factor, int nlevels, float* sf, float* isf, float* s2, float* is2,
                      int* npl, int* umax) {
  if (nfeatures < 0 || nlevels <= 0 || nlevels > SDORB_MAX_LEVELS || !(scale_factor > 0.f)) return SDORB_ERR_BAD_ARG;
  Tables t;
  build_tables(nfeatures, scale_factor, nlevels, &t);
  for (int i = 0; i < nlevels; ++i) {
    if (sf) sf[i] = t.scale[i];
    if (isf) isf[i] = t.inv_scale[i];
    if (s2) s2[i] = t.sigma2[i];
    if (is2) is2[i] = t.inv_sigma2[i];
    if (npl) npl[i] = t.n_per_level[i];
  }
  if (umax)
    for (int i = 0; i <= SDORB_HALF_PATCH; ++i) umax[i] = t.umax[i];
  return SDORB_OK;
}

int sdorb_host_level_geometry(int nfeatures, float scale_factor, int nlevels, int th_fast, int width, int height,
                              sdorb_level_geom* out) {
  if (!out || nfeatures < 0 || nlevels <= 0 || nlevels > SDORB_MAX_LEVELS || !(scale_factor > 0.f))
    return SDORB_ERR_BAD_ARG;
  Tables t;
  build_tables(nfeatures, scale_factor, nlevels, &t);
  FrameGeom g;
  std::vector<ResizeTap> taps;
  const int ge = build_frame_geom(t, nfeatures, th_fast, width, height, &g, &taps);
  if (ge) return geom_err(ge);
  for (int l = 0; l < nlevels; ++l) {
    const LevelGeom& L = g.lv[l];
    out[l] = sdorb_level_geom{L.w, L.h, L.n_desired, L.cols, L.rows, L.cell_w, L.cell_h, L.n_features_cell, L.scaled_patch_size};
  }
  return SDORB_OK;
}

}  // extern "C"

namespace {
size_t up64(size_t v) { return (v + 63) / 64 * 64; }

// One single-frame call after the upload of the image, in two parts that are each captured once per geometry as a CUDA graph
// (or enqueued on the stream as they are): part 0 = the pyramid kernels, part 1 = FAST .. describe and the copies of the
// results, the count and the error flag into the pinned block.  Frame 0 of staging slot 0 is the frame.  Between the two the
// caller forks the read-back of the pyramid levels onto s_out with ordinary stream calls: a host thread can wait on an event
// recorded by a stream call, while an event-record NODE inside a graph only changes the event's state when it executes.
int enqueue_single_part(sdorb_handle* h, int capacity, int part) {
  cudaStream_t s = h->s_compute;
  const LevelGeom& L0 = h->geom.lv[0];
  BatchPlanes pl{};
  pl.img0 = h->d_stage_in[0];
  pl.img0_frame_stride = L0.plane_bytes;
  pl.img0_pitch = h->sg_tight ? L0.w : L0.pitch;
  int rc = enqueue_pass(h, pl, 1, h->d_kps[0], h->d_desc[0], h->d_counts[0], capacity, s, part == 0 ? PASS_PYRAMID : PASS_REST);
  if (rc || part == 0) return rc;
  uint8_t* r = h->h_res;
  const size_t bK = sizeof(sdorb_keypoint) * (size_t)capacity, bD = (size_t)32 * capacity;
  CU(cudaMemcpyAsync(r, h->d_kps[0], bK, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(r + up64(bK), h->d_desc[0], bD, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(r + up64(bK) + up64(bD), h->d_counts[0], sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(r + up64(bK) + up64(bD) + 64, h->d_error, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  return SDORB_OK;
}

// Runs one part: replays its graph (capturing it first if needed), or enqueues it directly when graphs are off / profiling is on.
int run_single_part(sdorb_handle* h, int capacity, int part) {
  cudaStream_t s = h->s_compute;
  if (!h->use_graph || h->profiling) return enqueue_single_part(h, capacity, part);
  sdorb_handle::SingleGraph& sg = h->sg[part];
  if (sg.exec && sg.capacity != capacity) {
    cudaGraphExecDestroy(sg.exec);
    sg = sdorb_handle::SingleGraph{};
  }
  if (!sg.exec) {
    // every pointer in the capture belongs to the handle and is stable until the geometry or the capacity changes
    const int64_t l0 = h->launches;
    int64_t st0[SDORB_NUM_STAGES];
    for (int i = 0; i < SDORB_NUM_STAGES; ++i) st0[i] = h->stage_launches[i];
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_single_part(h, capacity, part);
    const cudaError_t ce = cudaStreamEndCapture(s, &graph);
    sg.launches = h->launches - l0;
    h->launches = l0;
    for (int i = 0; i < SDORB_NUM_STAGES; ++i) {
      sg.stage_launches[i] = h->stage_launches[i] - st0[i];
      h->stage_launches[i] = st0[i];
    }
    if (rc || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      if (rc) return rc;
      h->cuda_error = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce);
      return SDORB_ERR_CUDA;
    }
    const cudaError_t ie = cudaGraphInstantiate(&sg.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      sg.exec = nullptr;
      h->cuda_error = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie);
      return SDORB_ERR_CUDA;
    }
    sg.capacity = capacity;
  }
  CU(cudaGraphLaunch(sg.exec, s));
  h->launches += sg.launches;
  for (int i = 0; i < SDORB_NUM_STAGES; ++i) h->stage_launches[i] += sg.stage_launches[i];
  return SDORB_OK;
}
}  // namespace

extern "C" {

int sdorb_extract(sdorb_handle* h, const uint8_t* image, int width, int height, size_t stride, sdorb_keypoint* keypoints,
                  uint8_t* descriptors, int capacity, int* count, const sdorb_pyr_view* pyramid) {
  if (!h) return SDORB_ERR_BAD_ARG;
  if (!image || width <= 0 || height <= 0) return SDORB_OK;  // empty image: outputs untouched (src/ORBextractor.cc:622-623)
  if (!keypoints || !descriptors || !count || stride < (size_t)width) return SDORB_ERR_BAD_ARG;
  if (capacity < sdorb_max_keypoints(h)) return SDORB_ERR_CAPACITY;
  DeviceGuard guard(h->device);
  int rc = ensure_geometry(h, width, height);
  if (rc) return rc;
  rc = ensure_host_staging(h, capacity);
  if (rc) return rc;
  const FrameGeom& g = h->geom;
  const LevelGeom& L0 = g.lv[0];
  const int nl = g.nlevels;
  const bool with_pyr = pyramid != nullptr;
  const size_t bK = sizeof(sdorb_keypoint) * (size_t)capacity, bD = (size_t)32 * capacity;
  const size_t need_res = up64(bK) + up64(bD) + 128;
  if (need_res > h->h_res_bytes) {
    free_single_graphs(h);
    if (h->h_res) cudaFreeHost(h->h_res);
    h->h_res = nullptr;
    h->h_res_bytes = 0;
    CU(cudaMallocHost(&h->h_res, need_res));
    h->h_res_bytes = need_res;
  }
  int64_t pad_off[SDORB_MAX_LEVELS];
  const size_t pad_bytes = (size_t)padded_pyramid_layout(g, pad_off);
  if (with_pyr) {
    if (pad_bytes > h->h_pyr_bytes) {
      if (h->h_pyr) cudaFreeHost(h->h_pyr);
      h->h_pyr = nullptr;
      h->h_pyr_bytes = 0;
      CU(cudaMallocHost(&h->h_pyr, pad_bytes));
      h->h_pyr_bytes = pad_bytes;
    }
    if (pad_bytes > h->d_pyr_pad_bytes) {
      dfree(h->d_pyr_pad);
      h->d_pyr_pad_bytes = 0;
      CU(sd_malloc(&h->d_pyr_pad, pad_bytes));
      h->d_pyr_pad_bytes = pad_bytes;
    }
    for (int l = 0; l < nl; ++l) {
      const LevelGeom& L = g.lv[l];
      const sdorb_pyr_view& v = pyramid[l];
      if (v.data && (v.width != L.w || v.height != L.h || v.stride < (size_t)L.w)) return SDORB_ERR_BAD_ARG;
    }
  }
  cudaStream_t s = h->s_compute;
  const bool tight = stride == (size_t)width && width % 16 == 0;
  if (tight != h->sg_tight) {
    free_single_graphs(h);
    h->sg_tight = tight;
  }
  if (tight) {  // a contiguous image whose rows stay 16-byte aligned: one linear copy, the kernels read it with pitch == width
    CU(cudaMemcpyAsync(h->d_stage_in[0], image, (size_t)width * height, cudaMemcpyHostToDevice, s));
  } else {
    // a 2-D copy from pageable memory is issued row by row (measured: +70 us at 752x480): repack the rows into a pinned
    // buffer with the device pitch and send them as one linear copy
    const size_t need_in = (size_t)L0.pitch * height;
    if (need_in > h->h_in_bytes) {
      if (h->h_in) cudaFreeHost(h->h_in);
      h->h_in = nullptr;
      h->h_in_bytes = 0;
      CU(cudaMallocHost(&h->h_in, need_in));
      h->h_in_bytes = need_in;
    }
    for (int y = 0; y < height; ++y) memcpy(h->h_in + (size_t)y * L0.pitch, image + (size_t)y * stride, (size_t)width);
    CU(cudaMemcpyAsync(h->d_stage_in[0], h->h_in, need_in, cudaMemcpyHostToDevice, s));
  }
  auto drain = [&](int code) {  // nothing of a failed call may stay in flight
    cudaStreamSynchronize(s);
    cudaStreamSynchronize(h->s_out);
    cudaGetLastError();
    return code;
  };
  rc = run_single_part(h, capacity, 0);
  if (rc) return drain(rc);
  if (with_pyr && cudaEventRecord(h->ev_fork, s) != cudaSuccess) return drain(SDORB_ERR_CUDA);
  rc = run_single_part(h, capacity, 1);  // queued right behind the pyramid: the compute chain has no gap
  if (rc) return drain(rc);
  if (with_pyr) {
    // imagePyramid is complete after part 0.  On s_out, beside FAST .. describe: one kernel lays all levels out in the padded
    // form the reference returns (19 px of BORDER_REFLECT_101 around each), one copy brings them to the pinned buffer.
    cudaError_t e = cudaStreamWaitEvent(h->s_out, h->ev_fork, 0);
    if (e == cudaSuccess) {
      BatchPlanes pl{};
      pl.img0 = h->d_stage_in[0];
      pl.img0_frame_stride = L0.plane_bytes;
      pl.img0_pitch = tight ? width : L0.pitch;
      pl.pyr = h->d_pyr;
      pl.batch_cap = h->prm.max_batch;
      launch_pack_padded(h->d_geom, g, pl, 1, h->d_pyr_pad, h->s_out);
      ++h->launches;
      ++h->stage_launches[SDORB_STAGE_PYRAMID];
      e = cudaGetLastError();
    }
    if (e == cudaSuccess && nl > 1)
      e = cudaMemcpyAsync(h->h_pyr + pad_off[1], h->d_pyr_pad + pad_off[1], pad_bytes - (size_t)pad_off[1], cudaMemcpyDeviceToHost, h->s_out);
    if (e == cudaSuccess) e = cudaEventRecord(h->ev_pyr_host, h->s_out);
    if (e != cudaSuccess) {
      h->cuda_error = std::string("pyramid read-back: ") + cudaGetErrorString(e);
      return drain(SDORB_ERR_CUDA);
    }
  }
  if (with_pyr) {
    // imagePyramid (src/ORBextractor.cc:684-696), laid into the caller's views while FAST .. describe still run.  A view
    // with the reference's own layout (19 px of border, rows w + 38 apart) takes its whole padded buffer in one copy.
    // Level 0 is the caller's own image: host to host, while the other levels are still on their way.
    if (pyramid[0].data) {
      const sdorb_pyr_view& v = pyramid[0];
      if (v.stride == stride && stride == (size_t)width)
        memcpy(v.data, image, (size_t)width * height);
      else
        for (int y = 0; y < height; ++y) memcpy(v.data + (size_t)y * v.stride, image + (size_t)y * stride, (size_t)width);
      if (v.border > 0) sdorb_fill_border_reflect101(v.data, width, height, v.stride, v.border);
    }
    if (cudaEventSynchronize(h->ev_pyr_host) != cudaSuccess) {
      h->cuda_error = "pyramid read-back: cudaEventSynchronize failed";
      return drain(SDORB_ERR_CUDA);
    }
    for (int l = 1; l < nl; ++l) {
      const LevelGeom& L = g.lv[l];
      const sdorb_pyr_view& v = pyramid[l];
      if (!v.data) continue;
      const size_t pw = (size_t)L.w + 2 * SDORB_EDGE;
      const uint8_t* src = h->h_pyr + pad_off[l];
      if (v.border == SDORB_EDGE && v.stride == pw) {
        memcpy(v.data - (size_t)SDORB_EDGE * pw - SDORB_EDGE, src, pw * ((size_t)L.h + 2 * SDORB_EDGE));
      } else if (v.border >= 0 && v.border <= SDORB_EDGE) {  // any smaller border is a sub-rectangle of the padded level
        const int b = v.border;
        for (int y = -b; y < L.h + b; ++y)
          memcpy(v.data + (ptrdiff_t)y * (ptrdiff_t)v.stride - b, src + (size_t)(y + SDORB_EDGE) * pw + SDORB_EDGE - b, (size_t)L.w + 2 * (size_t)b);
      } else {
        for (int y = 0; y < L.h; ++y) memcpy(v.data + (size_t)y * v.stride, src + (size_t)(y + SDORB_EDGE) * pw + SDORB_EDGE, (size_t)L.w);
        sdorb_fill_border_reflect101(v.data, L.w, L.h, v.stride, v.border);
      }
    }
  }
  CU(cudaStreamSynchronize(s));
  const uint8_t* r = h->h_res;
  int32_t n = 0, err = 0;
  memcpy(&n, r + up64(bK) + up64(bD), sizeof(n));
  memcpy(&err, r + up64(bK) + up64(bD) + 64, sizeof(err));
  if (err) {
    int zero = 0;
    CU(cudaMemcpy(h->d_error, &zero, sizeof(int), cudaMemcpyHostToDevice));
    return -err;
  }
  if (n < 0 || n > capacity) return SDORB_ERR_OVERFLOW;
  memcpy(keypoints, r, sizeof(sdorb_keypoint) * (size_t)n);
  memcpy(descriptors, r + up64(bK), (size_t)32 * n);
  *count = n;
  return SDORB_OK;
}

}  // extern "C"

extern "C" {

namespace {
// scratch for the host-memory forms of the small batched entry points: one growing device buffer, carved by the caller
int ensure_match_buf(sdorb_handle* h, size_t need) {
  if (need > h->match_buf_bytes) {
    if (h->d_match_buf) sd_free(h->d_match_buf);
    h->d_match_buf = nullptr;
    h->match_buf_bytes = 0;
    CU(sd_malloc(&h->d_match_buf, need));
    h->match_buf_bytes = need;
  }
  return SDORB_OK;
}
size_t up256(size_t v) { return (v + 255) / 256 * 256; }
// Lays the host arrays of one call out in the handle's scratch buffer: sizes first, then copies.
struct Stager {
  struct Item { const void* host_in; void* host_out; size_t bytes, off; };
  std::vector<Item> items;
  size_t total = 0;
  size_t add(const void* in, void* out, size_t bytes) {
    items.push_back({in, out, bytes, total});
    total += up256(bytes);
    return items.size() - 1;
  }
  uint8_t* dev(sdorb_handle* h, size_t i) const { return (uint8_t*)h->d_match_buf + items[i].off; }
  int upload(sdorb_handle* h, cudaStream_t s) {
    int rc = ensure_match_buf(h, total);
    if (rc) return rc;
    for (auto& it : items)
      if (it.host_in) CU(cudaMemcpyAsync((uint8_t*)h->d_match_buf + it.off, it.host_in, it.bytes, cudaMemcpyHostToDevice, s));
    return SDORB_OK;
  }
  int download(sdorb_handle* h, cudaStream_t s) {
    for (auto& it : items)
      if (it.host_out) CU(cudaMemcpyAsync(it.host_out, (uint8_t*)h->d_match_buf + it.off, it.bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return SDORB_OK;
  }
};
}  // namespace

static int match_common(sdorb_handle* h, const uint8_t* A, const int32_t* nA, int strideA, const uint8_t* Bm,
                        const int32_t* nB, int strideB, int npairs, float ratio, int th_low, sdorb_match* out, int mem,
                        void* stream, bool greedy) {
  if (!h || npairs < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (npairs == 0) return SDORB_OK;
  if (!A || !nA || !Bm || !nB || !out || strideA <= 0 || strideB <= 0 || strideB > 65535) return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  cudaStream_t s = (mem == SDORB_MEM_DEVICE && stream) ? (cudaStream_t)stream : h->s_compute;
  const uint8_t *dA = A, *dB = Bm;
  const int32_t *dnA = nA, *dnB = nB;
  sdorb_match* dout = out;
  const size_t bytesA = (size_t)npairs * strideA * 32, bytesB = (size_t)npairs * strideB * 32;
  const size_t bytesN = sizeof(int32_t) * (size_t)npairs, bytesO = sizeof(sdorb_match) * (size_t)npairs * strideA;
  if (mem == SDORB_MEM_HOST) {
    for (int p = 0; p < npairs; ++p)
      if (nA[p] < 0 || nA[p] > strideA || nB[p] < 0 || nB[p] > strideB) return SDORB_ERR_BAD_ARG;
    auto up = up256;
    const int rcb = ensure_match_buf(h, up(bytesA) + up(bytesB) + 2 * up(bytesN) + up(bytesO));
    if (rcb) return rcb;
    uint8_t* base = (uint8_t*)h->d_match_buf;
    uint8_t* pA = base;
    uint8_t* pB = pA + up(bytesA);
    int32_t* pnA = (int32_t*)(pB + up(bytesB));
    int32_t* pnB = (int32_t*)((uint8_t*)pnA + up(bytesN));
    dout = (sdorb_match*)((uint8_t*)pnB + up(bytesN));
    CU(cudaMemcpyAsync(pA, A, bytesA, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(pB, Bm, bytesB, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(pnA, nA, bytesN, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(pnB, nB, bytesN, cudaMemcpyHostToDevice, s));
    dA = pA;
    dB = pB;
    dnA = pnA;
    dnB = pnB;
  } else if (((uintptr_t)A | (uintptr_t)Bm) % 16) {
    return SDORB_ERR_BAD_ARG;  // descriptor slabs must be 16-byte aligned on the device
  }
  {
    StageScope st(h, s, SDORB_STAGE_MATCH);
    if (greedy)
      launch_match_greedy(dA, dnA, strideA, dB, dnB, strideB, npairs, ratio, th_low, dout, s);
    else
      launch_match(dA, dnA, strideA, dB, dnB, strideB, npairs, ratio, th_low, dout, s);
    st.launched();
  }
  CU(cudaGetLastError());
  if (mem == SDORB_MEM_HOST) {
    CU(cudaMemcpyAsync(out, dout, bytesO, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  return SDORB_OK;
}

int sdorb_match_batch(sdorb_handle* h, const uint8_t* A, const int32_t* nA, int strideA, const uint8_t* Bm,
                      const int32_t* nB, int strideB, int npairs, float ratio, int th_low, sdorb_match* out, int mem,
                      void* stream) {
  return match_common(h, A, nA, strideA, Bm, nB, strideB, npairs, ratio, th_low, out, mem, stream, false);
}

int sdorb_match_greedy_batch(sdorb_handle* h, const uint8_t* A, const int32_t* nA, int strideA, const uint8_t* Bm,
                             const int32_t* nB, int strideB, int npairs, float ratio, int th_low, sdorb_match* out,
                             int mem, void* stream) {
  return match_common(h, A, nA, strideA, Bm, nB, strideB, npairs, ratio, th_low, out, mem, stream, true);
}

int sdorb_hamming_matrix(sdorb_handle* h, const uint8_t* A, int nA, const uint8_t* Bm, int nB, uint16_t* out, int mem,
                         void* stream) {
  if (!h || nA < 0 || nB < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (nA == 0 || nB == 0) return SDORB_OK;
  if (!A || !Bm || !out) return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  cudaStream_t s = (mem == SDORB_MEM_DEVICE && stream) ? (cudaStream_t)stream : h->s_compute;
  if (mem == SDORB_MEM_DEVICE) {
    if (((uintptr_t)A | (uintptr_t)Bm) % 16) return SDORB_ERR_BAD_ARG;
    StageScope st(h, s, SDORB_STAGE_MATCH);
    launch_hamming_matrix(A, nA, Bm, nB, out, s);
    st.launched();
    CU(cudaGetLastError());
    return SDORB_OK;
  }
  // host form: staged through the handle's scratch buffer (no allocation per call; DescriptorDistance calls this per pair)
  Stager st;
  const size_t iA = st.add(A, nullptr, (size_t)nA * 32), iB = st.add(Bm, nullptr, (size_t)nB * 32);
  const size_t iO = st.add(nullptr, out, sizeof(uint16_t) * (size_t)nA * nB);
  int rc = st.upload(h, s);
  if (rc) return rc;
  {
    StageScope sc(h, s, SDORB_STAGE_MATCH);
    launch_hamming_matrix(st.dev(h, iA), nA, st.dev(h, iB), nB, (uint16_t*)st.dev(h, iO), s);
    sc.launched();
  }
  CU(cudaGetLastError());
  return st.download(h, s);
}

int sdorb_distinctive_batch(sdorb_handle* h, const uint8_t* desc, const int32_t* offsets, int nsets, int32_t* best_idx,
                            int32_t* best_median, int mem, void* stream) {
  if (!h || nsets < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (nsets == 0) return SDORB_OK;
  if (!desc || !offsets || !best_idx) return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  cudaStream_t s = (mem == SDORB_MEM_DEVICE && stream) ? (cudaStream_t)stream : h->s_compute;
  if (mem == SDORB_MEM_DEVICE) {
    if ((uintptr_t)desc % 16) return SDORB_ERR_BAD_ARG;
    StageScope st(h, s, SDORB_STAGE_MATCH);
    launch_distinctive(desc, offsets, nsets, best_idx, best_median, s);
    st.launched();
    CU(cudaGetLastError());
    return SDORB_OK;
  }
  for (int i = 0; i < nsets; ++i)
    if (offsets[i] < 0 || offsets[i + 1] < offsets[i] || offsets[i + 1] - offsets[i] > 65535) return SDORB_ERR_BAD_ARG;
  const size_t rows = (size_t)offsets[nsets], bD = std::max<size_t>(rows * 32, 32), bO = sizeof(int32_t) * ((size_t)nsets + 1),
               bR = sizeof(int32_t) * (size_t)nsets;
  auto up = up256;
  const int rcb = ensure_match_buf(h, up(bD) + up(bO) + 2 * up(bR));
  if (rcb) return rcb;
  uint8_t* pD = (uint8_t*)h->d_match_buf;
  int32_t* pO = (int32_t*)(pD + up(bD));
  int32_t* pI = (int32_t*)((uint8_t*)pO + up(bO));
  int32_t* pM = (int32_t*)((uint8_t*)pI + up(bR));
  if (rows) CU(cudaMemcpyAsync(pD, desc, rows * 32, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(pO, offsets, bO, cudaMemcpyHostToDevice, s));
  {
    StageScope st(h, s, SDORB_STAGE_MATCH);
    launch_distinctive(pD, pO, nsets, pI, pM, s);
    st.launched();
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(best_idx, pI, bR, cudaMemcpyDeviceToHost, s));
  if (best_median) CU(cudaMemcpyAsync(best_median, pM, bR, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SDORB_OK;
}

int sdorb_assign_grid_batch(sdorb_handle* h, const sdorb_keypoint* kps, const int32_t* counts, int nframes, int capacity,
                            float min_x, float min_y, float inv_w, float inv_h, int32_t* cell_start, int32_t* indices, int mem,
                            void* stream) {
  if (!h || nframes < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (nframes == 0) return SDORB_OK;
  if (!kps || !counts || !cell_start || !indices || capacity <= 0 || capacity > 65535) return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  cudaStream_t s = (mem == SDORB_MEM_DEVICE && stream) ? (cudaStream_t)stream : h->s_compute;
  const size_t ncs = (size_t)SDORB_GRID_COLS * SDORB_GRID_ROWS + 1;
  if (mem == SDORB_MEM_DEVICE) {
    StageScope st(h, s, SDORB_STAGE_MATCH);
    launch_assign_grid(kps, counts, nframes, capacity, min_x, min_y, inv_w, inv_h, cell_start, indices, s);
    st.launched();
    CU(cudaGetLastError());
    return SDORB_OK;
  }
  const size_t bK = sizeof(sdorb_keypoint) * (size_t)nframes * capacity, bC = sizeof(int32_t) * (size_t)nframes,
               bS = sizeof(int32_t) * ncs * nframes, bI = sizeof(int32_t) * (size_t)nframes * capacity;
  int rc = ensure_match_buf(h, up256(bK) + up256(bC) + up256(bS) + up256(bI));
  if (rc) return rc;
  uint8_t* base = (uint8_t*)h->d_match_buf;
  sdorb_keypoint* dK = (sdorb_keypoint*)base;
  int32_t* dC = (int32_t*)(base + up256(bK));
  int32_t* dS = (int32_t*)((uint8_t*)dC + up256(bC));
  int32_t* dI = (int32_t*)((uint8_t*)dS + up256(bS));
  CU(cudaMemcpyAsync(dK, kps, bK, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(dC, counts, bC, cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(dI, 0xFF, bI, s));  // unused tail of every frame's index list reads -1
  {
    StageScope st(h, s, SDORB_STAGE_MATCH);
    launch_assign_grid(dK, dC, nframes, capacity, min_x, min_y, inv_w, inv_h, dS, dI, s);
    st.launched();
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(cell_start, dS, bS, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(indices, dI, bI, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SDORB_OK;
}

// ---- guided matchers -------------------------------------------------------------------------------------------------
namespace {
const int kSearchMaxCapacity = 16384;
}  // namespace

int sdorb_search_for_initialization_batch(sdorb_handle* h, const sdorb_keypoint* kps1, const uint8_t* desc1, const int32_t* n1,
                                          const sdorb_keypoint* kps2, const uint8_t* desc2, const int32_t* n2,
                                          const sdorb_frame_grid* grid2, int npairs, int capacity, float* prev_matched,
                                          int window_size, float nnratio, int check_orientation, int32_t* matches12,
                                          int32_t* nmatches, int mem, void* stream) {
  if (!h || npairs < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (npairs == 0) return SDORB_OK;
  if (!kps1 || !desc1 || !n1 || !kps2 || !desc2 || !n2 || !grid2 || !grid2->cell_start || !grid2->indices || !prev_matched ||
      !matches12 || !nmatches || capacity <= 0 || capacity > kSearchMaxCapacity)
    return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  cudaStream_t s = (mem == SDORB_MEM_DEVICE && stream) ? (cudaStream_t)stream : h->s_compute;
  SearchInitArgs a;
  a.capacity = capacity;
  a.window_size = window_size;
  a.th_low = 50;  // ORBmatcher::TH_LOW, src/ORBmatcher.cc:37
  a.check_orientation = check_orientation ? 1 : 0;
  a.nnratio = nnratio;
  a.grid.min_x = grid2->min_x;
  a.grid.min_y = grid2->min_y;
  a.grid.inv_w = grid2->inv_w;
  a.grid.inv_h = grid2->inv_h;
  const size_t P = (size_t)npairs, C = (size_t)capacity, ncs = (size_t)SDORB_GRID_COLS * SDORB_GRID_ROWS + 1;
  Stager st;
  if (mem == SDORB_MEM_DEVICE) {
    if (((uintptr_t)desc1 | (uintptr_t)desc2) % 16) return SDORB_ERR_BAD_ARG;
    a.kps1 = kps1; a.desc1 = desc1; a.n1 = n1; a.kps2 = kps2; a.desc2 = desc2; a.n2 = n2;
    a.grid.cell_start = grid2->cell_start; a.grid.indices = grid2->indices;
    a.prev_matched = prev_matched; a.matches12 = matches12; a.nmatches = nmatches;
  } else {
    const size_t iK1 = st.add(kps1, nullptr, sizeof(sdorb_keypoint) * P * C), iD1 = st.add(desc1, nullptr, 32 * P * C),
                 iN1 = st.add(n1, nullptr, 4 * P), iK2 = st.add(kps2, nullptr, sizeof(sdorb_keypoint) * P * C),
                 iD2 = st.add(desc2, nullptr, 32 * P * C), iN2 = st.add(n2, nullptr, 4 * P),
                 iCS = st.add(grid2->cell_start, nullptr, 4 * ncs * P), iIX = st.add(grid2->indices, nullptr, 4 * P * C),
                 iPM = st.add(prev_matched, prev_matched, 8 * P * C), iM = st.add(nullptr, matches12, 4 * P * C),
                 iNM = st.add(nullptr, nmatches, 4 * P);
    int rc = st.upload(h, s);
    if (rc) return rc;
    a.kps1 = (void*)st.dev(h, iK1); a.desc1 = (uint8_t*)st.dev(h, iD1); a.n1 = (int32_t*)st.dev(h, iN1);
    a.kps2 = (void*)st.dev(h, iK2); a.desc2 = (uint8_t*)st.dev(h, iD2); a.n2 = (int32_t*)st.dev(h, iN2);
    a.grid.cell_start = (int32_t*)st.dev(h, iCS); a.grid.indices = (int32_t*)st.dev(h, iIX);
    a.prev_matched = (float*)st.dev(h, iPM); a.matches12 = (int32_t*)st.dev(h, iM); a.nmatches = (int32_t*)st.dev(h, iNM);
  }
  {
    StageScope sc(h, s, SDORB_STAGE_MATCH);
    launch_search_init(a, npairs, s);
    sc.launched();
  }
  CU(cudaGetLastError());
  if (mem == SDORB_MEM_HOST) return st.download(h, s);
  return SDORB_OK;
}

int sdorb_search_by_projection_batch(sdorb_handle* h, const sdorb_projection_search* q, int npairs, int capacity, int32_t* assigned,
                                     int32_t* nmatches, int mem, void* stream) {
  if (!h || npairs < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (npairs == 0) return SDORB_OK;
  if (!q || !q->kps_last || !q->kps_last_un || !q->proj || !q->flags_last || !q->desc_mp || !q->n_last || !q->kps_cur_un ||
      !q->desc_cur || !q->u_right_cur || !q->occupied_cur || !q->n_cur || !q->grid_cur.cell_start || !q->grid_cur.indices ||
      !q->scale_factors || q->nlevels <= 0 || q->nlevels > SDORB_MAX_LEVELS || q->mode < 0 || q->mode > 2 || !assigned ||
      !nmatches || capacity <= 0 || capacity > kSearchMaxCapacity)
    return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  cudaStream_t s = (mem == SDORB_MEM_DEVICE && stream) ? (cudaStream_t)stream : h->s_compute;
  SearchProjArgs a;
  a.capacity = capacity;
  a.mode = q->mode;
  a.th_high = q->orb_dist > 0 ? q->orb_dist : 100;  // ORBmatcher::TH_HIGH, src/ORBmatcher.cc:36, or the KeyFrame overload's ORBdist
  a.check_orientation = q->check_orientation ? 1 : 0;
  a.th = q->th;
  a.mbf = q->mbf;
  a.min_x = q->bounds[0]; a.max_x = q->bounds[1]; a.min_y = q->bounds[2]; a.max_y = q->bounds[3];
  for (int l = 0; l < SDORB_MAX_LEVELS; ++l) a.scale_factors[l] = q->scale_factors[std::min(l, q->nlevels - 1)];
  a.grid.min_x = q->grid_cur.min_x; a.grid.min_y = q->grid_cur.min_y;
  a.grid.inv_w = q->grid_cur.inv_w; a.grid.inv_h = q->grid_cur.inv_h;
  const size_t P = (size_t)npairs, C = (size_t)capacity, ncs = (size_t)SDORB_GRID_COLS * SDORB_GRID_ROWS + 1;
  Stager st;
  if (mem == SDORB_MEM_DEVICE) {
    if (((uintptr_t)q->desc_mp | (uintptr_t)q->desc_cur) % 16) return SDORB_ERR_BAD_ARG;
    a.kps_last = q->kps_last; a.kps_last_un = q->kps_last_un; a.proj = q->proj; a.flags_last = q->flags_last;
    a.desc_mp = q->desc_mp; a.n_last = q->n_last; a.kps_cur_un = q->kps_cur_un; a.desc_cur = q->desc_cur;
    a.u_right_cur = q->u_right_cur; a.occupied_cur = q->occupied_cur; a.n_cur = q->n_cur;
    a.grid.cell_start = q->grid_cur.cell_start; a.grid.indices = q->grid_cur.indices;
    a.assigned = assigned; a.nmatches = nmatches;
  } else {
    const size_t bK = sizeof(sdorb_keypoint) * P * C;
    const size_t iKL = st.add(q->kps_last, nullptr, bK), iKU = st.add(q->kps_last_un, nullptr, bK),
                 iPR = st.add(q->proj, nullptr, 12 * P * C), iFL = st.add(q->flags_last, nullptr, P * C),
                 iDM = st.add(q->desc_mp, nullptr, 32 * P * C), iNL = st.add(q->n_last, nullptr, 4 * P),
                 iKC = st.add(q->kps_cur_un, nullptr, bK), iDC = st.add(q->desc_cur, nullptr, 32 * P * C),
                 iUR = st.add(q->u_right_cur, nullptr, 4 * P * C), iOC = st.add(q->occupied_cur, nullptr, P * C),
                 iNC = st.add(q->n_cur, nullptr, 4 * P), iCS = st.add(q->grid_cur.cell_start, nullptr, 4 * ncs * P),
                 iIX = st.add(q->grid_cur.indices, nullptr, 4 * P * C), iAS = st.add(nullptr, assigned, 4 * P * C),
                 iNM = st.add(nullptr, nmatches, 4 * P);
    int rc = st.upload(h, s);
    if (rc) return rc;
    a.kps_last = (void*)st.dev(h, iKL); a.kps_last_un = (void*)st.dev(h, iKU); a.proj = (float*)st.dev(h, iPR);
    a.flags_last = (uint8_t*)st.dev(h, iFL); a.desc_mp = (uint8_t*)st.dev(h, iDM); a.n_last = (int32_t*)st.dev(h, iNL);
    a.kps_cur_un = (void*)st.dev(h, iKC); a.desc_cur = (uint8_t*)st.dev(h, iDC); a.u_right_cur = (float*)st.dev(h, iUR);
    a.occupied_cur = (uint8_t*)st.dev(h, iOC); a.n_cur = (int32_t*)st.dev(h, iNC);
    a.grid.cell_start = (int32_t*)st.dev(h, iCS); a.grid.indices = (int32_t*)st.dev(h, iIX);
    a.assigned = (int32_t*)st.dev(h, iAS); a.nmatches = (int32_t*)st.dev(h, iNM);
  }
  {
    StageScope sc(h, s, SDORB_STAGE_MATCH);
    launch_search_projection(a, npairs, s);
    sc.launched();
  }
  CU(cudaGetLastError());
  if (mem == SDORB_MEM_HOST) return st.download(h, s);
  return SDORB_OK;
}

int sdorb_search_map_points_batch(sdorb_handle* h, const sdorb_map_point_search* q, int nframes, int capacity_mp, int capacity,
                                  int32_t* assigned, int32_t* nmatches, int mem, void* stream) {
  if (!h || nframes < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (nframes == 0) return SDORB_OK;
  const bool sim3 = q && q->sim3_form != 0;
  if (!q || !q->proj || (!sim3 && !q->view_cos) || !q->level || !q->flags || !q->desc_mp || !q->n_mp || !q->kps_un || !q->desc ||
      (!sim3 && !q->u_right) || !q->occupied || !q->n_frame || !q->grid.cell_start || !q->grid.indices || !q->scale_factors || q->nlevels <= 0 ||
      q->nlevels > SDORB_MAX_LEVELS || !assigned || !nmatches || capacity <= 0 || capacity > kSearchMaxCapacity || capacity_mp <= 0)
    return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  cudaStream_t s = (mem == SDORB_MEM_DEVICE && stream) ? (cudaStream_t)stream : h->s_compute;
  SearchPointsArgs a;
  a.capacity = capacity;
  a.capacity_mp = capacity_mp;
  a.th_high = sim3 ? 50 : 100;  // ORBmatcher::TH_LOW (:243) / TH_HIGH (:100), src/ORBmatcher.cc:36-37
  a.sim3_form = sim3 ? 1 : 0;
  a.th = q->th;
  a.nnratio = q->nnratio;
  for (int l = 0; l < SDORB_MAX_LEVELS; ++l) a.scale_factors[l] = q->scale_factors[std::min(l, q->nlevels - 1)];
  a.grid.min_x = q->grid.min_x; a.grid.min_y = q->grid.min_y;
  a.grid.inv_w = q->grid.inv_w; a.grid.inv_h = q->grid.inv_h;
  const size_t P = (size_t)nframes, C = (size_t)capacity, M = (size_t)capacity_mp, ncs = (size_t)SDORB_GRID_COLS * SDORB_GRID_ROWS + 1;
  Stager st;
  if (mem == SDORB_MEM_DEVICE) {
    if (((uintptr_t)q->desc_mp | (uintptr_t)q->desc) % 16) return SDORB_ERR_BAD_ARG;
    a.proj = q->proj; a.view_cos = q->view_cos; a.level = q->level; a.flags = q->flags; a.desc_mp = q->desc_mp; a.n_mp = q->n_mp;
    a.kps = q->kps_un; a.desc = q->desc; a.u_right = q->u_right; a.occupied = q->occupied; a.n_frame = q->n_frame;
    a.grid.cell_start = q->grid.cell_start; a.grid.indices = q->grid.indices;
    a.assigned = assigned; a.nmatches = nmatches;
  } else {
    // view_cos / u_right are not read in the Sim3 form: nothing is uploaded for them (their slots stay unread)
    const size_t iPR = st.add(q->proj, nullptr, 12 * P * M), iVC = st.add(sim3 ? nullptr : q->view_cos, nullptr, sim3 ? 4 : 4 * P * M),
                 iLV = st.add(q->level, nullptr, 4 * P * M), iFL = st.add(q->flags, nullptr, P * M),
                 iDM = st.add(q->desc_mp, nullptr, 32 * P * M), iNM = st.add(q->n_mp, nullptr, 4 * P),
                 iK = st.add(q->kps_un, nullptr, sizeof(sdorb_keypoint) * P * C), iD = st.add(q->desc, nullptr, 32 * P * C),
                 iUR = st.add(sim3 ? nullptr : q->u_right, nullptr, sim3 ? 4 : 4 * P * C), iOC = st.add(q->occupied, nullptr, P * C),
                 iNF = st.add(q->n_frame, nullptr, 4 * P), iCS = st.add(q->grid.cell_start, nullptr, 4 * ncs * P),
                 iIX = st.add(q->grid.indices, nullptr, 4 * P * C), iAS = st.add(nullptr, assigned, 4 * P * C),
                 iNO = st.add(nullptr, nmatches, 4 * P);
    int rc = st.upload(h, s);
    if (rc) return rc;
    a.proj = (float*)st.dev(h, iPR); a.view_cos = (float*)st.dev(h, iVC); a.level = (int32_t*)st.dev(h, iLV);
    a.flags = (uint8_t*)st.dev(h, iFL); a.desc_mp = (uint8_t*)st.dev(h, iDM); a.n_mp = (int32_t*)st.dev(h, iNM);
    a.kps = (void*)st.dev(h, iK); a.desc = (uint8_t*)st.dev(h, iD); a.u_right = (float*)st.dev(h, iUR);
    a.occupied = (uint8_t*)st.dev(h, iOC); a.n_frame = (int32_t*)st.dev(h, iNF);
    a.grid.cell_start = (int32_t*)st.dev(h, iCS); a.grid.indices = (int32_t*)st.dev(h, iIX);
    a.assigned = (int32_t*)st.dev(h, iAS); a.nmatches = (int32_t*)st.dev(h, iNO);
  }
  {
    StageScope sc(h, s, SDORB_STAGE_MATCH);
    launch_search_points(a, nframes, s);
    sc.launched();
  }
  CU(cudaGetLastError());
  if (mem == SDORB_MEM_HOST) return st.download(h, s);
  return SDORB_OK;
}

int sdorb_search_by_points_batch(sdorb_handle* h, const sdorb_keypoint* kps1, const uint8_t* desc1, const uint8_t* valid1,
                                 const int32_t* n1, const sdorb_keypoint* kps2, const uint8_t* desc2, const uint8_t* valid2,
                                 const int32_t* n2, int npairs, int capacity, float nnratio, int check_orientation, int32_t* matches12,
                                 int32_t* nmatches, int mem, void* stream) {
  if (!h || npairs < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (npairs == 0) return SDORB_OK;
  if (!kps1 || !desc1 || !valid1 || !n1 || !kps2 || !desc2 || !valid2 || !n2 || !matches12 || !nmatches || capacity <= 0 ||
      capacity > kSearchMaxCapacity)
    return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  cudaStream_t s = (mem == SDORB_MEM_DEVICE && stream) ? (cudaStream_t)stream : h->s_compute;
  SearchByPointsArgs a;
  a.capacity = capacity;
  a.th_low = 50;  // ORBmatcher::TH_LOW, src/ORBmatcher.cc:37
  a.check_orientation = check_orientation ? 1 : 0;
  a.nnratio = nnratio;
  const size_t P = (size_t)npairs, C = (size_t)capacity;
  Stager st;
  if (mem == SDORB_MEM_DEVICE) {
    if (((uintptr_t)desc1 | (uintptr_t)desc2) % 16) return SDORB_ERR_BAD_ARG;
    a.kps1 = kps1; a.desc1 = desc1; a.valid1 = valid1; a.n1 = n1; a.kps2 = kps2; a.desc2 = desc2; a.valid2 = valid2; a.n2 = n2;
    a.matches12 = matches12; a.nmatches = nmatches;
  } else {
    const size_t bK = sizeof(sdorb_keypoint) * P * C;
    const size_t iK1 = st.add(kps1, nullptr, bK), iD1 = st.add(desc1, nullptr, 32 * P * C), iV1 = st.add(valid1, nullptr, P * C),
                 iN1 = st.add(n1, nullptr, 4 * P), iK2 = st.add(kps2, nullptr, bK), iD2 = st.add(desc2, nullptr, 32 * P * C),
                 iV2 = st.add(valid2, nullptr, P * C), iN2 = st.add(n2, nullptr, 4 * P), iM = st.add(nullptr, matches12, 4 * P * C),
                 iNM = st.add(nullptr, nmatches, 4 * P);
    int rc = st.upload(h, s);
    if (rc) return rc;
    a.kps1 = (void*)st.dev(h, iK1); a.desc1 = (uint8_t*)st.dev(h, iD1); a.valid1 = (uint8_t*)st.dev(h, iV1); a.n1 = (int32_t*)st.dev(h, iN1);
    a.kps2 = (void*)st.dev(h, iK2); a.desc2 = (uint8_t*)st.dev(h, iD2); a.valid2 = (uint8_t*)st.dev(h, iV2); a.n2 = (int32_t*)st.dev(h, iN2);
    a.matches12 = (int32_t*)st.dev(h, iM); a.nmatches = (int32_t*)st.dev(h, iNM);
  }
  {
    StageScope sc(h, s, SDORB_STAGE_MATCH);
    launch_search_by_points(a, npairs, s);
    sc.launched();
  }
  CU(cudaGetLastError());
  if (mem == SDORB_MEM_HOST) return st.download(h, s);
  return SDORB_OK;
}

namespace {
// Staging of one sdorb_fuse_search (host memory) / pointer pass-through (device memory), shared by the Fuse and SearchBySim3 entries.
struct FuseStage {
  size_t iPR, iLV, iFL, iDM, iNM, iK, iD, iUR, iCS, iIX;
  static bool valid(const sdorb_fuse_search* q, int capacity_mp, int capacity) {
    return q && q->proj && q->level && q->flags && q->desc_mp && q->n_mp && q->kps_un && q->desc && q->grid.cell_start && q->grid.indices &&
           q->scale_factors && q->nlevels > 0 && q->nlevels <= SDORB_MAX_LEVELS && capacity > 0 && capacity <= kSearchMaxCapacity &&
           capacity_mp > 0 && (!q->check_reprojection || (q->u_right && q->inv_level_sigma2));
  }
  static void scalars(const sdorb_fuse_search* q, int capacity_mp, int capacity, int th_dist, FuseSearchArgs& a) {
    a.capacity = capacity;
    a.capacity_mp = capacity_mp;
    a.th_dist = th_dist;
    a.check_reprojection = q->check_reprojection ? 1 : 0;
    a.th = q->th;
    for (int l = 0; l < SDORB_MAX_LEVELS; ++l) {
      a.scale_factors[l] = q->scale_factors[std::min(l, q->nlevels - 1)];
      a.inv_sigma2[l] = q->check_reprojection ? q->inv_level_sigma2[std::min(l, q->nlevels - 1)] : 0.f;
    }
    a.grid.min_x = q->grid.min_x; a.grid.min_y = q->grid.min_y;
    a.grid.inv_w = q->grid.inv_w; a.grid.inv_h = q->grid.inv_h;
  }
  static bool device_ok(const sdorb_fuse_search* q) { return (((uintptr_t)q->desc_mp | (uintptr_t)q->desc) % 16) == 0; }
  static void device(const sdorb_fuse_search* q, FuseSearchArgs& a) {
    a.proj = q->proj; a.level = q->level; a.flags = q->flags; a.desc_mp = q->desc_mp; a.n_mp = q->n_mp;
    a.kps = q->kps_un; a.desc = q->desc; a.u_right = q->u_right;
    a.grid.cell_start = q->grid.cell_start; a.grid.indices = q->grid.indices;
  }
  void add(Stager& st, const sdorb_fuse_search* q, size_t P, size_t M, size_t C) {
    const size_t ncs = (size_t)SDORB_GRID_COLS * SDORB_GRID_ROWS + 1;
    const bool ur = q->check_reprojection != 0;  // mvuRight is only read by the reprojection gate
    iPR = st.add(q->proj, nullptr, 12 * P * M); iLV = st.add(q->level, nullptr, 4 * P * M); iFL = st.add(q->flags, nullptr, P * M);
    iDM = st.add(q->desc_mp, nullptr, 32 * P * M); iNM = st.add(q->n_mp, nullptr, 4 * P);
    iK = st.add(q->kps_un, nullptr, sizeof(sdorb_keypoint) * P * C); iD = st.add(q->desc, nullptr, 32 * P * C);
    iUR = st.add(ur ? q->u_right : nullptr, nullptr, ur ? 4 * P * C : 4); iCS = st.add(q->grid.cell_start, nullptr, 4 * ncs * P);
    iIX = st.add(q->grid.indices, nullptr, 4 * P * C);
  }
  void resolve(sdorb_handle* h, const Stager& st, FuseSearchArgs& a) const {
    a.proj = (float*)st.dev(h, iPR); a.level = (int32_t*)st.dev(h, iLV); a.flags = (uint8_t*)st.dev(h, iFL);
    a.desc_mp = (uint8_t*)st.dev(h, iDM); a.n_mp = (int32_t*)st.dev(h, iNM);
    a.kps = (void*)st.dev(h, iK); a.desc = (uint8_t*)st.dev(h, iD); a.u_right = (float*)st.dev(h, iUR);
    a.grid.cell_start = (int32_t*)st.dev(h, iCS); a.grid.indices = (int32_t*)st.dev(h, iIX);
  }
};
}  // namespace

int sdorb_fuse_search_batch(sdorb_handle* h, const sdorb_fuse_search* q, int nframes, int capacity_mp, int capacity, int32_t* best_idx,
                            int32_t* best_dist, int mem, void* stream) {
  if (!h || nframes < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (nframes == 0) return SDORB_OK;
  if (!FuseStage::valid(q, capacity_mp, capacity) || !best_idx || !best_dist || q->th_dist < 0 || q->th_dist > 256 ||
      nframes > SDORB_MAX_GRID_BATCH)
    return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  cudaStream_t s = (mem == SDORB_MEM_DEVICE && stream) ? (cudaStream_t)stream : h->s_compute;
  FuseSearchArgs a;
  FuseStage::scalars(q, capacity_mp, capacity, q->th_dist, a);
  const size_t P = (size_t)nframes, C = (size_t)capacity, M = (size_t)capacity_mp;
  Stager st;
  if (mem == SDORB_MEM_DEVICE) {
    if (!FuseStage::device_ok(q)) return SDORB_ERR_BAD_ARG;
    FuseStage::device(q, a);
    a.best_idx = best_idx; a.best_dist = best_dist;
  } else {
    FuseStage fs;
    fs.add(st, q, P, M, C);
    const size_t iBI = st.add(nullptr, best_idx, 4 * P * M), iBD = st.add(nullptr, best_dist, 4 * P * M);
    int rc = st.upload(h, s);
    if (rc) return rc;
    fs.resolve(h, st, a);
    a.best_idx = (int32_t*)st.dev(h, iBI); a.best_dist = (int32_t*)st.dev(h, iBD);
  }
  {
    StageScope sc(h, s, SDORB_STAGE_MATCH);
    launch_fuse_search(a, nframes, s);
    sc.launched();
  }
  CU(cudaGetLastError());
  if (mem == SDORB_MEM_HOST) return st.download(h, s);
  return SDORB_OK;
}

int sdorb_search_by_sim3_batch(sdorb_handle* h, const sdorb_fuse_search* q12, const sdorb_fuse_search* q21, int npairs, int capacity,
                               int32_t* match1, int32_t* match2, int32_t* matches12, int32_t* nfound, int mem, void* stream) {
  if (!h || npairs < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (npairs == 0) return SDORB_OK;
  if (!FuseStage::valid(q12, capacity, capacity) || !FuseStage::valid(q21, capacity, capacity) || q12->check_reprojection ||
      q21->check_reprojection || !match1 || !match2 || !matches12 || !nfound || npairs > SDORB_MAX_GRID_BATCH)
    return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  cudaStream_t s = (mem == SDORB_MEM_DEVICE && stream) ? (cudaStream_t)stream : h->s_compute;
  FuseSearchArgs a12, a21;
  FuseStage::scalars(q12, capacity, capacity, 100, a12);  // ORBmatcher::TH_HIGH, src/ORBmatcher.cc:36 (:843, :921)
  FuseStage::scalars(q21, capacity, capacity, 100, a21);
  a12.best_dist = a21.best_dist = nullptr;
  const size_t P = (size_t)npairs, C = (size_t)capacity;
  int32_t *d_m12 = matches12, *d_nf = nfound;
  Stager st;
  if (mem == SDORB_MEM_DEVICE) {
    if (!FuseStage::device_ok(q12) || !FuseStage::device_ok(q21)) return SDORB_ERR_BAD_ARG;
    FuseStage::device(q12, a12);
    FuseStage::device(q21, a21);
    a12.best_idx = match1; a21.best_idx = match2;
  } else {
    FuseStage f12, f21;
    f12.add(st, q12, P, C, C);
    f21.add(st, q21, P, C, C);
    const size_t iM1 = st.add(nullptr, match1, 4 * P * C), iM2 = st.add(nullptr, match2, 4 * P * C),
                 iM12 = st.add(nullptr, matches12, 4 * P * C), iNF = st.add(nullptr, nfound, 4 * P);
    int rc = st.upload(h, s);
    if (rc) return rc;
    f12.resolve(h, st, a12);
    f21.resolve(h, st, a21);
    a12.best_idx = (int32_t*)st.dev(h, iM1); a21.best_idx = (int32_t*)st.dev(h, iM2);
    d_m12 = (int32_t*)st.dev(h, iM12); d_nf = (int32_t*)st.dev(h, iNF);
  }
  {
    StageScope sc(h, s, SDORB_STAGE_MATCH);
    launch_fuse_search(a12, npairs, s);
    sc.launched();
    launch_fuse_search(a21, npairs, s);
    sc.launched();
    launch_sim3_agreement(a12.best_idx, a21.best_idx, a12.n_mp, capacity, d_m12, d_nf, npairs, s);
    sc.launched();
  }
  CU(cudaGetLastError());
  if (mem == SDORB_MEM_HOST) return st.download(h, s);
  return SDORB_OK;
}

int sdorb_search_for_triangulation_batch(sdorb_handle* h, const sdorb_triangulation_search* q, int npairs, int capacity,
                                         int32_t* matches12, int32_t* nmatches, int mem, void* stream) {
  if (!h || npairs < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (npairs == 0) return SDORB_OK;
  if (!q || !q->kps1_un || !q->desc1 || !q->has_mp1 || !q->u_right1 || !q->n1 || !q->kps2_un || !q->desc2 || !q->has_mp2 ||
      !q->u_right2 || !q->n2 || !q->F12 || !q->epipole || !q->scale_factors || !q->level_sigma2 || q->nlevels <= 0 ||
      q->nlevels > SDORB_MAX_LEVELS || !matches12 || !nmatches || capacity <= 0 || capacity > kSearchMaxCapacity)
    return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  cudaStream_t s = (mem == SDORB_MEM_DEVICE && stream) ? (cudaStream_t)stream : h->s_compute;
  SearchTriArgs a;
  a.capacity = capacity;
  a.th_low = 50;  // ORBmatcher::TH_LOW, src/ORBmatcher.cc:37
  a.check_orientation = q->check_orientation ? 1 : 0;
  for (int l = 0; l < SDORB_MAX_LEVELS; ++l) {
    const int ll = std::min(l, q->nlevels - 1);
    // dsqr < 3.84 * sigma2 is a double comparison of a float (src/ORBmatcher.cc:143): the same as dsqr < the smallest float >= the product
    const double thr = 3.84 * (double)q->level_sigma2[ll];
    float g = (float)thr;
    if ((double)g < thr) g = std::nextafter(g, INFINITY);
    a.gate[l] = g;
    a.eplim[l] = 100 * q->scale_factors[ll];
  }
  const size_t P = (size_t)npairs, C = (size_t)capacity;
  Stager st;
  if (mem == SDORB_MEM_DEVICE) {
    if (((uintptr_t)q->desc1 | (uintptr_t)q->desc2) % 16 || (uintptr_t)q->F12 % 8) return SDORB_ERR_BAD_ARG;
    a.kps1 = q->kps1_un; a.desc1 = q->desc1; a.has_mp1 = q->has_mp1; a.u_right1 = q->u_right1; a.n1 = q->n1;
    a.kps2 = q->kps2_un; a.desc2 = q->desc2; a.has_mp2 = q->has_mp2; a.u_right2 = q->u_right2; a.n2 = q->n2;
    a.F12 = q->F12; a.epipole = q->epipole; a.matches12 = matches12; a.nmatches = nmatches;
  } else {
    const size_t bK = sizeof(sdorb_keypoint) * P * C;
    const size_t iK1 = st.add(q->kps1_un, nullptr, bK), iD1 = st.add(q->desc1, nullptr, 32 * P * C), iM1 = st.add(q->has_mp1, nullptr, P * C),
                 iR1 = st.add(q->u_right1, nullptr, 4 * P * C), iN1 = st.add(q->n1, nullptr, 4 * P), iK2 = st.add(q->kps2_un, nullptr, bK),
                 iD2 = st.add(q->desc2, nullptr, 32 * P * C), iM2 = st.add(q->has_mp2, nullptr, P * C),
                 iR2 = st.add(q->u_right2, nullptr, 4 * P * C), iN2 = st.add(q->n2, nullptr, 4 * P), iF = st.add(q->F12, nullptr, 72 * P),
                 iE = st.add(q->epipole, nullptr, 8 * P), iM = st.add(nullptr, matches12, 4 * P * C), iNM = st.add(nullptr, nmatches, 4 * P);
    int rc = st.upload(h, s);
    if (rc) return rc;
    a.kps1 = (void*)st.dev(h, iK1); a.desc1 = (uint8_t*)st.dev(h, iD1); a.has_mp1 = (uint8_t*)st.dev(h, iM1);
    a.u_right1 = (float*)st.dev(h, iR1); a.n1 = (int32_t*)st.dev(h, iN1);
    a.kps2 = (void*)st.dev(h, iK2); a.desc2 = (uint8_t*)st.dev(h, iD2); a.has_mp2 = (uint8_t*)st.dev(h, iM2);
    a.u_right2 = (float*)st.dev(h, iR2); a.n2 = (int32_t*)st.dev(h, iN2);
    a.F12 = (double*)st.dev(h, iF); a.epipole = (float*)st.dev(h, iE);
    a.matches12 = (int32_t*)st.dev(h, iM); a.nmatches = (int32_t*)st.dev(h, iNM);
  }
  {
    StageScope sc(h, s, SDORB_STAGE_MATCH);
    launch_search_triangulation(a, npairs, s);
    sc.launched();
  }
  CU(cudaGetLastError());
  if (mem == SDORB_MEM_HOST) return st.download(h, s);
  return SDORB_OK;
}

int sdorb_host_image_bounds(int cols, int rows, const float* K, const float* dist, int ndist, float* bounds) {
  if (!K || !bounds || cols <= 0 || rows <= 0 || ndist < 0 || ndist > 12 || (ndist > 0 && !dist)) return SDORB_ERR_BAD_ARG;
  host_image_bounds(cols, rows, K, dist, ndist, bounds);
  return SDORB_OK;
}

int sdorb_undistort_keypoints_batch(sdorb_handle* h, const sdorb_keypoint* kps, const int32_t* counts, int nframes, int capacity,
                                    const float* K, const float* dist, int ndist, sdorb_keypoint* out, int mem, void* stream) {
  if (!h || nframes < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (nframes == 0) return SDORB_OK;
  if (!kps || !counts || !out || !K || capacity <= 0 || ndist < 0 || ndist > 12 || (ndist > 0 && !dist) || nframes > SDORB_MAX_GRID_BATCH)
    return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  cudaStream_t s = (mem == SDORB_MEM_DEVICE && stream) ? (cudaStream_t)stream : h->s_compute;
  if (mem == SDORB_MEM_DEVICE) {
    StageScope st(h, s, SDORB_STAGE_MATCH);
    launch_undistort(kps, counts, nframes, capacity, K, dist, ndist, out, s);
    st.launched();
    CU(cudaGetLastError());
    return SDORB_OK;
  }
  const size_t bK = sizeof(sdorb_keypoint) * (size_t)nframes * capacity, bC = sizeof(int32_t) * (size_t)nframes;
  int rc = ensure_match_buf(h, 2 * up256(bK) + up256(bC));
  if (rc) return rc;
  uint8_t* base = (uint8_t*)h->d_match_buf;
  sdorb_keypoint* dK = (sdorb_keypoint*)base;
  sdorb_keypoint* dO = (sdorb_keypoint*)(base + up256(bK));
  int32_t* dC = (int32_t*)((uint8_t*)dO + up256(bK));
  CU(cudaMemcpyAsync(dK, kps, bK, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(dC, counts, bC, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(dO, dK, bK, cudaMemcpyDeviceToDevice, s));  // entries beyond counts[f] pass through unchanged
  {
    StageScope st(h, s, SDORB_STAGE_MATCH);
    launch_undistort(dK, dC, nframes, capacity, K, dist, ndist, dO, s);
    st.launched();
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, dO, bK, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SDORB_OK;
}

int sdorb_stereo_from_rgbd_batch(sdorb_handle* h, const sdorb_keypoint* kps, const sdorb_keypoint* kps_un, const int32_t* counts,
                                 int nframes, int capacity, const float* depth, int width, int height, size_t row_stride,
                                 size_t frame_stride, float mbf, float* u_right, float* z, int mem, void* stream) {
  if (!h || nframes < 0 || (mem != SDORB_MEM_HOST && mem != SDORB_MEM_DEVICE)) return SDORB_ERR_BAD_ARG;
  if (nframes == 0) return SDORB_OK;
  if (!kps || !kps_un || !counts || !depth || !u_right || !z || capacity <= 0 || width <= 0 || height <= 0 ||
      row_stride < (size_t)width || (nframes > 1 && frame_stride < row_stride * (size_t)(height - 1) + (size_t)width) ||
      nframes > SDORB_MAX_GRID_BATCH)
    return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  cudaStream_t s = (mem == SDORB_MEM_DEVICE && stream) ? (cudaStream_t)stream : h->s_compute;
  if (mem == SDORB_MEM_DEVICE) {
    StageScope st(h, s, SDORB_STAGE_MATCH);
    launch_stereo_rgbd(kps, kps_un, counts, nframes, capacity, depth, width, height, (int64_t)row_stride, (int64_t)frame_stride, mbf,
                       u_right, z, s);
    st.launched();
    CU(cudaGetLastError());
    return SDORB_OK;
  }
  const size_t bK = sizeof(sdorb_keypoint) * (size_t)nframes * capacity, bC = sizeof(int32_t) * (size_t)nframes,
               bD = sizeof(float) * (size_t)nframes * width * height, bO = sizeof(float) * (size_t)nframes * capacity;
  int rc = ensure_match_buf(h, 2 * up256(bK) + up256(bC) + up256(bD) + 2 * up256(bO));
  if (rc) return rc;
  uint8_t* base = (uint8_t*)h->d_match_buf;
  sdorb_keypoint* dK = (sdorb_keypoint*)base;
  sdorb_keypoint* dU = (sdorb_keypoint*)(base + up256(bK));
  int32_t* dC = (int32_t*)((uint8_t*)dU + up256(bK));
  float* dD = (float*)((uint8_t*)dC + up256(bC));
  float* dR = (float*)((uint8_t*)dD + up256(bD));
  float* dZ = (float*)((uint8_t*)dR + up256(bO));
  CU(cudaMemcpyAsync(dK, kps, bK, cudaMemcpyHostToDevice, s));
  if (kps_un != kps) CU(cudaMemcpyAsync(dU, kps_un, bK, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(dC, counts, bC, cudaMemcpyHostToDevice, s));
  for (int f = 0; f < nframes; ++f)  // repack to tight rows
    CU(cudaMemcpy2DAsync(dD + (size_t)f * width * height, sizeof(float) * (size_t)width, depth + (size_t)f * frame_stride,
                         sizeof(float) * row_stride, sizeof(float) * (size_t)width, height, cudaMemcpyHostToDevice, s));
  {
    StageScope st(h, s, SDORB_STAGE_MATCH);
    launch_stereo_rgbd(dK, kps_un != kps ? dU : dK, dC, nframes, capacity, dD, width, height, width, (int64_t)width * height, mbf, dR,
                       dZ, s);
    st.launched();
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(u_right, dR, bO, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(z, dZ, bO, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SDORB_OK;
}

int sdorb_set_profiling(sdorb_handle* h, int enabled) {
  if (!h) return SDORB_ERR_BAD_ARG;
  h->profiling = enabled != 0;
  return SDORB_OK;
}

int sdorb_get_stage_times(sdorb_handle* h, double* ms, int64_t* launches, int reset) {
  if (!h) return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  CU(cudaDeviceSynchronize());
  for (auto& p : h->pending) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) h->stage_ms[p.stage] += t;
    h->event_pool.push_back(p.a);
    h->event_pool.push_back(p.b);
  }
  h->pending.clear();
  for (int i = 0; i < SDORB_NUM_STAGES; ++i) {
    if (ms) ms[i] = h->stage_ms[i];
    if (launches) launches[i] = h->stage_launches[i];
    if (reset) {
      h->stage_ms[i] = 0;
      h->stage_launches[i] = 0;
    }
  }
  return SDORB_OK;
}

int64_t sdorb_kernel_launches(const sdorb_handle* h) { return h ? h->launches : 0; }

int64_t sdorb_debug_guard_check(sdorb_handle* h) {
  if (!h) return SDORB_ERR_BAD_ARG;
  if (!guard_on()) return -1000;  // not a guarded run
  DeviceGuard guard(h->device);
  if (cudaDeviceSynchronize() != cudaSuccess) return SDORB_ERR_CUDA;
  std::lock_guard<std::mutex> lk(g_guard_mu);
  h->cuda_error = g_guard_freed_msg;
  int64_t bad = g_guard_bad_freed;
  g_guard_bad_freed = 0;
  g_guard_freed_msg.clear();
  for (const auto& kv : g_guard_live) {
    const int64_t n = guard_scan(kv.second, &h->cuda_error);
    if (n < 0) return SDORB_ERR_CUDA;
    bad += n;
  }
  return bad;
}

int sdorb_debug_pipe_probe(sdorb_handle* h, int pipe, double* warp_instr_per_s, double* warp_instr_per_clk_per_sm) {
  if (!h || pipe < 0 || pipe > 7) return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  const int e = run_pipe_probe(pipe, h->s_compute, warp_instr_per_s, warp_instr_per_clk_per_sm);
  h->launches += 2;
  if (e) {
    h->cuda_error = std::string("pipe probe: ") + cudaGetErrorString((cudaError_t)e);
    return SDORB_ERR_CUDA;
  }
  return SDORB_OK;
}

int sdorb_debug_nth_element(sdorb_handle* h, uint32_t* entries, int n, int nth) {
  if (!h || !entries || n <= 0 || nth < 0 || nth >= n || n > 65535) return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  uint32_t* d = nullptr;
  CU(sd_malloc(&d, sizeof(uint32_t) * (size_t)n));
  int rc = SDORB_OK;
  if (cudaMemcpy(d, entries, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) rc = SDORB_ERR_CUDA;
  if (rc == SDORB_OK) {
    launch_debug_nth_element(d, n, nth, h->s_compute);
    h->launches += 1;
    if (cudaStreamSynchronize(h->s_compute) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = SDORB_ERR_CUDA;
  }
  if (rc == SDORB_OK && cudaMemcpy(entries, d, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = SDORB_ERR_CUDA;
  sd_free(d);
  return rc;
}

int64_t sdorb_debug_read(sdorb_handle* h, int what, int frame, int level, void* dst, size_t capacity) {
  if (!h || !dst || h->gw == 0 || frame < 0 || frame >= h->prm.max_batch || level < 0 || level >= h->geom.nlevels)
    return SDORB_ERR_BAD_ARG;
  DeviceGuard guard(h->device);
  CU(cudaDeviceSynchronize());
  const FrameGeom& g = h->geom;
  const LevelGeom& L = g.lv[level];
  const size_t B = (size_t)h->prm.max_batch;
  switch (what) {
    case SDORB_DBG_PYRAMID_LEVEL:
    case SDORB_DBG_BLURRED_LEVEL: {
      const size_t bytes = (size_t)L.w * L.h;
      if (capacity < bytes) return SDORB_ERR_CAPACITY;
      const uint8_t* base = what == SDORB_DBG_BLURRED_LEVEL ? h->d_blur : h->d_pyr;
      if (what == SDORB_DBG_PYRAMID_LEVEL && level == 0) return SDORB_ERR_UNSUPPORTED;  // level 0 is the caller's image
      CU(cudaMemcpy2D(dst, L.w, base + (size_t)L.plane_base * B + (size_t)frame * L.plane_bytes, L.pitch, L.w, L.h,
                      cudaMemcpyDeviceToHost));
      return (int64_t)bytes;
    }
    case SDORB_DBG_CELL_COUNTS: {
      const size_t n = (L.cols > 0 && L.rows > 0) ? (size_t)L.cols * L.rows : 0;
      if (capacity < n * 4) return SDORB_ERR_CAPACITY;
      if (n) CU(cudaMemcpy(dst, h->d_cell_seen + (size_t)frame * g.cells_total + L.cell_base, n * 4, cudaMemcpyDeviceToHost));
      return (int64_t)(n * 4);
    }
    case SDORB_DBG_LEVEL_SELECTED: {
      int32_t n = 0;
      CU(cudaMemcpy(&n, h->d_sel_count + (size_t)frame * g.nlevels + level, 4, cudaMemcpyDeviceToHost));
      if (capacity < (size_t)n * 4) return SDORB_ERR_CAPACITY;
      if (n) CU(cudaMemcpy(dst, h->d_sel + (size_t)frame * g.sel_total + L.sel_base, (size_t)n * 4, cudaMemcpyDeviceToHost));
      return (int64_t)n * 4;
    }
    default: return SDORB_ERR_BAD_ARG;
  }
}

}  // extern "C"
