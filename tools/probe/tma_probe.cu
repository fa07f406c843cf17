// tma_probe.cu -- does a TMA 2-D tile load (cp.async.bulk.tensor.2d) beat the hand-written staging loop of the scan kernels?
//
// The north star asks for "TMA 2D tiles where they win".  The FAST kernel (kernels_fast.cu) stages a 136 x 68-byte tile of a
// pyramid level into shared memory with one LDG.32 + STS.32 per lane and row; the blur and pyramid kernels keep their rows in
// registers and use no shared memory at all.  This probe runs ONLY the staging step of a FAST-shaped tile walk over 512 frames of
// 640 x 480 (157 MB, larger than the 126 MB L2) in both forms, consumes the tile from shared memory with one XOR per staged word
// (so the loads cannot be elided), and reports the time and the staged GB/s of each:
//   manual : 128 threads, every lane loads aligned words of its rows, __syncthreads, consume
//   tma    : thread 0 arms an mbarrier and issues one cp.async.bulk.tensor.2d, everybody waits on the barrier, consume.
//            A TMA box must START ON A 16-BYTE BOUNDARY in global memory (an unaligned x coordinate raises "illegal
//            instruction": first version of this probe).  FAST tiles start at column 12 + 120 * tx -- 4-byte aligned only --
//            so the box is 144 bytes wide from (x0 & ~15) and the tile is read at a word offset: 12.5 % more bytes staged.
// What to compare the result with: fast_tiles_kernel needs ~1.7 ms for the same 512 frames (ncu, profiles/r2_ncu_full_summary.csv)
// with the integer ALU pipe 89 % busy.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

constexpr int W = 640, H = 480, TW = 128, TH = 68, OW = 120, OH = 60;
constexpr int TWB = 144;  // TMA box width: TW plus the up to 12 bytes in front of an unaligned tile start, rounded to 16
constexpr int TILES_X = (W - 38 + OW - 1) / OW, TILES_Y = (H - 38 + OH - 1) / OH;

__global__ void __launch_bounds__(128, 8) stage_manual(const uint8_t* __restrict__ img, uint32_t* __restrict__ out) {
  __shared__ __align__(128) uint32_t tile[TH][TW / 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tx = blockIdx.x % TILES_X, ty = blockIdx.x / TILES_X, frame = blockIdx.y;
  const int x0 = 12 + tx * OW, y0 = 15 + ty * OH;
  const uint8_t* src = img + (size_t)frame * W * H;
  uint32_t v[TH / 4];
#pragma unroll
  for (int j = 0; j < TH / 4; ++j) {  // all loads of the thread in flight together, as in fast_tile()
    const int gy = y0 + warp + 4 * j, gx = x0 + 4 * lane;
    v[j] = (gy < H && gx + 4 <= W) ? *reinterpret_cast<const uint32_t*>(src + (size_t)gy * W + gx) : 0u;
  }
#pragma unroll
  for (int j = 0; j < TH / 4; ++j) tile[warp + 4 * j][lane] = v[j];
  __syncthreads();
  uint32_t acc = 0;
#pragma unroll 4
  for (int r = warp; r < TH; r += 4) acc ^= tile[r][(lane + r) & 31];
  out[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 128 + threadIdx.x] = acc;
}

// The TMA form, written with libcu++'s wrappers (cuda::barrier + cuda::device::experimental::cp_async_bulk_tensor_2d_global_to_shared,
// the form of the CUDA programming guide): SASS shows UTMALDG.2D + SYNCS (mbarrier).
namespace cde = cuda::device::experimental;
using block_barrier = cuda::barrier<cuda::thread_scope_block>;

__global__ void __launch_bounds__(128, 8) stage_tma(const __grid_constant__ CUtensorMap tmap, uint32_t* __restrict__ out) {
  __shared__ alignas(128) uint32_t tile[TH][TWB / 4];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ block_barrier bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tx = blockIdx.x % TILES_X, ty = blockIdx.x / TILES_X, frame = blockIdx.y;
  const int x0 = 12 + tx * OW, y0 = 15 + ty * OH + frame * H;  // frames are stacked rows of one 640 x (480 * N) tensor
  if (threadIdx.x == 0) {
    init(&bar, blockDim.x);
    cde::fence_proxy_async_shared_cta();  // the barrier's initialisation becomes visible to the async (TMA) proxy
  }
  __syncthreads();
  block_barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&tile[0][0], &tmap, x0 & ~15, y0, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(tile));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  const int wo = (x0 & 15) >> 2;  // the tile's first word inside the box
  uint32_t acc = 0;
#pragma unroll 4
  for (int r = warp; r < TH; r += 4) acc ^= tile[r][wo + ((lane + r) & 31)];
  out[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 128 + threadIdx.x] = acc;
}

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);          \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 512;
  uint8_t* img;
  uint32_t *out_a, *out_b;
  const size_t bytes = (size_t)N * W * H, nout = (size_t)N * TILES_X * TILES_Y * 128;
  CK(cudaMalloc(&img, bytes + 4096));
  CK(cudaMalloc(&out_a, nout * 4));
  CK(cudaMalloc(&out_b, nout * 4));
  uint8_t* h = (uint8_t*)malloc(bytes);
  uint32_t s = 12345;
  for (size_t i = 0; i < bytes; ++i) h[i] = (uint8_t)((s = s * 1664525u + 1013904223u) >> 24);
  CK(cudaMemcpy(img, h, bytes, cudaMemcpyHostToDevice));
  // tensor map: u8, 2-D {W, H * N}, box {TW, TH}; out-of-bounds bytes read as zero, like the staging loop
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) {
    printf("cuTensorMapEncodeTiled not available\n");
    return 1;
  }
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H * N}, strides[1] = {(cuuint64_t)W};
  const cuuint32_t box[2] = {TWB, TH}, estr[2] = {1, 1};
  const CUresult r = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return 1;
  }
  const dim3 grid(TILES_X * TILES_Y, N);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  stage_manual<<<grid, 128>>>(img, out_a);
  CK(cudaDeviceSynchronize());
  printf("manual staging kernel ran\n");
  stage_tma<<<grid, 128>>>(tmap, out_b);
  CK(cudaDeviceSynchronize());
  printf("TMA staging kernel ran\n");
  float ms_a = 1e9f, ms_b = 1e9f;
  for (int rep = 0; rep < 6; ++rep) {
    float ms;
    CK(cudaEventRecord(e0));
    stage_manual<<<grid, 128>>>(img, out_a);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < ms_a) ms_a = ms;
    CK(cudaEventRecord(e0));
    stage_tma<<<grid, 128>>>(tmap, out_b);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < ms_b) ms_b = ms;
  }
  CK(cudaGetLastError());
  // both forms must have staged the same bytes (the last tile column / row reads zeros beyond the image in both... except that
  // a frame's bottom tiles see the next frame's rows through the stacked tensor; compare the interior tiles only)
  uint32_t *ha = (uint32_t*)malloc(nout * 4), *hb = (uint32_t*)malloc(nout * 4);
  CK(cudaMemcpy(ha, out_a, nout * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb, out_b, nout * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0, cmp = 0;
  for (int f = 0; f < N; ++f)
    for (int t = 0; t < TILES_X * (TILES_Y - 1); ++t)
      if (t % TILES_X != TILES_X - 1)
        for (int k = 0; k < 128; ++k, ++cmp) bad += ha[((size_t)f * TILES_X * TILES_Y + t) * 128 + k] != hb[((size_t)f * TILES_X * TILES_Y + t) * 128 + k];
  const double staged = (double)N * TILES_X * TILES_Y * TW * TH;
  printf("frames %d (%.0f MB of pixels, %.0f MB staged in %d x %d tiles of %d x %d bytes)\n", N, bytes / 1e6, staged / 1e6, TILES_X, TILES_Y, TW, TH);
  printf("manual LDG.32 + STS.32 staging : %.3f ms  %.0f GB/s staged\n", ms_a, staged / ms_a / 1e6);
  printf("TMA cp.async.bulk.tensor.2d    : %.3f ms  %.0f GB/s staged\n", ms_b, staged / ms_b / 1e6);
  printf("interior tiles compared: %zu words, mismatches %zu\n", cmp, bad);
  return bad ? 2 : 0;
}
