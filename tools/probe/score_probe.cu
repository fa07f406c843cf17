// score_probe.cu -- can the FAST score network use the (idle) FMA pipe next to the (saturated) integer ALU pipe?
//
// fast_tiles_kernel (sdslam_b200/csrc/kernels_fast.cu) spends 234 ALU-pipe instructions per 4-pixel word: 128 min / max of the arc
// network, 23 PRMT that cut the ring windows out of the staged byte rows, the rest epilogue and non-max suppression.  ncu shows the
// ALU pipe 89 % busy and the FMA pipe idle (profiles/r2_ncu_full_summary.csv).  This probe runs ONLY the scoring phase of one
// 128 x 62-pixel tile, over and over, in two forms, checks that both give the same score bytes, and times them:
//   A  the shipped form: byte tile, 7 x 3 words per thread, PRMT windows, VIMNMX(3).U16x2 network on (pixel << 8 | junk) lanes;
//   B  an fp16-lane tile: every pixel is staged ONCE as the half-precision number 1024 + pixel (bits 0x6400 | pixel: a byte
//      permute), two copies (pairs starting at even / odd columns), so every ring sample of a pixel pair is one aligned LDS.32 and
//      no PRMT.  Positive halves order like their bit patterns, so the 3-input VIMNMX3.U16x2 instructions of the second network
//      stage work on them unchanged, and the FIRST stage -- 16 (min, max) pairs of two ring samples each -- can run on the FMA
//      pipe, exactly (all values are integers below 2048):  d = relu(a - b) [HFMA2.RELU], max = b + d, min = a - d [HADD2].
//      HL / PP select how the (lo2, hi2) and the (pmin, pmax) pairs are computed: 0 = two VIMNMX, 1 = one VIMNMX + two HADD2,
//      2 = HFMA2.RELU + two HADD2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o bin/score_probe score_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int OH = 60, SROWS = OH + 2, PROWS = SROWS + 6, SWORDS = 32, PWORDS = SWORDS + 2, TWORDS = SWORDS + 2, NT = 128;
constexpr int HW = 68;  // words per row of one fp16-lane copy: 136 staged pixels / 2

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }

// ---------------------------------------------------------------- form A (as shipped)
template <int P, int DX>
__device__ __forceinline__ uint32_t ring_at(const uint32_t w0, const uint32_t w1, const uint32_t w2) {
  constexpr int s = 3 + P + DX;
  if (s == 0) return w0;
  if (s == 4) return w1;
  if (s == 8) return w2;
  if (s < 4) {
    constexpr uint32_t sel = s | ((s + 1) << 4) | ((s + 2) << 8) | ((s + 3) << 12);
    return prmt(w0, w1, sel);
  } else {
    constexpr uint32_t t = s - 4, sel = t | ((t + 1) << 4) | ((t + 2) << 8) | ((t + 3) << 12);
    return prmt(w1, w2, sel);
  }
}

// second stage of the network, shared by both forms: X = min over arcs of the arc maximum, Y = max over arcs of the arc minimum
__device__ __forceinline__ void stage2(const uint32_t (&lo2)[8], const uint32_t (&hi2)[8], const uint32_t (&pmin)[8],
                                       const uint32_t (&pmax)[8], uint32_t& X, uint32_t& Y) {
  uint32_t wmax[8], wmin[8];
#pragma unroll
  for (int q = 0; q < 8; q += 2) {
    const uint32_t chi = __vimax3_u16x2(hi2[(q + 1) & 7], hi2[(q + 2) & 7], hi2[(q + 3) & 7]);
    const uint32_t clo = __vimin3_u16x2(lo2[(q + 1) & 7], lo2[(q + 2) & 7], lo2[(q + 3) & 7]);
    wmax[q] = __vimax3_u16x2(chi, hi2[q], pmin[q]);
    wmin[q] = __vimin3_u16x2(clo, lo2[q], pmax[q]);
    wmax[q + 1] = __vimax3_u16x2(chi, hi2[(q + 4) & 7], pmin[q + 1]);
    wmin[q + 1] = __vimin3_u16x2(clo, lo2[(q + 4) & 7], pmax[q + 1]);
  }
  X = __vimin3_u16x2(wmax[0], wmax[1], wmax[2]), Y = __vimax3_u16x2(wmin[0], wmin[1], wmin[2]);
  X = __vimin3_u16x2(X, wmax[3], wmax[4]);
  Y = __vimax3_u16x2(Y, wmin[3], wmin[4]);
  X = __vimin3_u16x2(X, wmax[5], wmax[6]);
  Y = __vimax3_u16x2(Y, wmin[5], wmin[6]);
  X = __vminu2(X, wmax[7]);
  Y = __vmaxu2(Y, wmin[7]);
}

template <int P>
__device__ __forceinline__ uint32_t score_pair_a(const uint32_t (&W)[7][3], const uint32_t th2) {
  uint32_t r[16];
  r[0] = ring_at<P, 0>(W[6][0], W[6][1], W[6][2]);
  r[1] = ring_at<P, 1>(W[6][0], W[6][1], W[6][2]);
  r[2] = ring_at<P, 2>(W[5][0], W[5][1], W[5][2]);
  r[3] = ring_at<P, 3>(W[4][0], W[4][1], W[4][2]);
  r[4] = ring_at<P, 3>(W[3][0], W[3][1], W[3][2]);
  r[5] = ring_at<P, 3>(W[2][0], W[2][1], W[2][2]);
  r[6] = ring_at<P, 2>(W[1][0], W[1][1], W[1][2]);
  r[7] = ring_at<P, 1>(W[0][0], W[0][1], W[0][2]);
  r[8] = ring_at<P, 0>(W[0][0], W[0][1], W[0][2]);
  r[9] = ring_at<P, -1>(W[0][0], W[0][1], W[0][2]);
  r[10] = ring_at<P, -2>(W[1][0], W[1][1], W[1][2]);
  r[11] = ring_at<P, -3>(W[2][0], W[2][1], W[2][2]);
  r[12] = ring_at<P, -3>(W[3][0], W[3][1], W[3][2]);
  r[13] = ring_at<P, -3>(W[4][0], W[4][1], W[4][2]);
  r[14] = ring_at<P, -2>(W[5][0], W[5][1], W[5][2]);
  r[15] = ring_at<P, -1>(W[6][0], W[6][1], W[6][2]);
  uint32_t lo2[8], hi2[8], pmax[8], pmin[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int j = 2 * q + 1;
    lo2[q] = __vminu2(r[j], r[(j + 1) & 15]);
    hi2[q] = __vmaxu2(r[j], r[(j + 1) & 15]);
    pmin[q] = __vminu2(r[j - 1], r[(j + 8) & 15]);
    pmax[q] = __vmaxu2(r[j - 1], r[(j + 8) & 15]);
  }
  uint32_t X, Y;
  stage2(lo2, hi2, pmin, pmax, X, Y);
  const uint32_t Xc = prmt(X, 0u, 0x4341), Yc = prmt(Y, 0u, 0x4341);
  const uint32_t Vc = prmt(W[3][1], 0u, P == 0 ? 0x4240 : 0x4341);
  const uint32_t A = Vc + 0x01000100u - Xc, B = Yc + 0x01000100u - Vc;
  return __vimax3_u16x2(A, B, th2) - th2;
}

__global__ void __launch_bounds__(NT, 1024 / NT) score_a(const uint8_t* __restrict__ img, uint32_t* __restrict__ out, int iters, int th) {
  __shared__ __align__(16) uint32_t s_pix[PROWS][PWORDS];
  __shared__ __align__(16) uint32_t s_t[SROWS][TWORDS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(img) + (size_t)(blockIdx.x & 63) * PROWS * PWORDS;
  for (int i = tid; i < PROWS * PWORDS; i += NT) s_pix[i / PWORDS][i % PWORDS] = src[i];
  for (int i = tid; i < SROWS * TWORDS; i += NT) s_t[i / TWORDS][i % TWORDS] = 0;
  __syncthreads();
  const uint32_t th2 = (uint32_t)(th + 256) * 0x00010001u;
  const int k = lane;
  for (int it = 0; it < iters; ++it) {
    for (int base = 0; base < SROWS; base += NT / 32) {
      const int sr = base + warp;
      if (sr < SROWS) {
        uint32_t W[7][3];
#pragma unroll
        for (int dy = 0; dy < 7; ++dy) {
          W[dy][0] = s_pix[sr + dy][k];
          W[dy][1] = s_pix[sr + dy][k + 1];
          W[dy][2] = s_pix[sr + dy][k + 2];
        }
        const uint32_t t0 = score_pair_a<0>(W, th2), t1 = score_pair_a<1>(W, th2);
        s_t[sr][k + 1] = prmt(t0, t1, 0x6240);
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < SROWS * TWORDS; i += NT) out[(size_t)blockIdx.x * SROWS * TWORDS + i] = s_t[i / TWORDS][i % TWORDS];
}

// ---------------------------------------------------------------- form B (fp16-lane tile, first stage on the FMA pipe)
__device__ __forceinline__ uint32_t h_add(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h_sub(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h_relu_sub(uint32_t a, uint32_t b) {  // max(a - b, 0) = relu(b * -1 + a)
  uint32_t d;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(0xBC00BC00u), "r"(a));
  return d;
}
template <int MODE>
__device__ __forceinline__ void minmax(const uint32_t a, const uint32_t b, uint32_t& lo, uint32_t& hi) {
  if (MODE == 0) {
    lo = __vminu2(a, b);
    hi = __vmaxu2(a, b);
  } else if (MODE == 1) {
    lo = __vminu2(a, b);
    hi = h_add(a, h_sub(b, lo));  // a + b - min, each step exact
  } else {
    const uint32_t d = h_relu_sub(a, b);
    hi = h_add(b, d);
    lo = h_sub(a, d);
  }
}

// word offset and copy of ring column DX for pair q: staged pixel 2q + 4 + DX
__host__ __device__ constexpr int col_copy(int dx) { return dx & 1; }
__host__ __device__ constexpr int col_word(int dx) { return 2 + (dx - (dx & 1)) / 2; }

template <int HL, int PP>
__device__ __forceinline__ uint32_t score_pair_b(const uint32_t* __restrict__ base /* &s_h[0][sr][q] */, const uint32_t th_h, const uint32_t th_back) {
  constexpr int DXS[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  constexpr int DYS[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = base[col_copy(DXS[i]) * PROWS * HW + (DYS[i] + 3) * HW + col_word(DXS[i])];
  const uint32_t v = base[3 * HW + col_word(0)];
  uint32_t lo2[8], hi2[8], pmax[8], pmin[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int j = 2 * q + 1;
    minmax<HL>(r[j], r[(j + 1) & 15], lo2[q], hi2[q]);
    minmax<PP>(r[j - 1], r[(j + 8) & 15], pmin[q], pmax[q]);
  }
  uint32_t X, Y;
  stage2(lo2, hi2, pmin, pmax, X, Y);
  const uint32_t A = h_relu_sub(v, X), B = h_relu_sub(Y, v);  // max(v - X, 0), max(Y - v, 0): non-negative halves
  const uint32_t m = __vimax3_u16x2(A, B, th_h);
  return h_add(m, th_back);  // m - th + 1024: bits 0x6400 | t
}

template <int HL, int PP>
__global__ void __launch_bounds__(NT, 5) score_b(const uint8_t* __restrict__ img, uint32_t* __restrict__ out, int iters, int th) {
  __shared__ __align__(16) uint32_t s_h[2][PROWS][HW];
  __shared__ __align__(16) uint32_t s_t[SROWS][TWORDS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(img) + (size_t)(blockIdx.x & 63) * PROWS * PWORDS;
  for (int i = tid; i < PROWS * PWORDS; i += NT) {  // one byte word -> two words of each copy
    const int row = i / PWORDS, kk = i % PWORDS;
    const uint32_t w = src[i], nx = kk + 1 < PWORDS ? src[i + 1] : 0u, c = 0x64646464u;
    s_h[0][row][2 * kk] = prmt(w, c, 0x4140);
    s_h[0][row][2 * kk + 1] = prmt(w, c, 0x4342);
    s_h[1][row][2 * kk] = prmt(w, c, 0x4241);
    s_h[1][row][2 * kk + 1] = prmt(prmt(w, nx, 0x0043), c, 0x4140);
  }
  for (int i = tid; i < SROWS * TWORDS; i += NT) s_t[i / TWORDS][i % TWORDS] = 0;
  __syncthreads();
  __half2 thh = __floats2half2_rn((float)th, (float)th), tbk = __floats2half2_rn(1024.f - th, 1024.f - th);
  const uint32_t th_h = *reinterpret_cast<uint32_t*>(&thh), th_back = *reinterpret_cast<uint32_t*>(&tbk);
  uint16_t* s_t16 = reinterpret_cast<uint16_t*>(&s_t[0][0]);
  for (int it = 0; it < iters; ++it) {
    for (int base = 0; base < SROWS; base += NT / 32) {
      const int sr = base + warp;
      if (sr < SROWS) {
        const uint32_t* b0 = &s_h[0][sr][lane];
        const uint32_t t0 = score_pair_b<HL, PP>(b0, th_h, th_back), t1 = score_pair_b<HL, PP>(b0 + 32, th_h, th_back);
        s_t16[sr * TWORDS * 2 + 2 + lane] = (uint16_t)prmt(t0, 0u, 0x4420);       // pixels 2 lane, 2 lane + 1
        s_t16[sr * TWORDS * 2 + 2 + 32 + lane] = (uint16_t)prmt(t1, 0u, 0x4420);  // pixels 64 + 2 lane, ...
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < SROWS * TWORDS; i += NT) out[(size_t)blockIdx.x * SROWS * TWORDS + i] = s_t[i / TWORDS][i % TWORDS];
}

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

__global__ void clock_kernel(long long* out) {
  unsigned long long g0, g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  const long long t0 = clock64();
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  } while (g1 - g0 < 2000000ull);
  out[0] = clock64() - t0;
  out[1] = (long long)(g1 - g0);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 200, th = 20;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int ctas = sms * 8 * 2;
  const size_t tile_words = (size_t)PROWS * PWORDS, nin = 64 * tile_words, nout = (size_t)ctas * SROWS * TWORDS;
  std::vector<uint32_t> h(nin);
  uint32_t s = 777;
  for (size_t i = 0; i < nin; ++i) {  // smooth + noise: neighbouring pixels correlate, as in an image
    uint32_t w = 0;
    for (int b = 0; b < 4; ++b) {
      s = s * 1664525u + 1013904223u;
      const int base = 128 + (int)(100 * __builtin_sin((double)(i % tile_words) * 0.07)), noise = (int)((s >> 24) % 61) - 30;
      int v = base + noise;
      v = v < 0 ? 0 : v > 255 ? 255 : v;
      w |= (uint32_t)v << (8 * b);
    }
    h[i] = w;
  }
  uint8_t* img;
  uint32_t *out_a, *out_b;
  long long* clk;
  CK(cudaMalloc(&img, nin * 4 + 64));
  CK(cudaMalloc(&out_a, nout * 4));
  CK(cudaMalloc(&out_b, nout * 4));
  CK(cudaMalloc(&clk, 16));
  CK(cudaMemcpy(img, h.data(), nin * 4, cudaMemcpyHostToDevice));
  clock_kernel<<<1, 1>>>(clk);
  long long hc[2];
  CK(cudaMemcpy(hc, clk, 16, cudaMemcpyDeviceToHost));
  const double ghz = (double)hc[0] / (double)hc[1];
  printf("SM clock %.3f GHz, %d SMs, %d CTAs x %d iterations of a %d-row tile\n", ghz, sms, ctas, iters, SROWS);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<uint32_t> ra(nout), rb(nout);
  auto report = [&](const char* name, float ms, bool check) {
    const double warp_rows = (double)ctas * iters * SROWS;
    const double cyc = (double)ms * 1e-3 * ghz * 1e9 * sms * 4 / warp_rows;
    size_t bad = 0, nz = 0;
    if (check)
      for (size_t i = 0; i < nout; ++i) bad += ra[i] != rb[i], nz += ra[i] != 0;
    printf("%-44s %8.3f ms  %6.1f SMSP-cycles per warp-row (32 words)", name, ms, cyc);
    if (check) printf("  mismatching words vs A: %zu (non-zero words %zu of %zu)", bad, nz, nout);
    printf("\n");
  };
#define RUN(NAME, KERNEL, OUT, HOSTV, CHECK)                          \
  do {                                                                \
    float best = 1e9f;                                                \
    for (int rep = 0; rep < 4; ++rep) {                               \
      float ms;                                                       \
      CK(cudaEventRecord(e0));                                        \
      KERNEL<<<ctas, NT>>>(img, OUT, iters, th);                      \
      CK(cudaEventRecord(e1));                                        \
      CK(cudaEventSynchronize(e1));                                   \
      CK(cudaGetLastError());                                         \
      CK(cudaEventElapsedTime(&ms, e0, e1));                          \
      if (rep && ms < best) best = ms;                                \
    }                                                                 \
    CK(cudaMemcpy(HOSTV.data(), OUT, nout * 4, cudaMemcpyDeviceToHost)); \
    report(NAME, best, CHECK);                                        \
  } while (0)
  RUN("A  byte tile, PRMT windows, all VIMNMX", score_a, out_a, ra, false);
  RUN("B00 fp16 tile, all VIMNMX", (score_b<0, 0>), out_b, rb, true);
  RUN("B10 fp16 tile, lo2/hi2: 1 VIMNMX + 2 HADD2", (score_b<1, 0>), out_b, rb, true);
  RUN("B11 fp16 tile, both pairs: 1 VIMNMX + 2 HADD2", (score_b<1, 1>), out_b, rb, true);
  RUN("B20 fp16 tile, lo2/hi2: HFMA2.RELU + 2 HADD2", (score_b<2, 0>), out_b, rb, true);
  RUN("B21 fp16 tile, HFMA2 form / 1 VIMNMX form", (score_b<2, 1>), out_b, rb, true);
  RUN("B22 fp16 tile, both pairs: HFMA2.RELU + 2 HADD2", (score_b<2, 2>), out_b, rb, true);
  RUN("B12 fp16 tile, 1 VIMNMX form / HFMA2 form", (score_b<1, 2>), out_b, rb, true);
  return 0;
}
