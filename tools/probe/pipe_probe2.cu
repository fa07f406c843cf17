// Cleaner probe: independent chains per instruction type, operands change every iteration.
#include <cuda_fp16.h>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned hmax2u(unsigned a, unsigned b) { __half2 x = *(__half2*)&a, y = *(__half2*)&b; x = __hmax2(x, y); return *(unsigned*)&x; }
__device__ __forceinline__ unsigned hmin2u(unsigned a, unsigned b) { __half2 x = *(__half2*)&a, y = *(__half2*)&b; x = __hmin2(x, y); return *(unsigned*)&x; }
template <int NV, int NH, int NP, int NI>
__global__ void k(unsigned* out, const unsigned* in, int iters) {
  unsigned v[8], h[8], p[8], m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = in[threadIdx.x + i]; h[i] = (in[threadIdx.x + 8 + i] & 0x03ff03ffu) | 0x04000400u; p[i] = in[threadIdx.x + 16 + i]; m[i] = in[threadIdx.x + 24 + i]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < NV) v[i] = __vimax3_u16x2(v[i], v[(i + 1) & 7], v[(i + 2) & 7]) ;
      if (i < NH) h[i] = (it & 1) ? hmax2u(h[i], h[(i + 1) & 7]) : hmin2u(h[i], h[(i + 3) & 7]);
      if (i < NP) p[i] = __byte_perm(p[i], p[(i + 1) & 7], 0x5140);
      if (i < NI) m[i] = m[i] * 5 + m[(i + 1) & 7];
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] + h[i] + p[i] + m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NV, int NH, int NP, int NI>
void run(const char* name) {
  unsigned *d, *in; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0x35, 4096);
  const int iters = 4096;
  k<NV, NH, NP, NI><<<148 * 8, 256>>>(d, in, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<NV, NH, NP, NI><<<148 * 8, 256>>>(d, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr = 148.0 * 8 * 8 * iters * (NV + NH + NP + NI);
  printf("%-36s %8.3f ms  %6.2f warp-instr/clk/SM\n", name, ms, warp_instr / (ms * 1e-3) / 148 / 1.92e9);
}
int main() {
  run<8, 0, 0, 0>("8 VIMNMX3");
  run<0, 8, 0, 0>("8 HMNMX2");
  run<0, 0, 8, 0>("8 PRMT");
  run<0, 0, 0, 8>("8 IMAD");
  run<8, 8, 0, 0>("8 VIMNMX3 + 8 HMNMX2");
  run<4, 8, 0, 0>("4 VIMNMX3 + 8 HMNMX2");
  run<8, 0, 8, 0>("8 VIMNMX3 + 8 PRMT");
  run<8, 0, 0, 8>("8 VIMNMX3 + 8 IMAD");
  run<0, 8, 0, 8>("8 HMNMX2 + 8 IMAD");
  run<0, 8, 8, 0>("8 HMNMX2 + 8 PRMT");
  run<8, 8, 0, 8>("8 VIMNMX3 + 8 HMNMX2 + 8 IMAD");
  return 0;
}
