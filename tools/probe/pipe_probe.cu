// Which pipe runs packed 16-bit min/max on sm_100a?  Times N dependent-free chains of VIMNMX.U16x2, VIMNMX3.U16x2, HMNMX2, PRMT, IMAD
// alone and interleaved (development probe, not part of the library).
#include <cuda_fp16.h>
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(unsigned* out, unsigned seed, int iters) {
  unsigned a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + 1) + i * 0x01010101u; b[i] = a[i] ^ 0x00110011u; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = __vmaxu2(a[i], b[i]) + 1;                         // VIMNMX.U16x2 (+IADD)
      if (MODE == 1) a[i] = __vimax3_u16x2(a[i], b[i], a[(i + 1) & 7]);       // VIMNMX3
      if (MODE == 2) { __half2 x = *(__half2*)&a[i], y = *(__half2*)&b[i]; x = __hmax2(x, y); a[i] = *(unsigned*)&x; a[i] ^= it; }  // HMNMX2 (+LOP)
      if (MODE == 3) { a[i] = __vimax3_u16x2(a[i], b[i], a[(i + 1) & 7]); __half2 x = *(__half2*)&b[i], y = *(__half2*)&a[(i + 3) & 7]; x = __hmax2(x, y); b[i] = *(unsigned*)&x; }  // both
      if (MODE == 4) { __half2 x = *(__half2*)&b[i], y = *(__half2*)&a[(i + 3) & 7]; x = __hmax2(x, y); b[i] = *(unsigned*)&x; }  // HMNMX2 only, chained like mode 3
      if (MODE == 5) a[i] = __byte_perm(a[i], b[i], 0x5140 + (it & 1));        // PRMT
      if (MODE == 6) a[i] = a[i] * 3 + b[i];                                   // IMAD
      if (MODE == 7) { a[i] = __vimax3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = b[i] * 3 + a[(i + 3) & 7]; }  // VIMNMX3 + IMAD
      if (MODE == 8) { a[i] = __vimax3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = __byte_perm(b[i], a[(i + 3) & 7], 0x5140); }  // VIMNMX3 + PRMT
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int ops_per_iter) {
  unsigned* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  const int iters = 4096;
  k<MODE><<<148 * 8, 256>>>(d, 7, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, 7, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr = 148.0 * 8 * 8 * iters * 8 * ops_per_iter;  // blocks * warps * iters * unroll * ops
  printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (at 1.92 GHz)\n", name, ms, warp_instr / (ms * 1e-3) / 148 / 1.92e9);
  cudaFree(d);
}
int main() {
  run<0>("VIMNMX.U16x2 + IADD", 2);
  run<1>("VIMNMX3.U16x2", 1);
  run<2>("HMNMX2 + LOP3", 2);
  run<4>("HMNMX2", 1);
  run<3>("VIMNMX3 + HMNMX2", 2);
  run<5>("PRMT", 1);
  run<6>("IMAD", 1);
  run<7>("VIMNMX3 + IMAD", 2);
  run<8>("VIMNMX3 + PRMT", 2);
  return 0;
}
