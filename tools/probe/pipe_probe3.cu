// Does the FMA pipe take packed 16-bit max / min work beside the ALU pipe?  Lanes hold 0x6400 | byte: as u16 they order
// like the byte (VIMNMX.U16x2 works on them), as fp16 they are 1024 + byte, so  max(a,b) = b + relu(a - b)  and
// min(a,b) = a - relu(a - b)  are exact with HFMA2.RELU + HADD2 (FMA pipe).
#include <cuda_fp16.h>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned relu_sub(unsigned a, unsigned b) {  // relu(a - b) per fp16 lane
  unsigned d;
  asm("{ .reg .b32 one, nb; mov.b32 one, 0x3c003c00; xor.b32 nb, %2, 0x80008000; fma.rn.relu.f16x2 %0, %1, one, nb; }" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ unsigned relu_sub2(unsigned a, unsigned b) {  // relu(a - b): b * -1 + a
  unsigned d;
  asm("{ .reg .b32 m1; mov.b32 m1, 0xbc00bc00; fma.rn.relu.f16x2 %0, %2, m1, %1; }" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ unsigned hadd2u(unsigned a, unsigned b) {
  unsigned d;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ unsigned hsub2u(unsigned a, unsigned b) {
  unsigned d;
  asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ unsigned fmax2(unsigned a, unsigned b) { return hadd2u(b, relu_sub2(a, b)); }
__device__ __forceinline__ unsigned fmin2(unsigned a, unsigned b) { return hsub2u(a, relu_sub2(a, b)); }

template <int NV, int NF>
__global__ void k(unsigned* out, const unsigned* in, int iters) {
  unsigned v[8], h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = (in[threadIdx.x + i] & 0x00ff00ffu) | 0x64006400u;
    h[i] = (in[threadIdx.x + 8 + i] & 0x00ff00ffu) | 0x64006400u;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < NV) v[i] = (it & 1) ? __vimax3_u16x2(v[i], v[(i + 1) & 7], v[(i + 2) & 7]) : __vimin3_u16x2(v[i], v[(i + 3) & 7], v[(i + 5) & 7]);
      if (i < NF) h[i] = (it & 1) ? fmax2(h[i], h[(i + 1) & 7]) : fmin2(h[i], h[(i + 3) & 7]);
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] + h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// exactness: all byte pairs
__global__ void check(int* bad) {
  const unsigned a = threadIdx.x, b = blockIdx.x;
  const unsigned A = 0x64006400u | a | (b << 16), B = 0x64006400u | b | (a << 16);
  const unsigned mx = fmax2(A, B), mn = fmin2(A, B);
  if (mx != __vmaxu2(A, B) || mn != __vminu2(A, B)) atomicAdd(bad, 1);
}
template <int NV, int NF>
void run(const char* name) {
  unsigned *d, *in; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0x35, 4096);
  const int iters = 4096;
  k<NV, NF><<<148 * 8, 256>>>(d, in, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<NV, NF><<<148 * 8, 256>>>(d, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double clk = ms * 1e-3 * 1.965e9;  // per SM: 8 CTAs x 8 warps = 64 warps, 16 per SMSP
  printf("%-40s %8.3f ms  %6.2f clk per (iteration x warp) per SMSP\n", name, ms, clk / (16.0 * iters));
}
int main() {
  int* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
  check<<<256, 256>>>(bad);
  int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  printf("fp16 max/min vs u16 max/min on all byte pairs: %d mismatches\n", hb);
  run<8, 0>("8 VIMNMX3");
  run<0, 8>("8 fp16 min/max (16 FMA-pipe instr)");
  run<0, 4>("4 fp16 min/max (8 FMA-pipe instr)");
  run<8, 2>("8 VIMNMX3 + 2 fp16 (4 FMA)");
  run<8, 4>("8 VIMNMX3 + 4 fp16 (8 FMA)");
  run<8, 8>("8 VIMNMX3 + 8 fp16 (16 FMA)");
  run<4, 8>("4 VIMNMX3 + 8 fp16 (16 FMA)");
  return 0;
}
