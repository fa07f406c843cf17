// kernels_probe.cu -- micro-benchmarks of the two execution pipes the hot kernels of this library are bound by, so that bench.py
// quotes MEASURED peaks on the device it runs on ("popc peak to be measured"):
//   pipe 0  POPC (the pipe match_kernel saturates: DescriptorDistance, /root/reference/src/ORBmatcher.cc:1459-1473)
//   pipe 1  the integer ALU pipe with VIMNMX3.U16x2 (the arc min / max network of fast_tiles_kernel)
//   pipe 2  PRMT (the ring / tap windows of the FAST, pyramid and blur kernels; same ALU pipe)
//   pipe 3  HFMA2.RELU, pipe 4 HADD2 (the FMA pipe, idle in the integer kernels), pipe 7 IMAD.HI
//   pipe 5  eight VIMNMX3 + eight HFMA2.RELU per iteration, pipe 6 the same plus four LDS: can work be moved across pipes?
// Eight independent dependency chains per thread, 8 x 256 threads per SM: enough ILP and warps to saturate a pipe.  The rate is
// reported per second (CUDA events around the launch) and per SM clock, with the SM clock itself measured inside the kernel
// (clock64 ticks per %globaltimer nanosecond, median over the CTAs) instead of taken from nvidia-smi.
#include "kernels.cuh"

namespace sdorb {

template <int PIPE>
__global__ void __launch_bounds__(256) pipe_probe_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, int iters,
                                                         long long* __restrict__ cycles) {
  __shared__ uint32_t s_tile[1024];
  uint32_t v[8], u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = in[(threadIdx.x + 8 * i) & 1023];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = v[i] * 0x9E3779B1u;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_tile[i] = in[i] & 0xFFFEFFFEu;
  __syncthreads();
  unsigned long long g0, g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PIPE == 0) asm volatile("popc.b32 %0, %0;" : "+r"(v[i]));
      if (PIPE == 1) v[i] = __vimax3_u16x2(v[i], v[(i + 1) & 7], v[(i + 2) & 7]);
      if (PIPE == 2) v[i] = __byte_perm(v[i], v[(i + 1) & 7], 0x5140);
      if (PIPE == 3 || PIPE == 5 || PIPE == 6)  // HFMA2.RELU on the FMA pipe
        asm volatile("fma.rn.relu.f16x2 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(v[(i + 1) & 7]), "r"(v[(i + 2) & 7]));
      if (PIPE == 4) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(v[i]) : "r"(v[(i + 1) & 7]));
      if (PIPE == 7) v[i] = __umulhi(v[i], v[(i + 1) & 7]);
    }
    if (PIPE == 5 || PIPE == 6) {  // the same number of ALU-pipe instructions on eight more chains: do the two pipes co-issue?
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = __vimax3_u16x2(u[i], u[(i + 1) & 7], u[(i + 2) & 7]);
    }
    if (PIPE == 6) {  // ... and a shared-memory load per two of them (the mix of a kernel that reads its operands from a tile)
#pragma unroll
      for (int i = 0; i < 4; ++i) u[i] ^= s_tile[(threadIdx.x + 33 * i + (u[i + 4] & 1)) & 1023];
    }
  }
  const long long t1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] + u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) {  // ticks and nanoseconds of this CTA's loop
    cycles[2 * blockIdx.x] = t1 - t0;
    cycles[2 * blockIdx.x + 1] = (long long)(g1 - g0);
  }
}

// Runs the probe on stream s; returns 0 or a cudaError_t.  rate_per_s: warp-instructions per second over the whole GPU;
// per_clk_sm: warp-instructions per SM clock per SM (median CTA duration in clock64 ticks as the denominator).
int run_pipe_probe(int pipe, cudaStream_t s, double* rate_per_s, double* per_clk_sm) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ctas = sms * 8, iters = 8192;
  uint32_t *out = nullptr, *in = nullptr;
  long long* cyc = nullptr;
  cudaError_t e = cudaMalloc(&out, sizeof(uint32_t) * (size_t)ctas * 256);
  if (e == cudaSuccess) e = cudaMalloc(&in, sizeof(uint32_t) * 1024);
  if (e == cudaSuccess) e = cudaMalloc(&cyc, sizeof(long long) * 2 * (size_t)ctas);
  if (e == cudaSuccess) e = cudaMemsetAsync(in, 0x35, sizeof(uint32_t) * 1024, s);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  float ms = 0.f;
  if (e == cudaSuccess) {
    auto launch = [&](int n) {
      if (pipe == 0) pipe_probe_kernel<0><<<ctas, 256, 0, s>>>(out, in, n, cyc);
      else if (pipe == 1) pipe_probe_kernel<1><<<ctas, 256, 0, s>>>(out, in, n, cyc);
      else if (pipe == 2) pipe_probe_kernel<2><<<ctas, 256, 0, s>>>(out, in, n, cyc);
      else if (pipe == 3) pipe_probe_kernel<3><<<ctas, 256, 0, s>>>(out, in, n, cyc);
      else if (pipe == 4) pipe_probe_kernel<4><<<ctas, 256, 0, s>>>(out, in, n, cyc);
      else if (pipe == 5) pipe_probe_kernel<5><<<ctas, 256, 0, s>>>(out, in, n, cyc);
      else if (pipe == 6) pipe_probe_kernel<6><<<ctas, 256, 0, s>>>(out, in, n, cyc);
      else pipe_probe_kernel<7><<<ctas, 256, 0, s>>>(out, in, n, cyc);
    };
    launch(64);  // warm-up
    cudaEventRecord(e0, s);
    launch(iters);
    cudaEventRecord(e1, s);
    e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
  }
  if (e == cudaSuccess) {
    std::vector<long long> h(2 * (size_t)ctas);
    e = cudaMemcpy(h.data(), cyc, sizeof(long long) * 2 * (size_t)ctas, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) {
      std::vector<double> ghz;
      for (int c = 0; c < ctas; ++c)
        if (h[2 * (size_t)c + 1] > 0) ghz.push_back((double)h[2 * (size_t)c] / (double)h[2 * (size_t)c + 1]);
      std::nth_element(ghz.begin(), ghz.begin() + ghz.size() / 2, ghz.end());
      const double clk_hz = ghz.empty() ? 0.0 : ghz[ghz.size() / 2] * 1e9;
      const int per_iter = pipe == 5 ? 16 : pipe == 6 ? 20 : 8;  // instructions per thread and iteration
      const double warp_instr = (double)ctas * 8 /*warps*/ * per_iter * (double)iters;
      const double rate = warp_instr / ((double)ms * 1e-3);
      if (rate_per_s) *rate_per_s = rate;
      if (per_clk_sm) *per_clk_sm = clk_hz > 0 ? rate / sms / clk_hz : 0.0;
    }
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(out);
  cudaFree(in);
  cudaFree(cyc);
  return (int)e;
}

}  // namespace sdorb
