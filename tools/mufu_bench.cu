// Micro-benchmark of the XU (MUFU) pipe on B200 -- written at the end of round 1; run in round 2 (profiles/r02_mufu_bench.txt): 16 elements / clk / SM for every MUFU form.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_bench tools/mufu_bench.cu && /tmp/mufu_bench
//
// Why: the fused GroupNorm+SiLU producers of conv_tc6 evaluate silu(z) = hz*tanh(hz) + hz with `tanh.approx.f16x2`, which
// ptxas splits into TWO `MUFU.TANH.F16` (cuobjdump: 88 MUFU for 44 packed pairs per thread and chunk), and ncu reports the XU
// pipe as the busiest unit of the fused kernel (profiles/r01_conv_tc6_ncu.txt).  If MUFU.TANH.F16 issues slower than the fp32
// MUFU.TANH / MUFU.EX2, a different formulation of the same activation frees the producers; if all run at 4 lanes per clock
// and SM sub-partition, the producers are at the XU floor (1 MUFU per activation x 1.33 halo) and only a smaller halo or
// a cheaper activation formulation can help.  Prints warp-instructions per clock per SM for each candidate.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int ITERS = 4096, UNROLL = 8;

template <int KIND>
__global__ void __launch_bounds__(1024) bench(float* out, float seed) {
  float x[UNROLL];
  uint32_t h[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    x[u] = seed + 0.001f * (threadIdx.x + u);
    __half2 v = __floats2half2_rn(x[u], x[u] * 0.5f);
    h[u] = *reinterpret_cast<uint32_t*>(&v);
  }
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (KIND == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[u]));
      if (KIND == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[u]));
      if (KIND == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[u]));
      if (KIND == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[u]));
      if (KIND == 4) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[u]));
      if (KIND == 5) asm volatile("tanh.approx.f16 %0, %0;" : "+h"(*reinterpret_cast<uint16_t*>(&h[u])));
      if (KIND == 6) asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(h[u]));       // FMA-pipe reference
    }
  }
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) s += x[u] + (float)h[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
static void run(const char* name, int elems_per_instr) {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out;
  cudaMalloc(&out, (size_t)sms * 2 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<KIND><<<sms * 2, 1024>>>(out, 0.3f);                 // warm-up
  cudaEventRecord(e0);
  bench<KIND><<<sms * 2, 1024>>>(out, 0.3f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr = (double)sms * 2 * 32 * ITERS * UNROLL;           // 32 warps per block
  const double per_us = warp_instr / (ms * 1e3);
  printf("%-22s %8.3f ms  %8.1f warp-instr/us/SM  = %5.2f warp-instr/clk/SM at the max clock (%d MHz), %5.1f elements/clk/SM\n", name, ms,
         per_us / sms, per_us / sms / (khz * 1e-3), khz / 1000, per_us / sms / (khz * 1e-3) * 32 * elems_per_instr);
  cudaFree(out);
}

int main() {
  run<0>("tanh.approx.f32", 1);
  run<1>("ex2.approx.ftz.f32", 1);
  run<2>("rcp.approx.ftz.f32", 1);
  run<3>("tanh.approx.f16x2", 2);
  run<4>("ex2.approx.f16x2", 2);
  run<5>("tanh.approx.f16", 1);
  run<6>("fma.rn.f16x2 (FMA pipe)", 2);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
