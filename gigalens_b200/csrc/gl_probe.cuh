// gl_probe.cuh -- measured FP32 FMA peak of the device (the roofline denominator of the ray-tracing kernels).
//
// Independent register chains of fused multiply-adds, no memory traffic: every SM runs 8 warps x 4 CTAs of
// GLPB_CHAINS independent accumulators, so the FMA pipe is the only limiter.  Two variants are timed with CUDA
// events, scalar FFMA and packed FFMA2 (the instruction the hot kernels issue), and the caller reports the larger
// one as the peak.  SURVEY.md section 8d: "measure an FP32 FMA peak on the box and use that".
#pragma once
#include <cuda_runtime.h>

#define GLPB_THREADS 256
#define GLPB_CHAINS 8
#define GLPB_ITERS 2048

template <bool PACKED>
__global__ void __launch_bounds__(GLPB_THREADS) k_fp32_probe(float seed, float* __restrict__ sink) {
  const float a0 = 1.0f + 1e-7f * seed, b0 = 1e-7f * (float)threadIdx.x;
  float r = 0.f;
  if constexpr (PACKED) {
    float2 acc[GLPB_CHAINS];
#pragma unroll
    for (int k = 0; k < GLPB_CHAINS; ++k) acc[k] = make_float2(b0 + k, b0 - k);
    const float2 a = make_float2(a0, a0 - 1e-7f), b = make_float2(b0, -b0);
    for (int it = 0; it < GLPB_ITERS; ++it) {
#pragma unroll
      for (int k = 0; k < GLPB_CHAINS; ++k) acc[k] = __ffma2_rn(acc[k], a, b);
    }
#pragma unroll
    for (int k = 0; k < GLPB_CHAINS; ++k) r += acc[k].x + acc[k].y;
  } else {
    float acc[2 * GLPB_CHAINS];
#pragma unroll
    for (int k = 0; k < 2 * GLPB_CHAINS; ++k) acc[k] = b0 + k;
    for (int it = 0; it < GLPB_ITERS; ++it) {
#pragma unroll
      for (int k = 0; k < 2 * GLPB_CHAINS; ++k) acc[k] = fmaf(acc[k], a0, b0);
    }
#pragma unroll
    for (int k = 0; k < 2 * GLPB_CHAINS; ++k) r += acc[k];
  }
  if (r == 123.456f) sink[0] = r;   // never true: keeps the chains alive without a store on the timed path
}

// Best-of-`reps` TFLOP/s (FMA = 2 flops) of both variants on the current device.
static inline cudaError_t gl_fp32_probe_run(int sm_count, int reps, float* tflops_ffma, float* tflops_ffma2, float* sink) {
  cudaEvent_t e0, e1;
  cudaError_t e;
  if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
  if ((e = cudaEventCreate(&e1)) != cudaSuccess) return e;
  const int grid = sm_count * 8 * 16;   // 16 waves of 8 CTAs per SM
  const double flops = 2.0 * 2.0 * GLPB_CHAINS * (double)GLPB_ITERS * GLPB_THREADS * grid;   // both variants: 16 fp32 FMA / thread / iteration
  float best[2] = {0.f, 0.f};
  for (int v = 0; v < 2; ++v)
    for (int r = 0; r < reps + 1; ++r) {
      cudaEventRecord(e0, 0);
      if (v) k_fp32_probe<true><<<grid, GLPB_THREADS>>>((float)r, sink);
      else k_fp32_probe<false><<<grid, GLPB_THREADS>>>((float)r, sink);
      cudaEventRecord(e1, 0);
      if ((e = cudaEventSynchronize(e1)) != cudaSuccess) return e;
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      const float tf = (float)(flops / (ms * 1e-3) / 1e12);
      if (r > 0 && tf > best[v]) best[v] = tf;   // first launch of each variant is a warm-up
    }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops_ffma = best[0];
  *tflops_ffma2 = best[1];
  return cudaGetLastError();
}
