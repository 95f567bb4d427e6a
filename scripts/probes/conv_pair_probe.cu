// conv_pair_probe.cu -- steady-state throughput of the two register-tile correlation schemes of gl_conv.cuh on synthetic
// shared-memory tiles (no HBM traffic, no TMA): (a) row pairs with broadcast tap pairs (corr_rows2, the round-1 kernel),
// (b) column pairs with scalar taps (corr_rows_cp).  Same thread tile (6 x 4 outputs), same tile geometry as C2
// (A = 13 taps per phase, 4 phases, 160 threads per CTA, 4 CTAs per SM).  Prints ms and TFLOP/s of useful work for both.
#include <cstdio>
#include <vector>
#include "../../gigalens_b200/csrc/gl_conv.cuh"

#define A 13
#define NTHR 160
#define PITCH 76
#define ROWS_IN 72
#define TILE (ROWS_IN * PITCH)

template <int MODE>
__global__ void __launch_bounds__(NTHR) k_probe(const float* __restrict__ gin, const float* __restrict__ gw, float* __restrict__ out, int reps, unsigned zero) {
  extern __shared__ __align__(128) float sm[];
  float* s_in = sm;
  float* s_w = sm + TILE;
  constexpr int WT = MODE == 0 ? 4 * A * 2 * glc_ulen(A) : 4 * A * glc_wpitch(A);
  for (int i = threadIdx.x; i < TILE; i += NTHR) s_in[i] = gin[i];
  for (int i = threadIdx.x; i < WT; i += NTHR) s_w[i] = gw[i];
  __syncthreads();
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  const int origin = ty * GLC_RY * PITCH + tx * GLC_RX + 2;
  float tot = 0.f;
  for (int it = 0; it < reps; ++it) {
    if constexpr (MODE == 0) {
      float2 acc2[GLC_RP][GLC_RX];
      for (int r = 0; r < GLC_RP; ++r) for (int c = 0; c < GLC_RX; ++c) acc2[r][c] = make_float2(0.f, 0.f);
      for (int q = 0; q < 4; ++q) corr_rows2<A, 2>(s_in + origin, PITCH, s_w + q * A * 2 * glc_ulen(A), acc2);
      for (int r = 0; r < GLC_RP; ++r) for (int c = 0; c < GLC_RX; ++c) tot += acc2[r][c].x + acc2[r][c].y;
    } else {
      float2 acc[GLC_RY][GLC_CP];
      for (int r = 0; r < GLC_RY; ++r) for (int c = 0; c < GLC_CP; ++c) acc[r][c] = make_float2(0.f, 0.f);
      for (int q = 0; q < 4; ++q) corr_rows_cp<A, 2>(s_in + origin, PITCH, s_w + q * A * glc_wpitch(A), acc, zero);
      for (int r = 0; r < GLC_RY; ++r) for (int c = 0; c < GLC_CP; ++c) tot += acc[r][c].x + acc[r][c].y;
    }
    s_in[threadIdx.x] += 1e-9f * tot;   // keep iterations dependent
    __syncthreads();
  }
  out[blockIdx.x * NTHR + threadIdx.x] = tot;
}

int main() {
  const int reps = 256, grid = 148 * 4;
  std::vector<float> hin(TILE), hw(8192);
  for (int i = 0; i < TILE; ++i) hin[i] = 0.001f * (i % 97);
  for (int i = 0; i < 8192; ++i) hw[i] = 0.01f * ((i * 7) % 13);
  float *din, *dw, *dout;
  cudaMalloc(&din, TILE * 4); cudaMalloc(&dw, 8192 * 4); cudaMalloc(&dout, grid * NTHR * 4);
  cudaMemcpy(din, hin.data(), TILE * 4, cudaMemcpyHostToDevice); cudaMemcpy(dw, hw.data(), 8192 * 4, cudaMemcpyHostToDevice);
  const size_t smem = (TILE + 4 * A * 2 * glc_ulen(A)) * 4 + 22 * 1024;   // pad to the C2 kernel's footprint (4 CTAs per SM)
  cudaFuncSetAttribute(k_probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    float best = 1e30f;
    for (int t = 0; t < 8; ++t) {
      cudaEventRecord(e0);
      if (mode == 0) k_probe<0><<<grid, NTHR, smem>>>(din, dw, dout, reps, 0u); else k_probe<1><<<grid, NTHR, smem>>>(din, dw, dout, reps, 0u);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (t > 0 && ms < best) best = ms;
    }
    // useful work: per image-equivalent 60 x 60 outputs x 4 phases x 169 taps x 2 flops (150 of the 160 threads map to real outputs)
    const double flops = (double)grid * reps * 3600.0 * 4 * 169 * 2;
    printf("mode %d (%s): %.3f ms  %.1f TFLOP/s useful  err=%s\n", mode, mode ? "column pairs, scalar taps" : "row pairs, tap pairs", best,
           flops / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
