// Stand-alone check of the tcgen05 Gram kernel (gl_gram_tc.cuh) against a host fp64 Gram (TEST ONLY).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 tests/cuda/gram_tc_check.cu -o tests/cuda/_build/gram_tc_check
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../gigalens_b200/csrc/gl_gram_tc.cuh"

static int run_case(int D, int npx, int bs, int mode, bool verbose) {
  const int Dx = D + 1;
  std::vector<float> R((size_t)bs * D * npx), w(npx), obs(npx);
  srand(1234 + D + npx);
  auto rnd = []() { return (float)rand() / RAND_MAX - 0.5f; };
  for (int p = 0; p < npx; ++p) { w[p] = mode == 0 ? 1.f : 0.5f + (float)rand() / RAND_MAX; obs[p] = mode == 0 ? 0.f : rnd() * 3.f; }
  for (int b = 0; b < bs; ++b)
    for (int c = 0; c < D; ++c)
      for (int p = 0; p < npx; ++p)
        R[((size_t)b * D + c) * npx + p] = mode == 0 ? ((p == c) ? (float)(c + 1) : 0.f) : rnd() * (1.f + c % 5);
  float *dR, *dw, *dobs, *dG; int* derr;
  cudaMalloc(&dR, R.size() * 4); cudaMalloc(&dw, npx * 4); cudaMalloc(&dobs, npx * 4); cudaMalloc(&dG, (size_t)bs * Dx * Dx * 4); cudaMalloc(&derr, 4);
  cudaMemcpy(dR, R.data(), R.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dw, w.data(), npx * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dobs, obs.data(), npx * 4, cudaMemcpyHostToDevice); cudaMemset(dG, 0, (size_t)bs * Dx * Dx * 4); cudaMemset(derr, 0, 4);
  cudaError_t e = gl_gram_tc_init();
  if (e == cudaSuccess) e = gl_launch_gram_tc(bs, D, npx, dR, dw, dobs, dG, derr, 0);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
  std::vector<float> G((size_t)bs * Dx * Dx); int err = 0;
  cudaMemcpy(G.data(), dG, G.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&err, derr, 4, cudaMemcpyDeviceToHost);
  double worst = 0, scale = 0;
  for (int b = 0; b < bs; ++b)
    for (int i = 0; i < Dx; ++i)
      for (int j = 0; j < Dx; ++j) {
        double s = 0;
        for (int p = 0; p < npx; ++p) {
          double a = (i < D ? (double)R[((size_t)b * D + i) * npx + p] : (double)obs[p]) * w[p];
          double c = (j < D ? (double)R[((size_t)b * D + j) * npx + p] : (double)obs[p]) * w[p];
          s += a * c;
        }
        worst = fmax(worst, fabs(s - G[((size_t)b * Dx + i) * Dx + j])); scale = fmax(scale, fabs(s));
      }
  printf("D=%d npx=%d bs=%d mode=%d: timeout=%d max|err|=%.3e scale=%.3e rel=%.3e\n", D, npx, bs, mode, err, worst, scale, worst / scale);
  if (verbose) {
    const int n = Dx < 12 ? Dx : 12;
    for (int i = 0; i < n; ++i) { for (int j = 0; j < n; ++j) printf("%8.2f ", G[(size_t)i * Dx + j]); printf("\n"); }
  }
  cudaFree(dR); cudaFree(dw); cudaFree(dobs); cudaFree(dG); cudaFree(derr);
  return (worst / scale < 2e-6 && !err) ? 0 : 1;
}

// `gram_tc_check time [bs]`: device time of the C3-shaped launch (D = 66, 3600 pixels), CUDA events, best of 5
static int run_timing(int bs) {
  const int D = 66, npx = 3600, Dx = D + 1;
  float *dR, *dw, *dobs, *dG;
  cudaMalloc(&dR, (size_t)bs * D * npx * 4); cudaMalloc(&dw, npx * 4); cudaMalloc(&dobs, npx * 4); cudaMalloc(&dG, (size_t)bs * Dx * Dx * 4);
  cudaMemset(dR, 0x3c, (size_t)bs * D * npx * 4); cudaMemset(dw, 0x3c, npx * 4); cudaMemset(dobs, 0x3c, npx * 4);
  if (gl_gram_tc_init() != cudaSuccess) return 2;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    if (gl_launch_gram_tc(bs, D, npx, dR, dw, dobs, dG, nullptr, 0) != cudaSuccess) return 2;
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  printf("timing: D=%d npx=%d bs=%d: %.3f ms  (%.0f GB/s of operand reads)\n", D, npx, bs, best, (double)bs * D * npx * 4 / best / 1e6);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : 2;
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 't') return run_timing(argc > 2 ? atoi(argv[2]) : 2048);
  int bad = 0;
  bad |= run_case(11, 32, 1, 0, true);     // one stage, one-hot rows: G = diag((c+1)^2)
  bad |= run_case(11, 64, 1, 0, true);
  bad |= run_case(11, 32, 1, 1, false);
  bad |= run_case(66, 3600, 3, 1, false);
  bad |= run_case(22, 3600, 5, 1, false);
  bad |= run_case(100, 1000, 2, 1, false);
  printf(bad ? "FAIL\n" : "PASS\n");
  return bad;
}
