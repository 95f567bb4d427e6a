// Stand-alone probe of the TMA halo-tile load used by k_conv_fwd_tma (development aid).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, int bw, int bh, int c0, int c1, int c2, float* out) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar)), "r"(bw * bh * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(su32(sm)), "l"(&tm), "r"(su32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  unsigned ok = 0;
  for (int spin = 0; spin < (1 << 20) && !ok; ++spin)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(su32(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = ok ? sm[i] : -1.f;
}
int main(int argc, char** argv) {
  const int n = atoi(argv[1]), nimg = 8, bw = atoi(argv[2]), bh = atoi(argv[3]); const int c0 = atoi(argv[4]), c1 = atoi(argv[5]);
  std::vector<float> h((size_t)n * n * nimg);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
  float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, bw * bh * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  printf("entry point: %d %d %p\n", (int)e, (int)q, f);
  CUtensorMap tm;
  const cuuint64_t gdim[3] = {n, n, nimg}; const cuuint64_t gstr[2] = {n * 4, n * n * 4};
  const cuuint32_t box[3] = {bw, bh, 1}; const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((enc_fn)f)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  k<<<1, 128, bw * bh * 4>>>(tm, bw, bh, c0, c1, 3, o);
  e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> ho(bw * bh); cudaMemcpy(ho.data(), o, bw * bh * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < bh; ++r2) for (int c = 0; c < bw; ++c) {
    const int ri = r2 + c1, ci = c + c0;
    const float want = (ri >= 0 && ri < n && ci >= 0 && ci < n) ? h[(size_t)3 * n * n + ri * n + ci] : 0.f;
    if (ho[r2 * bw + c] != want) { if (bad < 5) printf("mismatch at %d,%d: %f vs %f\n", r2, c, ho[r2 * bw + c], want); ++bad; }
  }
  printf("n=%d box=%dx%d c=(%d,%d) mismatches: %d\n", n, bw, bh, c0, c1, bad);
  return 0;
}
