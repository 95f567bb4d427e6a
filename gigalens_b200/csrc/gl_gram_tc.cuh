// gl_gram_tc.cuh -- the batched normal-equation matrix of lstsq_simulate on the 5th-generation tensor
// cores (tcgen05 + TMEM), the one place north_star puts tensor cores.
//
//   gram[b] = A A^T,  A = [X Y]^T  ((D+1) x P, row c = channel c of R[b] times W, last row = obs * W)
//   (tf/simulator.py:232-235: X^T X and X^T Y in one product.)
//
// One CTA (4 warps) per sample.  A is K-major in HBM already (every channel image is contiguous in the
// pixel index), so the same shared-memory tile serves as both MMA operands: D[128 x N] += A_tile[128 x 8] *
// A_tile[N x 8]^T with M = 128 (rows >= D+1 are zero), N = D+1 rounded up to 16, kind::tf32, fp32
// accumulators in TMEM.  TF32 keeps 11 significand bits, far short of the 1e-5 parity bound after the
// pseudo-inverse, so each fp32 value is split v = hi + lo (both TF32-representable) and every K-slice
// issues three MMAs: hi*hi + hi*lo + lo*hi ("3xTF32"; the dropped lo*lo term is 2^-22 relative).
//
// Shared-memory operand layout: the canonical K-major, no-swizzle UMMA layout -- 8-row x 16-byte core
// matrices; here a K-chunk (4 pixels = 16 bytes) of all 128 rows forms one 2 KB panel, so
// LBO (next 16 bytes in K) = 2048 B and SBO (next 8 rows) = 128 B.  Loads are float4 per (row, chunk) with
// each quarter-warp writing 8 consecutive rows of one panel (conflict-free 16-byte stores) and each warp
// reading 64 contiguous bytes of 8 channel rows (full sectors).  Two stages: the CUDA cores convert /
// split stage s+1 while the tensor core consumes stage s; tcgen05.commit -> mbarrier frees a stage.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GTC_THREADS 128
#define GTC_KC 32                                  // pixels per stage
#define GTC_PANEL 2048                             // bytes: 128 rows x 16 B
#define GTC_TILE (GTC_PANEL * (GTC_KC / 4))        // one hi or lo tile: 16 KB
#define GTC_STAGE (2 * GTC_TILE)                   // hi + lo
#define GTC_SMEM (2 * GTC_STAGE + 64)              // two stages + 4 barriers + tmem pointer

__device__ __forceinline__ uint32_t gtc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float gtc_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t gtc_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);                 // start address            [0,14)
  d |= (uint64_t)(GTC_PANEL >> 4) << 16;                    // leading byte offset (K)  [16,30)
  d |= (uint64_t)(128 >> 4) << 32;                          // stride byte offset (M/N) [32,46)
  d |= (uint64_t)1 << 46;                                   // descriptor version (Blackwell)
  return d;                                                 // base offset 0, layout type 0 = no swizzle
}
__device__ __forceinline__ void gtc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void gtc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(gtc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool gtc_wait(uint64_t* bar, uint32_t parity) {   // bounded spin: a broken pipeline must not hang the GPU
  const uint32_t a = gtc_smem_u32(bar);
  for (int it = 0; it < (1 << 22); ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

// The tensor core's fp32 accumulation is not round-to-nearest: over a 3600-pixel chain the Gram drifts by
// ~4e-5 relative (measured, linear in the chain length), which the pseudo-inverse amplifies past the 1e-5
// parity bound.  So the chain is cut into segments of GTC_SEG stages (64 pixels): two TMEM accumulators
// ping-pong, and while the tensor core works on segment g the threads drain segment g-1 (tcgen05.ld) and add
// it to per-thread fp32 registers with round-to-nearest adds (thread r owns Gram row r).
#define GTC_SEG 2
#define GTC_ACC_COLS 128                           // TMEM column stride between the two accumulators

template <int NB>
__device__ __forceinline__ void gtc_drain(uint32_t taddr, float (&acc)[NB * 16]) {
#pragma unroll
  for (int cb = 0; cb < NB; ++cb) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr + (uint32_t)(cb * 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[cb * 16 + q] += __uint_as_float(r[q]);
  }
}

// grid = ksplit x (number of samples in the chunk), block = 128; NB = N / 16 column blocks (N = D+1 rounded up to 16).
// ksplit = 1: one CTA per sample, plain stores.  ksplit = 2: two CTAs per sample, each sums half of the pixel segments and ADDS its
// rows to the (pre-zeroed) Gram with fp32 atomics -- two addends commute, so the result does not depend on which CTA arrives
// first and stays bit-reproducible.  The kernel is latency-bound (one 128-thread CTA per sample converts, stores and issues;
// 11 % of the warp slots busy): twice the CTAs hide twice the latency.
// err_flag (may be null): set to 1 if a barrier wait timed out.
template <int NB>
__global__ void __launch_bounds__(GTC_THREADS) k_gram_tc(int D, int npx, const float* __restrict__ R, const float* __restrict__ w,
                                                        const float* __restrict__ obs, float* __restrict__ gram,
                                                        int* __restrict__ err_flag, int ksplit) {
  extern __shared__ __align__(1024) unsigned char gtc_smem[];
  unsigned char* smem = gtc_smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * GTC_STAGE);     // [0,1]: stage free, [2,3]: accumulator a complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 2 * GTC_STAGE + 32);
  const int tid = threadIdx.x, warp = tid >> 5, b = blockIdx.x / ksplit, half = blockIdx.x - b * ksplit;
  const int Dx = D + 1;
  constexpr int N = NB * 16;
  const float* Rb = R + (size_t)b * D * npx;
  // zero both stages once: rows > D stay zero for the whole kernel
  for (int i = tid; i < 2 * GTC_STAGE / 16; i += GTC_THREADS) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(gtc_smem_u32(bars + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(gtc_smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int nstage_all = (npx + GTC_KC - 1) / GTC_KC;
  const int nseg_all = (nstage_all + GTC_SEG - 1) / GTC_SEG;
  // this CTA's share of the segments: [seg0, seg1), stages [st0, st1)
  const int seg0 = (nseg_all * half) / ksplit, seg1 = (nseg_all * (half + 1)) / ksplit;
  const int st0 = seg0 * GTC_SEG, st1 = (seg1 * GTC_SEG < nstage_all) ? seg1 * GTC_SEG : nstage_all;
  const int nstage = st1 - st0;                      // stages of this CTA (local index st below)
  const int nseg = seg1 - seg0;
  const int nitem = ((Dx + 7) / 8) * 8 * (GTC_KC / 4);   // (row, 16-byte chunk) items per stage, whole 8-row groups
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);   // TMEM lane == accumulator row; warp w owns lanes 32w..32w+31
  float acc[NB * 16];
#pragma unroll
  for (int i = 0; i < NB * 16; ++i) acc[i] = 0.f;
  bool ok = true;
  uint32_t phase[2] = {0u, 0u}, aphase[2] = {0u, 0u};
  // Software pipeline of the operand loads: the raw values of stage st+1 are fetched into registers (float4 per
  // (row, chunk) item, up to GTC_MAXI items per thread) before stage st is converted and stored, so the global-load
  // latency overlaps the split / store / MMA issue of the current stage.
  constexpr int GTC_MAXI = (NB * 16 + 7) / 8 * 8 * (GTC_KC / 4) / GTC_THREADS + 1;
  float4 pre[GTC_MAXI], prw[GTC_MAXI];
  const bool vec_ok = (npx % 4) == 0;               // channel rows are then 16-byte aligned
  auto fetch = [&](int stage) {
    const int p0 = (st0 + stage) * GTC_KC;
#pragma unroll
    for (int k = 0; k < GTC_MAXI; ++k) {
      const int it = tid + k * GTC_THREADS;
      const int r0 = it & 7, kc = (it >> 3) & 7, row = (it >> 6) * 8 + r0;
      const int p = p0 + kc * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f), wv = v;
      if (it < nitem && row < Dx && p < npx) {
        const float* src = (row < D) ? Rb + (size_t)row * npx + p : obs + p;
        if (vec_ok && p + 3 < npx) {
          v = __ldg(reinterpret_cast<const float4*>(src));
          wv = __ldg(reinterpret_cast<const float4*>(w + p));
        } else {
          v.x = __ldg(src); wv.x = __ldg(w + p);
          if (p + 1 < npx) { v.y = __ldg(src + 1); wv.y = __ldg(w + p + 1); }
          if (p + 2 < npx) { v.z = __ldg(src + 2); wv.z = __ldg(w + p + 2); }
          if (p + 3 < npx) { v.w = __ldg(src + 3); wv.w = __ldg(w + p + 3); }
        }
      }
      pre[k] = v; prw[k] = wv;
    }
  };
  fetch(0);
  for (int st = 0; st < nstage; ++st) {
    const int s = st & 1, seg = st / GTC_SEG, a = seg & 1;
    unsigned char* hi = smem + s * GTC_STAGE;
    unsigned char* lo = hi + GTC_TILE;
    if (st >= 2) {                                  // the MMAs that read this stage two iterations ago must have retired
      ok = gtc_wait(bars + s, phase[s]) && ok;
      phase[s] ^= 1u;
    }
#pragma unroll
    for (int k = 0; k < GTC_MAXI; ++k) {
      // item -> (row group, chunk, row in group): a quarter-warp covers 8 consecutive rows of one panel
      const int it = tid + k * GTC_THREADS;
      const int r0 = it & 7, kc = (it >> 3) & 7, row = (it >> 6) * 8 + r0;
      if (it >= nitem || row >= Dx) continue;
      float v[4] = {pre[k].x, pre[k].y, pre[k].z, pre[k].w};
      const float wv[4] = {prw[k].x, prw[k].y, prw[k].z, prw[k].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (v[q] != v[q]) v[q] = 0.f;               // NaN scrub (:228)
        v[q] *= wv[q];
      }
      float4 h, l;
      h.x = gtc_tf32(v[0]); h.y = gtc_tf32(v[1]); h.z = gtc_tf32(v[2]); h.w = gtc_tf32(v[3]);
      l.x = gtc_tf32(v[0] - h.x); l.y = gtc_tf32(v[1] - h.y); l.z = gtc_tf32(v[2] - h.z); l.w = gtc_tf32(v[3] - h.w);
      const int off = kc * GTC_PANEL + row * 16;
      *reinterpret_cast<float4*>(hi + off) = h;
      *reinterpret_cast<float4*>(lo + off) = l;
    }
    if (st + 1 < nstage) fetch(st + 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ah = gtc_smem_u32(hi), al = gtc_smem_u32(lo);
      const uint32_t dacc = tmem + (uint32_t)(a * GTC_ACC_COLS);
      const bool first = (st % GTC_SEG) == 0;       // a new segment overwrites its accumulator
#pragma unroll
      for (int j = 0; j < GTC_KC / 8; ++j) {        // K = 8 per MMA = two 16-byte chunks = two panels
        const uint64_t dh = gtc_desc(ah + j * 2 * GTC_PANEL), dl = gtc_desc(al + j * 2 * GTC_PANEL);
        gtc_mma(dacc, dh, dh, idesc, (first && j == 0) ? 0u : 1u);
        gtc_mma(dacc, dh, dl, idesc, 1u);
        gtc_mma(dacc, dl, dh, idesc, 1u);
      }
      gtc_commit(bars + s);
      if ((st % GTC_SEG) == GTC_SEG - 1 || st == nstage - 1) gtc_commit(bars + 2 + a);
    }
    // drain the previous segment while the tensor core works on this one
    if ((st % GTC_SEG) == 0 && seg >= 1) {
      const int pa = a ^ 1;
      ok = gtc_wait(bars + 2 + pa, aphase[pa]) && ok;
      aphase[pa] ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      gtc_drain<NB>(lane_base + (uint32_t)(pa * GTC_ACC_COLS), acc);
    }
  }
  {   // last segment
    const int pa = (nseg - 1) & 1;
    ok = gtc_wait(bars + 2 + pa, aphase[pa]) && ok;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    gtc_drain<NB>(lane_base + (uint32_t)(pa * GTC_ACC_COLS), acc);
  }
  float* out = gram + (size_t)b * Dx * Dx;
  if (ok && nstage > 0) {
    if (tid < Dx) {
      if (ksplit == 1) {
#pragma unroll
        for (int q = 0; q < NB * 16; ++q)
          if (q < Dx) out[(size_t)tid * Dx + q] = acc[q];
      } else {
#pragma unroll
        for (int q = 0; q < NB * 16; ++q)
          if (q < Dx) atomicAdd(out + (size_t)tid * Dx + q, acc[q]);
      }
    }
  } else if (!ok) {   // a barrier wait timed out: fail loudly (NaN Gram -> NaN amplitudes and likelihood), never silently
    for (int i = tid; i < Dx * Dx; i += GTC_THREADS) out[i] = __int_as_float(0x7fc00000);
    if (tid == 0 && err_flag) *err_flag = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem) : "memory");
}

// once per device (gl_plan_create / the stand-alone probe): opt every instance in to its dynamic shared memory
static inline cudaError_t gl_gram_tc_init() {
  cudaError_t e;
#define GTC_INIT(n) if ((e = cudaFuncSetAttribute(k_gram_tc<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, GTC_SMEM)) != cudaSuccess) return e;
  GTC_INIT(1) GTC_INIT(2) GTC_INIT(3) GTC_INIT(4) GTC_INIT(5) GTC_INIT(6) GTC_INIT(7)
#undef GTC_INIT
  return cudaSuccess;
}

// host-side launch: pick the column-block count for D (gl_gram_tc_init() must have run on this device)
// ksplit = 2 adds into `gram`: the launcher zeroes it first.
static inline cudaError_t gl_launch_gram_tc(int nb, int D, int npx, const float* R, const float* w, const float* obs, float* gram,
                                            int* err_flag, cudaStream_t st, int ksplit = 2) {
  const int NBv = (D + 1 + 15) / 16;
  if (npx < 4 * GTC_KC * GTC_SEG) ksplit = 1;       // too few segments to split
  if (ksplit > 1) {
    cudaError_t e = cudaMemsetAsync(gram, 0, (size_t)nb * (D + 1) * (D + 1) * sizeof(float), st);
    if (e != cudaSuccess) return e;
  }
#define GTC_CASE(n)                                                                                                 \
  case n: {                                                                                                         \
    k_gram_tc<n><<<nb * ksplit, GTC_THREADS, GTC_SMEM, st>>>(D, npx, R, w, obs, gram, err_flag, ksplit);            \
    return cudaGetLastError();                                                                                      \
  }
  switch (NBv) {
    GTC_CASE(1) GTC_CASE(2) GTC_CASE(3) GTC_CASE(4) GTC_CASE(5) GTC_CASE(6) GTC_CASE(7)
    default: return cudaErrorInvalidValue;
  }
#undef GTC_CASE
}
