// gl_kernels.cu -- sm_100a kernels and the C ABI of libgigalens_b200.so (include/gigalens_b200.h).
//
// Data flow of one log-prob + gradient evaluation of a batch (all fp32, one stream):
//
//   z[bs][d] --k_sample_fwd--> params[P][bs], derived[bs][ND], logprior[bs]
//   derived  --k_raytrace_fwd--> ss[bs][hs][hs]                  (ray-shoot + light, NaN scrub)
//   ss       --k_conv_fwd-----> image[bs][n][n], like partials, dL/dimage
//   dL/dimage--k_conv_bwd-----> dL/dss[bs][hs][hs]
//   dL/dss   --k_raytrace_bwd-> gpart[bs][chunks][NG]             (recompute + hand adjoint)
//   gpart    --k_sample_bwd---> dparams[P][bs] / dz[bs][d], loglike, red_chi2, logp
//
// No framework autodiff graph and no CPU fallback anywhere: every entry point needs the GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <cmath>
#include <cstring>
#include <type_traits>
#include <string>
#include <vector>

#include "gl_build.h"
#include "gl_conv.cuh"
#include "gl_lstsq.cuh"
#include "gl_positions.cuh"
#include "gl_gram_tc.cuh"
#include "gl_probe.cuh"

// ---------------------------------------------------------------------------------------------
// error handling / bookkeeping
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static int64_t g_launch_count = 0;

static int gl_fail(const std::string& msg) { g_last_error = msg; return 1; }
#define GL_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) return gl_fail(std::string(#call) + ": " + cudaGetErrorString(_e));     \
  } while (0)
#define GL_LAUNCH_CHECK(name)                                                                      \
  do {                                                                                             \
    ++g_launch_count;                                                                              \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess) return gl_fail(std::string("launch ") + name + ": " + cudaGetErrorString(_e)); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// guarded allocations (GL_GUARD=1 in the environment): the stand-in for compute-sanitizer's memcheck, which is closed on the
// GPU pool this was built on.  Every device allocation of the library gets GL_GUARD_BYTES of red zone on both sides, filled
// with 0xFF (a float NaN, an int -1: a kernel that READS out of bounds poisons its result and fails the parity tests);
// gl_guard_check() verifies that no kernel WROTE into any red zone.  Off (the default) this is a plain cudaMalloc / cudaFree.
// ---------------------------------------------------------------------------------------------
#include <cstdlib>
#include <map>
#include <mutex>
#define GL_GUARD_BYTES ((size_t)65536)
static std::mutex g_guard_mu;
static std::map<void*, size_t> g_guard_live;   // user pointer -> user size
static bool gl_guard_on() {
  static const bool on = [] { const char* e = getenv("GL_GUARD"); return e && e[0] && e[0] != '0'; }();
  return on;
}
static cudaError_t gl_malloc(void** out, size_t bytes) {
  if (!gl_guard_on()) return cudaMalloc(out, bytes);
  const size_t user = (bytes + 255) & ~(size_t)255;
  char* base = nullptr;
  cudaError_t e = cudaMalloc((void**)&base, user + 2 * GL_GUARD_BYTES);
  if (e != cudaSuccess) return e;
  if ((e = cudaMemset(base, 0xFF, GL_GUARD_BYTES)) != cudaSuccess) return e;
  if ((e = cudaMemset(base + GL_GUARD_BYTES + bytes, 0xFF, user - bytes + GL_GUARD_BYTES)) != cudaSuccess) return e;
  *out = base + GL_GUARD_BYTES;
  std::lock_guard<std::mutex> lk(g_guard_mu);
  g_guard_live[*out] = bytes;
  return cudaSuccess;
}
static std::string g_guard_sticky;              // first corrupted red zone seen when an allocation was freed
// verify both red zones of one allocation (device synchronised by the caller); "" = intact
static std::string gl_guard_verify(const char* user, size_t bytes) {
  static std::vector<unsigned char> h(GL_GUARD_BYTES + 256);
  const size_t tail = ((bytes + 255) & ~(size_t)255) - bytes + GL_GUARD_BYTES;
  for (int side = 0; side < 2; ++side) {
    const char* src = side ? user + bytes : user - GL_GUARD_BYTES;
    const size_t n = side ? tail : GL_GUARD_BYTES;
    if (cudaMemcpy(h.data(), src, n, cudaMemcpyDeviceToHost) != cudaSuccess) return "red zone read-back failed";
    for (size_t i = 0; i < n; ++i)
      if (h[i] != 0xFF) {
        char msg[160];
        snprintf(msg, sizeof msg, "red zone %s an allocation of %zu bytes was written (offset %td)", side ? "after" : "before", bytes,
                 side ? (ptrdiff_t)i : (ptrdiff_t)i - (ptrdiff_t)GL_GUARD_BYTES);
        return msg;
      }
  }
  return "";
}
static cudaError_t gl_free(void* q) {
  if (!gl_guard_on() || !q) return cudaFree(q);
  {
    std::lock_guard<std::mutex> lk(g_guard_mu);
    auto it = g_guard_live.find(q);
    if (it != g_guard_live.end()) {       // an allocation is checked one last time before it goes away
      cudaDeviceSynchronize();
      const std::string e = gl_guard_verify((const char*)q, it->second);
      if (!e.empty() && g_guard_sticky.empty()) g_guard_sticky = e + " [found at free]";
      g_guard_live.erase(it);
    }
  }
  return cudaFree((char*)q - GL_GUARD_BYTES);
}
#define cudaMalloc(p, n) gl_malloc((void**)(p), (n))
#define cudaFree(q) gl_free((void*)(q))

// ---------------------------------------------------------------------------------------------
// prior / bijector leaves
// ---------------------------------------------------------------------------------------------
struct GlLeaf { int dist; int slot; float a, b, low, high, log_norm; };

__device__ __forceinline__ float gl_softplus(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float gl_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// x = bijector(z); lp = log p(x) + log|dx/dz|; also d x/d z and d lp/d z.
__device__ __forceinline__ void gl_leaf_eval(const GlLeaf& L, float z, float& x, float& lp, float& dxdz, float& dlpdz) {
  const float LOG_2PI = 1.8378770664093453f;
  switch (L.dist) {
    case GL_DIST_NORMAL: {
      x = z; dxdz = 1.f;
      const float u = (x - L.a) / L.b;
      lp = -0.5f * u * u - 0.5f * LOG_2PI - logf(L.b);
      dlpdz = -u / L.b;
    } break;
    case GL_DIST_LOGNORMAL: {
      x = expf(z); dxdz = x;
      // log x = z exactly: -log x - log s - .5 log 2pi - .5 ((log x - m)/s)^2 + fldj(z) = z
      const float lx = logf(x);
      const float u = (lx - L.a) / L.b;
      lp = -lx - logf(L.b) - 0.5f * LOG_2PI - 0.5f * u * u + z;
      dlpdz = -u / L.b;   // (-1 + 1) from -log x + fldj cancel
    } break;
    default: {  // Uniform / TruncatedNormal: Sigmoid(low, high)
      const float lo = (L.dist == GL_DIST_UNIFORM) ? L.a : L.low;
      const float hi = (L.dist == GL_DIST_UNIFORM) ? L.b : L.high;
      const float diff = hi - lo;
      const float sg = gl_sigmoid(z), sgm = gl_sigmoid(-z);
      x = (z < 0.f) ? lo + diff * sg : hi - diff * sgm;
      dxdz = diff * sg * sgm;
      const float fldj = logf(diff) - gl_softplus(-z) - gl_softplus(z);
      const float dfldj = sgm - sg;
      if (L.dist == GL_DIST_UNIFORM) {
        lp = -logf(diff) + fldj;
        dlpdz = dfldj;
      } else {
        const float u = (x - L.a) / L.b;
        lp = -0.5f * u * u - 0.5f * LOG_2PI - logf(L.b) - L.log_norm + fldj;
        dlpdz = -u / L.b * dxdz + dfldj;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// per-sample kernels
// ---------------------------------------------------------------------------------------------
// z -> params (+ logprior); one thread per sample.
__global__ void k_unconstrain(int bs, int d, const GlLeaf* __restrict__ leaves, const float* __restrict__ z,
                              float* __restrict__ params, float* __restrict__ logprior) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bs) return;
  float lp = 0.f;
  for (int k = 0; k < d; ++k) {
    const GlLeaf L = leaves[k];
    float x, l, dx, dl;
    gl_leaf_eval(L, z[(size_t)b * d + k], x, l, dx, dl);
    lp += l;
    if (params && L.slot >= 0) params[(size_t)L.slot * bs + b] = x;
  }
  if (logprior) logprior[b] = lp;
}

// Chain rule of the bijector for a caller-supplied d(.)/d(params): dz = dparams[slot] * dx/dz (+ d(log prior + fldj)/dz).
// Lets a driver differentiate the likelihood terms and the prior separately (tempered SMC targets,
// tf/inference.py:289-302); one thread per sample.
__global__ void k_chain_grad(int bs, int d, const GlLeaf* __restrict__ leaves, const float* __restrict__ z,
                             const float* __restrict__ dparams, int with_prior, float* __restrict__ logprior,
                             float* __restrict__ dz) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bs) return;
  float lp = 0.f;
  for (int k = 0; k < d; ++k) {
    const GlLeaf L = leaves[k];
    float x, l, dx, dl;
    gl_leaf_eval(L, z[(size_t)b * d + k], x, l, dx, dl);
    lp += l;
    const float gp = (dparams && L.slot >= 0) ? dparams[(size_t)L.slot * bs + b] : 0.f;
    dz[(size_t)b * d + k] = gp * dx + (with_prior ? dl : 0.f);
  }
  if (logprior) logprior[b] = lp;
}

// batch maximum of the EPL series ratio f (epl.py:33,37) for the reference-exact trip count
__global__ void k_epl_fmax(GlProgram P, int bs, const float* __restrict__ params, const float* __restrict__ member_factor,
                           int* __restrict__ fmax_bits) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bs) return;
  for (int i = 0; i < P.n_lens; ++i) {
    if (P.prof[i].type != GLT_EPL) continue;
    float raw[GL_MAX_RAW];
    gl_gather_raw<float, float>(P.prof[i], params, bs, b, member_factor, 0, raw);
    float phi, q, c;
    ellip_fwd(raw[2], raw[3], 1.f, phi, q, c);
    const float f = (1.f - q) / (1.f + q);
    if (f > 0.f) atomicMax(fmax_bits + i, __float_as_int(f));   // f >= 0: int order == float order
  }
}

// params -> derived vector; one thread per sample.
__global__ void k_prep(GlProgram P, int bs, const float* __restrict__ params, const float* __restrict__ member_factor,
                       const int* __restrict__ amp_slot, const float* __restrict__ epl_fmax, float* __restrict__ derived) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bs) return;
  gl_sample_prep<float, float>(P, params, bs, b, member_factor, amp_slot, epl_fmax, derived + (size_t)b * P.der_total);
}

// partial sums -> loglike, red_chi2, dparams / dz, logp; one thread per sample.  Combines the pixel
// likelihood with the image-position likelihood as ForwardProbModel.log_prob does (tf/model.py:150-163):
// log-likes add, the reduced chi^2 is the mean of the included terms.
__global__ void k_sample_bwd(GlProgram P, int bs, const float* __restrict__ params, const float* __restrict__ member_factor,
                             const int* __restrict__ amp_slot, const float* __restrict__ derived, const float* __restrict__ gpart, int nchunk,
                             float* __restrict__ gsum /*[bs][g_total] scratch*/, const float* __restrict__ like_part,
                             int ntile, float n_pix_used, float* __restrict__ loglike, float* __restrict__ red_chi2,
                             float* __restrict__ dparams, int d, const GlLeaf* __restrict__ leaves,
                             const float* __restrict__ z, const float* __restrict__ logprior, float* __restrict__ logp,
                             float* __restrict__ dz, int include_pixels, const float* __restrict__ pos_ll,
                             const float* __restrict__ pos_chi2, const float* __restrict__ dpos) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bs) return;
  float ll = 0.f, rc = 0.f;
  int n_chi = 0;
  if (include_pixels) {
    if (like_part) {
      float chi2 = 0.f, norm = 0.f;
      for (int t = 0; t < ntile; ++t) {
        chi2 += like_part[((size_t)b * ntile + t) * 2];
        norm += like_part[((size_t)b * ntile + t) * 2 + 1];
      }
      ll = -0.5f * (chi2 + norm);
      rc = chi2 / n_pix_used;
    } else {
      ll = loglike[b];   // lstsq path: k_lstsq_image already wrote log-like and red chi^2
      rc = red_chi2 ? red_chi2[b] : 0.f;
    }
    ++n_chi;
  }
  if (pos_ll) { ll += pos_ll[b]; rc += pos_chi2[b]; ++n_chi; }
  if (loglike) loglike[b] = ll;
  if (red_chi2) red_chi2[b] = rc / (float)n_chi;
  if (logp) logp[b] = ll + (logprior ? logprior[b] : 0.f);
  if (!dparams) return;
  if (include_pixels) {
    float* g = gsum + (size_t)b * P.g_total;
    for (int k = 0; k < P.g_total; ++k) {
      float s = 0.f;
      for (int c = 0; c < nchunk; ++c) s += gpart[((size_t)b * nchunk + c) * P.g_total + k];
      g[k] = s;
    }
    gl_sample_prep_bwd<float, float>(P, params, bs, b, member_factor, amp_slot, derived + (size_t)b * P.der_total, g, dparams);
  } else {
    for (int k = 0; k < P.n_params; ++k) dparams[(size_t)k * bs + b] = 0.f;
  }
  if (dpos)
    for (int k = 0; k < P.n_params; ++k) dparams[(size_t)k * bs + b] += dpos[(size_t)k * bs + b];
  if (dz) {
    for (int k = 0; k < d; ++k) {
      const GlLeaf L = leaves[k];
      float x, l, dx, dl;
      gl_leaf_eval(L, z[(size_t)b * d + k], x, l, dx, dl);
      const float gp = (L.slot >= 0) ? dparams[(size_t)L.slot * bs + b] : 0.f;
      dz[(size_t)b * d + k] = gp * dx + dl;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ray-shooting + light: forward and adjoint.  One thread per (sample, ss pixel), PPT pixels per
// thread per batch for ILP, per-sample derived constants staged in shared memory.
//   grid = (chunks, bs); a CTA walks pixel batches  chunk, chunk + chunks, ...
// ---------------------------------------------------------------------------------------------
// Dynamic shared memory above 48 KB needs cudaFuncAttributeMaxDynamicSharedMemorySize; the limit applies to static +
// dynamic bytes, and the kernels carry up to ~1 KB of static shared memory (reduction scratch, mbarriers), so every
// launch opts in from 40 KB of dynamic shared memory upwards (a 48.4 KB tile + 1 KB static used to fail to launch).
#define GLK_THREADS 256
// shared-memory layout of the adjoint ray-tracing kernels (floats): [der_total][nwarps * g_total][series scratch]
//   accumulators: one row of g_total floats per warp, or (rows = true, packed kernels) a g_total x 33 staging tile per warp
__host__ __device__ inline int gl_scr_offset(const GlProgram& P, bool rows = false) {
  return (P.der_total + (rows ? 33 : 1) * (GLK_THREADS / 32) * P.g_total + 3) & ~3;
}
__host__ __device__ inline int gl_bwd_smem_floats(const GlProgram& P, int ppt, bool rows = false) {
  return gl_scr_offset(P, rows) + (P.scr_prof >= 0 ? GL_EPL_NSTATE * ppt * GLK_THREADS : 0);
}
// packed adjoint kernel: per-thread prefetch slots (x, y, dL/dss pairs of the next batch) behind the series scratch
__host__ __device__ inline int gl_pre_offset(const GlProgram& P, bool rows, int ppt) { return gl_bwd_smem_floats(P, ppt, rows); }
__host__ __device__ inline int gl_bwd_p_smem_floats(const GlProgram& P, int ppt, bool rows, bool tape = false) {
  return gl_pre_offset(P, rows, ppt) + (tape ? 11 : 3) * ppt * GLK_THREADS;
}

template <int PPT, unsigned F>
__global__ void __launch_bounds__(GLK_THREADS) k_raytrace_fwd(GlProgram P, int npix, const float* __restrict__ grid_x,
                                                              const float* __restrict__ grid_y,
                                                              const unsigned char* __restrict__ ss_mask,
                                                              const float* __restrict__ derived, int no_deflection,
                                                              float* __restrict__ ss_img, int* __restrict__ nan_count) {
  extern __shared__ __align__(16) float s_der[];
  const int b = blockIdx.y;
  const float* dsrc = derived + (size_t)b * P.der_total;
  for (int i = threadIdx.x; i < P.der_total; i += blockDim.x) s_der[i] = dsrc[i];
  __syncthreads();
  const int per_batch = GLK_THREADS * PPT;
  const int nbatch = (npix + per_batch - 1) / per_batch;
  float* dst = ss_img + (size_t)b * npix;
  for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
    if (batch * per_batch + (int)(threadIdx.x & ~31u) >= npix) break;   // this warp has no pixel left (ragged last batch)
    float x[PPT], y[PPT], v[PPT];
    int pix[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      pix[j] = batch * per_batch + j * GLK_THREADS + threadIdx.x;
      const int p = pix[j] < npix ? pix[j] : 0;
      x[j] = __ldg(grid_x + p); y[j] = __ldg(grid_y + p);
    }
    gl_pix_image<float, PPT, F>(P, s_der, x, y, no_deflection != 0, v);
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      if (pix[j] < npix) {
        float o = v[j];
        if (o != o) { o = 0.f; if (nan_count) atomicAdd(nan_count + b, 1); }   // tf.where(is_nan(img), 0, img)  (tf/simulator.py:140)
        if (ss_mask && !ss_mask[pix[j]]) o = 0.f;              // pixels outside pix_region are never evaluated (:34-44)
        dst[pix[j]] = o;
      }
    }
  }
}

// Warp-level reduction of the <= 8 dvar cotangents of one profile: a reduce-scatter butterfly (each
// exchange halves the number of live values per lane: 4+2+1 exchanges, then two plain steps), 9
// shuffles instead of 8 x 5.  Afterwards lane L (L % 4 == 0) holds the warp total of value
// v = 4*bit4(L) + 2*bit3(L) + bit2(L) and adds it to this warp's accumulator row in shared memory.
struct DevFlush {
  float* s_acc;   // this warp's accumulator row in shared memory
  int lane;
  __device__ __forceinline__ void operator()(const GlF2* acc2, int n, int off) {   // two-pixel lanes: fold first
    float acc[GL_MAX_DVARS];
#pragma unroll
    for (int k = 0; k < GL_MAX_DVARS; ++k) acc[k] = acc2[k].x + acc2[k].y;
    (*this)(acc, n, off);
  }
  __device__ __forceinline__ void operator()(const float* acc, int n, int off) {
    const unsigned FULL = 0xffffffffu;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float a4[4], a2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = b4 ? acc[i] : acc[i + 4], keep = b4 ? acc[i + 4] : acc[i];
      a4[i] = keep + __shfl_xor_sync(FULL, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = b3 ? a4[i] : a4[i + 2], keep = b3 ? a4[i + 2] : a4[i];
      a2[i] = keep + __shfl_xor_sync(FULL, send, 8);
    }
    float a1 = (b2 ? a2[1] : a2[0]) + __shfl_xor_sync(FULL, b2 ? a2[0] : a2[1], 4);
    a1 += __shfl_xor_sync(FULL, a1, 2);
    a1 += __shfl_xor_sync(FULL, a1, 1);
    const int v = (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
    if ((lane & 3) == 0 && v < n) s_acc[off + v] += a1;
  }
};

// tf.where(is_nan(img), 0, img) (tf/simulator.py:140) in the adjoint: a scrubbed pixel passes no gradient.  The
// forward kernels count scrubbed pixels per sample; for the (rare) samples that had any, this kernel re-evaluates
// the forward and zeroes the cotangent of the scrubbed pixels before the adjoint kernel reads it.  CTAs of clean
// samples exit at once, and the adjoint kernels stay free of the forward code (registers, instruction cache).
template <int PPT, unsigned F>
__global__ void __launch_bounds__(GLK_THREADS) k_nan_cotangent_mask(GlProgram P, int npix, const float* __restrict__ grid_x,
                                                                    const float* __restrict__ grid_y,
                                                                    const float* __restrict__ derived, int no_deflection,
                                                                    float* __restrict__ gss, const int* __restrict__ nan_count, int bs) {
  extern __shared__ __align__(16) float s_der[];
  // a few CTAs walk the samples (the common case is "no sample has a scrubbed pixel": 4096 one-CTA-per-sample launches cost 8 us
  // to find that out, 592 CTAs that read the flags cost 3)
  for (int b = blockIdx.y; b < bs; b += gridDim.y) {
    if (nan_count[b] == 0) continue;       // uniform over the CTA
    __syncthreads();                       // the previous sample's s_der readers are done
    const float* dsrc = derived + (size_t)b * P.der_total;
    for (int i = threadIdx.x; i < P.der_total; i += blockDim.x) s_der[i] = dsrc[i];
    __syncthreads();
    const int per_batch = GLK_THREADS * PPT;
    const int nbatch = (npix + per_batch - 1) / per_batch;
    float* g = gss + (size_t)b * npix;
    for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
      if (batch * per_batch + (int)(threadIdx.x & ~31u) >= npix) break;   // this warp has no pixel left (ragged last batch)
      float x[PPT], y[PPT], v[PPT];
      int pix[PPT];
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        pix[j] = batch * per_batch + j * GLK_THREADS + threadIdx.x;
        const int p = pix[j] < npix ? pix[j] : 0;
        x[j] = __ldg(grid_x + p); y[j] = __ldg(grid_y + p);
      }
      gl_pix_image<float, PPT, F>(P, s_der, x, y, no_deflection != 0, v);
#pragma unroll
      for (int j = 0; j < PPT; ++j)
        if (pix[j] < npix && v[j] != v[j]) g[pix[j]] = 0.f;
    }
  }
}

// Sample selection: a CTA works on sample b iff (nan_count[b] & sel_mask) == sel_want.  Forward-model path: 0, 0 = every sample
// (nan_count is a count there).  lstsq path: nan_count[b] holds FLAGS -- bit 0: the forward pass scrubbed a NaN component value (the
// reference scrubs per component, tf/simulator.py:200: such samples go to the SCRUB instance, which masks the cotangent per light
// profile), bit 1: the sample's amplitudes come from the eigen-solve, which runs on a side stream while the other samples'' adjoint
// proceeds (gl_lstsq_loglike_core).
template <int PPT, unsigned F, bool SCRUB = false>
__global__ void __launch_bounds__(GLK_THREADS) k_raytrace_bwd(GlProgram P, int npix, const float* __restrict__ grid_x,
                                                              const float* __restrict__ grid_y,
                                                              const unsigned char* __restrict__ ss_mask,
                                                              const float* __restrict__ derived, int no_deflection,
                                                              const float* __restrict__ gss, float* __restrict__ gpart,
                                                              const int* __restrict__ nan_count, int sel_mask, int sel_want) {
  extern __shared__ __align__(16) float smem[];
  float* s_der = smem;                               // [der_total]
  float* s_acc = smem + P.der_total;                 // [nwarps][g_total]
  const int b = blockIdx.y;
  if (sel_mask && (nan_count[b] & sel_mask) != sel_want) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = GLK_THREADS / 32;
  const float* dsrc = derived + (size_t)b * P.der_total;
  for (int i = threadIdx.x; i < P.der_total; i += blockDim.x) s_der[i] = dsrc[i];
  for (int i = threadIdx.x; i < nw * P.g_total; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  DevFlush flush{s_acc + warp * P.g_total, lane};
  // series state of the EPL entry, parked by the forward sweep for the adjoint sweep: [GL_EPL_NSTATE * PPT][threads]
  float* scr = (P.scr_prof >= 0) ? smem + gl_scr_offset(P) + threadIdx.x : nullptr;
  const int per_batch = GLK_THREADS * PPT;
  const int nbatch = (npix + per_batch - 1) / per_batch;
  const float* gsrc = gss + (size_t)b * npix;
  for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
    if (batch * per_batch + (int)(threadIdx.x & ~31u) >= npix) break;   // this warp has no pixel left (ragged last batch)
    float x[PPT], y[PPT], gs[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int pix = batch * per_batch + j * GLK_THREADS + threadIdx.x;
      const bool ok = pix < npix && (!ss_mask || ss_mask[pix]);
      const int p = pix < npix ? pix : 0;
      x[j] = __ldg(grid_x + p); y[j] = __ldg(grid_y + p);
      gs[j] = ok ? __ldg(gsrc + p) : 0.f;
    }
    gl_pix_image_bwd<float, PPT, F, DevFlush, SCRUB>(P, s_der, x, y, gs, no_deflection != 0, flush, scr, GLK_THREADS);
  }
  __syncthreads();
  float* out = gpart + ((size_t)b * gridDim.x + blockIdx.x) * P.g_total;
  for (int k = threadIdx.x; k < P.g_total; k += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += s_acc[w * P.g_total + k];
    out[k] = s;
  }
}

// Two-pixel packed variants (lane type GlF2): each thread owns PPT/2 pairs of ADJACENT pixels, every
// per-pixel +, *, fma issues as one FFMA2 / FMUL2 / FADD2.  Used when the program only contains
// profiles whose arithmetic is written over the lane type (feature set FS0) and npix is even.
// TAPE (programs with a forward-mode scaling-relation group, gradient requested): also writes beta and the group's 2 x 3 Jacobian,
// tape[b][8][npix] = (beta_x, beta_y, Jx0, Jy0, Jx1, Jy1, Jx2, Jy2), so that the adjoint kernel skips the member loop.
template <int PPT, unsigned F, bool BS = false, bool TAPE = false>   // BS: straight-line driver of the benchmark-shape program (gl_program.h)
__global__ void __launch_bounds__(GLK_THREADS, TAPE ? 2 : 1) k_raytrace_fwd_p(GlProgram P, int npix, const float* __restrict__ grid_x,
                                                                const float* __restrict__ grid_y,
                                                                const unsigned char* __restrict__ ss_mask,
                                                                const float* __restrict__ derived, int no_deflection,
                                                                float* __restrict__ ss_img, int* __restrict__ nan_count,
                                                                float* __restrict__ tape) {
  extern __shared__ __align__(16) float s_der[];
  const int b = blockIdx.y;
  const float* dsrc = derived + (size_t)b * P.der_total;
  for (int i = threadIdx.x; i < P.der_total; i += blockDim.x) s_der[i] = dsrc[i];
  __syncthreads();
  constexpr int NV = PPT / 2;
  const int npair = npix >> 1;
  const int per_batch = GLK_THREADS * NV;
  const int nbatch = (npair + per_batch - 1) / per_batch;
  float2* dst = reinterpret_cast<float2*>(ss_img + (size_t)b * npix);
  const float2* gx2 = reinterpret_cast<const float2*>(grid_x);
  const float2* gy2 = reinterpret_cast<const float2*>(grid_y);
  for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
    if (batch * per_batch + (int)(threadIdx.x & ~31u) >= npair) break;   // this warp has no pixel left (ragged last batch)
    GlF2 x[NV], y[NV], v[NV];
    int pr[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      pr[j] = batch * per_batch + j * GLK_THREADS + threadIdx.x;
      const int p = pr[j] < npair ? pr[j] : 0;
      const float2 a = __ldg(gx2 + p), c = __ldg(gy2 + p);
      x[j] = GlF2(a.x, a.y); y[j] = GlF2(c.x, c.y);
    }
    if constexpr (TAPE) {
      GlF2 bx[NV], by[NV], Jx[3][NV], Jy[3][NV];
      gl_pix_image_tape<GlF2, NV, F>(P, s_der, x, y, v, bx, by, Jx, Jy);
      float2* tp = reinterpret_cast<float2*>(tape + (size_t)b * 8 * npix);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (pr[j] < npair) {
          tp[pr[j]] = make_float2(bx[j].x, bx[j].y);
          tp[(size_t)npair + pr[j]] = make_float2(by[j].x, by[j].y);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            tp[(size_t)(2 + 2 * k) * npair + pr[j]] = make_float2(Jx[k][j].x, Jx[k][j].y);
            tp[(size_t)(3 + 2 * k) * npair + pr[j]] = make_float2(Jy[k][j].x, Jy[k][j].y);
          }
        }
      }
    } else if constexpr (BS) gl_pix_image_bs<GlF2, NV>(P, s_der, x, y, v);
    else gl_pix_image<GlF2, NV, F>(P, s_der, x, y, no_deflection != 0, v);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (pr[j] < npair) {
        float o0 = v[j].x, o1 = v[j].y;
        if (o0 != o0) { o0 = 0.f; if (nan_count) atomicAdd(nan_count + b, 1); }
        if (o1 != o1) { o1 = 0.f; if (nan_count) atomicAdd(nan_count + b, 1); }
        if (ss_mask) { if (!ss_mask[2 * pr[j]]) o0 = 0.f; if (!ss_mask[2 * pr[j] + 1]) o1 = 0.f; }
        dst[pr[j]] = make_float2(o0, o1);
      }
    }
  }
}

// Staged flush: every lane parks its (lane-folded) partial cotangents in the warp's staging tile
// stage[dvar][lane] (pitch 33: conflict-free both ways); once per pixel batch lane k sums row k -- 32 LDS + 32 FADD
// for ALL dvars of the program instead of one 9-shuffle butterfly (with its selects) per profile -- and keeps the
// running total of dvar k in a register.  Needs nwarps * g_total * 33 floats of shared memory and g_total <= 64.
#define GLK_STAGE_PITCH 33
struct DevFlushStage {
  float* stage;   // &stage[0][lane] of this warp
  __device__ __forceinline__ void operator()(const GlF2* acc2, int n, int off) {
#pragma unroll
    for (int k = 0; k < GL_MAX_DVARS; ++k)
      if (k < n) stage[(off + k) * GLK_STAGE_PITCH] = acc2[k].x + acc2[k].y;
  }
};

template <int PPT, unsigned F, bool ROWS, bool BS = false, bool TAPE = false>
__global__ void __launch_bounds__(GLK_THREADS, 2) k_raytrace_bwd_p(GlProgram P, int npix, const float* __restrict__ grid_x,
                                                                const float* __restrict__ grid_y,
                                                                const unsigned char* __restrict__ ss_mask,
                                                                const float* __restrict__ derived, int no_deflection,
                                                                const float* __restrict__ gss, float* __restrict__ gpart,
                                                                const int* __restrict__ nan_count, int sel_mask, int sel_want,
                                                                const float* __restrict__ tape) {
  extern __shared__ __align__(16) float smem[];
  float* s_der = smem;
  float* s_acc = smem + P.der_total;   // ROWS: staging tiles [nwarps][g_total][33], else accumulator rows [nwarps][g_total]
  const int b = blockIdx.y;
  if (sel_mask && (nan_count[b] & sel_mask) != sel_want) return;   // lstsq path: sample selection (see k_raytrace_bwd)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = GLK_THREADS / 32;
  const float* dsrc = derived + (size_t)b * P.der_total;
  for (int i = threadIdx.x; i < P.der_total; i += blockDim.x) s_der[i] = dsrc[i];
  const int nacc = ROWS ? nw * P.g_total * GLK_STAGE_PITCH : nw * P.g_total;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  constexpr int NV = PPT / 2;
  GlF2* scr = (P.scr_prof >= 0) ? reinterpret_cast<GlF2*>(smem + gl_scr_offset(P, ROWS)) + threadIdx.x : nullptr;   // [GL_EPL_NSTATE * NV][threads]
  const int npair = npix >> 1;
  const int per_batch = GLK_THREADS * NV;
  const int nbatch = (npair + per_batch - 1) / per_batch;
  const float2* gsrc = reinterpret_cast<const float2*>(gss + (size_t)b * npix);
  const float2* gx2 = reinterpret_cast<const float2*>(grid_x);
  const float2* gy2 = reinterpret_cast<const float2*>(grid_y);
  // The batch inputs (grid x, y and dL/d(ss) of this thread's pixel pairs) are prefetched one batch ahead with
  // cp.async into a per-thread shared-memory slot: the slot is read into registers at the top of a batch and
  // refilled for the next one at once, so the HBM / L2 latency of the loads hides behind a whole batch of
  // arithmetic (the first consumer of the plain loads was the top stall of the kernel, 7 % of its samples).
  constexpr int NSLOT = TAPE ? 11 : 3;   // per pixel pair: x, y, dL/d(ss) (+ the 8 tape planes)
  float2* pre = reinterpret_cast<float2*>(smem + gl_pre_offset(P, ROWS, PPT)) + threadIdx.x;   // [NSLOT * NV][threads]
  const float2* tp = TAPE ? reinterpret_cast<const float2*>(tape + (size_t)b * 8 * npix) : nullptr;
  auto prefetch = [&](int batch) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int pr = batch * per_batch + j * GLK_THREADS + threadIdx.x;
      const unsigned nb = (batch < nbatch && pr < npair) ? 8u : 0u;    // 0 source bytes = zero fill
      const int p = nb ? pr : 0;
      const unsigned d0 = (unsigned)__cvta_generic_to_shared(pre + (NSLOT * j) * GLK_THREADS);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d0), "l"(gx2 + p), "r"(nb) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d0 + 8u * GLK_THREADS), "l"(gy2 + p), "r"(nb) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d0 + 16u * GLK_THREADS), "l"(gsrc + p), "r"(nb) : "memory");
      if constexpr (TAPE) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d0 + 8u * GLK_THREADS * (3 + k)), "l"(tp + (size_t)k * npair + p), "r"(nb) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto sweep = [&](auto& flush, auto&& after_batch) {
    prefetch(blockIdx.x);
    for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
      if (batch * per_batch + (int)(threadIdx.x & ~31u) >= npair) break;   // this warp has no pixel left (ragged last batch)
      GlF2 x[NV], y[NV], gs[NV];
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      GlF2 tbx[NV], tby[NV], tJx[3][NV], tJy[3][NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float2 a = pre[(NSLOT * j) * GLK_THREADS], c = pre[(NSLOT * j + 1) * GLK_THREADS];
        float2 gv = pre[(NSLOT * j + 2) * GLK_THREADS];
        if constexpr (TAPE) {
          const float2 t0 = pre[(NSLOT * j + 3) * GLK_THREADS], t1 = pre[(NSLOT * j + 4) * GLK_THREADS];
          tbx[j] = GlF2(t0.x, t0.y); tby[j] = GlF2(t1.x, t1.y);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float2 u0 = pre[(NSLOT * j + 5 + 2 * k) * GLK_THREADS], u1 = pre[(NSLOT * j + 6 + 2 * k) * GLK_THREADS];
            tJx[k][j] = GlF2(u0.x, u0.y); tJy[k][j] = GlF2(u1.x, u1.y);
          }
        }
        x[j] = GlF2(a.x, a.y); y[j] = GlF2(c.x, c.y);
        if (ss_mask) {
          const int pr = batch * per_batch + j * GLK_THREADS + threadIdx.x;
          if (pr < npair) { if (!ss_mask[2 * pr]) gv.x = 0.f; if (!ss_mask[2 * pr + 1]) gv.y = 0.f; }
        }
        gs[j] = GlF2(gv.x, gv.y);
      }
      prefetch(batch + gridDim.x);
      if constexpr (TAPE)
        gl_pix_image_bwd<GlF2, NV, F, std::remove_reference_t<decltype(flush)>, false, true>(P, s_der, x, y, gs, false, flush, (GlF2*)nullptr, 0,
                                                                                              tbx, tby, tJx, tJy);
      else if constexpr (BS) gl_pix_image_bwd_bs<GlF2, NV>(P, s_der, x, y, gs, flush, scr, GLK_THREADS);
      else gl_pix_image_bwd<GlF2, NV, F>(P, s_der, x, y, gs, no_deflection != 0, flush, scr, GLK_THREADS);
      after_batch();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  };
  float* out = gpart + ((size_t)b * gridDim.x + blockIdx.x) * P.g_total;
  if constexpr (ROWS) {
    float* stage = s_acc + warp * P.g_total * GLK_STAGE_PITCH;
    DevFlushStage flush{stage + lane};
    float tot0 = 0.f, tot1 = 0.f;          // running totals of dvars lane and lane + 32
    auto row_sum = [&](int k) {
      const float* r = stage + k * GLK_STAGE_PITCH;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 4) { s0 += r[i]; s1 += r[i + 1]; s2 += r[i + 2]; s3 += r[i + 3]; }
      return (s0 + s1) + (s2 + s3);
    };
    sweep(flush, [&]() {
      __syncwarp();
      if (lane < P.g_total) tot0 += row_sum(lane);
      if (lane + 32 < P.g_total) tot1 += row_sum(lane + 32);
      __syncwarp();
    });
    __syncthreads();
    if (lane < P.g_total) s_acc[warp * P.g_total + lane] = tot0;          // reuse the head of the staging area: [nwarps][g_total]
    if (lane + 32 < P.g_total) s_acc[warp * P.g_total + lane + 32] = tot1;
    __syncthreads();
  } else {
    DevFlush flush{s_acc + warp * P.g_total, lane};
    sweep(flush, []() {});
    __syncthreads();
  }
  for (int k = threadIdx.x; k < P.g_total; k += blockDim.x) {
    float s2 = 0.f;
    for (int w = 0; w < nw; ++w) s2 += s_acc[w * P.g_total + k];
    out[k] = s2;
  }
}

// phase-major -> natural row-major (diagnostic output of gl_simulate_ss only)
__global__ void k_unpermute(int npix, int bs, const int* __restrict__ perm, const float* __restrict__ src, float* __restrict__ dst) {
  const int b = blockIdx.y;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < npix; k += gridDim.x * blockDim.x)
    dst[(size_t)b * npix + perm[k]] = src[(size_t)b * npix + k];
}

// deflection / beta / surface brightness at arbitrary points shared by all samples
// mode 0: beta (x - alpha), 1: alpha, 2: surface brightness (lens light at theta + source at beta),
// 3: the unit-amplitude linear light components (what light() of a use_lstsq profile returns, sersic.py:31-35,
//    shapelets.py:62-63,72-73): out0[b][c][p]
template <unsigned F>
__global__ void k_points(GlProgram P, int npts, const float* __restrict__ px, const float* __restrict__ py,
                         const float* __restrict__ derived, int mode, float* __restrict__ out0, float* __restrict__ out1) {
  extern __shared__ __align__(16) float s_der[];
  const int b = blockIdx.y;
  const float* dsrc = derived + (size_t)b * P.der_total;
  for (int i = threadIdx.x; i < P.der_total; i += blockDim.x) s_der[i] = dsrc[i];
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += gridDim.x * blockDim.x) {
    float x[1] = {px[p]}, y[1] = {py[p]};
    if (mode == 2) {
      float v[1];
      gl_pix_image<float, 1, F>(P, s_der, x, y, false, v);
      out0[(size_t)b * npts + p] = v[0];
    } else if (mode == 3) {
      float bx[1], by[1];
      gl_pix_beta<float, 1, F>(P, s_der, x, y, bx, by);
      gl_point_components<float, F>(P, s_der, x[0], y[0], bx[0], by[0], out0 + (size_t)b * P.depth * npts + p, npts, true);
    } else {
      float bx[1], by[1];
      gl_pix_beta<float, 1, F>(P, s_der, x, y, bx, by);
      if (mode == 1) { bx[0] = x[0] - bx[0]; by[0] = y[0] - by[0]; }
      out0[(size_t)b * npts + p] = bx[0];
      out1[(size_t)b * npts + p] = by[0];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
// the conv geometry whose tap table travels as a kernel parameter (BASELINE.json: 13 x 13 PSF at ss = 2 => 13 taps per phase)
#define GLC_A_HOT 13
#define GLC_SS_HOT 2
#define GLC_NF_HOT (GLC_SS_HOT * GLC_SS_HOT * GLC_A_HOT * 2 * glc_ulen(GLC_A_HOT))
static_assert(GLC_NF_HOT * 4 <= GLC_CONST_TAP_BYTES, "hot tap table must fit the kernel parameter space");
struct gl_plan {
  int device = 0;
  int bs = 0;
  GlProgram prog;
  int n = 0, ss = 1, hs = 0, npix = 0;
  float conversion_factor = 1.f;
  bool has_epl = false;
  int feat_idx = 3;
  int epl_batch_max = 0;
  int straight_line = 1;     // benchmark-shape programs run the straight-line pixel drivers (0 = interpreter, for A/B)
  int row_flush = 1;         // packed adjoint kernels: staged shared-memory flush (0 = warp butterfly per profile, for A/B)
  // static inputs
  float* d_grid_x = nullptr; float* d_grid_y = nullptr;
  unsigned char* d_ss_mask = nullptr; unsigned char* d_mask = nullptr;
  float* d_member_factor = nullptr;
  int* d_amp_slot = nullptr;
  int* d_perm = nullptr;
  int* d_nan = nullptr;      // [bs] NaN-scrubbed ss pixels of the last forward pass
  // tape of the forward-mode group (cluster models): beta + 2 x 3 Jacobian per ss pixel, written by the forward kernel when a
  // gradient is requested and read by the adjoint kernel, which then never walks the member loop again
  float* d_tape = nullptr;   // [bs][8][npix]
  int use_tape = 1;          // 0 = recompute in the adjoint kernel (A/B; also the automatic fallback when the tape does not fit)
  bool tape_valid = false;   // the last forward pass wrote the tape (same parameters as the adjoint that follows)
  // image-position likelihood (gl_plan_set_positions)
  int gram_tc = 1;           // lstsq normal equations on the tensor cores (tcgen05, 3xTF32); 0 = FP32-FMA k_gram (A/B testing)
  int gram_ksplit = 2;       // CTAs per sample of the tensor-core Gram (2 = split the pixel segments in two halves; 1 for A/B)
  int include_pixels = 1, include_positions = 0;
  int pos_npts = 0, pos_nsys = 0;
  float pos_n_position = 0.f;
  int* d_pos_off = nullptr;
  float* d_pos_x = nullptr; float* d_pos_y = nullptr; float* d_pos_ex = nullptr; float* d_pos_ey = nullptr;
  float* d_pos_ll = nullptr; float* d_pos_chi = nullptr; float* d_pos_grad = nullptr;   // [bs], [bs], [P][bs]
  float* d_tables = nullptr;
  float* d_wf = nullptr; float* d_wb = nullptr;     // forward / flipped taps [nph][A][wpitch]
  int A = 1, pad = 0;
  GlConvGeom gf, gb;
  int bwd_rc_start = 0;
  size_t smem_cf = 0, smem_cb = 0, smem_cf_tma = 0;
  bool conv_tma_ok = false;  // geometry admits the TMA-staged forward kernel
  int conv_tma = 1;          // stage the forward conv tiles with TMA (cp.async.bulk.tensor); 0 = cp.async loader (A/B)
  CUtensorMap tmap_f; const float* tmap_f_base = nullptr; int tmap_f_nimg = -1; bool tmap_f_ok = false;
  CUtensorMap tmap_f2; const float* tmap_f2_base = nullptr; int tmap_f2_nimg = -1; bool tmap_f2_ok = false;   // second lstsq slot
  bool conv_tma_b_ok = false; size_t smem_cb_tma = 0;   // same for the adjoint (TMA load of dL/d(image), TMA store of dL/d(ss))
  CUtensorMap tmap_bi, tmap_bo; const float* tmap_bi_base = nullptr; const float* tmap_bo_base = nullptr; int tmap_b_nimg = -1; bool tmap_b_ok = false;
  int conv_threads_f = 0, conv_threads_b = 0;
  // tap tables as by-value kernel parameters (uniform-datapath taps, gl_conv.cuh GlTapsConst): the hot geometry A = 13, ss = 2
  int conv_const = 0;        // 1 = tap table through the uniform datapath (measured 3.5x SLOWER than broadcast LDS: kept for A/B only)
  bool conv_const_ok = false;
  GlTapsC<GLC_NF_HOT> ct_f, ct_b;
  size_t smem_cf_tma_c = 0, smem_cb_tma_c = 0;
  // likelihood
  bool has_like = false;
  float* d_obs = nullptr; float* d_err = nullptr;
  float bg2 = 0.f, inv_exp = 0.f, n_pix_used = 0.f, n_mask_used = 0.f;
  // prior
  int d = 0;
  GlLeaf* d_leaves = nullptr;
  // workspace
  float* d_params = nullptr;     // [P][bs] (for the z entry points)
  float* d_dparams = nullptr;    // [P][bs]
  float* d_derived = nullptr;    // [bs][der_total]
  float* d_ss = nullptr;         // [bs][npix]  ss image, reused for dL/dss
  float* d_img = nullptr;        // [bs][n*n]
  float* d_gimg = nullptr;       // [bs][n*n]
  float* d_like_part = nullptr;  // [bs][tiles][2]
  float* d_gpart = nullptr;      // [bs][chunks][g_total]
  float* d_gsum = nullptr;       // [bs][g_total]
  float* d_logprior = nullptr;   // [bs]
  float* d_fmax = nullptr;       // [GL_MAX_PROF]
  float* d_z = nullptr; float* d_logp = nullptr; float* d_chi = nullptr; float* d_dz = nullptr;  // *_host staging
  int chunks = 1;
  int sm_count = 148;
  // lstsq workspace (allocated on first use)
  int lstsq = 0;
  int no_deflection = 0;
  int use_packed = 1;
  // optional per-stage timing with CUDA events on the launch stream (bench.py roofline)
  int tm_slots = 0, tm_calls = 0, tm_have0 = 0;
  std::vector<cudaEvent_t> tm_ev;   // [tm_slots][GL_NSTAGE + 1]
  int lq_chunk = 0;
  int lq_chunk_req = 0;
  float* d_comps = nullptr; float* d_R = nullptr; float* d_gram = nullptr; float* d_coef = nullptr; float* d_w = nullptr;
  int* d_solve_queue = nullptr;   // [1 + chunk]: count, then the samples that need the eigen-solve
  // second slot of the chunk pipeline: chunks alternate between two streams, so one chunk's latency-bound tail (Gram, Cholesky,
  // the eigen-solve of the few singular samples) runs under the other chunk's convolution
  float* d_comps2 = nullptr; float* d_R2 = nullptr; float* d_gram2 = nullptr; int* d_solve_queue2 = nullptr;
  cudaStream_t lq_stream[2] = {nullptr, nullptr};      // per slot: ray-shooting + convolution (normal priority)
  cudaStream_t lq_stream_hi[2] = {nullptr, nullptr};   // per slot: Gram / solve / image tail (highest priority: its few CTAs get
                                                       // the SM slots the other slot's convolution frees, instead of queueing behind it)
  cudaEvent_t lq_fork = nullptr, lq_done[2] = {nullptr, nullptr}, lq_conv[2] = {nullptr, nullptr};
  int lq_pipeline = 1;       // 0 = all chunks on the caller's stream (A/B)
  // Gradient path, single chunk: the eigen-solve of the few samples with a singular Gram matrix (one sequential ~2 ms Jacobi each) runs
  // on lq_tail_stream while the caller's stream takes the other samples through image / adjoint conv / ray-tracing adjoint; the late
  // samples follow in a second, short pass (gl_lstsq_loglike_core).  lq_tail_pending: the caller's stream still has to wait for lq_tail_done.
  cudaStream_t lq_tail_stream = nullptr;
  cudaEvent_t lq_tail_fork = nullptr, lq_tail_done = nullptr;
  int lq_hide_tail = 1;      // 0 = eigen-solve in line (A/B)
  bool lq_tail_pending = false;
  float* d_ll = nullptr;
};

static void gl_free_plan(gl_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  float* fl[] = {p->d_grid_x, p->d_grid_y, p->d_member_factor, p->d_wf, p->d_wb, p->d_obs, p->d_err, p->d_params,
                 p->d_dparams, p->d_derived, p->d_ss, p->d_img, p->d_gimg, p->d_like_part, p->d_gpart, p->d_gsum,
                 p->d_logprior, p->d_fmax, p->d_z, p->d_logp, p->d_chi, p->d_dz, p->d_comps, p->d_R, p->d_gram, p->d_coef, p->d_w, p->d_ll};
  for (float* q : fl) if (q) cudaFree(q);
  if (p->d_ss_mask) cudaFree(p->d_ss_mask);
  if (p->d_mask) cudaFree(p->d_mask);
  if (p->d_leaves) cudaFree(p->d_leaves);
  if (p->d_amp_slot) cudaFree(p->d_amp_slot);
  if (p->d_perm) cudaFree(p->d_perm);
  if (p->d_nan) cudaFree(p->d_nan);
  if (p->d_tape) cudaFree(p->d_tape);
  if (p->d_solve_queue) cudaFree(p->d_solve_queue);
  if (p->d_solve_queue2) cudaFree(p->d_solve_queue2);
  for (float* q : {p->d_comps2, p->d_R2, p->d_gram2}) if (q) cudaFree(q);
  for (int k = 0; k < 2; ++k) {
    if (p->lq_stream[k]) cudaStreamDestroy(p->lq_stream[k]);
    if (p->lq_stream_hi[k]) cudaStreamDestroy(p->lq_stream_hi[k]);
    if (p->lq_done[k]) cudaEventDestroy(p->lq_done[k]);
    if (p->lq_conv[k]) cudaEventDestroy(p->lq_conv[k]);
  }
  if (p->lq_fork) cudaEventDestroy(p->lq_fork);
  if (p->lq_tail_stream) cudaStreamDestroy(p->lq_tail_stream);
  if (p->lq_tail_fork) cudaEventDestroy(p->lq_tail_fork);
  if (p->lq_tail_done) cudaEventDestroy(p->lq_tail_done);
  for (void* q : {(void*)p->d_pos_off, (void*)p->d_pos_x, (void*)p->d_pos_y, (void*)p->d_pos_ex, (void*)p->d_pos_ey, (void*)p->d_pos_ll,
                  (void*)p->d_pos_chi, (void*)p->d_pos_grad}) if (q) cudaFree(q);
  for (cudaEvent_t e : p->tm_ev) cudaEventDestroy(e);
  if (p->d_tables) cudaFree(p->d_tables);
  delete p;
}

template <class T>
static cudaError_t gl_upload(T** dst, const T* src, size_t count) {
  cudaError_t e = cudaMalloc((void**)dst, count * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice);
}

// Kernel instantiations by feature set (gl_math.cuh GLF_*): the first set covering the program is used.
#define GL_FS0 (GLF_EPL | GLF_SHEAR | GLF_SERSIC)
#define GL_FS1 (GLF_EPL | GLF_SHEAR | GLF_SERSIC | GLF_SHAPELETS)
#define GL_FS2 (GLF_NFW | GLF_DPIE | GLF_SHEAR | GLF_SERSIC)
#define GL_FS3 (GLF_ALL)
static const unsigned kFeatSets[] = {GL_FS0, GL_FS1, GL_FS2, GL_FS3};
#define GL_FEAT_DISPATCH(idx, ...)                              \
  switch (idx) {                                                \
    case 0: { constexpr unsigned F = GL_FS0; __VA_ARGS__; } break; \
    case 1: { constexpr unsigned F = GL_FS1; __VA_ARGS__; } break; \
    case 2: { constexpr unsigned F = GL_FS2; __VA_ARGS__; } break; \
    default: { constexpr unsigned F = GL_FS3; __VA_ARGS__; } break; \
  }

#define GL_NSTAGE 7   // unconstrain, prep, raytrace_fwd, conv_fwd, conv_bwd, raytrace_bwd, sample_bwd
#define GL_TM(p, st, k)                                                                         \
  do {                                                                                           \
    if ((p)->tm_slots > 0)                                                                       \
      cudaEventRecord((p)->tm_ev[(size_t)((p)->tm_calls % (p)->tm_slots) * (GL_NSTAGE + 1) + (k)], st); \
  } while (0)

static const int kConvA[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 13, 16, 20, 25, 32};

// choose the CTA tile of the conv kernels: cover `extent` outputs with ntx*RX (nty*RY) per tile,
// at most 256 threads and a bounded smem footprint
static void gl_pick_tiles(int extent, int A, int nph_in, GlConvGeom& g) {
  auto ceil_div = [](int a, int b) { return (a + b - 1) / b; };
  int best_ntx = 1, best_nty = 1; double best_cost = 1e30;
  for (int tiles_x = 1; tiles_x <= 64; ++tiles_x) {
    const int ntx = ceil_div(ceil_div(extent, tiles_x), GLC_RX);
    for (int tiles_y = 1; tiles_y <= 64; ++tiles_y) {
      const int nty = ceil_div(ceil_div(extent, tiles_y), GLC_RY);
      const int threads = ntx * nty;
      if (threads > 256) continue;
      const int in_rows = nty * GLC_RY + A - 1, in_pitch = (ntx * GLC_RX + A - 1 + 3) & ~3;
      const size_t smem = (size_t)(nph_in * in_rows * in_pitch + g.ss * g.ss * A * 2 * glc_ulen(A)) * 4;
      if (smem > 100 * 1024) continue;
      const int warps = ceil_div(threads, 32);
      // cost: issued warp-work (incl. idle lanes / edge waste) + halo traffic
      const double work = (double)tiles_x * tiles_y * warps * 32;
      const double halo = (double)tiles_x * tiles_y * in_rows * in_pitch * nph_in / 169.0;
      const double small = threads < 96 ? 1.15 : 1.0;
      const double cost = (work + halo) * small;
      if (cost < best_cost) { best_cost = cost; best_ntx = ntx; best_nty = nty; g.tiles_x = tiles_x; g.tiles_y = tiles_y; }
    }
  }
  g.ntx = best_ntx; g.nty = best_nty;
  g.tw = g.ntx * GLC_RX; g.th = g.nty * GLC_RY;
  g.tiles_x = ceil_div(extent, g.tw); g.tiles_y = ceil_div(extent, g.th);
  g.in_rows = g.th + A - 1; g.in_pitch = (g.tw + A - 1 + 3) & ~3;
  // Strip loads are LDS.128 by consecutive threads (tx fastest): rows of threads are padded to a
  // multiple of 8 (one quarter-warp = one 128-byte wavefront) when that costs no extra warp, so a
  // quarter-warp never straddles two strip rows and the loads stay conflict-free.
  g.tpr = g.ntx;
  {
    const int padded = (g.ntx + 7) & ~7;
    if (padded * g.nty <= 256 && ceil_div(padded * g.nty, 32) <= ceil_div(g.ntx * g.nty, 32) + 1) g.tpr = padded;
  }
}

// Every kernel of the library receives the device's full opt-in dynamic shared-memory limit ONCE per device, at the first
// gl_plan_create.  The attribute is a permission (occupancy follows from the bytes a launch actually asks for), so no launch
// path touches function attributes afterwards and two plans with different footprints cannot undo each other's setting.
template <class K>
static cudaError_t gl_optin(K kernel, int optin) {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, kernel);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)a.sharedSizeBytes);
}
template <int A>
static cudaError_t gl_optin_conv(int optin) {
  cudaError_t e;
  if ((e = gl_optin(k_conv_fwd_tma<A>, optin)) != cudaSuccess) return e;
  if ((e = gl_optin(k_conv_fwd<A>, optin)) != cudaSuccess) return e;
  if ((e = gl_optin(k_conv_bwd_tma<A>, optin)) != cudaSuccess) return e;
  return gl_optin(k_conv_bwd<A>, optin);
}
template <unsigned F>
static cudaError_t gl_optin_feat(int optin) {
  cudaError_t e;
  if ((e = gl_optin(k_raytrace_fwd<4, F>, optin)) != cudaSuccess) return e;
  if ((e = gl_optin(k_nan_cotangent_mask<4, F>, optin)) != cudaSuccess) return e;
  if ((e = gl_optin(k_raytrace_bwd<4, F>, optin)) != cudaSuccess) return e;
  if ((e = gl_optin(k_raytrace_comps<4, F>, optin)) != cudaSuccess) return e;
  return gl_optin(k_points<F>, optin);
}
static int gl_device_init(int device) {
  static bool done[64] = {false};
  if (device >= 0 && device < 64 && done[device]) return 0;
  int optin = 0;
  GL_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
#define GL_OPT(call) GL_CUDA(call)
  GL_OPT(gl_optin_feat<GL_FS0>(optin)); GL_OPT(gl_optin_feat<GL_FS1>(optin));
  GL_OPT(gl_optin_feat<GL_FS2>(optin)); GL_OPT(gl_optin_feat<GL_FS3>(optin));
  GL_OPT(gl_optin(k_raytrace_bwd<4, GL_FS3, true>, optin));
  GL_OPT(gl_optin(k_raytrace_fwd_p<4, GL_FS0, true>, optin)); GL_OPT(gl_optin(k_raytrace_fwd_p<4, GL_FS0>, optin));
  GL_OPT(gl_optin(k_raytrace_fwd_p<4, GL_FS2>, optin)); GL_OPT(gl_optin(k_raytrace_fwd_p<4, GL_FS2, false, true>, optin));
  GL_OPT(gl_optin(k_raytrace_bwd_p<4, GL_FS2, true, false, true>, optin)); GL_OPT(gl_optin(k_raytrace_bwd_p<4, GL_FS2, false, false, true>, optin));
  GL_OPT(gl_optin(k_raytrace_bwd_p<4, GL_FS0, true, true>, optin));
  GL_OPT(gl_optin(k_raytrace_bwd_p<4, GL_FS0, true>, optin)); GL_OPT(gl_optin(k_raytrace_bwd_p<4, GL_FS0, false>, optin));
  GL_OPT(gl_optin(k_raytrace_bwd_p<4, GL_FS2, true>, optin)); GL_OPT(gl_optin(k_raytrace_bwd_p<4, GL_FS2, false>, optin));
  GL_OPT(gl_optin_conv<1>(optin)); GL_OPT(gl_optin_conv<2>(optin)); GL_OPT(gl_optin_conv<3>(optin)); GL_OPT(gl_optin_conv<4>(optin));
  GL_OPT(gl_optin_conv<5>(optin)); GL_OPT(gl_optin_conv<6>(optin)); GL_OPT(gl_optin_conv<7>(optin)); GL_OPT(gl_optin_conv<8>(optin));
  GL_OPT(gl_optin_conv<9>(optin)); GL_OPT(gl_optin_conv<11>(optin)); GL_OPT(gl_optin_conv<13>(optin)); GL_OPT(gl_optin_conv<16>(optin));
  GL_OPT(gl_optin_conv<20>(optin)); GL_OPT(gl_optin_conv<25>(optin)); GL_OPT(gl_optin_conv<32>(optin));
  GL_OPT(gl_optin(k_conv_fwd_tma<GLC_A_HOT, GLC_NF_HOT>, optin)); GL_OPT(gl_optin(k_conv_bwd_tma<GLC_A_HOT, GLC_NF_HOT>, optin));
  GL_OPT(gl_optin(k_positions, optin)); GL_OPT(gl_optin(k_hessian, optin)); GL_OPT(gl_optin(k_pinv_solve, optin));
  GL_OPT(gl_optin(k_gram, optin));
  GL_OPT(gl_gram_tc_init());
#undef GL_OPT
  if (device >= 0 && device < 64) done[device] = true;
  return 0;
}

extern "C" {

static int gl_lstsq_reserve(gl_plan* p, int chunk);

const char* gl_last_error(void) { return g_last_error.c_str(); }
int32_t gl_abi_version(void) { return GL_ABI_VERSION; }

// Red zones of every live device allocation of the library (GL_GUARD=1): returns the number of live guarded allocations,
// or -1 with gl_last_error() naming the first corrupted one; 0 when guarding is off.  Synchronises the device.
int32_t gl_guard_check(void) {
  if (!gl_guard_on()) return 0;
  if (cudaDeviceSynchronize() != cudaSuccess) { gl_fail(std::string("gl_guard_check: ") + cudaGetErrorString(cudaGetLastError())); return -1; }
  std::lock_guard<std::mutex> lk(g_guard_mu);
  if (!g_guard_sticky.empty()) { gl_fail(g_guard_sticky); return -1; }
  for (auto& kv : g_guard_live) {
    const std::string e = gl_guard_verify((const char*)kv.first, kv.second);
    if (!e.empty()) { gl_fail(e); return -1; }
  }
  return (int32_t)g_guard_live.size();
}

// Positive control of the red zones: a kernel writes one float past the end of (before the start of) a guarded buffer and
// gl_guard_check() must report it.  Returns 0 when both overruns are detected and a clean buffer passes, 1 otherwise
// (2 when guarding is off).
__global__ void k_guard_selftest(float* p, long idx) { p[idx] = 1.0f; }
int32_t gl_guard_selftest(void) {
  if (!gl_guard_on()) return 2;
  const std::string keep = g_last_error;
  int bad = 0;
  for (int side = 0; side < 3; ++side) {
    float* buf = nullptr;
    if (cudaMalloc(&buf, 1000 * sizeof(float)) != cudaSuccess) return 1;
    k_guard_selftest<<<1, 1>>>(buf, side == 0 ? 0 : side == 1 ? 1000 : -1);
    const int32_t r = gl_guard_check();
    if ((side == 0) != (r >= 0)) ++bad;     // in-bounds write: clean; either overrun: reported
    cudaFree(buf);
    { std::lock_guard<std::mutex> lk(g_guard_mu); g_guard_sticky.clear(); }   // the deliberate overrun is not a finding
  }
  g_last_error = keep;
  return bad ? 1 : 0;
}
int64_t gl_launch_count(void) { return g_launch_count; }
int32_t gl_plan_depth(const gl_plan* plan) { return plan ? plan->prog.depth : 0; }

void gl_plan_destroy(gl_plan* plan) { gl_free_plan(plan); }

int gl_plan_create(const gl_model_desc* model, const gl_sim_config* sim, int32_t bs, int32_t device, gl_plan** out) {
  if (!out) return gl_fail("gl_plan_create: out is NULL");
  *out = nullptr;
  if (!model || !sim) return gl_fail("gl_plan_create: NULL descriptor");
  if (bs <= 0) return gl_fail("gl_plan_create: bs must be positive");
  if (sim->num_pix <= 0 || sim->supersample <= 0) return gl_fail("gl_plan_create: bad num_pix / supersample");
  if (!sim->grid_x || !sim->grid_y) return gl_fail("gl_plan_create: grid_x / grid_y are required");
  if (sim->psf && (sim->psf_n <= 0 || sim->psf_n > 255)) return gl_fail("gl_plan_create: bad psf_n");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return gl_fail("gl_plan_create: no CUDA device available (this library has no CPU path)");
  if (device < 0 || device >= ndev) return gl_fail("gl_plan_create: device index out of range");
  GL_CUDA(cudaSetDevice(device));
  if (gl_device_init(device)) return 1;

  GlBuilt built;
  std::string err = gl_build_program(model, built);
  if (!err.empty()) return gl_fail("gl_plan_create: " + err);

  gl_plan* p = new gl_plan();
  p->device = device; p->bs = bs; p->prog = built.prog;
  p->prog.epl_tol = 1e-9f;   // see "epl_tol_exp10" in include/gigalens_b200.h
  p->n = sim->num_pix; p->ss = sim->supersample; p->hs = p->n * p->ss; p->npix = p->hs * p->hs;
  p->conversion_factor = sim->conversion_factor;
  for (int i = 0; i < p->prog.n_lens; ++i) if (p->prog.prof[i].type == GLT_EPL) p->has_epl = true;
  {
    unsigned need = 0;
    for (int i = 0; i < p->prog.n_prof; ++i) need |= gl_feature_of(p->prog.prof[i].type);
    for (int k = 3; k >= 0; --k) if ((kFeatSets[k] & need) == need) p->feat_idx = k;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  p->sm_count = prop.multiProcessorCount;

#define GL_TRY(call)                                                                               \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) { gl_free_plan(p); return gl_fail(std::string(#call) + ": " + cudaGetErrorString(_e)); } \
  } while (0)

  // The supersampled buffers are PHASE-MAJOR (gl_conv.cuh): element ((py*ss+px)*n + ri)*n + cj holds
  // ss pixel (ss*ri + py, ss*cj + px).  The ray-tracing kernels just walk that index space with
  // permuted coordinate tables, so the permutation costs nothing at run time.
  std::vector<int> perm((size_t)p->npix);   // phase-major index -> natural row-major index
  {
    const int ss = p->ss, n = p->n, hs = p->hs;
    for (int py = 0; py < ss; ++py)
      for (int px = 0; px < ss; ++px)
        for (int ri = 0; ri < n; ++ri)
          for (int cj = 0; cj < n; ++cj)
            perm[(((size_t)py * ss + px) * n + ri) * n + cj] = (ss * ri + py) * hs + (ss * cj + px);
    std::vector<float> gx((size_t)p->npix), gy((size_t)p->npix);
    for (size_t k = 0; k < perm.size(); ++k) { gx[k] = sim->grid_x[perm[k]]; gy[k] = sim->grid_y[perm[k]]; }
    GL_TRY(gl_upload(&p->d_grid_x, gx.data(), gx.size()));
    GL_TRY(gl_upload(&p->d_grid_y, gy.data(), gy.size()));
    GL_TRY(gl_upload(&p->d_perm, perm.data(), perm.size()));
  }
  if (sim->mask) {
    std::vector<unsigned char> m((size_t)p->n * p->n), ms((size_t)p->npix);
    for (int i = 0; i < p->n * p->n; ++i) m[i] = sim->mask[i] ? 1 : 0;
    for (size_t k = 0; k < perm.size(); ++k) {
      const int i = perm[k] / p->hs, j = perm[k] % p->hs;
      ms[k] = m[(size_t)(i / p->ss) * p->n + j / p->ss];
    }
    GL_TRY(gl_upload(&p->d_mask, m.data(), m.size()));
    GL_TRY(gl_upload(&p->d_ss_mask, ms.data(), ms.size()));
    size_t c = 0; for (unsigned char v : m) c += v;
    p->n_mask_used = (float)c;   // count_nonzero(img_region)  (tf/model.py:100)
  } else {
    p->n_mask_used = (float)((size_t)p->n * p->n);
  }
  if (!built.member_factor.empty()) GL_TRY(gl_upload(&p->d_member_factor, built.member_factor.data(), built.member_factor.size()));
  if (!built.amp_slot.empty()) GL_TRY(gl_upload(&p->d_amp_slot, built.amp_slot.data(), built.amp_slot.size()));
  if (!built.tables.empty()) {
    GL_TRY(gl_upload(&p->d_tables, built.tables.data(), built.tables.size()));
    for (int i = 0; i < p->prog.n_prof; ++i)
      if (built.table_off[i] >= 0) p->prog.prof[i].table = p->d_tables + built.table_off[i];
  }

  // folded kernel Keff = box_ss (*) K, polyphase taps W[py][px][a][b] = Keff[ss a + py][ss b + px] / ss^2
  {
    const int K = sim->psf ? sim->psf_n : 1, ss = p->ss;
    const int E = K + ss - 1;
    std::vector<double> keff((size_t)E * E, 0.0);
    for (int i = 0; i < K; ++i)
      for (int j = 0; j < K; ++j) {
        const double kv = sim->psf ? (double)sim->psf[(size_t)i * K + j] : 1.0;
        for (int sy = 0; sy < ss; ++sy)
          for (int sx = 0; sx < ss; ++sx) keff[(size_t)(i + sy) * E + (j + sx)] += kv / (double)(ss * ss);
      }
    const int Aexact = (E + ss - 1) / ss;
    int A = -1;
    for (int a : kConvA) if (a >= Aexact) { A = a; break; }
    if (A < 0) { gl_free_plan(p); return gl_fail("gl_plan_create: PSF too large for the conv kernels (taps per phase > 32)"); }
    p->A = A; p->pad = (K - 1) / 2;
    const int wpitch = (A + 3) & ~3, nph = ss * ss;
    // packed tap tables for corr_rows2 (gl_conv.cuh): per phase and tap column b two copies of the
    // reversed, zero-padded column  U[j] = w[A-1-(j-(RY-1))][b]
    const int UL = glc_ulen(A), UTAB = A * 2 * UL;
    std::vector<float> wf((size_t)nph * UTAB, 0.f), wb((size_t)nph * UTAB, 0.f);
    auto put = [&](std::vector<float>& tab, int ph, int a, int b2, float w) {
      const int j = (A - 1 - a) + (GLC_RY - 1);
      float* base = tab.data() + (size_t)ph * UTAB + (size_t)b2 * 2 * UL;
      base[j] = w;                       // copy 0: U[j]
      if (j >= 1) base[UL + j - 1] = w;  // copy 1: U shifted left by one
    };
    for (int py = 0; py < ss; ++py)
      for (int px = 0; px < ss; ++px)
        for (int a = 0; a < A; ++a)
          for (int b2 = 0; b2 < A; ++b2) {
            const int u = ss * a + py, v = ss * b2 + px;
            const float w = (u < E && v < E) ? (float)keff[(size_t)u * E + v] : 0.f;
            const int ph = py * ss + px;
            put(wf, ph, a, b2, w);
            put(wb, ph, A - 1 - a, A - 1 - b2, w);   // adjoint: flipped taps
          }
    GL_TRY(gl_upload(&p->d_wf, wf.data(), wf.size()));
    GL_TRY(gl_upload(&p->d_wb, wb.data(), wb.size()));
    if (A == GLC_A_HOT && ss == GLC_SS_HOT && (int)wf.size() == GLC_NF_HOT) {
      memcpy(p->ct_f.w, wf.data(), wf.size() * sizeof(float));
      memcpy(p->ct_b.w, wb.data(), wb.size() * sizeof(float));
      p->conv_const_ok = true;
    }
    GlConvGeom g{};
    g.n = p->n; g.hs = p->hs; g.ss = ss; g.A = A; g.pad = p->pad; g.wpitch = wpitch;
    p->gf = g; p->gb = g;
    gl_pick_tiles(p->n, A, 2, p->gf);   // two staged phase tiles (double buffer)
    // adjoint runs over padded-phase coordinates r in [pad/ss, (hs-1+pad)/ss]
    p->bwd_rc_start = p->pad / ss;
    const int nr = (p->hs - 1 + p->pad) / ss - p->bwd_rc_start + 1;
    gl_pick_tiles(nr, A, 1, p->gb);
    p->gb.rc0 = p->bwd_rc_start;
    p->conv_threads_f = ((p->gf.tpr * p->gf.nty + 31) / 32) * 32;
    p->conv_threads_b = ((p->gb.tpr * p->gb.nty + 31) / 32) * 32;
    p->gf.tma_pitch = p->gf.in_pitch + 4;
    p->gf.phase_stride = (p->gf.in_rows * p->gf.tma_pitch + 31) & ~31;   // TMA destinations are 128-byte aligned
    p->conv_tma_ok = (p->n % 4) == 0 && p->gf.tma_pitch <= 256 && p->gf.in_rows <= 256;
    for (int px = 0; px < ss; ++px) if (((px + p->pad) / ss) & 1) p->conv_tma_ok = false;   // strips must stay 8-byte aligned
    p->smem_cf_tma = (size_t)(2 * p->gf.phase_stride + nph * UTAB) * sizeof(float);
    p->smem_cf_tma_c = (size_t)(2 * p->gf.phase_stride) * sizeof(float);
    p->smem_cf = (size_t)(2 * p->gf.in_rows * p->gf.in_pitch + nph * UTAB) * sizeof(float);
    p->smem_cb = (size_t)(p->gb.in_rows * p->gb.in_pitch + nph * UTAB) * sizeof(float);
    p->gb.tma_pitch = p->gb.in_pitch + 4;
    p->gb.phase_stride = (p->gb.in_rows * p->gb.tma_pitch + 31) & ~31;
    p->gb.out_stride = (p->gb.tw * p->gb.th + 31) & ~31;
    p->gb.band_rows = 0; p->gb.band_stride = 0;
    if (32 % p->gb.tpr == 0 && p->gb.nty % (32 / p->gb.tpr) == 0) {   // a warp = complete thread rows: per-warp band stores
      p->gb.band_rows = GLC_RY * (32 / p->gb.tpr);
      p->gb.band_stride = (p->gb.band_rows * p->gb.tw + 31) & ~31;
      p->gb.out_stride = p->gb.band_stride * (p->conv_threads_b / 32);
    }
    p->smem_cb_tma = (size_t)(p->gb.phase_stride + p->gb.out_stride + nph * UTAB) * sizeof(float);
    p->smem_cb_tma_c = (size_t)(p->gb.phase_stride + p->gb.out_stride) * sizeof(float);
    p->conv_tma_b_ok = (p->n % 4) == 0 && p->gb.tma_pitch <= 256 && p->gb.in_rows <= 256 && p->gb.th <= 256 && (p->gb.rc0 & 1) == 0;
    for (int px = 0; px < ss; ++px) {   // output tile origins rc0 + dx(px) + k * tw on 16-byte boundaries
      const int fx = px - p->pad;
      const int dx = (fx >= 0) ? fx / ss : -((-fx + ss - 1) / ss);
      if ((p->gb.rc0 + dx) & 3) p->conv_tma_b_ok = false;
    }
  }

  // chunks per sample for the ray-tracing kernels: enough CTAs for ~4 waves, at most one per pixel batch
  {
    const int per_batch = GLK_THREADS * 4;
    const int nbatch = (p->npix + per_batch - 1) / per_batch;
    int chunks = (p->sm_count * 8 * 4 + bs - 1) / bs;
    if (chunks > nbatch) chunks = nbatch;
    if (chunks < 1) chunks = 1;
    p->chunks = chunks;
  }
  const size_t P_ = (size_t)(p->prog.n_params > 0 ? p->prog.n_params : 1);
  const int ntile = p->gf.tiles_x * p->gf.tiles_y;
  GL_TRY(cudaMalloc((void**)&p->d_params, P_ * bs * sizeof(float)));
  GL_TRY(cudaMalloc((void**)&p->d_dparams, P_ * bs * sizeof(float)));
  GL_TRY(cudaMalloc((void**)&p->d_derived, (size_t)bs * p->prog.der_total * sizeof(float)));
  GL_TRY(cudaMalloc((void**)&p->d_ss, (size_t)bs * p->npix * sizeof(float)));
  GL_TRY(cudaMalloc((void**)&p->d_img, (size_t)bs * p->n * p->n * sizeof(float)));
  GL_TRY(cudaMalloc((void**)&p->d_gimg, (size_t)bs * p->n * p->n * sizeof(float)));
  GL_TRY(cudaMalloc((void**)&p->d_like_part, (size_t)bs * ntile * 2 * sizeof(float)));
  GL_TRY(cudaMalloc((void**)&p->d_gpart, (size_t)bs * p->chunks * (p->prog.g_total + 1) * sizeof(float)));
  GL_TRY(cudaMalloc((void**)&p->d_gsum, (size_t)bs * (p->prog.g_total + 1) * sizeof(float)));
  GL_TRY(cudaMalloc((void**)&p->d_logprior, (size_t)bs * sizeof(float)));
  GL_TRY(cudaMalloc((void**)&p->d_nan, (size_t)bs * sizeof(int)));
  GL_TRY(cudaMalloc((void**)&p->d_fmax, GL_MAX_PROF * sizeof(float)));
  GL_TRY(cudaMemset(p->d_logprior, 0, (size_t)bs * sizeof(float)));
  GL_TRY(cudaMemset(p->d_nan, 0, (size_t)bs * sizeof(int)));
#undef GL_TRY
  // tape of the forward-mode group: 32 bytes per ss pixel and sample, if it fits 60 % of the memory that is free now
  if (p->prog.has_fwdmode && p->feat_idx == 2 && (p->npix % 2) == 0) {
    const size_t need = (size_t)bs * 8 * p->npix * sizeof(float);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && need <= free_b / 5 * 3) {
      if (cudaMalloc((void**)&p->d_tape, need) != cudaSuccess) { p->d_tape = nullptr; cudaGetLastError(); }
    }
  }
  {   // models built for lstsq_simulate (some light profile has use_lstsq) get their workspace now; others on request
    bool wants = false;
    for (int i = p->prog.n_lens; i < p->prog.n_prof; ++i) wants = wants || (p->prog.prof[i].flags & GL_FLAG_USE_LSTSQ);
    if (wants && gl_lstsq_reserve(p, 0)) { gl_free_plan(p); return 1; }
  }
  *out = p;
  return 0;
}

int gl_plan_set_option(gl_plan* p, const char* name, int32_t value) {
  if (!p || !name) return gl_fail("gl_plan_set_option: NULL argument");
  if (!strcmp(name, "epl_batch_max")) { p->epl_batch_max = value; return 0; }
  if (!strcmp(name, "row_flush")) { p->row_flush = value; return 0; }
  if (!strcmp(name, "conv_tma")) { p->conv_tma = value; return 0; }
  if (!strcmp(name, "conv_const_taps")) { p->conv_const = value; return 0; }
  if (!strcmp(name, "tape")) { p->use_tape = value; return 0; }
  if (!strcmp(name, "lstsq_pipeline")) { p->lq_pipeline = value; return 0; }
  if (!strcmp(name, "lstsq_hide_tail")) { p->lq_hide_tail = value; return 0; }
  if (!strcmp(name, "straight_line")) { p->straight_line = value; return 0; }
  if (!strcmp(name, "epl_tol_exp10")) {   // EPL series terms below 10^-value are dropped (12 = the reference's constant, epl.py:37)
    if (value < 6 || value > 30) return gl_fail("gl_plan_set_option: epl_tol_exp10 must be in [6, 30]");
    p->prog.epl_tol = powf(10.f, -(float)value); return 0;
  }
  if (!strcmp(name, "lstsq")) { p->lstsq = value; return 0; }
  if (!strcmp(name, "no_deflection")) { p->no_deflection = value; return 0; }
  if (!strcmp(name, "components")) {   // gl_simulate only: 1 = lens light, 2 = source light, 3 = both
    if (value < 1 || value > 3) return gl_fail("gl_plan_set_option: components must be 1, 2 or 3");
    p->prog.comp_mask = value; return 0;
  }
  if (!strcmp(name, "timing")) {   // value = number of calls to keep event sets for (0 = off)
    for (cudaEvent_t e : p->tm_ev) cudaEventDestroy(e);
    p->tm_ev.clear(); p->tm_slots = 0; p->tm_calls = 0;
    if (value > 0) {
      p->tm_ev.resize((size_t)value * (GL_NSTAGE + 1));
      for (auto& e : p->tm_ev) if (cudaEventCreate(&e) != cudaSuccess) return gl_fail("gl_plan_set_option: cudaEventCreate failed");
      p->tm_slots = value;
    }
    return 0;
  }
  if (!strcmp(name, "include_pixels")) { p->include_pixels = value != 0; return 0; }       // ForwardProbModel(include_pixels=...)
  if (!strcmp(name, "include_positions")) {                                               // ForwardProbModel(include_positions=...)
    if (value && p->pos_npts == 0) return gl_fail("gl_plan_set_option: include_positions needs gl_plan_set_positions first");
    p->include_positions = value != 0; return 0;
  }
  if (!strcmp(name, "gram_tc")) { p->gram_tc = value != 0; return 0; }
  if (!strcmp(name, "gram_ksplit")) { p->gram_ksplit = value >= 2 ? 2 : 1; return 0; }
  if (!strcmp(name, "packed_math")) { p->use_packed = value; return 0; }   // 0: scalar-lane kernels (A/B testing)
  if (!strcmp(name, "lstsq_chunk")) {   // samples per pass of the lstsq component stack (0 = size by memory budget): re-reserves
    if (value < 0) return gl_fail("gl_plan_set_option: lstsq_chunk must be >= 0");
    return gl_lstsq_reserve(p, value);
  }
  return gl_fail(std::string("gl_plan_set_option: unknown option ") + name);
}

int gl_plan_set_likelihood(gl_plan* p, const gl_like_config* like) {
  if (!p || !like || !like->observed) return gl_fail("gl_plan_set_likelihood: NULL argument");
  GL_CUDA(cudaSetDevice(p->device));
  const size_t nn = (size_t)p->n * p->n;
  if (p->d_obs) { cudaFree(p->d_obs); p->d_obs = nullptr; }
  if (p->d_err) { cudaFree(p->d_err); p->d_err = nullptr; }
  GL_CUDA(gl_upload(&p->d_obs, like->observed, nn));
  if (like->error_map) GL_CUDA(gl_upload(&p->d_err, like->error_map, nn));
  else if (!(like->exp_time != 0.f)) return gl_fail("gl_plan_set_likelihood: exp_time must be non-zero without an error map");
  p->bg2 = like->background_rms * like->background_rms;
  p->inv_exp = like->error_map ? 0.f : 1.f / like->exp_time;
  p->n_pix_used = p->n_mask_used;   // count_nonzero(img_region)  (tf/model.py:100), counted at plan creation
  if (p->d_w) { cudaFree(p->d_w); p->d_w = nullptr; }
  if (like->error_map) {
    std::vector<float> w(nn);
    for (size_t i = 0; i < nn; ++i) w[i] = 1.f / like->error_map[i];   // W = 1 / err_map (tf/simulator.py:232)
    GL_CUDA(gl_upload(&p->d_w, w.data(), nn));
  }
  p->has_like = true;
  return 0;
}


/* Centroids of the multiply-imaged sources: ForwardProbModel(centroids_x, centroids_y, centroids_errors_x,
 * centroids_errors_y) (tf/model.py:69-74); host arrays, images of all systems concatenated. */
int gl_plan_set_positions(gl_plan* p, int32_t n_systems, const int32_t* n_images, const float* x, const float* y,
                          const float* err_x, const float* err_y) {
  if (!p || n_systems <= 0 || !n_images || !x || !y || !err_x || !err_y) return gl_fail("gl_plan_set_positions: bad argument");
  GL_CUDA(cudaSetDevice(p->device));
  std::vector<int> off(n_systems + 1, 0);
  for (int s = 0; s < n_systems; ++s) {
    if (n_images[s] < 1 || n_images[s] > 64) return gl_fail("gl_plan_set_positions: a system needs 1..64 images");
    off[s + 1] = off[s] + n_images[s];
  }
  const int npts = off[n_systems];
  for (int i = 0; i < npts; ++i)
    if (!(err_x[i] != 0.f) || !(err_y[i] != 0.f)) return gl_fail("gl_plan_set_positions: centroid errors must be non-zero");
  for (void** q : {(void**)&p->d_pos_off, (void**)&p->d_pos_x, (void**)&p->d_pos_y, (void**)&p->d_pos_ex, (void**)&p->d_pos_ey})
    if (*q) { cudaFree(*q); *q = nullptr; }
  GL_CUDA(gl_upload(&p->d_pos_off, off.data(), off.size()));
  GL_CUDA(gl_upload(&p->d_pos_x, x, (size_t)npts));
  GL_CUDA(gl_upload(&p->d_pos_y, y, (size_t)npts));
  GL_CUDA(gl_upload(&p->d_pos_ex, err_x, (size_t)npts));
  GL_CUDA(gl_upload(&p->d_pos_ey, err_y, (size_t)npts));
  if (!p->d_pos_ll) {
    GL_CUDA(cudaMalloc((void**)&p->d_pos_ll, (size_t)p->bs * sizeof(float)));
    GL_CUDA(cudaMalloc((void**)&p->d_pos_chi, (size_t)p->bs * sizeof(float)));
    GL_CUDA(cudaMalloc((void**)&p->d_pos_grad, (size_t)p->bs * (p->prog.n_params > 0 ? p->prog.n_params : 1) * sizeof(float)));
  }
  p->pos_npts = npts; p->pos_nsys = n_systems;
  p->pos_n_position = 2.f * (float)npts;   // n_position = 2 * size(concat(centroids_x))  (tf/model.py:74)
  p->include_positions = 1;
  return 0;
}

int gl_plan_set_prior(gl_plan* p, const gl_prior_leaf* leaves, int32_t n_leaves) {
  if (!p || !leaves || n_leaves <= 0) return gl_fail("gl_plan_set_prior: bad arguments");
  GL_CUDA(cudaSetDevice(p->device));
  std::vector<GlLeaf> L(n_leaves);
  for (int k = 0; k < n_leaves; ++k) {
    const gl_prior_leaf& s = leaves[k];
    if (s.slot >= p->prog.n_params) return gl_fail("gl_plan_set_prior: leaf slot out of range");
    if (s.dist < GL_DIST_NORMAL || s.dist > GL_DIST_TRUNCNORMAL) return gl_fail("gl_plan_set_prior: unknown distribution");
    L[k].dist = s.dist; L[k].slot = s.slot; L[k].a = s.a; L[k].b = s.b; L[k].low = s.low; L[k].high = s.high;
    L[k].log_norm = 0.f;
    if (s.dist == GL_DIST_TRUNCNORMAL) {
      const double al = ((double)s.low - s.a) / s.b, be = ((double)s.high - s.a) / s.b;
      const double Z = 0.5 * (std::erf(be / std::sqrt(2.0)) - std::erf(al / std::sqrt(2.0)));
      L[k].log_norm = (float)std::log(Z);
    }
  }
  if (p->d_leaves) { cudaFree(p->d_leaves); p->d_leaves = nullptr; }
  GL_CUDA(gl_upload(&p->d_leaves, L.data(), L.size()));
  p->d = n_leaves;
  for (float** q : {&p->d_z, &p->d_logp, &p->d_chi, &p->d_dz}) if (*q) { cudaFree(*q); *q = nullptr; }
  GL_CUDA(cudaMalloc((void**)&p->d_z, (size_t)p->bs * p->d * sizeof(float)));
  GL_CUDA(cudaMalloc((void**)&p->d_dz, (size_t)p->bs * p->d * sizeof(float)));
  GL_CUDA(cudaMalloc((void**)&p->d_logp, (size_t)p->bs * sizeof(float)));
  GL_CUDA(cudaMalloc((void**)&p->d_chi, (size_t)p->bs * sizeof(float)));
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------
static int gl_run_prep(gl_plan* p, const float* params, cudaStream_t st) {
  const int tb = 32, gb = (p->bs + tb - 1) / tb;   // one warp per CTA: a batch of a few thousand samples spreads over all SMs
  const float* fmax = nullptr;
  if (p->has_epl && p->epl_batch_max) {
    GL_CUDA(cudaMemsetAsync(p->d_fmax, 0, GL_MAX_PROF * sizeof(float), st));
    k_epl_fmax<<<gb, tb, 0, st>>>(p->prog, p->bs, params, p->d_member_factor, (int*)p->d_fmax);
    GL_LAUNCH_CHECK("k_epl_fmax");
    fmax = p->d_fmax;
  }
  k_prep<<<gb, tb, 0, st>>>(p->prog, p->bs, params, p->d_member_factor, p->d_amp_slot, fmax, p->d_derived);
  GL_LAUNCH_CHECK("k_prep");
  return 0;
}

// stats_positions of every sample (one warp per sample, FP64): log-like, chi2/n_position and, when
// dparams != null, d(log-like)/d(params) [P][bs].
static int gl_run_positions(gl_plan* p, const float* params, float* loglike, float* red_chi2, float* dparams, cudaStream_t st) {
  if (p->pos_npts == 0) return gl_fail("image-position likelihood requested but gl_plan_set_positions was never called");
  const size_t smem = sizeof(double) * ((size_t)p->prog.der_total + 12 * (size_t)p->pos_npts + 2 * (size_t)p->pos_nsys +
                                        (size_t)(GLP_THREADS + 1) * p->prog.g_total);
  if (smem > 200 * 1024) return gl_fail("gl_run_positions: model too large for the positions kernel's shared memory");
  k_positions<<<p->bs, GLP_THREADS, smem, st>>>(p->prog, p->bs, params, p->d_member_factor, p->d_amp_slot, p->pos_npts, p->pos_nsys,
                                                p->d_pos_off, p->d_pos_x, p->d_pos_y, p->d_pos_ex, p->d_pos_ey, p->pos_n_position,
                                                loglike, red_chi2, dparams);
  GL_LAUNCH_CHECK("k_positions");
  return 0;
}

static int gl_run_raytrace_fwd(gl_plan* p, float* ss_out, int no_deflection, cudaStream_t st, bool want_tape = false) {
  dim3 grid(p->chunks, p->bs);
  GL_CUDA(cudaMemsetAsync(p->d_nan, 0, (size_t)p->bs * sizeof(int), st));
  const size_t smem = (size_t)p->prog.der_total * sizeof(float);
  p->tape_valid = false;
  if (want_tape && p->d_tape && p->use_tape && p->use_packed && p->feat_idx == 2 && !no_deflection && p->prog.comp_mask == 3) {
    k_raytrace_fwd_p<4, GL_FS2, false, true><<<grid, GLK_THREADS, smem, st>>>(p->prog, p->npix, p->d_grid_x, p->d_grid_y, p->d_ss_mask,
                                                                              p->d_derived, 0, ss_out, p->d_nan, p->d_tape);
    GL_LAUNCH_CHECK("k_raytrace_fwd_p<tape>");
    p->tape_valid = true;
    return 0;
  }
  if ((p->feat_idx == 0 || p->feat_idx == 2) && (p->npix % 2) == 0 && p->use_packed) {   // two pixels per lane slot (FFMA2)
    if (p->feat_idx == 0 && p->straight_line && !no_deflection && gl_is_benchmark_shape(p->prog)) {
      k_raytrace_fwd_p<4, GL_FS0, true><<<grid, GLK_THREADS, smem, st>>>(p->prog, p->npix, p->d_grid_x, p->d_grid_y, p->d_ss_mask,
                                                                         p->d_derived, no_deflection, ss_out, p->d_nan, nullptr);
    } else if (p->feat_idx == 0) {
      k_raytrace_fwd_p<4, GL_FS0><<<grid, GLK_THREADS, smem, st>>>(p->prog, p->npix, p->d_grid_x, p->d_grid_y, p->d_ss_mask,
                                                                   p->d_derived, no_deflection, ss_out, p->d_nan, nullptr);
    } else {
      k_raytrace_fwd_p<4, GL_FS2><<<grid, GLK_THREADS, smem, st>>>(p->prog, p->npix, p->d_grid_x, p->d_grid_y, p->d_ss_mask,
                                                                   p->d_derived, no_deflection, ss_out, p->d_nan, nullptr);
    }
    GL_LAUNCH_CHECK("k_raytrace_fwd_p");
    return 0;
  }
  GL_FEAT_DISPATCH(p->feat_idx, {
    k_raytrace_fwd<4, F><<<grid, GLK_THREADS, smem, st>>>(p->prog, p->npix, p->d_grid_x, p->d_grid_y, p->d_ss_mask, p->d_derived,
                                                          no_deflection, ss_out, p->d_nan);
  })
  GL_LAUNCH_CHECK("k_raytrace_fwd");
  return 0;
}

// mode 0: forward-model path (every sample; NaN-scrubbed pixels handled by k_nan_cotangent_mask).  lstsq path (flags in d_nan, see
// k_raytrace_bwd): mode 1 = every sample, 2 = the samples solved by the Cholesky path, 3 = the samples solved by the eigen-solve;
// in each the samples with scrubbed components go to the SCRUB instance.
static int gl_run_raytrace_bwd(gl_plan* p, float* gss, int no_deflection, cudaStream_t st, int mode = 0) {
  dim3 grid(p->chunks, p->bs);
  const bool per_component = mode != 0;
  const int sel_mask = mode == 0 ? 0 : (mode == 1 ? 1 : 3);
  const int sel_want = mode == 3 ? 2 : 0;              // main kernel: no NaN flag (and the wanted eigen-solve flag)
  if (per_component) {
    const size_t smem_s = (size_t)gl_bwd_smem_floats(p->prog, 4) * sizeof(float);
    k_raytrace_bwd<4, GL_FS3, true><<<grid, GLK_THREADS, smem_s, st>>>(p->prog, p->npix, p->d_grid_x, p->d_grid_y, p->d_ss_mask, p->d_derived,
                                                                       no_deflection, gss, p->d_gpart, p->d_nan, sel_mask, sel_want | 1);
    GL_LAUNCH_CHECK("k_raytrace_bwd<scrub>");
  } else {   // NaN-scrubbed pixels pass no gradient: exits immediately for samples without any (nan_count from the forward pass)
    const size_t smem_m = (size_t)p->prog.der_total * sizeof(float);
    GL_FEAT_DISPATCH(p->feat_idx, {
      k_nan_cotangent_mask<4, F><<<dim3(1, p->bs < 592 ? p->bs : 592), GLK_THREADS, smem_m, st>>>(p->prog, p->npix, p->d_grid_x, p->d_grid_y,
                                                                                         p->d_derived, no_deflection, gss, p->d_nan, p->bs);
    })
    GL_LAUNCH_CHECK("k_nan_cotangent_mask");
  }
  size_t smem = (size_t)gl_bwd_smem_floats(p->prog, 4) * sizeof(float);
  if ((p->feat_idx == 0 || p->feat_idx == 2) && (p->npix % 2) == 0 && p->use_packed) {
    // staged flush (DevFlushStage) when two CTAs of it still fit one SM
    const size_t smem_rows = (size_t)gl_bwd_p_smem_floats(p->prog, 4, true) * sizeof(float);
    smem = (size_t)gl_bwd_p_smem_floats(p->prog, 4, false) * sizeof(float);
    const bool rows = p->row_flush && smem_rows <= 108 * 1024 && p->prog.g_total <= 64;
    if (p->tape_valid && p->feat_idx == 2 && !no_deflection && !per_component) {   // beta and the group Jacobian come from the forward kernel's tape
      const size_t smem_t = (size_t)gl_bwd_p_smem_floats(p->prog, 4, true, true) * sizeof(float);
      const size_t smem_tn = (size_t)gl_bwd_p_smem_floats(p->prog, 4, false, true) * sizeof(float);
      if (p->row_flush && smem_t <= 108 * 1024 && p->prog.g_total <= 64)
        k_raytrace_bwd_p<4, GL_FS2, true, false, true><<<grid, GLK_THREADS, smem_t, st>>>(p->prog, p->npix, p->d_grid_x, p->d_grid_y, p->d_ss_mask,
                                                                                       p->d_derived, 0, gss, p->d_gpart, p->d_nan, 0, 0, p->d_tape);
      else
        k_raytrace_bwd_p<4, GL_FS2, false, false, true><<<grid, GLK_THREADS, smem_tn, st>>>(p->prog, p->npix, p->d_grid_x, p->d_grid_y, p->d_ss_mask,
                                                                                        p->d_derived, 0, gss, p->d_gpart, p->d_nan, 0, 0, p->d_tape);
      GL_LAUNCH_CHECK("k_raytrace_bwd_p<tape>");
      return 0;
    }
#define GL_BWD_P(FS, ROWS, SM)                                                                                              \
    {                                                                                                                       \
      k_raytrace_bwd_p<4, FS, ROWS><<<grid, GLK_THREADS, (SM), st>>>(p->prog, p->npix, p->d_grid_x, p->d_grid_y, p->d_ss_mask,        \
                                                                   p->d_derived, no_deflection, gss, p->d_gpart, p->d_nan, sel_mask, sel_want, nullptr); \
    }
    if (p->feat_idx == 0 && rows && p->straight_line && !no_deflection && gl_is_benchmark_shape(p->prog)) {
      k_raytrace_bwd_p<4, GL_FS0, true, true><<<grid, GLK_THREADS, smem_rows, st>>>(p->prog, p->npix, p->d_grid_x, p->d_grid_y, p->d_ss_mask,
                                                                                  p->d_derived, no_deflection, gss, p->d_gpart, p->d_nan, sel_mask, sel_want, nullptr);
    } else
    if (p->feat_idx == 0) { if (rows) GL_BWD_P(GL_FS0, true, smem_rows) else GL_BWD_P(GL_FS0, false, smem) }
    else { if (rows) GL_BWD_P(GL_FS2, true, smem_rows) else GL_BWD_P(GL_FS2, false, smem) }
#undef GL_BWD_P
    GL_LAUNCH_CHECK("k_raytrace_bwd_p");
    return 0;
  }
  GL_FEAT_DISPATCH(p->feat_idx, {
    k_raytrace_bwd<4, F><<<grid, GLK_THREADS, smem, st>>>(p->prog, p->npix, p->d_grid_x, p->d_grid_y, p->d_ss_mask, p->d_derived,
                                                          no_deflection, gss, p->d_gpart, p->d_nan, sel_mask, sel_want);
  })
  GL_LAUNCH_CHECK("k_raytrace_bwd");
  return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry-point query (no link against libcuda)
typedef CUresult (*gl_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static gl_tmap_encode_fn gl_tmap_encoder() {
  static gl_tmap_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (gl_tmap_encode_fn)f;
  }
  return fn;
}
// 3-D fp32 tensor map (n, n, nimg) over contiguous n x n images, box (box_w, box_h, 1), zero fill outside
static bool gl_make_image_tmap(CUtensorMap* tm, const float* base, int n, size_t nimg, int box_w, int box_h) {
  gl_tmap_encode_fn enc = gl_tmap_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)nimg};
  const cuuint64_t gstr[2] = {(cuuint64_t)n * 4, (cuuint64_t)n * n * 4};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int A>
static int gl_launch_conv_fwd_A(gl_plan* p, const float* ss, float scale, float* img, bool like, float* gimg, cudaStream_t st, int nimg) {
  GlLikeArgs la{};
  la.enabled = like ? 1 : 0;
  if (like) { la.observed = p->d_obs; la.error_map = p->d_err; la.mask = p->d_mask; la.bg2 = p->bg2; la.inv_exp = p->inv_exp; }
  dim3 grid((unsigned)(p->gf.tiles_x * p->gf.tiles_y) * (unsigned)nimg);
  const int nph = p->ss * p->ss;
  if (p->conv_tma && p->conv_tma_ok && ((uintptr_t)ss % 16) == 0) {
    if (ss == p->d_comps2 && p->d_comps2) {   // second lstsq slot: its own cache entry (copied into the primary one for the launch)
      if (p->tmap_f2_base != ss || p->tmap_f2_nimg != nimg) {
        p->tmap_f2_ok = gl_make_image_tmap(&p->tmap_f2, ss, p->n, (size_t)nimg * nph, p->gf.tma_pitch, p->gf.in_rows);
        p->tmap_f2_base = ss; p->tmap_f2_nimg = nimg;
      }
      p->tmap_f = p->tmap_f2; p->tmap_f_ok = p->tmap_f2_ok; p->tmap_f_base = nullptr; p->tmap_f_nimg = -1;
    } else
    if (p->tmap_f_base != ss || p->tmap_f_nimg != nimg) {   // the map depends on the source buffer only
      p->tmap_f_ok = gl_make_image_tmap(&p->tmap_f, ss, p->n, (size_t)nimg * nph, p->gf.tma_pitch, p->gf.in_rows);
      p->tmap_f_base = ss; p->tmap_f_nimg = nimg;
    }
    if (p->tmap_f_ok) {
      if constexpr (A == GLC_A_HOT) {
        if (p->conv_const && p->conv_const_ok) {
          k_conv_fwd_tma<A, GLC_NF_HOT><<<grid, p->conv_threads_f, p->smem_cf_tma_c, st>>>(p->tmap_f, p->gf, p->d_wf, scale, img, la,
                                                                                        like ? p->d_like_part : nullptr, gimg, p->ct_f);
          GL_LAUNCH_CHECK("k_conv_fwd_tma<const taps>");
          return 0;
        }
      }
      k_conv_fwd_tma<A><<<grid, p->conv_threads_f, p->smem_cf_tma, st>>>(p->tmap_f, p->gf, p->d_wf, scale, img, la,
                                                                        like ? p->d_like_part : nullptr, gimg, GlTapsC<0>{});
      GL_LAUNCH_CHECK("k_conv_fwd_tma");
      return 0;
    }
  }
  k_conv_fwd<A><<<grid, p->conv_threads_f, p->smem_cf, st>>>(p->gf, ss, p->d_wf, scale, img, la, like ? p->d_like_part : nullptr, gimg);
  GL_LAUNCH_CHECK("k_conv_fwd");
  return 0;
}
template <int A>
static int gl_launch_conv_bwd_A(gl_plan* p, const float* gimg, float scale, float* gss, cudaStream_t st, int nimg, const int* img_list) {
  dim3 grid((unsigned)(p->gb.tiles_x * p->gb.tiles_y) * (unsigned)nimg);
  const int nph = p->ss * p->ss;
  if (p->conv_tma && p->conv_tma_b_ok && ((A - 1) & 1) == 0 && ((uintptr_t)gimg % 16) == 0 && ((uintptr_t)gss % 16) == 0) {
    if (p->tmap_bi_base != gimg || p->tmap_bo_base != gss || p->tmap_b_nimg != nimg) {
      p->tmap_b_ok = gl_make_image_tmap(&p->tmap_bi, gimg, p->n, (size_t)nimg, p->gb.tma_pitch, p->gb.in_rows) &&
                     gl_make_image_tmap(&p->tmap_bo, gss, p->n, (size_t)nimg * nph, p->gb.tw, p->gb.band_rows > 0 ? p->gb.band_rows : p->gb.th);
      p->tmap_bi_base = gimg; p->tmap_bo_base = gss; p->tmap_b_nimg = nimg;
    }
    if (p->tmap_b_ok) {
      if constexpr (A == GLC_A_HOT) {
        if (p->conv_const && p->conv_const_ok) {
          k_conv_bwd_tma<A, GLC_NF_HOT><<<grid, p->conv_threads_b, p->smem_cb_tma_c, st>>>(p->tmap_bi, p->tmap_bo, p->gb, p->d_wb, scale, p->ct_b, img_list);
          GL_LAUNCH_CHECK("k_conv_bwd_tma<const taps>");
          return 0;
        }
      }
      k_conv_bwd_tma<A><<<grid, p->conv_threads_b, p->smem_cb_tma, st>>>(p->tmap_bi, p->tmap_bo, p->gb, p->d_wb, scale, GlTapsC<0>{}, img_list);
      GL_LAUNCH_CHECK("k_conv_bwd_tma");
      return 0;
    }
  }
  k_conv_bwd<A><<<grid, p->conv_threads_b, p->smem_cb, st>>>(p->gb, gimg, p->d_wb, scale, nullptr, gss, img_list);
  GL_LAUNCH_CHECK("k_conv_bwd");
  return 0;
}
#define GL_CONV_DISPATCH(fn, ...)                                                                  \
  switch (p->A) {                                                                                  \
    case 1: return fn<1>(__VA_ARGS__); case 2: return fn<2>(__VA_ARGS__); case 3: return fn<3>(__VA_ARGS__);   \
    case 4: return fn<4>(__VA_ARGS__); case 5: return fn<5>(__VA_ARGS__); case 6: return fn<6>(__VA_ARGS__);   \
    case 7: return fn<7>(__VA_ARGS__); case 8: return fn<8>(__VA_ARGS__); case 9: return fn<9>(__VA_ARGS__);   \
    case 11: return fn<11>(__VA_ARGS__); case 13: return fn<13>(__VA_ARGS__); case 16: return fn<16>(__VA_ARGS__); \
    case 20: return fn<20>(__VA_ARGS__); case 25: return fn<25>(__VA_ARGS__); case 32: return fn<32>(__VA_ARGS__); \
  }                                                                                                \
  return gl_fail("conv dispatch: unsupported tap count")
static int gl_run_conv_fwd(gl_plan* p, const float* ss, float scale, float* img, bool like, float* gimg, cudaStream_t st,
                           int nimg = -1) {
  if (nimg < 0) nimg = p->bs;
  GL_CONV_DISPATCH(gl_launch_conv_fwd_A, p, ss, scale, img, like, gimg, st, nimg);
}
// img_list (device, optional): [0] = count, [1..] = the images to process -- the launch still covers nimg images, CTAs beyond
// the count exit at once (the count is only known on the device)
static int gl_run_conv_bwd(gl_plan* p, const float* gimg, float scale, float* gss, cudaStream_t st, int nimg = -1,
                           const int* img_list = nullptr) {
  if (nimg < 0) nimg = p->bs;
  GL_CONV_DISPATCH(gl_launch_conv_bwd_A, p, gimg, scale, gss, st, nimg, img_list);
}

// ---------------------------------------------------------------------------------------------
// data-path entry points
// ---------------------------------------------------------------------------------------------
extern "C" {

int gl_simulate_ss(gl_plan* p, const float* params_dev, float* ss_dev, void* stream) {
  if (!p || !params_dev || !ss_dev) return gl_fail("gl_simulate_ss: NULL argument");
  GL_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (gl_run_prep(p, params_dev, st)) return 1;
  if (gl_run_raytrace_fwd(p, p->d_ss, p->no_deflection, st)) return 1;
  dim3 grid((p->npix + 255) / 256 > 256 ? 256 : (p->npix + 255) / 256, p->bs);
  k_unpermute<<<grid, 256, 0, st>>>(p->npix, p->bs, p->d_perm, p->d_ss, ss_dev);
  GL_LAUNCH_CHECK("k_unpermute");
  return 0;
}

int gl_simulate(gl_plan* p, const float* params_dev, float* image_dev, void* stream) {
  if (!p || !params_dev || !image_dev) return gl_fail("gl_simulate: NULL argument");
  GL_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (gl_run_prep(p, params_dev, st)) return 1;
  if (gl_run_raytrace_fwd(p, p->d_ss, p->no_deflection, st)) return 1;
  return gl_run_conv_fwd(p, p->d_ss, p->conversion_factor, image_dev, false, nullptr, st);
}

int gl_beta(gl_plan* p, const float* params_dev, int32_t npts, const float* x_dev, const float* y_dev, float* bx, float* by,
            void* stream) {
  if (!p || !params_dev || !x_dev || !y_dev || !bx || !by || npts <= 0) return gl_fail("gl_beta: bad argument");
  GL_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (gl_run_prep(p, params_dev, st)) return 1;
  dim3 grid((npts + 127) / 128 > 1024 ? 1024 : (npts + 127) / 128, p->bs);
  GL_FEAT_DISPATCH(p->feat_idx, { k_points<F><<<grid, 128, (size_t)p->prog.der_total * sizeof(float), st>>>(p->prog, npts, x_dev, y_dev, p->d_derived, 0, bx, by); })
  GL_LAUNCH_CHECK("k_points");
  return 0;
}

/* mode 1: total deflection (alpha_x, alpha_y); mode 2: surface brightness (out1 unused) */
int gl_eval_points(gl_plan* p, const float* params_dev, int32_t npts, const float* x_dev, const float* y_dev, int32_t mode,
                   float* out0, float* out1, void* stream) {
  if (!p || !params_dev || !x_dev || !y_dev || !out0 || npts <= 0) return gl_fail("gl_eval_points: bad argument");
  if (mode < 0 || mode > 3 || (mode < 2 && !out1)) return gl_fail("gl_eval_points: bad mode");
  if (mode == 3 && p->prog.depth <= 0) return gl_fail("gl_eval_points: mode 3 needs a light profile with use_lstsq");
  GL_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (gl_run_prep(p, params_dev, st)) return 1;
  dim3 grid((npts + 127) / 128 > 1024 ? 1024 : (npts + 127) / 128, p->bs);
  GL_FEAT_DISPATCH(p->feat_idx, { k_points<F><<<grid, 128, (size_t)p->prog.der_total * sizeof(float), st>>>(p->prog, npts, x_dev, y_dev, p->d_derived, mode, out0, out1); })
  GL_LAUNCH_CHECK("k_points");
  return 0;
}

/* LensSimulator.magnification / convergence / shear inputs (tf/simulator.py:80-107): the Hessian
 * (f_xx, f_xy, f_yx, f_yy) of the summed deflection at npts points shared by all samples; out [bs][npts]. */
int gl_hessian(gl_plan* p, const float* params_dev, int32_t npts, const float* x_dev, const float* y_dev, float* fxx, float* fxy,
               float* fyx, float* fyy, void* stream) {
  if (!p || !params_dev || !x_dev || !y_dev || !fxx || !fxy || !fyx || !fyy || npts <= 0) return gl_fail("gl_hessian: bad argument");
  GL_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)p->prog.der_total * sizeof(double);
  if (smem > 200 * 1024) return gl_fail("gl_hessian: model too large");
  dim3 grid((npts + 127) / 128 > 64 ? 64 : (npts + 127) / 128, p->bs);
  k_hessian<<<grid, 128, smem, st>>>(p->prog, p->bs, params_dev, p->d_member_factor, p->d_amp_slot, npts, x_dev, y_dev, fxx, fxy, fyx, fyy);
  GL_LAUNCH_CHECK("k_hessian");
  return 0;
}

/* ForwardProbModel.stats_positions alone (tf/model.py:103-124): loglike / red_chi2 [bs], dparams [P][bs] or NULL. */
int gl_positions_loglike_grad(gl_plan* p, const float* params_dev, float* loglike_dev, float* red_chi2_dev, float* dparams_dev,
                              void* stream) {
  if (!p || !params_dev || !loglike_dev || !red_chi2_dev) return gl_fail("gl_positions_loglike_grad: NULL argument");
  GL_CUDA(cudaSetDevice(p->device));
  return gl_run_positions(p, params_dev, loglike_dev, red_chi2_dev, dparams_dev, (cudaStream_t)stream);
}

static int gl_lstsq_loglike_core(gl_plan* p, const float* params, float* loglike, float* red_chi2, float* dparams,
                                 const float* z, float* logp, float* dz, cudaStream_t st);
static int gl_loglike_core(gl_plan* p, const float* params, float* loglike, float* red_chi2, float* dparams,
                           const float* z, float* logp, float* dz, cudaStream_t st) {
  const bool pix = p->include_pixels != 0, pos = p->include_positions != 0;
  if (!pix && !pos) return gl_fail("log-likelihood requested with include_pixels = include_positions = 0");
  if (pix && !p->has_like) return gl_fail("log-likelihood requested but gl_plan_set_likelihood was never called");
  if (p->lstsq) {
    if (pos || !pix) return gl_fail("the lstsq (BackwardProbModel) path has no image-position term (tf/model.py:242-273)");
    return gl_lstsq_loglike_core(p, params, loglike, red_chi2, dparams, z, logp, dz, st);
  }
  const bool grad = dparams != nullptr;
  GL_TM(p, st, 1);
  if (pix && gl_run_prep(p, params, st)) return 1;
  GL_TM(p, st, 2);
  if (pix && gl_run_raytrace_fwd(p, p->d_ss, p->no_deflection, st, grad)) return 1;
  GL_TM(p, st, 3);
  if (pix && gl_run_conv_fwd(p, p->d_ss, p->conversion_factor, p->d_img, true, grad ? p->d_gimg : nullptr, st)) return 1;
  GL_TM(p, st, 4);
  if (grad && pix) {
    if (gl_run_conv_bwd(p, p->d_gimg, p->conversion_factor, p->d_ss, st)) return 1;
    GL_TM(p, st, 5);
    if (gl_run_raytrace_bwd(p, p->d_ss, p->no_deflection, st)) return 1;
  } else {
    GL_TM(p, st, 5);
  }
  GL_TM(p, st, 6);
  if (pos && gl_run_positions(p, params, p->d_pos_ll, p->d_pos_chi, grad ? p->d_pos_grad : nullptr, st)) return 1;
  const int tb = 32, gb = (p->bs + tb - 1) / tb;   // one warp per CTA: a batch of a few thousand samples spreads over all SMs
  k_sample_bwd<<<gb, tb, 0, st>>>(p->prog, p->bs, params, p->d_member_factor, p->d_amp_slot, p->d_derived, p->d_gpart, p->chunks, p->d_gsum,
                                  p->d_like_part, p->gf.tiles_x * p->gf.tiles_y, p->n_pix_used, loglike, red_chi2, dparams,
                                  p->d, p->d_leaves, z, z ? p->d_logprior : nullptr, logp, dz, pix ? 1 : 0,
                                  pos ? p->d_pos_ll : nullptr, pos ? p->d_pos_chi : nullptr, (pos && grad) ? p->d_pos_grad : nullptr);
  GL_LAUNCH_CHECK("k_sample_bwd");
  GL_TM(p, st, 7);
  if (p->tm_slots > 0) p->tm_calls++;
  return 0;
}

/* Sum of per-stage device times (ms) over the calls recorded since "timing" was switched on; stages:
 * unconstrain, prep, raytrace_fwd, conv_fwd, conv_bwd, raytrace_bwd, sample_bwd.  Synchronises. */
int gl_plan_get_timings(gl_plan* p, float* ms_out, int32_t* ncalls_out) {
  if (!p || !ms_out) return gl_fail("gl_plan_get_timings: NULL argument");
  GL_CUDA(cudaSetDevice(p->device));
  GL_CUDA(cudaDeviceSynchronize());
  for (int k = 0; k < GL_NSTAGE; ++k) ms_out[k] = 0.f;
  const int n = p->tm_calls < p->tm_slots ? p->tm_calls : p->tm_slots;
  for (int c = 0; c < n; ++c)
    for (int k = 0; k < GL_NSTAGE; ++k) {
      float ms = 0.f;
      const cudaEvent_t* ev = &p->tm_ev[(size_t)c * (GL_NSTAGE + 1)];
      if (k == 0) { if (!p->tm_have0) continue; }
      if (cudaEventElapsedTime(&ms, ev[k], ev[k + 1]) == cudaSuccess) ms_out[k] += ms;
    }
  if (ncalls_out) *ncalls_out = n;
  p->tm_calls = 0;
  return 0;
}

int gl_loglike_grad(gl_plan* p, const float* params_dev, float* loglike_dev, float* red_chi2_dev, float* dparams_dev,
                    void* stream) {
  if (!p || !params_dev) return gl_fail("gl_loglike_grad: NULL argument");
  GL_CUDA(cudaSetDevice(p->device));
  return gl_loglike_core(p, params_dev, loglike_dev, red_chi2_dev, dparams_dev, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int gl_unconstrain(gl_plan* p, const float* z_dev, float* params_dev, float* logprior_dev, void* stream) {
  if (!p || !z_dev) return gl_fail("gl_unconstrain: NULL argument");
  if (!p->d_leaves) return gl_fail("gl_unconstrain: gl_plan_set_prior was never called");
  GL_CUDA(cudaSetDevice(p->device));
  const int tb = 32, gb = (p->bs + tb - 1) / tb;   // one warp per CTA: a batch of a few thousand samples spreads over all SMs
  k_unconstrain<<<gb, tb, 0, (cudaStream_t)stream>>>(p->bs, p->d, p->d_leaves, z_dev, params_dev, logprior_dev);
  GL_LAUNCH_CHECK("k_unconstrain");
  return 0;
}

/* One Adam update of the MAP driver (tf/inference.py:34-37: optimizer.apply_gradients on the per-sample loss) in one launch:
 * g = grad * grad_scale (non-finite -> 0 when scrub_nan), m = b1 m + (1 - b1) g, v = b2 v + (1 - b2) g^2,
 * x -= alpha m / (sqrt(v) + eps) with alpha = lr sqrt(1 - b2^t) / (1 - b1^t) folded by the caller (Keras' formulation).
 * Element-wise on n floats; the torch formulation of the same update is eight launches. */
__global__ void k_adam(long n, float* __restrict__ x, const float* __restrict__ grad, float grad_scale, float* __restrict__ m,
                       float* __restrict__ v, float b1, float b2, float omb1, float omb2, float alpha, float eps, int scrub_nan) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float g = grad[i] * grad_scale;
    if (scrub_nan && !(fabsf(g) <= 3.4028234e38f)) g = 0.f;
    const float mi = m[i] * b1 + g * omb1;          // omb = 1 - beta, formed in double on the host (1.f - 0.999f is off by 1.3e-5)
    const float vi = v[i] * b2 + (g * g) * omb2;
    m[i] = mi; v[i] = vi;
    x[i] = x[i] - alpha * (mi / (sqrtf(vi) + eps));
  }
}
int gl_adam_step(float* x_dev, const float* grad_dev, float* m_dev, float* v_dev, int64_t n, double grad_scale, double beta1, double beta2,
                 double alpha, double eps, int32_t scrub_nan, void* stream) {
  if (!x_dev || !grad_dev || !m_dev || !v_dev || n <= 0) return gl_fail("gl_adam_step: bad argument");
  const int threads = 256;
  long blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_adam<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>((long)n, x_dev, grad_dev, (float)grad_scale, m_dev, v_dev, (float)beta1,
                                                                 (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)alpha,
                                                                 (float)eps, scrub_nan);
  GL_LAUNCH_CHECK("k_adam");
  return 0;
}

/* dz = (d params / d z)^T dparams (+ d log_prior / d z when with_prior); dparams_dev may be NULL (prior gradient only). */
int gl_chain_grad(gl_plan* p, const float* z_dev, const float* dparams_dev, int32_t with_prior, float* logprior_dev, float* dz_dev,
                  void* stream) {
  if (!p || !z_dev || !dz_dev) return gl_fail("gl_chain_grad: NULL argument");
  if (!p->d_leaves) return gl_fail("gl_chain_grad: gl_plan_set_prior was never called");
  GL_CUDA(cudaSetDevice(p->device));
  const int tb = 32, gb = (p->bs + tb - 1) / tb;   // one warp per CTA: a batch of a few thousand samples spreads over all SMs
  k_chain_grad<<<gb, tb, 0, (cudaStream_t)stream>>>(p->bs, p->d, p->d_leaves, z_dev, dparams_dev, with_prior, logprior_dev, dz_dev);
  GL_LAUNCH_CHECK("k_chain_grad");
  return 0;
}

int gl_logprob_grad(gl_plan* p, const float* z_dev, float* logp_dev, float* red_chi2_dev, float* dz_dev, void* stream) {
  if (!p || !z_dev) return gl_fail("gl_logprob_grad: NULL argument");
  if (!p->d_leaves) return gl_fail("gl_logprob_grad: gl_plan_set_prior was never called");
  GL_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int tb = 32, gb = (p->bs + tb - 1) / tb;   // one warp per CTA: a batch of a few thousand samples spreads over all SMs
  GL_TM(p, st, 0);
  p->tm_have0 = 1;
  k_unconstrain<<<gb, tb, 0, st>>>(p->bs, p->d, p->d_leaves, z_dev, p->d_params, p->d_logprior);
  GL_LAUNCH_CHECK("k_unconstrain");
  return gl_loglike_core(p, p->d_params, nullptr, red_chi2_dev, dz_dev ? p->d_dparams : nullptr, z_dev, logp_dev, dz_dev, st);
}

int gl_logprob_grad_host(gl_plan* p, const float* z_host, float* logp_host, float* red_chi2_host, float* dz_host) {
  if (!p || !z_host) return gl_fail("gl_logprob_grad_host: NULL argument");
  if (!p->d_leaves) return gl_fail("gl_logprob_grad_host: gl_plan_set_prior was never called");
  GL_CUDA(cudaSetDevice(p->device));
  const size_t nz = (size_t)p->bs * p->d * sizeof(float), nb = (size_t)p->bs * sizeof(float);
  GL_CUDA(cudaMemcpyAsync(p->d_z, z_host, nz, cudaMemcpyHostToDevice, 0));
  if (gl_logprob_grad(p, p->d_z, p->d_logp, p->d_chi, dz_host ? p->d_dz : nullptr, nullptr)) return 1;
  if (logp_host) GL_CUDA(cudaMemcpyAsync(logp_host, p->d_logp, nb, cudaMemcpyDeviceToHost, 0));
  if (red_chi2_host) GL_CUDA(cudaMemcpyAsync(red_chi2_host, p->d_chi, nb, cudaMemcpyDeviceToHost, 0));
  if (dz_host) GL_CUDA(cudaMemcpyAsync(dz_host, p->d_dz, nz, cudaMemcpyDeviceToHost, 0));
  GL_CUDA(cudaStreamSynchronize(0));
  return 0;
}

int gl_simulate_host(gl_plan* p, const float* params_host, float* image_host) {
  if (!p || !params_host || !image_host) return gl_fail("gl_simulate_host: NULL argument");
  GL_CUDA(cudaSetDevice(p->device));
  const size_t np = (size_t)p->prog.n_params * p->bs * sizeof(float);
  GL_CUDA(cudaMemcpyAsync(p->d_params, params_host, np, cudaMemcpyHostToDevice, 0));
  if (gl_simulate(p, p->d_params, p->d_img, nullptr)) return 1;
  GL_CUDA(cudaMemcpyAsync(image_host, p->d_img, (size_t)p->bs * p->n * p->n * sizeof(float), cudaMemcpyDeviceToHost, 0));
  GL_CUDA(cudaStreamSynchronize(0));
  return 0;
}

// lstsq workspace (gl_plan_reserve_lstsq; plan creation calls it for models with use_lstsq components): the
// component stack of `chunk` samples, sized by a memory budget when chunk == 0.  Never called from a data-path entry point.
static int gl_lstsq_reserve(gl_plan* p, int chunk) {
  const int D = p->prog.depth, npx = p->n * p->n;
  if (D <= 0) return gl_fail("lstsq: the model has no linear light component");
  GL_CUDA(cudaSetDevice(p->device));
  if (p->d_comps) {
    GL_CUDA(cudaDeviceSynchronize());
    for (float** q : {&p->d_R, &p->d_gram, &p->d_coef, &p->d_ll, &p->d_comps, &p->d_R2, &p->d_gram2, &p->d_comps2}) { if (*q) cudaFree(*q); *q = nullptr; }
    cudaFree(p->d_solve_queue); p->d_solve_queue = nullptr;
    if (p->d_solve_queue2) { cudaFree(p->d_solve_queue2); p->d_solve_queue2 = nullptr; }
  }
  const size_t per_sample = (size_t)D * p->npix * sizeof(float);
  // component-stack budget: 32 GB of the 180 GB HBM3e, but never more than 40 % of what is free right now
  size_t budget = (size_t)32 << 30, free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b / 5 * 2 < budget) budget = free_b / 5 * 2;
  size_t cb = budget / (per_sample + (size_t)D * npx * sizeof(float));
  if (cb < 1) cb = 1;
  // One chunk when the whole batch fits the budget (measured at C3, bs 2048: 23.3 ms; every extra chunk pays the ~2.5 ms
  // eigen-solve tail of its singular samples again).  When memory forces chunks, they alternate between two slots so that one
  // chunk's tail runs under the other's convolution (4 chunks: 23.8 ms pipelined vs 31.3 ms back to back).
  bool two = cb < (size_t)p->bs || (chunk > 0 && chunk < p->bs);
  if (two && chunk == 0) { cb /= 2; if (cb < 1) cb = 1; }
  if (chunk > 0) cb = (size_t)chunk;
  if (cb >= (size_t)p->bs) { cb = p->bs; two = false; }
  if (p->bs < 16) two = false;
  p->lq_chunk = (int)cb;
  GL_CUDA(cudaMalloc((void**)&p->d_R, cb * (size_t)D * npx * sizeof(float)));
  GL_CUDA(cudaMalloc((void**)&p->d_gram, cb * (size_t)(D + 1) * (D + 1) * sizeof(float)));
  GL_CUDA(cudaMalloc((void**)&p->d_solve_queue, (cb + 1) * sizeof(int)));
  GL_CUDA(cudaMalloc((void**)&p->d_coef, (size_t)p->bs * D * sizeof(float)));
  GL_CUDA(cudaMalloc((void**)&p->d_ll, (size_t)p->bs * 2 * sizeof(float)));
  GL_CUDA(cudaMalloc((void**)&p->d_comps, cb * per_sample));
  if (two) {
    GL_CUDA(cudaMalloc((void**)&p->d_R2, cb * (size_t)D * npx * sizeof(float)));
    GL_CUDA(cudaMalloc((void**)&p->d_gram2, cb * (size_t)(D + 1) * (D + 1) * sizeof(float)));
    GL_CUDA(cudaMalloc((void**)&p->d_solve_queue2, (cb + 1) * sizeof(int)));
    GL_CUDA(cudaMalloc((void**)&p->d_comps2, cb * per_sample));
    int prio_lo = 0, prio_hi = 0;
    GL_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int k = 0; k < 2; ++k) {
      if (!p->lq_stream[k]) GL_CUDA(cudaStreamCreateWithPriority(&p->lq_stream[k], cudaStreamNonBlocking, prio_lo));
      if (!p->lq_stream_hi[k]) GL_CUDA(cudaStreamCreateWithPriority(&p->lq_stream_hi[k], cudaStreamNonBlocking, prio_hi));
      if (!p->lq_done[k]) GL_CUDA(cudaEventCreateWithFlags(&p->lq_done[k], cudaEventDisableTiming));
      if (!p->lq_conv[k]) GL_CUDA(cudaEventCreateWithFlags(&p->lq_conv[k], cudaEventDisableTiming));
    }
    if (!p->lq_fork) GL_CUDA(cudaEventCreateWithFlags(&p->lq_fork, cudaEventDisableTiming));
  }
  if (!p->lq_tail_stream) {
    int prio_lo = 0, prio_hi = 0;
    GL_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    GL_CUDA(cudaStreamCreateWithPriority(&p->lq_tail_stream, cudaStreamNonBlocking, prio_hi));
    GL_CUDA(cudaEventCreateWithFlags(&p->lq_tail_fork, cudaEventDisableTiming));
    GL_CUDA(cudaEventCreateWithFlags(&p->lq_tail_done, cudaEventDisableTiming));
  }
  return 0;
}
static int gl_lstsq_alloc(gl_plan* p) {
  if (p->d_comps) return 0;
  if (p->prog.depth <= 0) return gl_fail("lstsq: the model has no linear light component");
  return gl_fail("lstsq: no workspace -- call gl_plan_reserve_lstsq first (plan creation does it for models with use_lstsq components)");
}

static int gl_lstsq_forward(gl_plan* p, const float* params, float* image, float* coeffs_out, float* loglike,
                            float* red_chi2, bool want_gimg, cudaStream_t st, float* stack_out = nullptr) {
  if (!p->has_like || !p->d_err) return gl_fail("lstsq: needs gl_plan_set_likelihood with observed image and error_map");
  const int D = p->prog.depth, npx = p->n * p->n;
  if (D <= 0) return gl_fail("lstsq: the model has no linear light component");
  if (D + 1 > 128 || D > 104) return gl_fail("lstsq: more than 104 linear components are not supported yet");
  if (gl_lstsq_alloc(p)) return 1;
  if (gl_run_prep(p, params, st)) return 1;
  const size_t smem_der = (size_t)p->prog.der_total * sizeof(float);
  const int nt = (D + 1 + 3) / 4;
  const size_t smem_gram = (size_t)GLL_PT * nt * 4 * sizeof(float);
  const int npair = ((D + 1) & ~1) / 2;
  const size_t smem_solve = (size_t)(2 * D * (D | 1) + 2 * (npair + 1)) * sizeof(double) + (size_t)(2 * (npair + 1) + 2) * sizeof(int) + 4 * sizeof(double);
  GL_CUDA(cudaMemsetAsync(p->d_nan, 0, (size_t)p->bs * sizeof(int), st));   // component values scrubbed per sample (read by the adjoint)
  // Chunks alternate between two slots (buffers + stream): chunk k+1's ray-shooting and convolution fill the SMs while chunk k's
  // Gram / Cholesky / eigen-solve tail (one CTA per sample, a few long-running CTAs for the singular samples) drains.
  const bool pipe = p->lq_pipeline && p->d_comps2 && p->lq_chunk < p->bs;
  if (pipe) {
    GL_CUDA(cudaEventRecord(p->lq_fork, st));
    for (int k = 0; k < 2; ++k) {
      GL_CUDA(cudaStreamWaitEvent(p->lq_stream[k], p->lq_fork, 0));
      GL_CUDA(cudaStreamWaitEvent(p->lq_stream_hi[k], p->lq_fork, 0));
    }
  }
  int ci = 0;
  for (int b0 = 0; b0 < p->bs; b0 += p->lq_chunk, ++ci) {
    const int nb = (p->bs - b0 < p->lq_chunk) ? p->bs - b0 : p->lq_chunk;
    const int slot = pipe ? (ci & 1) : 0;
    cudaStream_t cs = pipe ? p->lq_stream[slot] : st;        // ray-shooting + convolution
    cudaStream_t ts = pipe ? p->lq_stream_hi[slot] : st;     // the latency-bound tail
    float* comps = slot ? p->d_comps2 : p->d_comps;
    float* Rbuf = slot ? p->d_R2 : p->d_R;
    float* gram = slot ? p->d_gram2 : p->d_gram;
    int* queue = slot ? p->d_solve_queue2 : p->d_solve_queue;
    if (pipe && ci >= 2) GL_CUDA(cudaStreamWaitEvent(cs, p->lq_done[slot], 0));   // the slot's previous tail has read R / Gram
    dim3 grid(p->chunks, nb);
    GL_FEAT_DISPATCH(p->feat_idx, {
      k_raytrace_comps<4, F><<<grid, GLL_THREADS, smem_der, cs>>>(p->prog, p->npix, p->d_grid_x, p->d_grid_y, p->d_ss_mask,
                                                                 p->d_derived + (size_t)b0 * p->prog.der_total, p->no_deflection, comps, p->d_nan + b0);
    })
    GL_LAUNCH_CHECK("k_raytrace_comps");
    if (gl_run_conv_fwd(p, comps, 1.f, Rbuf, false, nullptr, cs, nb * D)) return 1;
    if (stack_out)   // return_stacked: the convolved, pooled unit-amplitude components, nothing solved
      GL_CUDA(cudaMemcpyAsync(stack_out + (size_t)b0 * D * npx, Rbuf, (size_t)nb * D * npx * sizeof(float), cudaMemcpyDeviceToDevice, cs));
    if (pipe) {
      GL_CUDA(cudaEventRecord(p->lq_conv[slot], cs));
      GL_CUDA(cudaStreamWaitEvent(ts, p->lq_conv[slot], 0));
    }
    if (!stack_out) {
      if (p->gram_tc) {
        GL_CUDA(gl_launch_gram_tc(nb, D, npx, Rbuf, p->d_w, p->d_obs, gram, nullptr, ts, p->gram_ksplit));
        ++g_launch_count;
      } else {
        k_gram<<<nb, GLL_THREADS, smem_gram, ts>>>(D, npx, Rbuf, p->d_w, p->d_obs, gram);
        GL_LAUNCH_CHECK("k_gram");
      }
      GL_CUDA(cudaMemsetAsync(queue, 0, sizeof(int), ts));   // [0] = count, [1..] = sample indices
      float* coef = p->d_coef + (size_t)b0 * D;
      float* im = image ? image + (size_t)b0 * npx : nullptr;
      float* llp = loglike ? loglike + b0 : nullptr;
      float* chip = red_chi2 ? red_chi2 + b0 : nullptr;
      float* gim = want_gimg ? p->d_gimg + (size_t)b0 * npx : nullptr;
      const size_t smem_img = (size_t)D * sizeof(float);
      // Gradient requested, one chunk: the eigen-solve leaves the caller's stream.  Phase 0 flags the samples it queues (bit 1 of
      // d_nan); the tail stream -- highest priority, and its kernels become ready first, so its few CTAs are resident before the
      // caller's stream fills the SMs -- runs the eigen-solve, the image / likelihood and the amplitude patch of the queued samples,
      // while the caller's stream goes on with everything else for the other samples.
      const bool hide = want_gimg && !pipe && nb == p->bs && p->lq_hide_tail && p->lq_tail_stream != nullptr;
      k_pinv_solve<<<nb, 128, smem_solve, ts>>>(D, gram, 1e-6, 16, coef, 0, queue + 1, queue, hide ? p->d_nan + b0 : nullptr);
      GL_LAUNCH_CHECK("k_pinv_solve");
      if (hide) {
        cudaStream_t tl = p->lq_tail_stream;
        GL_CUDA(cudaEventRecord(p->lq_tail_fork, ts));
        GL_CUDA(cudaStreamWaitEvent(tl, p->lq_tail_fork, 0));
        k_pinv_solve<<<nb, 512, smem_solve, tl>>>(D, gram, 1e-6, 16, coef, 1, queue + 1, queue, nullptr);
        GL_LAUNCH_CHECK("k_pinv_solve");
        k_lstsq_image<<<nb, GLL_THREADS, smem_img, tl>>>(D, npx, Rbuf, coef, p->d_obs, p->d_err, im, llp, chip, gim, nullptr, queue);
        GL_LAUNCH_CHECK("k_lstsq_image");
        k_patch_amps<<<(nb + 31) / 32, 32, 0, tl>>>(p->prog, nb, coef, p->d_derived, nullptr, queue);
        GL_LAUNCH_CHECK("k_patch_amps");
        GL_CUDA(cudaEventRecord(p->lq_tail_done, tl));
        p->lq_tail_pending = true;
        k_lstsq_image<<<nb, GLL_THREADS, smem_img, ts>>>(D, npx, Rbuf, coef, p->d_obs, p->d_err, im, llp, chip, gim, p->d_nan + b0, nullptr);
        GL_LAUNCH_CHECK("k_lstsq_image");
      } else {
        k_pinv_solve<<<nb, 512, smem_solve, ts>>>(D, gram, 1e-6, 16, coef, 1, queue + 1, queue, nullptr);
        GL_LAUNCH_CHECK("k_pinv_solve");
        k_lstsq_image<<<nb, GLL_THREADS, smem_img, ts>>>(D, npx, Rbuf, coef, p->d_obs, p->d_err, im, llp, chip, gim, nullptr, nullptr);
        GL_LAUNCH_CHECK("k_lstsq_image");
      }
    }
    if (pipe) GL_CUDA(cudaEventRecord(p->lq_done[slot], ts));
  }
  if (pipe) {
    for (int k = 0; k < 2 && k < ci; ++k) GL_CUDA(cudaStreamWaitEvent(st, p->lq_done[k], 0));
  }
  if (coeffs_out) GL_CUDA(cudaMemcpyAsync(coeffs_out, p->d_coef, (size_t)p->bs * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

// log-like (+ gradient) of BackwardProbModel; shares the tail (k_sample_bwd) with the forward model.
static int gl_lstsq_loglike_core(gl_plan* p, const float* params, float* loglike, float* red_chi2, float* dparams,
                                 const float* z, float* logp, float* dz, cudaStream_t st) {
  if (gl_lstsq_alloc(p)) return 1;   // before taking workspace pointers
  float* ll = loglike ? loglike : p->d_ll;
  float* chi = red_chi2 ? red_chi2 : p->d_ll + p->bs;
  const bool grad = dparams != nullptr;
  if (gl_lstsq_forward(p, params, nullptr, nullptr, ll, chi, grad, st)) return 1;
  const int tb = 32, gb = (p->bs + tb - 1) / tb;   // one warp per CTA: a batch of a few thousand samples spreads over all SMs
  if (grad) {
    const bool late = p->lq_tail_pending;   // set by gl_lstsq_forward: some samples' amplitudes come from the tail stream
    p->lq_tail_pending = false;
    k_patch_amps<<<gb, tb, 0, st>>>(p->prog, p->bs, p->d_coef, p->d_derived, late ? p->d_nan : nullptr, nullptr);
    GL_LAUNCH_CHECK("k_patch_amps");
    if (gl_run_conv_bwd(p, p->d_gimg, 1.f, p->d_ss, st)) return 1;   // (late samples: stale dL/dimage in, overwritten below)
    if (gl_run_raytrace_bwd(p, p->d_ss, p->no_deflection, st, late ? 2 : 1)) return 1;
    if (late) {
      // second pass for the late samples only: their image, likelihood, dL/dimage and patched amplitudes are complete once the
      // tail stream is done (normally long before: ~2 ms of eigen-solve against ~4 ms of adjoint work on this stream)
      GL_CUDA(cudaStreamWaitEvent(st, p->lq_tail_done, 0));
      if (gl_run_conv_bwd(p, p->d_gimg, 1.f, p->d_ss, st, -1, p->d_solve_queue)) return 1;
      if (gl_run_raytrace_bwd(p, p->d_ss, p->no_deflection, st, 3)) return 1;
    }
  }
  k_sample_bwd<<<gb, tb, 0, st>>>(p->prog, p->bs, params, p->d_member_factor, p->d_amp_slot, p->d_derived, p->d_gpart, p->chunks,
                                  p->d_gsum, nullptr, 0, 1.f, ll, chi, dparams, p->d, p->d_leaves, z,
                                  z ? p->d_logprior : nullptr, logp, dz, 1, nullptr, nullptr, nullptr);
  GL_LAUNCH_CHECK("k_sample_bwd");
  return 0;
}

int gl_lstsq_simulate(gl_plan* p, const float* params_dev, float* image_dev, float* coeffs_dev, void* stream) {
  if (!p || !params_dev) return gl_fail("gl_lstsq_simulate: NULL argument");
  GL_CUDA(cudaSetDevice(p->device));
  return gl_lstsq_forward(p, params_dev, image_dev, coeffs_dev, nullptr, nullptr, false, (cudaStream_t)stream);
}
int gl_lstsq_stack(gl_plan* p, const float* params_dev, float* stack_dev, void* stream) {
  if (!p || !params_dev || !stack_dev) return gl_fail("gl_lstsq_stack: NULL argument");
  GL_CUDA(cudaSetDevice(p->device));
  return gl_lstsq_forward(p, params_dev, nullptr, nullptr, nullptr, nullptr, false, (cudaStream_t)stream, stack_dev);
}
int gl_lstsq_loglike_grad(gl_plan* p, const float* params_dev, float* loglike_dev, float* red_chi2_dev, float* dparams_dev,
                          void* stream) {
  if (!p || !params_dev) return gl_fail("gl_lstsq_loglike_grad: NULL argument");
  GL_CUDA(cudaSetDevice(p->device));
  return gl_lstsq_loglike_core(p, params_dev, loglike_dev, red_chi2_dev, dparams_dev, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int gl_plan_reserve_lstsq(gl_plan* p, int32_t chunk) {
  if (!p || chunk < 0) return gl_fail("gl_plan_reserve_lstsq: bad argument");
  return gl_lstsq_reserve(p, chunk);
}

int gl_fp32_peak(int32_t device, float* tflops_ffma, float* tflops_ffma2) {
  if (!tflops_ffma || !tflops_ffma2) return gl_fail("gl_fp32_peak: NULL argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return gl_fail("gl_fp32_peak: no such CUDA device");
  GL_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  GL_CUDA(cudaGetDeviceProperties(&prop, device));
  float* sink = nullptr;
  GL_CUDA(cudaMalloc((void**)&sink, sizeof(float)));
  cudaError_t e = gl_fp32_probe_run(prop.multiProcessorCount, 5, tflops_ffma, tflops_ffma2, sink);
  cudaFree(sink);
  g_launch_count += 12;
  if (e != cudaSuccess) return gl_fail(std::string("gl_fp32_peak: ") + cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
