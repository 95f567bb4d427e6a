// gl_lstsq.cuh -- linear light-amplitude solve of LensSimulator.lstsq_simulate
// (src/gigalens/tf/simulator.py:158-240; coherent tensor layout src/gigalens/jax/simulator.py:171-195).
//
//   comps[b][c][ss pixel]   unit-amplitude light components on the ray-shooting grid (k_raytrace_comps)
//   R[b][c][pixel]          after PSF conv + pooling (k_conv_fwd with bs' = bs*D, scale 1), NaN scrubbed
//   X = R * W, Y = obs * W,  W = 1/err_map        (:232-234)
//   gram[b] = [X Y]^T [X Y]                         (k_gram: (D+1)^2 matrix, last column = X^T Y)
//   coeffs  = pinv(X^T X, rcond=1e-6) X^T Y         (k_pinv_solve: symmetric Jacobi eigen-solve in fp64)
//   image   = sum_c R_c coeffs_c                    (k_lstsq_image, + Gaussian log-likelihood and dL/dimage)
//
// Gradient w.r.t. the non-linear parameters: coeffs minimise the same chi^2 the likelihood measures,
// so (envelope theorem) dL/dtheta = <dL/dimage, sum_c coeffs_c dR_c/dtheta>: the adjoint is the
// ordinary simulate-adjoint with the amplitudes frozen at `coeffs` (k_patch_amps writes them into the
// derived blocks).  Exact while no singular value is truncated; with an active rcond cut TF
// differentiates through the SVD with the mask fixed and the two differ at first order in the cut
// directions (documented in DESIGN.md).
#pragma once
#include <cuda_runtime.h>

#include "gl_program.h"

#define GLL_THREADS 256

template <int PPT, unsigned F>
__global__ void __launch_bounds__(GLL_THREADS) k_raytrace_comps(GlProgram P, int npix, const float* __restrict__ grid_x,
                                                                const float* __restrict__ grid_y,
                                                                const unsigned char* __restrict__ ss_mask,
                                                                const float* __restrict__ derived, int no_deflection,
                                                                float* __restrict__ comps, int* __restrict__ nan_count) {
  extern __shared__ __align__(16) float s_der[];
  const int b = blockIdx.y;
  const float* dsrc = derived + (size_t)b * P.der_total;
  for (int i = threadIdx.x; i < P.der_total; i += blockDim.x) s_der[i] = dsrc[i];
  __syncthreads();
  const int per_batch = GLL_THREADS * PPT;
  const int nbatch = (npix + per_batch - 1) / per_batch;
  float* dst = comps + (size_t)b * P.depth * npix;
  for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
    float x[PPT], y[PPT], bx[PPT], by[PPT];
    int pix[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      pix[j] = batch * per_batch + j * GLL_THREADS + threadIdx.x;
      const int p = pix[j] < npix ? pix[j] : 0;
      x[j] = __ldg(grid_x + p); y[j] = __ldg(grid_y + p);
    }
    if (no_deflection) {
#pragma unroll
      for (int j = 0; j < PPT; ++j) { bx[j] = x[j]; by[j] = y[j]; }
    } else {
      gl_pix_beta<float, PPT, F>(P, s_der, x, y, bx, by);
    }
    for (int j = 0; j < PPT; ++j) {
      if (pix[j] >= npix) continue;
      const bool keep = !ss_mask || ss_mask[pix[j]];
      const int n_nan = gl_point_components<float, F>(P, s_der, x[j], y[j], bx[j], by[j], dst + pix[j], npix, keep);
      if (n_nan && nan_count) atomicOr(nan_count + b, 1);   // flag bit 0: the adjoint treats this sample per component (SCRUB)
    }
  }
}

// gram[b] = [X Y]^T [X Y] with X[p][c] = R[b][c][p] * w[p], Y[p] = obs[p] * w[p].
//   grid = bs, block = 256; 4x4 register tiles over the (D+1)^2 outputs (D + 1 <= 128).
#define GLL_PT 32
__global__ void __launch_bounds__(GLL_THREADS) k_gram(int D, int npx, const float* __restrict__ R, const float* __restrict__ w,
                                                      const float* __restrict__ obs, float* __restrict__ gram) {
  extern __shared__ __align__(16) float s_x[];   // [GLL_PT][Dp]
  const int Dx = D + 1, nt = (Dx + 3) >> 2, Dp = nt * 4;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* Rb = R + (size_t)b * D * npx;
  const int ntile = nt * nt;
  float acc[4][4][4];   // up to 4 tiles per thread
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;
  for (int p0 = 0; p0 < npx; p0 += GLL_PT) {
    for (int e = tid; e < Dp * GLL_PT; e += GLL_THREADS) {
      const int c = e / GLL_PT, pp = e - c * GLL_PT, p = p0 + pp;
      float v = 0.f;
      if (p < npx) {
        if (c < D) { v = __ldg(Rb + (size_t)c * npx + p); if (v != v) v = 0.f; v *= w[p]; }   // NaN scrub (:228)
        else if (c == D) v = obs[p] * w[p];
      }
      s_x[pp * Dp + c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int tile = tid + t * GLL_THREADS;
      if (tile < ntile) {
        const int ti = tile / nt, tj = tile - ti * nt;
#pragma unroll 4
        for (int pp = 0; pp < GLL_PT; ++pp) {
          const float4 a = *reinterpret_cast<const float4*>(s_x + pp * Dp + ti * 4);
          const float4 c4 = *reinterpret_cast<const float4*>(s_x + pp * Dp + tj * 4);
          const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[t][i][j] = fmaf(av[i], bv[j], acc[t][i][j]);
        }
      }
    }
    __syncthreads();
  }
  float* out = gram + (size_t)b * Dx * Dx;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int tile = tid + t * GLL_THREADS;
    if (tile < ntile) {
      const int ti = tile / nt, tj = tile - ti * nt;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = ti * 4 + i, c = tj * 4 + j;
          if (r < Dx && c < Dx) out[(size_t)r * Dx + c] = acc[t][i][j];
        }
    }
  }
}

// coeffs = pinv(G, rcond) h for the symmetric PSD G = X^T X, in fp64 in shared memory.  Fast path: a
// Cholesky solve when a cheap certificate proves that no eigenvalue falls under the rcond cut (then
// pinv == inverse).  General path: cyclic parallel Jacobi (round-robin pairing, two-sided rotations),
// then V diag(1/lambda_i > cut) V^T h.
// tf.linalg.pinv keeps singular values > rcond * max (src/gigalens/tf/simulator.py:235).
//   grid = bs, block = 128, smem = 2*D*D doubles + small
// `flags` (phase 0, optional): bit 1 of flags[b] is set for the samples left to phase 1.
// Two launches per chunk: phase 0 (128 threads per sample) runs the Cholesky fast path and appends the samples whose
// certificate fails to `queue`; phase 1 (512 threads per sample, grid = chunk, CTAs beyond *count exit at once) runs
// the eigen-solve for those only.  In the C3 prior ~0.6 % of the draws have a genuinely singular Gram matrix (cond
// ~5e8); in a single launch those few CTAs ran ~10x longer than the rest and the tail was 90 % of the kernel time.
__global__ void __launch_bounds__(512) k_pinv_solve(int D, const float* __restrict__ gram, double rcond, int max_sweeps,
                                                    float* __restrict__ coeffs, int phase, int* __restrict__ queue,
                                                    int* __restrict__ count, int* __restrict__ flags) {
  extern __shared__ __align__(16) double s_d[];
  const int LD = D | 1;            // odd leading dimension: column walks (stride LD doubles) hit distinct banks
  double* A = s_d;                 // [D][LD]
  double* V = A + D * LD;          // [D][LD]
  double* cs = V + D * LD;         // [m/2 + 1][2] rotation (c, s) per pair; reused for y at the end
  int* pq = reinterpret_cast<int*>(cs + 2 * ((D + 1) / 2 + 1));   // [m/2 + 1][2]
  double* red = reinterpret_cast<double*>(pq + 2 * ((D + 1) / 2 + 1) + 2);   // [4]
  const int tid = threadIdx.x, nthr = blockDim.x;
  if (phase == 1 && (int)blockIdx.x >= *count) return;
  const int b = (phase == 1) ? queue[blockIdx.x] : (int)blockIdx.x;
  const int Dx = D + 1;
  const float* G = gram + (size_t)b * Dx * Dx;
  if (phase == 0)
  // ---- fast path: when G is provably well conditioned (lambda_min > rcond * lambda_max) the
  // pseudo-inverse IS the inverse, and a Cholesky solve (D^3/3 flops) replaces the eigen-solve
  // (~60 D^3).  Certificate:  lambda_max <= tr(G)  and  lambda_min >= 1 / tr(G^-1) = 1 / ||L^-1||_F^2.
  // Thread i owns ROW i of the factor (D <= 128 threads work): per column j every row i >= j forms its dot product
  // s_i = G_ij - sum_{k<j} L_ik L_jk (row j is a broadcast read, row i the thread's own: conflict-free), thread j takes the
  // square root, the others divide: two barriers per column and no index arithmetic -- the first version (right-looking
  // rank-1 updates over a flattened index with an integer division per element, half the threads idle, then L^-1 with
  // every thread at a different row: 10 M bank conflicts) took 340 us per sample, this one ~40.
  {
    double* L = V;       // Cholesky factor (lower)
    double* Li = A;      // L^-1 (lower), column c owned by thread c
    for (int e = tid; e < D * D; e += nthr) {
      const int i = e / D, j = e - i * D;
      L[i * LD + j] = 0.5 * ((double)G[(size_t)i * Dx + j] + (double)G[(size_t)j * Dx + i]);
    }
    if (tid == 0) red[3] = 1.0;   // ok flag
    __syncthreads();
    const int i = tid;
    for (int j = 0; j < D; ++j) {
      double sdot = 0.0;
      if (i >= j && i < D) {
        const double* Ri = L + i * LD;
        const double* Rj = L + j * LD;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int k = 0;
        for (; k + 3 < j; k += 4) {
          s0 += Ri[k] * Rj[k]; s1 += Ri[k + 1] * Rj[k + 1]; s2 += Ri[k + 2] * Rj[k + 2]; s3 += Ri[k + 3] * Rj[k + 3];
        }
        for (; k < j; ++k) s0 += Ri[k] * Rj[k];
        sdot = Ri[j] - ((s0 + s1) + (s2 + s3));
      }
      if (i == j) {
        if (!(sdot > 0.0)) red[3] = 0.0;
        L[j * LD + j] = sqrt(sdot > 0.0 ? sdot : 1.0);
      }
      __syncthreads();
      if (i > j && i < D) L[i * LD + j] = sdot / L[j * LD + j];
      __syncthreads();
    }
    // L^-1 by forward substitution, column c per thread; every thread walks the SAME rows r and inner index k (entries above
    // its own diagonal are zeros), so L_rk is one broadcast read and column c of L^-1 is this thread's private bank
    double tri = 0.0;
    if (i < D) {
      for (int r = 0; r < D; ++r) {
        double a0 = (r == i) ? 1.0 : 0.0, a1 = 0.0;
        const double* Rr = L + r * LD;
        int k = 0;
        for (; k + 1 < r; k += 2) { a0 -= Rr[k] * Li[k * LD + i]; a1 -= Rr[k + 1] * Li[(k + 1) * LD + i]; }
        if (k < r) a0 -= Rr[k] * Li[k * LD + i];
        const double v = (r >= i) ? (a0 + a1) / Rr[r] : 0.0;
        Li[r * LD + i] = v;
        tri += v * v;
      }
    }
    // traces: tr(G) (<= lambda_max bound) and tr(G^-1) = ||L^-1||_F^2
    double trg = (i < D) ? (double)G[(size_t)i * Dx + i] : 0.0;
    for (int o = 16; o > 0; o >>= 1) { trg += __shfl_xor_sync(0xffffffffu, trg, o); tri += __shfl_xor_sync(0xffffffffu, tri, o); }
    double* part = cs;   // [nwarps][2] (cs is free until the solve below)
    if ((tid & 31) == 0) { part[2 * (tid >> 5)] = trg; part[2 * (tid >> 5) + 1] = tri; }
    __syncthreads();
    if (tid == 0) {
      double a = 0.0, c2 = 0.0;
      for (int w2 = 0; w2 < (nthr >> 5); ++w2) { a += part[2 * w2]; c2 += part[2 * w2 + 1]; }
      red[0] = a; red[1] = c2;
    }
    __syncthreads();
    const bool well = red[3] > 0.5 && red[1] > 0.0 && (1.0 / red[1]) > rcond * red[0];
    if (well) {
      double* yv = cs;   // D doubles fit (see below)
      __syncthreads();   // everyone has read `part` (aliases yv)
      for (int r = tid; r < D; r += nthr) {   // y = L^-1 h
        double acc = 0.0;
        for (int k = 0; k <= r; ++k) acc += Li[r * LD + k] * (double)G[(size_t)k * Dx + D];
        yv[r] = acc;
      }
      __syncthreads();
      for (int k = tid; k < D; k += nthr) {   // c = L^-T y
        double acc = 0.0;
        for (int r = k; r < D; ++r) acc += Li[r * LD + k] * yv[r];
        coeffs[(size_t)b * D + k] = (float)acc;
      }
      return;
    }
    if (tid == 0) {
      queue[atomicAdd(count, 1)] = b;   // leave it to the phase-1 launch
      if (flags) atomicOr(flags + b, 2);   // flag bit 1: amplitudes arrive late (the eigen-solve runs on a side stream, gl_lstsq_forward)
    }
    return;
  }
  // ---- general path: symmetric Jacobi eigen-decomposition, pinv with the rcond cut
  for (int e = tid; e < D * D; e += nthr) {
    const int i = e / D, j = e - i * D;
    // symmetrise (the two triangles of the fp32 Gram matrix can differ by rounding)
    A[i * LD + j] = 0.5 * ((double)G[(size_t)i * Dx + j] + (double)G[(size_t)j * Dx + i]);
    V[i * LD + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  const int m = (D + 1) & ~1, npair = m / 2;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    // convergence: off-diagonal mass relative to the diagonal (1e-11 in norm: eigenvalues are then
    // converged far below the 1e-6 rcond cut and below fp32 resolution of the result)
    if (tid < 32) {
      double off = 0.0, dg = 0.0;
      for (int e = tid; e < D * D; e += 32) {
        const int i = e / D, j = e - i * D;
        const double v = A[i * LD + j] * A[i * LD + j];
        if (i == j) dg += v; else off += v;
      }
      for (int o = 16; o > 0; o >>= 1) { off += __shfl_xor_sync(0xffffffffu, off, o); dg += __shfl_xor_sync(0xffffffffu, dg, o); }
      if (tid == 0) { red[0] = off; red[1] = dg; }
    }
    __syncthreads();
    if (red[0] <= 1e-22 * red[1]) break;
    for (int round = 0; round < m - 1; ++round) {
      if (tid < npair) {
        int p, q;
        if (tid == 0) { p = m - 1; q = round; }
        else { p = (round + tid) % (m - 1); q = (round - tid + (m - 1)) % (m - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < D) {
          const double apq = A[p * LD + q];
          if (fabs(apq) > 1e-300) {
            const double tau = (A[q * LD + q] - A[p * LD + p]) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = rsqrt(1.0 + t * t);
            s = t * c;
          }
        }
        cs[2 * tid] = c; cs[2 * tid + 1] = s;
        pq[2 * tid] = p; pq[2 * tid + 1] = q;
      }
      __syncthreads();
      // two-sided rotation A <- J^T A J in ONE pass: the 2x2 block (rows of pair r) x (columns of pair r2) depends
      // only on itself, so every thread owns whole blocks and no intermediate barrier is needed.
      // (flat over the npair x npair blocks: every lane has work -- the former 8 x 64 mapping left 31 of 64 lanes idle)
      for (int e = tid; e < npair * npair; e += nthr) {
        const int r = e / npair, r2 = e - r * npair;
        const int p = pq[2 * r], q = pq[2 * r + 1];
        const double c1 = cs[2 * r], s1 = cs[2 * r + 1];
        const bool vq = q < D;
        const int p2 = pq[2 * r2], q2 = pq[2 * r2 + 1];
        const double c2 = cs[2 * r2], s2 = cs[2 * r2 + 1];
        const bool vq2 = q2 < D;
        const double app = A[p * LD + p2], apq2 = vq2 ? A[p * LD + q2] : 0.0;
        const double aqp = vq ? A[q * LD + p2] : 0.0, aqq = (vq && vq2) ? A[q * LD + q2] : 0.0;
        // columns (J on the right), then rows (J^T on the left)
        const double tpp = c2 * app - s2 * apq2, tpq = s2 * app + c2 * apq2;
        const double tqp = c2 * aqp - s2 * aqq, tqq = s2 * aqp + c2 * aqq;
        A[p * LD + p2] = c1 * tpp - s1 * tqp;
        if (vq2) A[p * LD + q2] = c1 * tpq - s1 * tqq;
        if (vq) A[q * LD + p2] = s1 * tpp + c1 * tqp;
        if (vq && vq2) A[q * LD + q2] = s1 * tpq + c1 * tqq;
      }
      // eigenvectors: V <- V J   (k fastest across threads: stride LD, conflict-free)
      for (int e = tid; e < npair * D; e += nthr) {
        const int pr = e / D, k = e - pr * D;
        const int p = pq[2 * pr], q = pq[2 * pr + 1];
        if (q >= D) continue;
        const double c = cs[2 * pr], s = cs[2 * pr + 1];
        const double vkp = V[k * LD + p], vkq = V[k * LD + q];
        V[k * LD + p] = c * vkp - s * vkq; V[k * LD + q] = s * vkp + c * vkq;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // lambda_i = A_ii; y = V^T h; coeffs = V (y / lambda) over kept eigenvalues
  double* y = cs;   // cs holds 2*(npair+1) >= D + 2 doubles
  if (tid == 0) {
    double lmax = 0.0;
    for (int i = 0; i < D; ++i) lmax = fmax(lmax, fabs(A[i * LD + i]));
    red[2] = lmax;
  }
  __syncthreads();
  const double cut = rcond * red[2];
  for (int i = tid; i < D; i += nthr) {
    double acc = 0.0;
    for (int k = 0; k < D; ++k) acc += V[k * LD + i] * (double)G[(size_t)k * Dx + D];
    const double lam = A[i * LD + i];
    y[i] = (fabs(lam) > cut) ? acc / lam : 0.0;
  }
  __syncthreads();
  for (int k = tid; k < D; k += nthr) {
    double acc = 0.0;
    for (int i = 0; i < D; ++i) acc += V[k * LD + i] * y[i];
    coeffs[(size_t)b * D + k] = (float)acc;
  }
}

// image = sum_c R_c coeffs_c; Independent(Normal(obs, err)).log_prob(image) (tf/model.py:221-226,266-273)
// and dL/dimage.  grid = bs, block = 256.
__global__ void __launch_bounds__(GLL_THREADS) k_lstsq_image(int D, int npx, const float* __restrict__ R,
                                                             const float* __restrict__ coeffs, const float* __restrict__ obs,
                                                             const float* __restrict__ err, float* __restrict__ image,
                                                             float* __restrict__ loglike, float* __restrict__ red_chi2,
                                                             float* __restrict__ gimg, const int* __restrict__ flags,
                                                             const int* __restrict__ list) {
  extern __shared__ float s_c[];   // [D]
  __shared__ float s_red[2][GLL_THREADS / 32];
  // sample selection (gradient path with the eigen-solve on a side stream): `list` = only the listed samples ([0] = count,
  // [1..] = indices; CTAs beyond the count exit), else `flags` = skip the samples whose amplitudes arrive late (bit 1)
  int b = blockIdx.x;
  if (list) { if (b >= list[0]) return; b = list[1 + b]; }
  else if (flags && (flags[b] & 2)) return;
  const int tid = threadIdx.x;
  for (int c = tid; c < D; c += blockDim.x) s_c[c] = coeffs[(size_t)b * D + c];
  __syncthreads();
  const float* Rb = R + (size_t)b * D * npx;
  float ll = 0.f, chi = 0.f;
  auto finish = [&](int p, float v) {
    if (image) image[(size_t)b * npx + p] = v;
    if (obs) {
      const float e = err[p];
      const float q = (v - obs[p]) / e;
      chi += q * q;
      ll += -0.5f * q * q - 0.9189385332046727f - logf(e);
      if (gimg) gimg[(size_t)b * npx + p] = -q / e;
    }
  };
  if ((npx & 3) == 0) {
    // 4 pixels per thread as one float4 per channel, 6 channels per trip: 6 independent 16-byte loads in flight per thread
    // (the scalar loop had one dependent load + FMA at a time: 8 % issue utilisation, 1.2 ms for 1.9 GB)
    const float4* R4 = reinterpret_cast<const float4*>(Rb);
    const int n4 = npx >> 2;
    for (int p4 = tid; p4 < n4; p4 += blockDim.x) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      int c = 0;
      for (; c + 5 < D; c += 6) {
        float4 r[6];
#pragma unroll
        for (int u = 0; u < 6; ++u) r[u] = __ldg(R4 + (size_t)(c + u) * n4 + p4);
#pragma unroll
        for (int u = 0; u < 6; ++u) {
          const float k = s_c[c + u];
          v.x = fmaf((r[u].x == r[u].x) ? r[u].x : 0.f, k, v.x); v.y = fmaf((r[u].y == r[u].y) ? r[u].y : 0.f, k, v.y);
          v.z = fmaf((r[u].z == r[u].z) ? r[u].z : 0.f, k, v.z); v.w = fmaf((r[u].w == r[u].w) ? r[u].w : 0.f, k, v.w);
        }
      }
      for (; c < D; ++c) {
        const float4 r = __ldg(R4 + (size_t)c * n4 + p4);
        const float k = s_c[c];
        v.x = fmaf((r.x == r.x) ? r.x : 0.f, k, v.x); v.y = fmaf((r.y == r.y) ? r.y : 0.f, k, v.y);
        v.z = fmaf((r.z == r.z) ? r.z : 0.f, k, v.z); v.w = fmaf((r.w == r.w) ? r.w : 0.f, k, v.w);
      }
      finish(4 * p4, v.x); finish(4 * p4 + 1, v.y); finish(4 * p4 + 2, v.z); finish(4 * p4 + 3, v.w);
    }
  } else {
    for (int p = tid; p < npx; p += blockDim.x) {
      float v = 0.f;
      for (int c = 0; c < D; ++c) {
        float r = __ldg(Rb + (size_t)c * npx + p);
        if (r != r) r = 0.f;
        v = fmaf(r, s_c[c], v);
      }
      finish(p, v);
    }
  }
  if (obs && (loglike || red_chi2)) {
    for (int o = 16; o > 0; o >>= 1) { ll += __shfl_xor_sync(0xffffffffu, ll, o); chi += __shfl_xor_sync(0xffffffffu, chi, o); }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = ll; s_red[1][tid >> 5] = chi; }
    __syncthreads();
    if (tid == 0) {
      float a = 0.f, c2 = 0.f;
      for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) { a += s_red[0][w2]; c2 += s_red[1][w2]; }
      if (loglike) loglike[b] = a;
      if (red_chi2) red_chi2[b] = c2 / (float)npx;
    }
  }
}

// write the solved amplitudes into the derived blocks (Sersic Ie, Shapelets amplitudes) so that the
// ordinary adjoint kernels differentiate the combined image with the amplitudes frozen.
// Sample selection as in k_lstsq_image.
__global__ void k_patch_amps(GlProgram P, int bs, const float* __restrict__ coeffs, float* __restrict__ derived,
                             const int* __restrict__ flags, const int* __restrict__ list) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bs) return;
  if (list) { if (b >= list[0]) return; b = list[1 + b]; }
  else if (flags && (flags[b] & 2)) return;
  float* der = derived + (size_t)b * P.der_total;
  const float* c = coeffs + (size_t)b * P.depth;
  for (int i = P.n_lens; i < P.n_prof; ++i) {
    const GlProf& pr = P.prof[i];
    if (pr.type == GLT_SERSIC || pr.type == GLT_SERSIC_ELLIPSE) der[pr.der_off + SER_IE] = c[pr.comp_off];
    else if (pr.type == GLT_CORE_SERSIC) der[pr.der_off + CS_IE] = c[pr.comp_off];
    else if (pr.type == GLT_SHAPELETS) {
      const int L = shp_layers(pr.n_max);
      for (int k = 0; k < L; ++k) der[pr.der_off + shp_amp_off(pr.n_max) + k] = c[pr.comp_off + k];
    }
  }
}
