// gl_positions.cuh -- image-position likelihood and lensing-Hessian kernels.
//
// ForwardProbModel.stats_positions (src/gigalens/tf/model.py:103-124) constrains cluster models with
// the observed positions of multiply-imaged sources: every image is ray-traced to the source plane,
// the scatter about the barycentre is measured in units of centroid_error / magnification, and the
// magnification is 1/det(I - H) with H the Hessian of the summed deflection
// (LensSimulator.magnification, tf/simulator.py:80-91).  The workload is tens of points per sample, so
// these kernels are written for exactness, not throughput: the whole chain -- parameter conversion,
// deflection, Hessian, likelihood, and the parameter gradient -- runs in FP64 on forward-mode dual
// numbers (gl_math.cuh, GlDual) from the fp32 raw parameters, one warp per sample.  Near critical
// curves |mu| reaches 1e2-1e3 and an fp32 evaluation of the complex-log dPIE form loses the 1e-5
// parity bound; FP64 costs nothing at this size.
#pragma once
#include "gl_program.h"

#define GLP_THREADS 32
constexpr unsigned GLF_LENSES = GLF_EPL | GLF_SHEAR | GLF_SIE | GLF_SIS | GLF_NFW | GLF_DPIS | GLF_DPIE | GLF_TNFW | GLF_DPIEP;

struct GlRowFlush {   // per-thread cotangent row in shared memory: deterministic, no atomics
  double* row;
  __device__ void operator()(const double* acc, int n, int off) {
    for (int k = 0; k < n; ++k) row[off + k] += acc[k];
  }
};

// One warp per sample.  Shared memory (doubles): der[der_total] | pt[12 * npts] (bx, by, H[4], gbx, gby, gH[4])
// | sys[2 * n_sys] | rows[GLP_THREADS][g_total] | gsum[g_total].
__global__ void __launch_bounds__(GLP_THREADS) k_positions(GlProgram P, int bs, const float* __restrict__ params,
                                                           const float* __restrict__ member_factor, const int* __restrict__ amp_slot,
                                                           int npts, int n_sys, const int* __restrict__ sys_off,
                                                           const float* __restrict__ px, const float* __restrict__ py,
                                                           const float* __restrict__ ex, const float* __restrict__ ey, float n_position,
                                                           float* __restrict__ loglike, float* __restrict__ red_chi2,
                                                           float* __restrict__ dparams /*[P][bs] or null*/) {
  extern __shared__ __align__(16) double sm[];
  double* der = sm;
  double* bx = der + P.der_total; double* by = bx + npts; double* H = by + npts;
  double* gbx = H + 4 * npts; double* gby = gbx + npts; double* gH = gby + npts;
  double* sys = gH + 4 * npts;
  double* rows = sys + 2 * n_sys;
  double* gsum = rows + (size_t)GLP_THREADS * P.g_total;
  const int b = blockIdx.x, t = threadIdx.x;
  if (t == 0) gl_sample_prep<double, float>(P, params, bs, b, member_factor, amp_slot, nullptr, der);
  __syncthreads();
  for (int p = t; p < npts; p += GLP_THREADS)
    gl_point_hessian<double, GLF_LENSES>(P, der, (double)px[p], (double)py[p], bx[p], by[p], H + 4 * p);
  __syncthreads();
  for (int s = t; s < n_sys; s += GLP_THREADS) {
    const int o = sys_off[s], n = sys_off[s + 1] - o;
    double chi2 = 0.0, norm = 0.0;
    double exd[64], eyd[64];   // a system has at most 64 images (checked by gl_plan_set_positions)
    for (int i = 0; i < n; ++i) { exd[i] = (double)ex[o + i]; eyd[i] = (double)ey[o + i]; }
    gl_positions_system<double>(n, bx + o, by + o, H + 4 * o, exd, eyd, chi2, norm, dparams ? gbx + o : nullptr, gby + o, gH + 4 * o);
    sys[2 * s] = chi2; sys[2 * s + 1] = norm;
  }
  __syncthreads();
  if (t == 0) {
    double chi2 = 0.0, norm = 0.0;
    for (int s = 0; s < n_sys; ++s) { chi2 += sys[2 * s]; norm += sys[2 * s + 1]; }
    loglike[b] = (float)(-0.5 * (chi2 + norm));
    red_chi2[b] = (float)(chi2 / (double)n_position);
  }
  if (!dparams) return;
  double* row = rows + (size_t)t * P.g_total;
  for (int k = 0; k < P.g_total; ++k) row[k] = 0.0;
  GlRowFlush fl{row};
  for (int p = t; p < npts; p += GLP_THREADS)
    gl_point_positions_bwd<double, GLF_LENSES>(P, der, (double)px[p], (double)py[p], gbx[p], gby[p], gH + 4 * p, fl);
  __syncthreads();
  for (int k = t; k < P.g_total; k += GLP_THREADS) {
    double s = 0.0;
    for (int r = 0; r < GLP_THREADS; ++r) s += rows[(size_t)r * P.g_total + k];
    gsum[k] = s;
  }
  __syncthreads();
  if (t == 0) gl_sample_prep_bwd<double, float>(P, params, bs, b, member_factor, amp_slot, der, gsum, dparams);
}

// beta and Hessian (f_xx, f_xy, f_yx, f_yy) at npts points shared by all samples; out [bs][npts] each.
// grid = (point blocks, bs); thread 0 converts the sample's parameters once per block.
__global__ void __launch_bounds__(128) k_hessian(GlProgram P, int bs, const float* __restrict__ params,
                                                 const float* __restrict__ member_factor, const int* __restrict__ amp_slot, int npts,
                                                 const float* __restrict__ px, const float* __restrict__ py, float* __restrict__ fxx,
                                                 float* __restrict__ fxy, float* __restrict__ fyx, float* __restrict__ fyy) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.y;
  if (threadIdx.x == 0) gl_sample_prep<double, float>(P, params, bs, b, member_factor, amp_slot, nullptr, sm);
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += gridDim.x * blockDim.x) {
    double bxv, byv, H[4];
    gl_point_hessian<double, GLF_LENSES>(P, sm, (double)px[p], (double)py[p], bxv, byv, H);
    const size_t o = (size_t)b * npts + p;
    fxx[o] = (float)H[0]; fxy[o] = (float)H[1]; fyx[o] = (float)H[2]; fyy[o] = (float)H[3];
  }
}
