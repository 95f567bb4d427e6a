// hostcheck.cpp -- TEST-ONLY host build of the device arithmetic (gl_math.cuh / gl_program.h).
//
// Compiled by g++ (no CUDA) from tests/test_hostcheck.py into tests/hostcheck/_build/.  It runs
// the exact template code the sm_100a kernels instantiate -- prep, per-pixel forward, per-pixel
// adjoint, prep adjoint -- in fp32 and fp64 with plain loops, so every formula and every hand
// adjoint can be checked against the oracle's autograd on a machine without a GPU.  Nothing
// under gigalens_b200/ loads this library; it is not a CPU fallback.
#include <cstring>
#include <string>
#include <vector>

#include "../../gigalens_b200/csrc/gl_build.h"

static thread_local std::string g_err;

template <class T>
struct HostFlush {
  T* g;
  void operator()(const T* acc, int n, int off) { for (int k = 0; k < n; ++k) g[off + k] += acc[k]; }
};

template <class T>
static int run(const gl_model_desc* m, int bs, const T* params, int npix, const T* gx, const T* gy, int no_deflection,
               int epl_batch_max, T* ss_out, const T* g_ss, T* gparams, T* beta_out, T* comps_out) {
  GlBuilt B;
  std::string e = gl_build_program(m, B);
  if (!e.empty()) { g_err = e; return 1; }
  GlProgram& P = B.prog;
  for (int i = 0; i < P.n_prof; ++i)
    if (B.table_off[i] >= 0) P.prof[i].table = B.tables.data() + B.table_off[i];
  const float* mf = B.member_factor.empty() ? nullptr : B.member_factor.data();
  const int* as = B.amp_slot.empty() ? nullptr : B.amp_slot.data();
  std::vector<float> fmax(P.n_prof, -1.f);
  if (epl_batch_max) {
    for (int i = 0; i < P.n_lens; ++i) {
      if (P.prof[i].type != GLT_EPL) continue;
      T best = T(0);
      for (int b = 0; b < bs; ++b) {
        T raw[GL_MAX_RAW];
        gl_gather_raw<T, T>(P.prof[i], params, bs, b, mf, 0, raw);
        T phi, q, c;
        ellip_fwd(raw[2], raw[3], T(1), phi, q, c);
        T f = (T(1) - q) / (T(1) + q);
        if (f > best) best = f;
      }
      fmax[i] = (float)best;
    }
  }
  std::vector<T> der(P.der_total), g(P.g_total > 0 ? P.g_total : 1);
  for (int b = 0; b < bs; ++b) {
    gl_sample_prep<T, T>(P, params, bs, b, mf, as, epl_batch_max ? fmax.data() : nullptr, der.data());
    std::fill(g.begin(), g.end(), T(0));
    HostFlush<T> fl{g.data()};
    for (int p = 0; p < npix; ++p) {
      T x[1] = {gx[p]}, y[1] = {gy[p]}, v[1];
      if (ss_out) {
        gl_pix_image<T, 1, GLF_ALL>(P, der.data(), x, y, no_deflection != 0, v);
        ss_out[(size_t)b * npix + p] = gl_isnan(v[0]) ? T(0) : v[0];
      }
      if (comps_out) {   // [bs][depth][npix]
        T bx[1], by[1];
        if (no_deflection) { bx[0] = x[0]; by[0] = y[0]; } else gl_pix_beta<T, 1, GLF_ALL>(P, der.data(), x, y, bx, by);
        gl_point_components<T, GLF_ALL>(P, der.data(), x[0], y[0], bx[0], by[0], comps_out + (size_t)b * P.depth * npix + p, npix, true);
      }
      if (beta_out) {
        T bx[1], by[1];
        gl_pix_beta<T, 1, GLF_ALL>(P, der.data(), x, y, bx, by);
        beta_out[((size_t)b * 2 + 0) * npix + p] = bx[0];
        beta_out[((size_t)b * 2 + 1) * npix + p] = by[0];
      }
      if (g_ss && gparams) {
        T gs[1] = {g_ss[(size_t)b * npix + p]};
        T scr[GL_EPL_NSTATE];   // series state handed from the forward sweep to the adjoint, as in the kernels
        gl_pix_image_bwd<T, 1, GLF_ALL>(P, der.data(), x, y, gs, no_deflection != 0, fl, scr, 1);
      }
    }
    if (g_ss && gparams) gl_sample_prep_bwd<T, T>(P, params, bs, b, mf, as, der.data(), g.data(), gparams);
  }
  return 0;
}

// The two-pixel packed lane type (GlF2) through the same drivers: host build of the code path the
// k_raytrace_*_p kernels instantiate (EPL | SHEAR | SERSIC, and the cluster set NFW | DPIE | SHEAR | SERSIC).
struct HostFlush2 {
  float* g;
  void operator()(const GlF2* acc, int n, int off) { for (int k = 0; k < n; ++k) g[off + k] += acc[k].x + acc[k].y; }
};
static int run_packed(const gl_model_desc* m, int bs, const float* params, int npix, const float* gx, const float* gy,
                      float* ss_out, const float* g_ss, float* gparams, int straight_line = 0) {
  constexpr unsigned F = GLF_EPL | GLF_SHEAR | GLF_SERSIC | GLF_NFW | GLF_DPIE;   // union of the packed kernels' feature sets
  GlBuilt B;
  std::string e = gl_build_program(m, B);
  if (!e.empty()) { g_err = e; return 1; }
  GlProgram& P = B.prog;
  for (int i = 0; i < P.n_prof; ++i)
    if ((gl_feature_of(P.prof[i].type) & F) == 0) { g_err = "packed lanes: unsupported profile"; return 1; }
  if (npix % 2) { g_err = "packed lanes need an even pixel count"; return 1; }
  if (straight_line && !gl_is_benchmark_shape(P)) { g_err = "straight-line drivers: not the benchmark-shape program"; return 1; }
  const float* mf = B.member_factor.empty() ? nullptr : B.member_factor.data();
  std::vector<float> der(P.der_total), g(P.g_total > 0 ? P.g_total : 1);
  for (int b = 0; b < bs; ++b) {
    gl_sample_prep<float, float>(P, params, bs, b, mf, nullptr, nullptr, der.data());
    std::fill(g.begin(), g.end(), 0.f);
    HostFlush2 fl{g.data()};
    for (int p = 0; p < npix; p += 2) {
      GlF2 x[1] = {GlF2(gx[p], gx[p + 1])}, y[1] = {GlF2(gy[p], gy[p + 1])}, v[1];
      if (ss_out) {
        if (straight_line) gl_pix_image_bs<GlF2, 1>(P, der.data(), x, y, v);
        else gl_pix_image<GlF2, 1, F>(P, der.data(), x, y, false, v);
        ss_out[(size_t)b * npix + p] = (v[0].x != v[0].x) ? 0.f : v[0].x;
        ss_out[(size_t)b * npix + p + 1] = (v[0].y != v[0].y) ? 0.f : v[0].y;
      }
      if (g_ss && gparams) {
        GlF2 gs[1] = {GlF2(g_ss[(size_t)b * npix + p], g_ss[(size_t)b * npix + p + 1])};
        GlF2 scr[GL_EPL_NSTATE];
        if (straight_line) gl_pix_image_bwd_bs<GlF2, 1>(P, der.data(), x, y, gs, fl, scr, 1);
        else gl_pix_image_bwd<GlF2, 1, F>(P, der.data(), x, y, gs, false, fl, scr, 1);
      }
    }
    if (g_ss && gparams) gl_sample_prep_bwd<float, float>(P, params, bs, b, mf, nullptr, der.data(), g.data(), gparams);
  }
  return 0;
}

// stats_positions (tf/model.py:103-124) through the dual-number point drivers: beta, Hessian, log-like,
// chi2 and the parameter gradient for n_sys image systems (points concatenated, n_img per system).
template <class T>
static int run_positions(const gl_model_desc* m, int bs, const T* params, int n_sys, const int* n_img, const T* px, const T* py,
                         const T* ex, const T* ey, int use_fwdmode, T* beta_out, T* hess_out, T* loglike, T* chi2_out, T* gparams) {
  GlBuilt B;
  std::string e = gl_build_program(m, B, use_fwdmode != 0);
  if (!e.empty()) { g_err = e; return 1; }
  GlProgram& P = B.prog;
  const float* mf = B.member_factor.empty() ? nullptr : B.member_factor.data();
  const int* as = B.amp_slot.empty() ? nullptr : B.amp_slot.data();
  int npts = 0;
  for (int s = 0; s < n_sys; ++s) npts += n_img[s];
  std::vector<T> der(P.der_total), g(P.g_total > 0 ? P.g_total : 1), bx(npts), by(npts), H(4 * npts), gbx(npts), gby(npts), gH(4 * npts);
  for (int b = 0; b < bs; ++b) {
    gl_sample_prep<T, T>(P, params, bs, b, mf, as, nullptr, der.data());
    for (int p = 0; p < npts; ++p) {
      gl_point_hessian<T, GLF_ALL>(P, der.data(), px[p], py[p], bx[p], by[p], &H[4 * p]);
      if (beta_out) { beta_out[((size_t)b * 2 + 0) * npts + p] = bx[p]; beta_out[((size_t)b * 2 + 1) * npts + p] = by[p]; }
      if (hess_out) for (int c = 0; c < 4; ++c) hess_out[((size_t)b * 4 + c) * npts + p] = H[4 * p + c];
    }
    T chi2 = T(0), norm = T(0);
    int o = 0;
    for (int s = 0; s < n_sys; ++s) {
      gl_positions_system<T>(n_img[s], &bx[o], &by[o], &H[4 * o], ex + o, ey + o, chi2, norm, &gbx[o], &gby[o], &gH[4 * o]);
      o += n_img[s];
    }
    if (loglike) loglike[b] = -(chi2 + norm) / T(2);
    if (chi2_out) chi2_out[b] = chi2 / T(2 * npts);
    if (gparams) {
      std::fill(g.begin(), g.end(), T(0));
      HostFlush<T> fl{g.data()};
      for (int p = 0; p < npts; ++p)
        gl_point_positions_bwd<T, GLF_ALL>(P, der.data(), px[p], py[p], gbx[p], gby[p], &gH[4 * p], fl);
      gl_sample_prep_bwd<T, T>(P, params, bs, b, mf, as, der.data(), g.data(), gparams);
    }
  }
  return 0;
}

extern "C" {
int glh_positions_f64(const gl_model_desc* m, int bs, const double* params, int n_sys, const int* n_img, const double* px,
                      const double* py, const double* ex, const double* ey, int use_fwdmode, double* beta_out, double* hess_out,
                      double* loglike, double* chi2, double* gparams) {
  return run_positions<double>(m, bs, params, n_sys, n_img, px, py, ex, ey, use_fwdmode, beta_out, hess_out, loglike, chi2, gparams);
}
int glh_positions_f32(const gl_model_desc* m, int bs, const float* params, int n_sys, const int* n_img, const float* px,
                      const float* py, const float* ex, const float* ey, int use_fwdmode, float* beta_out, float* hess_out,
                      float* loglike, float* chi2, float* gparams) {
  return run_positions<float>(m, bs, params, n_sys, n_img, px, py, ex, ey, use_fwdmode, beta_out, hess_out, loglike, chi2, gparams);
}
}

extern "C" {
int glh_run_f32x2(const gl_model_desc* m, int bs, const float* params, int npix, const float* gx, const float* gy,
                  float* ss_out, const float* g_ss, float* gparams) {
  return run_packed(m, bs, params, npix, gx, gy, ss_out, g_ss, gparams);
}
// the straight-line drivers of the benchmark-shape program (gl_pix_image_bs / gl_pix_image_bwd_bs), packed lanes
int glh_run_f32x2_bs(const gl_model_desc* m, int bs, const float* params, int npix, const float* gx, const float* gy,
                     float* ss_out, const float* g_ss, float* gparams) {
  return run_packed(m, bs, params, npix, gx, gy, ss_out, g_ss, gparams, 1);
}
const char* glh_last_error() { return g_err.c_str(); }
int glh_depth(const gl_model_desc* m) { GlBuilt B; std::string e = gl_build_program(m, B); if (!e.empty()) { g_err = e; return -1; } return B.prog.depth; }
int glh_run_f64(const gl_model_desc* m, int bs, const double* params, int npix, const double* gx, const double* gy,
                int no_deflection, int epl_batch_max, double* ss_out, const double* g_ss, double* gparams, double* beta_out,
                double* comps_out) {
  return run<double>(m, bs, params, npix, gx, gy, no_deflection, epl_batch_max, ss_out, g_ss, gparams, beta_out, comps_out);
}
int glh_run_f32(const gl_model_desc* m, int bs, const float* params, int npix, const float* gx, const float* gy,
                int no_deflection, int epl_batch_max, float* ss_out, const float* g_ss, float* gparams, float* beta_out,
                float* comps_out) {
  return run<float>(m, bs, params, npix, gx, gy, no_deflection, epl_batch_max, ss_out, g_ss, gparams, beta_out, comps_out);
}
}
