// gl_program.h -- the flat "program" a physical model compiles to, and the per-sample /
// per-pixel drivers that interpret it.  Shared by the CUDA library (device, fp32) and by the
// test-only host harness (g++, fp32/fp64).
//
// A program is the reference's PhysicalModel (src/gigalens/tf/model.py:290-306) flattened:
// entries [0, n_lens) are deflectors, then lens-light, then source-light profiles.  Each entry
// owns a slice of the per-sample derived vector (der_off) and of the per-sample dvar-cotangent
// vector (g_off).  An entry with n_members > 0 is a scaling-relation sum over a galaxy catalogue
// (src/gigalens/tf/profiles/mass/scaling_relation.py:61-70): it owns n_members consecutive
// derived blocks and n_members consecutive dvar blocks.
#pragma once
#include "gl_math.cuh"

#define GL_MAX_PROF 24

struct GlProf {
  int type;
  unsigned flags;
  int der_off;     // offset of the (first) derived block
  int der_size;    // size of one derived block (per member)
  int g_off;       // offset of the (first) dvar block
  int n_dvars;     // dvars per member
  int ts;          // EPL table stride
  int niter;       // EPL iteration cap
  int n_max;       // Shapelets order
  int n_members;   // 0 = plain profile
  int member_off;  // offset into the member-factor array ([n_raw][n_members] per entry)
  int amp_off;     // Shapelets: offset into the amp-slot array
  int fwdmode;     // scaling-relation group differentiated in forward mode (3 dvars = the base scaling params)
  int comp_off;    // light profiles: index of the first linear component (lstsq stack channel)
  const float* table;  // Shapelets(interpolate=True): [n_max+1][6000] basis table (device / host pointer)
  int slot[GL_MAX_RAW];
  float constant[GL_MAX_RAW];
};

struct GlProgram {
  int n_lens, n_ll, n_sl;
  int n_prof;
  int der_total;   // floats per sample in the derived vector
  int g_total;     // floats per sample in the dvar-cotangent vector
  int n_params;    // P
  int depth;       // number of linear light components (lstsq)
  int has_fwdmode; // some scaling-relation group is differentiated in forward mode
  int comp_mask;   // which light groups gl_pix_image adds: bit 0 lens light, bit 1 source light (simulate_* variants)
  int scr_prof;    // lens entry whose series state the adjoint kernels carry from their forward sweep (first plain EPL), or -1
  float epl_tol;   // EPL series: terms below this are dropped (1e-12 = the reference's constant, epl.py:37)
  GlProf prof[GL_MAX_PROF];
};

// ---------------------------------------------------------------------------------------------
// per-sample: raw gather, prep, prep adjoint
// ---------------------------------------------------------------------------------------------
// Gather the raw parameters of entry `pr`, member m, sample b from params[P][bs].
template <class T, class TP>
GL_HD void gl_gather_raw(const GlProf& pr, const TP* params, int bs, int b, const float* member_factor, int m, T* raw) {
  const int nraw = gl_n_raw(pr.type);
#pragma unroll
  for (int k = 0; k < GL_MAX_RAW; ++k) {
    if (k < nraw) {
      T v = (pr.slot[k] >= 0) ? T(params[(size_t)pr.slot[k] * bs + b]) : T(pr.constant[k]);
      if (pr.n_members > 0) v *= T(member_factor[pr.member_off + k * pr.n_members + m]);
      raw[k] = v;
    } else {
      raw[k] = T(0);
    }
  }
}

// params -> derived vector of one sample.  epl_fmax: per-entry batch maximum of f, or null for
// per-sample trip counts.
template <class T, class TP>
GL_HD void gl_sample_prep(const GlProgram& P, const TP* params, int bs, int b, const float* member_factor,
                          const int* amp_slot, const float* epl_fmax, T* der) {
  for (int i = 0; i < P.n_prof; ++i) {
    const GlProf& pr = P.prof[i];
    if (pr.type == GLT_SHAPELETS) {
      T raw[GL_MAX_RAW];
      gl_gather_raw<T, TP>(pr, params, bs, b, member_factor, 0, raw);
      T* d = der + pr.der_off;
      shp_prep<T>(raw, d, pr.n_max);
      T* amp = d + shp_amp_off(pr.n_max);
      const int L = shp_layers(pr.n_max);
      const bool lstsq = (pr.flags & 1u) != 0;   // amplitudes are solved for; filled in after the solve
      for (int k = 0; k < L; ++k) amp[k] = lstsq ? T(0) : T(params[(size_t)amp_slot[pr.amp_off + k] * bs + b]);
      continue;
    }
    const int nm = pr.n_members > 0 ? pr.n_members : 1;
    for (int m = 0; m < nm; ++m) {
      T raw[GL_MAX_RAW];
      gl_gather_raw<T, TP>(pr, params, bs, b, member_factor, m, raw);
      T* dm = der + pr.der_off + m * pr.der_size;
      gl_prep<T>(pr.type, pr.flags, pr.niter, raw, dm, epl_fmax ? T(epl_fmax[i]) : T(-1), T(P.epl_tol));
      if (pr.n_members > 0 && pr.type == GLT_DPIE) {
        // M[j][k] = d(scale, rc, rt)_j / d(base theta_E, r_core, r_cut)_k for this member
        for (int jj = 0; jj < 3; ++jj) {
          T gu[GL_MAX_DVARS], graw[GL_MAX_RAW];
          for (int k = 0; k < GL_MAX_DVARS; ++k) gu[k] = T(0);
          for (int k = 0; k < GL_MAX_RAW; ++k) graw[k] = T(0);
          gu[DPG_SCALE + jj] = T(1);
          dpie_prep_bwd<T>(raw, dm, gu, graw, true);
          for (int k = 0; k < 3; ++k)
            dm[DP_M + 3 * jj + k] = graw[k] * T(member_factor[pr.member_off + k * pr.n_members + m]);
        }
        dm[DP_FAST] = T(0); dm[DP_ITHETA] = T(0);
      }
    }
    if (pr.fwdmode) {
      // sparse chain matrix for every member (r_core < r_cut, so the sort is the identity) and a non-zero base theta_E:
      // dpie_fwd_jac<FAST> applies (gl_pix_beta_jac reads the flag from member 0's block)
      T* d0 = der + pr.der_off;
      bool fast = true;
      for (int m = 0; m < nm; ++m) {
        const T* M = der + pr.der_off + m * pr.der_size + DP_M;
        fast = fast && M[3] == T(0) && M[5] == T(0) && M[6] == T(0) && M[7] == T(0);
      }
      const T theta = (pr.slot[0] >= 0) ? T(params[(size_t)pr.slot[0] * bs + b]) : T(pr.constant[0]);
      fast = fast && theta != T(0) && theta == theta;
      d0[DP_FAST] = fast ? T(1) : T(0);
      d0[DP_ITHETA] = fast ? T(1) / theta : T(0);
    }
  }
}

// dvar cotangents g[] of one sample -> gparams[P][bs] column b (overwritten).
template <class T, class TP>
GL_HD void gl_sample_prep_bwd(const GlProgram& P, const TP* params, int bs, int b, const float* member_factor,
                              const int* amp_slot, const T* der, const T* g, TP* gparams) {
  for (int k = 0; k < P.n_params; ++k) gparams[(size_t)k * bs + b] = TP(0);
  for (int i = 0; i < P.n_prof; ++i) {
    const GlProf& pr = P.prof[i];
    const int nraw = gl_n_raw(pr.type);
    if (pr.type == GLT_SHAPELETS) {
      T raw[GL_MAX_RAW];
      gl_gather_raw<T, TP>(pr, params, bs, b, member_factor, 0, raw);
      const T* gg = g + pr.g_off;   // cx, cy, invbeta, amplitudes
      T graw[3] = {-gg[SHPG_IB] / (raw[0] * raw[0]), gg[SHPG_CX], gg[SHPG_CY]};
      for (int k = 0; k < 3; ++k)
        if (pr.slot[k] >= 0) gparams[(size_t)pr.slot[k] * bs + b] += TP(graw[k]);
      if (!(pr.flags & 1u)) {
        const int L = shp_layers(pr.n_max);
        for (int k = 0; k < L; ++k) gparams[(size_t)amp_slot[pr.amp_off + k] * bs + b] += TP(gg[SHPG_AMP + k]);
      }
      continue;
    }
    const int nm = pr.n_members > 0 ? pr.n_members : 1;
    T gbase[GL_MAX_RAW];
    for (int k = 0; k < GL_MAX_RAW; ++k) gbase[k] = T(0);
    if (pr.fwdmode) {
      for (int k = 0; k < 3; ++k)
        if (pr.slot[k] >= 0) gparams[(size_t)pr.slot[k] * bs + b] += TP(g[pr.g_off + k]);
      continue;
    }
    for (int m = 0; m < nm; ++m) {
      T raw[GL_MAX_RAW], graw[GL_MAX_RAW];
      gl_gather_raw<T, TP>(pr, params, bs, b, member_factor, m, raw);
      for (int k = 0; k < GL_MAX_RAW; ++k) graw[k] = T(0);
      gl_prep_bwd<T>(pr.type, pr.flags, raw, der + pr.der_off + m * pr.der_size, g + pr.g_off + m * pr.n_dvars, graw);
      for (int k = 0; k < nraw; ++k) {
        T fac = pr.n_members > 0 ? T(member_factor[pr.member_off + k * pr.n_members + m]) : T(1);
        gbase[k] += graw[k] * fac;
      }
    }
    for (int k = 0; k < nraw; ++k)
      if (pr.slot[k] >= 0) gparams[(size_t)pr.slot[k] * bs + b] += TP(gbase[k]);
  }
}

// ---------------------------------------------------------------------------------------------
// per-pixel drivers (NP pixels of one sample at a time; der = derived vector of the sample)
// ---------------------------------------------------------------------------------------------
// beta = theta - sum_i alpha_i(theta)   (src/gigalens/tf/simulator.py:72-78)
template <class T, int NP, unsigned F>
GL_HD void gl_pix_beta(const GlProgram& P, const typename gl_scalar_of<T>::type* der, const T* x, const T* y, T* bx, T* by,
                       T* scr = nullptr, int scr_stride = 0) {
#pragma unroll
  for (int j = 0; j < NP; ++j) { bx[j] = x[j]; by[j] = y[j]; }
  for (int i = 0; i < P.n_lens; ++i) {
    const GlProf& pr = P.prof[i];
    if (pr.n_members > 0) {
      // a scaling-relation group is summed over its members first (scaling_relation.py:61-70 reduces over the member axis) and
      // then subtracted: the order every driver of this file uses, so that forward, taped forward and adjoint agree bit for bit
      T gx[NP], gy[NP];
#pragma unroll
      for (int j = 0; j < NP; ++j) { gx[j] = T(0); gy[j] = T(0); }
      for (int m = 0; m < pr.n_members; ++m) {
        T ax[NP], ay[NP];
        gl_lens_fwd<T, NP, F>(pr.type, pr.ts, der + pr.der_off + m * pr.der_size, x, y, ax, ay);
#pragma unroll
        for (int j = 0; j < NP; ++j) { gx[j] += ax[j]; gy[j] += ay[j]; }
      }
#pragma unroll
      for (int j = 0; j < NP; ++j) { bx[j] -= gx[j]; by[j] -= gy[j]; }
      continue;
    }
    T ax[NP], ay[NP];
    // the entry that owns the series scratch (P.scr_prof) parks its series state for the adjoint
    gl_lens_fwd<T, NP, F>(pr.type, pr.ts, der + pr.der_off, x, y, ax, ay, (i == P.scr_prof) ? scr : (T*)nullptr, scr_stride);
#pragma unroll
    for (int j = 0; j < NP; ++j) { bx[j] -= ax[j]; by[j] -= ay[j]; }
  }
}

// beta as above, plus the 2x3 Jacobian of the forward-mode group's deflection w.r.t. its base params
template <class T, int NP, unsigned F>
GL_HD void gl_pix_beta_jac(const GlProgram& P, const typename gl_scalar_of<T>::type* der, const T* x, const T* y, T* bx, T* by,
                           T (*Jx)[NP], T (*Jy)[NP]) {
#pragma unroll
  for (int j = 0; j < NP; ++j) { bx[j] = x[j]; by[j] = y[j]; }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < NP; ++j) { Jx[k][j] = T(0); Jy[k][j] = T(0); }
  for (int i = 0; i < P.n_lens; ++i) {
    const GlProf& pr = P.prof[i];
    const int nm = pr.n_members > 0 ? pr.n_members : 1;
    if constexpr ((F & GLF_DPIE) != 0) {
      if (pr.fwdmode && der[pr.der_off + DP_FAST] != typename gl_scalar_of<T>::type(0)) {
        // sparse chain matrix: accumulate the group deflection and Jacobian columns 1, 2; column 0 = deflection / theta_E
        T gx[NP], gy[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) { gx[j] = T(0); gy[j] = T(0); }
        for (int m = 0; m < nm; ++m) {
          T ax[NP], ay[NP];
          dpie_fwd_jac_fast<T, NP>(der + pr.der_off + m * pr.der_size, x, y, ax, ay, Jx, Jy);
#pragma unroll
          for (int j = 0; j < NP; ++j) { gx[j] += ax[j]; gy[j] += ay[j]; }
        }
        const T ith = T(der[pr.der_off + DP_ITHETA]);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          bx[j] -= gx[j]; by[j] -= gy[j];
          Jx[0][j] += gx[j] * ith; Jy[0][j] += gy[j] * ith;
        }
        continue;
      }
    }
    T gx[NP], gy[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) { gx[j] = T(0); gy[j] = T(0); }
    for (int m = 0; m < nm; ++m) {
      T ax[NP], ay[NP];
      bool done = false;
      if constexpr ((F & GLF_DPIE) != 0) {
        if (pr.fwdmode) { dpie_fwd_jac<T, NP>(der + pr.der_off + m * pr.der_size, x, y, ax, ay, Jx, Jy); done = true; }
      }
      if (!done) gl_lens_fwd<T, NP, F>(pr.type, pr.ts, der + pr.der_off + m * pr.der_size, x, y, ax, ay);
#pragma unroll
      for (int j = 0; j < NP; ++j) { gx[j] += ax[j]; gy[j] += ay[j]; }
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) { bx[j] -= gx[j]; by[j] -= gy[j]; }
  }
}

// Supersampled surface brightness before the NaN scrub (tf/simulator.py:124-138), non-lstsq mode.
template <class T, int NP, unsigned F>
GL_HD void gl_pix_image(const GlProgram& P, const typename gl_scalar_of<T>::type* der, const T* x, const T* y, bool no_deflection, T* out) {
  T bx[NP], by[NP];
  if (no_deflection) {
#pragma unroll
    for (int j = 0; j < NP; ++j) { bx[j] = x[j]; by[j] = y[j]; }
  } else {
    gl_pix_beta<T, NP, F>(P, der, x, y, bx, by);
  }
#pragma unroll
  for (int j = 0; j < NP; ++j) out[j] = T(0);
  for (int i = P.n_lens; i < P.n_prof; ++i) {
    const GlProf& pr = P.prof[i];
    const bool src = i >= P.n_lens + P.n_ll;
    if (!(P.comp_mask & (src ? 2 : 1))) continue;
    const T* px = src ? bx : x;
    const T* py = src ? by : y;
    switch (pr.type) {
      case GLT_SERSIC: case GLT_SERSIC_ELLIPSE:
        if constexpr ((F & GLF_SERSIC) != 0) sersic_fwd<T, NP>(der + pr.der_off, px, py, out);
        break;
      case GLT_CORE_SERSIC:
        if constexpr ((F & GLF_CORESERSIC) != 0) core_sersic_fwd<T, NP>(der + pr.der_off, px, py, out);
        break;
      case GLT_SHAPELETS:
        if constexpr ((F & GLF_SHAPELETS) != 0)
        for (int j = 0; j < NP; ++j)
          out[j] += shp_point<T>(der + pr.der_off, pr.table, (pr.flags & 2u) != 0, pr.n_max, px[j], py[j], (T*)nullptr, 0,
                                 (const T*)nullptr, (T*)nullptr, (T*)nullptr, (T*)nullptr, (T*)nullptr);
        break;
      default: break;
    }
  }
}

// gl_pix_image for programs with a forward-mode group, keeping what the adjoint needs ("tape"): beta and the 2 x 3 Jacobian of the
// group deflection w.r.t. its base parameters.  With the tape in HBM the adjoint kernel never walks the member loop again (at the
// cluster config the loop is 84 % of both ray-tracing kernels; 8 floats per ss pixel cost 1.6 ms of HBM time per 1024 samples).
template <class T, int NP, unsigned F>
GL_HD void gl_pix_image_tape(const GlProgram& P, const typename gl_scalar_of<T>::type* der, const T* x, const T* y, T* out,
                             T* bx, T* by, T (*Jx)[NP], T (*Jy)[NP]) {
  gl_pix_beta_jac<T, NP, F>(P, der, x, y, bx, by, Jx, Jy);
#pragma unroll
  for (int j = 0; j < NP; ++j) out[j] = T(0);
  for (int i = P.n_lens; i < P.n_prof; ++i) {
    const GlProf& pr = P.prof[i];
    const bool src = i >= P.n_lens + P.n_ll;
    if (!(P.comp_mask & (src ? 2 : 1))) continue;
    if (pr.type == GLT_SERSIC || pr.type == GLT_SERSIC_ELLIPSE) {
      if constexpr ((F & GLF_SERSIC) != 0) sersic_fwd<T, NP>(der + pr.der_off, src ? bx : x, src ? by : y, out);
    }
  }
}

// The D unit-amplitude linear light components at one pixel (lstsq stack before the convolution,
// tf/simulator.py:183-200 with the layout of jax/simulator.py:171-175): out[c * stride], NaN scrubbed
// (:200); `keep` = false (pixel outside pix_region) writes zeros.
template <class T, unsigned F>
GL_HD int gl_point_components(const GlProgram& P, const T* der, T x, T y, T bx, T by, T* out, int stride, bool keep) {
  int n_nan = 0;   // values scrubbed at this pixel (the adjoint passes them no gradient)
  for (int i = P.n_lens; i < P.n_prof; ++i) {
    const GlProf& pr = P.prof[i];
    const bool src = i >= P.n_lens + P.n_ll;
    const T px = src ? bx : x, py = src ? by : y;
    switch (pr.type) {
      case GLT_SERSIC: case GLT_SERSIC_ELLIPSE: if constexpr ((F & GLF_SERSIC) != 0) {
        T v[1] = {T(0)}, xx[1] = {px}, yy[1] = {py};
        sersic_fwd<T, 1>(der + pr.der_off, xx, yy, v);
        if (keep && gl_isnan(v[0])) ++n_nan;
        out[pr.comp_off * stride] = (keep && !gl_isnan(v[0])) ? v[0] : T(0);
      } break;
      case GLT_CORE_SERSIC: if constexpr ((F & GLF_CORESERSIC) != 0) {
        T v[1] = {T(0)}, xx[1] = {px}, yy[1] = {py};
        core_sersic_fwd<T, 1>(der + pr.der_off, xx, yy, v);
        if (keep && gl_isnan(v[0])) ++n_nan;
        out[pr.comp_off * stride] = (keep && !gl_isnan(v[0])) ? v[0] : T(0);
      } break;
      case GLT_SHAPELETS: if constexpr ((F & GLF_SHAPELETS) != 0) {
        T* o = out + (size_t)pr.comp_off * stride;
        shp_point<T>(der + pr.der_off, pr.table, (pr.flags & 2u) != 0, pr.n_max, px, py, o, stride, (const T*)nullptr,
                     (T*)nullptr, (T*)nullptr, (T*)nullptr, (T*)nullptr, keep, &n_nan);   // scrubs and counts as it stores
      } break;
      default: break;
    }
  }
  return n_nan;
}

// Adjoint of gl_pix_image: gS is the cotangent of the (scrubbed) surface brightness at the NP
// pixels.  `flush(acc, n, off)` receives the NP-pixel partial cotangent of dvars [off, off+n) --
// the host harness adds it into a vector, the CUDA kernel warp-reduces it into shared memory.
// SCRUB (lstsq path, samples that had NaN components): the reference scrubs the component stack per component
// (tf/simulator.py:200), so a light profile whose own value is NaN at a pixel receives a zero cotangent there while
// the other profiles keep theirs.
// TAPE: beta and the forward-mode Jacobian come from the forward kernel's tape (gl_pix_image_tape) instead of being recomputed.
template <class T, int NP, unsigned F, class Flush, bool SCRUB = false, bool TAPE = false>
GL_HD void gl_pix_image_bwd(const GlProgram& P, const typename gl_scalar_of<T>::type* der, const T* x, const T* y, const T* gS,
                            bool no_deflection, Flush& flush, T* scr = nullptr, int scr_stride = 0,
                            const T* tape_bx = nullptr, const T* tape_by = nullptr, const T (*tape_Jx)[NP] = nullptr,
                            const T (*tape_Jy)[NP] = nullptr) {
  T bx[NP], by[NP], Gx[NP], Gy[NP];
  T Jx[3][NP], Jy[3][NP];
  bool have_jac = false;
  const T* const gS_all = gS;   // SCRUB re-points gS at a per-profile masked copy
  (void)gS_all;
  if constexpr (TAPE) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      bx[j] = tape_bx[j]; by[j] = tape_by[j];
#pragma unroll
      for (int k = 0; k < 3; ++k) { Jx[k][j] = tape_Jx[k][j]; Jy[k][j] = tape_Jy[k][j]; }
    }
    have_jac = true; scr = nullptr;
  } else
  if (no_deflection) {
#pragma unroll
    for (int j = 0; j < NP; ++j) { bx[j] = x[j]; by[j] = y[j]; }
  } else {
    if constexpr ((F & GLF_DPIE) != 0) {
      if (P.has_fwdmode) { gl_pix_beta_jac<T, NP, F>(P, der, x, y, bx, by, Jx, Jy); have_jac = true; }
    }
    if (have_jac) scr = nullptr;
    else gl_pix_beta<T, NP, F>(P, der, x, y, bx, by, scr, scr_stride);
  }
#pragma unroll
  for (int j = 0; j < NP; ++j) { Gx[j] = T(0); Gy[j] = T(0); }
  for (int i = P.n_lens; i < P.n_prof; ++i) {
    const GlProf& pr = P.prof[i];
    const bool src = i >= P.n_lens + P.n_ll;
    T acc[GL_MAX_DVARS];
#pragma unroll
    for (int k = 0; k < GL_MAX_DVARS; ++k) acc[k] = T(0);
    T gM[NP];
    if constexpr (SCRUB) {   // this profile's own forward value decides which pixels pass a cotangent to it
      T v[NP];
#pragma unroll
      for (int j = 0; j < NP; ++j) v[j] = T(0);
      if (pr.type == GLT_SHAPELETS) {
        if constexpr ((F & GLF_SHAPELETS) != 0)
          for (int j = 0; j < NP; ++j)
            v[j] = shp_point<T>(der + pr.der_off, pr.table, (pr.flags & 2u) != 0, pr.n_max, src ? bx[j] : x[j], src ? by[j] : y[j],
                                (T*)nullptr, 0, (const T*)nullptr, (T*)nullptr, (T*)nullptr, (T*)nullptr, (T*)nullptr);
      } else if (pr.type == GLT_CORE_SERSIC) {
        if constexpr ((F & GLF_CORESERSIC) != 0) core_sersic_fwd<T, NP>(der + pr.der_off, src ? bx : x, src ? by : y, v);
      } else {
        if constexpr ((F & GLF_SERSIC) != 0) sersic_fwd<T, NP>(der + pr.der_off, src ? bx : x, src ? by : y, v);
      }
#pragma unroll
      for (int j = 0; j < NP; ++j) gM[j] = gl_isnan(v[j]) ? T(0) : gS_all[j];
      gS = gM;
    }
    switch (pr.type) {
      case GLT_SERSIC: case GLT_SERSIC_ELLIPSE:
        if constexpr ((F & GLF_SERSIC) != 0) {
          if (src) sersic_bwd<T, NP>(der + pr.der_off, bx, by, gS, acc, Gx, Gy);
          else sersic_bwd<T, NP>(der + pr.der_off, x, y, gS, acc, (T*)nullptr, (T*)nullptr);
        }
        break;
      case GLT_CORE_SERSIC: if constexpr ((F & GLF_CORESERSIC) != 0) {
        T a9[GL_MAX_DVARS];      // the 9th dvar (Ie) travels in its own flush
#pragma unroll
        for (int k = 0; k < GL_MAX_DVARS; ++k) a9[k] = T(0);
        if (src) core_sersic_bwd<T, NP>(der + pr.der_off, bx, by, gS, acc, a9, Gx, Gy);
        else core_sersic_bwd<T, NP>(der + pr.der_off, x, y, gS, acc, a9, (T*)nullptr, (T*)nullptr);
        flush(a9, 1, pr.g_off + CSG_IE);
      } break;
      case GLT_SHAPELETS: if constexpr ((F & GLF_SHAPELETS) != 0) {
        const bool want_amp = !(pr.flags & 1u);
        const int L = shp_layers(pr.n_max);
        T gamp[(GL_SHP_MAXN + 1) * (GL_SHP_MAXN + 2) / 2];
        if (want_amp) for (int k = 0; k < L; ++k) gamp[k] = T(0);
        for (int j = 0; j < NP; ++j)
          shp_point<T>(der + pr.der_off, pr.table, (pr.flags & 2u) != 0, pr.n_max, src ? bx[j] : x[j], src ? by[j] : y[j],
                       (T*)nullptr, 0, gS + j, acc, want_amp ? gamp : (T*)nullptr, src ? Gx + j : (T*)nullptr,
                       src ? Gy + j : (T*)nullptr);
        if (want_amp) {
          for (int k0 = 0; k0 < L; k0 += GL_MAX_DVARS) {
            T a8[GL_MAX_DVARS];
#pragma unroll
            for (int k = 0; k < GL_MAX_DVARS; ++k) a8[k] = (k0 + k < L) ? gamp[k0 + k] : T(0);
            flush(a8, (L - k0) < GL_MAX_DVARS ? (L - k0) : GL_MAX_DVARS, pr.g_off + SHPG_AMP + k0);
          }
        }
      } break;
      default: break;
    }
    flush(acc, pr.type == GLT_SHAPELETS ? 3 : (pr.n_dvars < GL_MAX_DVARS ? pr.n_dvars : GL_MAX_DVARS), pr.g_off);
  }
  if (no_deflection) return;
  // d(beta)/d(lens) = -d(alpha): cotangent of each deflection is (-Gx, -Gy)
#pragma unroll
  for (int j = 0; j < NP; ++j) { Gx[j] = -Gx[j]; Gy[j] = -Gy[j]; }
  for (int i = 0; i < P.n_lens; ++i) {
    const GlProf& pr = P.prof[i];
    if constexpr ((F & GLF_DPIE) != 0) {
      if (have_jac && pr.fwdmode) {   // forward-mode group: contract the carried Jacobian with the cotangent
        T acc[GL_MAX_DVARS];
#pragma unroll
        for (int k = 0; k < GL_MAX_DVARS; ++k) acc[k] = T(0);
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int j = 0; j < NP; ++j) acc[k] += gl_fma(Gx[j], Jx[k][j], Gy[j] * Jy[k][j]);
        flush(acc, 3, pr.g_off);
        continue;
      }
    }
    const int nm = pr.n_members > 0 ? pr.n_members : 1;
    for (int m = 0; m < nm; ++m) {
      T acc[GL_MAX_DVARS];
#pragma unroll
      for (int k = 0; k < GL_MAX_DVARS; ++k) acc[k] = T(0);
      gl_lens_bwd<T, NP, F>(pr.type, pr.ts, der + pr.der_off + m * pr.der_size, x, y, Gx, Gy, acc, (i == P.scr_prof) ? scr : (const T*)nullptr, scr_stride);
      flush(acc, pr.n_dvars, pr.g_off + m * pr.n_dvars);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Straight-line drivers for the GIGA-Lens benchmark shape (tf-demo.ipynb / BASELINE configs[1]):
//   lenses [EPL, Shear], lens light [Sersic | SersicEllipse], source light [Sersic | SersicEllipse].
// Same building blocks and the same order of floating-point operations as the interpreting drivers above, but no
// profile loop, no type switch and compile-time flush widths, so the whole pixel body is a few large basic blocks
// the scheduler can interleave (the interpreter's loop control, cotangent zeroing and descriptor loads were ~10 % of
// the adjoint kernel's instructions).  Every other program runs the interpreter.
// ---------------------------------------------------------------------------------------------
GL_HD bool gl_is_benchmark_shape(const GlProgram& P) {
  auto ser = [](int t) { return t == GLT_SERSIC || t == GLT_SERSIC_ELLIPSE; };
  return P.n_lens == 2 && P.n_ll == 1 && P.n_sl == 1 && P.prof[0].type == GLT_EPL && P.prof[0].n_members == 0 &&
         P.prof[1].type == GLT_SHEAR && P.prof[1].n_members == 0 && ser(P.prof[2].type) && ser(P.prof[3].type) &&
         P.comp_mask == 3 && P.scr_prof == 0;
}
template <class T, int NP>
GL_HD void gl_pix_image_bs(const GlProgram& P, const typename gl_scalar_of<T>::type* der, const T* x, const T* y, T* out) {
  T ax[NP], ay[NP], sx[NP], sy[NP], bx[NP], by[NP];
  epl_fwd<T, NP>(der + P.prof[0].der_off, P.prof[0].ts, x, y, ax, ay);
  shear_fwd<T, NP>(der + P.prof[1].der_off, x, y, sx, sy);
#pragma unroll
  for (int j = 0; j < NP; ++j) { bx[j] = (x[j] - ax[j]) - sx[j]; by[j] = (y[j] - ay[j]) - sy[j]; out[j] = T(0); }
  sersic_fwd<T, NP>(der + P.prof[2].der_off, x, y, out);
  sersic_fwd<T, NP>(der + P.prof[3].der_off, bx, by, out);
}
template <class T, int NP, class Flush>
GL_HD void gl_pix_image_bwd_bs(const GlProgram& P, const typename gl_scalar_of<T>::type* der, const T* x, const T* y, const T* gS,
                               Flush& flush, T* scr, int scr_stride) {
  T ax[NP], ay[NP], sx[NP], sy[NP], bx[NP], by[NP], Gx[NP], Gy[NP];
  epl_fwd_save<T, NP>(der + P.prof[0].der_off, P.prof[0].ts, x, y, ax, ay, scr, scr_stride);
  shear_fwd<T, NP>(der + P.prof[1].der_off, x, y, sx, sy);
#pragma unroll
  for (int j = 0; j < NP; ++j) { bx[j] = (x[j] - ax[j]) - sx[j]; by[j] = (y[j] - ay[j]) - sy[j]; Gx[j] = T(0); Gy[j] = T(0); }
  T acc[GL_MAX_DVARS];
#pragma unroll
  for (int k = 0; k < GL_MAX_DVARS; ++k) acc[k] = T(0);
  sersic_bwd<T, NP>(der + P.prof[2].der_off, x, y, gS, acc, (T*)nullptr, (T*)nullptr);
  flush(acc, 8, P.prof[2].g_off);
#pragma unroll
  for (int k = 0; k < GL_MAX_DVARS; ++k) acc[k] = T(0);
  sersic_bwd<T, NP>(der + P.prof[3].der_off, bx, by, gS, acc, Gx, Gy);
  flush(acc, 8, P.prof[3].g_off);
#pragma unroll
  for (int j = 0; j < NP; ++j) { Gx[j] = -Gx[j]; Gy[j] = -Gy[j]; }
#pragma unroll
  for (int k = 0; k < GL_MAX_DVARS; ++k) acc[k] = T(0);
  epl_bwd_load<T, NP>(der + P.prof[0].der_off, P.prof[0].ts, x, y, Gx, Gy, acc, scr, scr_stride);
  flush(acc, 8, P.prof[0].g_off);
#pragma unroll
  for (int k = 0; k < GL_MAX_DVARS; ++k) acc[k] = T(0);
  shear_bwd<T, NP>(der + P.prof[1].der_off, x, y, Gx, Gy, acc);
  flush(acc, 2, P.prof[1].g_off);
}

// ---------------------------------------------------------------------------------------------
// image-position likelihood (ForwardProbModel.stats_positions, src/gigalens/tf/model.py:103-124):
// point drivers on the forward-mode lane type GlDual.  Tens of points per sample -- not a hot path,
// but it needs d(beta)/d(theta) (the lensing Hessian, LensSimulator.magnification,
// tf/simulator.py:80-91) and the parameter gradient of a function of that Hessian.
// ---------------------------------------------------------------------------------------------
// beta and the Hessian H = (f_xx, f_xy, f_yx, f_yy) = (d ax/dx, d ax/dy, d ay/dx, d ay/dy) of the summed
// deflection at one point: two forward-mode sweeps (tf/profile.py:9-30 does the same with a tape).
template <class S, unsigned F>
GL_HD void gl_point_hessian(const GlProgram& P, const S* der, S x, S y, S& bx, S& by, S* H) {
  typedef GlDual<S> D;
  D X[1], Y[1], BX[1], BY[1];
  X[0] = D(x, S(1)); Y[0] = D(y, S(0));
  gl_pix_beta<D, 1, F>(P, der, X, Y, BX, BY);
  bx = BX[0].v; by = BY[0].v;
  H[0] = S(1) - BX[0].d; H[2] = -BY[0].d;
  X[0] = D(x, S(0)); Y[0] = D(y, S(1));
  gl_pix_beta<D, 1, F>(P, der, X, Y, BX, BY);
  H[1] = -BX[0].d; H[3] = S(1) - BY[0].d;
}

// One multiply-imaged source system with n images (tf/model.py:107-122): source-plane scatter about
// the barycentre, errors err = centroid_error / magnification.  Adds to chi2 / norm and writes the
// cotangents of log_like = -(chi2 + norm)/2 w.r.t. beta and H of every image.
template <class S>
GL_HD void gl_positions_system(int n, const S* bx, const S* by, const S* H, const S* ex, const S* ey, S& chi2, S& norm,
                               S* gbx, S* gby, S* gH) {
  const S two_pi = S(6.283185307179586);
  S mx = S(0), my = S(0);
  for (int i = 0; i < n; ++i) { mx += bx[i]; my += by[i]; }
  mx /= S(n); my /= S(n);
  S sax = S(0), say = S(0);
  for (int i = 0; i < n; ++i) {
    const S* h = H + 4 * i;
    const S det = (S(1) - h[0]) * (S(1) - h[3]) - h[1] * h[2];
    const S mu = S(1) / det;
    const S dx = bx[i] - mx, dy = by[i] - my;
    const S errx = ex[i] / mu, erry = ey[i] / mu;
    const S rx = dx / errx, ry = dy / erry;
    chi2 += rx * rx + ry * ry;
    norm += gl_log(two_pi * errx * errx) + gl_log(two_pi * erry * erry);
    if (gbx) {
      const S ax = -rx / errx, ay = -ry / erry;            // d log_like / d (beta - barycentre)
      gbx[i] = ax; gby[i] = ay; sax += ax; say += ay;
      const S gmu = -(rx * rx + ry * ry) / mu + S(2) / mu;   // chi2 ~ mu^2, norm ~ -4 log|mu|
      const S gdet = -gmu * mu * mu;
      gH[4 * i + 0] = -gdet * (S(1) - h[3]);
      gH[4 * i + 3] = -gdet * (S(1) - h[0]);
      gH[4 * i + 1] = -gdet * h[2];
      gH[4 * i + 2] = -gdet * h[1];
    }
  }
  if (gbx) for (int i = 0; i < n; ++i) { gbx[i] -= sax / S(n); gby[i] -= say / S(n); }
}

// Parameter gradient of one image point: (gbx, gby) = cotangent of beta, gH = cotangent of H.  Runs
// the hand adjoint of every deflector on dual numbers along x and along y; the tangent of the adjoint
// output is   d^2 alpha/(d theta_dir d p)^T . gH[., dir]  +  d alpha/d p ^T . (-g_beta / 2)   per sweep.
// flush(acc, n, off) receives scalar partial cotangents of dvars [off, off+n), as in gl_pix_image_bwd.
template <class S, unsigned F, class Flush>
GL_HD void gl_point_positions_bwd(const GlProgram& P, const S* der, S x, S y, S gbx, S gby, const S* gH, Flush& flush) {
  typedef GlDual<S> D;
  for (int dir = 0; dir < 2; ++dir) {
    D X[1], Y[1], GX[1], GY[1];
    X[0] = D(x, dir == 0 ? S(1) : S(0)); Y[0] = D(y, dir == 1 ? S(1) : S(0));
    GX[0] = D(gH[dir], S(-0.5) * gbx);
    GY[0] = D(gH[2 + dir], S(-0.5) * gby);
    for (int i = 0; i < P.n_lens; ++i) {
      const GlProf& pr = P.prof[i];
      const int nm = pr.n_members > 0 ? pr.n_members : 1;
      S base[GL_MAX_DVARS];
#pragma unroll
      for (int k = 0; k < GL_MAX_DVARS; ++k) base[k] = S(0);
      for (int m = 0; m < nm; ++m) {
        D acc[GL_MAX_DVARS];
#pragma unroll
        for (int k = 0; k < GL_MAX_DVARS; ++k) acc[k] = D(S(0));
        const S* dm = der + pr.der_off + m * pr.der_size;
        gl_lens_bwd<D, 1, F>(pr.type, pr.ts, dm, X, Y, GX, GY, acc);
        if (pr.fwdmode) {   // scaling-relation member: chain (scale, rc, rt) -> the group's three base parameters
          const S* M = dm + DP_M;
          for (int jj = 0; jj < 3; ++jj)
            for (int k = 0; k < 3; ++k) base[k] += acc[DPG_SCALE + jj].d * M[3 * jj + k];
        } else {
          S a[GL_MAX_DVARS];
#pragma unroll
          for (int k = 0; k < GL_MAX_DVARS; ++k) a[k] = acc[k].d;
          flush(a, pr.n_dvars, pr.g_off + m * pr.n_dvars);
        }
      }
      if (pr.fwdmode) flush(base, 3, pr.g_off);
    }
  }
}
