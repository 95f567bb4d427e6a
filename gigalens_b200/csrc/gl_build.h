// gl_build.h -- host-only: validate a gl_model_desc and flatten it into a GlProgram.
// Used by the CUDA library (gl_plan_create) and by the test-only host harness.
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "../../include/gigalens_b200.h"
#include "gl_program.h"

struct GlBuilt {
  GlProgram prog;
  std::vector<float> member_factor;  // concatenated [n_raw][n_members] blocks
  std::vector<int> amp_slot;         // concatenated Shapelets amplitude slots
  std::vector<float> tables;         // concatenated Shapelets interpolation tables
  std::vector<int> table_off;        // per profile: offset into tables, or -1
};

// lenstronomy Shapelets.phi_n on linspace(-5, 5, 6000), cast to fp32 (shapelets.py:39-40,50-51)
inline void gl_shapelets_table(int n_max, std::vector<float>& out) {
  const int N = GL_SHP_TABLE_N;
  const size_t base = out.size();
  out.resize(base + (size_t)(n_max + 1) * N);
  for (int i = 0; i < N; ++i) {
    const double x = -5.0 + 10.0 * (double)i / (double)(N - 1);
    double hm2 = 0.0, hm1 = 0.75112554446494248;   // pi^(-1/4)
    const double gauss = std::exp(-x * x / 2.0);
    out[base + i] = (float)(hm1 * gauss);
    for (int n = 1; n <= n_max; ++n) {
      const double hn = std::sqrt(2.0 / n) * x * hm1 - std::sqrt((n - 1.0) / n) * hm2;
      out[base + (size_t)n * N + i] = (float)(hn * gauss);
      hm2 = hm1; hm1 = hn;
    }
  }
}

inline int gl_shapelets_layers(int n_max) { return (n_max + 1) * (n_max + 2) / 2; }

inline bool gl_is_mass(int t) { return t >= GLT_EPL && t <= GLT_DPIEP; }
inline bool gl_is_light(int t) { return t == GLT_SERSIC || t == GLT_SERSIC_ELLIPSE || t == GLT_SHAPELETS || t == GLT_CORE_SERSIC; }

// Returns "" on success, else an error message.
inline std::string gl_build_program(const gl_model_desc* m, GlBuilt& out, bool use_fwdmode = true) {
  if (!m) return "model descriptor is NULL";
  GlProgram& P = out.prog;
  P = GlProgram();
  P.n_lens = m->n_lens; P.n_ll = m->n_lens_light; P.n_sl = m->n_source_light;
  P.n_prof = P.n_lens + P.n_ll + P.n_sl;
  P.n_params = m->n_params;
  if (P.n_lens < 0 || P.n_ll < 0 || P.n_sl < 0) return "negative profile count";
  if (P.n_prof > GL_MAX_PROF) return "too many profiles (max " + std::to_string(GL_MAX_PROF) + ")";
  if (m->n_params < 0) return "negative n_params";
  int der = 0, g = 0, depth = 0;
  P.scr_prof = -1;
  P.epl_tol = 1e-12f;
  for (int i = 0; i < P.n_prof; ++i) {
    const gl_profile_desc* src = i < P.n_lens ? &m->lens[i]
                                 : i < P.n_lens + P.n_ll ? &m->lens_light[i - P.n_lens]
                                                         : &m->source_light[i - P.n_lens - P.n_ll];
    GlProf& pr = P.prof[i];
    pr.type = src->type;
    pr.flags = src->flags;
    const bool is_lens = i < P.n_lens;
    if (is_lens && !gl_is_mass(pr.type)) return "profile " + std::to_string(i) + ": not a mass profile type";
    if (!is_lens && !gl_is_light(pr.type)) return "profile " + std::to_string(i) + ": not a light profile type";
    pr.niter = src->niter > 0 ? src->niter : 50;
    if (pr.type == GLT_EPL && pr.niter > 200) return "EPL niter cap above 200 is not supported";
    pr.ts = epl_table_stride(pr.niter);
    if (pr.type == GLT_EPL && src->n_members == 0 && is_lens && P.scr_prof < 0) P.scr_prof = i;
    pr.n_max = src->n_max;
    pr.n_members = src->n_members;
    if (pr.n_members < 0) return "negative n_members";
    if (pr.n_members > 0 && !is_lens) return "scaling-relation sums are only defined for mass profiles";
    const int nraw = gl_n_raw(pr.type);
    for (int k = 0; k < GL_MAX_RAW; ++k) {
      pr.slot[k] = k < nraw ? src->slot[k] : -1;
      pr.constant[k] = k < nraw ? src->constant[k] : 0.f;
      if (pr.slot[k] >= m->n_params) return "profile " + std::to_string(i) + ": slot out of range";
    }
    // a light profile fitted by least squares has no amplitude parameter (profile.py:36-41)
    if ((pr.flags & GL_FLAG_USE_LSTSQ) && (pr.type == GLT_SERSIC || pr.type == GLT_SERSIC_ELLIPSE || pr.type == GLT_CORE_SERSIC)) {
      pr.slot[nraw - 1] = -1; pr.constant[nraw - 1] = 1.f;
    }
    pr.member_off = (int)out.member_factor.size();
    if (pr.n_members > 0) {
      if (!src->member_factor) return "n_members > 0 but member_factor is NULL";
      out.member_factor.insert(out.member_factor.end(), src->member_factor, src->member_factor + (size_t)nraw * pr.n_members);
    }
    pr.amp_off = (int)out.amp_slot.size();
    out.table_off.push_back(-1);
    pr.table = nullptr;
    pr.comp_off = depth;
    if (pr.type == GLT_SHAPELETS) {
      if (pr.n_max < 0 || pr.n_max > GL_SHP_MAXN) return "Shapelets n_max out of range [0, " + std::to_string(GL_SHP_MAXN) + "]";
      if (pr.flags & GL_FLAG_INTERPOLATE) { out.table_off.back() = (int)out.tables.size(); gl_shapelets_table(pr.n_max, out.tables); }
      const int nl = gl_shapelets_layers(pr.n_max);
      if (!(pr.flags & GL_FLAG_USE_LSTSQ)) {
        if (!src->amp_slot) return "Shapelets without use_lstsq needs amp_slot";
        for (int k = 0; k < nl; ++k) {
          if (src->amp_slot[k] < 0 || src->amp_slot[k] >= m->n_params) return "Shapelets amp_slot out of range";
          out.amp_slot.push_back(src->amp_slot[k]);
        }
      }
      depth += nl;
    } else if (!is_lens) {
      depth += 1;
    }
    pr.der_size = gl_der_size(pr.type, pr.niter, pr.n_max);
    pr.n_dvars = gl_n_dvars(pr.type);
    if (pr.type == GLT_SHAPELETS && !(pr.flags & GL_FLAG_USE_LSTSQ)) pr.n_dvars += gl_shapelets_layers(pr.n_max);
    pr.der_off = der;
    pr.g_off = g;
    const int nm = pr.n_members > 0 ? pr.n_members : 1;
    der += pr.der_size * nm;
    // The first dPIE scaling-relation group is differentiated in forward mode (gl_math.cuh dpie_fwd_jac):
    // its cotangent block is just the three base scaling parameters.
    pr.fwdmode = 0;
    if (pr.n_members > 0 && pr.type == GLT_DPIE && !P.has_fwdmode && use_fwdmode) { pr.fwdmode = 1; P.has_fwdmode = 1; }
    g += pr.fwdmode ? 3 : pr.n_dvars * nm;
  }
  P.der_total = (der + 3) & ~3;
  P.g_total = g;
  P.depth = depth;
  P.comp_mask = 3;
  return "";
}
