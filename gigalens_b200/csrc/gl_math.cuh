// gl_math.cuh -- per-sample and per-pixel arithmetic of the forward model and its hand adjoint.
//
// Everything here is a template over the scalar type and is `__host__ __device__`, so the very
// same source is (a) instantiated in fp32 inside the sm_100a kernels and (b) compiled by g++ in
// fp32/fp64 into the test-only host harness (tests/hostcheck) that checks every formula and every
// adjoint against the oracle's autograd without a GPU.  The product never calls the host build.
//
// Structure of one profile type X:
//   X_prep     raw params (one sample)            -> derived block d[] (constants + tables)
//   X_fwd      d[], NP pixel coordinates           -> deflection / surface brightness
//   X_bwd      d[], coordinates, output cotangent  -> += g[] (cotangent of the "dvars" of d[])
//                                                     (+ coordinate cotangent for source light)
//   X_prep_bwd raw, d[], g[]                       -> graw[] (cotangent of the raw params)
// The per-pixel kernels never see raw parameters: clamps, `where`s and trigonometry of the
// parameter conversion run once per sample in X_prep and their gradient conventions
//  are applied once per sample in X_prep_bwd.
#pragma once

#include <math.h>

#ifdef __CUDACC__
#define GL_HD __host__ __device__ __forceinline__
#else
#define GL_HD inline
#endif

// ---------------------------------------------------------------------------------------------
// scalar math wrappers (fp32 uses the accurate CUDA libm: results must hold 1e-5 vs the oracle)
// ---------------------------------------------------------------------------------------------
GL_HD float gl_sqrt(float x) { return sqrtf(x); }
GL_HD double gl_sqrt(double x) { return sqrt(x); }
GL_HD float gl_exp(float x) { return expf(x); }
GL_HD double gl_exp(double x) { return exp(x); }
GL_HD float gl_log(float x) { return logf(x); }
GL_HD double gl_log(double x) { return log(x); }
GL_HD float gl_pow(float x, float y) { return powf(x, y); }
GL_HD double gl_pow(double x, double y) { return pow(x, y); }
GL_HD float gl_atan2(float y, float x) { return atan2f(y, x); }
GL_HD double gl_atan2(double y, double x) { return atan2(y, x); }
GL_HD float gl_atan(float x) { return atanf(x); }
GL_HD double gl_atan(double x) { return atan(x); }
GL_HD float gl_atanh(float x) { return atanhf(x); }
GL_HD double gl_atanh(double x) { return atanh(x); }
GL_HD float gl_acosh(float x) { return acoshf(x); }
GL_HD double gl_acosh(double x) { return acosh(x); }
GL_HD float gl_acos(float x) { return acosf(x); }
GL_HD double gl_acos(double x) { return acos(x); }
GL_HD float gl_cos(float x) { return cosf(x); }
GL_HD double gl_cos(double x) { return cos(x); }
GL_HD float gl_sin(float x) { return sinf(x); }
GL_HD double gl_sin(double x) { return sin(x); }
GL_HD float gl_abs(float x) { return fabsf(x); }
GL_HD double gl_abs(double x) { return fabs(x); }
GL_HD float gl_ceil(float x) { return ceilf(x); }
GL_HD double gl_ceil(double x) { return ceil(x); }
GL_HD float gl_min(float a, float b) { return fminf(a, b); }
GL_HD double gl_min(double a, double b) { return fmin(a, b); }
GL_HD float gl_max(float a, float b) { return fmaxf(a, b); }
GL_HD double gl_max(double a, double b) { return fmax(a, b); }
GL_HD bool gl_isnan(float x) { return x != x; }
GL_HD bool gl_isnan(double x) { return x != x; }
GL_HD float gl_fma(float a, float b, float c) { return fmaf(a, b, c); }
GL_HD double gl_fma(double a, double b, double c) { return a * b + c; }   // host harness: plain (no contraction)

// Hot-loop special functions.  On the device (fp32) these are single MUFU operations:
//   lg2.approx (abs error <= 2^-22.6), ex2.approx (rel error <= 2^-22), rcp/rsqrt.approx (1-2 ulp).
// They are only used where the log feeds an exponent of magnitude <= ~1 (x^y = 2^(y log2 x) with
// |y| <= 1, Sersic 1/n and EPL gamma-2), so the result error stays at the 1e-7 level of powf/expf
// themselves; GPU parity tests hold 1e-5 with them.  The fp64 / host instantiation is exact.
#if defined(__CUDA_ARCH__)
#if defined(GL_ACCURATE_LOGEXP)   // experiment switch (scripts/dz_outliers.py): full-precision libm instead of the MUFU approximations
GL_HD float gl_log2_fast(float x) { return log2f(x); }
GL_HD float gl_exp2_fast(float x) { return exp2f(x); }
#else
GL_HD float gl_log2_fast(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
GL_HD float gl_exp2_fast(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif
// rcp / rsqrt as the bare MUFU op (.ftz): rsqrtf() and __fdividef() wrap it in range tests and rescaling
// multiplies for denormal or huge operands (1 FSETP + 2 FMUL per call -- a fifth of the dPIE member loop),
// which the operands here (squared radii and norms on the arcsecond scale) never are.
#if defined(GL_ACCURATE_RCP)      // experiment switch: IEEE division / square root instead of the MUFU approximations
GL_HD float gl_rcp_fast(float x) { return __fdiv_rn(1.0f, x); }
GL_HD float gl_div_fast(float a, float b) { return __fdiv_rn(a, b); }
GL_HD float gl_rsqrt_fast(float x) { return __fdiv_rn(1.0f, __fsqrt_rn(x)); }
#else
GL_HD float gl_rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
GL_HD float gl_div_fast(float a, float b) { return a * gl_rcp_fast(b); }
GL_HD float gl_rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif
#else
GL_HD float gl_log2_fast(float x) { return log2f(x); }
GL_HD float gl_exp2_fast(float x) { return exp2f(x); }
GL_HD float gl_div_fast(float a, float b) { return a / b; }
GL_HD float gl_rsqrt_fast(float x) { return 1.0f / sqrtf(x); }
#endif
GL_HD double gl_log2_fast(double x) { return log2(x); }
GL_HD double gl_exp2_fast(double x) { return exp2(x); }
GL_HD double gl_div_fast(double a, double b) { return a / b; }
GL_HD double gl_rsqrt_fast(double x) { return 1.0 / sqrt(x); }
// atan2 with |error| <= 1e-7 (the level of atan2f): minimax polynomial of atan(t)/t in t^2 on [0,1]
// (max abs error 9.5e-8 in fp32 Horner), t = min/max via one rcp, octant fix-up by selects.  Used in
// the dPIE member loop, where the libm atan2f (~45 instructions with its slow paths) dominated.
#if defined(__CUDA_ARCH__)
GL_HD float gl_atan2_fast(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float t = (mx > 0.f) ? mn * gl_rcp_fast(mx) : 0.f;
  const float s = t * t;
  float p = 0.0029327620286494493f;
  p = fmaf(p, s, -0.016413191333413124f);
  p = fmaf(p, s, 0.04327824339270592f);
  p = fmaf(p, s, -0.07556900382041931f);
  p = fmaf(p, s, 0.10667487233877182f);
  p = fmaf(p, s, -0.14211106300354004f);
  p = fmaf(p, s, 0.19993694126605988f);
  p = fmaf(p, s, -0.3333313763141632f);
  p = fmaf(p, s, 1.0f);
  float r = p * t;
  r = (ay > ax) ? 1.5707963267948966f - r : r;
  r = (x < 0.f) ? 3.141592653589793f - r : r;
  return copysignf(r, y);
}
#else
GL_HD float gl_atan2_fast(float y, float x) { return atan2f(y, x); }
#endif
GL_HD double gl_atan2_fast(double y, double x) { return atan2(y, x); }
#define GL_LN2 0.69314718055994531
#define GL_LOG2E 1.4426950408889634


// ---------------------------------------------------------------------------------------------
// lane types.  The per-pixel functions of the hot profiles (EPL, SHEAR, SERSIC) are templates over
// a lane type V and the scalar type S = gl_scalar_t<V> of the per-sample constants:
//   V = S = float / double : one pixel per lane slot (generic kernels, host harness)
//   V = GlF2, S = float    : TWO pixels per lane slot; every +, *, fma is one packed FFMA2 / FMUL2 /
//                            FADD2 (sm_100+), i.e. half the issue slots of the issue-bound kernels.
// Per-pixel conditionals are written as selects (gl_where_*), so the same source serves both.
// ---------------------------------------------------------------------------------------------
struct alignas(8) GlF2 {
  float x, y;
  GL_HD GlF2() {}
  GL_HD GlF2(float a) : x(a), y(a) {}
  GL_HD GlF2(float a, float b) : x(a), y(b) {}
};
#if defined(__CUDA_ARCH__)
GL_HD GlF2 operator+(GlF2 a, GlF2 b) { float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)); return GlF2(r.x, r.y); }
GL_HD GlF2 operator*(GlF2 a, GlF2 b) { float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y)); return GlF2(r.x, r.y); }
GL_HD GlF2 gl_fma(GlF2 a, GlF2 b, GlF2 c) {
  float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
  return GlF2(r.x, r.y);
}
#else
GL_HD GlF2 operator+(GlF2 a, GlF2 b) { return GlF2(a.x + b.x, a.y + b.y); }
GL_HD GlF2 operator*(GlF2 a, GlF2 b) { return GlF2(a.x * b.x, a.y * b.y); }
GL_HD GlF2 gl_fma(GlF2 a, GlF2 b, GlF2 c) { return GlF2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
GL_HD GlF2 operator-(GlF2 a) { return GlF2(-a.x, -a.y); }
GL_HD GlF2 operator-(GlF2 a, GlF2 b) { return a + (-b); }
GL_HD GlF2& operator+=(GlF2& a, GlF2 b) { a = a + b; return a; }
GL_HD GlF2& operator-=(GlF2& a, GlF2 b) { a = a - b; return a; }
GL_HD GlF2& operator*=(GlF2& a, GlF2 b) { a = a * b; return a; }
GL_HD GlF2 gl_log2_fast(GlF2 a) { return GlF2(gl_log2_fast(a.x), gl_log2_fast(a.y)); }
GL_HD GlF2 gl_exp2_fast(GlF2 a) { return GlF2(gl_exp2_fast(a.x), gl_exp2_fast(a.y)); }
GL_HD GlF2 gl_rsqrt_fast(GlF2 a) { return GlF2(gl_rsqrt_fast(a.x), gl_rsqrt_fast(a.y)); }
GL_HD GlF2 gl_div_fast(GlF2 a, GlF2 b) { return GlF2(gl_div_fast(a.x, b.x), gl_div_fast(a.y, b.y)); }
#if defined(__CUDA_ARCH__)
GL_HD GlF2 gl_atan2_fast(GlF2 y, GlF2 x) {   // the polynomial of the scalar version on packed lanes; octant fix-up per lane
  const float ax0 = fabsf(x.x), ay0 = fabsf(y.x), ax1 = fabsf(x.y), ay1 = fabsf(y.y);
  const float mx0 = fmaxf(ax0, ay0), mx1 = fmaxf(ax1, ay1);
  const GlF2 t = GlF2(fminf(ax0, ay0), fminf(ax1, ay1)) * GlF2(mx0 > 0.f ? gl_rcp_fast(mx0) : 0.f, mx1 > 0.f ? gl_rcp_fast(mx1) : 0.f);
  const GlF2 s = t * t;
  GlF2 p = GlF2(0.0029327620286494493f);
  p = gl_fma(p, s, GlF2(-0.016413191333413124f));
  p = gl_fma(p, s, GlF2(0.04327824339270592f));
  p = gl_fma(p, s, GlF2(-0.07556900382041931f));
  p = gl_fma(p, s, GlF2(0.10667487233877182f));
  p = gl_fma(p, s, GlF2(-0.14211106300354004f));
  p = gl_fma(p, s, GlF2(0.19993694126605988f));
  p = gl_fma(p, s, GlF2(-0.3333313763141632f));
  p = gl_fma(p, s, GlF2(1.0f));
  const GlF2 r = p * t;
  float r0 = (ay0 > ax0) ? 1.5707963267948966f - r.x : r.x, r1 = (ay1 > ax1) ? 1.5707963267948966f - r.y : r.y;
  r0 = (x.x < 0.f) ? 3.141592653589793f - r0 : r0; r1 = (x.y < 0.f) ? 3.141592653589793f - r1 : r1;
  return GlF2(copysignf(r0, y.x), copysignf(r1, y.y));
}
#else
GL_HD GlF2 gl_atan2_fast(GlF2 y, GlF2 x) { return GlF2(gl_atan2_fast(y.x, x.x), gl_atan2_fast(y.y, x.y)); }
#endif
GL_HD GlF2 gl_min(GlF2 a, GlF2 b) { return GlF2(fminf(a.x, b.x), fminf(a.y, b.y)); }
GL_HD GlF2 gl_max(GlF2 a, GlF2 b) { return GlF2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
// per-lane selects: (a > t ? vt : vf), (lo <= a <= hi ? vt : vf)
GL_HD GlF2 gl_where_gt(GlF2 a, float t, GlF2 vt, GlF2 vf) { return GlF2(a.x > t ? vt.x : vf.x, a.y > t ? vt.y : vf.y); }
GL_HD GlF2 gl_where_in(GlF2 a, float lo, float hi, GlF2 vt, GlF2 vf) {
  return GlF2((a.x >= lo && a.x <= hi) ? vt.x : vf.x, (a.y >= lo && a.y <= hi) ? vt.y : vf.y);
}
GL_HD GlF2 gl_where_lt(GlF2 a, float t, GlF2 vt, GlF2 vf) { return GlF2(a.x < t ? vt.x : vf.x, a.y < t ? vt.y : vf.y); }
GL_HD float gl_where_lt(float a, float t, float vt, float vf) { return a < t ? vt : vf; }
GL_HD float gl_hsum(GlF2 a) { return a.x + a.y; }
GL_HD float gl_where_gt(float a, float t, float vt, float vf) { return a > t ? vt : vf; }
GL_HD double gl_where_gt(double a, double t, double vt, double vf) { return a > t ? vt : vf; }
GL_HD float gl_where_in(float a, float lo, float hi, float vt, float vf) { return (a >= lo && a <= hi) ? vt : vf; }
GL_HD double gl_where_in(double a, double lo, double hi, double vt, double vf) { return (a >= lo && a <= hi) ? vt : vf; }
GL_HD float gl_hsum(float a) { return a; }
GL_HD double gl_hsum(double a) { return a; }
template <class V> struct gl_scalar_of { typedef V type; };
template <> struct gl_scalar_of<GlF2> { typedef float type; };
template <class V> struct gl_lanes_of { enum { value = 1 }; };
template <> struct gl_lanes_of<GlF2> { enum { value = 2 }; };
GL_HD float gl_lane(float a, int) { return a; }
GL_HD double gl_lane(double a, int) { return a; }
GL_HD float gl_lane(GlF2 a, int i) { return i ? a.y : a.x; }

// ---------------------------------------------------------------------------------------------
// GlDual<S>: a lane type carrying a value and ONE directional derivative (forward-mode AD).
// Instantiating the per-pixel code with V = GlDual<S> (the per-sample constants stay plain S) gives
//   * X_fwd<GlDual>: the deflection and its derivative along a coordinate direction, i.e. one column
//     of the lensing Hessian (MassProfile.hessian, src/gigalens/tf/profile.py:9-30 and the analytic
//     versions in sis.py / shear.py / nfw.py / piemd.py) -- magnification, convergence, shear maps;
//   * X_bwd<GlDual>: the directional derivative of the hand adjoint, which is exactly the parameter
//     gradient of anything that depends on the Hessian (the image-position likelihood with its
//     magnification-scaled errors, tf/model.py:103-124).
// Only the tiny point workloads use it (tens of image positions per sample), never the pixel path.
// Operators are in-class friends so that mixed S (op) GlDual expressions convert implicitly.
// ---------------------------------------------------------------------------------------------
template <class S>
struct GlDual {
  S v, d;
  GL_HD GlDual() {}
  GL_HD GlDual(S a) : v(a), d(S(0)) {}
  GL_HD GlDual(S a, S b) : v(a), d(b) {}
  friend GL_HD GlDual operator+(GlDual a, GlDual b) { return GlDual(a.v + b.v, a.d + b.d); }
  friend GL_HD GlDual operator-(GlDual a, GlDual b) { return GlDual(a.v - b.v, a.d - b.d); }
  friend GL_HD GlDual operator*(GlDual a, GlDual b) { return GlDual(a.v * b.v, a.v * b.d + a.d * b.v); }
  friend GL_HD GlDual operator/(GlDual a, GlDual b) { S r = a.v / b.v; return GlDual(r, (a.d - r * b.d) / b.v); }
  friend GL_HD GlDual operator-(GlDual a) { return GlDual(-a.v, -a.d); }
  friend GL_HD GlDual& operator+=(GlDual& a, GlDual b) { a.v += b.v; a.d += b.d; return a; }
  friend GL_HD GlDual& operator-=(GlDual& a, GlDual b) { a.v -= b.v; a.d -= b.d; return a; }
  friend GL_HD GlDual& operator*=(GlDual& a, GlDual b) { a = a * b; return a; }
  friend GL_HD bool operator<(GlDual a, GlDual b) { return a.v < b.v; }
  friend GL_HD bool operator>(GlDual a, GlDual b) { return a.v > b.v; }
  friend GL_HD bool operator<=(GlDual a, GlDual b) { return a.v <= b.v; }
  friend GL_HD bool operator>=(GlDual a, GlDual b) { return a.v >= b.v; }
  friend GL_HD bool operator==(GlDual a, GlDual b) { return a.v == b.v; }
  friend GL_HD GlDual gl_sqrt(GlDual a) { S r = gl_sqrt(a.v); return GlDual(r, a.d / (S(2) * r)); }
  friend GL_HD GlDual gl_rsqrt_fast(GlDual a) { S r = gl_rsqrt_fast(a.v); return GlDual(r, S(-0.5) * r * r * r * a.d); }
  friend GL_HD GlDual gl_exp(GlDual a) { S e = gl_exp(a.v); return GlDual(e, e * a.d); }
  friend GL_HD GlDual gl_exp2_fast(GlDual a) { S e = gl_exp2_fast(a.v); return GlDual(e, e * a.d * S(GL_LN2)); }
  friend GL_HD GlDual gl_log(GlDual a) { return GlDual(gl_log(a.v), a.d / a.v); }
  friend GL_HD GlDual gl_log2_fast(GlDual a) { return GlDual(gl_log2_fast(a.v), a.d / a.v * S(GL_LOG2E)); }
  friend GL_HD GlDual gl_atan2(GlDual y, GlDual x) { return GlDual(gl_atan2(y.v, x.v), (x.v * y.d - y.v * x.d) / (x.v * x.v + y.v * y.v)); }
  friend GL_HD GlDual gl_atan2_fast(GlDual y, GlDual x) { return GlDual(gl_atan2_fast(y.v, x.v), (x.v * y.d - y.v * x.d) / (x.v * x.v + y.v * y.v)); }
  friend GL_HD GlDual gl_atan(GlDual a) { return GlDual(gl_atan(a.v), a.d / (S(1) + a.v * a.v)); }
  friend GL_HD GlDual gl_atanh(GlDual a) { return GlDual(gl_atanh(a.v), a.d / (S(1) - a.v * a.v)); }
  friend GL_HD GlDual gl_acosh(GlDual a) { return GlDual(gl_acosh(a.v), a.d / gl_sqrt(a.v * a.v - S(1))); }
  friend GL_HD GlDual gl_acos(GlDual a) { return GlDual(gl_acos(a.v), -a.d / gl_sqrt(S(1) - a.v * a.v)); }
  friend GL_HD GlDual gl_cos(GlDual a) { return GlDual(gl_cos(a.v), -gl_sin(a.v) * a.d); }
  friend GL_HD GlDual gl_sin(GlDual a) { return GlDual(gl_sin(a.v), gl_cos(a.v) * a.d); }
  friend GL_HD GlDual gl_abs(GlDual a) { return a.v < S(0) ? -a : a; }
  friend GL_HD GlDual gl_min(GlDual a, GlDual b) { return (b.v < a.v) ? b : a; }
  friend GL_HD GlDual gl_max(GlDual a, GlDual b) { return (b.v > a.v) ? b : a; }
  friend GL_HD GlDual gl_div_fast(GlDual a, GlDual b) { return a / b; }
  friend GL_HD GlDual gl_fma(GlDual a, GlDual b, GlDual c) { return GlDual(gl_fma(a.v, b.v, c.v), gl_fma(a.v, b.d, gl_fma(a.d, b.v, c.d))); }
  friend GL_HD bool gl_isnan(GlDual a) { return a.v != a.v; }
  friend GL_HD GlDual gl_where_gt(GlDual a, S t, GlDual vt, GlDual vf) { return a.v > t ? vt : vf; }
  friend GL_HD GlDual gl_where_in(GlDual a, S lo, S hi, GlDual vt, GlDual vf) { return (a.v >= lo && a.v <= hi) ? vt : vf; }
};
template <class S> struct gl_scalar_of<GlDual<S>> { typedef S type; };

// Profile type ids (same values as include/gigalens_b200.h).
enum {
  GLT_EPL = 1, GLT_SHEAR = 2, GLT_SIE = 3, GLT_SIS = 4, GLT_NFW = 5, GLT_NFW_ELLIPSE = 6, GLT_DPIS = 7, GLT_DPIE = 8,
  GLT_TNFW = 9, GLT_DPIEP = 10,
  GLT_SERSIC = 32, GLT_SERSIC_ELLIPSE = 33, GLT_SHAPELETS = 34, GLT_CORE_SERSIC = 35
};

// Feature bits: which profile families a kernel instantiation contains code for.  The per-pixel
// drivers are templates over a feature mask F; cases outside F are compiled out, so a program that
// uses only {EPL, SHEAR, SERSIC} runs a kernel whose register allocation is not inflated by the dPIE
// adjoint or the Shapelets scratch arrays.  GLF_ALL is the generic interpreter.
enum {
  GLF_EPL = 1, GLF_SHEAR = 2, GLF_SIE = 4, GLF_SIS = 8, GLF_NFW = 16, GLF_DPIS = 32, GLF_DPIE = 64, GLF_SERSIC = 128,
  GLF_SHAPELETS = 256, GLF_TNFW = 512, GLF_DPIEP = 1024, GLF_CORESERSIC = 2048, GLF_ALL = 4095
};
GL_HD unsigned gl_feature_of(int type) {
  switch (type) {
    case GLT_EPL: return GLF_EPL; case GLT_SHEAR: return GLF_SHEAR; case GLT_SIE: return GLF_SIE; case GLT_SIS: return GLF_SIS;
    case GLT_NFW: case GLT_NFW_ELLIPSE: return GLF_NFW; case GLT_DPIS: return GLF_DPIS; case GLT_DPIE: return GLF_DPIE;
    case GLT_SERSIC: case GLT_SERSIC_ELLIPSE: return GLF_SERSIC; case GLT_SHAPELETS: return GLF_SHAPELETS;
    case GLT_TNFW: return GLF_TNFW; case GLT_DPIEP: return GLF_DPIEP; case GLT_CORE_SERSIC: return GLF_CORESERSIC;
  }
  return 0;
}

#define GL_MAX_DVARS 8      // accumulators per profile in the pixel adjoint (CORE_SERSIC flushes its 9th, Ie, separately)
#define GL_MAX_RAW 10       // == GL_MAX_PROFILE_PARAMS of include/gigalens_b200.h (CORE_SERSIC has 10 raw parameters)

// ---------------------------------------------------------------------------------------------
// ellipticity (e1, e2) -> (phi, q) and its adjoint  
//   phi = atan2(e2, e1)/2,  c = min(|e|, cmax)  [EPL: clip(|e|, 0, 1)],  q = (1-c)/(1+c)
// ---------------------------------------------------------------------------------------------
template <class T>
GL_HD void ellip_fwd(T e1, T e2, T cmax, T& phi, T& q, T& c) {
  phi = gl_atan2(e2, e1) / T(2);
  T craw = gl_sqrt(e1 * e1 + e2 * e2);
  c = gl_min(craw, cmax);
  q = (T(1) - c) / (T(1) + c);
}
// tf.minimum / clip_by_value pass the gradient on the un-clamped side (ties included).
// At e = (0,0) TF yields NaN (0/0); measure zero under every prior, we return 0.
template <class T>
GL_HD void ellip_bwd(T e1, T e2, T cmax, T gphi, T gq, T& ge1, T& ge2) {
  T r2 = e1 * e1 + e2 * e2;
  if (!(r2 > T(0))) { ge1 = T(0); ge2 = T(0); return; }
  T craw = gl_sqrt(r2);
  T c = gl_min(craw, cmax);
  T gc = gq * (-T(2) / ((T(1) + c) * (T(1) + c)));
  T gcraw = (craw <= cmax) ? gc : T(0);
  ge1 = gcraw * e1 / craw - gphi * e2 / (T(2) * r2);
  ge2 = gcraw * e2 / craw + gphi * e1 / (T(2) * r2);
}

// =============================================================================================
// EPL  (src/gigalens/tf/profiles/mass/epl.py:19-57)
//   raw  : theta_E, gamma, e1, e2, center_x, center_y
//   d[]  : cx, cy, cos(phi), sin(phi), q, b, t, pref0 = 2b/(1+q), f, N (trip count), log2 b, R_N,
//          then two tables of stride ts:  Bq_m = A_{m+1}/f,  Bt_m = dA_{m+1}/dt   (m = 0..N-1)
//   dvars: cx, cy, phi, q, b, t, f, pref0
// The reference iterates Omega_n = c_n Rot(2 ang) Omega_{n-1}, c_n = -f (2n-s)/(2n+s), s = 2-t, and
// sums Omega_0..Omega_N.  With u = e^{i ang}, w = u^2 and A_n = prod_{k<=n} c_k (per-sample
// constants) that sum is u * P(w), P(w) = sum_n A_n w^n = 1 + f w Q(w), Q = sum_m Bq_m w^m, which we
// evaluate by complex Horner: 4 FMA per trip instead of the recurrence's 8, and no per-pixel state
// beyond (Q_re, Q_im).
//
// Adjoint.  dP/dt needs its own series T = w sum_m Bt_m w^m.  dP/df (and w dP/dw = f dP/df) does
// not: c_n (2n+s) = -f (2n-s) makes the truncated sum satisfy the first-order relation
//     2 D (1 + f w) = f w [ R_N w^N - (s Q + t P) ],   D = w dP/dw = f dP/df,  R_N = A_N (2N+2-s),
// so dP/df = w [R_N w^N - (s Q + t P)] / (2 (1 + f w)) in closed form (well conditioned: no
// division by f, no P - 1).  R_N w^N is the footprint of the truncation; it is below rounding unless
// the iteration cap cuts the series short, and is evaluated (binary powering) only for such samples.
// =============================================================================================
enum { EPL_CX = 0, EPL_CY, EPL_C, EPL_S, EPL_Q, EPL_B, EPL_T, EPL_PREF0, EPL_F, EPL_N, EPL_LOG2B, EPL_RN, EPL_TAB = 12 };
enum { EPLG_CX = 0, EPLG_CY, EPLG_PHI, EPLG_Q, EPLG_B, EPLG_T, EPLG_F, EPLG_PREF0 };
#define GL_EPL_NSTATE 4   // per-pixel series state (Q_re, Q_im, T'_re, T'_im) a forward pass can hand to the adjoint

GL_HD int epl_table_stride(int niter_cap) { return ((niter_cap + 1) + 3) & ~3; }
GL_HD int epl_der_size(int niter_cap) { return EPL_TAB + 2 * epl_table_stride(niter_cap); }

// Trip count of `tf.while_loop(i < niter, i0 = 1.0, maximum_iterations=cap)` (epl.py:37,47-54);
// tol = 1e-12 is the reference's constant.
template <class T>
GL_HD int epl_trip_count(T fmax, int cap, T tol) {
  T niter = gl_log(tol) / gl_log(fmax) + T(2);
  if (!(niter > T(1))) return 0;  // also catches NaN / -inf
  T n = gl_ceil(niter) - T(1);
  if (n > T(cap)) return cap;
  return (int)n;
}

// fmax_batch < 0  => use this sample's own f (default; differs from the reference's batch-global
// count only by series terms below tol, SURVEY.md §7 "Iteration count semantics").
template <class T>
GL_HD void epl_prep(const T* raw, T* d, int niter_cap, T fmax_batch, T tol) {
  T theta_E = raw[0], gamma = raw[1], e1 = raw[2], e2 = raw[3];
  T phi, q, c;
  ellip_fwd(e1, e2, T(1), phi, q, c);
  T theta_E_conv = theta_E / gl_sqrt((T(1) + q * q) / (T(2) * q));
  T b = theta_E_conv * gl_sqrt((T(1) + q * q) / T(2));
  T t = gamma - T(1);
  T f = (T(1) - q) / (T(1) + q);
  d[EPL_CX] = raw[4]; d[EPL_CY] = raw[5];
  d[EPL_C] = gl_cos(phi); d[EPL_S] = gl_sin(phi);
  d[EPL_Q] = q; d[EPL_B] = b; d[EPL_T] = t;
  d[EPL_PREF0] = (T(2) * b) / (T(1) + q);
  d[EPL_F] = f;
  int N = epl_trip_count(fmax_batch < T(0) ? f : fmax_batch, niter_cap, tol);
  d[EPL_N] = T(N);
  d[EPL_LOG2B] = gl_log(b) * T(GL_LOG2E);
  const int ts = epl_table_stride(niter_cap);
  T* Bq = d + EPL_TAB; T* Bt = Bq + ts;
  T s = T(2) - t;
  T a = T(1), at = T(0), bq = T(0);   // A_n, dA_n/dt, A_n / f
  for (int n = 1; n <= ts; ++n) {
    if (n <= N) {
      T tn = T(2 * n);
      T kap = -(tn - s) / (tn + s);           // c_n = f * kap
      T cn = f * kap;
      T dcdt = -f * T(4 * n) / ((tn + s) * (tn + s));
      at = at * cn + a * dcdt;
      bq = (n == 1) ? kap : bq * cn;
      a = a * cn;
      Bq[n - 1] = bq; Bt[n - 1] = at;
    } else {
      Bq[n - 1] = T(0); Bt[n - 1] = T(0);
    }
  }
  // truncation footprint in the closed-form f-derivative; dropped when below a tenth of an ulp of the O(1) sum
  const T RN = a * (T(2 * N + 2) - s);
  const T eps = sizeof(T) == 4 ? T(1.2e-7) : T(2.3e-16);
  d[EPL_RN] = (gl_abs(RN) > T(0.1) * eps) ? RN : T(0);
}

template <class T>
GL_HD void epl_prep_bwd(const T* raw, const T* d, const T* g, T* graw) {
  T e1 = raw[2], e2 = raw[3];
  T q = d[EPL_Q], b = d[EPL_B];
  T gq = g[EPLG_Q], gb = g[EPLG_B];
  T opq = T(1) + q;
  gb += g[EPLG_PREF0] * T(2) / opq;
  gq -= g[EPLG_PREF0] * T(2) * b / (opq * opq);
  gq -= g[EPLG_F] * T(2) / (opq * opq);
  // b = theta_E * sqrt(q) (algebraically; epl.py:24-25)
  T sq = gl_sqrt(q);
  graw[0] = gb * sq;
  gq += (q > T(0)) ? gb * b / (T(2) * q) : T(0);
  graw[1] = g[EPLG_T];
  ellip_bwd(e1, e2, T(1), g[EPLG_PHI], gq, graw[2], graw[3]);
  graw[4] = g[EPLG_CX];
  graw[5] = g[EPLG_CY];
}

// Shared per-pixel geometry of the EPL forward / adjoint.  R = clip(sqrt(R2), 1e-10, 1e10) enters
// only through log2(R2) (prefactor) and 1/sqrt(R2) (cos/sin of the polar angle), so no sqrt is taken.
// V is the lane type (float, double, or the two-pixel pack GlF2), S its scalar type.
template <class V>
struct EplGeom {
  V xr, yr, qx, R2, ir, l2r2, Cs, Ss, wr, wi;   // ir = 1/R0 (0 when R0 == 0), l2r2 = log2(clamped R^2)
};
template <class V, class S>
GL_HD void epl_geom(const S* d, V x, V y, EplGeom<V>& G) {
  V dx = x - V(d[EPL_CX]), dy = y - V(d[EPL_CY]);
  V c = V(d[EPL_C]), s = V(d[EPL_S]);
  G.xr = gl_fma(dx, c, dy * s);
  G.yr = gl_fma(dy, c, -(dx * s));
  G.qx = V(d[EPL_Q]) * G.xr;
  G.R2 = gl_fma(G.qx, G.qx, G.yr * G.yr);
  G.ir = gl_where_gt(G.R2, S(0), gl_rsqrt_fast(G.R2), V(S(0)));
  G.Cs = gl_where_gt(G.R2, S(0), G.qx * G.ir, V(S(1)));   // cos/sin of atan2(yr, q xr); atan2(0,0) = 0
  G.Ss = G.yr * G.ir;
  G.l2r2 = gl_log2_fast(gl_min(gl_max(G.R2, V(S(1e-20))), V(S(1e20))));
  G.wr = gl_fma(G.Cs, G.Cs, -(G.Ss * G.Ss));
  G.wi = V(S(2)) * G.Cs * G.Ss;
}

// Complex Horner step P <- P w + a for TWO pixels at once with Blackwell's packed fp32 FMA (FFMA2,
// sm_100+): the series loops are issue-bound, and one FFMA2 retires two FMAs per issue slot.
#if defined(__CUDA_ARCH__)
#define GL_HAVE_F32X2 1
struct GlC2 { float2 r, i; };   // (re, im) of two pixels
__device__ __forceinline__ void gl_horner2(GlC2& P, const float2 wr, const float2 wi, const float a) {
  const float2 a2 = make_float2(a, a);
  const float2 nPi = make_float2(-P.i.x, -P.i.y);
  const float2 pr = __ffma2_rn(P.r, wr, __ffma2_rn(nPi, wi, a2));
  const float2 pi = __ffma2_rn(P.r, wi, __fmul2_rn(P.i, wr));
  P.r = pr; P.i = pi;
}
#endif

// Q = sum_m Bq_m w^m and (WANT_T) T' = sum_m Bt_m w^m, m = 0..N-1, at NP lanes.
template <class V, int NP, bool WANT_T>
GL_HD void epl_series(const typename gl_scalar_of<V>::type* d, int ts, const EplGeom<V>* G, V* Qr, V* Qi, V* Tr, V* Ti) {
  typedef typename gl_scalar_of<V>::type S;
  const int N = (int)d[EPL_N];
  const S* Bq = d + EPL_TAB; const S* Bt = Bq + ts;
  if (N <= 0) {
#pragma unroll
    for (int j = 0; j < NP; ++j) { Qr[j] = V(S(0)); Qi[j] = V(S(0)); if constexpr (WANT_T) { Tr[j] = V(S(0)); Ti[j] = V(S(0)); } }
    return;
  }
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    Qr[j] = V(Bq[N - 1]); Qi[j] = V(S(0));
    if constexpr (WANT_T) { Tr[j] = V(Bt[N - 1]); Ti[j] = V(S(0)); }
  }
#ifdef GL_HAVE_F32X2
  if constexpr (sizeof(V) == 4 && (NP % 2) == 0) {   // scalar float lanes: pack pixel pairs for the series only
    GlC2 Q2[NP / 2], T2[NP / 2]; float2 wr2[NP / 2], wi2[NP / 2];
#pragma unroll
    for (int h = 0; h < NP / 2; ++h) {
      Q2[h].r = make_float2(Qr[2 * h], Qr[2 * h + 1]); Q2[h].i = make_float2(0.f, 0.f);
      if constexpr (WANT_T) { T2[h].r = make_float2(Tr[2 * h], Tr[2 * h + 1]); T2[h].i = make_float2(0.f, 0.f); }
      wr2[h] = make_float2(G[2 * h].wr, G[2 * h + 1].wr); wi2[h] = make_float2(G[2 * h].wi, G[2 * h + 1].wi);
    }
    for (int n = N - 2; n >= 0; --n) {
      const float a = Bq[n], at = Bt[n];
#pragma unroll
      for (int h = 0; h < NP / 2; ++h) {
        gl_horner2(Q2[h], wr2[h], wi2[h], a);
        if constexpr (WANT_T) gl_horner2(T2[h], wr2[h], wi2[h], at);
      }
    }
#pragma unroll
    for (int h = 0; h < NP / 2; ++h) {
      Qr[2 * h] = Q2[h].r.x; Qr[2 * h + 1] = Q2[h].r.y; Qi[2 * h] = Q2[h].i.x; Qi[2 * h + 1] = Q2[h].i.y;
      if constexpr (WANT_T) { Tr[2 * h] = T2[h].r.x; Tr[2 * h + 1] = T2[h].r.y; Ti[2 * h] = T2[h].i.x; Ti[2 * h + 1] = T2[h].i.y; }
    }
    return;
  }
#endif
#pragma unroll 2
  for (int n = N - 2; n >= 0; --n) {   // unrolled by two so the loop-carried pairs ping-pong without register copies
    const V a = V(Bq[n]);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      V qr = gl_fma(Qr[j], G[j].wr, gl_fma(-Qi[j], G[j].wi, a));
      V qi = gl_fma(Qr[j], G[j].wi, Qi[j] * G[j].wr);
      Qr[j] = qr; Qi[j] = qi;
    }
    if constexpr (WANT_T) {
      const V at = V(Bt[n]);
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        V tr = gl_fma(Tr[j], G[j].wr, gl_fma(-Ti[j], G[j].wi, at));
        V ti = gl_fma(Tr[j], G[j].wi, Ti[j] * G[j].wr);
        Tr[j] = tr; Ti[j] = ti;
      }
    }
  }
}

// deflection from the geometry and Q (P = 1 + f w Q)
template <class V, int NP>
GL_HD void epl_alpha(const typename gl_scalar_of<V>::type* d, const EplGeom<V>* G, const V* Qr, const V* Qi, V* ax, V* ay) {
  typedef typename gl_scalar_of<V>::type S;
  const V c = V(d[EPL_C]), s = V(d[EPL_S]), l2b = V(d[EPL_LOG2B]), tm1 = V(d[EPL_T] - S(1)), pref0 = V(d[EPL_PREF0]);
  const V f = V(d[EPL_F]);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const V zr = f * G[j].wr, zi = f * G[j].wi;
    const V Pr = gl_fma(zr, Qr[j], gl_fma(-zi, Qi[j], V(S(1)))), Pi = gl_fma(zr, Qi[j], zi * Qr[j]);
    V fx = gl_fma(G[j].Cs, Pr, -(G[j].Ss * Pi));
    V fy = gl_fma(G[j].Cs, Pi, G[j].Ss * Pr);
    V pref = pref0 * gl_exp2_fast(tm1 * gl_fma(V(S(-0.5)), G[j].l2r2, l2b));   // (b/R)^(t-1)
    fx = fx * pref; fy = fy * pref;
    ax[j] = gl_fma(fx, c, -(fy * s));
    ay[j] = gl_fma(fx, s, fy * c);
  }
}

template <class V, int NP>
GL_HD void epl_fwd(const typename gl_scalar_of<V>::type* d, int ts, const V* x, const V* y, V* ax, V* ay) {
  typedef typename gl_scalar_of<V>::type S;
  EplGeom<V> G[NP];
  V Qr[NP], Qi[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) epl_geom<V, S>(d, x[j], y[j], G[j]);
  epl_series<V, NP, false>(d, ts, G, Qr, Qi, (V*)nullptr, (V*)nullptr);
  epl_alpha<V, NP>(d, G, Qr, Qi, ax, ay);
}

// Forward that also runs the T' series and parks (Q, T') in scr[k * scr_stride], k < GL_EPL_NSTATE * NP,
// for epl_bwd_load: the adjoint kernels need the deflection first (source-plane position) and the
// series again later, and the series is the expensive part.
template <class V, int NP>
GL_HD void epl_fwd_save(const typename gl_scalar_of<V>::type* d, int ts, const V* x, const V* y, V* ax, V* ay, V* scr, int scr_stride) {
  typedef typename gl_scalar_of<V>::type S;
  EplGeom<V> G[NP];
  V Qr[NP], Qi[NP], Tr[NP], Ti[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) epl_geom<V, S>(d, x[j], y[j], G[j]);
  epl_series<V, NP, true>(d, ts, G, Qr, Qi, Tr, Ti);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    scr[(4 * j + 0) * scr_stride] = Qr[j]; scr[(4 * j + 1) * scr_stride] = Qi[j];
    scr[(4 * j + 2) * scr_stride] = Tr[j]; scr[(4 * j + 3) * scr_stride] = Ti[j];
  }
  epl_alpha<V, NP>(d, G, Qr, Qi, ax, ay);
}

// Adjoint given the series values: (gax, gay) is the cotangent of the deflection; accumulates into
// g[EPLG_*] (lane-wise).
template <class V, int NP>
GL_HD void epl_bwd_core(const typename gl_scalar_of<V>::type* d, const EplGeom<V>* G, const V* Qr, const V* Qi, const V* Tqr,
                        const V* Tqi, const V* gax, const V* gay, V* g) {
  typedef typename gl_scalar_of<V>::type S;
  const V c = V(d[EPL_C]), s = V(d[EPL_S]), q = V(d[EPL_Q]), l2b = V(d[EPL_LOG2B]), tm1 = V(d[EPL_T] - S(1));
  const V pref0 = V(d[EPL_PREF0]), f = V(d[EPL_F]), f2 = V(S(2) * d[EPL_F]);
  const V tm1_over_b = V((d[EPL_T] - S(1)) / d[EPL_B]);
  const V tt = V(d[EPL_T]), ss = V(S(2) - d[EPL_T]);
  const S RN = d[EPL_RN];
  const int N = (int)d[EPL_N];
  // R_N w^N for the (rare) samples whose series is cut short by the iteration cap -- evaluated for all lanes first so
  // that the per-pixel code below stays one straight-line block the scheduler can interleave across the NP lanes
  V wNr[NP], wNi[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) { wNr[j] = V(S(0)); wNi[j] = V(S(0)); }
  if (RN != S(0)) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      V pr = V(S(1)), pi = V(S(0)), br = G[j].wr, bi = G[j].wi;
      for (int e = N; e > 0; e >>= 1) {
        if (e & 1) { const V t0 = gl_fma(pr, br, -(pi * bi)); pi = gl_fma(pr, bi, pi * br); pr = t0; }
        const V b0 = gl_fma(br, br, -(bi * bi)); bi = V(S(2)) * br * bi; br = b0;
      }
      wNr[j] = V(RN) * pr; wNi[j] = V(RN) * pi;
    }
  }
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const EplGeom<V>& E = G[j];
    const V zr = f * E.wr, zi = f * E.wi;                                           // f w
    const V Pr = gl_fma(zr, Qr[j], gl_fma(-zi, Qi[j], V(S(1)))), Pi = gl_fma(zr, Qi[j], zi * Qr[j]);
    const V Tr = gl_fma(E.wr, Tqr[j], -(E.wi * Tqi[j])), Ti = gl_fma(E.wr, Tqi[j], E.wi * Tqr[j]);   // dP/dt = w T'
    // dP/df = w M / (2 (1 + f w)),  M = R_N w^N - (s Q + t P)
    const V Mr = wNr[j] - gl_fma(ss, Qr[j], tt * Pr), Mi = wNi[j] - gl_fma(ss, Qi[j], tt * Pi);
    const V dr = zr + V(S(1)), di = zi;
    const V inv = gl_div_fast(V(S(0.5)), gl_fma(dr, dr, di * di));
    const V Er = gl_fma(E.wr, Mr, -(E.wi * Mi)), Ei = gl_fma(E.wr, Mi, E.wi * Mr);
    const V Fr = gl_fma(Er, dr, Ei * di) * inv, Fi = gl_fma(Ei, dr, -(Er * di)) * inv;
    // forward values
    V fx = gl_fma(E.Cs, Pr, -(E.Ss * Pi)), fy = gl_fma(E.Cs, Pi, E.Ss * Pr);      // u P
    V l2br = gl_fma(V(S(-0.5)), E.l2r2, l2b);            // log2(b/R)
    V pw = gl_exp2_fast(tm1 * l2br);
    V pref = pref0 * pw;
    V Fx = fx * pref, Fy = fy * pref;
    V ax = gl_fma(Fx, c, -(Fy * s)), ay = gl_fma(Fx, s, Fy * c);
    // back-rotation by -phi
    g[EPLG_PHI] += gl_fma(gay[j], ax, -(gax[j] * ay));
    V gFx = gl_fma(gax[j], c, gay[j] * s), gFy = gl_fma(gay[j], c, -(gax[j] * s));
    // prefactor
    V gpref = gl_fma(gFx, fx, gFy * fy);
    V gr = gFx * pref, gi = gFy * pref;               // cotangent of u P (complex pair)
    g[EPLG_PREF0] += gpref * pw;
    V gpwpw = gpref * pref;                           // gpw * pw
    g[EPLG_T] += gpwpw * (l2br * V(S(GL_LN2)));
    g[EPLG_B] += gpwpw * tm1_over_b;
    V ir2 = E.ir * E.ir;
    V gR_over_R = gl_where_in(E.R2, S(1e-20), S(1e20), -(gpwpw * tm1 * ir2), V(S(0)));   // gR / R0 (0 where R is clamped)
    // series: F = u P(w; f, t).  <g, u X> = Re(conj(g) u X) for X = dP/df, dP/dt
    V hr = gl_fma(gr, E.Cs, gi * E.Ss), hi = gl_fma(gi, E.Cs, -(gr * E.Ss));   // conj(u) g
    g[EPLG_F] += gl_fma(hr, Fr, hi * Fi);
    g[EPLG_T] += gl_fma(hr, Tr, hi * Ti);
    // d/d(ang): dF = i u (P + 2 w P_w) d(ang), with w P_w = f dP/df
    V Zr = gl_fma(f2, Fr, Pr), Zi = gl_fma(f2, Fi, Pi);
    V gang = gl_fma(hi, Zr, -(hr * Zi));              // Re(conj(g) i u Z)
    // ang = atan2(yr, qx), R0 = hypot(qx, yr):  d(ang) = (qx dyr - yr dqx)/R0^2, dR0 = (qx dqx + yr dyr)/R0
    V ga = gang * ir2;
    V gqx = gl_fma(gR_over_R, E.qx, -(ga * E.yr));
    V gyr = gl_fma(gR_over_R, E.yr, ga * E.qx);
    g[EPLG_Q] += gqx * E.xr;
    V gxr = gqx * q;
    g[EPLG_PHI] += gl_fma(gxr, E.yr, -(gyr * E.xr));
    V gdx = gl_fma(gxr, c, -(gyr * s)), gdy = gl_fma(gxr, s, gyr * c);
    g[EPLG_CX] -= gdx;
    g[EPLG_CY] -= gdy;
  }
}

template <class V, int NP>
GL_HD void epl_bwd(const typename gl_scalar_of<V>::type* d, int ts, const V* x, const V* y, const V* gax, const V* gay, V* g) {
  typedef typename gl_scalar_of<V>::type S;
  EplGeom<V> G[NP];
  V Qr[NP], Qi[NP], Tr[NP], Ti[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) epl_geom<V, S>(d, x[j], y[j], G[j]);
  epl_series<V, NP, true>(d, ts, G, Qr, Qi, Tr, Ti);
  epl_bwd_core<V, NP>(d, G, Qr, Qi, Tr, Ti, gax, gay, g);
}

// Adjoint from the series state epl_fwd_save parked for the same pixels.
template <class V, int NP>
GL_HD void epl_bwd_load(const typename gl_scalar_of<V>::type* d, int ts, const V* x, const V* y, const V* gax, const V* gay, V* g,
                        const V* scr, int scr_stride) {
  typedef typename gl_scalar_of<V>::type S;
  (void)ts;
  EplGeom<V> G[NP];
  V Qr[NP], Qi[NP], Tr[NP], Ti[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    epl_geom<V, S>(d, x[j], y[j], G[j]);
    Qr[j] = scr[(4 * j + 0) * scr_stride]; Qi[j] = scr[(4 * j + 1) * scr_stride];
    Tr[j] = scr[(4 * j + 2) * scr_stride]; Ti[j] = scr[(4 * j + 3) * scr_stride];
  }
  epl_bwd_core<V, NP>(d, G, Qr, Qi, Tr, Ti, gax, gay, g);
}

// =============================================================================================
// SHEAR  (tf/profiles/mass/shear.py:14-16)   raw = d = dvars = (gamma1, gamma2)
// =============================================================================================
template <class V, int NP>
GL_HD void shear_fwd(const typename gl_scalar_of<V>::type* d, const V* x, const V* y, V* ax, V* ay) {
  const V g1 = V(d[0]), g2 = V(d[1]);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    ax[j] = gl_fma(g1, x[j], g2 * y[j]);
    ay[j] = gl_fma(g2, x[j], -(g1 * y[j]));
  }
}
template <class V, int NP>
GL_HD void shear_bwd(const typename gl_scalar_of<V>::type* d, const V* x, const V* y, const V* gax, const V* gay, V* g) {
  (void)d;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    g[0] += gl_fma(gax[j], x[j], -(gay[j] * y[j]));
    g[1] += gl_fma(gax[j], y[j], gay[j] * x[j]);
  }
}

// =============================================================================================
// SIS  (tf/profiles/mass/sis.py:12-17)   raw = d = dvars = (theta_E, cx, cy)
// =============================================================================================
template <class T, int NP>
GL_HD void sis_fwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, T* ax, T* ay) {
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - d[1], dy = y[j] - d[2];
    T R = gl_sqrt(dx * dx + dy * dy);
    T a = (R == T(0)) ? T(0) : d[0] / R;
    ax[j] = a * dx; ay[j] = a * dy;
  }
}
template <class T, int NP>
GL_HD void sis_bwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, const T* gax, const T* gay, T* g) {
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - d[1], dy = y[j] - d[2];
    T R2 = dx * dx + dy * dy;
    if (R2 > T(0)) {
      T R = gl_sqrt(R2);
      T a = d[0] / R;
      T ga = gax[j] * dx + gay[j] * dy;
      g[0] += ga / R;
      T gR = -ga * a / R;
      T gdx = gax[j] * a + gR * dx / R, gdy = gay[j] * a + gR * dy / R;
      g[1] -= gdx; g[2] -= gdy;
    }
  }
}

// =============================================================================================
// SIE  (tf/profiles/mass/sie.py:13-42, core s = 0)
//   raw  : theta_E, e1, e2, cx, cy
//   d[]  : cx, cy, cos, sin, q, b, w = sqrt(1-q^2)
//   dvars: cx, cy, phi, q, b, w
// =============================================================================================
enum { SIE_CX = 0, SIE_CY, SIE_C, SIE_S, SIE_Q, SIE_B, SIE_W, SIE_SIZE = 8 };
enum { SIEG_CX = 0, SIEG_CY, SIEG_PHI, SIEG_Q, SIEG_B, SIEG_W };
template <class T>
GL_HD void sie_prep(const T* raw, T* d) {
  T phi, q, c;
  ellip_fwd(raw[1], raw[2], T(0.9999), phi, q, c);
  T theta_E_conv = raw[0] / gl_sqrt((T(1) + q * q) / (T(2) * q));
  d[SIE_CX] = raw[3]; d[SIE_CY] = raw[4];
  d[SIE_C] = gl_cos(phi); d[SIE_S] = gl_sin(phi);
  d[SIE_Q] = q;
  d[SIE_B] = theta_E_conv * gl_sqrt((T(1) + q * q) / T(2));
  d[SIE_W] = gl_sqrt(T(1) - q * q);
  d[7] = T(0);
}
template <class T>
GL_HD void sie_prep_bwd(const T* raw, const T* d, const T* g, T* graw) {
  T q = d[SIE_Q], b = d[SIE_B], w = d[SIE_W];
  T gq = g[SIEG_Q] - g[SIEG_W] * q / w;
  graw[0] = g[SIEG_B] * gl_sqrt(q);
  gq += g[SIEG_B] * b / (T(2) * q);
  ellip_bwd(raw[1], raw[2], T(0.9999), g[SIEG_PHI], gq, graw[1], graw[2]);
  graw[3] = g[SIEG_CX]; graw[4] = g[SIEG_CY];
}
template <class T, int NP>
GL_HD void sie_fwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, T* ax, T* ay) {
  T c = d[SIE_C], s = d[SIE_S], q = d[SIE_Q], w = d[SIE_W], bw = d[SIE_B] / d[SIE_W];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - d[SIE_CX], dy = y[j] - d[SIE_CY];
    T xr = dx * c + dy * s, yr = -dx * s + dy * c;
    T psi = gl_sqrt(q * q * (xr * xr) + yr * yr);
    T fx = bw * gl_atan(w * xr / psi);
    T fy = bw * gl_atanh(w * yr / psi);
    ax[j] = fx * c - fy * s; ay[j] = fx * s + fy * c;
  }
}
template <class T, int NP>
GL_HD void sie_bwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, const T* gax, const T* gay, T* g) {
  T c = d[SIE_C], s = d[SIE_S], q = d[SIE_Q], w = d[SIE_W], b = d[SIE_B];
  T bw = b / w;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - d[SIE_CX], dy = y[j] - d[SIE_CY];
    T xr = dx * c + dy * s, yr = -dx * s + dy * c;
    T psi2 = q * q * (xr * xr) + yr * yr;
    T psi = gl_sqrt(psi2);
    T ux = w * xr / psi, uy = w * yr / psi;
    T atx = gl_atan(ux), aty = gl_atanh(uy);
    T fx = bw * atx, fy = bw * aty;
    T ax = fx * c - fy * s, ay = fx * s + fy * c;
    g[SIEG_PHI] += -gax[j] * ay + gay[j] * ax;
    T gfx = gax[j] * c + gay[j] * s, gfy = -gax[j] * s + gay[j] * c;
    T gbw = gfx * atx + gfy * aty;
    g[SIEG_B] += gbw / w;
    T gw = -gbw * bw / w;
    T gux = gfx * bw / (T(1) + ux * ux);
    T guy = gfy * bw / (T(1) - uy * uy);
    // u = w * r / psi
    gw += (gux * xr + guy * yr) / psi;
    T gxr = gux * w / psi, gyr = guy * w / psi;
    T gpsi = -(gux * ux + guy * uy) / psi;
    // psi = sqrt(q^2 xr^2 + yr^2)
    T gpsi2 = gpsi / (T(2) * psi);
    g[SIEG_Q] += gpsi2 * T(2) * q * xr * xr;
    gxr += gpsi2 * T(2) * q * q * xr;
    gyr += gpsi2 * T(2) * yr;
    g[SIEG_W] += gw;
    g[SIEG_PHI] += gxr * yr - gyr * xr;
    T gdx = gxr * c - gyr * s, gdy = gxr * s + gyr * c;
    g[SIEG_CX] -= gdx; g[SIEG_CY] -= gdy;
  }
}

// =============================================================================================
// NFW / NFW_ELLIPSE  (tf/profiles/mass/nfw.py:15-52, 106-134)
//   raw  : Rs, alpha_Rs, [e1, e2,] cx, cy
//   d[]  : cx, cy, cos, sin, s1 = sqrt(1-e), s2 = sqrt(1+e), Rs_c = max(Rs, 1e-7), pref = 4 rho0 Rs_c
//   dvars: cx, cy, phi, s1, s2, Rs_c, pref
//   (spherical NFW: cos = 1, sin = 0, s1 = s2 = 1 and their cotangents are dropped)
// =============================================================================================
enum { NFW_CX = 0, NFW_CY, NFW_C, NFW_S, NFW_S1, NFW_S2, NFW_RS, NFW_PREF, NFW_SIZE = 8 };
enum { NFWG_CX = 0, NFWG_CY, NFWG_PHI, NFWG_S1, NFWG_S2, NFWG_RS, NFWG_PREF };
#define GL_NFW_RMIN 0.0000001
#define GL_NFW_C 0.000001
#define GL_ONE_MINUS_LN2 0.30685281944005469
template <class T>
GL_HD void nfw_prep(const T* raw, T* d, bool ellipse) {
  T Rs = raw[0], alpha_Rs = raw[1];
  T rho0 = alpha_Rs / (T(4) * Rs * Rs * T(GL_ONE_MINUS_LN2));
  T Rs_c = gl_max(T(GL_NFW_RMIN), Rs);
  d[NFW_RS] = Rs_c;
  d[NFW_PREF] = T(4) * rho0 * Rs_c;
  if (ellipse) {
    T phi, q, c;
    ellip_fwd(raw[2], raw[3], T(0.9999), phi, q, c);
    T e = gl_abs(T(1) - q * q) / (T(1) + q * q);
    d[NFW_C] = gl_cos(phi); d[NFW_S] = gl_sin(phi);
    d[NFW_S1] = gl_sqrt(T(1) - e); d[NFW_S2] = gl_sqrt(T(1) + e);
    d[NFW_CX] = raw[4]; d[NFW_CY] = raw[5];
  } else {
    d[NFW_C] = T(1); d[NFW_S] = T(0); d[NFW_S1] = T(1); d[NFW_S2] = T(1);
    d[NFW_CX] = raw[2]; d[NFW_CY] = raw[3];
  }
}
template <class T>
GL_HD void nfw_prep_bwd(const T* raw, const T* d, const T* g, T* graw, bool ellipse) {
  T Rs = raw[0], alpha_Rs = raw[1];
  T k = T(4) * T(GL_ONE_MINUS_LN2);
  T Rs_c = d[NFW_RS];
  // pref = 4 rho0 Rs_c, rho0 = alpha_Rs / (k Rs^2)
  T rho0 = alpha_Rs / (k * Rs * Rs);
  T gRs_c = g[NFWG_RS] + g[NFWG_PREF] * T(4) * rho0;
  T grho0 = g[NFWG_PREF] * T(4) * Rs_c;
  graw[1] = grho0 / (k * Rs * Rs);
  graw[0] = -grho0 * T(2) * rho0 / Rs + ((Rs >= T(GL_NFW_RMIN)) ? gRs_c : T(0));
  if (ellipse) {
    T phi, q, c;
    ellip_fwd(raw[2], raw[3], T(0.9999), phi, q, c);
    T e = gl_abs(T(1) - q * q) / (T(1) + q * q);
    T ge = -g[NFWG_S1] / (T(2) * d[NFW_S1]) + g[NFWG_S2] / (T(2) * d[NFW_S2]);
    // e = (1-q^2)/(1+q^2) for q <= 1 (always, since c >= 0): de/dq = -4q/(1+q^2)^2
    T opq2 = T(1) + q * q;
    T gq = ge * (-T(4) * q / (opq2 * opq2));
    (void)e;
    ellip_bwd(raw[2], raw[3], T(0.9999), g[NFWG_PHI], gq, graw[2], graw[3]);
    graw[4] = g[NFWG_CX]; graw[5] = g[NFWG_CY];
  } else {
    graw[2] = g[NFWG_CX]; graw[3] = g[NFWG_CY];
  }
}
// g(X) of nfw.py:34-52 and its derivative; X == 1 keeps the reference's 1.0.
template <class T>
GL_HD T nfw_g(T X, T& dg) {
  if (X < T(1)) {
    T om = T(1) - X * X;
    T r = gl_sqrt(om);
    T ach = gl_acosh(T(1) / X);
    dg = T(1) / X + X * ach / (om * r) - T(1) / (X * om);
    return gl_log(X / T(2)) + ach / r;
  } else if (X > T(1)) {
    T om = X * X - T(1);
    T r = gl_sqrt(om);
    T ac = gl_acos(T(1) / X);
    dg = T(1) / X - X * ac / (om * r) + T(1) / (X * om);
    return gl_log(X / T(2)) + ac / r;
  }
  dg = T(0);
  return T(1);
}
template <class T, int NP>
GL_HD void nfw_fwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, T* ax, T* ay) {
  T c = d[NFW_C], s = d[NFW_S], s1 = d[NFW_S1], s2 = d[NFW_S2];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - d[NFW_CX], dy = y[j] - d[NFW_CY];
    T xr = (dx * c + dy * s) * s1, yr = (-dx * s + dy * c) * s2;
    T R = gl_max(T(GL_NFW_RMIN), gl_sqrt(xr * xr + yr * yr));
    T X = gl_max(T(GL_NFW_C), R / d[NFW_RS]);
    T dg;
    T a = d[NFW_PREF] * nfw_g(X, dg) / ((R / d[NFW_RS]) * (R / d[NFW_RS]));
    T fx = a * xr * s1, fy = a * yr * s2;
    ax[j] = fx * c - fy * s; ay[j] = fx * s + fy * c;
  }
}
template <class T, int NP>
GL_HD void nfw_bwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, const T* gax, const T* gay, T* g) {
  T c = d[NFW_C], s = d[NFW_S], s1 = d[NFW_S1], s2 = d[NFW_S2], Rs = d[NFW_RS], pref = d[NFW_PREF];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - d[NFW_CX], dy = y[j] - d[NFW_CY];
    T xr0 = dx * c + dy * s, yr0 = -dx * s + dy * c;
    T xr = xr0 * s1, yr = yr0 * s2;
    T R0 = gl_sqrt(xr * xr + yr * yr);
    T R = gl_max(T(GL_NFW_RMIN), R0);
    T X0 = R / Rs;
    T X = gl_max(T(GL_NFW_C), X0);
    T dg;
    T gx = nfw_g(X, dg);
    T a = pref * gx / (X0 * X0);
    T fx = a * xr * s1, fy = a * yr * s2;
    T ax = fx * c - fy * s, ay = fx * s + fy * c;
    g[NFWG_PHI] += -gax[j] * ay + gay[j] * ax;
    T gfx = gax[j] * c + gay[j] * s, gfy = -gax[j] * s + gay[j] * c;
    // fx = a xr s1, fy = a yr s2
    T ga = gfx * xr * s1 + gfy * yr * s2;
    T gxr = gfx * a * s1, gyr = gfy * a * s2;
    g[NFWG_S1] += gfx * a * xr;
    g[NFWG_S2] += gfy * a * yr;
    // a = pref g(X) / X0^2
    g[NFWG_PREF] += ga * gx / (X0 * X0);
    T gX0 = -T(2) * ga * a / X0 + ((X0 >= T(GL_NFW_C)) ? ga * pref * dg / (X0 * X0) : T(0));
    // X0 = R / Rs
    g[NFWG_RS] += -gX0 * X0 / Rs;
    T gR = gX0 / Rs;
    if (R0 >= T(GL_NFW_RMIN) && R0 > T(0)) {
      gxr += gR * xr / R0; gyr += gR * yr / R0;
    }
    g[NFWG_S1] += gxr * xr0;
    g[NFWG_S2] += gyr * yr0;
    T gxr0 = gxr * s1, gyr0 = gyr * s2;
    g[NFWG_PHI] += gxr0 * yr0 - gyr0 * xr0;
    T gdx = gxr0 * c - gyr0 * s, gdy = gxr0 * s + gyr0 * c;
    g[NFWG_CX] -= gdx; g[NFWG_CY] -= gdy;
  }
}

// ---- branch-free fp32 g(X) for the float lane types (float and the two-pixel pack GlF2) -------------------------------------
// g(X) = ln(X/2) + h(X),  h = acosh(1/X)/sqrt(1-X^2) (X < 1), acos(1/X)/sqrt(X^2-1) (X > 1)          (nfw.py:34-52)
// The libm transcription above (logf, acoshf / acosf, sqrt, divisions, three-way branch) cost ~150 scalar instructions per pixel
// and, being scalar, four times that per thread batch: after the member loop left the adjoint kernel (tape) it was three
// quarters of that kernel's arithmetic.  With u = (1-X)/(1+X):  acosh(1/X) = 2 atanh(sqrt u), acos(1/X) = 2 atan(sqrt -u) and
// sqrt|1-X^2| = (1+X) sqrt|u|, so  h = 2 S(u)/(1+X)  with ONE analytic function S(u) = sum u^n/(2n+1) on both sides of X = 1.
// Three per-lane regimes, all evaluated and selected (no branch):
//   A  |u| <= 0.1 (0.818 <= X <= 1.222): S = 1 + u S1, S1 = sum u^n/(2n+3) (8 terms, 6e-10).  No 0/0 at X = 1: neither in g nor in
//      dg/dX = X (h-1)/(1-X^2), where (h-1)/(1-X^2) = (2 S1 + 1 + X)/(1+X)^3.
//   B  X < 0.818: with r = sqrt(1-X^2), w = 1 - r = X^2/(1+r):  g = [ln(1 - w/2) - w ln(X/2)]/r.  The transcription subtracts two
//      O(|ln X|) numbers to get an O(X^2 |ln X|) result (relative error 2e-3 at X = 0.01 in fp32); this form has no cancellation.
//      ln(1 - z) is 2 atanh-series terms of q = z/(2-z) for z <= 0.1 and lg2 otherwise.
//   C  X > 1.222: h = 2 atan(t)/((1+X) t), t = sqrt(-u) in (0.316, 1), atan by the minimax polynomial of gl_atan2_fast.
// X == 1 exactly keeps the reference's g = 1, dg = 0.  Checked against the fp64 oracle on the host harness.
template <class V>
GL_HD void nfw_g_fast(V X, V& g, V& dg) {
  const V one(1.f), two(2.f);
  const V X2 = X * X;
  const V L = V((float)GL_LN2) * gl_log2_fast(X * V(0.5f));
  const V i1x = gl_div_fast(one, one + X);
  const V u = (one - X) * i1x;
  // A
  V S1 = V(1.f / 17.f);
  S1 = gl_fma(S1, u, V(1.f / 15.f)); S1 = gl_fma(S1, u, V(1.f / 13.f)); S1 = gl_fma(S1, u, V(1.f / 11.f));
  S1 = gl_fma(S1, u, V(1.f / 9.f)); S1 = gl_fma(S1, u, V(1.f / 7.f)); S1 = gl_fma(S1, u, V(1.f / 5.f));
  S1 = gl_fma(S1, u, V(1.f / 3.f));
  const V hA = two * gl_fma(u, S1, one) * i1x;
  const V qA = gl_fma(two, S1, one + X) * (i1x * i1x * i1x);
  // B
  const V omB = gl_max(one - X2, V(1e-3f));               // clamped: lanes of the other regimes must stay finite
  const V irB = gl_rsqrt_fast(omB), rB = omB * irB;
  const V w = X2 * gl_div_fast(one, one + rB), z = w * V(0.5f);
  const V q = z * gl_div_fast(one, two - z), q2 = q * q;
  const V ln_s = -(two * q) * gl_fma(q2, gl_fma(q2, gl_fma(q2, V(1.f / 7.f), V(0.2f)), V(1.f / 3.f)), one);
  const V ln_l = V((float)GL_LN2) * gl_log2_fast(one - z);
  const V ln1mz = gl_where_gt(z, 0.1f, ln_l, ln_s);
  const V gB = (ln1mz - w * L) * irB;
  const V qB = (gB - L - one) * (irB * irB);
  // C
  const V nu = gl_max(-u, V(1e-3f));
  const V it = gl_rsqrt_fast(nu), t = nu * it;
  const V hC = two * gl_atan2_fast(t, one) * it * i1x;
  const V qC = gl_div_fast(hC - one, gl_min(one - X2, V(-1e-3f)));
  // select
  const V gAC = L + gl_where_gt(u, -0.1f, hA, hC);
  g = gl_where_gt(u, 0.1f, gB, gAC);
  const V qq = gl_where_gt(u, 0.1f, qB, gl_where_gt(u, -0.1f, qA, qC));
  dg = X * qq;
  // X == 1: g = 1, dg = 0
  g = gl_where_in(X, 1.f, 1.f, one, g);
  dg = gl_where_in(X, 1.f, 1.f, V(0.f), dg);
}
template <class V, int NP>
GL_HD void nfw_fwd_fast(const float* d, const V* x, const V* y, V* ax, V* ay) {
  const V c(d[NFW_C]), s(d[NFW_S]), s1(d[NFW_S1]), s2(d[NFW_S2]), cx(d[NFW_CX]), cy(d[NFW_CY]), pref(d[NFW_PREF]);
  const V iRs(1.f / d[NFW_RS]);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const V dx = x[j] - cx, dy = y[j] - cy;
    const V xr = gl_fma(dx, c, dy * s) * s1, yr = gl_fma(dy, c, -(dx * s)) * s2;
    const V R2 = gl_max(gl_fma(xr, xr, yr * yr), V((float)(GL_NFW_RMIN * GL_NFW_RMIN)));   // R = max(r_min, sqrt(.))
    const V R = R2 * gl_rsqrt_fast(R2);
    const V X0 = R * iRs, X = gl_max(V((float)GL_NFW_C), X0);
    V g, dg;
    nfw_g_fast<V>(X, g, dg);
    const V a = pref * g * gl_div_fast(V(1.f), X0 * X0);
    const V fx = a * xr * s1, fy = a * yr * s2;
    ax[j] = gl_fma(fx, c, -(fy * s)); ay[j] = gl_fma(fx, s, fy * c);
  }
}
template <class V, int NP>
GL_HD void nfw_bwd_fast(const float* d, const V* x, const V* y, const V* gax, const V* gay, V* g) {
  const V c(d[NFW_C]), s(d[NFW_S]), s1(d[NFW_S1]), s2(d[NFW_S2]), cx(d[NFW_CX]), cy(d[NFW_CY]), pref(d[NFW_PREF]);
  const V iRs(1.f / d[NFW_RS]), zero(0.f);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const V dx = x[j] - cx, dy = y[j] - cy;
    const V xr0 = gl_fma(dx, c, dy * s), yr0 = gl_fma(dy, c, -(dx * s));
    const V xr = xr0 * s1, yr = yr0 * s2;
    const V R2u = gl_fma(xr, xr, yr * yr);
    const V R2 = gl_max(R2u, V((float)(GL_NFW_RMIN * GL_NFW_RMIN)));
    const V iR = gl_rsqrt_fast(R2), R = R2 * iR;
    const V X0 = R * iRs, X = gl_max(V((float)GL_NFW_C), X0);
    V gx, dg;
    nfw_g_fast<V>(X, gx, dg);
    const V iX0 = gl_div_fast(V(1.f), X0), iX02 = iX0 * iX0;
    const V a = pref * gx * iX02;
    const V fx = a * xr * s1, fy = a * yr * s2;
    const V ax = gl_fma(fx, c, -(fy * s)), ay = gl_fma(fx, s, fy * c);
    g[NFWG_PHI] += gl_fma(gay[j], ax, -(gax[j] * ay));
    const V gfx = gl_fma(gax[j], c, gay[j] * s), gfy = gl_fma(gay[j], c, -(gax[j] * s));
    const V ga = gl_fma(gfx * xr, s1, gfy * yr * s2);                      // fx = a xr s1, fy = a yr s2
    V gxr = gfx * a * s1, gyr = gfy * a * s2;
    g[NFWG_S1] += gfx * a * xr;
    g[NFWG_S2] += gfy * a * yr;
    g[NFWG_PREF] += ga * gx * iX02;                                       // a = pref g(X) / X0^2
    const V gX0 = gl_fma(V(-2.f) * ga, a * iX0, gl_where_lt(X0, (float)GL_NFW_C, zero, ga * pref * dg * iX02));
    g[NFWG_RS] -= gX0 * X0 * iRs;                                         // X0 = R / Rs
    const V gR = gX0 * iRs;
    // R = max(r_min, R0): the clamp passes no gradient; above it iR = 1/R0
    const V gRi = gl_where_lt(R2u, (float)(GL_NFW_RMIN * GL_NFW_RMIN), zero, gR * iR);
    gxr = gl_fma(gRi, xr, gxr); gyr = gl_fma(gRi, yr, gyr);
    g[NFWG_S1] += gxr * xr0;
    g[NFWG_S2] += gyr * yr0;
    const V gxr0 = gxr * s1, gyr0 = gyr * s2;
    g[NFWG_PHI] += gl_fma(gxr0, yr0, -(gyr0 * xr0));
    g[NFWG_CX] -= gl_fma(gxr0, c, -(gyr0 * s));
    g[NFWG_CY] -= gl_fma(gxr0, s, gyr0 * c);
  }
}

// Lane dispatch: fp64 and the forward-mode dual lanes (positions kernels, host harness) run the libm transcription; the fp32 lanes
// of the pixel kernels (float, and GlF2 = two pixels per FFMA2) run the branch-free form.
template <class T, int NP>
struct NfwLane {
  typedef typename gl_scalar_of<T>::type S;
  static GL_HD void fwd(const S* d, const T* x, const T* y, T* ax, T* ay) { nfw_fwd<T, NP>(d, x, y, ax, ay); }
  static GL_HD void bwd(const S* d, const T* x, const T* y, const T* gax, const T* gay, T* g) { nfw_bwd<T, NP>(d, x, y, gax, gay, g); }
};
template <int NP>
struct NfwLane<float, NP> {
  static GL_HD void fwd(const float* d, const float* x, const float* y, float* ax, float* ay) { nfw_fwd_fast<float, NP>(d, x, y, ax, ay); }
  static GL_HD void bwd(const float* d, const float* x, const float* y, const float* gax, const float* gay, float* g) {
    nfw_bwd_fast<float, NP>(d, x, y, gax, gay, g);
  }
};
template <int NP>
struct NfwLane<GlF2, NP> {
  static GL_HD void fwd(const float* d, const GlF2* x, const GlF2* y, GlF2* ax, GlF2* ay) { nfw_fwd_fast<GlF2, NP>(d, x, y, ax, ay); }
  static GL_HD void bwd(const float* d, const GlF2* x, const GlF2* y, const GlF2* gax, const GlF2* gay, GlF2* g) {
    nfw_bwd_fast<GlF2, NP>(d, x, y, gax, gay, g);
  }
};

// =============================================================================================
// dPIS / dPIE  (tf/profiles/mass/piemd.py:33-60, 105-119, 183-255)
//   raw  : theta_E, r_core, r_cut, [e1, e2,] cx, cy
//   d[]  : cx, cy, cos, sin, scale = theta_E r_cut/(r_cut - r_core), rc, rt (sorted, floored),
//          e, (dPIE only; the pixel code derives sqrt(e), q, ... from e)
//   dvars: cx, cy, phi, scale, rc, rt, e
// =============================================================================================
enum { DP_CX = 0, DP_CY, DP_C, DP_S, DP_SCALE, DP_RC, DP_RT, DP_E,
       DP_SQE, DP_Q, DP_IQ, DP_IOPE2, DP_IOME2, DP_ZCI, DP_RC2, DP_RT2,
       DP_M = 16, DP_FAST = 25, DP_ITHETA = 26, DP_SIZE = 28 };   // DP_FAST / DP_ITHETA: member 0 of a forward-mode group only (gl_sample_prep)   // DP_M: 3x3 d(scale, rc, rt)/d(base scaling params) of a scaling-relation member   // second row: per-member constants of e, rc, rt
enum { DPG_CX = 0, DPG_CY, DPG_PHI, DPG_SCALE, DPG_RC, DPG_RT, DPG_E };
#define GL_DPIE_RMIN 0.0001
// _sort_ra_rs (piemd.py:51-60): returns the sorted/floored radii and the selection pattern needed
// by the adjoint.  sel bit0: r_core < r_cut (kept), bit1: r_core floored, bit2: r_cut bumped.
template <class T>
GL_HD void dpie_sort(T r_core, T r_cut, T& rc, T& rt) {
  T a = (r_core < r_cut) ? r_core : r_cut;
  T b = (a > r_cut) ? a : r_cut;
  a = gl_max(T(GL_DPIE_RMIN), a);
  b = (b > a + T(GL_DPIE_RMIN)) ? b : b + T(GL_DPIE_RMIN);
  rc = a; rt = b;
}
template <class T>
GL_HD void dpie_sort_bwd(T r_core, T r_cut, T grc, T grt, T& g_core, T& g_cut) {
  // a0 = where(r_core < r_cut, r_core, r_cut); b0 = where(a0 > r_cut, a0, r_cut) == r_cut always
  // (a0 <= r_cut), a1 = max(rmin, a0), b1 = b0 or b0 + rmin (derivative 1 either way).
  T a0 = (r_core < r_cut) ? r_core : r_cut;
  T ga0 = (a0 >= T(GL_DPIE_RMIN)) ? grc : T(0);
  g_core = T(0); g_cut = grt;
  if (r_core < r_cut) g_core += ga0; else g_cut += ga0;
}
template <class T>
GL_HD void dpie_prep(const T* raw, T* d, bool ellipse) {
  T rc, rt;
  dpie_sort(raw[1], raw[2], rc, rt);
  d[DP_SCALE] = raw[0] * rt / (rt - rc);
  d[DP_RC] = rc; d[DP_RT] = rt;
  if (ellipse) {
    T phi, q, e;
    ellip_fwd(raw[3], raw[4], T(0.9999), phi, q, e);
    d[DP_C] = gl_cos(phi); d[DP_S] = gl_sin(phi); d[DP_E] = e;
    d[DP_CX] = raw[5]; d[DP_CY] = raw[6];
    // constants of complex_deriv_dual (piemd.py:202-207) hoisted out of the pixel loop
    d[DP_SQE] = gl_sqrt(e);
    d[DP_Q] = q; d[DP_IQ] = T(1) / q;
    d[DP_IOPE2] = T(1) / ((T(1) + e) * (T(1) + e));
    d[DP_IOME2] = T(1) / ((T(1) - e) * (T(1) - e));
    d[DP_ZCI] = -T(0.5) * (T(1) - e * e) / d[DP_SQE];
  } else {
    d[DP_C] = T(1); d[DP_S] = T(0); d[DP_E] = T(0);
    d[DP_CX] = raw[3]; d[DP_CY] = raw[4];
    d[DP_SQE] = T(0); d[DP_Q] = T(1); d[DP_IQ] = T(1); d[DP_IOPE2] = T(1); d[DP_IOME2] = T(1); d[DP_ZCI] = T(0);
  }
  d[DP_RC2] = rc * rc; d[DP_RT2] = rt * rt;
}
template <class T>
GL_HD void dpie_prep_bwd(const T* raw, const T* d, const T* g, T* graw, bool ellipse) {
  T rc = d[DP_RC], rt = d[DP_RT], theta_E = raw[0];
  T dif = rt - rc;
  graw[0] = g[DPG_SCALE] * rt / dif;
  T grt = g[DPG_RT] + g[DPG_SCALE] * theta_E * (-rc) / (dif * dif);
  T grc = g[DPG_RC] + g[DPG_SCALE] * theta_E * rt / (dif * dif);
  dpie_sort_bwd(raw[1], raw[2], grc, grt, graw[1], graw[2]);
  if (ellipse) {
    // e = min(|e|, 0.9999) is ellip_fwd's `c`; q = (1-c)/(1+c) => feed gq = 0 and add gc by hand.
    T e1 = raw[3], e2 = raw[4];
    T r2 = e1 * e1 + e2 * e2;
    if (r2 > T(0)) {
      T craw = gl_sqrt(r2);
      T gcraw = (craw <= T(0.9999)) ? g[DPG_E] : T(0);
      graw[3] = gcraw * e1 / craw - g[DPG_PHI] * e2 / (T(2) * r2);
      graw[4] = gcraw * e2 / craw + g[DPG_PHI] * e1 / (T(2) * r2);
    } else { graw[3] = T(0); graw[4] = T(0); }
    graw[5] = g[DPG_CX]; graw[6] = g[DPG_CY];
  } else {
    graw[3] = g[DPG_CX]; graw[4] = g[DPG_CY];
  }
}
template <class T, int NP>
GL_HD void dpis_fwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, T* ax, T* ay) {
  T rc = d[DP_RC], rt = d[DP_RT], scale = d[DP_SCALE];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - d[DP_CX], dy = y[j] - d[DP_CY];
    T r2 = dx * dx + dy * dy;
    T fa = gl_sqrt(r2 + rc * rc) - rc - gl_sqrt(r2 + rt * rt) + rt;
    T ar = scale / r2 * fa;
    ax[j] = ar * dx; ay[j] = ar * dy;
  }
}
template <class T, int NP>
GL_HD void dpis_bwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, const T* gax, const T* gay, T* g) {
  T rc = d[DP_RC], rt = d[DP_RT], scale = d[DP_SCALE];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - d[DP_CX], dy = y[j] - d[DP_CY];
    T r2 = dx * dx + dy * dy;
    T sc = gl_sqrt(r2 + rc * rc), st = gl_sqrt(r2 + rt * rt);
    T fa = sc - rc - st + rt;
    T ar = scale / r2 * fa;
    T gar = gax[j] * dx + gay[j] * dy;
    T gdx = gax[j] * ar, gdy = gay[j] * ar;
    g[DPG_SCALE] += gar * fa / r2;
    T gfa = gar * scale / r2;
    T gr2 = -gar * ar / r2 + gfa * (T(0.5) / sc - T(0.5) / st);
    g[DPG_RC] += gfa * (rc / sc - T(1));
    g[DPG_RT] += gfa * (T(1) - rt / st);
    gdx += gr2 * T(2) * dx; gdy += gr2 * T(2) * dy;
    g[DPG_CX] -= gdx; g[DPG_CY] -= gdy;
  }
}

// complex_deriv_dual (piemd.py:201-255) on rotated coordinates; returns the unscaled (re, im).
// Per-member constants (sqrt(e), q, 1/q, 1/(1+-e)^2, zci, rc^2, rt^2) come from the derived block;
// the per-pixel work is 2 rsqrt, 1 rcp, 1 lg2 and one atan2.
template <class T>
GL_HD T gl_sqrt_pos(T x) { return x * gl_rsqrt_fast(x); }   // x > 0
template <class T>
struct DpieFw {
  T sc, st, isc, ist, a, b_, c_, d_, e_, f_, aa, bb, cc, dd, inorm, aaa, bbb, inorm2, zr_re, zr_im;   // isc = 1/sc, ist = 1/st
};
// T is the lane type (float, double, GlDual, or the two-pixel pack GlF2), S the scalar type of the constants.
template <class T>
GL_HD void dpie_core_fwd(const typename gl_scalar_of<T>::type* d, T x, T y, DpieFw<T>& W, T& re, T& im) {
  typedef typename gl_scalar_of<T>::type S;
  const S two_sqe = S(2) * d[DP_SQE];
  const T rem2 = gl_fma(x * x, T(d[DP_IOPE2]), y * y * T(d[DP_IOME2]));
  const T sc2 = T(d[DP_RC2]) + rem2, st2 = T(d[DP_RT2]) + rem2;   // > 0
  W.isc = gl_rsqrt_fast(sc2); W.ist = gl_rsqrt_fast(st2);        // the adjoints need the reciprocals as well
  W.sc = sc2 * W.isc; W.st = st2 * W.ist;
  const T yq = y * T(d[DP_IQ]);
  W.a = T(d[DP_Q]) * x;                             // znum_rc_re
  W.b_ = gl_fma(T(two_sqe), W.sc, -yq);             // znum_rc_im
  W.c_ = x;                                         // zden_rc_re
  W.d_ = T(two_sqe * d[DP_RC]) - y;                 // zden_rc_im
  W.e_ = gl_fma(T(two_sqe), W.st, -yq);             // znum_rcut_im
  W.f_ = T(two_sqe * d[DP_RT]) - y;                 // zden_rcut_im
  W.aa = gl_fma(W.a, W.c_, -(W.b_ * W.f_));
  W.bb = gl_fma(W.a, W.f_, W.b_ * W.c_);
  W.cc = gl_fma(W.a, W.c_, -(W.d_ * W.e_));
  W.dd = gl_fma(W.a, W.d_, W.c_ * W.e_);
  W.inorm = gl_div_fast(T(S(1)), gl_fma(W.cc, W.cc, W.dd * W.dd));
  W.aaa = gl_fma(W.aa, W.cc, W.bb * W.dd) * W.inorm;
  W.bbb = gl_fma(W.bb, W.cc, -(W.aa * W.dd)) * W.inorm;
  const T norm2 = gl_fma(W.aaa, W.aaa, W.bbb * W.bbb);
  W.inorm2 = gl_div_fast(T(S(1)), norm2);
  W.zr_re = T(S(0.5 * GL_LN2)) * gl_log2_fast(norm2);
  W.zr_im = gl_atan2_fast(W.bbb, W.aaa);
  re = -(T(d[DP_ZCI]) * W.zr_im);
  im = T(d[DP_ZCI]) * W.zr_re;
}
template <class T, int NP>
GL_HD void dpie_fwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, T* ax, T* ay) {
  const T c = T(d[DP_C]), s = T(d[DP_S]), scale = T(d[DP_SCALE]), cx = T(d[DP_CX]), cy = T(d[DP_CY]);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - cx, dy = y[j] - cy;
    T xr = gl_fma(dx, c, dy * s), yr = gl_fma(dy, c, -(dx * s));
    DpieFw<T> W; T re, im;
    dpie_core_fwd<T>(d, xr, yr, W, re, im);
    const T Ar = gl_fma(re, c, -(im * s)), Ai = gl_fma(re, s, im * c);   // (re, im) rotated back: alpha_m = scale * A
    ax[j] = scale * Ar;
    ay[j] = scale * Ai;
  }
}
template <class T, int NP>
GL_HD void dpie_bwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, const T* gax, const T* gay, T* g) {
  typedef typename gl_scalar_of<T>::type S;
  const S e_s = d[DP_E], sqe_s = d[DP_SQE];
  const T c = T(d[DP_C]), s = T(d[DP_S]), scale = T(d[DP_SCALE]), rc = T(d[DP_RC]), rt = T(d[DP_RT]);
  const T sqe = T(sqe_s), q = T(d[DP_Q]), iq = T(d[DP_IQ]), iope2 = T(d[DP_IOPE2]), iome2 = T(d[DP_IOME2]), zci = T(d[DP_ZCI]);
  const T cx = T(d[DP_CX]), cy = T(d[DP_CY]), iq2 = T(d[DP_IQ] * d[DP_IQ]), two = T(S(2)), half = T(S(0.5));
  // d(constant)/de, applied per pixel to fold every e-dependence into the single dvar e
  const S dsqe_s = S(0.5) / sqe_s;
  const T dsqe = T(dsqe_s);
  const T dzci = T(S(0.5) * (S(1) - e_s * e_s) / (sqe_s * sqe_s) * dsqe_s + e_s / sqe_s);   // d zci / d e
  const T dq = T(-S(2) / ((S(1) + e_s) * (S(1) + e_s)));
  const T diope2 = T(-S(2) * d[DP_IOPE2] / (S(1) + e_s)), diome2 = T(S(2) * d[DP_IOME2] / (S(1) - e_s));
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - cx, dy = y[j] - cy;
    T xr = gl_fma(dx, c, dy * s), yr = gl_fma(dy, c, -(dx * s));
    DpieFw<T> W; T re, im;
    dpie_core_fwd<T>(d, xr, yr, W, re, im);
    T ux = gl_fma(re, c, -(im * s)), uy = gl_fma(re, s, im * c);     // rotated back, unscaled
    g[DPG_SCALE] += gl_fma(gax[j], ux, gay[j] * uy);
    T gux = gax[j] * scale, guy = gay[j] * scale;
    g[DPG_PHI] += gl_fma(guy, ux, -(gux * uy));
    T gre = gl_fma(gux, c, guy * s), gim = gl_fma(guy, c, -(gux * s));
    // re = -zci zr_im ; im = zci zr_re
    T gzci = gl_fma(gim, W.zr_re, -(gre * W.zr_im));
    T gzr_im = -(gre * zci), gzr_re = gim * zci;
    // zr_re = 0.5 log(norm2), zr_im = atan2(bbb, aaa)
    T gaaa = gl_fma(gzr_re, W.aaa, -(gzr_im * W.bbb)) * W.inorm2;
    T gbbb = gl_fma(gzr_re, W.bbb, gzr_im * W.aaa) * W.inorm2;
    // aaa = (aa cc + bb dd)/norm, bbb = (bb cc - aa dd)/norm
    T gaa = gl_fma(gaaa, W.cc, -(gbbb * W.dd)) * W.inorm;
    T gbb = gl_fma(gaaa, W.dd, gbbb * W.cc) * W.inorm;
    T gnorm = -(gl_fma(gaaa, W.aaa, gbbb * W.bbb) * W.inorm);
    T gcc = gl_fma(gl_fma(gaaa, W.aa, gbbb * W.bb), W.inorm, gnorm * two * W.cc);
    T gdd = gl_fma(gl_fma(gaaa, W.bb, -(gbbb * W.aa)), W.inorm, gnorm * two * W.dd);
    // aa = a c - b f ; bb = a f + b c ; cc = a c - d e ; dd = a d + c e
    T ga = gl_fma(gaa + gcc, W.c_, gl_fma(gbb, W.f_, gdd * W.d_));
    T gb = gl_fma(gbb, W.c_, -(gaa * W.f_));
    T gc = gl_fma(gaa + gcc, W.a, gl_fma(gbb, W.b_, gdd * W.e_));
    T gd = gl_fma(gdd, W.a, -(gcc * W.e_));
    T ge_ = gl_fma(gdd, W.c_, -(gcc * W.d_));
    T gf = gl_fma(gbb, W.a, -(gaa * W.b_));
    // a = q x ; b = 2 sqe sc - y/q ; c = x ; d = 2 rc sqe - y ; e = 2 sqe st - y/q ; f = 2 rt sqe - y
    T gbe = gb + ge_;
    T gq = gl_fma(ga, xr, gbe * yr * iq2);
    T gx = gl_fma(ga, q, gc);
    T gy = -gl_fma(gbe, iq, gd + gf);
    T gsqe = two * gl_fma(gb, W.sc, gl_fma(ge_, W.st, gl_fma(gd, rc, gf * rt)));
    T gsc = gb * two * sqe, gst = ge_ * two * sqe;
    T isc = W.isc, ist = W.ist;
    g[DPG_RC] += gl_fma(gd * two, sqe, gsc * rc * isc);
    g[DPG_RT] += gl_fma(gf * two, sqe, gst * rt * ist);
    T grem2 = half * gl_fma(gsc, isc, gst * ist);
    // rem2 = x^2 iope2 + y^2 iome2
    gx = gl_fma(grem2 * two * xr, iope2, gx);
    gy = gl_fma(grem2 * two * yr, iome2, gy);
    T ge = grem2 * gl_fma(xr * xr, diope2, yr * yr * diome2);
    ge = gl_fma(gzci, dzci, ge);
    ge = gl_fma(gsqe, dsqe, ge);
    ge = gl_fma(gq, dq, ge);
    g[DPG_E] += ge;
    g[DPG_PHI] += gl_fma(gx, yr, -(gy * xr));
    T gdx = gl_fma(gx, c, -(gy * s)), gdy = gl_fma(gx, s, gy * c);
    g[DPG_CX] -= gdx; g[DPG_CY] -= gdy;
  }
}

// Forward-mode variant for scaling-relation members (scaling_relation.py:57-70): a catalogue member
// has exactly three free inputs -- the group's base (theta_E, r_core, r_cut) -- so instead of the
// reverse sweep (forward recompute + adjoint per member and a cotangent flush per member) the beta
// pass carries the 2x3 Jacobian of the group deflection w.r.t. the base parameters:
//     J[:, k] += sum_j  d(alpha_m)/d(scale, rc, rt)_j * M_m[j][k],
// with M_m = d(scale, rc, rt)_m / d(base)_k precomputed per (sample, member) in the derived block.
// d(alpha)/d(rc|rt) follows from z = num/den, zr = log z:  dz = (num' - z den')/den, dzr = dz / z.
// FAST (every member of the group has r_core < r_cut, the usual case; decided per sample in gl_sample_prep): the chain matrix M is
// sparse -- d(rc)/d(theta_E) = d(rt)/d(theta_E) = d(rc)/d(r_cut) = d(rt)/d(r_core) = 0 -- and the deflection is linear in theta_E, so
// column 0 of the Jacobian is (group deflection) / theta_E (formed by the caller) and columns 1, 2 take two terms each.
// The tangents come from the FACTORED form of z = (num_rc / den_rc) (den_rt / num_rt): log z = log num_rc - log den_rc - log num_rt
// + log den_rt, where only the imaginary parts depend on rc / rt (b' = 2 sqe rc / sc, d' = 2 sqe =: k;  e' = 2 sqe rt / st, f' = k):
//   d log z / d rc = i b'/num_rc - i k/den_rc = (b' b/N1 - k d/N2) + i (b' a/N1 - k c/N2)
//   d log z / d rt = -i e'/num_rt + i k/den_rt = (k f/N4 - e' e/N3) + i (k c/N4 - e' a/N3),     N1..N4 = squared moduli of the factors,
// and (re, im) = i zci log z: 20 packed operations and four MUFU rcp for both tangents instead of 45 through the quotient form.
// The forward value itself stays on dpie_core_fwd (bit-identical to dpie_fwd).
template <class T, int NP>
GL_HD void dpie_fwd_jac_fast(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, T* ax, T* ay, T (*Jx)[NP], T (*Jy)[NP]) {
  typedef typename gl_scalar_of<T>::type S;
  const T c = T(d[DP_C]), s = T(d[DP_S]), scale = T(d[DP_SCALE]), cx = T(d[DP_CX]), cy = T(d[DP_CY]);
  const S zs = d[DP_ZCI] * d[DP_SCALE], k = S(2) * d[DP_SQE];
  const T krc = T(k * d[DP_RC] * zs), krt = T(k * d[DP_RT] * zs), kz = T(k * zs), one = T(S(1));
  const S* M = d + DP_M;
  const T m1 = T(M[1]), m2 = T(M[2]), m4 = T(M[4]), m8 = T(M[8]);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - cx, dy = y[j] - cy;
    T xr = gl_fma(dx, c, dy * s), yr = gl_fma(dy, c, -(dx * s));
    DpieFw<T> W; T re, im;
    dpie_core_fwd<T>(d, xr, yr, W, re, im);
    const T Ar = gl_fma(re, c, -(im * s)), Ai = gl_fma(re, s, im * c);
    ax[j] = scale * Ar;
    ay[j] = scale * Ai;
    // tangents of scale * (re, im) w.r.t. rc and rt
    const T a2 = W.a * W.a, c2 = W.c_ * W.c_;
    const T iN1 = gl_div_fast(one, gl_fma(W.b_, W.b_, a2)), iN2 = gl_div_fast(one, gl_fma(W.d_, W.d_, c2));
    const T iN3 = gl_div_fast(one, gl_fma(W.e_, W.e_, a2)), iN4 = gl_div_fast(one, gl_fma(W.f_, W.f_, c2));
    const T P1 = krc * W.isc * iN1, P2 = kz * iN2, Q1 = krt * W.ist * iN3, Q2 = kz * iN4;
    const T tre0 = gl_fma(P2, W.c_, -(P1 * W.a)), tim0 = gl_fma(P1, W.b_, -(P2 * W.d_));
    const T tre1 = gl_fma(Q1, W.a, -(Q2 * W.c_)), tim1 = gl_fma(Q2, W.f_, -(Q1 * W.e_));
    const T B0r = gl_fma(tre0, c, -(tim0 * s)), B0i = gl_fma(tre0, s, tim0 * c);
    const T B1r = gl_fma(tre1, c, -(tim1 * s)), B1i = gl_fma(tre1, s, tim1 * c);
    Jx[1][j] = gl_fma(m1, Ar, gl_fma(m4, B0r, Jx[1][j]));
    Jy[1][j] = gl_fma(m1, Ai, gl_fma(m4, B0i, Jy[1][j]));
    Jx[2][j] = gl_fma(m2, Ar, gl_fma(m8, B1r, Jx[2][j]));
    Jy[2][j] = gl_fma(m2, Ai, gl_fma(m8, B1i, Jy[2][j]));
  }
}
template <class T, int NP>
GL_HD void dpie_fwd_jac(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, T* ax, T* ay, T (*Jx)[NP], T (*Jy)[NP]) {
  typedef typename gl_scalar_of<T>::type S;
  const T c = T(d[DP_C]), s = T(d[DP_S]), scale = T(d[DP_SCALE]), zci = T(d[DP_ZCI]), two_sqe = T(S(2) * d[DP_SQE]);
  const T cx = T(d[DP_CX]), cy = T(d[DP_CY]), bp0 = T(S(2) * d[DP_SQE] * d[DP_RC]), ep0 = T(S(2) * d[DP_SQE] * d[DP_RT]);
  const S* M = d + DP_M;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - cx, dy = y[j] - cy;
    T xr = gl_fma(dx, c, dy * s), yr = gl_fma(dy, c, -(dx * s));
    DpieFw<T> W; T re, im;
    dpie_core_fwd<T>(d, xr, yr, W, re, im);
    const T Ar = gl_fma(re, c, -(im * s)), Ai = gl_fma(re, s, im * c);   // (re, im) rotated back: alpha_m = scale * A
    ax[j] = scale * Ar;
    ay[j] = scale * Ai;
    // tangents of (re, im) w.r.t. rc and rt
    T tre[2], tim[2];
    const T zsc = zci * scale * W.inorm * W.inorm2;      // tangents below carry the member's scale
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      T naa, nbb, ncc, ndd;   // num' = naa + i nbb, den' = ncc + i ndd
      if (w == 0) {           // d/d rc: b' = 2 sqe rc/sc, d' = 2 sqe
        const T bp = bp0 * W.isc;
        naa = -(bp * W.f_); nbb = bp * W.c_; ncc = -(two_sqe * W.e_); ndd = W.a * two_sqe;
      } else {                // d/d rt: e' = 2 sqe rt/st, f' = 2 sqe
        const T ep = ep0 * W.ist;
        naa = -(W.b_ * two_sqe); nbb = W.a * two_sqe; ncc = -(W.d_ * ep); ndd = W.c_ * ep;
      }
      // t = num' - z den'
      const T tr = naa - gl_fma(W.aaa, ncc, -(W.bbb * ndd));
      const T ti = nbb - gl_fma(W.aaa, ndd, W.bbb * ncc);
      // dz = t conj(den) / |den|^2,  d(log z) = dz conj(z) / |z|^2: both real scalings (and zci) applied once
      const T dzr_ = gl_fma(tr, W.cc, ti * W.dd);
      const T dzi_ = gl_fma(ti, W.cc, -(tr * W.dd));
      const T lr = gl_fma(dzr_, W.aaa, dzi_ * W.bbb);
      const T li = gl_fma(dzi_, W.aaa, -(dzr_ * W.bbb));
      tre[w] = -(zsc * li); tim[w] = zsc * lr;
    }
    // rotate the two tangents back once; then every base parameter k is a 3-term combination per component
    const T B0r = gl_fma(tre[0], c, -(tim[0] * s)), B0i = gl_fma(tre[0], s, tim[0] * c);
    const T B1r = gl_fma(tre[1], c, -(tim[1] * s)), B1i = gl_fma(tre[1], s, tim[1] * c);
    {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        Jx[k][j] += gl_fma(T(M[k]), Ar, gl_fma(T(M[3 + k]), B0r, T(M[6 + k]) * B1r));
        Jy[k][j] += gl_fma(T(M[k]), Ai, gl_fma(T(M[3 + k]), B0i, T(M[6 + k]) * B1i));
      }
    }
  }
}

// =============================================================================================
// TNFW  (tf/profiles/mass/tnfw.py:10-62; the docs call it experimental)
//   raw  : Rs, alpha_Rs, r_trunc, cx, cy
//   d[]  : cx, cy, Rs, pref = 4 rho0 Rs = alpha_Rs / (Rs (1 - ln 2)), tau = r_trunc / Rs, and the tau-only
//          constants pre = tau^2/(tau^2+1)^2, c0 = tau pi + (tau^2-1) ln tau, c1 = (tau^2-1)/tau, ln tau
//   dvars: cx, cy, Rs, pref, tau
//   alpha = pref g(X, tau) / X^2 (x, y),  X = max(R, 0.001 Rs) / Rs,  s = sqrt(tau^2 + X^2),
//   L = ln(X / (tau + s)),  g = pre [ (tau^2 + 2 X^2 - 1) F(X) + c0 + s (c1 L - pi) ]
// =============================================================================================
enum { TNFW_CX = 0, TNFW_CY, TNFW_RS, TNFW_PREF, TNFW_TAU, TNFW_PRE, TNFW_C0, TNFW_C1, TNFW_LNT, TNFW_SIZE = 12 };
enum { TNFWG_CX = 0, TNFWG_CY, TNFWG_RS, TNFWG_PREF, TNFWG_TAU };
#define GL_PI 3.14159265358979323846
template <class T>
GL_HD void tnfw_prep(const T* raw, T* d) {
  const T Rs = raw[0], tau = raw[2] / raw[0];
  d[TNFW_CX] = raw[3]; d[TNFW_CY] = raw[4]; d[TNFW_RS] = Rs;
  d[TNFW_PREF] = raw[1] / (Rs * T(GL_ONE_MINUS_LN2));   // 4 rho0 Rs, rho0 = alpha_Rs / (4 Rs^2 (1 + ln 0.5))
  d[TNFW_TAU] = tau;
  const T t2 = tau * tau, lnt = gl_log(tau);
  d[TNFW_PRE] = t2 / ((t2 + T(1)) * (t2 + T(1)));
  d[TNFW_C0] = tau * T(GL_PI) + (t2 - T(1)) * lnt;
  d[TNFW_C1] = (t2 - T(1)) / tau;
  d[TNFW_LNT] = lnt;
  d[9] = T(0); d[10] = T(0); d[11] = T(0);
}
template <class T>
GL_HD void tnfw_prep_bwd(const T* raw, const T* d, const T* g, T* graw) {
  const T Rs = raw[0];
  graw[1] = g[TNFWG_PREF] / (Rs * T(GL_ONE_MINUS_LN2));
  graw[2] = g[TNFWG_TAU] / Rs;
  graw[0] = g[TNFWG_RS] - g[TNFWG_PREF] * d[TNFW_PREF] / Rs - g[TNFWG_TAU] * d[TNFW_TAU] / Rs;
  graw[3] = g[TNFWG_CX]; graw[4] = g[TNFWG_CY];
}
// F(X) of tnfw.py:42-62 (1 at X == 1 exactly, where the scatter leaves the initial value) and dF/dX
template <class T>
GL_HD T tnfw_F(T X, T& dF) {
  T F;
  if (X < T(1)) { T r = gl_sqrt(T(1) - X * X); F = gl_atanh(r) / r; }
  else if (X > T(1)) { T r = gl_sqrt(X * X - T(1)); F = gl_atan(r) / r; }
  else { dF = T(0); return T(1); }
  dF = (T(1) - X * X * F) / (X * (X * X - T(1)));
  return F;
}
// The bracket of g(X, tau) cancels to O(X^2 log X) for small X while its terms are O(1) (the same
// cancellation as NFW's ln(X/2) + F(X), with more terms): in fp32 the deflection inside X ~ 0.03 has no
// correct digits left.  The TNFW pixel functions therefore run in the promoted type (fp32 -> FP64, which
// is full-rate enough on B200 for an optional profile) and recompute the tau-only constants there.
template <class T> struct gl_promote { typedef T type; };
template <> struct gl_promote<float> { typedef double type; };
template <> struct gl_promote<GlDual<float> > { typedef GlDual<double> type; };
GL_HD double gl_up(float a) { return (double)a; }
GL_HD double gl_up(double a) { return a; }
GL_HD GlDual<double> gl_up(GlDual<float> a) { return GlDual<double>((double)a.v, (double)a.d); }
GL_HD GlDual<double> gl_up(GlDual<double> a) { return a; }
template <class T> struct gl_narrow_t { template <class C> static GL_HD T f(C c) { return (T)c; } };
template <class S> struct gl_narrow_t<GlDual<S> > { template <class C> static GL_HD GlDual<S> f(C c) { return GlDual<S>((S)c.v, (S)c.d); } };

template <class T>
struct TnfwW { T R0, X, s, L, F, dF, B, gx, a; bool floored; };
// dd[]: the derived block in the compute precision CS, tau-only constants recomputed
template <class S, class CS>
GL_HD void tnfw_consts(const S* d, CS* dd) {
  dd[TNFW_CX] = CS(d[TNFW_CX]); dd[TNFW_CY] = CS(d[TNFW_CY]); dd[TNFW_RS] = CS(d[TNFW_RS]);
  dd[TNFW_PREF] = CS(d[TNFW_PREF]);
  const CS tau = CS(d[TNFW_TAU]), t2 = tau * tau, lnt = gl_log(tau);
  dd[TNFW_TAU] = tau;
  dd[TNFW_PRE] = t2 / ((t2 + CS(1)) * (t2 + CS(1)));
  dd[TNFW_C0] = (t2 - CS(1)) * lnt;            // tau*pi is folded into pi (tau - s) = -pi X^2 / (tau + s) per pixel
  dd[TNFW_C1] = (t2 - CS(1)) / tau;
  dd[TNFW_LNT] = lnt;
}
template <class T, class S>
GL_HD void tnfw_core(const S* d, T dx, T dy, TnfwW<T>& W) {
  const T Rs = d[TNFW_RS], tau = d[TNFW_TAU];
  W.R0 = gl_sqrt(dx * dx + dy * dy);
  const T Rmin = T(0.001) * Rs;
  W.floored = W.R0 < Rmin;
  const T R = W.floored ? Rmin : W.R0;
  W.X = R / Rs;
  W.s = gl_sqrt(tau * tau + W.X * W.X);
  W.L = gl_log(W.X / (tau + W.s));
  W.F = tnfw_F(W.X, W.dF);
  W.B = (tau * tau + T(2) * W.X * W.X - T(1)) * W.F + T(d[TNFW_C0]) - T(GL_PI) * W.X * W.X / (tau + W.s) + W.s * T(d[TNFW_C1]) * W.L;
  W.gx = T(d[TNFW_PRE]) * W.B;
  W.a = T(d[TNFW_PREF]) * W.gx / (W.X * W.X);
}
template <class T, int NP>
GL_HD void tnfw_fwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, T* ax, T* ay) {
  typedef typename gl_promote<T>::type C;
  typedef typename gl_scalar_of<C>::type CS;
  CS dd[TNFW_SIZE];
  tnfw_consts(d, dd);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    C dx = gl_up(x[j]) - dd[TNFW_CX], dy = gl_up(y[j]) - dd[TNFW_CY];
    TnfwW<C> W;
    tnfw_core(dd, dx, dy, W);
    ax[j] = gl_narrow_t<T>::f(W.a * dx); ay[j] = gl_narrow_t<T>::f(W.a * dy);
  }
}
template <class T, int NP>
GL_HD void tnfw_bwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, const T* gax, const T* gay, T* g) {
  typedef typename gl_promote<T>::type C;
  typedef typename gl_scalar_of<C>::type CS;
  CS dd[TNFW_SIZE];
  tnfw_consts(d, dd);
  const C Rs = dd[TNFW_RS], tau = dd[TNFW_TAU], pref = dd[TNFW_PREF], pre = dd[TNFW_PRE], c1 = dd[TNFW_C1], lnt = dd[TNFW_LNT];
  const C t2 = tau * tau;
  const C dpre = C(2) * tau * (C(1) - t2) / ((t2 + C(1)) * (t2 + C(1)) * (t2 + C(1)));
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    C dx = gl_up(x[j]) - dd[TNFW_CX], dy = gl_up(y[j]) - dd[TNFW_CY];
    const C gx_ = gl_up(gax[j]), gy_ = gl_up(gay[j]);
    TnfwW<C> W;
    tnfw_core(dd, dx, dy, W);
    const C X = W.X, s = W.s, L = W.L;
    C ga = gx_ * dx + gy_ * dy;
    C gdx = gx_ * W.a, gdy = gy_ * W.a;
    g[TNFWG_PREF] += gl_narrow_t<T>::f(ga * W.gx / (X * X));
    const C ggx = ga * pref / (X * X);
    const C inner = c1 * L - C(GL_PI);
    const C dBdX = C(4) * X * W.F + (t2 + C(2) * X * X - C(1)) * W.dF + (X / s) * inner + s * c1 * (C(1) / X - X / (s * (tau + s)));
    const C dBdt = C(2) * tau * W.F + C(GL_PI) + C(2) * tau * lnt + c1 + (tau / s) * inner + s * (-c1 / s + L * (C(1) + C(1) / t2));
    const C gX = -C(2) * ga * W.a / X + ggx * pre * dBdX;
    C gRs = -gX * X / Rs;
    const C gR = gX / Rs;
    if (W.floored) {
      gRs += gR * C(0.001);
    } else {
      gdx += gR * dx / W.R0; gdy += gR * dy / W.R0;
    }
    g[TNFWG_TAU] += gl_narrow_t<T>::f(ggx * (dpre * W.B + pre * dBdt));
    g[TNFWG_RS] += gl_narrow_t<T>::f(gRs);
    g[TNFWG_CX] -= gl_narrow_t<T>::f(gdx); g[TNFWG_CY] -= gl_narrow_t<T>::f(gdy);
  }
}

// =============================================================================================
// dPIEP  (tf/profiles/mass/piep.py:17-56): dPIS with the ellipticity in the potential -- the
// spherical deflection evaluated at (x sqrt(1-e), y sqrt(1+e)) in the rotated frame and rescaled,
// exactly the construction of NFW_ELLIPSE around NFW.
//   raw  : theta_E, Ra, Rs, e1, e2, cx, cy        (deriv argument order, piep.py:33)
//   d[]  : cx, cy, cos, sin, s1 = sqrt(1-e), s2 = sqrt(1+e), scale = theta_E rt/(rt - rc), rc, rt
//   dvars: cx, cy, phi, s1, s2, scale, rc, rt
// =============================================================================================
enum { PP_CX = 0, PP_CY, PP_C, PP_S, PP_S1, PP_S2, PP_SCALE, PP_RC, PP_RT, PP_SIZE = 12 };
enum { PPG_CX = 0, PPG_CY, PPG_PHI, PPG_S1, PPG_S2, PPG_SCALE, PPG_RC, PPG_RT };
template <class T>
GL_HD void dpiep_prep(const T* raw, T* d) {
  T rc, rt;
  dpie_sort(raw[1], raw[2], rc, rt);
  d[PP_SCALE] = raw[0] * rt / (rt - rc);
  d[PP_RC] = rc; d[PP_RT] = rt;
  T phi, q, c;
  ellip_fwd(raw[3], raw[4], T(0.9999), phi, q, c);
  T e = gl_abs(T(1) - q * q) / (T(1) + q * q);
  d[PP_C] = gl_cos(phi); d[PP_S] = gl_sin(phi);
  d[PP_S1] = gl_sqrt(T(1) - e); d[PP_S2] = gl_sqrt(T(1) + e);
  d[PP_CX] = raw[5]; d[PP_CY] = raw[6];
  d[9] = T(0); d[10] = T(0); d[11] = T(0);
}
template <class T>
GL_HD void dpiep_prep_bwd(const T* raw, const T* d, const T* g, T* graw) {
  const T rc = d[PP_RC], rt = d[PP_RT], theta_E = raw[0];
  const T dif = rt - rc;
  graw[0] = g[PPG_SCALE] * rt / dif;
  const T grt = g[PPG_RT] + g[PPG_SCALE] * theta_E * (-rc) / (dif * dif);
  const T grc = g[PPG_RC] + g[PPG_SCALE] * theta_E * rt / (dif * dif);
  dpie_sort_bwd(raw[1], raw[2], grc, grt, graw[1], graw[2]);
  T phi, q, c;
  ellip_fwd(raw[3], raw[4], T(0.9999), phi, q, c);
  const T ge = -g[PPG_S1] / (T(2) * d[PP_S1]) + g[PPG_S2] / (T(2) * d[PP_S2]);
  const T opq2 = T(1) + q * q;
  const T gq = ge * (-T(4) * q / (opq2 * opq2));
  ellip_bwd(raw[3], raw[4], T(0.9999), g[PPG_PHI], gq, graw[3], graw[4]);
  graw[5] = g[PPG_CX]; graw[6] = g[PPG_CY];
}
template <class T, int NP>
GL_HD void dpiep_fwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, T* ax, T* ay) {
  const T c = d[PP_C], s = d[PP_S], s1 = d[PP_S1], s2 = d[PP_S2], scale = d[PP_SCALE], rc = d[PP_RC], rt = d[PP_RT];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - d[PP_CX], dy = y[j] - d[PP_CY];
    T xr = (dx * c + dy * s) * s1, yr = (-dx * s + dy * c) * s2;
    T r2 = xr * xr + yr * yr;
    T fa = gl_sqrt(r2 + rc * rc) - rc - gl_sqrt(r2 + rt * rt) + rt;
    T ar = scale / r2 * fa;
    T fx = ar * xr * s1, fy = ar * yr * s2;
    ax[j] = fx * c - fy * s; ay[j] = fx * s + fy * c;
  }
}
template <class T, int NP>
GL_HD void dpiep_bwd(const typename gl_scalar_of<T>::type* d, const T* x, const T* y, const T* gax, const T* gay, T* g) {
  const T c = d[PP_C], s = d[PP_S], s1 = d[PP_S1], s2 = d[PP_S2], scale = d[PP_SCALE], rc = d[PP_RC], rt = d[PP_RT];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    T dx = x[j] - d[PP_CX], dy = y[j] - d[PP_CY];
    T xr0 = dx * c + dy * s, yr0 = -dx * s + dy * c;
    T xr = xr0 * s1, yr = yr0 * s2;
    T r2 = xr * xr + yr * yr;
    T sc = gl_sqrt(r2 + rc * rc), st = gl_sqrt(r2 + rt * rt);
    T fa = sc - rc - st + rt;
    T ar = scale / r2 * fa;
    T fx = ar * xr * s1, fy = ar * yr * s2;
    T ax = fx * c - fy * s, ay = fx * s + fy * c;
    g[PPG_PHI] += -gax[j] * ay + gay[j] * ax;
    T gfx = gax[j] * c + gay[j] * s, gfy = -gax[j] * s + gay[j] * c;
    T gar = gfx * xr * s1 + gfy * yr * s2;
    T gxr = gfx * ar * s1, gyr = gfy * ar * s2;
    g[PPG_S1] += gfx * ar * xr;
    g[PPG_S2] += gfy * ar * yr;
    g[PPG_SCALE] += gar * fa / r2;
    T gfa = gar * scale / r2;
    T gr2 = -gar * ar / r2 + gfa * (T(0.5) / sc - T(0.5) / st);
    g[PPG_RC] += gfa * (rc / sc - T(1));
    g[PPG_RT] += gfa * (T(1) - rt / st);
    gxr += gr2 * T(2) * xr; gyr += gr2 * T(2) * yr;
    g[PPG_S1] += gxr * xr0;
    g[PPG_S2] += gyr * yr0;
    T gxr0 = gxr * s1, gyr0 = gyr * s2;
    g[PPG_PHI] += gxr0 * yr0 - gyr0 * xr0;
    T gdx = gxr0 * c - gyr0 * s, gdy = gxr0 * s + gyr0 * c;
    g[PPG_CX] -= gdx; g[PPG_CY] -= gdy;
  }
}

// =============================================================================================
// SERSIC / SERSIC_ELLIPSE  (tf/profiles/light/sersic.py:29-80)
//   raw  : R_sersic, n_sersic, [e1, e2,] cx, cy, Ie      (Ie == 1 under use_lstsq, sersic.py:31,76)
//   d[]  : cx, cy, cos, sin, sq = sqrt(q), isq = 1/sq, 1/R_sersic, 1/n, bn, Ie
//   dvars: cx, cy, phi, sq, invRs, invn, bn, Ie
// =============================================================================================
enum { SER_CX = 0, SER_CY, SER_C, SER_S, SER_SQ, SER_ISQ, SER_IRS, SER_IN, SER_BN, SER_IE, SER_NBL, SER_SIZE = 12 };
enum { SERG_CX = 0, SERG_CY, SERG_PHI, SERG_SQ, SERG_IRS, SERG_IN, SERG_BN, SERG_IE };
template <class T>
GL_HD void sersic_prep(const T* raw, T* d, bool ellipse, bool use_lstsq) {
  T Rs = raw[0], n = raw[1];
  d[SER_IRS] = T(1) / Rs;
  d[SER_IN] = T(1) / n;
  d[SER_BN] = T(1.9992) * n - T(0.3271);
  if (ellipse) {
    T phi, q, c;
    ellip_fwd(raw[2], raw[3], T(0.9999), phi, q, c);
    d[SER_C] = gl_cos(phi); d[SER_S] = gl_sin(phi);
    d[SER_SQ] = gl_sqrt(q); d[SER_ISQ] = T(1) / gl_sqrt(q);
    d[SER_CX] = raw[4]; d[SER_CY] = raw[5];
    d[SER_IE] = use_lstsq ? T(1) : raw[6];
  } else {
    d[SER_C] = T(1); d[SER_S] = T(0); d[SER_SQ] = T(1); d[SER_ISQ] = T(1);
    d[SER_CX] = raw[2]; d[SER_CY] = raw[3];
    d[SER_IE] = use_lstsq ? T(1) : raw[4];
  }
  d[SER_NBL] = -d[SER_BN] * T(GL_LOG2E); d[11] = T(0);
}
template <class T>
GL_HD void sersic_prep_bwd(const T* raw, const T* d, const T* g, T* graw, bool ellipse, bool use_lstsq) {
  T Rs = raw[0], n = raw[1];
  graw[0] = -g[SERG_IRS] / (Rs * Rs);
  graw[1] = -g[SERG_IN] / (n * n) + T(1.9992) * g[SERG_BN];
  if (ellipse) {
    T gq = g[SERG_SQ] / (T(2) * d[SER_SQ]);
    ellip_bwd(raw[2], raw[3], T(0.9999), g[SERG_PHI], gq, graw[2], graw[3]);
    graw[4] = g[SERG_CX]; graw[5] = g[SERG_CY];
    graw[6] = use_lstsq ? T(0) : g[SERG_IE];
  } else {
    graw[2] = g[SERG_CX]; graw[3] = g[SERG_CY];
    graw[4] = use_lstsq ? T(0) : g[SERG_IE];
  }
}
// (R/Rs)^(1/n) = 2^((0.5/n) log2(R^2/Rs^2)) and exp(-bn (p-1)) = 2^(-bn log2e (p-1)): no sqrt, no powf.
template <class V, int NP>
GL_HD void sersic_fwd(const typename gl_scalar_of<V>::type* d, const V* x, const V* y, V* out) {
  typedef typename gl_scalar_of<V>::type S;
  const V c = V(d[SER_C]), s = V(d[SER_S]), irs2 = V(d[SER_IRS] * d[SER_IRS]), hin = V(S(0.5) * d[SER_IN]);
  const V cx = V(d[SER_CX]), cy = V(d[SER_CY]), sq = V(d[SER_SQ]), isq = V(d[SER_ISQ]), ie = V(d[SER_IE]), nbl = V(d[SER_NBL]);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    V dx = x[j] - cx, dy = y[j] - cy;
    V xt1 = gl_fma(c, dx, s * dy) * sq;
    V xt2 = gl_fma(c, dy, -(s * dx)) * isq;
    V u2 = gl_fma(xt1, xt1, xt2 * xt2) * irs2;
    V p = gl_exp2_fast(hin * gl_log2_fast(u2));
    out[j] = gl_fma(ie, gl_exp2_fast(nbl * (p - V(S(1)))), out[j]);
  }
}
// gI: cotangent of the surface brightness.  gx/gy (may be null): += cotangent of the coordinates.
template <class V, int NP>
GL_HD void sersic_bwd(const typename gl_scalar_of<V>::type* d, const V* x, const V* y, const V* gI, V* g, V* gx, V* gy) {
  typedef typename gl_scalar_of<V>::type S;
  const V c = V(d[SER_C]), s = V(d[SER_S]), sq = V(d[SER_SQ]), isq = V(d[SER_ISQ]), irs = V(d[SER_IRS]), bn = V(d[SER_BN]);
  const V irs2 = irs * irs, hin = V(S(0.5) * d[SER_IN]), cx = V(d[SER_CX]), cy = V(d[SER_CY]), ie = V(d[SER_IE]), nbl = V(d[SER_NBL]);
  const V isq2 = isq * isq, two_irs = V(S(2)) * irs;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    V dx = x[j] - cx, dy = y[j] - cy;
    V xr = gl_fma(c, dx, s * dy), yr = gl_fma(c, dy, -(s * dx));
    V xt1 = xr * sq, xt2 = yr * isq;
    V r2 = gl_fma(xt1, xt1, xt2 * xt2);
    V u2 = r2 * irs2;
    V L2 = gl_log2_fast(u2);
    V p = gl_exp2_fast(hin * L2);
    V pm1 = p - V(S(1));
    V E = gl_exp2_fast(nbl * pm1);
    g[SERG_IE] += gI[j] * E;
    V garg = gI[j] * ie * E;
    g[SERG_BN] -= garg * pm1;
    V gpp = -(garg * bn * p);                      // gp * p
    // R = 0 (u2 == 0): the reference's gradient is NaN there (measure zero); contributes 0 here
    V L2s = gl_where_gt(u2, S(0), L2, V(S(0)));
    V gu2 = gl_where_gt(u2, S(0), gl_div_fast(gpp * hin, u2), V(S(0)));
    g[SERG_IN] += gpp * (V(S(0.5 * GL_LN2)) * L2s);
    g[SERG_IRS] += gu2 * r2 * two_irs;
    V gr2 = gu2 * irs2 * V(S(2));
    V gxt1 = gr2 * xt1, gxt2 = gr2 * xt2;
    g[SERG_SQ] += gl_fma(gxt1, xr, -(gxt2 * yr * isq2));
    V gxr = gxt1 * sq, gyr = gxt2 * isq;
    g[SERG_PHI] += gl_fma(gxr, yr, -(gyr * xr));
    V gdx = gl_fma(gxr, c, -(gyr * s)), gdy = gl_fma(gxr, s, gyr * c);
    g[SERG_CX] -= gdx; g[SERG_CY] -= gdy;
    if (gx) { gx[j] += gdx; gy[j] += gdy; }
  }
}

// =============================================================================================
// CORE_SERSIC  (tf/profiles/light/sersic.py:83-132), the formula AS WRITTEN there:
//   I = Ie (1 + (Rb/R)^alpha)^(gamma/alpha) exp(-bn (R^alpha + Rb^alpha) / (R_sersic^alpha alpha n) - 1)
//   -- `R_sersic ** alpha ** 1.0` is R_sersic^alpha, 1/(alpha n) is a divisor rather than an exponent and the -1 sits
//   outside the bn product (SURVEY App. B7 reads this as a slip of the reference; parity means reproducing it, and the
//   golden vectors of tests/golden/reference_golden.npz come from executing exactly that code).
//   raw  : R_sersic, n_sersic, Rb, alpha, gamma, e1, e2, cx, cy, Ie      (Ie == 1 under use_lstsq, :117)
//   d[]  : cx, cy, cos, sin, sq, isq, alpha, goa = gamma/alpha, Rba = Rb^alpha, K = bn / (R_sersic^alpha alpha n), Ie
//   dvars: cx, cy, phi, sq, alpha (through R^alpha only), goa, Rba, K  |  Ie (a 9th accumulator, flushed on its own)
// Only the generic interpreter instantiates it (float lanes and the fp64 host harness): libm log / exp, no fast paths.
// =============================================================================================
enum { CS_CX = 0, CS_CY, CS_C, CS_S, CS_SQ, CS_ISQ, CS_AL, CS_GOA, CS_RBA, CS_K, CS_IE, CS_SIZE = 12 };
enum { CSG_CX = 0, CSG_CY, CSG_PHI, CSG_SQ, CSG_AL, CSG_GOA, CSG_RBA, CSG_K, CSG_IE };
template <class T>
GL_HD void core_sersic_prep(const T* raw, T* d, bool use_lstsq) {
  const T Rs = raw[0], n = raw[1], Rb = raw[2], al = raw[3], ga = raw[4];
  T phi, q, c;
  ellip_fwd(raw[5], raw[6], T(0.9999), phi, q, c);
  d[CS_C] = gl_cos(phi); d[CS_S] = gl_sin(phi);
  d[CS_SQ] = gl_sqrt(q); d[CS_ISQ] = T(1) / gl_sqrt(q);
  d[CS_CX] = raw[7]; d[CS_CY] = raw[8];
  d[CS_AL] = al; d[CS_GOA] = ga / al;
  d[CS_RBA] = gl_pow(Rb, al);
  const T bn = T(1.9992) * n - T(0.3271);
  d[CS_K] = bn / (gl_pow(Rs, al) * al * n);
  d[CS_IE] = use_lstsq ? T(1) : raw[9];
  d[11] = T(0);
}
template <class T>
GL_HD void core_sersic_prep_bwd(const T* raw, const T* d, const T* g, T* graw, bool use_lstsq) {
  const T Rs = raw[0], n = raw[1], Rb = raw[2], al = raw[3], ga = raw[4];
  const T K = d[CS_K], Rba = d[CS_RBA];
  const T rsa = gl_pow(Rs, al);
  graw[0] = -g[CSG_K] * K * al / Rs;
  graw[1] = g[CSG_K] * (T(1.9992) / (rsa * al * n) - K / n);
  graw[2] = g[CSG_RBA] * al * Rba / Rb;
  graw[3] = g[CSG_AL] - g[CSG_GOA] * ga / (al * al) + g[CSG_RBA] * Rba * gl_log(Rb) - g[CSG_K] * K * (gl_log(Rs) + T(1) / al);
  graw[4] = g[CSG_GOA] / al;
  const T gq = g[CSG_SQ] / (T(2) * d[CS_SQ]);
  ellip_bwd(raw[5], raw[6], T(0.9999), g[CSG_PHI], gq, graw[5], graw[6]);
  graw[7] = g[CSG_CX]; graw[8] = g[CSG_CY];
  graw[9] = use_lstsq ? T(0) : g[CSG_IE];
}
template <class V, int NP>
GL_HD void core_sersic_fwd(const typename gl_scalar_of<V>::type* d, const V* x, const V* y, V* out) {
  typedef typename gl_scalar_of<V>::type S;
  const V c = V(d[CS_C]), s = V(d[CS_S]), cx = V(d[CS_CX]), cy = V(d[CS_CY]), sq = V(d[CS_SQ]), isq = V(d[CS_ISQ]);
  const V hal = V(S(0.5) * d[CS_AL]), goa = V(d[CS_GOA]), rba = V(d[CS_RBA]), K = V(d[CS_K]), ie = V(d[CS_IE]);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    V dx = x[j] - cx, dy = y[j] - cy;
    V xt1 = gl_fma(c, dx, s * dy) * sq;
    V xt2 = gl_fma(c, dy, -(s * dx)) * isq;
    V r2 = gl_fma(xt1, xt1, xt2 * xt2);
    V Ra = gl_exp(hal * gl_log(r2));                       // R^alpha
    V A = gl_exp(goa * gl_log(V(S(1)) + rba / Ra));        // (1 + (Rb/R)^alpha)^(gamma/alpha)
    out[j] = gl_fma(ie * A, gl_exp(-(K * (Ra + rba)) - V(S(1))), out[j]);
  }
}
// gI: cotangent of the surface brightness; g: the 8 dvars above; gIe: += cotangent of Ie.
template <class V, int NP>
GL_HD void core_sersic_bwd(const typename gl_scalar_of<V>::type* d, const V* x, const V* y, const V* gI, V* g, V* gIe, V* gx, V* gy) {
  typedef typename gl_scalar_of<V>::type S;
  const V c = V(d[CS_C]), s = V(d[CS_S]), cx = V(d[CS_CX]), cy = V(d[CS_CY]), sq = V(d[CS_SQ]), isq = V(d[CS_ISQ]);
  const V al = V(d[CS_AL]), hal = V(S(0.5) * d[CS_AL]), goa = V(d[CS_GOA]), rba = V(d[CS_RBA]), K = V(d[CS_K]), ie = V(d[CS_IE]);
  const V isq2 = isq * isq;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    V dx = x[j] - cx, dy = y[j] - cy;
    V xr = gl_fma(c, dx, s * dy), yr = gl_fma(c, dy, -(s * dx));
    V xt1 = xr * sq, xt2 = yr * isq;
    V r2 = gl_fma(xt1, xt1, xt2 * xt2);
    V lr2 = gl_log(r2);
    V Ra = gl_exp(hal * lr2);
    V t = rba / Ra;
    V l1t = gl_log(V(S(1)) + t);
    V AE = gl_exp(goa * l1t) * gl_exp(-(K * (Ra + rba)) - V(S(1)));
    gIe[0] += gI[j] * AE;
    V garg = gI[j] * ie * AE;                               // cotangent of ln I
    // R = 0: the reference's value is inf / NaN there (measure zero); contributes 0 here, like sersic_bwd
    V inv1t = V(S(1)) / (V(S(1)) + t);
    g[CSG_GOA] += gl_where_gt(r2, S(0), garg * l1t, V(S(0)));
    g[CSG_K] -= gl_where_gt(r2, S(0), garg * (Ra + rba), V(S(0)));
    g[CSG_RBA] += gl_where_gt(r2, S(0), garg * (goa * inv1t / Ra - K), V(S(0)));
    V gRa = gl_where_gt(r2, S(0), -garg * (goa * t * inv1t / Ra + K), V(S(0)));     // d lnI / d Ra, times garg
    g[CSG_AL] += gl_where_gt(r2, S(0), gRa * Ra * V(S(0.5)) * lr2, V(S(0)));
    V gr2 = gl_where_gt(r2, S(0), gRa * Ra * al / r2, V(S(0)));                     // 2 d/d(r2): gxt = gr2 * xt
    V gxt1 = gr2 * xt1, gxt2 = gr2 * xt2;
    g[CSG_SQ] += gl_fma(gxt1, xr, -(gxt2 * yr * isq2));
    V gxr = gxt1 * sq, gyr = gxt2 * isq;
    g[CSG_PHI] += gl_fma(gxr, yr, -(gyr * xr));
    V gdx = gl_fma(gxr, c, -(gyr * s)), gdy = gl_fma(gxr, s, gyr * c);
    g[CSG_CX] -= gdx; g[CSG_CY] -= gdy;
    if (gx) { gx[j] += gdx; gy[j] += gdy; }
  }
}

// =============================================================================================
// SHAPELETS  (tf/profiles/light/shapelets.py:20-85)
//   raw  : beta, center_x, center_y  (+ n_layers amplitudes, staged separately)
//   d[]  : cx, cy, 1/beta, n_max, then tables c1[n], c2[n], c3[n] (n = 0..n_max) of the normalised
//          Hermite recurrence, then the n_layers amplitudes in component order
//   dvars: cx, cy, 1/beta, amp_0 .. amp_{L-1}
// Basis (interpolate=False, :67-85): psi_n(u) = H_n(u) / sqrt(2^n sqrt(pi) n!) via the normalised
// recurrence psi_n = sqrt(2/n) u psi_{n-1} - sqrt((n-1)/n) psi_{n-2} (same polynomials as the
// reference's H_n recurrence times its prefactor; psi_n' = sqrt(2n) psi_{n-1}); component k of
// order o = n1 + n2 is exp(-(u^2+v^2)/2) psi_{n1}(u) psi_{n2}(v), k = o(o+1)/2 + n2  (:26-46).
// interpolate=True (:54-66): psi_n(u) exp(-u^2/2) is read by linear interpolation from a 6000-point
// table on [-5, 5] (tfp.math.interp_regular_1d_grid semantics, 0 outside).
// =============================================================================================
#define GL_SHP_MAXN 20
#define GL_SHP_NMAX_FAST 10   // orders up to this one (BASELINE's n_max = 10, the demo's 8) run the unrolled, register-resident instance
#define GL_SHP_TABLE_N 6000
enum { SHP_CX = 0, SHP_CY, SHP_IB, SHP_NMAX, SHP_TAB = 4 };
enum { SHPG_CX = 0, SHPG_CY, SHPG_IB, SHPG_AMP = 3 };
GL_HD int shp_layers(int nmax) { return (nmax + 1) * (nmax + 2) / 2; }
GL_HD int shp_amp_off(int nmax) { return SHP_TAB + 3 * (nmax + 1); }
GL_HD int shp_der_size(int nmax) { return (shp_amp_off(nmax) + shp_layers(nmax) + 3) & ~3; }

template <class T>
GL_HD void shp_prep(const T* raw, T* d, int nmax) {
  d[SHP_CX] = raw[1]; d[SHP_CY] = raw[2]; d[SHP_IB] = T(1) / raw[0]; d[SHP_NMAX] = T(nmax);
  T* c1 = d + SHP_TAB; T* c2 = c1 + (nmax + 1); T* c3 = c2 + (nmax + 1);
  c1[0] = T(0.75112554446494248);   // pi^(-1/4) = psi_0
  c2[0] = T(0); c3[0] = T(0);
  for (int n = 1; n <= nmax; ++n) {
    c1[n] = gl_sqrt(T(2) / T(n));
    c2[n] = gl_sqrt(T(n - 1) / T(n));
    c3[n] = gl_sqrt(T(2 * n));
  }
}

// h[n] = basis function n at u, dh[n] = its derivative (dh may be null).  Returns false when the
// point is outside the interpolation table (all components are 0 there).
// The loops below are bounded by the compile-time NMAXC (and cut at the run-time order by uniform tests): fully unrolled, the basis
// arrays hu / hv / dhu / dhv live in registers.  With run-time bounds they were indexed dynamically and lived in local memory
// (256 bytes of stack per thread): k_raytrace_comps executed ~2500 instructions per pixel for ~600 of real work.
template <class T, int NMAXC>
GL_HD void shp_basis_rec(const T* d, int nmax, T u, T* h, T* dh) {
  constexpr int UNR = (NMAXC <= 10) ? NMAXC + 1 : 1;   // unrolled for the fast instance only
  const T* c1 = d + SHP_TAB; const T* c2 = c1 + (nmax + 1); const T* c3 = c2 + (nmax + 1);
  h[0] = c1[0];
  if (dh) dh[0] = T(0);
  T hm2 = T(0), hm1 = c1[0];
#pragma unroll UNR
  for (int n = 1; n <= NMAXC; ++n) {
    if (n > nmax) break;
    T hn = c1[n] * u * hm1 - c2[n] * hm2;
    h[n] = hn;
    if (dh) dh[n] = c3[n] * hm1;
    hm2 = hm1; hm1 = hn;
  }
}
template <class T, int NMAXC>
GL_HD void shp_basis_interp(const float* table, int nmax, T u, T* h, T* dh) {
  constexpr int UNR = (NMAXC <= 10) ? NMAXC + 1 : 1;
  const T nm1 = T(GL_SHP_TABLE_N - 1);
  T xi_un = (u - T(-5)) / (T(5) - T(-5)) * nm1;
  bool outside = (xi_un < T(0)) || (xi_un > nm1) || gl_isnan(xi_un);
  T xi = gl_min(gl_max(xi_un, T(0)), nm1);
  T fl = T((int)xi);                                  // floor (xi >= 0)
  T above = gl_min(fl + T(1), nm1);
  T below = gl_max(above - T(1), T(0));
  int ib = (int)below, ia = (int)above;
  T t = xi - below;
  T slope_scale = nm1 / T(10);
#pragma unroll UNR
  for (int n = 0; n <= NMAXC; ++n) {
    if (n > nmax) break;
    T yb = T(table[n * GL_SHP_TABLE_N + ib]), ya = T(table[n * GL_SHP_TABLE_N + ia]);
    h[n] = outside ? T(0) : t * ya + (T(1) - t) * yb;
    if (dh) dh[n] = outside ? T(0) : (ya - yb) * slope_scale;
  }
}

// Surface brightness of a Shapelets profile at one point, its unit-amplitude components (comps != null: component k goes to
// comps[k * stride], NaN-scrubbed and zeroed where !keep like tf/simulator.py:200, scrubbed values counted in *n_nan), and the
// adjoint (gI != null).  NMAXC: compile-time bound of the order (see above).
template <class T, int NMAXC>
GL_HD T shp_point_ct(const T* d, const float* table, bool interp, int nmax, T x, T y, T* comps, int stride,
                     const T* gI, T* g, T* gamp, T* gx, T* gy, bool keep, int* n_nan) {
  constexpr int UNR = (NMAXC <= 10) ? NMAXC + 1 : 1;
  T hu[NMAXC + 1], hv[NMAXC + 1], dhu[NMAXC + 1], dhv[NMAXC + 1];
  const T ib = d[SHP_IB];
  const T dx = x - d[SHP_CX], dy = y - d[SHP_CY];
  const T u = dx * ib, v = dy * ib;
  T fac;
  if (interp) {
    shp_basis_interp<T, NMAXC>(table, nmax, u, hu, gI ? dhu : (T*)nullptr);
    shp_basis_interp<T, NMAXC>(table, nmax, v, hv, gI ? dhv : (T*)nullptr);
    fac = T(1);
  } else {
    shp_basis_rec<T, NMAXC>(d, nmax, u, hu, gI ? dhu : (T*)nullptr);
    shp_basis_rec<T, NMAXC>(d, nmax, v, hv, gI ? dhv : (T*)nullptr);
    fac = gl_exp(-(u * u + v * v) / T(2));
  }
  const T* amp = d + shp_amp_off(nmax);
  T I = T(0), Iu = T(0), Iv = T(0);
  int nn = 0;
#pragma unroll UNR
  for (int o = 0; o <= NMAXC; ++o) {
    if (o > nmax) break;
#pragma unroll UNR
    for (int n2 = 0; n2 <= o; ++n2) {
      const int n1 = o - n2, k = o * (o + 1) / 2 + n2;
      const T B = hu[n1] * hv[n2];
      if (comps) {
        const T c = fac * B;
        const bool bad = gl_isnan(c);
        nn += (keep && bad) ? 1 : 0;
        comps[(size_t)k * stride] = (keep && !bad) ? c : T(0);
        continue;
      }
      I += amp[k] * B;
      if (gI) {
        Iu += amp[k] * dhu[n1] * hv[n2];
        Iv += amp[k] * hu[n1] * dhv[n2];
        if (gamp) gamp[k] += gI[0] * fac * B;
      }
    }
  }
  if (comps) { if (n_nan) *n_nan += nn; return T(0); }
  I *= fac;
  if (gI) {
    T dIdu = fac * Iu, dIdv = fac * Iv;
    if (!interp) { dIdu -= u * I; dIdv -= v * I; }
    const T gu = gI[0] * dIdu, gv = gI[0] * dIdv;
    g[SHPG_IB] += gu * dx + gv * dy;
    const T gdx = gu * ib, gdy = gv * ib;
    g[SHPG_CX] -= gdx; g[SHPG_CY] -= gdy;
    if (gx) { *gx += gdx; *gy += gdy; }
  }
  return I;
}
template <class T>
GL_HD T shp_point(const T* d, const float* table, bool interp, int nmax, T x, T y, T* comps, int stride,
                  const T* gI, T* g, T* gamp, T* gx, T* gy, bool keep = true, int* n_nan = nullptr) {
  if (nmax <= GL_SHP_NMAX_FAST) return shp_point_ct<T, GL_SHP_NMAX_FAST>(d, table, interp, nmax, x, y, comps, stride, gI, g, gamp, gx, gy, keep, n_nan);
  return shp_point_ct<T, GL_SHP_MAXN>(d, table, interp, nmax, x, y, comps, stride, gI, g, gamp, gx, gy, keep, n_nan);
}

// ---------------------------------------------------------------------------------------------
// generic per-type tables
// ---------------------------------------------------------------------------------------------
GL_HD int gl_n_raw(int type) {
  switch (type) {
    case GLT_EPL: return 6; case GLT_SHEAR: return 2; case GLT_SIE: return 5; case GLT_SIS: return 3;
    case GLT_NFW: return 4; case GLT_NFW_ELLIPSE: return 6; case GLT_DPIS: return 5; case GLT_DPIE: return 7;
    case GLT_SERSIC: return 5; case GLT_SERSIC_ELLIPSE: return 7; case GLT_SHAPELETS: return 3;
    case GLT_TNFW: return 5; case GLT_DPIEP: return 7; case GLT_CORE_SERSIC: return 10;
  }
  return 0;
}
GL_HD int gl_n_dvars(int type) {
  switch (type) {
    case GLT_EPL: return 8; case GLT_SHEAR: return 2; case GLT_SIE: return 6; case GLT_SIS: return 3;
    case GLT_NFW: case GLT_NFW_ELLIPSE: return 7; case GLT_DPIS: case GLT_DPIE: return 7;
    case GLT_SERSIC: case GLT_SERSIC_ELLIPSE: return 8; case GLT_SHAPELETS: return 3;   // + n_layers amplitudes without use_lstsq
    case GLT_TNFW: return 5; case GLT_DPIEP: return 8; case GLT_CORE_SERSIC: return 9;
  }
  return 0;
}
GL_HD int gl_der_size(int type, int niter, int nmax) {
  switch (type) {
    case GLT_EPL: return epl_der_size(niter); case GLT_SHEAR: return 4; case GLT_SIE: return SIE_SIZE; case GLT_SIS: return 4;
    case GLT_NFW: case GLT_NFW_ELLIPSE: return NFW_SIZE; case GLT_DPIS: case GLT_DPIE: return DP_SIZE;
    case GLT_SERSIC: case GLT_SERSIC_ELLIPSE: return SER_SIZE; case GLT_SHAPELETS: return shp_der_size(nmax);
    case GLT_TNFW: return TNFW_SIZE; case GLT_DPIEP: return PP_SIZE; case GLT_CORE_SERSIC: return CS_SIZE;
  }
  return 0;
}

// raw -> derived for the simple (non-scaled, non-shapelets) types
template <class T>
GL_HD void gl_prep(int type, unsigned flags, int niter, const T* raw, T* d, T epl_fmax, T epl_tol) {
  switch (type) {
    case GLT_EPL: epl_prep(raw, d, niter, epl_fmax, epl_tol); break;
    case GLT_SHEAR: d[0] = raw[0]; d[1] = raw[1]; d[2] = T(0); d[3] = T(0); break;
    case GLT_SIE: sie_prep(raw, d); break;
    case GLT_SIS: d[0] = raw[0]; d[1] = raw[1]; d[2] = raw[2]; d[3] = T(0); break;
    case GLT_NFW: nfw_prep(raw, d, false); break;
    case GLT_NFW_ELLIPSE: nfw_prep(raw, d, true); break;
    case GLT_DPIS: dpie_prep(raw, d, false); break;
    case GLT_DPIE: dpie_prep(raw, d, true); break;
    case GLT_TNFW: tnfw_prep(raw, d); break;
    case GLT_DPIEP: dpiep_prep(raw, d); break;
    case GLT_SERSIC: sersic_prep(raw, d, false, (flags & 1u) != 0); break;
    case GLT_SERSIC_ELLIPSE: sersic_prep(raw, d, true, (flags & 1u) != 0); break;
    case GLT_CORE_SERSIC: core_sersic_prep(raw, d, (flags & 1u) != 0); break;
    default: break;
  }
}
template <class T>
GL_HD void gl_prep_bwd(int type, unsigned flags, const T* raw, const T* d, const T* g, T* graw) {
  switch (type) {
    case GLT_EPL: epl_prep_bwd(raw, d, g, graw); break;
    case GLT_SHEAR: graw[0] = g[0]; graw[1] = g[1]; break;
    case GLT_SIE: sie_prep_bwd(raw, d, g, graw); break;
    case GLT_SIS: graw[0] = g[0]; graw[1] = g[1]; graw[2] = g[2]; break;
    case GLT_NFW: nfw_prep_bwd(raw, d, g, graw, false); break;
    case GLT_NFW_ELLIPSE: nfw_prep_bwd(raw, d, g, graw, true); break;
    case GLT_DPIS: dpie_prep_bwd(raw, d, g, graw, false); break;
    case GLT_DPIE: dpie_prep_bwd(raw, d, g, graw, true); break;
    case GLT_TNFW: tnfw_prep_bwd(raw, d, g, graw); break;
    case GLT_DPIEP: dpiep_prep_bwd(raw, d, g, graw); break;
    case GLT_SERSIC: sersic_prep_bwd(raw, d, g, graw, false, (flags & 1u) != 0); break;
    case GLT_SERSIC_ELLIPSE: sersic_prep_bwd(raw, d, g, graw, true, (flags & 1u) != 0); break;
    case GLT_CORE_SERSIC: core_sersic_prep_bwd(raw, d, g, graw, (flags & 1u) != 0); break;
    default: break;
  }
}

// deflection of one lens entry at NP points (d = derived block of the entry)
template <class T, int NP, unsigned F>
GL_HD void gl_lens_fwd(int type, int ts, const typename gl_scalar_of<T>::type* d, const T* x, const T* y, T* ax, T* ay,
                       T* scr = nullptr, int scr_stride = 0) {
#pragma unroll
  for (int j = 0; j < NP; ++j) { ax[j] = T(0); ay[j] = T(0); }
  switch (type) {
    case GLT_EPL: if constexpr ((F & GLF_EPL) != 0) {
      if (scr) epl_fwd_save<T, NP>(d, ts, x, y, ax, ay, scr, scr_stride); else epl_fwd<T, NP>(d, ts, x, y, ax, ay);
    } break;
    case GLT_SHEAR: if constexpr ((F & GLF_SHEAR) != 0) shear_fwd<T, NP>(d, x, y, ax, ay); break;
    case GLT_SIE: if constexpr ((F & GLF_SIE) != 0) sie_fwd<T, NP>(d, x, y, ax, ay); break;
    case GLT_SIS: if constexpr ((F & GLF_SIS) != 0) sis_fwd<T, NP>(d, x, y, ax, ay); break;
    case GLT_NFW: case GLT_NFW_ELLIPSE: if constexpr ((F & GLF_NFW) != 0) NfwLane<T, NP>::fwd(d, x, y, ax, ay); break;
    case GLT_DPIS: if constexpr ((F & GLF_DPIS) != 0) dpis_fwd<T, NP>(d, x, y, ax, ay); break;
    case GLT_DPIE: if constexpr ((F & GLF_DPIE) != 0) dpie_fwd<T, NP>(d, x, y, ax, ay); break;
    case GLT_TNFW: if constexpr ((F & GLF_TNFW) != 0) tnfw_fwd<T, NP>(d, x, y, ax, ay); break;
    case GLT_DPIEP: if constexpr ((F & GLF_DPIEP) != 0) dpiep_fwd<T, NP>(d, x, y, ax, ay); break;
    default: break;
  }
}
template <class T, int NP, unsigned F>
GL_HD void gl_lens_bwd(int type, int ts, const typename gl_scalar_of<T>::type* d, const T* x, const T* y, const T* gax, const T* gay, T* g,
                       const T* scr = nullptr, int scr_stride = 0) {
  switch (type) {
    case GLT_EPL: if constexpr ((F & GLF_EPL) != 0) {
      if (scr) epl_bwd_load<T, NP>(d, ts, x, y, gax, gay, g, scr, scr_stride); else epl_bwd<T, NP>(d, ts, x, y, gax, gay, g);
    } break;
    case GLT_SHEAR: if constexpr ((F & GLF_SHEAR) != 0) shear_bwd<T, NP>(d, x, y, gax, gay, g); break;
    case GLT_SIE: if constexpr ((F & GLF_SIE) != 0) sie_bwd<T, NP>(d, x, y, gax, gay, g); break;
    case GLT_SIS: if constexpr ((F & GLF_SIS) != 0) sis_bwd<T, NP>(d, x, y, gax, gay, g); break;
    case GLT_NFW: case GLT_NFW_ELLIPSE: if constexpr ((F & GLF_NFW) != 0) NfwLane<T, NP>::bwd(d, x, y, gax, gay, g); break;
    case GLT_DPIS: if constexpr ((F & GLF_DPIS) != 0) dpis_bwd<T, NP>(d, x, y, gax, gay, g); break;
    case GLT_DPIE: if constexpr ((F & GLF_DPIE) != 0) dpie_bwd<T, NP>(d, x, y, gax, gay, g); break;
    case GLT_TNFW: if constexpr ((F & GLF_TNFW) != 0) tnfw_bwd<T, NP>(d, x, y, gax, gay, g); break;
    case GLT_DPIEP: if constexpr ((F & GLF_DPIEP) != 0) dpiep_bwd<T, NP>(d, x, y, gax, gay, g); break;
    default: break;
  }
}
