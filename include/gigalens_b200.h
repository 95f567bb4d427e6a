/*
 * gigalens_b200 -- C ABI of the B200-native forward model (libgigalens_b200.so).
 *
 * The reference (furcelay/gigalens) is pure Python on TensorFlow/JAX and has no FFI of its
 * own; the entry points below are what a binding for its hot path would bind.  Each one
 * cites the reference interface it replaces (paths relative to /root/reference).
 * INTEGRATION.md shows the reference-side ctypes stub.
 *
 * Conventions
 *   - Plain C: opaque plan handle, plain pointers and sizes, `int` status (0 = ok), message via
 *     gl_last_error() (thread-local).  No exceptions cross the ABI, no torch types.
 *   - "dev" pointers are CUDA device pointers owned by the caller; "host" pointers are ordinary
 *     host memory.  The plan owns only its workspace and its copies of the static inputs
 *     (grid, mask, PSF, observation, catalogue).  Memory is allocated only by the setup calls
 *     (gl_plan_create, gl_plan_set_likelihood / _prior / _positions, gl_plan_reserve_lstsq); the data-path
 *     entry points (simulate, log-likelihood, log-prob, lstsq) allocate nothing and set no function attributes.
 *   - Kernels are launched asynchronously on the caller's `stream` (a cudaStream_t passed as
 *     void*; NULL = legacy default stream).  No entry point synchronises except the *_host ones.
 *   - A plan is bound to one device and one batch size and is not thread-safe.
 *   - Parameters cross as SoA fp32 `params[P][bs]` (row = one free parameter, the natural layout
 *     after the reference's tfb.Split, src/gigalens/tf/model.py:78-85).  All arithmetic is fp32
 *     like the reference (tf.float32 throughout, e.g. src/gigalens/tf/simulator.py:27-32).
 */
#ifndef GIGALENS_B200_H
#define GIGALENS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GL_ABI_VERSION 2           /* 2: GL_MAX_PROFILE_PARAMS 8 -> 10 (gl_profile_desc layout), GL_CORE_SERSIC, gl_eval_points mode 3 */
#define GL_MAX_PROFILE_PARAMS 10

/* Profile type ids.  The raw-parameter order of each type is fixed here and mirrors the
 * keyword names of the reference's `deriv` / `light`. */
typedef enum {
  GL_EPL = 1,          /* theta_E, gamma, e1, e2, center_x, center_y      tf/profiles/mass/epl.py:12-57 */
  GL_SHEAR = 2,        /* gamma1, gamma2                                  tf/profiles/mass/shear.py:8-16 */
  GL_SIE = 3,          /* theta_E, e1, e2, center_x, center_y             tf/profiles/mass/sie.py:6-42 */
  GL_SIS = 4,          /* theta_E, center_x, center_y                     tf/profiles/mass/sis.py:6-17 */
  GL_NFW = 5,          /* Rs, alpha_Rs, center_x, center_y                tf/profiles/mass/nfw.py:7-52 */
  GL_NFW_ELLIPSE = 6,  /* Rs, alpha_Rs, e1, e2, center_x, center_y        tf/profiles/mass/nfw.py:99-134 */
  GL_DPIS = 7,         /* theta_E, r_core, r_cut, center_x, center_y      tf/profiles/mass/piemd.py:26-60 */
  GL_DPIE = 8,         /* theta_E, r_core, r_cut, e1, e2, center_x, center_y  tf/profiles/mass/piemd.py:98-119,183-255 */
  GL_TNFW = 9,         /* Rs, alpha_Rs, r_trunc, center_x, center_y       tf/profiles/mass/tnfw.py:10-62 */
  GL_DPIEP = 10,       /* theta_E, Ra, Rs, e1, e2, center_x, center_y     tf/profiles/mass/piep.py:17-56 */
  GL_SERSIC = 32,          /* R_sersic, n_sersic, center_x, center_y, Ie          tf/profiles/light/sersic.py:22-35 */
  GL_SERSIC_ELLIPSE = 33,  /* R_sersic, n_sersic, e1, e2, center_x, center_y, Ie  tf/profiles/light/sersic.py:67-80 */
  GL_SHAPELETS = 34,       /* beta, center_x, center_y (+ amplitudes)              tf/profiles/light/shapelets.py:17-75 */
  GL_CORE_SERSIC = 35      /* R_sersic, n_sersic, Rb, alpha, gamma, e1, e2, center_x, center_y, Ie   tf/profiles/light/sersic.py:83-132
                            * (the formula exactly as written there, see gl_math.cuh) */
} gl_profile_type;

#define GL_FLAG_USE_LSTSQ 1u   /* LightProfile.use_lstsq  (src/gigalens/profile.py:36-41) */
#define GL_FLAG_INTERPOLATE 2u /* Shapelets(interpolate=True)  (shapelets.py:20,54-66) */

/* One profile of the physical model.
 *
 * Raw parameter k of sample b is  base_k(b) * member_factor[k*n_members + g]  where
 * base_k(b) = params[slot[k]][b] if slot[k] >= 0 else constant[k], and the factor is 1 when
 * n_members == 0.  n_members > 0 makes this entry a scaling-relation sum over a galaxy
 * catalogue (tf/profiles/mass/scaling_relation.py:57-70): scaled parameters carry
 * (L_g/L*)^power_k in member_factor, catalogue columns carry constant[k]=1 and the catalogue
 * value in member_factor, and the deflection is summed over the n_members members. */
typedef struct {
  int32_t type;                           /* gl_profile_type */
  uint32_t flags;                         /* GL_FLAG_* */
  int32_t slot[GL_MAX_PROFILE_PARAMS];    /* row of params[P][bs], or -1 = constant */
  float constant[GL_MAX_PROFILE_PARAMS];  /* used when slot[k] < 0 */
  int32_t niter;                          /* GL_EPL: iteration cap (EPL(niter=50), epl.py:15) */
  int32_t n_max;                          /* GL_SHAPELETS: maximum order */
  const int32_t* amp_slot;                /* GL_SHAPELETS without use_lstsq: [n_layers] rows of params, host */
  int32_t n_members;                      /* 0, or G for a scaling-relation sum */
  const float* member_factor;             /* host, [n_raw_params(type)][n_members], NULL if n_members == 0 */
} gl_profile_desc;

/* PhysicalModel(lenses, lens_light, source_light, *_constants): src/gigalens/tf/model.py:290-306 */
typedef struct {
  int32_t n_lens, n_lens_light, n_source_light;
  const gl_profile_desc* lens;
  const gl_profile_desc* lens_light;
  const gl_profile_desc* source_light;
  int32_t n_params;                       /* P: rows of params[P][bs] */
} gl_model_desc;

/* SimulatorConfig + the host-side setup of LensSimulator.__init__:
 * src/gigalens/simulator.py:11-29, src/gigalens/tf/simulator.py:14-70 */
typedef struct {
  int32_t num_pix;          /* n: the image is n x n */
  int32_t supersample;      /* ss: ray-shooting grid is (n*ss) x (n*ss) */
  const float* grid_x;      /* host [(n*ss)^2] row-major: LensWCS.pix2angle x of every ss pixel (tf/simulator.py:45) */
  const float* grid_y;      /* host [(n*ss)^2] */
  const float* psf;         /* host [psf_n*psf_n]: the kernel exactly as handed to tf.nn.conv2d, i.e.
                               subgrid_kernel(kernel, ss, odd=True)[::-1, ::-1] (tf/simulator.py:62-70); NULL = no PSF */
  int32_t psf_n;
  const uint8_t* mask;      /* host [n*n] pix_region (non-zero = used), NULL = all pixels (tf/simulator.py:34-44) */
  float conversion_factor;  /* det(transform_pix2angle), tf/simulator.py:27-29 */
} gl_sim_config;

/* Pixel likelihood of ForwardProbModel.stats_pixels (src/gigalens/tf/model.py:89-101) or, with
 * fixed_error_map, the Independent(Normal(obs, err)) of BackwardProbModel (:221-226,266-267). */
typedef struct {
  const float* observed;    /* host [n*n] */
  const float* error_map;   /* host [n*n] or NULL => sqrt(background_rms^2 + im_sim/exp_time) */
  float background_rms;
  float exp_time;
} gl_like_config;

/* Prior + default event-space bijector of one flattened leaf (the reference
 * delegates these to TFP: tf/model.py:76-87,148,164-166). */
typedef enum { GL_DIST_NORMAL = 0, GL_DIST_LOGNORMAL = 1, GL_DIST_UNIFORM = 2, GL_DIST_TRUNCNORMAL = 3 } gl_dist_type;
typedef struct {
  int32_t dist;        /* gl_dist_type */
  int32_t slot;        /* row of params[P][bs] this leaf feeds (simulator slot order) */
  float a, b;          /* Normal/LogNormal/TruncNormal: loc, scale.  Uniform: low, high */
  float low, high;     /* TruncNormal bounds */
} gl_prior_leaf;

typedef struct gl_plan gl_plan;

/* --- lifecycle ------------------------------------------------------------------------- */
/* LensSimulator(phys_model, sim_config, bs): tf/simulator.py:14-70 */
int gl_plan_create(const gl_model_desc* model, const gl_sim_config* sim, int32_t bs, int32_t device, gl_plan** out);
/* ForwardProbModel(prior, observed_image, background_rms, exp_time, error_map): tf/model.py:32-66 */
int gl_plan_set_likelihood(gl_plan* plan, const gl_like_config* like);
/* prior / bijector leaves in z-column order (tf/model.py:76-87); d = n_leaves */
int gl_plan_set_prior(gl_plan* plan, const gl_prior_leaf* leaves, int32_t n_leaves);
/* ForwardProbModel(centroids_x, centroids_y, centroids_errors_x, centroids_errors_y): tf/model.py:69-74.
 * One entry of n_images per multiply-imaged source; x, y, err_x, err_y are HOST arrays with the images of
 * all systems concatenated (1..64 images per system).  Switches "include_positions" on. */
int gl_plan_set_positions(gl_plan* plan, int32_t n_systems, const int32_t* n_images, const float* x, const float* y,
                          const float* err_x, const float* err_y);
/* Options.  "epl_batch_max" = 1: EPL series length from the batch maximum of f exactly like
 * tf/profiles/mass/epl.py:37 (default 0: per-sample length, identical to fp32 rounding).
 * "epl_tol_exp10" = k: the EPL series stops at terms below 10^-k.  The reference's constant is 1e-12
 * (epl.py:37, k = 12) although its sums are fp32; the default here is k = 9: a dropped tail is below
 * 2e-9 of the O(1) sum (1/30 ulp), and the terms the reference adds beyond that are absorbed by its
 * own fp32 additions.  k = 12 with "epl_batch_max" = 1 reproduces the reference's trip count exactly.
 * "row_flush" = 0: the packed adjoint kernels reduce dvar cotangents with a warp butterfly per profile
 * instead of the staged shared-memory transposition (A/B measurement aid; results agree to fp32
 * summation order).  "conv_tma" = 0: the conv kernels stage their tiles with cp.async / plain stores
 * instead of TMA (A/B; bit-identical results).  "straight_line" = 0: programs of the benchmark shape
 * (lenses [EPL, Shear], one Sersic(Ellipse) lens light, one Sersic(Ellipse) source) run the interpreting
 * pixel drivers like every other program instead of the straight-line ones (A/B; bit-identical results).
 * "no_deflection" = 1: evaluate source light at the image-plane position (simulate(..., no_deflection=True),
 * tf/simulator.py:125-126).  "components" = 1 | 2 | 3: gl_simulate adds only the lens light / only the
 * source light / both (simulate_lens_light, simulate_images, simulate_source: tf/simulator.py:242-328).
 * "lstsq" = 1: the log-likelihood entry points use the linear-amplitude solve
 * (BackwardProbModel, tf/model.py:242-273).
 * "include_pixels" / "include_positions" = 0 | 1: which terms gl_loglike_grad / gl_logprob_grad add
 * (ForwardProbModel(include_pixels, include_positions), tf/model.py:43-44,150-163): log-likes add and the
 * reduced chi^2 is the mean of the included terms.
 * "lstsq_hide_tail" = 0: the gradient path of the lstsq model runs the eigen-solve of the samples with a singular
 * Gram matrix in line instead of on its side stream (A/B; bit-identical results). */
int gl_plan_set_option(gl_plan* plan, const char* name, int32_t value);
/* Measurement aid: after gl_plan_set_option(plan, "timing", n) every log-likelihood call records CUDA
 * events around its kernels on the launch stream; this returns the summed device time (ms) of the 7
 * stages (unconstrain, prep, raytrace_fwd, conv_fwd, conv_bwd, raytrace_bwd, sample_bwd) over the last
 * <= n calls and resets the counter.  Synchronises the device. */
int gl_plan_get_timings(gl_plan* plan, float* ms_out, int32_t* ncalls_out);
void gl_plan_destroy(gl_plan* plan);
const char* gl_last_error(void);
int32_t gl_abi_version(void);
/* Debug aid (no reference counterpart): with GL_GUARD=1 in the environment every device allocation of the library carries
 * 64 KB red zones filled with 0xFF; this verifies them (returns the number of live guarded allocations, -1 + gl_last_error()
 * if a kernel wrote into one, 0 when guarding is off).  Synchronises the device. */
int32_t gl_guard_check(void);
/* Positive control: deliberately overruns a guarded scratch buffer by one float on either side; 0 = both detected. */
int32_t gl_guard_selftest(void);
/* number of kernels the library has launched in this process (bench.py's gpu_launches) */
int64_t gl_launch_count(void);

/* --- simulate -------------------------------------------------------------------------- */
/* LensSimulator.simulate(params): tf/simulator.py:109-156.  image_dev [bs][n][n]. */
int gl_simulate(gl_plan* plan, const float* params_dev, float* image_dev, void* stream);
/* The supersampled pre-convolution image after the NaN scrub (tf/simulator.py:124-140),
 * ss_dev [bs][n*ss][n*ss].  Diagnostic / parity entry point. */
int gl_simulate_ss(gl_plan* plan, const float* params_dev, float* ss_dev, void* stream);
/* LensSimulator.beta(x, y, lens_params) at arbitrary points: tf/simulator.py:72-78.
 * x,y [npts] shared by all samples; beta_* [bs][npts]. */
int gl_beta(gl_plan* plan, const float* params_dev, int32_t npts, const float* x_dev, const float* y_dev,
            float* beta_x_dev, float* beta_y_dev, void* stream);

/* Total deflection (mode 1: out0 = alpha_x, out1 = alpha_y), beta (mode 0) or surface brightness
 * (mode 2: out0 only) at arbitrary points: MassProfile.deriv / LightProfile.light as the reference's
 * profile tests call them (tests/test_profiles.py:14-111).  out* [bs][npts].
 * Mode 3 (out0 only, [bs][depth][npts]): the unit-amplitude linear components that light() of a use_lstsq profile
 * returns (tf/profiles/light/sersic.py:31-35, shapelets.py:62-63,72-73), NaN-scrubbed like the lstsq stack. */
int gl_eval_points(gl_plan* plan, const float* params_dev, int32_t npts, const float* x_dev, const float* y_dev,
                   int32_t mode, float* out0_dev, float* out1_dev, void* stream);

/* --- likelihood and gradient ----------------------------------------------------------- */
/* ForwardProbModel.stats_pixels + d(log_like)/d(params): tf/model.py:89-101 with the gradient
 * tf.GradientTape would give (tf/inference.py:34-37).  loglike/red_chi2 [bs]; dparams [P][bs] or
 * NULL for forward only. */
int gl_loglike_grad(gl_plan* plan, const float* params_dev, float* loglike_dev, float* red_chi2_dev,
                    float* dparams_dev, void* stream);
/* ForwardProbModel.log_prob(simulator, z) and its gradient: tf/model.py:126-167.
 * z [bs][d] row-major (unconstrained); logp/red_chi2 [bs]; dz [bs][d] or NULL. */
int gl_logprob_grad(gl_plan* plan, const float* z_dev, float* logp_dev, float* red_chi2_dev, float* dz_dev,
                    void* stream);
/* bij.forward(z) -> params[P][bs] and log_prior(z) [bs] (either output may be NULL): tf/model.py:148,164-166,183-185 */
int gl_unconstrain(gl_plan* plan, const float* z_dev, float* params_dev, float* logprior_dev, void* stream);
/* Chain rule of the bijector alone: dz[bs][d] = (d params/d z)^T dparams_dev[P][bs] (+ d(log_prior + fldj)/dz when
 * with_prior != 0); dparams_dev may be NULL (prior gradient only); logprior_dev [bs] or NULL.  The pieces a
 * tempered target prior + aux + beta (like - aux) needs (tf/inference.py:289-302). */
int gl_chain_grad(gl_plan* plan, const float* z_dev, const float* dparams_dev, int32_t with_prior, float* logprior_dev,
                  float* dz_dev, void* stream);
/* One Adam update of ModellingSequence.MAP (tf/inference.py:34-37, optimizer.apply_gradients on the per-sample loss) in one launch,
 * element-wise on n floats (all device pointers; x, m, v updated in place): g = grad * grad_scale (non-finite -> 0 when
 * scrub_nan), m = b1 m + (1 - b1) g, v = b2 v + (1 - b2) g^2, x -= alpha m / (sqrt(v) + eps); the caller folds the bias
 * correction into alpha = lr sqrt(1 - b2^t) / (1 - b1^t) (Keras' Adam).  The scalars are doubles so that 1 - beta is formed before
 * the float32 cast (1.f - 0.999f is off by 1.3e-5). */
int gl_adam_step(float* x_dev, const float* grad_dev, float* m_dev, float* v_dev, int64_t n, double grad_scale, double beta1, double beta2,
                 double alpha, double eps, int32_t scrub_nan, void* stream);

/* Host-buffer convenience wrappers (pinned or pageable host memory): copy in, run, copy out,
 * synchronise.  These are the calls a reference-side binding would make per optimiser step. */
int gl_logprob_grad_host(gl_plan* plan, const float* z_host, float* logp_host, float* red_chi2_host, float* dz_host);
int gl_simulate_host(gl_plan* plan, const float* params_host, float* image_host);

/* --- lensing Hessian and image-position likelihood (FP64 forward-mode duals) ------------- */
/* Hessian of the summed deflection, (f_xx, f_xy, f_yx, f_yy) = (d ax/dx, d ax/dy, d ay/dx, d ay/dy), at npts
 * points shared by all samples: MassProfile.hessian (tf/profile.py:9-30 and the analytic sis.py:19-29,
 * shear.py:18-26, nfw.py:78-94, piemd.py:62-83,121-138), summed as LensSimulator.magnification /
 * convergence / shear do (tf/simulator.py:80-107).  out [bs][npts] each. */
int gl_hessian(gl_plan* plan, const float* params_dev, int32_t npts, const float* x_dev, const float* y_dev,
               float* fxx_dev, float* fxy_dev, float* fyx_dev, float* fyy_dev, void* stream);
/* ForwardProbModel.stats_positions(simulator, params): tf/model.py:103-124.  loglike_dev / red_chi2_dev [bs];
 * dparams_dev [P][bs] = d(log-like)/d(params), or NULL. */
int gl_positions_loglike_grad(gl_plan* plan, const float* params_dev, float* loglike_dev, float* red_chi2_dev,
                              float* dparams_dev, void* stream);

/* --- linear light-amplitude solve ------------------------------------------------------ */
/* LensSimulator.lstsq_simulate(params, observed_image, err_map, return_coeffs): tf/simulator.py:158-240
 * (layout jax/simulator.py:171-195).  image_dev [bs][n][n] or NULL; coeffs_dev [bs][D] or NULL. */
int gl_lstsq_simulate(gl_plan* plan, const float* params_dev, float* image_dev, float* coeffs_dev, void* stream);
/* lstsq_simulate(..., return_stacked=True): the D convolved, down-sampled unit-amplitude light components
 * (tf/simulator.py:203-229), stack_dev [bs][D][n][n] (the reference's layout is [bs][n][n][D]). */
int gl_lstsq_stack(gl_plan* plan, const float* params_dev, float* stack_dev, void* stream);
/* BackwardProbModel.log_prob likelihood part + gradient w.r.t. the non-linear params: tf/model.py:242-273 */
int gl_lstsq_loglike_grad(gl_plan* plan, const float* params_dev, float* loglike_dev, float* red_chi2_dev,
                          float* dparams_dev, void* stream);
int32_t gl_plan_depth(const gl_plan* plan); /* D: number of linear light components */
/* Workspace of the lstsq entry points: the component stack [chunk][D][(n*ss)^2] and its convolved copy; chunk = samples
 * per pass, 0 = sized by a memory budget (at most 32 GB / 40 % of the free memory).  gl_plan_create reserves it for models
 * with a GL_FLAG_USE_LSTSQ light profile; any other model that wants lstsq_simulate (the reference treats every light
 * profile as a linear component there, tf/simulator.py:183-200) calls this once first.  Re-reserving synchronises. */
int gl_plan_reserve_lstsq(gl_plan* plan, int32_t chunk);

/* --- measurement aid --------------------------------------------------------------------- */
/* Measured FP32 FMA peak of `device` in TFLOP/s (FMA = 2): independent register chains, no memory; scalar FFMA and
 * packed FFMA2 variants, best of 5 launches each, CUDA-event timed.  bench.py's roofline denominator. */
int gl_fp32_peak(int32_t device, float* tflops_ffma, float* tflops_ffma2);

#ifdef __cplusplus
}
#endif
#endif /* GIGALENS_B200_H */
