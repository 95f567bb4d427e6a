"""ctypes binding of ``libgigalens_b200.so`` (declared in ``include/gigalens_b200.h``).

The library is the product: there is no Python or CPU fallback.  ``load()`` raises if the
shared object is missing, and every compute entry point fails inside the library when no
CUDA device is present.
"""
import ctypes as C
import os

GL_MAX_PROFILE_PARAMS = 10
GL_ABI_VERSION = 2

# gl_profile_type
GL_EPL, GL_SHEAR, GL_SIE, GL_SIS, GL_NFW, GL_NFW_ELLIPSE, GL_DPIS, GL_DPIE = 1, 2, 3, 4, 5, 6, 7, 8
GL_TNFW, GL_DPIEP = 9, 10
GL_SERSIC, GL_SERSIC_ELLIPSE, GL_SHAPELETS, GL_CORE_SERSIC = 32, 33, 34, 35
GL_FLAG_USE_LSTSQ, GL_FLAG_INTERPOLATE = 1, 2
GL_DIST_NORMAL, GL_DIST_LOGNORMAL, GL_DIST_UNIFORM, GL_DIST_TRUNCNORMAL = 0, 1, 2, 3

# raw-parameter order per type (must match the enum comments in the header)
RAW_ORDER = {
    GL_EPL: ["theta_E", "gamma", "e1", "e2", "center_x", "center_y"],
    GL_SHEAR: ["gamma1", "gamma2"],
    GL_SIE: ["theta_E", "e1", "e2", "center_x", "center_y"],
    GL_SIS: ["theta_E", "center_x", "center_y"],
    GL_NFW: ["Rs", "alpha_Rs", "center_x", "center_y"],
    GL_NFW_ELLIPSE: ["Rs", "alpha_Rs", "e1", "e2", "center_x", "center_y"],
    GL_DPIS: ["theta_E", "r_core", "r_cut", "center_x", "center_y"],
    GL_DPIE: ["theta_E", "r_core", "r_cut", "e1", "e2", "center_x", "center_y"],
    GL_TNFW: ["Rs", "alpha_Rs", "r_trunc", "center_x", "center_y"],
    GL_DPIEP: ["theta_E", "Ra", "Rs", "e1", "e2", "center_x", "center_y"],
    GL_SERSIC: ["R_sersic", "n_sersic", "center_x", "center_y", "Ie"],
    GL_SERSIC_ELLIPSE: ["R_sersic", "n_sersic", "e1", "e2", "center_x", "center_y", "Ie"],
    GL_SHAPELETS: ["beta", "center_x", "center_y"],
    GL_CORE_SERSIC: ["R_sersic", "n_sersic", "Rb", "alpha", "gamma", "e1", "e2", "center_x", "center_y", "Ie"],
}


class ProfileDesc(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("flags", C.c_uint32),
        ("slot", C.c_int32 * GL_MAX_PROFILE_PARAMS),
        ("constant", C.c_float * GL_MAX_PROFILE_PARAMS),
        ("niter", C.c_int32),
        ("n_max", C.c_int32),
        ("amp_slot", C.POINTER(C.c_int32)),
        ("n_members", C.c_int32),
        ("member_factor", C.POINTER(C.c_float)),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ("n_lens", C.c_int32),
        ("n_lens_light", C.c_int32),
        ("n_source_light", C.c_int32),
        ("lens", C.POINTER(ProfileDesc)),
        ("lens_light", C.POINTER(ProfileDesc)),
        ("source_light", C.POINTER(ProfileDesc)),
        ("n_params", C.c_int32),
    ]


class SimConfig(C.Structure):
    _fields_ = [
        ("num_pix", C.c_int32),
        ("supersample", C.c_int32),
        ("grid_x", C.POINTER(C.c_float)),
        ("grid_y", C.POINTER(C.c_float)),
        ("psf", C.POINTER(C.c_float)),
        ("psf_n", C.c_int32),
        ("mask", C.POINTER(C.c_uint8)),
        ("conversion_factor", C.c_float),
    ]


class LikeConfig(C.Structure):
    _fields_ = [
        ("observed", C.POINTER(C.c_float)),
        ("error_map", C.POINTER(C.c_float)),
        ("background_rms", C.c_float),
        ("exp_time", C.c_float),
    ]


class PriorLeaf(C.Structure):
    _fields_ = [
        ("dist", C.c_int32),
        ("slot", C.c_int32),
        ("a", C.c_float),
        ("b", C.c_float),
        ("low", C.c_float),
        ("high", C.c_float),
    ]


# every symbol include/gigalens_b200.h declares (tests check they are all exported)
EXPORTED_SYMBOLS = [
    "gl_plan_create", "gl_plan_set_likelihood", "gl_plan_set_prior", "gl_plan_destroy", "gl_last_error",
    "gl_abi_version", "gl_guard_check", "gl_guard_selftest", "gl_launch_count", "gl_simulate", "gl_simulate_ss", "gl_beta", "gl_eval_points",
    "gl_loglike_grad", "gl_logprob_grad", "gl_unconstrain", "gl_logprob_grad_host", "gl_simulate_host",
    "gl_lstsq_simulate", "gl_lstsq_loglike_grad", "gl_plan_depth", "gl_plan_set_option", "gl_plan_get_timings",
    "gl_plan_set_positions", "gl_hessian", "gl_positions_loglike_grad", "gl_lstsq_stack", "gl_chain_grad", "gl_adam_step",
    "gl_plan_reserve_lstsq", "gl_fp32_peak",
]

_LIB = None


def library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgigalens_b200.so")


def load():
    """Load the CUDA library; raise loudly when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: gigalens_b200 has no CPU fallback. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc)."
        )
    lib = C.CDLL(path)
    vp, fp, i32 = C.c_void_p, C.c_void_p, C.c_int32
    lib.gl_last_error.restype = C.c_char_p
    lib.gl_abi_version.restype = C.c_int32
    if lib.gl_abi_version() != GL_ABI_VERSION:
        raise RuntimeError(f"{path} has ABI version {lib.gl_abi_version()}, this binding needs {GL_ABI_VERSION}: rebuild it "
                           "(`python -c 'import __graft_entry__ as g; g.build(force=True)'`)")
    lib.gl_guard_check.restype = C.c_int32
    lib.gl_guard_selftest.restype = C.c_int32
    lib.gl_launch_count.restype = C.c_int64
    lib.gl_plan_depth.restype = C.c_int32
    lib.gl_plan_depth.argtypes = [vp]
    lib.gl_plan_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(SimConfig), i32, i32, C.POINTER(vp)]
    lib.gl_plan_set_likelihood.argtypes = [vp, C.POINTER(LikeConfig)]
    lib.gl_plan_set_prior.argtypes = [vp, C.POINTER(PriorLeaf), i32]
    lib.gl_plan_set_option.argtypes = [vp, C.c_char_p, i32]
    lib.gl_plan_get_timings.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    lib.gl_plan_destroy.argtypes = [vp]
    lib.gl_plan_destroy.restype = None
    lib.gl_simulate.argtypes = [vp, fp, fp, vp]
    lib.gl_simulate_ss.argtypes = [vp, fp, fp, vp]
    lib.gl_beta.argtypes = [vp, fp, i32, fp, fp, fp, fp, vp]
    lib.gl_eval_points.argtypes = [vp, fp, i32, fp, fp, i32, fp, fp, vp]
    lib.gl_loglike_grad.argtypes = [vp, fp, fp, fp, fp, vp]
    lib.gl_logprob_grad.argtypes = [vp, fp, fp, fp, fp, vp]
    lib.gl_unconstrain.argtypes = [vp, fp, fp, fp, vp]
    lib.gl_logprob_grad_host.argtypes = [vp, fp, fp, fp, fp]
    lib.gl_simulate_host.argtypes = [vp, fp, fp]
    lib.gl_lstsq_simulate.argtypes = [vp, fp, fp, fp, vp]
    lib.gl_lstsq_loglike_grad.argtypes = [vp, fp, fp, fp, fp, vp]
    lib.gl_chain_grad.argtypes = [vp, fp, fp, i32, fp, fp, vp]
    lib.gl_adam_step.argtypes = [fp, fp, fp, fp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, i32, vp]
    lib.gl_lstsq_stack.argtypes = [vp, fp, fp, vp]
    lib.gl_plan_set_positions.argtypes = [vp, i32, fp, fp, fp, fp, fp]
    lib.gl_hessian.argtypes = [vp, fp, i32, fp, fp, fp, fp, fp, fp, vp]
    lib.gl_positions_loglike_grad.argtypes = [vp, fp, fp, fp, fp, vp]
    lib.gl_plan_reserve_lstsq.argtypes = [vp, i32]
    lib.gl_fp32_peak.argtypes = [i32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    for name in EXPORTED_SYMBOLS:
        if getattr(lib, name).restype is C.c_int:
            getattr(lib, name).restype = C.c_int
    _LIB = lib
    return lib


def check(status, lib=None):
    if status != 0:
        lib = lib or load()
        raise RuntimeError("gigalens_b200: " + lib.gl_last_error().decode())
