"""Shared test helpers: build the same physical model for the product (descriptor objects) and
for the oracle (torch restatement), draw seeded parameters, and drive the host harness."""
import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np
import torch

from gigalens_b200 import _cabi
from gigalens_b200.model import PhysicalModel
from gigalens_b200.profiles.light import sersic as gl_sersic, shapelets as gl_shapelets
from gigalens_b200.profiles.mass import (dpie_subhalo as gl_sub, epl as gl_epl, nfw as gl_nfw, piemd as gl_piemd,
                                         scaling_relation as gl_sr, shear as gl_shear, sie as gl_sie, sis as gl_sis,
                                         tnfw as gl_tnfw, piep as gl_piep)
from gigalens_b200.simulator import CompiledModel
from oracle import profiles as OP

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def to_oracle_profile(p, dtype=torch.float32):
    if isinstance(p, gl_sr.ScalingRelation):
        return OP.ScalingRelation(to_oracle_profile(p.profile, dtype), p.scaling_params, p.lum_star, p.power,
                                  p.galaxy_cat, dtype=dtype)
    table = {
        gl_epl.EPL: lambda: OP.EPL(p.niter), gl_shear.Shear: OP.Shear, gl_sie.SIE: OP.SIE, gl_sis.SIS: OP.SIS,
        gl_nfw.NFW: OP.NFW, gl_nfw.NFW_ELLIPSE: OP.NFW_ELLIPSE, gl_piemd.DPIS: OP.DPIS, gl_piemd.DPIE: OP.DPIE,
        gl_tnfw.TNFW: OP.TNFW, gl_piep.DPIEP: OP.DPIEP,
        gl_sersic.SersicEllipse: lambda: OP.SersicEllipse(p.use_lstsq), gl_sersic.Sersic: lambda: OP.Sersic(p.use_lstsq),
        gl_sersic.CoreSersic: lambda: OP.CoreSersic(p.use_lstsq),
        gl_shapelets.Shapelets: lambda: OP.Shapelets(p.n_max, p.use_lstsq, p.interpolate, dtype=dtype),
    }
    return table[type(p)]()


def to_oracle_model(pm, dtype=torch.float32):
    return SimpleNamespace(
        lenses=[to_oracle_profile(p, dtype) for p in pm.lenses],
        lens_light=[to_oracle_profile(p, dtype) for p in pm.lens_light],
        source_light=[to_oracle_profile(p, dtype) for p in pm.source_light],
        lenses_constants=pm.lenses_constants, lens_light_constants=pm.lens_light_constants,
        source_light_constants=pm.source_light_constants)


# ---------------------------------------------------------------- parameter draws
RANGES = {  # generic, well-conditioned draws per parameter name
    "theta_E": (0.8, 1.6), "gamma": (1.6, 2.4), "e1": (-0.25, 0.25), "e2": (-0.25, 0.25),
    "center_x": (-0.2, 0.2), "center_y": (-0.2, 0.2), "gamma1": (-0.08, 0.08), "gamma2": (-0.08, 0.08),
    "Rs": (0.5, 3.0), "alpha_Rs": (0.5, 2.0), "r_core": (0.02, 0.2), "r_cut": (1.0, 6.0),
    "r_trunc": (2.0, 8.0), "Ra": (0.02, 0.2),
    "R_sersic": (0.3, 1.2), "n_sersic": (0.8, 5.0), "Ie": (20.0, 300.0), "beta": (0.08, 0.2),
    "Rb": (0.05, 0.3), "alpha": (1.0, 3.0), "gamma": (1.6, 2.4),
}


def draw_matrix(cm: CompiledModel, bs, seed=0):
    """fp64 [P][bs] matrix of seeded parameter draws in slot order."""
    rng = np.random.default_rng(seed)
    mat = np.empty((max(1, cm.n_params), bs))
    for i, (_, _, name) in enumerate(cm.slot_keys):
        if name.startswith("amp"):
            mat[i] = rng.normal(0, 50, size=bs)
        else:
            lo, hi = RANGES[name]
            mat[i] = rng.uniform(lo, hi, size=bs)
    return mat


def matrix_to_pytree(cm: CompiledModel, mat, dtype, requires_grad=False):
    t = torch.as_tensor(np.asarray(mat), dtype=dtype)
    if requires_grad:
        t = t.clone().requires_grad_(True)
    return cm.unflatten(t), t


# ---------------------------------------------------------------- host harness
_HOST = None


def hostcheck_lib():
    """Build (g++) and load the TEST-ONLY host instantiation of gl_math.cuh / gl_program.h."""
    global _HOST
    if _HOST is not None:
        return _HOST
    src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
    out_dir = os.path.join(HERE, "hostcheck", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libglhostcheck.so")
    deps = [src] + [os.path.join(ROOT, "gigalens_b200", "csrc", f) for f in ("gl_math.cuh", "gl_program.h", "gl_build.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    lib = C.CDLL(so)
    lib.glh_last_error.restype = C.c_char_p
    _HOST = lib
    return lib


def host_run(cm: CompiledModel, mat, gx, gy, g_ss=None, dtype=np.float64, no_deflection=False, epl_batch_max=True,
             want_beta=False, want_comps=False):
    """Run prep -> pixel forward (-> pixel adjoint -> prep adjoint) on the host harness.
    Returns dict(ss=[bs][npix], gparams=[P][bs] or None, beta=[bs][2][npix] or None)."""
    lib = hostcheck_lib()
    fn = lib.glh_run_f64 if dtype == np.float64 else lib.glh_run_f32
    mat = np.ascontiguousarray(mat, dtype=dtype)
    P, bs = mat.shape
    gx = np.ascontiguousarray(gx, dtype=dtype)
    gy = np.ascontiguousarray(gy, dtype=dtype)
    npix = gx.size
    ss = np.zeros((bs, npix), dtype=dtype)
    gparams = np.zeros((P, bs), dtype=dtype) if g_ss is not None else None
    beta = np.zeros((bs, 2, npix), dtype=dtype) if want_beta else None
    comps = None
    if want_comps:
        depth = lib.glh_depth(C.byref(cm.desc))
        comps = np.zeros((bs, depth, npix), dtype=dtype)
    g_ss_c = np.ascontiguousarray(g_ss, dtype=dtype) if g_ss is not None else None
    vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = fn(C.byref(cm.desc), C.c_int(bs), vp(mat), C.c_int(npix), vp(gx), vp(gy), C.c_int(int(no_deflection)),
            C.c_int(int(epl_batch_max)), vp(ss), vp(g_ss_c), vp(gparams), vp(beta), vp(comps))
    if rc != 0:
        raise RuntimeError(lib.glh_last_error().decode())
    return dict(ss=ss, gparams=gparams, beta=beta, comps=comps)


def host_positions(cm: CompiledModel, mat, systems, dtype=np.float64, use_fwdmode=True, want_grad=True):
    """stats_positions on the host harness.  systems = [(x, y, err_x, err_y), ...] one tuple of (n_img,) arrays
    per multiply-imaged source.  Returns dict(beta=[bs][2][npts], hess=[bs][4][npts], loglike, chi2, gparams)."""
    lib = hostcheck_lib()
    fn = lib.glh_positions_f64 if dtype == np.float64 else lib.glh_positions_f32
    mat = np.ascontiguousarray(mat, dtype=dtype)
    P, bs = mat.shape
    n_img = np.asarray([len(s[0]) for s in systems], dtype=np.int32)
    cat = [np.ascontiguousarray(np.concatenate([np.asarray(s[k], dtype=dtype) for s in systems])) for k in range(4)]
    npts = int(n_img.sum())
    beta = np.zeros((bs, 2, npts), dtype=dtype)
    hess = np.zeros((bs, 4, npts), dtype=dtype)
    ll, chi2 = np.zeros(bs, dtype=dtype), np.zeros(bs, dtype=dtype)
    gparams = np.zeros((P, bs), dtype=dtype) if want_grad else None
    vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = fn(C.byref(cm.desc), C.c_int(bs), vp(mat), C.c_int(len(systems)), vp(n_img), vp(cat[0]), vp(cat[1]), vp(cat[2]), vp(cat[3]),
            C.c_int(int(use_fwdmode)), vp(beta), vp(hess), vp(ll), vp(chi2), vp(gparams))
    if rc != 0:
        raise RuntimeError(lib.glh_last_error().decode())
    return dict(beta=beta, hess=hess, loglike=ll, chi2=chi2, gparams=gparams)


def host_run_packed(cm: CompiledModel, mat, gx, gy, g_ss=None, straight_line=False):
    """Same as host_run(dtype=float32) but through the two-pixel packed lane type (GlF2); ``straight_line`` runs the
    benchmark-shape drivers (gl_pix_image_bs / gl_pix_image_bwd_bs) instead of the interpreter."""
    lib = hostcheck_lib()
    mat = np.ascontiguousarray(mat, dtype=np.float32)
    P, bs = mat.shape
    gx = np.ascontiguousarray(gx, dtype=np.float32)
    gy = np.ascontiguousarray(gy, dtype=np.float32)
    npix = gx.size
    ss = np.zeros((bs, npix), dtype=np.float32)
    gparams = np.zeros((P, bs), dtype=np.float32) if g_ss is not None else None
    g_ss_c = np.ascontiguousarray(g_ss, dtype=np.float32) if g_ss is not None else None
    vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = (lib.glh_run_f32x2_bs if straight_line else lib.glh_run_f32x2)(C.byref(cm.desc), C.c_int(bs), vp(mat), C.c_int(npix), vp(gx), vp(gy), vp(ss), vp(g_ss_c), vp(gparams))
    if rc != 0:
        raise RuntimeError(lib.glh_last_error().decode())
    return dict(ss=ss, gparams=gparams)


# ---------------------------------------------------------------- parity rule
def ulp_perturb(mat, seed=99):
    """The same fp32 inputs moved by +-1/2 ulp(fp32) relative: an input error no fp32 computation can
    distinguish from its own first rounding.  Used to probe the conditioning of a test case."""
    rng = np.random.default_rng(seed)
    return np.asarray(mat, dtype=np.float64) * (1.0 + rng.choice([-1.0, 1.0], size=np.shape(mat)) * 2.0 ** -24)


PARITY_LOG = []   # one record per assert_parity call; tests/conftest.py writes it out at session end


def assert_parity(cuda, o32, o64, tol, what, o64_perturbed=None, factor=4.0, axis=None):
    """The parity rule (DESIGN.md "Parity metric").

    error(a) = max|a - o64| / max|o64| over `axis` (None = the whole array); the worst slice decides.
    The CUDA (or float-instantiated) result must satisfy  error <= tol  -- BASELINE.json's 1e-5 for
    images / log-likelihoods, 1e-4 for gradients -- except where fp32 arithmetic itself cannot hold
    `tol` for that input (cuspy Sersic cores under the lens mapping, NFW at X ~ 1, the dPIE log
    ratio): there the bound is `factor` x the fp32 noise floor of the case, measured as the larger of
    (i) the fp32 oracle's own error against the fp64 oracle and (ii) the change of the fp64 oracle
    when its inputs move by half an fp32 ulp (`o64_perturbed`: one result or a list -- e.g. also the fp64 oracle with
    the ray positions beta moved by one fp32 rounding per pixel, oracle_bridge.logprob_and_grad(beta_noise=...)).
    The median slice must hold `tol` outright.

    Every call is recorded in PARITY_LOG (how many slices needed the noise-floor rule and how far above `tol`
    they landed), so a green run says how much of it was green outright: `profiles/r02_parity.json`."""
    cuda, o32, o64 = (np.asarray(v, dtype=np.float64) for v in (cuda, o32, o64))
    red = (lambda v: np.max(v)) if axis is None else (lambda v: np.max(v, axis=axis))
    scale = red(np.abs(o64))
    e_c = np.atleast_1d(red(np.abs(cuda - o64)) / scale)
    floor = np.atleast_1d(red(np.abs(o32 - o64)) / scale)
    if o64_perturbed is not None:   # one perturbed result, or a list of them (the floor is the largest change)
        for pert in (o64_perturbed if isinstance(o64_perturbed, (list, tuple)) else [o64_perturbed]):
            floor = np.maximum(floor, np.atleast_1d(red(np.abs(np.asarray(pert, dtype=np.float64) - o64)) / scale))
    over = e_c > tol                                   # slices that needed the noise-floor rule
    bad = e_c > np.maximum(tol, factor * floor)
    with np.errstate(divide="ignore", invalid="ignore"):
        over_floor = np.where(floor > 0, e_c / floor, np.inf)
    PARITY_LOG.append(dict(
        test=os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], what=str(what), tol=float(tol), factor=float(factor),
        slices=int(e_c.size), slices_over_tol=int(over.sum()), max_err=float(np.max(e_c)), median_err=float(np.median(e_c)),
        max_err_over_tol=float(np.max(e_c) / tol),
        max_err_over_floor_of_slices_over_tol=float(np.max(over_floor[over])) if over.any() else 0.0,
        max_floor=float(np.max(floor)), failed=bool(bad.any())))
    assert not bad.any(), (what, np.nonzero(bad)[0][:8], e_c[bad][:8], floor[bad][:8])
    if axis is not None and np.size(e_c) >= 16:
        assert np.median(e_c) <= tol, (what, "median", float(np.median(e_c)))
    return float(np.max(e_c))


def parity_summary():
    """Per-test roll-up of PARITY_LOG for the committed report."""
    tests = {}
    for r in PARITY_LOG:
        t = tests.setdefault(r["test"], dict(checks=0, slices=0, slices_over_tol=0, max_err_over_tol=0.0,
                                             max_err_over_floor_of_slices_over_tol=0.0, worst=None, failed=False))
        t["checks"] += 1
        t["slices"] += r["slices"]
        t["slices_over_tol"] += r["slices_over_tol"]
        t["failed"] = t["failed"] or r["failed"]
        if r["max_err_over_tol"] > t["max_err_over_tol"]:
            t["max_err_over_tol"] = r["max_err_over_tol"]
            t["worst"] = r["what"]
        t["max_err_over_floor_of_slices_over_tol"] = max(t["max_err_over_floor_of_slices_over_tol"],
                                                         r["max_err_over_floor_of_slices_over_tol"])
    tot_s = sum(t["slices"] for t in tests.values())
    tot_o = sum(t["slices_over_tol"] for t in tests.values())
    return dict(rule="error = max|cuda - oracle64| / max|oracle64| per slice; bound = max(tol, 4 x fp32 noise floor of the slice)",
                total_checks=len(PARITY_LOG), total_slices=tot_s, slices_over_tol=tot_o,
                fraction_over_tol=(tot_o / tot_s) if tot_s else 0.0,
                max_err_over_floor_of_slices_over_tol=max([t["max_err_over_floor_of_slices_over_tol"] for t in tests.values()] or [0.0]),
                tests=tests)


# ---------------------------------------------------------------- cases shared with tests/golden/make_reference_golden.py
def spec_profile(cls, ctor):
    """Product-side descriptor object of a ``(class name, constructor kwargs)`` spec of tests/golden/reference_cases.py."""
    table = {"EPL": gl_epl.EPL, "Shear": gl_shear.Shear, "SIE": gl_sie.SIE, "SIS": gl_sis.SIS, "NFW": gl_nfw.NFW,
             "NFW_ELLIPSE": gl_nfw.NFW_ELLIPSE, "DPIS": gl_piemd.DPIS, "DPIE": gl_piemd.DPIE, "TNFW": gl_tnfw.TNFW,
             "DPIEP": gl_piep.DPIEP, "DPIESubhalo": gl_sub.DPIESubhalo, "Sersic": gl_sersic.Sersic,
             "SersicEllipse": gl_sersic.SersicEllipse, "CoreSersic": gl_sersic.CoreSersic, "Shapelets": gl_shapelets.Shapelets}
    if cls == "ScaledSIS":
        return gl_sr.ScalingRelation(gl_sis.SIS(), ["theta_E"], ctor["lum_star"], ctor["scaling_params_power"], ctor["galaxy_catalogue"])
    return table[cls](**ctor)


def spec_model(model, constants=None):
    kw = {} if constants is None else dict(lenses_constants=constants["lens_mass"], lens_light_constants=constants["lens_light"],
                                           source_light_constants=constants["source_light"])
    return PhysicalModel(*[[spec_profile(c, k) for c, k in model[g]] for g in ("lens_mass", "lens_light", "source_light")], **kw)
