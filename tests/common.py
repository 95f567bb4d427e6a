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
                                         scaling_relation as gl_sr, shear as gl_shear, sie as gl_sie, sis as gl_sis)
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
        gl_sersic.SersicEllipse: lambda: OP.SersicEllipse(p.use_lstsq), gl_sersic.Sersic: lambda: OP.Sersic(p.use_lstsq),
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
    "R_sersic": (0.3, 1.2), "n_sersic": (0.8, 5.0), "Ie": (20.0, 300.0), "beta": (0.08, 0.2),
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
             want_beta=False):
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
    g_ss_c = np.ascontiguousarray(g_ss, dtype=dtype) if g_ss is not None else None
    vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = fn(C.byref(cm.desc), C.c_int(bs), vp(mat), C.c_int(npix), vp(gx), vp(gy), C.c_int(int(no_deflection)),
            C.c_int(int(epl_batch_max)), vp(ss), vp(g_ss_c), vp(gparams), vp(beta))
    if rc != 0:
        raise RuntimeError(lib.glh_last_error().decode())
    return dict(ss=ss, gparams=gparams, beta=beta)
