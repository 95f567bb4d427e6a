"""Profile descriptors: the reference's ``Parameterized`` / ``LightProfile`` / ``MassProfile``
interface (``src/gigalens/profile.py:5-82``) re-hosted as *descriptors* of CUDA profile types.

A profile object carries its type id and parameter names; the arithmetic lives in
``csrc/gl_math.cuh``.  ``deriv`` / ``light`` still work on arbitrary points (they build a
one-profile plan and evaluate it on the GPU), which is what the reference's profile tests use.
"""
from abc import ABC
from typing import List

import numpy as np


class Parameterized(ABC):
    """``src/gigalens/profile.py:5-21``."""

    _name: str
    _params: List[str]
    _type_id: int = 0

    def __init__(self, *args, **kwargs):
        self.name = self._name
        self.params = list(self._params)

    def __str__(self):
        return self.name


def _eval_points(kind, profile, x, y, kwargs):
    """Evaluate one profile at points on the GPU through the C ABI (gl_eval_points)."""
    import torch

    from . import model as _model
    from . import simulator as _sim

    x = np.asarray(x, dtype=np.float32)
    y = np.asarray(y, dtype=np.float32)
    shape = np.broadcast(x, y).shape
    xf = np.broadcast_to(x, shape).reshape(-1)
    yf = np.broadcast_to(y, shape).reshape(-1)
    names = [p for p in profile.params if p in kwargs]
    consts = {}
    if kind in ("mass", "hessian"):
        pm = _model.PhysicalModel([profile], [], [])
        params = {"lens_mass": [{k: kwargs[k] for k in names}]}
    else:
        pm = _model.PhysicalModel([], [profile], [])
        params = {"lens_light": [{k: kwargs[k] for k in names}]}
    del consts
    bs = 1
    for v in kwargs.values():
        bs = max(bs, int(np.size(v)))
    sim = _sim.LensSimulator(pm, _sim.SimulatorConfig(delta_pix=1.0, num_pix=2), bs=bs)
    if kind == "hessian":
        H = sim.hessian(xf, yf, params["lens_mass"])
        return tuple(h.reshape((bs,) + shape).squeeze(0) if bs == 1 else h.reshape((bs,) + shape) for h in H)
    if kind == "light" and getattr(profile, "use_lstsq", False):
        # the reference returns the linear components with a leading axis (sersic.py:35, shapelets.py:62-63,72-73)
        v = sim.eval_points(params, xf, yf, mode=3)[0]                       # (bs, depth, npts)
        v = v.permute(1, 0, 2).reshape((v.shape[1], bs) + shape)
        return v.squeeze(1) if bs == 1 else v
    out = sim.eval_points(params, xf, yf, mode=1 if kind == "mass" else 2)
    if kind == "mass":
        ax, ay = out
        ax = ax.reshape((bs,) + shape).squeeze(0) if bs == 1 else ax.reshape((bs,) + shape)
        ay = ay.reshape((bs,) + shape).squeeze(0) if bs == 1 else ay.reshape((bs,) + shape)
        return ax, ay
    v = out[0]
    return v.reshape((bs,) + shape).squeeze(0) if bs == 1 else v.reshape((bs,) + shape)


class LightProfile(Parameterized, ABC):
    """``src/gigalens/profile.py:24-60``.  Only the constructor path of ``use_lstsq`` is trusted
    (the reference's setter is broken)."""

    _amp = ""

    def __init__(self, use_lstsq=False, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._use_lstsq = bool(use_lstsq)
        self.depth = 1
        if not self._use_lstsq:
            self.params.append(self._amp)

    @property
    def use_lstsq(self):
        return self._use_lstsq

    def light(self, x, y, **kwargs):
        """Surface brightness at points ``(x, y)`` (shared by all samples); parameters are scalars or
        ``(bs,)`` arrays.  Returns a CUDA tensor of shape ``x.shape`` (``(bs,)+x.shape`` when bs>1); a ``use_lstsq`` profile
        returns its unit-amplitude components with a leading axis of length ``depth``, like the reference."""
        return _eval_points("light", self, x, y, kwargs)


class MassProfile(Parameterized, ABC):
    """``src/gigalens/profile.py:63-82``."""

    def deriv(self, x, y, **kwargs):
        """Deflection ``(alpha_x, alpha_y)`` at points ``(x, y)``; see :meth:`LightProfile.light`."""
        return _eval_points("mass", self, x, y, kwargs)

    def hessian(self, x, y, **kwargs):
        """``(f_xx, f_xy, f_yx, f_yy)``: the Jacobian of :meth:`deriv` (``tf/profile.py:9-30``; the analytic
        ``hessian`` overrides of SIS / Shear / NFW / dPIE are the same function).  The reference's analytic
        ``DPIS.hessian`` scales kappa by ``(r_core + r_cut)/r_cut`` relative to its own ``deriv``
        (``piemd.py:72-73``); this returns the Jacobian."""
        return _eval_points("hessian", self, x, y, kwargs)

    def convergence(self, x, y, **kwargs):
        """``tf/profile.py:32-36``."""
        f_xx, _, _, f_yy = self.hessian(x, y, **kwargs)
        return (f_xx + f_yy) / 2

    def shear(self, x, y, **kwargs):
        """``tf/profile.py:38-43``."""
        f_xx, f_xy, _, f_yy = self.hessian(x, y, **kwargs)
        return (f_xx - f_yy) / 2, f_xy
