"""TEST INFRASTRUCTURE -- stand-in for the few ``tensorflow_probability`` names the reference's profile / simulator /
model files touch at import or call time (see ``oracle/tfshim/tensorflow/__init__.py`` for the purpose).

Real arithmetic lives in two places only: ``math.interp_regular_1d_grid`` (regular-grid linear interpolation with fill
values, restated from TFP's documented behaviour) and ``distributions.Normal / Independent`` (Gaussian log-density).
The bijector classes are constructible placeholders: ``ForwardProbModel.__init__`` builds its ``pack_bij`` chain, but the
golden generator calls ``stats_pixels`` / ``stats_positions`` with explicit parameter dictionaries and never runs a bijector
(the TFP prior / bijector arithmetic stays restated in ``oracle/model.py``: parity unpinned)."""
import math as _pymath
import types as _types

import torch as _torch

import tensorflow as _tf


def _interp_regular_1d_grid(x, x_ref_min, x_ref_max, y_ref, axis=-1, fill_value="constant_extension",
                            fill_value_below=None, fill_value_above=None, grid_regularizing_transform=None, name=None):
    """y_ref is sampled on ``linspace(x_ref_min, x_ref_max, ny)`` along ``axis``; the result has shape
    ``y_ref.shape[:axis] + x.shape + y_ref.shape[axis+1:]``; outside the grid the fill values apply."""
    x = _tf._t(x)
    y_ref = _tf._t(y_ref).to(x.dtype)
    axis = axis % y_ref.dim()
    ny = y_ref.shape[axis]
    fb = fill_value if fill_value_below is None else fill_value_below
    fa = fill_value if fill_value_above is None else fill_value_above
    idx_unclipped = (x - x_ref_min) / (x_ref_max - x_ref_min) * (ny - 1)
    nan = _torch.isnan(idx_unclipped)
    idx_unclipped = _torch.where(nan, _torch.zeros_like(idx_unclipped), idx_unclipped)
    idx = _torch.clamp(idx_unclipped, 0.0, float(ny - 1))
    below = _torch.floor(idx)
    above = _torch.clamp(below + 1, max=float(ny - 1))
    below = _torch.clamp(above - 1, min=0.0)
    t = idx - below
    flat_b, flat_a = below.long().reshape(-1), above.long().reshape(-1)
    out_shape = tuple(y_ref.shape[:axis]) + tuple(x.shape) + tuple(y_ref.shape[axis + 1:])
    yb = _torch.index_select(y_ref, axis, flat_b).reshape(out_shape)
    ya = _torch.index_select(y_ref, axis, flat_a).reshape(out_shape)
    expand = (None,) * axis + (Ellipsis,) + (None,) * (y_ref.dim() - axis - 1)
    t_e, iu_e, nan_e = t[expand], idx_unclipped[expand], nan[expand]
    y = t_e * ya + (1 - t_e) * yb
    if fb != "constant_extension":
        y = _torch.where(iu_e < 0, _torch.full_like(y, float(fb)), y)
    if fa != "constant_extension":
        y = _torch.where(iu_e > ny - 1, _torch.full_like(y, float(fa)), y)
    return _torch.where(nan_e, _torch.full_like(y, float("nan")), y)


math = _types.SimpleNamespace(interp_regular_1d_grid=_interp_regular_1d_grid)


class _Placeholder:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs

    def _no(self, *a, **k):
        raise NotImplementedError("tfshim: TFP bijector / distribution arithmetic is not provided (restated in oracle/model.py)")

    forward = inverse = forward_log_det_jacobian = sample = _no


class _Normal:
    def __init__(self, loc, scale, **_kw):
        self.loc, self.scale = _tf._t(loc), _tf._t(scale)

    def log_prob(self, x):
        z = (_tf._t(x) - self.loc) / self.scale
        return -0.5 * z * z - _torch.log(self.scale) - 0.5 * _pymath.log(2.0 * _pymath.pi)


class _Independent:
    def __init__(self, distribution, reinterpreted_batch_ndims=0, **_kw):
        self.distribution, self.n = distribution, int(reinterpreted_batch_ndims)

    def log_prob(self, x):
        lp = self.distribution.log_prob(x)
        return lp.sum(dim=tuple(range(-self.n, 0))) if self.n else lp


def _ns(**kw):
    return _types.SimpleNamespace(**kw)


_names_d = ["Distribution", "MultivariateNormalTriL", "MultivariateNormalDiag", "JointDistributionNamed",
            "JointDistributionSequential", "LogNormal", "TruncatedNormal", "Uniform"]
distributions = _ns(Normal=_Normal, Independent=_Independent, **{n: type(n, (_Placeholder,), {}) for n in _names_d})
_names_b = ["Bijector", "Chain", "pack_sequence_as", "Split", "Reshape", "Transpose", "Exp", "FillScaleTriL", "Identity"]
bijectors = _ns(**{n: type(n, (_Placeholder,), {}) for n in _names_b})
util = _ns(TransformedVariable=_Placeholder)
mcmc = _ns()
vi = _ns()
experimental = _ns(mcmc=_ns())
optimizer = _ns()
