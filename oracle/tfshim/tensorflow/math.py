"""``tensorflow.math`` of the torch-backed stand-in (TEST INFRASTRUCTURE; see ``oracle/tfshim/tensorflow/__init__.py``)."""
import torch as _torch

from tensorflow import _t, _un, _map_dtype, sqrt, exp, sin, cos, abs, atan2, pow, maximum, minimum, reduce_mean  # noqa: A004,F401

log = _un(_torch.log)
atan = _un(_torch.atan)
atanh = _un(_torch.atanh)
acos = _un(_torch.acos)
acosh = _un(_torch.acosh)
lgamma = _un(_torch.lgamma)
is_nan = _un(_torch.isnan)


def reduce_prod(x, axis=None):
    x = _t(x)
    return x.prod() if axis is None else x.prod(dim=axis)


def count_nonzero(x, axis=None, dtype=None):
    n = _torch.count_nonzero(_t(x)) if axis is None else _torch.count_nonzero(_t(x), dim=axis)
    return n.to(_map_dtype(dtype) or _torch.int64)
