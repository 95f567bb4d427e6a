"""TEST INFRASTRUCTURE -- a torch-backed stand-in for the small part of the ``tensorflow`` API that
``/root/reference/src/gigalens/tf/**`` touches (about 60 functions), so that the reference's own, unmodified
source files can be IMPORTED AND EXECUTED in this container (where tensorflow cannot be installed) to produce
golden vectors for the oracle and the CUDA path (``tests/golden/make_reference_golden.py``).

What this pins and what it does not: the arithmetic that runs is the reference's (every clamp, ``where``, sign,
operation order and loop of its profile / simulator / model files); the array backend underneath is torch, not
TensorFlow, so TF's own kernels (conv2d, avg_pool2d, pinv, atan2 ...) are represented by torch's -- IEEE-equivalent for
the elementwise operations, restated from the documented semantics for the structural ones (``SAME`` padding, NHWC
layout, scatter_nd, ``tf.where`` with one argument, ``tf.nest.flatten``'s sorted keys, ``tf.while_loop``).

A "tf tensor" here IS a ``torch.Tensor``.  ``tf.float32`` is the module attribute ``float32``; the golden generator
flips it to ``torch.float64`` (``set_float(torch.float64)``) to run the same reference code as a double-precision arbiter.
Nothing under ``gigalens_b200/`` imports this package."""
import builtins
import math as _pymath

import numpy as _np
import torch as _torch

Tensor = _torch.Tensor
float32 = _torch.float32
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64
bool = _torch.bool   # noqa: A001  (tf.bool)
newaxis = None
_FLOAT = _torch.float32   # what a python float / float numpy array becomes


def set_float(dtype):
    """Switch what ``tf.float32`` means (float32: the reference's arithmetic; float64: the arbiter)."""
    global float32, _FLOAT
    float32 = dtype
    _FLOAT = dtype


if not hasattr(_torch.Tensor, "get_shape"):
    _torch.Tensor.get_shape = lambda self: self.shape


def _map_dtype(dtype):
    if dtype is None:
        return None
    if dtype in (_torch.float32, _torch.float64):
        # tf.float32 was captured by value in a default argument somewhere, or the caller passes ours: honour the switch
        return _FLOAT
    return dtype


def _t(v, dtype=None):
    """Anything -> torch tensor with TF's conversion rules for the cases the reference hits: python floats and float
    arrays become the working float type unless a dtype is given; ints stay integers; lists of tensors are stacked."""
    dtype = _map_dtype(dtype)
    if isinstance(v, _torch.Tensor):
        return v if dtype is None or v.dtype == dtype else v.to(dtype)
    if isinstance(v, (list, tuple)) and len(v) and any(isinstance(e, (_torch.Tensor, list, tuple)) for e in v):
        parts = [_t(e, dtype) for e in v]
        want = dtype
        if want is None:
            want = _FLOAT if any(p.is_floating_point() for p in parts) else parts[0].dtype
        return _torch.stack([p.to(want) for p in parts], dim=0)
    a = _np.asarray(v)
    if a.dtype == object:
        raise TypeError(f"cannot convert {type(v)} to a tensor")
    if dtype is None:
        if a.dtype.kind == "f":
            dtype = _FLOAT if not isinstance(v, _np.ndarray) or a.dtype == _np.float32 else None
            # a float64 numpy ARRAY keeps float64 (tf.constant(np_f64) does); python floats follow the working type
            if dtype is None:
                dtype = _torch.float64
        elif a.dtype.kind == "b":
            dtype = _torch.bool
        elif a.dtype.kind in "iu":
            dtype = _torch.int32 if not isinstance(v, _np.ndarray) and not isinstance(v, _np.generic) else _torch.int64
    return _torch.as_tensor(_np.ascontiguousarray(a)).to(dtype)


def _like(v, ref):
    """Second operand of a binary op: python scalars take the dtype of the tensor operand (TF does the same)."""
    if isinstance(v, _torch.Tensor):
        return v
    if isinstance(ref, _torch.Tensor):
        return _torch.as_tensor(v, dtype=ref.dtype if ref.is_floating_point() or not isinstance(v, builtins.float) else _FLOAT)
    return _t(v)


def _pair(a, b):
    if isinstance(a, _torch.Tensor):
        return a, _like(b, a)
    if isinstance(b, _torch.Tensor):
        return _like(a, b), b
    return _t(a), _t(b)


def function(fn=None, **_kw):
    """``@tf.function``: eager execution of the same python (tracing changes nothing about the arithmetic)."""
    if fn is None:
        return lambda f: f
    return fn


def constant(value, dtype=None, shape=None, name=None):
    out = _t(value, dtype)
    return out if shape is None else out.reshape(shape)


def convert_to_tensor(value, dtype=None, name=None):
    return _t(value, dtype)


def identity(x, name=None):
    return _t(x)


def Variable(value, dtype=None, **_kw):
    return _t(value, dtype).clone().requires_grad_(True)


def cast(x, dtype):
    return _t(x).to(_map_dtype(dtype))


def stop_gradient(x):
    return _t(x).detach()


def zeros(shape, dtype=None):
    return _torch.zeros(_shape_tuple(shape), dtype=_map_dtype(dtype) or _FLOAT)


def eye(n, dtype=None):
    return _torch.eye(int(n), dtype=_map_dtype(dtype) or _FLOAT)


def zeros_like(x, dtype=None):
    return _torch.zeros_like(_t(x), dtype=_map_dtype(dtype))


def ones_like(x, dtype=None):
    return _torch.ones_like(_t(x), dtype=_map_dtype(dtype))


def range(start, limit=None, delta=1, dtype=None):   # noqa: A001
    if limit is None:
        start, limit = 0, start
    return _torch.arange(start, limit, delta, dtype=_map_dtype(dtype))


def linspace(start, stop, num):
    # integer end points are promoted to float64 (the reference calls tf.linspace(-5, 5, 6000) and hands the result to numpy code)
    dt = _FLOAT if isinstance(start, builtins.float) or isinstance(stop, builtins.float) else _torch.float64
    return _torch.linspace(builtins.float(start), builtins.float(stop), int(num), dtype=_torch.float64).to(dt)


def _shape_tuple(shape):
    if isinstance(shape, (builtins.int, _np.integer)):      # tf.zeros(n)
        return (builtins.int(shape),)
    if isinstance(shape, _torch.Tensor) and shape.dim() == 0:
        return (builtins.int(shape),)
    if isinstance(shape, _torch.Tensor):
        return tuple(int(s) for s in shape.tolist())
    return tuple(int(s) for s in shape)


def shape(x):
    return _torch.as_tensor(tuple(_t(x).shape), dtype=_torch.int32)


def size(x, out_type=None):
    n = _t(x).numel()
    return _torch.as_tensor(n, dtype=_map_dtype(out_type) or _torch.int32)


def TensorShape(dims):
    return tuple(dims)


def reshape(x, shape):
    return _t(x).reshape(_shape_tuple(shape))


def transpose(x, perm=None):
    x = _t(x)
    return x.permute(*perm) if perm is not None else x.permute(*reversed(builtins.range(x.dim())))


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def squeeze(x, axis=None):
    return _t(x).squeeze() if axis is None else _t(x).squeeze(axis)


def stack(values, axis=0):
    vals = [_t(v) for v in values]
    vals = _torch.broadcast_tensors(*vals) if len({tuple(v.shape) for v in vals}) > 1 else vals
    return _torch.stack(list(vals), dim=axis)


def concat(values, axis=0):
    vals = [_t(v) for v in values]
    want = _FLOAT if any(v.is_floating_point() for v in vals) else vals[0].dtype
    return _torch.cat([v.to(want) if v.is_floating_point() else v for v in vals], dim=axis)


def repeat(x, repeats, axis=None):
    x = _t(x) if not isinstance(x, _np.ndarray) else _torch.as_tensor(_np.ascontiguousarray(x))
    if isinstance(repeats, (list, tuple)):
        repeats = repeats[0] if len(repeats) == 1 else _torch.as_tensor(repeats)
    if isinstance(repeats, _torch.Tensor) and repeats.dim() == 0:
        repeats = int(repeats)
    return _torch.repeat_interleave(x, repeats, dim=axis)


def gather(params, indices, axis=0):
    params = _t(params)
    idx = _t(indices).long()
    out = _torch.index_select(params, axis, idx.reshape(-1))
    return out.reshape(tuple(params.shape[:axis]) + tuple(idx.shape) + tuple(params.shape[axis + 1:]))


def where(condition, x=None, y=None):
    if x is None and y is None:
        c = condition if isinstance(condition, _torch.Tensor) else _torch.as_tensor(_np.ascontiguousarray(condition))
        return _torch.nonzero(c)          # (K, ndim) int64, row-major order, like tf.where(cond)
    x, y = _pair(x, y)
    return _torch.where(_t(condition), x, y)


def _nd_index(indices):
    idx = _t(indices).long()
    return tuple(idx[:, k] for k in builtins.range(idx.shape[1]))


def tensor_scatter_nd_add(tensor, indices, updates):
    return _t(tensor).index_put(_nd_index(indices), _t(updates).to(tensor.dtype), accumulate=True)


def tensor_scatter_nd_update(tensor, indices, updates):
    return _t(tensor).index_put(_nd_index(indices), _t(updates).to(tensor.dtype), accumulate=False)


def scatter_nd(indices, updates, shape):
    updates = _t(updates)
    return _torch.zeros(_shape_tuple(shape), dtype=updates.dtype).index_put(_nd_index(indices), updates, accumulate=True)


def reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)


def reduce_max(x, axis=None, keepdims=False):
    x = _t(x)
    return x.max() if axis is None else x.amax(dim=axis, keepdim=keepdims)


def clip_by_value(x, lo, hi):
    x = _t(x)
    return _torch.minimum(_torch.maximum(x, _like(lo, x)), _like(hi, x))


def maximum(a, b):
    return _torch.maximum(*_pair(a, b))


def minimum(a, b):
    return _torch.minimum(*_pair(a, b))


def _un(f):
    def op(x, name=None):
        return f(_t(x))
    return op


sqrt = _un(_torch.sqrt)
exp = _un(_torch.exp)
sin = _un(_torch.sin)
cos = _un(_torch.cos)
abs = _un(_torch.abs)   # noqa: A001


def atan2(y, x, name=None):
    return _torch.atan2(*_pair(y, x))


def pow(x, y, name=None):   # noqa: A001
    return _torch.pow(*_pair(x, y))


def einsum(eq, *ops):
    ops = [_t(o) for o in ops]
    want = _FLOAT if any(o.dtype != ops[0].dtype for o in ops) else ops[0].dtype
    eq = eq.replace(" ", "")
    return _torch.einsum(eq, *[o.to(want) for o in ops])


def while_loop(cond, body, loop_vars, shape_invariants=None, maximum_iterations=None, swap_memory=False, name=None,
               parallel_iterations=10, back_prop=True):
    """Eager ``tf.while_loop``: python loop with the same (cond, body, loop_vars, maximum_iterations) contract."""
    state = loop_vars
    it = 0
    while builtins.bool(cond(*state)) and (maximum_iterations is None or it < maximum_iterations):
        out = body(*state)
        state = type(loop_vars)(out) if isinstance(loop_vars, (list, tuple)) else out
        it += 1
    return state


class GradientTape:
    """``tf.GradientTape`` on torch autograd (only what ``tf/profile.py:22-28`` and ``tf/inference.py`` use)."""

    def __init__(self, watch_accessed_variables=True, persistent=False):
        self.persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, x):
        if not x.requires_grad:
            x.requires_grad_(True)

    def gradient(self, target, sources):
        single = isinstance(sources, _torch.Tensor)
        srcs = [sources] if single else list(sources)
        grads = _torch.autograd.grad(target.sum(), srcs, retain_graph=True, create_graph=True, allow_unused=True)
        grads = [zeros_like(s) if g is None else g for g, s in zip(grads, srcs)]
        return grads[0] if single else grads




class _NN:
    """NHWC, cross-correlation, ``SAME`` zero padding (odd kernels / kernel == stride pools: what the reference uses)."""

    @staticmethod
    def _same_pad(k):
        total = k - 1
        return total // 2, total - total // 2     # TF puts the extra element at the end

    @staticmethod
    def conv2d(input, filters, strides=1, padding="SAME", **_kw):
        assert padding == "SAME"
        x = _t(input).permute(0, 3, 1, 2)                   # NHWC -> NCHW
        w = _t(filters).to(x.dtype).permute(3, 2, 0, 1)     # (kh,kw,cin,cout) -> (cout,cin,kh,kw)
        pt, pb = _NN._same_pad(w.shape[2])
        pl, pr = _NN._same_pad(w.shape[3])
        x = _torch.nn.functional.pad(x, (pl, pr, pt, pb))
        return _torch.nn.functional.conv2d(x, w).permute(0, 2, 3, 1)

    @staticmethod
    def depthwise_conv2d(input, filter, strides=(1, 1, 1, 1), padding="SAME", **_kw):
        assert padding == "SAME"
        x = _t(input).permute(0, 3, 1, 2)
        w = _t(filter).to(x.dtype)                          # (kh,kw,cin,mult)
        kh, kw, cin, mult = w.shape
        w = w.permute(2, 3, 0, 1).reshape(cin * mult, 1, kh, kw)
        pt, pb = _NN._same_pad(kh)
        pl, pr = _NN._same_pad(kw)
        x = _torch.nn.functional.pad(x, (pl, pr, pt, pb))
        return _torch.nn.functional.conv2d(x, w, groups=cin).permute(0, 2, 3, 1)

    @staticmethod
    def avg_pool2d(input, ksize, strides, padding="SAME", **_kw):
        k, s = int(ksize), int(strides)
        x = _t(input).permute(0, 3, 1, 2)
        assert k == s and x.shape[2] % k == 0 and x.shape[3] % k == 0, "only the exact-tiling case the reference uses"
        return _torch.nn.functional.avg_pool2d(x, k, s).permute(0, 2, 3, 1)


nn = _NN()


class _Linalg:
    @staticmethod
    def det(a):
        return _torch.linalg.det(_t(a))

    @staticmethod
    def pinv(a, rcond=None):
        """``tf.linalg.pinv``: SVD, singular values <= rcond * max(s) are dropped (documented behaviour)."""
        a = _t(a)
        u, s, vh = _torch.linalg.svd(a, full_matrices=False)
        if rcond is None:
            rcond = 10.0 * builtins.max(a.shape[-2:]) * _torch.finfo(a.dtype).eps
        cutoff = rcond * s.amax(dim=-1, keepdim=True)
        sinv = _torch.where(s > cutoff, 1.0 / _torch.where(s > cutoff, s, _torch.ones_like(s)), _torch.zeros_like(s))
        return (vh.transpose(-1, -2) * sinv.unsqueeze(-2)) @ u.transpose(-1, -2)

    @staticmethod
    def cholesky(a):
        return _torch.linalg.cholesky(_t(a))


linalg = _Linalg()


class _Nest:
    @staticmethod
    def flatten(structure):
        """``tf.nest.flatten``: depth-first, dict values in SORTED KEY order."""
        out = []

        def rec(s):
            if isinstance(s, dict):
                for k in sorted(s):
                    rec(s[k])
            elif isinstance(s, (list, tuple)):
                for e in s:
                    rec(e)
            else:
                out.append(s)
        rec(structure)
        return out


nest = _Nest()


class _Random:
    @staticmethod
    def set_seed(seed):
        _torch.manual_seed(int(seed))


random = _Random()
pi = _pymath.pi

from . import math  # noqa: E402,F401  (a real submodule: the reference does ``from tensorflow.math import atan2``)
