"""Prior specification shim: the handful of ``tensorflow_probability.distributions`` classes the
reference's model specs use (``docsrc/source/model-spec.rst``, ``tests/conftest.py:21-73``), with
the TFP semantics the hot path depends on:

* flattening order of ``tf.nest.flatten``: lists in order, dict keys sorted;
* default event-space bijectors: Normal -> Identity, LogNormal -> Exp, Uniform and
  TruncatedNormal -> Sigmoid(low, high);
* log-densities.

These objects are *specifications*: ``log_prob`` and the bijector run inside the CUDA library
(``gl_unconstrain`` / ``gl_logprob_grad``); only sampling (numpy, host) happens here.
"""
import math

import numpy as np

from . import _cabi


class Distribution:
    dist_id = -1

    def leaf(self):
        """(dist id, a, b, low, high) as handed to gl_prior_leaf."""
        raise NotImplementedError

    def sample_np(self, rng, n):
        raise NotImplementedError

    def inverse_np(self, x):
        """Unconstraining map (inverse of the default event-space bijector), host numpy."""
        raise NotImplementedError

    def forward_np(self, z):
        raise NotImplementedError


class Normal(Distribution):
    dist_id = _cabi.GL_DIST_NORMAL

    def __init__(self, loc, scale):
        self.loc, self.scale = float(loc), float(scale)

    def leaf(self):
        return (self.dist_id, self.loc, self.scale, 0.0, 0.0)

    def sample_np(self, rng, n):
        return rng.normal(self.loc, self.scale, size=n)

    def inverse_np(self, x):
        return np.asarray(x, dtype=np.float64)

    def forward_np(self, z):
        return np.asarray(z, dtype=np.float64)


class LogNormal(Distribution):
    dist_id = _cabi.GL_DIST_LOGNORMAL

    def __init__(self, loc, scale):
        self.loc, self.scale = float(loc), float(scale)

    def leaf(self):
        return (self.dist_id, self.loc, self.scale, 0.0, 0.0)

    def sample_np(self, rng, n):
        return np.exp(rng.normal(self.loc, self.scale, size=n))

    def inverse_np(self, x):
        return np.log(np.asarray(x, dtype=np.float64))

    def forward_np(self, z):
        return np.exp(np.asarray(z, dtype=np.float64))


class _Bounded(Distribution):
    low: float
    high: float

    def inverse_np(self, x):
        u = (np.asarray(x, dtype=np.float64) - self.low) / (self.high - self.low)
        return np.log(u) - np.log1p(-u)

    def forward_np(self, z):
        z = np.asarray(z, dtype=np.float64)
        return self.low + (self.high - self.low) / (1.0 + np.exp(-z))


class Uniform(_Bounded):
    dist_id = _cabi.GL_DIST_UNIFORM

    def __init__(self, low, high):
        self.low, self.high = float(low), float(high)

    def leaf(self):
        return (self.dist_id, self.low, self.high, self.low, self.high)

    def sample_np(self, rng, n):
        return rng.uniform(self.low, self.high, size=n)


class TruncatedNormal(_Bounded):
    dist_id = _cabi.GL_DIST_TRUNCNORMAL

    def __init__(self, loc, scale, low, high):
        self.loc, self.scale, self.low, self.high = float(loc), float(scale), float(low), float(high)

    def leaf(self):
        return (self.dist_id, self.loc, self.scale, self.low, self.high)

    def sample_np(self, rng, n):
        out = np.empty(n)
        filled = 0
        while filled < n:
            d = rng.normal(self.loc, self.scale, size=2 * (n - filled) + 8)
            d = d[(d > self.low) & (d < self.high)][: n - filled]
            out[filled:filled + len(d)] = d
            filled += len(d)
        return out


def _flatten(struct, path=()):
    if isinstance(struct, JointDistribution):
        struct = struct.model
    if isinstance(struct, dict):
        out = []
        for k in sorted(struct):
            out += _flatten(struct[k], path + (k,))
        return out
    if isinstance(struct, (list, tuple)):
        out = []
        for i, v in enumerate(struct):
            out += _flatten(v, path + (i,))
        return out
    return [(path, struct)]


def _map_structure(struct, leaves_iter):
    if isinstance(struct, JointDistribution):
        struct = struct.model
    if isinstance(struct, dict):
        vals = {k: _map_structure(struct[k], leaves_iter) for k in sorted(struct)}
        return {k: vals[k] for k in struct}
    if isinstance(struct, (list, tuple)):
        return [_map_structure(v, leaves_iter) for v in struct]
    return next(leaves_iter)


class JointDistribution(Distribution):
    """Nested dict / list of independent leaves (``tfd.JointDistributionNamed`` /
    ``tfd.JointDistributionSequential`` as the reference's specs use them)."""

    def __init__(self, model):
        self.model = model

    @property
    def leaves(self):
        return _flatten(self.model)

    def sample(self, sample_shape=(), seed=None):
        """Pytree of float32 numpy arrays of shape ``sample_shape`` (an int, a tuple or () like TFP; a tuple such as
        ``(num_particles, num_ensembles)`` -- ``tf/inference.py:199`` -- gives leaves of exactly that shape)."""
        n = int(np.prod(sample_shape)) if np.size(sample_shape) else 1
        rng = np.random.default_rng(seed)
        cols = [np.asarray(d.sample_np(rng, n), dtype=np.float32) for _, d in self.leaves]
        if not np.size(sample_shape):
            cols = [c[0] for c in cols]
        elif np.ndim(sample_shape) == 1 and len(sample_shape) > 1:
            cols = [c.reshape(tuple(int(k) for k in sample_shape)) for c in cols]
        return _map_structure(self.model, iter(cols))

    def flatten_values(self, values):
        """Pytree of values with the prior's structure -> list of leaves in flatten order."""
        return [v for _, v in _flatten_values(self.model, values)]

    def pack(self, leaves):
        return _map_structure(self.model, iter(leaves))


def _flatten_values(struct, values, path=()):
    if isinstance(struct, JointDistribution):
        struct = struct.model
    if isinstance(struct, dict):
        out = []
        for k in sorted(struct):
            out += _flatten_values(struct[k], values[k], path + (k,))
        return out
    if isinstance(struct, (list, tuple)):
        out = []
        for i, v in enumerate(struct):
            out += _flatten_values(v, values[i], path + (i,))
        return out
    return [(path, values)]


class JointDistributionNamed(JointDistribution):
    pass


class JointDistributionSequential(JointDistribution):
    pass
