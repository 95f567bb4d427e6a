"""``PhysicalModel`` / ``ForwardProbModel`` / ``BackwardProbModel`` with the reference's API
surface (``src/gigalens/model.py:7-73``, ``src/gigalens/tf/model.py:12-306``).

``log_prob(simulator, z)`` returns ``(log_prob, red_chi2)`` exactly like the reference; its
gradient comes from the hand-written adjoint kernels (``gl_logprob_grad``), exposed both as
``log_prob_and_grad`` and through ``torch.autograd`` (``z.requires_grad``) so that optimiser and
sampler code written against autograd keeps working without any framework graph of the model.
"""
import ctypes as C
from typing import Dict, List

import numpy as np

from . import _cabi
from . import distributions as tfd
from .simulator import GROUPS


class PhysicalModelBase:
    """``src/gigalens/model.py:7-44``."""

    def __init__(self, lenses, lens_light, source_light, lenses_constants: List[Dict] = None,
                 lens_light_constants: List[Dict] = None, source_light_constants: List[Dict] = None):
        self.lenses = list(lenses)
        self.lens_light = list(lens_light)
        self.source_light = list(source_light)
        self.lenses_constants = lenses_constants if lenses_constants is not None else [dict() for _ in self.lenses]
        self.lens_light_constants = lens_light_constants if lens_light_constants is not None \
            else [dict() for _ in self.lens_light]
        self.source_light_constants = source_light_constants if source_light_constants is not None \
            else [dict() for _ in self.source_light]


class PhysicalModel(PhysicalModelBase):
    """``src/gigalens/tf/model.py:276-306`` (constants are cast to fp32)."""

    def __init__(self, lenses, lens_light, source_light, lenses_constants=None, lens_light_constants=None,
                 source_light_constants=None):
        super().__init__(lenses, lens_light, source_light, lenses_constants, lens_light_constants,
                         source_light_constants)
        cast = lambda ds: [{k: np.float32(v) for k, v in d.items()} for d in ds]
        self.lenses_constants = cast(self.lenses_constants)
        self.lens_light_constants = cast(self.lens_light_constants)
        self.source_light_constants = cast(self.source_light_constants)


class _HostBijector:
    """``prob_model.bij`` as scripts use it between runs (``bij.inverse(prior.sample(n))`` to start a driver,
    ``bij.forward(z)`` to read a result: ``tests/tf/test_model.py:10-16,36-37``): the default event-space bijector
    chained with the pack / split of ``tf/model.py:76-87``, on the host in numpy.  Inside the hot path the same map
    runs in ``k_unconstrain`` (``ProbabilisticModel.bij_forward`` returns device tensors)."""

    def __init__(self, owner):
        self._owner = owner

    def inverse(self, params):
        return self._owner.bij_inverse(params)

    def forward(self, z):
        z = np.asarray(z.detach().cpu() if hasattr(z, "detach") else z, dtype=np.float64)
        z = z.reshape(-1, self._owner.size)
        cols = [np.asarray(d.forward_np(z[:, k]), dtype=np.float32) for k, (_, d) in enumerate(self._owner._leaves)]
        return self._owner.prior.pack(cols)


class ProbabilisticModel:
    """``src/gigalens/model.py:47-73`` plus the prior plumbing shared by both concrete models."""

    def __init__(self, prior, bij=None, *args):
        if not isinstance(prior, tfd.JointDistribution):
            prior = tfd.JointDistribution(prior)
        self.prior = prior
        self._leaves = prior.leaves  # [(path, distribution)] in tf.nest.flatten order
        self.size = len(self._leaves)
        self.bij = bij if bij is not None else _HostBijector(self)

    # -- path of a prior leaf -> simulator slot key
    @staticmethod
    def _slot_key(path):
        g = path[0]
        group = GROUPS[g] if isinstance(g, int) else g
        return (group, int(path[1]), path[2])

    def _leaf_array(self, simulator):
        arr = (_cabi.PriorLeaf * self.size)()
        slots = simulator.compiled.slots
        used = set()
        for k, (path, dist) in enumerate(self._leaves):
            key = self._slot_key(path)
            if key not in slots:
                raise KeyError(f"prior leaf {path} does not correspond to a free parameter of the physical model")
            did, a, b, lo, hi = dist.leaf()
            arr[k].dist, arr[k].slot, arr[k].a, arr[k].b, arr[k].low, arr[k].high = did, slots[key], a, b, lo, hi
            used.add(slots[key])
        if len(used) != simulator.compiled.n_params:
            missing = [k for k, s in slots.items() if s not in used]
            raise KeyError(f"free parameters without a prior: {missing}")
        return arr

    def _bind(self, simulator):
        """Install this model's prior (and likelihood data) on the simulator's plan once."""
        lib = simulator._lib
        if simulator._prior_owner is not self:
            arr = self._leaf_array(simulator)
            _cabi.check(lib.gl_plan_set_prior(simulator._plan, arr, self.size), lib)
            simulator._prior_owner = self
        if simulator._like_owner is not self:
            self._install_likelihood(simulator)
            simulator._like_owner = self

    def _install_likelihood(self, simulator):
        raise NotImplementedError

    # -- bijector on the host (pack / unpack like tf/model.py:76-87)
    def bij_inverse(self, params, bs=None):
        """Physical pytree (leaves scalar, ``(bs,)`` or any leading shape, flattened in C order) -> unconstrained
        ``z`` numpy ``(bs, d)``."""
        vals = self.prior.flatten_values(params)
        host = lambda v: np.asarray(v.detach().cpu() if hasattr(v, "detach") else v)
        cols = [np.atleast_1d(d.inverse_np(host(v))).reshape(-1) for (_, d), v in zip(self._leaves, vals)]
        n = max(len(c) for c in cols)
        return np.stack([np.broadcast_to(c, (n,)) for c in cols], 1).astype(np.float32)

    def bij_forward(self, simulator, z):
        """``bij.forward(z)``: unconstrained ``z (bs, d)`` -> params pytree of CUDA ``(bs,)`` tensors."""
        torch = simulator._torch
        self._bind(simulator)
        z = torch.as_tensor(z, dtype=torch.float32, device=simulator.device).contiguous()
        mat = torch.empty((max(1, simulator.compiled.n_params), simulator.bs), dtype=torch.float32, device=simulator.device)
        _cabi.check(simulator._lib.gl_unconstrain(simulator._plan, z.data_ptr(), mat.data_ptr(), None,
                                                  simulator._stream()), simulator._lib)
        return simulator.compiled.unflatten(mat)

    def _eval(self, simulator, z, want_grad):
        torch = simulator._torch
        self._bind(simulator)
        z = z.to(device=simulator.device, dtype=torch.float32).contiguous()
        if z.shape != (simulator.bs, self.size):
            raise ValueError(f"z must have shape ({simulator.bs}, {self.size}), got {tuple(z.shape)}")
        logp = torch.empty((simulator.bs,), dtype=torch.float32, device=simulator.device)
        chi = torch.empty_like(logp)
        dz = torch.empty_like(z) if want_grad else None
        _cabi.check(simulator._lib.gl_logprob_grad(simulator._plan, z.data_ptr(), logp.data_ptr(), chi.data_ptr(),
                                                   dz.data_ptr() if want_grad else None, simulator._stream()),
                    simulator._lib)
        return logp, chi, dz

    def log_prob_and_grad(self, simulator, z):
        """``(log_prob, red_chi2, d log_prob / d z)``: one fused forward + hand-adjoint pass."""
        torch = simulator._torch
        return self._eval(simulator, torch.as_tensor(z), True)

    def log_prob(self, simulator, z):
        """Reference ``log_prob(simulator, z) -> (log_like + log_prior, red_chi2)`` (``tf/model.py:126-167``,
        ``:242-273``); differentiable through torch.autograd via the hand-written adjoint."""
        torch = simulator._torch
        z = torch.as_tensor(z)
        if torch.is_grad_enabled() and z.requires_grad:
            return _autograd_wrap(torch, lambda zz, g: self._eval(simulator, zz, g), z)
        logp, chi, _ = self._eval(simulator, z, False)
        return logp, chi

    def log_prior(self, simulator, z):
        """``tf/model.py:183-185``: prior.log_prob(bij.forward(z)) + forward_log_det_jacobian."""
        torch = simulator._torch
        self._bind(simulator)
        z = torch.as_tensor(z, dtype=torch.float32, device=simulator.device).contiguous()
        lp = torch.empty((simulator.bs,), dtype=torch.float32, device=simulator.device)
        _cabi.check(simulator._lib.gl_unconstrain(simulator._plan, z.data_ptr(), None, lp.data_ptr(),
                                                  simulator._stream()), simulator._lib)
        return lp


    def log_prior_and_grad(self, simulator, z):
        """``log_prior(z)`` and its gradient w.r.t. ``z`` (bijector chain rule only, ``gl_chain_grad``)."""
        torch = simulator._torch
        self._bind(simulator)
        z = torch.as_tensor(z, dtype=torch.float32, device=simulator.device).contiguous()
        lp = torch.empty((simulator.bs,), dtype=torch.float32, device=simulator.device)
        dz = torch.empty_like(z)
        _cabi.check(simulator._lib.gl_chain_grad(simulator._plan, z.data_ptr(), None, 1, lp.data_ptr(), dz.data_ptr(),
                                                 simulator._stream()), simulator._lib)
        return lp, dz

    def chain_to_z(self, simulator, z, dparams):
        """``(d params / d z)^T dparams``: pull a ``[P][bs]`` parameter gradient back to ``z`` (no prior term)."""
        torch = simulator._torch
        self._bind(simulator)
        dz = torch.empty_like(z)
        _cabi.check(simulator._lib.gl_chain_grad(simulator._plan, z.data_ptr(), dparams.data_ptr(), 0, None, dz.data_ptr(),
                                                 simulator._stream()), simulator._lib)
        return dz


def _autograd_wrap(torch, fn, z):
    """Attach the hand-written gradient to autograd when ``z`` requires grad."""

    class _F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, zz):
            logp, chi, dz = fn(zz.detach(), True)
            ctx.save_for_backward(dz)
            ctx.mark_non_differentiable(chi)
            return logp, chi

        @staticmethod
        def backward(ctx, g_logp, _g_chi):
            (dz,) = ctx.saved_tensors
            return g_logp[:, None] * dz

    return _F.apply(z)


class ForwardProbModel(ProbabilisticModel):
    """``src/gigalens/tf/model.py:12-194``: pixel likelihood with the *simulated* image as the variance
    estimate (``stats_pixels``) and/or the image-position likelihood of multiply-imaged sources
    (``stats_positions``).  Signature and defaults are the reference's (``tf/model.py:32-44``, ``include_positions=True``).
    The reference then iterates ``centroids_x`` unconditionally and raises ``TypeError`` when no centroids were given --
    which is how its own notebooks and ``tests/tf/test_model.py`` call it -- so here ``include_positions`` without
    centroids simply means "no position term"."""

    def __init__(self, prior, observed_image=None, background_rms=None, exp_time=None, error_map=None,
                 centroids_x=None, centroids_y=None, centroids_errors_x=None, centroids_errors_y=None,
                 include_pixels=True, include_positions=True):
        super().__init__(prior)
        if include_positions and centroids_x is None:
            include_positions = False
        if not include_pixels and not include_positions:
            raise ValueError("include_pixels=False and include_positions=False leave no likelihood term")
        self.include_pixels = bool(include_pixels)
        self.include_positions = bool(include_positions)
        self.observed_image = self.error_map = self.background_rms = self.exp_time = None
        if self.include_pixels:  # tf/model.py:62-68
            self.observed_image = np.ascontiguousarray(observed_image, dtype=np.float32)
            self.error_map = None if error_map is None else np.ascontiguousarray(error_map, dtype=np.float32)
            self.background_rms = None if background_rms is None else float(np.float32(background_rms))
            self.exp_time = None if exp_time is None else float(np.float32(exp_time))
            if self.error_map is None and (self.background_rms is None or self.exp_time is None):
                raise ValueError("give either error_map or background_rms and exp_time")
        self.centroids_x = self.centroids_y = self.centroids_errors_x = self.centroids_errors_y = None
        self.centroids_x_batch = self.centroids_y_batch = None
        if self.include_positions:  # tf/model.py:69-74
            if any(v is None for v in (centroids_x, centroids_y, centroids_errors_x, centroids_errors_y)):
                raise ValueError("include_positions=True needs centroids_x, centroids_y, centroids_errors_x, centroids_errors_y")
            f32 = lambda group: [np.atleast_1d(np.asarray(c, dtype=np.float32)) for c in group]
            self.centroids_x, self.centroids_y = f32(centroids_x), f32(centroids_y)
            self.centroids_errors_x, self.centroids_errors_y = f32(centroids_errors_x), f32(centroids_errors_y)
            self.n_position = 2 * sum(c.size for c in self.centroids_x)

    def init_centroids(self, bs):
        """``tf/model.py:185-194``: the centroids tiled over the batch, ``(n_img, bs)`` per system, as
        ``centroids_x_batch`` / ``centroids_y_batch``.  The likelihood kernels share one copy between all samples and do
        not read these; they exist so that scripts which feed them to ``simulator.beta`` / ``magnification`` (which
        accept this tiled layout and answer in it) keep working."""
        if self.include_positions:
            self.centroids_x_batch = [np.repeat(cx[:, None], bs, axis=-1) for cx in self.centroids_x]
            self.centroids_y_batch = [np.repeat(cy[:, None], bs, axis=-1) for cy in self.centroids_y]

    def _install_likelihood(self, simulator):
        if self.include_pixels:
            lc = _cabi.LikeConfig()
            lc.observed = self.observed_image.ctypes.data_as(C.POINTER(C.c_float))
            if self.error_map is not None:
                lc.error_map = self.error_map.ctypes.data_as(C.POINTER(C.c_float))
            lc.background_rms = self.background_rms or 0.0
            lc.exp_time = self.exp_time or 1.0
            n = simulator.numPix
            if self.observed_image.shape != (n, n):
                raise ValueError(f"observed_image must be ({n}, {n})")
            _cabi.check(simulator._lib.gl_plan_set_likelihood(simulator._plan, C.byref(lc)), simulator._lib)
        if self.include_positions:
            simulator.set_positions(self.centroids_x, self.centroids_y, self.centroids_errors_x, self.centroids_errors_y)
        simulator.set_option("lstsq", 0)
        simulator.set_option("include_pixels", int(self.include_pixels))
        simulator.set_option("include_positions", int(self.include_positions))

    def _own(self, simulator):
        if simulator._like_owner is not self:
            self._install_likelihood(simulator)
            simulator._like_owner = self

    def stats_positions(self, simulator, params):
        """``tf/model.py:103-124`` -> ``(log_like, red_chi2)``, both ``(bs,)``."""
        if not self.include_positions:
            raise ValueError("this model was built with include_positions=False")
        self._own(simulator)
        return simulator.positions_loglike(params)

    def stats_pixels(self, simulator, params):
        """``tf/model.py:89-101`` -> ``(log_like, red_chi2)``, both ``(bs,)``."""
        torch = simulator._torch
        if not self.include_pixels:
            raise ValueError("this model was built with include_pixels=False")
        self._own(simulator)
        mat = simulator._params_matrix(params)
        ll = torch.empty((simulator.bs,), dtype=torch.float32, device=simulator.device)
        chi = torch.empty_like(ll)
        if self.include_positions:
            simulator.set_option("include_positions", 0)
        try:
            _cabi.check(simulator._lib.gl_loglike_grad(simulator._plan, mat.data_ptr(), ll.data_ptr(), chi.data_ptr(), None,
                                                       simulator._stream()), simulator._lib)
        finally:
            if self.include_positions:
                simulator.set_option("include_positions", 1)
        return ll, chi

    def term_and_grad(self, simulator, z, term):
        """One likelihood term and its gradient w.r.t. ``z``: ``term`` is ``'pixels'`` (``stats_pixels``),
        ``'positions'`` (``stats_positions``) or ``'none'`` (zeros) -- the target / auxiliary choices of
        ``ModellingSequence.SMC`` (``tf/inference.py:208-214``).  Returns ``(log_like (bs,), dz (bs, d))``."""
        torch = simulator._torch
        z = torch.as_tensor(z, dtype=torch.float32, device=simulator.device).contiguous()
        if term == "none":
            return torch.zeros(simulator.bs, device=simulator.device), torch.zeros_like(z)
        params = self.bij_forward(simulator, z)
        self._own(simulator)
        mat = simulator._params_matrix(params)
        if term == "positions":
            if not self.include_positions:
                raise ValueError("term 'positions' needs a model built with centroids")
            ll, _, g = simulator.positions_loglike(mat, want_grad=True)
        elif term == "pixels":
            if not self.include_pixels:
                raise ValueError("term 'pixels' needs a model built with an observed image")
            ll = torch.empty((simulator.bs,), dtype=torch.float32, device=simulator.device)
            chi, g = torch.empty_like(ll), torch.empty_like(mat)
            if self.include_positions:
                simulator.set_option("include_positions", 0)
            try:
                _cabi.check(simulator._lib.gl_loglike_grad(simulator._plan, mat.data_ptr(), ll.data_ptr(), chi.data_ptr(),
                                                           g.data_ptr(), simulator._stream()), simulator._lib)
            finally:
                if self.include_positions:
                    simulator.set_option("include_positions", 1)
        else:
            raise ValueError(f"unknown likelihood term {term!r}")
        return ll, self.chain_to_z(simulator, z, g)

    def loglike_and_grad(self, simulator, params):
        """log-likelihood (pixels and/or positions as configured), red_chi2 and d(log_like)/d(params) as a
        ``[P][bs]`` matrix."""
        torch = simulator._torch
        self._own(simulator)
        mat = simulator._params_matrix(params)
        ll = torch.empty((simulator.bs,), dtype=torch.float32, device=simulator.device)
        chi = torch.empty_like(ll)
        g = torch.empty_like(mat)
        _cabi.check(simulator._lib.gl_loglike_grad(simulator._plan, mat.data_ptr(), ll.data_ptr(), chi.data_ptr(),
                                                   g.data_ptr(), simulator._stream()), simulator._lib)
        return ll, chi, g

    def log_like(self, simulator, z):
        """``tf/model.py:169-180``."""
        params = self.bij_forward(simulator, z)
        self._own(simulator)
        mat = simulator._params_matrix(params)
        torch = simulator._torch
        ll = torch.empty((simulator.bs,), dtype=torch.float32, device=simulator.device)
        chi = torch.empty_like(ll)
        _cabi.check(simulator._lib.gl_loglike_grad(simulator._plan, mat.data_ptr(), ll.data_ptr(), chi.data_ptr(), None,
                                                   simulator._stream()), simulator._lib)
        return ll


class BackwardProbModel(ProbabilisticModel):
    """``src/gigalens/tf/model.py:197-273``: fixed noise map from the observation, linear light
    amplitudes solved by least squares inside ``lstsq_simulate``."""

    def __init__(self, prior, observed_image, background_rms, exp_time):
        super().__init__(prior)
        obs = np.ascontiguousarray(observed_image, dtype=np.float32)
        self.observed_image = obs
        self.err_map = np.sqrt(np.float32(background_rms) ** 2 + np.clip(obs, 0, np.inf) / np.float32(exp_time)).astype(np.float32)

    def _install_likelihood(self, simulator):
        lc = _cabi.LikeConfig()
        lc.observed = self.observed_image.ctypes.data_as(C.POINTER(C.c_float))
        lc.error_map = self.err_map.ctypes.data_as(C.POINTER(C.c_float))
        lc.background_rms, lc.exp_time = 0.0, 1.0
        n = simulator.numPix
        if self.observed_image.shape != (n, n):
            raise ValueError(f"observed_image must be ({n}, {n})")
        _cabi.check(simulator._lib.gl_plan_set_likelihood(simulator._plan, C.byref(lc)), simulator._lib)
        if not simulator._lstsq_reserved:
            simulator.reserve_lstsq()
        simulator.set_option("lstsq", 1)   # log-prob entry points use the linear-amplitude solve
        simulator.set_option("include_pixels", 1)
        simulator.set_option("include_positions", 0)

    def loglike_and_grad(self, simulator, params):
        """log-likelihood, red_chi2 and d(log_like)/d(non-linear params) ``[P][bs]``."""
        torch = simulator._torch
        if simulator._like_owner is not self:
            self._install_likelihood(simulator)
            simulator._like_owner = self
        mat = simulator._params_matrix(params)
        ll = torch.empty((simulator.bs,), dtype=torch.float32, device=simulator.device)
        chi = torch.empty_like(ll)
        g = torch.empty_like(mat)
        _cabi.check(simulator._lib.gl_lstsq_loglike_grad(simulator._plan, mat.data_ptr(), ll.data_ptr(), chi.data_ptr(),
                                                         g.data_ptr(), simulator._stream()), simulator._lib)
        return ll, chi, g
