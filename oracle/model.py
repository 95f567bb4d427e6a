"""Oracle restatement of the probabilistic models (TEST INFRASTRUCTURE).

Follows ``src/gigalens/tf/model.py:12-273``.  The TFP pieces the reference leans on
(``JointDistribution*`` flattening, default event-space bijectors, log-densities) live in
third-party code that is not installed here; they are restated from their documented
behaviour and are PARITY UNPINNED against TFP.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

_LOG_2PI = math.log(2 * math.pi)


# ----------------------------------------------------------------- distributions


class Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale = float(loc), float(scale)

    def forward(self, z):  # Identity bijector
        return z

    def fldj(self, z):
        return torch.zeros_like(z)

    def log_prob(self, x):
        return -0.5 * ((x - self.loc) / self.scale) ** 2 - 0.5 * _LOG_2PI - math.log(self.scale)

    def sample(self, rng, n):
        return rng.normal(self.loc, self.scale, size=n)

    def inverse(self, x):
        return x


class LogNormal:
    def __init__(self, loc, scale):
        self.loc, self.scale = float(loc), float(scale)

    def forward(self, z):  # Exp bijector
        return torch.exp(z)

    def fldj(self, z):
        return z

    def log_prob(self, x):
        lx = torch.log(x)
        return -lx - math.log(self.scale) - 0.5 * _LOG_2PI - 0.5 * ((lx - self.loc) / self.scale) ** 2

    def sample(self, rng, n):
        return np.exp(rng.normal(self.loc, self.scale, size=n))

    def inverse(self, x):
        return torch.log(x)


class _Bounded:
    low: float
    high: float

    def forward(self, z):  # Sigmoid(low, high)
        diff = self.high - self.low
        left = self.low + diff * torch.sigmoid(z)
        right = self.high - diff * torch.sigmoid(-z)
        return torch.where(z < 0, left, right)

    def fldj(self, z):
        return math.log(self.high - self.low) - F.softplus(-z) - F.softplus(z)

    def inverse(self, x):
        u = (x - self.low) / (self.high - self.low)
        return torch.log(u) - torch.log1p(-u)


class Uniform(_Bounded):
    def __init__(self, low, high):
        self.low, self.high = float(low), float(high)

    def log_prob(self, x):
        return torch.full_like(x, -math.log(self.high - self.low))

    def sample(self, rng, n):
        return rng.uniform(self.low, self.high, size=n)


class TruncatedNormal(_Bounded):
    def __init__(self, loc, scale, low, high):
        self.loc, self.scale, self.low, self.high = float(loc), float(scale), float(low), float(high)
        a, b = (self.low - self.loc) / self.scale, (self.high - self.loc) / self.scale
        ndtr = lambda v: 0.5 * (1 + math.erf(v / math.sqrt(2)))
        self.log_norm = math.log(ndtr(b) - ndtr(a))

    def log_prob(self, x):
        return (-0.5 * ((x - self.loc) / self.scale) ** 2 - 0.5 * _LOG_2PI - math.log(self.scale) - self.log_norm)

    def sample(self, rng, n):
        out = np.empty(n)
        filled = 0
        while filled < n:
            d = rng.normal(self.loc, self.scale, size=2 * (n - filled) + 8)
            d = d[(d > self.low) & (d < self.high)][: n - filled]
            out[filled:filled + len(d)] = d
            filled += len(d)
        return out


# ----------------------------------------------------------------- structure


def flatten_structure(struct, path=()):
    """``tf.nest.flatten`` order: lists in order, dict keys sorted."""
    if isinstance(struct, dict):
        out = []
        for k in sorted(struct):
            out += flatten_structure(struct[k], path + (k,))
        return out
    if isinstance(struct, (list, tuple)):
        out = []
        for i, v in enumerate(struct):
            out += flatten_structure(v, path + (i,))
        return out
    return [(path, struct)]


def pack_like(struct, leaves):
    it = iter(leaves)

    def rec(s):
        if isinstance(s, dict):
            vals = {k: rec(s[k]) for k in sorted(s)}
            return {k: vals[k] for k in s}
        if isinstance(s, (list, tuple)):
            return [rec(v) for v in s]
        return next(it)

    return rec(struct)


class JointPrior:
    """A nested dict/list of independent leaf distributions (the fork's model-spec shape:
    ``{'lens_mass': [..], 'lens_light': [..], 'source_light': [..]}``)."""

    def __init__(self, struct):
        self.struct = struct
        self.leaves = flatten_structure(struct)
        self.size = len(self.leaves)

    def sample(self, n, seed=0):
        rng = np.random.default_rng(seed)
        cols = [np.asarray(d.sample(rng, n), dtype=np.float32) for _, d in self.leaves]
        return np.stack(cols, 1)  # (n, d) physical values, flatten order

    def forward(self, z):
        """bij.forward: z (bs, d) -> (params pytree of (bs,) leaves, list of leaves)."""
        leaves = [d.forward(z[:, k]) for k, (_, d) in enumerate(self.leaves)]
        return pack_like(self.struct, leaves), leaves

    def inverse(self, x):
        return torch.stack([d.inverse(x[:, k]) for k, (_, d) in enumerate(self.leaves)], 1)

    def log_prior(self, z):
        """prior.log_prob(params) + unconstraining_bij.fldj  (tf/model.py:164-166)."""
        lp = torch.zeros(z.shape[0], dtype=z.dtype)
        for k, (_, d) in enumerate(self.leaves):
            x = d.forward(z[:, k])
            lp = lp + d.log_prob(x) + d.fldj(z[:, k])
        return lp


# ----------------------------------------------------------------- prob models


class ForwardProbModel:
    """``tf/model.py:12-194``: pixel likelihood (``stats_pixels``) and image-position likelihood
    (``stats_positions``).  ``include_positions`` defaults to False here because most oracle callers have
    no centroids; the reference's default is True (``:44``)."""

    def __init__(self, prior: JointPrior, observed_image=None, background_rms=None, exp_time=None, error_map=None,
                 centroids_x=None, centroids_y=None, centroids_errors_x=None, centroids_errors_y=None,
                 include_pixels=True, include_positions=False, dtype=torch.float32):
        self.prior = prior
        self.dtype = dtype
        self.include_pixels = include_pixels
        self.include_positions = include_positions
        self.observed_image = self.error_map = self.background_rms = self.exp_time = None
        if include_pixels:
            self.observed_image = torch.as_tensor(np.asarray(observed_image, dtype=np.float32)).to(dtype)
            self.error_map = None if error_map is None else torch.as_tensor(np.asarray(error_map, dtype=np.float32)).to(dtype)
            self.background_rms = None if background_rms is None else float(np.float32(background_rms))
            self.exp_time = None if exp_time is None else float(np.float32(exp_time))
        if include_positions:  # :69-74
            f32 = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float32)).to(dtype)
            self.centroids_x = [f32(c) for c in centroids_x]
            self.centroids_y = [f32(c) for c in centroids_y]
            self.centroids_errors_x = [f32(c) for c in centroids_errors_x]
            self.centroids_errors_y = [f32(c) for c in centroids_errors_y]
            self.n_position = 2 * sum(int(c.numel()) for c in self.centroids_x)

    def init_centroids(self, bs):
        # :185-194  centroids repeated over the batch axis: (n_img,) -> (n_img, bs)
        if self.include_positions:
            self.centroids_x_batch = [c[:, None].repeat(1, bs) for c in self.centroids_x]
            self.centroids_y_batch = [c[:, None].repeat(1, bs) for c in self.centroids_y]

    def stats_positions(self, simulator, params):
        # :103-124
        chi2, log_like = 0.0, 0.0
        for cx, cy, cex, cey in zip(self.centroids_x_batch, self.centroids_y_batch, self.centroids_errors_x,
                                    self.centroids_errors_y):
            beta_centroids = torch.stack(simulator.beta(cx, cy, params["lens_mass"]), dim=0)  # (xy, images, bs)
            beta_centroids = beta_centroids.permute(2, 0, 1)  # batch size, xy, images
            beta_barycentre = beta_centroids.mean(dim=2, keepdim=True)
            magnifications = simulator.magnification(cx, cy, params["lens_mass"]).permute(1, 0)  # batch size, images
            err_map = torch.stack([cex / magnifications, cey / magnifications], dim=1)  # batch size, xy, images
            chi2_i = (((beta_centroids - beta_barycentre) / err_map) ** 2).sum((-2, -1))
            normalization_i = torch.log(2 * np.pi * err_map ** 2).sum((-2, -1))
            log_like = log_like + -1 / 2 * (chi2_i + normalization_i)
            chi2 = chi2 + chi2_i
        return log_like, chi2 / self.n_position

    def stats_pixels_from_image(self, im_sim, img_region):
        # tf/model.py:91-101
        if self.error_map is not None:
            err_map = self.error_map
        else:
            err_map = torch.sqrt(self.background_rms ** 2 + im_sim / self.exp_time)
        chi2 = (((im_sim - self.observed_image) / err_map) ** 2 * img_region).sum((-2, -1))
        normalization = (torch.log(2 * np.pi * err_map ** 2) * img_region).sum((-2, -1))
        log_like = -1 / 2 * (chi2 + normalization)
        red_chi2 = chi2 / torch.count_nonzero(img_region).to(im_sim.dtype)
        return log_like, red_chi2

    def stats_pixels(self, simulator, params):
        im_sim = simulator.simulate(params)
        if im_sim.dim() == 2:
            im_sim = im_sim[None]
        return self.stats_pixels_from_image(im_sim, simulator.img_region)

    def _stats(self, simulator, params):
        # :150-163
        log_like, red_chi2, n_chi = 0.0, 0.0, 0
        if self.include_pixels:
            ll, rc = self.stats_pixels(simulator, params)
            log_like, red_chi2, n_chi = log_like + ll, red_chi2 + rc, n_chi + 1
        if self.include_positions:
            ll, rc = self.stats_positions(simulator, params)
            log_like, red_chi2, n_chi = log_like + ll, red_chi2 + rc, n_chi + 1
        return log_like, red_chi2 / n_chi

    def log_prob(self, simulator, z):
        params, _ = self.prior.forward(z)
        log_like, red_chi2 = self._stats(simulator, params)
        return log_like + self.prior.log_prior(z), red_chi2

    def log_like(self, simulator, z):
        params, _ = self.prior.forward(z)
        return self._stats(simulator, params)[0]


class BackwardProbModel:
    """``tf/model.py:197-273``."""

    def __init__(self, prior: JointPrior, observed_image, background_rms, exp_time, dtype=torch.float32):
        self.prior = prior
        self.dtype = dtype
        obs = torch.as_tensor(np.asarray(observed_image, dtype=np.float32)).to(dtype)
        self.observed_image = obs
        self.err_map = torch.sqrt(float(background_rms) ** 2 + torch.clamp(obs, min=0) / float(exp_time))  # :221-223

    def log_prob(self, simulator, z):
        params, _ = self.prior.forward(z)
        im_sim = simulator.lstsq_simulate(params, self.observed_image, self.err_map)
        if im_sim.dim() == 2:
            im_sim = im_sim[None]
        r = (im_sim - self.observed_image) / self.err_map
        # Independent(Normal(obs, err)).log_prob(im_sim), reinterpreted over both pixel axes
        log_like = (-0.5 * r ** 2 - 0.5 * _LOG_2PI - torch.log(self.err_map)).sum((-2, -1))
        return log_like + self.prior.log_prior(z), (r ** 2).mean((-2, -1))
