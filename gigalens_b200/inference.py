"""``ModellingSequence``: MAP / SVI / HMC drivers with the reference's signatures
(``src/gigalens/inference.py:10-139``, ``src/gigalens/tf/inference.py:17-182``), re-hosted on torch.

Every driver is a thin caller of the hot path ``prob_model.log_prob_and_grad(simulator, z)`` (one
fused CUDA forward + hand-adjoint pass); the optimiser / sampler state is a handful of small
``(bs, d)`` tensors.  Multi-GPU: one process per GPU (torchrun), each rank owns a contiguous shard of
the samples / chains; NCCL carries only the small reductions the algorithms need:

* MAP  - none (per-sample Adam); ``gather=True`` all-gathers the final ``z`` and log-probs.
* SVI  - one all-reduce(sum) per step of ``[ELBO, grad_mu (d), grad_L (d(d+1)/2)]``
         (``src/gigalens/jax/inference.py:126-128``).
* HMC  - per step all-reduce of the acceptance statistic (dual averaging) and the ChEES moments,
         so that adaptation equals the single-device run.
* SMC  - per stage one all-gather of the particles and their cached terms (resampling mixes particles
         across ranks) and all-reduces of the tempering / tuning statistics.
"""
import math

import numpy as np

from .simulator import LensSimulator


class PolynomialDecay:
    """``tf.keras.optimizers.schedules.PolynomialDecay`` (used by the reference notebooks)."""

    def __init__(self, initial_learning_rate, decay_steps, end_learning_rate=0.0001, power=1.0):
        self.lr0, self.steps, self.lr1, self.power = initial_learning_rate, decay_steps, end_learning_rate, power

    def __call__(self, step):
        s = min(step, self.steps)
        return (self.lr0 - self.lr1) * (1 - s / self.steps) ** self.power + self.lr1


class Adam:
    """Minimal Adam on one tensor (Keras defaults: beta1 0.9, beta2 0.999, eps 1e-7); ``lr`` may be a
    float or a schedule ``step -> lr``.  Element-wise, so sharding the batch changes nothing."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.t, self.m, self.v = 0, None, None

    def step(self, x, grad, grad_scale=1.0, scrub_nan=False):
        """``x`` is updated in place with the gradient ``grad * grad_scale`` (non-finite entries zeroed when ``scrub_nan``).
        CUDA float32 tensors take the fused ``gl_adam_step`` kernel (one launch instead of eight); anything else the torch
        formulation of the same update."""
        import torch

        if self.m is None:
            self.m, self.v = torch.zeros_like(x), torch.zeros_like(x)
        lr = self.lr(self.t) if callable(self.lr) else self.lr
        self.t += 1
        if (x.is_cuda and x.dtype == torch.float32 and grad.dtype == torch.float32 and x.is_contiguous() and grad.is_contiguous()
                and grad.device == x.device):
            import ctypes as C

            from . import _cabi

            lib = _cabi.load()
            alpha = lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
            with torch.cuda.device(x.device):
                _cabi.check(lib.gl_adam_step(x.data_ptr(), grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), x.numel(),
                                             float(grad_scale), float(self.b1), float(self.b2), float(alpha), float(self.eps),
                                             int(bool(scrub_nan)), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), lib)
            return x
        if grad_scale != 1.0:
            grad = grad * grad_scale
        if scrub_nan:
            grad = torch.nan_to_num(grad, nan=0.0, posinf=0.0, neginf=0.0)
        self.m.mul_(self.b1).add_(grad, alpha=1 - self.b1)
        self.v.mul_(self.b2).addcmul_(grad, grad, value=1 - self.b2)
        alpha = lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        x.addcdiv_(self.m, self.v.sqrt().add_(self.eps), value=-alpha)
        return x


def _dist():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def _shard(n, rank, world):
    """Contiguous shard of n items for this rank."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class SurrogateMVN:
    """``tfd.MultivariateNormalTriL(loc, FillScaleTriL(diag_bijector=Exp, diag_shift=1e-6))``
    (``tf/inference.py:65-74``) or the diagonal variant (``:76-83``)."""

    def __init__(self, loc, scale_tril):
        self.loc, self.scale_tril = loc, scale_tril

    def mean(self):
        return self.loc

    def covariance(self):
        return self.scale_tril @ self.scale_tril.T

    def sample(self, n, generator=None):
        import torch

        eps = torch.randn((n, self.loc.shape[0]), device=self.loc.device, dtype=self.loc.dtype, generator=generator)
        return self.loc + eps @ self.scale_tril.T


class ModellingSequence:
    """``src/gigalens/tf/inference.py:17`` on the CUDA hot path."""

    def __init__(self, phys_model, prob_model, sim_config, simulator_cls=LensSimulator):
        self.phys_model = phys_model
        self.prob_model = prob_model
        self.sim_config = sim_config
        self._simulator_cls = simulator_cls   # injectable so the sharding / collective logic is testable on CPU
        self._coll_events = None              # measurement aid: CUDA event pairs around every all-reduce (bench.py)
        self._sim = None                      # the most recent simulator (one plan = tens of GB at the cluster config: reused
                                              # by the next driver call with the same batch size instead of being re-allocated)

    def _simulator(self, bs):
        """A simulator for `bs` samples: the previous call's when the batch size matches (plans are immutable), else a new one
        (the old plan is released first, so two workspaces never coexist)."""
        if self._sim is not None and getattr(self._sim, "bs", None) == bs:
            return self._sim
        self._sim = None
        self._sim = self._simulator_cls(self.phys_model, self.sim_config, bs=bs)
        return self._sim

    def time_collectives(self, on=True):
        """Record CUDA events on the current stream around every all-reduce of SVI / HMC from now on."""
        self._coll_events = [] if on else None

    def collective_ms(self):
        """Device time (ms) spent in (waiting for) the all-reduces recorded so far and their count; synchronises."""
        import torch

        if not self._coll_events:
            return 0.0, 0
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in self._coll_events), len(self._coll_events)

    def _all_reduce(self, dist, t):
        if self._coll_events is None or not t.is_cuda:
            dist.all_reduce(t)
            return t
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_reduce(t)
        e1.record()
        self._coll_events.append((e0, e1))
        return t

    # ------------------------------------------------------------------ MAP
    def MAP(self, optimizer=None, start=None, n_samples=500, num_steps=350, seed=0, gather=True, callback=None,
            scrub_nan_gradients=False):
        """``tf/inference.py:18-45``: per-sample Adam ascent of ``log_prob / event_size``.

        ``optimizer``: an :class:`Adam` (or anything with ``step(z, grad)``); default ``Adam(1e-2)``.
        ``start``: physical-parameter pytree with ``n_samples`` leaves, default: prior draws.
        ``scrub_nan_gradients``: the reference applies gradients as they come, so a sample whose gradient turns NaN
        stays NaN (Adam is element-wise: its neighbours are unaffected) and drops out through ``np.nanmin``; ``True``
        zeroes non-finite gradients instead and keeps such a sample where it is.
        Returns the final unconstrained ``z`` ``(n_samples, d)`` (all samples when ``gather``)."""
        import torch

        dist, rank, world = _dist()
        pm = self.prob_model
        optimizer = optimizer or Adam(1e-2)
        start = pm.prior.sample(n_samples, seed=seed) if start is None else start
        z_all = pm.bij_inverse(start)
        lo, hi = _shard(n_samples, rank, world)
        sim = self._simulator(hi - lo)
        z = torch.as_tensor(z_all[lo:hi], device=sim.device).clone()
        # tf/inference.py:27-31: pixels count when include_pixels, image positions add n_position
        event_size = 0.0
        if getattr(pm, "include_pixels", True):
            event_size += float(torch.count_nonzero(sim.img_region))
        if getattr(pm, "include_positions", False):
            event_size += float(pm.n_position)
        self.last_red_chi2 = None
        for step in range(num_steps):
            logp, red_chi2, dz = pm.log_prob_and_grad(sim, z)
            # d mean(-logp / event_size) / dz; scale, NaN scrub and the Adam update are one kernel on the GPU
            if isinstance(optimizer, Adam):
                optimizer.step(z, dz, grad_scale=-1.0 / (event_size * n_samples), scrub_nan=scrub_nan_gradients)
            else:
                grad = dz.mul_(-1.0 / (event_size * n_samples))
                if scrub_nan_gradients:
                    grad = torch.nan_to_num_(grad, nan=0.0, posinf=0.0, neginf=0.0)
                optimizer.step(z, grad)
            if callback is not None:
                callback(step, red_chi2)
        self.last_red_chi2 = red_chi2
        if gather and world > 1:
            sizes = [_shard(n_samples, r, world)[1] - _shard(n_samples, r, world)[0] for r in range(world)]
            pad = torch.zeros((max(sizes), z.shape[1]), device=z.device)
            pad[: z.shape[0]] = z
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad)     # equal-sized blocks; trim the padding of the short shards
            z = torch.cat([q[:n] for q, n in zip(parts, sizes)], 0)
        return z

    # ------------------------------------------------------------------ SVI
    def SVI(self, optimizer=None, start_mean=None, n_vi=250, init_scales=1e-3, num_steps=500, seed=2, full_rank=True,
            start=None):
        """``tf/inference.py:47-93``: fit a (full-rank) Gaussian surrogate by maximising the ELBO with
        reparameterised samples; returns ``(q_z, losses)``.  ``start`` is accepted as an alias of ``start_mean``
        (the reference's own ``tests/tf/test_model.py:53`` calls it that way)."""
        import torch

        if start_mean is None:
            start_mean = start
        if start_mean is None:
            raise TypeError("SVI needs start_mean")

        dist, rank, world = _dist()
        pm = self.prob_model
        optimizer = optimizer or Adam(1e-3)
        lo, hi = _shard(n_vi, rank, world)
        sim = self._simulator(hi - lo)
        dev = sim.device
        mu = torch.as_tensor(np.asarray(start_mean.detach().cpu() if torch.is_tensor(start_mean) else start_mean),
                             dtype=torch.float32, device=dev).reshape(-1).clone()
        d = mu.numel()
        scale = torch.eye(d, device=dev) * init_scales if np.size(init_scales) == 1 \
            else torch.as_tensor(np.asarray(init_scales), dtype=torch.float32, device=dev)
        tril_idx = torch.tril_indices(d, d, device=dev)
        diag_mask = tril_idx[0] == tril_idx[1]
        # unconstrained parameters of FillScaleTriL(Exp, shift 1e-6): log(diag - 1e-6), off-diagonals as is
        if full_rank:
            raw = scale[tril_idx[0], tril_idx[1]].clone()
            raw[diag_mask] = torch.log(raw[diag_mask] - 1e-6)
        else:
            raw = torch.log(torch.diagonal(scale).clone())
        theta = torch.cat([mu, raw]).requires_grad_(True)
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed * 1000003 + rank)
        loss_buf = torch.zeros(num_steps, device=dev)   # the ELBO trace stays on the device: no host sync per step

        def build(th):
            m, r = th[:d], th[d:]
            if full_rank:
                L = torch.zeros((d, d), device=dev)
                vals = torch.where(diag_mask, torch.exp(r) + 1e-6, r)
                L = L.index_put((tril_idx[0], tril_idx[1]), vals)
            else:
                L = torch.diag(torch.exp(r))
            return m, L

        for step in range(num_steps):
            m, L = build(theta)
            eps = torch.randn((hi - lo, d), device=dev, generator=gen)
            z = m + eps @ L.T
            logp, _, dz = pm.log_prob_and_grad(sim, z.detach())
            # loss = mean_i[log q(z_i) - log p(z_i)] over all n_vi samples (tfp.vi.fit_surrogate_posterior);
            # log q(z(eps)) = -|eps|^2/2 - sum log L_ii - d/2 log 2pi.  The hot-path gradient enters
            # through the surrogate term <stop_grad(dlogp/dz), z>.
            logq = -0.5 * (eps ** 2).sum(1) - torch.log(torch.diagonal(L)).sum() - 0.5 * d * math.log(2 * math.pi)
            good = torch.isfinite(logp) & torch.isfinite(dz).all(1)    # drop samples whose value or gradient is not finite
            dz = torch.where(good[:, None], dz, torch.zeros_like(dz))
            surrogate = (logq.sum() - (dz * z).sum()) / n_vi
            (g,) = torch.autograd.grad(surrogate, theta)
            loss = (logq - torch.where(good, logp, torch.zeros_like(logp))).sum() / n_vi
            packed = torch.cat([loss.detach().reshape(1), g])
            if world > 1:
                self._all_reduce(dist, packed)   # the one collective of SVI: [ELBO, grad_mu, grad_L]
            loss_buf[step] = packed[0]
            with torch.no_grad():
                optimizer.step(theta, torch.nan_to_num(packed[1:], nan=0.0, posinf=0.0, neginf=0.0))
        m, L = build(theta.detach())
        return SurrogateMVN(m, L), loss_buf.tolist()

    # ------------------------------------------------------------------ HMC
    def HMC(self, q_z, init_eps=0.3, init_l=3, n_hmc=50, num_burnin_steps=250, num_results=750, max_leapfrog_steps=30,
            adapt_rate=0.05, adapt_mode="dual", seed=3, target_accept=0.75):
        """``tf/inference.py:95-182``: HMC preconditioned with the SVI covariance (momentum precision =
        Sigma, ``:131-138``), dual-averaging (or simple) step-size adaptation and ChEES trajectory-length
        adaptation during the first 80% of burn-in (``:140-158``).  Returns ``(samples, stats)`` with
        samples ``(num_results, n_hmc, d)`` (this rank's chains)."""
        import torch

        if adapt_mode not in ("dual", "simple"):
            raise ValueError(f"Invalid adaptation mode {adapt_mode}, the options are 'simple' and 'dual'")
        dist, rank, world = _dist()
        pm = self.prob_model
        lo, hi = _shard(n_hmc, rank, world)
        nloc = hi - lo
        sim = self._simulator(nloc)
        dev = sim.device
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed * 1000003 + rank)
        cov = q_z.covariance().to(dev)
        if not bool(torch.isfinite(cov).all()):
            raise FloatingPointError("HMC: the surrogate covariance is not finite (SVI diverged)")
        cov = 0.5 * (cov + cov.T)
        Lc, info = torch.linalg.cholesky_ex(cov)
        if int(info) != 0:   # fp32 round-off on a nearly singular surrogate: add a relative jitter
            cov = cov + 1e-6 * torch.diag(torch.diagonal(cov))
            Lc = torch.linalg.cholesky(cov)
        d = cov.shape[0]
        z = q_z.sample(nloc, generator=gen).to(dev)
        logp, _, grad = pm.log_prob_and_grad(sim, z)
        logp, grad = logp.clone(), grad.clone()

        def allsum(t):
            if world > 1:
                self._all_reduce(dist, t)
            return t

        n_adapt = int(num_burnin_steps * 0.8)
        log_eps, log_eps_bar, h_bar, mu_da = math.log(init_eps), 0.0, 0.0, math.log(10 * init_eps)
        gamma_da, t0_da, kappa_da = 0.05, 10.0, 0.75     # tfp.mcmc.DualAveragingStepSizeAdaptation defaults
        # ChEES adapts log T (T = trajectory length) with a scalar Adam (lr 0.025) on the host: the adaptation state is a
        # handful of floats, fed by ONE packed device->host read per adaptation step; the sampling phase reads nothing.
        log_T, log_T_avg = math.log(init_eps * init_l), math.log(init_eps * init_l)
        ad_m = ad_v = 0.0
        samples = torch.empty((num_results, nloc, d), device=dev)
        n_total = num_burnin_steps + num_results
        acc_trace = torch.zeros(n_total, device=dev)
        stats = {"accept_prob": [], "step_size": [], "num_leapfrog": []}
        n_evals = 0
        nloc_t = torch.tensor([float(nloc)], device=dev)
        for it in range(n_total):
            eps = math.exp(log_eps)
            # jittered trajectory (Halton sequence, shared by all chains / ranks); TFP keeps the jitter for the sampling
            # phase and applies it to the averaged trajectory length there
            u = (it * 0.6180339887498949) % 1.0
            T_now = math.exp(log_T if it < n_adapt else log_T_avg)
            n_leap = int(max(1, min(max_leapfrog_steps, math.ceil(2.0 * u * T_now / eps))))
            xi = torch.randn((nloc, d), device=dev, generator=gen)
            p0 = torch.linalg.solve_triangular(Lc.T, xi.T, upper=True).T          # p ~ N(0, Sigma^-1)
            zc, pc, gc, lpc = z.clone(), p0.clone(), grad, logp
            pc = pc + 0.5 * eps * gc
            for leap in range(n_leap):
                zc = zc + eps * (pc @ cov)                                           # dz/dt = Sigma p
                lpc, _, gc = pm.log_prob_and_grad(sim, zc)
                n_evals += nloc
                pc = pc + (eps if leap < n_leap - 1 else 0.5 * eps) * gc
            k0 = 0.5 * ((p0 @ cov) * p0).sum(1)
            k1 = 0.5 * ((pc @ cov) * pc).sum(1)
            log_acc = (lpc - k1) - (logp - k0)
            log_acc = torch.where(torch.isfinite(log_acc), log_acc, torch.full_like(log_acc, -float("inf")))
            acc_prob = torch.exp(torch.clamp(log_acc, max=0.0))
            accept = torch.log(torch.rand(nloc, device=dev, generator=gen)) < log_acc
            if it < n_adapt:
                # ChEES (tfp GradientBasedTrajectoryLengthAdaptation): moments over ALL chains
                tot = allsum(torch.cat([zc.sum(0), z.sum(0), nloc_t]))
                mean_prop, mean_prev = tot[:d] / tot[-1], tot[d:2 * d] / tot[-1]
                xc, yc = zc - mean_prop, z - mean_prev
                v = pc @ cov
                dsq = (xc ** 2).sum(1) - (yc ** 2).sum(1)
                gi = 2.0 * u * dsq * (xc * v).sum(1)
                gi = torch.where(torch.isfinite(gi), gi, torch.zeros_like(gi))
                num, den, cnt = allsum(torch.cat([(acc_prob * gi).sum().reshape(1), acc_prob.sum().reshape(1), nloc_t])).tolist()
                acc_trace[it] = den / cnt
                g_T = num / max(den, 1e-20)
                # Adam ascent on log T (d crit / d log T = T * d crit / dT), normalised like TFP by T^2
                g = -(g_T / math.exp(log_T))
                t_ad = it + 1
                ad_m = 0.9 * ad_m + 0.1 * g
                ad_v = 0.999 * ad_v + 0.001 * g * g
                log_T -= 0.025 * math.sqrt(1 - 0.999 ** t_ad) / (1 - 0.9 ** t_ad) * ad_m / (math.sqrt(ad_v) + 1e-7)
                log_T = min(log_T, math.log(max_leapfrog_steps * eps))
                w_avg = t_ad ** -0.5                                                # averaged trajectory length used after adaptation
                log_T_avg = log_T if it == 0 else (1 - w_avg) * log_T_avg + w_avg * log_T
                mean_acc = den / cnt
                if adapt_mode == "dual":
                    t_da = it + 1
                    h_bar = (1 - 1 / (t_da + t0_da)) * h_bar + (target_accept - mean_acc) / (t_da + t0_da)
                    log_eps = mu_da - math.sqrt(t_da) / gamma_da * h_bar
                    eta = t_da ** (-kappa_da)
                    log_eps_bar = eta * log_eps + (1 - eta) * log_eps_bar
                else:
                    log_eps += math.log1p(adapt_rate) if mean_acc > target_accept else -math.log1p(adapt_rate)
                if it == n_adapt - 1 and adapt_mode == "dual":
                    log_eps = log_eps_bar
            else:
                acc_trace[it] = acc_prob.mean()     # this rank's chains; no host read in the sampling phase
            z = torch.where(accept[:, None], zc, z)
            logp = torch.where(accept, lpc, logp)
            grad = torch.where(accept[:, None], gc, grad)
            stats["step_size"].append(eps)
            stats["num_leapfrog"].append(n_leap)
            if it >= num_burnin_steps:
                samples[it - num_burnin_steps] = z
        stats["accept_prob"] = acc_trace.tolist()
        stats["n_evals"] = n_evals
        return samples, stats


    # ------------------------------------------------------------------ SMC
    def SMC(self, start=None, num_particles=1000, num_ensembles=1, num_leapfrog_steps=10, post_sampling_steps=100,
            ess_threshold_ratio=0.8, max_sampling_per_stage=8, target="pixels", auxiliar="positions", seed=1,
            max_stage=100, min_sampling_per_stage=1, optimal_accept=0.651):
        """``tf/inference.py:184-302``: adaptively tempered sequential Monte Carlo from the prior to the
        posterior with an auxiliary likelihood, followed by ``post_sampling_steps`` HMC steps.

        The stage target is ``prior + aux + beta (like - aux)`` (``make_tempered_target_log_prob_fn_with_auxiliar``,
        ``:289-302``): with ``auxiliar='positions'`` the particles are first pulled to models that reproduce the image
        positions and the pixel likelihood is switched on gradually.  The outer algorithm restates
        ``tfp.experimental.mcmc.sample_sequential_monte_carlo`` (tensorflow-probability >= 0.19, not vendored in the
        reference): per ensemble, the next inverse temperature is found by bisection so that the effective sample size
        of the incremental weights ``exp((beta' - beta) like)`` equals ``ess_threshold_ratio N`` (the reference call
        hard-codes 0.8 and ``max_num_steps=8`` whatever its own arguments say -- those are the defaults here);
        systematic resampling; HMC mutation with ``num_leapfrog_steps`` and per-particle step size
        ``scaling_i * std(z) / num_leapfrog_steps`` (``gen_make_hmc_kernel_fn``); ``simple_heuristic_tuning`` of the
        scalings and of the number of mutation steps (``optimal_accept=0.651``).  Like TFP, the incremental weights use
        the target likelihood only (the auxiliary term enters the mutation target, not the weights).

        Returns ``(samples, info)``: samples ``(post_sampling_steps, n_local, d)`` of this rank's particles, or the
        final particle cloud ``(num_particles, num_ensembles, d)`` when ``post_sampling_steps == 0``."""
        import torch

        dist, rank, world = _dist()
        pm = self.prob_model
        E, P_ = int(num_ensembles), int(num_particles)
        N = E * P_
        if start is None:
            z_all = np.asarray(pm.bij_inverse(pm.prior.sample(N, seed=seed)), dtype=np.float32)
        else:
            flat = np.asarray(start.cpu() if hasattr(start, "cpu") else start, dtype=np.float32)
            flat = flat.reshape(-1, flat.shape[-1])
            z_all = flat[np.random.default_rng(seed).integers(0, flat.shape[0], size=N)]   # tf.random.categorical(:201-203)
        d = z_all.shape[1]
        lo, hi = _shard(N, rank, world)
        nloc = hi - lo
        sim = self._simulator(nloc)
        dev = sim.device
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed * 1000003 + rank)
        shared = torch.Generator(device="cpu")
        shared.manual_seed(seed)                                   # resampling offsets: identical on every rank
        ens_all = torch.arange(N, device=dev) % E                  # particle n belongs to ensemble n % E (reshape (P, E))
        sizes = [_shard(N, r, world)[1] - _shard(N, r, world)[0] for r in range(world)]

        def gather(t):
            if world == 1:
                return t
            pad = torch.zeros((max(sizes),) + tuple(t.shape[1:]), device=dev, dtype=t.dtype)
            pad[: t.shape[0]] = t
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad)
            return torch.cat([q[:n] for q, n in zip(parts, sizes)], 0)

        def terms(zz):
            lp, glp = pm.log_prior_and_grad(sim, zz)
            ll, gll = pm.term_and_grad(sim, zz, target)
            la, gla = pm.term_and_grad(sim, zz, auxiliar)
            scrub = lambda v, g: (torch.where(torch.isfinite(v), v, torch.full_like(v, -float("inf"))),
                                  torch.nan_to_num(g, nan=0.0, posinf=0.0, neginf=0.0))
            (lp, glp), (ll, gll), (la, gla) = scrub(lp, glp), scrub(ll, gll), scrub(la, gla)
            return torch.stack([lp, ll, la], 1).clone(), torch.stack([glp, gll, gla], 1).clone()   # (n, 3), (n, 3, d)

        def tempered(val, grd, b):           # b: (n,) inverse temperature of each particle's ensemble
            v = val[:, 0] + val[:, 2] + b * (val[:, 1] - val[:, 2])
            g = grd[:, 0] + grd[:, 2] + b[:, None] * (grd[:, 1] - grd[:, 2])
            return torch.where(torch.isnan(v), torch.full_like(v, -float("inf")), v), g

        def hmc_steps(z, val, grd, b, eps, n_steps):
            """n_steps HMC transitions (identity mass) of the local particles; returns the log mean accept prob."""
            acc_sum = torch.zeros(z.shape[0], device=dev)
            n_eval = 0
            for _ in range(n_steps):
                lp0, g0 = tempered(val, grd, b)
                p0 = torch.randn(z.shape, device=dev, generator=gen)
                zc, pc = z.clone(), p0 + 0.5 * eps * g0
                for leap in range(num_leapfrog_steps):
                    zc = zc + eps * pc
                    vc, gc = terms(zc)
                    n_eval += z.shape[0]
                    lpc, gcur = tempered(vc, gc, b)
                    pc = pc + (eps if leap < num_leapfrog_steps - 1 else 0.5 * eps) * gcur
                log_acc = (lpc - 0.5 * (pc ** 2).sum(1)) - (lp0 - 0.5 * (p0 ** 2).sum(1))
                log_acc = torch.where(torch.isfinite(log_acc), log_acc, torch.full_like(log_acc, -float("inf")))
                accept = torch.log(torch.rand(z.shape[0], device=dev, generator=gen)) < log_acc
                z = torch.where(accept[:, None], zc, z)
                val = torch.where(accept[:, None], vc, val)
                grd = torch.where(accept[:, None, None], gc, grd)
                acc_sum += torch.exp(torch.clamp(log_acc, max=0.0))
            return z, val, grd, torch.log(torch.clamp(acc_sum / max(n_steps, 1), min=1e-30)), n_eval

        z = torch.as_tensor(z_all[lo:hi], device=dev).clone()
        val, grd = terms(z)
        beta = torch.zeros(E, device=dev)
        log_scal = torch.full((nloc,), math.log(min(1.0, 2.38 ** 2 / d)), device=dev)
        n_steps, stage, n_evals = int(max_sampling_per_stage), 0, nloc
        log_evidence = torch.zeros(E, device=dev)
        info = {"inverse_temperature": [], "num_steps": [], "accept_prob": []}
        log_acc_loc = None
        while bool((beta < 1).any()) and stage < max_stage:
            # ---- all particles on every rank (small: N x (1 + 3 + 3d + d) floats)
            Z, V, G, LS = gather(z), gather(val), gather(grd), gather(log_scal)
            ll = V[:, 1].double()
            new_beta = beta.clone()
            idx_all = torch.arange(N, device=dev)
            for e in range(E):
                m = ens_all == e
                le = ll[m]
                le = torch.where(torch.isfinite(le), le, torch.full_like(le, -1e300))

                def ess_ratio(delta):
                    lw = delta * le
                    lw = lw - torch.logsumexp(lw, 0)
                    return float(torch.exp(-torch.logsumexp(2 * lw, 0))) / le.numel()

                b0, room = float(beta[e]), 1.0 - float(beta[e])
                if room <= 0:
                    continue
                if ess_ratio(room) >= ess_threshold_ratio:
                    delta = room
                else:
                    a_, c_ = 0.0, room
                    for _ in range(50):
                        mid = 0.5 * (a_ + c_)
                        if ess_ratio(mid) >= ess_threshold_ratio:
                            a_ = mid
                        else:
                            c_ = mid
                    delta = max(a_, 1e-12)
                new_beta[e] = min(1.0, b0 + delta)
                lw = delta * le
                log_evidence[e] += float(torch.logsumexp(lw, 0) - math.log(le.numel()))
                w = torch.softmax(lw, 0)
                # systematic resampling (tfe.mcmc.resample_systematic): one uniform offset per ensemble
                n_e = le.numel()
                u0 = float(torch.rand(1, generator=shared))
                pts = (torch.arange(n_e, device=dev, dtype=torch.float64) + u0) / n_e
                pick = torch.searchsorted(torch.cumsum(w, 0), pts).clamp_(max=n_e - 1)
                idx_all[m] = idx_all[m][pick]
            beta = new_beta
            sel = idx_all[lo:hi]
            z, val, grd, log_scal = Z[sel].clone(), V[sel].clone(), G[sel].clone(), LS[sel].clone()
            # ---- tuning (simple_heuristic_tuning) from the previous stage's acceptance
            if log_acc_loc is not None:
                LA = gather(log_acc_loc)[sel]
                tot = torch.zeros(2 * E + 1, device=dev)
                ens_loc = ens_all[lo:hi]
                tot[:E].index_add_(0, ens_loc, torch.exp(log_scal))
                tot[E:2 * E].index_add_(0, ens_loc, torch.exp(LA))
                tot[2 * E] = float(nloc)
                if world > 1:
                    dist.all_reduce(tot)
                per_e = N / E
                avg_scal, avg_acc = tot[:E] / per_e, tot[E:2 * E] / per_e
                log_scal = math.log(0.5) + torch.logaddexp(torch.log(avg_scal[ens_loc]) + (avg_acc[ens_loc] - optimal_accept),
                                                           log_scal + (torch.exp(LA) - optimal_accept))
                a = float(avg_acc.mean().clamp(1e-6, 1 - 1e-6))
                n_steps = int(min(max_sampling_per_stage, max(min_sampling_per_stage, math.ceil(math.log1p(-0.99) / math.log1p(-a)))))
            # ---- mutation: step size = scaling * std(z) / L, std over ALL particles of the ensemble
            Zs = gather(z)
            std = torch.stack([Zs[ens_all == e].std(0, unbiased=False) for e in range(E)], 0)      # (E, d)
            ens_loc = ens_all[lo:hi]
            eps = torch.exp(log_scal)[:, None] * std[ens_loc] / float(num_leapfrog_steps)
            z, val, grd, log_acc_loc, ne = hmc_steps(z, val, grd, beta[ens_loc], eps, n_steps)
            n_evals += ne
            stage += 1
            info["inverse_temperature"].append(beta.cpu().tolist())
            info["num_steps"].append(n_steps)
            info["accept_prob"].append(float(torch.exp(log_acc_loc).mean()))
        info.update(stages=stage, log_evidence=log_evidence.cpu().tolist())
        if post_sampling_steps <= 0:
            info["n_evals"] = n_evals
            return gather(z).reshape(P_, E, d), info
        # ---- HMC on the full posterior prior + like with the final step sizes (:262-286)
        Zs = gather(z)
        std = torch.stack([Zs[ens_all == e].std(0, unbiased=False) for e in range(E)], 0)
        ens_loc = ens_all[lo:hi]
        eps = torch.exp(log_scal)[:, None] * std[ens_loc] / float(num_leapfrog_steps)
        ones = torch.ones(nloc, device=dev)
        samples = torch.empty((post_sampling_steps, nloc, d), device=dev)
        for it in range(post_sampling_steps):
            z, val, grd, la_, ne = hmc_steps(z, val, grd, ones, eps, 1)
            n_evals += ne
            samples[it] = z
        info["n_evals"] = n_evals
        return samples, info
