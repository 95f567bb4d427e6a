"""Bridge from product-side workload objects to the CPU oracle (TEST INFRASTRUCTURE).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference legs
import this module."""
import numpy as np
import torch

from gigalens_b200 import distributions as tfd
from oracle import model as OM
from oracle.simulator import OracleSimulator

import common


def to_oracle_prior(prior):
    def conv(s):
        if isinstance(s, tfd.JointDistribution):
            return conv(s.model)
        if isinstance(s, dict):
            return {k: conv(v) for k, v in s.items()}
        if isinstance(s, (list, tuple)):
            return [conv(v) for v in s]
        if isinstance(s, tfd.TruncatedNormal):
            return OM.TruncatedNormal(s.loc, s.scale, s.low, s.high)
        if isinstance(s, tfd.Uniform):
            return OM.Uniform(s.low, s.high)
        if isinstance(s, tfd.LogNormal):
            return OM.LogNormal(s.loc, s.scale)
        if isinstance(s, tfd.Normal):
            return OM.Normal(s.loc, s.scale)
        raise TypeError(type(s))

    return OM.JointPrior(conv(prior))


def build_oracle(wl, bs, dtype=torch.float32):
    sc = wl["sim_config"]
    om = common.to_oracle_model(wl["phys_model"], dtype)
    sim = OracleSimulator(om, sc.delta_pix, sc.num_pix, sc.supersample, kernel=sc.kernel,
                          transform_pix2angle=sc.transform_pix2angle, pix_region=sc.pix_region, bs=bs, dtype=dtype)
    prior = to_oracle_prior(wl["prior"])
    cen = wl.get("centroids")   # dict(x=[...], y=[...], ex=[...], ey=[...]): one array per multiply-imaged source
    pm = OM.ForwardProbModel(prior, wl.get("observed"), background_rms=wl.get("background_rms"), exp_time=wl.get("exp_time"),
                             error_map=wl.get("error_map"), dtype=dtype, include_pixels=wl.get("include_pixels", True),
                             include_positions=cen is not None,
                             **({} if cen is None else dict(centroids_x=cen["x"], centroids_y=cen["y"],
                                                            centroids_errors_x=cen["ex"], centroids_errors_y=cen["ey"])))
    pm.init_centroids(bs)
    return sim, pm


def find_images(wl, truth, beta_s, half_width, n_grid=48, newton=30):
    """Image positions of a point source at beta_s under the oracle's lens model at `truth` (fp64): coarse grid
    search for sign changes of beta - beta_s, Newton refinement with the oracle Hessian, duplicates merged."""
    osim, _ = build_oracle({**wl, "centroids": None, "include_pixels": True}, 1, torch.float64)
    lens = [{k: torch.as_tensor([float(v)], dtype=torch.float64) for k, v in d.items()} for d in truth["lens_mass"]]
    g = np.linspace(-half_width, half_width, n_grid)
    X, Y = np.meshgrid(g, g)
    xt = torch.as_tensor(X.reshape(-1, 1)); yt = torch.as_tensor(Y.reshape(-1, 1))
    bx, by = osim.beta(xt, yt, lens)
    d2 = ((bx[:, 0] - beta_s[0]) ** 2 + (by[:, 0] - beta_s[1]) ** 2).numpy().reshape(n_grid, n_grid)
    seeds = []
    for i in range(1, n_grid - 1):
        for j in range(1, n_grid - 1):
            if d2[i, j] == d2[i - 1:i + 2, j - 1:j + 2].min():
                seeds.append((X[i, j], Y[i, j]))
    out = []
    for sx, sy in seeds:
        th = np.array([sx, sy], dtype=np.float64)
        ok = False
        for _ in range(newton):
            x1 = torch.as_tensor([[th[0]]]); y1 = torch.as_tensor([[th[1]]])
            bx, by = osim.beta(x1, y1, lens)
            H = [float(h.detach()) for h in osim.hessian(x1, y1, lens)]
            r = np.array([float(bx) - beta_s[0], float(by) - beta_s[1]])
            A = np.array([[1 - H[0], -H[1]], [-H[2], 1 - H[3]]])
            step = np.linalg.solve(A, r)
            th = th - step
            if np.hypot(*step) < 1e-12:
                ok = True
                break
        if ok and np.all(np.abs(th) < half_width) and not any(np.hypot(*(th - o)) < 1e-6 for o in out):
            out.append(th)
    return np.array(out)


def logprob_and_grad(wl, z, dtype=torch.float32, beta_noise=None, noise_seed=0):
    """(logp[bs], red_chi2[bs], dlogp/dz[bs][d]) from the oracle with autograd.
    ``beta_noise`` (arcsec, rms): add fixed Gaussian noise of that size to every ray's source-plane position -- the size of one
    fp32 rounding of beta is ~1.2e-7 arcsec at |beta| ~ 2 arcsec.  Probes how far pixel-level rounding of the ray positions moves
    the result: a ray that lands next to the cusp of a high-n Sersic source carries d(light)/d(beta) ~ R^(1/n - 1)."""
    z = np.asarray(z)
    sim, pm = build_oracle(wl, z.shape[0], dtype)
    if beta_noise:
        exact = sim.beta
        gen = torch.Generator().manual_seed(noise_seed)

        def noisy_beta(x, y, lens_params):
            bx, by = exact(x, y, lens_params)
            return (bx + beta_noise * torch.randn(bx.shape, generator=gen, dtype=bx.dtype),
                    by + beta_noise * torch.randn(by.shape, generator=gen, dtype=by.dtype))

        sim.beta = noisy_beta
    zt = torch.as_tensor(z, dtype=dtype).clone().requires_grad_(True)
    logp, chi2 = pm.log_prob(sim, zt)
    logp.sum().backward()
    return logp.detach().numpy(), chi2.detach().numpy(), zt.grad.numpy()


def loglike_and_grad_matrix(wl, cm, mat, dtype=torch.float32):
    """log-like, red_chi2, image and d(log_like)/d(params [P][bs]) from the oracle."""
    bs = mat.shape[1]
    sim, pm = build_oracle(wl, bs, dtype)
    params, leaf = common.matrix_to_pytree(cm, mat, dtype, requires_grad=True)
    im = sim.simulate(params)
    if im.dim() == 2:
        im = im[None]
    ll, chi2 = pm.stats_pixels_from_image(im, sim.img_region)
    ll.sum().backward()
    return ll.detach().numpy(), chi2.detach().numpy(), im.detach().numpy(), leaf.grad.numpy()


def build_oracle_backward(wl, bs, dtype=torch.float32):
    sc = wl["sim_config"]
    om = common.to_oracle_model(wl["phys_model"], dtype)
    sim = OracleSimulator(om, sc.delta_pix, sc.num_pix, sc.supersample, kernel=sc.kernel,
                          transform_pix2angle=sc.transform_pix2angle, pix_region=sc.pix_region, bs=bs, dtype=dtype)
    pm = OM.BackwardProbModel(to_oracle_prior(wl["prior"]), wl["observed"], wl["background_rms"], wl["exp_time"], dtype=dtype)
    return sim, pm


def backward_logprob_and_grad(wl, z, dtype=torch.float32):
    """BackwardProbModel.log_prob (lstsq path) and its autograd gradient through torch.linalg.pinv."""
    z = np.asarray(z)
    sim, pm = build_oracle_backward(wl, z.shape[0], dtype)
    zt = torch.as_tensor(z, dtype=dtype).clone().requires_grad_(True)
    logp, chi2 = pm.log_prob(sim, zt)
    logp.sum().backward()
    return logp.detach().numpy(), chi2.detach().numpy(), zt.grad.numpy()
