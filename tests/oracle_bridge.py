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
    pm = OM.ForwardProbModel(prior, wl["observed"], background_rms=wl.get("background_rms"), exp_time=wl.get("exp_time"),
                             error_map=wl.get("error_map"), dtype=dtype)
    return sim, pm


def logprob_and_grad(wl, z, dtype=torch.float32):
    """(logp[bs], red_chi2[bs], dlogp/dz[bs][d]) from the oracle with autograd."""
    z = np.asarray(z)
    sim, pm = build_oracle(wl, z.shape[0], dtype)
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
