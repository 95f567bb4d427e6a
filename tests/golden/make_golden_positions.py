"""Writes tests/golden/positions_golden.npz: fp64 oracle outputs of the image-position likelihood
(ForwardProbModel.stats_positions, tf/model.py:103-124) and of the lensing Hessian for the C2 lens model.

Like c2_golden.npz these come from the oracle restatement (the reference cannot run here); they pin the
oracle against drift and travel to the GPU box."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import common  # noqa: E402
import oracle_bridge  # noqa: E402
from gigalens_b200 import workloads  # noqa: E402
from gigalens_b200.model import ProbabilisticModel  # noqa: E402

wl = workloads.c2_workload()
cen = dict(x=[], y=[], ex=[], ey=[])
for k, beta_s in enumerate([(0.05, 0.03), (-0.12, 0.08)]):
    img = oracle_bridge.find_images(wl, workloads.DEMO_TRUTH, beta_s, 2.0)
    img = img + np.random.default_rng(40 + k).normal(0, 0.01, img.shape)
    cen["x"].append(img[:, 0].astype(np.float32)); cen["y"].append(img[:, 1].astype(np.float32))
    cen["ex"].append(np.full(len(img), 0.01, np.float32)); cen["ey"].append(np.full(len(img), 0.012, np.float32))
bs = 4
z0 = ProbabilisticModel(wl["prior"]).bij_inverse(workloads.DEMO_TRUTH)
z = (z0 + np.random.default_rng(3).normal(0, 0.01, size=(bs, z0.shape[1]))).astype(np.float32)
out = {}
for tag, inc_pix in (("pos", False), ("both", True)):
    w = dict(wl, centroids=cen, include_pixels=inc_pix)
    lp, chi, dz = oracle_bridge.logprob_and_grad(w, z.astype(np.float64), torch.float64)
    lp32, _, dz32 = oracle_bridge.logprob_and_grad(w, z, torch.float32)
    lpp, _, dzp = oracle_bridge.logprob_and_grad(w, common.ulp_perturb(z), torch.float64)
    out.update({f"logp_{tag}": lp, f"chi2_{tag}": chi, f"dz_{tag}": dz, f"logp32_{tag}": lp32, f"dz32_{tag}": dz32,
                f"logp_pert_{tag}": lpp, f"dz_pert_{tag}": dzp})
# Hessian of the summed deflection at the image positions, truth parameters
osim, _ = oracle_bridge.build_oracle(wl, 1, torch.float64)
lens = [{k: torch.as_tensor([float(v)], dtype=torch.float64) for k, v in d.items()} for d in workloads.DEMO_TRUTH["lens_mass"]]
X = torch.as_tensor(np.concatenate(cen["x"]).astype(np.float64))[:, None]
Y = torch.as_tensor(np.concatenate(cen["y"]).astype(np.float64))[:, None]
H = np.stack([h.detach().numpy()[:, 0] for h in osim.hessian(X, Y, lens)], 0)
mu = osim.magnification(X, Y, lens).detach().numpy()[:, 0]
np.savez_compressed(os.path.join(HERE, "positions_golden.npz"), z=z, n_img=np.asarray([len(c) for c in cen["x"]]),
                    cx=np.concatenate(cen["x"]), cy=np.concatenate(cen["y"]), ex=np.concatenate(cen["ex"]), ey=np.concatenate(cen["ey"]),
                    hessian_truth=H, magnification_truth=mu, **out)
print({k: np.asarray(v).shape for k, v in out.items()}, H.shape, mu)
