"""Writes tests/golden/c2_golden.npz: fp64 oracle outputs for the C2 workload on seeded inputs.

The reference itself cannot run here (no tensorflow / jax / lenstronomy), so these vectors come
from the oracle restatement, not from the reference: they pin the oracle against drift and travel
to the GPU box, where /root/reference does not exist."""
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
from gigalens_b200.simulator import CompiledModel  # noqa: E402

wl = workloads.c2_workload()
bs = 6
prior = wl["prior"]
z = ProbabilisticModel(prior).bij_inverse(prior.sample(bs, seed=1248)).astype(np.float32)
logp, chi, dz = oracle_bridge.logprob_and_grad(wl, z.astype(np.float64), torch.float64)
# the same inputs through the fp32 oracle and through fp64 with inputs moved by 1/2 ulp(fp32): the
# fp32 noise floor of these inputs (gradients of a 14400-pixel chi^2 cancel heavily in fp32)
logp32, chi32, dz32 = oracle_bridge.logprob_and_grad(wl, z, torch.float32)
logp_p, chi_p, dz_p = oracle_bridge.logprob_and_grad(wl, common.ulp_perturb(z), torch.float64)
sim, pm = oracle_bridge.build_oracle(wl, bs, torch.float64)
params, _ = pm.prior.forward(torch.as_tensor(z.astype(np.float64)))
img = sim.simulate(params).numpy()
cm = CompiledModel(wl["phys_model"])
mat = cm.flatten(workloads.DEMO_TRUTH, 1, torch, "cpu").numpy().astype(np.float64)
sim1, _ = oracle_bridge.build_oracle(wl, 1, torch.float64)
p1, _ = common.matrix_to_pytree(cm, mat, torch.float64)
truth_img = sim1.simulate(p1).numpy()
np.savez_compressed(os.path.join(HERE, "c2_golden.npz"), z=z, logp=logp, red_chi2=chi, dz=dz,
                    logp_fp32=logp32, red_chi2_fp32=chi32, dz_fp32=dz32, logp_pert=logp_p, dz_pert=dz_p,
                    image=img.astype(np.float32), truth_image=truth_img.astype(np.float32))
print("wrote c2_golden.npz", logp)
