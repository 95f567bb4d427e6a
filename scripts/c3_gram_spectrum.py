"""Spectrum of the C3 normal equations over a prior batch (development aid): how many samples have eigenvalues under the
pinv cut, how many fail the trace certificate, and how far the smallest eigenvalues sit from the cut."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gigalens_b200 import workloads
from gigalens_b200.model import BackwardProbModel
from gigalens_b200.simulator import LensSimulator
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
wl = workloads.c3_workload(observed=workloads.c3_observation())
sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
pm = BackwardProbModel(wl["prior"], wl["observed"], wl["background_rms"], wl["exp_time"])
z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(bs, seed=0)), device="cuda")
params = pm.bij_forward(sim, z)
st = sim.lstsq_simulate(params, wl["observed"], pm.err_map, return_stacked=True).double()      # (bs, n, n, D)
W = (1.0 / torch.as_tensor(pm.err_map, device="cuda", dtype=torch.float64))[None, :, :, None]
X = (st * W).reshape(bs, -1, st.shape[-1])
ev_all = []
for b0 in range(0, bs, 256):
    G = X[b0:b0 + 256].transpose(1, 2) @ X[b0:b0 + 256]
    ev_all.append(torch.linalg.eigvalsh(G).cpu().numpy())
    if b0 == 0:
        tr = torch.diagonal(G, dim1=1, dim2=2).sum(1); tri = torch.diagonal(torch.linalg.inv(G), dim1=1, dim2=2).sum(1)
        print("trace certificate fails (first 256):", int(((1 / tri) <= 1e-6 * tr).sum()))
ev = np.concatenate(ev_all)
ratio = ev[:, 0] / ev[:, -1]
print("lambda_min/lambda_max percentiles 0, 0.1, 1, 5, 50:", np.percentile(ratio, [0, 0.1, 1, 5, 50]))
nb = (ev <= 1e-6 * ev[:, -1:]).sum(1)
print("samples by number of eigenvalues under the cut:", np.bincount(nb))
idx = np.nonzero(nb)[0]
for b in idx[:12]:
    print(b, "smallest 4 / lmax:", ev[b, :4] / ev[b, -1])
print("samples with ratio in [1e-6, 1e-4]:", int(((ratio > 1e-6) & (ratio < 1e-4)).sum()))
