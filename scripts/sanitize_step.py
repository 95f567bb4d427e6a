"""One small log-prob + gradient step of C2, C3 (lstsq) and C4 (taped cluster path) at bs = 8 and the image-position term: the
workload compute-sanitizer runs over (memcheck / racecheck / synccheck, one tool per GPU call; summaries under profiles/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gigalens_b200 import workloads
from gigalens_b200.model import BackwardProbModel, ForwardProbModel
from gigalens_b200.simulator import LensSimulator

bs = 8
wl = workloads.c2_workload()
sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
pm = ForwardProbModel(wl["prior"], wl["observed"], background_rms=0.2, exp_time=100.0,
                      centroids_x=[np.array([1.1, -1.1, 0.05, -0.05], np.float32)], centroids_y=[np.array([0.05, -0.05, 1.1, -1.1], np.float32)],
                      centroids_errors_x=[np.full(4, 0.05, np.float32)], centroids_errors_y=[np.full(4, 0.05, np.float32)])
z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(bs, seed=0)), device="cuda")
out = pm.log_prob_and_grad(sim, z)
print("C2 + positions", float(out[0][0]), bool(torch.isfinite(out[2]).all()))
for opt in ("conv_tma", "straight_line", "packed_math"):
    sim.set_option(opt, 0)
    out = pm.log_prob_and_grad(sim, z)
    print("C2", opt, "= 0", float(out[0][0]))
    sim.set_option(opt, 1)
wl = workloads.c3_workload(n_max=6)
sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
pm = BackwardProbModel(wl["prior"], wl["observed"], wl["background_rms"], wl["exp_time"])
z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(bs, seed=0)), device="cuda")
out = pm.log_prob_and_grad(sim, z)
print("C3 (n_max 6, tcgen05 Gram)", float(out[0][0]), bool(torch.isfinite(out[2]).all()))
wl = workloads.c4_workload(num_pix=48, observed=np.ones((48, 48), np.float32))
sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
pm = ForwardProbModel(wl["prior"], wl["observed"], background_rms=wl["background_rms"], exp_time=wl["exp_time"])
z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(bs, seed=0)), device="cuda")
out = pm.log_prob_and_grad(sim, z)
print("C4 48x48 (tape)", float(out[0][0]), bool(torch.isfinite(out[2]).all()))
sim.set_option("tape", 0)
out = pm.log_prob_and_grad(sim, z)
print("C4 48x48 (recompute)", float(out[0][0]))
torch.cuda.synchronize()
print("done")
