"""CPU cost of one log_prob_and_grad call (tiny batch: the GPU work is negligible), development aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigalens_b200 import workloads
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator
wl = workloads.c2_workload(); n = 8
pm = ForwardProbModel(wl["prior"], wl["observed"], background_rms=0.2, exp_time=100.0)
sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=n)
z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(n, seed=0)), device="cuda")
for _ in range(20): pm.log_prob_and_grad(sim, z)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(500): pm.log_prob_and_grad(sim, z)
torch.cuda.synchronize(); print(f"bs={n}: {(time.perf_counter() - t0) / 500 * 1e6:.1f} us per call (wall, incl. tiny kernels)")
