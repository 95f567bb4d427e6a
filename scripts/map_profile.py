"""Where a MAP step's time goes (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigalens_b200 import workloads
from gigalens_b200.inference import Adam
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator

wl = workloads.c2_workload(); n = 4096
pm = ForwardProbModel(wl["prior"], wl["observed"], background_rms=0.2, exp_time=100.0)
sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=n)
z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(n, seed=0)), device="cuda").clone()
opt = Adam(1e-2)
def timed(fn, it=200):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    return t / it * 1e3, t_cpu / it * 1e3
def a(): return pm.log_prob_and_grad(sim, z)
def b():
    logp, chi, dz = pm.log_prob_and_grad(sim, z)
    g = dz.mul_(-1.0 / (3600 * n)); return torch.nan_to_num_(g, nan=0.0, posinf=0.0, neginf=0.0)
def c(): opt.step(z, b())
for name, fn in (("logprob+grad", a), ("+scale/nan_to_num", b), ("+Adam (z moves)", c)):
    ms, cpu = timed(fn)
    print(f"{name:22s} {ms:.3f} ms/step  (CPU enqueue {cpu:.3f} ms/step)")
ms, cpu = timed(a)
print(f"{'logprob+grad at MAP z':22s} {ms:.3f} ms/step")
