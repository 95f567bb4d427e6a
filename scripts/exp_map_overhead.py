"""Where the MAP driver's time goes on C2 (bs 4096): the fused log-prob + gradient call alone, with the torch Adam update, and through
ModellingSequence.MAP.  Development aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigalens_b200 import workloads
from gigalens_b200.inference import Adam, ModellingSequence
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator

n, steps = 4096, 300
wl = workloads.c2_workload()
pm = ForwardProbModel(wl["prior"], wl["observed"], background_rms=0.2, exp_time=100.0)
sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=n)
z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(n, seed=0)), device="cuda")


def timed(fn, label):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(steps): fn()
    e1.record(); t_issue = time.perf_counter() - t0
    torch.cuda.synchronize(); t_wall = time.perf_counter() - t0
    print(f"{label}: device {e0.elapsed_time(e1) / steps:.4f} ms/step, host issue {t_issue / steps * 1e3:.4f} ms/step, wall {t_wall / steps * 1e3:.4f} ms/step", flush=True)


timed(lambda: pm.log_prob_and_grad(sim, z), "log_prob_and_grad only")
opt = Adam(1e-3)
zz = z.clone()


def step():
    logp, chi2, dz = pm.log_prob_and_grad(sim, zz)
    opt.step(zz, dz.mul_(-1.0 / (3600 * n)))


timed(step, "log_prob_and_grad + torch Adam")
seq = ModellingSequence(wl["phys_model"], pm, wl["sim_config"])
seq._sim = sim
for k in (1, 2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    seq.MAP(Adam(1e-3), n_samples=n, num_steps=steps, seed=0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"ModellingSequence.MAP call {k}: {dt / steps * 1e3:.4f} ms/step ({n * steps / dt / 1e6:.3f} M evals/s)", flush=True)
