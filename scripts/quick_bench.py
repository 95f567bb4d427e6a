"""Quick device timing of the C2 log-prob + gradient step (development aid, not bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gigalens_b200 import workloads
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wl = workloads.c2_workload()
sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
pm = ForwardProbModel(wl["prior"], wl["observed"], background_rms=0.2, exp_time=100.0)
z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(bs, seed=0)), device="cuda")
for grad in (False, True):
    fn = (lambda: pm.log_prob_and_grad(sim, z)) if grad else (lambda: pm.log_prob(sim, z))
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"bs={bs} grad={grad}: {ms:.3f} ms/step -> {bs/ms*1e3:.0f} evals/s; logp[0]={float(out[0][0]):.3f}")
