"""End-to-end MAP throughput on C2 (driver-level: ModellingSequence.MAP with Adam), development aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigalens_b200 import workloads
from gigalens_b200.inference import Adam, ModellingSequence, PolynomialDecay
from gigalens_b200.model import ForwardProbModel

n, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 100
wl = workloads.c2_workload()
prob = ForwardProbModel(wl["prior"], wl["observed"], background_rms=0.2, exp_time=100.0)
seq = ModellingSequence(wl["phys_model"], prob, wl["sim_config"])
seq.MAP(Adam(PolynomialDecay(1e-2, steps, 1e-3)), n_samples=n, num_steps=5, seed=0)     # warm-up (plan creation, allocator)
torch.cuda.synchronize(); t0 = time.perf_counter()
z = seq.MAP(Adam(PolynomialDecay(1e-2, steps, 1e-3)), n_samples=n, num_steps=steps, seed=0)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"MAP n_samples={n} steps={steps}: {dt:.3f} s -> {n * steps / dt:.0f} logprob+grad evals/s "
      f"({dt / steps * 1e3:.3f} ms/step incl. plan creation); best red_chi2 {float(seq.last_red_chi2.min()):.4f}")
