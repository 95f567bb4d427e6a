"""Where does the fixed cost of ModellingSequence.MAP go (plan creation, prior draws, ...)?  Development aid."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigalens_b200 import workloads
from gigalens_b200.inference import Adam, ModellingSequence
from gigalens_b200.model import ForwardProbModel
wl = workloads.c2_workload()
prob = ForwardProbModel(wl["prior"], wl["observed"], background_rms=0.2, exp_time=100.0)
seq = ModellingSequence(wl["phys_model"], prob, wl["sim_config"])
seq.MAP(Adam(1e-2), n_samples=4096, num_steps=2, seed=0)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter()
seq.MAP(Adam(1e-2), n_samples=4096, num_steps=2, seed=0)
torch.cuda.synchronize()
print("MAP(2 steps):", time.perf_counter() - t0, "s")
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
