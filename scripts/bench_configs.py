"""Throughput of the other BASELINE.json configs (development / reporting aid; bench.py is the contract):
  C3  EPL+shear / Shapelets(n_max=10) via lstsq_simulate, BackwardProbModel, bs=2048
  C4  NFW + 30-member dPIE scaling relation + shear / SersicEllipse, 200x200, ss=2, bs=1024
Usage: python scripts/bench_configs.py [c3|c4|all] [bs]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gigalens_b200 import workloads
from gigalens_b200.model import BackwardProbModel, ForwardProbModel
from gigalens_b200.simulator import LensSimulator


def timeit(fn, reps):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


which = sys.argv[1] if len(sys.argv) > 1 else "all"
res = {}
if which in ("c3", "all"):
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    wl = workloads.c3_workload(observed=workloads.c3_observation())
    sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
    pm = BackwardProbModel(wl["prior"], wl["observed"], wl["background_rms"], wl["exp_time"])
    z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(bs, seed=0)), device="cuda")
    ms_f, out = timeit(lambda: pm.log_prob(sim, z), 3)
    ms, out = timeit(lambda: pm.log_prob_and_grad(sim, z), 3)
    res["c3"] = dict(workload=wl["name"], bs=bs, ms_fwd=ms_f, ms_fwd_bwd=ms, evals_per_s=bs / ms * 1e3,
                     best_red_chi2=float(torch.nan_to_num(out[1], nan=1e30).min()), finite=float(torch.isfinite(out[0]).float().mean()))
    print(json.dumps(res["c3"]), flush=True)
    del sim
if which in ("c4", "all"):
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    wl = workloads.c4_workload(observed=workloads.c4_observation())
    sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
    pm = ForwardProbModel(wl["prior"], wl["observed"], background_rms=wl["background_rms"], exp_time=wl["exp_time"])
    z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(bs, seed=0)), device="cuda")
    ms_f, out = timeit(lambda: pm.log_prob(sim, z), 3)
    ms, out = timeit(lambda: pm.log_prob_and_grad(sim, z), 3)
    res["c4"] = dict(workload=wl["name"], bs=bs, ms_fwd=ms_f, ms_fwd_bwd=ms, evals_per_s=bs / ms * 1e3,
                     finite=float(torch.isfinite(out[0]).float().mean()))
    print(json.dumps(res["c4"]), flush=True)
