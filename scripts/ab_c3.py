"""C3 (lstsq) step time with the eigen-solve on the tail stream / in line and with forced chunks: python scripts/ab_c3.py [bs [chunk ...]]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigalens_b200 import workloads
from gigalens_b200.model import BackwardProbModel
from gigalens_b200.simulator import LensSimulator
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
wl = workloads.c3_workload(observed=workloads.c3_observation())
ref = None
CH = [int(a) for a in sys.argv[2:]]
for opts in ([{}, {"lstsq_hide_tail": 0}, {"lstsq_chunk": bs // 2}] if not CH else [{}] + [{"lstsq_chunk": c} for c in CH]):
    sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
    pm = BackwardProbModel(wl["prior"], wl["observed"], wl["background_rms"], wl["exp_time"])
    z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(bs, seed=0)), device="cuda")
    pm._bind(sim)
    for k, v in opts.items():
        sim.set_option(k, v)
    for _ in range(2):
        out = pm.log_prob_and_grad(sim, z)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        out = pm.log_prob_and_grad(sim, z)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    res = [t.clone() for t in out]
    same = None if ref is None else [bool(torch.equal(x, y)) for x, y in zip(res, ref)]
    if ref is None: ref = res
    print(json.dumps({"options": opts, "ms_per_step": ms, "evals_per_s": bs / ms * 1e3, "bit_identical_to_first": same}), flush=True)
    del sim
