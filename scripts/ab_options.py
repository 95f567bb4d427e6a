"""A/B of plan options on one workload: per-kernel CUDA-event times (gl_plan_get_timings) and step time.
   python scripts/ab_options.py c2|c4 [bs] opt=val[,opt=val] [opt=val ...]      (each argument after bs = one variant)"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigalens_b200 import _cabi, workloads
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator

STAGES = ["unconstrain", "prep", "raytrace_fwd", "conv_fwd", "conv_bwd", "raytrace_bwd", "sample_bwd"]
which, bs = sys.argv[1], int(sys.argv[2])
variants = sys.argv[3:] or [""]
wl = workloads.c2_workload() if which == "c2" else workloads.c4_workload(observed=workloads.c4_observation())
lib = _cabi.load()
ref = None
for v in variants:
    sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
    for kv in filter(None, v.split(",")):
        k, val = kv.split("=")
        sim.set_option(k, int(val))
    pm = ForwardProbModel(wl["prior"], wl["observed"], background_rms=wl["background_rms"], exp_time=wl["exp_time"])
    z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(bs, seed=0)), device="cuda")
    reps = 30 if which == "c2" else 5
    for _ in range(3):
        out = pm.log_prob_and_grad(sim, z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = pm.log_prob_and_grad(sim, z)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sim.set_option("timing", reps)
    for _ in range(reps):
        pm.log_prob_and_grad(sim, z)
    st = (C.c_float * 7)(); n = C.c_int32(0)
    _cabi.check(lib.gl_plan_get_timings(sim._plan, st, C.byref(n)), lib)
    res = [t.clone() for t in out]
    same = None if ref is None else [bool(torch.equal(a, b)) for a, b in zip(res, ref)]
    if ref is None:
        ref = res
    print(json.dumps({"variant": v or "default", "ms_per_step": ms, "evals_per_s": bs / ms * 1e3,
                      "kernel_ms": {s: st[i] / max(1, n.value) for i, s in enumerate(STAGES)}, "bit_identical_to_first": same}), flush=True)
    del sim
