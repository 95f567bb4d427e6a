"""A/B timing of plan options on the C2 log-prob + gradient step with per-kernel CUDA-event times (development aid).
usage: python scripts/ab_bench.py [bs] name=value[,name=value...] ..."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gigalens_b200 import workloads, _cabi
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator

STAGES = ["unconstrain", "prep", "raytrace_fwd", "conv_fwd", "conv_bwd", "raytrace_bwd", "sample_bwd"]
bs = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
variants = [a for a in sys.argv[1:] if "=" in a] or ["row_flush=1"]
wl = workloads.c2_workload()
pm = ForwardProbModel(wl["prior"], wl["observed"], background_rms=0.2, exp_time=100.0)
z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(bs, seed=0)), device="cuda")
lib = _cabi.load()
ref = None
for var in variants:
    sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
    for kv in var.split(","):
        k, v = kv.split("="); sim.set_option(k, int(v))
    for _ in range(3): out = pm.log_prob_and_grad(sim, z)
    torch.cuda.synchronize()
    n = 20
    sim.set_option("timing", n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = pm.log_prob_and_grad(sim, z)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    stage = (C.c_float * 7)(); nc = C.c_int32(0)
    _cabi.check(lib.gl_plan_get_timings(sim._plan, stage, C.byref(nc)), lib)
    sim.set_option("timing", 0)
    st = "  ".join(f"{s}={stage[i] / max(nc.value, 1):.3f}" for i, s in enumerate(STAGES))
    res = [o.double().cpu().numpy() for o in out]
    msg = ""
    if ref is None: ref = res
    else:
        dl = np.max(np.abs(res[0] - ref[0]) / np.abs(ref[0]))
        g, g0 = res[-1], ref[-1]
        dg = np.max(np.max(np.abs(g - g0), axis=0) / np.max(np.abs(g0), axis=0))
        msg = f"  | vs first: logp rel {dl:.2e}, dz rel(col max) {dg:.2e}"
    print(f"[{var}] {ms:.3f} ms/step -> {bs / ms * 1e3:.0f} evals/s{msg}\n    {st}", flush=True)
