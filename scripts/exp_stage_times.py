"""Per-kernel stage times of the C2 step with an experimental build of the library (GL_EXP_LIB=path): development aid."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gigalens_b200 import _cabi, workloads
if os.environ.get("GL_EXP_LIB"):
    _cabi.library_path = lambda: os.path.join(ROOT, os.environ["GL_EXP_LIB"])
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator
bs = 4096
wl = workloads.c2_workload()
sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
pm = ForwardProbModel(wl["prior"], wl["observed"], background_rms=0.2, exp_time=100.0)
z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(bs, seed=0)), device="cuda")
for _ in range(5): pm.log_prob_and_grad(sim, z)
n = 20
sim.set_option("timing", n)
for _ in range(n): pm.log_prob_and_grad(sim, z)
stage = (C.c_float * 7)(); nc = C.c_int32(0)
lib = _cabi.load()
_cabi.check(lib.gl_plan_get_timings(sim._plan, stage, C.byref(nc)), lib)
names = ["unconstrain", "prep", "raytrace_fwd", "conv_fwd", "conv_bwd", "raytrace_bwd", "sample_bwd"]
print(os.environ.get("GL_EXP_LIB", "default"), " ".join(f"{nm}={stage[k] / nc.value * 1e3:.1f}us" for k, nm in enumerate(names)))
