"""Forward conv with and without TMA staging on the same input (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gigalens_b200 import workloads
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
wl = workloads.c2_workload()
pm = ForwardProbModel(wl["prior"], wl["observed"], background_rms=0.2, exp_time=100.0)
z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(bs, seed=0)), device="cuda")
outs = []
for tma in (0, 1):
    sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
    sim.set_option("conv_tma", tma)
    out = pm.log_prob_and_grad(sim, z)
    torch.cuda.synchronize()
    outs.append([o.double().cpu().numpy() for o in out])
    print("tma", tma, "logp", outs[-1][0][:3], flush=True)
print("max |dlogp|", np.max(np.abs(outs[0][0] - outs[1][0])), "max |ddz|", np.max(np.abs(outs[0][-1] - outs[1][-1])))
