"""How accurate is the fp32 source-plane position beta on the ray-shooting grid: CUDA (MUFU-based) vs the torch-fp32 oracle, both
against the fp64 oracle, for chosen samples of the C2 bench batch.  python scripts/beta_error.py [sample ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_bridge
from gigalens_b200 import workloads
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator
idx = [int(a) for a in sys.argv[1:]] or [768, 2288, 2928, 0, 16, 32, 48, 64]
wl = workloads.c2_workload()
pmod = ForwardProbModel(wl["prior"], wl["observed"], background_rms=wl["background_rms"], exp_time=wl["exp_time"])
draw = wl["prior"].sample(4096, seed=0)
sub = {g: [{k: np.asarray(v)[idx] for k, v in d.items()} for d in draw[g]] for g in draw}
bs = len(idx)
sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
o64, _ = oracle_bridge.build_oracle(wl, bs, torch.float64)
o32, _ = oracle_bridge.build_oracle(wl, bs, torch.float32)
X64, Y64 = o64.img_X[:, 0].numpy(), o64.img_Y[:, 0].numpy()
X32 = o32.img_X[:, 0].numpy()
assert np.array_equal(X64.astype(np.float32), X32)
lens = lambda dt: [{k: torch.as_tensor(v.astype(np.float32)).to(dt) for k, v in d.items()} for d in sub["lens_mass"]]
bx64, by64 = (t.numpy() for t in o64.beta(o64.img_X, o64.img_Y, lens(torch.float64)))      # (N, bs)
bx32, by32 = (t.numpy().astype(np.float64) for t in o32.beta(o32.img_X, o32.img_Y, lens(torch.float32)))
lm = [{k: torch.as_tensor(v.astype(np.float32), device="cuda") for k, v in d.items()} for d in sub["lens_mass"]]
cbx, cby = sim.beta(X32, o32.img_Y[:, 0].numpy(), lm)
cbx, cby = cbx.cpu().numpy().astype(np.float64).T, cby.cpu().numpy().astype(np.float64).T            # (N, bs)
for j, b in enumerate(idx):
    ec = np.hypot(cbx[:, j] - bx64[:, j], cby[:, j] - by64[:, j]); eo = np.hypot(bx32[:, j] - bx64[:, j], by32[:, j] - by64[:, j])
    print(f"sample {b:5d}: |d beta| CUDA rms {np.sqrt((ec**2).mean()):.2e} max {ec.max():.2e}   torch-fp32 rms {np.sqrt((eo**2).mean()):.2e} max {eo.max():.2e}"
          f"   ratio rms {np.sqrt((ec**2).mean() / (eo**2).mean()):.2f}")
