"""Per-component view of the z-gradient of chosen samples of the C2 bench batch: CUDA vs fp64 / fp32 / ulp-perturbed oracle
(development aid for the per-sample parity test).  python scripts/dz_outliers.py [sample ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import common, oracle_bridge
from gigalens_b200 import _cabi, workloads
if os.environ.get("GL_EXP_LIB"):      # an experimental build of the library (e.g. -DGL_ACCURATE_LOGEXP under scripts/probes/_exp/)
    _cabi.library_path = lambda: os.path.join(ROOT, os.environ["GL_EXP_LIB"])
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator
np.set_printoptions(linewidth=220, precision=3)
idx = [int(a) for a in sys.argv[1:]] or [768, 2288, 2928, 0, 16]
wl = workloads.c2_workload()
bs = 4096
pmod = ForwardProbModel(wl["prior"], wl["observed"], background_rms=wl["background_rms"], exp_time=wl["exp_time"])
draw = wl["prior"].sample(bs, seed=0)
z = pmod.bij_inverse(draw)
outs = {}
for tag, opts in (("default", {}),):
    sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
    pmod._bind(sim) if hasattr(pmod, "_bind") else None
    for k, v in opts.items():
        sim.set_option(k, v)
    zt = torch.as_tensor(z, device="cuda")
    outs[tag] = [t.cpu().numpy() for t in pmod.log_prob_and_grad(sim, zt)]
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): pmod.log_prob_and_grad(sim, zt)
    e1.record(); torch.cuda.synchronize()
    print(f"[{tag}] step {e0.elapsed_time(e1) / 20:.3f} ms")
zs = z[idx]
r = oracle_bridge.logprob_and_grad(wl, zs.astype(np.float64), torch.float64)
s = oracle_bridge.logprob_and_grad(wl, zs, torch.float32)
p = oracle_bridge.logprob_and_grad(wl, common.ulp_perturb(zs), torch.float64)
keys = sim.compiled.slot_keys
print("params:", [f"{k[0][:2]}{k[1]}.{k[2]}" for k in keys])
for j, b in enumerate(idx):
    sc = np.abs(r[2][j]).max()
    print(f"--- sample {b}: logp {r[0][j]:.6g} red_chi2 {r[1][j]:.4g}  max|dz| {sc:.4g}")
    print(" dz64      ", r[2][j])
    for tag in outs:
        print(f" err {tag:8s}", (outs[tag][2][b] - r[2][j]) / sc)
    print(" err fp32orc ", (s[2][j] - r[2][j]) / sc)
    print(" err perturb ", (p[2][j] - r[2][j]) / sc)
    vals = {k2: np.asarray(v)[b] for k2, v in [(f"{g[:2]}{i}.{n}", draw[g][i][n]) for g in draw for i in range(len(draw[g])) for n in draw[g][i]]}
    print(" params", {k2: float(f"{v:.4g}") for k2, v in vals.items()})
