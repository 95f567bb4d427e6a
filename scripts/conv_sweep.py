"""TMA-staged vs cp.async conv kernels over a sweep of image sizes, PSF sizes and supersampling factors:
images, log-likelihoods and gradients must be bit-identical (development aid / GPU regression sweep)."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gigalens_b200 import workloads
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator, SimulatorConfig

wl = workloads.c2_workload()
rng = np.random.default_rng(0)
bad = 0
for num_pix, K, ss in itertools.product([12, 20, 36, 60, 64, 100, 200], [3, 5, 9, 13, 21], [1, 2, 3, 4]):
    if num_pix * ss > 400:
        continue
    g = np.exp(-0.5 * (np.arange(K) - K // 2) ** 2 / (0.15 * K + 0.5) ** 2)
    psf = np.outer(g, g) * (1 + 0.05 * rng.standard_normal((K, K)))
    psf = (psf / psf.sum()).astype(np.float32)
    cfg = SimulatorConfig(delta_pix=3.9 / num_pix, num_pix=num_pix, supersample=ss, kernel=psf)
    bs = 3
    obs = rng.normal(1.0, 0.3, (num_pix, num_pix))
    pm = ForwardProbModel(wl["prior"], obs, background_rms=0.2, exp_time=100.0)
    z = torch.as_tensor(pm.bij_inverse(wl["prior"].sample(bs, seed=1)), device="cuda")
    outs = []
    try:
      for tma in (1, 0):
        sim = LensSimulator(wl["phys_model"], cfg, bs=bs)
        sim.set_option("conv_tma", tma)
        res = [t.cpu().numpy() for t in pm.log_prob_and_grad(sim, z)]
        res.append(sim.simulate(pm.bij_forward(sim, z)).cpu().numpy())
        outs.append(res)
    except RuntimeError as e:
        bad += 1
        print(f"ERROR num_pix={num_pix} K={K} ss={ss} tma={tma}: {e}", flush=True)
        continue
    same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(*outs))
    finite = bool(np.isfinite(outs[0][3]).all())
    if not same or not finite:
        bad += 1
        print(f"MISMATCH num_pix={num_pix} K={K} ss={ss} same={same} finite={finite}", flush=True)
print("conv sweep done, mismatches:", bad)
