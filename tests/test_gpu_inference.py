"""GPU smoke tests of the inference drivers, mirroring the reference's ``tests/tf/test_model.py``:
bijector round trip, fldj shape, MAP / SVI with learning rate 0 leave the parameters unchanged and
a positive rate changes them, HMC returns ``num_results`` samples.  Data fixture as in the
reference's ``tests/conftest.py:83-85``: a 20x20 all-zeros image, bg 0.1, exp 100, delta_pix 0.05."""
import numpy as np
import pytest
import torch

from gigalens_b200 import workloads
from gigalens_b200.inference import Adam, ModellingSequence, PolynomialDecay
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator, SimulatorConfig

pytestmark = pytest.mark.gpu


@pytest.fixture
def setup():
    prior = workloads.demo_prior()
    phys = workloads.demo_phys_model()
    prob = ForwardProbModel(prior, np.zeros((20, 20)), background_rms=0.1, exp_time=100)
    cfg = SimulatorConfig(delta_pix=0.05, num_pix=20)
    return prior, phys, prob, cfg


def test_bij_round_trip_and_fldj_shape(setup):
    prior, phys, prob, cfg = setup           # tests/tf/test_model.py:10-26
    sample = prior.sample(5, seed=0)
    sim = LensSimulator(phys, cfg, bs=5)
    z = prob.bij_inverse(sample)
    back = prob.bij_forward(sim, torch.as_tensor(z, device="cuda"))
    flat_a = sim.compiled.flatten(back, 5, torch, sim.device).cpu().numpy()
    flat_b = sim.compiled.flatten(sample, 5, torch, "cpu").numpy()
    assert np.allclose(flat_a, flat_b, rtol=1e-5, atol=1e-6)
    assert prob.log_prior(sim, torch.as_tensor(z, device="cuda")).numel() == 5


def test_map_lr_zero_and_positive(setup):
    prior, phys, prob, cfg = setup           # tests/tf/test_model.py:29-43
    seq = ModellingSequence(phys, prob, cfg)
    start = prior.sample(2, seed=1)
    z0 = prob.bij_inverse(start)
    ret = seq.MAP(Adam(0.0), start, n_samples=2, num_steps=5, seed=0)
    assert np.allclose(ret.cpu().numpy(), z0)
    ret = seq.MAP(Adam(1e-3), start, n_samples=2, num_steps=5, seed=0)
    assert not np.allclose(ret.cpu().numpy(), z0)


def test_vi_lr_zero_and_positive_then_hmc(setup):
    prior, phys, prob, cfg = setup           # tests/tf/test_model.py:46-72
    seq = ModellingSequence(phys, prob, cfg)
    start = prob.bij_inverse(prior.sample(2, seed=2))[0]
    q_z, losses = seq.SVI(optimizer=Adam(0.0), start_mean=start, n_vi=5, num_steps=5)
    assert np.allclose(q_z.mean().cpu().numpy(), start)
    q_z2, losses = seq.SVI(optimizer=Adam(1e-3), start_mean=start, n_vi=5, num_steps=5)
    assert not np.allclose(q_z2.mean().cpu().numpy(), start) and len(losses) == 5
    samples, stats = seq.HMC(q_z, n_hmc=3, init_eps=0.3, init_l=3, max_leapfrog_steps=5, num_burnin_steps=3, num_results=5)
    assert len(samples) == 5 and samples.shape == (5, 3, 22)


def test_map_improves_fit_on_demo_image():
    """tf-demo.ipynb cell 12 in miniature: MAP from prior draws lowers the best reduced chi^2."""
    wl = workloads.c2_workload()
    prob = ForwardProbModel(wl["prior"], wl["observed"], background_rms=0.2, exp_time=100.0)
    seq = ModellingSequence(wl["phys_model"], prob, wl["sim_config"])
    hist = []
    seq.MAP(Adam(PolynomialDecay(1e-2, 150, 2e-3)), n_samples=64, num_steps=150, seed=0,
            callback=lambda i, chi: hist.append(float(torch.nan_to_num(chi, nan=1e30).min())))
    assert hist[-1] < 0.2 * hist[0] and hist[-1] < 10.0


def _positions_setup():
    """C2 model with one quadruply imaged point source: centroids from a 4-image cross around the Einstein ring."""
    prior, phys = workloads.demo_prior(), workloads.demo_phys_model()
    cx, cy = [np.array([1.1, -1.1, 0.05, -0.05], np.float32)], [np.array([0.05, -0.05, 1.1, -1.1], np.float32)]
    err = [np.full(4, 0.05, np.float32)]
    prob = ForwardProbModel(prior, workloads.load_demo_image(), background_rms=0.2, exp_time=100.0,
                            centroids_x=cx, centroids_y=cy, centroids_errors_x=err, centroids_errors_y=err)
    return prior, phys, prob, workloads.demo_sim_config()


def test_separable_terms_add_up_to_log_prob():
    """prior + pixels + positions (ModellingSequence.SMC's building blocks, tf/inference.py:208-233) equal the fused
    log_prob and its gradient."""
    prior, phys, prob, cfg = _positions_setup()
    bs = 6
    sim = LensSimulator(phys, cfg, bs=bs)
    z = torch.as_tensor(prob.bij_inverse(prior.sample(bs, seed=4)), device="cuda")
    logp, _, dz = prob.log_prob_and_grad(sim, z)
    lp, glp = prob.log_prior_and_grad(sim, z)
    l1, g1 = prob.term_and_grad(sim, z, "pixels")
    l2, g2 = prob.term_and_grad(sim, z, "positions")
    l0, g0 = prob.term_and_grad(sim, z, "none")
    assert torch.allclose(lp, prob.log_prior(sim, z)) and float(l0.abs().max()) == 0 and float(g0.abs().max()) == 0
    tot, gtot = (lp + l1 + l2).cpu().numpy(), (glp + g1 + g2).cpu().numpy()
    assert np.allclose(tot, logp.cpu().numpy(), rtol=2e-6)
    scale = np.abs(dz.cpu().numpy()).max(0)
    assert np.all(np.abs(gtot - dz.cpu().numpy()) <= 1e-5 * scale)


def test_smc_with_position_auxiliary_runs_and_tempers_to_one():
    prior, phys, prob, cfg = _positions_setup()
    seq = ModellingSequence(phys, prob, cfg)
    samples, info = seq.SMC(num_particles=48, num_ensembles=1, num_leapfrog_steps=3, post_sampling_steps=4,
                            max_sampling_per_stage=2, max_stage=40, seed=1)
    assert samples.shape == (4, 48, 22) and bool(torch.isfinite(samples).all())
    betas = np.asarray(info["inverse_temperature"])[:, 0]
    assert np.all(np.diff(betas) >= 0) and 0 < betas[0] <= 1 and info["stages"] <= 40
    cloud, info2 = seq.SMC(num_particles=32, num_leapfrog_steps=2, post_sampling_steps=0, max_sampling_per_stage=1,
                           max_stage=3, target="positions", auxiliar="none", seed=2)
    assert cloud.shape == (32, 1, 22) and info2["stages"] <= 3


def test_fused_adam_kernel_matches_the_torch_formulation():
    """gl_adam_step (one launch) against the eight-launch torch formulation of Keras' Adam, with the gradient scale and the
    NaN scrub folded in (tf/inference.py:34-37)."""
    import torch
    from gigalens_b200.inference import Adam, PolynomialDecay
    g = torch.Generator(device="cuda").manual_seed(0)
    x0 = torch.randn(4096, 22, device="cuda", generator=g)
    a, b = x0.clone(), x0.clone().cpu()
    oa, ob = Adam(PolynomialDecay(1e-2, 10, 1e-3)), Adam(PolynomialDecay(1e-2, 10, 1e-3))
    for t in range(12):
        grad = torch.randn(4096, 22, device="cuda", generator=g) * 10.0 ** float(t % 5 - 2)
        if t == 3:
            grad[5, 7] = float("nan"); grad[9, 1] = float("inf")
        oa.step(a, grad, grad_scale=-0.37, scrub_nan=True)       # CUDA float32 contiguous: the fused kernel
        ob.step(b, grad.cpu(), grad_scale=-0.37, scrub_nan=True)  # CPU tensors: the torch formulation
    assert torch.isfinite(a).all()

    def rel(u, w):
        return float((u.cpu() - w).abs().max() / w.abs().max())

    # twelve updates: a few float32 ulps (FMA contraction on the GPU, separate roundings on the CPU)
    errs = dict(x=rel(a, b), m=rel(oa.m, ob.m), v=rel(oa.v, ob.v))
    assert all(e < 2e-6 for e in errs.values()), errs
