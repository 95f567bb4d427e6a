"""world_size-2 gloo tests (CPU) of the multi-process driver logic: sample sharding, the SVI
gradient all-reduce and the HMC adaptation reductions.  The hot path is replaced by an analytic
Gaussian target so no GPU is needed; the collectives and the sharding are the real code."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gigalens_b200.inference import Adam, ModellingSequence, PolynomialDecay, _shard

D = 5
MU = np.linspace(-1, 1, D).astype(np.float32)
SIG = np.linspace(0.5, 1.5, D).astype(np.float32)


class FakePrior:
    def sample(self, n, seed=0):
        return np.random.default_rng(seed).normal(size=(n, D)).astype(np.float32)


class FakeProb:
    """log p(z) = -1/2 sum ((z - mu)/sig)^2 with the ForwardProbModel evaluation interface."""
    prior = FakePrior()

    def bij_inverse(self, x):
        return np.asarray(x, dtype=np.float32)

    def log_prob_and_grad(self, sim, z):
        mu, sig = torch.as_tensor(MU), torch.as_tensor(SIG)
        u = (z - mu) / sig
        return -0.5 * (u ** 2).sum(1), (u ** 2).mean(1), -u / sig


    # the separable interface ModellingSequence.SMC uses: prior N(0, 2^2), target N(mu, sig^2), broad auxiliary
    def log_prior_and_grad(self, sim, z):
        return -0.5 * ((z / 2) ** 2).sum(1), -z / 4

    def term_and_grad(self, sim, z, term):
        if term == "none":
            return torch.zeros(z.shape[0]), torch.zeros_like(z)
        mu, sig = torch.as_tensor(MU), torch.as_tensor(SIG) * (3.0 if term == "positions" else 1.0)
        u = (z - mu) / sig
        return -0.5 * (u ** 2).sum(1), -u / sig


class FakeSim:
    def __init__(self, phys_model, sim_config, bs):
        self.bs, self.device, self.img_region = bs, torch.device("cpu"), torch.ones(4, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq = ModellingSequence(None, FakeProb(), None, simulator_cls=FakeSim)
    z = seq.MAP(Adam(PolynomialDecay(1e-1, 250, 1e-3)), n_samples=7, num_steps=300, seed=1)
    q, losses = seq.SVI(Adam(5e-2), start_mean=z[0], n_vi=16, init_scales=0.3, num_steps=150, seed=4)
    samples, stats = seq.HMC(q, init_eps=0.3, init_l=3, n_hmc=6, num_burnin_steps=40, num_results=20, max_leapfrog_steps=10)
    out[rank] = dict(z=z.numpy(), mean=q.mean().numpy(), cov=q.covariance().numpy(), losses=losses,
                     samples=samples.numpy(), eps=stats["step_size"], nleap=stats["num_leapfrog"])
    dist.destroy_process_group()


def _smc_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq = ModellingSequence(None, FakeProb(), None, simulator_cls=FakeSim)
    samples, info = seq.SMC(num_particles=300, num_ensembles=2, num_leapfrog_steps=5, post_sampling_steps=20, seed=3)
    cloud, info0 = seq.SMC(num_particles=64, num_ensembles=1, num_leapfrog_steps=4, post_sampling_steps=0, seed=5)
    out[rank] = dict(samples=samples.numpy(), betas=info["inverse_temperature"], steps=info["num_steps"],
                     evidence=info["log_evidence"], cloud=cloud.numpy())
    dist.destroy_process_group()


def _smc_posterior():
    pv = 1.0 / (1.0 / 4.0 + 1.0 / SIG.astype(np.float64) ** 2)
    pm = pv * (MU / SIG.astype(np.float64) ** 2)
    s2 = SIG.astype(np.float64) ** 2
    logz = np.sum(0.5 * np.log(2 * np.pi * s2) - 0.5 * np.log(2 * np.pi * (4 + s2)) - 0.5 * MU ** 2 / (4 + s2))
    return pm, np.sqrt(pv), logz


def test_smc_recovers_gaussian_posterior_and_evidence():
    """Tempered SMC with an auxiliary likelihood on an analytic target: posterior moments and log-evidence."""
    seq = ModellingSequence(None, FakeProb(), None, simulator_cls=FakeSim)
    samples, info = seq.SMC(num_particles=500, num_ensembles=2, num_leapfrog_steps=5, post_sampling_steps=30, seed=3)
    assert samples.shape == (30, 1000, D) and info["inverse_temperature"][-1] == [1.0, 1.0]
    betas = np.asarray(info["inverse_temperature"])
    assert np.all(np.diff(betas, axis=0) >= 0) and info["stages"] == len(betas) >= 2
    z = samples.reshape(-1, D).numpy()
    pm, ps, logz = _smc_posterior()
    assert np.allclose(z.mean(0), pm, atol=0.05) and np.allclose(z.std(0), ps, rtol=0.1)
    assert all(1 <= n <= 8 for n in info["num_steps"])
    # target only, no auxiliary term; starting cloud given explicitly
    start = np.random.default_rng(0).normal(size=(50, 3, D)).astype(np.float32) * 2
    cloud, info = seq.SMC(start=start, num_particles=400, num_ensembles=1, num_leapfrog_steps=4, post_sampling_steps=0,
                          auxiliar="none", seed=2)
    assert cloud.shape == (400, 1, D) and np.allclose(cloud.reshape(-1, D).numpy().mean(0), pm, atol=0.15)
    # the incremental weights use the target likelihood only (as TFP does), so the evidence estimate is
    # meaningful without an auxiliary term -- here the start cloud is N(0, 2^2) = the prior
    assert np.allclose(info["log_evidence"], logz, atol=0.35)


@pytest.mark.timeout(300)
def test_two_rank_gloo_smc():
    """Particles sharded over two ranks; resampling mixes them through the all-gather; both ranks agree on the
    tempering schedule and together reproduce the posterior."""
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    out = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_smc_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    a, b = out[0], out[1]
    assert a["betas"] == b["betas"] and a["steps"] == b["steps"] and a["evidence"] == b["evidence"]
    assert a["samples"].shape == (20, 300, D) and b["samples"].shape == (20, 300, D)
    assert np.array_equal(a["cloud"], b["cloud"]) and a["cloud"].shape == (64, 1, D)
    z = np.concatenate([a["samples"], b["samples"]], 1).reshape(-1, D)
    pm, ps, logz = _smc_posterior()
    assert np.allclose(z.mean(0), pm, atol=0.06) and np.allclose(z.std(0), ps, rtol=0.12)


def test_shard_partitions():
    for n in (1, 7, 16, 4096):
        for world in (1, 2, 3, 8):
            spans = [_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_drivers_match_single_process():
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    out2 = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out2)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    out1 = mgr.dict()
    p = ctx.Process(target=_worker, args=(0, 1, _free_port(), out1))
    p.start()
    p.join(240)
    assert p.exitcode == 0
    a, b, single = out2[0], out2[1], out1[0]
    # MAP has no collective: the gathered result equals the single-process run and converges to mu
    assert np.allclose(a["z"], b["z"]) and np.allclose(a["z"], single["z"], atol=1e-6)
    assert np.allclose(a["z"], MU[None, :], atol=5e-2)
    # SVI: both ranks hold the same surrogate (one all-reduce per step) and it fits the Gaussian target
    assert np.allclose(a["mean"], b["mean"]) and np.allclose(a["cov"], b["cov"]) and a["losses"] == b["losses"]
    assert np.allclose(a["mean"], MU, atol=0.15) and np.allclose(np.sqrt(np.diag(a["cov"])), SIG, rtol=0.35)
    assert a["losses"][-1] < a["losses"][0]
    # HMC: chains are sharded (3 + 3), adaptation state is common to the ranks
    assert a["samples"].shape == (20, 3, D) and b["samples"].shape == (20, 3, D) and single["samples"].shape == (20, 6, D)
    assert a["eps"] == b["eps"] and a["nleap"] == b["nleap"]
    allz = np.concatenate([a["samples"], b["samples"]], 1).reshape(-1, D)
    assert np.all(np.abs(allz.mean(0) - MU) < 1.0)
