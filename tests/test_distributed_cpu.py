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
