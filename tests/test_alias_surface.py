"""The ``gigalens.*`` import surface: a script written against the reference keeps its imports and only swaps the
TensorFlow-Probability prior spec for the shim.  The model spec below restates the reference's own fixtures
(``tests/conftest.py:21-87``) and the flows of ``tests/tf/test_model.py`` through those module paths."""
import math

import numpy as np
import pytest
import torch

from gigalens import distributions as tfd            # the one import a reference script swaps
from gigalens.model import PhysicalModel              # tests/conftest.py:8
from gigalens.simulator import SimulatorConfig        # tests/tf/test_model.py:5
from gigalens.tf.inference import ModellingSequence   # tests/tf/test_model.py:6
from gigalens.tf.model import BackwardProbModel, ForwardProbModel   # tests/tf/test_model.py:7
from gigalens.tf.profiles.light import sersic         # tests/conftest.py:9
from gigalens.tf.profiles.mass import epl, shear      # tests/conftest.py:10
from gigalens.tf.simulator import LensSimulator


def default_prior():
    """tests/conftest.py:21-73, verbatim structure: a Sequential of Sequentials of Named."""
    lens_prior = tfd.JointDistributionSequential([
        tfd.JointDistributionNamed(dict(
            theta_E=tfd.LogNormal(math.log(1.25), 0.25), gamma=tfd.TruncatedNormal(2, 0.25, 1, 3), e1=tfd.Normal(0, 0.1),
            e2=tfd.Normal(0, 0.1), center_x=tfd.Normal(0, 0.05), center_y=tfd.Normal(0, 0.05))),
        tfd.JointDistributionNamed(dict(gamma1=tfd.Normal(0, 0.05), gamma2=tfd.Normal(0, 0.05))),
    ])
    lens_light_prior = tfd.JointDistributionSequential([
        tfd.JointDistributionNamed(dict(
            R_sersic=tfd.LogNormal(math.log(1.0), 0.15), n_sersic=tfd.Uniform(2, 6), e1=tfd.TruncatedNormal(0, 0.1, -0.3, 0.3),
            e2=tfd.TruncatedNormal(0, 0.1, -0.3, 0.3), center_x=tfd.Normal(0, 0.05), center_y=tfd.Normal(0, 0.05),
            Ie=tfd.LogNormal(math.log(500.0), 0.3)))])
    source_light_prior = tfd.JointDistributionSequential([
        tfd.JointDistributionNamed(dict(
            R_sersic=tfd.LogNormal(math.log(0.25), 0.15), n_sersic=tfd.Uniform(0.5, 4), e1=tfd.TruncatedNormal(0, 0.15, -0.5, 0.5),
            e2=tfd.TruncatedNormal(0, 0.15, -0.5, 0.5), center_x=tfd.Normal(0, 0.25), center_y=tfd.Normal(0, 0.25),
            Ie=tfd.LogNormal(math.log(150.0), 0.5)))])
    return tfd.JointDistributionSequential([lens_prior, lens_light_prior, source_light_prior])


def default_physmodel():
    return PhysicalModel([epl.EPL(), shear.Shear()], [sersic.SersicEllipse()], [sersic.SersicEllipse()])   # conftest.py:77-80


DEFAULT_DATA = (np.zeros((20, 20)), 0.1, 100)   # conftest.py:84-85


def test_alias_modules_export_the_cuda_classes():
    import gigalens_b200.inference
    import gigalens_b200.model
    import gigalens_b200.simulator
    from gigalens.tf.profiles.mass import dpie_subhalo, nfw, piemd, piep, scaling_relation, sie, sis, tnfw
    from gigalens.tf.profiles.light import shapelets

    assert LensSimulator is gigalens_b200.simulator.LensSimulator
    assert ForwardProbModel is gigalens_b200.model.ForwardProbModel and BackwardProbModel is gigalens_b200.model.BackwardProbModel
    assert ModellingSequence is gigalens_b200.inference.ModellingSequence
    for mod, names in ((nfw, ("NFW", "NFW_ELLIPSE")), (piemd, ("DPIS", "DPIE")), (piep, ("DPIEP",)), (tnfw, ("TNFW",)), (sie, ("SIE",)),
                       (sis, ("SIS",)), (scaling_relation, ("ScalingRelation",)), (dpie_subhalo, ("DPIESubhalo",)), (shapelets, ("Shapelets",))):
        for n in names:
            assert hasattr(mod, n)
    import gigalens_b200.profiles.light.sersic as b_sersic
    assert sersic.CoreSersic is b_sersic.CoreSersic and sersic.CoreSersic._params == [
        "R_sersic", "n_sersic", "Rb", "alpha", "gamma", "e1", "e2", "center_x", "center_y"]   # tf/profiles/light/sersic.py:83-97


def test_reference_fixture_spec_builds_and_bijector_round_trips():
    """tests/tf/test_model.py:10-26 on the host: ForwardProbModel(default_prior, ones, 1, 1) with the reference's positional
    arguments and defaults; bij.forward(bij.inverse(sample)) == sample; one z row per sample."""
    prior = default_prior()
    model = ForwardProbModel(prior, np.ones((20, 20)), 1, 1)
    assert model.include_pixels and not model.include_positions      # no centroids given: no position term
    sample = prior.sample(5, seed=0)
    z = model.bij.inverse(sample)
    assert z.shape == (5, 22)
    back = model.bij.forward(z)
    flat = lambda t: prior.flatten_values(t)
    for a, b in zip(flat(back), flat(sample)):
        assert np.allclose(a, b, rtol=1e-5, atol=1e-6)
    # list-of-lists top level maps onto the simulator's (group, index, name) slots
    from gigalens_b200.simulator import CompiledModel
    cm = CompiledModel(default_physmodel())
    keys = {model._slot_key(path) for path, _ in model._leaves}
    assert keys == set(cm.slot_keys)


def test_prior_sample_shapes_follow_tfp():
    prior = default_prior()
    s = prior.sample((6, 2), seed=1)                        # tf/inference.py:199: prior.sample((num_particles, num_ensembles))
    assert s[0][0]["theta_E"].shape == (6, 2)
    assert np.shape(prior.sample(seed=0)[0][0]["theta_E"]) == ()
    z = ForwardProbModel(prior, *DEFAULT_DATA).bij.inverse(s)
    assert z.shape == (12, 22)


@pytest.mark.gpu
def test_reference_test_model_flows_through_the_alias_imports():
    """tests/tf/test_model.py:29-72 (test_map, test_vi, test_hmc) with the reference's call signatures."""
    from gigalens.tf.inference import Adam

    prior, phys = default_prior(), default_physmodel()
    prob_model = ForwardProbModel(prior, *DEFAULT_DATA)
    sim_config = SimulatorConfig(delta_pix=0.05, num_pix=20)
    model_seq = ModellingSequence(phys, prob_model, sim_config)
    start = prob_model.prior.sample(2, seed=0)
    flat = lambda t: np.concatenate([np.ravel(v) for v in prior.flatten_values(t)])
    ret = model_seq.MAP(Adam(0), start, n_samples=2, num_steps=5, seed=0)
    assert np.allclose(flat(start), flat(prob_model.bij.forward(ret)), rtol=1e-5, atol=1e-6)
    ret = model_seq.MAP(Adam(1e-3), start, n_samples=2, num_steps=5, seed=0)
    assert not np.allclose(flat(start), flat(prob_model.bij.forward(ret)))
    z0 = prob_model.bij.inverse(prior.sample(2, seed=3))[0]
    q_z, losses = model_seq.SVI(optimizer=Adam(0), start=z0, n_vi=5, num_steps=5)
    assert np.allclose(q_z.mean().cpu().numpy(), z0)
    q_z2, losses = model_seq.SVI(optimizer=Adam(1e-3), start=z0, n_vi=5, num_steps=5)
    assert not np.allclose(q_z2.mean().cpu().numpy(), z0)
    samples, stats = model_seq.HMC(q_z, n_hmc=3, init_eps=0.3, init_l=3, max_leapfrog_steps=5, num_burnin_steps=3, num_results=5)
    assert len(samples) == 5
    # the tiled-centroid layout of init_centroids goes through beta / magnification and comes back as (N, bs)
    sim = LensSimulator(phys, sim_config, bs=4)
    params = prob_model.bij_forward(sim, torch.as_tensor(prob_model.bij.inverse(prior.sample(4, seed=5)), device="cuda"))
    x = np.repeat(np.array([0.7, -0.4, 0.1], np.float32)[:, None], 4, axis=-1)
    y = np.repeat(np.array([0.2, 0.9, -0.8], np.float32)[:, None], 4, axis=-1)
    bx, by = sim.beta(x, y, params["lens_mass"] if isinstance(params, dict) else params[0])
    assert bx.shape == (3, 4)
    bx2, _ = sim.beta(x[:, 0], y[:, 0], params["lens_mass"] if isinstance(params, dict) else params[0])
    assert bx2.shape == (4, 3) and torch.equal(bx2.T, bx)
    with pytest.raises(ValueError):
        sim.beta(x + np.arange(4, dtype=np.float32), y, params["lens_mass"] if isinstance(params, dict) else params[0])
