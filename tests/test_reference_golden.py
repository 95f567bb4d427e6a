"""The oracle against golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE FILES (tests/golden/
make_reference_golden.py: the unmodified ``/root/reference/src/gigalens/tf/**`` run on the torch-backed tensorflow stand-in
of ``oracle/tfshim``).  This is what pins the restatement in ``oracle/`` to the reference's code: same float32 arithmetic in
the same order must agree to a few ulp, the float64 runs to round-off.  CPU only; the CUDA path is compared with the same
vectors in tests/test_gpu_reference_golden.py."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import common
from oracle import model as OM
from oracle.simulator import OracleSimulator

sys.path.insert(0, os.path.join(common.HERE, "golden"))
import reference_cases as RC  # noqa: E402

GOLD = np.load(os.path.join(common.HERE, "golden", "reference_golden.npz"))
PSF = np.load(os.path.join(common.ROOT, "gigalens_b200", "assets", "psf.npy")).astype(np.float32)
DEMO = np.load(os.path.join(common.ROOT, "gigalens_b200", "assets", "demo.npy")).astype(np.float32)
DT = {"f32": torch.float32, "f64": torch.float64}
# same operations in the same order on the same backend: float32 may differ where the oracle fuses or reorders an
# expression (a few ulp of the largest value); float64 to round-off
TOL = {"f32": 2e-6, "f64": 1e-12}


def close(a, ref, tol, what):
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    if a.size == ref.size:      # bs = 1: the reference squeezes the batch axis away (tf/simulator.py:156)
        a = a.reshape(ref.shape)
    elif ref.ndim < a.ndim:     # Shear.hessian returns the bare (bs,) parameters (shear.py:18-26); the oracle broadcasts over the points
        ref = np.broadcast_to(ref[:, None, :], a.shape)
    assert a.shape == ref.shape, (what, a.shape, ref.shape)
    assert np.array_equal(np.isnan(a), np.isnan(ref)), (what, "NaN pattern")
    assert np.array_equal(a[np.isinf(ref)], ref[np.isinf(ref)]) and np.array_equal(np.isinf(a), np.isinf(ref)), (what, "inf pattern")
    m = np.isfinite(ref)
    scale = max(np.max(np.abs(ref[m])), 1e-30) if m.any() else 1.0
    err = np.max(np.abs(a[m] - ref[m])) / scale if m.any() else 0.0
    assert err <= tol, (what, err)
    return err


def T(a, dt, grad=False):
    t = torch.as_tensor(np.asarray(a, dtype=np.float64)).to(dt)
    return t.requires_grad_(True) if grad else t


PROFILE_CASES = RC.profile_cases()


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("key", sorted(PROFILE_CASES))
def test_oracle_profiles_match_the_executed_reference(key, tag):
    c = PROFILE_CASES[key]
    dt = DT[tag]
    obj = common.to_oracle_profile(common.spec_profile(c["cls"], c["ctor"]), dt)
    p = {k: T(v, dt) for k, v in c["params"].items()}
    x, y = T(c["x"], dt), T(c["y"], dt)
    if c["kind"] == "mass":
        close(torch.stack(obj.deriv(x, y, **p)).detach().numpy(), GOLD[f"prof/{key}/{tag}/deriv"], TOL[tag], key + " deriv")
        h = torch.stack([v.detach() for v in obj.hessian(x.clone(), y.clone(), **p)]).numpy()
        # second derivatives through autodiff amplify the float32 differences of the forward pass
        close(h, GOLD[f"prof/{key}/{tag}/hessian"], 50 * TOL[tag], key + " hessian")
    else:
        close(obj.light(x, y, **p).detach().numpy(), GOLD[f"prof/{key}/{tag}/light"], TOL[tag], key + " light")


SIM_CASES = RC.simulator_cases(PSF, DEMO)


def build_oracle_case(c, dt):
    keys = RC.grad_keys(c["params"])
    bs = len(c["params"][keys[0][0]][keys[0][1]][keys[0][2]])
    pm = common.spec_model(c["model"], c.get("constants"))
    s = c["sim"]
    osim = OracleSimulator(common.to_oracle_model(pm, dt), s["delta_pix"], s["num_pix"], s["supersample"], kernel=s["kernel"],
                           pix_region=s["pix_region"], transform_pix2angle=s.get("transform_pix2angle"), bs=bs, dtype=dt)
    cen = c.get("centroids")
    kw = dict(error_map=c["error_map"]) if "error_map" in c else dict(background_rms=c["noise"]["background_rms"], exp_time=c["noise"]["exp_time"])
    if cen is not None:
        kw.update(centroids_x=cen["x"], centroids_y=cen["y"], centroids_errors_x=cen["ex"], centroids_errors_y=cen["ey"])
    opm = OM.ForwardProbModel(OM.JointPrior({}), c["observed"], dtype=dt, include_pixels=True, include_positions=cen is not None, **kw)
    opm.init_centroids(bs)
    params = {g: [{k: T(v, dt, grad=True) for k, v in d.items()} for d in c["params"][g]] for g in c["params"]}
    leaves = [params[g][i][k] for g, i, k in keys]
    return pm, osim, opm, params, leaves, bs


def test_the_executed_reference_reproduces_the_demo_image_pin():
    """tf-demo.ipynb cells 5-9 with the reference's own simulate / stats_pixels: reduced chi^2 of simulate(truth) against
    demo.npy (SURVEY 8c measured 0.981 with a throwaway numpy restatement; the notebook's MAP reaches 0.9794)."""
    chi2 = GOLD["sim/c2/f32/red_chi2"][0]
    assert 0.975 < chi2 < 0.99, chi2
    assert abs(GOLD["sim/c2/f32/image"][0].max() - 44.0347) < 2e-3


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("key", sorted(SIM_CASES))
def test_oracle_simulator_and_likelihood_match_the_executed_reference(key, tag):
    c = SIM_CASES[key]
    dt = DT[tag]
    pm, osim, opm, params, leaves, bs = build_oracle_case(c, dt)
    pre = f"sim/{key}/{tag}"
    tol = TOL[tag]
    img = osim.simulate(params)
    close(img.detach().numpy().reshape(GOLD[f"{pre}/image"].shape), GOLD[f"{pre}/image"], tol, "image")
    if not c.get("image_and_likelihood_only"):
        close(osim.simulate(params, no_deflection=True).detach().numpy(), GOLD[f"{pre}/image_no_deflection"], tol, "no_deflection")
    if c.get("variants"):
        close(osim.simulate_source(params).detach().numpy(), GOLD[f"{pre}/source"], tol, "simulate_source")
        close(osim.simulate_lens_light(params).detach().numpy(), GOLD[f"{pre}/lens_light"], tol, "simulate_lens_light")
        close(osim.simulate_images(params).detach().numpy(), GOLD[f"{pre}/images"], tol, "simulate_images")
    ll, chi2 = opm.stats_pixels(osim, params)
    close(ll.detach().numpy(), GOLD[f"{pre}/loglike"], tol, "log-like")
    close(chi2.detach().numpy(), GOLD[f"{pre}/red_chi2"], tol, "red chi2")
    g = torch.autograd.grad(ll.sum(), leaves, allow_unused=True)
    g = np.stack([np.zeros(bs) if v is None else v.numpy() for v in g])
    gref = GOLD[f"{pre}/grad"]
    g64 = GOLD[f"sim/{key}/f64/grad"]
    for (grp, i, k), a, r, r64 in zip(RC.grad_keys(c["params"]), g, gref, g64):
        # float32 autodiff through 160 000 rays sums rounding noise in an order-dependent way: at the cluster geometry the
        # reference's OWN float32 gradient is 1e-2 (one row 0.7) away from its float64 gradient, so two float32 evaluations
        # can only agree to that noise floor
        floor = float(np.max(np.abs(r - r64)) / max(np.max(np.abs(r64)), 1e-30)) if tag == "f32" else 0.0
        close(a, r, max(100 * tol, 2 * floor), f"grad {grp}[{i}].{k}")
    if c.get("image_and_likelihood_only"):
        return
    px, py = GOLD[f"sim/{key}/points"]
    X, Y = T(px[:, None].repeat(bs, axis=1), dt), T(py[:, None].repeat(bs, axis=1), dt)
    lens = [{k: v.detach() for k, v in d.items()} for d in params["lens_mass"]]
    close(torch.stack(osim.beta(X, Y, lens)).numpy(), GOLD[f"{pre}/beta"], tol, "beta")
    close(osim.magnification(X.clone(), Y.clone(), lens).detach().numpy(), GOLD[f"{pre}/magnification"], 100 * tol, "magnification")
    close(osim.convergence(X.clone(), Y.clone(), lens).detach().numpy(), GOLD[f"{pre}/convergence"], 100 * tol, "convergence")
    close(torch.stack([v.detach() for v in osim.shear(X.clone(), Y.clone(), lens)]).numpy(), GOLD[f"{pre}/shear"], 100 * tol, "shear")
    if c.get("centroids") is not None:
        llp, chi2p = opm.stats_positions(osim, params)
        close(llp.detach().numpy(), GOLD[f"{pre}/pos_loglike"], 100 * tol, "positions log-like")
        close(chi2p.detach().numpy(), GOLD[f"{pre}/pos_red_chi2"], 100 * tol, "positions red chi2")
        gp = torch.autograd.grad(llp.sum(), leaves, allow_unused=True)
        gp = np.stack([np.zeros(bs) if v is None else v.numpy() for v in gp])
        for (grp, i, k), a, r in zip(RC.grad_keys(c["params"]), gp, GOLD[f"{pre}/pos_grad"]):
            close(a, r, 1000 * tol, f"positions grad {grp}[{i}].{k}")
        # the combination of the terms (tf/model.py:126-181) with stand-in prior / bijector values
        ll_all, rc_all = opm._stats(osim, params)
        prior = T(RC.FAKE_LOG_PRIOR[:bs] + RC.FAKE_FLDJ[:bs], dt)
        close((ll_all + prior).detach().numpy(), GOLD[f"{pre}/logprob_total"], 100 * tol, "log_prob (pixels + positions + prior)")
        close(ll_all.detach().numpy(), GOLD[f"{pre}/loglike_total"], 100 * tol, "log_like (pixels + positions)")
        close(rc_all.detach().numpy(), GOLD[f"{pre}/red_chi2_total"], 100 * tol, "red chi2 averaged over the included terms")


def lstsq_oracle(dt, which="sersic"):
    c = RC.lstsq_sersic_case(PSF) if which == "sersic" else RC.lstsq_tail_case(PSF)
    pm = common.spec_model(c["model"])
    s = c["sim"]
    osim = OracleSimulator(common.to_oracle_model(pm, dt), s["delta_pix"], s["num_pix"], s["supersample"], kernel=s["kernel"], bs=2, dtype=dt)
    params = {g: [{k: T(v, dt) for k, v in d.items()} for d in c["params"][g]] for g in c["params"]}
    return c, pm, osim, params, T(c["observed"], dt), T(c["err_map"], dt)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_oracle_lstsq_simulate_matches_the_executed_reference(tag):
    """lstsq_simulate with the reference's own simulator object, beta, light calls and source lines :204-240 (depthwise conv, pooling,
    NaN scrubs, weights, normal equations, pinv(rcond=1e-6), recombination); only the broken scatter of :183-203 is restated in the
    generator.  Supersampled (ss = 2), PSF, three single-component linear profiles (what the reference's reshape can take)."""
    c, pm, osim, params, obs, err = lstsq_oracle(DT[tag])
    close(osim.lstsq_simulate(params, obs, err, return_stacked=True).numpy(), GOLD[f"lstsq/{tag}/stack"], TOL[tag], "stack")
    close(osim.lstsq_simulate(params, obs, err, return_coeffs=True).numpy(), GOLD[f"lstsq/{tag}/coeffs"], {"f32": 1e-4, "f64": 1e-10}[tag], "coeffs")
    close(osim.lstsq_simulate(params, obs, err).numpy(), GOLD[f"lstsq/{tag}/image"], {"f32": 1e-5, "f64": 1e-11}[tag], "image")


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_oracle_lstsq_tail_matches_the_executed_reference_lines(tag):
    """tf/simulator.py:231-240 (weights, normal equations, pinv(rcond=1e-6), recombination), executed from the reference file on the
    oracle's component stack of a model with a 15-component Shapelets set, against the oracle's own tail (measured bit-identical)."""
    c, pm, osim, params, obs, err = lstsq_oracle(DT[tag], "shapelets")
    close(osim.lstsq_simulate(params, obs, err, return_coeffs=True).numpy(), GOLD[f"lstsq_tail/{tag}/coeffs"], TOL[tag], "coeffs")
    close(osim.lstsq_simulate(params, obs, err).numpy(), GOLD[f"lstsq_tail/{tag}/image"], TOL[tag], "image")


def test_tensorflow_stand_in_matches_independent_statements_of_the_tf_semantics():
    """oracle/tfshim/selfcheck.py: SAME-padded NHWC conv / depthwise conv / pooling against explicit numpy loops, the scatter
    family and one-argument where against numpy indexing, nest.flatten's key order, while_loop's contract, GradientTape,
    pinv(rcond) against numpy.linalg, interp_regular_1d_grid against numpy.interp.  A subprocess: the module name ``tensorflow``
    must not leak into the test process."""
    out = subprocess.run([sys.executable, os.path.join(common.ROOT, "oracle", "tfshim", "selfcheck.py")], capture_output=True, text=True,
                         cwd=os.path.join(common.ROOT, "oracle"))
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr


def test_c4_case_uses_the_benchmark_catalogue():
    from gigalens_b200 import workloads
    assert RC.c4_catalogue() == workloads.cluster_catalogue(30, 7)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/gigalens"), reason="the reference tree exists only in the build container")
def test_committed_fixture_is_what_the_reference_produces_today(tmp_path):
    """Re-run the generator (a subprocess: the reference package shadows the repo's `gigalens` alias) and compare."""
    out = tmp_path / "g.npz"
    subprocess.check_call([sys.executable, os.path.join(common.HERE, "golden", "make_reference_golden.py"), str(out)], cwd=str(tmp_path))
    new = np.load(out)
    assert sorted(new.files) == sorted(GOLD.files)
    for k in new.files:      # not bit-exact: torch's CPU reductions depend on the thread count of the machine
        a, g = np.asarray(new[k], dtype=np.float64), np.asarray(GOLD[k], dtype=np.float64)
        fin = np.isfinite(g)
        assert a.shape == g.shape and np.array_equal(np.isfinite(a), fin), k
        scale = max(float(np.max(np.abs(g[fin]))), 1e-30) if fin.any() else 1.0
        tol = 1e-3 if ("/f32/" in k and "grad" in k) else 1e-5 if "/f32/" in k else 1e-10
        assert not fin.any() or float(np.max(np.abs(a[fin] - g[fin]))) <= tol * scale, k
