"""The CUDA path (through the C ABI) against golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE FILES
(tests/golden/make_reference_golden.py; see tests/test_reference_golden.py for the oracle's check against the same vectors).
Truth = the reference run with ``tf.float32`` := float64; noise floor = the reference's own float32 run (plus the
ulp-perturbed float64 oracle, which test_reference_golden.py ties to the reference at 1e-14); tolerances of BASELINE.json."""
import numpy as np
import pytest
import torch

import common
from common import assert_parity
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator, SimulatorConfig
import test_reference_golden as R

pytestmark = pytest.mark.gpu
GOLD, RC = R.GOLD, R.RC


def _pts(c):
    return c["x"][:, 0].astype(np.float32), c["y"][:, 0].astype(np.float32)


@pytest.mark.parametrize("key", sorted(R.PROFILE_CASES))
def test_cuda_profiles_match_the_executed_reference(key):
    c = R.PROFILE_CASES[key]
    prof = common.spec_profile(c["cls"], c["ctor"])
    x, y = _pts(c)
    p = {k: v.astype(np.float32) for k, v in c["params"].items()}
    g32 = lambda q: (GOLD[f"prof/{key}/f32/{q}"], GOLD[f"prof/{key}/f64/{q}"])   # noqa: E731
    fin = lambda a: np.where(np.isfinite(a), a, 0.0)   # noqa: E731  (|e| = 1 overflows float32 in the reference too)
    if c["kind"] == "mass":
        a = np.stack([t.cpu().numpy() for t in prof.deriv(x=x, y=y, **p)])                 # (2, bs, N)
        r32, r64 = (np.transpose(v, (0, 2, 1)) for v in g32("deriv"))
        if key == "EDGE_SIE_round":
            ok = np.isfinite(r32).all(axis=(0, 2))
            a, r32, r64 = a[:, ok], r32[:, ok], r64[:, ok]
        assert_parity(a, fin(r32), r64, 1e-5, f"{key} deriv", axis=(0, 2))
        if c["cls"] != "DPIS" and not key.startswith("EDGE_"):
            # DPIS: the reference's analytic hessian is not the Jacobian of its deriv (piemd.py:72-73); the product returns the Jacobian
            h = np.stack([t.cpu().numpy() for t in prof.hessian(x=x, y=y, **p)])
            h32, h64 = g32("hessian")
            if h64.ndim == 2:      # Shear.hessian: bare (bs,) parameters
                h32, h64 = (np.broadcast_to(v[:, None, :], (4, len(x), v.shape[-1])) for v in (h32, h64))
            h32, h64 = (np.transpose(v, (0, 2, 1)) for v in (h32, h64))
            assert_parity(h, h32, h64, 1e-5, f"{key} hessian", axis=(0, 2))
    else:
        a = prof.light(x=x, y=y, **p).cpu().numpy()
        r32, r64 = g32("light")
        if r64.ndim == 3:          # use_lstsq: (D, N, bs) component stack
            r32, r64 = (np.transpose(v, (0, 2, 1)) for v in (r32, r64))          # (D, bs, N), the product's layout
            assert a.shape == r64.shape, (a.shape, r64.shape)
            assert_parity(a, r32, r64, 1e-5, f"{key} light components", axis=(0, 2))
        else:
            assert_parity(a, r32.T, r64.T, 1e-5, f"{key} light", axis=1)


def _matrix(cm, params, bs):
    mat = np.empty((max(1, cm.n_params), bs), dtype=np.float32)
    for i, (g, j, k) in enumerate(cm.slot_keys):
        mat[i] = params[g][j][k]
    return mat


def _perturbed_oracle(c):
    """The float64 oracle on inputs moved by half a float32 ulp: the conditioning probe of the parity rule."""
    cp = dict(c, params={g: [{k: common.ulp_perturb(v, seed=7 + i) for k, v in d.items()} for i, d in enumerate(c["params"][g])]
                         for g in c["params"]})
    return R.build_oracle_case(cp, torch.float64)


@pytest.mark.parametrize("key", sorted(R.SIM_CASES))
def test_cuda_simulator_and_likelihood_match_the_executed_reference(key):
    c = R.SIM_CASES[key]
    s = c["sim"]
    n = s["num_pix"]
    pm = common.spec_model(c["model"], c.get("constants"))
    keys = RC.grad_keys(c["params"])
    bs = len(c["params"][keys[0][0]][keys[0][1]][keys[0][2]])
    sc = SimulatorConfig(delta_pix=s["delta_pix"], num_pix=n, supersample=s["supersample"],
                         kernel=None if s["kernel"] is None else s["kernel"].astype(np.float32),
                         pix_region=None if s["pix_region"] is None else s["pix_region"].astype(np.float32),
                         transform_pix2angle=s.get("transform_pix2angle"))
    sim = LensSimulator(pm, sc, bs=bs)
    cm = sim.compiled
    mat = _matrix(cm, c["params"], bs)
    dev = torch.as_tensor(mat, device="cuda")
    params = cm.unflatten(dev)
    g = lambda q, t: GOLD[f"sim/{key}/{t}/{q}"].reshape(bs, n, n)   # noqa: E731
    _, osim_p, opm_p, params_p, leaves_p, _ = _perturbed_oracle(c)

    def img_check(got, q, pert):
        assert_parity(got.cpu().numpy().reshape(bs, n, n), g(q, "f32"), g(q, "f64"), 1e-5, f"{key} {q}",
                      pert.detach().numpy().reshape(bs, n, n), axis=(1, 2))

    img_check(sim.simulate(params), "image", osim_p.simulate(params_p))
    if not c.get("image_and_likelihood_only"):
        img_check(sim.simulate(params, no_deflection=True), "image_no_deflection", osim_p.simulate(params_p, no_deflection=True))
    if c.get("variants"):
        img_check(sim.simulate_source(params), "source", osim_p.simulate_source(params_p))
        img_check(sim.simulate_lens_light(params), "lens_light", osim_p.simulate_lens_light(params_p))
        img_check(sim.simulate_images(params), "images", osim_p.simulate_images(params_p))
    # pixel likelihood and its gradient (tf/model.py:89-101 + autodiff through the reference's simulate)
    kw = dict(error_map=c["error_map"].astype(np.float32)) if "error_map" in c else dict(background_rms=c["noise"]["background_rms"],
                                                                                         exp_time=c["noise"]["exp_time"])
    pmod = ForwardProbModel({"lens_mass": []}, c["observed"].astype(np.float32), **kw)
    ll, chi2, grad = (t.cpu().numpy() for t in pmod.loglike_and_grad(sim, dev))
    ll_p, chi2_p = opm_p.stats_pixels(osim_p, params_p)
    gp = torch.autograd.grad(ll_p.sum(), leaves_p, allow_unused=True)
    def G(q, t):      # bs = 1: the reference squeezes the batch axis away (tf/simulator.py:156) -- put it back
        v = GOLD[f"sim/{key}/{t}/{q}"]
        return v.reshape((bs,) if v.ndim == 0 else v.shape)
    assert_parity(ll[:, None], G("loglike", "f32")[:, None], G("loglike", "f64")[:, None], 1e-5, f"{key} log-like",
                  ll_p.detach().numpy()[:, None], axis=1)
    assert_parity(chi2[:, None], G("red_chi2", "f32")[:, None], G("red_chi2", "f64")[:, None], 1e-5, f"{key} red chi2",
                  chi2_p.detach().numpy()[:, None], axis=1)
    row = {k: i for i, k in enumerate(cm.slot_keys)}
    for j, k in enumerate(keys):
        pert = np.zeros(bs) if gp[j] is None else gp[j].numpy()
        assert_parity(grad[row[k]], G("grad", "f32")[j], G("grad", "f64")[j], 1e-4, f"{key} grad {k}", pert)
    if c.get("image_and_likelihood_only"):
        return
    # points
    px, py = (v.astype(np.float32) for v in GOLD[f"sim/{key}/points"])
    lens = params["lens_mass"]
    b = np.stack([t.cpu().numpy() for t in sim.beta(px, py, lens)])                       # (2, bs, npts)
    assert_parity(b, np.transpose(G("beta", "f32"), (0, 2, 1)), np.transpose(G("beta", "f64"), (0, 2, 1)), 1e-5, f"{key} beta", axis=(0, 2))
    if key != "nopsf_ss1":      # contains DPIS: see the profile test above
        kap = sim.convergence(px, py, lens).cpu().numpy()
        assert_parity(kap, G("convergence", "f32").T, G("convergence", "f64").T, 1e-5, f"{key} convergence", axis=1)
        sh = np.stack([t.cpu().numpy() for t in sim.shear(px, py, lens)])
        assert_parity(sh, np.transpose(G("shear", "f32"), (0, 2, 1)), np.transpose(G("shear", "f64"), (0, 2, 1)), 1e-5, f"{key} shear", axis=(0, 2))
        mu, mu32, mu64 = sim.magnification(px, py, lens).cpu().numpy(), G("magnification", "f32").T, G("magnification", "f64").T
        ok = np.abs(mu64) < 50      # away from critical curves, where 1/det amplifies the float32 rounding of the Hessian
        assert np.allclose(mu[ok], mu64[ok], rtol=max(2e-4, 4 * float(np.max(np.abs(mu32[ok] / mu64[ok] - 1)))))
    cen = c.get("centroids")
    if cen is not None:         # image-position likelihood (tf/model.py:103-124), positions alone
        f = lambda v: [a.astype(np.float32) for a in v]   # noqa: E731
        ppos = ForwardProbModel({"lens_mass": []}, centroids_x=f(cen["x"]), centroids_y=f(cen["y"]), centroids_errors_x=f(cen["ex"]),
                                centroids_errors_y=f(cen["ey"]), include_pixels=False)
        llp, chip, gpos = (t.cpu().numpy() for t in ppos.loglike_and_grad(sim, dev))
        o_ll, o_chi = opm_p.stats_positions(osim_p, params_p)
        o_g = torch.autograd.grad(o_ll.sum(), leaves_p, allow_unused=True)
        assert_parity(llp[:, None], G("pos_loglike", "f32")[:, None], G("pos_loglike", "f64")[:, None], 1e-5, f"{key} positions log-like",
                      o_ll.detach().numpy()[:, None], axis=1)
        assert_parity(chip[:, None], G("pos_red_chi2", "f32")[:, None], G("pos_red_chi2", "f64")[:, None], 1e-5, f"{key} positions red chi2",
                      o_chi.detach().numpy()[:, None], axis=1)
        for j, k in enumerate(keys):
            if k[0] != "lens_mass":
                continue
            pert = np.zeros(bs) if o_g[j] is None else o_g[j].numpy()
            assert_parity(gpos[row[k]], G("pos_grad", "f32")[j], G("pos_grad", "f64")[j], 1e-4, f"{key} positions grad {k}", pert)
        # both terms together: log-likes add, the reduced chi^2 is the mean of the included terms (tf/model.py:150-163)
        pboth = ForwardProbModel({"lens_mass": []}, c["observed"].astype(np.float32), centroids_x=f(cen["x"]), centroids_y=f(cen["y"]),
                                 centroids_errors_x=f(cen["ex"]), centroids_errors_y=f(cen["ey"]), include_pixels=True, **kw)
        llb, chib, _ = (t.cpu().numpy() for t in pboth.loglike_and_grad(sim, dev))
        o_llb, o_chib = opm_p._stats(osim_p, params_p)
        assert_parity(llb[:, None], G("loglike_total", "f32")[:, None], G("loglike_total", "f64")[:, None], 1e-5, f"{key} log-like (both terms)",
                      o_llb.detach().numpy()[:, None], axis=1)
        assert_parity(chib[:, None], G("red_chi2_total", "f32")[:, None], G("red_chi2_total", "f64")[:, None], 1e-5,
                      f"{key} red chi2 (mean of both terms)", o_chib.detach().numpy()[:, None], axis=1)


def test_cuda_lstsq_simulate_matches_the_executed_reference():
    """lstsq_simulate through the C ABI against the reference's own simulator object, beta, light calls and source lines :204-240
    (tests/golden/make_reference_golden.py run_lstsq): component stack, amplitudes and recombined image."""
    c, pm, osim_p, params_p, obs, err = R.lstsq_oracle(torch.float64)
    s = c["sim"]
    n, bs = s["num_pix"], 2
    sim = LensSimulator(pm, SimulatorConfig(delta_pix=s["delta_pix"], num_pix=n, supersample=s["supersample"],
                                            kernel=s["kernel"].astype(np.float32)), bs=bs)
    cm = sim.compiled
    params = cm.unflatten(torch.as_tensor(_matrix(cm, c["params"], bs), device="cuda"))
    o, e = c["observed"].astype(np.float32), c["err_map"].astype(np.float32)
    stack = sim.lstsq_simulate(params, o, e, return_stacked=True).cpu().numpy()
    coef = sim.lstsq_simulate(params, o, e, return_coeffs=True).cpu().numpy()
    img = sim.lstsq_simulate(params, o, e).cpu().numpy().reshape(bs, n, n)
    pert = {g: [{k: torch.as_tensor(common.ulp_perturb(v.numpy(), seed=5 + i)) for k, v in d.items()} for i, d in enumerate(params_p[g])]
            for g in params_p}
    with torch.no_grad():
        sp = osim_p.lstsq_simulate(pert, obs, err, return_stacked=True).numpy()
        cp = osim_p.lstsq_simulate(pert, obs, err, return_coeffs=True).numpy()
        ip = osim_p.lstsq_simulate(pert, obs, err).numpy().reshape(bs, n, n)
    G = lambda q, t: GOLD[f"lstsq/{t}/{q}"]   # noqa: E731
    assert stack.shape == G("stack", "f64").shape
    assert_parity(stack, G("stack", "f32"), G("stack", "f64"), 1e-5, "lstsq stack", sp, axis=(1, 2, 3))
    assert_parity(coef, G("coeffs", "f32"), G("coeffs", "f64"), 1e-4, "lstsq coefficients", cp, axis=1)
    assert_parity(img, G("image", "f32").reshape(bs, n, n), G("image", "f64").reshape(bs, n, n), 1e-5, "lstsq image", ip, axis=(1, 2))
