"""CPU tests of the device arithmetic: gl_math.cuh / gl_program.h compiled by g++ into the
test-only host harness (tests/hostcheck) and compared with the oracle's autograd.  fp64 pins
every formula and hand adjoint to ~1e-12; fp32 shows the float instantiation the kernels use
stays within the fp32 oracle's own rounding.  (This harness is not a product path.)"""
import numpy as np
import pytest
import torch

import common
from common import CompiledModel, PhysicalModel, draw_matrix, host_run, matrix_to_pytree, to_oracle_model
from gigalens_b200.profiles.light import sersic, shapelets
from gigalens_b200.profiles.mass import dpie_subhalo, epl, nfw, piemd, piep, shear, sie, sis, tnfw
from oracle.simulator import OracleSimulator

SRC = lambda: [sersic.SersicEllipse()]


def _catalogue(G=5, seed=7):
    rng = np.random.default_rng(seed)
    e = rng.normal(0, 0.1, size=(2, G))
    return dict(lum=rng.lognormal(0, 0.5, G).tolist(), center_x=rng.uniform(-1, 1, G).tolist(),
                center_y=rng.uniform(-1, 1, G).tolist(), e1=(e[0] + 0.05).tolist(), e2=(e[1] - 0.04).tolist())


MODELS = {
    "c2": lambda: PhysicalModel([epl.EPL(50), shear.Shear()], [sersic.SersicEllipse()], SRC()),
    "sis": lambda: PhysicalModel([sis.SIS()], [sersic.Sersic()], SRC()),
    "sie": lambda: PhysicalModel([sie.SIE()], [], SRC()),
    "nfw": lambda: PhysicalModel([nfw.NFW()], [], SRC()),
    "nfw_ellipse": lambda: PhysicalModel([nfw.NFW_ELLIPSE()], [], SRC()),
    "dpis": lambda: PhysicalModel([piemd.DPIS()], [], SRC()),
    "dpie": lambda: PhysicalModel([piemd.DPIE()], [], SRC()),
    "tnfw": lambda: PhysicalModel([tnfw.TNFW()], [], SRC()),
    "dpiep": lambda: PhysicalModel([piep.DPIEP()], [], SRC()),
    "constants": lambda: PhysicalModel([epl.EPL(20), shear.Shear()], [], [sersic.Sersic()],
                                       lenses_constants=[{"gamma": 2.1, "center_x": 0.05}, {}],
                                       source_light_constants=[{"n_sersic": 1.5}]),
    # strongly elliptical lens with a short iteration cap: the truncated series differs from the infinite one at the
    # 1e-2 level, so the closed-form f-derivative needs its truncation term R_N w^N (gl_math.cuh, EPL)
    "epl_capped": lambda: PhysicalModel([epl.EPL(12)], [], SRC(), lenses_constants=[{"e1": 0.45, "e2": -0.3}]),
    "shapelets": lambda: PhysicalModel([epl.EPL(30), shear.Shear()], [sersic.Sersic()], [shapelets.Shapelets(4, interpolate=False)]),
    "shapelets_interp": lambda: PhysicalModel([sis.SIS()], [], [shapelets.Shapelets(5, interpolate=True)]),
    "cluster": lambda: PhysicalModel([nfw.NFW(), dpie_subhalo.DPIESubhalo(1.0, _catalogue()), shear.Shear()], [], SRC()),
    # the reference's core-Sersic expression as written (sersic.py:83-132), as lens light and as source (10 raw parameters, 9 dvars)
    "core_sersic": lambda: PhysicalModel([sis.SIS(), shear.Shear()], [sersic.CoreSersic()], [sersic.CoreSersic()]),
}


def _reference(pm, mat, dtype, seed, num_pix=12, delta=None):
    if delta is None:  # Shapelets live on the beta ~ 0.1 scale; the other models on the arcsecond scale
        delta = 0.08 if any(type(p).__name__ == "Shapelets" for p in pm.source_light) else 0.25
    bs = mat.shape[1]
    cm = CompiledModel(pm)
    sim = OracleSimulator(to_oracle_model(pm, dtype), delta, num_pix, 1, bs=bs, dtype=dtype)
    params, leaf = matrix_to_pytree(cm, mat, dtype, True)
    S = sim.simulate_ss(params).reshape(-1, bs).T
    G = np.random.default_rng(seed + 1).normal(size=tuple(S.shape))
    (S * torch.as_tensor(G, dtype=dtype)).sum().backward()
    return cm, sim.img_X[:, 0].numpy(), sim.img_Y[:, 0].numpy(), S.detach().numpy(), G, leaf.grad.numpy()


@pytest.mark.parametrize("name", sorted(MODELS))
def test_formulas_and_adjoints_fp64(name):
    pm = MODELS[name]()
    mat = draw_matrix(CompiledModel(pm), 3, seed=1)
    cm, gx, gy, S, G, gref = _reference(pm, mat, torch.float64, 1)
    out = host_run(cm, mat, gx, gy, g_ss=G, dtype=np.float64)
    # the reference rounds the Shapelets prefactor table to fp32 (shapelets.py:47-48); the kernels use the
    # exact normalised recurrence, so that one model agrees to fp32-constant precision only
    t_img, t_g = (1e-7, 5e-5) if name == "shapelets" else (1e-11, 1e-9)
    assert np.max(np.abs(out["ss"] - S)) / np.max(np.abs(S)) < t_img
    for k in range(cm.n_params):
        assert np.max(np.abs(out["gparams"][k] - gref[k])) / np.max(np.abs(gref[k])) < t_g, cm.slot_keys[k]


@pytest.mark.parametrize("name", sorted(MODELS))
def test_float_instantiation_tracks_fp32_oracle(name):
    pm = MODELS[name]()
    mat = draw_matrix(CompiledModel(pm), 3, seed=2).astype(np.float32)
    cm, gx, gy, S64, G, g64 = _reference(pm, mat.astype(np.float64), torch.float64, 2)
    _, _, _, S32, _, g32 = _reference(pm, mat.astype(np.float64), torch.float32, 2)
    _, _, _, S64p, _, g64p = _reference(pm, common.ulp_perturb(mat), torch.float64, 2)
    out = host_run(cm, mat, gx, gy, g_ss=G, dtype=np.float32)
    common.assert_parity(out["ss"], S32, S64, 1e-5, "ss image", S64p, axis=1)
    for k in range(cm.n_params):
        common.assert_parity(out["gparams"][k], g32[k], g64[k], 1e-4, f"grad {cm.slot_keys[k]}", g64p[k])


def test_epl_per_sample_trip_count_equals_batch_max():
    """Per-sample series length (the CUDA default) vs the reference's batch-global length."""
    pm = MODELS["c2"]()
    mat = draw_matrix(CompiledModel(pm), 6, seed=3)
    cm, gx, gy, S, G, gref = _reference(pm, mat, torch.float64, 3)
    a = host_run(cm, mat, gx, gy, g_ss=G, dtype=np.float64, epl_batch_max=True)
    b = host_run(cm, mat, gx, gy, g_ss=G, dtype=np.float64, epl_batch_max=False)
    assert np.max(np.abs(a["ss"] - b["ss"])) / np.max(np.abs(S)) < 1e-10
    assert np.max(np.abs(a["gparams"] - b["gparams"]) / np.max(np.abs(gref), axis=1, keepdims=True)) < 1e-9


def test_beta_matches_oracle():
    pm = MODELS["c2"]()
    cm = CompiledModel(pm)
    mat = draw_matrix(cm, 2, seed=4)
    sim = OracleSimulator(to_oracle_model(pm, torch.float64), 0.25, 12, 1, bs=2, dtype=torch.float64)
    params, _ = matrix_to_pytree(cm, mat, torch.float64)
    bx, by = sim.beta(sim.img_X, sim.img_Y, params["lens_mass"])
    out = host_run(cm, mat, sim.img_X[:, 0].numpy(), sim.img_Y[:, 0].numpy(), dtype=np.float64, want_beta=True)
    assert np.allclose(out["beta"][:, 0], bx.numpy().T, atol=1e-12) and np.allclose(out["beta"][:, 1], by.numpy().T, atol=1e-12)


@pytest.mark.parametrize("interpolate", [False, True])
def test_lstsq_component_stack(interpolate):
    """Unit-amplitude linear components (the lstsq stack before the convolution) vs the oracle."""
    pm = PhysicalModel([epl.EPL(30), shear.Shear()], [sersic.SersicEllipse(use_lstsq=True)],
                       [shapelets.Shapelets(5, use_lstsq=True, interpolate=interpolate)])
    cm = CompiledModel(pm)
    assert cm.depth == 1 + 21
    mat = draw_matrix(cm, 2, seed=3)
    sim = OracleSimulator(to_oracle_model(pm, torch.float64), 0.08, 12, 1, bs=2, dtype=torch.float64)
    params, _ = matrix_to_pytree(cm, mat, torch.float64)
    st = sim.lstsq_stack(params)   # (bs, ny, nx, D); no PSF and ss = 1 -> the raw components
    ref = st.reshape(2, -1, st.shape[-1]).permute(0, 2, 1).numpy()
    out = host_run(cm, mat, sim.img_X[:, 0].numpy(), sim.img_Y[:, 0].numpy(), want_comps=True)
    assert np.max(np.abs(out["comps"] - ref)) / np.max(np.abs(ref)) < (1e-12 if interpolate else 1e-7)


@pytest.mark.parametrize("name", ["c2", "constants", "cluster", "nfw", "dpie"])
def test_packed_lane_instantiation_matches_scalar(name):
    """The two-pixel packed lane type (what k_raytrace_*_p instantiate) runs the same source as the
    scalar lanes; on the host both are plain fp32, so results must agree to rounding."""
    pm = MODELS[name]()
    mat = draw_matrix(CompiledModel(pm), 3, seed=5).astype(np.float32)
    cm, gx, gy, S64, G, g64 = _reference(pm, mat.astype(np.float64), torch.float64, 5)
    a = host_run(cm, mat, gx, gy, g_ss=G, dtype=np.float32)
    b = common.host_run_packed(cm, mat, gx, gy, g_ss=G)
    assert np.max(np.abs(a["ss"] - b["ss"])) <= 1e-6 * np.max(np.abs(a["ss"]))
    for k in range(cm.n_params):
        sc = np.max(np.abs(g64[k]))
        assert np.max(np.abs(a["gparams"][k] - b["gparams"][k])) <= 2e-5 * sc, cm.slot_keys[k]
        assert np.max(np.abs(b["gparams"][k] - g64[k])) <= max(1e-4, 10 * np.max(np.abs(a["gparams"][k] - g64[k]))) * sc
    if name == "c2":   # the straight-line drivers of the benchmark shape: same blocks, same operation order => identical
        c = common.host_run_packed(cm, mat, gx, gy, g_ss=G, straight_line=True)
        assert np.array_equal(b["ss"], c["ss"]) and np.array_equal(b["gparams"], c["gparams"])


# ---------------------------------------------------------------------------------------------
# image-position likelihood: Hessian by forward-mode duals, gradient by duals through the hand adjoint
# ---------------------------------------------------------------------------------------------
def _systems(seed=5):
    rng = np.random.default_rng(seed)
    out = []
    for n in (4, 2, 3):
        r, t = rng.uniform(0.7, 1.6, n), rng.uniform(0, 2 * np.pi, n)
        out.append((r * np.cos(t), r * np.sin(t), rng.uniform(0.01, 0.03, n), rng.uniform(0.01, 0.03, n)))
    return out


def _positions_reference(pm, mat, systems, dtype, analytic=True):
    """Oracle stats_positions (tf/model.py:103-124) + autograd; analytic=False swaps every profile's hessian for
    the autodiff default of tf/profile.py:9-30."""
    from oracle import model as OM, profiles as OP
    bs = mat.shape[1]
    cm = CompiledModel(pm)
    om = to_oracle_model(pm, dtype)
    if not analytic:
        for lens in om.lenses:
            lens.hessian = OP.MassBase.hessian.__get__(lens)
            inner = getattr(lens, "profile", None)
            if inner is not None:
                inner.hessian = OP.MassBase.hessian.__get__(inner)
    sim = OracleSimulator(om, 0.25, 4, 1, bs=bs, dtype=dtype)
    prob = OM.ForwardProbModel(None, include_pixels=False, include_positions=True, dtype=dtype,
                               centroids_x=[s[0] for s in systems], centroids_y=[s[1] for s in systems],
                               centroids_errors_x=[s[2] for s in systems], centroids_errors_y=[s[3] for s in systems])
    for k in ("centroids_x", "centroids_y", "centroids_errors_x", "centroids_errors_y"):  # keep the fp64 inputs un-rounded
        setattr(prob, k, [torch.as_tensor(np.asarray(s[i], dtype=np.float64)).to(dtype) for s in systems]
                if (i := ["centroids_x", "centroids_y", "centroids_errors_x", "centroids_errors_y"].index(k)) >= 0 else None)
    prob.init_centroids(bs)
    params, leaf = matrix_to_pytree(cm, mat, dtype, True)
    ll, chi2 = prob.stats_positions(sim, params)
    ll.sum().backward()
    X = torch.cat(prob.centroids_x_batch, 0)
    Y = torch.cat(prob.centroids_y_batch, 0)
    params2, _ = matrix_to_pytree(cm, mat, dtype)
    H = [h.detach().numpy().T for h in sim.hessian(X, Y, params2["lens_mass"])]
    return dict(loglike=ll.detach().numpy(), chi2=chi2.detach().numpy(), gparams=leaf.grad.numpy(), hess=np.stack(H, 1))


POS_MODELS = ["c2", "sis", "sie", "nfw", "nfw_ellipse", "dpie", "tnfw", "dpiep", "constants", "cluster"]


@pytest.mark.parametrize("name", POS_MODELS)
@pytest.mark.parametrize("fwdmode", [True, False])
def test_positions_likelihood_and_gradient_fp64(name, fwdmode):
    if not fwdmode and name != "cluster":
        pytest.skip("forward-mode groups exist only in scaling-relation models")
    pm = MODELS[name]()
    cm = CompiledModel(pm)
    mat = draw_matrix(cm, 3, seed=6)
    systems = _systems()
    ref = _positions_reference(pm, mat, systems, torch.float64)
    out = common.host_positions(cm, mat, systems, np.float64, use_fwdmode=fwdmode)
    assert np.max(np.abs(out["hess"] - ref["hess"])) < 1e-9 * max(1.0, np.max(np.abs(ref["hess"])))
    assert np.allclose(out["loglike"], ref["loglike"], rtol=1e-9) and np.allclose(out["chi2"], ref["chi2"], rtol=1e-9)
    lens_rows = [k for k, key in enumerate(cm.slot_keys) if key[0] == "lens_mass"]
    for k in lens_rows:
        assert np.max(np.abs(out["gparams"][k] - ref["gparams"][k])) <= 1e-8 * np.max(np.abs(ref["gparams"][k])), cm.slot_keys[k]
    other = [k for k in range(cm.n_params) if k not in lens_rows]
    assert np.all(out["gparams"][other] == 0)


def test_positions_dpis_follows_the_jacobian_of_deriv():
    """The reference's analytic DPIS.hessian (piemd.py:62-83) carries a factor (r_core + r_cut)/r_cut on kappa
    that its own deriv (and DPIE.hessian in the round limit) does not have; the CUDA path differentiates deriv."""
    pm = MODELS["dpis"]()
    cm = CompiledModel(pm)
    mat = draw_matrix(cm, 3, seed=6)
    systems = _systems()
    out = common.host_positions(cm, mat, systems, np.float64)
    ad = _positions_reference(pm, mat, systems, torch.float64, analytic=False)
    an = _positions_reference(pm, mat, systems, torch.float64, analytic=True)
    assert np.max(np.abs(out["hess"] - ad["hess"])) < 1e-10
    assert np.allclose(out["loglike"], ad["loglike"], rtol=1e-9)
    assert np.max(np.abs(out["gparams"] - ad["gparams"])) <= 1e-8 * np.max(np.abs(ad["gparams"]))
    # the analytic version differs in kappa only: f_xx + f_yy scaled, f_xx - f_yy and f_xy equal
    assert np.allclose(an["hess"][:, 0] - an["hess"][:, 3], ad["hess"][:, 0] - ad["hess"][:, 3], atol=1e-12)
    assert np.max(np.abs((an["hess"][:, 0] + an["hess"][:, 3]) - (ad["hess"][:, 0] + ad["hess"][:, 3]))) > 1e-4


@pytest.mark.parametrize("name", ["c2"])
def test_positions_float_instantiation(name):
    """GlDual<float> compiles and tracks the fp32 oracle; the CUDA kernels instantiate GlDual<double> (the cluster
    model's complex-log dPIE form loses the 1e-5 bound in fp32 near critical curves), which the fp64 tests pin."""
    pm = MODELS[name]()
    cm = CompiledModel(pm)
    mat = draw_matrix(cm, 3, seed=6).astype(np.float32)
    systems = [tuple(np.asarray(a, dtype=np.float32) for a in s) for s in _systems()]
    ref = _positions_reference(pm, mat.astype(np.float64), [tuple(a.astype(np.float64) for a in s) for s in systems], torch.float64)
    sys64 = [tuple(a.astype(np.float64) for a in s) for s in systems]
    r32 = _positions_reference(pm, mat.astype(np.float64), sys64, torch.float32)
    rpt = _positions_reference(pm, common.ulp_perturb(mat), sys64, torch.float64)
    out = common.host_positions(cm, mat, systems, np.float32)
    # random image positions land near critical curves (|mu| ~ 1e2), which amplifies fp32 rounding: same
    # metric as everywhere else -- 1e-5 / 1e-4, or a small multiple of the fp32 oracle's own error
    common.assert_parity(out["loglike"][:, None], r32["loglike"][:, None], ref["loglike"][:, None], 1e-5, "positions log-like",
                         rpt["loglike"][:, None], axis=1)
    lens_rows = [k for k, key in enumerate(cm.slot_keys) if key[0] == "lens_mass"]
    for k in lens_rows:
        common.assert_parity(out["gparams"][k], r32["gparams"][k], ref["gparams"][k], 1e-4, f"positions grad {cm.slot_keys[k]}",
                             rpt["gparams"][k])


def test_nfw_branch_free_fp32_form_accuracy_and_gradient():
    """The fp32 lanes evaluate g(X) of nfw.py:34-52 branch-free (gl_math.cuh nfw_g_fast: three regimes in u = (1-X)/(1+X)).
    Deflection along a radius from X = 1e-4 to 10, dense around X = 1, against the fp64 oracle: the float instantiation must be
    at least as accurate as the fp32 oracle (the libm transcription of the reference) everywhere and hold 5e-6 outright --
    for small X the transcription cancels two O(|ln X|) terms and is off by 1e-3 -- and the adjoint must agree with fp64 autograd."""
    pm = PhysicalModel([nfw.NFW()], [], [sersic.Sersic()])
    cm = CompiledModel(pm)
    mat = draw_matrix(cm, 1, seed=1)
    i_rs = [i for i, k in enumerate(cm.slot_keys) if k[2] == "Rs"][0]
    i_cx = [i for i, k in enumerate(cm.slot_keys) if k[0] == "lens_mass" and k[2] == "center_x"][0]
    i_cy = [i for i, k in enumerate(cm.slot_keys) if k[0] == "lens_mass" and k[2] == "center_y"][0]
    mat[i_rs], mat[i_cx], mat[i_cy] = 2.0, 0.0, 0.0
    mat = mat.astype(np.float32).astype(np.float64)
    X = np.concatenate([np.logspace(-4, 1, 400), 1 + np.logspace(-7, -1, 60), 1 - np.logspace(-7, -1, 60), [0.818, 0.819, 1.222, 1.223]])
    ang = np.linspace(0.1, 6.0, X.size)
    gx, gy = (2.0 * X * np.cos(ang)).astype(np.float32).astype(np.float64), (2.0 * X * np.sin(ang)).astype(np.float32).astype(np.float64)

    def oracle_alpha(dt):
        om = to_oracle_model(pm, dt)
        p, _ = matrix_to_pytree(cm, mat, dt)
        ax, ay = om.lenses[0].deriv(torch.as_tensor(gx, dtype=dt)[:, None], torch.as_tensor(gy, dtype=dt)[:, None], **p["lens_mass"][0])
        return ax[:, 0].double().numpy(), ay[:, 0].double().numpy()

    a64, a32 = oracle_alpha(torch.float64), oracle_alpha(torch.float32)
    out = host_run(cm, mat, gx, gy, dtype=np.float32, want_beta=True)
    ax, ay = gx - out["beta"][0, 0], gy - out["beta"][0, 1]
    mag = np.hypot(*a64)
    err = np.hypot(ax - a64[0], ay - a64[1]) / mag
    err32 = np.hypot(a32[0] - a64[0], a32[1] - a64[1]) / mag
    # beta = x - alpha is rounded to fp32 in the harness: allow its ulp on top
    ulp = 1.2e-7 * np.hypot(gx, gy) / mag
    assert np.all(err <= 5e-6 + 2 * ulp), (float(err.max()), float(X[np.argmax(err)]))
    assert np.median(err) < 5e-7
    assert err32.max() > 1e-4            # the reference's own fp32 form loses digits at small X ...
    assert np.all(err <= np.maximum(3 * err32, 5e-6 + 2 * ulp))   # ... and the branch-free form is never meaningfully worse
    packed = common.host_run_packed(cm, mat.astype(np.float32), gx, gy)
    assert np.max(np.abs(packed["ss"] - host_run(cm, mat, gx, gy, dtype=np.float32)["ss"])) <= 1e-6 * np.max(np.abs(packed["ss"]))
