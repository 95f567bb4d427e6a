"""CPU tests pinning the oracle (no GPU): the framework-free known-answer tests recoverable from
the reference's own tests (tests/test_profiles.py, tests/tf/test_model.py), cross-profile
identities, the demo.npy image-level pin, golden vectors, and fp64 finite differences."""
import math
import os

import numpy as np
import pytest
import torch

import common
import oracle_bridge
from gigalens_b200 import workloads
from oracle import model as OM
from oracle import profiles as OP
from oracle.simulator import OracleSimulator, subgrid_kernel

D = torch.float64
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pts(n=2000, seed=0, dtype=D):
    rng = np.random.default_rng(seed)
    return torch.as_tensor(rng.normal(size=n), dtype=dtype), torch.as_tensor(rng.normal(size=n), dtype=dtype)


def test_sersic_half_light_identity():
    # reference tests/test_profiles.py:17-26
    se = OP.SersicEllipse()
    a = se.light(torch.tensor(0.0, dtype=D), torch.tensor(1.0, dtype=D), R_sersic=1.0, n_sersic=2.0, e1=0.0, e2=0.0,
                 center_x=0.0, center_y=0.0, Ie=5.0)
    assert math.isclose(float(a), 5.0)


def test_epl_isothermal_round_is_sis():
    # tests/test_profiles.py:55-58 with :71-74 : EPL(theta_E=1, gamma=2, e=0) == SIS == (x/r, y/r)
    x, y = pts()
    ax, ay = OP.EPL(100).deriv(x, y, theta_E=1.0, gamma=2.0, e1=0.0, e2=0.0, center_x=0.0, center_y=0.0)
    sx, sy = OP.SIS().deriv(x, y, theta_E=1.0, center_x=0.0, center_y=0.0)
    r = torch.sqrt(x ** 2 + y ** 2)
    assert torch.allclose(ax, sx, rtol=1e-5, atol=1e-4) and torch.allclose(ay, sy, rtol=1e-5, atol=1e-4)
    assert torch.allclose(sx, x / r) and torch.allclose(sy, y / r)


def test_epl_gamma2_equals_sie():
    # both use b = theta_E sqrt(q): the framework-free test of the EPL series with e != 0
    x, y = pts()
    kw = dict(theta_E=1.2, e1=0.1, e2=-0.1, center_x=0.0, center_y=0.0)
    ax, ay = OP.EPL(100).deriv(x, y, gamma=2.0, **kw)
    sx, sy = OP.SIE().deriv(x, y, **kw)
    assert torch.allclose(ax, sx, rtol=0, atol=1e-12) and torch.allclose(ay, sy, rtol=0, atol=1e-12)


def test_sie_small_e_is_sis():
    # tests/test_profiles.py:87-90
    x, y = pts()
    ax, ay = OP.SIE().deriv(x, y, theta_E=1.0, e1=1e-3, e2=1e-3, center_x=0.0, center_y=0.0)
    sx, sy = OP.SIS().deriv(x, y, theta_E=1.0, center_x=0.0, center_y=0.0)
    assert torch.allclose(ax, sx, rtol=1e-2, atol=1e-2) and torch.allclose(ay, sy, rtol=1e-2, atol=1e-2)


def test_shear_closed_form():
    # tests/test_profiles.py:103-111
    x, y = pts()
    ax, ay = OP.Shear().deriv(x, y, gamma1=0.1, gamma2=0.1)
    assert torch.allclose(ax, 0.1 * x + 0.1 * y) and torch.allclose(ay, 0.1 * x - 0.1 * y)
    ax, ay = OP.Shear().deriv(x, y, gamma1=0.0, gamma2=0.0)
    assert float(ax.abs().max()) == 0 and float(ay.abs().max()) == 0


def test_shapelets_interpolate_matches_recurrence():
    # tests/test_profiles.py:36-47: both modes are compared with the same external answer
    rng = np.random.default_rng(1)
    x = torch.as_tensor(rng.normal(size=(5, 5, 1)), dtype=torch.float32)
    y = torch.as_tensor(rng.normal(size=(5, 5, 1)), dtype=torch.float32)
    a, b = OP.Shapelets(5, interpolate=True), OP.Shapelets(5, interpolate=False)
    amps = {n: torch.as_tensor(rng.normal(size=1), dtype=torch.float32) for n in a.amp_names}
    va = a.light(x, y, center_x=0.0, center_y=0.0, beta=1.0, **amps)
    vb = b.light(x, y, center_x=0.0, center_y=0.0, beta=1.0, **amps)
    assert np.allclose(va.numpy(), vb.numpy(), rtol=1e-5, atol=1e-4)
    # orthonormality of the basis: integral phi_n phi_m = delta_nm
    g = np.linspace(-12, 12, 48001)
    for n in range(6):
        for m in range(6):
            v = np.trapezoid(OP.shapelet_phi_n_np(n, g) * OP.shapelet_phi_n_np(m, g), g)
            assert abs(v - (n == m)) < 1e-9


def test_nfw_ellipse_round_is_nfw():
    x, y = pts()
    kw = dict(Rs=1.3, alpha_Rs=0.9, center_x=0.05, center_y=-0.02)
    ax, ay = OP.NFW_ELLIPSE().deriv(x, y, e1=1e-9, e2=0.0, **kw)
    sx, sy = OP.NFW().deriv(x, y, **kw)
    assert torch.allclose(ax, sx, atol=1e-7) and torch.allclose(ay, sy, atol=1e-7)


def test_dpie_small_e_tends_to_dpis():
    x, y = pts()
    kw = dict(theta_E=1.0, r_core=0.05, r_cut=3.0, center_x=0.0, center_y=0.0)
    sx, sy = OP.DPIS().deriv(x, y, **kw)
    for e in (1e-3, 1e-4):
        ax, ay = OP.DPIE().deriv(x, y, e1=e, e2=0.0, **kw)
        assert float((ax - sx).abs().max()) < e and float((ay - sy).abs().max()) < e


def test_scaling_relation_single_member_is_dpie():
    x, y = pts(500)
    x, y = x[:, None].repeat(1, 2), y[:, None].repeat(1, 2)
    cat = dict(lum=[1.0], center_x=[0.3], center_y=[-0.2], e1=[0.1], e2=[0.05])
    sr = OP.DPIESubhalo(1.0, cat, dtype=D)
    th, rc, rt = (torch.tensor(v, dtype=D) for v in ([0.8, 0.9], [0.05, 0.06], [5.0, 4.0]))
    ax, ay = sr.deriv(x, y, theta_E=th, r_core=rc, r_cut=rt)
    f = lambda v: torch.tensor(np.float32(v), dtype=D)
    bx, by = OP.DPIE().deriv(x, y, theta_E=th, r_core=rc, r_cut=rt, e1=f(0.1), e2=f(0.05), center_x=f(0.3), center_y=f(-0.2))
    assert torch.allclose(ax, bx, atol=1e-12) and torch.allclose(ay, by, atol=1e-12)


def test_subgrid_kernel_checkpoint():
    k = subgrid_kernel(workloads.load_psf(), 2, odd=True)
    assert k.shape == (25, 25) and abs(k.sum() - 1) < 1e-12 and abs(k.max() - 0.18792085) < 1e-7


def test_demo_image_level_pin():
    # tf-demo.ipynb cells 5-9: chi^2 of simulate(truth) against the reference asset demo.npy must be ~1;
    # wrong conventions are rejected (un-flipped PSF 1.08, no PSF 1.49, transposed grid 21.9 -- SURVEY §8c)
    wl = workloads.c2_workload()
    sim, _ = oracle_bridge.build_oracle(wl, 1, torch.float32)
    t = {k: [{kk: torch.tensor([vv], dtype=torch.float32) for kk, vv in d.items()} for d in v]
         for k, v in workloads.DEMO_TRUTH.items()}
    im = sim.simulate(t).numpy()
    err = np.sqrt(0.2 ** 2 + im / 100)
    chi2 = np.mean(((im - wl["observed"]) / err) ** 2)
    assert 0.93 < chi2 < 1.03
    assert abs(im.max() - 44.0347) < 5e-3


def test_bijector_round_trip_and_fldj_shape():
    # tests/tf/test_model.py:10-26
    prior = oracle_bridge.to_oracle_prior(workloads.demo_prior())
    x = torch.as_tensor(prior.sample(5, seed=0), dtype=D)
    z = prior.inverse(x)
    _, leaves = prior.forward(z)
    assert torch.allclose(torch.stack(leaves, 1), x, rtol=1e-10)
    assert prior.log_prior(z).shape == (5,)
    # flatten order: dict keys sorted (lens_light < lens_mass < source_light; inner names sorted)
    assert [p for p, _ in prior.leaves][:3] == [("lens_light", 0, "Ie"), ("lens_light", 0, "R_sersic"), ("lens_light", 0, "center_x")]


def test_prior_densities_integrate_to_one():
    for d, lo, hi in [(OM.Normal(0.3, 0.2), -3, 3), (OM.LogNormal(0.1, 0.3), 1e-6, 8), (OM.Uniform(2, 6), 2, 6),
                      (OM.TruncatedNormal(2, 0.25, 1, 3), 1, 3)]:
        g = torch.linspace(lo, hi, 200001, dtype=D)
        assert abs(float(torch.trapz(torch.exp(d.log_prob(g)), g)) - 1) < 1e-5
        # change of variables: density of z integrates to one too
        z = torch.linspace(-12, 12, 200001, dtype=D)
        lp = d.log_prob(d.forward(z)) + d.fldj(z)
        assert abs(float(torch.trapz(torch.exp(lp), z)) - 1) < 2e-4


def test_prior_leaf_densities_and_bijectors_match_scipy_and_autodiff():
    """Independent pin of the restated TFP leaf arithmetic (no TFP here): ``log_prob`` of Normal / LogNormal /
    TruncatedNormal / Uniform against scipy.stats, the default event-space bijectors (Identity / Exp / Sigmoid(low, high))
    against their defining maps, and ``forward_log_det_jacobian`` against log|d forward / dz| from autodiff."""
    from scipy import stats
    cases = [(OM.Normal(0.3, 0.7), stats.norm(0.3, 0.7), lambda z: z),
             (OM.LogNormal(math.log(1.25), 0.25), stats.lognorm(s=0.25, scale=1.25), lambda z: torch.exp(z)),
             (OM.TruncatedNormal(2.0, 0.25, 1.0, 3.0), stats.truncnorm((1.0 - 2.0) / 0.25, (3.0 - 2.0) / 0.25, loc=2.0, scale=0.25),
              lambda z: 1.0 + 2.0 * torch.sigmoid(z)),
             (OM.Uniform(0.5, 4.0), stats.uniform(0.5, 3.5), lambda z: 0.5 + 3.5 * torch.sigmoid(z))]
    z = torch.linspace(-4.0, 4.0, 41, dtype=torch.float64).requires_grad_(True)
    for d, ref, fwd in cases:
        x = d.forward(z)
        assert torch.allclose(x, fwd(z), rtol=1e-13, atol=1e-13), type(d).__name__
        assert np.allclose(d.log_prob(x).detach().numpy(), ref.logpdf(x.detach().numpy()), rtol=1e-11, atol=1e-11), type(d).__name__
        (dx,) = torch.autograd.grad(x.sum(), z)
        assert np.allclose(d.fldj(z).detach().numpy(), np.log(np.abs(dx.numpy())), rtol=1e-10, atol=1e-10), type(d).__name__
        assert torch.allclose(d.inverse(x), z, atol=1e-8), type(d).__name__


def test_logprob_autograd_matches_finite_differences():
    wl = workloads.c2_workload()
    prior = workloads.demo_prior()
    from gigalens_b200.model import ProbabilisticModel

    z = ProbabilisticModel(prior).bij_inverse(prior.sample(2, seed=2)).astype(np.float64)
    lp, _, g = oracle_bridge.logprob_and_grad(wl, z, torch.float64)
    sim, pm = oracle_bridge.build_oracle(wl, 2, torch.float64)
    for k in (0, 5, 9, 12, 17, 21):
        h = 1e-6
        zp, zm = z.copy(), z.copy()
        zp[:, k] += h
        zm[:, k] -= h
        fp = pm.log_prob(sim, torch.as_tensor(zp))[0].numpy()
        fm = pm.log_prob(sim, torch.as_tensor(zm))[0].numpy()
        fd = (fp - fm) / (2 * h)
        assert np.allclose(fd, g[:, k], rtol=2e-5, atol=1e-4 * np.abs(g[:, k]).max()), k


def test_golden_vectors_reproduce():
    """tests/golden/c2_golden.npz was written by tests/golden/make_golden.py from this oracle; it pins
    the oracle against accidental edits and is what the GPU tests compare with on the box."""
    gold = np.load(os.path.join(GOLDEN, "c2_golden.npz"))
    wl = workloads.c2_workload()
    lp, chi, dz = oracle_bridge.logprob_and_grad(wl, gold["z"].astype(np.float64), torch.float64)
    assert np.allclose(lp, gold["logp"], rtol=1e-12) and np.allclose(dz, gold["dz"], rtol=1e-9, atol=1e-9)
    assert np.allclose(chi, gold["red_chi2"], rtol=1e-12)


def _positions_golden():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "positions_golden.npz"))
    split = np.cumsum(g["n_img"])[:-1]
    cen = dict(x=np.split(g["cx"], split), y=np.split(g["cy"], split), ex=np.split(g["ex"], split), ey=np.split(g["ey"], split))
    return g, cen


@pytest.mark.parametrize("tag,include_pixels", [("pos", False), ("both", True)])
def test_positions_golden_pins_the_oracle(tag, include_pixels):
    """tests/golden/positions_golden.npz (make_golden_positions.py): stats_positions alone and added to the pixel term."""
    import oracle_bridge
    from gigalens_b200 import workloads
    g, cen = _positions_golden()
    wl = dict(workloads.c2_workload(), centroids=cen, include_pixels=include_pixels)
    lp, chi, dz = oracle_bridge.logprob_and_grad(wl, g["z"].astype(np.float64), torch.float64)
    assert np.allclose(lp, g[f"logp_{tag}"], rtol=1e-12) and np.allclose(chi, g[f"chi2_{tag}"], rtol=1e-12)
    assert np.allclose(dz, g[f"dz_{tag}"], rtol=1e-9, atol=1e-9 * np.abs(g[f"dz_{tag}"]).max())


def test_hessian_golden_through_the_device_arithmetic():
    """Golden Hessian / magnification at the image positions (truth parameters) vs the dual-number point driver (fp64)."""
    import common
    from gigalens_b200 import workloads
    from gigalens_b200.simulator import CompiledModel
    g, cen = _positions_golden()
    wl = workloads.c2_workload()
    cm = CompiledModel(wl["phys_model"])
    mat = cm.flatten(workloads.DEMO_TRUTH, 1, torch, "cpu").numpy().astype(np.float64)
    systems = [(cen["x"][s].astype(np.float64), cen["y"][s].astype(np.float64), cen["ex"][s].astype(np.float64),
                cen["ey"][s].astype(np.float64)) for s in range(len(cen["x"]))]
    out = common.host_positions(cm, mat, systems, np.float64, want_grad=False)
    H = out["hess"][0]
    # (flatten() rounds the truth parameters to fp32; the golden values used the python floats)
    assert np.max(np.abs(H - g["hessian_truth"])) < 2e-7
    mu = 1.0 / ((1 - H[0]) * (1 - H[3]) - H[1] * H[2])
    assert np.allclose(mu, g["magnification_truth"], rtol=5e-6)


# ---------------------------------------------------------------------------------------------
# Independent pins (no reference code involved): closed forms from the literature evaluated with scipy,
# and the two identities every lensing deflection obeys -- alpha is a gradient (curl-free) and its
# divergence is twice the convergence.
# ---------------------------------------------------------------------------------------------
def test_epl_series_matches_hypergeometric_closed_form():
    """Tessore & Metcalf (2015) eq. 13: in the frame of the ellipse, with R = sqrt(q^2 x^2 + y^2), phi = atan2(y, q x),
    alpha = 2b/(1+q) (b/R)^(t-1) e^{i phi} 2F1(1, t/2; 2 - t/2; -f e^{2 i phi}),  f = (1-q)/(1+q).  scipy's hyp2f1 is an
    implementation the series of epl.py:39-54 has never seen."""
    from scipy.special import hyp2f1

    x, y = pts(400, seed=3)
    for gamma, e1, e2 in [(2.0, 0.15, -0.1), (1.6, -0.2, 0.25), (2.4, 0.05, 0.3), (1.9, 0.0, 0.0)]:
        kw = dict(theta_E=1.3, gamma=gamma, e1=e1, e2=e2, center_x=0.07, center_y=-0.04)
        ax, ay = OP.EPL(200).deriv(x, y, **kw)
        phi_e = 0.5 * math.atan2(e2, e1)
        c = min(math.hypot(e1, e2), 1.0)
        q = (1 - c) / (1 + c)
        b, t, f = 1.3 * math.sqrt(q), gamma - 1.0, (1 - q) / (1 + q)
        dx, dy = x.numpy() - 0.07, y.numpy() + 0.04
        xr = dx * math.cos(phi_e) + dy * math.sin(phi_e)
        yr = -dx * math.sin(phi_e) + dy * math.cos(phi_e)
        R = np.sqrt((q * xr) ** 2 + yr ** 2)
        ang = np.arctan2(yr, q * xr)
        a = 2 * b / (1 + q) * (b / R) ** (t - 1) * np.exp(1j * ang) * hyp2f1(1.0, t / 2, 2 - t / 2, -f * np.exp(2j * ang))
        a = a * np.exp(1j * phi_e)   # back to the sky frame
        assert np.max(np.abs(ax.numpy() - a.real)) < 1e-10 and np.max(np.abs(ay.numpy() - a.imag)) < 1e-10


def _div_curl(prof, kw, x, y):
    x = x.clone().requires_grad_(True); y = y.clone().requires_grad_(True)
    ax, ay = prof.deriv(x, y, **kw)
    axx, axy = torch.autograd.grad(ax.sum(), [x, y], retain_graph=True)
    ayx, ayy = torch.autograd.grad(ay.sum(), [x, y])
    return (axx + ayy).detach().numpy(), (axy - ayx).detach().numpy(), x.detach().numpy(), y.detach().numpy()


def test_deflections_are_curl_free_with_the_analytic_convergence():
    x, y = pts(300, seed=5)
    x, y = 2.0 * x, 2.0 * y
    # EPL: kappa = (2 - t)/2 (b / sqrt(q^2 x'^2 + y'^2))^t in the frame of the ellipse (Tessore & Metcalf eq. 1)
    kw = dict(theta_E=1.1, gamma=2.2, e1=0.2, e2=0.1, center_x=0.0, center_y=0.0)
    div, curl, xn, yn = _div_curl(OP.EPL(200), kw, x, y)
    phi_e = 0.5 * math.atan2(0.1, 0.2); c = math.hypot(0.2, 0.1); q = (1 - c) / (1 + c); t = 1.2; b = 1.1 * math.sqrt(q)
    xr = xn * math.cos(phi_e) + yn * math.sin(phi_e); yr = -xn * math.sin(phi_e) + yn * math.cos(phi_e)
    kappa = (2 - t) / 2 * (b / np.sqrt((q * xr) ** 2 + yr ** 2)) ** t
    assert np.max(np.abs(curl)) < 1e-9 and np.max(np.abs(div - 2 * kappa) / (2 * kappa)) < 1e-9
    # SIE = the gamma = 2 member of the same family
    kw = dict(theta_E=0.9, e1=-0.15, e2=0.2, center_x=0.1, center_y=0.0)
    div, curl, xn, yn = _div_curl(OP.SIE(), kw, x, y)
    phi_e = 0.5 * math.atan2(0.2, -0.15); c = math.hypot(0.15, 0.2); q = (1 - c) / (1 + c); b = 0.9 * math.sqrt(q)
    dx = xn - 0.1
    xr = dx * math.cos(phi_e) + yn * math.sin(phi_e); yr = -dx * math.sin(phi_e) + yn * math.cos(phi_e)
    kappa = 0.5 * b / np.sqrt((q * xr) ** 2 + yr ** 2)
    assert np.max(np.abs(curl)) < 1e-9 and np.max(np.abs(div - 2 * kappa) / (2 * kappa)) < 1e-8
    # NFW (Bartelmann 1996): kappa(X) = 2 rho0 Rs (1 - F(X)) / (X^2 - 1), F = acosh(1/X)/sqrt(1-X^2) | acos(1/X)/sqrt(X^2-1),
    # with rho0 = alpha_Rs / (4 Rs^2 (1 + ln 1/2))  (nfw.py:15-52)
    Rs, aRs = 1.7, 0.9
    div, curl, xn, yn = _div_curl(OP.NFW(), dict(Rs=Rs, alpha_Rs=aRs, center_x=0.0, center_y=0.0), x, y)
    X = np.sqrt(xn ** 2 + yn ** 2) / Rs
    F = np.where(X < 1, np.arccosh(1 / np.minimum(X, 1 - 1e-12)) / np.sqrt(np.abs(1 - X ** 2)),
                 np.arccos(1 / np.maximum(X, 1 + 1e-12)) / np.sqrt(np.abs(X ** 2 - 1)))
    rho0 = aRs / (4 * Rs ** 2 * (1 + math.log(0.5)))
    kappa = 2 * rho0 * Rs * (1 - F) / (X ** 2 - 1)
    assert np.max(np.abs(curl)) < 1e-9 and np.max(np.abs(div - 2 * kappa) / np.abs(2 * kappa)) < 1e-7
    # the elliptical families without a closed-form kappa here: a potential must exist (curl-free)
    for prof, kw in [(OP.DPIE(), dict(theta_E=1.0, r_core=0.1, r_cut=2.5, e1=0.2, e2=-0.1, center_x=0.0, center_y=0.1)),
                     (OP.NFW_ELLIPSE(), dict(Rs=1.5, alpha_Rs=1.0, e1=0.1, e2=0.15, center_x=0.0, center_y=0.0)),
                     (OP.DPIS(), dict(theta_E=1.0, r_core=0.1, r_cut=2.5, center_x=0.0, center_y=0.1))]:
        div, curl, _, _ = _div_curl(prof, kw, x, y)
        assert np.max(np.abs(curl)) < 1e-8 * max(1.0, np.max(np.abs(div))), type(prof).__name__
        assert np.all(div > 0)       # positive surface density everywhere


def test_sersic_total_flux_matches_the_closed_form():
    """Integral of I = Ie exp(-bn ((R/Rs)^(1/n) - 1)) over the plane: 2 pi n Rs^2 Ie e^bn bn^(-2n) Gamma(2n) for any axis
    ratio (the elliptical radius of sersic.py:66-80 is area preserving: x sqrt(q), y / sqrt(q))."""
    from scipy.special import gamma as Gamma

    for n, e1, e2 in [(1.0, 0.0, 0.0), (2.0, 0.2, -0.1), (0.7, -0.3, 0.2)]:
        Rs, Ie = 0.4, 3.0
        bn = 1.9992 * n - 0.3271
        want = 2 * math.pi * n * Rs ** 2 * Ie * math.exp(bn) * bn ** (-2 * n) * Gamma(2 * n)
        # polar grid in the stretched frame would hide the ellipticity convention: integrate on a fine Cartesian grid
        L, N = 14.0, 2801
        g = torch.linspace(-L, L, N, dtype=D)
        X, Y = torch.meshgrid(g, g, indexing="xy")
        I = OP.SersicEllipse().light(X, Y, R_sersic=Rs, n_sersic=n, e1=e1, e2=e2, center_x=0.013, center_y=-0.021, Ie=Ie)
        got = float(I.sum()) * (2 * L / (N - 1)) ** 2
        assert abs(got - want) / want < 3e-3, (n, got, want)      # the cusp at R = 0 limits the grid quadrature


def test_shapelet_basis_is_orthonormal():
    """int B_k B_l dx dy = beta^2 delta_kl for the (n1, n2) ordering of shapelets.py:26-46 (Refregier 2003)."""
    n_max, beta = 4, 0.3
    sh = OP.Shapelets(n_max, use_lstsq=True, interpolate=False, dtype=D)
    L, N = 12.0 * beta, 1201
    g = torch.linspace(-L, L, N, dtype=D)
    X, Y = torch.meshgrid(g, g, indexing="xy")
    B = sh.light(X, Y, center_x=0.0, center_y=0.0, beta=beta).reshape(sh.n_layers, -1)
    G = (B @ B.T).numpy() * (2 * L / (N - 1)) ** 2 / beta ** 2
    # the reference rounds the normalisation table to fp32 (shapelets.py:47-48): identity to 1e-6
    assert np.max(np.abs(G - np.eye(sh.n_layers))) < 1e-6
    assert sorted(zip(sh.N1, sh.N2)) == sorted((a, b) for a in range(n_max + 1) for b in range(n_max + 1 - a))
