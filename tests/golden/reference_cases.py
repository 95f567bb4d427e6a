"""Case specifications shared by ``make_reference_golden.py`` (which EXECUTES THE REFERENCE'S OWN SOURCE FILES on them,
through the torch-backed ``oracle/tfshim`` stand-in for tensorflow) and by the tests that compare the oracle and the
CUDA path with the stored results.  Plain python / numpy only: this module imports neither the reference, nor the
oracle, nor the product, so every side builds its own objects from the same description.

A profile is ``(class name, constructor kwargs)``; parameter values are float32-representable so that the float32 and
the float64 run of the reference see identical inputs."""
import numpy as np

# parameter ranges of the seeded draws (well inside every profile's domain; clamps are exercised by the EDGE cases below)
RANGES = {
    "theta_E": (0.8, 1.6), "gamma": (1.6, 2.4), "e1": (-0.25, 0.25), "e2": (-0.25, 0.25),
    "center_x": (-0.2, 0.2), "center_y": (-0.2, 0.2), "gamma1": (-0.08, 0.08), "gamma2": (-0.08, 0.08),
    "Rs": (0.5, 3.0), "alpha_Rs": (0.5, 2.0), "r_core": (0.02, 0.2), "r_cut": (1.0, 6.0),
    "r_trunc": (2.0, 8.0), "Ra": (0.02, 0.2),
    "R_sersic": (0.3, 1.2), "n_sersic": (0.8, 5.0), "Ie": (20.0, 300.0), "beta": (0.08, 0.2),
    "Rb": (0.05, 0.3), "alpha": (1.0, 3.0),
}

PARAMS = {   # parameter names per class, in the reference's ``_params`` order (+ the amplitude for light profiles)
    "EPL": ["theta_E", "gamma", "e1", "e2", "center_x", "center_y"],
    "Shear": ["gamma1", "gamma2"],
    "SIE": ["theta_E", "e1", "e2", "center_x", "center_y"],
    "SIS": ["theta_E", "center_x", "center_y"],
    "NFW": ["Rs", "alpha_Rs", "center_x", "center_y"],
    "NFW_ELLIPSE": ["Rs", "alpha_Rs", "e1", "e2", "center_x", "center_y"],
    "DPIS": ["theta_E", "r_core", "r_cut", "center_x", "center_y"],
    "DPIE": ["theta_E", "r_core", "r_cut", "center_x", "center_y", "e1", "e2"],
    "TNFW": ["Rs", "alpha_Rs", "r_trunc", "center_x", "center_y"],
    "DPIEP": ["theta_E", "Ra", "Rs", "center_x", "center_y", "e1", "e2"],
    "DPIESubhalo": ["theta_E", "r_core", "r_cut"],
    "Sersic": ["R_sersic", "n_sersic", "center_x", "center_y", "Ie"],
    "SersicEllipse": ["R_sersic", "n_sersic", "e1", "e2", "center_x", "center_y", "Ie"],
    "CoreSersic": ["R_sersic", "n_sersic", "Rb", "alpha", "gamma", "e1", "e2", "center_x", "center_y", "Ie"],
    "Shapelets": ["beta", "center_x", "center_y"],
    "ScaledSIS": ["theta_E"],      # ScalingRelation(profile=SIS(), scaling_params=["theta_E"], ...): a generic scaling relation
}


# stand-in values of prior.log_prob(params) and the bijector's forward_log_det_jacobian in the executed ForwardProbModel.log_prob
FAKE_LOG_PRIOR = np.array([0.5, -1.25, 3.0, -0.75])
FAKE_FLDJ = np.array([0.25, 2.0, -1.5, 0.125])


def f32(v):
    return np.asarray(np.asarray(v, dtype=np.float32), dtype=np.float64)


def catalogue(G, seed, half_width=1.0):
    """Member-galaxy catalogue of a scaling relation (``scaling_relation.py:8-55``): luminosities and the per-member
    constants of the dPIE (centre, ellipticity; |e| kept away from the e = 0 singularity of ``piemd.py:207``)."""
    rng = np.random.default_rng(seed)
    e = rng.normal(0, 0.1, size=(2, G))
    return dict(lum=f32(rng.lognormal(0, 0.5, G)).tolist(), center_x=f32(rng.uniform(-half_width, half_width, G)).tolist(),
                center_y=f32(rng.uniform(-half_width, half_width, G)).tolist(), e1=f32(e[0] + 0.05).tolist(),
                e2=f32(e[1] - 0.04).tolist())


def draw(cls, ctor, bs, rng):
    names = list(PARAMS[cls])
    if cls in ("Sersic", "SersicEllipse", "CoreSersic") and ctor.get("use_lstsq"):
        names.remove("Ie")
    out = {}
    for n in names:
        lo, hi = (0.1, 0.9) if (cls == "CoreSersic" and n == "gamma") else RANGES[n]
        out[n] = f32(rng.uniform(lo, hi, size=bs))
    if cls == "DPIEP":                       # Rs is the cut radius there
        out["Rs"] = f32(rng.uniform(1.0, 6.0, size=bs))
    if cls == "Shapelets" and not ctor.get("use_lstsq"):
        n_layers = (ctor["n_max"] + 1) * (ctor["n_max"] + 2) // 2
        w = len(str(n_layers))
        for k in range(n_layers):
            out[f"amp{str(k).zfill(w)}"] = f32(rng.normal(0, 50.0 / np.sqrt(k + 1.0), size=bs))
    return out


# ------------------------------------------------------------------------------------------------- profile-level cases
def profile_cases():
    """``deriv`` / ``hessian`` of every deflector and ``light`` of every light profile on scattered points.
    Returns {key: dict(cls, ctor, kind, x, y, params)} with x, y of shape (N, bs) and params of shape (bs,)."""
    cases = {}
    bs, n = 3, 48
    mass = [("EPL", dict(niter=50)), ("Shear", {}), ("SIE", {}), ("SIS", {}), ("NFW", {}), ("NFW_ELLIPSE", {}), ("DPIS", {}),
            ("DPIE", {}), ("TNFW", {}), ("DPIEP", {}),
            ("DPIESubhalo", dict(lum_star=1.0, galaxy_catalogue=catalogue(5, 3)))]
    light = [("Sersic", dict(use_lstsq=False)), ("SersicEllipse", dict(use_lstsq=False)), ("SersicEllipse", dict(use_lstsq=True)),
             ("CoreSersic", dict(use_lstsq=False)),
             ("Shapelets", dict(n_max=4, use_lstsq=False, interpolate=False)),
             ("Shapelets", dict(n_max=4, use_lstsq=False, interpolate=True)),
             ("Shapelets", dict(n_max=6, use_lstsq=True, interpolate=False)),
             ("Shapelets", dict(n_max=6, use_lstsq=True, interpolate=True))]
    for i, (cls, ctor) in enumerate(mass + light):
        rng = np.random.default_rng(1000 + i)
        kind = "mass" if i < len(mass) else "light"
        scale = 0.4 if cls == "Shapelets" else 2.5        # shapelets: stay inside the +-5 beta table most of the time
        x = f32(rng.uniform(-scale, scale, size=(n, 1))).repeat(bs, axis=1)
        y = f32(rng.uniform(-scale, scale, size=(n, 1))).repeat(bs, axis=1)
        tag = cls + "".join(f"_{k}{int(v)}" for k, v in sorted(ctor.items()) if isinstance(v, (bool, int)) and k != "niter")
        cases[tag] = dict(cls=cls, ctor=ctor, kind=kind, x=x, y=y, params=draw(cls, ctor, bs, rng))
    # EDGE: the clamps and special branches the reference has
    z3 = f32(np.zeros(3))
    x = f32([[0.0], [1.0], [1e-9], [0.3], [-2.0], [2.0]]).repeat(3, axis=1)
    y = f32([[0.0], [0.0], [0.0], [-0.4], [1.0], [0.0]]).repeat(3, axis=1)
    cases["EDGE_SIS_centre"] = dict(cls="SIS", ctor={}, kind="mass", x=x, y=y,
                                    params=dict(theta_E=f32([1.0, 1.2, 0.9]), center_x=z3, center_y=z3))
    cases["EDGE_EPL_round_and_extreme"] = dict(                 # e = 0 (f = 0: niter = 2 - log(1e-12)/inf), e ~ 1 clip, gamma at the ends
        cls="EPL", ctor=dict(niter=50), kind="mass", x=x[1:], y=y[1:],
        params=dict(theta_E=f32([1.0, 1.2, 0.9]), gamma=f32([2.0, 1.2, 2.8]), e1=f32([0.0, 0.6, -0.3]), e2=f32([0.0, 0.5, 0.4]),
                    center_x=z3, center_y=z3))
    cases["EDGE_NFW_X_eq_1"] = dict(                             # X == 1 exactly (g left at 1.0), X below the 1e-6 floor, R below r_min
        cls="NFW", ctor={}, kind="mass", x=x, y=y,
        params=dict(Rs=f32([1.0, 2.0, 0.5]), alpha_Rs=f32([1.0, 0.7, 1.5]), center_x=z3, center_y=z3))
    cases["EDGE_DPIE_radii_sorted"] = dict(                      # r_core > r_cut (swapped), r_core below r_min, r_cut == r_core
        cls="DPIE", ctor={}, kind="mass", x=x[1:], y=y[1:],
        params=dict(theta_E=f32([1.0, 1.2, 0.9]), r_core=f32([3.0, 1e-6, 0.5]), r_cut=f32([0.1, 2.0, 0.5]),
                    center_x=z3, center_y=z3, e1=f32([0.1, -0.2, 0.05]), e2=f32([0.05, 0.1, -0.3])))
    cases["EDGE_SIE_round"] = dict(cls="SIE", ctor={}, kind="mass", x=x[1:], y=y[1:],
                                   params=dict(theta_E=f32([1.0, 1.2, 0.9]), e1=f32([1e-3, 0.7, -0.8]), e2=f32([0.0, 0.7, 0.6]),
                                               center_x=z3, center_y=z3))
    cases["EDGE_Sersic_centre"] = dict(cls="SersicEllipse", ctor=dict(use_lstsq=False), kind="light", x=x, y=y,
                                       params=dict(R_sersic=f32([1.0, 0.5, 0.25]), n_sersic=f32([2.0, 4.0, 0.5]), e1=f32([0.0, 0.2, -0.1]),
                                                   e2=f32([0.0, 0.1, 0.3]), center_x=z3, center_y=z3, Ie=f32([5.0, 100.0, 30.0])))
    return cases


# ------------------------------------------------------------------------------------------------- simulator / model cases
DEMO_TRUTH = {          # tf-demo.ipynb cell 5
    "lens_mass": [{"theta_E": 1.1, "gamma": 2.0, "e1": 0.1, "e2": 0.1, "center_x": 0.1, "center_y": 0.0},
                  {"gamma1": -0.01, "gamma2": 0.03}],
    "lens_light": [{"R_sersic": 0.8, "n_sersic": 2.5, "e1": 0.09534746574143645, "e2": 0.14849487967198177, "center_x": 0.1,
                    "center_y": 0.0, "Ie": 499.3695906504067}],
    "source_light": [{"R_sersic": 0.25, "n_sersic": 1.5, "e1": 0.0, "e2": 0.0, "center_x": 0.09566681002252231,
                      "center_y": -0.0639623054267272, "Ie": 149.58828877085668}],
}


def _draw_model(model, bs, rng):
    return {g: [draw(cls, ctor, bs, rng) for cls, ctor in model[g]] for g in ("lens_mass", "lens_light", "source_light")}


def simulator_cases(psf, demo):
    """{key: dict(model, sim, params, observed, noise | error_map, [centroids], [lstsq])}.  ``psf`` / ``demo`` are the reference's
    ``assets/psf.npy`` and ``assets/demo.npy``."""
    cases = {}
    # C2: the demo model at the notebook's truth and one seeded draw, 60 x 60, ss = 2, 13 x 13 PSF, observation demo.npy
    model = dict(lens_mass=[("EPL", dict(niter=50)), ("Shear", {})], lens_light=[("SersicEllipse", dict(use_lstsq=False))],
                 source_light=[("SersicEllipse", dict(use_lstsq=False))])
    rng = np.random.default_rng(2001)
    p = _draw_model(model, 2, rng)
    for g in p:
        for d, t in zip(p[g], DEMO_TRUTH[g]):
            for k in d:
                d[k][0] = f32(t[k])
    p["source_light"][0]["center_x"][1] = f32(0.05)
    p["source_light"][0]["center_y"][1] = f32(-0.03)
    cases["c2"] = dict(model=model, sim=dict(delta_pix=0.065, num_pix=60, supersample=2, kernel=f32(psf), pix_region=None), params=p,
                       observed=f32(demo), noise=dict(background_rms=0.2, exp_time=100.0), variants=True)
    # masked grid, ss = 3, truncated PSF, explicit error map, SIE + NFW_ELLIPSE + shear, Sersic lens light, two source components
    model = dict(lens_mass=[("SIE", {}), ("NFW_ELLIPSE", {}), ("Shear", {})], lens_light=[("Sersic", dict(use_lstsq=False))],
                 source_light=[("SersicEllipse", dict(use_lstsq=False)), ("Shapelets", dict(n_max=3, use_lstsq=False, interpolate=False))])
    rng = np.random.default_rng(2002)
    n = 24
    p = _draw_model(model, 3, rng)
    p["lens_mass"][1]["alpha_Rs"] = f32(p["lens_mass"][1]["alpha_Rs"] * 0.3)
    mask = (rng.uniform(size=(n, n)) > 0.2).astype(np.float64)
    cases["masked_ss3"] = dict(model=model, sim=dict(delta_pix=0.12, num_pix=n, supersample=3, kernel=f32(psf[3:10, 3:10]), pix_region=mask),
                               params=p, observed=f32(rng.normal(0, 1, size=(n, n)) + 5), error_map=f32(rng.uniform(0.5, 1.5, size=(n, n))),
                               variants=True)
    # no PSF, ss = 1, dPIS + TNFW + dPIEP deflectors
    model = dict(lens_mass=[("DPIS", {}), ("TNFW", {}), ("DPIEP", {})], lens_light=[], source_light=[("Sersic", dict(use_lstsq=False))])
    rng = np.random.default_rng(2003)
    n = 16
    p = _draw_model(model, 2, rng)
    for d in p["lens_mass"]:
        for k in ("theta_E", "alpha_Rs"):
            if k in d:
                d[k] = f32(d[k] * 0.4)
    cases["nopsf_ss1"] = dict(model=model, sim=dict(delta_pix=0.2, num_pix=n, supersample=1, kernel=None, pix_region=None), params=p,
                              observed=f32(rng.normal(0, 1, size=(n, n)) + 3), noise=dict(background_rms=0.3, exp_time=50.0), variants=False)
    # cluster: NFW + dPIE scaling relation (G = 8, +-4 arcsec) + shear, with two multiply-imaged sources for stats_positions
    model = dict(lens_mass=[("NFW", {}), ("DPIESubhalo", dict(lum_star=1.0, galaxy_catalogue=catalogue(8, 5, 3.5))), ("Shear", {})],
                 lens_light=[], source_light=[("SersicEllipse", dict(use_lstsq=False))])
    rng = np.random.default_rng(2004)
    n = 40
    p = _draw_model(model, 2, rng)
    p["lens_mass"][0]["Rs"] = f32([4.0, 5.0])
    p["lens_mass"][0]["alpha_Rs"] = f32([2.5, 3.0])
    p["lens_mass"][1]["theta_E"] = f32([0.3, 0.25])
    p["lens_mass"][1]["r_core"] = f32([0.05, 0.04])
    p["lens_mass"][1]["r_cut"] = f32([2.0, 2.5])
    cen = dict(x=[f32([2.9, -2.2, 0.4, -1.0]), f32([1.6, -2.5])], y=[f32([0.7, -1.9, 2.8, 2.4]), f32([-2.2, 1.1])],
               ex=[f32([0.04, 0.04, 0.05, 0.03]), f32([0.05, 0.05])], ey=[f32([0.04, 0.05, 0.04, 0.03]), f32([0.04, 0.06])])
    cases["cluster"] = dict(model=model, sim=dict(delta_pix=0.2, num_pix=n, supersample=2, kernel=f32(psf), pix_region=None), params=p,
                            observed=f32(rng.normal(0, 0.2, size=(n, n)) + 0.5), noise=dict(background_rms=0.2, exp_time=100.0),
                            centroids=cen, variants=False)
    # a rotated, sheared, non-symmetric pixel -> angle transform (simulator.py:32-55: the einsum of pix2angle applies the TRANSPOSE of
    # the supersampled transform; conversion_factor = det of the un-supersampled one, tf/simulator.py:27-29)
    model = dict(lens_mass=[("EPL", dict(niter=50)), ("Shear", {})], lens_light=[("Sersic", dict(use_lstsq=False))],
                 source_light=[("SersicEllipse", dict(use_lstsq=False))])
    rng = np.random.default_rng(2008)
    n = 18
    p = _draw_model(model, 2, rng)
    ang = 0.4
    T = 0.11 * np.array([[np.cos(ang), -1.15 * np.sin(ang)], [0.85 * np.sin(ang), np.cos(ang)]])
    cases["rotated_wcs"] = dict(model=model, sim=dict(delta_pix=0.11, num_pix=n, supersample=2, kernel=f32(psf[4:9, 4:9]), pix_region=None,
                                                      transform_pix2angle=T), params=p,
                                observed=f32(rng.normal(0, 1, size=(n, n)) + 4), noise=dict(background_rms=0.25, exp_time=80.0), variants=False)
    # per-profile constants merged into the call (tf/simulator.py:75-76,129-138), an EPL whose series is cut by `niter` (maximum_iterations,
    # epl.py:50) at strong ellipticity, a generic ScalingRelation around SIS with its own power / L*, pooling without a PSF, and a
    # table-interpolated Shapelets source whose argument leaves the +-5 table (fill value 0, shapelets.py:58-60)
    model = dict(lens_mass=[("EPL", dict(niter=6)),
                            ("ScaledSIS", dict(lum_star=1.3, scaling_params_power={"theta_E": 0.6},
                                               galaxy_catalogue={k: v for k, v in catalogue(4, 9, 0.8).items() if k in ("lum", "center_x", "center_y")})),
                            ("Shear", {})],
                 lens_light=[], source_light=[("Sersic", dict(use_lstsq=False)), ("Shapelets", dict(n_max=3, use_lstsq=False, interpolate=True))])
    rng = np.random.default_rng(2009)
    n = 20
    p = _draw_model(model, 3, rng)
    p["lens_mass"][0]["e1"] = f32([0.45, -0.3, 0.1])
    p["lens_mass"][0]["e2"] = f32([-0.3, 0.4, 0.05])
    p["lens_mass"][1]["theta_E"] = f32([0.12, 0.2, 0.08])
    consts = dict(lens_mass=[{"gamma": 2.15, "center_x": 0.04}, {}, {}], lens_light=[], source_light=[{"n_sersic": 1.5}, {}])
    for g in consts:
        for d, cdict in zip(p[g], consts[g]):
            for k in cdict:
                d.pop(k)
    cases["constants_capped"] = dict(model=model, constants=consts, sim=dict(delta_pix=0.1, num_pix=n, supersample=2, kernel=None, pix_region=None),
                                     params=p, observed=f32(rng.normal(0, 1, size=(n, n)) + 4), noise=dict(background_rms=0.3, exp_time=60.0),
                                     variants=False)
    # image-position likelihood through a deflector WITHOUT an analytic Hessian: EPL + shear, so magnification comes from the autodiff
    # default of tf/profile.py:9-30 through the EPL while_loop, and its parameter gradient from second-order autodiff
    model = dict(lens_mass=[("EPL", dict(niter=50)), ("Shear", {})], lens_light=[], source_light=[("SersicEllipse", dict(use_lstsq=False))])
    rng = np.random.default_rng(2010)
    n = 20
    p = _draw_model(model, 3, rng)
    p["lens_mass"][0]["theta_E"] = f32([1.1, 1.15, 1.05])
    cen = dict(x=[f32([1.05, -0.95, 0.2, -0.3])], y=[f32([0.35, -0.5, 1.1, -1.0])], ex=[f32([0.03, 0.04, 0.03, 0.05])],
               ey=[f32([0.04, 0.03, 0.05, 0.03])])
    cases["epl_positions"] = dict(model=model, sim=dict(delta_pix=0.15, num_pix=n, supersample=1, kernel=None, pix_region=None), params=p,
                                  observed=f32(rng.normal(0, 1, size=(n, n)) + 3), noise=dict(background_rms=0.3, exp_time=50.0),
                                  centroids=cen, variants=False)
    # C4 at BASELINE geometry (configs[3]): 200 x 200, ss = 2, +-10 arcsec, NFW + 30-member dPIE scaling relation + shear;
    # one sample near the prior medians of workloads.c4_prior() (image, likelihood and gradient only: the fixture stays small)
    model = dict(lens_mass=[("NFW", {}), ("DPIESubhalo", dict(lum_star=1.0, galaxy_catalogue=c4_catalogue())), ("Shear", {})],
                 lens_light=[], source_light=[("SersicEllipse", dict(use_lstsq=False))])
    rng = np.random.default_rng(2006)
    n = 200
    p = dict(lens_mass=[dict(Rs=f32([10.5]), alpha_Rs=f32([7.6]), center_x=f32([0.3]), center_y=f32([-0.2])),
                        dict(theta_E=f32([0.85]), r_core=f32([0.052]), r_cut=f32([4.6])),
                        dict(gamma1=f32([0.03]), gamma2=f32([-0.02]))],
             lens_light=[],
             source_light=[dict(R_sersic=f32([0.27]), n_sersic=f32([1.8]), e1=f32([0.1]), e2=f32([-0.12]), center_x=f32([0.6]),
                                center_y=f32([-0.4]), Ie=f32([140.0]))])
    cases["c4_baseline"] = dict(model=model, sim=dict(delta_pix=0.1, num_pix=n, supersample=2, kernel=f32(psf), pix_region=None), params=p,
                                observed=f32(np.abs(rng.normal(0, 0.3, size=(n, n)))), noise=dict(background_rms=0.2, exp_time=100.0),
                                variants=False, image_and_likelihood_only=True)
    return cases


def lstsq_tail_case(psf):
    """``lstsq_simulate`` cannot run as written (see make_reference_golden.py), but its tail -- weights, normal equations,
    ``pinv(rcond=1e-6)``, recombination: ``tf/simulator.py:231-240`` -- can: the generator executes those source lines on a
    component stack (built by the oracle from this model, whose pieces are pinned individually)."""
    model = dict(lens_mass=[("EPL", dict(niter=50)), ("Shear", {})], lens_light=[("SersicEllipse", dict(use_lstsq=True))],
                 source_light=[("Shapelets", dict(n_max=4, use_lstsq=True, interpolate=False))])
    rng = np.random.default_rng(2007)
    n = 30
    p = _draw_model(model, 2, rng)
    p["source_light"][0]["beta"] = f32([0.15, 0.12])
    obs = f32(np.abs(rng.normal(0, 1, size=(n, n))) * 3 + 1)
    err = f32(np.sqrt(0.2 ** 2 + np.clip(obs, 0, np.inf) / 100.0))
    return dict(model=model, sim=dict(delta_pix=0.13, num_pix=n, supersample=2, kernel=f32(psf[4:9, 4:9]), pix_region=None), params=p,
                observed=obs, err_map=err)


def lstsq_sersic_case(psf):
    """The part of ``lstsq_simulate`` that executes in the reference when every linear component is a single-component profile
    (``self.depth`` counts PROFILES, ``tf/simulator.py:57-59``, so a Shapelets set does not fit its reshape): SersicEllipse lens light
    + two Sersic-type sources, ss = 2, PSF."""
    model = dict(lens_mass=[("EPL", dict(niter=50)), ("Shear", {})], lens_light=[("SersicEllipse", dict(use_lstsq=True))],
                 source_light=[("SersicEllipse", dict(use_lstsq=True)), ("Sersic", dict(use_lstsq=True))])
    rng = np.random.default_rng(2011)
    n = 30
    p = _draw_model(model, 2, rng)
    p["source_light"][1]["center_x"] = f32([0.3, -0.25])
    p["source_light"][1]["R_sersic"] = f32([0.2, 0.35])
    obs = f32(np.abs(rng.normal(0, 1, size=(n, n))) * 3 + 1)
    err = f32(np.sqrt(0.2 ** 2 + np.clip(obs, 0, np.inf) / 100.0))
    return dict(model=model, sim=dict(delta_pix=0.13, num_pix=n, supersample=2, kernel=f32(psf[4:9, 4:9]), pix_region=None), params=p,
                observed=obs, err_map=err)


def c4_catalogue(G=30, seed=7):
    """``gigalens_b200.workloads.cluster_catalogue``, restated here so that this module stays import-free;
    tests/test_reference_golden.py asserts the two are identical."""
    rng = np.random.default_rng(seed)
    cx, cy = rng.uniform(-9, 9, G), rng.uniform(-9, 9, G)
    e1, e2 = rng.normal(0, 0.1, G), rng.normal(0, 0.1, G)
    mod = np.sqrt(e1 ** 2 + e2 ** 2)
    scale = np.clip(mod, 0.02, 0.6) / np.maximum(mod, 1e-12)
    e1, e2 = e1 * scale, e2 * scale
    return dict(lum=rng.lognormal(0, 0.5, G).tolist(), center_x=cx.tolist(), center_y=cy.tolist(), e1=e1.tolist(), e2=e2.tolist())


def grad_keys(params):
    """Fixed order of the gradient leaves: group, component index, sorted parameter name."""
    return [(g, i, k) for g in ("lens_mass", "lens_light", "source_light") for i, d in enumerate(params[g]) for k in sorted(d)]
