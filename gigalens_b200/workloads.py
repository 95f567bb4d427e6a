"""The benchmark / parity workloads of BASELINE.json, defined once so that
``bench.py``, ``__graft_entry__.smoke()`` and the tests build identical inputs.

Model specs mirror the reference's demo notebooks and test fixtures:
C1/C2 = ``tests/conftest.py:21-73`` == ``tf-demo.ipynb`` cells 2, 5, 6.
"""
import math
import os

import numpy as np

from . import distributions as tfd
from .model import PhysicalModel
from .profiles.light import sersic, shapelets
from .profiles.mass import dpie_subhalo, epl, nfw, shear
from .simulator import SimulatorConfig

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def load_psf():
    return np.load(os.path.join(ASSETS, "psf.npy")).astype(np.float32)


def load_demo_image():
    return np.load(os.path.join(ASSETS, "demo.npy")).astype(np.float32)


# tf-demo.ipynb cell 5
DEMO_TRUTH = {
    "lens_mass": [
        {"theta_E": 1.1, "gamma": 2.0, "e1": 0.1, "e2": 0.1, "center_x": 0.1, "center_y": 0.0},
        {"gamma1": -0.01, "gamma2": 0.03},
    ],
    "lens_light": [
        {"R_sersic": 0.8, "n_sersic": 2.5, "e1": 0.09534746574143645, "e2": 0.14849487967198177, "center_x": 0.1,
         "center_y": 0.0, "Ie": 499.3695906504067}
    ],
    "source_light": [
        {"R_sersic": 0.25, "n_sersic": 1.5, "e1": 0.0, "e2": 0.0, "center_x": 0.09566681002252231,
         "center_y": -0.0639623054267272, "Ie": 149.58828877085668}
    ],
}


def demo_prior():
    """tests/conftest.py:21-73 (the fork's dict top level)."""
    lens_mass = [
        dict(theta_E=tfd.LogNormal(math.log(1.25), 0.25), gamma=tfd.TruncatedNormal(2, 0.25, 1, 3),
             e1=tfd.Normal(0, 0.1), e2=tfd.Normal(0, 0.1), center_x=tfd.Normal(0, 0.05), center_y=tfd.Normal(0, 0.05)),
        dict(gamma1=tfd.Normal(0, 0.05), gamma2=tfd.Normal(0, 0.05)),
    ]
    lens_light = [
        dict(R_sersic=tfd.LogNormal(math.log(1.0), 0.15), n_sersic=tfd.Uniform(2, 6),
             e1=tfd.TruncatedNormal(0, 0.1, -0.3, 0.3), e2=tfd.TruncatedNormal(0, 0.1, -0.3, 0.3),
             center_x=tfd.Normal(0, 0.05), center_y=tfd.Normal(0, 0.05), Ie=tfd.LogNormal(math.log(500.0), 0.3))
    ]
    source_light = [
        dict(R_sersic=tfd.LogNormal(math.log(0.25), 0.15), n_sersic=tfd.Uniform(0.5, 4),
             e1=tfd.TruncatedNormal(0, 0.15, -0.5, 0.5), e2=tfd.TruncatedNormal(0, 0.15, -0.5, 0.5),
             center_x=tfd.Normal(0, 0.25), center_y=tfd.Normal(0, 0.25), Ie=tfd.LogNormal(math.log(150.0), 0.5))
    ]
    return tfd.JointDistributionNamed(dict(lens_mass=lens_mass, lens_light=lens_light, source_light=source_light))


def demo_phys_model():
    return PhysicalModel([epl.EPL(50), shear.Shear()], [sersic.SersicEllipse(use_lstsq=False)],
                         [sersic.SersicEllipse(use_lstsq=False)])


def demo_sim_config():
    return SimulatorConfig(delta_pix=0.065, num_pix=60, supersample=2, kernel=load_psf())


DEMO_NOISE = dict(background_rms=0.2, exp_time=100.0)


def c2_workload():
    """BASELINE.json configs[0]/[1]: EPL+shear, SersicEllipse x2, 60x60, ss=2, 13x13 PSF."""
    return dict(name="C2: EPL+Shear / SersicEllipse x2, 60x60, ss=2, PSF 13x13 (tf-demo.ipynb)",
                phys_model=demo_phys_model(), sim_config=demo_sim_config(), prior=demo_prior(),
                observed=load_demo_image(), **DEMO_NOISE)


def c3_prior():
    """shapelets-demo.ipynb cell 4 lens prior + shear; Shapelets source with linear amplitudes solved."""
    lens_mass = [
        dict(theta_E=tfd.LogNormal(math.log(1.0), 0.25), gamma=tfd.TruncatedNormal(2, 0.25, 1, 3),
             e1=tfd.Normal(0, 0.1), e2=tfd.Normal(0, 0.1), center_x=tfd.Normal(0, 0.025), center_y=tfd.Normal(0, 0.025)),
        dict(gamma1=tfd.Normal(0, 0.05), gamma2=tfd.Normal(0, 0.05)),
    ]
    source_light = [dict(beta=tfd.LogNormal(math.log(0.1), 0.15), center_x=tfd.Normal(0, 0.01), center_y=tfd.Normal(0, 0.01))]
    return tfd.JointDistributionNamed(dict(lens_mass=lens_mass, source_light=source_light))


def c3_workload(n_max=10, interpolate=False, observed=None):
    """BASELINE.json configs[2]: EPL+shear lens, Shapelets(n_max) source via lstsq_simulate, 60x60, ss=2,
    PSF, BackwardProbModel (bg 0.1, exp 200).  `observed` defaults to the demo image; bench.py replaces
    it by a simulation of a seeded truth (`c3_observation`)."""
    pm = PhysicalModel([epl.EPL(50), shear.Shear()], [], [shapelets.Shapelets(n_max, use_lstsq=True, interpolate=interpolate)])
    return dict(name=f"C3: EPL+Shear / Shapelets(n_max={n_max}) via lstsq_simulate, 60x60, ss=2, PSF 13x13",
                phys_model=pm, sim_config=demo_sim_config(), prior=c3_prior(),
                observed=load_demo_image() if observed is None else observed, background_rms=0.1, exp_time=200.0)


def c3_observation(n_max=10, seed=121, noise_seed=1):
    """Noisy simulation of a seeded truth with amplitudes ~ N(0, 500/sqrt(k+1)) (mirrors
    shapelets-demo.ipynb cells 6-7); runs on the GPU through this repo's own simulate."""
    from .simulator import LensSimulator

    src = shapelets.Shapelets(n_max, use_lstsq=False, interpolate=False)
    pm = PhysicalModel([epl.EPL(50), shear.Shear()], [], [src])
    rng = np.random.default_rng(seed)
    truth = {"lens_mass": [dict(theta_E=1.05, gamma=2.05, e1=0.08, e2=-0.05, center_x=0.01, center_y=-0.01),
                           dict(gamma1=0.02, gamma2=-0.01)],
             "source_light": [dict(beta=0.1, center_x=0.005, center_y=-0.004,
                                   **{n: float(rng.normal(0, 500 / np.sqrt(k + 1))) for k, n in enumerate(src._amp_names)})]}
    sim = LensSimulator(pm, demo_sim_config(), bs=1)
    img = sim.simulate(truth).cpu().numpy().astype(np.float64)
    nrng = np.random.default_rng(noise_seed)
    noise = nrng.normal(size=img.shape) * np.sqrt(0.1 ** 2 + np.clip(img, 0, None) / 200.0)
    return (img + noise).astype(np.float32)


def cluster_catalogue(G=30, seed=7):
    """member galaxies in a +-9 arcsec box, |e| in [0.02, 0.6], lum ~ LogN(0, 0.5)."""
    rng = np.random.default_rng(seed)
    cx, cy = rng.uniform(-9, 9, G), rng.uniform(-9, 9, G)
    e1, e2 = rng.normal(0, 0.1, G), rng.normal(0, 0.1, G)
    mod = np.sqrt(e1 ** 2 + e2 ** 2)
    scale = np.clip(mod, 0.02, 0.6) / np.maximum(mod, 1e-12)
    e1, e2 = e1 * scale, e2 * scale
    return dict(lum=rng.lognormal(0, 0.5, G).tolist(), center_x=cx.tolist(), center_y=cy.tolist(), e1=e1.tolist(), e2=e2.tolist())


def c4_prior():
    lens_mass = [
        dict(Rs=tfd.LogNormal(math.log(10.0), 0.2), alpha_Rs=tfd.LogNormal(math.log(8.0), 0.2),
             center_x=tfd.Normal(0, 0.5), center_y=tfd.Normal(0, 0.5)),
        dict(theta_E=tfd.LogNormal(math.log(0.8), 0.2), r_core=tfd.LogNormal(math.log(0.05), 0.1),
             r_cut=tfd.LogNormal(math.log(5.0), 0.2)),
        dict(gamma1=tfd.Normal(0, 0.05), gamma2=tfd.Normal(0, 0.05)),
    ]
    source_light = [
        dict(R_sersic=tfd.LogNormal(math.log(0.25), 0.15), n_sersic=tfd.Uniform(0.5, 4),
             e1=tfd.TruncatedNormal(0, 0.15, -0.5, 0.5), e2=tfd.TruncatedNormal(0, 0.15, -0.5, 0.5),
             center_x=tfd.Normal(0, 1.0), center_y=tfd.Normal(0, 1.0), Ie=tfd.LogNormal(math.log(150.0), 0.5))
    ]
    return tfd.JointDistributionNamed(dict(lens_mass=lens_mass, source_light=source_light))


def c4_phys_model(G=30):
    return PhysicalModel([nfw.NFW(), dpie_subhalo.DPIESubhalo(1.0, cluster_catalogue(G)), shear.Shear()], [],
                         [sersic.SersicEllipse()])


def c4_sim_config(num_pix=200):
    return SimulatorConfig(delta_pix=0.1, num_pix=num_pix, supersample=2, kernel=load_psf())


def c4_workload(num_pix=200, G=30, observed=None):
    """BASELINE.json configs[3]: NFW halo + G member-galaxy dPIE deflectors on a scaling relation + shear,
    SersicEllipse source, 200x200, ss=2 (SURVEY.md §8d fixes the details BASELINE.json leaves open)."""
    obs = np.zeros((num_pix, num_pix), dtype=np.float32) if observed is None else observed
    return dict(name=f"C4: NFW + {G}-member dPIE scaling relation + Shear / SersicEllipse, {num_pix}x{num_pix}, ss=2, PSF 13x13",
                phys_model=c4_phys_model(G), sim_config=c4_sim_config(num_pix), prior=c4_prior(), observed=obs,
                background_rms=0.2, exp_time=100.0)


def c4_observation(num_pix=200, G=30, seed=11, noise_seed=12):
    """Noisy simulation at a seeded prior draw (runs on the GPU through this repo's own simulate)."""
    from .simulator import LensSimulator

    wl = c4_workload(num_pix, G)
    sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=1)
    img = sim.simulate(wl["prior"].sample(1, seed=seed)).cpu().numpy().astype(np.float64)
    nrng = np.random.default_rng(noise_seed)
    noise = nrng.normal(size=img.shape) * np.sqrt(0.2 ** 2 + np.clip(img, 0, None) / 100.0)
    return (img + noise).astype(np.float32)
