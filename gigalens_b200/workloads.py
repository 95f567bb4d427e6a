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
