from ... import _cabi
from ...profile import LightProfile


class Sersic(LightProfile):
    """Spherical Sersic (reference ``tf/profiles/light/sersic.py:9-35``)."""

    _name = "SERSIC"
    _params = ["R_sersic", "n_sersic", "center_x", "center_y"]
    _amp = "Ie"
    _type_id = _cabi.GL_SERSIC

    def __init__(self, use_lstsq=False):
        super().__init__(use_lstsq=use_lstsq)


class SersicEllipse(Sersic):
    """Elliptical Sersic (reference ``tf/profiles/light/sersic.py:66-80``)."""

    _name = "SERSIC_ELLIPSE"
    _params = ["R_sersic", "n_sersic", "e1", "e2", "center_x", "center_y"]
    _amp = "Ie"
    _type_id = _cabi.GL_SERSIC_ELLIPSE


class CoreSersic(Sersic):
    """Core-Sersic (reference ``tf/profiles/light/sersic.py:83-132``), evaluated exactly as the reference writes it:
    ``Ie (1 + (Rb/R)^alpha)^(gamma/alpha) exp(-bn (R^alpha + Rb^alpha) / (R_sersic^alpha alpha n_sersic) - 1)``
    with ``bn = 1.9992 n_sersic - 0.3271`` and ``R`` the elliptical distance of ``Sersic.distance``."""

    _name = "CORE_SERSIC"
    _params = ["R_sersic", "n_sersic", "Rb", "alpha", "gamma", "e1", "e2", "center_x", "center_y"]
    _amp = "Ie"
    _type_id = _cabi.GL_CORE_SERSIC
