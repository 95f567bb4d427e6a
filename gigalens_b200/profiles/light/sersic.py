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
