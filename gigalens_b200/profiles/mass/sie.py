from ... import _cabi
from ...profile import MassProfile


class SIE(MassProfile):
    """Singular isothermal ellipsoid (reference ``tf/profiles/mass/sie.py:5-42``; core forced to 0)."""

    _name = "SIE"
    _params = ["theta_E", "e1", "e2", "center_x", "center_y"]
    _type_id = _cabi.GL_SIE
