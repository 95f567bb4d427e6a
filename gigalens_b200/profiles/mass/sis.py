from ... import _cabi
from ...profile import MassProfile


class SIS(MassProfile):
    """Singular isothermal sphere (reference ``tf/profiles/mass/sis.py:5-17``)."""

    _name = "SIS"
    _params = ["theta_E", "center_x", "center_y"]
    _type_id = _cabi.GL_SIS
