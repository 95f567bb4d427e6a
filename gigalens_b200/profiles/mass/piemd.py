from ... import _cabi
from ...profile import MassProfile


class DPIS(MassProfile):
    """Dual pseudo-isothermal sphere (reference ``tf/profiles/mass/piemd.py:21-60``)."""

    _name = "dPIS"
    _params = ["theta_E", "r_core", "r_cut", "center_x", "center_y"]
    _type_id = _cabi.GL_DPIS


class DPIE(MassProfile):
    """Dual pseudo-isothermal elliptical profile (reference ``tf/profiles/mass/piemd.py:97-119,183-255``).
    Singular at e = 0 like the reference (1/sqrt(e) prefactor)."""

    _name = "dPIE"
    _params = ["theta_E", "r_core", "r_cut", "center_x", "center_y", "e1", "e2"]
    _type_id = _cabi.GL_DPIE
