from ... import _cabi
from ...profile import MassProfile


class DPIEP(MassProfile):
    """Dual pseudo-isothermal profile with the ellipticity in the potential (reference
    ``tf/profiles/mass/piep.py:17-56``): dPIS evaluated on stretched rotated coordinates.  The reference
    names it "dPIE" like the elliptical-mass profile; the parameter names (``Ra``, ``Rs``) differ."""

    _name = "dPIE"
    _params = ["theta_E", "Ra", "Rs", "center_x", "center_y", "e1", "e2"]
    _type_id = _cabi.GL_DPIEP
