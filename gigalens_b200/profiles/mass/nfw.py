from ... import _cabi
from ...profile import MassProfile


class NFW(MassProfile):
    """Spherical NFW (reference ``tf/profiles/mass/nfw.py:5-52``)."""

    _name = "NFW"
    _params = ["Rs", "alpha_Rs", "center_x", "center_y"]
    _type_id = _cabi.GL_NFW


class NFW_ELLIPSE(MassProfile):
    """NFW with ellipticity in the potential (reference ``tf/profiles/mass/nfw.py:97-134``)."""

    _name = "NFW_ELLIPSE"
    _params = ["Rs", "alpha_Rs", "e1", "e2", "center_x", "center_y"]
    _type_id = _cabi.GL_NFW_ELLIPSE
