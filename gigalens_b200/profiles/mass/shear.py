from ... import _cabi
from ...profile import MassProfile


class Shear(MassProfile):
    """External shear (reference ``tf/profiles/mass/shear.py:5-16``)."""

    _name = "SHEAR"
    _params = ["gamma1", "gamma2"]
    _type_id = _cabi.GL_SHEAR
