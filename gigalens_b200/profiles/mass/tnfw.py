from ... import _cabi
from ...profile import MassProfile


class TNFW(MassProfile):
    """Truncated NFW halo (reference ``tf/profiles/mass/tnfw.py:10-62``; the reference docs call it
    experimental, ``docsrc/source/profiles.rst:41-46``)."""

    _name = "TNFW"
    _params = ["Rs", "alpha_Rs", "r_trunc", "center_x", "center_y"]
    _type_id = _cabi.GL_TNFW
