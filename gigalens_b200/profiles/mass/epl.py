from ... import _cabi
from ...profile import MassProfile


class EPL(MassProfile):
    """Elliptical power law (reference ``tf/profiles/mass/epl.py:5-17``); kernel: ``epl_fwd`` / ``epl_bwd``.

    Attributes:
        niter (int): cap on the number of terms of the angular series (``maximum_iterations``).
    """

    _name = "EPL"
    _params = ["theta_E", "gamma", "e1", "e2", "center_x", "center_y"]
    _type_id = _cabi.GL_EPL

    def __init__(self, niter=50):
        super().__init__()
        self.niter = int(niter)
