from ... import _cabi
from ...profile import LightProfile


class Shapelets(LightProfile):
    """Hermite-basis light profile (reference ``tf/profiles/light/shapelets.py:11-85``).

    Component order (n1, n2) = (0,0),(1,0),(0,1),(2,0),(1,1),(0,2),...; amplitude names are
    ``amp{k}`` zero-padded to ``len(str(n_layers))`` digits (``shapelets.py:26-46``).
    """

    _name = "SHAPELETS"
    _params = ["beta", "center_x", "center_y"]
    _type_id = _cabi.GL_SHAPELETS

    def __init__(self, n_max, use_lstsq=False, interpolate=True):
        super().__init__(use_lstsq=use_lstsq)
        if not use_lstsq:
            self.params.pop()  # drop the empty amp name appended by LightProfile
        self.n_max = int(n_max)
        self.n_layers = int((n_max + 1) * (n_max + 2) / 2)
        self.interpolate = bool(interpolate)
        self.N1, self.N2, self._amp_names = [], [], []
        decimal_places = len(str(self.n_layers))
        n1 = n2 = 0
        for i in range(self.n_layers):
            self._amp_names.append(f"amp{str(i).zfill(decimal_places)}")
            self.N1.append(n1)
            self.N2.append(n2)
            if n1 == 0:
                n1, n2 = n2 + 1, 0
            else:
                n1, n2 = n1 - 1, n2 + 1
        if not use_lstsq:
            self.params += self._amp_names
        self.depth = self.n_layers
