"""TEST INFRASTRUCTURE -- placeholder: ``PixelGrid`` is imported by the reference's ``simulator.py:6`` but only used by the
legacy ``get_coords`` helper, which is out of scope."""


class PixelGrid:
    def __init__(self, *a, **k):
        raise NotImplementedError("tfshim: lenstronomy PixelGrid is not provided")
