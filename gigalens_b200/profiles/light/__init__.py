from . import sersic, shapelets  # noqa: F401
