from . import mass, light  # noqa: F401
