"""``src/gigalens/tf/profiles/mass/nfw.py``."""
from gigalens_b200.profiles.mass.nfw import NFW, NFW_ELLIPSE  # noqa: F401
