"""``src/gigalens/tf/profiles/mass/piemd.py``."""
from gigalens_b200.profiles.mass.piemd import DPIE, DPIS  # noqa: F401
