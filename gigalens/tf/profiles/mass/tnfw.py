"""``src/gigalens/tf/profiles/mass/tnfw.py``."""
from gigalens_b200.profiles.mass.tnfw import TNFW  # noqa: F401
