"""``src/gigalens/tf/profiles/mass/piep.py``."""
from gigalens_b200.profiles.mass.piep import DPIEP  # noqa: F401
