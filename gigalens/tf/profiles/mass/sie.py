"""``src/gigalens/tf/profiles/mass/sie.py``."""
from gigalens_b200.profiles.mass.sie import SIE  # noqa: F401
