"""``src/gigalens/tf/profiles/mass/epl.py``."""
from gigalens_b200.profiles.mass.epl import EPL  # noqa: F401
