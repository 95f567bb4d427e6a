"""``src/gigalens/tf/profiles/mass/sis.py``."""
from gigalens_b200.profiles.mass.sis import SIS  # noqa: F401
