"""``src/gigalens/tf/profiles/mass/shear.py``."""
from gigalens_b200.profiles.mass.shear import Shear  # noqa: F401
