"""``src/gigalens/tf/profiles/mass/dpie_subhalo.py``."""
from gigalens_b200.profiles.mass.dpie_subhalo import DPIESubhalo  # noqa: F401
