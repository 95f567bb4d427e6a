"""``src/gigalens/profile.py``: ``Parameterized``, ``LightProfile``, ``MassProfile``."""
from gigalens_b200.profile import LightProfile, MassProfile, Parameterized  # noqa: F401
