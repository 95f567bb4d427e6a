"""``src/gigalens/tf/profile.py``: ``MassProfile`` with ``hessian`` / ``convergence`` / ``shear``."""
from gigalens_b200.profile import LightProfile, MassProfile  # noqa: F401
