"""``src/gigalens/tf/simulator.py``: ``LensSimulator``."""
from gigalens_b200.simulator import LensSimulator  # noqa: F401
