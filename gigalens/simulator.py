"""``src/gigalens/simulator.py``: ``SimulatorConfig``, ``LensWCS``, ``LensSimulatorInterface``."""
from gigalens_b200.simulator import LensSimulatorInterface, LensWCS, SimulatorConfig  # noqa: F401
