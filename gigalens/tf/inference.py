"""``src/gigalens/tf/inference.py``: ``ModellingSequence`` (MAP / SVI / HMC / SMC)."""
from gigalens_b200.inference import Adam, ModellingSequence, PolynomialDecay  # noqa: F401
