"""``src/gigalens/tf/model.py``: ``ForwardProbModel``, ``BackwardProbModel``, ``PhysicalModel``."""
from gigalens_b200.model import BackwardProbModel, ForwardProbModel, PhysicalModel  # noqa: F401
