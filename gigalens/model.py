"""``src/gigalens/model.py``: ``PhysicalModelBase``, ``ProbabilisticModel``; ``PhysicalModel`` as the reference's
``tests/conftest.py:8`` imports it."""
from gigalens_b200.model import PhysicalModel, PhysicalModelBase, ProbabilisticModel  # noqa: F401
