"""``src/gigalens/inference.py``: the driver interface (here the concrete class serves as its own interface)."""
from gigalens_b200.inference import ModellingSequence as ModellingSequenceInterface  # noqa: F401
