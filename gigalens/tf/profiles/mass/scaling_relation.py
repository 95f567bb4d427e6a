"""``src/gigalens/tf/profiles/mass/scaling_relation.py``."""
from gigalens_b200.profiles.mass.scaling_relation import ScalingRelation  # noqa: F401
