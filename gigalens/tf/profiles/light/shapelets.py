"""``src/gigalens/tf/profiles/light/shapelets.py``."""
from gigalens_b200.profiles.light.shapelets import Shapelets  # noqa: F401
