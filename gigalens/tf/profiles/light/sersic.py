"""``src/gigalens/tf/profiles/light/sersic.py``."""
from gigalens_b200.profiles.light.sersic import CoreSersic, Sersic, SersicEllipse  # noqa: F401
