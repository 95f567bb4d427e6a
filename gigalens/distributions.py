"""Stand-in for ``tensorflow_probability.distributions`` in prior specs (``tfd.Normal``, ``LogNormal``, ``Uniform``,
``TruncatedNormal``, ``JointDistributionNamed`` / ``Sequential``): SURVEY.md App. C."""
from gigalens_b200.distributions import *  # noqa: F401,F403
from gigalens_b200.distributions import (JointDistribution, JointDistributionNamed, JointDistributionSequential, LogNormal,  # noqa: F401
                                         Normal, TruncatedNormal, Uniform)
