"""Import-compatibility surface: ``gigalens.*`` module paths of the reference (furcelay/gigalens,
``src/gigalens/``) re-exporting the B200-native classes of :mod:`gigalens_b200`.

A script written against the reference keeps its ``gigalens.tf.simulator`` / ``gigalens.tf.model`` /
``gigalens.tf.inference`` / ``gigalens.tf.profiles.{mass,light}.*`` / ``gigalens.model`` / ``gigalens.simulator``
imports unchanged; only its TensorFlow-Probability prior spec is swapped for the shim
(``from gigalens import distributions as tfd``, same constructor names and argument order), see INTEGRATION.md.
Nothing here computes: every class is the one defined in ``gigalens_b200`` (CUDA behind the C ABI, no CPU path).
The JAX substrate (``gigalens.jax``) and the Taylor-series profiles (``*_series``) are out of scope.
"""
from gigalens_b200 import __version__  # noqa: F401
