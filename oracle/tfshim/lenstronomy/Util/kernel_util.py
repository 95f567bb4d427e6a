"""TEST INFRASTRUCTURE -- ``lenstronomy.Util.kernel_util.subgrid_kernel`` is third-party code that is not vendored in the
reference and not installable here; the name resolves to the oracle's restatement of its published algorithm
(``oracle/simulator.py``), so this piece stays "restated", not pinned."""
from oracle.simulator import subgrid_kernel  # noqa: F401
