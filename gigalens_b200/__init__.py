"""gigalens_b200 -- B200-native batched differentiable forward model with the API surface of
GIGA-Lens (furcelay/gigalens, ``cluster-lens`` fork).

Host code (this package) keeps the reference's ``SimulatorConfig`` / ``LensSimulator`` /
``PhysicalModel`` / ``ForwardProbModel`` / ``BackwardProbModel`` / ``ModellingSequence`` names and
argument meaning; all arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/gigalens_b200.h``.  There is no CPU fallback.
"""
__version__ = "0.1.0"
