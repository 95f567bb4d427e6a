"""CPU oracle for the gigalens hot path -- TEST INFRASTRUCTURE ONLY.

This package is a torch-CPU restatement of the reference's TensorFlow substrate
(``/root/reference/src/gigalens/tf/**`` plus ``src/gigalens/simulator.py``), written
function by function with the reference file:line each one follows.  It exists to
*check* the CUDA path; it is never the thing measured or shipped.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  Nothing under ``gigalens_b200/`` imports it.

PARITY PINNING STATUS
---------------------
tensorflow, tensorflow_probability, jax and lenstronomy are not installable here (no network), and the reference's own
tests hold no numeric golden vectors for images or log-probabilities (they compare against lenstronomy at run time).
The oracle is pinned in two ways.

1. **To the reference's own source code, executed.**  ``oracle/tfshim`` is a torch-backed stand-in for the ~60 ``tf.*``
   functions the reference's profile / simulator / model files call; with it on ``sys.path`` the UNMODIFIED files under
   ``/root/reference/src/gigalens/`` import and run (``tests/golden/make_reference_golden.py``), once with ``tf.float32`` =
   float32 and once = float64.  The committed vectors (``tests/golden/reference_golden.npz``: ``deriv`` / ``hessian`` / ``light``
   of every profile incl. the clamp / branch edge cases, ``simulate`` and its variants on five models up to BASELINE's cluster
   geometry, ``stats_pixels`` / ``stats_positions`` with autodiff gradients, ``beta`` / magnification / convergence / shear) are
   reproduced by this oracle to 7e-15 (float64) and to float32 round-off (``tests/test_reference_golden.py``), and by the
   CUDA path within the parity rule (``tests/test_gpu_reference_golden.py``).  What that pins: every clamp, ``where``, sign,
   operation order and loop of the reference's files.  What it does not: TensorFlow's own kernels (represented by torch's).
2. By reference-held fixtures and independent mathematics: the framework-free known-answer tests of
   ``tests/test_profiles.py``, cross-profile identities, the ``demo.npy`` image-level pin (reduced chi^2 of
   ``simulate(truth)`` = 0.981 -- the executed reference gives the same number), the EPL series against the hypergeometric
   closed form, div(alpha) = 2 kappa, Sersic total flux, Shapelets orthonormality, fp64 finite differences.

Still **parity unpinned** (third-party arithmetic that is neither vendored in the reference nor installable): the TFP prior /
bijector values (``oracle/model.py``), lenstronomy's ``subgrid_kernel`` and ``phi_n`` (restated; the stand-in borrows the
restatements), and ``tf.linalg.pinv``'s own kernel.  The reference's ``lstsq_simulate`` cannot execute as a whole in either substrate
(``tf/simulator.py:183-203`` scatters into a zero-sized buffer); (1) executes everything else of it -- the reference's simulator
object, ``beta``, ``light`` calls and its source lines ``:204-240`` -- restating only those three scatter lines.  DESIGN.md says the same.
"""
