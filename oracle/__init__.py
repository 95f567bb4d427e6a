"""CPU oracle for the gigalens hot path -- TEST INFRASTRUCTURE ONLY.

This package is a torch-CPU restatement of the reference's TensorFlow substrate
(``/root/reference/src/gigalens/tf/**`` plus ``src/gigalens/simulator.py``), written
function by function with the reference file:line each one follows.  It exists to
*check* the CUDA path; it is never the thing measured or shipped.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  Nothing under ``gigalens_b200/`` imports it.

PARITY PINNING STATUS
---------------------
The reference cannot be imported here (tensorflow, tensorflow_probability, jax and
lenstronomy are not installed and there is no network), and the reference's own tests
hold no numeric golden vectors for images or log-probabilities (they compare against
lenstronomy at run time).  The oracle is therefore pinned by:

* the framework-free known-answer tests recoverable from ``tests/test_profiles.py``
  (Sersic half-light identity, EPL(gamma=2,e=0) == SIS == (x/r, y/r), Shear closed form,
  Shapelets interpolate == recurrence),
* cross-profile identities (EPL(gamma=2) == SIE, NFW_ELLIPSE(e=0) == NFW, dPIE(e->0) -> dPIS,
  ScalingRelation(G=1, L=L*) == DPIE),
* the one image-level pin that exists: reduced chi^2 of ``simulate(truth)`` against the
  reference asset ``demo.npy`` (tf-demo.ipynb cells 5-9) must be ~1,
* autograd-vs-finite-difference checks in float64.

Everything that lives in third-party code (TFP bijectors/distributions, TF conv/pool/pinv
semantics, lenstronomy ``subgrid_kernel``/``phi_n``) is restated from its published
behaviour and is **parity unpinned** against the real libraries; DESIGN.md says the same.
"""
