"""CPU oracle for the tn4ml hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``parity`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
(``tn4ml_b200``) never does.

Pinned by the reference's own source: the reference (tn4ml v1.1.1) holds no golden vectors or
known-answer tests for this path and cannot be imported as it is (jax, quimb, autoray, optax are
not installed), so ``tools/make_reference_golden.py`` executes ``/root/reference/tn4ml``
unmodified over minimal stand-ins for that stack (``tools/ref_shim``) and stores what its functions
return under ``tests/golden/reference_*.npz``; ``tests/test_reference_golden.py`` holds the oracle to
those vectors (feature maps, normaliser, MPS / classifier / SMPO conventions, loss heads, finite
differences of the executed losses).  Independent routes (dense contraction, finite differences, two
gradient routes, property tests, closed forms) are in ``tests/test_oracle.py``.
"""
