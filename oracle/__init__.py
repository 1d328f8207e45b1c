"""CPU oracle for the tn4ml hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
(``tn4ml_b200``) never does.

PARITY UNPINNED: the reference (tn4ml v1.1.1) holds no golden vectors or
known-answer tests for this path, and it cannot be imported here (jax, quimb,
autoray, optax are not installed).  The restatement is instead
validated against itself by four independent routes, see ``tests/test_oracle.py``.
"""
