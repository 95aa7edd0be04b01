"""CPU oracle for the rec-attend hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` may be imported by the product package
(``rec-attend-public_b200/``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, as the checker or
as the timed CPU baseline.

Parity status:
  * ``oracle.hungarian``  — PINNED by the reference's own known-answer tests
    (/root/reference/hungarian_tf_tests.py, fixtures in tests/golden/hungarian_kat.json).
  * ``oracle.model``      — the graph as a whole is PARITY UNPINNED: it runs only on
    TensorFlow 0.12 / Python 2.7, neither of which exists here, and the reference
    ships no golden outputs for it; it is a structure-faithful restatement validated by
    semantic unit tests (tests/test_model_oracle.py).  Its math library (the functions
    restating modellib.py) IS pinned: tests/golden/make_modellib_golden.py executes the
    reference's own modellib.py, unmodified, over a numpy stand-in for the ~30 TF ops
    it uses, and tests/test_modellib_golden.py holds the oracle to those vectors.
"""
