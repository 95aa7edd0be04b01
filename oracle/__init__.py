"""CPU oracle for the rec-attend hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` may be imported by the product package
(``rec-attend-public_b200/``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, as the checker or
as the timed CPU baseline.

Parity status:
  * ``oracle.hungarian``  — PINNED by the reference's own known-answer tests
    (/root/reference/hungarian_tf_tests.py, fixtures in tests/golden/hungarian_kat.json).
  * ``oracle.model``      — PINNED TO THE REFERENCE'S OWN PYTHON SOURCE for the full model:
    TensorFlow 0.12 / Python 2.7 cannot run here, but full_model.py / nnlib.py /
    modellib.py / image_ops.py parse as Python 3, so tests/golden/make_*_golden.py execute
    them UNMODIFIED over a numpy stand-in for the TensorFlow-0.12 ops they use
    (tests/golden/tf012_shim) and the oracle must reproduce the results
    (full_model.get_model end to end to 1e-7 in float64; nnlib layer factories; modellib
    function by function).  On trust: TensorFlow's own kernel semantics, one line each in
    the shim.  box_model_forward and fg_model_forward are pinned the same way (the reference's
    box_model.get_model / fg_model.get_model; the latter needs its missing `image_ops_old`
    import aliased to the reference's image_ops.py).
"""
