"""CPU oracle for the POD-UQNN hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy restatement of the reference's algorithm for the path
named in BASELINE.json (snapshot fields -> POD -> projection/reconstruction ->
ensemble predict).  Every function cites the reference file:line it follows.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, and there only as the checker or
the timed CPU baseline.  The product package ``pod_uqnn_b200`` never imports
``oracle`` and has no CPU fallback.

Parity pinning:
  * POD / snapshot loops / projection: pinned against the UNMODIFIED reference
    imported from ``/root/reference`` in the build container
    (``oracle/reference_loader.py``); the outputs are committed as fixtures in
    ``tests/golden/`` by ``oracle/gen_golden.py``.  The reference itself has no
    tests or golden vectors, and its SVD is third-party LAPACK ``dgesdd``
    (numba -> scipy cython_lapack, versions unpinned in requirements.txt:4-5).
  * Ensemble predict (varneuralnetwork.py / podnnmodel.py:314-379): TensorFlow
    2.1 / TFP 0.9 are absent from the image and cannot be installed, so
    ``oracle/ensemble.py`` is a line-by-line numpy restatement --
    PARITY UNPINNED for that part (no reference run, no reference tests).
"""
