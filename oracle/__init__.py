"""CPU oracle for the BPR-MF / funk-SVD hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of what the reference
(NotFoundGG/recommend-lib, "Daisy") computes on the hot path.  It exists so
that the CUDA path can be checked against it.  It is **not** part of the
product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  Nothing
in ``recommend_lib_b200/`` imports this package.

Parity pinning
--------------
The reference ships no tests, golden vectors or known-answer fixtures
(SURVEY.md section 4), so this oracle is pinned against *outputs of the
reference itself run in the build container*:

* ``tests/golden/make_golden.py`` imports the unmodified ``BPR`` class
  (``BPRMFRecommender.py:28-50``) and ``_bpr_topk`` (``util/metrics.py:46-66``)
  from ``/root/reference`` and records step / eval fixtures under
  ``tests/golden/``;
* ``oracle/build_ref.py`` compiles the reference's own
  ``util/matrix_factorization.pyx`` (where it lies) into ``oracle/_ref/`` and the
  same script records ``SVD`` / ``RSVD`` fit fixtures.

``tests/test_oracle_golden.py`` checks every function here against those
fixtures, on CPU.
"""
