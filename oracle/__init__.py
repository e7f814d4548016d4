"""CPU oracle for the render-layer + bilateral-solver hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import it, and only as the checker (or as the timed
CPU baseline), never as the thing shipped.  The product path
(``inverserenderingofindoorscene_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Parity status: the reference ships no tests or golden vectors, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the
build container by ``oracle/make_golden.py`` (imports ``/root/reference/models.py``
and ``/root/reference/BilateralGrid.py`` unmodified) and committed as small
fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks the
oracle against those fixtures on every run (no reference needed at test time).
"""
