"""CPU oracle for the annchor_b200 hot path -- TEST INFRASTRUCTURE ONLY.

A numpy + plain-C restatement of gchq/annchor's ``Annchor.fit()`` path
(reference v1.1.0).  It exists to check the CUDA product; nothing under
``annchor_b200/`` imports it.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``.

Parity status: PINNED.  ``tests/test_oracle.py`` checks this package against
the reference's own known-answer tests, against its bundled exact 100-NN
fixtures, and against stage-by-stage outputs of the unmodified reference
``fit()`` captured by ``tests/golden/make_golden.py`` (run where
``/root/reference`` exists; the vectors are committed under ``tests/golden/``).
"""
from .clib import lib, build  # noqa: F401
from . import metrics  # noqa: F401
from .pipeline import OracleAnnchor, OracleBruteForce, compare_neighbor_graphs  # noqa: F401
