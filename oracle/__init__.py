"""TEST INFRASTRUCTURE ONLY — CPU oracle for the KD-loss hot path.

Nothing in the product package may import this directory.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and only as the checker or as the timed CPU baseline.

Parity status: the reference ships NO golden vectors or tests for this path
(SURVEY.md §4, §8c), so the oracle is pinned by running the reference's own
Python functions imported from ``/root/reference`` in the dev container
(``oracle/ref_import.py``) and freezing their outputs under ``tests/golden/``
(``oracle/make_goldens.py``).
"""
