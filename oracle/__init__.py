"""Deep-FIR hot-path ORACLE — test infrastructure only.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
(or as the timed CPU baseline) — never on the shipped GPU path.
"""
