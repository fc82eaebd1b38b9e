"""CPU oracles for the DPRT hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker or as the timed CPU baseline.  Nothing under
``dpft_b200/`` imports it (tests/test_boundary.py enforces that).
"""
