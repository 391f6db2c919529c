"""CPU oracle for the refractive rendering hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``samplenerfro_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use it, and only as the checker or the timed CPU baseline.
"""
