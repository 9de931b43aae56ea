"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the orbital-optimization hot path.  Nothing under
``auto_oo_b200`` may import this package; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may.
"""
