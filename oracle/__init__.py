"""CPU oracle for the GIST aggregation hot path — TEST INFRASTRUCTURE ONLY.

Nothing in ``gist_b200`` imports this package.  Allowed users: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl reference legs.
See oracle/README.md for the parity status of each piece.
"""
