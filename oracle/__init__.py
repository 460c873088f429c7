"""TEST INFRASTRUCTURE ONLY: CPU oracle for the fully Bayesian GP hot path.

Nothing in the product package imports from here.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg do.
"""
