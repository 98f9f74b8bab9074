"""ORACLE — test infrastructure only.

CPU float64 restatement of the reference's MPC path (mpc.py, conic_ipm.py) and
postprocessing path (postprocessing.py).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package; nothing under ``adacharge_b200/`` does.
"""
