"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement / harness of the reference's fitting path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import anything from here, and only as the checker or the reported CPU
baseline -- never as part of the product path (``bodyfitting_b200``).
"""
