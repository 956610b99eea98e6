"""ORACLE (test infrastructure, not product code): CPU float64 restatement of the reference's ``World3D.step`` hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may import this package; the product
(``diffsdfsim_b200``) never does.  Pinned against the unmodified reference through ``tests/golden``.
"""
