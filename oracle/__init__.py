"""CPU oracle of the RAGraph hot path -- TEST INFRASTRUCTURE ONLY (see ragraph_oracle.py).

Importable from ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of ``bench.py``;
nothing under ``ragraph_b200/`` may import it.
"""
