"""CPU checkers for the FIR hot path.  TEST INFRASTRUCTURE ONLY.

Importable from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline legs — never from the product package (warmup-fir-filter_b200/).
"""
