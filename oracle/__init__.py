"""ORACLE - test infrastructure only.

CPU restatements of the reference's perturbation hot path used to check the CUDA engine.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this package; the product package never does (it fails loudly without its CUDA library).
"""
