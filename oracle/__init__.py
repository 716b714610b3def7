"""CPU oracle of the pileup-and-call hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; nothing under ``trueconsense_b200/`` does
(tests/test_layout.py checks).  PARITY UNPINNED at the pysam/htslib boundary: see
oracle/pileup_oracle.c's header and DESIGN.md.
"""
