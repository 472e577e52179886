"""CPU oracle of the Prompt-Free-Diffusion hot path — TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (pfd_b200/) must never import it.  Parity pinned against the unmodified
reference by tools/make_golden.py (see tests/golden/oracle_pin_report.json).
"""
