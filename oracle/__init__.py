"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the SelfC-large rescaling hot path (see selfc_oracle.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product path (selfc_b200/) never
does: it fails loudly when the CUDA library is missing.
"""
