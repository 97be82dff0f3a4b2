"""CPU oracle for the `fade annotate` realignment path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (fade_b200) never does.  PARITY UNPINNED: see fade_oracle.h.
"""
