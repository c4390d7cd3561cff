"""CPU oracle for the de novo k-mer hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  PARITY UNPINNED: see oracle/dnk_oracle.c header.
"""
from .oracle import *  # noqa: F401,F403
