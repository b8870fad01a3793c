"""CPU oracle for the BGGE hot paths (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product path (bgge_b200) never does.
"""
from .bgge_oracle import (getK, BGGE, eig_block, setDEC, TapeDraws, RandomDraws,  # noqa: F401
                          quantile7, gb_kernel, gk_distance, gk_kernel, expand,
                          tape_sizes, design_codes)
