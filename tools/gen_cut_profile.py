#!/usr/bin/env python
"""A few cuts of a random-cost problem (multi-term bases + feasibility mask) at a size where the sweep dominates, for an ncu
capture of k_sweep_tma_gen."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import variant_bench

print(variant_bench.run(D=6144, N=65536, Q=0, rvd=4, phi=2, reps=6))
