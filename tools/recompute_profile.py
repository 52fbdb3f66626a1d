#!/usr/bin/env python
"""A few cuts of a 3-random-RHS problem (pgp2's width) at 16 384 x 131 072 for an ncu capture of k_sweep_recompute<3>."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import recompute_probe
print(recompute_probe.run(16384, 131072, 3, reps=3))
