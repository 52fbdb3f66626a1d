#!/usr/bin/env python
"""A few cuts at real-problem sizes, for an ncu launch list (per-kernel durations of prep / sweep / merge)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
import stochasticdecomposition_b200 as sd

for D, N, rv, n1 in ((1000, 1000, 86, 89), (5000, 5000, 86, 89)):
    prob, pis, obsv, weights, xs = bench.make_workload(D, N, rv, n1, 0, 8)
    k = int(weights.sum())
    t = bench.load_tables(sd.load_library(), prob, pis, obsv, weights, D, N, k, 8)
    for s in range(6):
        t.sd_cut(xs[s], k, 1, 0.0)
    oi, onew = t.calc_omega(obsv[N], 1e-3)
    t.stochastic_updates(oi, onew, pis[D], 0.0, k, 1e-3)
    t.close()
print("ok")
