#!/usr/bin/env python
"""Sweep time against the number of basis chunks (SDGPU_CHUNKS) at one table size: run once per value.
usage: SDGPU_CHUNKS=n tools/chunk_probe.py D N"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import stochasticdecomposition_b200 as sd  # noqa: E402

D, N = int(sys.argv[1]), int(sys.argv[2])
prob, pis, obsv, weights, xs = bench.make_workload(D, N, 86, 89, 0, 8)
k = int(weights.sum())
t = bench.load_tables(sd.load_library(), prob, pis, obsv, weights, D, N, k, 8)
t.set_timing(True)
sw, ct = [], []
for s in range(24):
    t.sd_cut(xs[s % 8], k, 1, 0.0, want_istar=False)
    if s >= 4:
        st = t.stats(); sw.append(st["last_sweep_ms"]); ct.append(st["last_cut_ms"])
print(json.dumps({"D": D, "N": N, "chunks_env": os.environ.get("SDGPU_CHUNKS"), "sweep_us": round(float(np.median(sw)) * 1e3, 1),
                  "cut_us": round(float(np.median(ct)) * 1e3, 1), "GBps": round(8.0 * D * N / (float(np.median(sw)) * 1e-3) / 1e9, 1)}), flush=True)
t.close()
