#!/usr/bin/env python
"""Sweep that recomputes delta.pib from (lambda, omega) (variant 3) against the streaming kernels (variant 0 with the automatic
recompute switched off by Rb, i.e. variants 1 / 2) for 1..8 random right-hand sides.  One JSON line per Rb and size."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import stochasticdecomposition_b200 as sd  # noqa: E402


def run(D, N, Rb, reps=10):
    prob, pis, obsv, weights, xs = bench.make_workload(D, N, Rb, 4, 0, 4)
    k = int(weights.sum())
    t = bench.load_tables(sd.load_library(), prob, pis, obsv, weights, D, N, k, 4)
    t.set_timing(True)
    out = {"D": D, "N": N, "Rb": Rb}
    cuts = {}
    for name, v in (("ldg", 1), ("tma", 2), ("recompute", 3)):
        t.set_sweep_variant(v)
        cuts[name] = t.sd_cut(xs[0], k, 1, 0.0)
        ms = []
        for s in range(reps):
            t.sd_cut(xs[s % 4], k, 1, 0.0, want_istar=False)
            ms.append(t.stats()["last_sweep_ms"])
        m = float(np.median(ms))
        out[f"{name}_sweep_us"] = round(m * 1e3, 1)
        out[f"{name}_pairs_per_s"] = float(f"{D * N / (m * 1e-3):.4g}")
    out["recompute_fp64_ops_per_s"] = float(f"{(2 * Rb + 2) * D * N / (out['recompute_sweep_us'] * 1e-6):.4g}")
    for a in ("ldg", "tma"):
        assert np.array_equal(cuts[a].iStar, cuts["recompute"].iStar) and cuts[a].alpha == cuts["recompute"].alpha
    out["identical"] = True
    t.close()
    return out


if __name__ == "__main__":
    for D, N in ((16384, 131072), (5000, 5000)):
        for Rb in (1, 2, 3, 4, 5, 6, 8):
            print(json.dumps(run(D, N, Rb)), flush=True)
