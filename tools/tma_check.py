#!/usr/bin/env python
"""Sweep variant 2 (TMA bulk ring) against variant 1 (LDG): identical iStar / cut, and timing of both."""
import json
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
import stochasticdecomposition_b200 as sd

def run(D, N, rv=64, n1=40, reps=10):
    prob, pis, obsv, weights, xs = bench.make_workload(D, N, rv, n1, 0, 4)
    k = int(weights.sum())
    t = bench.load_tables(sd.load_library(), prob, pis, obsv, weights, D, N, k, 4)
    t.set_timing(True)
    out = {"D": D, "N": N}
    cuts = {}
    for v in (1, 2):
        t.set_sweep_variant(v)
        for pe in (0, 1):
            cuts[(v, pe)] = t.sd_cut(xs[0], k, pe, 0.0)
        ms = []
        for s in range(reps):
            t.sd_cut(xs[s], k, 1, 0.0, want_istar=False)
            ms.append(t.stats()["last_sweep_ms"])
        out[f"sweep_ms_v{v}"] = round(float(np.median(ms)), 4)
        out[f"GBps_v{v}"] = round(8 * D * N / (np.median(ms) * 1e-3) / 1e9, 1)
    for pe in (0, 1):
        a, b = cuts[(1, pe)], cuts[(2, pe)]
        assert np.array_equal(a.iStar, b.iStar), ("iStar differs", pe)
        assert a.alpha == b.alpha and np.array_equal(a.beta, b.beta), ("cut differs", pe)
    out["identical"] = True
    t.close()
    return out

if __name__ == "__main__":
    import os
    shapes = [(int(a), int(b)) for a, b in (x.split("x") for x in sys.argv[1:])] or [(100, 700), (3000, 5000), (8192, 65536), (65536, 131072)]
    for D, N in shapes:
        r = run(D, N)
        r["cfg"] = os.environ.get("SDGPU_TMA_CFG", "0")
        print(json.dumps(r), flush=True)
