#!/usr/bin/env python
"""Per-call latency of the library at real-problem sizes (N, D <= a few thousand): what the reference's
"Argmax time" counter (twoSD.h:93) would see per iteration."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
import stochasticdecomposition_b200 as sd


def probe(D, N, rv, n1, reps=30):
    prob, pis, obsv, weights, xs = bench.make_workload(D, N, rv, n1, 0, reps + 8)
    k = int(weights.sum())
    t = bench.load_tables(sd.load_library(), prob, pis, obsv, weights, D, N, k, reps + 8)
    out = {"D": D, "N": N, "rv": rv, "n1": n1}
    t.set_timing(True)
    for want_istar in (True, False):
        for s in range(5):
            t.sd_cut(xs[s], k, 1, 0.0, want_istar=want_istar)
        w, dev, swp = [], [], []
        for s in range(reps):
            t0 = time.perf_counter()
            t.sd_cut(xs[s % 64], k, 1, 0.0, want_istar=want_istar)
            w.append(time.perf_counter() - t0)
            st = t.stats()
            dev.append(st["last_cut_ms"]); swp.append(st["last_sweep_ms"])
        tag = "istar" if want_istar else "noistar"
        out[f"cut_wall_us_{tag}"] = round(float(np.median(w)) * 1e6, 1)
        out[f"cut_dev_us_{tag}"] = round(float(np.median(dev)) * 1e3, 1)
        out[f"sweep_us_{tag}"] = round(float(np.median(swp)) * 1e3, 1)
    # the same call with pre-built ctypes arguments: what a C host sees (no numpy / wrapper overhead), event timing off
    import ctypes as C
    from stochasticdecomposition_b200._abi import CCut, _pf64, _pi32
    t.set_timing(False)
    beta = np.zeros(prob.prevCols + 1); istar = np.zeros(N + reps + 16, np.int32)
    cut = CCut(0.0, _pf64(beta), _pi32(istar), 0, 0, 0.0, 0.0)
    fn, ctx = t.api._fn("sd_cut"), t.ctx
    xs_c = [np.ascontiguousarray(xs[i]) for i in range(8)]
    xp = [_pf64(a) for a in xs_c]
    ref = C.byref(cut)
    for s in range(5):
        fn(ctx, xp[s % 8], k, 1, 0.0, ref)
    w = []
    for s in range(reps * 3):
        t0 = time.perf_counter(); fn(ctx, xp[s % 8], k, 1, 0.0, ref); w.append(time.perf_counter() - t0)
    out["cut_wall_us_raw_c_call"] = round(float(np.median(w)) * 1e6, 1)
    t.set_timing(True)
    out["launches_per_cut"] = t.stats()["last_cut_launches"]
    out["sweep_GBps"] = round(8 * D * N / (out["sweep_us_istar"] * 1e-6) / 1e9, 1)
    its = []
    for i in range(reps):
        t0 = time.perf_counter()
        oi, onew = t.calc_omega(obsv[N + i], 1e-3)
        t.stochastic_updates(oi, onew, pis[D + i], 0.0, k, 1e-3)
        its.append(time.perf_counter() - t0)
    out["update_wall_us"] = round(float(np.median(its)) * 1e6, 1)
    t.close()
    return out


if __name__ == "__main__":
    shapes = [(64, 64, 3, 4), (1000, 1000, 86, 89), (5000, 5000, 86, 89), (7500, 5000, 118, 121), (16384, 16384, 86, 89)]
    for D, N, rv, n1 in shapes:
        print(json.dumps(probe(D, N, rv, n1)), flush=True)
