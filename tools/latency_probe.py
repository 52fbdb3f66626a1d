#!/usr/bin/env python
"""Per-call latency of the library at real-problem sizes (N, D <= a few thousand): what the reference's "Argmax time" counter
(twoSD.h:93) would see per iteration.  Raw C calls with pre-built ctypes arguments (what a C host sees), wall clock, medians.

For every shape the cut and the table update are timed under each combination of the three latency features of round 2 --
programmatic dependent launch between the kernels of a cut (SDGPU_PDL), the alternating direction of the load-based sweep
(SDGPU_ALTDIR: L2 re-use on tables a little larger than L2) and the fused update launch (SDGPU_FUSED_UPDATE) -- and the cuts of all
combinations are compared bit for bit (alpha, beta, iStar): the features must not change a single result bit."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
import stochasticdecomposition_b200 as sd
from stochasticdecomposition_b200._abi import CCut, _pf64, _pi32


def probe(D, N, rv, n1, pdl, alt, fused, l2keep=0, chunks=0, reps=40):
    os.environ["SDGPU_PDL"], os.environ["SDGPU_ALTDIR"], os.environ["SDGPU_FUSED_UPDATE"] = str(pdl), str(alt), str(fused)
    os.environ["SDGPU_CHUNKS"] = str(chunks)
    api = sd.load_library()
    prob, pis, obsv, weights, xs = bench.make_workload(D, N, rv, n1, 0, reps + 8)
    k = int(weights.sum())
    t = bench.load_tables(api, prob, pis, obsv, weights, D, N, k, reps + 8)
    out = {"D": D, "N": N, "rv": rv, "n1": n1, "pdl": pdl, "altdir": alt, "fused_update": fused, "chunks": chunks}
    beta = np.zeros(prob.prevCols + 1); istar = np.zeros(N + reps + 16, np.int32)
    cut = CCut(0.0, _pf64(beta), _pi32(istar), 0, 0, 0.0, 0.0)
    fn, ctx = api._fn("sd_cut"), t.ctx
    xs_c = [np.ascontiguousarray(xs[i]) for i in range(8)]
    xp = [_pf64(a) for a in xs_c]
    ref = C.byref(cut)
    # device-side split with event timing (PDL is off while events sit between the kernels)
    t.set_timing(True)
    for s in range(5):
        fn(ctx, xp[s % 8], k, 1, 0.0, ref)
    split = {"prep": [], "sweep": [], "merge": [], "cut": []}
    for s in range(reps):
        fn(ctx, xp[s % 8], k, 1, 0.0, ref)
        st = t.stats()
        split["prep"].append(st["last_prep_ms"]); split["sweep"].append(st["last_sweep_ms"]); split["merge"].append(st["last_merge_ms"]); split["cut"].append(st["last_cut_ms"])
    out.update({f"dev_{k_}_us": round(float(np.median(v)) * 1e3, 1) for k_, v in split.items()})
    out["launches_per_cut"] = t.stats()["last_cut_launches"]
    # wall clock of the raw C call, event timing off
    t.set_timing(False)
    for s in range(6):
        fn(ctx, xp[s % 8], k, 1, 0.0, ref)
    w = []
    for s in range(reps * 3):
        t0 = time.perf_counter(); fn(ctx, xp[s % 8], k, 1, 0.0, ref); w.append(time.perf_counter() - t0)
    out["cut_wall_us"] = round(float(np.median(w)) * 1e6, 1)
    out["cut_wall_us_p10"] = round(float(np.percentile(w, 10)) * 1e6, 1)
    # results for the bit-for-bit comparison across feature combinations: one cut per x, twice (both sweep directions)
    sig = []
    for s in range(4):
        fn(ctx, xp[s % 2], k, 1, 0.0, ref)
        sig.append((cut.alpha, beta.tobytes(), istar[:N].tobytes(), cut.cummOld, cut.cummAll))
    out["_sig"] = sig
    # the table update of one iteration: calcOmega, then delta column + calcLambda + calcSigma + delta row + basis record
    calc_omega, upd, basis = api._fn("calc_omega"), api._fn("update_dual_col"), api._fn("basis_find_or_append")
    obs_c = [np.ascontiguousarray(obsv[N + i]) for i in range(reps)]
    pi_c = [np.ascontiguousarray(pis[D + i]) for i in range(reps)]
    flag, li, nl, si, ns, nb = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    sig32 = (C.c_int32 * 1)(0)
    wo, wu = [], []
    for i in range(reps):
        t0 = time.perf_counter()
        oi = calc_omega(ctx, _pf64(obs_c[i]), 1e-3, C.byref(flag))
        t1 = time.perf_counter()
        upd(ctx, oi if flag.value else -1, _pf64(pi_c[i]), 0.0, k, 1e-3, C.byref(li), C.byref(nl), C.byref(si), C.byref(ns))
        sig32[0] = si.value
        basis(ctx, ns.value, oi, k, 1, 0, sig32, None, C.byref(nb))
        t2 = time.perf_counter()
        wo.append(t1 - t0); wu.append(t2 - t1)
    out["calc_omega_wall_us"] = round(float(np.median(wo)) * 1e6, 1)
    out["stochastic_updates_wall_us"] = round(float(np.median(wu)) * 1e6, 1)
    out["update_wall_us"] = round(float(np.median(np.add(wo, wu))) * 1e6, 1)
    out["_counts"] = t.counts()
    t.close()
    return out


if __name__ == "__main__":
    shapes = [(1000, 1000, 86, 89), (5000, 5000, 86, 89), (7500, 5000, 118, 121)]
    if "--quick" in sys.argv:
        shapes = shapes[1:2]
    if "--floor" in sys.argv or "--quick" not in sys.argv:
        api = sd.load_library()
        floor = {f"{l}_launch_{'poll' if m else 'sync'}_us": round(api.launch_roundtrip(m, l, 400), 2) for l in (1, 2, 3) for m in (0, 1)}
        print(json.dumps({"launch_roundtrip_floor": floor}), flush=True)
        if "--floor" in sys.argv:
            sys.exit(0)
    combos = [(0, 0, 0, 0, 0), (1, 0, 0, 0, 0), (0, 1, 0, 0, 0), (1, 1, 0, 0, 0), (1, 1, 1, 0, 0)]
    for D, N, rv, n1 in shapes:
        base = None
        for pdl, alt, fused, l2keep, chunks in combos:
            r = probe(D, N, rv, n1, pdl, alt, fused, l2keep, chunks)
            sig, counts = r.pop("_sig"), r.pop("_counts")
            if base is None:
                base = (sig, counts)
            r["bit_identical_to_baseline"] = bool(sig == base[0] and counts == base[1])
            print(json.dumps(r), flush=True)
