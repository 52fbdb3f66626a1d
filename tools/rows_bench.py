#!/usr/bin/env python
"""Timings of the table-append kernels and of the "next" rows (bootstrap reformCuts, checkBasisFeasibility), CUDA library
vs the reference's own C on the same box.  One JSON line per item."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import oracle_loader  # noqa: E402
import stochasticdecomposition_b200 as sd  # noqa: E402
from stochasticdecomposition_b200._abi import Caps  # noqa: E402
from stochasticdecomposition_b200.synthetic import make_problem  # noqa: E402


def med(f, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def appends(D, N, rv=256, n1=89, reps=20):
    prob, pis, obsv, weights, xs = bench.make_workload(D, N, rv, n1, 0, 3 * reps + 8)
    k = int(weights.sum())
    t0 = time.perf_counter()
    t = bench.load_tables(sd.load_library(), prob, pis, obsv, weights, D, N, k, 3 * reps + 8)
    build = time.perf_counter() - t0
    it = iter(range(10 ** 6))
    def new_obs():
        i = next(it)
        oi, onew = t.calc_omega(obsv[N + i], 1e-3)
        t.calc_delta(True, oi)
        t.counts()
    def new_dual():
        i = next(it)
        t.update_dual(pis[D + i], 0.0, k, 1e-3)
    o_s, d_s = med(new_obs, reps), med(new_dual, reps)
    R = prob.rvRowCnt
    out = {"item": "append", "duals": D, "observations": N, "R": R,
           "bulk_build_s": round(build, 3), "bulk_fp64_TFLOPs": round(2.0 * D * N * R / build / 1e12, 2),
           "new_observation_us": round(o_s * 1e6, 1), "new_observation_alg_GBps": round(8.0 * (D * R + R + D) / o_s / 1e9, 1),
           "new_dual_us": round(d_s * 1e6, 1), "new_dual_alg_GBps": round(8.0 * (N * R + R + N + D * R) / d_s / 1e9, 1)}
    t.close()
    return out


def bootstrap(K=2000, n_cuts=20, reps=50):
    prob = make_problem(9, rows=175, cols=706, n1=89, n1c=89, R=86, Rb=86)
    rng = np.random.default_rng(1)
    pis = rng.uniform(-1, 1, (K, prob.rows + 1)); pis[:, 0] = 0
    obs = rng.normal(0, 1, (K, prob.numRV + 1)); obs[:, 0] = 0
    iters = np.arange(1, K + 1, dtype=np.int32)
    istars = [rng.integers(0, K, K).astype(np.int32) for _ in range(n_cuts)]
    observ = rng.integers(0, K, (reps, K)).astype(np.int32)
    res = {}
    for name, api in (("gpu", sd.load_library()), ("reference", oracle_loader.reference())):
        t = api.create(prob, Caps(K + 2, K + 2, K + 2, K + 2, 1))
        if name == "gpu":
            t.omega_append_bulk(obs, None); t.update_dual_bulk(pis, None, iters, -1.0); t.calc_delta_block(0, K, 0, K)
            t.basis_append_bulk(iters, np.arange(K, dtype=np.int32))
        else:
            from stochasticdecomposition_b200._abi import _pf64, _pi32
            li, si = np.zeros(K, np.int32), np.zeros(K, np.int32)
            api._fn("bulk_load")(t.ctx, K, _pf64(obs), None, K, _pf64(pis), None, _pi32(iters), -1.0, _pi32(li), _pi32(si))
            for b in range(K):
                t.basis_append(int(iters[b]), True, [b])
        t.reform_cuts_batch(istars, observ[:2], 0, 0)
        res[name] = med(lambda: t.reform_cuts_batch(istars, observ, 0, 0), 3)
        res[name + "_out"] = t.reform_cuts_batch(istars, observ[:3], 0, 0)
    a, b = res["gpu_out"], res["reference_out"]
    assert np.abs(a[0] - b[0]).max() <= 1e-9 * np.abs(b[0]).max() and np.abs(a[1] - b[1]).max() <= 1e-9 * np.abs(b[1]).max()
    return {"item": "bootstrap reformCuts (optimal.c:96-103)", "k": K, "cuts": n_cuts, "replications": reps, "n1": 89,
            "gpu_ms": round(res["gpu"] * 1e3, 3), "reference_cpu_ms": round(res["reference"] * 1e3, 3),
            "speedup": round(res["reference"] / res["gpu"], 1), "parity": "within 1e-9"}


def feasibility(B=1500, N=1500, rvd=4):
    import feas_scenario  # noqa: F401  (same data conventions)
    prob = make_problem(13, rows=528, cols=1259, n1=121, n1c=121, R=118, Rb=118, Q=0, rvd=rvd)
    rng = np.random.default_rng(2)
    rows, cols = prob.rows, prob.cols
    obs = rng.normal(0, 1, (N, prob.numRV + 1)); obs[:, 0] = 0; obs[:, prob.rvOffset[2] + 1:] *= 0.15
    rvdOmCols = np.concatenate([[0], np.sort(rng.choice(np.arange(1, cols + 1), rvd, replace=False))]).astype(np.int32)
    senx = bytes(rng.choice([ord("G"), ord("L"), ord("E")], rows).tolist())
    out = {}
    data = []
    for b in range(B):
        pl = int(rng.integers(0, rvd + 1))
        data.append((pl, rng.normal(0, 0.1, rows + 1), rng.uniform(-0.5, 0.5, (pl, rows + 1)), np.abs(rng.normal(0.25, 0.1, cols + 1)),
                     rng.uniform(-0.5, 0.5, (cols, pl)), np.concatenate([[0], rng.integers(0, 3, cols)]).astype(np.int32),
                     [0] + sorted(rng.choice(np.arange(1, rvd + 1), pl, replace=False).tolist())))
    flags = {}
    for name, api in (("gpu", sd.load_library()), ("reference", oracle_loader.reference())):
        t = api.create(prob, Caps(B * (1 + rvd) + 2, B * (1 + rvd) + 2, B + 2, N + 2, 1 + rvd))
        t.set_cost_coords(rvdOmCols, senx)
        for i in range(N):
            t.calc_omega(obs[i], -1.0 if False else 1e-9)
        pi = np.zeros(rows + 1); pi[1] = 1.0
        s0 = t.update_dual(pi, 0.0, 1, 1e-3)[2]
        for b, (pl, piDet, phi, gBar, psi, cstat, om) in enumerate(data):
            bi = t.basis_append(1, True, [s0] * (pl + 1), om if pl else None)
            t.basis_set_feas_data(bi, piDet, phi if pl else None, gBar, psi.ravel() if pl else None, cstat)
        out[name + "_obs_ms"] = med(lambda: t.check_feasibility_obs(N - 1, 1e-3), 3) * 1e3
        out[name + "_basis_ms"] = med(lambda: t.check_feasibility_basis(B - 1, 1e-3), 3) * 1e3
        flags[name] = (t.check_feasibility_obs(N - 1, 1e-3).copy(), t.check_feasibility_basis(B - 1, 1e-3).copy())
    assert np.array_equal(flags["gpu"][0], flags["reference"][0]) and np.array_equal(flags["gpu"][1], flags["reference"][1])
    return {"item": "checkBasisFeasibility (randCost.c:202-258), storm shape", "bases": B, "observations": N, "rvdOmCnt": rvd,
            **{k: round(v, 3) for k, v in out.items()}, "feasible_share": round(float(flags["gpu"][1].mean()), 3), "parity": "flags identical"}


if __name__ == "__main__":
    print(json.dumps(appends(65536, 131072)), flush=True)
    print(json.dumps(appends(5000, 5000, rv=86)), flush=True)
    print(json.dumps(bootstrap()), flush=True)
    print(json.dumps(feasibility()), flush=True)
