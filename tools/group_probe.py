#!/usr/bin/env python
"""Several sigmas / bases per lambda row: the grouped sweep (automatic, variant 0) against the plain TMA ring (variant 2) and the
LDG kernel (variant 1).  One JSON line per group size."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import stochasticdecomposition_b200 as sd  # noqa: E402
from stochasticdecomposition_b200._abi import Caps  # noqa: E402
from stochasticdecomposition_b200.synthetic import make_problem  # noqa: E402


FRESH_X = float(os.environ.get("PROBE_FRESH_X", "0.5"))      # half-width of the per-cut perturbation of x (0: the same x every cut)


def run(D, N, g, reps=8, R=40, n1=63):
    prob = make_problem(bench.SEED, rows=R + 8, cols=2 * R, n1=n1, n1c=n1, R=R, Rb=R, Q=0)
    rng = np.random.default_rng(bench.SEED)
    pis = rng.uniform(-1, 1, (D, prob.rows + 1)); pis[rng.random(pis.shape) < 0.3] = 0; pis[:, 0] = 0
    obs = rng.normal(0, 1, (N, prob.numRV + 1)); obs[:, 0] = 0
    w = (1 + rng.poisson(0.25, N)).astype(np.int32)
    k = int(w.sum())
    iters = np.ceil((np.arange(D) + 1) * (k / D)).astype(np.int32)
    t = sd.load_library(os.environ.get("PROBE_LIB")).create(prob, Caps(D + 2, g * D + 2, g * D + 2, N + 2, 1))
    t.omega_append_bulk(obs, w)
    t.update_dual_bulk(pis, None, iters, -1.0)
    t.calc_delta_block(0, D, 0, N)
    t.basis_append_bulk(iters, np.arange(D, dtype=np.int32))
    for e in range(1, g):                                   # further sigmas (and bases) on the same lambda rows
        for d in range(D):
            si, new = t.calc_sigma(pis[d], float(e) + 0.5, d, False, int(iters[d]), 1e-3)
            t.basis_append(int(iters[d]), True, [si])
    t.set_timing(True)
    x = rng.uniform(0, 1, prob.prevCols + 1); x[0] = 0
    out = {"lambda_rows": D, "bases": g * D, "observations": N, "x_perturbation": FRESH_X}
    cuts = {}
    for name, v in (("ldg", 1), ("tma", 2), ("auto", 0), ("grouped", 4)):
        t.set_sweep_variant(v)
        cuts[name] = t.sd_cut(x, k, 1, 0.0)
        ms = []
        xr = np.random.default_rng(7)
        for s in range(reps):
            if FRESH_X:                                     # a new first-stage point per cut (the seeded ring must not be timed on a repeated x)
                xs = x + xr.uniform(-FRESH_X, FRESH_X, x.shape); xs[0] = 0
            else:
                xs = x
            t.sd_cut(xs, k, 1, 0.0, want_istar=False)
            ms.append(t.stats()["last_sweep_ms"])
        m = float(np.median(ms))
        out[f"{name}_variant"] = t.stats()["last_sweep_variant"]
        out[f"{name}_cut_ms"] = round(t.stats()["last_cut_ms"], 4)
        out[f"{name}_sweep_ms"] = round(m, 4)
        out[f"{name}_pairs_per_s"] = float(f"{g * D * N / (m * 1e-3):.4g}")
        out[f"{name}_GBps_per_distinct_row"] = round(8.0 * D * N / (m * 1e-3) / 1e9, 1)
    for a in ("ldg", "tma", "grouped"):
        assert np.array_equal(cuts[a].iStar, cuts["auto"].iStar) and cuts[a].alpha == cuts["auto"].alpha
    out["identical"] = True
    t.close()
    return out


if __name__ == "__main__":
    shapes = [(int(a), int(b), int(c)) for a, b, c in (x.split("x") for x in sys.argv[1:])] or [(4096, 131072, g) for g in (1, 2, 4, 8)]
    for D, N, g in shapes:
        print(json.dumps(run(D, N, g)), flush=True)
