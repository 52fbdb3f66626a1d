#!/usr/bin/env python
"""The synthetic cut-formation grid of SURVEY.md section 8d on one GPU: duals x observations x random elements x Q,
pairs/s and algorithmic GB/s of the sweep kernel against the measured HBM peak.  Configurations whose delta table
exceeds --max-gib are skipped (they need more GPUs).  One JSON line per configuration."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import stochasticdecomposition_b200 as sd  # noqa: E402
from stochasticdecomposition_b200._abi import Caps  # noqa: E402
from stochasticdecomposition_b200.synthetic import make_problem  # noqa: E402


def run(D, N, R, Q, n1, two_sigma=False, reps=5):
    prob = make_problem(bench.SEED, rows=max(R, 8) + 8, cols=2 * max(R, 8), n1=n1, n1c=n1, R=R, Rb=R, Q=Q)
    rng = np.random.default_rng(bench.SEED)
    pis = rng.uniform(-1, 1, (D, prob.rows + 1)); pis[rng.random(pis.shape) < 0.3] = 0; pis[:, 0] = 0
    w = (1 + rng.poisson(0.25, N)).astype(np.int32)
    k = int(w.sum())
    iters = np.ceil((np.arange(D) + 1) * (k / D)).astype(np.int32)
    S = 2 * D if two_sigma else D
    t0 = time.perf_counter()
    t = sd.load_library().create(prob, Caps(D + 2, S + 2, S + 2, N + 2, 1))
    chunk = 1 << 17
    for o0 in range(0, N, chunk):                                     # observations generated chunk by chunk (host memory)
        n = min(chunk, N - o0)
        obs = rng.normal(0, 1, (n, prob.numRV + 1)); obs[:, 0] = 0
        t.omega_append_bulk(obs, w[o0:o0 + n])
    t.update_dual_bulk(pis, None, iters, -1.0)
    t.calc_delta_block(0, D, 0, N)
    t.basis_append_bulk(iters, np.arange(D, dtype=np.int32))
    if two_sigma:                                                     # a second sigma (and basis) per lambda: same rows read twice
        for d in range(D):
            si, new = t.calc_sigma(pis[d], 1.0 + d, d, False, int(iters[d]), 1e-3)
            t.basis_append(int(iters[d]), True, [si])
    setup = time.perf_counter() - t0
    t.set_timing(True)
    x = rng.uniform(0, 1, prob.prevCols + 1); x[0] = 0
    sweep, cut = [], []
    for s in range(reps + 2):
        c = t.sd_cut(x, k, 1, 0.0, want_istar=False)
        assert c is not None
        st = t.stats()
        if s >= 2:
            sweep.append(st["last_sweep_ms"]); cut.append(st["last_cut_ms"])
    st = t.stats()
    nb = t.counts()["basis"]
    peak, _ = bench.measured_peak()
    sm, cm = float(np.median(sweep)), float(np.median(cut))
    out = {"duals": D, "bases": nb, "observations": N, "R": R, "Q": Q, "n1": n1, "delta_GiB": round(8 * (1 + Q) * D * N / 2**30, 2),
           "variant": {1: "ldg", 2: "tma", 3: "general", 4: "tma_gen", 5: "recompute", 6: "tma_grouped"}[st["last_sweep_variant"]], "sweep_ms": round(sm, 4), "cut_ms": round(cm, 4),
           "pairs_per_s": round(nb * N / (cm * 1e-3), 0), "sweep_alg_GBps": round(st["last_sweep_bytes"] / (sm * 1e-3) / 1e9, 1),
           "frac_of_measured_peak": round(st["last_sweep_bytes"] / (sm * 1e-3) / 1e9 / peak, 3), "setup_s": round(setup, 2)}
    if out["variant"] == "recompute":
        out["note"] = ("FP64-pipe bound: delta.pib is recomputed from (lambda, omega), the table is not read; sweep_alg_GBps is the rate "
                       "an HBM-bound sweep would need for the same time")
    t.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-gib", type=float, default=140.0)
    args = ap.parse_args()
    shapes = {3: 4, 40: 63, 86: 89, 256: 121}                        # R -> n1 (pgp2, 20term, ssn, storm-like widths)
    grid = []
    for D in (1024, 4096, 16384, 65536):
        for N in (16384, 131072, 1048576):
            grid.append((D, N, 86, 0, 89, False))
    grid += [(16384, 131072, 3, 0, 4, False), (16384, 131072, 40, 0, 63, False), (16384, 131072, 256, 0, 121, False),
             (4096, 131072, 40, 8, 63, False), (16384, 131072, 40, 8, 63, False), (4096, 131072, 86, 0, 89, True)]
    for D, N, R, Q, n1, two in grid:
        gib = 8 * (1 + Q) * D * N / 2**30
        if gib > args.max_gib:
            print(json.dumps({"duals": D, "observations": N, "Q": Q, "skipped": f"delta table {gib:.0f} GiB exceeds one GPU"}), flush=True)
            continue
        print(json.dumps(run(D, N, R, Q, n1, two)), flush=True)
