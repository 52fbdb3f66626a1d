#!/usr/bin/env python
"""Sweep bandwidth of the non-default code paths: random technology-matrix elements (Q > 0: delta carries 1+Q planes per
pair) and random-cost bases (multi-term scores + feasibility mask).  One JSON line per configuration."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import stochasticdecomposition_b200 as sd  # noqa: E402
from stochasticdecomposition_b200._abi import Caps  # noqa: E402
from stochasticdecomposition_b200.synthetic import make_problem  # noqa: E402


def run(D, N, Q, rvd=0, phi=0, reps=8, R=40, n1=63, variant=0):
    prob = make_problem(3, rows=max(R, 8) + 20, cols=200, n1=n1, n1c=n1, R=R, Rb=R, Q=Q, rvd=rvd)
    rng = np.random.default_rng(5)
    pis = rng.uniform(-1, 1, (D, prob.rows + 1)); pis[:, 0] = 0
    obs = rng.normal(0, 1, (N, prob.numRV + 1)); obs[:, 0] = 0
    w = (1 + rng.poisson(0.25, N)).astype(np.int32)
    k = int(w.sum())
    iters = np.ceil((np.arange(D) + 1) * (k / D)).astype(np.int32)
    t = sd.load_library().create(prob, Caps(D + 2, D + 2, D + 2, N + 2, 1 + rvd))
    t.omega_append_bulk(obs, w)
    t.update_dual_bulk(pis, None, iters, -1.0)
    t.calc_delta_block(0, D, 0, N)
    if phi == 0:
        t.basis_append_bulk(iters, np.arange(D, dtype=np.int32))
        nb, pairs_rows = D, D
    else:
        nb = D // (1 + phi)
        for b in range(nb):
            sig = [b * (1 + phi) + j for j in range(1 + phi)]
            t.basis_append(int(iters[sig[0]]), True, sig, [0] + list(range(1, phi + 1)))
        pairs_rows = nb * (1 + phi)
    t.set_timing(True)
    t.set_sweep_variant(variant)
    x = rng.uniform(0, 1, prob.prevCols + 1); x[0] = 0
    ms = []
    for _ in range(reps):
        c = t.sd_cut(x, k, 1, 0.0, want_istar=False)
        assert c is not None
        ms.append(t.stats()["last_sweep_ms"])
    st = t.stats()
    m = float(np.median(ms))
    byts = 8.0 * (1 + Q) * pairs_rows * N + (nb * N if rvd else 0)
    out = {"env": {k: v for k, v in os.environ.items() if k.startswith("SDGPU_")}, "D": D, "N": N, "Q": Q, "rvdOmCnt": rvd, "phiLength": phi, "bases": nb, "variant": st["last_sweep_variant"],
           "sweep_ms": round(m, 4), "alg_GBps": round(byts / (m * 1e-3) / 1e9, 1), "pairs_per_s": round(nb * N / (m * 1e-3), 0),
           "cut_ms": round(st["last_cut_ms"], 4)}
    t.close()
    return out


if __name__ == "__main__":
    for cfg in (dict(D=16384, N=131072, Q=0), dict(D=8192, N=131072, Q=2), dict(D=4096, N=131072, Q=8), dict(D=5000, N=5000, Q=8),
                dict(D=6000, N=5000, Q=0, rvd=4, phi=2), dict(D=6000, N=5000, Q=2, rvd=4, phi=2), dict(D=5000, N=5000, Q=0, rvd=4, phi=0),
                dict(D=6000, N=5000, Q=0, rvd=4, phi=2, variant=1), dict(D=5000, N=5000, Q=0, rvd=4, phi=0, variant=1),
                dict(D=6144, N=65536, Q=0, rvd=4, phi=2), dict(D=6144, N=65536, Q=0, rvd=4, phi=2, variant=1), dict(D=6144, N=65536, Q=2, rvd=8, phi=1),
                dict(D=6144, N=65536, Q=0, rvd=4, phi=0), dict(D=6144, N=65536, Q=0, rvd=4, phi=0, variant=1),
                dict(D=6144, N=16384, Q=0, rvd=4, phi=0), dict(D=6144, N=16384, Q=0, rvd=4, phi=0, variant=1)):
        if len(sys.argv) > 1 and sys.argv[1] == "rc" and not cfg.get("rvd"):
            continue
        print(json.dumps(run(**cfg)), flush=True)
