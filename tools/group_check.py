#!/usr/bin/env python
"""ONE process, one host thread, all visible GPUs: a synthetic SD trace through the sdgpu_group_* API (observations sharded
round-robin, cut all-reduced through NVLink peer memory inside the cut kernel), compared with the single-table CPU oracle.
Prints GROUP_OK on success."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_loader  # noqa: E402
import stochasticdecomposition_b200 as sd  # noqa: E402
from replay import pi_eval_flag, replay  # noqa: E402
from stochasticdecomposition_b200._abi import Caps, Group  # noqa: E402
from stochasticdecomposition_b200.synthetic import make_problem, make_trace  # noqa: E402


def main():
    G = torch.cuda.device_count()
    prob = make_problem(5, rows=40, cols=60, n1=14, n1c=11, R=17, Rb=13, Q=2)
    K = 70
    trace = make_trace(prob, K, seed=11, dual_pool=20, obs_pool=30)
    n = 2 * K + 2
    grp = Group(sd.load_library(), prob, Caps(n, n, n, K + 1, 1), list(range(G)))
    cuts = []
    for it in range(K):
        k = it + 1
        oi, onew = grp.calc_omega(trace.observ[it], 1e-3)
        for sv in ((0, 1) if trace.two_solves[it] else (0,)):
            grp.stochastic_updates(oi, onew, trace.duals[it, sv], trace.mubBar[it, sv], k, 1e-3)
            onew = False
            cuts.append(grp.sd_cut(trace.xs[it, sv], k, pi_eval_flag(k), 0.0))
    single = replay(oracle_loader.oracle(), prob, trace, Caps(n, n, n, K + 1, 1))
    assert grp.counts() == single.counts, (grp.counts(), single.counts)
    assert len(cuts) == len(single.cuts)
    for c, ref in zip(cuts, single.cuts):
        assert np.array_equal(c.iStar, ref.iStar)
        scale = max(abs(ref.alpha), np.abs(ref.beta[1:]).max())
        assert abs(c.alpha - ref.alpha) <= 1e-9 * abs(ref.alpha) and np.abs(c.beta - ref.beta).max() <= 1e-9 * scale
        assert abs(c.cummAll - ref.cummAll) <= 1e-9 * max(abs(ref.cummAll), 1e-300)
    grp.close()
    print(f"GROUP_OK devices={G} cuts={len(cuts)}", flush=True)


if __name__ == "__main__":
    main()
