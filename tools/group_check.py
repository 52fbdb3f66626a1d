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
    total = 0
    # three replications through sdgpu_group_reset (setup.c:242-246): the second one forms only two cuts, so that a restarted
    # exchange sequence would meet the first replication's stale flags (the sequence number is monotonic for that reason)
    for rep, (Kr, seed) in enumerate(((K, 11), (2, 12), (40, 13))):
        tr = trace if rep == 0 else make_trace(prob, Kr, seed=seed, dual_pool=20, obs_pool=30)
        if rep > 0:
            grp.reset()
        cuts = []
        for it in range(Kr):
            k = it + 1
            oi, onew = grp.calc_omega(tr.observ[it], 1e-3)
            for sv in ((0, 1) if tr.two_solves[it] else (0,)):
                grp.stochastic_updates(oi, onew, tr.duals[it, sv], tr.mubBar[it, sv], k, 1e-3)
                onew = False
                cuts.append(grp.sd_cut(tr.xs[it, sv], k, pi_eval_flag(k), 0.0))
        single = replay(oracle_loader.oracle(), prob, tr, Caps(n, n, n, K + 1, 1))
        assert grp.counts() == single.counts, (grp.counts(), single.counts)
        assert len(cuts) == len(single.cuts)
        for c, ref in zip(cuts, single.cuts):
            assert np.array_equal(c.iStar, ref.iStar), rep
            scale = max(abs(ref.alpha), np.abs(ref.beta[1:]).max())
            assert abs(c.alpha - ref.alpha) <= 1e-9 * abs(ref.alpha) and np.abs(c.beta - ref.beta).max() <= 1e-9 * scale, rep
            assert abs(c.cummAll - ref.cummAll) <= 1e-9 * max(abs(ref.cummAll), 1e-300), rep
        total += len(cuts)
    grp.close()
    print(f"GROUP_OK devices={G} cuts={total} replications=3", flush=True)


if __name__ == "__main__":
    main()
