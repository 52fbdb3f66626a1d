#!/usr/bin/env python
"""Run under torchrun (one rank per GPU): a synthetic SD trace through ShardedTables on the CUDA library with the
in-library NCCL all-reduce, compared on rank 0 with the single-process CPU oracle (iStar after the gather
bit-exact, cut coefficients within 1e-9).  Also exercises the torch.distributed all-reduce path on the device
buffer.  Prints MULTIGPU_OK on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_loader  # noqa: E402
import stochasticdecomposition_b200 as sd  # noqa: E402
from replay import pi_eval_flag, replay  # noqa: E402
from stochasticdecomposition_b200._abi import Caps  # noqa: E402
from stochasticdecomposition_b200.sharding import ShardedTables  # noqa: E402
from stochasticdecomposition_b200.synthetic import make_problem, make_trace  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    prob = make_problem(5, rows=40, cols=60, n1=14, n1c=11, R=17, Rb=13, Q=2)
    K = 60
    trace = make_trace(prob, K, seed=11, dual_pool=20, obs_pool=0)
    n = 2 * K + 2
    for mode in ("library_nccl", "torch", "peer", "peer_then_nccl", "both_pick_nccl"):
        sh = ShardedTables(sd.load_library().create(prob, Caps(n, n, n, K + 1, 1), local), rank, world)
        if mode == "library_nccl":
            sh.attach_library_nccl()
        if mode == "peer":
            sh.attach_peer_exchange()
        if mode == "peer_then_nccl":                # attaching NCCL afterwards must leave the peer exchange alone (automatic choice: peer)
            sh.attach_peer_exchange()
            sh.attach_library_nccl()
        if mode == "both_pick_nccl":                # both attached, NCCL selected explicitly
            sh.attach_library_nccl()
            sh.attach_peer_exchange()
            sh.t.set_collective(1)
        cuts = []
        for it in range(K):
            k = it + 1
            oi, onew = sh.calc_omega(trace.observ[it], 1e-3)
            for sv in ((0, 1) if trace.two_solves[it] else (0,)):
                sh.stochastic_updates(oi, onew, trace.duals[it, sv], trace.mubBar[it, sv], k, 1e-3)
                onew = False
                cut = sh.sd_cut(trace.xs[it, sv], k, pi_eval_flag(k), 0.0)
                cuts.append((cut.alpha, cut.beta.copy(), sh.gather_istar(cut.iStar), cut.cummOld, cut.cummAll))
        if rank == 0:
            single = replay(oracle_loader.oracle(), prob, trace, Caps(n, n, n, K + 1, 1))
            assert len(cuts) == len(single.cuts)
            for (alpha, beta, istar, cold, call), ref in zip(cuts, single.cuts):
                assert np.array_equal(istar, ref.iStar), mode
                scale = max(abs(ref.alpha), np.abs(ref.beta[1:]).max())
                assert abs(alpha - ref.alpha) <= 1e-9 * abs(ref.alpha), mode
                assert np.abs(beta - ref.beta).max() <= 1e-9 * scale, mode
                assert abs(call - ref.cummAll) <= 1e-9 * max(abs(ref.cummAll), 1e-300), mode
        sh.t.close()
        dist.barrier()
    if rank == 0:
        print(f"MULTIGPU_OK world={world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
