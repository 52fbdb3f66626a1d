"""world_size-2 (and 3) CPU tests of the multi-GPU host logic over gloo: observations dealt round-robin, dual-side
tables replicated, one all-reduce of n1+4 doubles per cut.  The CPU checker stands in for the device tables, so
what is under test is sharding.py: global first-match dedup of observations, owner-only delta columns, replicated
find-or-append, the partial/all-reduce/finish split and the iStar gather.  The sharded run must reproduce the
single-process run: every index identical, iStar identical after the gather, cut coefficients within 1e-9."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, K, q):
    sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_loader
    from replay import pi_eval_flag
    from stochasticdecomposition_b200._abi import Caps
    from stochasticdecomposition_b200.sharding import ShardedTables
    from stochasticdecomposition_b200.synthetic import make_problem, make_trace
    prob = make_problem(5, rows=20, cols=30, n1=8, n1c=6, R=9, Rb=6, Q=2)
    trace = make_trace(prob, K, seed=11, dual_pool=9, obs_pool=14)
    n = 2 * K + 2
    sh = ShardedTables(oracle_loader.oracle().create(prob, Caps(n, n, n, K + 1, 1)), rank, world)
    rec = {"omega": [], "basis": [], "cuts": []}
    for it in range(K):
        k = it + 1
        oi, onew = sh.calc_omega(trace.observ[it], 1e-3)
        rec["omega"].append((oi, onew))
        for sv in ((0, 1) if trace.two_solves[it] else (0,)):
            bi, bnew = sh.stochastic_updates(oi, onew, trace.duals[it, sv], trace.mubBar[it, sv], k, 1e-3)
            onew = False
            rec["basis"].append((bi, bnew))
            cut = sh.sd_cut(trace.xs[it, sv], k, pi_eval_flag(k), 0.0)
            full = sh.gather_istar(cut.iStar)
            rec["cuts"].append((cut.alpha, cut.beta.copy(), full, cut.cummOld, cut.cummAll))
    if rank == 0:
        q.put(rec)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_matches_single_process(world):
    sys.path.insert(0, HERE)
    import oracle_loader
    from replay import replay
    from stochasticdecomposition_b200._abi import Caps
    from stochasticdecomposition_b200.synthetic import make_problem, make_trace
    oracle_loader.oracle()                                    # build once before forking workers
    K = 30
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    rec = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prob = make_problem(5, rows=20, cols=30, n1=8, n1c=6, R=9, Rb=6, Q=2)
    trace = make_trace(prob, K, seed=11, dual_pool=9, obs_pool=14)
    n = 2 * K + 2
    single = replay(oracle_loader.oracle(), prob, trace, Caps(n, n, n, K + 1, 1))
    assert [o for o, _ in rec["omega"]] == single.omega_idx and [f for _, f in rec["omega"]] == single.omega_new
    assert [b for b, _ in rec["basis"]] == single.basis_idx and [f for _, f in rec["basis"]] == single.basis_new
    assert len(rec["cuts"]) == len(single.cuts)
    for (alpha, beta, istar, cold, call), ref in zip(rec["cuts"], single.cuts):
        assert np.array_equal(istar, ref.iStar)
        scale = max(abs(ref.alpha), np.abs(ref.beta[1:]).max())
        assert abs(alpha - ref.alpha) <= 1e-9 * abs(ref.alpha)
        assert np.abs(beta - ref.beta).max() <= 1e-9 * scale
        assert abs(cold - ref.cummOld) <= 1e-9 * max(abs(ref.cummOld), 1e-300)
        assert abs(call - ref.cummAll) <= 1e-9 * max(abs(ref.cummAll), 1e-300)


def test_index_maps():
    from stochasticdecomposition_b200.sharding import global_index, local_slot, owner_of, shard_counts
    for world in (1, 2, 4, 8):
        seen = set()
        for o in range(100):
            r, l = owner_of(o, world), local_slot(o, world)
            assert global_index(l, r, world) == o
            seen.add((r, l))
        assert len(seen) == 100
        assert sum(shard_counts(100, world)) == 100
        assert shard_counts(5, 4) == [2, 1, 1, 1]
