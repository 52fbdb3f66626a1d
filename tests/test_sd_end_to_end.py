"""End-to-end SD runs (tools/sd_highs_host.py: the reference's host loop restated around HiGHS, synthetic instances with
the reference problems' shapes) with two table backends driven in lock step on identical inputs.

  CPU : the reference build (oracle/_ref) and the port must agree bit for bit at every iteration of a whole SD run;
  GPU : the CUDA library and the port must agree -- every index and iStar exactly, every cut within 1e-9 relative --
        and independent runs (one backend each) must end at the same incumbent.
LP solves are HiGHS, not CPLEX; the master is solved exactly through its dual (see the harness)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

import oracle_loader  # noqa: E402
from sd_highs_host import Lockstep, SDHost, caps_for, make_slp  # noqa: E402


def _caps(slp, K):
    from stochasticdecomposition_b200._abi import Caps
    if not slp.rvd:
        return caps_for(K)
    n = (1 + slp.rvd) * 2 * K + 8                         # every solve may add 1 + phiLength lambda / sigma rows
    return Caps(n, n, 2 * K + 2, K + 1, 1 + slp.rvd)


def _lockstep(apis, shape, K, rtol, seed=3):
    slp = make_slp(shape)
    prob = slp.problem()
    tabs = Lockstep([a.create(prob, _caps(slp, K)) for a in apis], rtol=rtol)
    host = SDHost(slp, tabs, seed=seed)
    st = host.run(K)
    assert tabs.checked >= K
    return st, host


@pytest.mark.skipif(not oracle_loader.have_reference(), reason="reference build unavailable")
@pytest.mark.parametrize("shape,K", [("pgp2", 150), ("20term", 40), ("20term_T", 40), ("randcost_small", 70)])
def test_reference_and_port_agree_over_a_whole_run(shape, K):
    _lockstep([oracle_loader.reference(), oracle_loader.oracle()], shape, K, rtol=0.0)


@pytest.mark.skipif(not oracle_loader.have_reference(), reason="reference build unavailable")
@pytest.mark.parametrize("shape,K", [("ssn", 160), ("storm_rc", 40)])
def test_reference_and_port_agree_through_reset(shape, K):
    """the shapes of the long GPU lock-step runs (tests/test_sd_long.py), here reference build vs port, bit for bit, over two
    replications separated by cleanCellType (setup.c:242-246)"""
    slp = make_slp(shape)
    prob = slp.problem()
    tabs = Lockstep([a.create(prob, _caps(slp, K)) for a in (oracle_loader.reference(), oracle_loader.oracle())], rtol=0.0)
    for rep in range(2):
        if rep:
            tabs.reset()
            assert tabs.counts() == {"omega": 0, "lambda": 0, "sigma": 0, "basis": 0}
        SDHost(slp, tabs, seed=3 + 7 * rep).run(K)
    assert tabs.checked >= 2 * K


def test_sd_converges_on_pgp2_shape():
    """sanity of the harness itself: candidate and incumbent estimates approach each other"""
    st, host = _lockstep([oracle_loader.oracle()], "pgp2", 300, rtol=0.0)
    tail = st.history[-20:]
    gap = np.mean([abs(c - i) / max(abs(i), 1e-9) for _, c, i, _, _ in tail])
    assert gap < 0.05, gap
    assert len(host.cuts) <= host.maxCuts


@pytest.mark.gpu
@pytest.mark.parametrize("shape,K", [("pgp2", 300), ("20term", 120), ("20term_T", 120), ("ssn", 120), ("randcost_small", 150)])
def test_cuda_and_port_agree_over_a_whole_run(shape, K):
    import stochasticdecomposition_b200 as sd
    _lockstep([sd.load_library(), oracle_loader.oracle()], shape, K, rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,K,variant", [("ssn", 80, 2), ("20term_T", 80, 2), ("randcost_small", 100, 2), ("20term", 80, 4), ("pgp2", 150, 1)])
def test_cuda_sweep_families_agree_over_a_whole_run(shape, K, variant):
    """The same lock-step run with one sweep family forced: the TMA rings (plain, random T, term-linear random cost), the grouped
    ring, the load-based kernels where the default would recompute."""
    import stochasticdecomposition_b200 as sd
    from replay import ForcedVariant
    _lockstep([ForcedVariant(sd.load_library(), variant), oracle_loader.oracle()], shape, K, rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,K", [("pgp2", 300), ("20term_T", 100), ("ssn", 100)])
def test_independent_runs_reach_the_same_incumbent(shape, K):
    import stochasticdecomposition_b200 as sd
    out = []
    for api in (sd.load_library(), oracle_loader.oracle()):
        slp = make_slp(shape)
        host = SDHost(slp, api.create(slp.problem(), caps_for(K)), seed=5)
        out.append(host.run(K))
    a, b = out
    assert abs(a.incumb_est - b.incumb_est) <= 1e-9 * max(abs(b.incumb_est), 1e-300), (a.incumb_est, b.incumb_est)
    assert np.abs(a.incumbX - b.incumbX).max() <= 1e-9 * max(np.abs(b.incumbX).max(), 1e-300)
    assert a.lp_solves == b.lp_solves and a.incumbent_changes == b.incumbent_changes


@pytest.mark.parametrize("shape", ["pgp2", "20term_T", "randcost_small"])
def test_stoch_check_invariants_against_the_lp(shape):
    """The reference's STOCH_CHECK blocks (cuts.c:64-76, subprob.c:75-80), as assertions: with the tables an SD run has
    built, the argmax value of ANY stored observation at x is a lower bound on the true recourse value h(x, w) (LP solved
    here), it is attained (equality) for the observation / x pair whose dual vertex was just stored, and the cut height at x
    is the weighted mean of those argmax values.  This checks the path against the LP itself, not against another
    implementation of the path."""
    K = 80
    slp = make_slp(shape)
    t = oracle_loader.oracle().create(slp.problem(), _caps(slp, K + 1))
    host = SDHost(slp, t, seed=11)
    host.run(K)
    x1 = host.candidX
    # a fresh solve at (x, last observation), stored through the normal update path -> its estimate must equal the LP objective
    last = len(host.obs_store) - 1
    obj, pi = host.sub.solve(x1[1:], host.obs_store[last][1:])
    if slp.rvd:
        host.random_cost_updates(last, False, pi)
        assert any(len(b) for b in [host.basis_keys]) and t.counts()["sigma"] > t.counts()["basis"]      # phi columns were stored
    else:
        t.stochastic_updates(last, False, pi, 0.0, host.k, 1e-3)
    istar, val = t.compute_istar(x1, last, host.k, 0, 0)
    assert istar >= 0 and abs(val - obj) <= 1e-7 * max(1.0, abs(obj)), (val, obj)
    # every stored observation: argmax value <= true recourse value
    vals, ws = [], []
    for o in range(len(host.obs_store)):
        true_obj, _ = host.sub.solve(x1[1:], host.obs_store[o][1:])
        _, v = t.compute_istar(x1, o, host.k, 0, 0)
        assert v <= true_obj + 1e-6 * max(1.0, abs(true_obj)), (o, v, true_obj)
        vals.append(v); ws.append(t.get_omega(o)[1])
    # the cut at x, evaluated at x, is the weight-averaged argmax value (cuts.c:142-168,184-188)
    cut = t.sd_cut(x1, host.k, 0, 0.0)
    height = cut.alpha - float(np.dot(cut.beta[1:], x1[1:]))
    mean = float(np.dot(vals, ws)) / host.k
    assert abs(height - mean) <= 1e-9 * max(1.0, abs(mean)), (height, mean)
