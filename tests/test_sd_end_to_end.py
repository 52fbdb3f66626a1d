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


def _lockstep(apis, shape, K, rtol, seed=3):
    slp = make_slp(shape)
    prob = slp.problem()
    tabs = Lockstep([a.create(prob, caps_for(K)) for a in apis], rtol=rtol)
    host = SDHost(slp, tabs, seed=seed)
    st = host.run(K)
    assert tabs.checked >= K
    return st, host


@pytest.mark.skipif(not oracle_loader.have_reference(), reason="reference build unavailable")
@pytest.mark.parametrize("shape,K", [("pgp2", 150), ("20term", 40), ("20term_T", 40)])
def test_reference_and_port_agree_over_a_whole_run(shape, K):
    _lockstep([oracle_loader.reference(), oracle_loader.oracle()], shape, K, rtol=0.0)


def test_sd_converges_on_pgp2_shape():
    """sanity of the harness itself: candidate and incumbent estimates approach each other"""
    st, host = _lockstep([oracle_loader.oracle()], "pgp2", 300, rtol=0.0)
    tail = st.history[-20:]
    gap = np.mean([abs(c - i) / max(abs(i), 1e-9) for _, c, i, _, _ in tail])
    assert gap < 0.05, gap
    assert len(host.cuts) <= host.maxCuts


@pytest.mark.gpu
@pytest.mark.parametrize("shape,K", [("pgp2", 300), ("20term", 120), ("20term_T", 120), ("ssn", 120)])
def test_cuda_and_port_agree_over_a_whole_run(shape, K):
    import stochasticdecomposition_b200 as sd
    _lockstep([sd.load_library(), oracle_loader.oracle()], shape, K, rtol=1e-9)


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
