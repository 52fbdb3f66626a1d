"""Pins the restated oracle (oracle/sd_oracle.c) to the reference's own code: the same synthetic SD traces
are replayed through oracle/_ref/libsdref.so (stocUpdate.c / cuts.c / optimal.c compiled from
/root/reference against the shim) and through the port; tables, flags, iStar and cut coefficients must be
BIT-identical (both are sequential CPU code with the same operation order)."""
import numpy as np
import pytest

import oracle_loader
from replay import assert_records_match, assert_tables_identical, replay
from stochasticdecomposition_b200._abi import Caps
from stochasticdecomposition_b200.synthetic import make_problem, make_trace, problem_for

pytestmark = pytest.mark.skipif(not oracle_loader.have_reference(), reason="reference build unavailable")

CASES = {
    # name: (problem kwargs, K, dual_pool, obs_pool, phi_len, replay kwargs)
    "pgp2_dedup": (dict(rows=7, cols=16, n1=4, n1c=4, R=3, Rb=3), 60, 6, 9, 0, {}),
    "rhs_only_fresh": (dict(rows=30, cols=40, n1=12, n1c=9, R=11, Rb=8), 40, 0, 0, 0, {}),
    "rhs_subset_rows": (dict(rows=25, cols=40, n1=10, n1c=10, R=12, Rb=5), 35, 10, 0, 0, dict(lb=-3.5)),
    "T_random": (dict(rows=30, cols=40, n1=12, n1c=9, R=11, Rb=8, Q=5), 40, 12, 20, 0, {}),
    "T_shared_cols_rvCols": (dict(rows=20, cols=30, n1=8, n1c=6, R=9, Rb=6, Q=6, shared_T_cols=True, distinct_rvCols=True), 30, 8, 0, 0, {}),
    "no_pi_eval": (dict(rows=20, cols=30, n1=8, n1c=6, R=9, Rb=6), 30, 8, 10, 0, dict(dual_stability=0)),
    "pi_cycle3_start10": (dict(rows=20, cols=30, n1=8, n1c=6, R=9, Rb=6), 40, 8, 10, 0, dict(pi_eval_start=10, pi_cycle=3)),
    "random_cost": (dict(rows=24, cols=30, n1=9, n1c=7, R=10, Rb=7, Q=0, rvd=3), 30, 0, 0, 2, {}),
    "random_cost_T": (dict(rows=24, cols=30, n1=9, n1c=7, R=10, Rb=7, Q=3, rvd=3), 30, 8, 12, 1, {}),
    "random_cost_pool_ties": (dict(rows=24, cols=30, n1=9, n1c=7, R=10, Rb=7, Q=2, rvd=3), 40, 5, 0, 1, dict(feas_density=0.6)),
    "random_cost_sparse_mask_null_cuts": (dict(rows=24, cols=30, n1=9, n1c=7, R=10, Rb=7, rvd=3), 30, 0, 0, 2, dict(feas_density=0.25)),
    "infeasible_bases": (dict(rows=20, cols=30, n1=8, n1c=6, R=9, Rb=6, rvd=2), 30, 0, 0, 1, dict(infeasible_every=7)),
}


def roomy_caps(K, phi_len=0):
    # the reference's own sizing (setup.c:136-139) is exceeded by two solves per iteration with phi columns and
    # it never bounds-checks; tests size the tables for the worst case instead
    n = 2 * K * (1 + phi_len) + 2
    return Caps(n, n, 2 * K + 2, K + 1, 1 + phi_len)


@pytest.mark.parametrize("name", sorted(CASES))
def test_trace_bit_identical(name):
    pk, K, dpool, opool, phi_len, rk = CASES[name]
    prob = make_problem(1000 + len(name), **pk)
    trace = make_trace(prob, K, seed=77 + K, dual_pool=dpool, obs_pool=opool, phi_len=phi_len)
    caps = roomy_caps(K, phi_len)
    ref = replay(oracle_loader.reference(), prob, trace, caps, **rk)
    port = replay(oracle_loader.oracle(), prob, trace, caps, **rk)
    assert any(c is not None for c in ref.cuts)
    assert_records_match(ref, port, exact_cut=True, ratio_only=True)
    assert_tables_identical(ref.tables, port.tables)


def test_shapes_of_named_problems_smoke():
    for name in ("pgp2", "20term_T", "ssn"):
        prob = problem_for(name)
        K = 12
        trace = make_trace(prob, K, seed=5, dual_pool=5, obs_pool=6)
        caps = roomy_caps(K)
        ref = replay(oracle_loader.reference(), prob, trace, caps)
        port = replay(oracle_loader.oracle(), prob, trace, caps)
        assert_records_match(ref, port, exact_cut=True, ratio_only=True)


def test_compute_istar_single_observation():
    prob = make_problem(3, rows=20, cols=30, n1=8, n1c=6, R=9, Rb=6, Q=2)
    K = 25
    trace = make_trace(prob, K, seed=9, dual_pool=7)
    caps = roomy_caps(K)
    ref = replay(oracle_loader.reference(), prob, trace, caps)
    port = replay(oracle_loader.oracle(), prob, trace, caps)
    x = trace.xs[3, 0]
    for obs in range(ref.counts["omega"]):
        for pi_eval in (0, 1):
            for is_new in (0, 1):
                a = ref.tables.compute_istar(x, obs, K, pi_eval, is_new)
                b = port.tables.compute_istar(x, obs, K, pi_eval, is_new)
                assert a[0] == b[0] and np.float64(a[1]).tobytes() == np.float64(b[1]).tobytes()


def test_dual_stability_tail_and_variance():
    """cuts.c:171-182 + calcVariance: run the reference SDCut with its real config gates over a trace and
    compare pi_ratio[] / dualStableFlag with sd_cut + dual_stability of the port."""
    import ctypes as C
    from stochasticdecomposition_b200._abi import CCut, _pf64, _pi32
    prob = make_problem(11, rows=12, cols=20, n1=5, n1c=5, R=4, Rb=4)
    K, scan = 40, 8
    trace = make_trace(prob, K, seed=3, dual_pool=4, obs_pool=5)
    caps = roomy_caps(K)
    ref_api, port_api = oracle_loader.reference(), oracle_loader.oracle()
    tr, tp = ref_api.create(prob, caps), port_api.create(prob, caps)
    ratio_r, ratio_p = np.zeros(scan), np.zeros(scan)
    for it in range(K):
        k = it + 1
        for t in (tr, tp):
            oi, onew = t.calc_omega(trace.observ[it], 1e-3)
            t.stochastic_updates(oi, onew, trace.duals[it, 0], trace.mubBar[it, 0], k, 1e-3)
        x = np.ascontiguousarray(trace.xs[it, 0])
        beta = np.zeros(prob.prevCols + 1)
        cut = CCut(0.0, _pf64(beta), None, 0, 0, 0.0, 0.0)
        flag = C.c_int(0)
        st = ref_api._fn("sd_cut_cfg")(tr.ctx, _pf64(x), k, 1, 2, 1, scan, 0.0, C.byref(cut), _pf64(ratio_r), C.byref(flag))
        assert st == 0
        pc = tp.sd_cut(x, k, k > 2, 0.0)
        stable = None
        if k > 2:
            stable = port_api._fn("dual_stability")(pc.cummOld, pc.cummAll, k, 2, scan, _pf64(ratio_p))
            assert stable == flag.value
        assert np.array_equal(ratio_r.view(np.int64), ratio_p.view(np.int64))
    v_r = ref_api._fn("calc_variance")(_pf64(ratio_r), scan)
    v_p = port_api._fn("calc_variance")(_pf64(ratio_p), scan)
    assert v_r == v_p


def test_cut_heights_and_reform():
    prob = make_problem(21, rows=20, cols=30, n1=8, n1c=6, R=9, Rb=6, Q=3)
    K = 30
    trace = make_trace(prob, K, seed=4, dual_pool=9, obs_pool=12)
    caps = roomy_caps(K)
    ref = replay(oracle_loader.reference(), prob, trace, caps, lb=-2.0)
    port = replay(oracle_loader.oracle(), prob, trace, caps, lb=-2.0)
    cuts = [c for c in ref.cuts if c is not None][-6:]
    alpha = np.array([c.alpha for c in cuts]); beta = np.stack([c.beta for c in cuts])
    ns = np.array([c.numSamples for c in cuts], np.int32); ai = alpha - 1.0
    xk = trace.xs[-1, 0]
    a = ref.tables.cut_heights(alpha, beta, ns, ai, K, xk, -2.0)
    b = port.tables.cut_heights(alpha, beta, ns, ai, K, xk, -2.0)
    assert a[0] == b[0] >= 0
    for u, v in zip(a[1:], b[1:]):
        assert np.array_equal(u.view(np.int64), v.view(np.int64))
    rng = np.random.default_rng(8)
    last = cuts[-1]
    for lbtype, lbv in ((0, 0.0), (1, -7.9)):
        observ = rng.integers(0, ref.counts["omega"] + 3, size=K).astype(np.int32)   # some beyond omegaCnt (optimal.c:205)
        ar, br = ref.tables.reform_cut(last.iStar, observ, K, lbtype, lbv)
        ap, bp = port.tables.reform_cut(last.iStar, observ, K, lbtype, lbv)
        assert np.float64(ar).tobytes() == np.float64(ap).tobytes()
        assert np.array_equal(br.view(np.int64), bp.view(np.int64))


def test_bootstrap_batch_reform():
    prob = make_problem(23, rows=20, cols=30, n1=8, n1c=6, R=9, Rb=6, Q=2)
    K = 40
    trace = make_trace(prob, K, seed=14, dual_pool=9, obs_pool=0)
    caps = roomy_caps(K)
    ref = replay(oracle_loader.reference(), prob, trace, caps)
    port = replay(oracle_loader.oracle(), prob, trace, caps)
    cuts = [c for c in ref.cuts if c is not None][-5:]
    rng = np.random.default_rng(3)
    observ = rng.integers(0, ref.counts["omega"], size=(7, K)).astype(np.int32)
    ar, br = ref.tables.reform_cuts_batch([c.iStar for c in cuts], observ, 1, -3.0)
    ap, bp = port.tables.reform_cuts_batch([c.iStar for c in cuts], observ, 1, -3.0)
    assert np.array_equal(ar.view(np.int64), ap.view(np.int64)) and np.array_equal(br.view(np.int64), bp.view(np.int64))


def test_check_basis_feasibility_matches_reference():
    """randCost.c:202-258 compiled from the reference vs the port, then cuts formed under the resulting mask"""
    import feas_scenario
    feas_scenario.compare(feas_scenario.run(oracle_loader.reference()), feas_scenario.run(oracle_loader.oracle()), exact=True)


def test_feasibility_cut_pool_matches_reference():
    """the reference's own updtFeasCutPool + addCut2Pool (cuts.c:465-517,643-655) against raw cuts from the port + the host dedup"""
    import feas_pool
    sr, ar, br = feas_pool.run(oracle_loader.reference(), use_reference_pool=True)
    sp, ap, bp = feas_pool.run(oracle_loader.oracle())
    assert sr == sp and sr[-1] > 5
    assert np.array_equal(ar.view(np.int64), ap.view(np.int64)) and np.array_equal(br.view(np.int64), bp.view(np.int64))


def test_feasibility_cut_pool_behind_the_abi_matches_reference():
    """sdgpu_feas_pool_update / _check (the port's restatement) against the reference's own updtFeasCutPool + addCut2Pool
    (cuts.c:465-517,643-655) and checkFeasCutPool (cuts.c:521-567; the cuts it hands to addCut2Master are recorded)"""
    import feas_pool
    r = feas_pool.run_device_pool(oracle_loader.reference())
    p = feas_pool.run_device_pool(oracle_loader.oracle())
    assert r[0] == p[0] and r[0][-1] > 100 and r[4] == p[4]
    assert np.array_equal(r[1].view(np.int64), p[1].view(np.int64)) and np.array_equal(r[2].view(np.int64), p[2].view(np.int64))
    seen = set()
    for (ar, ir), (ap, ip) in zip(r[3], p[3]):
        assert ir == ip
        assert np.array_equal(ar, np.isin(ap, [1, 3]).astype(np.int32))       # the reference only shows which cuts it adds to the master
        seen |= set(ap.tolist())
    assert seen == {0, 1, 2, 3}
    s2, a2, b2 = feas_pool.run(oracle_loader.oracle(), K=60)                   # the host-side pool of the older API gives the same pool
    assert s2 == p[0] and np.array_equal(a2, p[1]) and np.array_equal(b2, p[2])


def test_omp_flavour_same_istar():
    prob = make_problem(31, rows=30, cols=40, n1=12, n1c=9, R=11, Rb=8, Q=2)
    K = 40
    trace = make_trace(prob, K, seed=6, dual_pool=12)
    caps = roomy_caps(K)
    seq = replay(oracle_loader.oracle(), prob, trace, caps)
    omp = replay(oracle_loader.oracle(), prob, trace, caps, sd_cut_variant="sd_cut_omp")
    assert_records_match(seq, omp, exact_cut=False)
