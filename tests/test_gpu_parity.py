"""GPU parity (run on the B200 box with -m gpu): the CUDA library against the CPU oracle through the same C
ABI, on the same seeded traces.  Bar (BASELINE.json north_star): every index bit-exact (observation, lambda,
sigma, basis, iStar with the reference's tie-break), tables bit-identical (same operation order, no FMA),
cut coefficients / cummOld / cummAll within 1e-9 relative (the only place the summation order differs)."""
import numpy as np
import pytest

import oracle_loader
import stochasticdecomposition_b200 as sd
from replay import assert_records_match, assert_tables_identical, replay
from stochasticdecomposition_b200._abi import Caps
from stochasticdecomposition_b200.synthetic import make_problem, make_trace, problem_for
from test_oracle_vs_ref import CASES, roomy_caps

pytestmark = pytest.mark.gpu
RTOL = 1e-9   # BASELINE.json: "cut coefficients ... within 1e-9 relative"


@pytest.mark.parametrize("variant", [0, 2])
@pytest.mark.parametrize("name", sorted(CASES))
def test_trace_parity(name, variant):
    """variant 0: the size-based default (these small traces take the LDG / per-term-gather kernels); variant 2: the TMA rings
    forced (k_sweep_tma, k_sweep_tma_q, and for the random-cost cases the term-linear k_sweep_tma_gen)."""
    pk, K, dpool, opool, phi_len, rk = CASES[name]
    prob = make_problem(1000 + len(name), **pk)
    trace = make_trace(prob, K, seed=77 + K, dual_pool=dpool, obs_pool=opool, phi_len=phi_len)
    caps = roomy_caps(K, phi_len)
    port = replay(oracle_loader.oracle(), prob, trace, caps, **rk)
    gpu = replay(sd.load_library(), prob, trace, caps, sweep_variant=variant, **rk)
    assert_records_match(port, gpu, exact_cut=False, rtol=RTOL)
    assert_tables_identical(port.tables, gpu.tables)
    if variant == 2 and pk.get("rvd", 0) > 0:
        assert gpu.tables.stats()["last_sweep_variant"] == 4


@pytest.mark.parametrize("shape", ["pgp2", "20term_T", "ssn"])
def test_named_shapes(shape):
    prob = problem_for(shape)
    K = 40
    trace = make_trace(prob, K, seed=5, dual_pool=9, obs_pool=14)
    caps = roomy_caps(K)
    port = replay(oracle_loader.oracle(), prob, trace, caps)
    gpu = replay(sd.load_library(), prob, trace, caps)
    assert_records_match(port, gpu, exact_cut=False, rtol=RTOL)


def _bulk_tables(api, prob, pis, mub, iters, obs, weights, caps, tol=-1.0):
    t = api.create(prob, caps)
    t.omega_append_bulk(obs, weights)
    li, si = t.update_dual_bulk(pis, mub, iters, tol)
    c = t.counts()
    if tol < 0:
        t.calc_delta_block(0, c["lambda"], 0, c["omega"])
    for s in si if tol < 0 else sorted(set(si.tolist())):
        t.basis_append(int(iters[0]) if tol >= 0 else int(iters[s]), True, [int(s)])
    return t, li, si


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("D,N,Q", [(700, 1500, 0), (300, 1100, 3), (2100, 600, 0), (150, 700, 90)])
def test_bulk_multi_tile_multi_chunk(D, N, Q, variant):
    """Several observation tiles and several basis chunks; ties forced by duplicated duals (lowest index must
    win) and all-zero observations; both pi_eval modes; weights > 1.  The last shape has 90 random T elements: past what a ring stage
    holds, the load-based sweep with all the x entries in shared memory."""
    prob = make_problem(9, rows=40, cols=60, n1=14, n1c=11, R=17, Rb=13, Q=Q) if Q <= 14 else \
        make_problem(9, rows=120, cols=200, n1=100, n1c=60, R=40, Rb=30, Q=Q)
    rng = np.random.default_rng(D + N)
    pis = rng.uniform(-1, 1, (D, prob.rows + 1)) * (rng.random((D, prob.rows + 1)) > 0.3)
    dup = rng.choice(np.arange(1, D), size=max(1, D // 50), replace=False)
    for d in dup:
        pis[d] = pis[rng.integers(0, d)]                       # exact copy of an earlier dual -> equal scores
    mub = np.zeros(D)
    iters = np.ceil((np.arange(D) + 1) * (1.25 * N) / D).astype(np.int32)   # monotone ck, ~90/10 window split
    obs = rng.normal(0, 1, (N, prob.numRV + 1)); obs[:, 0] = 0
    obs[rng.choice(N, 16, replace=False)] = 0.0
    weights = (1 + rng.poisson(0.25, N)).astype(np.int32)
    k = int(weights.sum())
    caps = Caps(D + 2, D + 2, D + 2, N + 3, 1)
    to, lo, so = _bulk_tables(oracle_loader.oracle(), prob, pis, mub, iters, obs, weights, caps)
    tg, lg, sg = _bulk_tables(sd.load_library(), prob, pis, mub, iters, obs, weights, caps)
    assert np.array_equal(lo, lg) and np.array_equal(so, sg)
    tg.set_sweep_variant(variant)           # 1 = LDG streaming, 2 = TMA bulk ring (falls back to LDG when Q > 0)
    for plane in range(Q + 1):
        a = to.get_delta_block(0, D, 0, N, plane); b = tg.get_delta_block(0, D, 0, N, plane)
        assert np.array_equal(a.view(np.int64), b.view(np.int64)), f"delta plane {plane} differs"
    x = rng.uniform(0, 1, prob.prevCols + 1); x[0] = 0
    for pi_eval in (0, 1):
        co = to.sd_cut(x, k, pi_eval, 0.0); cg = tg.sd_cut(x, k, pi_eval, 0.0)
        assert co is not None and cg is not None
        assert np.array_equal(co.iStar, cg.iStar), np.nonzero(co.iStar != cg.iStar)[0][:10]
        assert len(set(co.iStar.tolist())) > 3
        scale = max(abs(co.alpha), np.abs(co.beta[1:]).max())
        assert abs(co.alpha - cg.alpha) <= RTOL * abs(co.alpha)
        assert np.abs(co.beta - cg.beta).max() <= RTOL * scale
        assert abs(co.cummOld - cg.cummOld) <= RTOL * max(abs(co.cummOld), 1e-300)
        assert abs(co.cummAll - cg.cummAll) <= RTOL * max(abs(co.cummAll), 1e-300)
    # the one-at-a-time appends after a bulk load agree with the bulk block (same cell arithmetic)
    newobs = rng.normal(0, 1, prob.numRV + 1)
    for t in (to, tg):
        oi, new = t.calc_omega(newobs, 1e-3)
        assert new and oi == N
        t.calc_delta(True, oi)
    a = to.get_delta_block(0, D, N, N + 1); b = tg.get_delta_block(0, D, N, N + 1)
    assert np.array_equal(a.view(np.int64), b.view(np.int64))


@pytest.mark.parametrize("D,N,extra", [(500, 1500, 2), (900, 700, 1), (300, 1100, 5), (2600, 600, 3)])
def test_grouped_sweep_shared_lambda_rows(D, N, extra, monkeypatch):
    """Several sigmas / bases per lambda row (calcSigma appends a sigma whenever pi x bBar or pi x Cbar differ while the lambda is
    already stored): the grouped sweep (variant 4) walks the bases sorted by (row, basis index) and keeps the running maximum
    lexicographically.  iStar equals the oracle's, and the cut is bit-identical to the plain kernels'.  Bases are interleaved
    (second sigmas appended after all first ones) so the sorted walk is far from basis order; duplicated duals force ties
    across groups.  The last shape has more than one descriptor batch (256 entries) per chunk."""
    if D == 2600:
        monkeypatch.setenv("SDGPU_CHUNKS", "3")          # ~3 500 entries per chunk: fourteen descriptor batches, a ragged last one
    prob = make_problem(17, rows=40, cols=60, n1=14, n1c=11, R=17, Rb=13, Q=0)
    rng = np.random.default_rng(D + N + extra)
    pis = rng.uniform(-1, 1, (D, prob.rows + 1)) * (rng.random((D, prob.rows + 1)) > 0.3)
    for d in rng.choice(np.arange(1, D), size=max(1, D // 50), replace=False):
        pis[d] = pis[rng.integers(0, d)]
    iters = np.ceil((np.arange(D) + 1) * (1.25 * N) / D).astype(np.int32)
    obs = rng.normal(0, 1, (N, prob.numRV + 1)); obs[:, 0] = 0
    obs[rng.choice(N, 16, replace=False)] = 0.0
    weights = (1 + rng.poisson(0.25, N)).astype(np.int32)
    k = int(weights.sum())
    nb = D * (1 + extra)
    caps = Caps(D + 2, nb + 2, nb + 2, N + 3, 1)
    tabs = []
    for api in (oracle_loader.oracle(), sd.load_library()):
        t, li, si = _bulk_tables(api, prob, pis, np.zeros(D), iters, obs, weights, caps)
        for e in range(extra):
            for d in range(D):                  # every call appends a sigma (mubBar differs), so basis index == sigma index stays true (cuts.c:161)
                s2, new = t.calc_sigma(pis[d], 1.0 + e + 0.25 * (d % 3), int(li[d]), False, int(iters[(d * 7 + e) % D]), 1e-3)
                assert new
                t.basis_append(int(iters[(d * 7 + e) % D]), True, [s2])
        tabs.append(t)
    to, tg = tabs
    assert to.counts() == tg.counts() and tg.counts()["basis"] > D
    x = rng.uniform(0, 1, prob.prevCols + 1); x[0] = 0
    for pi_eval in (0, 1):
        co = to.sd_cut(x, k, pi_eval, 0.0)
        tg.set_sweep_variant(1); c1 = tg.sd_cut(x, k, pi_eval, 0.0)
        tg.set_sweep_variant(4); c4 = tg.sd_cut(x, k, pi_eval, 0.0)
        assert tg.stats()["last_sweep_variant"] == 6
        assert co is not None and c1 is not None and c4 is not None
        assert np.array_equal(co.iStar, c4.iStar), np.nonzero(co.iStar != c4.iStar)[0][:10]
        assert np.array_equal(c1.iStar, c4.iStar)
        assert c1.alpha == c4.alpha and np.array_equal(c1.beta, c4.beta) and c1.cummOld == c4.cummOld and c1.cummAll == c4.cummAll
        assert (co.iStar >= D).any(), "no second-sigma basis ever won: the test does not exercise the groups"
    # a basis appended after the grouping was built is inserted in place
    s3, new = tg.calc_sigma(pis[3], 77.0, 3, False, 1, 1e-3); tg.basis_append(1, True, [s3])
    s3o, _ = to.calc_sigma(pis[3], 77.0, 3, False, 1, 1e-3); to.basis_append(1, True, [s3o])
    co = to.sd_cut(x, k, 1, 0.0); c4 = tg.sd_cut(x, k, 1, 0.0)
    assert np.array_equal(co.iStar, c4.iStar)


@pytest.mark.parametrize("Rb,R", [(1, 4), (3, 3), (4, 9), (5, 12), (8, 8)])
def test_recompute_sweep_matches_streaming(Rb, R):
    """Very few random right-hand sides: the sweep that recomputes delta.pib from (lambda, omega) instead of streaming the table
    (k_sweep_recompute, variant 3; automatic for Rb <= 4).  Same iStar as the oracle, and the cut is bit-identical to the one the
    streaming kernel forms from the stored table (the recomputed entry has the stored entry's bits)."""
    D, N = 900, 1700
    prob = make_problem(31 + Rb, rows=30, cols=40, n1=10, n1c=8, R=R, Rb=Rb, Q=0)
    rng = np.random.default_rng(100 + Rb)
    pis = rng.uniform(-1, 1, (D, prob.rows + 1)) * (rng.random((D, prob.rows + 1)) > 0.3)
    for d in rng.choice(np.arange(1, D), size=D // 40, replace=False):
        pis[d] = pis[rng.integers(0, d)]
    iters = np.ceil((np.arange(D) + 1) * (1.25 * N) / D).astype(np.int32)
    obs = rng.normal(0, 1, (N, prob.numRV + 1)); obs[:, 0] = 0
    obs[rng.choice(N, 16, replace=False)] = 0.0
    weights = (1 + rng.poisson(0.25, N)).astype(np.int32)
    k = int(weights.sum())
    caps = Caps(D + 2, D + 2, D + 2, N + 3, 1)
    to, lo, so = _bulk_tables(oracle_loader.oracle(), prob, pis, np.zeros(D), iters, obs, weights, caps)
    tg, lg, sg = _bulk_tables(sd.load_library(), prob, pis, np.zeros(D), iters, obs, weights, caps)
    x = rng.uniform(0, 1, prob.prevCols + 1); x[0] = 0
    for pi_eval in (0, 1):
        co = to.sd_cut(x, k, pi_eval, 0.0)
        tg.set_sweep_variant(1); c1 = tg.sd_cut(x, k, pi_eval, 0.0)
        assert tg.stats()["last_sweep_variant"] == 1
        tg.set_sweep_variant(3); c3 = tg.sd_cut(x, k, pi_eval, 0.0)
        assert tg.stats()["last_sweep_variant"] == 5
        tg.set_sweep_variant(0); c0 = tg.sd_cut(x, k, pi_eval, 0.0)
        assert tg.stats()["last_sweep_variant"] == (5 if Rb <= 4 else 1)
        assert co is not None and c1 is not None and c3 is not None
        assert np.array_equal(co.iStar, c3.iStar), np.nonzero(co.iStar != c3.iStar)[0][:10]
        assert np.array_equal(c1.iStar, c3.iStar) and np.array_equal(c0.iStar, c3.iStar)
        assert c1.alpha == c3.alpha and np.array_equal(c1.beta, c3.beta) and c1.cummOld == c3.cummOld and c1.cummAll == c3.cummAll
        assert abs(co.alpha - c3.alpha) <= RTOL * abs(co.alpha)
        assert np.abs(co.beta - c3.beta).max() <= RTOL * max(abs(co.alpha), np.abs(co.beta[1:]).max())


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("S,N,Q,phi,density", [(900, 1500, 0, 2, 0.8), (600, 1100, 2, 1, 0.5), (1300, 700, 0, 0, 0.9), (640, 1030, 3, 3, 0.7)])
def test_random_cost_multi_tile_multi_chunk(S, N, Q, phi, density, variant):
    """Random-cost shape (rvdOmCnt > 0) over several observation tiles and basis chunks: multi-term bases whose terms straddle
    ring stages and descriptor batches, a sparse obsFeasible mask, infeasible bases, duplicated duals (ties), both windows.
    variant 1 = the per-term gather kernels, variant 2 = the term-linear TMA ring; iStar must be bit-exact either way."""
    rvd = 4
    prob = make_problem(21, rows=40, cols=60, n1=14, n1c=11, R=17, Rb=13, Q=Q, rvd=rvd)
    rng = np.random.default_rng(S + N + phi)
    pis = rng.uniform(-1, 1, (S, prob.rows + 1)) * (rng.random((S, prob.rows + 1)) > 0.3)
    for d in rng.choice(np.arange(1, S), size=max(1, S // 40), replace=False):
        pis[d] = pis[rng.integers(0, d)]
    nb = S // (1 + phi)
    iters = np.ceil((np.arange(S) + 1) * (1.25 * N) / S).astype(np.int32)
    obs = rng.normal(0, 1, (N, prob.numRV + 1)); obs[:, 0] = 0
    obs[rng.choice(N, 8, replace=False)] = 0.0
    weights = (1 + rng.poisson(0.25, N)).astype(np.int32)
    k = int(weights.sum())
    feas_basis = rng.random(nb) > 0.05
    omega_idx = [[0] + [int(v) for v in rng.integers(1, rvd + 1, phi)] for _ in range(nb)]
    masks = rng.random((nb, N)) < density
    caps = Caps(S + 2, S + 2, nb + 2, N + 3, 1 + phi)
    tabs = []
    for api in (oracle_loader.oracle(), sd.load_library()):
        t = api.create(prob, caps)
        t.omega_append_bulk(obs, weights)
        li, si = t.update_dual_bulk(pis, None, iters, -1.0)
        t.calc_delta_block(0, S, 0, N)
        for b in range(nb):
            sig = [int(si[b * (1 + phi) + j]) for j in range(1 + phi)]
            t.basis_append(int(iters[b * (1 + phi)]), bool(feas_basis[b]), sig, omega_idx[b])
            if feas_basis[b]:
                t.basis_set_obs_feasible_row(b, masks[b].tolist())
        tabs.append(t)
    to, tg = tabs
    tg.set_sweep_variant(variant)
    x = rng.uniform(0, 1, prob.prevCols + 1); x[0] = 0
    for pi_eval in (0, 1):
        co = to.sd_cut(x, k, pi_eval, 0.0); cg = tg.sd_cut(x, k, pi_eval, 0.0)
        assert (co is None) == (cg is None)
        assert tg.stats()["last_sweep_variant"] == (4 if variant == 2 else (3 if phi > 0 else 1))
        if co is None:
            continue
        assert np.array_equal(co.iStar, cg.iStar), np.nonzero(co.iStar != cg.iStar)[0][:10]
        assert len(set(co.iStar.tolist())) > 3
        scale = max(abs(co.alpha), np.abs(co.beta[1:]).max())
        assert abs(co.alpha - cg.alpha) <= RTOL * abs(co.alpha)
        assert np.abs(co.beta - cg.beta).max() <= RTOL * scale
        assert abs(co.cummOld - cg.cummOld) <= RTOL * max(abs(co.cummOld), 1e-300)
        assert abs(co.cummAll - cg.cummAll) <= RTOL * max(abs(co.cummAll), 1e-300)


def test_bulk_dedup_chain_matches_single_calls():
    prob = make_problem(4, rows=30, cols=40, n1=10, n1c=8, R=12, Rb=9, Q=2)
    rng = np.random.default_rng(3)
    pool = rng.uniform(-1, 1, (25, prob.rows + 1))
    pis = pool[rng.integers(0, 25, 120)] + rng.uniform(-3e-4, 3e-4, (120, prob.rows + 1))
    obs = rng.normal(0, 1, (70, prob.numRV + 1))
    iters = np.arange(1, 121, dtype=np.int32)
    caps = Caps(130, 130, 130, 80, 1)
    res = []
    for api in (oracle_loader.oracle(), sd.load_library()):
        t = api.create(prob, caps)
        t.omega_append_bulk(obs, None)
        li, si = t.update_dual_bulk(pis, None, iters, 1e-3)
        res.append((t, li, si))
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    assert res[0][0].counts() == res[1][0].counts()
    assert res[0][0].counts()["lambda"] < 120
    assert_tables_identical(res[0][0], res[1][0])


def test_reset_and_reuse():
    """cleanCellType (setup.c:242-246): tables are emptied between replications and must refill identically."""
    prob = problem_for("pgp2")
    K = 30
    trace = make_trace(prob, K, seed=12, dual_pool=5, obs_pool=6)
    caps = roomy_caps(K)
    api = sd.load_library()
    first = replay(api, prob, trace, caps)
    t = first.tables
    t.reset()
    assert t.counts() == {"omega": 0, "lambda": 0, "sigma": 0, "basis": 0}
    second = replay(api, prob, trace, caps)
    assert_records_match(first, second, exact_cut=True)


def test_null_cut_and_errors():
    prob = problem_for("pgp2")
    api = sd.load_library()
    t = api.create(prob, Caps(8, 8, 8, 8, 1))
    x = np.zeros(prob.prevCols + 1)
    obs = np.arange(prob.numRV + 1, dtype=float)
    t.calc_omega(obs, 1e-3)
    assert t.sd_cut(x, 1, 0, 0.0) is None                      # no basis at all -> NULL cut (cuts.c:136-139)
    pi = np.linspace(-1, 1, prob.rows + 1)
    t.stochastic_updates(0, True, pi, 0.0, 5, 1e-3)            # ck = 5 > numSamples = 1: not eligible
    assert t.sd_cut(x, 1, 0, 0.0) is None
    assert t.sd_cut(x, 5, 0, 0.0) is not None
    with pytest.raises(sd.SdError, match="capacity"):
        for i in range(20):
            t.calc_omega(obs + i + 1, 1e-3)
    with pytest.raises(sd.SdError):
        t.calc_delta(True, 99)


def test_heights_reform_istar_single():
    prob = make_problem(21, rows=20, cols=30, n1=8, n1c=6, R=9, Rb=6, Q=3)
    K = 30
    trace = make_trace(prob, K, seed=4, dual_pool=9, obs_pool=12)
    caps = roomy_caps(K)
    port = replay(oracle_loader.oracle(), prob, trace, caps, lb=-2.0)
    gpu = replay(sd.load_library(), prob, trace, caps, lb=-2.0)
    cuts = [c for c in port.cuts if c is not None][-6:]
    alpha = np.array([c.alpha for c in cuts]); beta = np.stack([c.beta for c in cuts])
    ns = np.array([c.numSamples for c in cuts], np.int32); ai = alpha - 1.0
    xk = trace.xs[-1, 0]
    a = port.tables.cut_heights(alpha, beta, ns, ai, K, xk, -2.0)
    b = gpu.tables.cut_heights(alpha, beta, ns, ai, K, xk, -2.0)
    assert a[0] == b[0] >= 0
    for u, v in zip(a[1:], b[1:]):
        assert np.array_equal(u.view(np.int64), v.view(np.int64))      # sequential per cut: bit-identical
    rng = np.random.default_rng(8)
    last = cuts[-1]
    for lbtype, lbv in ((0, 0.0), (1, -7.9)):
        observ = rng.integers(0, port.counts["omega"] + 3, size=K).astype(np.int32)
        ar, br = port.tables.reform_cut(last.iStar, observ, K, lbtype, lbv)
        ag, bg = gpu.tables.reform_cut(last.iStar, observ, K, lbtype, lbv)
        assert abs(ar - ag) <= RTOL * max(abs(ar), 1e-300)
        assert np.abs(br - bg).max() <= RTOL * max(np.abs(br).max(), 1e-300)
    x = trace.xs[3, 0]
    for obs in range(port.counts["omega"]):
        for pi_eval in (0, 1):
            for is_new in (0, 1):
                p = port.tables.compute_istar(x, obs, K, pi_eval, is_new)
                g = gpu.tables.compute_istar(x, obs, K, pi_eval, is_new)
                assert p[0] == g[0]
                if p[0] >= 0:
                    assert np.float64(p[1]).tobytes() == np.float64(g[1]).tobytes()


@pytest.mark.parametrize("case", ["T_random", "random_cost_T"])
def test_bootstrap_batch_reform(case):
    """reformCuts over cuts x bootstrap replications (optimal.c:96-103) in one launch, against the oracle"""
    pk, K, dpool, opool, phi_len, rk = CASES[case]
    prob = make_problem(1000 + len(case), **pk)
    trace = make_trace(prob, K, seed=77 + K, dual_pool=dpool, obs_pool=opool, phi_len=phi_len)
    caps = roomy_caps(K, phi_len)
    port = replay(oracle_loader.oracle(), prob, trace, caps, **rk)
    gpu = replay(sd.load_library(), prob, trace, caps, **rk)
    cuts = [c for c in port.cuts if c is not None][-6:]
    rng = np.random.default_rng(3)
    for k in (K, 5000):                                       # 5000 > one staging chunk
        observ = rng.integers(0, port.counts["omega"] + 2, size=(9, k)).astype(np.int32)
        ap, bp = port.tables.reform_cuts_batch([c.iStar for c in cuts], observ, 1, -3.0)
        ag, bg = gpu.tables.reform_cuts_batch([c.iStar for c in cuts], observ, 1, -3.0)
        assert np.abs(ap - ag).max() <= RTOL * np.abs(ap).max()
        assert np.abs(bp - bg).max() <= RTOL * np.abs(bp).max()


def test_check_basis_feasibility_on_device():
    """checkBasisFeasibility (randCost.c:202-258) evaluated by the library: flags identical to the oracle, cuts under the mask agree"""
    import feas_scenario
    feas_scenario.compare(feas_scenario.run(oracle_loader.oracle()), feas_scenario.run(sd.load_library()), exact=False)


def test_feasibility_cuts_on_device():
    """raw feasibility cuts (cuts.c:473-490) from the device: the pool built from them is bit-identical to the oracle's"""
    import feas_pool
    sp, ap, bp = feas_pool.run(oracle_loader.oracle())
    sg, ag, bg = feas_pool.run(sd.load_library())
    assert sp == sg and sp[-1] > 5
    assert np.array_equal(ap.view(np.int64), ag.view(np.int64)) and np.array_equal(bp.view(np.int64), bg.view(np.int64))


def test_feasibility_cut_pool_on_device():
    """updtFeasCutPool with the duplicate test of addCut2Pool (cuts.c:465-517,643-655) and checkFeasCutPool (cuts.c:521-567) entirely
    in the library: pool sizes after every update, the pool itself and every action code equal the oracle's"""
    import feas_pool
    p = feas_pool.run_device_pool(oracle_loader.oracle())
    g = feas_pool.run_device_pool(sd.load_library())
    assert p[0] == g[0] and p[0][-1] > 100 and p[4] == g[4]
    assert np.array_equal(p[1].view(np.int64), g[1].view(np.int64)) and np.array_equal(p[2].view(np.int64), g[2].view(np.int64))
    for (ap, ip), (ag, ig) in zip(p[3], g[3]):
        assert ip == ig and np.array_equal(ap, ag)


def test_feasibility_cut_pool_many_raw_cuts():
    """more raw cuts than one resolution batch (4 096), heavy duplication (a pool of 6 dual rays, near-duplicate observations): the
    first-come-first-kept rule with a tolerance is not transitive, so the kept set depends on the order -- it must be the oracle's"""
    from stochasticdecomposition_b200.synthetic import make_problem
    prob = make_problem(62, rows=14, cols=20, n1=6, n1c=5, R=7, Rb=5, Q=2)
    rng = np.random.default_rng(4)
    N, D = 700, 24
    base = rng.normal(0, 1.0, (40, prob.numRV + 1))
    obs = base[rng.integers(0, 40, N)] + rng.uniform(-6e-4, 6e-4, (N, prob.numRV + 1))       # clusters about as wide as the tolerance
    obs[:, 0] = 0.0
    rays = rng.uniform(-1, 1, (6, prob.rows + 1))
    pis = rays[rng.integers(0, 6, D)] + rng.uniform(-3e-4, 3e-4, (D, prob.rows + 1))
    pis[:, 0] = 0.0
    out = []
    ix, cx = rng.normal(0, 1, prob.prevCols + 1), rng.normal(0, 1, prob.prevCols + 1)
    for api in (oracle_loader.oracle(), sd.load_library()):
        t = api.create(prob, Caps(D + 4, D + 4, D + 4, N + 4, 1))
        fUpdt, sizes = [0, 0], []
        for half in range(2):
            for o in obs[half * N // 2:(half + 1) * N // 2]:
                oi, onew = t.calc_omega(o, 1e-7)
                if onew:
                    t.calc_delta(True, oi)
            for d in range(half * D // 2, (half + 1) * D // 2):
                li, nl, si, ns = t.update_dual(pis[d], 0.0, d + 1, 1e-7)
                t.basis_append(d + 1, d % 4 == 0, [si])                    # three of four are infeasible rays
            sizes.append(t.feas_pool_update(fUpdt, 1e-3))
        a, b = t.feas_pool()
        act, inf = t.feas_pool_check(a[:5] + 2e-4, b[:5], ix, cx, 1e-3)
        out.append((sizes, a.copy(), b.copy(), act.copy(), inf, t.counts()))
    (sp, ap, bp, cp, ip, np_), (sg, ag, bg, cg, ig, ng) = out
    assert np_ == ng and np_["omega"] * 18 > 8192, np_                          # > two batches of raw cuts in the second update
    assert sp == sg and 40 < sp[-1] < np_["omega"] * 18
    assert np.array_equal(ap.view(np.int64), ag.view(np.int64)) and np.array_equal(bp.view(np.int64), bg.view(np.int64))
    assert np.array_equal(cp, cg) and ip == ig


@pytest.mark.parametrize("name", ["pgp2_dedup", "T_random", "random_cost_pool_ties"])
def test_latency_features_do_not_change_a_bit(name, monkeypatch):
    """Programmatic dependent launch, the alternating sweep direction and the fused update launch (delta column || lambda scan -> sigma
    in the committing block) are pure scheduling: with all of them off and with all of them on, every table entry, every index and
    every cut must have the same bits."""
    pk, K, dpool, opool, phi_len, rk = CASES[name]
    prob = make_problem(1000 + len(name), **pk)
    trace = make_trace(prob, K, seed=77 + K, dual_pool=dpool, obs_pool=opool, phi_len=phi_len)
    caps = roomy_caps(K, phi_len)
    runs = []
    for on in ("0", "1"):
        for knob in ("SDGPU_PDL", "SDGPU_ALTDIR", "SDGPU_FUSED_UPDATE"):
            monkeypatch.setenv(knob, on)
        runs.append(replay(sd.load_library(), prob, trace, caps, **rk))
    off, on = runs
    assert_records_match(off, on, exact_cut=True)
    assert_tables_identical(off.tables, on.tables)


def test_multi_tile_both_sweep_directions():
    """the load-based sweep alternates its direction from cut to cut (descending order keeps the lowest index with '>='): consecutive
    cuts at the same x must be bit-identical, with duplicated duals (exact ties) in every chunk, and equal the oracle's iStar"""
    D, N = 900, 1300
    prob = make_problem(5, rows=30, cols=40, n1=12, n1c=9, R=11, Rb=9, Q=0)
    rng = np.random.default_rng(2)
    pis = rng.uniform(-1, 1, (D, prob.rows + 1)); pis[:, 0] = 0
    for d in rng.choice(np.arange(1, D), size=D // 10, replace=False):
        pis[d] = pis[rng.integers(0, d)]
    obs = rng.normal(0, 1, (N, prob.numRV + 1)); obs[:, 0] = 0
    w = (1 + rng.poisson(0.3, N)).astype(np.int32)
    iters = np.ceil((np.arange(D) + 1) * (int(w.sum()) / D)).astype(np.int32)
    x = rng.uniform(0, 1, prob.prevCols + 1); x[0] = 0
    t, li, si = _bulk_tables(sd.load_library(), prob, pis, None, iters, obs, w, Caps(D + 2, D + 2, D + 2, N + 3, 1))
    t.set_sweep_variant(1)
    cuts = [t.sd_cut(x, int(w.sum()), pe, 0.0) for pe in (1, 1, 1, 0, 0)]
    for c in cuts[1:3]:
        assert np.array_equal(c.iStar, cuts[0].iStar) and c.alpha == cuts[0].alpha and np.array_equal(c.beta, cuts[0].beta)
    assert np.array_equal(cuts[4].iStar, cuts[3].iStar) and cuts[4].alpha == cuts[3].alpha
    port, _, _ = _bulk_tables(oracle_loader.oracle(), prob, pis, None, iters, obs, w, Caps(D + 2, D + 2, D + 2, N + 3, 1))
    for pe, c in ((1, cuts[0]), (0, cuts[3])):
        pc = port.sd_cut(x, int(w.sum()), pe, 0.0)
        assert np.array_equal(pc.iStar, c.iStar)


def test_instrumentation_and_collective_selection_errors():
    """sdgpu_fp64_peak returns plausible rates (separately rounded multiply + add below the fused rate); sdgpu_set_collective refuses an
    exchange that is not attached; the per-phase split of a cut is reported when timing is on"""
    api = sd.load_library()
    mul_add, fma = api.fp64_peak(0, 2)
    assert 1e12 < mul_add < fma < 1e14, (mul_add, fma)
    prob = problem_for("pgp2")
    t = api.create(prob, Caps(8, 8, 8, 8, 1))
    for mode in (1, 2):
        with pytest.raises(sd.SdError):
            t.set_collective(mode)
    t.set_collective(0)
    t.set_timing(True)
    rng = np.random.default_rng(0)
    ob = rng.normal(0, 1, prob.numRV + 1); ob[0] = 0
    pi = rng.uniform(-1, 1, prob.rows + 1); pi[0] = 0
    oi, onew = t.calc_omega(ob, 1e-3)
    t.stochastic_updates(oi, onew, pi, 0.0, 1, 1e-3)
    x = rng.uniform(0, 1, prob.prevCols + 1); x[0] = 0
    assert t.sd_cut(x, 1, 1, 0.0) is not None
    st = t.stats()
    assert st["last_cut_ms"] > 0 and st["last_sweep_ms"] > 0 and st["last_merge_ms"] > 0 and st["last_collective_ms"] >= 0
    assert abs(st["last_prep_ms"] + st["last_sweep_ms"] + st["last_merge_ms"] + st["last_collective_ms"] - st["last_cut_ms"]) < 0.05 * st["last_cut_ms"] + 1e-3


@pytest.mark.parametrize("Q", [0, 2])
def test_delta_table_mapped_on_demand(Q):
    """A capacity far beyond the GPU (300 000 duals x 400 000 observations: ~1 TB of delta) is accepted: the table's address range is
    reserved and physical memory is mapped under the part in use only (csrc/vmem.cu; SURVEY.md section 7 "reserve VA once, map slabs").
    Bulk build, then single appends that cross a 512-row slab boundary and open a new observation tile; every delta entry, iStar and
    the cut equal the port's, and the mapped bytes stay what the rows and tiles in use need."""
    prob = make_problem(23, rows=40, cols=60, n1=14, n1c=11, R=17, Rb=13, Q=Q)
    rng = np.random.default_rng(77 + Q)
    D, N = 500, 500
    pis = rng.uniform(-1, 1, (D + 40, prob.rows + 1)) * (rng.random((D + 40, prob.rows + 1)) > 0.3)
    obs = rng.normal(0, 1, (N + 40, prob.numRV + 1)); obs[:, 0] = 0
    weights = (1 + rng.poisson(0.25, N)).astype(np.int32)
    iters = np.ceil((np.arange(D) + 1) * (1.25 * N) / D).astype(np.int32)
    to, lo, so = _bulk_tables(oracle_loader.oracle(), prob, pis[:D], np.zeros(D), iters, obs[:N], weights, Caps(D + 50, D + 50, D + 50, N + 50, 1))
    tg, lg, sg = _bulk_tables(sd.load_library(), prob, pis[:D], np.zeros(D), iters, obs[:N], weights, Caps(300000, 300000, 300000, 400000, 1))
    on_demand, reserved, mapped = tg.delta_memory()
    row_bytes = (1 + Q) * 4096
    assert on_demand and reserved >= 8 * (1 + Q) * 300000 * 400000
    assert mapped == 1 * 512 * row_bytes, mapped                       # one tile, one slab of 512 rows
    k = int(weights.sum())
    for i in range(40):                                                # duals 500..539 cross the 512-row slab, observations 500..539 open tile 1
        for t in (to, tg):
            oi, new = t.calc_omega(obs[N + i], 1e-3)
            assert new and oi == N + i
            t.calc_delta(True, oi)
            li, nl, si, ns = t.update_dual(pis[D + i], 0.0, k + i, 1e-3)
            assert nl and li == D + i
            t.basis_append(k + i, True, [si])
    assert tg.delta_memory()[2] == 2 * 1024 * row_bytes                # two tiles x two slabs
    assert to.counts() == tg.counts()
    for plane in range(Q + 1):
        a = to.get_delta_block(0, D + 40, 0, N + 40, plane); b = tg.get_delta_block(0, D + 40, 0, N + 40, plane)
        assert np.array_equal(a.view(np.int64), b.view(np.int64)), f"delta plane {plane} differs"
    x = rng.uniform(0, 1, prob.prevCols + 1); x[0] = 0
    for variant in (1, 2):
        tg.set_sweep_variant(variant)
        for pi_eval in (0, 1):
            co = to.sd_cut(x, k + 40, pi_eval, 0.0); cg = tg.sd_cut(x, k + 40, pi_eval, 0.0)
            assert np.array_equal(co.iStar, cg.iStar)
            assert abs(co.alpha - cg.alpha) <= RTOL * abs(co.alpha) and np.abs(co.beta - cg.beta).max() <= RTOL * max(abs(co.alpha), np.abs(co.beta[1:]).max())
    tg.reset()                                                         # a new replication keeps the mappings (cleanCellType, setup.c:242-246)
    assert tg.delta_memory()[2] == 2 * 1024 * row_bytes and tg.counts()["lambda"] == 0
    tg.close()
