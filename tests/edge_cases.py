"""Edge cases of the table / cut path, runnable on any bound library: empty and ragged inputs, tile boundaries, minimal
dimensions, repeated and NaN observations, all-infeasible bases, capacity edges, two live contexts."""
import numpy as np

from stochasticdecomposition_b200._abi import Caps, SdError
from stochasticdecomposition_b200.synthetic import make_problem


def _fill(t, prob, n_obs, n_dual, seed, feas=True, tol=1e-3, obs_scale=3.0):
    rng = np.random.default_rng(seed)
    for i in range(n_obs):
        o = rng.normal(0, obs_scale, prob.numRV + 1); o[0] = 0
        oi, new = t.calc_omega(o, tol)
        if new:
            t.calc_delta(True, oi)
    out = []
    for i in range(n_dual):
        pi = rng.uniform(-1, 1, prob.rows + 1); pi[0] = 0
        li, nl, si, ns = t.update_dual(pi, 0.1 * i, i + 1, tol)
        out.append(t.basis_find_or_append(ns, 0, i + 1, feas, [si], None) if n_obs else (t.basis_append(i + 1, feas, [si]), True))
    return out


def _cut_tuple(c):
    if c is None:
        return None
    with np.errstate(divide="ignore", invalid="ignore"):        # the reference only exposes cummOld / cummAll as their ratio (cuts.c:172)
        ratio = float(np.float64(c.cummOld) / np.float64(c.cummAll))
    return (c.alpha, c.beta.copy(), c.iStar.copy(), ratio, c.omegaCnt)


def case_no_observations(api):
    prob = make_problem(1, rows=10, cols=14, n1=5, n1c=4, R=4, Rb=3, Q=1)
    t = api.create(prob, Caps(8, 8, 8, 8, 1))
    _fill(t, prob, 0, 3, 1)
    x = np.linspace(0, 1, prob.prevCols + 1)
    return [_cut_tuple(t.sd_cut(x, 3, pe, 0.0)) for pe in (0, 1)], t.counts()


def case_tile_boundaries(api):
    prob = make_problem(2, rows=9, cols=12, n1=4, n1c=4, R=3, Rb=3)
    res = []
    for n in (1, 511, 512, 513, 1025):
        t = api.create(prob, Caps(40, 40, 40, n + 1, 1))
        rng = np.random.default_rng(n)
        obs = rng.normal(0, 2, (n, prob.numRV + 1)); obs[:, 0] = 0
        t.omega_append_bulk(obs, (1 + rng.poisson(0.5, n)).astype(np.int32))
        pis = rng.uniform(-1, 1, (33, prob.rows + 1)); pis[:, 0] = 0
        li, si = t.update_dual_bulk(pis, None, np.arange(1, 34, dtype=np.int32), 1e-3)
        for s in sorted(set(si.tolist())):
            t.basis_append(int(s) + 1, True, [int(s)])
        x = rng.uniform(0, 1, prob.prevCols + 1); x[0] = 0
        res.append([_cut_tuple(t.sd_cut(x, 40, pe, -0.5)) for pe in (0, 1)])
        t.close()
    return res


def case_minimal_dims(api):
    prob = make_problem(3, rows=1, cols=2, n1=1, n1c=1, R=1, Rb=1)
    t = api.create(prob, Caps(30, 30, 30, 30, 1))
    _fill(t, prob, 20, 12, 3)
    x = np.array([0.0, 0.7])
    return [_cut_tuple(t.sd_cut(x, 25, pe, 0.0)) for pe in (0, 1)], t.counts()


def case_no_random_rhs_only_T(api):
    prob = make_problem(4, rows=8, cols=10, n1=5, n1c=5, R=4, Rb=0, Q=3)
    t = api.create(prob, Caps(30, 30, 30, 30, 1))
    _fill(t, prob, 15, 10, 4)
    x = np.linspace(0, 2, prob.prevCols + 1); x[0] = 0
    return [_cut_tuple(t.sd_cut(x, 20, pe, 0.0)) for pe in (0, 1)], t.counts()


def case_all_bases_infeasible(api):
    prob = make_problem(5, rows=8, cols=10, n1=4, n1c=3, R=4, Rb=4)
    t = api.create(prob, Caps(20, 20, 20, 20, 1))
    _fill(t, prob, 6, 5, 5, feas=False)
    x = np.zeros(prob.prevCols + 1)
    return t.sd_cut(x, 6, 1, 0.0), t.counts()


def case_repeats_and_nan(api):
    """calcOmega's DBL_ABS test never flags a NaN difference as a mismatch (stocUpdate.c:331 through equalVector)"""
    prob = make_problem(6, rows=8, cols=10, n1=4, n1c=3, R=4, Rb=4)
    t = api.create(prob, Caps(20, 20, 20, 20, 1))
    base = np.array([0.0, 1.0, -2.0, 0.5, 3.0])
    seq = [base, base + 5e-4, base + 2e-3, base, np.array([0.0, np.nan, -2.0, 0.5, 3.0]), np.array([0.0, np.inf, 0, 0, 0]), base * 0]
    out = [t.calc_omega(o, 1e-3) for o in seq]
    ws = [t.get_omega(i)[1] for i in range(t.counts()["omega"])]
    return out, ws


def case_capacity_edges(api):
    prob = make_problem(7, rows=8, cols=10, n1=4, n1c=3, R=4, Rb=4)
    t = api.create(prob, Caps(3, 3, 3, 4, 1))
    rng = np.random.default_rng(7)
    got = []
    for i in range(4):
        o = rng.normal(0, 3, prob.numRV + 1); o[0] = 0
        got.append(t.calc_omega(o, 1e-3))
    try:
        t.calc_omega(rng.normal(0, 3, prob.numRV + 1), 1e-3)
        got.append("no error")
    except SdError:
        got.append("omega full")
    for i in range(3):
        pi = rng.uniform(-1, 1, prob.rows + 1)
        got.append(t.update_dual(pi, 0.0, i + 1, 1e-3))
    try:
        t.update_dual(rng.uniform(-1, 1, prob.rows + 1), 0.0, 9, 1e-3)
        got.append("no error")
    except SdError:
        got.append("lambda full")
    got.append(t.update_dual(pi, 0.0, 10, 1e-3))            # a repeat still resolves after the overflow
    return got, t.counts()


def case_two_contexts(api):
    pa = make_problem(8, rows=8, cols=10, n1=4, n1c=3, R=4, Rb=4)
    pb = make_problem(9, rows=11, cols=12, n1=6, n1c=5, R=5, Rb=3, Q=2)
    ta, tb = api.create(pa, Caps(20, 20, 20, 20, 1)), api.create(pb, Caps(20, 20, 20, 20, 1))
    ra, rb = np.random.default_rng(1), np.random.default_rng(2)
    out = []
    for i in range(10):
        for t, p, r in ((ta, pa, ra), (tb, pb, rb)):
            o = r.normal(0, 3, p.numRV + 1); o[0] = 0
            oi, new = t.calc_omega(o, 1e-3)
            pi = r.uniform(-1, 1, p.rows + 1); pi[0] = 0
            bi, bn = t.stochastic_updates(oi, new, pi, 0.0, i + 1, 1e-3)
            x = r.uniform(0, 1, p.prevCols + 1); x[0] = 0
            out.append((oi, new, bi, bn, _cut_tuple(t.sd_cut(x, i + 1, 1, 0.0))))
    return out


CASES = {f.__name__[5:]: f for f in (case_no_observations, case_tile_boundaries, case_minimal_dims, case_no_random_rhs_only_T,
                                     case_all_bases_infeasible, case_repeats_and_nan, case_capacity_edges, case_two_contexts)}


def same(a, b, exact, rtol=1e-9):
    """structural comparison: ints / bools / strings exact, floats exact or within rtol, arrays elementwise"""
    if isinstance(a, (list, tuple)):
        assert type(a) is type(b) and len(a) == len(b), (a, b)
        for u, v in zip(a, b):
            same(u, v, exact, rtol)
    elif isinstance(a, dict):
        assert a == b
    elif isinstance(a, np.ndarray):
        if a.dtype.kind in "iu" or exact:
            assert np.array_equal(a, b, equal_nan=True), (a, b)
        else:
            scale = max(float(np.abs(a).max()) if a.size else 0.0, 1e-300)
            assert np.abs(a - b).max() <= rtol * scale, (a, b)
    elif isinstance(a, float):
        if exact or a != a:
            assert a == b or (a != a and b != b), (a, b)
        else:
            assert abs(a - b) <= rtol * max(abs(a), 1e-300), (a, b)
    else:
        assert a == b, (a, b)
