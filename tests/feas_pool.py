"""Host half of updtFeasCutPool (cuts.c:465-517) + the FEASIBILITY branch of addCut2Pool (cuts.c:643-655), around the
library's raw feasibility-cut generator.  Used by the parity tests; the C version is in integration/sdgpu_hooks.c."""
import numpy as np


class FeasCutPool:
    def __init__(self, tol):
        self.alpha, self.beta, self.tol = [], [], tol
        self.fUpdt = [0, 0]                                    # cell->fUpdt, twoSD.h:130

    def _add(self, a, b):
        for pa, pb in zip(self.alpha, self.beta):              # cuts.c:645-653
            d = a - pa
            if (d if d > 0 else -d) < self.tol and not (np.abs(b[1:] - pb[1:]) > self.tol).any():
                return False
        self.alpha.append(float(a)); self.beta.append(b.copy())
        return True

    def update(self, t):
        c = t.counts()
        a, b = t.feas_cuts(self.fUpdt[1], c["omega"], 0, self.fUpdt[0])            # cuts.c:472-489
        for i in range(len(a)):
            self._add(a[i], b[i])
        self.fUpdt[1] = c["omega"]
        a, b = t.feas_cuts(0, c["omega"], self.fUpdt[0], c["basis"])               # cuts.c:494-511
        for i in range(len(a)):
            self._add(a[i], b[i])
        self.fUpdt[0] = c["basis"]
        return len(self.alpha)


def run(api, use_reference_pool=False, K=40, seed=9):
    import ctypes as C
    from replay import replay
    from stochasticdecomposition_b200._abi import Caps, _pf64
    from stochasticdecomposition_b200.synthetic import make_problem, make_trace
    prob = make_problem(61, rows=16, cols=24, n1=7, n1c=5, R=8, Rb=6, Q=3, distinct_rvCols=True)
    trace = make_trace(prob, K, seed=seed, dual_pool=10, obs_pool=14)
    n = 2 * K + 2
    t = api.create(prob, Caps(n, n, n, K + 1, 1))
    pool = FeasCutPool(1e-3)
    fUpdt = (C.c_int * 2)(0, 0)
    sizes = []
    ra, rb = np.zeros(4096), np.zeros((4096, prob.prevCols + 1))
    for it in range(K):
        k = it + 1
        oi, onew = t.calc_omega(trace.observ[it], 1e-3)
        feas = (k % 3 != 0)                                    # every third solve is an infeasible subproblem (subprob.c:47-52)
        t.stochastic_updates(oi, onew, trace.duals[it, 0], trace.mubBar[it, 0], k, 1e-3, feas)
        if not feas or k % 5 == 0:                             # formFeasCut (cuts.c:450-460) runs on infeasibility or a new observation
            if use_reference_pool:
                cnt = api._fn("updt_feas_cut_pool")(t.ctx, fUpdt, 1e-3, 4096, _pf64(ra), _pf64(rb))
                assert cnt >= 0
                sizes.append(cnt)
            else:
                sizes.append(pool.update(t))
    if use_reference_pool:
        return sizes, ra[:sizes[-1]].copy(), rb[:sizes[-1]].copy()
    return sizes, np.array(pool.alpha), np.array(pool.beta)


def run_device_pool(api, K=60, seed=9):
    """The whole of updtFeasCutPool / addCut2Pool / checkFeasCutPool behind the C ABI (sdgpu_feas_pool_*): the pool lives in the
    library.  Returns the pool sizes after every update, the final pool, and the (added-to-master?, infeasIncumb) answers of
    checkFeasCutPool at several (incumbent, candidate) pairs, with a few pool cuts already "in the master"."""
    from stochasticdecomposition_b200._abi import Caps
    from stochasticdecomposition_b200.synthetic import make_problem, make_trace
    prob = make_problem(61, rows=16, cols=24, n1=7, n1c=5, R=8, Rb=6, Q=3, distinct_rvCols=True)
    trace = make_trace(prob, K, seed=seed, dual_pool=10, obs_pool=14)
    n = 2 * K + 2
    t = api.create(prob, Caps(n, n, n, K + 1, 1))
    fUpdt, sizes, checks = [0, 0], [], []
    rng = np.random.default_rng(seed)
    for it in range(K):
        k = it + 1
        oi, onew = t.calc_omega(trace.observ[it], 1e-3)
        feas = (k % 3 != 0)
        t.stochastic_updates(oi, onew, trace.duals[it, 0], trace.mubBar[it, 0], k, 1e-3, feas)
        if not feas or k % 5 == 0:
            sizes.append(t.feas_pool_update(fUpdt, 1e-3))
            alpha, beta = t.feas_pool()
            pick = rng.choice(len(alpha), size=min(3, len(alpha)), replace=False) if len(alpha) else []
            fA = np.array([alpha[p] + 4e-4 for p in pick])                  # within tolerance of a pool cut: counts as "already in the master"
            fB = np.array([beta[p] for p in pick]).reshape(len(pick), prob.prevCols + 1)
            ix, cx = rng.normal(0, 1.5, prob.prevCols + 1), rng.normal(0, 1.5, prob.prevCols + 1)
            act, inf = t.feas_pool_check(fA, fB, ix, cx, 1e-3)
            checks.append((np.asarray(act).copy(), inf))
    alpha, beta = t.feas_pool()
    return sizes, alpha.copy(), beta.copy(), checks, fUpdt
