#!/usr/bin/env python
"""A stochastic-decomposition host loop around the table / cut library, for end-to-end parity and iterations/s.

TEST AND BENCHMARK HARNESS -- not part of the product.  The reference's host (algo.c, cuts.c:22-89, soln.c,
master.c) drives IBM CPLEX and reads SMPS files; neither exists offline.  This module restates that host loop in
Python with **HiGHS** (scipy's bundled core, `scipy.optimize._highspy._core`) for the single subproblem LP and the
single regularised master QP per iteration, on synthetic two-stage SLPs with the reference problems' shapes.  Every
number produced with it must be read as "HiGHS, not CPLEX; synthetic instance, not the SMPS file".

What it mirrors (file:line under /root/reference/twoSD_src):
  solveCell main loop            algo.c:127-183     sample, calcOmega, formSDCut (candidate, and incumbent every TAU),
                                                   checkImprovement, solveQPMaster
  formSDCut                      cuts.c:22-89       solve subproblem -> stochasticUpdates -> SDCut -> addCut2Pool
  addCut2Pool / reduceCuts       cuts.c:277-320,616-661
  checkImprovement / replaceIncumbent   soln.c:24-95
  solveQPMaster                  master.c:18-88     (master kept in x-space: eta coefficient k/j, master.c:152)
  pi_eval gate, dual stability   cuts.c:112,171-182
The table and cut arithmetic itself is NOT here: it happens behind the C ABI (`Tables`), in whichever library the
caller binds -- the CUDA library, the CPU oracle, or several at once in lock step (`Lockstep`).
"""
from __future__ import annotations

import os
import sys
import time
from dataclasses import dataclass, field

import numpy as np
from scipy.optimize._highspy import _core as hs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from stochasticdecomposition_b200._abi import Caps, Problem  # noqa: E402

CANDIDATE, INCUMBENT = 0, 1


@dataclass
class Config:
    """config.sd defaults (reference config.sd:1-136) that the loop reads"""
    TOLERANCE: float = 1e-3
    TAU: int = 2
    CUT_MULT: int = 1
    MIN_QUAD_SCALAR: float = 1e-3
    MAX_QUAD_SCALAR: float = 1e4
    R1: float = 0.2
    R2: float = 0.95
    R3: float = 2.0
    DUAL_STABILITY: int = 1
    PI_EVAL_START: int = 0
    PI_CYCLE: int = 1
    SCAN_LEN: int = 256


# ----------------------------------------------------------------------------------------------------------------
# synthetic two-stage SLP with complete recourse
# ----------------------------------------------------------------------------------------------------------------
class SyntheticSLP:
    """min c.x + E[h(x, w)] over x in R^n1 (no first-stage constraints);  h = min d.y s.t. W y = rbar + w - T x, y >= 0 with
    W = [W0 | I | -I] (penalised slack / surplus columns => every subproblem is feasible and bounded below by 0; the
    penalties dominate c, so the problem is bounded)."""

    def __init__(self, seed: int, n1: int, rows: int, core_cols: int, R: int, levels: int = 0, density: float = 0.08, Q: int = 0,
                 rvd: int = 0):
        rng = np.random.default_rng(seed)
        self.n1, self.rows, self.R, self.levels, self.Q, self.rvd = n1, rows, R, levels, Q, rvd
        self.c = rng.uniform(0.5, 1.5, n1)
        self.xu = np.full(n1, 10.0)
        self.budget = 2.5 * n1
        W0 = rng.uniform(0.2, 1.0, (rows, core_cols)) * (rng.random((rows, core_cols)) < density)
        for j in range(core_cols):                      # no empty columns
            if not W0[:, j].any():
                W0[rng.integers(rows), j] = rng.uniform(0.2, 1.0)
        self.W = np.hstack([W0, np.eye(rows), -np.eye(rows)])
        self.cols = self.W.shape[1]
        self.d = np.concatenate([rng.uniform(1.0, 2.0, core_cols), np.full(rows, 8.0), np.full(rows, 3.0)])
        T = rng.uniform(0.1, 1.0, (rows, n1)) * (rng.random((rows, n1)) < density)
        for j in range(n1):                             # every first-stage column appears (cntCcols == n1)
            if not T[:, j].any():
                T[rng.integers(rows), j] = rng.uniform(0.1, 1.0)
        self.T = T
        self.rbar = rng.uniform(2.0, 6.0, rows)
        self.rv_rows = np.sort(rng.choice(rows, size=R, replace=False))      # 0-based rows with a random right-hand side
        self.scale = rng.uniform(0.5, 2.0, R)
        # Q random technology-matrix elements T[r_e][c_e] + dT_e(w), on rows that already carry a random right-hand side
        tr = rng.choice(self.rv_rows, size=Q, replace=True) if Q else np.zeros(0, np.int64)
        tc = rng.choice(n1, size=Q, replace=False) if Q else np.zeros(0, np.int64)
        order = np.lexsort((tr, tc))
        self.t_rows, self.t_cols = tr[order], tc[order]
        self.t_scale = rng.uniform(0.05, 0.3, Q)
        # rvd random cost coefficients d_j + dd_j(w) on core columns (the v2.0 "randCost" path, randCost.c); |dd| < min d
        self.d_cols = np.sort(rng.choice(core_cols, size=rvd, replace=False)) if rvd else np.zeros(0, np.int64)
        self.d_scale = rng.uniform(0.2, 0.6, rvd)
        if levels:
            self.level_vals = np.sort(rng.uniform(-1.0, 1.0, (R, levels)), axis=1)
            self.level_vals -= self.level_vals.mean(axis=1, keepdims=True)   # zero mean: observations are deviations (algo.c:148-149)
        self.lb = 0.0

    def sample(self, rng) -> np.ndarray:
        """one observation of the random right-hand side, as a deviation from its mean, 1-based"""
        if self.levels:
            w = self.level_vals[np.arange(self.R), rng.integers(0, self.levels, self.R)] * self.scale
            wt = (rng.integers(0, self.levels, self.Q) - (self.levels - 1) / 2.0) / max(1, self.levels - 1) * 2.0 * self.t_scale
            wd = (rng.integers(0, self.levels, self.rvd) - (self.levels - 1) / 2.0) / max(1, self.levels - 1) * 2.0 * self.d_scale
        else:
            w = rng.uniform(-1.0, 1.0, self.R) * self.scale
            wt = rng.uniform(-1.0, 1.0, self.Q) * self.t_scale
            wd = rng.uniform(-1.0, 1.0, self.rvd) * self.d_scale
        return np.concatenate([[0.0], w, wt, wd])

    def problem(self) -> Problem:
        """numType / coordType / bBar / Cbar as the tables want them (1-based)"""
        one = lambda a, dt: np.concatenate([np.zeros(1, dt), np.asarray(a, dt)])
        cc = np.nonzero(self.T.any(axis=0))[0] + 1
        tr, tc = np.nonzero(self.T)
        order = np.lexsort((tr, tc))                    # column-major nnz order, as a column-wise sparse matrix would list it
        tr, tc = tr[order], tc[order]
        rvrows = self.rv_rows + 1
        return Problem(rows=self.rows, cols=self.cols, prevCols=self.n1, CCols=one(cc, np.int32), rvRows=one(rvrows, np.int32),
                       rvbOmRows=one(rvrows, np.int32), rvCOmCols=one(self.t_cols + 1, np.int32), rvCOmRows=one(self.t_rows + 1, np.int32),
                       rvCols=one(self.t_cols + 1, np.int32),
                       bBar_col=one(np.arange(1, self.rows + 1), np.int32), bBar_val=one(self.rbar, np.float64),
                       Cbar_col=one(tc + 1, np.int32), Cbar_row=one(tr + 1, np.int32), Cbar_val=one(self.T[tr, tc], np.float64),
                       rvdOmCnt=self.rvd, rvOffset=(0, self.R, self.R + self.Q))

    def cost_coords(self):
        """coord->rvdOmCols (1-based) and the row senses prob->sp->senx, for checkBasisFeasibility (randCost.c:202)"""
        return np.concatenate([[0], self.d_cols + 1]).astype(np.int32), b"E" * self.rows


SHAPES = {
    "pgp2": dict(n1=4, rows=7, core_cols=9, R=3, levels=4, density=0.5),
    "20term": dict(n1=63, rows=124, core_cols=516, R=40, levels=2, density=0.06),
    "20term_T": dict(n1=63, rows=124, core_cols=516, R=40, levels=3, density=0.06, Q=8),     # RHS + technology-matrix randomness
    "ssn": dict(n1=89, rows=175, core_cols=356, R=86, levels=5, density=0.05),
    "storm": dict(n1=121, rows=528, core_cols=203, R=118, levels=5, density=0.02),
    "randcost_small": dict(n1=8, rows=14, core_cols=24, R=8, levels=3, density=0.3, Q=2, rvd=3),   # storm-style random cost, small enough for tests
    "storm_rc": dict(n1=121, rows=528, core_cols=203, R=118, levels=5, density=0.02, rvd=4),        # BASELINE config 4: storm's shape with random cost coefficients
}


def make_slp(name: str, seed: int = 20240607) -> SyntheticSLP:
    return SyntheticSLP(seed, **SHAPES[name])


# ----------------------------------------------------------------------------------------------------------------
# HiGHS wrappers
# ----------------------------------------------------------------------------------------------------------------
def _new_highs():
    h = hs._Highs()
    h.setOptionValue("output_flag", False)
    h.setOptionValue("threads", 1)
    return h


def _csc(M):
    M = np.asarray(M)
    start, index, value = [0], [], []
    for j in range(M.shape[1]):
        nz = np.nonzero(M[:, j])[0]
        index.extend(nz.tolist()); value.extend(M[nz, j].tolist()); start.append(len(index))
    return np.array(start, np.int32), np.array(index, np.int32), np.array(value, np.float64)


class Subproblem:
    """the single second-stage LP: min d.y s.t. W y = rhs, y >= 0; only the right-hand side changes (subprob.c:96-128)"""

    def __init__(self, slp: SyntheticSLP):
        self.slp, self.h = slp, _new_highs()
        self.h.setOptionValue("presolve", "off")          # subprob.c:43 (PARAM_PREIND off)
        self.h.setOptionValue("solver", "simplex")
        lp = hs.HighsLp()
        lp.num_col_, lp.num_row_ = slp.cols, slp.rows
        lp.col_cost_ = slp.d
        lp.col_lower_, lp.col_upper_ = np.zeros(slp.cols), np.full(slp.cols, hs.kHighsInf)
        lp.row_lower_, lp.row_upper_ = slp.rbar.copy(), slp.rbar.copy()
        lp.a_matrix_.format_ = hs.MatrixFormat.kColwise
        lp.a_matrix_.start_, lp.a_matrix_.index_, lp.a_matrix_.value_ = _csc(slp.W)
        self.h.passModel(lp)
        self.solves = 0

    def solve(self, x: np.ndarray, w: np.ndarray):
        """x: first-stage point [n1]; w: observation deviations [R + Q].  Returns (objective, row duals pi 1-based [rows+1])."""
        slp = self.slp
        rhs = slp.rbar - slp.T @ x                                  # computeRHS subprob.c:96-128
        rhs[slp.rv_rows] += w[:slp.R]
        for e in range(slp.Q):
            rhs[slp.t_rows[e]] -= w[slp.R + e] * x[slp.t_cols[e]]
        for j in range(slp.rvd):                                    # computeCostCoeff subprob.c:131-168
            self.h.changeColCost(int(slp.d_cols[j]), float(slp.d[slp.d_cols[j]] + w[slp.R + slp.Q + j]))
        for i in range(slp.rows):
            self.h.changeRowBounds(i, rhs[i], rhs[i])
        self.h.run()
        self.solves += 1
        if self.h.getModelStatus() != hs.HighsModelStatus.kOptimal:
            raise RuntimeError(f"subproblem not optimal: {self.h.modelStatusToString(self.h.getModelStatus())}")
        sol = self.h.getSolution()
        pi = np.concatenate([[0.0], np.asarray(sol.row_dual)])
        return self.h.getInfo().objective_function_value, pi


def basis_info(sub: "Subproblem", wd: np.ndarray, pi: np.ndarray):
    """What newBasis / calcBasis / decomposeDualSolution (randCost.c:19-200) extract from the solver after a solve with random
    costs: the basis code, for every BASIC column with a random cost its row of the basis inverse (phi) and its position in
    the cost block (omegaIdx), the deterministic part of the dual (piDet = pi - sum phi * dd), the deterministic reduced costs
    (gBar), the tableau entries of the phi rows (psi) and the column statuses.  wd = this observation's cost deltas [rvd]."""
    slp, h = sub.slp, sub.h
    rows, cols = slp.rows, slp.cols
    b = h.getBasis()
    cstat = np.array([int(v) for v in b.col_status], np.int32)
    rstat = np.array([int(v) for v in b.row_status], np.int32)
    key = (cstat.tobytes(), rstat.tobytes())                        # encodeIntvec of cstat / rstat (randCost.c:171-172)
    head = h.getBasicVariables()[1]
    phis, om, heads = [], [], []
    for i, col in enumerate(slp.d_cols):                            # randCost.c:36-49
        pos = np.nonzero(head == col)[0]
        if len(pos):
            phis.append(np.concatenate([[0.0], h.getBasisInverseRow(int(pos[0]))[1]]))
            om.append(i + 1); heads.append(int(pos[0]))
    piDet = pi.copy()                                               # decomposeDualSolution randCost.c:182-200
    for n, ph in enumerate(phis):
        piDet[1:] -= ph[1:] * wd[om[n] - 1]
    basic_cost = np.array([slp.d[c] if c >= 0 else 0.0 for c in head])
    gBar = np.zeros(cols + 1); psi = np.zeros((cols, len(phis)))
    if cols <= 256:
        for i in range(cols):                                       # randCost.c:78-89, one tableau column B^-1 A_i at a time as the reference asks the solver
            colv = h.getReducedColumn(i)[1]
            gBar[i + 1] = slp.d[i] - float(np.dot(colv, basic_cost))
            for n, p in enumerate(heads):
                psi[i, n] = colv[p]
    else:
        # the same quantities without 1 259 solver calls per basis (storm shape): (B^-1 A_i) . c_B = A_i . (B^-T c_B) = A_i . piDet, and the
        # tableau entry (B^-1 A_i)[p_n] = (row p_n of B^-1) . A_i = phi_n . A_i -- two matrix products.  Equal in exact arithmetic; every
        # table backend of a lock-step run is fed the same numbers either way.
        gBar[1:] = slp.d - slp.W.T @ piDet[1:]
        for n, ph in enumerate(phis):
            psi[:, n] = slp.W.T @ ph[1:]
    return dict(key=key, phi=phis, omegaIdx=om, piDet=piDet, gBar=gBar, psi=psi, cstat=np.concatenate([[0], cstat]).astype(np.int32))


@dataclass
class OneCut:                     # twoSD.h:69-80
    alpha: float
    beta: np.ndarray              # 1-based, beta[0] = 1
    numSamples: int
    omegaCnt: int
    iStar: np.ndarray | None
    type: int = CANDIDATE
    rowNum: int = -1              # row in the most recent master solve, -1 if not in it yet


def cut_height(cut: OneCut, k: int, x1: np.ndarray, lb: float) -> float:
    """cutHeight cuts.c:213-227 (x1 is 1-based)"""
    t_over_k = cut.numSamples / k
    h = cut.alpha - float(np.dot(cut.beta[1:], x1[1:]))
    h *= t_over_k
    h += (1 - t_over_k) * lb
    return h


def max_cut_height(cuts, k, x1, lb) -> float:
    """maxCutHeight cuts.c:197-209"""
    sm = -1e20
    for c in cuts:
        h = cut_height(c, k, x1, lb)
        if sm < h:
            sm = h
    return sm


def solve_simplex_qp(H: np.ndarray, q: np.ndarray, tol: float = 1e-11, max_iter: int = 2000) -> np.ndarray:
    """min 1/2 t'Ht - q't over the unit simplex (H PSD), by a primal active-set method on the support of t.
    Used for the dual of the regularised master; sizes are <= n1 + 4."""
    m = len(q)
    t = np.zeros(m)
    j0 = int(np.argmax(q - 0.5 * np.diag(H)))
    t[j0] = 1.0
    S = [j0]
    ridge = 1e-13 * max(1.0, float(np.trace(H)) / m)
    for _ in range(max_iter):
        k = len(S)
        K = np.zeros((k + 1, k + 1))
        K[:k, :k] = H[np.ix_(S, S)] + ridge * np.eye(k)
        K[:k, k] = 1.0
        K[k, :k] = 1.0
        rhs = np.concatenate([q[S], [1.0]])
        sol = np.linalg.lstsq(K, rhs, rcond=None)[0]
        cand, nu = sol[:k], sol[k]
        if (cand >= -tol).all():
            t[:] = 0.0
            t[S] = np.maximum(cand, 0.0)
            t /= t.sum()
            grad = H @ t - q                      # KKT: grad_j + nu >= 0 off the support
            viol = grad + nu
            viol[S] = 0.0
            j = int(np.argmin(viol))
            if viol[j] >= -tol * max(1.0, float(np.abs(grad).max())):
                return t
            S.append(j)
            continue
        cur = t[S]
        step = cand - cur                         # move towards the subspace minimiser until a component hits zero
        neg = step < 0
        ratios = np.where(neg, cur / np.where(neg, -step, 1.0), np.inf)
        r = float(ratios.min())
        new = cur + min(1.0, r) * step
        drop = int(np.argmin(ratios))
        new[drop] = 0.0
        t[:] = 0.0
        t[S] = np.maximum(new, 0.0)
        del S[drop]
        if not S:
            S = [int(np.argmax(t))] if t.any() else [j0]
        t /= t.sum() if t.sum() > 0 else 1.0
    return t


class Master:
    """The regularised master of SD (master.c:18-88), kept in x-space, for a first stage without constraints:

        min  c.x + eta + quad/2 |x - xbar|^2   s.t.  eta >= r_j (alpha_j - beta_j.x) + (1 - r_j) lb,  r_j = numSamples_j / k
                                                     (the k/j coefficient of changeEtaCol master.c:152 and the shift of updateRHS
                                                      master.c:174, divided through), eta >= lb

    solved exactly through its dual, a QP over the unit simplex of cut multipliers theta:
        x(theta) = xbar - (c - G'theta)/quad,   G_j = r_j beta_j,   a_j = r_j alpha_j + (1 - r_j) lb.
    (HiGHS' own QP solver cycles on these highly degenerate models, so it is not used here; the subproblem LP is HiGHS.)
    The multipliers are what reduceCuts (cuts.c:290) reads as the cut duals."""

    def __init__(self, slp: SyntheticSLP):
        self.slp, self.solves = slp, 0

    def solve(self, cuts, k: int, xbar: np.ndarray, quad: float, lb: float):
        slp, n1 = self.slp, self.slp.n1
        m = len(cuts) + 1
        G = np.zeros((m, n1))
        a = np.zeros(m)
        a[0] = lb                                                     # eta >= lb
        for j, cth in enumerate(cuts):
            r = cth.numSamples / k
            G[1 + j] = r * cth.beta[1:]
            a[1 + j] = r * cth.alpha + (1.0 - r) * lb
            cth.rowNum = 1 + j
        H = (G @ G.T) / quad
        q = (G @ slp.c) / quad + (a - G @ xbar)
        theta = solve_simplex_qp(H, q)
        x = xbar - (slp.c - G.T @ theta) / quad
        self.solves += 1
        return x, theta


# ----------------------------------------------------------------------------------------------------------------
# several table backends in lock step
# ----------------------------------------------------------------------------------------------------------------
class Lockstep:
    """Drives several `Tables` with identical calls and checks that they agree: indices and flags exactly, iStar exactly,
    cut coefficients / cummOld / cummAll within `rtol` relative (1e-9, BASELINE.json).  The first backend's answers are used."""

    def __init__(self, backends, rtol=1e-9):
        self.b, self.rtol, self.checked = list(backends), rtol, 0
        self.problem = self.b[0].problem

    def _same(self, outs, what):
        for o in outs[1:]:
            assert o == outs[0], f"{what}: backends disagree: {outs}"
        return outs[0]

    def calc_omega(self, observ, tol):
        return self._same([t.calc_omega(observ, tol) for t in self.b], "calc_omega")

    def reset(self):
        """cleanCellType (setup.c:242-246) on every backend"""
        for t in self.b:
            t.reset()

    def stochastic_updates(self, *a, **kw):
        return self._same([t.stochastic_updates(*a, **kw) for t in self.b], "stochastic_updates")

    def counts(self):
        return self._same([t.counts() for t in self.b], "counts")

    def _all(self, name, *a, **kw):
        outs = [getattr(t, name)(*a, **kw) for t in self.b]
        for o in outs[1:]:
            if isinstance(outs[0], np.ndarray):
                assert np.array_equal(o, outs[0]), f"{name}: backends disagree"
            elif isinstance(outs[0], tuple) and len(outs[0]) == 2 and isinstance(outs[0][1], float):
                assert o[0] == outs[0][0] and (o[0] < 0 or abs(o[1] - outs[0][1]) <= max(self.rtol, 1e-15) * max(abs(outs[0][1]), 1e-300)), (name, o, outs[0])
            else:
                assert o == outs[0], f"{name}: backends disagree: {outs}"
        return outs[0]

    def calc_delta(self, *a): return self._all("calc_delta", *a)
    def set_cost_coords(self, *a): return self._all("set_cost_coords", *a)
    def basis_set_feas_data(self, *a): return self._all("basis_set_feas_data", *a)
    def check_feasibility_obs(self, *a): return self._all("check_feasibility_obs", *a)
    def check_feasibility_basis(self, *a): return self._all("check_feasibility_basis", *a)
    def compute_istar(self, *a): return self._all("compute_istar", *a)
    def get_omega(self, *a): return self.b[0].get_omega(*a)

    def sd_cut(self, X, numSamples, pi_eval_flag, lb):
        cuts = [t.sd_cut(X, numSamples, pi_eval_flag, lb) for t in self.b]
        ref = cuts[0]
        for c in cuts[1:]:
            assert (c is None) == (ref is None)
            if ref is None:
                continue
            assert np.array_equal(c.iStar, ref.iStar), f"iStar differs at {np.nonzero(c.iStar != ref.iStar)[0][:8]} (k={numSamples})"
            scale = max(abs(ref.alpha), float(np.abs(ref.beta[1:]).max()), 1e-300)
            assert abs(c.alpha - ref.alpha) <= self.rtol * max(abs(ref.alpha), 1e-300), (c.alpha, ref.alpha)
            assert np.abs(c.beta - ref.beta).max() <= self.rtol * scale
            # the reference build only exposes cummOld / cummAll as their ratio (cuts.c:172): compare the ratio
            with np.errstate(divide="ignore", invalid="ignore"):
                ra, rb = np.float64(c.cummOld) / np.float64(c.cummAll), np.float64(ref.cummOld) / np.float64(ref.cummAll)
            assert (np.isnan(ra) and np.isnan(rb)) or abs(ra - rb) <= self.rtol * max(abs(rb), 1e-300), (ra, rb)
        self.checked += 1
        return ref


# ----------------------------------------------------------------------------------------------------------------
# the SD loop
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class RunStats:
    iterations: int = 0
    seconds: float = 0.0
    argmax_seconds: float = 0.0        # time inside the table / cut library (the reference's "Argmax time", twoSD.h:93)
    subprob_seconds: float = 0.0
    master_seconds: float = 0.0
    lp_solves: int = 0
    incumbent_changes: int = 0
    incumb_est: float = 0.0
    candid_est: float = 0.0
    lower_bound_gap: float = 0.0
    dual_stable: bool = False
    incumbX: np.ndarray | None = None
    history: list = field(default_factory=list)


class SDHost:
    def __init__(self, slp: SyntheticSLP, tables, cfg: Config | None = None, seed: int = 1, check_lp_identity: bool = True):
        self.slp, self.t, self.cfg = slp, tables, cfg or Config()
        self.rng = np.random.default_rng(seed)
        self.sub, self.master = Subproblem(slp), Master(slp)
        self.check = check_lp_identity
        n1 = slp.n1
        self.maxCuts = self.cfg.CUT_MULT * n1 + 3                               # setup.c:126
        self.cuts: list[OneCut] = []
        self.k = 0
        self.candidX = np.zeros(n1 + 1)                                         # 1-based like the reference
        self.incumbX = np.zeros(n1 + 1)
        self.candidEst = self.incumbEst = slp.lb + float(slp.c @ self.candidX[1:])
        self.quad = self.cfg.MIN_QUAD_SCALAR
        self.iCutIdx, self.iCutUpdt, self.incumbChg = 0, 0, True
        self.gamma = self.normDk_1 = self.normDk = 0.0
        self.piM = np.zeros(0)
        self.pi_ratio = np.zeros(self.cfg.SCAN_LEN)
        self.dualStable = False
        self.obs_store: list[np.ndarray] = []                                   # host copy of omega->vals (subprob.c:24 reads it)
        self.basis_keys: dict = {}                                              # basis code -> basis index (stocUpdate.c:39-53)
        if slp.rvd:
            self.t.set_cost_coords(*slp.cost_coords())
        self.stats = RunStats()

    # ---- cuts.c:22-89 ------------------------------------------------------------------------------------------
    def form_sd_cut(self, x1, omegaIdx, newOmegaFlag, ctype):
        st, cfg, slp = self.stats, self.cfg, self.slp
        t0 = time.perf_counter()
        obj, pi = self.sub.solve(x1[1:], self.obs_store[omegaIdx][1:])
        st.subprob_seconds += time.perf_counter() - t0
        st.lp_solves += 1
        t0 = time.perf_counter()
        if slp.rvd:
            self.random_cost_updates(omegaIdx, newOmegaFlag, pi)
        else:
            self.t.stochastic_updates(omegaIdx, newOmegaFlag, pi, 0.0, self.k, cfg.TOLERANCE)             # subprob.c:70
        pi_eval = bool(cfg.DUAL_STABILITY and self.k > cfg.PI_EVAL_START and self.k % cfg.PI_CYCLE == 0)   # cuts.c:112
        cut = self.t.sd_cut(x1, self.k, pi_eval, slp.lb)                                                   # cuts.c:56
        st.argmax_seconds += time.perf_counter() - t0
        if cut is None:
            raise RuntimeError("SDCut returned NULL")
        if self.check:
            # STOCH_CHECK invariants: the argmax value at the observation just solved is the LP objective,
            # and the basis just returned attains it
            ist = int(cut.iStar[omegaIdx])
            assert 0 <= ist < self.t.counts()["basis"]
        if pi_eval:                                                                                          # cuts.c:171-182
            self.dualStable = bool(_dual_stability(cut.cummOld, cut.cummAll, self.k, cfg.PI_EVAL_START, cfg.SCAN_LEN, self.pi_ratio))
        oc = OneCut(cut.alpha, cut.beta.copy(), self.k, cut.omegaCnt, cut.iStar, ctype)
        return self.add_cut_to_pool(oc, ctype), obj

    # ---- stocUpdate.c:14-133 with random costs: the solver-side extraction stays on the host, everything else is library calls ---
    def random_cost_updates(self, omegaIdx, newOmegaFlag, pi):
        cfg, slp, t = self.cfg, self.slp, self.t
        if newOmegaFlag:                                                         # stocUpdate.c:24-31
            t.calc_delta(True, omegaIdx)
            if t.counts()["basis"]:
                t.check_feasibility_obs(omegaIdx, cfg.TOLERANCE)
        wd = self.obs_store[omegaIdx][1 + slp.R + slp.Q:]
        info = basis_info(self.sub, wd, pi)
        if info["key"] in self.basis_keys:                                       # stocUpdate.c:39-53: basis met before
            return self.basis_keys[info["key"]], False
        bi, bnew = t.stochastic_updates(omegaIdx, False, info["piDet"], 0.0, self.k, cfg.TOLERANCE, True, info["phi"], info["omegaIdx"])
        if bnew:                                                                 # stocUpdate.c:119-127
            t.basis_set_feas_data(bi, info["piDet"], np.array(info["phi"]) if info["phi"] else None, info["gBar"],
                                  info["psi"].ravel() if info["phi"] else None, info["cstat"])
            flags = t.check_feasibility_basis(bi, cfg.TOLERANCE)
            assert flags[omegaIdx], "a basis must be dual feasible at the observation it is optimal for"
            self.basis_keys[info["key"]] = bi
        return bi, bnew

    # ---- cuts.c:616-661, 277-360 ----------------------------------------------------------------------------------
    def add_cut_to_pool(self, cut: OneCut, ctype):
        if ctype == CANDIDATE:
            if len(self.cuts) >= self.maxCuts:
                self.reduce_cuts()
            self.cuts.append(cut)
            return len(self.cuts) - 1
        if len(self.cuts) >= self.maxCuts:
            self.drop_cut(self.iCutIdx)
        if self.cuts:
            self.cuts[min(self.iCutIdx, len(self.cuts) - 1)].type = CANDIDATE
        self.cuts.append(cut)
        self.iCutIdx = len(self.cuts) - 1
        self.iCutUpdt = self.k
        return self.iCutIdx

    def reduce_cuts(self):
        cfg = self.cfg
        minObs, oldest = self.k, len(self.cuts)
        for idx, c in enumerate(self.cuts):
            if c.type == INCUMBENT:
                continue
            loose = c.rowNum >= 0 and c.rowNum < len(self.piM) and abs(self.piM[c.rowNum]) <= cfg.TOLERANCE
            if c.numSamples < minObs and loose:
                minObs, oldest = c.numSamples, idx
        if oldest == len(self.cuts):
            minH, oldest = cut_height(self.cuts[0], self.k, self.candidX, self.slp.lb), 0
            for idx in range(1, len(self.cuts)):
                if self.cuts[idx].type == INCUMBENT:
                    continue
                hgt = cut_height(self.cuts[idx], self.k, self.candidX, self.slp.lb)
                if hgt < minH:
                    minH, oldest = hgt, idx
        self.drop_cut(oldest)

    def drop_cut(self, idx):
        last = self.cuts.pop()
        if idx < len(self.cuts):
            self.cuts[idx] = last
        if self.iCutIdx == len(self.cuts):
            self.iCutIdx = idx

    # ---- soln.c:24-95 ------------------------------------------------------------------------------------------------
    def check_improvement(self, candidCut):
        cfg, slp = self.cfg, self.slp
        candidEst = float(slp.c @ self.candidX[1:]) + max_cut_height(self.cuts, self.k, self.candidX, slp.lb)
        self.incumbEst = float(slp.c @ self.incumbX[1:]) + max_cut_height(self.cuts, self.k, self.incumbX, slp.lb)
        if (candidEst - self.incumbEst) < cfg.R1 * self.gamma:
            self.incumbX = self.candidX.copy()
            self.incumbEst = candidEst
            if self.normDk > cfg.TOLERANCE and self.normDk >= cfg.R3 * self.normDk_1:
                self.quad *= cfg.R2 * cfg.R3 * self.normDk_1 / self.normDk
                self.quad = max(cfg.MIN_QUAD_SCALAR, min(cfg.MAX_QUAD_SCALAR, self.quad))
            self.iCutUpdt, self.incumbChg = self.k, True
            self.normDk_1 = self.normDk
            self.gamma = 0.0
            self.cuts[self.iCutIdx].type = CANDIDATE
            self.cuts[candidCut].type = INCUMBENT
            self.iCutIdx = candidCut
            self.incumbChg = False
            self.stats.incumbent_changes += 1
        else:
            self.quad = min(cfg.MAX_QUAD_SCALAR, self.quad / cfg.R2)
            self.normDk_1 = self.normDk

    # ---- master.c:18-88 ----------------------------------------------------------------------------------------------
    def solve_master(self):
        slp = self.slp
        t0 = time.perf_counter()
        x, dual = self.master.solve(self.cuts, self.k, self.incumbX[1:], self.quad, slp.lb)
        self.stats.master_seconds += time.perf_counter() - t0
        d2 = float(np.sum((x - self.incumbX[1:]) ** 2))
        self.candidX = np.concatenate([[0.0], x])
        if self.k == 1:
            self.normDk_1 = d2
        self.normDk = d2
        self.piM = dual
        self.candidEst = float(slp.c @ x) + max_cut_height(self.cuts, self.k, self.candidX, slp.lb)
        self.gamma = self.candidEst - self.incumbEst

    # ---- algo.c:127-183 -----------------------------------------------------------------------------------------------
    def iterate(self):
        cfg = self.cfg
        self.k += 1
        observ = self.slp.sample(self.rng)
        t0 = time.perf_counter()
        omegaIdx, newOmega = self.t.calc_omega(observ, cfg.TOLERANCE)                  # algo.c:152
        self.stats.argmax_seconds += time.perf_counter() - t0
        if newOmega:
            assert omegaIdx == len(self.obs_store)
            self.obs_store.append(observ)
        candCut, _ = self.form_sd_cut(self.candidX, omegaIdx, newOmega, CANDIDATE)      # algo.c:155
        if (self.k - self.iCutUpdt) % cfg.TAU == 0:                                     # algo.c:161
            self.form_sd_cut(self.incumbX, omegaIdx, False, INCUMBENT)
        if not self.incumbChg and self.k > 1:                                           # algo.c:167
            self.check_improvement(candCut)
        self.incumbChg = False if self.k > 1 else self.incumbChg
        self.solve_master()                                                             # algo.c:174
        self.stats.history.append((self.k, self.candidEst, self.incumbEst, self.quad, len(self.cuts)))

    def run(self, iterations: int) -> RunStats:
        t0 = time.perf_counter()
        for _ in range(iterations):
            self.iterate()
        st = self.stats
        st.seconds += time.perf_counter() - t0
        st.iterations = self.k
        st.incumb_est, st.candid_est, st.dual_stable = self.incumbEst, self.candidEst, self.dualStable
        st.incumbX = self.incumbX.copy()
        return st


def _dual_stability(cummOld, cummAll, numSamples, piEvalStart, scanLen, pi_ratio):
    """cuts.c:171-182 + calcVariance cuts.c:366-396 (host scalar arithmetic)"""
    with np.errstate(divide="ignore", invalid="ignore"):
        pi_ratio[numSamples % scanLen] = np.float64(cummOld) / np.float64(cummAll)
    if numSamples - piEvalStart > scanLen:
        mean, vari = pi_ratio[0], 0.0
        for count in range(1, scanLen):
            temp = mean
            mean = mean + (pi_ratio[count] - mean) / (count + 1)
            vari = (1 - 1 / count) * vari + (count + 1) * (mean - temp) * (mean - temp)
        variance = vari
    else:
        variance = 1.0
    return not (abs(variance) >= .000002 or pi_ratio[numSamples % scanLen] < 0.95)


def caps_for(iterations: int, tau: int = 2) -> Caps:
    n = iterations + iterations // tau + 2                                              # setup.c:139
    return Caps(n, n, 2 * iterations + 2, iterations + 1, 1)
