"""Seeded synthetic two-stage-SLP shapes and SD traces.

The SMPS inputs of the reference's test problems (pgp2, 20term, ssn, storm) are not available offline, so
every run here uses synthetic data with those problems' *shapes*; the generator is deterministic in the
seed so the CPU checkers and the CUDA library see identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ._abi import Problem

# (rows, cols, n1, n1c, R, Rb, Q, rvdOmCnt) -- dimensions quoted in SURVEY.md section 8 (literature values)
SHAPES = {
    "pgp2": dict(rows=7, cols=16, n1=4, n1c=4, R=3, Rb=3, Q=0, rvd=0),
    "20term": dict(rows=124, cols=764, n1=63, n1c=63, R=40, Rb=40, Q=0, rvd=0),
    "20term_T": dict(rows=124, cols=764, n1=63, n1c=63, R=40, Rb=40, Q=8, rvd=0),
    "ssn": dict(rows=175, cols=706, n1=89, n1c=89, R=86, Rb=86, Q=0, rvd=0),
    "storm": dict(rows=528, cols=1259, n1=121, n1c=121, R=118, Rb=118, Q=0, rvd=4),
    "sweep256": dict(rows=256, cols=512, n1=89, n1c=89, R=256, Rb=256, Q=0, rvd=0),
}


def _one_based(a, dtype):
    return np.concatenate([np.zeros(1, dtype), np.asarray(a, dtype)])


def make_problem(seed: int, rows: int, cols: int, n1: int, n1c: int, R: int, Rb: int, Q: int = 0, rvd: int = 0,
                 distinct_rvCols: bool = False, shared_T_cols: bool = False) -> Problem:
    rng = np.random.default_rng(seed)
    CCols = np.sort(rng.choice(np.arange(1, n1 + 1), size=n1c, replace=False))
    rvRows = np.sort(rng.choice(np.arange(1, rows + 1), size=R, replace=False))
    rvbOmRows = np.sort(rng.choice(rvRows, size=Rb, replace=False))
    if Q:
        rvCOmRows = rng.choice(rvRows, size=Q, replace=True)
        if shared_T_cols:                      # several random T elements in the same column (vxMSparse accumulates)
            pool = rng.choice(np.arange(1, n1 + 1), size=max(1, Q // 2), replace=False)
            rvCOmCols = rng.choice(pool, size=Q, replace=True)
        else:
            rvCOmCols = rng.choice(np.arange(1, n1 + 1), size=Q, replace=False)
        order = np.lexsort((rvCOmRows, rvCOmCols))
        rvCOmRows, rvCOmCols = rvCOmRows[order], rvCOmCols[order]
        rvCols = rng.choice(np.arange(1, n1 + 1), size=Q, replace=False) if distinct_rvCols else rvCOmCols.copy()
    else:
        rvCOmRows = rvCOmCols = rvCols = np.zeros(0, np.int32)
    nb = max(1, rows // 2)
    bcol = np.sort(rng.choice(np.arange(1, rows + 1), size=nb, replace=False))
    bval = rng.normal(0.0, 1.0, nb)
    nnz = min(rows * n1c, max(n1c, 3 * rows))
    flat = rng.choice(rows * n1c, size=nnz, replace=False)      # unsorted: nnz order matters for vxMSparse
    crow = flat // n1c + 1
    ccol = CCols[flat % n1c]
    cval = rng.uniform(-1.0, 1.0, nnz)
    return Problem(
        rows=rows, cols=cols, prevCols=n1,
        CCols=_one_based(CCols, np.int32), rvRows=_one_based(rvRows, np.int32),
        rvbOmRows=_one_based(rvbOmRows, np.int32), rvCOmCols=_one_based(rvCOmCols, np.int32),
        rvCOmRows=_one_based(rvCOmRows, np.int32), rvCols=_one_based(rvCols, np.int32),
        bBar_col=_one_based(bcol, np.int32), bBar_val=_one_based(bval, np.float64),
        Cbar_col=_one_based(ccol, np.int32), Cbar_row=_one_based(crow, np.int32), Cbar_val=_one_based(cval, np.float64),
        rvdOmCnt=rvd, rvOffset=(0, Rb, Rb + Q))


def problem_for(name: str, seed: int = 20240607, **overrides) -> Problem:
    shape = dict(SHAPES[name])
    shape.update(overrides)
    return make_problem(seed, **shape)


@dataclass
class Trace:
    """A recorded SD run as the tables see it: one observation and one (or two) dual vertices per
    iteration, the x each cut is formed at, and the random-cost phi columns if any."""
    observ: np.ndarray      # [K][numRV+1]
    duals: np.ndarray       # [K][2][rows+1]   (candidate solve, incumbent solve)
    mubBar: np.ndarray      # [K][2]
    xs: np.ndarray          # [K][2][n1+1]
    two_solves: np.ndarray  # [K] bool: incumbent solve happens (algo.c:161)
    phi: np.ndarray | None = None       # [K][2][rvd][rows+1] or None
    phi_omega: np.ndarray | None = None  # [K][2][rvd] 1-based positions in the cost block


def make_trace(problem: Problem, K: int, seed: int, dual_pool: int = 0, obs_pool: int = 0, tol: float = 1e-3,
               tau: int = 2, phi_len: int = 0, obs_scale: float = 3.0) -> Trace:
    """dual_pool / obs_pool > 0 draw from a finite pool (with sub-tolerance jitter on some draws) so the dedup
    paths of calcOmega / calcLambda / calcSigma are exercised; 0 means every draw is fresh."""
    rng = np.random.default_rng(seed)
    rows, n1, nrv = problem.rows, problem.prevCols, problem.numRV

    def pool_or_fresh(shape_one, pool, scale, sparsify=0.0):
        def fresh():
            v = rng.uniform(-1.0, 1.0, shape_one) * scale
            if sparsify:
                v[rng.random(shape_one) < sparsify] = 0.0
            return v
        if pool <= 0:
            return lambda: fresh()
        store = [fresh() for _ in range(pool)]

        def draw():
            v = store[rng.integers(pool)].copy()
            r = rng.random()
            if r < 0.3:
                v += rng.uniform(-0.4, 0.4, shape_one) * tol        # within tolerance of the stored one
            elif r < 0.4:
                j = rng.integers(shape_one)
                v[j] += 1.5 * tol                                    # just outside tolerance in one coordinate
            return v
        return draw

    draw_pi = pool_or_fresh(rows + 1, dual_pool, 1.0, sparsify=0.3)
    draw_ob = pool_or_fresh(nrv + 1, obs_pool, obs_scale)
    observ = np.stack([draw_ob() for _ in range(K)])
    observ[:, 0] = 0.0
    duals = np.stack([np.stack([draw_pi(), draw_pi()]) for _ in range(K)])
    duals[:, :, 0] = 0.0
    mub = rng.normal(0.0, 1.0, (K, 2)) * (rng.random((K, 2)) < 0.5)
    xs = rng.uniform(0.0, 1.0, (K, 2, n1 + 1))
    xs[:, :, 0] = 0.0
    two = (np.arange(1, K + 1) % tau) == 0
    phi = phi_om = None
    if phi_len:
        phi = rng.uniform(-0.5, 0.5, (K, 2, phi_len, rows + 1)) * (rng.random((K, 2, phi_len, rows + 1)) < 0.3)
        phi[..., 0] = 0.0
        phi_om = np.stack([[np.sort(rng.choice(np.arange(1, problem.rvdOmCnt + 1), size=phi_len, replace=False))
                            for _ in range(2)] for _ in range(K)]).astype(np.int32)
    return Trace(observ, duals, mub, xs, two, phi, phi_om)
