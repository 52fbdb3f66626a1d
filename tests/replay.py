"""Replays a synthetic SD trace (stochasticdecomposition_b200.synthetic.Trace) through one bound library in
the order the reference's main loop touches the tables (algo.c:127-183 -> cuts.c:22-89 -> subprob.c:70 ->
stocUpdate.c:14-133 -> cuts.c:91-194) and records everything comparable."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from stochasticdecomposition_b200._abi import Api, Caps, Problem
from stochasticdecomposition_b200.synthetic import Trace


def feas_mask(b: int, o: int, density: float = 0.85) -> bool:
    """Deterministic stand-in for checkBasisFeasibility (randCost.c:202) in random-cost traces."""
    h = (b * 2654435761 + o * 40503 + 12345) & 0xFFFFFFFF
    h ^= h >> 15
    h = (h * 2246822519) & 0xFFFFFFFF
    h ^= h >> 13
    return (h % 1000) < int(density * 1000)


class ForcedVariant:
    """A bound library whose new contexts are switched to one sweep family (sdgpu_set_sweep_variant: 1 loads, 2 TMA rings,
    3 recompute, 4 grouped ring; a family the problem's shape rules out falls back to the load-based kernels)."""

    def __init__(self, api, variant):
        self._api, self._variant = api, variant

    def create(self, *args, **kwargs):
        t = self._api.create(*args, **kwargs)
        t.set_sweep_variant(self._variant)
        return t

    def __getattr__(self, name):
        return getattr(self._api, name)


@dataclass
class Record:
    omega_idx: list = field(default_factory=list)
    omega_new: list = field(default_factory=list)
    basis_idx: list = field(default_factory=list)
    basis_new: list = field(default_factory=list)
    cuts: list = field(default_factory=list)          # Cut or None, one per solve
    counts: dict = field(default_factory=dict)
    tables: object = None


def pi_eval_flag(k: int, dual_stability=1, pi_eval_start=0, pi_cycle=1) -> bool:
    return bool(dual_stability and k > pi_eval_start and k % pi_cycle == 0)     # cuts.c:112


def replay(api: Api, problem: Problem, trace: Trace, caps: Caps, tol=1e-3, lb=0.0, dual_stability=1,
           pi_eval_start=0, pi_cycle=1, infeasible_every=0, cut_every=1, device=0, sd_cut_variant="sd_cut",
           feas_density=0.85, sweep_variant=None) -> Record:
    t = api.create(problem, caps, device)
    if sweep_variant is not None and api.has("set_sweep_variant"):
        t.set_sweep_variant(sweep_variant)
    rec = Record(tables=t)
    rvd = problem.rvdOmCnt
    K = trace.observ.shape[0]
    for it in range(K):
        k = it + 1
        oi, onew = t.calc_omega(trace.observ[it], tol)                       # algo.c:152
        rec.omega_idx.append(oi); rec.omega_new.append(onew)
        solves = (0, 1) if trace.two_solves[it] else (0,)                    # algo.c:155,161
        for sv in solves:
            feas = not (infeasible_every and (k * 2 + sv) % infeasible_every == 0)
            phi = () if trace.phi is None else trace.phi[it, sv]
            phio = () if trace.phi is None else trace.phi_omega[it, sv]
            if onew and rvd:                                                 # stocUpdate.c:24-31
                t.calc_delta(True, oi)
                nb = t.counts()["basis"]
                if nb:
                    t.basis_set_obs_feasible_col(oi, [feas_mask(b, oi, feas_density) for b in range(nb)])
                bi, bnew = t.stochastic_updates(oi, False, trace.duals[it, sv], trace.mubBar[it, sv], k, tol, feas, phi, phio)
            else:
                bi, bnew = t.stochastic_updates(oi, onew, trace.duals[it, sv], trace.mubBar[it, sv], k, tol, feas, phi, phio)
            if bnew and feas and rvd:                                        # stocUpdate.c:119-127
                t.basis_set_obs_feasible_row(bi, [feas_mask(bi, o, feas_density) for o in range(t.counts()["omega"])])
            onew = False                                                     # subprob.c:72
            rec.basis_idx.append(bi); rec.basis_new.append(bnew)
            if k % cut_every == 0:
                rec.cuts.append(t.sd_cut(trace.xs[it, sv], k, pi_eval_flag(k, dual_stability, pi_eval_start, pi_cycle), lb,
                                         variant=sd_cut_variant))
    rec.counts = t.counts()
    return rec


def dump_tables(t) -> dict:
    c = t.counts()
    out = {"omega": [], "lambda": [], "sigma": [], "delta_pib": None, "delta_piC": None}
    for o in range(c["omega"]):
        v, w = t.get_omega(o)
        out["omega"].append((v[1:].copy(), w))
    for l in range(c["lambda"]):
        out["lambda"].append(t.get_lambda(l)[1:].copy())
    for s in range(c["sigma"]):
        pib, piC, li, ck = t.get_sigma(s)
        out["sigma"].append((pib, piC[1:].copy(), li, ck))
    Q = t.problem.rvCOmCnt
    dp = np.zeros((c["lambda"], c["omega"]))
    dc = np.zeros((c["lambda"], c["omega"], Q))
    for l in range(c["lambda"]):
        for o in range(c["omega"]):
            pib, piC = t.get_delta(l, o)
            dp[l, o] = pib
            dc[l, o] = piC[1:]
    out["delta_pib"], out["delta_piC"] = dp, dc
    return out


def same_bits(a, b) -> bool:
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool(np.array_equal(a.view(np.int64), b.view(np.int64)))


def assert_tables_identical(ta, tb):
    a, b = dump_tables(ta), dump_tables(tb)
    assert len(a["omega"]) == len(b["omega"]) and len(a["lambda"]) == len(b["lambda"]) and len(a["sigma"]) == len(b["sigma"])
    for (va, wa), (vb, wb) in zip(a["omega"], b["omega"]):
        assert same_bits(va, vb) and wa == wb
    for va, vb in zip(a["lambda"], b["lambda"]):
        assert same_bits(va, vb)
    for (pa, ca, la, ka), (pb, cb, lb_, kb) in zip(a["sigma"], b["sigma"]):
        assert same_bits(pa, pb) and same_bits(ca, cb) and la == lb_ and ka == kb
    assert same_bits(a["delta_pib"], b["delta_pib"]), "delta.pib differs"
    assert same_bits(a["delta_piC"], b["delta_piC"]), "delta.piC differs"


def assert_records_match(ra: Record, rb: Record, exact_cut: bool, rtol=1e-9, ratio_only=False):
    """Indices and flags always exact.  exact_cut: alpha/beta bit-identical (CPU vs CPU, same summation order);
    otherwise within rtol of the largest coefficient magnitude (BASELINE.json: 1e-9 relative)."""
    assert ra.omega_idx == rb.omega_idx and ra.omega_new == rb.omega_new
    assert ra.basis_idx == rb.basis_idx and ra.basis_new == rb.basis_new
    assert ra.counts == rb.counts
    assert len(ra.cuts) == len(rb.cuts)
    for n, (ca, cb) in enumerate(zip(ra.cuts, rb.cuts)):
        assert (ca is None) == (cb is None), f"cut {n}: NULL-ness differs"
        if ca is None:
            continue
        assert ca.omegaCnt == cb.omegaCnt and ca.numSamples == cb.numSamples
        assert np.array_equal(ca.iStar, cb.iStar), f"cut {n}: iStar differs at {np.nonzero(ca.iStar != cb.iStar)[0][:8]}"
        if exact_cut:
            assert same_bits(ca.alpha, cb.alpha) and same_bits(ca.beta, cb.beta), f"cut {n}: coefficients differ"
        else:
            scale = max(abs(ca.alpha), float(np.max(np.abs(ca.beta[1:]))) if len(ca.beta) > 1 else 0.0, 1e-300)
            assert abs(ca.alpha - cb.alpha) <= rtol * max(abs(ca.alpha), 1e-300) + 0.0, f"cut {n}: alpha {ca.alpha} vs {cb.alpha}"
            assert np.max(np.abs(ca.beta - cb.beta)) <= rtol * scale, f"cut {n}: beta differs by {np.max(np.abs(ca.beta - cb.beta))}"
        ra_ratio = ca.cummOld / ca.cummAll if ca.cummAll != 0 else float("nan")
        rb_ratio = cb.cummOld / cb.cummAll if cb.cummAll != 0 else float("nan")
        if ratio_only:
            assert same_bits(ra_ratio, rb_ratio) or (np.isnan(ra_ratio) and np.isnan(rb_ratio)), f"cut {n}: pi_ratio differs"
        elif exact_cut:
            assert same_bits(ca.cummOld, cb.cummOld) and same_bits(ca.cummAll, cb.cummAll)
        else:
            assert abs(ca.cummOld - cb.cummOld) <= rtol * max(abs(ca.cummOld), 1e-300)
            assert abs(ca.cummAll - cb.cummAll) <= rtol * max(abs(ca.cummAll), 1e-300)
