"""Long LP-driven lock-step runs (missing #3): whole SD runs of the HiGHS host (tools/sd_highs_host.py) far beyond one
512-observation tile, CUDA library and CPU port fed identical calls, automatic kernel selection.  Real LP duals are where exact
ties live (degenerate vertices, repeated sigma rows): the strict '>' of stocUpdate.c:178-181 and the old-beats-new rule of
cuts.c:125 are exercised on them at every cut -- iStar must be equal at EVERY cut, cut coefficients within 1e-9, the final
incumbent within 1e-9.

  ssn shape, K = 1 500       N ~ 1 500 observations = 3 tiles; crosses the fused-prologue boundary (tiles x chunks > 148, ~k = 1 024)
  20term + random T, K = 800 crosses the LDG -> bulk-copy ring threshold of the random-T sweep (pairs x (1+Q) >= 4M, ~k = 550)
  storm shape + random cost, two replications of K = 600 through sdgpu_reset (setup.c:242-246): multi-term bases, obsFeasible mask,
                             device-evaluated checkBasisFeasibility; the second replication forces the term-linear ring

The CPU side is the port with OpenMP over observations (same arithmetic per observation as the sequential port, which
tests/test_sd_end_to_end.py pins bit for bit to the reference build on the same kind of run at smaller K).  Marked `slow`."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

import oracle_loader  # noqa: E402
from sd_highs_host import Lockstep, SDHost, caps_for, make_slp  # noqa: E402
from stochasticdecomposition_b200._abi import Caps  # noqa: E402


class OmpPort:
    """the CPU port with its OpenMP cut (sdo_sd_cut_omp): bit-identical iStar, same sums per observation"""

    def __init__(self, t):
        self._t = t

    def sd_cut(self, X, numSamples, pi_eval_flag, lb, want_istar=True, variant="sd_cut"):
        return self._t.sd_cut(X, numSamples, pi_eval_flag, lb, want_istar=want_istar, variant="sd_cut_omp")

    def __getattr__(self, name):
        return getattr(self._t, name)


class Recorder:
    """notes which sweep family and which prologue each GPU cut used, and the largest table seen"""

    def __init__(self, t):
        self._t, self.variants, self.launches, self.max_obs = t, set(), set(), 0

    def sd_cut(self, *a, **kw):
        cut = self._t.sd_cut(*a, **kw)
        st = self._t.stats()
        self.variants.add(int(st["last_sweep_variant"])); self.launches.add(int(st["last_cut_launches"]))
        self.max_obs = max(self.max_obs, self._t.counts()["omega"])
        return cut

    def __getattr__(self, name):
        return getattr(self._t, name)


def _caps(slp, K):
    if not slp.rvd:
        return caps_for(K)
    n = (1 + slp.rvd) * 2 * K + 8
    return Caps(n, n, 2 * K + 2, K + 1, 1 + slp.rvd)


def _run(shape, K, reps=1, force_second=None, seed=3, force_first=None):
    import stochasticdecomposition_b200 as sd
    slp = make_slp(shape)
    prob = slp.problem()
    gpu = Recorder(sd.load_library().create(prob, _caps(slp, K)))
    if force_first is not None:
        gpu.set_sweep_variant(force_first)
    cpu = OmpPort(oracle_loader.oracle().create(prob, _caps(slp, K)))
    tabs = Lockstep([gpu, cpu], rtol=1e-9)
    out = []
    for rep in range(reps):
        if rep > 0:
            tabs.reset()                                   # cleanCellType, setup.c:242-246
            assert tabs.counts() == {"omega": 0, "lambda": 0, "sigma": 0, "basis": 0}
            if force_second is not None:
                gpu.set_sweep_variant(force_second)
        host = SDHost(slp, tabs, seed=seed + 7 * rep)
        st = host.run(K)
        out.append((st, tabs.counts()))
    assert tabs.checked >= K * reps
    return gpu, out


pytestmark = [pytest.mark.gpu, pytest.mark.slow]


def test_ssn_1500_iterations_lockstep():
    gpu, out = _run("ssn", 1500)
    st, counts = out[0]
    assert counts["omega"] > 2 * 512 and gpu.max_obs > 2 * 512, counts                  # three observation tiles
    assert {2, 3} <= gpu.launches, gpu.launches                                          # fused prologue early (2 launches), separate later (3)
    assert st.iterations == 1500 and np.isfinite(st.incumb_est)


def test_20term_randomT_800_iterations_lockstep():
    gpu, out = _run("20term_T", 800)
    st, counts = out[0]
    assert counts["omega"] > 512, counts
    assert {1, 2} <= gpu.variants, gpu.variants                                          # load-based sweep first, the bulk-copy ring from ~4M elements


def test_storm_random_cost_two_replications_through_reset():
    gpu, out = _run("storm_rc", 600, reps=2, force_second=2)
    (st1, c1), (st2, c2) = out
    assert c1["omega"] > 512 and c2["omega"] > 512 and c1["sigma"] > c1["basis"]         # phi columns were stored (multi-term bases)
    assert {3, 4} <= gpu.variants, gpu.variants                                          # per-term gathers, then the term-linear ring (forced in replication 2)
    assert st1.iterations == st2.iterations == 600


def test_ssn_800_iterations_on_the_grouped_ring():
    """the grouped ring (bases walked in (lambda row, basis) order, lexicographic running maximum, four entries per filter, chunk merge
    by basis index) forced on LP duals: the ties are the LP's.  (The synthetic instances rarely store two sigmas on one lambda row, so
    the sharing itself is covered by tests/test_gpu_parity.py::test_grouped_sweep_shared_lambda_rows; here the point is the order-free
    walk and the incremental re-sort as one basis after the other arrives.)"""
    gpu, out = _run("ssn", 800, force_first=4)
    st, counts = out[0]
    assert gpu.variants == {6}, gpu.variants
    assert counts["omega"] > 512 and st.iterations == 800


def test_pgp2_1000_iterations_recompute_sweep():
    """pgp2 shape: 3 random right-hand sides with a small discrete support -- every observation is found again by calcOmega
    (weights >> 1, N stays at the support size) and the automatic choice is the sweep that recomputes delta.pib from the factors"""
    gpu, out = _run("pgp2", 1000)
    st, counts = out[0]
    assert gpu.variants == {5}, gpu.variants
    assert counts["omega"] <= 64 and counts["basis"] >= 4 and st.iterations == 1000, counts
