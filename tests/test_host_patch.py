"""The host patch EXECUTES: integration/twoSD_src.patch is applied to a scratch copy of the reference
with patch(1), the patched stocUpdate.c / cuts.c / optimal.c / randCost.c are compiled together with integration/sdgpu_hooks.c
(oracle/Makefile target `hooks`), and the patched `stochasticUpdates` / `SDCut` / `updtFeasCutPool` / `reformCuts` run in lock step
with the UNPATCHED reference functions (oracle/_ref/libsdref.so) on the same recorded solves.

The solver is a replay LP (oracle/shim/solver_cplex.h: getBasis / getDual / getPrimal / getDualSlacks / getBasisHead /
getBasisInvRow / getBasisInvACol answer from one recorded solve); the records come from a whole SD run of the HiGHS host
(tools/sd_highs_host.py).  Checked at every solve: observation index and flag, the basis index stochasticUpdates returns, the
newBasisFlag flow, NULL cuts where cuts.c:136-139 returns NULL, iStar, alpha, beta, pi_ratio and dualStableFlag; at the end the two
host basis lists index for index (ck, weight, phiLength, feasFlag, mubBar, sigmaIdx, omegaIdx, obsFeasible), the feasibility-cut
pool and a reformed cut; and the "Argmax time" the reference's clock() bracketing (subprob.c:68-73, cuts.c:55-62) accumulates is
printed in the format of inout.c:57.

  CPU flavour: the patched host linked against the restated oracle (bit-identical to the reference => everything must be EQUAL);
  GPU flavour (-m gpu): the patched host linked against libsdgpu.so (indices and iStar equal, cut coefficients within 1e-9)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import oracle_loader  # noqa: E402
from sd_highs_host import SDHost, SyntheticSLP, caps_for, make_slp  # noqa: E402
from stochasticdecomposition_b200._abi import CCounts, CCut, CProblem, _f64, _i32, _pf64, _pi32, c_f64p, c_i32p, c_u8p  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")
c_intp = C.POINTER(C.c_int)


class CReplayLP(C.Structure):
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("basisDim", C.c_int), ("cstat", c_intp), ("rstat", c_intp),
                ("x", c_f64p), ("dj", c_f64p), ("pi", c_f64p), ("head", c_intp), ("binvRows", c_f64p), ("binvACols", c_f64p)]


def _lp_struct(rec):
    keep = {k: np.ascontiguousarray(rec[k], dtype=(np.int32 if k in ("cstat", "rstat", "head") else np.float64))
            for k in ("cstat", "rstat", "x", "dj", "pi", "head", "binvRows", "binvACols") if rec.get(k) is not None}
    ip = lambda k: keep[k].ctypes.data_as(c_intp) if k in keep else None
    dp = lambda k: keep[k].ctypes.data_as(c_f64p) if k in keep else None
    lp = CReplayLP(rec["rows"], rec["cols"], rec["rows"], ip("cstat"), ip("rstat"), dp("x"), dp("dj"), dp("pi"), ip("head"), dp("binvRows"), dp("binvACols"))
    return lp, keep


# ---- recording a whole SD run --------------------------------------------------------------------------------------------------
class _LoggingTables:
    """passes everything through to the driving tables and logs the observation of every calcOmega call"""

    def __init__(self, t, events):
        self._t, self._events = t, events

    def calc_omega(self, observ, tol):
        self._events.append(("omega", np.array(observ, dtype=np.float64)))
        return self._t.calc_omega(observ, tol)

    def __getattr__(self, name):
        return getattr(self._t, name)


class RecordingHost(SDHost):
    def __init__(self, slp, tables, **kw):
        self.events = []
        super().__init__(slp, _LoggingTables(tables, self.events), **kw)
        solve = self.sub.solve

        def solve_and_snapshot(x, w):
            out = solve(x, w)
            self._snap = self._snapshot()
            return out
        self.sub.solve = solve_and_snapshot

    def _snapshot(self):
        h, slp = self.sub.h, self.slp
        sol, bas = h.getSolution(), h.getBasis()
        rec = dict(rows=slp.rows, cols=slp.cols, cstat=[int(v) for v in bas.col_status], rstat=[int(v) for v in bas.row_status],
                   x=np.asarray(sol.col_value), dj=np.asarray(sol.col_dual), pi=np.asarray(sol.row_dual), head=None, binvRows=None, binvACols=None)
        if slp.rvd:
            head = np.asarray(h.getBasicVariables()[1])
            rec["head"] = head
            inv = np.zeros((slp.rows, slp.rows))
            for col in slp.d_cols:                              # only the rows calcBasis asks for (randCost.c:36-47)
                pos = np.nonzero(head == col)[0]
                if len(pos):
                    inv[pos[0]] = h.getBasisInverseRow(int(pos[0]))[1]
            rec["binvRows"] = inv
            rec["binvACols"] = np.stack([h.getReducedColumn(i)[1] for i in range(slp.cols)])
        return rec

    def form_sd_cut(self, x1, omegaIdx, newOmegaFlag, ctype):
        out = super().form_sd_cut(x1, omegaIdx, newOmegaFlag, ctype)
        self.events.append(("solve", dict(x=np.array(x1), omegaIdx=omegaIdx, newOmega=bool(newOmegaFlag), k=self.k, feas=True, lp=self._snap)))
        return out


def record(shape, K, seed=3):
    slp = make_slp(shape)
    from test_sd_end_to_end import _caps
    host = RecordingHost(slp, oracle_loader.oracle().create(slp.problem(), _caps(slp, K)), seed=seed)
    host.run(K)
    return slp, host.events


# ---- the two hosts ---------------------------------------------------------------------------------------------------------------
class Host:
    """the same flat calls on the unpatched reference build (prefix sdref_) or on the patched host (prefix sdhk_)"""

    def __init__(self, lib, prefix, slp, max_iter, tau=2, scan_len=64):
        self.lib, self.pre, self.slp, self.scan_len = lib, prefix, slp, scan_len
        self.prob = slp.problem()
        self.cp = self.prob.to_c()
        self.n1 = self.prob.prevCols
        rvdOmCols, senx = slp.cost_coords()
        rvdOmCols = _i32(rvdOmCols)
        dcol, dval = _i32(np.arange(0, slp.cols + 1)), _f64(np.concatenate([[0.0], slp.d]))
        self.ctx = C.c_void_p()
        self.max_iter = max_iter
        if prefix == "sdref_":
            from stochasticdecomposition_b200._abi import Caps
            length = (slp.rvd if slp.rvd else 1) * max_iter + max_iter // tau + 1                     # setup.c:136-139
            caps = Caps(length, length, 2 * max_iter + 1, max_iter, 1 + slp.rvd).to_c()
            assert self._f("create")(C.byref(self.cp), C.byref(caps), 0, C.byref(self.ctx)) == 0
            assert self._f("set_cost_coords")(self.ctx, _pi32(rvdOmCols), senx) == 0
            assert self._f("set_dbar")(self.ctx, slp.cols, _pi32(dcol), _pf64(dval)) == 0
        else:
            self._f("last_error").restype = C.c_char_p
            st = self._f("create")(C.byref(self.cp), _pi32(rvdOmCols), senx, slp.cols, _pi32(dcol), _pf64(dval), max_iter, tau, 0, C.byref(self.ctx))
            assert st == 0, self._f("last_error")()
        self.pi_ratio = np.zeros(scan_len)
        self.stable = C.c_int(0)
        if prefix == "sdhk_":
            self._f("end_iteration").restype = C.c_double

    def _f(self, name):
        return getattr(self.lib, self.pre + name)

    def close(self):
        self._f("destroy").restype = None
        self._f("destroy")(self.ctx)

    def calc_omega(self, observ, tol=1e-3):
        o, flag = _f64(observ), C.c_int(0)
        return self._f("calc_omega")(self.ctx, _pf64(o), C.c_double(tol), C.byref(flag)), bool(flag.value)

    def stochastic_updates(self, ev, tol=1e-3):
        lp, keep = _lp_struct(ev["lp"])
        nb = C.c_int(1)                               # the reference only ever clears the flag
        st = self._f("stochastic_updates")(self.ctx, C.byref(lp), ev["omegaIdx"], int(ev["newOmega"]), ev["k"], C.c_double(tol), int(ev["feas"]), C.byref(nb))
        return st, bool(nb.value)

    def sd_cut(self, X, k, lb=0.0, n_obs=0):
        x = _f64(X)
        beta, istar = np.zeros(self.n1 + 1), np.full(max(n_obs, 1), -7, np.int32)
        cut = CCut(0.0, _pf64(beta), _pi32(istar), 0, 0, 0.0, 0.0)
        st = self._f("sd_cut_cfg")(self.ctx, _pf64(x), k, 1, 0, 1, self.scan_len, C.c_double(lb), C.byref(cut), _pf64(self.pi_ratio), C.byref(self.stable))
        if st == -1:
            return None
        assert st == 0, st
        return cut.alpha, beta, istar[:cut.omegaCnt].copy(), self.pi_ratio.copy(), bool(self.stable.value)

    def counts(self):
        c = CCounts()
        assert self._f("get_counts")(self.ctx, C.byref(c)) == 0
        return {"omega": c.omega, "lambda": c.lambda_, "sigma": c.sigma, "basis": c.basis}

    def basis_info(self, b, n_obs):
        ck, w, pl, ff, mub = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_double(0.0)
        sig, om, feas = np.zeros(self.slp.rvd + 2, np.int32), np.zeros(self.slp.rvd + 2, np.int32), np.zeros(max(n_obs, 1), np.uint8)
        assert self._f("basis_info")(self.ctx, b, C.byref(ck), C.byref(w), C.byref(pl), C.byref(ff), C.byref(mub), _pi32(sig), _pi32(om), feas.ctypes.data_as(c_u8p)) == 0
        p = pl.value
        return (ck.value, w.value, p, ff.value, mub.value, sig[:p + 1].tolist(), om[1:p + 1].tolist(), feas[:n_obs].tolist())

    def feas_pool(self, tol=1e-3):
        fu = (C.c_int * 2)(0, 0)
        alpha, beta = np.zeros(4096), np.zeros((4096, self.n1 + 1))
        n = self._f("updt_feas_cut_pool")(self.ctx, fu, C.c_double(tol), 4096, _pf64(alpha), _pf64(beta))
        assert n >= 0
        return alpha[:n].copy(), beta[:n].copy()

    def reform_cut(self, istar, observ, k):
        a, beta = C.c_double(0.0), np.zeros(self.n1 + 1)
        ist, ob = _i32(istar), _i32(observ)
        assert self._f("reform_cut")(self.ctx, _pi32(ist), len(ist), _pi32(ob), k, 0, 0, C.byref(a), _pf64(beta)) == 0
        return a.value, beta


def _libs(flavour):
    ref = os.path.join(REFDIR, "libsdref.so")
    host = os.path.join(REFDIR, f"libsdhost_{flavour}.so")
    if os.path.isdir(oracle_loader.REFERENCE_SRC):
        oracle_loader.build()
    if not (os.path.exists(ref) and os.path.exists(host)):
        pytest.skip("the reference build / patched host were not built here (no /root/reference) and did not travel")
    if flavour == "gpu":
        import stochasticdecomposition_b200 as sd
        sd.load_library()                          # fail loudly if libsdgpu.so is missing
    return C.CDLL(ref), C.CDLL(host)


def _lockstep(flavour, slp, events, max_iter, rtol, capsys=None):
    ref_lib, host_lib = _libs(flavour)
    ref, pat = Host(ref_lib, "sdref_", slp, max_iter), Host(host_lib, "sdhk_", slp, max_iter)
    cuts = nulls = 0
    last = None
    for kind, ev in events:
        if kind == "omega":
            a, b = ref.calc_omega(ev), pat.calc_omega(ev)
            assert a == b and a[0] >= 0, (a, b)
            continue
        a, b = ref.stochastic_updates(ev), pat.stochastic_updates(ev)
        assert a == b and a[0] >= 0, f"stochasticUpdates: reference {a}, patched host {b} (k={ev['k']})"
        n_obs = ref.counts()["omega"]
        ca, cb = ref.sd_cut(ev["x"], ev["k"], n_obs=n_obs), pat.sd_cut(ev["x"], ev["k"], n_obs=n_obs)
        assert (ca is None) == (cb is None), "NULL cut on one side only (cuts.c:136-139)"
        if ca is None:
            nulls += 1
            continue
        assert np.array_equal(ca[2], cb[2]), f"iStar differs at k={ev['k']}"
        scale = max(abs(ca[0]), float(np.abs(ca[1][1:]).max()), 1e-300)
        assert abs(ca[0] - cb[0]) <= rtol * max(abs(ca[0]), 1e-300) and np.abs(ca[1] - cb[1]).max() <= rtol * scale, (ca[0], cb[0])
        assert ca[1][0] == cb[1][0] == 1.0
        pa, pb = ca[3][ev["k"] % ref.scan_len], cb[3][ev["k"] % ref.scan_len]
        assert (np.isnan(pa) and np.isnan(pb)) or abs(pa - pb) <= max(rtol, 1e-15) * max(abs(pa), 1e-300), (pa, pb)
        assert ca[4] == cb[4], "dualStableFlag differs"
        pat._f("end_iteration")(pat.ctx)
        cuts += 1
        last = (ca[2], ev["k"])
    cr, cp_ = ref.counts(), pat.counts()
    assert cr == cp_, (cr, cp_)
    for b in range(cr["basis"]):                                         # the host lists, index for index
        ia, ib = ref.basis_info(b, cr["omega"]), pat.basis_info(b, cr["omega"])
        assert ia == ib, (b, ia, ib)
    fa, fb = ref.feas_pool(), pat.feas_pool()
    assert len(fa[0]) == len(fb[0])
    if len(fa[0]):
        assert np.abs(fa[0] - fb[0]).max() <= rtol * max(np.abs(fa[0]).max(), 1e-300) and np.abs(fa[1] - fb[1]).max() <= rtol * max(np.abs(fa[1]).max(), 1e-300)
    if last is not None:
        rng = np.random.default_rng(5)
        observ = rng.integers(0, cr["omega"], last[1]).astype(np.int32)
        ra, rb = ref.reform_cut(last[0], observ, last[1]), pat.reform_cut(last[0], observ, last[1])
        assert abs(ra[0] - rb[0]) <= rtol * max(abs(ra[0]), 1e-300) and np.abs(ra[1] - rb[1]).max() <= rtol * max(np.abs(ra[1]).max(), 1e-300)
    pat._f("print_summary")(pat.ctx)
    ref.close(); pat.close()
    return cuts, nulls, cr


SCENARIOS = [("pgp2", 150), ("20term_T", 40), ("randcost_small", 90)]


def _with_infeasible_tail(events):
    """one more solve, reported infeasible (subprob.c:46-52): its dual ray is stored with feasFlag = false (stocUpdate.c:66-75,128-129)
    and feeds the feasibility-cut pool.  It has to be the LAST solve: the reference's own dedup loop dereferences obsFeasible[b]
    of an infeasible basis."""
    kind, ev = [e for e in events if e[0] == "solve"][-1]
    lp = dict(ev["lp"])
    lp["pi"] = np.asarray(lp["pi"]) * 1.25 + 0.01
    lp["cstat"] = list(reversed(lp["cstat"]))
    tail = dict(ev, lp=lp, feas=False, newOmega=False)
    return list(events) + [("solve", tail)]


@pytest.mark.parametrize("shape,K", SCENARIOS)
def test_patched_host_matches_reference_cpu(shape, K, capfd):
    slp, events = record(shape, K)
    if not slp.rvd:
        events = _with_infeasible_tail(events)
    cuts, nulls, counts = _lockstep("cpu", slp, events, max_iter=K + 1, rtol=0.0)
    assert cuts >= K and nulls == 0 and counts["basis"] >= 2
    out = capfd.readouterr().out
    assert "Total time for argmax operation    : " in out and f"Number of unique observations      : {counts['omega']}" in out


def test_null_cut_where_the_reference_returns_null_cpu():
    """first solve of a run reported infeasible: the only stored basis is infeasible, no observation has a maximiser, SDCut returns
    NULL (cuts.c:136-139) on both sides"""
    slp, events = record("pgp2", 2)
    first = [("omega", events[0][1]), ("solve", dict(events[1][1], feas=False))]
    cuts, nulls, counts = _lockstep("cpu", slp, first, max_iter=8, rtol=0.0)
    assert cuts == 0 and nulls == 1 and counts["basis"] == 1


@pytest.mark.gpu
@pytest.mark.parametrize("shape,K", SCENARIOS + [("ssn", 60)])
def test_patched_host_matches_reference_gpu(shape, K, capfd):
    slp, events = record(shape, K)
    if not slp.rvd:
        events = _with_infeasible_tail(events)
    cuts, nulls, counts = _lockstep("gpu", slp, events, max_iter=K + 1, rtol=1e-9)
    assert cuts >= K and nulls == 0
    assert "Total time for argmax operation    : " in capfd.readouterr().out


@pytest.mark.gpu
def test_null_cut_where_the_reference_returns_null_gpu():
    slp, events = record("pgp2", 2)
    first = [("omega", events[0][1]), ("solve", dict(events[1][1], feas=False))]
    cuts, nulls, counts = _lockstep("gpu", slp, first, max_iter=8, rtol=1e-9)
    assert cuts == 0 and nulls == 1
