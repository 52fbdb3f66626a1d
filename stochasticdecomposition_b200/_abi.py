"""ctypes view of the C ABI declared in include/sdgpu.h.

`Api(cdll, prefix)` binds one shared library that exports the sdgpu entry points under `prefix`
(`sdgpu_` for the product library).  The test suite binds its CPU checkers with the same class and a
different prefix, so a parity test is literally "same calls, compare outputs".  Nothing in this module
loads a library by itself.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

SDGPU_NONE = -1
SDGPU_ERR = -2

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)
c_intp = C.POINTER(C.c_int)


class CNum(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "rows", "cols", "prevCols", "cntCcols", "rvRowCnt", "rvbOmCnt", "rvCOmCnt", "rvdOmCnt", "numRV")]


class CCoord(C.Structure):
    _fields_ = [("CCols", c_i32p), ("rvRows", c_i32p), ("rvbOmRows", c_i32p), ("rvCOmCols", c_i32p),
                ("rvCOmRows", c_i32p), ("rvCols", c_i32p), ("rvOffset", C.c_int32 * 3)]


class CSparseVec(C.Structure):
    _fields_ = [("cnt", C.c_int32), ("col", c_i32p), ("val", c_f64p)]


class CSparseMat(C.Structure):
    _fields_ = [("cnt", C.c_int32), ("col", c_i32p), ("row", c_i32p), ("val", c_f64p)]


class CCaps(C.Structure):
    _fields_ = [("maxLambda", C.c_int64), ("maxSigma", C.c_int64), ("maxBasis", C.c_int64),
                ("maxOmega", C.c_int64), ("maxTerms", C.c_int32)]


class CProblem(C.Structure):
    _fields_ = [("num", CNum), ("coord", CCoord), ("bBar", CSparseVec), ("Cbar", CSparseMat)]


class CCounts(C.Structure):
    _fields_ = [("omega", C.c_int64), ("lambda_", C.c_int64), ("sigma", C.c_int64), ("basis", C.c_int64)]


class CCut(C.Structure):
    _fields_ = [("alpha", C.c_double), ("beta", c_f64p), ("iStar", c_i32p), ("omegaCnt", C.c_int32),
                ("numSamples", C.c_int32), ("cummOld", C.c_double), ("cummAll", C.c_double)]


class CStats(C.Structure):
    _fields_ = [("last_cut_ms", C.c_double), ("last_sweep_ms", C.c_double), ("last_cut_launches", C.c_int64),
                ("total_launches", C.c_int64), ("last_sweep_bytes", C.c_int64), ("last_sweep_variant", C.c_int64),
                ("last_prep_ms", C.c_double), ("last_merge_ms", C.c_double), ("last_collective_ms", C.c_double)]


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _pi32(a):
    return None if a is None else a.ctypes.data_as(c_i32p)


def _pf64(a):
    return None if a is None else a.ctypes.data_as(c_f64p)


@dataclass
class Problem:
    """The static description the reference passes around as numType / coordType / prob->bBar, prob->Cbar
    (SURVEY.md section 8b).  Every array is 1-based: element 0 is a dummy."""
    rows: int
    cols: int
    prevCols: int
    CCols: np.ndarray
    rvRows: np.ndarray
    rvbOmRows: np.ndarray
    rvCOmCols: np.ndarray
    rvCOmRows: np.ndarray
    rvCols: np.ndarray
    bBar_col: np.ndarray
    bBar_val: np.ndarray
    Cbar_col: np.ndarray
    Cbar_row: np.ndarray
    Cbar_val: np.ndarray
    rvdOmCnt: int = 0
    rvOffset: tuple = (0, 0, 0)
    _keep: list = field(default_factory=list, repr=False)

    @property
    def cntCcols(self) -> int: return len(self.CCols) - 1
    @property
    def rvRowCnt(self) -> int: return len(self.rvRows) - 1
    @property
    def rvbOmCnt(self) -> int: return len(self.rvbOmRows) - 1
    @property
    def rvCOmCnt(self) -> int: return len(self.rvCOmCols) - 1
    @property
    def numRV(self) -> int: return self.rvbOmCnt + self.rvCOmCnt + self.rvdOmCnt

    def to_c(self) -> CProblem:
        arrs = {k: _i32(getattr(self, k)) for k in
                ("CCols", "rvRows", "rvbOmRows", "rvCOmCols", "rvCOmRows", "rvCols", "bBar_col", "Cbar_col", "Cbar_row")}
        vals = {k: _f64(getattr(self, k)) for k in ("bBar_val", "Cbar_val")}
        self._keep = [arrs, vals]
        p = CProblem()
        p.num = CNum(self.rows, self.cols, self.prevCols, self.cntCcols, self.rvRowCnt, self.rvbOmCnt,
                     self.rvCOmCnt, self.rvdOmCnt, self.numRV)
        p.coord = CCoord(_pi32(arrs["CCols"]), _pi32(arrs["rvRows"]), _pi32(arrs["rvbOmRows"]),
                         _pi32(arrs["rvCOmCols"]), _pi32(arrs["rvCOmRows"]), _pi32(arrs["rvCols"]),
                         (C.c_int32 * 3)(*self.rvOffset))
        p.bBar = CSparseVec(len(arrs["bBar_col"]) - 1, _pi32(arrs["bBar_col"]), _pf64(vals["bBar_val"]))
        p.Cbar = CSparseMat(len(arrs["Cbar_col"]) - 1, _pi32(arrs["Cbar_col"]), _pi32(arrs["Cbar_row"]),
                            _pf64(vals["Cbar_val"]))
        return p


@dataclass
class Caps:
    """Capacity contract of setup.c:136-144."""
    maxLambda: int
    maxSigma: int
    maxBasis: int
    maxOmega: int
    maxTerms: int = 1

    @staticmethod
    def like_reference(max_iter: int, tau: int, rvdOmCnt: int = 0) -> "Caps":
        length = (rvdOmCnt if rvdOmCnt > 0 else 1) * max_iter + max_iter // tau + 1   # setup.c:136-139
        return Caps(length, length, max_iter, max_iter, 1 + rvdOmCnt)

    def to_c(self) -> CCaps:
        return CCaps(self.maxLambda, self.maxSigma, self.maxBasis, self.maxOmega, self.maxTerms)


@dataclass
class Cut:
    """oneCut of twoSD.h:69-80 (solver bookkeeping left to the host)."""
    alpha: float
    beta: np.ndarray          # [prevCols+1], beta[0] == 1.0
    iStar: np.ndarray | None  # [omegaCnt] int32
    omegaCnt: int
    numSamples: int
    cummOld: float
    cummAll: float


class SdError(RuntimeError):
    pass


class Api:
    """One bound library.  Methods mirror include/sdgpu.h one to one."""

    def __init__(self, cdll: C.CDLL, prefix: str):
        self.lib, self.prefix = cdll, prefix
        self._sig()

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name, None)

    def has(self, name) -> bool:
        return self._fn(name) is not None

    def _sig(self):
        vp, i, d, i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
        table = {
            "abi_version": (i, []),
            "create": (i, [C.POINTER(CProblem), C.POINTER(CCaps), i, C.POINTER(vp)]),
            "reset": (i, [vp]), "destroy": (None, [vp]), "last_error": (C.c_char_p, []),
            "get_counts": (i, [vp, C.POINTER(CCounts)]),
            "calc_omega": (i, [vp, c_f64p, d, c_intp]),
            "omega_find": (i, [vp, c_f64p, d]), "omega_append": (i, [vp, c_f64p, i]), "omega_bump": (i, [vp, i, i]),
            "omega_append_bulk": (i, [vp, i64, c_f64p, c_i32p]),
            "calc_lambda": (i, [vp, c_f64p, d, c_intp]),
            "calc_sigma": (i, [vp, c_f64p, d, i, i, i, d, c_intp]),
            "calc_delta": (i, [vp, i, i]),
            "calc_delta_block": (i, [vp, i64, i64, i64, i64]),
            "get_delta_block": (i, [vp, i64, i64, i64, i64, i, c_f64p]),
            "bulk_load": (i, [vp, i, c_f64p, c_i32p, i, c_f64p, c_f64p, c_i32p, d, c_i32p, c_i32p]),
            "update_dual": (i, [vp, c_f64p, d, i, d, c_intp, c_intp, c_intp, c_intp]),
            "update_dual_col": (i, [vp, i, c_f64p, d, i, d, c_intp, c_intp, c_intp, c_intp]),
            "update_dual_bulk": (i, [vp, i64, c_f64p, c_f64p, c_i32p, d, c_i32p, c_i32p]),
            "basis_append": (i, [vp, i, i, i, c_i32p, c_i32p]),
            "basis_append_bulk": (i, [vp, i64, c_i32p, c_i32p, c_i32p]),
            "basis_find_or_append": (i, [vp, i, i, i, i, i, c_i32p, c_i32p, c_intp]),
            "basis_set_obs_feasible": (i, [vp, i, i, i]),
            "basis_set_obs_feasible_row": (i, [vp, i, c_u8p]), "basis_set_obs_feasible_col": (i, [vp, i, c_u8p]),
            "set_cost_coords": (i, [vp, c_i32p, C.c_char_p]),
            "basis_set_feas_data": (i, [vp, i, c_f64p, c_f64p, c_f64p, c_f64p, c_i32p]),
            "check_feasibility_obs": (i, [vp, i, d, c_u8p]), "check_feasibility_basis": (i, [vp, i, d, c_u8p]),
            "feas_cuts": (i, [vp, i, i, i, i, i, c_f64p, c_f64p]),
            "updt_feas_cut_pool": (i, [vp, c_intp, d, i, c_f64p, c_f64p]),
            "feas_pool_update": (i, [vp, c_intp, d]), "feas_pool_size": (i, [vp]), "feas_pool_get": (i, [vp, i, i, c_f64p, c_f64p]),
            "feas_pool_check": (i, [vp, i, c_f64p, c_f64p, c_f64p, c_f64p, d, c_i32p, c_intp]),
            "compute_istar": (i, [vp, c_f64p, i, i, i, i, c_f64p]),
            "sd_cut": (i, [vp, c_f64p, i, i, d, C.POINTER(CCut)]),
            "sd_cut_omp": (i, [vp, c_f64p, i, i, d, C.POINTER(CCut), c_intp]),
            "sd_cut_cfg": (i, [vp, c_f64p, i, i, i, i, i, d, C.POINTER(CCut), c_f64p, c_intp]),
            "sd_cut_partial_host": (i, [vp, c_f64p, i, i, d, c_f64p, c_i32p]),
            "dual_stability": (i, [d, d, i, i, i, c_f64p]),
            "calc_variance": (d, [c_f64p, i]),
            "sd_cut_partial": (i, [vp, c_f64p, i, i, d]),
            "sd_cut_partial_buffer": (i, [vp, C.POINTER(vp), c_intp]),
            "sd_cut_finish": (i, [vp, i, C.POINTER(CCut)]),
            "peer_export": (i, [vp, i, vp]), "peer_attach": (i, [vp, i, i, vp]),
            "attach_nccl": (i, [vp, vp]), "nccl_unique_id": (i, [vp]), "nccl_init": (i, [vp, i, i, vp]),
            "cut_heights": (i, [vp, i, c_f64p, c_f64p, c_i32p, c_f64p, i, c_f64p, d, c_f64p, c_f64p, c_f64p]),
            "reform_cut": (i, [vp, c_i32p, i, c_i32p, i, i, i, c_f64p, c_f64p]),
            "reform_cuts_batch": (i, [vp, i, c_i32p, i, c_i32p, i, c_i32p, i, i, i, c_f64p, c_f64p]),
            "get_omega": (i, [vp, i, c_f64p, c_intp]), "get_lambda": (i, [vp, i, c_f64p]),
            "get_sigma": (i, [vp, i, c_f64p, c_f64p, c_intp, c_intp]), "get_delta": (i, [vp, i, i, c_f64p, c_f64p]),
            "group_create": (i, [C.POINTER(CProblem), C.POINTER(CCaps), i, c_intp, C.POINTER(vp)]),
            "group_destroy": (None, [vp]), "group_reset": (i, [vp]), "group_size": (i, [vp]), "group_member": (vp, [vp, i]),
            "group_get_counts": (i, [vp, C.POINTER(CCounts)]),
            "group_calc_omega": (i, [vp, c_f64p, d, c_intp]),
            "group_update_dual": (i, [vp, c_f64p, d, i, d, c_intp, c_intp, c_intp, c_intp]),
            "group_basis_find_or_append": (i, [vp, i, i, i, i, c_intp]),
            "group_sd_cut": (i, [vp, c_f64p, i, i, d, C.POINTER(CCut)]),
            "last_istar_device": (i, [vp, C.POINTER(vp), c_intp]),
            "get_stats": (i, [vp, C.POINTER(CStats)]),
            "set_sweep_variant": (i, [vp, i]), "set_timing": (i, [vp, i]), "set_stream": (i, [vp, vp]), "set_collective": (i, [vp, i]),
            "fp64_peak": (i, [i, i, c_f64p, c_f64p]), "launch_roundtrip": (i, [i, i, i, i, c_f64p]), "delta_memory": (i, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
            "plan_sweep_grid": (i, [i, i64, i64, i, c_intp, c_intp, c_intp]),
            "plan_sweep_kind": (i, [i, i, i, i, i, i, i, i64, i64, i64, i64, i, c_intp]),
        }
        for name, (res, args) in table.items():
            fn = self._fn(name)
            if fn is not None:
                fn.restype, fn.argtypes = res, args

    def plan_sweep_grid(self, sm_count: int, observations: int, bases: int, max_chunks: int = 64):
        """(tiles, chunkSize, nChunks) of the 2-D sweep grid for a table of this shape; host-only, no device needed."""
        t, cs, nc = C.c_int(0), C.c_int(0), C.c_int(0)
        if self._fn("plan_sweep_grid")(sm_count, observations, bases, max_chunks, C.byref(t), C.byref(cs), C.byref(nc)) != 0:
            raise SdError(self.error())
        return t.value, cs.value, nc.value

    def plan_sweep_kind(self, observations, bases, Q=0, rvd=0, Rb=10, max_phi=0, cost_cols=0, n1=10, n1c=10, terms=None,
                        distinct_rows=None, variant=0):
        """(sweep family id as in stats()["last_sweep_variant"], fused prologue?) for this shape; host-only."""
        fused = C.c_int(0)
        k = self._fn("plan_sweep_kind")(Q, rvd, Rb, max_phi, cost_cols, n1, n1c, bases, bases if terms is None else terms,
                                        bases if distinct_rows is None else distinct_rows, observations, variant, C.byref(fused))
        if k < 0:
            raise SdError(self.error())
        return k, bool(fused.value)

    def fp64_peak(self, device: int = 0, reps: int = 5):
        """(separately rounded DMUL+DADD operations / s, DFMA flops / s) of the device's FP64 pipe"""
        a, b = C.c_double(0.0), C.c_double(0.0)
        if self._fn("fp64_peak")(device, reps, C.byref(a), C.byref(b)) != 0:
            raise SdError(self.error())
        return a.value, b.value

    def launch_roundtrip(self, mode: int, launches: int = 1, reps: int = 200, device: int = 0) -> float:
        """median wall microseconds of `launches` empty kernels + a stream synchronise (mode 0) or a host spin on mapped memory (mode 1)"""
        a = C.c_double(0.0)
        if self._fn("launch_roundtrip")(device, mode, launches, reps, C.byref(a)) != 0:
            raise SdError(self.error())
        return a.value

    def error(self) -> str:
        fn = self._fn("last_error")
        msg = fn() if fn is not None else None
        return msg.decode() if msg else ""

    def create(self, problem: Problem, caps: Caps, device: int = 0) -> "Tables":
        ctx = C.c_void_p()
        cp, cc = problem.to_c(), caps.to_c()
        st = self._fn("create")(C.byref(cp), C.byref(cc), device, C.byref(ctx))
        if st != 0 or not ctx:
            raise SdError(f"{self.prefix}create failed ({st}): {self.error()}")
        return Tables(self, ctx, problem, caps)


class Group:
    """Several GPUs of this process behind one handle (include/sdgpu.h, sdgpu_group_*): observations dealt round-robin, dual-side
    tables replicated, the cut all-reduced through NVLink peer memory inside the cut kernel.  One host thread drives it."""

    def __init__(self, api: "Api", problem: Problem, capsPerDevice: Caps, devices):
        self.api, self.problem = api, problem
        h = C.c_void_p()
        cp, cc = problem.to_c(), capsPerDevice.to_c()
        devs = (C.c_int * len(devices))(*devices)
        st = api._fn("group_create")(C.byref(cp), C.byref(cc), len(devices), devs, C.byref(h))
        if st != 0 or not h:
            raise SdError(f"group_create failed ({st}): {api.error()}")
        self.h = h

    def _ok(self, st, what, allow_none=False):
        if st <= SDGPU_ERR or (st == SDGPU_NONE and not allow_none):
            raise SdError(f"group {what} failed ({st}): {self.api.error()}")
        return st

    def close(self):
        if self.h:
            self.api._fn("group_destroy")(self.h)
            self.h = None

    def reset(self):
        """cleanCellType (setup.c:242-246) for every member: counts to zero, memory and the peer exchange kept"""
        self._ok(self.api._fn("group_reset")(self.h), "reset")

    def counts(self):
        c = CCounts()
        self._ok(self.api._fn("group_get_counts")(self.h, C.byref(c)), "get_counts")
        return {"omega": c.omega, "lambda": c.lambda_, "sigma": c.sigma, "basis": c.basis}

    def calc_omega(self, observ, tol):
        o, flag = _f64(observ), C.c_int(0)
        idx = self._ok(self.api._fn("group_calc_omega")(self.h, _pf64(o), tol, C.byref(flag)), "calc_omega")
        return idx, bool(flag.value)

    def stochastic_updates(self, omegaIdx, newOmegaFlag, piDet, mubBar, currentIter, tol, feasFlag=True):
        p = _f64(piDet)
        li, nl, si, ns, nb = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        self._ok(self.api._fn("group_update_dual")(self.h, _pf64(p), mubBar, currentIter, tol, C.byref(li), C.byref(nl), C.byref(si),
                                                     C.byref(ns)), "update_dual")
        b = self._ok(self.api._fn("group_basis_find_or_append")(self.h, ns.value, currentIter, int(feasFlag), si.value, C.byref(nb)),
                     "basis_find_or_append")
        return b, bool(nb.value)

    def sd_cut(self, X, numSamples, pi_eval_flag, lb):
        x = _f64(X)
        n = self.counts()["omega"]
        beta, istar = np.zeros(self.problem.prevCols + 1), np.full(max(n, 1), -7, np.int32)
        cut = CCut(0.0, _pf64(beta), _pi32(istar), 0, 0, 0.0, 0.0)
        st = self.api._fn("group_sd_cut")(self.h, _pf64(x), numSamples, int(pi_eval_flag), lb, C.byref(cut))
        if st == SDGPU_NONE:
            return None
        self._ok(st, "sd_cut")
        return Cut(cut.alpha, beta, istar[:n], cut.omegaCnt, cut.numSamples, cut.cummOld, cut.cummAll)


class Tables:
    """A live table set (omega / lambda / sigma / delta / basis) behind one C context.  Method names are
    the reference's (calcOmega, calcLambda, calcSigma, calcDelta, computeIstar, SDCut, ...) in snake case."""

    def __init__(self, api: Api, ctx, problem: Problem, caps: Caps):
        self.api, self.ctx, self.problem, self.caps = api, ctx, problem, caps

    # -- plumbing ---------------------------------------------------------------------------------------
    def _call(self, name, *args):
        fn = self.api._fn(name)
        if fn is None:
            raise SdError(f"{self.api.prefix}{name} is not exported by this library")
        return fn(self.ctx, *args)

    def _check(self, st, what, allow_none=False):
        if st <= SDGPU_ERR or (st == SDGPU_NONE and not allow_none):
            raise SdError(f"{self.api.prefix}{what} failed ({st}): {self.api.error()}")
        return st

    def close(self):
        if self.ctx:
            self.api._fn("destroy")(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._check(self._call("reset"), "reset")

    def counts(self) -> dict:
        c = CCounts()
        self._check(self._call("get_counts", C.byref(c)), "get_counts")
        return {"omega": c.omega, "lambda": c.lambda_, "sigma": c.sigma, "basis": c.basis}

    # -- tables -----------------------------------------------------------------------------------------
    def calc_omega(self, observ, tol):
        o, flag = _f64(observ), C.c_int(0)
        idx = self._check(self._call("calc_omega", _pf64(o), tol, C.byref(flag)), "calc_omega")
        return idx, bool(flag.value)

    def omega_find(self, observ, tol):
        o = _f64(observ)
        return self._check(self._call("omega_find", _pf64(o), tol), "omega_find", allow_none=True)

    def omega_append(self, observ, weight=1):
        o = _f64(observ)
        return self._check(self._call("omega_append", _pf64(o), weight), "omega_append")

    def omega_bump(self, idx, by=1):
        self._check(self._call("omega_bump", idx, by), "omega_bump")

    def omega_append_bulk(self, vals, weights=None):
        v = _f64(vals)
        w = None if weights is None else _i32(weights)
        self._check(self._call("omega_append_bulk", v.shape[0], _pf64(v), _pi32(w)), "omega_append_bulk")

    def calc_lambda(self, Pi, tol):
        p, flag = _f64(Pi), C.c_int(0)
        idx = self._check(self._call("calc_lambda", _pf64(p), tol, C.byref(flag)), "calc_lambda")
        return idx, bool(flag.value)

    def calc_sigma(self, pi, mubBar, idxLambda, newLambdaFlag, currentIter, tol):
        p, flag = _f64(pi), C.c_int(0)
        idx = self._check(self._call("calc_sigma", _pf64(p), mubBar, idxLambda, int(newLambdaFlag), currentIter, tol,
                                     C.byref(flag)), "calc_sigma")
        return idx, bool(flag.value)

    def calc_delta(self, newOmegaFlag, elemIdx):
        self._check(self._call("calc_delta", int(newOmegaFlag), elemIdx), "calc_delta")

    def calc_delta_block(self, l0, l1, o0, o1):
        self._check(self._call("calc_delta_block", l0, l1, o0, o1), "calc_delta_block")

    def get_delta_block(self, l0, l1, o0, o1, plane=0):
        out = np.zeros((l1 - l0, o1 - o0))
        self._check(self._call("get_delta_block", l0, l1, o0, o1, plane, _pf64(out)), "get_delta_block")
        return out

    def update_dual(self, pi, mubBar, currentIter, tol):
        p = _f64(pi)
        li, nl, si, ns = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        self._check(self._call("update_dual", _pf64(p), mubBar, currentIter, tol, C.byref(li), C.byref(nl),
                               C.byref(si), C.byref(ns)), "update_dual")
        return li.value, bool(nl.value), si.value, bool(ns.value)

    def update_dual_col(self, newOmegaIdx, pi, mubBar, currentIter, tol):
        """the delta column of a new observation (index, or -1) + calcLambda + calcSigma + delta row in one device round trip"""
        if not self.api.has("update_dual_col"):                       # a checker without the fused entry point: the two calls it stands for
            if newOmegaIdx >= 0:
                self.calc_delta(True, newOmegaIdx)
            return self.update_dual(pi, mubBar, currentIter, tol)
        p = _f64(pi)
        li, nl, si, ns = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        self._check(self._call("update_dual_col", int(newOmegaIdx), _pf64(p), mubBar, currentIter, tol, C.byref(li), C.byref(nl),
                               C.byref(si), C.byref(ns)), "update_dual_col")
        return li.value, bool(nl.value), si.value, bool(ns.value)

    def update_dual_bulk(self, pis, mubBar=None, iters=None, tol=1e-3):
        p = _f64(pis)
        n = p.shape[0]
        mb = None if mubBar is None else _f64(mubBar)
        it = None if iters is None else _i32(iters)
        li, si = np.empty(n, np.int32), np.empty(n, np.int32)
        self._check(self._call("update_dual_bulk", n, _pf64(p), _pf64(mb), _pi32(it), tol, _pi32(li), _pi32(si)),
                    "update_dual_bulk")
        return li, si

    def basis_append(self, ck, feasFlag=True, sigmaIdx=(0,), omegaIdx=None):
        s = _i32(list(sigmaIdx))
        phi = len(s) - 1
        om = None if omegaIdx is None else _i32(list(omegaIdx))
        return self._check(self._call("basis_append", ck, int(feasFlag), phi, _pi32(s), _pi32(om)), "basis_append")

    def basis_append_bulk(self, ck, sigmaIdx, feas=None):
        c, s = _i32(ck), _i32(sigmaIdx)
        f = None if feas is None else _i32(feas)
        return self._check(self._call("basis_append_bulk", len(c), _pi32(c), _pi32(f), _pi32(s)), "basis_append_bulk")

    def basis_find_or_append(self, retainBasis, obsIdx, ck, feasFlag=True, sigmaIdx=(0,), omegaIdx=None):
        s = _i32(list(sigmaIdx))
        phi = len(s) - 1
        om = None if omegaIdx is None else _i32(list(omegaIdx))
        flag = C.c_int(0)
        idx = self._check(self._call("basis_find_or_append", int(retainBasis), obsIdx, ck, int(feasFlag), phi,
                                     _pi32(s), _pi32(om), C.byref(flag)), "basis_find_or_append")
        return idx, bool(flag.value)

    def basis_set_obs_feasible(self, basisIdx, obsIdx, flag):
        self._check(self._call("basis_set_obs_feasible", basisIdx, obsIdx, int(flag)), "basis_set_obs_feasible")

    def basis_set_obs_feasible_row(self, basisIdx, flags):
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        self._check(self._call("basis_set_obs_feasible_row", basisIdx, f.ctypes.data_as(c_u8p)), "basis_set_obs_feasible_row")

    def basis_set_obs_feasible_col(self, obsIdx, flags):
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        self._check(self._call("basis_set_obs_feasible_col", obsIdx, f.ctypes.data_as(c_u8p)), "basis_set_obs_feasible_col")

    # -- checkBasisFeasibility (randCost.c:202-258) ---------------------------------------------------------
    def set_cost_coords(self, rvdOmCols, senx: bytes):
        r = _i32(rvdOmCols)
        self._check(self._call("set_cost_coords", _pi32(r), senx), "set_cost_coords")

    def basis_set_feas_data(self, basisIdx, piDet, phi, gBar, psiVal, cstat):
        p, g, cs = _f64(piDet), _f64(gBar), _i32(cstat)
        ph = None if phi is None or len(phi) == 0 else _f64(phi)
        ps = None if psiVal is None or len(psiVal) == 0 else _f64(psiVal)
        self._check(self._call("basis_set_feas_data", basisIdx, _pf64(p), _pf64(ph), _pf64(g), _pf64(ps), _pi32(cs)), "basis_set_feas_data")

    def check_feasibility_obs(self, obsIdx, tol):
        f = np.zeros(max(1, self.counts()["basis"]), np.uint8)
        self._check(self._call("check_feasibility_obs", obsIdx, tol, f.ctypes.data_as(c_u8p)), "check_feasibility_obs")
        return f[:self.counts()["basis"]]

    def check_feasibility_basis(self, basisIdx, tol):
        f = np.zeros(max(1, self.counts()["omega"]), np.uint8)
        self._check(self._call("check_feasibility_basis", basisIdx, tol, f.ctypes.data_as(c_u8p)), "check_feasibility_basis")
        return f[:self.counts()["omega"]]

    def feas_cuts(self, obsFirst, obsLast, basisFirst, basisLast, maxOut=4096):
        """raw feasibility cuts (cuts.c:473-490 loop body) in the reference's loop order"""
        alpha, beta = np.zeros(maxOut), np.zeros((maxOut, self.problem.prevCols + 1))
        n = self._check(self._call("feas_cuts", obsFirst, obsLast, basisFirst, basisLast, maxOut, _pf64(alpha), _pf64(beta)), "feas_cuts")
        return alpha[:n], beta[:n]

    def feas_pool_update(self, fUpdt, tol):
        """updtFeasCutPool cuts.c:465-517 incl. the duplicate test of addCut2Pool cuts.c:643-655; fUpdt (list of 2) is updated in place"""
        fu = (C.c_int * 2)(*fUpdt)
        n = self._check(self._call("feas_pool_update", fu, tol), "feas_pool_update")
        fUpdt[0], fUpdt[1] = fu[0], fu[1]
        return n

    def feas_pool(self):
        n = self._check(self._call("feas_pool_size"), "feas_pool_size")
        alpha, beta = np.zeros(max(n, 1)), np.zeros((max(n, 1), self.problem.prevCols + 1))
        if n:
            self._check(self._call("feas_pool_get", 0, n, _pf64(alpha), _pf64(beta)), "feas_pool_get")
        return alpha[:n], beta[:n]

    def feas_pool_check(self, fAlpha, fBeta, incumbX, candidX, tol):
        """checkFeasCutPool cuts.c:521-567: (action per pool cut, infeasIncumb)"""
        n = self._check(self._call("feas_pool_size"), "feas_pool_size")
        fa, fb = _f64(fAlpha), _f64(fBeta)
        ix, cx = _f64(incumbX), _f64(candidX)
        act, inf = np.zeros(max(n, 1), np.int32), C.c_int(0)
        self._check(self._call("feas_pool_check", len(fa), _pf64(fa) if len(fa) else None, _pf64(fb) if len(fa) else None, _pf64(ix), _pf64(cx), tol,
                               _pi32(act), C.byref(inf)), "feas_pool_check")
        return act[:n], bool(inf.value)

    # -- cut formation ----------------------------------------------------------------------------------
    def compute_istar(self, X, obs, numSamples, pi_eval, isNew):
        x, am = _f64(X), C.c_double(0.0)
        idx = self._check(self._call("compute_istar", _pf64(x), obs, numSamples, int(pi_eval), int(isNew), C.byref(am)),
                          "compute_istar", allow_none=True)
        return idx, am.value

    def _cut_buffers(self, want_istar=True):
        beta = np.zeros(self.problem.prevCols + 1, np.float64)
        n = self.counts()["omega"]
        istar = np.full(max(n, 1), -7, np.int32) if want_istar else None
        cut = CCut(0.0, _pf64(beta), _pi32(istar), 0, 0, 0.0, 0.0)
        return cut, beta, istar, n

    @staticmethod
    def _cut_out(cut, beta, istar, n):
        return Cut(cut.alpha, beta, None if istar is None else istar[:n], cut.omegaCnt, cut.numSamples,
                   cut.cummOld, cut.cummAll)

    def sd_cut(self, X, numSamples, pi_eval_flag, lb, want_istar=True, variant="sd_cut"):
        """SDCut cuts.c:91-194.  Returns a Cut, or None where the reference returns NULL."""
        x = _f64(X)
        cut, beta, istar, n = self._cut_buffers(want_istar)
        if variant == "sd_cut_omp":
            nt = C.c_int(0)
            st = self._call(variant, _pf64(x), numSamples, int(pi_eval_flag), lb, C.byref(cut), C.byref(nt))
        else:
            st = self._call(variant, _pf64(x), numSamples, int(pi_eval_flag), lb, C.byref(cut))
        if st == SDGPU_NONE:
            return None
        self._check(st, variant)
        return self._cut_out(cut, beta, istar, n)

    def sd_cut_partial(self, X, numSamples, pi_eval_flag, lb):
        x = _f64(X)
        self._check(self._call("sd_cut_partial", _pf64(x), numSamples, int(pi_eval_flag), lb), "sd_cut_partial")

    def sd_cut_partial_buffer(self):
        p, n = C.c_void_p(), C.c_int(0)
        self._check(self._call("sd_cut_partial_buffer", C.byref(p), C.byref(n)), "sd_cut_partial_buffer")
        return p.value, n.value

    def sd_cut_finish(self, numSamples, want_istar=True):
        cut, beta, istar, n = self._cut_buffers(want_istar)
        st = self._call("sd_cut_finish", numSamples, C.byref(cut))
        if st == SDGPU_NONE:
            return None
        self._check(st, "sd_cut_finish")
        return self._cut_out(cut, beta, istar, n)

    def cut_heights(self, alpha, beta, numSamples, alphaIncumb, currIter, xk, lb):
        a, b, ns, x = _f64(alpha), _f64(beta), _i32(numSamples), _f64(xk)
        ai = None if alphaIncumb is None else _f64(alphaIncumb)
        n = len(a)
        h, e, r = np.zeros(n), np.zeros(n), np.zeros(n)
        best = self._check(self._call("cut_heights", n, _pf64(a), _pf64(b), _pi32(ns), _pf64(ai), currIter, _pf64(x), lb,
                                      _pf64(h), _pf64(e), _pf64(r)), "cut_heights", allow_none=True)
        return best, h, e, r

    def reform_cut(self, iStar, observ, k, lbType, lb):
        ob = _i32(observ)
        ist = None if iStar is None else _i32(iStar)
        n = 0 if ist is None else len(ist)
        alpha, beta = C.c_double(0.0), np.zeros(self.problem.prevCols + 1)
        self._check(self._call("reform_cut", _pi32(ist), n, _pi32(ob), k, lbType, int(lb), C.byref(alpha), _pf64(beta)),
                    "reform_cut")
        return alpha.value, beta

    def reform_cuts_batch(self, iStars, observ, lbType, lb):
        """iStars: list of int32 arrays (one per cut); observ: [nReps][k].  Returns alpha[nReps][nCuts], beta[nReps][nCuts][n1+1]."""
        ob = _i32(observ)
        nReps, k = ob.shape
        nCuts = len(iStars)
        oc = _i32([len(x) for x in iStars])
        stride = int(max(1, oc.max()))
        mat = np.zeros((nCuts, stride), np.int32)
        for n, x in enumerate(iStars):
            mat[n, :len(x)] = x
        alpha = np.zeros((nReps, nCuts))
        beta = np.zeros((nReps, nCuts, self.problem.prevCols + 1))
        self._check(self._call("reform_cuts_batch", nCuts, _pi32(mat), stride, _pi32(oc), nReps, _pi32(ob), k, lbType, int(lb),
                               _pf64(alpha), _pf64(beta)), "reform_cuts_batch")
        return alpha, beta

    # -- readers ----------------------------------------------------------------------------------------
    def get_omega(self, idx):
        v, w = np.zeros(self.problem.numRV + 1), C.c_int(0)
        self._check(self._call("get_omega", idx, _pf64(v), C.byref(w)), "get_omega")
        return v, w.value

    def get_lambda(self, idx):
        v = np.zeros(self.problem.rvRowCnt + 1)
        self._check(self._call("get_lambda", idx, _pf64(v)), "get_lambda")
        return v

    def get_sigma(self, idx):
        pib, piC, li, ck = C.c_double(0.0), np.zeros(self.problem.cntCcols + 1), C.c_int(0), C.c_int(0)
        self._check(self._call("get_sigma", idx, C.byref(pib), _pf64(piC), C.byref(li), C.byref(ck)), "get_sigma")
        return pib.value, piC, li.value, ck.value

    def get_delta(self, lambdaIdx, obsIdx):
        pib, piC = C.c_double(0.0), np.zeros(self.problem.rvCOmCnt + 1)
        self._check(self._call("get_delta", lambdaIdx, obsIdx, C.byref(pib), _pf64(piC)), "get_delta")
        return pib.value, piC

    def stats(self) -> dict:
        s = CStats()
        self._check(self._call("get_stats", C.byref(s)), "get_stats")
        return {k: getattr(s, k) for k, _ in CStats._fields_}

    def set_timing(self, on: bool = True):
        self._check(self._call("set_timing", int(on)), "set_timing")

    def delta_memory(self):
        """(mapped on demand?, bytes of address space reserved, bytes of physical memory mapped) of the delta table"""
        r, m = C.c_int64(0), C.c_int64(0)
        st = self._call("delta_memory", C.byref(r), C.byref(m))
        self._check(st, "delta_memory")
        return bool(st), r.value, m.value

    def set_sweep_variant(self, v: int):
        self._check(self._call("set_sweep_variant", v), "set_sweep_variant")

    def set_collective(self, mode: int):
        """0 automatic (peer exchange if attached, else NCCL), 1 NCCL, 2 NVLink peer exchange"""
        self._check(self._call("set_collective", mode), "set_collective")

    # -- the reference's stochasticUpdates, minus the CPLEX calls ----------------------------------------
    def stochastic_updates(self, omegaIdx, newOmegaFlag, piDet, mubBar, currentIter, tol, feasFlag=True,
                           phi=(), phiOmegaIdx=()):
        """stocUpdate.c:14-133 with the solver outputs (piDet, mubBar, phi columns) passed in.
        Returns (basisIdx, newBasisFlag).  The basis-code shortcut (:39-53) is the caller's."""
        li, nl, s0, ns = self.update_dual_col(omegaIdx if newOmegaFlag else -1, piDet, mubBar, currentIter, tol)   # :24-25, :78-85
        sig, retain = [s0], ns                                                # :87
        for col in phi:                                                       # :88-99
            li, nl, sk, ns = self.update_dual(col, 0.0, currentIter, tol)
            sig.append(sk)
            retain = retain or ns
        om = None if not len(phi) else [0] + list(phiOmegaIdx)
        return self.basis_find_or_append(retain, omegaIdx, currentIter, feasFlag, sig, om)   # :101-131
