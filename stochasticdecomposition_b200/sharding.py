"""Observation sharding for multi-GPU runs: one process per GPU, global observation o
lives on rank o % G at local slot o // G; lambda, sigma and the basis records are replicated (every rank
makes the same find-or-append calls and gets the same indices).  The only exchange step of a cut is an
all-reduce (sum) of the n1 + 4 doubles [alpha, beta[1..n1], cummOld, cummAll, missing].

Two ways to run that all-reduce:
  * in the library, on its own stream, with an NCCL communicator attached through sdgpu_nccl_init()
    (`attach_library_nccl`) -- the production path, one C call per cut;
  * through torch.distributed on the partial vector (`backend="torch"`): NCCL on GPUs, gloo on CPUs -- this is
    what the world_size-2 CPU tests drive, with the CPU checker standing in for the device tables.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._abi import Cut, SdError, Tables

_BIG = 2**62


def owner_of(o: int, world: int) -> int:
    return o % world


def local_slot(o: int, world: int) -> int:
    return o // world


def global_index(local: int, rank: int, world: int) -> int:
    return local * world + rank


def shard_counts(total: int, world: int) -> list[int]:
    """observations held by each rank when `total` observations were dealt round-robin"""
    return [(total - r + world - 1) // world for r in range(world)]


class ShardedTables:
    def __init__(self, tables: Tables, rank: int, world: int, group=None):
        if tables.problem.rvdOmCnt > 0:
            raise SdError("sharded mode supports rvdOmCnt == 0 (the obsFeasible mask is owner-local)")
        self.t, self.rank, self.world, self.group = tables, rank, world, group
        self.total_obs = 0
        self.library_nccl = False
        self._basis_sigmas, self._basis_feas = [], []          # mirror of (sigmaIdx[0], feasFlag) per basis

    # ---- plumbing ---------------------------------------------------------------------------------------
    def _dist(self):
        import torch.distributed as dist
        return dist

    def _allreduce_scalar_min(self, v: int) -> int:
        import torch
        dist = self._dist()
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        tsr = torch.tensor([v], dtype=torch.int64, device=dev)
        dist.all_reduce(tsr, op=dist.ReduceOp.MIN, group=self.group)
        return int(tsr.item())

    def attach_library_nccl(self):
        """Create an NCCL communicator inside libsdgpu.so (unique id from rank 0, shipped with torch.distributed)."""
        import torch
        dist = self._dist()
        api = self.t.api
        raw = (C.c_char * 128)()
        if self.rank == 0 and api._fn("nccl_unique_id")(raw) != 0:
            raise SdError("nccl_unique_id: " + api.error())
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone().to(dev)
        dist.broadcast(buf, 0, group=self.group)
        raw = (C.c_char * 128).from_buffer_copy(buf.cpu().numpy().tobytes())
        if api._fn("nccl_init")(self.t.ctx, self.world, self.rank, raw) != 0:
            raise SdError("nccl_init: " + api.error())
        self.library_nccl = True

    def attach_peer_exchange(self):
        """NVLink peer-memory all-reduce fused into the cut kernel: exchange CUDA IPC handles of the per-rank buffers."""
        import torch
        dist = self._dist()
        api = self.t.api
        raw = (C.c_char * 64)()
        err = ""
        if api._fn("peer_export")(self.t.ctx, self.world, raw) != 0:
            err = "peer_export: " + api.error()
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        mine = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone().to(dev)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=self.group)             # every rank takes part, whether or not its own export worked
        if not err:
            blob = b"".join(h.cpu().numpy().tobytes() for h in allh)
            buf = (C.c_char * len(blob)).from_buffer_copy(blob)
            if api._fn("peer_attach")(self.t.ctx, self.world, self.rank, buf) != 0:
                err = "peer_attach: " + api.error()
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)   # all ranks or none: a half-attached exchange would leave kernels spinning
        if int(ok.item()) != 1:
            raise SdError("peer exchange unavailable on at least one rank" + (": " + err if err else ""))
        self.library_nccl = True                      # sd_cut reduces inside the library from now on

    # ---- tables -----------------------------------------------------------------------------------------
    def calc_omega(self, observ, tol):
        """calcOmega stocUpdate.c:326-348 over the sharded observation set: the FIRST match in global order wins."""
        loc = self.t.omega_find(observ, tol)
        cand = global_index(loc, self.rank, self.world) if loc >= 0 else _BIG
        first = self._allreduce_scalar_min(cand)
        if first < _BIG:
            if owner_of(first, self.world) == self.rank:
                self.t.omega_bump(local_slot(first, self.world), 1)
            return first, False
        g = self.total_obs
        if owner_of(g, self.world) == self.rank:
            got = self.t.omega_append(observ, 1)
            assert got == local_slot(g, self.world)
        self.total_obs += 1
        return g, True

    def load_shard(self, vals, weights, total_after: int):
        """bulk load of this rank's observations (already dealt round-robin by the caller)"""
        self.t.omega_append_bulk(vals, weights)
        self.total_obs = total_after

    def stochastic_updates(self, omegaIdx, newOmegaFlag, piDet, mubBar, currentIter, tol, feasFlag=True):
        """stocUpdate.c:14-133 with replicated dual-side tables: the owner of a new observation fills its delta
        column, every rank runs the same calcLambda / calcSigma (same result) and fills its shard of a new delta row."""
        mine = owner_of(omegaIdx, self.world) == self.rank
        if newOmegaFlag and mine:
            self.t.calc_delta(True, local_slot(omegaIdx, self.world))
        li, nl, s0, ns = self.t.update_dual(piDet, mubBar, currentIter, tol)
        # obsFeasible is constant-true without random costs, so the dedup of stocUpdate.c:101-113 is observation independent;
        # any stored local observation can stand in for obsIdx (rank-local slot 0 exists whenever this rank holds one)
        b, new = self._basis_find_or_append(ns, currentIter, feasFlag, s0)
        if new:
            self._basis_sigmas.append(s0); self._basis_feas.append(bool(feasFlag))
        return b, new

    def _basis_find_or_append(self, newSigma, currentIter, feasFlag, s0):
        if self.t.counts()["omega"] > 0:
            return self.t.basis_find_or_append(newSigma, 0, currentIter, feasFlag, [s0], None)
        # this rank holds no observation yet: replay the observation-independent dedup on the host-visible basis list
        if not newSigma:
            for b, sig in enumerate(self._basis_sigmas):
                if sig == s0 and self._basis_feas[b]:
                    return b, False
        return self.t.basis_append(currentIter, feasFlag, [s0]), True

    # ---- cut --------------------------------------------------------------------------------------------
    def sd_cut(self, X, numSamples, pi_eval_flag, lb, want_istar=True):
        """SDCut over all shards.  Returns a Cut whose iStar covers this rank's observations only (local order)."""
        t = self.t
        if self.library_nccl:
            return t.sd_cut(X, numSamples, pi_eval_flag, lb, want_istar=want_istar)
        import torch
        dist = self._dist()
        n1 = t.problem.prevCols
        if t.api.has("sd_cut_partial_host"):                      # CPU checker standing in for the device tables (tests)
            part = np.zeros(n1 + 4)
            n = t.counts()["omega"]
            istar = np.full(max(n, 1), -7, np.int32)
            st = t.api._fn("sd_cut_partial_host")(t.ctx, np.ascontiguousarray(X, np.float64).ctypes.data_as(C.POINTER(C.c_double)),
                                                  numSamples, int(pi_eval_flag), lb, part.ctypes.data_as(C.POINTER(C.c_double)),
                                                  istar.ctypes.data_as(C.POINTER(C.c_int32)))
            if st < -1:
                raise SdError("sd_cut_partial_host: " + t.api.error())
            tsr = torch.from_numpy(part)
            dist.all_reduce(tsr, op=dist.ReduceOp.SUM, group=self.group)
            if part[n1 + 3] > 0:
                return None
            beta = np.zeros(n1 + 1)
            beta[1:] = part[1:n1 + 1] / numSamples                # cuts.c:184-188
            beta[0] = 1.0
            return Cut(part[0] / numSamples, beta, istar[:n] if want_istar else None, n, numSamples, part[n1 + 1], part[n1 + 2])
        t.sd_cut_partial(X, numSamples, pi_eval_flag, lb)
        ptr, n = t.sd_cut_partial_buffer()
        view = _DevVec(ptr, n)
        tsr = torch.as_tensor(view, device="cuda")
        torch.cuda.synchronize()
        dist.all_reduce(tsr, op=dist.ReduceOp.SUM, group=self.group)
        torch.cuda.synchronize()
        return t.sd_cut_finish(numSamples, want_istar=want_istar)

    def gather_istar(self, local_istar: np.ndarray) -> np.ndarray | None:
        """iStar in global observation order on rank 0 (chooseCuts / reformCuts need the whole vector, optimal.c:147)."""
        import torch
        dist = self._dist()
        counts = shard_counts(self.total_obs, self.world)
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        m = max(counts) if counts else 0
        mine = torch.full((max(m, 1),), -1, dtype=torch.int32, device=dev)
        mine[:len(local_istar)] = torch.as_tensor(np.ascontiguousarray(local_istar), device=dev)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        out = np.full(self.total_obs, -1, np.int32)
        for r, p in enumerate(parts):
            out[r::self.world] = p.cpu().numpy()[:counts[r]]
        return out


class _DevVec:
    """float64 device vector exposed through __cuda_array_interface__ (no copy)"""
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}
