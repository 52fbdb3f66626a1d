#!/usr/bin/env python
"""bench.py -- SD cut formation throughput (dual x observation argmax pairs / second) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (libsdgpu.so)
    python bench.py --impl reference [...]                          # the reference's own CPU SDCut (oracle/_ref)
    torchrun --nproc-per-node N ... bench.py --gpus N ...           # one rank per GPU, observations sharded

Workload (BASELINE.json config 5, the synthetic cut-formation sweep): `--duals` stored dual vertices x
`--obs-per-gpu` observations per GPU x `--rv` random right-hand-side elements, weak scaling in the observation
count -- the default 65 536 x 131 072 per GPU is the 64K x 1M x 256 configuration at 8 GPUs and a 64 GiB delta
table per GPU.  One "step" is one SDCut (two-window pi_eval mode, the shipped default config.sd) over the
resident tables: `value` times exactly that with CUDA events; `e2e` times one whole SD iteration through the
C ABI with host buffers (new observation -> calcOmega + delta column, new dual vertex -> calcLambda / calcSigma
/ delta row, then SDCut with x from the host and alpha / beta / iStar back on the host), wall clock.
The delta table (64 GiB) is far larger than L2 (126 MB), so no L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "dual x observation argmax pairs/sec (SD cut formation)"
UNIT = "pairs/s"
SEED = 20240607


# --------------------------------------------------------------------------------------------------------------
# synthetic tables
# --------------------------------------------------------------------------------------------------------------
def make_workload(duals: int, obs: int, rv: int, n1: int, rank: int, extra: int):
    """Same duals on every rank (replicated lambda / sigma), a rank-specific shard of observations."""
    from stochasticdecomposition_b200.synthetic import make_problem
    rows = max(rv, 8)
    prob = make_problem(SEED, rows=rows, cols=2 * rows, n1=n1, n1c=n1, R=rv, Rb=rv, Q=0, rvd=0)
    rng = np.random.default_rng(SEED)
    pis = rng.uniform(-1.0, 1.0, (duals + extra, rows + 1))
    pis[rng.random(pis.shape) < 0.3] = 0.0                         # 30 % exact zeros
    ncopy = max(1, duals // 100)                                   # tie stress: 1 % exact copies of an earlier dual
    dst = rng.choice(np.arange(1, duals), size=ncopy, replace=False)
    for d in dst:
        pis[d] = pis[rng.integers(0, d)]
    pis[:, 0] = 0.0
    orng = np.random.default_rng(SEED + 1000 * (rank + 1))
    obsv = orng.normal(0.0, 1.0, (obs + extra, prob.numRV + 1))
    obsv[:, 0] = 0.0
    if rank == 0:
        obsv[orng.choice(obs, size=min(16, obs), replace=False)] = 0.0    # 16 all-zero observations
    weights = (1 + orng.poisson(0.25, obs)).astype(np.int32)
    xs = rng.uniform(0.0, 1.0, (64, n1 + 1))
    xs[:, 0] = 0.0
    return prob, pis, obsv, weights, xs


def load_tables(api, prob, pis, obsv, weights, duals, obs, k_total, extra, device=0):
    from stochasticdecomposition_b200._abi import Caps
    caps = Caps(duals + extra + 8, duals + extra + 8, duals + extra + 8, obs + extra + 8, 1)
    t = api.create(prob, caps, device)
    t.omega_append_bulk(obsv[:obs], weights)
    iters = np.ceil((np.arange(duals) + 1) * (k_total / duals)).astype(np.int32)      # monotone ck: ~90/10 window split
    t.update_dual_bulk(pis[:duals], None, iters, -1.0)
    t.calc_delta_block(0, duals, 0, obs)
    t.basis_append_bulk(iters, np.arange(duals, dtype=np.int32))   # one basis per sigma (plain branch: basis index == sigma index)
    return t


# --------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc, self.first = index, [], None, 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        self.first = len(self.lines)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        timed = self.lines[self.first:] if len(self.lines) > self.first else self.lines
        for ln in timed:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(duals, obs, kernel=None):
    """dram bytes per sweep launch from the committed ncu capture, if one exists for this shape (and kernel)."""
    path = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    try:
        with open(path) as fh:
            rec = json.load(fh)
        for r in rec.get("captures", []):
            if r.get("duals") == duals and r.get("obs") == obs and (kernel is None or r.get("kernel", "").startswith(kernel)):
                return r.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------------------------------------------
# CPU legs (test infrastructure used as the measured-beside baseline only)
# --------------------------------------------------------------------------------------------------------------
def cpu_sample_dims(args):
    return min(args.cpu_duals, args.duals), min(args.cpu_obs, args.obs_per_gpu)


def run_cpu(kind: str, args, steps: int, warmup: int):
    """Times SDCut on a bounded sample of the workload.  kind = "reference": the reference's own cuts.c /
    stocUpdate.c (oracle/_ref, single thread -- the reference has no threading); "port_omp": the restated
    oracle with OpenMP over observations on all host cores."""
    import oracle_loader
    from stochasticdecomposition_b200._abi import CCut, Caps, _pf64, _pi32
    import ctypes as C
    D, N = cpu_sample_dims(args)
    prob, pis, obsv, weights, xs = make_workload(D, N, args.rv, args.n1, 0, 0)
    k_total = int(weights.sum())
    iters = np.ceil((np.arange(D) + 1) * (k_total / D)).astype(np.int32)
    caps = Caps(D + 8, D + 8, D + 8, N + 8, 1)
    t0 = time.perf_counter()
    if kind == "reference" and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsdref.so")) and not os.path.isdir("/root/reference"):
        kind = "port_single"             # the reference build did not travel to this box: fall back to the restated oracle, one thread
    if kind == "reference":
        api = oracle_loader.reference()
        t = api.create(prob, caps)
        li, si = np.zeros(D, np.int32), np.zeros(D, np.int32)
        o, p = np.ascontiguousarray(obsv[:N]), np.ascontiguousarray(pis[:D])
        st = api._fn("bulk_load")(t.ctx, N, _pf64(o), _pi32(weights), D, _pf64(p), None, _pi32(iters), -1.0, _pi32(li), _pi32(si))
        assert st == 0
        cores, variant = 1, "sd_cut"
    else:
        api = oracle_loader.oracle()
        t = api.create(prob, caps)
        t.omega_append_bulk(obsv[:N], weights)
        li, si = t.update_dual_bulk(pis[:D], None, iters, -1.0)
        t.calc_delta_block(0, D, 0, N)
        cores, variant = (os.cpu_count() or 1, "sd_cut_omp") if kind == "port_omp" else (1, "sd_cut")
    for b in range(int(si.max()) + 1):
        t.basis_append(int(iters[b]), True, [b])          # (the reference build has no bulk form; a few thousand calls)
    nb = t.counts()["basis"]
    build_s = time.perf_counter() - t0
    beta = np.zeros(prob.prevCols + 1)
    istar = np.zeros(N, np.int32)
    cut = CCut(0.0, _pf64(beta), _pi32(istar), 0, 0, 0.0, 0.0)
    fn = api._fn(variant)
    nt = C.c_int(0)
    times = []
    for s in range(warmup + steps):
        x = np.ascontiguousarray(xs[s % len(xs)])
        t1 = time.perf_counter()
        if variant == "sd_cut_omp":
            st = fn(t.ctx, _pf64(x), k_total, 1, 0.0, C.byref(cut), C.byref(nt))
        else:
            st = fn(t.ctx, _pf64(x), k_total, 1, 0.0, C.byref(cut))
        dt = time.perf_counter() - t1
        assert st == 0, st
        if s >= warmup:
            times.append(dt)
    if variant == "sd_cut_omp":
        cores = nt.value
    sec = float(np.mean(times))
    return {"value": nb * N / sec, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "reference" else "port",
            "sample": f"{nb} duals x {N} observations x {args.rv} random elements, pi_eval on, {steps} SDCut calls "
                      f"({sec * 1e3:.1f} ms each; table build {build_s:.1f} s not timed)",
            "ms_per_step": sec * 1e3, "pairs_per_step": nb * N}


def sd_iterations_leg(iterations: int):
    """BASELINE.json's third figure: SD iterations per second on an ssn-shaped instance (n1 = 89, 175 rows, 86 random right-hand
    sides), the same host loop (tools/sd_highs_host.py: HiGHS for the subproblem LP and the master, not CPLEX; synthetic instance, the
    SMPS file is not available offline) over this library's tables and over the reference's CPU tables (oracle/_ref)."""
    try:
        tools = os.path.join(ROOT, "tools")
        if tools not in sys.path:
            sys.path.insert(0, tools)
        import oracle_loader
        import stochasticdecomposition_b200 as sd
        from sd_iterations_bench import run
        g = run(sd.load_library(), "gpu", "ssn", iterations, 7)
        have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsdref.so")) or os.path.isdir("/root/reference")
        r = run(oracle_loader.reference() if have_ref else oracle_loader.oracle(), "reference" if have_ref else "port", "ssn", iterations, 7)
        return {"shape": "ssn-shaped synthetic instance", "iterations": iterations, "lp_solver": "HiGHS (scipy), not CPLEX",
                "gpu_tables_it_per_s": g["iterations_per_s"], "gpu_argmax_seconds": g["argmax_seconds"], "gpu_argmax_share": g["argmax_share"],
                "cpu_tables_it_per_s": r["iterations_per_s"], "cpu_argmax_seconds": r["argmax_seconds"], "cpu_argmax_share": r["argmax_share"],
                "cpu_tables_kind": r["backend"], "same_incumbent_estimate": abs(g["incumbent_estimate"] - r["incumbent_estimate"]) <= 1e-9 * max(1.0, abs(r["incumbent_estimate"]))}
    except Exception as exc:                                  # an optional extra: never let it take the bench line down
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    res = run_cpu("reference", args, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "sample": res["sample"],
                       "note": "reference's own SDCut/computeIstar (cuts.c, stocUpdate.c) built from /root/reference against the header "
                               "shim; single thread because the reference is single-threaded"},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(args):
    return (f"synthetic cut-formation sweep: {args.duals} dual vertices x {args.obs_per_gpu} observations per GPU x "
            f"{args.rv} random elements, n1={args.n1}, Q=0, pi_eval two-window mode")


# --------------------------------------------------------------------------------------------------------------
# multi-GPU self-check: a sharded cut must equal the unsharded one (cuts.c:116-168,184-188 is ONE sum over all observations)
# --------------------------------------------------------------------------------------------------------------
def attach_library_nccl(api, t, rank, world, dist):
    """hand the library its own NCCL communicator (unique id from rank 0, shipped with torch.distributed)"""
    import ctypes as C
    import torch
    idbuf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = (C.c_char * 128)()
        assert api._fn("nccl_unique_id")(raw) == 0, api.error()
        idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
    idbuf = idbuf.cuda()
    dist.broadcast(idbuf, 0)
    raw = (C.c_char * 128).from_buffer_copy(bytes(idbuf.cpu().numpy().tobytes()))
    assert api._fn("nccl_init")(t.ctx, world, rank, raw) == 0, api.error()


def multi_gpu_parity(api, rank, world, local, dist, duals=2048, obs_per_rank=4096):
    """Before anything is timed: a small table (2 048 duals x 4 096 x world observations, two random technology-matrix elements,
    duplicated duals and all-zero observations => exact score ties) is built twice -- sharded over the ranks (observation o on rank
    o % world) and whole on rank 0 -- and cuts are formed on both, in both pi_eval modes, with both sweep families that size allows,
    once with the NCCL all-reduce and once with the NVLink peer exchange fused into the cut kernel.  Required: the gathered iStar
    equals the unsharded iStar bit for bit, alpha / beta / cummOld / cummAll agree within 1e-9 relative, every rank holds the same
    cut.  Returns the record printed as "multi_gpu_parity"; any mismatch makes every rank exit non-zero."""
    import torch
    from stochasticdecomposition_b200._abi import Caps
    from stochasticdecomposition_b200.sharding import ShardedTables
    from stochasticdecomposition_b200.synthetic import make_problem
    N = obs_per_rank * world
    prob = make_problem(SEED + 7, rows=24, cols=48, n1=20, n1c=16, R=16, Rb=12, Q=2)
    rng = np.random.default_rng(SEED + 7)
    pis = rng.uniform(-1.0, 1.0, (duals, prob.rows + 1))
    pis[rng.random(pis.shape) < 0.3] = 0.0
    for d in rng.choice(np.arange(1, duals), size=duals // 50, replace=False):     # 2 % exact copies: equal scores, lowest index must win
        pis[d] = pis[rng.integers(0, d)]
    pis[:, 0] = 0.0
    obsv = rng.normal(0.0, 1.0, (N, prob.numRV + 1))
    obsv[:, 0] = 0.0
    obsv[rng.choice(N, size=16, replace=False)] = 0.0
    weights = (1 + rng.poisson(0.25, N)).astype(np.int32)
    k_total = int(weights.sum())
    iters = np.ceil((np.arange(duals) + 1) * (k_total / duals)).astype(np.int32)
    xs = rng.uniform(0.0, 1.0, (4, prob.prevCols + 1))
    xs[:, 0] = 0.0

    def build(ob, w):
        t = api.create(prob, Caps(duals + 8, duals + 8, duals + 8, len(ob) + 8, 1), local)
        t.omega_append_bulk(ob, w)
        t.update_dual_bulk(pis, None, iters, -1.0)
        t.calc_delta_block(0, duals, 0, len(ob))
        t.basis_append_bulk(iters, np.arange(duals, dtype=np.int32))
        return t

    sh = ShardedTables(build(obsv[rank::world], weights[rank::world]), rank, world)
    sh.total_obs = N
    whole = build(obsv, weights) if rank == 0 else None
    attach_library_nccl(api, sh.t, rank, world, dist)
    rec = {"ranks": world, "duals": duals, "observations": N, "cuts_checked": 0, "max_rel_err": 0.0, "tolerance": 1e-9}
    modes = [("nccl", 1), ("peer", 2)]
    try:
        sh.attach_peer_exchange()                # both exchanges attached; sdgpu_set_collective picks one per cut
    except Exception as exc:                     # (raised on every rank or on none) no CUDA IPC between these processes: NCCL only
        rec["peer"] = f"unavailable: {exc}"
        modes = modes[:1]
    ok_all = True
    for name, mode in modes:
        sh.t.set_collective(mode)
        ok = True
        for variant in (0, 1):                   # automatic (bulk-copy ring at this size) and the load-based sweep
            sh.t.set_sweep_variant(variant)
            if whole is not None:
                whole.set_sweep_variant(variant)
            for ci, pi_eval in enumerate((1, 0, 1)):
                x = xs[ci]
                cut = sh.t.sd_cut(x, k_total, pi_eval, 0.0)
                vec = np.concatenate([[cut.alpha], cut.beta[1:], [cut.cummOld, cut.cummAll]])
                mine = torch.from_numpy(vec).cuda()
                allv = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(allv, mine)
                same_everywhere = all(bool(torch.equal(v, allv[0])) for v in allv)
                istar = sh.gather_istar(cut.iStar)
                if rank == 0:
                    ref = whole.sd_cut(x, k_total, pi_eval, 0.0)
                    sab = max(abs(ref.alpha), float(np.abs(ref.beta[1:]).max()), 1e-300)      # beta entries against the cut's largest coefficient
                    err = max(abs(cut.alpha - ref.alpha) / max(abs(ref.alpha), 1e-300), float(np.abs(cut.beta - ref.beta).max()) / sab,
                              abs(cut.cummOld - ref.cummOld) / max(abs(ref.cummOld), 1e-300), abs(cut.cummAll - ref.cummAll) / max(abs(ref.cummAll), 1e-300))
                    rec["max_rel_err"] = max(rec["max_rel_err"], err)
                    good = np.array_equal(istar, ref.iStar) and err <= 1e-9 and same_everywhere
                    if not good:
                        print(f"bench.py: multi-GPU parity FAILED ({name}, variant {variant}, pi_eval {pi_eval}): iStar equal "
                              f"{np.array_equal(istar, ref.iStar)}, max rel err {err:.3e}, identical on all ranks {same_everywhere}", file=sys.stderr)
                    ok = ok and good
                    rec["cuts_checked"] += 1
        rec[name] = "ok" if ok else "MISMATCH"
        ok_all = ok_all and ok
    flag = torch.tensor([1 if ok_all else 0], dtype=torch.int32, device="cuda")
    dist.broadcast(flag, 0)
    sh.t.close()
    if whole is not None:
        whole.close()
    dist.barrier()
    if int(flag.item()) != 1:
        raise SystemExit(3)
    return rec


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------
def gpu_arm(args):
    # stdout must carry exactly one JSON line: NCCL and torchrun helpers print banners to fd 1, so route fd 1 to stderr until then
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import stochasticdecomposition_b200 as sd
    from stochasticdecomposition_b200._abi import CCut, _pf64, _pi32
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU leg)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

    api = sd.load_library()
    parity = None
    if world > 1 and not args.no_parity:
        parity = multi_gpu_parity(api, rank, world, local, dist)      # exits non-zero on every rank if a sharded cut differs

    strong = args.scaling == "strong"
    D = args.strong_duals if strong else args.duals
    N = (args.strong_obs // world) if strong else args.obs_per_gpu     # observations on this GPU
    extra = args.steps + args.warmup + 4
    prob, pis, obsv, weights, xs = make_workload(D, N, args.rv, args.n1, rank, extra)
    k_local = int(weights.sum())
    k_total = k_local
    if world > 1:
        kt = torch.tensor([k_local], dtype=torch.int64, device="cuda")
        dist.all_reduce(kt)
        k_total = int(kt.item())
    t_setup = time.perf_counter()
    t = load_tables(api, prob, pis, obsv, weights, D, N, k_total, extra, local)
    setup_s = time.perf_counter() - t_setup
    collectives = []
    if world > 1:
        # both exchanges are attached; sdgpu_set_collective picks the one a cut uses (the headline uses --collective; the strong-scaling
        # record times both side by side)
        from stochasticdecomposition_b200.sharding import ShardedTables
        attach_library_nccl(api, t, rank, world, dist)
        have_peer = True
        try:
            ShardedTables(t, rank, world).attach_peer_exchange()
        except Exception:                                     # consistent across ranks (see attach_peer_exchange)
            have_peer = False
        chosen = args.collective if args.collective != "auto" else ("peer" if world >= 4 else "nccl")
        collectives = ["nccl", "peer"] if (strong or args.both_collectives) else [chosen]
        if not have_peer:
            collectives = ["nccl"]
        t.set_collective({"nccl": 1, "peer": 2}[collectives[0]])

    stream = torch.cuda.Stream()
    t._check(api._fn("set_stream")(t.ctx, C.c_void_p(stream.cuda_stream)), "set_stream")
    t.set_timing(True)
    beta = np.zeros(prob.prevCols + 1)
    istar = np.zeros(N + extra + 8, np.int32)
    cut_dev = CCut(0.0, _pf64(beta), None, 0, 0, 0.0, 0.0)           # iStar stays device-resident
    cut_host = CCut(0.0, _pf64(beta), _pi32(istar), 0, 0, 0.0, 0.0)
    sd_cut = api._fn("sd_cut")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_cut(step, cut):
        x = np.ascontiguousarray(xs[step % len(xs)])
        st = sd_cut(t.ctx, _pf64(x), k_total, 1, 0.0, C.byref(cut))
        if st != 0:
            raise RuntimeError(f"sd_cut failed ({st}): {api.error()}")

    def max_over_ranks(v):
        if dist is None:
            return float(v)
        mt = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(mt, op=dist.ReduceOp.MAX)
        return float(mt.item())

    def timed_cuts():
        """K cuts over resident tables, CUDA events on the library's stream, max over ranks; plus the per-phase split of a cut"""
        barrier()
        l0 = t.stats()["total_launches"]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        split = {"prep": [], "sweep": [], "merge": [], "collective": []}
        for s in range(args.steps):
            one_cut(args.warmup + s, cut_dev)
            stt = t.stats()
            for key, name in (("prep", "last_prep_ms"), ("sweep", "last_sweep_ms"), ("merge", "last_merge_ms"), ("collective", "last_collective_ms")):
                split[key].append(stt[name])
        ev1.record(stream)
        barrier()
        launches = t.stats()["total_launches"] - l0
        ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
        return ms_step, launches, {k: float(np.mean(v)) for k, v in split.items()}

    # ---- value: K cuts over resident tables ------------------------------------------------------------------------
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    phys = local
    try:                                                       # nvidia-smi counts physical GPUs, CUDA counts the visible ones
        phys = int(vis.split(",")[local]) if vis else local
    except (ValueError, IndexError):
        phys = local
    sampler = ClockSampler(phys)
    sampler.start()                                            # nvidia-smi takes a while to come up (longer with 8 GPUs): start it before the warm-up,
    for s in range(args.warmup):                               # count only the samples taken inside the timed regions
        one_cut(s, cut_dev)
    barrier()
    sampler.mark()
    per_collective = {}
    ms_step, launches, split = timed_cuts()
    if collectives:
        per_collective[collectives[0]] = {"ms_per_step": ms_step, "split_ms_rank0": split, "launches_per_step": launches / args.steps}
        for name in collectives[1:]:
            t.set_collective({"nccl": 1, "peer": 2}[name])
            for s in range(args.warmup):
                one_cut(s, cut_dev)
            ms2, l2, sp2 = timed_cuts()
            per_collective[name] = {"ms_per_step": ms2, "split_ms_rank0": sp2, "launches_per_step": l2 / args.steps}
        best = min(per_collective, key=lambda n: per_collective[n]["ms_per_step"]) if strong else collectives[0]
        t.set_collective({"nccl": 1, "peer": 2}[best])
        ms_step, split, launches = per_collective[best]["ms_per_step"], per_collective[best]["split_ms_rank0"], per_collective[best]["launches_per_step"] * args.steps
        used_collective = best
    else:
        used_collective = None
    st = t.stats()
    nb = t.counts()["basis"]
    pairs_step = nb * N * world
    value = pairs_step / (ms_step * 1e-3)
    sweep_bytes = st["last_sweep_bytes"]
    sweep_avg_ms = split["sweep"]
    peak, peak_src = measured_peak()
    achieved = sweep_bytes / (sweep_avg_ms * 1e-3) / 1e9
    sweep_kernel = {1: "k_sweep_ldg", 2: "k_sweep_tma", 3: "k_sweep_general", 4: "k_sweep_tma_gen", 5: "k_sweep_recompute", 6: "k_sweep_tma_grp"}.get(st["last_sweep_variant"], "?")

    # ---- e2e: whole SD iterations through the C ABI with host buffers, wall clock --------------------------------
    def one_iteration(i):
        ob = np.ascontiguousarray(obsv[N + i])
        oi, onew = t.calc_omega(ob, 1e-3)
        bi, bnew = t.stochastic_updates(oi, onew, np.ascontiguousarray(pis[D + i]), 0.0, k_total, 1e-3)
        one_cut(i, cut_host)
        return oi, bi

    one_iteration(0)
    barrier()
    w0 = time.perf_counter()
    for i in range(1, 1 + args.steps):
        one_iteration(i)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - w0)
    clocks = sampler.stop()                                   # sampled through both timed regions (value and e2e)
    cnt = t.counts()
    e2e_value = (nb * N * world) / (e2e_s / args.steps)
    h2d = 8 * (prob.numRV + 1) + 8 * (prob.rows + 1) + 8 * (prob.prevCols + 1)
    d2h = 8 * (prob.prevCols + 4) + 4 * cnt["omega"] + 64

    cpu = cpu_omp = same_shape = None
    if rank == 0 and world == 1 and not args.no_cpu:
        t.close()                                              # the big table goes before the small legs (64 GiB back)
        t = None
        same_shape = same_shape_leg(api, args, local)
        cpu = run_cpu("reference", args, 3, 1)
        cpu_omp = run_cpu("port_omp", args, 3, 1)
        same_shape.update({"cpu_reference_value": cpu["value"], "cpu_cores": cpu["cores"],
                           "ratio_value": same_shape["gpu_value"] / cpu["value"], "ratio_e2e": same_shape["gpu_e2e_value"] / cpu["value"]})
    sdit = None
    if rank == 0 and world == 1 and not args.no_cpu and args.sd_iterations > 0:
        sdit = sd_iterations_leg(args.sd_iterations)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "duals": nb, "observations_total": N * world, "delta_table_GiB_per_gpu": round(8 * nb * N / 2**30, 2),
                       "l2_policy": "inputs (the delta stream of one step, 8 bytes x duals x observations per GPU) far exceed the 126 MB L2; no flush needed",
                       "timing": "value: CUDA events on the library stream around K SDCut calls; e2e: wall clock around K full SD iterations "
                                 "(calcOmega, calcLambda/calcSigma/calcDelta, SDCut) with host buffers",
                       "sharding": ("observations split across ranks, lambda/sigma/basis replicated, one all-reduce of n1+4 doubles per cut ("
                                    + ("NVLink peer-memory exchange fused into the cut kernel" if used_collective == "peer" else "NCCL") + ")") if world > 1 else "single GPU",
                       "setup_s": round(setup_s, 2)},
            "roofline": {"bound": "hbm", "kernel": sweep_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": ncu_traffic(nb, N, sweep_kernel),
                         "traffic_source": "committed ncu capture of this kernel at this shape (profiles/sweep_traffic.json), not measured in this run",
                         "nominal_peak": 8000.0, "frac_of_nominal": achieved / 8000.0, "bytes_per_launch": sweep_bytes, "avg_launch_ms": sweep_avg_ms,
                         "sweep_share_of_step": sweep_avg_ms / ms_step},
            "step_split_ms_rank0": split,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if strong:
            line["config"]["workload"] = (f"strong scaling: fixed table of {D} dual vertices x {N * world} observations x {args.rv} random elements "
                                          f"({round(8 * D * N * world / 2**30, 1)} GiB of delta), observations split {world} way(s)")
        if parity is not None:
            line["multi_gpu_parity"] = parity
        if per_collective:
            line["collectives"] = per_collective
            line["config"]["collective"] = used_collective
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline_omp_port"] = {k: cpu_omp[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["vs_reference_same_shape"] = same_shape
        if sdit is not None:
            line["sd_iterations_ssn"] = sdit
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if t is not None:
        t.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def same_shape_leg(api, args, device):
    """The GPU path on exactly the shape the CPU reference leg times (the headline compares a 64 GiB table with a CPU
    sample 2 000 times smaller): the same synthetic generator at the CPU sample's dimensions, device-timed cuts and host-to-host
    iterations.  The table (0.5 GiB at 4 096 x 16 384) is larger than L2 but only just; consecutive cuts do re-use what stays in
    L2, exactly as consecutive SDCut calls on a real problem do -- and exactly as the CPU leg's caches do."""
    import ctypes as C
    from stochasticdecomposition_b200._abi import CCut, _pf64, _pi32
    D, N = cpu_sample_dims(args)
    steps = 20
    prob, pis, obsv, weights, xs = make_workload(D, N, args.rv, args.n1, 0, steps + 8)
    k_total = int(weights.sum())
    t = load_tables(api, prob, pis, obsv, weights, D, N, k_total, steps + 8, device)
    t.set_timing(True)
    beta, istar = np.zeros(prob.prevCols + 1), np.zeros(N + steps + 16, np.int32)
    cut_dev = CCut(0.0, _pf64(beta), None, 0, 0, 0.0, 0.0)
    cut_host = CCut(0.0, _pf64(beta), _pi32(istar), 0, 0, 0.0, 0.0)
    fn = api._fn("sd_cut")
    xc = [np.ascontiguousarray(x) for x in xs]
    for s in range(5):
        assert fn(t.ctx, _pf64(xc[s]), k_total, 1, 0.0, C.byref(cut_dev)) == 0, api.error()
    dev_ms = []
    for s in range(steps):
        assert fn(t.ctx, _pf64(xc[s % len(xc)]), k_total, 1, 0.0, C.byref(cut_dev)) == 0, api.error()
        dev_ms.append(t.stats()["last_cut_ms"])
    nb = t.counts()["basis"]
    w0 = time.perf_counter()
    for i in range(steps):
        oi, onew = t.calc_omega(np.ascontiguousarray(obsv[N + i]), 1e-3)
        t.stochastic_updates(oi, onew, np.ascontiguousarray(pis[D + i]), 0.0, k_total, 1e-3)
        assert fn(t.ctx, _pf64(xc[i % len(xc)]), k_total, 1, 0.0, C.byref(cut_host)) == 0, api.error()
    e2e_s = (time.perf_counter() - w0) / steps
    variant = t.stats()["last_sweep_variant"]
    t.close()
    ms = float(np.mean(dev_ms))
    return {"workload": f"{nb} duals x {N} observations x {args.rv} random elements (the CPU reference leg's sample), pi_eval on",
            "gpu_ms_per_cut": ms, "gpu_value": nb * N / (ms * 1e-3), "gpu_e2e_ms_per_iteration": e2e_s * 1e3, "gpu_e2e_value": nb * N / e2e_s,
            "sweep_variant": int(variant), "unit": UNIT}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--duals", type=int, default=65536)
    ap.add_argument("--obs-per-gpu", type=int, default=131072)
    ap.add_argument("--rv", type=int, default=256)
    ap.add_argument("--n1", type=int, default=89)
    ap.add_argument("--cpu-duals", type=int, default=4096)
    ap.add_argument("--cpu-obs", type=int, default=16384)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--sd-iterations", type=int, default=600, help="ssn-shaped SD run for the iterations/s figure (0 = skip)")
    ap.add_argument("--collective", default="auto", choices=["auto", "nccl", "peer"],
                    help="the cut's one exchange: NCCL all-reduce, or the NVLink peer exchange fused into the cut kernel; auto = peer from 4 GPUs "
                         "(measured on the strong-scaling table: 1.336 against 1.415 ms per cut at 8 GPUs, equal at 4, NCCL 0.6 %% ahead at 2)")
    ap.add_argument("--both-collectives", action="store_true", help="time the cut with NCCL and with the NVLink peer exchange, side by side")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the headline): --obs-per-gpu observations on every GPU; strong: a fixed --strong-duals x --strong-obs table split over the GPUs")
    ap.add_argument("--strong-duals", type=int, default=8192)
    ap.add_argument("--strong-obs", type=int, default=1048576)
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity self-check (N > 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
