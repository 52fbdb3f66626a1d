#!/usr/bin/env python
"""Measures the FP64 pipe peak of the GPU for separately rounded multiply + add (what the library issues) and for DFMA, and quotes the
FP64-bound kernels against it: the recompute sweep (k_sweep_recompute, (2 Rb + 2) operations per pair + 1 compare) and the bulk delta
build (k_delta_block_rhs, 2 Rb operations per cell)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
import stochasticdecomposition_b200 as sd

api = sd.load_library()
mul_add, fma = api.fp64_peak(0, 5)
print(json.dumps({"fp64_mul_add_ops_per_s": mul_add, "fp64_fma_flops_per_s": fma, "note": "ops = separately rounded DMUL and DADD instructions x lanes; "
                  "the library never issues DFMA on the path (bit-exactness), so mul_add is its FP64 roofline"}), flush=True)
for rb in (3, 5):
    D, N = 16384, 131072
    prob, pis, obsv, weights, xs = bench.make_workload(D, N, rb, 89, 0, 8)
    k = int(weights.sum())
    from stochasticdecomposition_b200._abi import Caps
    t = api.create(prob, Caps(D + 16, D + 16, D + 16, N + 16, 1))
    t.omega_append_bulk(obsv[:N], weights)
    iters = np.ceil((np.arange(D) + 1) * (k / D)).astype(np.int32)
    t.update_dual_bulk(pis[:D], None, iters, -1.0)
    t0 = time.perf_counter(); t.calc_delta_block(0, D, 0, N); build_s = time.perf_counter() - t0
    t.basis_append_bulk(iters, np.arange(D, dtype=np.int32))
    t.set_sweep_variant(3)
    t.set_timing(True)
    for s in range(3):
        t.sd_cut(xs[s], k, 1, 0.0, want_istar=False)
    ms = []
    for s in range(10):
        t.sd_cut(xs[s % 8], k, 1, 0.0, want_istar=False)
        ms.append(t.stats()["last_sweep_ms"])
    sweep_ms = float(np.median(ms))
    ops = (2 * rb + 2) * D * N                       # rb multiplies + rb adds (the contraction), 2 adds (the score); compares not counted
    print(json.dumps({"kernel": f"k_sweep_recompute<{rb}>", "duals": D, "observations": N, "sweep_ms": sweep_ms, "pairs_per_s": D * N / (sweep_ms * 1e-3),
                      "fp64_ops_per_s": ops / (sweep_ms * 1e-3), "frac_of_measured_mul_add_peak": ops / (sweep_ms * 1e-3) / mul_add,
                      "variant": t.stats()["last_sweep_variant"]}), flush=True)
    t.close()
# bulk delta build at 256 random right-hand sides
D, N, rb = 16384, 131072, 256
prob, pis, obsv, weights, xs = bench.make_workload(D, N, rb, 89, 0, 8)
from stochasticdecomposition_b200._abi import Caps
t = api.create(prob, Caps(D + 16, D + 16, D + 16, N + 16, 1))
t.omega_append_bulk(obsv[:N], weights)
t.update_dual_bulk(pis[:D], None, None, -1.0)
t.calc_delta_block(0, 256, 0, 1024)
t0 = time.perf_counter(); t.calc_delta_block(0, D, 0, N); s = time.perf_counter() - t0
print(json.dumps({"kernel": "k_delta_block_rhs", "duals": D, "observations": N, "Rb": rb, "seconds_wall": s, "fp64_ops_per_s": 2.0 * rb * D * N / s,
                  "frac_of_measured_mul_add_peak": 2.0 * rb * D * N / s / mul_add}), flush=True)
t.close()
