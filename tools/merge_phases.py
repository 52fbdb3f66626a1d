#!/usr/bin/env python
"""Phase breakdown of k_cut_merge at real-problem sizes (debug build of the library with -DSD_PHASE_CLOCKS, written next to the
product library as libsdgpu_phase.so; never shipped).  Prints global-timer deltas in microseconds."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import stochasticdecomposition_b200 as sd  # noqa: E402
from stochasticdecomposition_b200 import build as B  # noqa: E402

LIB = os.path.join(B.HERE, "libsdgpu_phase.so")


def build():
    cmd = [B.nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-DSD_PHASE_CLOCKS",
           "-Xcompiler", "-fPIC,-O2,-fvisibility=default", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + B.CSRC]
    cmd += [os.path.join(B.CSRC, s) for s in B.SOURCES] + ["-o", LIB, "-ldl"]
    subprocess.run(cmd, check=True)


NAMES = {0: "start", 10: "old-window chunks merged", 11: "both windows merged, iStar known", 1: "alpha gathers done", 2: "4 block sums",
         3: "delta.piC sums", 4: "sigma.piC walk", 5: "last block elected", 6: "tile partials summed", 7: "scatter into beta", 8: "stored"}
ORDER = [0, 10, 11, 1, 2, 3, 4, 5, 6, 7, 8]

if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
        sys.exit(0)
    if "--lib" in sys.argv:
        LIB = sys.argv[sys.argv.index("--lib") + 1]
    api = sd.load_library(LIB)
    fn = api._fn("debug_phase_clocks")
    shapes = ((1000, 1000, 86, 89), (5000, 5000, 86, 89))
    if "--full" in sys.argv:                                  # the strong-scaling per-GPU shape and the headline shape
        shapes = ((8192, 131072, 256, 89), (65536, 131072, 256, 89))
    if "--strong1" in sys.argv:                               # the one-GPU shape of the strong-scaling table
        shapes = ((8192, 1048576, 256, 89),)
    for D, N, rv, n1 in shapes:
        prob, pis, obsv, weights, xs = bench.make_workload(D, N, rv, n1, 0, 8)
        k = int(weights.sum())
        t = bench.load_tables(api, prob, pis, obsv, weights, D, N, k, 8)
        t.set_timing(True)
        acc = []
        for s in range(12):
            t.sd_cut(xs[s], k, 1, 0.0)
            buf = (C.c_longlong * 16)()
            assert fn(buf) == 0
            if s >= 4:
                acc.append([buf[i] for i in range(16)])
        a = np.median(np.array(acc, dtype=np.float64) - np.array(acc, dtype=np.float64)[:, :1], axis=0)
        out = {"D": D, "N": N, "cut_dev_us": round(t.stats()["last_cut_ms"] * 1e3, 1), "sweep_us": round(t.stats()["last_sweep_ms"] * 1e3, 1),
               "merge_event_us": round(t.stats().get("last_merge_ms", 0.0) * 1e3, 1)}
        prev = 0.0
        for i in ORDER:
            out[NAMES[i]] = round((a[i] - prev) / 1e3, 2)
            prev = a[i]
        print(json.dumps(out), flush=True)
        t.close()
