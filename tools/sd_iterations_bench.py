#!/usr/bin/env python
"""SD iterations per second on an ssn-shaped instance (BASELINE.json metric 3), GPU tables vs the CPU reference path,
same host loop, same seed.  HiGHS solves the subproblem LP (the reference uses CPLEX); the instance is synthetic with
ssn's dimensions (n1 = 89, 175 rows, 86 random right-hand sides) because the SMPS file is not available offline.
Prints one JSON line per backend."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))

import oracle_loader  # noqa: E402
from sd_highs_host import SDHost, caps_for, make_slp  # noqa: E402


def run(api, label, shape, K, seed):
    slp = make_slp(shape)
    host = SDHost(slp, api.create(slp.problem(), caps_for(K)), seed=seed, check_lp_identity=False)
    t0 = time.perf_counter()
    marks = []
    for chunk in range(0, K, max(1, K // 5)):
        n = min(max(1, K // 5), K - chunk)
        a0, t1 = host.stats.argmax_seconds, time.perf_counter()
        host.run(n)
        marks.append({"k": host.k, "it_per_s": round(n / (time.perf_counter() - t1), 2),
                      "argmax_share": round((host.stats.argmax_seconds - a0) / (time.perf_counter() - t1), 4)})
    st = host.stats
    total = time.perf_counter() - t0
    c = host.t.counts()
    return {"backend": label, "shape": shape, "iterations": K, "seconds": round(total, 3), "iterations_per_s": round(K / total, 3),
            "argmax_seconds": round(st.argmax_seconds, 3), "subproblem_lp_seconds": round(st.subprob_seconds, 3),
            "master_seconds": round(st.master_seconds, 3), "argmax_share": round(st.argmax_seconds / total, 4),
            "lp_solves": st.lp_solves, "incumbent_estimate": st.incumb_est, "tables": c, "by_segment": marks,
            "note": "HiGHS LP (not CPLEX), synthetic ssn-shaped instance, Python host loop"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="ssn")
    ap.add_argument("--iterations", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--backends", default="gpu,reference")
    args = ap.parse_args()
    for b in args.backends.split(","):
        if b == "gpu":
            import stochasticdecomposition_b200 as sd
            api = sd.load_library()
        elif b == "reference":
            api = oracle_loader.reference()
        else:
            api = oracle_loader.oracle()
        print(json.dumps(run(api, b, args.shape, args.iterations, args.seed)), flush=True)
