#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE BUILD (oracle/_ref/libsdref.so = the reference's own
stocUpdate.c / cuts.c / optimal.c compiled from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Each file stores the inputs (problem, trace, replay options) and what the reference produced: every
find-or-append index and flag, every cut (alpha, beta, iStar, pi_ratio) and the final tables.  The GPU box has
no /root/reference; there the fixtures are the reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))

import oracle_loader  # noqa: E402
from replay import dump_tables, replay  # noqa: E402
from stochasticdecomposition_b200._abi import Caps  # noqa: E402
from stochasticdecomposition_b200.synthetic import make_problem, make_trace  # noqa: E402

CASES = {
    "pgp2_like": (dict(rows=7, cols=16, n1=4, n1c=4, R=3, Rb=3), 50, 6, 9, 0, {}),
    "rhs_T": (dict(rows=30, cols=40, n1=12, n1c=9, R=11, Rb=8, Q=5), 36, 12, 20, 0, dict(lb=-1.25)),
    "random_cost": (dict(rows=24, cols=30, n1=9, n1c=7, R=10, Rb=7, Q=2, rvd=3), 30, 6, 0, 2, dict(feas_density=0.6)),
    "ssn_like": (dict(rows=175, cols=706, n1=89, n1c=89, R=86, Rb=86), 28, 10, 0, 0, {}),
}


def caps_for(K, phi_len):
    n = 2 * K * (1 + phi_len) + 2
    return Caps(n, n, 2 * K + 2, K + 1, 1 + phi_len)


def main():
    ref = oracle_loader.reference()
    for name, (pk, K, dpool, opool, phi_len, rk) in CASES.items():
        prob = make_problem(4000 + len(name), **pk)
        trace = make_trace(prob, K, seed=500 + K, dual_pool=dpool, obs_pool=opool, phi_len=phi_len)
        rec = replay(ref, prob, trace, caps_for(K, phi_len), **rk)
        tab = dump_tables(rec.tables)
        out = {"K": K, "phi_len": phi_len, "pk_keys": np.array(list(pk.keys())), "pk_vals": np.array(list(pk.values()), dtype=np.int64),
               "problem_seed": 4000 + len(name), "trace_seed": 500 + K, "dual_pool": dpool, "obs_pool": opool,
               "rk_keys": np.array(list(rk.keys())), "rk_vals": np.array(list(rk.values()), dtype=np.float64),
               "observ": trace.observ, "duals": trace.duals, "mubBar": trace.mubBar, "xs": trace.xs, "two_solves": trace.two_solves,
               "omega_idx": np.array(rec.omega_idx), "omega_new": np.array(rec.omega_new), "basis_idx": np.array(rec.basis_idx),
               "basis_new": np.array(rec.basis_new), "cut_null": np.array([c is None for c in rec.cuts]),
               "lambda": np.array(tab["lambda"]), "sigma_pib": np.array([s[0] for s in tab["sigma"]]),
               "sigma_piC": np.array([s[1] for s in tab["sigma"]]), "sigma_lam": np.array([s[2] for s in tab["sigma"]]),
               "sigma_ck": np.array([s[3] for s in tab["sigma"]]), "delta_pib": tab["delta_pib"], "delta_piC": tab["delta_piC"],
               "omega_vals": np.array([o[0] for o in tab["omega"]]), "omega_w": np.array([o[1] for o in tab["omega"]])}
        if trace.phi is not None:
            out["phi"], out["phi_omega"] = trace.phi, trace.phi_omega
        for n, c in enumerate(rec.cuts):
            if c is None:
                continue
            out[f"cut{n}_alpha"] = np.float64(c.alpha); out[f"cut{n}_beta"] = c.beta; out[f"cut{n}_istar"] = c.iStar
            out[f"cut{n}_ratio"] = np.float64(c.cummOld / c.cummAll if c.cummAll != 0 else np.nan)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB", "cuts", len(rec.cuts))


if __name__ == "__main__":
    main()
