"""A random-cost scenario that exercises checkBasisFeasibility (randCost.c:202-258) through the C ABI: bases carry piDet /
phi / gBar / psi / cstat, the mask is computed by the library for new observations (all bases) and new bases (all
observations), and cuts are then formed under that mask."""
import numpy as np

from stochasticdecomposition_b200._abi import Caps
from stochasticdecomposition_b200.synthetic import make_problem


def run(api, seed=5, n_obs=40, n_basis=30, rvd=3):
    rng = np.random.default_rng(seed)
    prob = make_problem(77, rows=18, cols=26, n1=7, n1c=6, R=8, Rb=6, Q=2, rvd=rvd)
    rows, cols = prob.rows, prob.cols
    n = 4 * n_basis + 8
    t = api.create(prob, Caps(n, n, n_basis + 2, n_obs + 2, 1 + rvd))
    rvdOmCols = np.concatenate([[0], np.sort(rng.choice(np.arange(1, cols + 1), rvd, replace=False))]).astype(np.int32)
    senx = bytes(rng.choice([ord("G"), ord("L"), ord("E")], rows).tolist())
    t.set_cost_coords(rvdOmCols, senx)
    out = {"obs_flags": [], "basis_flags": [], "cuts": [], "basis_idx": []}
    obs = rng.normal(0, 1.0, (n_obs, prob.numRV + 1)); obs[:, 0] = 0
    obs[:, prob.rvOffset[2] + 1:] *= 0.15                      # cost deltas: small, so that feasibility is mixed
    nb = 0
    for it in range(n_obs):
        k = it + 1
        oi, onew = t.calc_omega(obs[it], 1e-3)
        if onew:
            t.calc_delta(True, oi)
            if nb:
                out["obs_flags"].append(t.check_feasibility_obs(oi, 1e-3).copy())
        if it % 4 != 3 and nb < n_basis:
            phi_len = int(rng.integers(0, rvd + 1))
            pi = rng.uniform(-1, 1, rows + 1) * (rng.random(rows + 1) < 0.6); pi[0] = 0
            li, nl, s0, ns = t.update_dual(pi, 0.0, k, 1e-3)
            sig, phis = [s0], []
            om = [0] + sorted(rng.choice(np.arange(1, rvd + 1), phi_len, replace=False).tolist())
            for c in range(phi_len):
                col = rng.uniform(-0.5, 0.5, rows + 1) * (rng.random(rows + 1) < 0.4); col[0] = 0
                phis.append(col)
                sig.append(t.update_dual(col, 0.0, k, 1e-3)[2])
            bi = t.basis_append(k, True, sig, om if phi_len else None)
            sense = np.frombuffer(senx, np.uint8)
            mag = np.abs(rng.normal(0, 0.12, rows))
            piDet = np.concatenate([[0.0], np.where(sense == ord("G"), mag, np.where(sense == ord("L"), -mag, rng.normal(0, 0.3, rows)))])
            gBar = np.abs(rng.normal(0.25, 0.1, cols + 1)) + 0.02; gBar[0] = 0
            psi = rng.uniform(-0.5, 0.5, (cols, phi_len)) * (rng.random((cols, phi_len)) < 0.5)
            cstat = np.concatenate([[0], rng.integers(0, 3, cols)]).astype(np.int32)
            neg = rng.random(cols + 1) < 0.05                      # a few negative reduced costs, legal only at upper bound
            gBar[neg] *= -1; cstat[neg & (rng.random(cols + 1) < 0.8)] = 2; gBar[0] = 0
            t.basis_set_feas_data(bi, piDet, np.array(phis) if phi_len else None, gBar, psi.ravel() if phi_len else None, cstat)
            out["basis_flags"].append(t.check_feasibility_basis(bi, 1e-3).copy())
            out["basis_idx"].append(bi)
            nb += 1
        if nb:
            x = rng.uniform(0, 1, prob.prevCols + 1); x[0] = 0
            out["cuts"].append(t.sd_cut(x, k, k % 2, 0.0))
    out["tables"] = t
    return out


def compare(a, b, exact, rtol=1e-9):
    assert a["basis_idx"] == b["basis_idx"]
    assert len(a["obs_flags"]) == len(b["obs_flags"]) and len(a["basis_flags"]) == len(b["basis_flags"])
    for u, v in zip(a["obs_flags"], b["obs_flags"]):
        assert np.array_equal(u, v)
    for u, v in zip(a["basis_flags"], b["basis_flags"]):
        assert np.array_equal(u, v)
    mixed = np.concatenate(a["basis_flags"])
    assert 0.05 < mixed.mean() < 0.95, f"feasibility is not mixed ({mixed.mean():.2f}); the scenario does not test anything"
    assert len(a["cuts"]) == len(b["cuts"])
    some = 0
    for ca, cb in zip(a["cuts"], b["cuts"]):
        assert (ca is None) == (cb is None)
        if ca is None:
            continue
        some += 1
        assert np.array_equal(ca.iStar, cb.iStar)
        if exact:
            assert ca.alpha == cb.alpha and np.array_equal(ca.beta, cb.beta)
        else:
            scale = max(abs(ca.alpha), np.abs(ca.beta[1:]).max())
            assert abs(ca.alpha - cb.alpha) <= rtol * max(abs(ca.alpha), 1e-300) and np.abs(ca.beta - cb.beta).max() <= rtol * scale
    assert some > 3
