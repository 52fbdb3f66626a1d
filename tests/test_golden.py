"""Golden vectors produced by the reference build (tests/golden/make_golden.py).  Without a GPU the CPU oracle is
checked against them (bit-exact); with a GPU (-m gpu) the CUDA library is (indices and tables bit-exact, cut
coefficients within 1e-9 relative)."""
import glob
import os

import numpy as np
import pytest

import oracle_loader
from replay import dump_tables, replay
from stochasticdecomposition_b200._abi import Caps
from stochasticdecomposition_b200.synthetic import Trace, make_problem

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _run(api, path, sweep_variant=None):
    g = np.load(path)
    pk = {k: int(v) for k, v in zip(g["pk_keys"], g["pk_vals"])}
    prob = make_problem(int(g["problem_seed"]), **pk)
    K, phi_len = int(g["K"]), int(g["phi_len"])
    rk = {k: (float(v) if k in ("lb", "feas_density") else int(v)) for k, v in zip(g["rk_keys"], g["rk_vals"])}
    trace = Trace(g["observ"], g["duals"], g["mubBar"], g["xs"], g["two_solves"],
                  g["phi"] if "phi" in g else None, g["phi_omega"] if "phi_omega" in g else None)
    n = 2 * K * (1 + phi_len) + 2
    rec = replay(api, prob, trace, Caps(n, n, 2 * K + 2, K + 1, 1 + phi_len), sweep_variant=sweep_variant, **rk)
    return g, rec


def _check(g, rec, exact):
    assert rec.omega_idx == g["omega_idx"].tolist() and rec.omega_new == g["omega_new"].tolist()
    assert rec.basis_idx == g["basis_idx"].tolist() and rec.basis_new == g["basis_new"].tolist()
    assert [c is None for c in rec.cuts] == g["cut_null"].tolist()
    tab = dump_tables(rec.tables)
    bits = lambda a: np.ascontiguousarray(a, np.float64).view(np.int64)
    assert np.array_equal(bits(np.array(tab["lambda"])), bits(g["lambda"]))
    assert np.array_equal(bits(np.array([s[0] for s in tab["sigma"]])), bits(g["sigma_pib"]))
    assert np.array_equal(bits(np.array([s[1] for s in tab["sigma"]])), bits(g["sigma_piC"]))
    assert [s[2] for s in tab["sigma"]] == g["sigma_lam"].tolist() and [s[3] for s in tab["sigma"]] == g["sigma_ck"].tolist()
    assert np.array_equal(bits(tab["delta_pib"]), bits(g["delta_pib"])) and np.array_equal(bits(tab["delta_piC"]), bits(g["delta_piC"]))
    assert [o[1] for o in tab["omega"]] == g["omega_w"].tolist()
    for n, c in enumerate(rec.cuts):
        if c is None:
            continue
        assert np.array_equal(c.iStar, g[f"cut{n}_istar"]), f"cut {n} iStar"
        a, b = float(g[f"cut{n}_alpha"]), g[f"cut{n}_beta"]
        ratio = c.cummOld / c.cummAll if c.cummAll != 0 else np.nan
        if exact:
            assert np.float64(c.alpha).tobytes() == np.float64(a).tobytes() and np.array_equal(bits(c.beta), bits(b))
            assert np.float64(ratio).tobytes() == g[f"cut{n}_ratio"].tobytes() or (np.isnan(ratio) and np.isnan(g[f"cut{n}_ratio"]))
        else:
            scale = max(abs(a), np.abs(b[1:]).max())
            assert abs(c.alpha - a) <= 1e-9 * abs(a) and np.abs(c.beta - b).max() <= 1e-9 * scale
            gr = float(g[f"cut{n}_ratio"])
            assert (np.isnan(ratio) and np.isnan(gr)) or abs(ratio - gr) <= 1e-9 * max(abs(gr), 1e-300)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_reference_vectors(path):
    g, rec = _run(oracle_loader.oracle(), path)
    _check(g, rec, exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4], ids=["auto", "loads", "tma_rings", "recompute", "grouped"])
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_cuda_reproduces_reference_vectors(path, variant):
    """Every sweep family against the vectors the reference build produced (a forced family falls back to the load-based kernels
    where the problem's shape rules it out, e.g. the recompute sweep with random T elements)."""
    import stochasticdecomposition_b200 as sd
    g, rec = _run(sd.load_library(), path, sweep_variant=variant)
    _check(g, rec, exact=False)


def test_fixtures_exist():
    assert len(FILES) >= 4
