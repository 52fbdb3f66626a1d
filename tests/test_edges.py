"""Edge cases (tests/edge_cases.py): CPU -- the port against the reference build, bit-exact; GPU -- the CUDA library against
the port (indices exact, coefficients 1e-9)."""
import pytest

import edge_cases
import oracle_loader
from replay import ForcedVariant


@pytest.mark.skipif(not oracle_loader.have_reference(), reason="reference build unavailable")
@pytest.mark.parametrize("name", sorted(set(edge_cases.CASES) - {"capacity_edges", "tile_boundaries", "no_observations"}))
def test_port_matches_reference(name):
    # (the reference never bounds-checks and has no bulk loaders, so the capacity / bulk cases are port-vs-GPU only)
    f = edge_cases.CASES[name]
    edge_cases.same(f(oracle_loader.reference()), f(oracle_loader.oracle()), exact=name != "two_contexts" or True)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 2, 3, 4], ids=["auto", "tma_rings", "recompute", "grouped"])
@pytest.mark.parametrize("name", sorted(edge_cases.CASES))
def test_cuda_matches_port(name, variant):
    import stochasticdecomposition_b200 as sd
    f = edge_cases.CASES[name]
    api = sd.load_library() if variant == 0 else ForcedVariant(sd.load_library(), variant)
    edge_cases.same(f(oracle_loader.oracle()), f(api), exact=False)
