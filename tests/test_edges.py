"""Edge cases (tests/edge_cases.py): CPU -- the port against the reference build, bit-exact; GPU -- the CUDA library against
the port (indices exact, coefficients 1e-9)."""
import pytest

import edge_cases
import oracle_loader


@pytest.mark.skipif(not oracle_loader.have_reference(), reason="reference build unavailable")
@pytest.mark.parametrize("name", sorted(set(edge_cases.CASES) - {"capacity_edges", "tile_boundaries", "no_observations"}))
def test_port_matches_reference(name):
    # (the reference never bounds-checks and has no bulk loaders, so the capacity / bulk cases are port-vs-GPU only)
    f = edge_cases.CASES[name]
    edge_cases.same(f(oracle_loader.reference()), f(oracle_loader.oracle()), exact=name != "two_contexts" or True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(edge_cases.CASES))
def test_cuda_matches_port(name):
    import stochasticdecomposition_b200 as sd
    f = edge_cases.CASES[name]
    edge_cases.same(f(oracle_loader.oracle()), f(sd.load_library()), exact=False)
