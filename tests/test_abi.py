"""CPU-side checks of the drop-in boundary: the CUDA library builds for sm_100a, loads without a GPU, exports
every entry point include/sdgpu.h declares, and refuses to work without a device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

import stochasticdecomposition_b200 as sd
from stochasticdecomposition_b200 import build as sdbuild
from stochasticdecomposition_b200.synthetic import problem_for

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sdgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdgpu_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    lib_path = sdbuild.build()
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in sdgpu.h but not exported: {missing}"


def test_binary_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", sdbuild.build()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sd.SdError, match="no CUDA device"):
        sd.create_tables(problem_for("pgp2"), sd.Caps.like_reference(50, 2))


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(sd.SdError, match="no CPU fallback"):
        sd.load_library(str(tmp_path / "libsdgpu.so"))


def test_dual_stability_tail_matches_oracle():
    """sdgpu_dual_stability is host scalar arithmetic (cuts.c:171-182); it must agree with the checker."""
    import numpy as np
    import oracle_loader
    from stochasticdecomposition_b200._abi import _pf64
    api, orc = sd.load_library(), oracle_loader.oracle()
    rng = np.random.default_rng(5)
    scan = 16
    ra, rb = np.zeros(scan), np.zeros(scan)
    for k in range(1, 80):
        old = rng.uniform(0.9, 1.0) * 100
        allv = 100.0 if k % 17 else 0.0
        fa = api._fn("dual_stability")(old, allv, k, 3, scan, _pf64(ra))
        fb = orc._fn("dual_stability")(old, allv, k, 3, scan, _pf64(rb))
        assert fa == fb
        assert np.array_equal(ra.view(np.int64), rb.view(np.int64))
