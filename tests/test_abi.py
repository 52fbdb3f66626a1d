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


def test_sweep_grid_plan_is_wave_aware():
    """Host-only check of the chunk-count rule ("Basis chunks"): with three resident CTAs per SM a grid must not
    end just past a whole number of waves; chunks cover every basis exactly once and never exceed the scratch limit."""
    import math
    api = sd.load_library()
    slots = 148 * 3
    for N, B in [(131072, 65536), (5000, 5000), (5000, 7500), (1000, 1000), (16384, 16384), (1048576, 4096), (1048576, 16384), (64, 64),
                 (400000, 512), (65536, 8192), (20000, 4096), (700, 2100), (300, 31), (513, 1)]:
        tiles, cs, nc = api.plan_sweep_grid(148, N, B, 64)
        assert tiles == (N + 511) // 512
        assert 1 <= nc <= 64 and cs >= 1 and (nc - 1) * cs < B <= nc * cs          # every basis in exactly one chunk
        assert nc == 1 or cs >= 16                                                # chunks stay long enough to amortise their start-up
        waves = tiles * nc / slots
        best_possible = max((tiles * c / slots) for c in range(1, min(64, max(1, (B + 31) // 32)) + 1))
        if best_possible >= 2.0:                                                  # multi-wave regime: the last wave is at least 60 % full
            assert math.ceil(waves) - waves <= 0.4 + 1e-9 or waves >= 30, (N, B, tiles, nc, waves)
    # the two cases that motivated the rule: 256 tiles must not get 7 chunks (4.04 waves); 10 tiles get one full wave, not 1.44
    assert api.plan_sweep_grid(148, 131072, 65536, 64)[2] != 7
    assert api.plan_sweep_grid(148, 5000, 5000, 64)[2] == 44
    # round 2: many tiles, few bases (one GPU's share of the strong-scaling table): chunks of >= ~600 rows, not 55 chunks of 149 rows
    t, cs, nc = api.plan_sweep_grid(148, 131072, 8192, 64)
    assert 8 <= nc <= 13 and cs >= 600, (cs, nc)
    assert api.plan_sweep_grid(148, 131072, 65536, 64)[2] >= 50                  # the bench table keeps its ~55 chunks of ~1 200 rows


def test_sweep_family_plan():
    """Host-only check of the kernel choice: 1 loads, 2 TMA ring, 3 per-term gathers, 4 term-linear ring,
    5 recompute, 6 grouped ring; and the fused prologue of tiny cuts."""
    api = sd.load_library()
    big, real = (131072, 65536), (5000, 5000)                                   # (observations, bases)
    assert api.plan_sweep_kind(*big) == (2, False)                              # the bench workload: plain TMA ring
    assert api.plan_sweep_kind(*real) == (1, False)                             # a late real-problem cut: LDG, separate prologue (440 CTAs)
    assert api.plan_sweep_kind(1000, 1000) == (1, True)                         # an early one: LDG with the prologue fused
    assert api.plan_sweep_kind(576, 300, Rb=3, n1=4, n1c=4) == (5, True)        # pgp2: recompute, fused
    assert api.plan_sweep_kind(*big, Rb=3) == (5, False) and api.plan_sweep_kind(*big, Rb=5) == (5, False)
    assert api.plan_sweep_kind(*real, Rb=5)[0] == 1 and api.plan_sweep_kind(*big, Rb=6)[0] == 2
    assert api.plan_sweep_kind(*big, distinct_rows=30000)[0] == 6               # several bases per lambda row: grouped ring
    assert api.plan_sweep_kind(*big, distinct_rows=60000)[0] == 2               # less than 15 % to gain: plain ring
    assert api.plan_sweep_kind(*real, distinct_rows=2500)[0] == 1               # below L2 size grouping buys nothing
    assert api.plan_sweep_kind(*real, Q=8)[0] == 2 and api.plan_sweep_kind(300, 300, Q=2)[0] == 1
    rc = dict(rvd=4, cost_cols=4)
    assert api.plan_sweep_kind(5000, 2000, max_phi=2, terms=6000, **rc)[0] == 4          # multi-term: the ring from ~4M pairs
    assert api.plan_sweep_kind(300, 200, max_phi=2, terms=600, **rc)[0] == 3
    assert api.plan_sweep_kind(*real, **rc)[0] == 1 and api.plan_sweep_kind(65536, 6144, **rc)[0] == 4   # mask only: LDG until ~50M pairs
    assert api.plan_sweep_kind(5000, 2000, max_phi=2, terms=6000, rvd=40, cost_cols=60)[0] == 3          # cost block does not fit shared memory
    for v, want in ((1, 1), (2, 2), (3, 1), (4, 6)):                            # forced families (3: Rb = 10 rules the recompute sweep out)
        assert api.plan_sweep_kind(*real, variant=v)[0] == want
