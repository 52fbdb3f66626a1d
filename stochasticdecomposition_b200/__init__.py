"""stochasticdecomposition_b200 -- B200-native cut formation for two-stage Stochastic Decomposition.

The product is the C-ABI CUDA library `libsdgpu.so` (include/sdgpu.h).  This package is its Python host
side: a ctypes binding whose method names mirror the reference's stocUpdate.c / cuts.c entry points, the
synthetic workload generator, and the observation-sharding helpers for multi-GPU runs.

There is no CPU fallback: `load_library()` raises if the CUDA library has not been built, and
`sdgpu_create` fails without a CUDA device.
"""
from __future__ import annotations

import ctypes
import os

from ._abi import Api, Caps, Cut, Problem, SdError, Tables  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdgpu.so")
_api = None


def load_library(path: str | None = None) -> Api:
    """Bind libsdgpu.so.  Fails loudly when the extension is missing -- nothing here computes on the CPU."""
    global _api
    if _api is not None and path is None:
        return _api
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise SdError(f"{p} not found: build it with `python -m stochasticdecomposition_b200.build` "
                      "(nvcc, sm_100a).  There is no CPU fallback.")
    api = Api(ctypes.CDLL(p), "sdgpu_")
    if api._fn("abi_version")() != 2:
        raise SdError("libsdgpu.so ABI version mismatch")
    if path is None:
        _api = api
    return api


def create_tables(problem: Problem, caps: Caps, device: int = 0) -> Tables:
    """newLambda/newSigma/newDelta/newOmega/newBasisType of setup.c:140-144, on `device`."""
    return load_library().create(problem, caps, device)
