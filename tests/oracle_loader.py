"""Loads the CPU checkers (TEST INFRASTRUCTURE).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / reference legs import this module; the product package never does.

  oracle()    -> Api over oracle/libsdoracle.so      (restated port, prefix sdo_)
  reference() -> Api over oracle/_ref/libsdref.so    (the reference's own .c files + shim, prefix sdref_)
"""
from __future__ import annotations

import ctypes
import functools
import os
import subprocess

from stochasticdecomposition_b200._abi import Api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REFERENCE_SRC = "/root/reference/twoSD_src"


def build(quiet: bool = True) -> None:
    """(Re)build the checkers: the port always, the reference build only where /root/reference exists."""
    out = subprocess.run(["make", "-C", ORACLE_DIR, "oracle", "ref", "hooks"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


@functools.lru_cache(maxsize=None)
def oracle() -> Api:
    path = os.path.join(ORACLE_DIR, "libsdoracle.so")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(ORACLE_DIR, "sd_oracle.c")):
        build()
    return Api(ctypes.CDLL(path), "sdo_")


def have_reference() -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libsdref.so")) or os.path.isdir(REFERENCE_SRC)


@functools.lru_cache(maxsize=None)
def reference() -> Api:
    path = os.path.join(ORACLE_DIR, "_ref", "libsdref.so")
    if os.path.isdir(REFERENCE_SRC):
        build()
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return Api(ctypes.CDLL(path), "sdref_")
