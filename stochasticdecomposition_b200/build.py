"""Builds libsdgpu.so in-tree with nvcc for sm_100a (cross-compiles on a box without a GPU).

    python -m stochasticdecomposition_b200.build            # build if sources are newer than the library
    python -m stochasticdecomposition_b200.build --force

-fmad=false: the table entries and the argmax scores must be rounded exactly like the reference's scalar C
(one rounding per multiply and per add), otherwise iStar is not bit-exact.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsdgpu.so")
SOURCES = ["tables.cu", "cut.cu", "feaspool.cu", "nccl_glue.cu", "group.cu", "peaks.cu", "vmem.cu"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "sdgpu.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
           "-ccbin", host_cxx, "-Xcompiler", "-fPIC,-O2,-fvisibility=default", "-shared",
           "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB + ".tmp", "-ldl"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out.stdout + out.stderr)
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        print(out.stdout + out.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
