"""The C ABI from plain C (tests/c/host_demo.c), as the reference's C host would call it: one binary linked against
libsdgpu.so, one against the CPU oracle, transcripts compared.  The CPU half also checks that the header compiles as C99."""
import os
import subprocess

import pytest

import oracle_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "host_demo.c")
CC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


def _build(tmp, prefix, libdir, libname):
    exe = os.path.join(tmp, f"host_demo_{prefix}")
    cmd = [CC, "-std=c99", "-O1", "-Wall", "-Werror", f"-DSD_PREFIX={prefix}", "-I" + os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L" + libdir, "-l" + libname, "-Wl,-rpath," + libdir, "-lm"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def _run(exe):
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_header_is_c99_and_oracle_transcript_is_sane(tmp_path):
    oracle_loader.oracle()
    lines = _run(_build(str(tmp_path), "sdo_", os.path.join(ROOT, "oracle"), "sdoracle"))
    assert lines[-1].startswith("counts ") and len(lines) > 50
    assert any("o=" in ln and "*" not in ln.split(" l=")[0] for ln in lines[:-1])      # some observation was a repeat


@pytest.mark.gpu
def test_c_host_gpu_matches_oracle(tmp_path):
    from stochasticdecomposition_b200 import build as sdbuild
    oracle_loader.oracle()
    sdbuild.build()
    cpu = _run(_build(str(tmp_path), "sdo_", os.path.join(ROOT, "oracle"), "sdoracle"))
    gpu = _run(_build(str(tmp_path), "sdgpu_", os.path.join(ROOT, "stochasticdecomposition_b200"), "sdgpu"))
    assert len(cpu) == len(gpu)
    for a, b in zip(cpu, gpu):
        ha, hb = a.split(" alpha=")[0], b.split(" alpha=")[0]
        assert ha == hb, (a, b)                                  # iteration, indices, flags, omegaCnt, iStar checksum: exact
        if " alpha=" in a:
            va = [float(t) for t in a.split(" alpha=")[1].replace("|", " ").split()]
            vb = [float(t) for t in b.split(" alpha=")[1].replace("|", " ").split()]
            scale = max(abs(v) for v in va[:-2]) or 1.0
            assert all(abs(p - q) <= 1e-9 * max(scale, abs(p)) for p, q in zip(va, vb)), (a, b)
