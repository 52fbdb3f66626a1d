"""The host patch must apply to the reference as it lies under /root/reference and the patched tree must compile: it is what a
maintainer gets.  (tests/test_host_patch.py goes further and RUNS the patched stocUpdate.c / cuts.c / optimal.c.)  Only possible
where /root/reference exists (not on the GPU box)."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/twoSD_src"
PATCH = os.path.join(ROOT, "integration", "twoSD_src.patch")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present")


@pytest.fixture(scope="module")
def patched_tree():
    work = tempfile.mkdtemp(prefix="sdgpu_patch_test.")
    shutil.copytree(REF, os.path.join(work, "twoSD_src"))
    out = subprocess.run(["patch", "-p1", "-i", PATCH], cwd=work, capture_output=True, text=True)
    assert out.returncode == 0 and "FAILED" not in out.stdout and "fuzz" not in out.stdout, out.stdout + out.stderr
    shutil.copy(os.path.join(ROOT, "integration", "sdgpu_hooks.c"), os.path.join(work, "twoSD_src"))
    yield os.path.join(work, "twoSD_src")
    shutil.rmtree(work, ignore_errors=True)


def _cc(tree, name, extra):
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cmd = [cc, "-std=gnu99", "-fsyntax-only", "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + tree, "-I" + os.path.join(ROOT, "include")] + extra + [os.path.join(tree, name)]
    return subprocess.run(cmd, capture_output=True, text=True)


@pytest.mark.parametrize("name", ["sdgpu_hooks.c", "stocUpdate.c", "cuts.c", "optimal.c", "randCost.c"])
def test_patched_hot_path_sources_compile(patched_tree, name):
    out = _cc(patched_tree, name, ["-w", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types", "-Werror=int-conversion"])
    assert out.returncode == 0, out.stderr


@pytest.mark.parametrize("name", ["setup.c", "algo.c", "subprob.c", "soln.c"])
def test_patched_call_sites_type_check(patched_tree, name):
    """these files also call spAlgorithms functions the header shim does not declare (meanProblem, setupProblem, ...), which is the
    unpatched files' situation too; what must hold is that every call the patch touched passes the right types"""
    out = _cc(patched_tree, name, ["-w", "-Werror=incompatible-pointer-types", "-Werror=int-conversion"])
    assert out.returncode == 0, out.stderr


def test_committed_patch_is_what_the_generator_writes():
    """integration/twoSD_src.patch is generated (integration/make_patch.py); the committed file must be current"""
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "integration", "make_patch.py"), "--no-write", "--keep", "/tmp/sdgpu_patch_regen"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    diff = subprocess.run(["diff", "-u", "-r", "a/twoSD_src", "b/twoSD_src"], cwd="/tmp/sdgpu_patch_regen", capture_output=True, text=True).stdout
    regen = [ln.split("\t")[0] + "\n" if ln.startswith(("--- ", "+++ ")) else ln for ln in diff.splitlines(keepends=True) if not ln.startswith("diff -u -r")]
    shutil.rmtree("/tmp/sdgpu_patch_regen", ignore_errors=True)
    assert regen == open(PATCH).readlines()
