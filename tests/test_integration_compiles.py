"""The host patch (integration/sdgpu_hooks.c) must compile against the reference's own headers: it is the code a
maintainer pastes into twoSD_src/.  Only possible where /root/reference exists (not on the GPU box)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/twoSD_src"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present")
def test_hooks_compile_against_reference_headers():
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cmd = [cc, "-std=gnu99", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types", "-fsyntax-only",
           "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + REF, "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "integration", "sdgpu_hooks.c")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
