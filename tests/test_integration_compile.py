"""INTEGRATION.md's reference-side bindings compile against metaMDBG's own headers (compile only; needs
/root/reference, so it is skipped on the GPU box)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"


@pytest.mark.ref
def test_reference_side_bindings_compile(tmp_path):
    if not os.path.exists(os.path.join(REF, "Commons.hpp")):
        pytest.skip("reference sources not present")
    out = subprocess.run(["/usr/bin/g++", "-std=gnu++20", "-fopenmp", "-w", "-c", "-I" + REF, "-I" + os.path.join(ROOT, "include"),
                          "-I" + os.path.join(ROOT, "metamdbg_b200", "host"),
                          os.path.join(ROOT, "tests", "cpp", "integration_binding.cpp"), "-o", str(tmp_path / "binding.o")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
