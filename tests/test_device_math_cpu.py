"""The sketch kernel's pure arithmetic, run on the CPU.

metamdbg_b200/csrc/common.cuh and bitmath.cuh are __host__ __device__: tests/cpp/device_math_test.cu compiles them
as host code with nvcc and checks (1) the 64-bit and 128-bit Murmur arithmetic against the oracle, (2) the
SIMD-in-register bit tricks against naive loops (all 65 536 keep masks, all 256 4-base words), (3) the 16-position
register roll against the oracle's l-mer iterator and (4) that the high-word candidate test never rejects a key the
exact `(double)hash < bound` test selects.  `MDBG_EXHAUSTIVE=1` walks all 2^32 keys at the two production densities
(~100 core-seconds; result recorded in DESIGN.md), the default samples every 257th key."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"),
                    reason="nvcc not available")
@pytest.mark.parametrize("variant", [0, 1])
def test_device_math_on_host(tmp_path, variant):
    from oracle import pyoracle
    pyoracle.build()
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "device_math_test")
    odir = os.path.join(ROOT, "oracle")
    cmd = [nvcc, "-O2", "-std=c++17", "-w", f"-DMDBG_TEST_VARIANT={variant}", "-Xcompiler", "-pthread", "-o", exe,
           os.path.join(ROOT, "tests", "cpp", "device_math_test.cu"), "-L" + odir, "-lmdbg_oracle",
           "-Xlinker", "-rpath=" + odir]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    args = [exe] + (["--exhaustive"] if os.environ.get("MDBG_EXHAUSTIVE") == "1" else [])
    run = subprocess.run(args, capture_output=True, text=True, timeout=3000)
    assert run.returncode == 0, run.stdout[-3000:]
    assert run.stdout.strip().endswith("OK")
    assert "misclassified 0" in run.stdout
