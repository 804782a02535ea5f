"""CPU-side checks of the drop-in boundary: the shared library loads, exports
every symbol include/mdbg_b200.h declares, and refuses to run without a GPU
(no CPU fallback).  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from metamdbg_b200 import _capi
    return _capi.load()


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mdbg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mdbg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    from metamdbg_b200 import _capi
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mdbg_b200.h but not exported"
    assert sorted(_capi.SYMBOLS) == names       # the ctypes table covers exactly the header


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "metamdbg_b200", "libmdbg_b200.so")],
                         capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_product_does_not_link_or_import_oracle():
    so = os.path.join(ROOT, "metamdbg_b200", "libmdbg_b200.so")
    ldd = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "mdbg_ref" not in ldd
    for dirpath, _, files in os.walk(os.path.join(ROOT, "metamdbg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "mdbg_oracle" not in txt and "libmdbg_ref" not in txt, f


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from metamdbg_b200 import Engine, MdbgError
    with pytest.raises(MdbgError) as e:
        Engine(15, 0.005, True)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_bad_params_rejected(lib):
    from metamdbg_b200 import _capi
    p = _capi.MdbgParams(17, 0.005, 1, None, 0)
    ctx = C.c_void_p()
    assert lib.mdbg_ctx_create(0, C.byref(p), C.byref(ctx)) != 0
    assert b"minimizer_size" in lib.mdbg_last_error(None)


def test_host_packer_cpp(tmp_path):
    """The host-side 2-bit packer (transfer compression of host batches) is plain C++: unit-tested on the CPU."""
    import __graft_entry__ as g
    g.build()
    exe = tmp_path / "pack_host_test"
    obj = os.path.join(ROOT, "metamdbg_b200", "csrc", "pack_host.o")
    assert os.path.exists(obj)
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-pthread", "-o", str(exe),
                    os.path.join(ROOT, "tests", "cpp", "pack_host_test.cpp"), obj], check=True)
    seen = set()
    for isa in ("scalar", "avx2", "avx512"):       # the request only lowers what the CPU supports
        out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120,
                             env=dict(os.environ, MDBG_PACK_ISA=isa))
        assert out.returncode == 0, out.stdout + out.stderr
        seen.add(out.stdout.split()[-1])
    assert "scalar" in seen


@pytest.mark.ref
def test_host_glue_scalars_vs_reference(tmp_path):
    """mdbg_host.hpp's read_stats scalars (N50, mean length, purge lastK) against the reference's own
    Utils::computeN50 / computeMeanLength / Commons::computeLastK, and the file helper's error paths."""
    from oracle import pyoracle
    pyoracle.build()
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "libmdbg_ref.so")):
        pytest.skip("oracle/_ref not built")
    exe = tmp_path / "host_glue_test"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "metamdbg_b200", "host"), "-o", str(exe),
                    os.path.join(ROOT, "tests", "cpp", "host_glue_test.cpp"), "-L" + ref_dir, "-lmdbg_ref",
                    "-Wl,-rpath," + ref_dir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
