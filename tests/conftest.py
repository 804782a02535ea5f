import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# tests/test_capi_emulated_cpu.py re-runs the GPU parity tests against the CPU-emulated build of the C-ABI library
# (tests/_emu.py: api.cu + the kernels compiled by g++ against tests/cpp/warp_emu.hpp).  The product never looks at
# this variable; it only redirects the ctypes loader inside that child pytest process.
EMULATED_LIB = os.environ.get("MDBG_EMU_LIB")
if EMULATED_LIB:
    from metamdbg_b200 import _capi
    _capi.LIB_PATH = EMULATED_LIB


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libmdbg_ref.so (reference sources compiled)")


def _have_device() -> bool:
    """A context can be created: a CUDA device (or the CPU emulator library of tests/test_capi_emulated_cpu.py)."""
    if EMULATED_LIB:
        return True
    try:
        from metamdbg_b200 import Engine
        Engine(15, 0.005, True).close()
        return True
    except Exception:                                   # noqa: BLE001  -- no library, no driver, no device
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing in the first of them
    (`-m "not gpu"` deselects them anyway; on the GPU box nothing is skipped)."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _have_device():
        return
    skip = pytest.mark.skip(reason="no CUDA device: libmdbg_b200 has no CPU path (run on the B200 box)")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.pyoracle import Reference
    try:
        return Reference()
    except (FileNotFoundError, OSError) as e:  # GPU box without a prebuilt _ref
        pytest.skip(f"oracle/_ref not available: {e}")
