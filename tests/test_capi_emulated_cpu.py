"""The whole C-ABI library, on the CPU emulator.

tests/_emu.py compiles the PRODUCT sources with g++ -- api.cu unchanged against a synchronous stand-in for the CUDA
runtime (tests/cpp/emu_stub/cuda_runtime.h: device memory = host memory, async work completes in enqueue order),
the four kernel files against tests/cpp/warp_emu.hpp, pack_host.cpp as it is -- into a throw-away
libmdbg_b200_emu.so, and a child pytest re-runs the GPU parity suites against it:

  * tests/test_gpu_parity.py   every test that talks to the library through host arrays (27 of 30; the three that
                               hand torch CUDA tensors to the device-pointer entry points need a real GPU)
  * tests/test_gpu_host_cpp.py the C++ host driver and metaMDBG's own readSelection stage with the GPU functor
                               plugged in (the binaries pick the emulated library up through LD_LIBRARY_PATH)

So the host-side sequencing of api.cu (pieces, 2-bit packed transfer, slot overflow re-run, store, purge, count,
rescue, next-k, finalize, error paths) is checked here without a GPU.  The emulated library is test infrastructure:
it is built into a temporary directory, nothing in metamdbg_b200/ can load it, and the product still refuses to
run without a CUDA device (tests/test_capi_cpu.py)."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _emu  # noqa: E402


def test_gpu_parity_suites_against_the_emulated_library(tmp_path):
    import __graft_entry__ as g
    g.build()                                    # the host driver binaries (they are re-pointed at run time)
    lib = _emu.build_emulated_library(tmp_path)
    shutil.copy(lib, os.path.join(str(tmp_path), "libmdbg_b200.so"))     # the soname the C++ binaries ask for
    env = dict(os.environ, MDBG_EMU_LIB=lib,
               LD_LIBRARY_PATH=str(tmp_path) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    run = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_gpu_host_cpp.py", "-m", "gpu",
                          "-q", "-x", "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True,
                         timeout=3000)
    tail = run.stdout[-2500:] + run.stderr[-1500:]
    assert run.returncode == 0, tail
    last = run.stdout.strip().splitlines()[-1]
    assert "passed" in last and "failed" not in last, tail
    n_passed = int(last.split(" passed")[0].split()[-1])
    assert n_passed >= 31, tail                  # 27 parity + 4 host-driver tests
