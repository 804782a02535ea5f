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
rescue, next-k, finalize, error paths, and the multi-rank owner merge over an in-process fake NCCL) is checked here
without a GPU.  The emulated library is test infrastructure:
it is built into a temporary directory, nothing in metamdbg_b200/ can load it, and the product still refuses to
run without a CUDA device (tests/test_capi_cpu.py)."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _emu  # noqa: E402


import pytest  # noqa: E402


@pytest.fixture(scope="module")
def emulated(tmp_path_factory):
    """(library path, environment for child processes): the emulated library under its own name and under the
    soname the C++ binaries ask for, plus the in-process fake libnccl.so.2, all in one temporary directory."""
    import __graft_entry__ as g
    g.build()                                    # the host driver binaries (they are re-pointed at run time)
    d = str(tmp_path_factory.mktemp("emu"))
    lib = _emu.build_emulated_library(d)
    shutil.copy(lib, os.path.join(d, "libmdbg_b200.so"))
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", os.path.join(d, "libnccl.so.2"),
                    os.path.join(ROOT, "tests", "cpp", "fake_nccl.cpp"), "-lpthread", "-ldl"], check=True)
    env = dict(os.environ, MDBG_EMU_LIB=lib, LD_LIBRARY_PATH=d + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    return lib, env


@pytest.mark.parametrize("n_ranks,k,n_reads", [(2, 4, 260), (3, 5, 260), (8, 4, 260), (8, 4, 5)])
def test_owner_merge_on_the_emulator(emulated, n_ranks, k, n_reads):
    """mdbg_comm_init + mdbg_count_merge with N ranks = N threads of one process, each with its own emulated
    context, exchanging through tests/cpp/fake_nccl.cpp (an in-process libnccl.so.2: barriers + memcpy).  The union
    of the ranks' tables is the oracle's table of the whole read set, every key sits on its owner, occurrences are
    conserved -- the N = 8 sequencing of the merge is exercised here although the round's GPU runs stopped at 4."""
    _, env = emulated
    # (8 ranks, 5 reads): three ranks hold no read at all but still take part in every collective and own keys
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu_multirank_child.py"), str(n_ranks), str(k),
                          str(n_reads)],
                         env=env, capture_output=True, text=True, timeout=900)
    assert run.returncode == 0 and run.stdout.strip().endswith("OK"), run.stdout[-2000:] + run.stderr[-3000:]


@pytest.mark.parametrize("n_ranks,last_k", [(3, 12), (8, 9)])
def test_multi_k_sweeps_over_several_ranks_on_the_emulator(emulated, n_ranks, last_k):
    """Four multi-k sweeps in a row on N ranks with the keys-only owner merge after every k (tests/emu_multirank_sweeps_child.py):
    the rotation of the three table buffers (trades, common size) leaves every table intact -- the owner-partitioned
    tables add up to the single-context ones in every sweep -- and stops allocating within three sweeps."""
    _, env = emulated
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu_multirank_sweeps_child.py"), str(n_ranks), str(last_k)],
                         env=env, capture_output=True, text=True, timeout=900)
    assert run.returncode == 0 and run.stdout.strip().endswith("OK"), run.stdout[-2000:] + run.stderr[-3000:]


def test_gpu_parity_suites_against_the_emulated_library(emulated):
    lib, env = emulated
    run = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_gpu_host_cpp.py", "tests/test_gpu_z_new_paths.py", "tests/test_gpu_zz_round2.py", "-m", "gpu",
                          "-q", "-x", "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True,
                         timeout=3000)
    tail = run.stdout[-2500:] + run.stderr[-1500:]
    assert run.returncode == 0, tail
    last = run.stdout.strip().splitlines()[-1]
    assert "passed" in last and "failed" not in last, tail
    n_passed = int(last.split(" passed")[0].split()[-1])
    assert n_passed >= 39, tail                  # parity + host-driver + the paths added after the last GPU run


def test_sketch_variant_1_through_the_emulated_library(emulated):
    """The sketch-heavy GPU parity tests once more with MDBG_SKETCH_VARIANT=1: variant 1 of the unrolled register
    block (k1v1:: in common.cuh / bitmath.cuh) behind the whole host-side sequencing."""
    lib, env = emulated
    env = dict(env, MDBG_SKETCH_VARIANT="1")
    run = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_gpu_z_new_paths.py", "-m", "gpu", "-q", "-x",
                          "-p", "no:cacheprovider", "-k", "sketch_hifi or sketch_ont or sketch_golden or sketch_edge or sketch_long or packed_host or full_path or piece_pipeline or sentinel"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=3000)
    tail = run.stdout[-2500:] + run.stderr[-1500:]
    assert run.returncode == 0, tail
    last = run.stdout.strip().splitlines()[-1]
    assert "passed" in last and "failed" not in last, tail
    assert int(last.split(" passed")[0].split()[-1]) >= 9, tail


def test_gpu_parity_suite_with_deferred_stream_execution(tmp_path):
    """The same library on the ADVERSARIAL serialisation of stream semantics (-DEMU_DEFERRED, see
    tests/cpp/emu_stub/cuda_runtime.h): asynchronous copies and kernels stay queued until the host, or another stream
    through an event, really waits for them; pinned memory is read / written only when the copy executes.  A host
    read of a result, or a reuse of a staging buffer, that is not ordered by a real synchronisation then sees
    poisoned or stale bytes and the parity tests fail.  (Removing the d2h-stream synchronisation of the piece
    pipeline makes this test fail while the eager emulator still passes.)"""
    os.environ["MDBG_EMU_EXTRA_FLAGS"] = "-DEMU_DEFERRED"
    try:
        lib = _emu.build_emulated_library(str(tmp_path))
    finally:
        del os.environ["MDBG_EMU_EXTRA_FLAGS"]
    env = dict(os.environ, MDBG_EMU_LIB=lib)
    # three ranks over the fake NCCL (it flushes the rank's stream through the library's emu_stream_synchronize hook)
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", str(tmp_path / "libnccl.so.2"),
                    os.path.join(ROOT, "tests", "cpp", "fake_nccl.cpp"), "-lpthread", "-ldl"], check=True)
    mr = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu_multirank_child.py"), "3", "4", "120"],
                        env=dict(env, LD_LIBRARY_PATH=str(tmp_path) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", "")),
                        capture_output=True, text=True, timeout=900)
    assert mr.returncode == 0 and mr.stdout.strip().endswith("OK"), mr.stdout[-2000:] + mr.stderr[-3000:]
    # the tests that exercise host-side sequencing (the kernels themselves are covered by the eager run above)
    pick = ("(piece or packed or count or rescue or next_k or multi_k or edge or full_path or side_outputs_vs or sentinel "
            "or variants or table_full or bad_host or empty or smoke or density) and not cpp_driver")
    run = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_gpu_z_new_paths.py", "tests/test_gpu_zz_round2.py", "-m", "gpu",
                          "-q", "-x", "-p", "no:cacheprovider", "-k", pick], cwd=ROOT, env=env, capture_output=True, text=True,
                         timeout=3000)
    tail = run.stdout[-2500:] + run.stderr[-1500:]
    assert run.returncode == 0, tail
    last = run.stdout.strip().splitlines()[-1]
    assert "passed" in last and "failed" not in last, tail
    assert int(last.split(" passed")[0].split()[-1]) >= 20, tail
