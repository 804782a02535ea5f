"""The CUDA kernels' SOURCE, executed on the CPU.

tests/cpp/warp_emu.hpp is a small SIMT emulator: one thread block = n fibers in one OS thread, warp collectives
(`__shfl*_sync`, `__ballot_sync`, `__any/__all_sync`, `__syncwarp`) complete when every lane of the warp has arrived
at the same call, `__syncthreads` when the whole block has; divergent, abandoned or mismatched collectives abort.
tests/_emu.py hands the product's .cu files to g++ with only host-side constructs rewritten
(`kernel<<<g, b, 0, s>>>(args)` -> `emu::launch(g, b, [&]{ kernel(args); })`, and kminmer.cu's two inline-PTX slot
primitives -> plain loads / compare-and-swap, exact in a single-threaded emulator).  Kernel bodies, device helpers
and the launch_* grid logic are compiled unchanged, driven the way api.cu drives them, and every result is compared
with the oracle.

This is a logic check of the GPU code that needs no GPU (and the way kernel changes are screened before GPU time
is spent on them); the `-m gpu` parity tests remain the authority for the compiled kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _emu  # noqa: E402


def test_sketch_scan_compact_kernels_in_emulator(tmp_path):
    """sketch_kernel<15>/<0> through launch_sketch (2 CTAs x 8 warps per 'SM'), the 3-kernel scan and compact_kernel:
    ragged / dirty / homopolymer / tandem / 40 kbp reads, HPC on and off, densities 0, 0.005, 0.05, 0.6 (slot
    overflow + exact re-run) and 1.0 (exact-hash fallback), generic l, 2-bit packed input, blacklist."""
    out = _emu.build_and_run(tmp_path, "sketch_emu_test.cpp", {"SKETCH_SOURCE": "sketch.cu"})
    assert "slot overflows exercised" in out


def test_sketch_kernel_does_not_depend_on_stale_shared_memory(tmp_path):
    """Same suite with -DMDBG_POISON_SMEM: every warp's ring is filled with non-code bytes when the kernel starts
    (a real SM hands a CTA whatever the previous CTA left in shared memory) and one launch is made per read, so a
    result that depends on ring bytes past `avail` differs from the oracle.  (Found on a B200: variant 1 packed the
    lane's 32 ring bytes with a multiply that let a garbage byte 31 carry into the codes of positions 14 / 15.)"""
    out = _emu.build_and_run(tmp_path, "sketch_emu_test.cpp", {"SKETCH_SOURCE": "sketch.cu"},
                             extra_flags=("-DMDBG_POISON_SMEM",))
    assert "slot overflows exercised" in out


def test_read_aux_kernel_in_emulator(tmp_path):
    """read_aux_kernel: exact error sums -> mean read quality (bit-exact float), DUST-like complexity (bit-exact
    double) + the low-complexity filter, per-minimizer minimum quality through the HPC -> raw coordinate re-scan."""
    out = _emu.build_and_run(tmp_path, "aux_emu_test.cpp", {"AUX_SOURCE": "aux.cu"})
    assert "qualities compared" in out


def test_minimizer_space_kernels_in_emulator(tmp_path):
    """purge.cu + kminmer.cu: purgePalindrome (flag / exact / compact), density re-threshold, insert (k = 4 and
    generic), stats, block-aggregated emit, rescue, previous-k table (from the device table and loaded) + next-k,
    and the owner merge (pack by owner, insert-add of foreign vectors) with 3 ranks in one process."""
    out = _emu.build_and_run(tmp_path, "table_emu_test.cpp", {"PURGE_SOURCE": "purge.cu", "KMINMER_SOURCE": "kminmer.cu"})
    assert "table entries" in out
