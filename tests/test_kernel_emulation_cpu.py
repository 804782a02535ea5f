"""The sketch kernel's SOURCE, executed on the CPU.

tests/cpp/warp_emu.hpp is a small SIMT emulator (one warp = 32 fibers, warp collectives exchange values when every
lane has arrived, divergent or abandoned collectives abort).  tests/cpp/sketch_emu_test.cpp includes
metamdbg_b200/csrc/sketch.cu -- unchanged except that the host-side `launch_*` functions (the only place with
`<<<...>>>` and CUDA runtime calls) are cut out by this test before compiling -- and runs `sketch_kernel<15>` /
`sketch_kernel<0>` over ragged, dirty, homopolymer, tandem-repeat and long reads: HPC on/off, densities 0 / 0.005 /
0.05 / 0.6 (slot overflow + exact re-run) / 1.0 (exact-hash fallback), generic l, 2-bit packed input, blacklist.
Every minimizer, position and strand is compared with the oracle.  This is a logic check of the GPU code that
needs no GPU; the `-m gpu` parity tests remain the authority for the compiled kernel."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def strip_launchers(src: str) -> str:
    out = re.sub(r"\n(?:static )?[A-Za-z_0-9]+ launch_[a-z_0-9]+\([^)]*\) \{\n.*?\n\}\n", "\n", src, flags=re.S)
    assert "<<<" not in out, "a kernel launch survived the stripping"
    return out


def test_sketch_kernel_source_in_warp_emulator(tmp_path):
    from oracle import pyoracle
    pyoracle.build()
    csrc = os.path.join(ROOT, "metamdbg_b200", "csrc")
    inc = tmp_path / "sketch_kernel_src.inc"
    inc.write_text(strip_launchers(open(os.path.join(csrc, "sketch.cu")).read()))
    exe = tmp_path / "sketch_emu_test"
    odir = os.path.join(ROOT, "oracle")
    cmd = ["/usr/bin/g++", "-O1", "-std=c++17", "-Wno-unknown-pragmas", "-I" + os.path.join(ROOT, "tests", "cpp", "emu_stub"),
           "-I" + os.path.join(ROOT, "tests", "cpp"), "-I" + csrc, f'-DSKETCH_SOURCE="{inc}"', "-o", str(exe),
           os.path.join(ROOT, "tests", "cpp", "sketch_emu_test.cpp"), "-L" + odir, "-lmdbg_oracle", "-Wl,-rpath," + odir]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0 and run.stdout.strip().endswith("OK"), run.stdout[-3000:] + run.stderr[-2000:]
