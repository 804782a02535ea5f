"""Build helper for the warp-emulator tests: turns a kernel source file into something g++ can compile against
tests/cpp/warp_emu.hpp WITHOUT editing the product source.  Only host-side constructs are rewritten:

  kernel<<<grid, block, 0, s>>>(args);   ->  emu::launch(grid, block, s, [=] { kernel(args); });
  the two inline-PTX slot primitives of kminmer.cu (ld.relaxed.v2.u64, atom.cas.b128) -> their plain C meaning
  (the emulator is single threaded, so a plain load and a plain compare-and-swap are exact).

Kernel bodies, device helpers and the grid-size logic of the launch_* functions are compiled as they are."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "metamdbg_b200", "csrc")

_C_SLOT_PRIMITIVES = '''
__device__ __forceinline__ void load_key(const Slot* s, uint64_t& lo, uint64_t& hi) { lo = s->lo; hi = s->hi; }
__device__ __forceinline__ void cas_key(Slot* s, uint64_t new_lo, uint64_t new_hi, uint64_t& old_lo, uint64_t& old_hi) {
    old_lo = s->lo; old_hi = s->hi;
    if (old_lo == 0 && old_hi == 0) { s->lo = new_lo; s->hi = new_hi; }
}
'''


def emulated_source(name: str) -> str:
    src = open(os.path.join(CSRC, name)).read()
    src, n = re.subn(r"(\w+(?:<[^<>;]*>)?)<<<([^;]*?),\s*([^;,]*?),\s*0,\s*s>>>\(([^;]*?)\);",
                     r"emu::launch(\2, \3, s, [=] { \1(\4); });", src, flags=re.S)
    assert n > 0 and "<<<" not in src, f"{name}: kernel launches not rewritten"
    if name == "kminmer.cu" and "asm volatile" in src:     # (sketch.cu keeps its PTX behind #ifdef __CUDACC__)
        src, n1 = re.subn(r"__device__ __forceinline__ void load_key\(.*?\n}\n", "", src, count=1, flags=re.S)
        src, n2 = re.subn(r"(// atom\.cas\.b128.*?\n)?__device__ __forceinline__ void cas_key\(.*?\n}\n",
                          _C_SLOT_PRIMITIVES, src, count=1, flags=re.S)
        assert n1 == 1 and n2 == 1 and "asm" not in src, f"{name}: inline PTX not replaced"
    return src


def build_and_run(tmp_path, test_cpp: str, sources: dict, extra_flags=(), timeout=900):
    """sources: {MACRO_NAME: 'file.cu'}; the test includes them through -DMACRO_NAME="path"."""
    from oracle import pyoracle
    pyoracle.build()
    defs = []
    for macro, name in sources.items():
        inc = tmp_path / (name.replace(".", "_") + ".inc")
        inc.write_text(emulated_source(name))
        defs.append(f'-D{macro}="{inc}"')
    exe = tmp_path / os.path.splitext(test_cpp)[0]
    odir = os.path.join(ROOT, "oracle")
    cmd = ["/usr/bin/g++", "-O1", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas",
           "-I" + os.path.join(ROOT, "tests", "cpp", "emu_stub"), "-I" + os.path.join(ROOT, "tests", "cpp"), "-I" + CSRC,
           *defs, *extra_flags, "-o", str(exe), os.path.join(ROOT, "tests", "cpp", test_cpp), "-L" + odir, "-lmdbg_oracle",
           "-lm", "-Wl,-rpath," + odir]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-4000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=timeout)
    assert run.returncode == 0 and run.stdout.strip().endswith("OK"), run.stdout[-3000:] + run.stderr[-2000:]
    return run.stdout


def build_emulated_library(out_dir) -> str:
    """The whole C-ABI library for the CPU emulator (TEST ONLY, never shipped or loaded by the product): api.cu is
    compiled unchanged against the synchronous runtime stub, the four kernel files with their launches rewritten,
    pack_host.cpp as it is.  Returns the path of libmdbg_b200_emu.so inside out_dir."""
    out_dir = str(out_dir)
    objs = []
    stub = os.path.join(ROOT, "tests", "cpp", "emu_stub")
    # -DMDBG_POISON_SMEM: every sketch launch starts from a ring full of non-code bytes, as on a real SM
    common = ["/usr/bin/g++", "-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas", "-w",
              "-DMDBG_POISON_SMEM", *os.environ.get("MDBG_EMU_EXTRA_FLAGS", "").split(),   # e.g. -fsanitize=address -g
              "-I" + stub, "-I" + os.path.join(ROOT, "tests", "cpp"), "-I" + CSRC, "-I" + os.path.join(ROOT, "include")]
    for name in ("sketch.cu", "kminmer.cu", "purge.cu", "aux.cu", "ingest.cu", "repeats.cu", "unitig.cu"):
        inc = os.path.join(out_dir, name.replace(".cu", "_emu.cpp"))
        with open(inc, "w") as f:
            f.write(emulated_source(name))
        objs.append(inc)
    api = os.path.join(out_dir, "api_emu.cpp")
    src = open(os.path.join(CSRC, "api.cu")).read()
    src = src.replace('#include "../../include/mdbg_b200.h"', '#include "mdbg_b200.h"')     # same header, found through -I
    with open(api, "w") as f:
        f.write(src)
    lib = os.path.join(out_dir, "libmdbg_b200_emu.so")
    cmd = common + ["-shared", "-o", lib, api, *objs, os.path.join(CSRC, "pack_host.cpp"), "-ldl", "-lpthread"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-4000:]
    return lib
