#!/usr/bin/env python
"""Where the sketch kernel's issue slots go: executed warp instructions and stall samples per code region, from the
SASS page of an `ncu --set full --import-source on` report.

    ncu -i prof.ncu-rep --page source --csv > source.csv
    python scripts/ncu_source_regions.py source.csv
Regions are found by landmarks in the SASS: the 16-byte global load of the fill phase, the two LDS.128 that start a
roll step, and the 16 `IMAD.WIDE ... 0x114253d5` (first Murmur multiply) of the unrolled hash block."""
import csv
import statistics
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    H, data = rows[h], rows[h + 1:]
    i_src, i_ex, i_smp = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
    ins = [(r[i_src].strip(), int(r[i_ex]), int(r[i_smp])) for r in data if len(r) > i_smp and r[i_ex].isdigit()]
    tot, tots = sum(x[1] for x in ins), sum(x[2] for x in ins)
    lds = [i for i, x in enumerate(ins) if x[0].startswith("LDS.128")]
    marks = [i for i, x in enumerate(ins) if "0x114253d5" in x[0] and "WIDE" in x[0]][:16]
    steps = statistics.median(x[1] for x in ins[marks[0]:marks[15]])
    print(f"{len(ins)} SASS instructions, {tot:.3e} warp instructions executed, {tots} samples")
    print(f"roll steps (512 positions per warp) executed: {steps:.0f}  ->  {tot / (steps * 16):.1f} instructions per l-mer")
    regions = [("read setup + fill (HPC keep mask, compaction into the ring)", 0, lds[0] - 5),
               ("roll step prologue (ring loads, validity)", lds[0] - 5, marks[0] - 30),
               ("unrolled hash block (16 l-mers per lane)", marks[0] - 30, marks[15] + 50),
               ("candidate list, flush (exact hash, blacklist, output), generic path", marks[15] + 50, len(ins))]
    for name, a, b in regions:
        e, s = sum(x[1] for x in ins[a:b]), sum(x[2] for x in ins[a:b])
        print(f"  {name:70s} {100 * e / tot:5.1f} % of executed   {100 * s / tots:5.1f} % of samples   "
              f"{e / (steps * 16):5.1f} instr per l-mer")


def main_packed():
    """sketch_packed_kernel: regions by execution-count plateaus (first kernel section of the CSV).  The unrolled hash block is
    the longest run of SASS instructions executed exactly once per roll step; what runs before it inside the per-read loop
    is stage + fill, what runs behind it is candidate handling."""
    rows = list(csv.reader(open(sys.argv[2])))
    hs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    H, data = rows[hs[0]], rows[hs[0] + 1:(hs[1] - 1 if len(hs) > 1 else len(rows))]
    i_src, i_ex, i_smp = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
    ins = [(r[i_src].strip(), int(r[i_ex]), int(r[i_smp])) for r in data if len(r) > i_smp and r[i_ex].isdigit()]
    tot, tots = sum(x[1] for x in ins), sum(x[2] for x in ins)
    best, i = (0, 0), 0
    while i < len(ins):
        j = i
        while j + 1 < len(ins) and ins[j + 1][1] == ins[i][1]:
            j += 1
        if ins[i][1] and j - i > best[1] - best[0]:
            best = (i, j + 1)
        i = j + 1
    a, b = best
    steps = ins[a][1]
    loop0 = min(i for i, x in enumerate(ins) if x[1] >= steps / 2)          # first instruction of the fill loop
    end = max(i for i, x in enumerate(ins) if x[1] >= steps / 2) + 1
    print(f"{len(ins)} SASS instructions, {tot:.3e} warp instructions executed, {tots} samples")
    print(f"roll steps (512 positions per warp) executed: {steps}  ->  {tot / (steps * 16):.1f} instructions per l-mer")
    for name, x, y in [("kernel / read setup (cursor, offsets, first bulk copies)", 0, loop0),
                       ("stage + fill (mbarrier waits, bulk-copy issue, table-driven HPC compaction, ring ORs)", loop0, a),
                       ("unrolled hash block (ring loads, 16 l-mers per lane, accept bits)", a, b),
                       ("candidate list, flush (exact hash, blacklist, output), loop control", b, end),
                       ("read epilogue (last flush, counts) and cold paths", end, len(ins))]:
        e, s = sum(v[1] for v in ins[x:y]), sum(v[2] for v in ins[x:y])
        print(f"  {name:90s} {100 * e / tot:5.1f} % of executed   {100 * s / tots:5.1f} % of samples   "
              f"{e / (steps * 16):5.1f} instr per l-mer")
    ops = {}
    for src, e, _ in ins[a:b]:
        op = src.split()[0] if not src.startswith("@") else src.split()[1]
        op = ".".join(op.split(".")[:2]) if op.startswith("IMAD") else op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    print("  hash block opcode mix (static = dynamic, straight-line): " +
          ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    if sys.argv[1] == "--packed":
        main_packed()
    else:
        main()
