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


if __name__ == "__main__":
    main()
