#!/usr/bin/env python
"""Static instruction mix of one kernel in an object file (cuobjdump -sass), grouped by the pipe the opcode issues
to on sm_100.  Used to compare formulations of the sketch kernel's hash arithmetic before spending GPU time.

    python scripts/sass_count.py metamdbg_b200/csrc/sketch.o sketch_kernelILi15
"""
import re
import subprocess
import sys
from collections import Counter

FMA = ("IMAD", "FFMA", "FMUL", "FADD", "IDP", "HFMA2")             # FMA pipes (IMAD* incl. IMAD.MOV/SHL/IADD)
ALU = ("LOP3", "SHF", "IADD3", "VIADD", "SEL", "ISETP", "VIMNMX", "VIADDMNMX", "LEA", "PLOP3", "POPC", "BREV", "FLO",
       "PRMT", "MOV", "CS2R", "SGXT", "BMSK", "IABS", "ICMP", "FSETP", "I2F", "F2I", "P2R", "R2P")


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    cur, counts = None, {}
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur and pat in cur:
            counts.setdefault(cur, Counter())[m.group(1) + m.group(2)] += 1
    for fn, c in counts.items():
        tot = sum(c.values())
        fma = sum(v for k, v in c.items() if k.split(".")[0] in FMA)
        alu = sum(v for k, v in c.items() if k.split(".")[0] in ALU)
        true_mul = sum(v for k, v in c.items() if k.split(".")[0] == "IMAD" and not any(
            t in k for t in (".MOV", ".SHL", ".IADD")))
        print(f"{fn}: {tot} instructions, FMA-pipe {fma} (true multiplies {true_mul}), ALU-pipe {alu}, other {tot - fma - alu}")
        for k, v in c.most_common(14):
            print(f"    {v:5d} {k}")


if __name__ == "__main__":
    main()
