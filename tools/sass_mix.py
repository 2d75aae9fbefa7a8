#!/usr/bin/env python3
"""Opcode mix of one kernel in a cuobjdump -sass dump (dev tool: ALU-pipe vs FMA-pipe instruction budget)."""
import re
import sys
from collections import Counter

ALU = {"LOP3", "SEL", "FSEL", "ISETP", "FSETP", "FMNMX", "FMNMX3", "IADD3", "LEA", "SHF", "VIMNMX", "VIMNMX3", "VIADD", "PRMT",
       "IABS", "PLOP3", "POPC", "FLO", "BREV", "SGXT", "BMSK", "HMNMX2", "HSETP2", "HSET2", "VIADDMNMX", "IADD", "MOV", "I2FP", "F2FP"}
FMA = {"IMAD", "FADD", "FMUL", "FFMA", "HADD2", "HMUL2", "HFMA2"}


def main(path, pat, lo=None, hi=None):
    txt = open(path).read()
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    for f in funcs:
        name = f.split("\n")[0]
        if pat not in name:
            continue
        ins = re.findall(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_]+)", f, re.M)
        if lo is not None:
            ins = [(a, o) for a, o in ins if int(lo, 16) <= int(a, 16) < int(hi, 16)]
        c = Counter(o for _, o in ins)
        n = len(ins)
        alu = sum(v for k, v in c.items() if k in ALU)
        fma = sum(v for k, v in c.items() if k in FMA)
        print(f"{name}: {n} instructions, ALU-pipe {alu}, FMA-pipe {fma}, other {n - alu - fma}")
        print("  " + "  ".join(f"{k}:{v}" for k, v in c.most_common(30)))


if __name__ == "__main__":
    main(*sys.argv[1:])
