#!/usr/bin/env python3
"""Split one kernel's SASS at BAR.SYNC and print per-segment opcode counts (dev tool)."""
import re
import sys
from collections import Counter

ALU = {"LOP3", "SEL", "FSEL", "ISETP", "FSETP", "FMNMX", "FMNMX3", "IADD3", "LEA", "SHF", "VIMNMX", "VIMNMX3", "VIADD", "PRMT",
       "IABS", "PLOP3", "POPC", "HMNMX2", "HSETP2", "HSET2", "MOV"}
FMA = {"IMAD", "FADD", "FMUL", "FFMA", "HADD2", "HMUL2", "HFMA2"}


def segments(path, pat):
    txt = open(path).read()
    f = [x for x in re.split(r"\n\s*Function : ", txt)[1:] if pat in x.split("\n")[0]][0]
    lines = [re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).strip() for l in f.split("\n") if re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", l)]
    segs, cur = [], []
    for l in lines:
        cur.append(l)
        if "BAR.SYNC" in l:
            segs.append(cur)
            cur = []
    segs.append(cur)
    return segs


def op(l):
    m = re.match(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z0-9_]+)", l)
    return m.group(1) if m else "?"


if __name__ == "__main__":
    segs = segments(sys.argv[1], sys.argv[2])
    show = int(sys.argv[3]) if len(sys.argv) > 3 else None
    for i, s in enumerate(segs):
        c = Counter(op(l) for l in s)
        alu = sum(v for k, v in c.items() if k in ALU)
        fma = sum(v for k, v in c.items() if k in FMA)
        print(f"seg {i:3d}: n={len(s):4d} alu={alu:4d} fma={fma:4d} lds={c['LDS']:3d} sts={c['STS']:3d} ldg={c['LDG']} stg={c['STG']}")
    if show is not None:
        print("\n".join(segs[show]))
