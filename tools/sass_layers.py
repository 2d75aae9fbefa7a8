#!/usr/bin/env python3
"""Static shape of the decode kernels in a `cuobjdump -sass` dump (dev tool, no GPU needed): total instructions, the span of
the unrolled layer loop (first to last layer barrier) in instructions and 128-byte lines, where the register spills sit
relative to that loop, and which barriers are reductions.  Used to judge a kernel change before spending GPU time:
    cuobjdump -sass ldpc_3gpp_matlab_b200/libnrldpc_b200.so > /tmp/k.sass && python tools/sass_layers.py /tmp/k.sass
"""
import re
import sys


def main(path, pat="decode_nms"):
    txt = open(path).read()
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n")[0]
        if pat not in name:
            continue
        lines = [l for l in f.split("\n") if re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", l)]
        bars = [(i, "RED" if "BAR.RED" in l else "SYNC") for i, l in enumerate(lines) if "BAR.SYNC" in l or "BAR.RED" in l]
        spills = [i for i, l in enumerate(lines) if re.search(r"\b(STL|LDL)\b", l)]
        rets = [i for i, l in enumerate(lines) if "RET." in l or re.search(r"\bEXIT\b", l)]
        print(f"{name}: {len(lines)} instructions ({len(lines) * 16 / 1024:.1f} KB), {len(bars)} barriers, returns at {rets}")
        if len(bars) > 34:
            lo, hi = bars[2][0], bars[34][0]      # after the two prologue barriers: 32 layer barriers of a full base graph 1
            in_loop = [i for i in spills if lo <= i <= hi]
            print(f"  layer loop (barriers 2..34): instructions {lo}..{hi} = {hi - lo} ({(hi - lo) * 16 / 128:.0f} lines of 128 B); "
                  f"spill instructions inside: {len(in_loop)} of {len(spills)}")
        print("  reducing barriers at", [i for i, k in bars if k == "RED"])


if __name__ == "__main__":
    main(*sys.argv[1:3])
