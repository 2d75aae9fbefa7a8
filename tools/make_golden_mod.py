#!/usr/bin/env python3
"""Dev-time generator of tests/golden/constellations.json (needs the read-only reference checkout; never run on
the GPU box).

Parses the toolbox CustomSymbolMapping vectors out of /root/reference/NRModulator.m:73-81 and turns them into
constellation points with MathWorks' documented conventions:
  * comm.RectangularQAMModulator, custom mapping: element i of the vector is the symbol integer sent to point i,
    points counted from the top-left corner down each column, columns left to right; 'Average power'
    normalisation to unit power; BitInput: first bit of a group is the MSB of the integer;
  * comm.PSKModulator(M, PhaseOffset): point i = exp(j*(PhaseOffset + 2*pi*i/M)); no custom mapping = binary order.
The result (symbol integer -> point, per modulation) is the reference's own definition of the maps; the script
asserts that it coincides with the TS 38.211 section 5.1 formulas the oracle and the kernels implement.
"""
import json
import math
import re
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/NRModulator.m")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "constellations.json"


def ts38211(Qm, s):
    b = [(s >> (Qm - 1 - k)) & 1 for k in range(Qm)]
    if Qm == 1:
        return ((1 - 2 * b[0]), (1 - 2 * b[0])), math.sqrt(2)
    if Qm == 2:
        return ((1 - 2 * b[0]), (1 - 2 * b[1])), math.sqrt(2)
    if Qm == 4:
        return ((1 - 2 * b[0]) * (2 - (1 - 2 * b[2])), (1 - 2 * b[1]) * (2 - (1 - 2 * b[3]))), math.sqrt(10)
    if Qm == 6:
        return ((1 - 2 * b[0]) * (4 - (1 - 2 * b[2]) * (2 - (1 - 2 * b[4]))),
                (1 - 2 * b[1]) * (4 - (1 - 2 * b[3]) * (2 - (1 - 2 * b[5])))), math.sqrt(42)
    return ((1 - 2 * b[0]) * (8 - (1 - 2 * b[2]) * (4 - (1 - 2 * b[4]) * (2 - (1 - 2 * b[6])))),
            (1 - 2 * b[1]) * (8 - (1 - 2 * b[3]) * (4 - (1 - 2 * b[5]) * (2 - (1 - 2 * b[7]))))), math.sqrt(170)


def main():
    text = REF.read_text()
    out = {}
    for name, Qm in (("BPSK", 1), ("QPSK", 2), ("16QAM", 4), ("64QAM", 6), ("256QAM", 8)):
        m = re.search(r"strcmp\(Modulation_, '%s'\)\s*\n\s*obj\.hMod = (comm\.\w+)\((.*?)\);" % name, text, re.S)
        assert m, name
        ctor, argstr = m.group(1), m.group(2)
        M = int(re.search(r"'ModulationOrder',(\d+)", argstr).group(1))
        assert M == 1 << Qm
        mv = re.search(r"'CustomSymbolMapping',\[([\d,\s]+)\]", argstr)
        mapping = [int(x) for x in re.split(r"[,\s]+", mv.group(1).strip())] if mv else list(range(M))
        assert sorted(mapping) == list(range(M))
        pts = {}
        if ctor == "comm.PSKModulator":
            assert "'PhaseOffset',pi/4" in argstr
            for i, sym in enumerate(mapping):
                ph = math.pi / 4 + 2 * math.pi * i / M
                pts[sym] = (math.cos(ph), math.sin(ph))
        else:
            assert "'NormalizationMethod','Average power'" in argstr
            side = int(round(math.sqrt(M)))
            norm = math.sqrt(2 * (M - 1) / 3)
            for i, sym in enumerate(mapping):
                col, row = divmod(i, side)
                pts[sym] = ((-(side - 1) + 2 * col) / norm, ((side - 1) - 2 * row) / norm)
        for s in range(M):
            (x, y), n = ts38211(Qm, s)
            assert abs(pts[s][0] - x / n) < 1e-12 and abs(pts[s][1] - y / n) < 1e-12, (name, s, pts[s], x / n, y / n)
        out[name] = {"Q_m": Qm, "points": [[pts[s][0], pts[s][1]] for s in range(M)]}
    OUT.write_text(json.dumps(out))
    print("wrote", OUT, {k: len(v["points"]) for k, v in out.items()}, "- reference vectors == TS 38.211 formulas")


if __name__ == "__main__":
    main()
