#!/usr/bin/env python3
"""Digest of `ncu -i report.ncu-rep --page source --csv` (optionally gzipped): executed static footprint, share of
stall_no_inst samples, their position inside 128-byte instruction lines, and a per-1000-instruction histogram.  This is
how profiles/r01_v8_decode_f32_et_source_digest.txt was made (tools/gpu_prof_et.sh brings the CSV back)."""
import csv
import gzip
import io
import sys
from collections import Counter


def main(path):
    fh = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
    rows = list(csv.reader(fh))
    hdr = rows[1]
    ia, ins, ine, ino = (hdr.index(k) for k in ("Address", "# Samples", "Instructions Executed", "stall_no_inst"))
    data = [(int(r[ia], 16), int(r[ins] or 0), int(r[ine] or 0), int(r[ino] or 0)) for r in rows[2:] if len(r) >= len(hdr)]
    tot_s, tot_no, tot_e = (sum(d[k] for d in data) for k in (1, 3, 2))
    ex = sum(1 for d in data if d[2] > 0)
    print(f"static instructions {len(data)}, executed at least once {ex} ({ex * 16 / 1024:.1f} KB)")
    print(f"warp samples {tot_s}, of which stall_no_inst {tot_no} ({100 * tot_no / max(1, tot_s):.1f} %)")
    c = Counter()
    for a, s, e, no in data:
        c[(a % 128) // 16] += no
    print("stall_no_inst by position inside the 128-byte line:", "  ".join(f"{k}:{100 * c[k] / max(1, tot_no):.1f}%" for k in range(8)))
    print("buckets of 1000 static instructions: samples %, stall_no_inst %, executed %")
    for i in range(0, len(data), 1000):
        ch = data[i:i + 1000]
        print(f"  {i:6d}  {100 * sum(d[1] for d in ch) / max(1, tot_s):5.1f}  {100 * sum(d[3] for d in ch) / max(1, tot_no):5.1f}  "
              f"{100 * sum(d[2] for d in ch) / max(1, tot_e):5.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
