#!/usr/bin/env python3
"""profiles/traffic.json from the ncu metric pass of the CURRENT binary (tools/gpu_r02_final.sh): DRAM bytes per decode
launch of the headline batch, ALU-pipe / issue utilisation of the same kernels from the --set full summaries, and the
sha256 of the kernel sources the capture belongs to (bench.py compares it with the tree it runs from)."""
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r02_v5"
out = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum of one decode kernel launch (ncu metric pass, %s kernels, headline batch "
                "4096 x BG1 Z=384); algorithmic bytes per launch = 4096 x 112896 = 462,422,016.  pipes: sm__pipe_alu_cycles_active / "
                "smsp__issue_active of the same kernels (ncu --set full, profiles/%s_decode_*_ncu_full.txt)" % (tag, tag),
       "_capture": tag, "_csrc_sha16": bench.csrc_digest(), "pipes": {}}
for dt, key in (("f32", "bg1_z384_r13_it8_b4096"), ("f16x2", "bg1_z384_r13_it8_b4096|f16x2")):
    rd = wr = None
    for line in open(ROOT / "profiles" / f"{tag}_dram_{dt}.csv"):
        m = re.search(r'"(dram__bytes_(read|write)\.sum)","[^"]*","([0-9.,]+)"', line)
        if m:
            v = float(m.group(3).replace(",", ""))
            unit = re.search(r'"dram__bytes_(?:read|write)\.sum","([^"]*)"', line).group(1)
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            if m.group(2) == "read":
                rd = v
            else:
                wr = v
    out[key] = int(rd + wr)
    out[key + "|read"] = int(rd)
    out[key + "|write"] = int(wr)
    pipes = {}
    for line in open(ROOT / "profiles" / f"{tag}_decode_{dt}_ncu_full.txt"):
        if line.startswith("sm__pipe_alu_cycles_active"):
            pipes["alu_pipe_pct"] = round(float(line.split()[1]), 2)
        if line.startswith("smsp__issue_active"):
            pipes["issue_active_pct"] = round(float(line.split()[1]), 2)
    out["pipes"][key] = pipes
(ROOT / "profiles" / "traffic.json").write_text(json.dumps(out, indent=2) + "\n")
print(json.dumps(out, indent=1))
