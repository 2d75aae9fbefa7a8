#!/usr/bin/env python3
"""Writes tests/golden/decode_bp.npz: known-answer vectors of oracle B (flooding sum-product, float64, parity-check
stop, whole H -- the restatement of the algorithm comm.LDPCDecoder runs at NRLDPCDecoder.m:120,265) on the LLRs already
committed in decode_nms.npz.  They pin oracle B against silent edits and give the CUDA sum-product kernel
(NRLDPC_ALG_BP) a committed target.  They are NOT outputs of the reference (no MATLAB here): decoder parity against
the toolbox itself stays unpinned.  Only decisions, iteration counts and parity flags are stored (the a-posteriori
values depend on the host libm's last bit)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

G = ROOT / "tests" / "golden"
src = np.load(G / "decode_nms.npz")
out = {}
for n in sorted({k.split("__")[0] for k in src.files}):
    bg, Z, iters, et, rows = src[n + "__cfg"].tolist()
    llr = src[n + "__llr"]
    for tag, early in (("stop", True), ("full", False)):
        r = O.decode_bp(bg, Z, llr, iters, early_term=early)
        out[f"{n}__{tag}__hard"] = np.packbits(r["hard"], axis=1)
        out[f"{n}__{tag}__iters"] = r["iters"]
        out[f"{n}__{tag}__ok"] = r["parity_ok"]
    out[n + "__cfg"] = np.array([bg, Z, iters], dtype=np.int32)
np.savez_compressed(G / "decode_bp.npz", **out)
print("written", G / "decode_bp.npz", {k: v.tolist() for k, v in out.items() if k.endswith("iters")})
