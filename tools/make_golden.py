#!/usr/bin/env python3
"""Writes tests/golden/*.  Two kinds of fixture:

  params.json / tables.json -- values DERIVED from the reference's formulas and tables
      (NRLDPC.m:297-543, get_3gpp_base_graph.m:13-328,333-529).  They were
      computed once from the reference text and are committed so the tests never read /root/reference.
  decode_nms.npz            -- small decoder known-answer vectors produced by the CPU oracle
      (oracle A).  They pin the oracle against silent edits and give the GPU tests a committed
      target.  They are NOT reference outputs: decoder parity is unpinned (no MATLAB here, the
      reference ships no decoder vectors) -- see oracle/nrldpc_oracle.c.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

G = ROOT / "tests" / "golden"
G.mkdir(parents=True, exist_ok=True)

# Appendix A.3 worked parameter sets (A, BG, R) -> derived values
params = [
    dict(A=20, BG=2, R="1/5", tb_L=16, B=36, C=1, K_prime=36, K_b=6, Z_c=6, i_LS=1, K=60, N=300, filler=24, G=100, E_r=[100]),
    dict(A=400, BG=2, R="1/5", tb_L=16, B=416, C=1, K_prime=416, K_b=8, Z_c=52, i_LS=6, K=520, N=2600, filler=104, G=2000, E_r=[2000]),
    dict(A=1000, BG=1, R="1/3", tb_L=16, B=1016, C=1, K_prime=1016, K_b=22, Z_c=48, i_LS=1, K=1056, N=3168, filler=40, G=3000, E_r=[3000]),
    dict(A=3842, BG=2, R="1/3", tb_L=24, B=3866, C=2, K_prime=1957, K_b=10, Z_c=208, i_LS=6, K=2080, N=10400, filler=123, G=11526, E_r=[5762, 5764]),
    dict(A=8000, BG=1, R="1/3", tb_L=24, B=8024, C=1, K_prime=8024, K_b=22, Z_c=384, i_LS=1, K=8448, N=25344, filler=424, G=24000, E_r=[24000]),
    dict(A=8424, BG=1, R="1/3", tb_L=24, B=8448, C=1, K_prime=8448, K_b=22, Z_c=384, i_LS=1, K=8448, N=25344, filler=0, G=25272, E_r=[25272]),
    dict(A=8424, BG=1, R="8/9", tb_L=24, B=8448, C=1, K_prime=8448, K_b=22, Z_c=384, i_LS=1, K=8448, N=25344, filler=0, G=9478, E_r=[9478]),
]
(G / "params.json").write_text(json.dumps(params, indent=1))

tables = {
    "lifting_sizes": [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 44, 48,
                      52, 56, 60, 64, 72, 80, 88, 96, 104, 112, 120, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288,
                      320, 352, 384],
    "bg1": dict(edges=316, rows=46, cols=68, sum_row=5472, sum_col=4866,
                sum_V=[34730, 49099, 42436, 29665, 31271, 47538, 20577, 34191],
                max_V=[255, 383, 319, 223, 283, 351, 207, 237],
                first=[0, 0, 250, 307, 73, 223, 211, 294, 0, 135], last=[45, 67, 0, 0, 0, 0, 0, 0, 0, 0],
                sha256="4f7508858e04dc7bbed58f8b13c39b4bdb67835790d7e3b7f19152f2e59578d2",
                row_deg=[19, 19, 19, 19, 3, 8, 9, 7, 10, 9, 7, 8, 7, 6, 7, 7, 6, 6, 6, 6, 6, 6, 5, 5, 6, 5, 5, 4, 5, 5,
                         5, 5, 5, 5, 5, 5, 5, 4, 5, 5, 4, 5, 4, 5, 5, 4]),
    "bg2": dict(edges=197, rows=42, cols=52, sum_row=3487, sum_col=2166,
                sum_V=[18025, 14069, 7888, 15505, 11140, 13530, 16802, 17943],
                max_V=[254, 190, 158, 222, 143, 175, 205, 239],
                first=[0, 0, 9, 174, 0, 72, 3, 156, 143, 145], last=[41, 51, 0, 0, 0, 0, 0, 0, 0, 0],
                sha256="a058c8507148dca1641ca1fc370f98739bca27532651119f9ae2212c7718bc92",
                row_deg=[8, 10, 8, 10, 4, 6, 6, 6, 4, 5, 5, 5, 4, 5, 5, 4, 5, 5, 4, 4, 4, 4, 3, 4, 4, 3, 5, 3, 4, 3, 5,
                         3, 4, 4, 4, 4, 4, 3, 4, 4, 4, 4]),
    "k0_num": {"1": [0, 17, 33, 56], "2": [0, 13, 25, 43]},
    "crc_check_123456789": {"CRC16": 0x31C3, "CRC24A": 0xCDE703, "CRC24B": 0x23EF52},
}
(G / "tables.json").write_text(json.dumps(tables, indent=1))

# decoder known-answer vectors (oracle A)
rng = np.random.default_rng(20261017)
cases = {}
for name, (bg, Z, B, E, esn0, iters, et, rows, fill) in {
    "bg2_z6_plumbing": (2, 6, 16, 100, 1.0, 8, 1, 13, 24),
    "bg2_z52_small": (2, 52, 4, 2000, -2.5, 8, 0, 33, 104),
    "bg1_z48": (1, 48, 3, 3000, -0.5, 8, 0, 0, 40),
    "bg1_z384_r13": (1, 384, 2, 25272, -0.6, 8, 0, 46, 0),
    "bg1_z384_r89_et": (1, 384, 2, 9478, 5.5, 20, 1, 5, 0),
    "bg1_z7_et": (1, 7, 8, 300, 1.0, 12, 1, 0, 0),
}.items():
    d = O.dims(bg, Z)
    info = rng.integers(0, 2, (B, d["K"]), dtype=np.uint8)
    if fill:
        info[:, d["K"] - fill:] = 0
    cw = O.encode(bg, Z, info)
    s2 = 10 ** (-esn0 / 10)
    y = (1 - 2.0 * cw) / np.sqrt(2) + rng.normal(0, np.sqrt(s2 / 2), cw.shape)
    llr = (2 * np.sqrt(2) * y / s2).astype(np.float32)
    llr[:, :2 * Z] = 0
    llr[:, 2 * Z + E:] = 0
    if fill:
        llr[:, d["K"] - fill:d["K"]] = np.inf
    ref = O.decode_nms(bg, Z, llr, iters, early_term=bool(et), n_rows=rows)
    cases[name + "__cfg"] = np.array([bg, Z, iters, et, rows], dtype=np.int32)
    cases[name + "__llr"] = llr.astype(np.float16).astype(np.float32) if False else llr
    cases[name + "__hard"] = np.packbits(ref["hard"], axis=1)
    cases[name + "__iters"] = ref["iters"]
    cases[name + "__ok"] = ref["parity_ok"]
    # APP checksum instead of the full tensor: xor and sum of the bit patterns per codeword
    u = ref["app"].view(np.uint32)
    cases[name + "__app_xor"] = np.bitwise_xor.reduce(u, axis=1)
    cases[name + "__app_sum"] = u.astype(np.uint64).sum(axis=1)
np.savez_compressed(G / "decode_nms.npz", **cases)
print("golden written:", sorted(p.name for p in G.iterdir()))
