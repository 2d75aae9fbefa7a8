#!/usr/bin/env python3
"""Writes tests/golden/matlab/inputs.mat: the committed decoder-input LLRs of tests/golden/decode_nms.npz (cw_tilde
layout of NRLDPCDecoder.m:262-264, +Inf filler kept) as MATLAB doubles, one struct per case, for
matlab/make_golden_vectors.m -- the script a Communications Toolbox licence holder runs to pin oracle B and the CUDA
NRLDPC_ALG_BP kernel to comm.LDPCDecoder itself (tests/test_matlab_golden.py consumes its output when present)."""
import sys
from pathlib import Path

import numpy as np
import scipy.io

ROOT = Path(__file__).resolve().parent.parent
G = ROOT / "tests" / "golden"
src = np.load(G / "decode_nms.npz")
names = sorted({k.split("__")[0] for k in src.files})
cases = np.zeros(len(names), dtype=[("name", object), ("BG", object), ("Z", object), ("iterations", object), ("cw_tilde", object)])
for i, n in enumerate(names):
    bg, Z, iters, et, rows = src[n + "__cfg"].tolist()
    cases[i] = (n, float(bg), float(Z), float(iters), src[n + "__llr"].astype(np.float64).T.copy())   # (n_cw x batch), one cw_tilde per column
scipy.io.savemat(G / "matlab" / "inputs.mat", {"cases": cases}, do_compression=True, oned_as="column")
print("wrote", G / "matlab" / "inputs.mat", names)
