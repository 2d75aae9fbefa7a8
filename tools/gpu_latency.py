#!/usr/bin/env python3
"""Latency of small synchronous host-memory calls (the reference calls step() with ONE codeword, NRLDPCDecoder.m:265).
Median wall time of nrldpc_decode with pinned host buffers, BG1 Z=384 rate 1/3, 8 iterations."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import torch
import bench
from ldpc_3gpp_matlab_b200 import capi

w = dict(bench.WORKLOADS[bench.DEFAULT_WORKLOAD]); w["batch"] = 256
st = torch.cuda.current_stream().cuda_stream
h0 = capi.Handle(w["bg"], w["Z"], 8, False)
info, llr = bench.make_inputs(h0, capi, torch, w, 1, st)
rows = []
for name, kw in (("nms_f32", {}), ("nms_f16x2", dict(llr_dtype=capi.F16X2)), ("bp_f64", dict(algorithm=capi.ALG_BP))):
    h = capi.Handle(w["bg"], w["Z"], 8, name == "bp_f64", **kw)
    for B in (1, 2, 8, 64, 256):
        lh = torch.empty((B, h.n_cw), dtype=torch.float32, pin_memory=True); lh.copy_(llr[:B])
        hh = torch.empty((B, h.K), dtype=torch.uint8, pin_memory=True)
        for _ in range(5):
            h.decode_raw(lh, B, hh, mem=capi.MEM_HOST)
        ts = []
        for _ in range(100 if name != "bp_f64" else 20):
            t0 = time.perf_counter(); h.decode_raw(lh, B, hh, mem=capi.MEM_HOST); ts.append(time.perf_counter() - t0)
        rows.append({"mode": name, "batch": B, "median_us": round(1e6 * float(np.median(ts)), 1), "min_us": round(1e6 * min(ts), 1)})
        print(rows[-1], flush=True)
    h.close()
json.dump(rows, open("gpurun_out/latency.json", "w"), indent=1)
