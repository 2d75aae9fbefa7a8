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

# the call the MEX gateway makes for one code block: nrldpc_decode64 on ordinary (pageable) float64 memory, decisions into
# pageable memory; and the same pinned-float32 call with the direct host-memory path switched off (A/B)
import os
llr64 = llr[:1].cpu().numpy().astype(np.float64)
hard_np = np.empty((1, h0.K), dtype=np.uint8)
for label, env in (("nms_f32 decode64 pageable float64 (gateway call)", None), ("nms_f32 pinned, NRLDPC_ZERO_COPY_MAX=0", "0")):
    if env is not None:
        os.environ["NRLDPC_ZERO_COPY_MAX"] = env
    hh = capi.Handle(w["bg"], w["Z"], 8, False)
    lh = torch.empty((1, hh.n_cw), dtype=torch.float32, pin_memory=True); lh.copy_(llr[:1])
    ph = torch.empty((1, hh.K), dtype=torch.uint8, pin_memory=True)
    call = (lambda: hh.decode64_raw(llr64, 1, hard_np, mem=capi.MEM_HOST)) if env is None else (lambda: hh.decode_raw(lh, 1, ph, mem=capi.MEM_HOST))
    for _ in range(5):
        call()
    ts = []
    for _ in range(100):
        t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
    rows.append({"mode": label, "batch": 1, "median_us": round(1e6 * float(np.median(ts)), 1), "min_us": round(1e6 * min(ts), 1)})
    print(rows[-1], flush=True)
    hh.close()
    os.environ.pop("NRLDPC_ZERO_COPY_MAX", None)

# where the time of a one-codeword call goes: the kernel alone on device buffers (CUDA events over back-to-back launches),
# and the reference's whole per-block RX chain (NRLDPCDecoder.m:133-140: rate recovery -> decode -> CRC) on device buffers
h = capi.Handle(w["bg"], w["Z"], 8, False)
hard = torch.empty((256, h.K), dtype=torch.uint8, device="cuda")
for B in (1, 8):
    for _ in range(10):
        h.decode_raw(llr, B, hard, n_rows=46, mem=capi.MEM_DEVICE, stream=st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        h.decode_raw(llr, B, hard, n_rows=46, mem=capi.MEM_DEVICE, stream=st)
    e1.record(); torch.cuda.synchronize()
    rows.append({"mode": "nms_f32 kernel only (device buffers, back-to-back)", "batch": B, "us_per_launch": round(1e3 * e0.elapsed_time(e1) / 200, 1)})
    print(rows[-1], flush=True)
E = w["E"]
rm = capi.Rm(E, 0, h.N, h.K, 2)
fl = torch.randn((1, E), dtype=torch.float32, device="cuda") * 3 + 2
out = torch.empty((1, h.n_cw), dtype=torch.float32, device="cuda")
ok = torch.zeros(1, dtype=torch.uint8, device="cuda")
ts = []
for i in range(120):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.rate_recover_raw(fl, 1, rm, None, out, mem=capi.MEM_DEVICE, stream=st)
    h.decode_raw(out, 1, hard, n_rows=46, mem=capi.MEM_DEVICE, stream=st)
    h.crc_raw(hard, 1, h.K, h.K, capi.CRC24A, ok=ok, stream=st)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
rows.append({"mode": "one-block RX chain on device buffers: rate_recover + decode + crc, 3 launches + sync", "batch": 1,
             "median_us": round(1e6 * float(np.median(ts[20:])), 1)})
print(rows[-1], flush=True)
h.close()
json.dump(rows, open("gpurun_out/latency.json", "w"), indent=1)
