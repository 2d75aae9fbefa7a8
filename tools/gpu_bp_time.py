#!/usr/bin/env python3
"""Times the sum-product (NRLDPC_ALG_BP) kernel on the headline batch (dev tool; NRLDPC_BP_THREADS selects the CTA width)."""
import sys, os
sys.path.insert(0, ".")
import torch
import bench
from ldpc_3gpp_matlab_b200 import capi

w = dict(bench.WORKLOADS[bench.DEFAULT_WORKLOAD])
w["batch"] = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
st = torch.cuda.current_stream().cuda_stream
h0 = capi.Handle(w["bg"], w["Z"], 8, False)
info, llr = bench.make_inputs(h0, capi, torch, w, 1, st)
h = capi.Handle(w["bg"], w["Z"], 8, True, algorithm=capi.ALG_BP)
B = w["batch"]
hard = torch.empty((B, h.K), dtype=torch.uint8, device="cuda")
it = torch.empty(B, dtype=torch.int32, device="cuda")
for n in range(1 + int(os.environ.get("BP_STEPS", 3))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h.decode_raw(llr, B, hard, iters=it, mem=capi.MEM_DEVICE, stream=st)
    e1.record()
    torch.cuda.synchronize()
    print("bp threads", os.environ.get("NRLDPC_BP_THREADS", "1024 (default)"), "batch", B, "ms", round(e0.elapsed_time(e1), 3),
          "Gb/s", round(B * h.K / e0.elapsed_time(e1) / 1e6, 4), "mean iters", float(it.float().mean()))
