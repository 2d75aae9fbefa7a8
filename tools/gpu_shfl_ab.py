#!/usr/bin/env python3
"""A/B of the check-node mappings for small lifting sizes: thread = check with a register scan (default kernels) against
lane = (check, edge) with a warp-shuffle butterfly (NRLDPC_DECODE_VARIANT=shfl, decode_kernel_shfl.cuh).
Per (BG, Z): kernel-only time of ONE codeword (device buffers, back-to-back launches -- the latency case where lanes would
otherwise be empty) and throughput of a large batch (the case the north star's metric is about).  Rate 1/3, 8 iterations."""
import json, math, os, sys
sys.path.insert(0, ".")
import torch
from ldpc_3gpp_matlab_b200 import capi

st = torch.cuda.current_stream().cuda_stream
rows_out = []
for bg in (1, 2):
    kcols, rows_all = (22, 46) if bg == 1 else (10, 42)
    for Z in (2, 4, 6, 8, 12, 16, 24, 30, 32):
        os.environ.pop("NRLDPC_DECODE_VARIANT", None)
        h0 = capi.Handle(bg, Z, 8, False)
        K, N, ncw = h0.K, h0.N, h0.n_cw
        B = max(64, int(100e6 / (ncw * 4)) // 4 * 4)
        g = torch.Generator(device="cuda").manual_seed(bg * 1000 + Z)
        info = torch.randint(0, 2, (B, K), dtype=torch.uint8, device="cuda", generator=g)
        cw = torch.empty((B, ncw), dtype=torch.uint8, device="cuda")
        h0.encode_raw(info, B, cw, mem=capi.MEM_DEVICE, stream=st)
        E = min(N, 2 * int(math.floor(K * 3 / 2 + 0.5)))
        n_rows = int(min(rows_all, max(4, -(-(E + 2 * Z) // Z) - kcols)))
        rm = capi.Rm(E, 0, N, K, 2)
        f = torch.empty((B, E), dtype=torch.uint8, device="cuda")
        fl = torch.empty((B, E), dtype=torch.float32, device="cuda")
        llr = torch.empty((B, ncw), dtype=torch.float32, device="cuda")
        h0.rate_match_raw(cw, B, rm, f, mem=capi.MEM_DEVICE, stream=st)
        h0.qpsk_awgn_llr_raw(f, B, E, 10 ** (-(1.0 if bg == 1 else 1.5) / 10), 1234, Z, fl, stream=st)
        h0.rate_recover_raw(fl, B, rm, None, llr, mem=capi.MEM_DEVICE, stream=st)
        h0.close()
        ref = None
        rec = {"bg": bg, "Z": Z, "batch": B, "n_rows": n_rows}
        for name, env in (("scan", {}), ("shfl", {"NRLDPC_DECODE_VARIANT": "shfl"}),
                          ("shfl_cw4x", {"NRLDPC_DECODE_VARIANT": "shfl", "NRLDPC_SHFL_CWPC": str(max(1, 128 // Z))})):
            for k in ("NRLDPC_DECODE_VARIANT", "NRLDPC_SHFL_CWPC"):
                os.environ.pop(k, None)
            os.environ.update(env)
            for et in (False, True):
                h = capi.Handle(bg, Z, 8, et)
                hard = torch.empty((B, K), dtype=torch.uint8, device="cuda")
                for nb, reps, key in ((1, 200, "us_1cw"), (B, 5, "gbps")):
                    for _ in range(3):
                        h.decode_raw(llr, nb, hard, n_rows=n_rows, mem=capi.MEM_DEVICE, stream=st)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(reps):
                        h.decode_raw(llr, nb, hard, n_rows=n_rows, mem=capi.MEM_DEVICE, stream=st)
                    e1.record(); torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / reps
                    rec[f"{name}{'_stop' if et else ''}_{key}"] = round(ms * 1e3, 1) if nb == 1 else round(B * K / (ms * 1e-3) / 1e9, 3)
                if not et:
                    if ref is None:
                        ref = hard.clone()
                    rec[f"{name}_same_bits"] = bool((hard == ref).all())
                h.close()
        print(json.dumps(rec), flush=True)
        rows_out.append(rec)
os.makedirs("gpurun_out/r02_shfl", exist_ok=True)
with open("gpurun_out/r02_shfl/ab.jsonl", "w") as fo:
    for r in rows_out:
        fo.write(json.dumps(r) + "\n")
