#!/usr/bin/env python3
"""One point of the config-5 sweep (same input generation as tools/sweep.py), for ncu captures and A/B timing.

    python tools/gpu_point.py --bg 1 --Z 208 --rate 1/3 --dtype f32 [--mb 100] [--reps 5] [--early-term]
Prints one JSON line: ms per launch, Gb/s, launch geometry is in the ncu capture.
"""
from __future__ import annotations

import argparse
import json
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

ESN0 = {1: {"1/3": 0.0, "1/2": 2.0, "2/3": 4.0, "8/9": 7.5}, 2: {"1/3": 0.5, "1/2": 2.5, "2/3": 4.5, "8/9": 8.5}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bg", type=int, default=1)
    ap.add_argument("--Z", type=int, default=384)
    ap.add_argument("--rate", default="1/3")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f16x2"])
    ap.add_argument("--mb", type=float, default=100.0)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--early-term", action="store_true")
    ap.add_argument("--esn0", type=float, default=None)
    args = ap.parse_args()

    import torch
    from ldpc_3gpp_matlab_b200 import capi

    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    bg, Z = args.bg, args.Z
    rn, rd = (int(x) for x in args.rate.split("/"))
    kcols, rows_all = (22, 46) if bg == 1 else (10, 42)
    h = capi.Handle(bg, Z, args.iters, args.early_term, device=0, llr_dtype=capi.F16X2 if args.dtype == "f16x2" else capi.F32)
    K, N, ncw = h.K, h.N, h.n_cw
    B = args.batch or max(64, int(args.mb * 1e6 / (ncw * 4)) // 2 * 2)
    g = torch.Generator(device="cuda").manual_seed(bg * 1000 + Z)
    info = torch.randint(0, 2, (B, K), dtype=torch.uint8, device="cuda", generator=g)
    cw = torch.empty((B, ncw), dtype=torch.uint8, device="cuda")
    h.encode_raw(info, B, cw, mem=capi.MEM_DEVICE, stream=stream)
    llr = torch.empty((B, ncw), dtype=torch.float32, device="cuda")
    hard = torch.empty((B, K), dtype=torch.uint8, device="cuda")
    iters = torch.empty(B, dtype=torch.int32, device="cuda")
    E = min(N, 2 * int(math.floor(K * rd / (2 * rn) + 0.5)))
    while (B * E) % 4:
        B += 1
    n_rows = int(min(rows_all, max(4, -(-(E + 2 * Z) // Z) - kcols)))
    rm = capi.Rm(E, 0, N, K, 2)
    f = torch.empty((B, E), dtype=torch.uint8, device="cuda")
    fl = torch.empty((B, E), dtype=torch.float32, device="cuda")
    esn0 = ESN0[bg][args.rate] if args.esn0 is None else args.esn0
    h.rate_match_raw(cw, B, rm, f, mem=capi.MEM_DEVICE, stream=stream)
    h.qpsk_awgn_llr_raw(f, B, E, 10 ** (-esn0 / 10), 1234, bg * 100000 + Z * 10 + rn, fl, stream=stream)
    h.rate_recover_raw(fl, B, rm, None, llr, mem=capi.MEM_DEVICE, stream=stream)

    def step():
        h.decode_raw(llr, B, hard, iters=iters if args.early_term else None, n_rows=n_rows, mem=capi.MEM_DEVICE, stream=stream)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    print(json.dumps({"bg": bg, "Z": Z, "rate": args.rate, "dtype": args.dtype, "batch": B, "n_rows": n_rows, "E": E, "ms": round(ms, 4),
                      "gbps": round(B * K / (ms * 1e-3) / 1e9, 3), "esn0": esn0,
                      "mean_iters": float(iters.float().mean()) if args.early_term else args.iters,
                      "bler": float((hard != info).any(dim=1).float().mean())}), flush=True)
    h.close()


if __name__ == "__main__":
    main()
