#!/usr/bin/env python3
"""BASELINE.json config 5: decode throughput over BG1+BG2 x all 51 lifting sizes x rates {1/3,1/2,2/3,8/9},
8 fixed layered-NMS iterations, as a Gb/s table (float32 and packed-half kernels).

    python tools/sweep.py [--mb 100] [--reps 5] [--out gpurun_out/sweep]           (one GPU)
    python -m torch.distributed.run --nproc-per-node N ... tools/sweep.py          (N GPUs, weak scaling)

Per point: K = 22Z / 10Z uniform random bits (no filler), encoded on device,
E = 2*round(K/(2R)) bits (capped at N) taken from k0 = 0, QPSK + AWGN + exact LLRs, rate recovery into the
decoder layout; active rows = max(4, ceil((E + 2Z)/Z) - kcols); batch sized to about --mb MB of LLRs per GPU.
Timing: CUDA events around --reps back-to-back launches after 3 warm-ups, max over ranks; Gb/s counts K bits.
"""
from __future__ import annotations

import argparse
import json
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

ALL_Z = sorted(a << j for a, n in zip((2, 3, 5, 7, 9, 11, 13, 15), (8, 8, 7, 6, 6, 6, 5, 5)) for j in range(n))
RATES = ((1, 3), (1, 2), (2, 3), (8, 9))
ESN0 = {1: {(1, 3): 0.0, (1, 2): 2.0, (2, 3): 4.0, (8, 9): 7.5}, 2: {(1, 3): 0.5, (1, 2): 2.5, (2, 3): 4.5, (8, 9): 8.5}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=100.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "sweep"))
    ap.add_argument("--zs", default="", help="comma-separated subset of lifting sizes")
    args = ap.parse_args()

    import torch
    from ldpc_3gpp_matlab_b200 import capi, dist as D

    rank, local_rank, world = D.init()
    torch.cuda.set_device(local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    zs = [int(z) for z in args.zs.split(",")] if args.zs else ALL_Z
    rows = []
    for bg in (1, 2):
        kcols, rows_all = (22, 46) if bg == 1 else (10, 42)
        for Z in zs:
            hs = {dt: capi.Handle(bg, Z, args.iters, False, device=local_rank, llr_dtype=v)
                  for dt, v in (("f32", capi.F32), ("f16x2", capi.F16X2))}
            h = hs["f32"]
            K, N, ncw = h.K, h.N, h.n_cw
            B = max(64, int(args.mb * 1e6 / (ncw * 4)) // 2 * 2)
            g = torch.Generator(device="cuda").manual_seed((D.rank_seed(bg * 1000 + Z, rank)) & 0x7FFFFFFF)
            info = torch.randint(0, 2, (B, K), dtype=torch.uint8, device="cuda", generator=g)
            cw = torch.empty((B, ncw), dtype=torch.uint8, device="cuda")
            h.encode_raw(info, B, cw, mem=capi.MEM_DEVICE, stream=stream)
            llr = torch.empty((B, ncw), dtype=torch.float32, device="cuda")
            hard = torch.empty((B, K), dtype=torch.uint8, device="cuda")
            for (rn, rd) in RATES:
                E = min(N, 2 * int(math.floor(K * rd / (2 * rn) + 0.5)))
                n_rows = int(min(rows_all, max(4, -(-(E + 2 * Z) // Z) - kcols)))
                rm = capi.Rm(E, 0, N, K, 2)
                f = torch.empty((B, E), dtype=torch.uint8, device="cuda")
                fl = torch.empty((B, E), dtype=torch.float32, device="cuda")
                h.rate_match_raw(cw, B, rm, f, mem=capi.MEM_DEVICE, stream=stream)
                if (B * E) % 4:
                    raise SystemExit("batch*E must be a multiple of 4")
                h.qpsk_awgn_llr_raw(f, B, E, 10 ** (-ESN0[bg][(rn, rd)] / 10), 1234 + rank, bg * 100000 + Z * 10 + rn, fl, stream=stream)
                h.rate_recover_raw(fl, B, rm, None, llr, mem=capi.MEM_DEVICE, stream=stream)
                rec = {"bg": bg, "Z": Z, "rate": f"{rn}/{rd}", "K": K, "E": E, "n_rows": n_rows, "batch_per_gpu": B, "n_gpus": world}
                for dt, hh in hs.items():
                    for _ in range(3):
                        hh.decode_raw(llr, B, hard, n_rows=n_rows, mem=capi.MEM_DEVICE, stream=stream)
                    D.barrier()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(args.reps):
                        hh.decode_raw(llr, B, hard, n_rows=n_rows, mem=capi.MEM_DEVICE, stream=stream)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = D.max_over_ranks(e0.elapsed_time(e1)) / args.reps
                    rec[dt + "_ms"] = round(ms, 4)
                    rec[dt + "_gbps"] = round(world * B * K / (ms * 1e-3) / 1e9, 3)
                    rec[dt + "_bler"] = round(float((hard != info).any(dim=1).float().mean()), 5)
                rows.append(rec)
                del f, fl
            for hh in hs.values():
                hh.close()
            del info, cw, llr, hard
            torch.cuda.empty_cache()
    if rank == 0:
        out = Path(args.out)
        out.parent.mkdir(parents=True, exist_ok=True)
        with open(str(out) + ".jsonl", "w") as fjs:
            for r in rows:
                fjs.write(json.dumps(r) + "\n")
        with open(str(out) + ".md", "w") as fmd:
            fmd.write(f"# Decode throughput sweep: {world} x B200, {args.iters} fixed iterations, ~{args.mb:.0f} MB of LLRs per GPU per launch\n\n")
            fmd.write("Gb/s of decoded information bits (K per codeword), whole job; float32 kernel / packed-half kernel.\n\n")
            for bg in (1, 2):
                fmd.write(f"## BG{bg}\n\n| Z | " + " | ".join(f"R={rn}/{rd} f32 | f16x2" for rn, rd in RATES) + " |\n|---|" + "---|---|" * len(RATES) + "\n")
                for Z in zs:
                    cells = []
                    for rn, rd in RATES:
                        r = next(x for x in rows if x["bg"] == bg and x["Z"] == Z and x["rate"] == f"{rn}/{rd}")
                        cells.append(f"{r['f32_gbps']:.2f} | {r['f16x2_gbps']:.2f}")
                    fmd.write(f"| {Z} | " + " | ".join(cells) + " |\n")
                fmd.write("\n")
        print(open(str(out) + ".md").read())
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
