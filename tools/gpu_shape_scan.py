#!/usr/bin/env python3
"""CTA-shape scan of the decode kernels (dev tool): for every requested lifting size, time the config-5 sweep point
(BG1/BG2, rate 1/3, 8 fixed iterations, ~100 MB of LLRs) for each codewords-per-CTA choice and occupancy cap
(NRLDPC_CWPC / NRLDPC_OCC_CAP are read at nrldpc_create), check that every shape returns identical decisions, and
print one JSON line per shape.  The table this produces is what choose_decode_cwpc() in nrldpc_b200.cu is fitted to."""
import argparse, json, math, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bg", type=int, default=1)
    ap.add_argument("--zs", default="all")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--mb", type=float, default=100.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--early-term", action="store_true")
    ap.add_argument("--caps", default="4,8")
    ap.add_argument("--esn0", type=float, default=None)
    ap.add_argument("--extra-caps", default="", help="occupancy caps above 8, tried for CTAs of at most 96 threads (plus CTA widths 32 / 64 / 96)")
    ap.add_argument("--only-extra", action="store_true", help="time only the default shape and the --extra-caps candidates")
    args = ap.parse_args()
    import torch
    from ldpc_3gpp_matlab_b200 import capi
    torch.cuda.set_device(0)
    st = torch.cuda.current_stream().cuda_stream
    bg = args.bg
    kcols, rows_all = (22, 46) if bg == 1 else (10, 42)
    all_z = sorted(a << j for a, n in zip((2, 3, 5, 7, 9, 11, 13, 15), (8, 8, 7, 6, 6, 6, 5, 5)) for j in range(n))
    for Z in (all_z if args.zs == "all" else [int(z) for z in args.zs.split(",")]):
        for k in ("NRLDPC_CWPC", "NRLDPC_OCC_CAP"):
            os.environ.pop(k, None)
        h0 = capi.Handle(bg, Z, 8, args.early_term, device=0)
        K, N, ncw = h0.K, h0.N, h0.n_cw
        B = max(64, int(args.mb * 1e6 / (ncw * 4)) // 2 * 2)
        g = torch.Generator(device="cuda").manual_seed(bg * 1000 + Z)
        info = torch.randint(0, 2, (B, K), dtype=torch.uint8, device="cuda", generator=g)
        cw = torch.empty((B, ncw), dtype=torch.uint8, device="cuda")
        h0.encode_raw(info, B, cw, mem=capi.MEM_DEVICE, stream=st)
        E = min(N, 2 * int(math.floor(K * 3 / 2 + 0.5)))
        while (B * E) % 4:
            B -= 2
        n_rows = int(min(rows_all, max(4, -(-(E + 2 * Z) // Z) - kcols)))
        rm = capi.Rm(E, 0, N, K, 2)
        f = torch.empty((B, E), dtype=torch.uint8, device="cuda")
        fl = torch.empty((B, E), dtype=torch.float32, device="cuda")
        llr = torch.empty((B, ncw), dtype=torch.float32, device="cuda")
        h0.rate_match_raw(cw[:B], B, rm, f, mem=capi.MEM_DEVICE, stream=st)
        h0.qpsk_awgn_llr_raw(f, B, E, 10 ** (-((0.0 if bg == 1 else 0.5) if args.esn0 is None else args.esn0) / 10), 1234, Z, fl, stream=st)
        h0.rate_recover_raw(fl, B, rm, None, llr, mem=capi.MEM_DEVICE, stream=st)
        h0.close()
        ref = None
        cmax = max(1, 384 // Z)
        # one candidate per CTA width: the most codewords that fit T = 128, 160, ... 384 threads (plus one codeword per CTA)
        cws = sorted({1, cmax} | {T // Z for T in range(128, 385, 32) if 1 <= T // Z <= cmax})
        caps = [int(x) for x in args.caps.split(",")]
        cands = [("default", None, None)]
        for c in cws:
            T = max(32, (c * Z + 31) // 32 * 32)
            for cap in caps:
                if cap > 4 and T > 192:      # a cap above 4 only matters for narrow CTAs
                    continue
                cands.append((f"cw{c}_cap{cap}", c, cap))
        if args.only_extra:
            cands = cands[:1]
        if args.extra_caps:
            for c in sorted({T // Z for T in (32, 64, 96) if 1 <= T // Z <= cmax} | {1}):
                T = max(32, (c * Z + 31) // 32 * 32) if c > 1 else Z
                if T > 96:
                    continue
                for cap in (int(x) for x in args.extra_caps.split(",")):
                    cands.append((f"cw{c}_cap{cap}", c, cap))
        for name, c, cap in cands:
            for k in ("NRLDPC_CWPC", "NRLDPC_OCC_CAP"):
                os.environ.pop(k, None)
            if c:
                os.environ["NRLDPC_CWPC"] = str(c); os.environ["NRLDPC_OCC_CAP"] = str(cap)
            try:
                h = capi.Handle(bg, Z, 8, args.early_term, device=0, llr_dtype=capi.F16X2 if args.dtype == "f16x2" else capi.F32)
            except Exception as e:
                print(json.dumps({"Z": Z, "shape": name, "error": str(e)[:80]})); continue
            hard = torch.empty((B, K), dtype=torch.uint8, device="cuda")
            try:
                for _ in range(3):
                    h.decode_raw(llr, B, hard, n_rows=n_rows, mem=capi.MEM_DEVICE, stream=st)
                torch.cuda.synchronize()
            except Exception as e:
                print(json.dumps({"Z": Z, "shape": name, "error": str(e)[:80]})); h.close(); continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                h.decode_raw(llr, B, hard, n_rows=n_rows, mem=capi.MEM_DEVICE, stream=st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            if ref is None:
                ref = hard.clone()
            same = bool((hard == ref).all())
            print(json.dumps({"bg": bg, "Z": Z, "dtype": args.dtype, "shape": name, "cwpc": c, "cap": cap, "batch": B, "ms": round(ms, 4),
                              "gbps": round(B * K / (ms * 1e-3) / 1e9, 3), "same_bits": same}), flush=True)
            h.close()


if __name__ == "__main__":
    main()
