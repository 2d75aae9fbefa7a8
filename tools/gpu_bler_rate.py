#!/usr/bin/env python3
"""Secondary figure of SURVEY.md 8(d): end-to-end encode -> rate match -> QPSK/AWGN/LLR -> rate recover -> decode -> CRC
frames per second of the device-resident BLER loop (plot_BLER_vs_SNR.m protocol), one GPU, headline code."""
import sys, time, json
sys.path.insert(0, ".")
import torch
from ldpc_3gpp_matlab_b200 import capi
from ldpc_3gpp_matlab_b200.bler import BlerSimulator

rows = []
for name, (A, R, BG, esn0, B) in {"cfgH_bg1_A8424_r13": (8424, 1 / 3, 1, -0.3, 4096), "cfgS_bg2_A400_r15": (400, 0.2, 2, -2.0, 32768),
                                  "cfgP_bg2_A20_r15": (20, 0.2, 2, 2.0, 65536)}.items():
    for alg, algname in ((capi.ALG_NMS, "layered NMS f32"), (capi.ALG_BP, "sum-product f64 (reference algorithm)")):
        sim = BlerSimulator(A, R, BG, iterations=8, early_termination=True, batch=B, seed=1, algorithm=alg)
        sim.run_batch(esn0)
        torch.cuda.synchronize()
        n = 5 if alg == capi.ALG_NMS else 2
        t0 = time.perf_counter()
        tot = 0
        for _ in range(n):
            c, _ = sim.run_batch(esn0)
            tot += c
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        rows.append({"config": name, "algorithm": algname, "batch": B, "esn0_db": esn0, "ms_per_batch": round(dt * 1e3, 3),
                     "frames_per_s": round(B / dt), "info_Gbps": round(B * A / dt / 1e9, 4), "bler": float(tot[1] / tot[0]),
                     "mean_iters": float(tot[3] / tot[0] / sim.C)})
        print(rows[-1], flush=True)
        sim.close()
json.dump(rows, open("gpurun_out/bler_rate.json", "w"), indent=1)
