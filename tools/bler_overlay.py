#!/usr/bin/env python3
"""BLER-vs-SNR overlay on identical noise: CUDA engine (float32) vs oracle A (same algorithm: must be IDENTICAL counts),
the packed-half CUDA kernel (checked bit-exact against oracle A16 on the first batch of every point), and the reference's
algorithm (flooding sum-product f64, whole H) both as the CUDA NRLDPC_ALG_BP kernel and as oracle B on the CPU: their
decisions and iteration counts must be IDENTICAL on the first batch of every point (oracle B is slow), the curve is then
counted with the CUDA kernel and reported as a dB delta against the layered min-sum default.  GPU box only.
Writes gpurun_out/bler_overlay.txt (copied to profiles/ by hand)."""
import sys, json, time
sys.path.insert(0, '.')
import numpy as np
import torch
from ldpc_3gpp_matlab_b200 import capi
from ldpc_3gpp_matlab_b200.bler import BlerSimulator
from oracle import oracle as O

def interp_db(rows, target, col):
    xs = [r[0] for r in rows]; ys = [max(r[col], 1e-9) for r in rows]
    for (x0, y0), (x1, y1) in zip(zip(xs, ys), zip(xs[1:], ys[1:])):
        if y0 >= target >= y1 and y0 != y1:
            return x0 + (x1 - x0) * (np.log10(y0) - np.log10(target)) / (np.log10(y0) - np.log10(y1))
    return None

out = []
for name, A, R, BG, snrs, nbatch, B in (("cfgP_bg2_A20_r15", 20, 0.2, 2, np.arange(0.0, 4.01, 0.5), 12, 8192),
                                       ("cfgS_bg2_A400_r15", 400, 0.2, 2, np.arange(-3.5, -1.49, 0.25), 2, 4096),
                                       ("cfgH_bg1_A8424_r13", 8424, 1 / 3, 1, np.arange(-0.9, 0.61, 0.15), 1, 2048)):
    sim = BlerSimulator(A, R, BG, iterations=8, early_termination=True, batch=B, seed=1)
    h16 = capi.Handle(BG, sim.Z, 8, True, llr_dtype=capi.F16X2)
    hard16 = torch.empty_like(sim.hard)
    hbp = capi.Handle(BG, sim.Z, 8, True, algorithm=capi.ALG_BP)
    hard_bp = torch.empty_like(sim.hard)
    iters_bp = torch.empty_like(sim.iters)
    rows = []
    for s in snrs:
        e_gpu = e_a = e_b = e_h = n = 0
        for bi in range(nbatch):
            sim.run_batch(float(s))
            llr = sim.llr.cpu().numpy(); info = sim.info.cpu().numpy(); hard = sim.hard.cpu().numpy()
            Kp = sim.Kp
            e_gpu += int((hard[:, :Kp] != info[:, :Kp]).any(1).sum())
            ra = O.decode_nms(BG, sim.Z, llr, 8, early_term=True, n_rows=sim.n_rows, want_app=False)
            e_a += int((ra["hard"][:, :Kp] != info[:, :Kp]).any(1).sum())
            assert (ra["hard"] == hard).all(), "CUDA and oracle A differ on identical LLRs"
            h16.decode_raw(sim.llr, sim.llr.shape[0], hard16, n_rows=sim.n_rows, mem=capi.MEM_DEVICE,
                           stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            hh = hard16.cpu().numpy()
            e_h += int((hh[:, :Kp] != info[:, :Kp]).any(1).sum())
            if bi == 0:
                r16 = O.decode_nms(BG, sim.Z, llr[:512], 8, early_term=True, n_rows=sim.n_rows, want_app=False, f16=True)
                assert (r16["hard"] == hh[:512]).all(), "packed-half CUDA and oracle A16 differ on identical LLRs"
            hbp.decode_raw(sim.llr, sim.llr.shape[0], hard_bp, iters=iters_bp, n_rows=0, mem=capi.MEM_DEVICE,
                           stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            hb = hard_bp.cpu().numpy()
            e_b += int((hb[:, :Kp] != info[:, :Kp]).any(1).sum())
            if bi == 0:
                nb = min(1024, llr.shape[0])
                rb = O.decode_bp(BG, sim.Z, llr[:nb], 8)
                assert (rb["hard"] == hb[:nb]).all() and (rb["iters"] == iters_bp.cpu().numpy()[:nb]).all(), \
                    "CUDA sum-product and oracle B differ on identical LLRs"
            n += B
        rows.append((float(s), e_gpu / n, e_a / n, e_h / n, e_b / n, n))
        print(name, rows[-1], flush=True)
    d = {"config": name, "rows": rows}
    for tgt in (1e-1, 1e-2):
        g, hf, b = interp_db(rows, tgt, 1), interp_db(rows, tgt, 3), interp_db(rows, tgt, 4)
        d[f"esn0_at_bler_{tgt:g}"] = {"cuda_layered_nms_f32": g, "cuda_layered_nms_f16x2": hf, "reference_algorithm_flooding_bp_f64": b,
                                      "delta_db_bp_minus_f32": (b - g) if g is not None and b is not None else None,
                                      "delta_db_f16x2_minus_f32": (hf - g) if g is not None and hf is not None else None}
    out.append(d)
    sim.close()
    h16.close()
    hbp.close()
txt = ["# BLER on identical noise, 8 iterations, early termination: columns EsN0_dB, BLER(CUDA f32), BLER(oracle A), BLER(CUDA f16x2), BLER(reference algorithm = CUDA NRLDPC_ALG_BP, identical to oracle B on the first <=1024 blocks of every point), blocks"]
for d in out:
    txt.append(f"## {d['config']}")
    for r in d["rows"]:
        txt.append("%6.2f\t%.4e\t%.4e\t%.4e\t%.4e\t%d" % r)
    for k, v in d.items():
        if k.startswith("esn0_at"):
            txt.append(f"{k}: {json.dumps(v)}")
open("gpurun_out/bler_overlay.txt", "w").write("\n".join(txt) + "\n")
print("\n".join(txt))
