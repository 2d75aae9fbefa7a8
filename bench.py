#!/usr/bin/env python3
"""bench.py -- headline benchmark of the NR LDPC decode hot path (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one decode of one batch of synthetic codewords: BG1, Z=384, K=8448, rate 1/3
(E=25272 of N=25344 bits sent, all 46 base rows active), 8 layered normalized-min-sum iterations,
early termination off, batch 4096 codewords per GPU (weak scaling: every rank decodes its own 4096).

  value   decoded information Gb/s with the LLRs already resident in HBM (CUDA events, max over ranks)
  e2e     same metric through the C ABI with pinned HOST buffers: H2D of the float32 LLRs, the kernel
          and D2H of the hard bits all inside the timed region
  roofline      algorithmic HBM bytes of the decode kernel / its measured launch time vs the measured
                copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference's algorithm (flooding sum-product, float64, parity-check stop: what
                comm.LDPCDecoder runs at NRLDPCDecoder.m:120,265) restated in C (oracle B), timed on
                this box's host cores on a bounded sample of the same LLRs
`--impl reference` times that CPU path alone (the reference itself is MATLAB + a closed toolbox and
cannot run here).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: bg, Z, E, n_rows, iters, early_term, batch/GPU, Es/N0 dB, filler
    "bg1_z384_r13_it8_b4096": dict(bg=1, Z=384, E=25272, n_rows=46, iters=8, early_term=0, batch=4096, esn0=-0.3, filler=0),
    "bg2_z52_r15_it8_b65536": dict(bg=2, Z=52, E=2000, n_rows=33, iters=8, early_term=0, batch=65536, esn0=-2.0, filler=104),
    "bg1_z384_r13_it8et_b4096": dict(bg=1, Z=384, E=25272, n_rows=46, iters=8, early_term=1, batch=4096, esn0=-0.3, filler=0),
    # the stop armed but never taken (no block converges at -3 dB): every CTA runs 8 iterations plus 8 failing syndromes, in phase
    "bg1_z384_r13_it8et_lowsnr_b4096": dict(bg=1, Z=384, E=25272, n_rows=46, iters=8, early_term=1, batch=4096, esn0=-3.0, filler=0),
    # four times the headline batch: how much of the stop path's cost is end-of-launch imbalance
    "bg1_z384_r13_it8_b16384": dict(bg=1, Z=384, E=25272, n_rows=46, iters=8, early_term=0, batch=16384, esn0=-0.3, filler=0),
    "bg1_z384_r13_it8et_b16384": dict(bg=1, Z=384, E=25272, n_rows=46, iters=8, early_term=1, batch=16384, esn0=-0.3, filler=0),
    "bg2_z52_r15_it8et_b65536": dict(bg=2, Z=52, E=2000, n_rows=33, iters=8, early_term=1, batch=65536, esn0=-2.0, filler=104),
    # config 3 with the stop armed but never taken (nothing converges at -6 dB): 8 iterations + 8 failing syndromes per block
    "bg2_z52_r15_it8et_lowsnr_b65536": dict(bg=2, Z=52, E=2000, n_rows=33, iters=8, early_term=1, batch=65536, esn0=-6.0, filler=104),
    "bg1_z384_r89_it20et_b4096": dict(bg=1, Z=384, E=9478, n_rows=5, iters=20, early_term=1, batch=4096, esn0=6.3, filler=0),
}
DEFAULT_WORKLOAD = "bg1_z384_r13_it8_b4096"
METRIC = "decoded info Gb/s @ BG1 Z=384 8-iter"
CPU_SAMPLE_CW = 256          # codewords per step of the --impl reference arm
CPU_BASELINE_CW = 2048       # codewords of the cpu_baseline leg (about 4 s on 16 cores, 60+ core-seconds)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md): an NVML polling thread
    (about 1 kHz), with the recipe's nvidia-smi loop as the fallback when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = str(gpu_index)
        self.rows = []          # (sm_mhz, power_w, reasons bitmask)
        self.max_mhz = None
        self.p = self.f = self.thread = None
        self.stop_flag = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.dev) / 1000.0
                except Exception:
                    pw = 0.0
                self.rows.append((float(mhz), pw, int(reasons)))
            except Exception:
                break
            time.sleep(0.001)

    def start(self):
        if self.nv is not None:
            import threading
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(2)
            rows = self.rows
            if rows:
                sm = sorted(r[0] for r in rows)
                out.update(sm_mhz=sm[len(sm) // 2], power_w_max=max(r[1] for r in rows), samples=len(rows), source="nvml")
                bits = 0
                for r in rows:
                    bits |= r[2]
                names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
                out["reasons"] = sorted(n for b, n in names.items() if bits & b)
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        rows = [r for r in rows if len(r) >= 9 and r[0] == self.idx]
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(rows[0][2]), power_w_max=max(float(r[3]) for r in rows),
                   samples=len(rows), source="nvidia-smi")
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        seen = set()
        for r in rows:
            for n, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    seen.add(n)
        out["reasons"] = sorted(seen)
        return out


def make_inputs(h, capi, torch, w, seed, stream):
    """Synthetic workload generated ON DEVICE through the library's own chain kernels:
    random info -> encode -> rate match -> QPSK + AWGN + exact LLR -> rate recover (decoder layout)."""
    B, E = w["batch"], w["E"]
    g = torch.Generator(device="cuda").manual_seed(seed)
    info = torch.randint(0, 2, (B, h.K), dtype=torch.uint8, device="cuda", generator=g)
    if w["filler"]:
        info[:, h.K - w["filler"]:] = 0
    cw = torch.empty((B, h.n_cw), dtype=torch.uint8, device="cuda")
    f = torch.empty((B, E), dtype=torch.uint8, device="cuda")
    fl = torch.empty((B, E), dtype=torch.float32, device="cuda")
    llr = torch.empty((B, h.n_cw), dtype=torch.float32, device="cuda")
    rm = capi.Rm(E, 0, h.N, h.K - w["filler"], 2)
    h.encode_raw(info, B, cw, mem=capi.MEM_DEVICE, stream=stream)
    h.rate_match_raw(cw, B, rm, f, mem=capi.MEM_DEVICE, stream=stream)
    h.qpsk_awgn_llr_raw(f, B, E, 10 ** (-w["esn0"] / 10), seed, 0, fl, stream=stream)
    h.rate_recover_raw(fl, B, rm, None, llr, mem=capi.MEM_DEVICE, stream=stream)
    torch.cuda.synchronize()
    del cw, f, fl
    return info, llr


def cpu_reference_time(w, llr_sample, threads, steps=1, warmup=0):
    """Oracle B (the reference's algorithm) on `threads` host threads.  Returns (sec/step, info bits/step)."""
    from oracle import oracle as O
    O.build()
    for _ in range(warmup):
        O.decode_bp(w["bg"], w["Z"], llr_sample, w["iters"], n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = O.decode_bp(w["bg"], w["Z"], llr_sample, w["iters"], n_threads=threads)
    dt = (time.perf_counter() - t0) / max(1, steps)
    cpu_reference_time.last = out
    return dt, llr_sample.shape[0] * O.dims(w["bg"], w["Z"])["K"]


def synth_llr_cpu(w, n, seed):
    """Host-side synthetic LLRs for the reference arm (no GPU involved): oracle encoder + numpy AWGN."""
    import numpy as np
    from oracle import oracle as O
    O.build()
    rng = np.random.default_rng(seed)
    d = O.dims(w["bg"], w["Z"])
    info = rng.integers(0, 2, (n, d["K"]), dtype=np.uint8)
    if w["filler"]:
        info[:, d["K"] - w["filler"]:] = 0
    cw = O.encode(w["bg"], w["Z"], info)
    s2 = 10 ** (-w["esn0"] / 10)
    y = (1 - 2.0 * cw) / np.sqrt(2) + rng.normal(0, np.sqrt(s2 / 2), cw.shape)
    llr = (2 * np.sqrt(2) * y / s2).astype(np.float32)
    llr[:, :2 * w["Z"]] = 0
    llr[:, 2 * w["Z"] + w["E"]:] = 0
    if w["filler"]:
        llr[:, d["K"] - w["filler"]:d["K"]] = np.inf
    return llr


CONFIG4 = "bg1_z384_r89_it20et_b4096"
CONFIG4_ESN0 = (6.25, 7.25)      # BLER ~ 1e-2 (oracle A, 3000 blocks: 0.035 at 6.2 dB, 0.0033 at 6.3 dB) and +1 dB


def run_config4(capi, torch, D, rank, local_rank, world, stream, steps=20, warmup=3):
    """BASELINE.json configs[3]: BG1 Z=384 rate 8/9 (E=9478, 5 active rows), <= 20 iterations with the reference's
    parity-check stop (NRLDPCDecoder.m:120), the batch sharded over the ranks (4096 blocks each, no data-path
    collective).  Time depends on Es/N0 under the stop, so two points are reported, each with its mean iteration count
    and the per-rank launch times: their spread is the inter-rank tail imbalance SURVEY 8(e) predicts."""
    w = WORKLOADS[CONFIG4]
    h = capi.Handle(w["bg"], w["Z"], w["iters"], True, device=local_rank)
    B, K = w["batch"], h.K
    hard = torch.empty((B, K), dtype=torch.uint8, device="cuda")
    iters_t = torch.empty(B, dtype=torch.int32, device="cuda")
    points = []
    for esn0 in CONFIG4_ESN0:
        ww = dict(w, esn0=esn0)
        info, llr = make_inputs(h, capi, torch, ww, (D.rank_seed(4, rank) + int(esn0 * 100)) & 0x7FFFFFFF, stream)

        def step():
            h.decode_raw(llr, B, hard, iters=iters_t, n_rows=w["n_rows"], mem=capi.MEM_DEVICE, stream=stream)
        for _ in range(warmup):
            step()
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ms_max, ms_min = D.max_over_ranks(ms), -D.max_over_ranks(-ms)
        tot = D.sum_counters([int(iters_t.sum()), int((hard != info).any(dim=1).sum()), B]).tolist()
        points.append({"esn0_db": esn0, "value": world * B * K / (ms_max * 1e-3) / 1e9, "unit": "Gb/s", "ms_per_step": ms_max,
                       "rank_ms_min": ms_min, "rank_ms_max": ms_max, "rank_imbalance": ms_max / ms_min - 1.0,
                       "mean_iters": tot[0] / tot[2], "bler": tot[1] / tot[2], "blocks": tot[2]})
        del info, llr
    h.close()
    return {"workload": CONFIG4, "n_gpus": world, "batch_per_gpu": B, "max_iters": w["iters"], "early_term": 1, "n_rows": w["n_rows"],
            "E": w["E"], "dtype": "f32", "steps": steps, "points": points,
            "note": "BASELINE config 4; device-resident, CUDA events, max over ranks; value counts K = 8448 bits per block"}


def run_bler_loop(torch, D, rank, local_rank, world, batches=8):
    """The Monte-Carlo loop of plot_BLER_vs_SNR.m:104-171 on device at one Es/N0 point of the headline code (A = 8424,
    BG1, R = 1/3, QPSK, 8 iterations with the parity-check stop): every rank simulates its own frames with its own random
    stream (:23-27) and the four counters {blocks, block errors, bit errors, iterations} are summed over the ranks with one
    32-byte all-reduce (NCCL).  At N > 1 rank 0 re-simulates the first batch of every rank alone and checks that the
    counters are identical (same seeds x ranks => same frames, whatever the number of GPUs)."""
    import numpy as np
    from ldpc_3gpp_matlab_b200.bler import BlerSimulator
    A, R, BG, esn0, B, seed = 8424, 1 / 3, 1, -0.3, 4096, 2
    sim = BlerSimulator(A, R, BG, iterations=8, early_termination=True, batch=B, seed=seed, device=local_rank, rank=rank, world=world)
    first, _ = sim.run_batch(esn0)
    for _ in range(2):
        sim.run_batch(esn0)
    D.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tot = np.zeros(4, dtype=np.int64)
    for _ in range(batches):
        c, _ = sim.run_batch(esn0)
        tot += c
    glob = D.sum_counters(tot).numpy()             # the reference's manual aggregation, as one all-reduce
    torch.cuda.synchronize()
    dt = D.max_over_ranks(time.perf_counter() - t0)
    sim.close()
    # per-rank first-batch counters, gathered by summing a one-hot table
    table = np.zeros((world, 4), dtype=np.int64)
    table[rank] = first
    table = D.sum_counters(table.ravel()).numpy().reshape(world, 4)
    same = None
    if rank == 0:
        same = True
        for r in range(1, world):
            s2 = BlerSimulator(A, R, BG, iterations=8, early_termination=True, batch=B, seed=seed, device=local_rank, rank=r, world=world)
            c, _ = s2.run_batch(esn0)
            s2.close()
            same = same and bool((c == table[r]).all())
    return {"code": "A=8424 BG1 R=1/3 QPSK, 8 iterations, parity-check stop, CRC24A checked on device", "esn0_db": esn0, "n_gpus": world,
            "batch_per_gpu": B, "batches_per_gpu": batches, "frames_per_s": world * batches * B / dt, "payload_gbps": world * batches * B * A / dt / 1e9,
            "ms_per_batch": dt / batches * 1e3, "counters": {"blocks": int(glob[0]), "block_errors": int(glob[1]), "bit_errors": int(glob[2]),
                                                           "iterations": int(glob[3])},
            "bler": float(glob[1] / glob[0]), "mean_iters": float(glob[3] / glob[0]),
            "collective": "one all_reduce(SUM) of 4 x int64 = 32 bytes per point (%s)" % ("nccl" if world > 1 else "single process"),
            "per_rank_counters_equal_single_gpu_rerun": same}


def csrc_digest():
    """sha256 over the kernel sources: ties an ncu capture (profiles/traffic.json) to the binary that was profiled."""
    import hashlib
    hsh = hashlib.sha256()
    for f in sorted((ROOT / "ldpc_3gpp_matlab_b200" / "csrc").glob("*")):
        if f.suffix in (".cu", ".cuh", ".inc", ".cpp", ".h"):
            hsh.update(f.read_bytes())
    return hsh.hexdigest()[:16]


def run_reference(args, w, wname):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    llr = synth_llr_cpu(w, CPU_SAMPLE_CW, 1234)
    dt, bits = cpu_reference_time(w, llr, threads, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    val = bits / dt / 1e9
    cfg_batch = w["batch"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gb/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wname, "bg": w["bg"], "Z": w["Z"], "K": 22 * w["Z"] if w["bg"] == 1 else 10 * w["Z"],
                   "E": w["E"], "iters": w["iters"], "early_term": 1, "batch_per_gpu": cfg_batch,
                   "algorithm": "flooding sum-product f64, parity-check stop (comm.LDPCDecoder as at NRLDPCDecoder.m:120), "
                                "C restatement: MATLAB and the toolbox are not runnable here"},
        "cpu_baseline": {"value": val, "unit": "Gb/s", "cores": threads, "kind": "port",
                         "sample": f"{CPU_SAMPLE_CW} codewords of the workload per step, OpenMP over codewords"},
        "e2e": {"value": val, "unit": "Gb/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--llr-dtype", default="f32", choices=["f32", "f16x2"],
                    help="decoder arithmetic: float32 (default, headline) or packed fp16, two codewords per thread")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the secondary packed-half measurement")
    ap.add_argument("--no-side", action="store_true", help="skip the config4 / bler_loop side keys")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args, w, args.workload)

    import numpy as np
    import torch
    from ldpc_3gpp_matlab_b200 import capi, dist as D

    rank, local_rank, world = D.init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys_gpu = int(visible.split(",")[local_rank]) if visible and visible.split(",")[local_rank].isdigit() else local_rank
    numa_bound = D.bind_to_gpu_numa_node(phys_gpu) if world > 1 else False
    stream = torch.cuda.current_stream().cuda_stream
    # host staging threads of the host-memory path (nrldpc_decode64 on pageable doubles): share the cores between the ranks
    os.environ.setdefault("NRLDPC_HOST_THREADS", str(max(1, min(32, (os.cpu_count() or 1) // world))))

    f16 = args.llr_dtype == "f16x2"
    h = capi.Handle(w["bg"], w["Z"], w["iters"], bool(w["early_term"]), device=local_rank,
                    llr_dtype=capi.F16X2 if f16 else capi.F32)
    B, K = w["batch"], h.K
    info, llr = make_inputs(h, capi, torch, w, D.rank_seed(0, rank) & 0x7FFFFFFF, stream)
    hard = torch.empty((B, K), dtype=torch.uint8, device="cuda")
    iters_t = torch.empty(B, dtype=torch.int32, device="cuda")

    def step():
        h.decode_raw(llr, B, hard, iters=iters_t if w["early_term"] else None, n_rows=w["n_rows"],
                     mem=capi.MEM_DEVICE, stream=stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    bler = float((hard != info).any(dim=1).float().mean())
    mean_iters = float(iters_t.float().mean()) if w["early_term"] else float(w["iters"])
    iters_hist = torch.bincount(iters_t, minlength=w["iters"] + 1).tolist() if w["early_term"] else None

    # ---- timed region: K steps, CUDA events on the launching stream, barrier + sync both sides ----
    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    l0 = h.launches
    D.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    D.barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = h.launches - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = D.max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    value = world * B * K / (ms_per_step * 1e-3) / 1e9

    # ---- secondary figure: the same workload through the packed-half kernel (not the headline) ----
    alt = None
    if not f16 and not args.no_alt:
        h2 = capi.Handle(w["bg"], w["Z"], w["iters"], bool(w["early_term"]), device=local_rank, llr_dtype=capi.F16X2)
        hard2 = torch.empty_like(hard)

        def step2():
            h2.decode_raw(llr, B, hard2, iters=iters_t if w["early_term"] else None, n_rows=w["n_rows"],
                          mem=capi.MEM_DEVICE, stream=stream)
        for _ in range(args.warmup):
            step2()
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step2()
        e1.record()
        torch.cuda.synchronize()
        ms2 = D.max_over_ranks(e0.elapsed_time(e1)) / args.steps
        alt = {"llr_dtype": "f16x2", "dtype": "f16", "value": world * B * K / (ms2 * 1e-3) / 1e9, "unit": "Gb/s",
               "ms_per_step": ms2, "bler_at_esn0": float((hard2 != info).any(dim=1).float().mean()),
               "note": "two codewords per thread in packed fp16 (bit-exact against its own binary16 oracle); "
                       "reported beside the float32 headline, not instead of it"}
        launches_alt = h2.launches
        h2.close()
        del hard2

    # ---- secondary figure: the REFERENCE's algorithm on the device (flooding sum-product, float64, parity-check
    # stop, whole H: NRLDPC_ALG_BP) on the same LLRs -- the like-for-like twin of the --impl reference arm
    bp = None
    if not f16 and not args.no_alt:
        hb = capi.Handle(w["bg"], w["Z"], w["iters"], True, device=local_rank, algorithm=capi.ALG_BP)
        hard_bp = torch.empty_like(hard)
        iters_bp = torch.empty(B, dtype=torch.int32, device="cuda")

        def step_bp():
            hb.decode_raw(llr, B, hard_bp, iters=iters_bp, n_rows=0, mem=capi.MEM_DEVICE, stream=stream)
        step_bp()
        D.barrier()
        torch.cuda.synchronize()
        n_bp = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_bp):
            step_bp()
        e1.record()
        torch.cuda.synchronize()
        ms_bp = D.max_over_ranks(e0.elapsed_time(e1)) / n_bp
        bp = {"algorithm": "flooding sum-product f64, parity-check stop, all base rows (NRLDPC_ALG_BP = comm.LDPCDecoder "
                           "as configured at NRLDPCDecoder.m:120)", "dtype": "f64", "value": world * B * K / (ms_bp * 1e-3) / 1e9,
              "unit": "Gb/s", "ms_per_step": ms_bp, "steps": n_bp, "mean_iters": float(iters_bp.float().mean()),
              "bler_at_esn0": float((hard_bp != info).any(dim=1).float().mean())}
        hb.close()

    # ---- e2e: pinned host buffers through the synchronous host-memory C-ABI call ----------------
    e2e = None
    if not args.no_e2e:
        llr_h = torch.empty((B, h.n_cw), dtype=torch.float32, pin_memory=True)
        llr_h.copy_(llr)
        hard_h = torch.empty((B, K), dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        for _ in range(2):
            h.decode_raw(llr_h, B, hard_h, n_rows=w["n_rows"], mem=capi.MEM_HOST)
        n_e2e = max(3, min(args.steps, 10))
        D.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            h.decode_raw(llr_h, B, hard_h, n_rows=w["n_rows"], mem=capi.MEM_HOST)
        torch.cuda.synchronize()
        dt = D.max_over_ranks((time.perf_counter() - t0) / n_e2e)
        assert bool((hard_h.cuda() == hard).all()), "host-path result differs from device-path result"
        # the PCIe floor of that call: the same pinned buffer copied to the device and nothing else (CUDA events)
        dst = torch.empty_like(llr)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dst.copy_(llr_h, non_blocking=True)
        torch.cuda.synchronize()
        c0.record()
        for _ in range(3):
            dst.copy_(llr_h, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_only_ms = c0.elapsed_time(c1) / 3
        del dst
        e2e = {"value": world * B * K / dt / 1e9, "unit": "Gb/s", "h2d_bytes_per_step": B * h.n_cw * 4,
               "d2h_bytes_per_step": B * K, "ms_per_step": dt * 1e3, "steps": n_e2e,
               "timer": "host perf_counter around the synchronous host-memory C-ABI call, max over ranks",
               "pipeline": "3 streams, chunked H2D / kernel / D2H overlap inside nrldpc_decode",
               "numa_bound": bool(numa_bound), "h2d_copy_alone_ms": h2d_only_ms,
               "h2d_copy_alone_gbs": B * h.n_cw * 4 / (h2d_only_ms * 1e-3) / 1e9,
               "note": "PCIe-bound: h2d_copy_alone_ms is the same pinned float32 buffer copied to the device with nothing else running"}
        # same call with the LLRs transported as binary16 (nrldpc_decode16): half the H2D bytes; reported beside the
        # float32-boundary figure above, which stays the e2e headline
        llr_h16 = torch.empty((B, h.n_cw), dtype=torch.float16, pin_memory=True)
        llr_h16.copy_(llr.clamp(-60000.0, 60000.0))
        torch.cuda.synchronize()
        for _ in range(2):
            h.decode16_raw(llr_h16, B, hard_h, n_rows=w["n_rows"], mem=capi.MEM_HOST)
        D.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            h.decode16_raw(llr_h16, B, hard_h, n_rows=w["n_rows"], mem=capi.MEM_HOST)
        torch.cuda.synchronize()
        dt16 = D.max_over_ranks((time.perf_counter() - t0) / n_e2e)
        e2e["f16_transport"] = {"value": world * B * K / dt16 / 1e9, "unit": "Gb/s", "h2d_bytes_per_step": B * h.n_cw * 2,
                                "d2h_bytes_per_step": B * K, "ms_per_step": dt16 * 1e3,
                                "bler_at_esn0": float((hard_h.cuda() != info).any(dim=1).float().mean())}
        del llr_h16
        # and as 8-bit integers (nrldpc_decode8, llr = q / 8): a quarter of the float32 bytes -- the transfer then hides behind the
        # kernel.  Quantisation is the caller's choice and costs BLER (reported); the float32 figure above stays the e2e headline.
        i8_scale = 0.125
        llr_h8 = torch.empty((B, h.n_cw), dtype=torch.int8, pin_memory=True)
        llr_h8.copy_((llr / i8_scale).round().clamp(-127, 126).to(torch.int8))
        torch.cuda.synchronize()
        for _ in range(2):
            h.decode8_raw(llr_h8, i8_scale, B, hard_h, n_rows=w["n_rows"], mem=capi.MEM_HOST)
        D.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            h.decode8_raw(llr_h8, i8_scale, B, hard_h, n_rows=w["n_rows"], mem=capi.MEM_HOST)
        torch.cuda.synchronize()
        dt8 = D.max_over_ranks((time.perf_counter() - t0) / n_e2e)
        e2e["i8_transport"] = {"value": world * B * K / dt8 / 1e9, "unit": "Gb/s", "h2d_bytes_per_step": B * h.n_cw,
                               "d2h_bytes_per_step": B * K, "ms_per_step": dt8 * 1e3, "scale": i8_scale,
                               "bler_at_esn0": float((hard_h.cuda() != info).any(dim=1).float().mean()),
                               "note": "LLRs quantised to int8 by the caller (q = round(8 * llr), |q| <= 127); decoded in float32"}
        del llr_h8
        # what the MEX gateway really calls (matlab/nrldpc_mex.cpp): nrldpc_decode64 on ORDINARY pageable float64 memory,
        # decisions into pageable memory.  Host threads narrow to float32 into a pinned ring, so PCIe still carries 4 B / LLR.
        llr64 = llr_h.numpy().astype(np.float64)
        hard_np = np.empty((B, K), dtype=np.uint8)
        for _ in range(2):
            h.decode64_raw(llr64, B, hard_np, n_rows=w["n_rows"], mem=capi.MEM_HOST)
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            h.decode64_raw(llr64, B, hard_np, n_rows=w["n_rows"], mem=capi.MEM_HOST)
        dt64 = D.max_over_ranks((time.perf_counter() - t0) / n_e2e)
        assert bool((torch.from_numpy(hard_np).cuda() == hard).all()), "float64 pageable host path differs from the device path"
        e2e["f64_pageable"] = {"value": world * B * K / dt64 / 1e9, "unit": "Gb/s", "ms_per_step": dt64 * 1e3,
                               "host_bytes_read_per_step": B * h.n_cw * 8, "h2d_bytes_per_step": B * h.n_cw * 4, "d2h_bytes_per_step": B * K,
                               "host_threads": int(os.environ["NRLDPC_HOST_THREADS"]), "ratio_to_pinned_f32": dt / dt64,
                               "call": "nrldpc_decode64(NRLDPC_MEM_HOST) on numpy (pageable) float64 LLRs and a pageable uint8 result: "
                                       "the call matlab/nrldpc_mex.cpp makes on mxGetPr memory"}
        del llr_h, hard_h, llr64, hard_np

    # ---- roofline of the dominant (only) kernel --------------------------------------------------
    peak, peak_src = peaks()
    bytes_per_cw = 4 * h.n_cw + K
    k_ms = sum(per_launch_ms) / len(per_launch_ms)
    achieved = B * bytes_per_cw / (k_ms * 1e-3) / 1e9
    traffic = pipes = traffic_src = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            tj = json.loads(tp.read_text())
            traffic = tj.get(args.workload + ("|f16x2" if f16 else ""))
            pipes = (tj.get("pipes") or {}).get(args.workload + ("|f16x2" if f16 else ""))
            traffic_src = {"file": "profiles/traffic.json", "capture": tj.get("_capture"), "captured_csrc_sha16": tj.get("_csrc_sha16"),
                           "current_csrc_sha16": csrc_digest(), "current_binary_profiled": tj.get("_csrc_sha16") == csrc_digest()}
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "decode_nms_h2_kernel" if f16 else "decode_nms_kernel", "kernel_ms": k_ms, "bytes_per_codeword": bytes_per_cw,
                "peak_source": peak_src,
                "binding_pipe": {"name": "SM ALU pipe (integer / logic / min-max / select)", "ncu": pipes,
                                 "source": "profiles/traffic.json <- ncu --set full capture of this kernel"},
                "note": "state stays on chip for all iterations; the kernel is bound by the SM's ALU pipe by construction "
                        "(SURVEY.md 8d), so the HBM fraction is small; binding_pipe carries the ncu utilisation of the pipe that binds"}

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = llr[:CPU_BASELINE_CW].cpu().numpy()
        dt, bits = cpu_reference_time(w, sample, threads)
        cpu = {"value": bits / dt / 1e9, "unit": "Gb/s", "cores": threads, "kind": "port",
               "sample": f"first {CPU_BASELINE_CW} codewords of the batch, oracle B = flooding sum-product f64 with parity-check "
                         f"stop (comm.LDPCDecoder's algorithm), OpenMP over codewords, {dt:.2f} s"}
        rb_all = cpu_reference_time.last
        # the same on ONE host thread (SURVEY 8d asks for both): the first 48 codewords of the sample
        dt1, bits1 = cpu_reference_time(w, sample[:48], 1)
        cpu["one_thread"] = {"value": bits1 / dt1 / 1e9, "unit": "Gb/s", "cores": 1, "sample": f"first 48 codewords, {dt1:.2f} s"}
        cpu_reference_time.last = rb_all
        from oracle import oracle as O
        t0 = time.perf_counter()
        ref = O.decode_nms(w["bg"], w["Z"], sample, w["iters"], early_term=bool(w["early_term"]), n_rows=w["n_rows"],
                           want_app=False, n_threads=threads, f16=f16)
        dta = time.perf_counter() - t0
        if bp is not None:     # the device's sum-product kernel against the CPU restatement on the same codewords
            rb = cpu_reference_time.last
            bp["matches_cpu_reference_bits"] = bool((rb["hard"] == hard_bp[:CPU_BASELINE_CW].cpu().numpy()).all() and
                                                    (rb["iters"] == iters_bp[:CPU_BASELINE_CW].cpu().numpy()).all())
        cpu["like_for_like_nms_" + ("f16" if f16 else "f32")] = {"value": bits / dta / 1e9, "unit": "Gb/s", "cores": threads,
                                        "matches_gpu_bits": bool((ref["hard"] == hard[:CPU_BASELINE_CW].cpu().numpy()).all())}

    # ---- side keys, at every N (so that the driver's scaling record carries them): BASELINE config 4 and the BLER loop
    config4 = bler_loop = None
    if not f16 and not args.no_side and args.workload == DEFAULT_WORKLOAD:
        config4 = run_config4(capi, torch, D, rank, local_rank, world, stream)
        bler_loop = run_bler_loop(torch, D, rank, local_rank, world)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Gb/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if f16 else "f32", "data": "synthetic",
            "config": {"workload": args.workload, "llr_dtype": args.llr_dtype, "bg": w["bg"], "Z": w["Z"], "K": K, "N": h.N, "E": w["E"],
                       "rate": round(K / w["E"], 4), "n_rows": w["n_rows"], "iters": w["iters"], "early_term": w["early_term"],
                       "alpha": 0.75, "algorithm": "layered normalized min-sum", "batch_per_gpu": B,
                       "global_batch": B * world, "parallelism": f"dp{world} (independent codeword shards, no data-path collective)",
                       "esn0_db": w["esn0"], "bler_at_esn0": bler, "mean_iters": mean_iters, "iters_hist": iters_hist,
                       "l2": f"inputs larger than L2 ({B * h.n_cw * 4 / 2**20:.0f} MiB LLRs per step vs 126 MB L2)"},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        }
        if args.workload != DEFAULT_WORKLOAD:
            line["metric"] = "decoded info Gb/s @ " + args.workload + " (parity-test configuration, not the headline)"
        if f16:
            line["metric"] += " (packed fp16 arithmetic)"
        if alt is not None:
            line["f16x2"] = alt
        if bp is not None:
            line["reference_algorithm_on_gpu"] = bp
        if config4 is not None:
            line["config4"] = config4
        if bler_loop is not None:
            line["bler_loop"] = bler_loop
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
