"""Device-resident BLER-vs-SNR Monte-Carlo loop: the protocol of plot_BLER_vs_SNR.m:104-171 with the
frame loop (:116) turned into batches that never leave the GPU; snr_vs_a() drives the same loop with the
protocol of plot_SNR_vs_A.m (required Es/N0 as a function of the block length).

Per batch: random information blocks -> LDPC encode -> rate match (bit selection + interleave) ->
QPSK + AWGN + exact LLR -> rate recover (+HARQ combine over the rv_id sequence, :124-137) -> decode
-> block-error count, every stage a C-ABI call on device pointers.  Batches shard across ranks
(one process per GPU, independent random streams per rank as the reference asks at :23-27); the only
collective is a sum of four counters per SNR point.

Block-error criterion.  The reference counts ~isequal(a, a_hat) with a_hat = [] whenever the TB CRC
or a CB CRC fails (NRLDPCDecoder.m:337-339, plot_BLER_vs_SNR.m:146).  With crc=True (default) exactly
that is evaluated on device: TB CRC attach, segmentation with CB CRC24B, and after decoding the CB and
TB CRC checks plus the payload comparison (nrldpc_crc).  With crc=False the K' bits of every code block
are drawn uniformly and a block error is "some code block has an error in its first K' decoded bits",
which is the same event except for undetected errors that leave the payload intact while corrupting
only CRC bits (the code is linear and the channel symmetric, so BLER does not depend on the word sent).

Results are written in the reference's file format ("%f\\t%e\\n" per SNR point, :165) so curves can be
overlaid file for file.
"""
from __future__ import annotations

import argparse
import math
import os
import sys
from pathlib import Path

import numpy as np

from . import capi, dist as D
from .nrldpc import NRLDPC, matlab_round


def active_rows(p: NRLDPC, E_max: int, rv_ids) -> int:
    """Base rows that can carry information for this rate-matching configuration (see nrldpc.NRLDPCDecoder)."""
    Z, kcols, rows_all = p.Z_c, (22 if p.BG == 1 else 10), (46 if p.BG == 1 else 42)
    nfill = max(0, p.K - max(p.K_prime, 2 * Z))
    hi = 0
    for rv in rv_ids:
        p.rv_id = rv
        span = p.k_0 + E_max + nfill
        hi = max(hi, p.N_cb if span >= p.N_cb else span)
    return int(min(rows_all, max(4, -(-(hi + 2 * Z) // Z) - kcols)))


class BlerSimulator:
    def __init__(self, A, R, BG, Q_m=2, rv_id_sequence=(0,), iterations=8, early_termination=True, alpha=0.75,
                 batch=4096, seed=0, device=0, rank=0, world=1, llr_dtype=capi.F32, decision_method=capi.DEMOD_LLR, crc=True,
                 algorithm=capi.ALG_NMS):
        import torch
        if Q_m not in (1, 2, 4, 6, 8):
            raise capi.UnsupportedParameters("Unsupported modulation")          # NRModulator.m:43
        self.Q_m, self.method, self.use_crc = int(Q_m), int(decision_method), bool(crc)
        self.torch = torch
        self.p = NRLDPC(A=A, BG=BG, G=matlab_round(A / R / Q_m) * Q_m, Q_m=Q_m)  # plot_BLER_vs_SNR.m:94
        self.p.validate_properties()
        p = self.p
        self.rvs = tuple(int(r) for r in rv_id_sequence)
        self.C, self.K, self.Kp, self.Z, self.N = p.C, p.K, p.K_prime, p.Z_c, p.N
        self.E_r = [int(e) for e in p.E_r]
        self.n_rows = active_rows(p, max(self.E_r), self.rvs)
        if algorithm == capi.ALG_BP:          # the reference decodes on the whole H (NRLDPCDecoder.m:120)
            self.n_rows = 46 if BG == 1 else 42
        p.rv_id = 0
        self.h = capi.Handle(BG, self.Z, iterations, early_termination, alpha, device=device, llr_dtype=llr_dtype,
                             algorithm=algorithm)
        self.B = int(batch)
        self.rank, self.world, self.seed = rank, world, seed
        self.stream_id = 0       # channel uses of this simulator: key of the AWGN generator
        self.n_batches = 0       # batches drawn: key of the information-bit generator (its own key space, bit 62 set)
        dev = "cuda"
        n = self.B * self.C
        self.info = torch.zeros((n, self.K), dtype=torch.uint8, device=dev)
        self.cw = torch.empty((n, self.h.n_cw), dtype=torch.uint8, device=dev)
        self.llr = torch.empty((n, self.h.n_cw), dtype=torch.float32, device=dev)
        self.hard = torch.empty((n, self.K), dtype=torch.uint8, device=dev)
        self.iters = torch.empty(n, dtype=torch.int32, device=dev)
        self.harq = torch.zeros((n, self.N), dtype=torch.float32, device=dev) if len(self.rvs) > 1 else None
        Emax = max(self.E_r)
        # CRC bookkeeping (NRLDPC.m:297-377): payload A, transport block B_ = A + L_tb, code-block CRC L_cb when C > 1
        self.A, self.Bsz = int(p.A), int(p.B)
        self.tb_kind = capi.CRC_KIND[p.transport_block_CRC]
        self.L_cb = 24 if self.C > 1 else 0
        self.tb = torch.zeros((self.B, self.Bsz), dtype=torch.uint8, device=dev)
        self.tb_hat = torch.zeros((self.B, self.Bsz), dtype=torch.uint8, device=dev)
        self.cb_flag = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.tb_flag = torch.zeros(self.B, dtype=torch.uint8, device=dev)
        self.f = torch.empty((self.B, Emax), dtype=torch.uint8, device=dev)
        self.fl = torch.empty((self.B, Emax), dtype=torch.float32, device=dev)
        self.counters = torch.zeros(4, dtype=torch.int64, device=dev)      # blocks, block errors, bit errors, iterations
        self.latch = torch.zeros(self.B, dtype=torch.uint8, device=dev)    # block decoded by some transmission so far
        self.cb_passed = torch.zeros(n, dtype=torch.uint8, device=dev)     # code_block_CRC_passed (NRLDPCDecoder.m:93)

    def close(self):
        self.h.close()

    def _channel_and_recover(self, f, E, rm, var, harq, llr_out, st):
        """f bits [B][E] -> decoder input llr_out [B][n_cw] (+ HARQ accumulate): modulate, AWGN, demodulate
        (plot_BLER_vs_SNR.m:130-132) and rate recovery (NRLDPCDecoder.m:172-242).  QPSK with the exact LLR runs as ONE kernel
        (nrldpc_qpsk_awgn_rate_recover): the received LLRs never go through HBM; bit-identical to the separate stages."""
        h, B = self.h, self.B
        self.stream_id += 1
        seed = D.rank_seed(self.seed, self.rank)
        if self.Q_m == 2 and self.method == capi.DEMOD_LLR and E % 4 == 0 and E * 4 <= 200 * 1024 and not os.environ.get("NRLDPC_BLER_UNFUSED"):
            h.qpsk_awgn_rate_recover_raw(f, B, rm, var, seed, self.stream_id, harq, llr_out, stream=st)
            return
        fl = self.fl[:, :E]
        if E != self.fl.shape[1]:
            fl = fl.contiguous()
        if self.Q_m == 2 and self.method == capi.DEMOD_LLR and (B * E) % 4 == 0:
            h.qpsk_awgn_llr_raw(f, B, E, var, seed, self.stream_id, fl, stream=st)
        else:                                    # modulate + AWGN + demodulate fused, any NRModulator setting
            h.mod_awgn_llr_raw(f, B * E, self.Q_m, var, self.method, seed, self.stream_id, fl, stream=st)
        h.rate_recover_raw(fl, B, rm, harq, llr_out, mem=capi.MEM_DEVICE, stream=st)

    def run_batch(self, esn0_db: float):
        """One batch of B transport blocks at Es/N0.  Returns [blocks, block_errors, bit_errors, iterations].
        Every stage is a C-ABI launch on device buffers; the block-error bookkeeping is one kernel per decoding attempt
        (nrldpc_bler_count) with a device-side latch, and the host reads four counters once per batch."""
        torch, h, B, C = self.torch, self.h, self.B, self.C
        st = torch.cuda.current_stream().cuda_stream
        var = 10 ** (-esn0_db / 10)                      # plot_BLER_vs_SNR.m:105-106
        self.n_batches += 1
        A, Kp, Lcb, K = self.A, self.Kp, self.L_cb, self.K
        if self.use_crc:
            # a -> b = [a ; TB CRC] (NRLDPCEncoder.m:70-89) -> C blocks of K'-L_cb bits + CB CRC24B (:92-124), on device.
            # One code block: the transport block IS the head of the code block row, so it is drawn and CRC'd in place.
            tb, tb_stride = (self.info, K) if C == 1 else (self.tb, self.Bsz)
            h.random_bits_raw(tb, B, A, tb_stride, D.rank_seed(self.seed, self.rank), (1 << 62) | self.n_batches, stream=st)   # a = round(rand(A,1)), :112
            h.crc_raw(tb, B, A, tb_stride, self.tb_kind, parity=tb.data_ptr() + A, parity_stride=tb_stride, stream=st)
            if C > 1:
                self.info.view(B, C, -1)[:, :, :Kp - Lcb] = self.tb.view(B, C, Kp - Lcb)
                h.crc_raw(self.info, B * C, Kp - Lcb, K, capi.CRC24B, parity=self.info.data_ptr() + Kp - Lcb,
                          parity_stride=K, stream=st)
        else:
            h.random_bits_raw(self.info, B * C, Kp, K, D.rank_seed(self.seed, self.rank), (1 << 62) | self.n_batches, stream=st)
        h.encode_raw(self.info, B * C, self.cw, mem=capi.MEM_DEVICE, stream=st)
        if self.harq is not None:
            self.harq.zero_()                            # reset(hDec), plot_BLER_vs_SNR.m:122
        self.counters.zero_()
        self.latch.zero_()
        if self.use_crc and C > 1:
            self.cb_passed.zero_()                       # reset(hDec): NRLDPCDecoder.m:353-354
            self.tb_hat.zero_()
        # code blocks are interleaved frame-major: block r of frame b sits at row b*C + r
        cw3 = self.cw.view(B, C, -1)
        llr3 = self.llr.view(B, C, -1)
        harq3 = self.harq.view(B, C, -1) if self.harq is not None else None
        for i_rv, rv in enumerate(self.rvs):             # HARQ loop, :124-137
            last = i_rv + 1 == len(self.rvs)
            self.p.rv_id = rv
            for r in range(C):
                E = self.E_r[r]
                rm = capi.Rm(E, int(self.p.k_0), int(self.p.N_cb), int(Kp), self.Q_m)
                cw_r = cw3[:, r].contiguous() if C > 1 else self.cw
                f = self.f[:, :E]
                if E != self.f.shape[1]:
                    f = f.contiguous()
                h.rate_match_raw(cw_r, B, rm, f, mem=capi.MEM_DEVICE, stream=st)
                if C > 1:
                    llr_r = torch.empty((B, h.n_cw), dtype=torch.float32, device="cuda")
                    hq = harq3[:, r].contiguous() if harq3 is not None else None
                    self._channel_and_recover(f, E, rm, var, hq, llr_r, st)
                    llr3[:, r] = llr_r
                    if hq is not None:
                        harq3[:, r] = hq
                else:
                    self._channel_and_recover(f, E, rm, var, self.harq, self.llr, st)
            h.decode_raw(self.llr, B * C, self.hard, iters=self.iters, n_rows=self.n_rows, mem=capi.MEM_DEVICE, stream=st)
            if self.use_crc:
                # a_hat = [] unless every CB CRC and the TB CRC pass (NRLDPCDecoder.m:300,336-339); a block error is
                # ~isequal(a, a_hat) (plot_BLER_vs_SNR.m:146)
                if C == 1:
                    tb_hat, tb_hat_stride, cb_passed = self.hard, K, None
                else:
                    # per code block, latched over the retransmissions (NRLDPCDecoder.m:296-309): a block whose CB CRC
                    # passes overwrites its part of b_hat_buffer and sets code_block_CRC_passed; block 0 may pass on rv0
                    # and block 1 on rv1
                    h.crc_raw(self.hard, B * C, Kp, K, capi.CRC24B, ok=self.cb_flag, stream=st)
                    pass_now = self.cb_flag.view(B, C).bool()
                    tb3 = self.tb_hat.view(B, C, Kp - Lcb)
                    torch.where(pass_now[:, :, None], self.hard.view(B, C, -1)[:, :, :Kp - Lcb], tb3, out=tb3)
                    self.cb_passed |= self.cb_flag
                    tb_hat, tb_hat_stride, cb_passed = self.tb_hat, self.Bsz, self.cb_passed
                h.crc_raw(tb_hat, B, self.Bsz, tb_hat_stride, self.tb_kind, ok=self.tb_flag, stream=st)
                h.bler_count_raw(self.hard, self.info, tb_hat, tb_hat_stride, tb, tb_stride, self.tb_flag, cb_passed, self.iters,
                                 B, C, Kp, A, self.latch, self.counters, do_latch=True, finalize=last, stream=st)
            else:
                # no CRCs: a block is decoded when the first K' bits of all its code blocks are right
                cb_ok = (self.hard[:, :Kp] == self.info[:, :Kp]).all(dim=1).view(B, C).all(dim=1)
                self.latch |= cb_ok.to(torch.uint8)
                self.counters[3] += self.iters.sum()
                if last:
                    self._finalize_torch()
            if not last and bool(self.latch.all()):      # every block decoded: no further retransmission (:124)
                if self.use_crc:
                    h.bler_count_raw(self.hard, self.info, None, 0, None, 0, None, None, None, B, C, Kp, A, self.latch, self.counters,
                                     do_latch=False, finalize=True, stream=st)
                else:
                    self._finalize_torch()
                break
        c = self.counters.cpu().numpy().astype(np.int64)     # the one host read of the batch
        return c, bool(c[1] < B)

    def _finalize_torch(self):
        B, C, Kp = self.B, self.C, self.Kp
        bad = self.latch == 0
        self.counters[0] += B
        self.counters[1] += bad.sum()
        self.counters[2] += ((self.hard[:, :Kp] != self.info[:, :Kp]).view(B, C, -1).sum(dim=(1, 2)) * bad).sum()

    def run_point(self, esn0_db, target_block_errors, max_blocks=None, found_start=True):
        """Batches until `target_block_errors` errors were seen across all ranks (plot_BLER_vs_SNR.m:116)."""
        tot = np.zeros(4, dtype=np.int64)
        while True:
            c, any_ok = self.run_batch(esn0_db)
            c = D.sum_counters(c).numpy()
            tot += c
            if not found_start:
                any_ok_all = bool(D.sum_counters([int(any_ok)])[0] > 0)
                if not any_ok_all:          # start detection, :139-144: nothing decodes yet at this SNR
                    return tot, False
                found_start = True
            if tot[1] >= target_block_errors or (max_blocks and tot[0] >= max_blocks):
                return tot, True


_MOD_NAME = {1: "BPSK", 2: "QPSK", 4: "16QAM", 6: "64QAM", 8: "256QAM"}


def sweep(A, R, BG, iterations=8, target_block_errors=100, target_BLER=1e-3, EsN0_start=0.0, EsN0_delta=0.5, seed=0,
          rv_id_sequence=(0,), batch=4096, max_blocks=None, early_termination=True, out_dir="results", log=print,
          Q_m=2, llr_dtype=capi.F32, algorithm=capi.ALG_NMS):
    """plot_BLER_vs_SNR.m:53-171 for one (A, R, BG): returns [(EsN0, BLER, blocks, errors, mean_iters)]."""
    rank, local_rank, world = D.init()
    import torch
    torch.cuda.set_device(local_rank)
    sim = BlerSimulator(A, R, BG, Q_m, rv_id_sequence, iterations, early_termination, 0.75, batch, seed, local_rank, rank, world,
                        llr_dtype=llr_dtype, algorithm=algorithm)
    rows, esn0, bler, found = [], float(EsN0_start), 1.0, False
    fid = None
    if rank == 0 and out_dir:
        Path(out_dir).mkdir(parents=True, exist_ok=True)
        name = f"BLER_vs_SNR_{A}_{R:g}_{BG}_{_MOD_NAME[Q_m]}_{iterations}_{target_block_errors}_{EsN0_start:g}_{seed}.txt"  # :79
        fid = open(Path(out_dir) / name, "w")
    while bler > target_BLER:
        tot, found = sim.run_point(esn0, target_block_errors, max_blocks, found)
        bler = tot[1] / tot[0] if found else 1.0
        if found and bler < 1 and rank == 0:
            if fid:
                fid.write("%f\t%e\n" % (esn0, bler))                                                  # :165
                fid.flush()
            rows.append((esn0, bler, int(tot[0]), int(tot[1]), tot[3] / max(1, tot[0] * sim.C)))
            log(f"EsN0 {esn0:6.2f} dB  BLER {bler:.3e}  blocks {tot[0]}  errors {tot[1]}  mean iters {rows[-1][4]:.2f}")
        if max_blocks and found and tot[1] == 0:
            break
        esn0 += EsN0_delta                                                                           # :169
    if fid:
        fid.close()
    sim.close()
    return rows


def interp_required_snr(prev_BLER, BLER, prev_EsN0, EsN0, target_BLER):
    """plot_SNR_vs_A.m:175: linear interpolation of Es/N0 over log10(BLER) between the last two simulated points, i.e.
    interp1(log10([prev_BLER, BLER]), [prev_EsN0, EsN0], log10(target_BLER)).  NaN where interp1 returns NaN: no previous
    point (the first SNR already met the target), a point without errors, or a target outside the bracket."""
    if prev_BLER is None or not (prev_BLER > 0) or not (BLER > 0) or prev_BLER == BLER:
        return float("nan")
    x0, x1, xt = math.log10(prev_BLER), math.log10(BLER), math.log10(target_BLER)
    if not (min(x0, x1) <= xt <= max(x0, x1)):
        return float("nan")
    return prev_EsN0 + (EsN0 - prev_EsN0) * (xt - x0) / (x1 - x0)


def snr_vs_a(A_list, R_list, BG, iterations=50, target_block_errors=100, target_BLER=1e-2, EsN0_start=-2.0, EsN0_delta=0.1,
             seed=0, rv_id_sequence=(0,), batch=4096, max_blocks=None, out_dir="results", log=print, Q_m=2,
             llr_dtype=capi.F32, algorithm=capi.ALG_NMS):
    """plot_SNR_vs_A.m:69-193 on device: for every coding rate and information block length, step Es/N0 up from
    EsN0_start until the BLER falls to target_BLER (:102-162, the frame loop being BlerSimulator batches), then
    interpolate the Es/N0 at which it equals the target (:175).  Unsupported (A, R) combinations are skipped as in
    :164-172.  One results file per rate in the reference's format ("%d\t%f\n", :186).  Returns {R: [(A, EsN0)]}."""
    rank, local_rank, world = D.init()
    import torch
    torch.cuda.set_device(local_rank)
    out = {}
    for R in R_list:
        fid = None
        if rank == 0 and out_dir:
            Path(out_dir).mkdir(parents=True, exist_ok=True)
            name = f"SNR_vs_A_{target_BLER:g}_{R:g}_{BG}_{_MOD_NAME[Q_m]}_{iterations}_{target_block_errors}_{seed}.txt"   # :79
            fid = open(Path(out_dir) / name, "w")
        rows = []
        for A in A_list:
            try:
                sim = BlerSimulator(A, R, BG, Q_m, rv_id_sequence, iterations, True, 0.75, batch, seed, local_rank, rank, world,
                                    llr_dtype=llr_dtype, algorithm=algorithm)
            except capi.UnsupportedParameters as e:                                    # :164-168
                log(f"A={A} R={R:g}: the requested combination of parameters is not supported ({e}); skipped")
                continue
            BLER, prev_BLER, found = 1.0, None, False
            EsN0, prev_EsN0 = float(EsN0_start) - float(EsN0_delta), None
            while BLER > target_BLER:                                                  # :102
                prev_EsN0, EsN0 = EsN0, EsN0 + float(EsN0_delta)
                tot, found = sim.run_point(EsN0, target_block_errors, max_blocks, found)
                prev_BLER, BLER = BLER, (tot[1] / tot[0] if found else 1.0)
            sim.close()
            req = interp_required_snr(prev_BLER, BLER, prev_EsN0, EsN0, target_BLER)
            rows.append((int(A), req))
            log(f"R={R:g} A={A}: required Es/N0 {req:.3f} dB (BLER {prev_BLER:.3e} @ {prev_EsN0:.2f} dB, {BLER:.3e} @ {EsN0:.2f} dB)")
            if fid:
                fid.write("%d\t%f\n" % (A, req))                                       # :186
                fid.flush()
        if fid:
            fid.close()
        out[R] = rows
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description="BLER vs SNR on device (protocol of plot_BLER_vs_SNR.m)")
    ap.add_argument("--A", type=int, default=3842); ap.add_argument("--R", type=float, default=1 / 3)
    ap.add_argument("--BG", type=int, default=2); ap.add_argument("--iterations", type=int, default=8)
    ap.add_argument("--target-block-errors", type=int, default=100); ap.add_argument("--target-BLER", type=float, default=1e-3)
    ap.add_argument("--EsN0-start", type=float, default=0.0); ap.add_argument("--EsN0-delta", type=float, default=0.5)
    ap.add_argument("--seed", type=int, default=0); ap.add_argument("--rv", type=int, nargs="+", default=[0])
    ap.add_argument("--batch", type=int, default=4096); ap.add_argument("--max-blocks", type=int, default=None)
    ap.add_argument("--out-dir", default="results")
    ap.add_argument("--modulation", default="QPSK", choices=sorted(_MOD_NAME.values()))
    ap.add_argument("--llr-dtype", default="f32", choices=["f32", "f16x2"])
    ap.add_argument("--algorithm", default="nms", choices=["nms", "bp"],
                    help="nms: layered normalized min-sum (default); bp: the reference's flooding sum-product in float64")
    ap.add_argument("--snr-vs-A", type=int, nargs="+", default=None, metavar="A",
                    help="run the plot_SNR_vs_A.m protocol over these block lengths instead (uses --R, --target-BLER ...)")
    a = ap.parse_args(argv)
    Q_m = {v: k for k, v in _MOD_NAME.items()}[a.modulation]
    if a.snr_vs_A:
        snr_vs_a(a.snr_vs_A, [a.R], a.BG, a.iterations, a.target_block_errors, a.target_BLER, a.EsN0_start, a.EsN0_delta,
                 a.seed, a.rv, a.batch, a.max_blocks, a.out_dir, Q_m=Q_m,
                 llr_dtype=capi.F16X2 if a.llr_dtype == "f16x2" else capi.F32,
                 algorithm=capi.ALG_BP if a.algorithm == "bp" else capi.ALG_NMS)
        return
    sweep(a.A, a.R, a.BG, a.iterations, a.target_block_errors, a.target_BLER, a.EsN0_start, a.EsN0_delta, a.seed, a.rv,
          a.batch, a.max_blocks, True, a.out_dir, Q_m=Q_m, llr_dtype=capi.F16X2 if a.llr_dtype == "f16x2" else capi.F32,
          algorithm=capi.ALG_BP if a.algorithm == "bp" else capi.ALG_NMS)


if __name__ == "__main__":
    main()
