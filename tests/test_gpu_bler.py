"""On-device BLER loop (plot_BLER_vs_SNR.m protocol) -- GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_bler_sweep_plumbing_config(tmp_path):
    """BASELINE config 1 (BG2, A=20, R=1/5, QPSK, 8 iterations): the curve falls monotonically, the
    results file has the reference's "%f\\t%e" format, and the device loop agrees with the oracle on
    the LLRs it generated."""
    from ldpc_3gpp_matlab_b200 import bler
    rows = bler.sweep(20, 0.2, 2, iterations=8, target_block_errors=200, target_BLER=2e-2, EsN0_start=0.0,
                      EsN0_delta=1.0, seed=3, batch=8192, out_dir=str(tmp_path), log=lambda *_: None)
    assert len(rows) >= 2
    blers = [r[1] for r in rows]
    assert all(b1 > b2 for b1, b2 in zip(blers, blers[1:]))
    assert 0.2 < blers[0] < 0.5 and blers[-1] <= 2e-2          # SURVEY 8c scratch curve: .307 @ 0 dB, .0163 @ 2 dB
    files = list(tmp_path.glob("BLER_vs_SNR_20_0.2_2_QPSK_8_200_0_3.txt"))
    assert len(files) == 1
    for line, r in zip(files[0].read_text().splitlines(), rows):
        snr, b = line.split("\t")
        assert abs(float(snr) - r[0]) < 1e-6 and abs(float(b) - r[1]) < 1e-6 * max(1, r[1])


def test_bler_batch_matches_oracle_on_same_llrs(O):
    from ldpc_3gpp_matlab_b200.bler import BlerSimulator
    sim = BlerSimulator(400, 0.2, 2, iterations=8, early_termination=True, batch=256, seed=5)
    assert (sim.Z, sim.K, sim.Kp, sim.n_rows) == (52, 520, 416, 33)
    c, _ = sim.run_batch(-2.5)
    llr, info, hard = sim.llr.cpu().numpy(), sim.info.cpu().numpy(), sim.hard.cpu().numpy()
    assert np.isinf(llr[:, 416:520]).all() and (llr[:, :104] == 0).all()
    ref = O.decode_nms(2, 52, llr, 8, early_term=True, n_rows=33)
    assert (ref["hard"] == hard).all()
    assert c[0] == 256 and c[1] == int((hard[:, :416] != info[:, :416]).any(1).sum())
    sim.close()


def test_bler_segmented_and_harq():
    """C=2 transport block (script default A=3842, BG2) and an rv sequence with HARQ combining."""
    from ldpc_3gpp_matlab_b200.bler import BlerSimulator
    sim = BlerSimulator(3842, 1 / 3, 2, iterations=8, batch=64, seed=1)
    assert sim.C == 2 and sim.E_r == [5762, 5764]
    c, _ = sim.run_batch(3.0)
    assert c[0] == 64 and c[1] == 0
    sim.close()
    one = BlerSimulator(4000, 4000 / 4800, 1, rv_id_sequence=(0,), iterations=12, batch=128, seed=2)
    four = BlerSimulator(4000, 4000 / 4800, 1, rv_id_sequence=(0, 2, 3, 1), iterations=12, batch=128, seed=2)
    c1, _ = one.run_batch(-1.0)
    c4, _ = four.run_batch(-1.0)
    assert c1[1] == 128 and c4[1] == 0          # rate 5/6 fails at -1 dB; four redundancy versions combine to rate ~0.21
    one.close(); four.close()


def test_bler_higher_order_modulations_and_packed_half():
    """Every NRModulator setting through the device loop (fused modulate + AWGN + exact-LLR kernel, Q_m-aware
    interleaver): error-free well above threshold, all blocks lost well below it; the packed-half decoder tracks
    the float32 one on the same noise."""
    from ldpc_3gpp_matlab_b200 import capi
    from ldpc_3gpp_matlab_b200.bler import BlerSimulator
    for Q_m, hi, lo in ((1, 2.0, -8.0), (4, 9.0, -2.0), (6, 14.0, 2.0), (8, 19.0, 6.0)):
        sim = BlerSimulator(2000, 0.5, 1, Q_m=Q_m, iterations=10, batch=128, seed=Q_m)
        assert sim.p.G % Q_m == 0
        c_hi, _ = sim.run_batch(hi)
        c_lo, _ = sim.run_batch(lo)
        assert c_hi[1] == 0 and c_lo[1] == 128, (Q_m, c_hi, c_lo)
        sim.close()
    a = BlerSimulator(2000, 0.5, 1, Q_m=4, iterations=8, batch=512, seed=9)
    b = BlerSimulator(2000, 0.5, 1, Q_m=4, iterations=8, batch=512, seed=9, llr_dtype=capi.F16X2)
    found = False
    for esn0 in np.arange(5.6, 8.5, 0.3):            # walk both decoders through the waterfall on identical noise
        ca, _ = a.run_batch(float(esn0))
        cb, _ = b.run_batch(float(esn0))
        assert abs(int(ca[1]) - int(cb[1])) <= 16, (esn0, ca, cb)
        found |= 0 < ca[1] < 512
    assert found
    a.close(); b.close()


def test_device_crc_matches_oracle_and_bler_criterion(O):
    """nrldpc_crc (attach + check) against the oracle CRCs for the three 3GPP polynomials, and the CRC-based block-error
    criterion of the device loop against the direct bit comparison on the same noise (segmented TB, C = 2)."""
    import torch
    from ldpc_3gpp_matlab_b200 import capi
    from ldpc_3gpp_matlab_b200.bler import BlerSimulator
    rng = np.random.default_rng(8)
    h = capi.Handle(2, 2, 1)
    st = torch.cuda.current_stream().cuda_stream
    for name, kind, L in (("CRC16", capi.CRC16, 16), ("CRC24A", capi.CRC24A, 24), ("CRC24B", capi.CRC24B, 24)):
        for n, pad in ((1, 5), (7, 5), (20, 5), (333, 5), (8424, 5), (8424, 0), (4000, 8), (264, 16 - L % 16)):
            B, stride = 37, n + L + pad          # odd strides: rows at every alignment; pad 0 / 8: word-aligned rows (table path)
            bits = rng.integers(0, 2, (B, stride), dtype=np.uint8)
            d = torch.from_numpy(bits).cuda()
            ok = torch.zeros(B, dtype=torch.uint8, device="cuda")
            h.crc_raw(d, B, n, stride, kind, parity=d.data_ptr() + n, parity_stride=stride, stream=st)
            h.crc_raw(d, B, n + L, stride, kind, ok=ok, stream=st)
            got = d.cpu().numpy()
            for b in range(0, B, 9):
                assert (got[b, n:n + L] == O.crc(name, bits[b, :n])).all(), (name, n, b)
            assert bool(ok.all())
            d[:, n // 2] ^= 1
            h.crc_raw(d, B, n + L, stride, kind, ok=ok, stream=st)
            assert not bool(ok.any())
    with pytest.raises(capi.UnsupportedParameters):
        h.crc_raw(d, 1, 8, 8, 7, ok=ok, stream=st)
    h.close()
    a = BlerSimulator(3842, 1 / 3, 2, iterations=8, batch=256, seed=4, crc=True)
    b = BlerSimulator(3842, 1 / 3, 2, iterations=8, batch=256, seed=4, crc=False)
    assert a.C == 2 and a.L_cb == 24 and a.Bsz == 3866
    for esn0 in (-1.6, -1.3, -1.0):
        ca, _ = a.run_batch(esn0)
        cb, _ = b.run_batch(esn0)
        assert ca[0] == cb[0] == 256 and ca[1] == cb[1], (esn0, ca, cb)     # symmetric decoder: same noise, same failures
    assert 0 < ca[1] + cb[1]
    a.close(); b.close()


def test_bler_reference_algorithm_mode_matches_oracle_b(O):
    """algorithm = NRLDPC_ALG_BP through the device loop: on the LLRs the loop generated, block decisions and
    iteration counts equal the CPU restatement of the reference's decoder (flooding sum-product, float64, whole H),
    so the BLER curve of this mode IS the reference algorithm's curve (0 dB apart), for BASELINE configs 1 and 3."""
    from ldpc_3gpp_matlab_b200 import capi
    from ldpc_3gpp_matlab_b200.bler import BlerSimulator
    for (A, R, BG, esn0s, Z, rows) in ((20, 0.2, 2, (0.0, 1.5, 3.0), 6, 42), (400, 0.2, 2, (-3.0, -2.4), 52, 42)):
        sim = BlerSimulator(A, R, BG, iterations=8, early_termination=True, batch=512, seed=11, algorithm=capi.ALG_BP)
        assert (sim.Z, sim.n_rows) == (Z, rows)
        seen_err = seen_ok = False
        for esn0 in esn0s:
            c, _ = sim.run_batch(esn0)
            llr, hard, iters = sim.llr.cpu().numpy(), sim.hard.cpu().numpy(), sim.iters.cpu().numpy()
            ref = O.decode_bp(BG, Z, llr, 8)
            assert (ref["hard"] == hard).all() and (ref["iters"] == iters).all(), (A, esn0)
            seen_err |= c[1] > 0
            seen_ok |= c[1] < c[0]
        assert seen_err and seen_ok
        sim.close()


def test_host_mirror_sum_product_option(O):
    """NRLDPCDecoder(algorithm='Sum-product'): same step()/reset() API, the reference's algorithm underneath."""
    from ldpc_3gpp_matlab_b200 import capi
    from ldpc_3gpp_matlab_b200.nrldpc import NRLDPCDecoder, NRLDPCEncoder
    rng = np.random.default_rng(21)
    enc = NRLDPCEncoder(BG=2, A=400, G=2000)
    dec = NRLDPCDecoder(BG=2, A=400, G=2000, iterations=8, algorithm="Sum-product")
    nms = NRLDPCDecoder(BG=2, A=400, G=2000, iterations=8)
    n_ok = 0
    for _ in range(6):
        a = rng.integers(0, 2, 400).astype(np.float64)
        g = enc.step(a)
        s2 = 10 ** (2.0 / 10)
        y = (1 - 2 * g) + rng.normal(0, np.sqrt(s2 / 2), g.shape)     # -2 dB: both decoders succeed most of the time
        g_tilde = 4 * y / s2
        a_hat = dec.step(g_tilde)
        assert a_hat.size in (0, 400)
        if a_hat.size:
            assert (a_hat == a).all()
            n_ok += 1
        nms.step(g_tilde)
    assert n_ok >= 3 and dec._active_rows(2000) == 42 and nms._active_rows(2000) == 33
    with pytest.raises(capi.UnsupportedParameters):
        NRLDPCDecoder(BG=2, A=400, G=2000, algorithm="bogus").step(g_tilde)
    enc.release(); dec.release(); nms.release()


def test_snr_vs_a_driver(tmp_path):
    """plot_SNR_vs_A.m protocol on device: required Es/N0 per block length, the reference's "%d\\t%f" file, an unsupported
    length skipped (the callers catch 'ldpc_3gpp_matlab:UnsupportedParameters', plot_SNR_vs_A.m:164-168)."""
    from ldpc_3gpp_matlab_b200 import bler
    out = bler.snr_vs_a([20, 200, 10 ** 7], [0.2], 2, iterations=8, target_block_errors=200, target_BLER=5e-2, EsN0_start=-4.0,
                        EsN0_delta=0.5, seed=1, batch=4096, out_dir=str(tmp_path), log=lambda *_: None)
    rows = out[0.2]
    assert [a for a, _ in rows] == [20, 200]                  # A = 10^7 at R = 1/5 is not supported and is skipped
    req = dict(rows)
    assert all(np.isfinite(v) for v in req.values())
    assert -4.0 <= req[200] < req[20] <= 4.0                   # longer blocks need less Es/N0
    f = tmp_path / "SNR_vs_A_0.05_0.2_2_QPSK_8_200_1.txt"
    lines = f.read_text().splitlines()
    assert len(lines) == 2
    for line, (a, v) in zip(lines, rows):
        x, y = line.split("\t")
        assert int(x) == a and abs(float(y) - v) < 1e-6


def test_fused_channel_rate_recover_equals_two_stages():
    """nrldpc_qpsk_awgn_rate_recover is bit-identical to nrldpc_qpsk_awgn_llr + nrldpc_rate_recover with the same
    (seed, stream id): every rv_id, limited-buffer N_cb, wrap-around repetition (E > N_cb), filler, HARQ accumulation."""
    import torch
    from ldpc_3gpp_matlab_b200 import capi
    st = torch.cuda.current_stream().cuda_stream
    for bg, Z, Kp, E, k0, Ncb in ((1, 384, 8448, 25272, 0, 25344), (1, 384, 8000, 24000, 17 * 384, 25344), (2, 52, 416, 2000, 13 * 52, 2600),
                                  (2, 52, 416, 6000, 25 * 52, 2000), (1, 96, 2000, 1200, 0, 6336)):
        h = capi.Handle(bg, Z, 4, False)
        B = 64
        rm = capi.Rm(E, k0, Ncb, Kp, 2)
        g = torch.Generator(device="cuda").manual_seed(Z + E)
        f = torch.randint(0, 2, (B, E), dtype=torch.uint8, device="cuda", generator=g)
        fl = torch.empty((B, E), dtype=torch.float32, device="cuda")
        harq_a = torch.randn((B, h.N), dtype=torch.float32, device="cuda", generator=g)
        harq_b = harq_a.clone()
        out_a = torch.empty((B, h.n_cw), dtype=torch.float32, device="cuda")
        out_b = torch.empty_like(out_a)
        for rep, hq in enumerate(((None, None), (harq_a, harq_b), (harq_a, harq_b))):
            h.qpsk_awgn_llr_raw(f, B, E, 0.7, 99, 5 + rep, fl, stream=st)
            h.rate_recover_raw(fl, B, rm, hq[0], out_a, mem=capi.MEM_DEVICE, stream=st)
            h.qpsk_awgn_rate_recover_raw(f, B, rm, 0.7, 99, 5 + rep, hq[1], out_b, stream=st)
            torch.cuda.synchronize()
            assert torch.equal(out_a.view(torch.int32), out_b.view(torch.int32)), (bg, Z, E, rep)
            assert torch.equal(harq_a.view(torch.int32), harq_b.view(torch.int32))
        h.close()


def test_bler_counters_fused_path_equals_stagewise_path(monkeypatch):
    """The batch loop with the fused channel + rate-recovery kernel and the device-side block-error bookkeeping returns the
    counters of the stage-by-stage path (NRLDPC_BLER_UNFUSED) on the same seeds, and those equal a host recount from the
    buffers: single block, segmented transport block (C = 2) with HARQ, and the no-CRC criterion."""
    from ldpc_3gpp_matlab_b200.bler import BlerSimulator
    for kw, esn0 in ((dict(A=8424, R=1 / 3, BG=1, batch=256), -0.35), (dict(A=3842, R=0.5, BG=2, batch=128, rv_id_sequence=(0, 2)), -3.0),
                     (dict(A=400, R=0.2, BG=2, batch=512, crc=False), -2.2), (dict(A=1000, R=0.8, BG=1, batch=128, rv_id_sequence=(0, 1, 2)), -1.5)):
        res = []
        for unfused in ("", "1"):
            if unfused:
                monkeypatch.setenv("NRLDPC_BLER_UNFUSED", "1")
            else:
                monkeypatch.delenv("NRLDPC_BLER_UNFUSED", raising=False)
            sim = BlerSimulator(iterations=8, seed=11, **kw)
            tot = np.zeros(4, dtype=np.int64)
            for _ in range(2):
                c, any_ok = sim.run_batch(esn0)
                tot += c
            hard, info = sim.hard.cpu().numpy(), sim.info.cpu().numpy()
            res.append((tot.copy(), hard, info))
            if len(sim.rvs) == 1:                      # last batch, host recount of the bit errors of failed blocks
                wrong = (hard[:, :sim.Kp] != info[:, :sim.Kp]).reshape(sim.B, sim.C, -1).sum(axis=(1, 2))
                latch = sim.latch.cpu().numpy()
                assert int(c[2]) == int((wrong * (latch == 0)).sum()) and int(c[1]) == int((latch == 0).sum())
                assert int(c[3]) == int(sim.iters.sum())
            sim.close()
        assert (res[0][0] == res[1][0]).all(), (kw, res[0][0], res[1][0])
        assert (res[0][1] == res[1][1]).all()
        assert res[0][0][0] == 2 * kw["batch"]
        if len(kw.get("rv_id_sequence", (0,))) == 1:
            assert 0 < res[0][0][1] < res[0][0][0]      # the point sits in the waterfall: both outcomes occur


def test_random_bits_kernel():
    """nrldpc_random_bits (the information blocks of the Monte-Carlo loop, plot_BLER_vs_SNR.m:112): only [0, n_bits) of each row
    is written, the bits depend on (seed, stream id, row, position) and not on alignment or launch geometry, different keys give
    different blocks, and the bits are uniform and independent enough for a Monte-Carlo source (mean, row/column balance,
    lag correlations within a few standard deviations)."""
    import torch
    from ldpc_3gpp_matlab_b200 import capi
    h = capi.Handle(2, 2, 1)
    st = torch.cuda.current_stream().cuda_stream
    ref = None
    for off, stride in ((0, 8448), (3, 8453), (16, 8432)):
        buf = torch.full((64 * stride + 64,), 7, dtype=torch.uint8, device="cuda")
        view = buf[off:off + 64 * stride]
        h.random_bits_raw(view, 64, 8424, stride, 1234, 5, stream=st)
        got = view.cpu().numpy().reshape(64, stride)
        assert set(np.unique(got[:, :8424])) == {0, 1}
        assert (got[:, 8424:] == 7).all() and (buf[:off].cpu().numpy() == 7).all() and (buf[off + 64 * stride:].cpu().numpy() == 7).all()
        if ref is None:
            ref = got[:, :8424].copy()
        assert (got[:, :8424] == ref).all(), (off, stride)
    small = torch.zeros((5, 40), dtype=torch.uint8, device="cuda")
    h.random_bits_raw(small, 5, 33, 40, 1234, 5, stream=st)
    assert (small.cpu().numpy()[:, :33] == ref[:5, :33]).all() and not small.cpu().numpy()[:, 33:].any()    # prefix property of a row
    other = torch.zeros((64, 8448), dtype=torch.uint8, device="cuda")
    h.random_bits_raw(other, 64, 8424, 8448, 1234, 6, stream=st)
    x = other.cpu().numpy()[:, :8424]
    assert 0.45 < (x != ref).mean() < 0.55
    big = torch.zeros((4096, 8448), dtype=torch.uint8, device="cuda")
    h.random_bits_raw(big, 4096, 8424, 8448, 99, 1, stream=st)
    b = big.cpu().numpy()[:, :8424].astype(np.float64)
    n = b.size
    assert abs(b.mean() - 0.5) < 5 * 0.5 / np.sqrt(n)
    assert np.abs(b.mean(axis=0) - 0.5).max() < 6 * 0.5 / np.sqrt(4096) and np.abs(b.mean(axis=1) - 0.5).max() < 6 * 0.5 / np.sqrt(8424)
    c = 2 * b - 1
    for lag in (1, 2, 16, 32, 128):
        assert abs((c[:, :-lag] * c[:, lag:]).mean()) < 5 / np.sqrt(c[:, lag:].size), lag
    assert abs((c[:-1] * c[1:]).mean()) < 5 / np.sqrt(c[1:].size)
    h.close()
