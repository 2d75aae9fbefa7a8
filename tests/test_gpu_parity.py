"""Parity tests proper: the sm_100a kernels, called through the C ABI, against the CPU oracle on the
same seeded inputs (bit-exact: hard bits, iteration counts, parity flags AND float32 APP bit patterns),
against the committed golden vectors, and through size-independent properties at BASELINE sizes."""
import math

import numpy as np
import pytest

from conftest import ALL_Z, make_core_pass_llr, make_llr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from ldpc_3gpp_matlab_b200 import capi
    capi.load()
    return capi


def _same_bits(a, b):
    return bool((a.view(np.uint32) == b.view(np.uint32)).all())


def test_decode_golden_vectors_on_gpu(capi, golden_decode):
    names = sorted({k.split("__")[0] for k in golden_decode.files})
    for n in names:
        bg, Z, iters, et, rows = golden_decode[n + "__cfg"].tolist()
        h = capi.Handle(bg, Z, iters, bool(et))
        out = h.decode(golden_decode[n + "__llr"], n_rows=rows, want_soft=True)
        assert (np.packbits(out["hard"], axis=1) == golden_decode[n + "__hard"]).all(), n
        assert (out["iters"] == golden_decode[n + "__iters"]).all(), n
        assert (out["parity_ok"] == golden_decode[n + "__ok"]).all(), n
        u = out["app"].view(np.uint32)
        assert (np.bitwise_xor.reduce(u, axis=1) == golden_decode[n + "__app_xor"]).all(), n
        assert (u.astype(np.uint64).sum(axis=1) == golden_decode[n + "__app_sum"]).all(), n
        h.close()


@pytest.mark.parametrize("bg", [1, 2])
def test_decode_bit_exact_all_51_lifting_sizes(capi, O, bg):
    """Every (BG, Z): ragged batch (not a multiple of codewords-per-CTA), fixed iterations and early stop."""
    rng = np.random.default_rng(100 + bg)
    for Z in ALL_Z:
        d = O.dims(bg, Z)
        B = max(3, min(40, 1500 // Z)) + 1
        E = int(d["N"] * rng.uniform(0.35, 1.0)) // 2 * 2
        info, llr = make_llr(O, bg, Z, B, E, rng.uniform(-1.0, 3.0), rng)
        for et in (False, True):
            ref = O.decode_nms(bg, Z, llr, 6, early_term=et)
            h = capi.Handle(bg, Z, 6, et)
            out = h.decode(llr, want_soft=True)
            h.close()
            assert (out["hard"] == ref["hard"]).all(), (bg, Z, et)
            assert _same_bits(out["app"], ref["app"]), (bg, Z, et)
            assert (out["iters"] == ref["iters"]).all() and (out["parity_ok"] == ref["parity_ok"]).all(), (bg, Z, et)


@pytest.mark.parametrize("bg", [1, 2])
def test_decode_f16x2_bit_exact_all_51_lifting_sizes(capi, O, bg):
    """Packed-half kernel (two codewords per thread) against oracle A16: every (BG, Z), odd batches (a pair with a
    missing second codeword), fixed iterations and early stop (codewords of one pair converge at different times)."""
    rng = np.random.default_rng(300 + bg)
    for Z in ALL_Z:
        d = O.dims(bg, Z)
        B = max(3, min(40, 1500 // Z)) | 1
        E = int(d["N"] * rng.uniform(0.35, 1.0)) // 2 * 2
        info, llr = make_llr(O, bg, Z, B, E, rng.uniform(-1.0, 3.0), rng)
        for et in (False, True):
            ref = O.decode_nms(bg, Z, llr, 6, early_term=et, f16=True)
            h = capi.Handle(bg, Z, 6, et, llr_dtype=capi.F16X2)
            out = h.decode(llr, want_soft=True)
            h.close()
            assert (out["hard"] == ref["hard"]).all(), (bg, Z, et)
            assert _same_bits(out["app"], ref["app"]), (bg, Z, et)
            assert (out["iters"] == ref["iters"]).all() and (out["parity_ok"] == ref["parity_ok"]).all(), (bg, Z, et)


@pytest.mark.parametrize("bg,Z,rows", [(1, 384, 46), (1, 384, 5), (2, 52, 33), (2, 6, 13), (1, 30, 4), (2, 208, 20)])
def test_decode_f16x2_special_values_and_row_trimming(capi, O, bg, Z, rows):
    """+inf / NaN filler, zeros, -0.0, saturating and sub-resolution magnitudes in the packed-half kernel."""
    rng = np.random.default_rng(Z + rows + 1)
    d = O.dims(bg, Z)
    B = 6
    llr = (rng.normal(0, 4, (B, d["ncw"]))).astype(np.float32)
    llr[:, :2 * Z] = 0
    llr[:, (d["kcols"] + rows) * Z:] = 0
    llr[0, 3 * Z:3 * Z + 17] = np.inf
    llr[1, 3 * Z:3 * Z + 17] = np.nan
    llr[2, 5 * Z:5 * Z + 9] = -np.inf
    llr[3, ::7] = 0.0
    llr[3, 1::11] = -0.0
    llr[4, 2 * Z::5] *= 1e30
    llr[4, 2 * Z + 1::9] *= 1e-6
    llr[5] *= 500.0
    for et in (False, True):
        for alpha in (0.75, 0.8):
            ref = O.decode_nms(bg, Z, llr, 7, early_term=et, n_rows=rows, alpha=alpha, f16=True)
            h = capi.Handle(bg, Z, 7, et, alpha=alpha, llr_dtype=capi.F16X2)
            out = h.decode(llr, n_rows=rows, want_soft=True)
            h.close()
            assert np.isfinite(out["app"]).all()
            assert (out["hard"] == ref["hard"]).all() and _same_bits(out["app"], ref["app"])
            assert (out["iters"] == ref["iters"]).all() and (out["parity_ok"] == ref["parity_ok"]).all()


def test_decode_f16x2_headline_config_matches_f32_decisions(capi, O):
    """BG1 Z=384 rate 1/3 at the benchmark operating point: the packed-half kernel equals oracle A16 bit for bit and
    its block decisions track the float32 kernel (stated tolerance: at most 2 of 64 blocks differ)."""
    rng = np.random.default_rng(77)
    info, llr = make_llr(O, 1, 384, 64, 25272, -0.3, rng)
    ref = O.decode_nms(1, 384, llr, 8, f16=True)
    h16 = capi.Handle(1, 384, 8, False, llr_dtype=capi.F16X2)
    out16 = h16.decode(llr, want_soft=True)
    h16.close()
    assert (out16["hard"] == ref["hard"]).all() and _same_bits(out16["app"], ref["app"])
    h32 = capi.Handle(1, 384, 8, False)
    out32 = h32.decode(llr)
    h32.close()
    e16 = (out16["hard"] != info).any(axis=1)
    e32 = (out32["hard"] != info).any(axis=1)
    assert int((e16 != e32).sum()) <= 2


@pytest.mark.parametrize("bg,Z,rows", [(1, 384, 46), (1, 384, 5), (1, 384, 13), (2, 52, 33), (2, 6, 13), (2, 384, 42),
                                       (1, 30, 4), (2, 208, 20)])
def test_decode_special_values_and_row_trimming(capi, O, bg, Z, rows):
    """+inf / NaN filler, exact zeros, huge and denormal magnitudes, -0.0; active-row trimming."""
    rng = np.random.default_rng(Z + rows)
    d = O.dims(bg, Z)
    B = 5
    llr = (rng.normal(0, 4, (B, d["ncw"]))).astype(np.float32)
    llr[:, :2 * Z] = 0
    llr[:, (d["kcols"] + rows) * Z:] = 0
    llr[0, 3 * Z:3 * Z + 17] = np.inf
    llr[1, 3 * Z:3 * Z + 17] = np.nan
    llr[2, 5 * Z:5 * Z + 9] = -np.inf
    llr[3, ::7] = 0.0
    llr[3, 1::11] = -0.0
    llr[4, 2 * Z::5] *= 1e30
    llr[4, 2 * Z + 1::9] *= 1e-42
    for et in (False, True):
        ref = O.decode_nms(bg, Z, llr, 7, early_term=et, n_rows=rows)
        h = capi.Handle(bg, Z, 7, et)
        out = h.decode(llr, n_rows=rows, want_soft=True)
        h.close()
        assert np.isfinite(out["app"]).all()
        assert (out["hard"] == ref["hard"]).all() and _same_bits(out["app"], ref["app"])
        assert (out["iters"] == ref["iters"]).all() and (out["parity_ok"] == ref["parity_ok"]).all()


@pytest.mark.parametrize("dtype", ["f32", "f16x2"])
@pytest.mark.parametrize("bg,Z,n_rows", [(1, 384, 46), (1, 384, 13), (1, 384, 5), (1, 224, 24), (2, 256, 42), (2, 352, 8),
                                         (1, 52, 46), (2, 52, 33), (1, 7, 46), (2, 12, 20), (1, 96, 4)])
def test_early_stop_core_checks_hold_extension_check_fails(capi, O, bg, Z, n_rows, dtype):
    """Two-stage 'Parity check satisfied' stop (NRLDPCDecoder.m:120): codewords whose core checks all hold while one
    extension check fails must run to the iteration limit with parity_ok = 0; clean ones stop after one iteration; both
    kinds share CTAs (small Z) and packed-half pairs.  Bit-exact against the oracle in every output."""
    rng = np.random.default_rng(7 * Z + n_rows + bg)
    f16 = dtype == "f16x2"
    B = 2 * max(1, 384 // Z) + 5
    info, llr, kind = make_core_pass_llr(O, bg, Z, B, n_rows, rng, wrong=2048.0 if f16 else 1048576.0)
    ref = O.decode_nms(bg, Z, llr, 3, early_term=True, n_rows=n_rows, f16=f16)
    h = capi.Handle(bg, Z, 3, True, llr_dtype=capi.F16X2 if f16 else capi.F32)
    out = h.decode(llr, n_rows=n_rows, want_soft=True)
    h.close()
    assert (out["hard"] == ref["hard"]).all() and _same_bits(out["app"], ref["app"])
    assert (out["iters"] == ref["iters"]).all() and (out["parity_ok"] == ref["parity_ok"]).all()
    assert (out["parity_ok"][kind == 0] == 1).all() and (out["iters"][kind == 0] == 1).all()
    if n_rows > 4:
        assert (out["parity_ok"][kind == 1] == 0).all() and (out["iters"][kind == 1] == 3).all()
    assert (out["hard"][kind != 1] == info[kind != 1]).all()


def test_decode_alpha_and_iteration_sweep(capi, O):
    rng = np.random.default_rng(9)
    info, llr = make_llr(O, 2, 104, 6, 4000, -1.5, rng)
    for alpha, iters in ((0.75, 1), (0.8125, 3), (1.0, 5), (0.7, 12)):
        ref = O.decode_nms(2, 104, llr, iters, alpha=alpha)
        h = capi.Handle(2, 104, iters, False, alpha=alpha)
        out = h.decode(llr, want_soft=True)
        h.close()
        assert (out["hard"] == ref["hard"]).all() and _same_bits(out["app"], ref["app"]), (alpha, iters)


def test_decode_empty_batch_and_shape_errors(capi):
    h = capi.Handle(1, 8, 4)
    out = h.decode(np.zeros((0, h.n_cw), np.float32))
    assert out["hard"].shape == (0, h.K)
    with pytest.raises(capi.NRLDPCError):
        h.decode(np.zeros((2, h.n_cw - 1), np.float32))
    with pytest.raises(capi.UnsupportedParameters):
        h.decode(np.zeros((2, h.n_cw), np.float32), n_rows=3)
    with pytest.raises(capi.UnsupportedParameters):
        h.decode(np.zeros((2, h.n_cw), np.float32), n_rows=47)
    h.close()


def test_decode_device_pointers_match_host_path(capi, O):
    import torch
    rng = np.random.default_rng(4)
    info, llr = make_llr(O, 1, 96, 37, 5000, 0.0, rng)
    h = capi.Handle(1, 96, 8, True)
    host = h.decode(llr, want_soft=True)
    t = torch.from_numpy(llr).cuda()
    hard = torch.zeros((37, h.K), dtype=torch.uint8, device="cuda")
    soft = torch.zeros((37, h.n_cw), dtype=torch.float32, device="cuda")
    iters = torch.zeros(37, dtype=torch.int32, device="cuda")
    ok = torch.zeros(37, dtype=torch.uint8, device="cuda")
    h.decode_raw(t, 37, hard, soft, iters, ok, mem=capi.MEM_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (hard.cpu().numpy() == host["hard"]).all() and _same_bits(soft.cpu().numpy(), host["app"])
    assert (iters.cpu().numpy() == host["iters"]).all() and (ok.cpu().numpy() == host["parity_ok"]).all()
    h.close()


@pytest.mark.parametrize("bg", [1, 2])
def test_encode_all_102_against_oracle(capi, O, bg):
    rng = np.random.default_rng(200 + bg)
    for Z in ALL_Z:
        d = O.dims(bg, Z)
        info = rng.integers(0, 2, (5, d["K"]), dtype=np.uint8)
        h = capi.Handle(bg, Z, 1)
        cw = h.encode(info)
        h.close()
        assert (cw == O.encode(bg, Z, info)).all(), (bg, Z)
        assert all(O.syndrome_weight(bg, Z, c) == 0 for c in cw)


def test_encode_linearity_and_zero(capi):
    """Size-independent properties at the headline size: enc(0)=0, enc(a^b)=enc(a)^enc(b)."""
    rng = np.random.default_rng(11)
    h = capi.Handle(1, 384, 1)
    a = rng.integers(0, 2, (64, h.K), dtype=np.uint8)
    b = rng.integers(0, 2, (64, h.K), dtype=np.uint8)
    assert not h.encode(np.zeros((2, h.K), np.uint8)).any()
    assert (h.encode(a ^ b) == (h.encode(a) ^ h.encode(b))).all()
    h.close()


@pytest.mark.parametrize("A,BG,R,Qm,rv,lbrm", [(20, 2, 0.2, 2, 0, 0), (400, 2, 0.2, 4, 1, 0), (1000, 1, 1 / 3, 6, 2, 0),
                                               (3842, 2, 1 / 3, 2, 3, 0), (8000, 1, 0.5, 8, 0, 0), (500, 1, 0.12, 1, 3, 0),
                                               (8424, 1, 1 / 3, 2, 2, 1), (8424, 1, 8 / 9, 2, 0, 0), (60, 2, 0.05, 2, 1, 0)])
def test_rate_match_and_recover_against_reference_loops(capi, O, A, BG, R, Qm, rv, lbrm):
    """GPU closed-form index maps vs the oracle's literal restatement of the reference's while-loops
    (NRLDPCEncoder.m:168-225, NRLDPCDecoder.m:172-242,262-264), including wrap-around soft combining,
    filler skipping, all rv_id, LBRM and a HARQ retransmission."""
    G = int(math.floor(A / R / Qm + 0.5)) * Qm
    p = O.params(BG, A, G, Q_m=Qm, rv_id=rv, I_LBRM=lbrm, TBS_LBRM=A)
    assert p is not None
    Z, K, Kp, N, E = p.Z_c, p.K, p.K_prime, p.N, p.E_r[0]
    rng = np.random.default_rng(A + rv)
    B = 3
    info = rng.integers(0, 2, (B, K), dtype=np.uint8)
    info[:, Kp:] = 0
    h = capi.Handle(BG, Z, 1)
    cw = h.encode(info)
    f = h.rate_match(cw, E, p.k_0, p.N_cb, Kp, Qm)
    harq_gpu = np.zeros((B, N), np.float32)
    harq_ref = [np.zeros(p.N_cb, np.float32) for _ in range(B)]
    for tx in range(2):
        llr_f = ((1 - 2.0 * f) * rng.uniform(0.5, 4.0, f.shape)).astype(np.float32)
        got = h.rate_recover(llr_f, E, p.k_0, p.N_cb, Kp, Qm, harq=harq_gpu)
        for b in range(B):
            d = O.cw_to_d(Z, K, Kp, N, cw[b])
            if tx == 0:
                assert (f[b] == O.interleave_tx(O.bit_selection_tx(d, p.N_cb, p.k_0, E), Qm)).all()
            d_t = O.bit_selection_rx(O.deinterleave_rx(llr_f[b], Qm), N, p.N_cb, p.k_0, Z, K, Kp, harq_buf=harq_ref[b])
            want = O.d_to_cw_llr(d_t, Z)
            assert (got[b].view(np.uint32) == want.view(np.uint32)).all(), (tx, b)
    h.close()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_rate_recover_random_geometries(capi, O, seed, monkeypatch):
    """Rate recovery over random rate-matching geometries (any E, k_0, N_cb, K', Q_m the C ABI accepts, including several laps of
    the circular buffer = wrapped repetitions, limited buffers, filler, rows that do and do not qualify for the TMA kernel), two
    transmissions with HARQ accumulation: the table-driven gather (default), the per-block closed form (NRLDPC_RR_TABLE=0) and the
    oracle's literal restatement of the reference's while-loop (NRLDPCDecoder.m:172-242,262-264) agree bit for bit."""
    rng = np.random.default_rng(9000 + seed)
    for case in range(14):
        bg = int(rng.integers(1, 3))
        Z = int(rng.choice([2, 6, 13, 22, 36, 52, 80, 120]))
        d = O.dims(bg, Z)
        K, N = d["K"], d["N"]
        Kp = int(rng.integers(max(1, K - 3 * Z), K + 1))
        N_cb = N if rng.random() < 0.5 else int(rng.integers(max(K - 2 * Z + Z, N // 3), N + 1))
        Qm = int(rng.choice([1, 2, 4, 6, 8]))
        E = int(rng.integers(1, max(2, int(2.7 * N_cb)) // Qm + 1)) * Qm
        k_0 = int(rng.integers(0, N_cb))
        B = 3
        # TX side of the same geometry: bit selection + interleaving against the literal loops (NRLDPCEncoder.m:168-225)
        info = rng.integers(0, 2, (B, K), dtype=np.uint8)
        info[:, Kp:] = 0
        h = capi.Handle(bg, Z, 1)
        cw = h.encode(info)
        f_tx = h.rate_match(cw, E, k_0, N_cb, Kp, Qm)
        h.close()
        for b in range(B):
            d_tx = O.cw_to_d(Z, K, Kp, N, cw[b])
            assert (f_tx[b] == O.interleave_tx(O.bit_selection_tx(d_tx, N_cb, k_0, E), Qm)).all(), ("tx", bg, Z, E, k_0, N_cb, Kp, Qm, b)
        outs = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("NRLDPC_RR_TABLE", mode)
            h = capi.Handle(bg, Z, 1)
            harq = np.zeros((B, N), np.float32)
            r2 = np.random.default_rng(case)
            got = []
            for tx in range(2):
                llr_f = r2.normal(0, 3, (B, E)).astype(np.float32)
                got.append((llr_f, h.rate_recover(llr_f, E, k_0, N_cb, Kp, Qm, harq=harq).copy()))
            outs[mode] = got
            h.close()
        harq_ref = [np.zeros(N_cb, np.float32) for _ in range(B)]
        for tx in range(2):
            llr_f, g1 = outs["1"][tx]
            g0 = outs["0"][tx][1]
            assert (g1.view(np.uint32) == g0.view(np.uint32)).all(), (bg, Z, E, k_0, N_cb, Kp, Qm, tx)
            for b in range(B):
                d_t = O.bit_selection_rx(O.deinterleave_rx(llr_f[b], Qm), N, N_cb, k_0, Z, K, Kp, harq_buf=harq_ref[b])
                want = O.d_to_cw_llr(d_t, Z)
                assert (g1[b].view(np.uint32) == want.view(np.uint32)).all(), (bg, Z, E, k_0, N_cb, Kp, Qm, tx, b)


def test_full_chain_system_objects(capi, O):
    """encode -> QPSK -> AWGN -> exact LLR -> decode through the NRLDPCEncoder/NRLDPCDecoder mirrors
    (plot_BLER_vs_SNR.m:118-146 for one frame per configuration), incl. C=2 segmentation and a_hat=[]."""
    from ldpc_3gpp_matlab_b200.nrldpc import NRLDPCDecoder, NRLDPCEncoder, matlab_round
    rng = np.random.default_rng(21)
    for A, BG, R, esn0 in ((20, 2, 0.2, 6.0), (3842, 2, 1 / 3, 3.0), (8424, 1, 1 / 3, 1.5), (8424, 1, 8 / 9, 9.0),
                           (12000, 1, 0.5, 4.0)):
        G = matlab_round(A / R / 2) * 2
        enc = NRLDPCEncoder(A=A, BG=BG, G=G, Q_m=2)
        dec = NRLDPCDecoder(A=A, BG=BG, G=G, Q_m=2, I_HARQ=1, iterations=8)
        a = rng.integers(0, 2, A).astype(np.float64)
        g = enc.step(a)
        assert g.shape == (G,) and set(np.unique(g)) <= {0.0, 1.0}
        var = 10 ** (-esn0 / 10)
        re, im = O.qpsk_mod(g.astype(np.uint8))
        re = re + rng.normal(0, math.sqrt(var / 2), re.shape).astype(np.float32)
        im = im + rng.normal(0, math.sqrt(var / 2), im.shape).astype(np.float32)
        dec.reset()
        a_hat = dec.step(O.qpsk_demod(re, im, var))
        assert a_hat.shape == (A,) and (a_hat == a).all(), (A, BG)
        # garbage in -> CRC failure -> empty output (NRLDPCDecoder.m:337-339)
        dec.reset()
        assert dec.step(rng.normal(0, 1, G)).size == 0
        with pytest.raises(capi.NRLDPCError):
            dec.step(np.zeros(G + 1))
        enc.release(); dec.release()


def test_harq_retransmissions_combine(capi, O):
    """rv_id sequence [0,2,3,1] at an SNR where one transmission fails but the combination succeeds."""
    from ldpc_3gpp_matlab_b200.nrldpc import NRLDPCDecoder, NRLDPCEncoder
    rng = np.random.default_rng(33)
    A, BG, G = 4000, 1, 4800
    enc = NRLDPCEncoder(A=A, BG=BG, G=G, Q_m=2)
    dec = NRLDPCDecoder(A=A, BG=BG, G=G, Q_m=2, I_HARQ=1, iterations=12)
    a = rng.integers(0, 2, A).astype(np.float64)
    var = 10 ** (-(-1.0) / 10)
    dec.reset()
    results = []
    for rv in (0, 2, 3, 1):
        enc.rv_id = rv; dec.rv_id = rv
        g = enc.step(a)
        re, im = O.qpsk_mod(g.astype(np.uint8))
        re = re + rng.normal(0, math.sqrt(var / 2), re.shape).astype(np.float32)
        im = im + rng.normal(0, math.sqrt(var / 2), im.shape).astype(np.float32)
        a_hat = dec.step(O.qpsk_demod(re, im, var))
        results.append(a_hat.size > 0 and bool((a_hat == a).all()))
    assert results[0] is False and results[-1] is True


def test_headline_size_round_trip_properties(capi):
    """BASELINE config 2 at full size (BG1 Z=384 K=8448 rate 1/3, batch 4096, 8 iterations): on-device
    encode -> rate match -> QPSK/AWGN/LLR -> rate recover -> decode; every block must come back at
    2 dB, and the all-zero LLR input must decode to all-zero bits."""
    import torch
    B, E = 4096, 25272
    h = capi.Handle(1, 384, 8, False)
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(5)
    info = torch.randint(0, 2, (B, h.K), dtype=torch.uint8, device="cuda", generator=g)
    cw = torch.empty((B, h.n_cw), dtype=torch.uint8, device="cuda")
    f = torch.empty((B, E), dtype=torch.uint8, device="cuda")
    fl = torch.empty((B, E), dtype=torch.float32, device="cuda")
    llr = torch.empty((B, h.n_cw), dtype=torch.float32, device="cuda")
    hard = torch.empty((B, h.K), dtype=torch.uint8, device="cuda")
    ok = torch.empty(B, dtype=torch.uint8, device="cuda")
    rm = capi.Rm(E, 0, h.N, h.K, 2)
    h.encode_raw(info, B, cw, mem=capi.MEM_DEVICE, stream=st)
    h.rate_match_raw(cw, B, rm, f, mem=capi.MEM_DEVICE, stream=st)
    h.qpsk_awgn_llr_raw(f, B, E, 10 ** (-0.2), 7, 0, fl, stream=st)
    h.rate_recover_raw(fl, B, rm, None, llr, mem=capi.MEM_DEVICE, stream=st)
    h.decode_raw(llr, B, hard, ok=ok, mem=capi.MEM_DEVICE, stream=st)
    torch.cuda.synchronize()
    assert torch.equal(cw[:, :h.K], info)
    assert torch.equal(hard, info) and bool(ok.all())
    # noise statistics of the channel leg: LLR = 2*sqrt(2)*(x+n)/var, so var(LLR | bit) = 4/var
    var = 10 ** (-0.2)
    s = (fl * (1 - 2.0 * f.float()))
    assert abs(float(s.mean()) - 2.0 / var) < 0.01 * 2.0 / var
    assert abs(float(s.var()) - 4.0 / var) < 0.01 * 4.0 / var
    llr.zero_()
    h.decode_raw(llr, B, hard, mem=capi.MEM_DEVICE, stream=st)
    torch.cuda.synchronize()
    assert not bool(hard.any())
    h.close()


@pytest.mark.parametrize("dtype", ["f32", "f16x2"])
@pytest.mark.parametrize("E,n_rows,iters,esn0", [(9478, 5, 20, 6.3), (25272, 46, 8, -0.35), (16896, 24, 8, 1.7)])
def test_full_size_parity_flag_is_exactly_the_syndrome(capi, E, n_rows, iters, esn0, dtype):
    """BASELINE config 4 (BG1 Z=384 rate 8/9, 20 iterations with the parity-check stop, batch 4096), the headline code and a
    rate-1/2 case with the stop, near their waterfalls so that converged and failed blocks both occur.  Size-independent
    property of 'Parity check satisfied' (NRLDPCDecoder.m:120): parity_ok = 1 exactly when the signs of the returned
    a-posteriori values form a codeword on the active part of H -- i.e. when re-encoding the decoded information bits
    reproduces them (the parity bits of the active columns are uniquely determined by the active checks); blocks that
    stop early report fewer iterations than the limit, the others exactly the limit."""
    import torch
    B, Z = 4096, 384
    h = capi.Handle(1, Z, iters, True, llr_dtype=capi.F16X2 if dtype == "f16x2" else capi.F32)
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(11)
    info = torch.randint(0, 2, (B, h.K), dtype=torch.uint8, device="cuda", generator=g)
    cw = torch.empty((B, h.n_cw), dtype=torch.uint8, device="cuda")
    f = torch.empty((B, E), dtype=torch.uint8, device="cuda")
    fl = torch.empty((B, E), dtype=torch.float32, device="cuda")
    llr = torch.empty((B, h.n_cw), dtype=torch.float32, device="cuda")
    hard = torch.empty((B, h.K), dtype=torch.uint8, device="cuda")
    app = torch.empty((B, h.n_cw), dtype=torch.float32, device="cuda")
    ok = torch.empty(B, dtype=torch.uint8, device="cuda")
    its = torch.empty(B, dtype=torch.int32, device="cuda")
    rm = capi.Rm(E, 0, h.N, h.K, 2)
    h.encode_raw(info, B, cw, mem=capi.MEM_DEVICE, stream=st)
    h.rate_match_raw(cw, B, rm, f, mem=capi.MEM_DEVICE, stream=st)
    h.qpsk_awgn_llr_raw(f, B, E, 10 ** (-esn0 / 10), 3, 0, fl, stream=st)
    h.rate_recover_raw(fl, B, rm, None, llr, mem=capi.MEM_DEVICE, stream=st)
    h.decode_raw(llr, B, hard, soft=app, iters=its, ok=ok, n_rows=n_rows, mem=capi.MEM_DEVICE, stream=st)
    recw = torch.empty_like(cw)
    h.encode_raw(hard, B, recw, mem=capi.MEM_DEVICE, stream=st)
    torch.cuda.synchronize()
    n_act = (22 + n_rows) * Z
    is_cw = ((app[:, :n_act] < 0).to(torch.uint8) == recw[:, :n_act]).all(dim=1)
    assert torch.equal(is_cw, ok.bool())
    assert int(ok.sum()) > 0
    if n_rows != 24:                                               # the two BASELINE operating points: BLER about 1e-2
        assert int(ok.sum()) < B                                   # converged and failed blocks both occur
    assert bool((its[~ok.bool()] == iters).all()) and bool((its[ok.bool()] <= iters).all()) and bool((its >= 1).all())
    assert bool((hard[ok.bool()] == info[ok.bool()]).all(dim=1).float().mean() > 0.99)   # undetected errors are rare
    h.close()


def test_small_z_high_batch_config3(capi, O):
    """BASELINE config 3 (BG2 Z=52, 104 filler, E=2000, 33 active rows, batch 65536): GPU batch equals the
    oracle on a strided sample and is idempotent (decoding twice gives the same bytes)."""
    import torch
    rng = np.random.default_rng(52)
    info, llr = make_llr(O, 2, 52, 512, 2000, -1.0, rng, filler=104)
    big = np.tile(llr, (128, 1))
    h = capi.Handle(2, 52, 8, False)
    out1 = h.decode(big, n_rows=33)
    out2 = h.decode(big, n_rows=33)
    assert (out1["hard"] == out2["hard"]).all()
    ref = O.decode_nms(2, 52, llr, 8, n_rows=33)
    for rep in (0, 57, 127):
        assert (out1["hard"][rep * 512:(rep + 1) * 512] == ref["hard"]).all()
    h.close()


@pytest.mark.parametrize("Qm", [1, 2, 4, 6, 8])
def test_modulate_demodulate_against_oracle(capi, O, Qm):
    """nrldpc_modulate bit-exact against the oracle map; nrldpc_demodulate (exact / max-log / hard) against the
    float64 two-dimensional definition within 2e-3 + 2e-4*|L| (float32 log-sum-exp); the fused channel kernel is
    bit-identical to modulate + awgn + demodulate; noise has the requested variance."""
    import torch
    rng = np.random.default_rng(40 + Qm)
    n_sym = 20001
    bits = rng.integers(0, 2, n_sym * Qm, dtype=np.uint8)
    h = capi.Handle(2, 6, 4)
    st = torch.cuda.current_stream().cuda_stream
    d_bits = torch.from_numpy(bits).cuda()
    sym = torch.empty((n_sym, 2), dtype=torch.float32, device="cuda")
    h.modulate_raw(d_bits, bits.size, Qm, sym, stream=st)
    ref = O.modulate(bits, Qm)
    got = sym.cpu().numpy()
    assert (got[:, 0] == ref.real.astype(np.float32)).all() and (got[:, 1] == ref.imag.astype(np.float32)).all()
    for var in (0.5, 0.02):
        noisy = sym.clone()
        h.awgn_raw(noisy, n_sym, var, 99, 3, stream=st)
        n = (noisy - sym).cpu().numpy().astype(np.float64)
        assert abs(n.mean()) < 4 * np.sqrt(var / 2 / n.size) and abs(n.var() - var / 2) < 0.03 * var / 2
        assert abs(np.mean(n[:, 0] * n[:, 1])) < 0.03 * var / 2
        rx = noisy.cpu().numpy()
        rxc = rx[:, 0] + 1j * rx[:, 1]
        for name, method in (("Log-likelihood ratio", capi.DEMOD_LLR), ("Approximate log-likelihood ratio", capi.DEMOD_APPROX),
                             ("Hard decision", capi.DEMOD_HARD)):
            out = torch.empty(n_sym * Qm, dtype=torch.float32, device="cuda")
            h.demodulate_raw(noisy, n_sym, Qm, var, method, out, stream=st)
            o = out.cpu().numpy().astype(np.float64)
            r = O.demodulate(rxc, Qm, var, name)
            if method == capi.DEMOD_HARD:
                assert (o != r).mean() < 1e-4        # a float32 / float64 tie-break on a decision boundary at most
            else:
                assert (np.abs(o - r) <= 2e-3 + 2e-4 * np.abs(r)).all(), (Qm, name, np.abs(o - r).max())
            fused = torch.empty(n_sym * Qm, dtype=torch.float32, device="cuda")
            h.mod_awgn_llr_raw(d_bits, bits.size, Qm, var, method, 99, 3, fused, stream=st)
            assert torch.equal(fused.view(torch.int32), out.view(torch.int32))
    with pytest.raises(capi.UnsupportedParameters):
        h.modulate_raw(d_bits, 6, 3, sym, stream=st)
    with pytest.raises(capi.NRLDPCError):
        h.modulate_raw(d_bits, 7, 2, sym, stream=st)
    h.close()


def test_modem_system_objects(capi, O):
    """NRModulator / NRDemodulator mirrors: property names, step protocol, tunable Variance, error identifier."""
    from ldpc_3gpp_matlab_b200.modem import NRModulator, NRDemodulator
    rng = np.random.default_rng(3)
    for name, Qm in (("BPSK", 1), ("QPSK", 2), ("16QAM", 4), ("64QAM", 6), ("256QAM", 8)):
        mod, dem = NRModulator(Modulation=name), NRDemodulator(Modulation=name)
        assert mod.Q_m == Qm and dem.ModulationOrder == 1 << Qm
        bits = rng.integers(0, 2, 120 * Qm, dtype=np.uint8)
        tx = mod.step(bits)
        assert np.allclose(tx, O.modulate(bits, Qm), atol=1e-7)
        dem.Variance = 0.01
        assert ((dem.step(tx) < 0) == (bits == 1)).all()
        dem2 = NRDemodulator(Modulation=name, DecisionMethod="Hard decision")
        assert (dem2.step(tx) == bits).all()
        for o in (mod, dem, dem2):
            o.release()
    with pytest.raises(capi.UnsupportedParameters):
        NRModulator(Modulation="8PSK").step(np.zeros(3, np.uint8))


def test_decode16_half_transport(capi, O):
    """nrldpc_decode16 (LLRs transported as binary16): host and device paths, both arithmetic modes.  Packed-half handle:
    bit-identical to nrldpc_decode on the float32 originals (the kernel rounds them the same way).  Float32 handle:
    equals the oracle on the binary16-rounded LLRs."""
    import torch
    rng = np.random.default_rng(61)
    for bg, Z, B, E in ((1, 384, 9, 25272), (2, 52, 77, 2000), (2, 6, 300, 100)):
        info, llr = make_llr(O, bg, Z, B, E, 0.5, rng)
        llr[0, 5 * Z:5 * Z + 3] = np.inf
        with np.errstate(over="ignore"):
            l16 = llr.astype(np.float16)
        for dt, f16 in ((capi.F16X2, True), (capi.F32, False)):
            h = capi.Handle(bg, Z, 6, True, llr_dtype=dt)
            hard = np.zeros((B, h.K), np.uint8); soft = np.zeros((B, h.n_cw), np.float32)
            it = np.zeros(B, np.int32); ok = np.zeros(B, np.uint8)
            h.decode16_raw(l16, B, hard, soft, it, ok)
            ref = O.decode_nms(bg, Z, l16.astype(np.float32), 6, early_term=True, f16=f16)
            assert (hard == ref["hard"]).all() and _same_bits(soft, ref["app"]) and (it == ref["iters"]).all()
            if f16:
                direct = h.decode(llr, want_soft=True)
                assert (direct["hard"] == hard).all() and _same_bits(direct["app"], soft)
            d16 = torch.from_numpy(l16).cuda()
            dh = torch.zeros((B, h.K), dtype=torch.uint8, device="cuda")
            h.decode16_raw(d16, B, dh, mem=capi.MEM_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert (dh.cpu().numpy() == hard).all()
            h.close()


def test_decode8_int8_transport(capi, O):
    """nrldpc_decode8 (LLRs transported as int8, llr = scale * q, q = 127 = filler): bit-identical to nrldpc_decode / the oracle on
    the float32 values scale * q -- host path (pinned and pageable), device path, both arithmetics, early stop, soft output,
    trimmed rows, extreme codes (-128, 126, 127, 0)."""
    import torch
    rng = np.random.default_rng(88)
    for bg, Z, B, E, scale in ((1, 384, 9, 25272, 0.125), (2, 52, 600, 2000, 0.25), (2, 6, 300, 100, 0.5), (1, 22, 41, 1200, 0.1)):
        info, llr = make_llr(O, bg, Z, B, E, 0.5, rng)
        q = np.clip(np.rint(llr / scale), -127, 126).astype(np.int8)
        q[0, 5 * Z:5 * Z + 3] = 127                      # filler marks
        q[1, 7 * Z] = -128
        q[1, 7 * Z + 1] = 126
        deq = (np.float32(scale) * q.astype(np.float32)).astype(np.float32)
        deq[q == 127] = np.inf
        rows = 0 if bg == 2 else max(4, -(-(E + 2 * Z) // Z) - 22)
        for dt, f16 in ((capi.F32, False), (capi.F16X2, True)):
            h = capi.Handle(bg, Z, 6, True, llr_dtype=dt)
            ref = O.decode_nms(bg, Z, deq, 6, early_term=True, f16=f16, n_rows=rows if rows else None)
            hard = np.zeros((B, h.K), np.uint8); soft = np.zeros((B, h.n_cw), np.float32)
            it = np.zeros(B, np.int32); ok = np.zeros(B, np.uint8)
            h.decode8_raw(q, scale, B, hard, soft, it, ok, n_rows=rows)            # pageable host memory
            assert (hard == ref["hard"]).all() and _same_bits(soft, ref["app"]), (bg, Z, f16)
            assert (it == ref["iters"]).all() and (ok == ref["parity_ok"]).all(), (bg, Z, f16)
            qp = torch.from_numpy(q).pin_memory()
            hp = torch.zeros((B, h.K), dtype=torch.uint8).pin_memory()
            h.decode8_raw(qp, scale, B, hp, n_rows=rows)                           # pinned host memory
            assert (hp.numpy() == ref["hard"]).all()
            dq = torch.from_numpy(q).cuda()
            dh = torch.zeros((B, h.K), dtype=torch.uint8, device="cuda")
            h.decode8_raw(dq, scale, B, dh, n_rows=rows, mem=capi.MEM_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert (dh.cpu().numpy() == ref["hard"]).all()
            with pytest.raises(capi.NRLDPCError):
                h.decode8_raw(q, 0.0, B, hard)
            h.close()


# ---- NRLDPC_ALG_BP: the reference's own algorithm (flooding sum-product, float64) on the device -------------
# Checker: oracle B (oracle/nrldpc_oracle.c, decode_bp_one), the restatement of MathWorks' documented comm.LDPCDecoder
# algorithm as configured at NRLDPCDecoder.m:120.  Kernel and oracle perform every +, -, * in the same order; they
# differ only in the tanh / atanh library (CUDA vs glibc, ~1 ulp each).  Stated bar: hard decisions, iteration counts
# and parity flags IDENTICAL; a-posteriori values within 1e-9 relative (float64 path) / float32 rounding of that.
BP_RTOL = 1e-9


def _bp_close(app, ref, rtol):
    fin = np.isfinite(ref)
    assert (np.isfinite(app) == fin).all()
    assert (app[~fin] == ref[~fin]).all()            # +-inf fillers stay +-inf
    return bool((np.abs(app[fin] - ref[fin]) <= rtol * np.maximum(1.0, np.abs(ref[fin]))).all())


@pytest.mark.parametrize("bg", [1, 2])
def test_decode_bp_matches_reference_algorithm_all_set_indices(capi, O, bg):
    """One lifting size per set index (plus the smallest and the largest), ragged batches, filler (+inf), punctured
    zeros, both termination rules; float64 buffers (nrldpc_decode64) as the reference passes them."""
    rng = np.random.default_rng(900 + bg)
    for Z in (2, 3, 5, 7, 9, 11, 13, 15, 36, 52, 96, 208, 384):
        d = O.dims(bg, Z)
        B = 5 if Z > 100 else 9
        E = int(d["N"] * rng.uniform(0.4, 1.0)) // 2 * 2
        filler = int(rng.integers(0, 3)) * (Z // 2)
        info, llr = make_llr(O, bg, Z, B, E, rng.uniform(0.0, 3.0), rng, filler=filler)
        llr64 = llr.astype(np.float64) * 1.000000001   # genuinely double-valued inputs
        for et in (True, False):
            ref = O.decode_bp(bg, Z, llr64, 6, early_term=et, want_app=True)
            h = capi.Handle(bg, Z, 6, et, algorithm=capi.ALG_BP)
            out = h.decode(llr64, want_soft=True)
            h.close()
            assert out["app"].dtype == np.float64
            assert (out["hard"] == ref["hard"]).all(), (bg, Z, et)
            assert (out["iters"] == ref["iters"]).all(), (bg, Z, et, out["iters"], ref["iters"])
            assert (out["parity_ok"] == ref["parity_ok"]).all(), (bg, Z, et)
            assert _bp_close(out["app"], ref["app"], BP_RTOL), (bg, Z, et)


@pytest.mark.parametrize("bg,Z,rows", [(1, 384, 46), (1, 384, 5), (2, 52, 33), (2, 6, 13), (1, 30, 4)])
def test_decode_bp_float32_buffers_special_values_and_row_trimming(capi, O, bg, Z, rows):
    """nrldpc_decode (float32 buffers) in BP mode: inputs are widened exactly; NaN and +inf mark filler, -inf, zeros,
    -0.0 and huge magnitudes go through the same arithmetic as in the restatement."""
    rng = np.random.default_rng(Z + rows + 5)
    d = O.dims(bg, Z)
    B = 6
    llr = (rng.normal(1.0, 3, (B, d["ncw"]))).astype(np.float32)
    llr[:, :2 * Z] = 0
    llr[:, (d["kcols"] + rows) * Z:] = 0
    llr[0, 3 * Z:3 * Z + Z // 2] = np.inf
    llr[1, 3 * Z:3 * Z + Z // 2] = np.nan
    llr[2, 5 * Z] = -np.inf
    llr[3, ::7] = 0.0
    llr[3, 1::11] = -0.0
    llr[4, 2 * Z::5] *= 1e20
    llr[5] *= 100.0
    ref_in = llr.astype(np.float64)
    ref_in[np.isnan(ref_in)] = np.inf     # NRLDPCDecoder.m:264
    for et in (True, False):
        ref = O.decode_bp(bg, Z, ref_in, 5, n_rows=rows, early_term=et, want_app=True)
        h = capi.Handle(bg, Z, 5, et, algorithm=capi.ALG_BP)
        out = h.decode(llr, n_rows=rows, want_soft=True)
        h.close()
        assert (out["hard"] == ref["hard"]).all() and (out["iters"] == ref["iters"]).all()
        assert (out["parity_ok"] == ref["parity_ok"]).all()
        assert out["app"].dtype == np.float32
        with np.errstate(over="ignore"):
            ref32 = ref["app"].astype(np.float32)
        assert _bp_close(out["app"].astype(np.float64), ref32.astype(np.float64), 1e-6)


def test_decode_bp_headline_code_and_device_buffers(capi, O):
    """BG1 Z=384 rate 1/3 near its waterfall through device pointers (torch tensors): identical block decisions and
    iteration counts to the restatement of the reference's decoder; then the same call on binary16 and float64
    device buffers (transport conversions in front of the same kernel)."""
    import torch
    rng = np.random.default_rng(4242)
    info, llr = make_llr(O, 1, 384, 24, 25272, -0.6, rng)
    ref = O.decode_bp(1, 384, llr, 12, early_term=True)
    h = capi.Handle(1, 384, 12, True, algorithm=capi.ALG_BP)
    dev = torch.device("cuda:0")
    t_llr = torch.from_numpy(llr).to(dev)
    hard = torch.zeros((24, h.K), dtype=torch.uint8, device=dev)
    iters = torch.zeros(24, dtype=torch.int32, device=dev)
    ok = torch.zeros(24, dtype=torch.uint8, device=dev)
    h.decode_raw(t_llr, 24, hard, None, iters, ok, mem=capi.MEM_DEVICE, stream=None)
    h.synchronize()
    assert (hard.cpu().numpy() == ref["hard"]).all()
    assert (iters.cpu().numpy() == ref["iters"]).all() and (ok.cpu().numpy() == ref["parity_ok"]).all()
    assert ref["iters"].min() < 12 and ref["iters"].max() >= 8    # the operating point exercises early stopping
    # float64 device buffers
    t64 = t_llr.double()
    hard2 = torch.zeros_like(hard)
    h.decode64_raw(t64, 24, hard2, None, iters, ok, mem=capi.MEM_DEVICE, stream=None)
    h.synchronize()
    assert (hard2.cpu().numpy() == ref["hard"]).all() and (iters.cpu().numpy() == ref["iters"]).all()
    # binary16 transport: the restatement sees the same rounded values
    l16 = llr.astype(np.float16)
    ref16 = O.decode_bp(1, 384, l16.astype(np.float32), 12, early_term=True)
    h.decode16_raw(torch.from_numpy(l16).to(dev), 24, hard2, None, iters, ok, mem=capi.MEM_DEVICE, stream=None)
    h.synchronize()
    assert (hard2.cpu().numpy() == ref16["hard"]).all() and (iters.cpu().numpy() == ref16["iters"]).all()
    h.close()


def test_decode64_with_default_algorithm_rounds_to_float32(capi, O):
    """nrldpc_decode64 in front of the min-sum kernel: doubles are rounded to float32 on the device; app_soft in
    float64 is refused (UnsupportedParameters), BP + packed-half is refused at create."""
    rng = np.random.default_rng(5)
    info, llr = make_llr(O, 2, 52, 11, 2000, 1.0, rng, filler=104)
    llr64 = llr.astype(np.float64) * (1 + 1e-12)
    ref = O.decode_nms(2, 52, llr64.astype(np.float32), 8, early_term=True)
    h = capi.Handle(2, 52, 8, True)
    out = h.decode(llr64)
    assert (out["hard"] == ref["hard"]).all() and (out["iters"] == ref["iters"]).all()
    with pytest.raises(capi.UnsupportedParameters):
        h.decode(llr64, want_soft=True)
    h.close()
    with pytest.raises(capi.UnsupportedParameters):
        capi.Handle(2, 52, 8, True, llr_dtype=capi.F16X2, algorithm=capi.ALG_BP)
    with pytest.raises(capi.UnsupportedParameters):
        capi.Handle(2, 52, 8, True, algorithm=7)


def test_decode_bp_golden_vectors_on_gpu(capi, golden_decode, golden_decode_bp):
    """The CUDA sum-product kernel against the committed oracle-B vectors (decisions, iteration counts, parity flags)."""
    for n in sorted({k.split("__")[0] for k in golden_decode_bp.files}):
        bg, Z, iters = golden_decode_bp[n + "__cfg"].tolist()
        for tag, early in (("stop", True), ("full", False)):
            h = capi.Handle(bg, Z, iters, early, algorithm=capi.ALG_BP)
            out = h.decode(golden_decode[n + "__llr"])
            h.close()
            assert (np.packbits(out["hard"], axis=1) == golden_decode_bp[f"{n}__{tag}__hard"]).all(), (n, tag)
            assert (out["iters"] == golden_decode_bp[f"{n}__{tag}__iters"]).all(), (n, tag)
            assert (out["parity_ok"] == golden_decode_bp[f"{n}__{tag}__ok"]).all(), (n, tag)


def test_config4_sample_against_oracle(capi, O):
    """BASELINE.json configs[3]: BG1 Z=384 rate 8/9 (E = 9478, 5 active rows), <= 20 iterations with the reference's
    parity-check stop (NRLDPCDecoder.m:120), 256 blocks straddling the waterfall (6.25 dB: BLER ~ 1e-2; 6.0 dB: about half
    the blocks fail and run all 20 iterations): decisions, iteration counts, parity flags and APP bit patterns equal
    oracle A; packed-half equals oracle A16."""
    rng = np.random.default_rng(4)
    a = make_llr(O, 1, 384, 192, 9478, 6.25, rng)
    b = make_llr(O, 1, 384, 64, 9478, 6.0, rng)
    info, llr = np.concatenate([a[0], b[0]]), np.concatenate([a[1], b[1]])
    ref = O.decode_nms(1, 384, llr, 20, early_term=True, n_rows=5)
    assert 1 < ref["iters"].min() < ref["iters"].max() == 20          # the sample exercises early and late stops
    h = capi.Handle(1, 384, 20, True)
    out = h.decode(llr, n_rows=5, want_soft=True)
    h.close()
    assert (out["hard"] == ref["hard"]).all() and _same_bits(out["app"], ref["app"])
    assert (out["iters"] == ref["iters"]).all() and (out["parity_ok"] == ref["parity_ok"]).all()
    ref16 = O.decode_nms(1, 384, llr, 20, early_term=True, n_rows=5, f16=True)
    h = capi.Handle(1, 384, 20, True, llr_dtype=capi.F16X2)
    out = h.decode(llr, n_rows=5, want_soft=True)
    h.close()
    assert (out["hard"] == ref16["hard"]).all() and _same_bits(out["app"], ref16["app"])
    assert (out["iters"] == ref16["iters"]).all() and (out["parity_ok"] == ref16["parity_ok"]).all()


def test_live_handles_of_different_sizes_do_not_invalidate_each_other(capi, O):
    """Two decoders / encoders alive at once whose kernels share an instantiation but need different amounts of dynamic
    shared memory (a receiver alternating between two transport-block sizes): the kernel attribute is process-wide, so a
    per-handle cache of it used to make the larger handle's next launch fail with cudaErrorInvalidValue."""
    rng = np.random.default_rng(12)
    big = [capi.Handle(1, 96, 6, False), capi.Handle(2, 176, 6, True)]
    small = [capi.Handle(1, 8, 6, False), capi.Handle(2, 20, 6, True)]
    data = {}
    for h in big + small:
        d = O.dims(h.bg, h.Z)
        info, llr = make_llr(O, h.bg, h.Z, 5, d["N"] // 2 * 2, 1.5, rng)
        data[id(h)] = (info, llr, O.decode_nms(h.bg, h.Z, llr, 6, early_term=(h.bg == 2)))
    for _ in range(3):                      # alternate: large, small, large, ...
        for hb, hs in zip(big, small):
            for h in (hb, hs):
                info, llr, ref = data[id(h)]
                out = h.decode(llr)
                assert (out["hard"] == ref["hard"]).all() and (out["iters"] == ref["iters"]).all()
                assert (h.encode(info) == O.encode(h.bg, h.Z, info)).all()
    for h in big + small:
        h.close()


@pytest.mark.parametrize("bg,Z,B", [(1, 384, 700), (2, 52, 3000), (1, 7, 1)])
def test_decode64_pageable_host_path(capi, O, bg, Z, B):
    """nrldpc_decode64 on ordinary (pageable) float64 memory with pageable outputs -- the call matlab/nrldpc_mex.cpp makes on
    mxGetPr memory (NRLDPCDecoder.m:262-265): host threads narrow to float32 into the pinned ring.  Several chunks per call
    (B above one persistent-grid wave), ragged last chunk; results equal the pinned float32 path and the oracle."""
    rng = np.random.default_rng(B)
    d = O.dims(bg, Z)
    info, llr = make_llr(O, bg, Z, B, d["N"] // 2 * 2, 0.5 if bg == 1 else 1.0, rng, filler=Z if Z > 7 else 0)
    llr64 = llr.astype(np.float64)
    h = capi.Handle(bg, Z, 8, True)
    a = h.decode(llr)                     # float32, numpy (pageable) buffers: staged copy
    b = h.decode(llr64)                   # float64: staged narrowing
    h.close()
    n = min(B, 64)
    ref = O.decode_nms(bg, Z, llr[:n], 8, early_term=True)
    for r in (a, b):
        assert (r["hard"][:n] == ref["hard"]).all() and (r["iters"][:n] == ref["iters"]).all() and (r["parity_ok"][:n] == ref["parity_ok"]).all()
    assert (a["hard"] == b["hard"]).all() and (a["iters"] == b["iters"]).all() and (a["parity_ok"] == b["parity_ok"]).all()
    # a payload-level property at full size: nearly every block decodes at this SNR
    assert (a["hard"] != info).any(axis=1).mean() < 0.05


@pytest.mark.parametrize("f16", [False, True])
def test_forced_cta_shapes_are_bit_identical(capi, O, f16, monkeypatch):
    """Every CTA shape the library may pick (codewords per CTA, resident CTAs per SM: decode_shapes.inc) is the same
    arithmetic: forced shapes -- one codeword per CTA with a partially filled last warp (the lane-masked CTA-uniform
    kernel), narrow CTAs sharing a scratch block, eight CTAs per SM -- against the oracle, fixed iterations and the
    parity-check stop, trimmed rows, soft output."""
    rng = np.random.default_rng(31)
    for bg, Z, rows in ((1, 52, 46), (2, 52, 33), (1, 208, 46), (2, 240, 20), (1, 20, 46), (2, 112, 42)):
        d = O.dims(bg, Z)
        B = 2 * max(3, 600 // Z) + 1
        info, llr = make_llr(O, bg, Z, B, (d["kcols"] - 2 + rows) * Z, rng.uniform(0.0, 2.5), rng, filler=Z)
        for et in (False, True):
            ref = O.decode_nms(bg, Z, llr, 6, early_term=et, n_rows=rows, f16=f16)
            for cw, cap in ((1, 8), (1, 2), (2, 4), (max(1, 384 // Z), 4), (max(1, 128 // Z), 8)):
                if cw > max(1, 384 // Z):
                    continue
                monkeypatch.setenv("NRLDPC_CWPC", str(cw))
                monkeypatch.setenv("NRLDPC_OCC_CAP", str(cap))
                h = capi.Handle(bg, Z, 6, et, llr_dtype=capi.F16X2 if f16 else capi.F32)
                out = h.decode(llr, n_rows=rows, want_soft=True)
                h.close()
                assert (out["hard"] == ref["hard"]).all() and _same_bits(out["app"], ref["app"]), (bg, Z, et, cw, cap)
                assert (out["iters"] == ref["iters"]).all() and (out["parity_ok"] == ref["parity_ok"]).all(), (bg, Z, et, cw, cap)


@pytest.mark.parametrize("bg,Z,rows,B", [(2, 52, 33, 3001), (1, 8, 46, 4000), (2, 6, 13, 5000), (1, 96, 20, 700), (2, 176, 42, 500), (1, 30, 46, 1500), (2, 4, 42, 6000)])
def test_slot_refill_under_the_stop(capi, O, bg, Z, rows, B, monkeypatch):
    """Multi-codeword CTAs with 'Parity check satisfied' (NRLDPCDecoder.m:120): slots are refilled one by one as their
    codewords converge (decode_nms_refill_kernel).  Batches of many CTA loads with a mix of operating points (codewords
    needing 1 ... max iterations side by side, some never converging), ragged tail, soft output: decisions, iteration
    counts, parity flags and APP bit patterns equal the oracle and the group kernel (NRLDPC_REFILL=0)."""
    rng = np.random.default_rng(Z + B)
    d = O.dims(bg, Z)
    E = (d["kcols"] - 2 + rows) * Z
    parts = [make_llr(O, bg, Z, B // 3 + (i < B % 3), E, esn0, rng, filler=(Z if Z > 7 else 0)) for i, esn0 in enumerate((6.0, 1.0, -2.5))]
    info = np.concatenate([p[0] for p in parts]); llr = np.concatenate([p[1] for p in parts])
    perm = rng.permutation(B)
    info, llr = info[perm], np.ascontiguousarray(llr[perm])
    ref = O.decode_nms(bg, Z, llr, 7, early_term=True, n_rows=rows)
    assert ref["iters"].min() < 3 and ref["iters"].max() == 7
    monkeypatch.setenv("NRLDPC_REFILL", "2")              # refill whenever possible (the default enables it for wide groups only)
    h = capi.Handle(bg, Z, 7, True)
    out = h.decode(llr, n_rows=rows, want_soft=True)
    out2 = h.decode(llr, n_rows=rows)                     # no soft output: other code path for the final records
    h.close()
    # prefetched refill (decode_kernel_refill.cuh; Z a multiple of 4, other sizes fall back) with 1, 2 and 5 mailboxes per CTA
    monkeypatch.setenv("NRLDPC_REFILL", "3")
    pre = []
    for spares in ("1", "2", "5"):
        monkeypatch.setenv("NRLDPC_REFILL_SPARES", spares)
        h = capi.Handle(bg, Z, 7, True)
        pre.append(h.decode(llr, n_rows=rows, want_soft=True))
        pre.append(h.decode(llr[:B // 7], n_rows=rows, want_soft=True))   # the same handle again, a batch that ends inside a CTA load
        h.close()
    monkeypatch.delenv("NRLDPC_REFILL_SPARES")
    for i, r in enumerate(pre):
        n = B if i % 2 == 0 else B // 7
        assert (r["hard"] == ref["hard"][:n]).all() and _same_bits(r["app"], ref["app"][:n]), (i, "prefetched refill")
        assert (r["iters"] == ref["iters"][:n]).all() and (r["parity_ok"] == ref["parity_ok"][:n]).all(), (i, "prefetched refill")
    monkeypatch.setenv("NRLDPC_REFILL", "0")              # whole groups, parity bits tracked in registers
    h = capi.Handle(bg, Z, 7, True)
    grp = h.decode(llr, n_rows=rows, want_soft=True)
    h.close()
    monkeypatch.delenv("NRLDPC_REFILL")
    h = capi.Handle(bg, Z, 7, True)
    dflt = h.decode(llr, n_rows=rows, want_soft=True)
    h.close()
    for r in (out, grp, dflt):
        assert (r["hard"] == ref["hard"]).all() and _same_bits(r["app"], ref["app"])
        assert (r["iters"] == ref["iters"]).all() and (r["parity_ok"] == ref["parity_ok"]).all()
    assert (out2["hard"] == ref["hard"]).all() and (out2["iters"] == ref["iters"]).all() and (out2["parity_ok"] == ref["parity_ok"]).all()


def test_device_mode_streams_and_graph_capture(capi, O):
    """NRLDPC_MEM_DEVICE decodes of one handle from two streams (the handle's scratch is shared: the library orders the
    launches) and a decode captured into a CUDA graph and replayed (the work counter must not depend on host-side state
    baked in at capture time): every result equals the oracle."""
    import torch
    rng = np.random.default_rng(41)
    bg, Z = 1, 96
    d = O.dims(bg, Z)
    info, llr_np = make_llr(O, bg, Z, 300, d["N"], 0.8, rng)
    ref = O.decode_nms(bg, Z, llr_np, 6, early_term=True)
    llr = torch.from_numpy(llr_np).cuda()
    h = capi.Handle(bg, Z, 6, True)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for i in range(6):
        st = (s1, s2)[i % 2]
        hard = torch.zeros((300, d["K"]), dtype=torch.uint8, device="cuda")
        it = torch.zeros(300, dtype=torch.int32, device="cuda")
        with torch.cuda.stream(st):
            h.decode_raw(llr, 300, hard, iters=it, mem=capi.MEM_DEVICE, stream=st.cuda_stream)
        outs.append((hard, it))
    torch.cuda.synchronize()
    for hard, it in outs:
        assert (hard.cpu().numpy() == ref["hard"]).all() and (it.cpu().numpy() == ref["iters"]).all()
    # graph capture + replay
    hard = torch.zeros((300, d["K"]), dtype=torch.uint8, device="cuda")
    it = torch.zeros(300, dtype=torch.int32, device="cuda")
    g = torch.cuda.CUDAGraph()
    cap_stream = torch.cuda.Stream()
    with torch.cuda.stream(cap_stream):
        h.decode_raw(llr, 300, hard, iters=it, mem=capi.MEM_DEVICE, stream=cap_stream.cuda_stream)   # warm-up outside capture
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=cap_stream):
        h.decode_raw(llr, 300, hard, iters=it, mem=capi.MEM_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        hard.zero_(); it.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert (hard.cpu().numpy() == ref["hard"]).all() and (it.cpu().numpy() == ref["iters"]).all()
    # an ordinary launch after the captured one starts from a clean counter again
    out = h.decode(llr_np)
    assert (out["hard"] == ref["hard"]).all() and (out["iters"] == ref["iters"]).all()
    h.close()


@pytest.mark.parametrize("bg", [1, 2])
def test_warp_shuffle_mapping_bit_exact_small_Z(capi, O, bg, monkeypatch):
    """The lane-per-edge kernel with the check-node reduction done by warp shuffles (decode_kernel_shfl.cuh,
    NRLDPC_DECODE_VARIANT=shfl, every Z <= 32): butterfly minima / ballot signs give the same bits as the register scan of
    the default kernel and as the oracle -- fixed iterations and the stop, trimmed rows, filler, soft output, ragged batches,
    one and several codewords per CTA, CTAs of 64 and 256 threads, a core-pass / extension-fail mix."""
    monkeypatch.setenv("NRLDPC_DECODE_VARIANT", "shfl")
    rng = np.random.default_rng(700 + bg)
    rows_all = 46 if bg == 1 else 42
    for Z in [z for z in ALL_Z if z <= 32]:
        d = O.dims(bg, Z)
        B = 37 if Z > 8 else 75
        rows = int(rng.integers(4, rows_all + 1)) if Z % 3 else rows_all
        E = (d["K"] // Z - 2 + rows) * Z
        info, llr = make_llr(O, bg, Z, B, E, rng.uniform(-1.0, 3.0), rng, filler=(Z if Z % 2 == 0 and Z > 4 else 0))
        for et in (False, True):
            ref = O.decode_nms(bg, Z, llr, 6, early_term=et, n_rows=rows)
            for cwpc, threads in (("0", "256"), ("1", "64"), ("5", "128")):
                monkeypatch.setenv("NRLDPC_SHFL_CWPC", cwpc)
                monkeypatch.setenv("NRLDPC_SHFL_THREADS", threads)
                h = capi.Handle(bg, Z, 6, et)
                out = h.decode(llr, n_rows=rows, want_soft=True)
                one = h.decode(llr[:1], n_rows=rows)
                h.close()
                assert (out["hard"] == ref["hard"]).all(), (bg, Z, et, cwpc)
                assert _same_bits(out["app"], ref["app"]), (bg, Z, et, cwpc)
                assert (out["iters"] == ref["iters"]).all() and (out["parity_ok"] == ref["parity_ok"]).all(), (bg, Z, et, cwpc)
                assert (one["hard"] == ref["hard"][:1]).all() and (one["iters"] == ref["iters"][:1]).all(), (bg, Z, et, cwpc)
    # codewords whose core checks hold while an extension check fails for ever
    Z, rows = 16, rows_all
    info, llr, kind = make_core_pass_llr(O, bg, Z, 48, rows, rng)
    ref = O.decode_nms(bg, Z, llr, 5, early_term=True, n_rows=rows)
    monkeypatch.setenv("NRLDPC_SHFL_CWPC", "0")
    monkeypatch.setenv("NRLDPC_SHFL_THREADS", "256")
    h = capi.Handle(bg, Z, 5, True)
    out = h.decode(llr, n_rows=rows, want_soft=True)
    h.close()
    assert (out["hard"] == ref["hard"]).all() and _same_bits(out["app"], ref["app"])
    assert (out["iters"] == ref["iters"]).all() and (out["parity_ok"] == ref["parity_ok"]).all()


def test_handle_lifetime_and_concurrent_host_threads(capi, O):
    """release(obj) gives everything back (NRLDPCDecoder.m releaseImpl -> nrldpc_destroy: 80 create / decode / destroy cycles of
    mixed sizes and arithmetics leave the device's free memory where it was), and distinct handles really are independent:
    two host threads, each with its own handle (different base graph, lifting size and termination rule), decode concurrently
    through the host-memory call and every result equals the oracle."""
    import threading
    import torch
    rng = np.random.default_rng(77)
    cases = []
    for bg, Z, et in ((1, 96, True), (2, 52, False)):
        d = O.dims(bg, Z)
        info, llr = make_llr(O, bg, Z, 300, d["N"], 1.0, rng)
        cases.append((bg, Z, et, llr, O.decode_nms(bg, Z, llr, 6, early_term=et)))
    h = capi.Handle(1, 384, 8, False); h.decode(np.zeros((2, 68 * 384), np.float32)); h.close()     # warm the context / module
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for i in range(80):
        bg, Z, et, llr, ref = cases[i % 2]
        h = capi.Handle(bg, Z, 6, et, llr_dtype=capi.F16X2 if i % 4 == 3 else capi.F32, algorithm=capi.ALG_BP if i % 8 == 5 else capi.ALG_NMS)
        out = h.decode(llr[:64 + i])
        if i % 4 != 3 and i % 8 != 5:
            assert (out["hard"] == ref["hard"][:64 + i]).all(), i
        h.close()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < (32 << 20), (free0, free1)
    errors = []

    def worker(case):
        try:
            bg, Z, et, llr, ref = case
            hh = capi.Handle(bg, Z, 6, et)
            for _ in range(25):
                out = hh.decode(llr, want_soft=True)
                assert (out["hard"] == ref["hard"]).all() and _same_bits(out["app"], ref["app"])
                assert (out["iters"] == ref["iters"]).all() and (out["parity_ok"] == ref["parity_ok"]).all()
            hh.close()
        except Exception as e:      # noqa: BLE001 - reported below
            errors.append(repr(e))
    ts = [threading.Thread(target=worker, args=(c,)) for c in cases]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
