"""Oracle known-answer and property tests -- CPU only."""
import numpy as np
import pytest

from conftest import make_llr


def _bits(s):
    return np.unpackbits(np.frombuffer(s, dtype=np.uint8))


def test_crc_check_values(O, golden_tables):
    """CRC catalogue check values for '123456789' (CRC-16/XMODEM, CRC-24/LTE-A, CRC-24/LTE-B)."""
    for kind, want in golden_tables["crc_check_123456789"].items():
        par = O.crc(kind, _bits(b"123456789"))
        got = int("".join(str(int(b)) for b in par), 2)
        assert got == want, kind


def test_decode_golden_vectors(O, golden_decode):
    names = sorted({k.split("__")[0] for k in golden_decode.files})
    assert len(names) >= 6
    for n in names:
        bg, Z, iters, et, rows = golden_decode[n + "__cfg"].tolist()
        out = O.decode_nms(bg, Z, golden_decode[n + "__llr"], iters, early_term=bool(et), n_rows=rows)
        assert (np.packbits(out["hard"], axis=1) == golden_decode[n + "__hard"]).all(), n
        assert (out["iters"] == golden_decode[n + "__iters"]).all(), n
        assert (out["parity_ok"] == golden_decode[n + "__ok"]).all(), n
        u = out["app"].view(np.uint32)
        assert (np.bitwise_xor.reduce(u, axis=1) == golden_decode[n + "__app_xor"]).all(), n
        assert (u.astype(np.uint64).sum(axis=1) == golden_decode[n + "__app_sum"]).all(), n


@pytest.mark.parametrize("bg,Z", [(1, 2), (1, 15), (2, 3), (2, 52), (1, 96), (2, 240)])
def test_noiseless_round_trip_both_decoders(O, bg, Z):
    rng = np.random.default_rng(Z)
    d = O.dims(bg, Z)
    info = rng.integers(0, 2, (3, d["K"]), dtype=np.uint8)
    cw = O.encode(bg, Z, info)
    llr = (4.0 * (1 - 2.0 * cw)).astype(np.float32)
    llr[:, :2 * Z] = 0  # punctured columns must be recovered
    a = O.decode_nms(bg, Z, llr, 10, early_term=True)
    b = O.decode_bp(bg, Z, llr, 10)
    assert (a["hard"] == info).all() and a["parity_ok"].all() and (a["iters"] <= 3).all()
    assert (b["hard"] == info).all() and b["parity_ok"].all()


def test_filler_and_row_trimming_equivalence(O):
    """Rows whose parity bit was not sent carry a zero degree-1 LLR: trimming them changes nothing."""
    rng = np.random.default_rng(7)
    bg, Z, E, fill = 2, 52, 2000, 104
    info, llr = make_llr(O, bg, Z, 6, E, -2.0, rng, filler=fill)
    full = O.decode_nms(bg, Z, llr, 8, n_rows=0)
    trim = O.decode_nms(bg, Z, llr, 8, n_rows=33)
    assert (full["hard"] == trim["hard"]).all()
    np.testing.assert_array_equal(np.abs(full["app"][:, :35 * Z]), np.abs(trim["app"][:, :35 * Z]))


def test_nms_beats_bp_is_reported_not_assumed(O):
    """Both decoders clear a comfortable SNR; the dB delta between them is a DESIGN.md number."""
    rng = np.random.default_rng(3)
    info, llr = make_llr(O, 2, 6, 200, 100, 6.0, rng, filler=24)
    a = O.decode_nms(2, 6, llr, 8, early_term=True, n_rows=13)
    b = O.decode_bp(2, 6, llr, 8)
    assert (a["hard"][:, :36] != info[:, :36]).any(1).mean() < 0.05
    assert (b["hard"][:, :36] != info[:, :36]).any(1).mean() < 0.05


def test_inf_nan_and_clamp(O):
    rng = np.random.default_rng(5)
    info, llr = make_llr(O, 1, 48, 4, 3000, 0.5, rng, filler=40)
    llr2 = llr.copy()
    llr2[np.isinf(llr2)] = np.nan            # NaN marks filler upstream (NRLDPCDecoder.m:224)
    a, b = O.decode_nms(1, 48, llr), O.decode_nms(1, 48, llr2)
    assert (a["hard"] == b["hard"]).all() and np.isfinite(a["app"]).all()
    assert (a["app"].view(np.uint32) == b["app"].view(np.uint32)).all()


@pytest.mark.parametrize("A,BG,R,Qm,rv", [(20, 2, 0.2, 2, 0), (400, 2, 0.2, 4, 1), (1000, 1, 1 / 3, 6, 2),
                                          (3842, 2, 1 / 3, 2, 3), (8000, 1, 0.5, 8, 0), (500, 1, 0.12, 1, 3)])
def test_rate_match_recover_round_trip(O, A, BG, R, Qm, rv):
    """TX bit_selection/interleave followed by RX deinterleave/bit_selection returns every sent bit to its
    own position; wrapped repetitions add (NRLDPCDecoder.m:230)."""
    import math
    G = int(math.floor(A / R / Qm + 0.5)) * Qm
    p = O.params(BG, A, G, Q_m=Qm, rv_id=rv)
    if p is None:
        pytest.skip("UnsupportedParameters")
    rng = np.random.default_rng(A)
    Z, K, Kp, N = p.Z_c, p.K, p.K_prime, p.N
    info = rng.integers(0, 2, K, dtype=np.uint8)
    info[Kp:] = 0
    cw = O.encode(BG, Z, info)
    d = O.cw_to_d(Z, K, Kp, N, cw)
    E = p.E_r[0]
    e = O.bit_selection_tx(d, p.N_cb, p.k_0, E)
    f = O.interleave_tx(e, Qm)
    assert set(np.unique(f)) <= {0, 1}
    llr_f = (1.0 - 2.0 * f).astype(np.float32)
    e_t = O.deinterleave_rx(llr_f, Qm)
    assert (e_t == 1.0 - 2.0 * e).all()
    d_t = O.bit_selection_rx(e_t, N, p.N_cb, p.k_0, Z, K, Kp)
    filler = np.isnan(d_t)
    assert filler.sum() == K - max(Kp, 2 * Z) if K > max(Kp, 2 * Z) else filler.sum() == 0
    sent = (d_t != 0) & ~filler
    assert (np.sign(d_t[sent]) == 1 - 2.0 * d[sent]).all()
    assert np.abs(d_t[~filler]).sum() == E       # every transmitted LLR landed somewhere, repeats added
    cwl = O.d_to_cw_llr(d_t, Z)
    assert (cwl[:2 * Z] == 0).all() and np.isinf(cwl[2 * Z:][filler]).all()


def test_qpsk_map_and_llr(O):
    bits = np.array([0, 0, 0, 1, 1, 0, 1, 1], np.uint8)
    re, im = O.qpsk_mod(bits)
    a = np.float32(1 / np.sqrt(2))
    np.testing.assert_allclose(re, [a, a, -a, -a]); np.testing.assert_allclose(im, [a, -a, a, -a])
    llr = O.qpsk_demod(re, im, 0.5)
    np.testing.assert_allclose(llr, (1 - 2.0 * bits) * 2 * np.sqrt(2) * a / 0.5, rtol=1e-6)


def test_f16_oracle_rounding_matches_numpy_float16(O):
    """Oracle A16's own binary16 rounding / widening against numpy.float16 (IEEE RNE), incl. subnormals,
    ties, overflow to inf and every finite bit pattern."""
    rng = np.random.default_rng(5)
    allh = np.arange(0, 0x10000, dtype=np.uint16)
    finite = allh[(allh & 0x7c00) != 0x7c00]
    wide = O.f16_widen(finite)
    assert (wide == finite.view(np.float16).astype(np.float64)).all()
    assert (O.f16_round(wide) == finite).all()
    x = np.concatenate([rng.normal(0, s, 20000) for s in (1e-7, 1e-4, 1.0, 300.0, 3e4)])
    # exact ties between neighbouring binary16 values, sums and 0.75 products of binary16 values
    a = finite[rng.integers(0, finite.size, 50000)].view(np.float16).astype(np.float64)
    b = finite[rng.integers(0, finite.size, 50000)].view(np.float16).astype(np.float64)
    x = np.concatenate([x, (a + np.nextafter(a.astype(np.float16), np.float16(np.inf)).astype(np.float64)) / 2, a + b, a - b, 0.75 * a,
                        [65504.0, 65519.9, 65520.0, 1e9, -65520.0, 2.0 ** -25, 2.0 ** -25 * 1.0001, 2.0 ** -24, 0.0, -0.0]])
    with np.errstate(over="ignore"):
        ref = x.astype(np.float16).view(np.uint16)
    assert (O.f16_round(x) == ref).all()


def test_f16_oracle_decodes_and_tracks_f32(O):
    """Oracle A16 recovers the information bits at moderate SNR and agrees with oracle A on almost every block."""
    from conftest import make_llr
    rng = np.random.default_rng(21)
    for bg, Z, E, esn0 in ((1, 48, 2800, 0.5), (2, 52, 2000, -1.0), (2, 6, 100, 3.5)):
        info, llr = make_llr(O, bg, Z, 48, E, esn0, rng)
        a = O.decode_nms(bg, Z, llr, 8, early_term=True)
        h = O.decode_nms(bg, Z, llr, 8, early_term=True, f16=True)
        assert np.isfinite(h["app"]).all() and np.abs(h["app"]).max() < 65504
        ok32 = (a["hard"] == info).all(axis=1)
        ok16 = (h["hard"] == info).all(axis=1)
        assert ok16.mean() > 0.8 and abs(int(ok16.sum()) - int(ok32.sum())) <= 3, (bg, Z, ok16.sum(), ok32.sum())
    # saturating inputs (+-inf, NaN filler, huge magnitudes) stay finite and decode
    d = O.dims(2, 52)
    info, llr = make_llr(O, 2, 52, 8, 2000, 8.0, rng)
    llr = llr * 1e4
    llr[:, d["K"] - 104:d["K"]] = np.inf
    h = O.decode_nms(2, 52, llr, 8, f16=True)
    assert np.isfinite(h["app"]).all() and np.abs(h["app"]).max() <= 2048 + 30 * 1536


def test_constellations_match_reference_mapping_vectors(O):
    """Oracle modulator (TS 38.211 formulas) against tests/golden/constellations.json, which holds the points the
    reference's own CustomSymbolMapping vectors define (NRModulator.m:73-81; generated by tools/make_golden_mod.py)."""
    import json
    from pathlib import Path
    gold = json.loads((Path(__file__).parent / "golden" / "constellations.json").read_text())
    assert sorted(gold) == ["16QAM", "256QAM", "64QAM", "BPSK", "QPSK"]
    for name, g in gold.items():
        Qm, pts = g["Q_m"], np.array(g["points"])
        M = 1 << Qm
        bits = ((np.arange(M)[:, None] >> np.arange(Qm - 1, -1, -1)) & 1).astype(np.uint8)
        sym = O.modulate(bits, Qm)
        assert np.allclose(sym.real, pts[:, 0], atol=1e-7) and np.allclose(sym.imag, pts[:, 1], atol=1e-7), name
        assert abs(np.mean(np.abs(sym) ** 2) - 1.0) < 1e-6, name      # 'Average power' normalisation
    with pytest.raises(ValueError):
        O.modulate(np.zeros(6, np.uint8), 3)


def test_demodulator_oracle_properties(O):
    """Exact LLR: closed forms for BPSK / QPSK, sign = hard decision, max-log agrees at high SNR, noiseless
    symbols decode to their own bits for every modulation."""
    rng = np.random.default_rng(11)
    for Qm in (1, 2, 4, 6, 8):
        bits = rng.integers(0, 2, 600 * Qm, dtype=np.uint8)
        tx = O.modulate(bits, Qm)
        hard = O.demodulate(tx, Qm, 0.1, "Hard decision")
        assert (hard == bits).all()
        var = 0.05
        rx = tx + rng.normal(0, np.sqrt(var / 2), tx.size) + 1j * rng.normal(0, np.sqrt(var / 2), tx.size)
        rx = rx.astype(np.complex64)
        llr = O.demodulate(rx, Qm, var)
        apx = O.demodulate(rx, Qm, var, "Approximate log-likelihood ratio")
        hd = O.demodulate(rx, Qm, var, "Hard decision")
        assert ((apx < 0) == (hd == 1)).all()          # max-log sign = nearest-point decision
        assert ((llr < 0) == (hd == 1)).mean() > 0.99
        assert np.abs(llr - apx).max() < np.log(1 << Qm) + 1e-9
        if Qm == 1:
            assert np.allclose(llr, 2 * np.sqrt(2) * (rx.real + rx.imag) / var, rtol=1e-6, atol=1e-6)
        if Qm == 2:
            ref = np.stack([rx.real, rx.imag], axis=1).ravel() * 2 * np.sqrt(2) / var
            assert np.allclose(llr, ref, rtol=1e-6, atol=1e-6)
            assert np.allclose(llr, O.qpsk_demod(rx.real.astype(np.float32), rx.imag.astype(np.float32), var), rtol=1e-5, atol=1e-5)


def test_oracle_b_variants_agree(O):
    """Oracle B: the float32-input export, the float64 export and the _ex export (termination selectable, APP returned)
    are the same decoder; with termination off it runs exactly max_iters flooding iterations; a noiseless word is returned."""
    rng = np.random.default_rng(12)
    info, llr = make_llr(O, 2, 52, 6, 2000, -2.0, rng, filler=104)
    a = O.decode_bp(2, 52, llr, 8)                                    # orc_decode_bp_f32
    b = O.decode_bp(2, 52, llr.astype(np.float64), 8, want_app=True)  # orc_decode_bp_ex
    assert (a["hard"] == b["hard"]).all() and (a["iters"] == b["iters"]).all() and (a["parity_ok"] == b["parity_ok"]).all()
    assert np.isinf(b["app"][:, 416:520]).all()                       # +inf filler stays +inf through every iteration
    assert ((b["app"][:, :520] < 0) == b["hard"].astype(bool)).all()
    c = O.decode_bp(2, 52, llr, 8, early_term=False)
    assert (c["iters"] == 8).all()
    done = a["parity_ok"] == 1
    assert done.any() and (c["hard"][done] == a["hard"][done]).all()  # a converged word stays converged (fixed point of hard)
    clean = np.where(O.encode(2, 52, info) == 0, 20.0, -20.0).astype(np.float32)
    clean[:, :104] = 0
    d = O.decode_bp(2, 52, clean, 8)
    assert (d["hard"] == info).all() and (d["iters"] <= 2).all()


def test_oracle_b_golden_vectors(O, golden_decode, golden_decode_bp):
    """Oracle B on the committed LLRs reproduces the committed decisions / iteration counts / parity flags
    (tools/make_golden_bp.py), with the reference's termination rule and with the iteration count fixed."""
    for n in sorted({k.split("__")[0] for k in golden_decode_bp.files}):
        bg, Z, iters = golden_decode_bp[n + "__cfg"].tolist()
        for tag, early in (("stop", True), ("full", False)):
            r = O.decode_bp(bg, Z, golden_decode[n + "__llr"], iters, early_term=early)
            assert (np.packbits(r["hard"], axis=1) == golden_decode_bp[f"{n}__{tag}__hard"]).all(), (n, tag)
            assert (r["iters"] == golden_decode_bp[f"{n}__{tag}__iters"]).all(), (n, tag)
            assert (r["parity_ok"] == golden_decode_bp[f"{n}__{tag}__ok"]).all(), (n, tag)


@pytest.mark.parametrize("bg,Z,n_rows", [(1, 384, 46), (2, 52, 33), (1, 7, 13)])
def test_core_checks_hold_extension_check_fails(O, bg, Z, n_rows):
    """The scenario the GPU's two-stage parity-check stop must get right, pinned on the oracle: a strongly wrong
    degree-1 extension parity bit leaves every core check satisfied and one extension check failing for ever."""
    from conftest import make_core_pass_llr
    rng = np.random.default_rng(Z + n_rows)
    info, llr, kind = make_core_pass_llr(O, bg, Z, 9, n_rows, rng)
    for f16 in (False, True):
        if f16:
            llr = np.clip(llr, -2048, 2048)
        ref = O.decode_nms(bg, Z, llr, 3, early_term=True, n_rows=n_rows, f16=f16)
        assert (ref["parity_ok"][kind == 0] == 1).all() and (ref["iters"][kind == 0] == 1).all()
        assert (ref["parity_ok"][kind == 1] == 0).all() and (ref["iters"][kind == 1] == 3).all()
        assert (ref["parity_ok"][kind == 2] == 1).all()
        assert (ref["hard"][kind != 1] == info[kind != 1]).all()
        # structural claim: on the final hard decisions of the kind-1 codewords no core check (rows 0..4Z-1 of H,
        # get_pcm.m:8) fails, and at least one active extension check does
        rows, cols = O.pcm(bg, Z)
        hard_all = (ref["app"] < 0).astype(np.uint8)
        syn = np.zeros((hard_all.shape[0], int(rows.max()) + 1), dtype=np.uint8)
        for b in range(hard_all.shape[0]):
            np.bitwise_xor.at(syn[b], rows, hard_all[b, cols])
        core_fail = syn[:, :4 * Z].any(axis=1)
        ext_fail = syn[:, 4 * Z:n_rows * Z].any(axis=1)
        assert not core_fail[kind == 1].any() and ext_fail[kind == 1].all()
        assert not core_fail[kind == 0].any() and not ext_fail[kind == 0].any()


def test_nms_revision_2_tracks_revision_1(O):
    """Oracle A revision 2 (degree-1 parity variables enter their check with the channel value) against
    the round-1 definition on identical noise: the two differ only by float rounding of app - c on those edges, so block
    decisions and iteration counts agree (stated tolerance: at most 1 block in 1000 differs in float32, 5 in binary16);
    both revisions also equal the independent numpy twin bit for bit."""
    from conftest import make_llr
    from oracle import twin as T
    rng = np.random.default_rng(21)
    try:
        for bg, Z, E, esn0, B, fill, rows in ((2, 6, 100, 2.0, 4000, 24, 13), (2, 52, 2000, -2.0, 1000, 104, 33), (1, 96, 6336, 0.0, 300, 0, 46)):
            info, llr = make_llr(O, bg, Z, B, E, esn0, rng, filler=fill)
            for f16, tol in ((False, 1e-3), (True, 5e-3)):
                err, its = {}, {}
                for rev in (1, 2):
                    O.set_nms_revision(rev)
                    r = O.decode_nms(bg, Z, llr, 8, early_term=True, n_rows=rows, f16=f16)
                    err[rev], its[rev] = (r["hard"] != info).any(axis=1), r["iters"]
                    dec = T.nms_layered_f16 if f16 else T.nms_layered
                    hard, app, it, ok = dec(bg, Z, llr[0], 8, True, rows, deg1_shortcut=(rev == 2))
                    assert (hard == r["hard"][0]).all() and it == r["iters"][0]
                    assert (app.astype(np.float32).view(np.uint32) == r["app"][0].view(np.uint32)).all()
                assert (err[1] != err[2]).mean() <= tol, (bg, Z, f16)
                assert abs(its[1].mean() - its[2].mean()) < 0.01
    finally:
        O.set_nms_revision(2)
