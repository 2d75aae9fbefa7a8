"""C oracle (oracle/nrldpc_oracle.c) == independent numpy twin (oracle/twin.py), bit for bit.

The reference ships no decoder vectors and its decoder is a closed toolbox (NRLDPCDecoder.m:120,265), so the C oracle
would otherwise be the single arbiter of every GPU parity test.  The twin shares no code, table or data structure
with it (own parse of get_3gpp_base_graph.m, H in matrix form per get_pcm.m:7-9, decoders on H's non-zeros).
"""
import hashlib

import numpy as np
import pytest

from oracle import twin as T

SHA = {1: "4f7508858e04dc7bbed58f8b13c39b4bdb67835790d7e3b7f19152f2e59578d2",
       2: "a058c8507148dca1641ca1fc370f98739bca27532651119f9ae2212c7718bc92"}
# one lifting size per set index i_LS = 0..7 (the smallest that keeps the run short) plus a few second members
SMALL_Z = [2, 3, 5, 7, 9, 11, 13, 15, 4, 6, 10, 16]


def _noisy(O, bg, Z, B, rng, esn0=1.0, filler=0, E=None):
    d = O.dims(bg, Z)
    info = rng.integers(0, 2, (B, d["K"]), dtype=np.uint8)
    if filler:
        info[:, d["K"] - filler:] = 0
    cw = O.encode(bg, Z, info)
    s2 = 10 ** (-esn0 / 10)
    y = (1 - 2.0 * cw) / np.sqrt(2) + rng.normal(0, np.sqrt(s2 / 2), cw.shape)
    llr = (2 * np.sqrt(2) * y / s2).astype(np.float32)
    llr[:, :2 * Z] = 0
    if E is not None:
        llr[:, 2 * Z + E:] = 0
    if filler:
        llr[:, d["K"] - filler:d["K"]] = np.inf
    return info, cw, llr


def test_twin_tables_are_the_reference_tables(O):
    """Three independent copies agree: the twin's parse (fixture or live reference), the C oracle's own header, and the
    sha256 of SURVEY Appendix A.2.  When the reference checkout is mounted the committed fixture must equal its parse."""
    fix = np.load(T.FIXTURE)
    for bg in (1, 2):
        t = T.tables()[bg]
        text = "\n".join(" ".join(str(int(v)) for v in row) for row in t)
        assert hashlib.sha256(text.encode()).hexdigest() == SHA[bg]
        assert (O.table(bg) == t).all()
        assert (fix["bg%d" % bg] == t).all()


@pytest.mark.parametrize("bg", [1, 2])
def test_twin_pcm_matches_oracle(O, bg):
    for Z in (2, 3, 7, 16, 52):
        g = T.graph(bg, Z)
        r, c = O.pcm(bg, Z)
        order = np.lexsort((c, r))
        assert (g.check_var == c[order]).all()
        assert (np.repeat(np.arange(g.M), g.check_deg) == r[order]).all()
    # the sparse construction used for large Z equals the dense one
    Hs = T.pcm(bg, 208)
    r, c = O.pcm(bg, 208)
    assert Hs.nnz == len(r) and (np.asarray(Hs[r, c]).ravel() == 1).all()


@pytest.mark.parametrize("bg", [1, 2])
def test_twin_encoder_matches_oracle(O, bg):
    rng = np.random.default_rng(11)
    for Z in (2, 3, 5, 7):
        d = O.dims(bg, Z)
        info = rng.integers(0, 2, d["K"], dtype=np.uint8)
        assert (T.encode(bg, Z, info) == O.encode(bg, Z, info[None])[0]).all()


@pytest.mark.parametrize("bg", [1, 2])
@pytest.mark.parametrize("f16", [False, True])
def test_twin_nms_small_Z_every_set(O, bg, f16):
    rng = np.random.default_rng(100 + bg)
    dec = T.nms_layered_f16 if f16 else T.nms_layered
    for Z in SMALL_Z:
        d = O.dims(bg, Z)
        for early, n_rows, filler in ((False, 0, 0), (True, 0, Z), (True, 7, 0)):
            E = None if not n_rows else (d["kcols"] - 2 + n_rows) * Z
            _, _, llr = _noisy(O, bg, Z, 3, rng, esn0=rng.uniform(-1.0, 4.0), filler=filler, E=E)
            llr[0, rng.integers(0, llr.shape[1], 3)] = [np.nan, -np.inf, -0.0]      # filler marker, saturation, signed zero
            ref = O.decode_nms(bg, Z, llr, 6, early_term=early, n_rows=n_rows, f16=f16)
            for b in range(llr.shape[0]):
                hard, app, it, ok = dec(bg, Z, llr[b], 6, early, n_rows)
                assert (hard == ref["hard"][b]).all(), (bg, Z, b)
                assert it == ref["iters"][b] and bool(ok) == bool(ref["parity_ok"][b]), (bg, Z, b)
                assert (app.astype(np.float32).view(np.uint32) == ref["app"][b].view(np.uint32)).all(), (bg, Z, b)


def test_twin_nms_z384(O):
    """Four codewords of the headline code near its waterfall (BASELINE config 2)."""
    rng = np.random.default_rng(5)
    _, _, llr = _noisy(O, 1, 384, 4, rng, esn0=-0.3, E=25272)
    ref = O.decode_nms(1, 384, llr, 8, early_term=True)
    for b in range(4):
        hard, app, it, ok = T.nms_layered(1, 384, llr[b], 8, True)
        assert (hard == ref["hard"][b]).all() and it == ref["iters"][b] and bool(ok) == bool(ref["parity_ok"][b])
        assert (app.view(np.uint32) == ref["app"][b].view(np.uint32)).all()
    ref16 = O.decode_nms(1, 384, llr[:2], 8, early_term=False, f16=True)
    for b in range(2):
        hard, app, it, ok = T.nms_layered_f16(1, 384, llr[b], 8, False)
        assert (hard == ref16["hard"][b]).all()
        assert (app.view(np.uint32) == ref16["app"][b].view(np.uint32)).all()


@pytest.mark.parametrize("bg", [1, 2])
def test_twin_bp_small_Z_every_set(O, bg):
    """Oracle B (the restatement of comm.LDPCDecoder's documented algorithm) against the twin: decisions, iteration
    counts, parity flags identical; a-posteriori values identical as float64 bit patterns (same libm, same order)."""
    rng = np.random.default_rng(200 + bg)
    for Z in SMALL_Z[:8] + [4]:
        d = O.dims(bg, Z)
        for early, filler in ((True, 0), (False, Z)):
            _, _, llr = _noisy(O, bg, Z, 2, rng, esn0=rng.uniform(0.0, 4.0), filler=filler)
            llr64 = llr.astype(np.float64)
            ref = O.decode_bp(bg, Z, llr64, 6, early_term=early, want_app=True)
            for b in range(2):
                hard, Q, it, ok = T.bp_flooding(bg, Z, llr64[b], 6, early)
                assert (hard == ref["hard"][b]).all() and it == ref["iters"][b] and bool(ok) == bool(ref["parity_ok"][b]), (bg, Z, b)
                assert np.array_equal(Q.view(np.uint64), ref["app"][b].view(np.uint64)), (bg, Z, b)


def test_twin_bp_z384(O):
    rng = np.random.default_rng(6)
    _, _, llr = _noisy(O, 1, 384, 1, rng, esn0=0.4, E=25272)
    llr64 = llr.astype(np.float64)
    ref = O.decode_bp(1, 384, llr64, 4, early_term=True, want_app=True)
    hard, Q, it, ok = T.bp_flooding(1, 384, llr64[0], 4, True)
    assert (hard == ref["hard"][0]).all() and it == ref["iters"][0] and bool(ok) == bool(ref["parity_ok"][0])
    assert np.array_equal(Q.view(np.uint64), ref["app"][0].view(np.uint64))
