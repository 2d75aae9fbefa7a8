"""ctypes front-end of the CPU oracle (oracle/nrldpc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this module; the product package never does.  Decoder parity is UNPINNED against the
reference (closed-source comm.LDPCDecoder, no golden vectors) -- see the C file's header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "libnrldpc_oracle.so"
    src = _HERE / "nrldpc_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.orc_pcm.restype = C.c_long
        _LIB.orc_syndrome_weight.restype = C.c_long
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Params(C.Structure):
    """Mirror of orc_params_t: every NRLDPC.m Dependent getter (NRLDPC.m:297-543)."""
    _fields_ = [(n, C.c_int) for n in (
        "BG", "A", "G", "Q_m", "N_L", "rv_id", "I_LBRM", "TBS_LBRM",
        "tb_L", "B", "K_cb", "C", "cb_L", "B_prime", "K_prime", "K_b", "Z_c", "i_LS", "K", "N",
        "N_ref", "N_cb", "k_0")] + [("E_r", C.c_int * 64), ("status", C.c_int)]


def params(BG, A, G, Q_m=2, N_L=1, rv_id=0, I_LBRM=0, TBS_LBRM=0):
    p = Params(BG=BG, A=A, G=G, Q_m=Q_m, N_L=N_L, rv_id=rv_id, I_LBRM=I_LBRM, TBS_LBRM=TBS_LBRM)
    rc = lib().orc_params(C.byref(p))
    return None if rc else p


def set_index(Z):
    return lib().orc_set_index(int(Z))


def lifting_size(K_b, K_prime):
    return lib().orc_lifting_size(int(K_b), int(K_prime))


def dims(bg, Z):
    rows, cols, kc = (46, 68, 22) if bg == 1 else (42, 52, 10)
    return dict(rows=rows, cols=cols, kcols=kc, K=kc * Z, N=(cols - 2) * Z, ncw=cols * Z,
                edges=316 if bg == 1 else 197)


def table(bg):
    out = np.zeros((316 if bg == 1 else 197, 10), dtype=np.int32)
    n = lib().orc_table(bg, _p(out, C.c_int))
    assert n == out.shape[0]
    return out


def pcm(bg, Z):
    """Ones of the lifted H (get_pcm.m:8) as (rows, cols) index arrays."""
    n = dims(bg, Z)["edges"] * Z
    r = np.zeros(n, dtype=np.int32)
    c = np.zeros(n, dtype=np.int32)
    got = lib().orc_pcm(bg, Z, _p(r, C.c_int), _p(c, C.c_int))
    assert got == n
    return r, c


def encode(bg, Z, info, method="qc"):
    info = np.ascontiguousarray(info, dtype=np.uint8)
    single = info.ndim == 1
    info2 = info.reshape(-1, info.shape[-1])
    d = dims(bg, Z)
    assert info2.shape[1] == d["K"]
    cw = np.zeros((info2.shape[0], d["ncw"]), dtype=np.uint8)
    fn = lib().orc_encode_qc if method == "qc" else lib().orc_encode_gf2
    for b in range(info2.shape[0]):
        rc = fn(bg, Z, _p(info2[b], C.c_uint8), _p(cw[b], C.c_uint8))
        if rc:
            raise RuntimeError(f"oracle encode rc={rc}")
    return cw[0] if single else cw


def syndrome_weight(bg, Z, cw, n_rows=0):
    cw = np.ascontiguousarray(cw, dtype=np.uint8)
    return lib().orc_syndrome_weight(bg, Z, n_rows, _p(cw, C.c_uint8))


def decode_nms(bg, Z, llr, max_iters=8, early_term=False, alpha=0.75, n_rows=0, want_app=True,
               n_threads=None, f16=False):
    """Oracle A (layered NMS, f32; f16=True: oracle A16, binary16 arithmetic).
    llr: [batch, cols*Z] float32 in cw_tilde layout."""
    llr = np.ascontiguousarray(llr, dtype=np.float32)
    d = dims(bg, Z)
    llr2 = llr.reshape(-1, d["ncw"])
    B = llr2.shape[0]
    hard = np.zeros((B, d["K"]), dtype=np.uint8)
    app = np.zeros((B, d["ncw"]), dtype=np.float32) if want_app else None
    iters = np.zeros(B, dtype=np.int32)
    ok = np.zeros(B, dtype=np.uint8)
    nt = n_threads or os.cpu_count() or 1
    fn = lib().orc_decode_nms_f16 if f16 else lib().orc_decode_nms
    rc = fn(bg, Z, n_rows, max_iters, int(early_term), C.c_float(alpha),
                              _p(llr2, C.c_float), C.c_long(B), _p(hard, C.c_uint8),
                              _p(app, C.c_float) if want_app else None, _p(iters, C.c_int32),
                              _p(ok, C.c_uint8), nt)
    if rc:
        raise RuntimeError(f"oracle decode_nms rc={rc}")
    return dict(hard=hard, app=app, iters=iters, parity_ok=ok)


def set_nms_revision(rev: int):
    """Oracle A / A16 revision: 2 (default) = degree-1 variables enter their check with the channel value; 1 = the round-1
    definition.  See the header of orc_decode_nms in nrldpc_oracle.c."""
    lib().orc_set_nms_revision(int(rev))


def f16_round(x):
    """binary16 bit patterns of float64 values, as oracle A16 rounds them (round to nearest even)."""
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    out = np.zeros(x.size, np.uint16)
    lib().orc_f16_round(_p(x, C.c_double), C.c_long(x.size), _p(out, C.c_uint16))
    return out


def f16_widen(h):
    h = np.ascontiguousarray(h, dtype=np.uint16).ravel()
    out = np.zeros(h.size, np.float64)
    lib().orc_f16_widen(_p(h, C.c_uint16), C.c_long(h.size), _p(out, C.c_double))
    return out


def decode_bp(bg, Z, llr, max_iters=8, n_rows=0, n_threads=None, early_term=True, want_app=False):
    """Oracle B (flooding sum-product, f64; early_term=True is the reference's parity-check termination).
    float32 input is widened exactly to float64, as the GPU boundary does."""
    d = dims(bg, Z)
    nt = n_threads or os.cpu_count() or 1
    llr = np.ascontiguousarray(llr)
    llr2 = llr.reshape(-1, d["ncw"])
    B = llr2.shape[0]
    hard = np.zeros((B, d["K"]), dtype=np.uint8)
    iters = np.zeros(B, dtype=np.int32)
    ok = np.zeros(B, dtype=np.uint8)
    app = None
    if llr2.dtype == np.float32 and early_term and not want_app:
        rc = lib().orc_decode_bp_f32(bg, Z, n_rows, max_iters, _p(llr2, C.c_float), C.c_long(B),
                                     _p(hard, C.c_uint8), _p(iters, C.c_int32), _p(ok, C.c_uint8), nt)
    else:
        llr2 = np.ascontiguousarray(llr2, dtype=np.float64)
        app = np.zeros((B, d["ncw"]), dtype=np.float64) if want_app else None
        rc = lib().orc_decode_bp_ex(bg, Z, n_rows, max_iters, int(bool(early_term)), _p(llr2, C.c_double), C.c_long(B),
                                    _p(hard, C.c_uint8), _p(app, C.c_double) if want_app else None,
                                    _p(iters, C.c_int32), _p(ok, C.c_uint8), nt)
    if rc:
        raise RuntimeError(f"oracle decode_bp rc={rc}")
    return dict(hard=hard, app=app, iters=iters, parity_ok=ok)


FILL = 0xFF


def cw_to_d(Z, K, K_prime, N, cw):
    cw = np.ascontiguousarray(cw, dtype=np.uint8)
    d = np.zeros(N, dtype=np.uint8)
    lib().orc_cw_to_d(Z, K, K_prime, N, _p(cw, C.c_uint8), _p(d, C.c_uint8))
    return d


def bit_selection_tx(d, N_cb, k_0, E):
    d = np.ascontiguousarray(d, dtype=np.uint8)
    e = np.zeros(E, dtype=np.uint8)
    lib().orc_bit_selection_tx(_p(d, C.c_uint8), N_cb, k_0, E, _p(e, C.c_uint8))
    return e


def interleave_tx(e, Q_m):
    e = np.ascontiguousarray(e, dtype=np.uint8)
    f = np.zeros_like(e)
    lib().orc_interleave_tx(_p(e, C.c_uint8), len(e), Q_m, _p(f, C.c_uint8))
    return f


def deinterleave_rx(f, Q_m):
    f = np.ascontiguousarray(f, dtype=np.float32)
    e = np.zeros_like(f)
    lib().orc_deinterleave_rx(_p(f, C.c_float), len(f), Q_m, _p(e, C.c_float))
    return e


def bit_selection_rx(e, N, N_cb, k_0, Z, K, K_prime, harq_buf=None):
    e = np.ascontiguousarray(e, dtype=np.float32)
    d = np.zeros(N, dtype=np.float32)
    hb = None
    if harq_buf is not None:
        assert harq_buf.dtype == np.float32 and harq_buf.flags.c_contiguous and len(harq_buf) == N_cb
        hb = _p(harq_buf, C.c_float)
    lib().orc_bit_selection_rx(_p(e, C.c_float), len(e), N, N_cb, k_0, Z, K, K_prime, hb, _p(d, C.c_float))
    return d


def d_to_cw_llr(d, Z):
    d = np.ascontiguousarray(d, dtype=np.float32)
    out = np.zeros(len(d) + 2 * Z, dtype=np.float32)
    lib().orc_d_to_cw_llr(_p(d, C.c_float), len(d), Z, _p(out, C.c_float))
    return out


def crc(kind, bits):
    """kind: 'CRC16' | 'CRC24A' | 'CRC24B' (get_3gpp_crc_polynomial.m:3-17). Returns parity bits."""
    k = {"CRC16": 0, "CRC24A": 1, "CRC24B": 2}[kind]
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    par = np.zeros(24, dtype=np.uint8)
    L = lib().orc_crc(k, _p(bits, C.c_uint8), len(bits), _p(par, C.c_uint8))
    return par[:L].copy()


def qpsk_mod(bits):
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    n = len(bits) // 2
    re = np.zeros(n, dtype=np.float32)
    im = np.zeros(n, dtype=np.float32)
    lib().orc_qpsk_mod(_p(bits, C.c_uint8), C.c_long(n), _p(re, C.c_float), _p(im, C.c_float))
    return re, im


def qpsk_demod(re, im, variance):
    re = np.ascontiguousarray(re, dtype=np.float32)
    im = np.ascontiguousarray(im, dtype=np.float32)
    llr = np.zeros(2 * len(re), dtype=np.float32)
    lib().orc_qpsk_demod(_p(re, C.c_float), _p(im, C.c_float), C.c_long(len(re)), C.c_float(variance), _p(llr, C.c_float))
    return llr


def modulate(bits, Q_m):
    """NRModulator.step for Q_m in {1,2,4,6,8}: complex64 symbols (TS 38.211 section 5.1 maps)."""
    bits = np.ascontiguousarray(bits, dtype=np.uint8).ravel()
    n = bits.size // Q_m
    re = np.zeros(n, np.float32); im = np.zeros(n, np.float32)
    if lib().orc_modulate(_p(bits, C.c_uint8), C.c_long(n), int(Q_m), _p(re, C.c_float), _p(im, C.c_float)):
        raise ValueError("Unsupported modulation")
    return re + 1j * im


METHODS = {"Log-likelihood ratio": 0, "Approximate log-likelihood ratio": 1, "Hard decision": 2}


def demodulate(sym, Q_m, variance, method="Log-likelihood ratio"):
    """NRDemodulator.step: float64, the literal two-dimensional definition (see nrldpc_oracle.c)."""
    sym = np.asarray(sym).ravel()
    re = np.ascontiguousarray(sym.real, dtype=np.float32); im = np.ascontiguousarray(sym.imag, dtype=np.float32)
    out = np.zeros(re.size * Q_m, np.float64)
    if lib().orc_demodulate(_p(re, C.c_float), _p(im, C.c_float), C.c_long(re.size), int(Q_m), C.c_double(variance),
                            METHODS[method], _p(out, C.c_double)):
        raise ValueError("Unsupported modulation / method / variance")
    return out
