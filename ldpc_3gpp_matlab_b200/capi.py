"""ctypes binding of the C ABI in include/nrldpc_b200.h (libnrldpc_b200.so).

This is the only bridge between the Python host mirror of the reference's System objects and
the sm_100a kernels.  There is no CPU fallback: if the shared library is missing the import of
:func:`load` raises, and if no B200 is present ``nrldpc_create`` fails with NRLDPC_ECUDA.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libnrldpc_b200.so"
if os.environ.get("NRLDPC_B200_LIB"):   # kernel experiments only: another build of the same library (tools/gpu_variants.sh)
    LIB_PATH = Path(os.environ["NRLDPC_B200_LIB"])

NRLDPC_OK = 0
NRLDPC_EUNSUPPORTED = -1
NRLDPC_ESHAPE = -2
NRLDPC_ECUDA = -3
NRLDPC_ENOMEM = -4
MEM_HOST = 0
MEM_DEVICE = 1
LLR_MAX = 1048576.0
F32 = 0
F16X2 = 1
ALG_NMS = 0   # layered normalized min-sum (default)
ALG_BP = 1    # the reference's flooding sum-product in float64 (comm.LDPCDecoder, NRLDPCDecoder.m:120)
DEMOD_LLR, DEMOD_APPROX, DEMOD_HARD = 0, 1, 2
CRC16, CRC24A, CRC24B = 0, 1, 2
CRC_KIND = {"CRC16": CRC16, "CRC24A": CRC24A, "CRC24B": CRC24B}
LLR_MAX_F16 = 2048.0

# every symbol include/nrldpc_b200.h declares (tests/test_abi.py checks the header against this)
SYMBOLS = (
    "nrldpc_create", "nrldpc_destroy", "nrldpc_last_error", "nrldpc_synchronize", "nrldpc_get_dims",
    "nrldpc_set_index", "nrldpc_lifting_size", "nrldpc_base_graph", "nrldpc_decode", "nrldpc_encode",
    "nrldpc_rate_match", "nrldpc_rate_recover", "nrldpc_qpsk_awgn_llr", "nrldpc_host_alloc",
    "nrldpc_host_free", "nrldpc_launch_count", "nrldpc_version",
    "nrldpc_modulate", "nrldpc_awgn", "nrldpc_demodulate", "nrldpc_mod_awgn_llr", "nrldpc_crc", "nrldpc_decode16", "nrldpc_decode64", "nrldpc_decode8",
    "nrldpc_qpsk_awgn_rate_recover", "nrldpc_bler_count", "nrldpc_random_bits",
)


class UnsupportedParameters(ValueError):
    """Python face of error('ldpc_3gpp_matlab:UnsupportedParameters', ...) (e.g. NRLDPC.m:242)."""
    identifier = "ldpc_3gpp_matlab:UnsupportedParameters"


class NRLDPCError(RuntimeError):
    """Python face of error('ldpc_3gpp_matlab:Error', ...) (e.g. NRLDPCDecoder.m:149)."""
    identifier = "ldpc_3gpp_matlab:Error"


class CudaError(RuntimeError):
    """CUDA failure or missing device/extension; never silently replaced by a CPU path."""


class Cfg(C.Structure):
    _fields_ = [("bg", C.c_int32), ("Z", C.c_int32), ("max_iters", C.c_int32), ("early_term", C.c_int32),
                ("alpha", C.c_float), ("device", C.c_int32), ("llr_dtype", C.c_int32), ("algorithm", C.c_int32)]


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("bg", "Z", "i_LS", "rows", "cols", "kcols", "edges", "K", "N", "n_cw")]


class Rm(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("E", "k_0", "N_cb", "K_prime", "Q_m")]


_lib = None


def load():
    """dlopen libnrldpc_b200.so and declare prototypes.  Raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise CudaError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    lib.nrldpc_create.argtypes = [C.POINTER(vp), C.POINTER(Cfg)]
    lib.nrldpc_create.restype = C.c_int
    lib.nrldpc_destroy.argtypes = [vp]
    lib.nrldpc_destroy.restype = None
    lib.nrldpc_last_error.argtypes = [vp]
    lib.nrldpc_last_error.restype = C.c_char_p
    lib.nrldpc_synchronize.argtypes = [vp]
    lib.nrldpc_get_dims.argtypes = [vp, C.POINTER(Dims)]
    lib.nrldpc_set_index.argtypes = [i32]
    lib.nrldpc_lifting_size.argtypes = [i32, i32]
    lib.nrldpc_base_graph.argtypes = [i32, i32, vp, vp, vp]
    lib.nrldpc_decode.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp, i32, vp]
    lib.nrldpc_decode16.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp, i32, vp]
    lib.nrldpc_decode8.argtypes = [vp, vp, C.c_float, i64, i32, vp, vp, vp, vp, i32, vp]
    lib.nrldpc_decode64.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp, i32, vp]
    lib.nrldpc_encode.argtypes = [vp, vp, i64, vp, i32, vp]
    lib.nrldpc_rate_match.argtypes = [vp, vp, i64, C.POINTER(Rm), vp, i32, vp]
    lib.nrldpc_rate_recover.argtypes = [vp, vp, i64, C.POINTER(Rm), vp, vp, i32, vp]
    lib.nrldpc_qpsk_awgn_llr.argtypes = [vp, vp, i64, i32, C.c_float, u64, u64, vp, vp]
    lib.nrldpc_qpsk_awgn_rate_recover.argtypes = [vp, vp, i64, C.POINTER(Rm), C.c_float, u64, u64, vp, vp, vp]
    lib.nrldpc_random_bits.argtypes = [vp, vp, i64, i32, i64, u64, u64, vp]
    lib.nrldpc_bler_count.argtypes = [vp, vp, vp, vp, i64, vp, i64, vp, vp, vp, i64, i32, i32, i32, vp, vp, i32, i32, vp]
    lib.nrldpc_modulate.argtypes = [vp, vp, i64, i32, vp, vp]
    lib.nrldpc_awgn.argtypes = [vp, vp, i64, C.c_float, u64, u64, vp]
    lib.nrldpc_demodulate.argtypes = [vp, vp, i64, i32, C.c_float, i32, vp, vp]
    lib.nrldpc_mod_awgn_llr.argtypes = [vp, vp, i64, i32, C.c_float, i32, u64, u64, vp, vp]
    lib.nrldpc_crc.argtypes = [vp, vp, i64, i32, i64, i32, vp, i64, vp, vp]
    lib.nrldpc_host_alloc.argtypes = [u64]
    lib.nrldpc_host_alloc.restype = vp
    lib.nrldpc_host_free.argtypes = [vp]
    lib.nrldpc_host_free.restype = None
    lib.nrldpc_launch_count.argtypes = [vp]
    lib.nrldpc_launch_count.restype = i64
    lib.nrldpc_version.restype = C.c_char_p
    _lib = lib
    return lib


def _raise(rc: int, msg: str):
    if rc == NRLDPC_EUNSUPPORTED:
        raise UnsupportedParameters(msg)
    if rc == NRLDPC_ESHAPE:
        raise NRLDPCError(msg)
    raise CudaError(f"nrldpc rc={rc}: {msg}")


def _ptr(x):
    """Raw address of a numpy array (host) / torch tensor (host or device) / int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        assert x.flags.c_contiguous
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        assert x.is_contiguous()
        return x.data_ptr()
    raise TypeError(type(x))


def set_index(Z: int) -> int:
    rc = load().nrldpc_set_index(int(Z))
    if rc < 0:
        raise UnsupportedParameters("Invalid lifting size.")
    return rc


def lifting_size(K_b: int, K_prime: int) -> int:
    rc = load().nrldpc_lifting_size(int(K_b), int(K_prime))
    if rc < 0:
        raise UnsupportedParameters("Invalid block length.")
    return rc


def base_graph(bg: int, i_LS: int):
    n = 316 if bg == 1 else 197
    r = np.zeros(n, np.int32); c = np.zeros(n, np.int32); s = np.zeros(n, np.int32)
    rc = load().nrldpc_base_graph(int(bg), int(i_LS), r.ctypes.data, c.ctypes.data, s.ctypes.data)
    if rc < 0:
        raise UnsupportedParameters("BG must be 1 or 2 and set_index must be between 0 and 7.")
    return r[:rc], c[:rc], s[:rc]


class Handle:
    """Owner of one nrldpc_t: a (BG, Z) code with its iteration policy on one GPU."""

    def __init__(self, bg: int, Z: int, max_iters: int = 8, early_term: bool = False, alpha: float = 0.75,
                 device: int = -1, llr_dtype: int = F32, algorithm: int = ALG_NMS):
        self._lib = load()
        self._h = C.c_void_p()
        cfg = Cfg(bg=int(bg), Z=int(Z), max_iters=int(max_iters), early_term=int(bool(early_term)),
                  alpha=float(alpha), device=int(device), llr_dtype=int(llr_dtype), algorithm=int(algorithm))
        rc = self._lib.nrldpc_create(C.byref(self._h), C.byref(cfg))
        if rc:
            _raise(rc, self._lib.nrldpc_last_error(None).decode())
        d = Dims()
        self._lib.nrldpc_get_dims(self._h, C.byref(d))
        self.dims = d
        self.bg, self.Z, self.K, self.N, self.n_cw, self.rows = d.bg, d.Z, d.K, d.N, d.n_cw, d.rows

    def close(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._h = None
            self._lib.nrldpc_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    def _check(self, rc):
        if rc:
            _raise(rc, self._lib.nrldpc_last_error(self._h).decode())

    @property
    def launches(self) -> int:
        return int(self._lib.nrldpc_launch_count(self._h))

    def synchronize(self):
        self._check(self._lib.nrldpc_synchronize(self._h))

    # raw pointer-level calls (host numpy arrays, or torch tensors on host/device) ---------------
    def decode_raw(self, llr, batch, hard, soft=None, iters=None, ok=None, n_rows=0, mem=MEM_HOST, stream=None):
        self._check(self._lib.nrldpc_decode(self._h, _ptr(llr), int(batch), int(n_rows), _ptr(hard), _ptr(soft),
                                            _ptr(iters), _ptr(ok), int(mem), stream))

    def decode8_raw(self, llr_q, scale, batch, hard, soft=None, iters=None, ok=None, n_rows=0, mem=MEM_HOST, stream=None):
        """nrldpc_decode8: LLRs as int8 (llr = scale * q, q = 127 = filler)."""
        self._check(self._lib.nrldpc_decode8(self._h, _ptr(llr_q), float(scale), int(batch), int(n_rows), _ptr(hard), _ptr(soft),
                                             _ptr(iters), _ptr(ok), mem, stream))

    def decode16_raw(self, llr_f16, batch, hard, soft=None, iters=None, ok=None, n_rows=0, mem=MEM_HOST, stream=None):
        """nrldpc_decode16: LLRs as IEEE binary16 (numpy float16 / torch.float16)."""
        self._check(self._lib.nrldpc_decode16(self._h, _ptr(llr_f16), int(batch), int(n_rows), _ptr(hard), _ptr(soft),
                                              _ptr(iters), _ptr(ok), int(mem), stream))

    def decode64_raw(self, llr_f64, batch, hard, soft=None, iters=None, ok=None, n_rows=0, mem=MEM_HOST, stream=None):
        """nrldpc_decode64: LLRs (and app_soft) as float64, the reference's own type (NRLDPCDecoder.m:262)."""
        self._check(self._lib.nrldpc_decode64(self._h, _ptr(llr_f64), int(batch), int(n_rows), _ptr(hard), _ptr(soft),
                                              _ptr(iters), _ptr(ok), int(mem), stream))

    def encode_raw(self, info, batch, cw, mem=MEM_HOST, stream=None):
        self._check(self._lib.nrldpc_encode(self._h, _ptr(info), int(batch), _ptr(cw), int(mem), stream))

    def rate_match_raw(self, cw, batch, rm: Rm, f, mem=MEM_HOST, stream=None):
        self._check(self._lib.nrldpc_rate_match(self._h, _ptr(cw), int(batch), C.byref(rm), _ptr(f), int(mem), stream))

    def rate_recover_raw(self, f, batch, rm: Rm, harq, llr_cw, mem=MEM_HOST, stream=None):
        self._check(self._lib.nrldpc_rate_recover(self._h, _ptr(f), int(batch), C.byref(rm), _ptr(harq),
                                                  _ptr(llr_cw), int(mem), stream))

    def qpsk_awgn_llr_raw(self, f_bits, batch, E, variance, seed, stream_id, f_llr, stream=None):
        self._check(self._lib.nrldpc_qpsk_awgn_llr(self._h, _ptr(f_bits), int(batch), int(E), float(variance),
                                                   int(seed), int(stream_id), _ptr(f_llr), stream))

    def qpsk_awgn_rate_recover_raw(self, f_bits, batch, rm: Rm, variance, seed, stream_id, harq, llr_cw, stream=None):
        self._check(self._lib.nrldpc_qpsk_awgn_rate_recover(self._h, _ptr(f_bits), int(batch), C.byref(rm), float(variance), int(seed),
                                                            int(stream_id), _ptr(harq), _ptr(llr_cw), stream))

    def bler_count_raw(self, hard, info, tb_hat, tb_hat_stride, tb, tb_stride, tb_ok, cb_passed, iters, n_tb, C_, K_prime, A, latch,
                       counters, do_latch=True, finalize=True, stream=None):
        self._check(self._lib.nrldpc_bler_count(self._h, _ptr(hard), _ptr(info), _ptr(tb_hat), int(tb_hat_stride), _ptr(tb), int(tb_stride),
                                                _ptr(tb_ok), _ptr(cb_passed), _ptr(iters), int(n_tb), int(C_), int(K_prime), int(A),
                                                _ptr(latch), _ptr(counters), int(bool(do_latch)), int(bool(finalize)), stream))

    def modulate_raw(self, bits, n_bits, Q_m, sym, stream=None):
        self._check(self._lib.nrldpc_modulate(self._h, _ptr(bits), int(n_bits), int(Q_m), _ptr(sym), stream))

    def awgn_raw(self, sym, n_sym, variance, seed, stream_id, stream=None):
        self._check(self._lib.nrldpc_awgn(self._h, _ptr(sym), int(n_sym), float(variance), int(seed), int(stream_id), stream))

    def demodulate_raw(self, sym, n_sym, Q_m, variance, method, llr, stream=None):
        self._check(self._lib.nrldpc_demodulate(self._h, _ptr(sym), int(n_sym), int(Q_m), float(variance), int(method),
                                                _ptr(llr), stream))

    def mod_awgn_llr_raw(self, bits, n_bits, Q_m, variance, method, seed, stream_id, llr, stream=None):
        self._check(self._lib.nrldpc_mod_awgn_llr(self._h, _ptr(bits), int(n_bits), int(Q_m), float(variance), int(method),
                                                  int(seed), int(stream_id), _ptr(llr), stream))

    def random_bits_raw(self, bits, rows, n_bits, stride, seed, stream_id, stream=None):
        """Uniform random bits (device memory), rows of n_bits at `stride`: round(rand(A,1)) of plot_BLER_vs_SNR.m:112."""
        self._check(self._lib.nrldpc_random_bits(self._h, _ptr(bits), int(rows), int(n_bits), int(stride), int(seed) & (2 ** 64 - 1),
                                                 int(stream_id) & (2 ** 64 - 1), stream))

    def crc_raw(self, bits, batch, n_bits, stride, kind, parity=None, parity_stride=0, ok=None, stream=None):
        self._check(self._lib.nrldpc_crc(self._h, _ptr(bits), int(batch), int(n_bits), int(stride), int(kind), _ptr(parity),
                                         int(parity_stride), _ptr(ok), stream))

    # numpy conveniences (host memory, synchronous) ----------------------------------------------
    def decode(self, llr, n_rows=0, want_soft=False):
        """float32 (default) or float64 LLRs (numpy float64 input goes through nrldpc_decode64)."""
        f64 = isinstance(llr, np.ndarray) and llr.dtype == np.float64
        llr = np.ascontiguousarray(llr, dtype=np.float64 if f64 else np.float32)
        if llr.shape[-1] != self.n_cw:
            raise NRLDPCError(f"llr should have {self.n_cw} entries per codeword (cw_tilde layout).")
        llr2 = llr.reshape(-1, self.n_cw)
        B = llr2.shape[0]
        hard = np.zeros((B, self.K), np.uint8)
        soft = np.zeros((B, self.n_cw), llr2.dtype) if want_soft else None
        iters = np.zeros(B, np.int32)
        ok = np.zeros(B, np.uint8)
        (self.decode64_raw if f64 else self.decode_raw)(llr2, B, hard, soft, iters, ok, n_rows=n_rows)
        return dict(hard=hard, app=soft, iters=iters, parity_ok=ok)

    def encode(self, info):
        info = np.ascontiguousarray(info, dtype=np.uint8)
        if info.shape[-1] != self.K:
            raise NRLDPCError(f"info should have K={self.K} entries per code block.")
        info2 = info.reshape(-1, self.K)
        cw = np.zeros((info2.shape[0], self.n_cw), np.uint8)
        self.encode_raw(info2, info2.shape[0], cw)
        return cw.reshape(info.shape[:-1] + (self.n_cw,))

    def rate_match(self, cw, E, k_0, N_cb, K_prime, Q_m):
        cw = np.ascontiguousarray(cw, dtype=np.uint8).reshape(-1, self.n_cw)
        f = np.zeros((cw.shape[0], E), np.uint8)
        self.rate_match_raw(cw, cw.shape[0], Rm(E, k_0, N_cb, K_prime, Q_m), f)
        return f

    def rate_recover(self, f, E, k_0, N_cb, K_prime, Q_m, harq=None):
        f = np.ascontiguousarray(f, dtype=np.float32).reshape(-1, E)
        out = np.zeros((f.shape[0], self.n_cw), np.float32)
        if harq is not None:
            assert harq.dtype == np.float32 and harq.flags.c_contiguous and harq.shape == (f.shape[0], self.N)
        self.rate_recover_raw(f, f.shape[0], Rm(E, k_0, N_cb, K_prime, Q_m), harq, out)
        return out
