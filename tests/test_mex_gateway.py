"""matlab/nrldpc_mex.cpp compiled against the stub MEX API (tests/stubs) and driven through mexFunction the way MATLAB
drives it at the reference's seam (obj.hLDPCDecoder, NRLDPCDecoder.m:117-121,265; obj.hLDPCEncoder,
NRLDPCEncoder.m:49,158): column-major (n_cw x batch) doubles in, logical (K x batch) out, handle as a uint64 scalar,
errors as MATLAB identifiers.  CPU tests cover loading and every error path that needs no device; the GPU tests decode
and encode through the gateway and compare with the oracle.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
SO = ROOT / "tests" / "stubs" / "libnrldpc_mex_test.so"
ID_UNSUPPORTED = "ldpc_3gpp_matlab:UnsupportedParameters"     # NRLDPC.m:242, caught at plot_BLER_vs_SNR.m:172-176
ID_ERROR = "ldpc_3gpp_matlab:Error"                            # NRLDPCDecoder.m:149
LOGICAL, DOUBLE, UINT64 = 3, 6, 15


class MexError(Exception):
    def __init__(self, ident, msg):
        super().__init__(f"{ident}: {msg}")
        self.identifier, self.message = ident, msg


class Mex:
    """mexFunction(nlhs, plhs, nrhs, prhs) with numpy values marshalled as mxArrays (column-major, like MATLAB)."""

    def __init__(self):
        subprocess.run(["make", "-C", str(SO.parent), "-s"], check=True)
        L = C.CDLL(str(SO))
        vp = C.c_void_p
        for name, res, args in (("shim_string", vp, [C.c_char_p]), ("shim_double", vp, [vp, C.c_size_t, C.c_size_t]),
                                ("shim_logical", vp, [vp, C.c_size_t, C.c_size_t]), ("shim_uint8", vp, [vp, C.c_size_t, C.c_size_t]),
                                ("shim_single", vp, [vp, C.c_size_t, C.c_size_t]), ("shim_int8", vp, [vp, C.c_size_t, C.c_size_t]),
                                ("shim_class", C.c_int, [vp]), ("mxGetM", C.c_size_t, [vp]), ("mxGetN", C.c_size_t, [vp]),
                                ("mxGetData", vp, [vp]), ("mxDestroyArray", None, [vp]),
                                ("shim_mex", C.c_int, [C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_char_p, C.c_char_p, C.c_size_t])):
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        self.L = L

    def _to_mx(self, v):
        L = self.L
        if isinstance(v, str):
            return L.shim_string(v.encode())
        if isinstance(v, Handle):
            return v.mx                          # handles travel as the uint64 array the gateway returned
        a = np.asarray(v)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        if a.ndim == 1:
            a = a.reshape(-1, 1)                 # MATLAB column vector
        m, n = a.shape
        f = np.asfortranarray(a)                 # column-major storage
        if a.dtype == np.bool_:
            f = np.asfortranarray(a.astype(np.uint8))
            return L.shim_logical(f.ctypes.data, m, n)
        if a.dtype == np.uint8:
            return L.shim_uint8(f.ctypes.data, m, n)
        if a.dtype == np.float32:
            return L.shim_single(f.ctypes.data, m, n)
        if a.dtype == np.int8:
            return L.shim_int8(f.ctypes.data, m, n)
        f = np.asfortranarray(a.astype(np.float64))
        return L.shim_double(f.ctypes.data, m, n)

    def _from_mx(self, p):
        L = self.L
        cls, m, n = L.shim_class(p), L.mxGetM(p), L.mxGetN(p)
        dt = {LOGICAL: np.uint8, DOUBLE: np.float64, UINT64: np.uint64}[cls]
        buf = (C.c_char * (m * n * np.dtype(dt).itemsize)).from_address(L.mxGetData(p))
        a = np.frombuffer(buf, dtype=dt).reshape((m, n), order="F").copy()
        return a.astype(np.bool_) if cls == LOGICAL else a

    def __call__(self, *args, nlhs=1):
        L = self.L
        prhs = (C.c_void_p * len(args))(*[self._to_mx(a) for a in args])
        plhs = (C.c_void_p * max(1, nlhs))()
        eid, emsg = C.create_string_buffer(512), C.create_string_buffer(512)
        rc = L.shim_mex(nlhs, plhs, len(args), prhs, eid, emsg, 512)
        for a, p in zip(args, prhs):
            if not isinstance(a, Handle):
                L.mxDestroyArray(p)
        if rc:
            raise MexError(eid.value.decode(), emsg.value.decode())
        outs = []
        for i in range(max(1, nlhs)):
            if not plhs[i]:
                outs.append(None)
                continue
            if L.shim_class(plhs[i]) == UINT64:
                outs.append(Handle(plhs[i]))      # keep the mxArray: it IS the handle value MATLAB would hold
            else:
                outs.append(self._from_mx(plhs[i]))
                L.mxDestroyArray(plhs[i])
        return outs[0] if nlhs <= 1 else outs


class Handle:
    def __init__(self, mx):
        self.mx = mx


@pytest.fixture(scope="module")
def mex():
    return Mex()


def test_gateway_compiles_and_rejects_bad_calls(mex):
    with pytest.raises(MexError) as e:
        mex(3.0)
    assert e.value.identifier == ID_ERROR
    with pytest.raises(MexError) as e:
        mex("frobnicate", 1.0)
    assert e.value.identifier == ID_ERROR and "unknown command" in e.value.message
    with pytest.raises(MexError) as e:
        mex("create", 1.0, 384.0)                   # too few arguments: an error, not a crash
    assert e.value.identifier == ID_ERROR
    # parameter errors surface with the identifier the reference's callers catch and skip
    for bad in (("create", 3.0, 384.0, 8.0, 1.0), ("create", 1.0, 17.0, 8.0, 1.0), ("create", 1.0, 384.0, 0.0, 1.0)):
        with pytest.raises(MexError) as e:
            mex(*bad)
        assert e.value.identifier == ID_UNSUPPORTED, bad
    with pytest.raises(MexError) as e:
        mex("decode", 1.0, np.zeros((4, 1)))        # not a handle
    assert e.value.identifier == ID_ERROR


def test_gateway_without_gpu_fails_loudly(mex):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(MexError) as e:
        mex("create", 1.0, 384.0, 8.0, 1.0)
    assert e.value.identifier == ID_ERROR and "no CPU path" in e.value.message


@pytest.mark.gpu
@pytest.mark.parametrize("bg,Z,E,esn0", [(1, 384, 25272, -0.2), (2, 52, 2000, -1.0), (2, 6, 100, 2.0)])
def test_gateway_decode_matches_oracle(mex, O, bg, Z, E, esn0):
    from conftest import make_llr
    rng = np.random.default_rng(7)
    filler = {384: 0, 52: 104, 6: 24}[Z]
    info, llr = make_llr(O, bg, Z, 6, E, esn0, rng, filler=filler)
    d = O.dims(bg, Z)
    cw_tilde = llr.astype(np.float64).T                      # (n_cw x batch), one cw_tilde per column, +Inf filler
    assert cw_tilde.shape == (d["ncw"], 6)
    # layered min-sum, parity-check stop (NRLDPCDecoder.m:120), three outputs
    h = mex("create", float(bg), float(Z), 8.0, 1.0)
    c_hat, n_it, ok = mex("decode", h, cw_tilde, 0.0, nlhs=3)
    ref = O.decode_nms(bg, Z, llr, 8, early_term=True)
    assert c_hat.dtype == np.bool_ and c_hat.shape == (d["K"], 6)          # logical K x batch
    assert (c_hat.T.astype(np.uint8) == ref["hard"]).all()
    assert (n_it.ravel() == ref["iters"]).all() and (ok.ravel().astype(np.uint8) == ref["parity_ok"]).all()
    # single column = the reference's calling pattern (one step per code block, NRLDPCDecoder.m:257-266)
    one = mex("decode", h, cw_tilde[:, 2])
    assert one.shape == (d["K"], 1) and (one[:, 0].astype(np.uint8) == ref["hard"][2]).all()
    with pytest.raises(MexError) as e:
        mex("decode", h, cw_tilde[:-1])
    assert e.value.identifier == ID_ERROR and "rows" in e.value.message
    with pytest.raises(MexError) as e:
        mex("decode", h, cw_tilde, 2.0)                                   # n_rows out of range
    assert e.value.identifier == ID_UNSUPPORTED
    # single(cw_tilde): same decisions (the min-sum kernels compute in float32 anyway); int8 codes with a scale: equal to the oracle
    # on the de-quantised values
    c32 = mex("decode", h, cw_tilde.astype(np.float32), 0.0)
    assert (c32.T.astype(np.uint8) == ref["hard"]).all()
    scale = 0.25
    q = np.clip(np.rint(np.nan_to_num(llr, posinf=0.0) / scale), -127, 126).astype(np.int8)
    q[np.isposinf(llr)] = 127
    deq = (np.float32(scale) * q.astype(np.float32)).astype(np.float32)
    deq[q == 127] = np.inf
    c8, it8 = mex("decode", h, np.ascontiguousarray(q.T), 0.0, scale, nlhs=2)
    ref8 = O.decode_nms(bg, Z, deq, 8, early_term=True)
    assert (c8.T.astype(np.uint8) == ref8["hard"]).all() and (it8.ravel() == ref8["iters"]).all()
    with pytest.raises(MexError) as e:
        mex("decode", h, np.ascontiguousarray(q.T), 0.0)                  # int8 without a scale
    assert e.value.identifier == ID_ERROR
    mex("destroy", h, nlhs=0)
    # the reference's own algorithm on the reference's own doubles
    h = mex("create", float(bg), float(Z), 8.0, 1.0, 0.75, 0.0, 1.0)
    c_hat, n_it = mex("decode", h, cw_tilde, nlhs=2)
    refb = O.decode_bp(bg, Z, llr.astype(np.float64), 8)
    assert (c_hat.T.astype(np.uint8) == refb["hard"]).all() and (n_it.ravel() == refb["iters"]).all()
    mex("destroy", h, nlhs=0)


@pytest.mark.gpu
def test_gateway_encode_matches_oracle(mex, O):
    rng = np.random.default_rng(8)
    for bg, Z in ((1, 384), (2, 52)):
        d = O.dims(bg, Z)
        info = rng.integers(0, 2, (5, d["K"]), dtype=np.uint8)
        h = mex("create", float(bg), float(Z), 8.0, 0.0)
        cw = mex("encode", h, info.T.astype(np.float64))               # K x batch double, as NRLDPCEncoder.m:158 passes it
        assert cw.dtype == np.float64 and cw.shape == (d["ncw"], 5)
        assert (cw.T == O.encode(bg, Z, info)).all()
        cw2 = mex("encode", h, info.T.astype(np.bool_))                 # logical input
        assert (cw2 == cw).all()
        with pytest.raises(MexError) as e:
            mex("encode", h, info.T[:-1].astype(np.float64))
        assert e.value.identifier == ID_ERROR
        mex("destroy", h, nlhs=0)
