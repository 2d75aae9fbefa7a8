"""Host-side mirror of the reference's System objects: NRLDPC (NRLDPC.m), NRLDPCEncoder
(NRLDPCEncoder.m) and NRLDPCDecoder (NRLDPCDecoder.m) -- same property names, same step / reset /
release protocol, same error identifiers -- with the LDPC arithmetic and the rate matching routed
through the C ABI (libnrldpc_b200.so) instead of comm.LDPCEncoder / comm.LDPCDecoder and the
interpreted bit_selection loops.  All C code blocks of a transport block go to the GPU as one batch.

Only host bookkeeping lives here (parameter derivation, CRC attach/check, segmentation,
concatenation).  There is no CPU implementation of encode / decode / rate matching in this
package: without the CUDA library those calls raise.
"""
from __future__ import annotations

import math

import numpy as np

from . import capi
from .capi import NRLDPCError, UnsupportedParameters

_CRC_POLY = {"CRC24A": (0x864CFB, 24), "CRC24B": (0x800063, 24), "CRC16": (0x1021, 16), "None": (0, 0)}
_CRC_TABLES: dict = {}


def get_3gpp_crc_polynomial(crc: str):
    """(polynomial without the leading term, L) -- get_3gpp_crc_polynomial.m:3-17."""
    if crc not in _CRC_POLY:
        raise UnsupportedParameters("Invalid CRC identifier.")
    return _CRC_POLY[crc]


def _crc_table(kind):
    if kind not in _CRC_TABLES:
        poly, L = _CRC_POLY[kind]
        mask, top = (1 << L) - 1, 1 << (L - 1)
        tab = []
        for byte in range(256):
            reg = byte << (L - 8)
            for _ in range(8):
                reg = ((reg << 1) ^ poly) & mask if reg & top else (reg << 1) & mask
            tab.append(reg)
        _CRC_TABLES[kind] = tab
    return _CRC_TABLES[kind]


def crc_bits(kind: str, bits) -> np.ndarray:
    """Parity bits (MSB first) of comm.CRCGenerator(Polynomial) with its defaults: zero initial state,
    direct method off, no reflection, no final XOR (NRLDPCEncoder.m:45-47,80,114)."""
    poly, L = _CRC_POLY[kind]
    bits = np.asarray(bits).astype(np.uint8) & 1
    pad = (-len(bits)) % 8  # leading zeros do not change a zero-initialised CRC
    by = np.packbits(np.concatenate([np.zeros(pad, np.uint8), bits]))
    tab, mask, reg = _crc_table(kind), (1 << L) - 1, 0
    for b in by.tolist():
        reg = ((reg << 8) & mask) ^ tab[((reg >> (L - 8)) ^ b) & 0xFF]
    return np.array([(reg >> (L - 1 - i)) & 1 for i in range(L)], dtype=np.uint8)


def matlab_round(x: float) -> int:
    """MATLAB round(): half away from zero (plot_BLER_vs_SNR.m:94)."""
    return int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


class NRLDPC:
    """Parameter base class: every TS 38.212 derived quantity as a read-only property
    (NRLDPC.m:90-228 declarations, :297-543 getters)."""

    _nontunable = ("BG", "A", "I_LBRM", "TBS_LBRM")
    _tunable = ("rv_id", "G", "Q_m", "N_L", "CBGTI")

    def __init__(self, **kw):
        object.__setattr__(self, "_locked", False)
        self.BG, self.A, self.I_LBRM, self.TBS_LBRM = 1, 44, 0, math.inf       # NRLDPC.m:28-47
        self.rv_id, self.G, self.Q_m, self.N_L, self.CBGTI = 0, 132, 1, 1, []  # NRLDPC.m:57-84
        self._set_properties(kw)

    def _set_properties(self, kw):
        for k, v in kw.items():
            if k not in self._nontunable + self._tunable + getattr(self, "_extra_props", ()):
                raise NRLDPCError(f"Unrecognized property '{k}'.")
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if name in self._nontunable or name in getattr(self, "_extra_nontunable", ()):
            if getattr(self, "_locked", False):
                raise NRLDPCError(f"Nontunable property '{name}' cannot be changed after step(); call release() first.")
        v = value
        if name == "BG" and (v < 1 or v > 2):                                   # NRLDPC.m:240-245
            raise UnsupportedParameters("BG selects the TS 38.212 base graph: 1 or 2.")
        if name == "A" and v < 0:                                               # :247-252
            raise UnsupportedParameters("A (payload bits) cannot be below zero.")
        if name == "TBS_LBRM" and v < 0:                                        # :254-259
            raise UnsupportedParameters("TBS_LBRM cannot be below zero.")
        if name == "rv_id" and (v < 0 or v > 3):                                # :263-268
            raise UnsupportedParameters("rv_id (redundancy version) is one of 0..3.")
        if name == "G" and v < 0:                                               # :270-275
            raise UnsupportedParameters("G (coded bits available) cannot be below zero.")
        if name == "Q_m" and v not in (1, 2, 4, 6, 8):                          # :278-283
            raise UnsupportedParameters("Q_m (bits per symbol) is one of 1, 2, 4, 6, 8.")
        if name == "N_L" and (v < 1 or v > 4):                                  # :289-294
            raise UnsupportedParameters("N_L (layers) is one of 1..4.")
        object.__setattr__(self, name, value)

    # --- Dependent getters ---------------------------------------------------------------------
    @property
    def transport_block_CRC(self):        # NRLDPC.m:297-303
        return "CRC24A" if self.A > 3824 else "CRC16"

    @property
    def transport_block_L(self):          # :311-313
        return _CRC_POLY[self.transport_block_CRC][1]

    @property
    def B(self):                          # :316-318
        return self.A + self.transport_block_L

    @property
    def K_cb(self):                       # :321-331
        return 8448 if self.BG == 1 else 3840

    @property
    def code_block_CRC(self):             # :347-353
        return "None" if self.B <= self.K_cb else "CRC24B"

    @property
    def code_block_L(self):               # :361-363
        return _CRC_POLY[self.code_block_CRC][1]

    @property
    def C(self):                          # :334-344
        return 1 if self.B <= self.K_cb else -(-self.B // (self.K_cb - self.code_block_L))

    @property
    def B_prime(self):                    # :366-377
        return self.B if self.B <= self.K_cb else self.B + self.C * self.code_block_L

    @property
    def K_prime(self):                    # :380-382 (validatePropertiesImpl guarantees divisibility)
        return self.B_prime // self.C if self.B_prime % self.C == 0 else self.B_prime / self.C

    @property
    def K_b(self):                        # :385-406
        if self.BG == 1:
            return 22
        kp = self.K_prime
        return 10 if kp > 640 else 9 if kp > 560 else 8 if kp > 192 else 6

    @property
    def Z_c(self):                        # :409-411 -> get_3gpp_lifting_size.m
        return capi.lifting_size(self.K_b, int(math.ceil(self.K_prime)))

    @property
    def K(self):                          # :414-425
        return self.Z_c * (22 if self.BG == 1 else 10)

    @property
    def i_LS(self):                       # :428-430 -> get_3gpp_set_index.m
        return capi.set_index(self.Z_c)

    @property
    def V(self):                          # :433-435: base graph as (rows, cols, shifts) edge lists
        return capi.base_graph(self.BG, self.i_LS)

    @property
    def N(self):                          # :443-454
        return self.Z_c * (66 if self.BG == 1 else 50)

    @property
    def N_ref(self):                      # :457-460
        return math.inf if math.isinf(self.TBS_LBRM) else int(math.floor(self.TBS_LBRM / (self.C * (2.0 / 3.0))))

    @property
    def N_cb(self):                       # :463-469
        return self.N if self.I_LBRM == 0 else int(min(self.N, self.N_ref))

    @property
    def CBGTI_flags(self):                # :471-477
        flags = np.ones(self.C, dtype=np.int64)
        for v in self.CBGTI:
            if v < self.C:
                flags[v] = 0
        return flags

    @property
    def C_prime(self):                    # :480-482
        return int(self.CBGTI_flags.sum())

    @property
    def E_r(self):                        # :485-507
        C_, Cp, flags, q = self.C, self.C_prime, self.CBGTI_flags, self.N_L * self.Q_m
        out, j = np.zeros(C_, dtype=np.int64), 0
        for r in range(C_):
            if flags[r] == 0:
                continue
            if j <= Cp - ((self.G // q) % Cp) - 1:
                out[r] = q * (self.G // (q * Cp))
            else:
                out[r] = q * (-(-self.G // (q * Cp)))
            j += 1
        return out

    @property
    def k_0(self):                        # :510-543
        num = {1: (0, 17, 33, 56), 2: (0, 13, 25, 43)}[self.BG][self.rv_id]
        den = 66 if self.BG == 1 else 50
        return (num * self.N_cb) // (den * self.Z_c) * self.Z_c

    def validate_properties(self):        # validatePropertiesImpl, :551-559
        if self.B_prime % self.C != 0:
            raise UnsupportedParameters("C code blocks must share B_prime evenly (B_prime mod C = 0).")
        if self.G % (self.Q_m * self.N_L) != 0:
            raise UnsupportedParameters("G has to be a whole number of Q_m*N_L groups.")

    # --- matlab.System protocol ------------------------------------------------------------------
    def step(self, x):
        if not self._locked:
            self.validate_properties()
            self._setup()
            object.__setattr__(self, "_locked", True)
        return self._step(x)

    __call__ = step

    def reset(self):
        if self._locked:
            self._reset()

    def release(self):
        self._release()
        object.__setattr__(self, "_locked", False)

    def _setup(self): ...
    def _reset(self): ...
    def _release(self): ...

    def _rm(self, r):
        return dict(E=int(self.E_r[r]), k_0=int(self.k_0), N_cb=int(self.N_cb), K_prime=int(self.K_prime), Q_m=int(self.Q_m))


def _col(x, n, name, ident="column vector"):
    x = np.asarray(x)
    if x.ndim == 2 and x.shape[1] == 1:
        x = x[:, 0]
    if x.ndim != 1 or x.shape[0] != n:
        raise NRLDPCError(f"{name} should be a {ident} of length {n}.")
    return x


class NRLDPCEncoder(NRLDPC):
    """g = step(enc, a): TS 38.212 5.1 -> 5.5 TX chain (NRLDPCEncoder.m:60-67)."""

    def _setup(self):                     # NRLDPCEncoder.m:44-50
        object.__setattr__(self, "_h", capi.Handle(self.BG, self.Z_c, 1, False))

    def _release(self):
        h = getattr(self, "_h", None)
        if h is not None:
            h.close()
            object.__setattr__(self, "_h", None)

    def _step(self, a):
        a = _col(a, self.A, "a").astype(np.uint8)
        A_, B_, C_, K_, Kp, L = self.A, self.B, self.C, self.K, self.K_prime, self.code_block_L
        # crc_calculation, NRLDPCEncoder.m:70-89
        b = np.concatenate([a, crc_bits(self.transport_block_CRC, a)])
        assert len(b) == B_
        # code_block_segmentation, :92-124 (filler encoded as 0, :153)
        c = np.zeros((C_, K_), dtype=np.uint8)
        s = 0
        for r in range(C_):
            c[r, :Kp - L] = b[s:s + Kp - L]
            s += Kp - L
            if C_ > 1:
                c[r, Kp - L:Kp] = crc_bits("CRC24B", c[r, :Kp - L])
        # LDPC_coding (:127-165) + bit_selection (:168-197) + bit_interleaving (:200-225) on the GPU
        cw = self._h.encode(c)
        E_r = self.E_r
        g = np.zeros(self.G, dtype=np.uint8)
        k = 0
        for r in range(C_):               # code_block_concatenation, :228-256
            if E_r[r] == 0:
                continue
            rm = self._rm(r)
            f = self._h.rate_match(cw[r:r + 1], **rm)[0]
            g[k:k + E_r[r]] = f
            k += E_r[r]
        return g.astype(np.float64)


class NRLDPCDecoder(NRLDPC):
    """a_hat = step(dec, g_tilde): RX chain (NRLDPCDecoder.m:133-140); returns an empty array when the
    transport-block CRC or any code-block CRC fails (:337-339).

    Engine-specific properties (not in the reference): ``early_termination`` (default True = the
    reference's 'Parity check satisfied', NRLDPCDecoder.m:120), ``alpha`` (min-sum normalisation),
    ``trim_rows`` (skip base rows whose parity bits were never received) and ``algorithm``:
    ``'Layered normalized min-sum'`` (default, the fast path) or ``'Sum-product'`` = the reference's own
    flooding sum-product in float64 on the whole H, as comm.LDPCDecoder runs it at :120,:265 (rows are never
    trimmed in this mode: the reference always passes the full matrix)."""

    _extra_props = ("I_HARQ", "iterations", "early_termination", "alpha", "trim_rows", "algorithm")
    ALGORITHMS = {"Layered normalized min-sum": capi.ALG_NMS, "Sum-product": capi.ALG_BP}
    _extra_nontunable = ("I_HARQ",)

    def __init__(self, **kw):
        object.__setattr__(self, "_locked", False)
        self.I_HARQ = 0                   # NRLDPCDecoder.m:34
        self.iterations = 50              # :41
        self.early_termination, self.alpha, self.trim_rows = True, 0.75, True
        self.algorithm = "Layered normalized min-sum"
        super().__init__(**kw)

    def _setup(self):                     # NRLDPCDecoder.m:107-130
        # `iterations` is read here only, as in the reference (:120): changing it later has no effect
        if self.algorithm not in self.ALGORITHMS:
            raise UnsupportedParameters("Valid values of algorithm are 'Layered normalized min-sum' and 'Sum-product'.")
        object.__setattr__(self, "_h", capi.Handle(self.BG, self.Z_c, int(self.iterations),
                                                   bool(self.early_termination), float(self.alpha),
                                                   algorithm=self.ALGORITHMS[self.algorithm]))
        self._reset()

    def _reset(self):                     # resetImpl, :343-356
        object.__setattr__(self, "d_tilde_buffer", np.zeros((self.C, self.N), dtype=np.float32))
        object.__setattr__(self, "b_hat_buffer", np.zeros(self.B, dtype=np.uint8))
        object.__setattr__(self, "code_block_CRC_passed", np.zeros(self.C, dtype=np.uint8))
        object.__setattr__(self, "_max_col", 0)

    def _release(self):
        h = getattr(self, "_h", None)
        if h is not None:
            h.close()
            object.__setattr__(self, "_h", None)

    def _active_rows(self, E_max):
        """Base rows that can carry information: a row whose own parity column was never received
        has a zero LLR on a degree-1 variable and sends zero messages for ever."""
        Z, kcols = self.Z_c, 22 if self.BG == 1 else 10
        rows_all = 46 if self.BG == 1 else 42
        if not self.trim_rows or self.algorithm == "Sum-product":
            return rows_all
        nfill = self.K - max(self.K_prime, 2 * Z) if self.K > max(self.K_prime, 2 * Z) else 0
        span = self.k_0 + E_max + nfill  # last circular-buffer index touched (filler is skipped)
        hi = self.N_cb if span >= self.N_cb else span
        object.__setattr__(self, "_max_col", max(self._max_col, hi))
        cols_needed = -(-(self._max_col + 2 * Z) // Z)
        return int(min(rows_all, max(4, cols_needed - kcols)))

    def _step(self, g_tilde):
        G_, C_, K_, Kp, L = self.G, self.C, self.K, self.K_prime, self.code_block_L
        g_tilde = np.asarray(g_tilde)
        if g_tilde.ndim == 2 and g_tilde.shape[1] == 1:
            g_tilde = g_tilde[:, 0]
        if g_tilde.ndim != 1 or g_tilde.shape[0] != G_:                       # NRLDPCDecoder.m:148-150
            raise NRLDPCError("g_tilde should be a column vector of length G.")
        g_tilde = g_tilde.astype(np.float32)
        E_r, flags = self.E_r, self.CBGTI_flags
        # code_block_concatenation (:143-169) on the host; bit_interleaving + bit_selection +
        # cw_tilde assembly (:172-242, :262-264) on the GPU, per code block (E_r may differ)
        llr_cw = np.zeros((C_, self._h.n_cw), dtype=np.float32)
        k = 0
        for r in range(C_):
            E = int(E_r[r])
            harq = self.d_tilde_buffer[r:r + 1] if self.I_HARQ != 0 else None
            if E == 0:
                if harq is not None:  # nothing received for this block: decoder input is the buffer
                    llr_cw[r, 2 * self.Z_c:] = harq[0]
                    f0, f1 = max(Kp - 2 * self.Z_c, 0), K_ - 2 * self.Z_c
                    llr_cw[r, 2 * self.Z_c + f0:2 * self.Z_c + f1] = np.inf
                continue
            rm = self._rm(r)
            llr_cw[r] = self._h.rate_recover(g_tilde[k:k + E], harq=harq, **rm)[0]
            k += E
        n_rows = self._active_rows(int(E_r.max()))
        # LDPC_coding, :245-268
        c_hat = self._h.decode(llr_cw, n_rows=n_rows)["hard"]
        # code_block_segmentation, :271-318
        b_hat = self.b_hat_buffer.copy() if self.I_HARQ != 0 else np.zeros(self.B, dtype=np.uint8)
        passed = self.code_block_CRC_passed.copy()
        s = 0
        for r in range(C_):
            failed = False
            if C_ > 1:
                failed = bool(crc_bits("CRC24B", c_hat[r, :Kp]).any())
            if not failed and flags[r] == 1:
                b_hat[s:s + Kp - L] = c_hat[r, :Kp - L]
                passed[r] = 1
            s += Kp - L
        if self.I_HARQ != 0:
            object.__setattr__(self, "b_hat_buffer", b_hat)
        object.__setattr__(self, "code_block_CRC_passed", passed)
        # crc_calculation, :321-340
        tb_failed = bool(crc_bits(self.transport_block_CRC, b_hat).any())
        if tb_failed or not passed.all():
            return np.zeros(0, dtype=np.float64)
        return b_hat[:self.A].astype(np.float64)
