"""Host-side mirror of NRModulator.m / NRDemodulator.m: same property names (Modulation, DecisionMethod,
Variance, ModulationOrder, Q_m), same step protocol and error identifier, with the mapping and the LLR
computation done by the sm_100a kernels behind the C ABI (nrldpc_modulate / nrldpc_demodulate) instead of
comm.PSK* / comm.RectangularQAM*.  No CPU implementation lives here: without the CUDA library step() raises.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .capi import UnsupportedParameters

_Q_M = {"BPSK": 1, "QPSK": 2, "16QAM": 4, "64QAM": 6, "256QAM": 8}           # NRModulator.m:47-63
_METHOD = {"Log-likelihood ratio": capi.DEMOD_LLR, "Approximate log-likelihood ratio": capi.DEMOD_APPROX,
           "Hard decision": capi.DEMOD_HARD}                                    # NRDemodulator.m:10


def _q_m(modulation):
    if modulation not in _Q_M:
        raise UnsupportedParameters("Unsupported modulation")                  # NRModulator.m:43,61
    return _Q_M[modulation]


class _Modem:
    def __init__(self, **kw):
        self.Modulation = "BPSK"                                               # NRModulator.m:4, NRDemodulator.m:4
        self.device = -1
        self._h = None
        for k, v in kw.items():
            if not hasattr(self, k):
                raise capi.NRLDPCError(f"unknown property {k}")
            setattr(self, k, v)

    @property
    def ModulationOrder(self):
        return 1 << _q_m(self.Modulation)

    @property
    def Q_m(self):
        return _q_m(self.Modulation)

    def _handle(self):
        if self._h is None:
            _q_m(self.Modulation)
            self._h = capi.Handle(2, 2, 1, False, device=self.device)         # any code: the modem calls ignore it
        return self._h

    def release(self):
        if self._h is not None:
            self._h.close()
            self._h = None

    def reset(self):
        pass


class NRModulator(_Modem):
    """tx = step(hMod, bits): bits (0/1) -> complex symbols, TS 38.211 section 5.1 (NRModulator.m:69-89)."""

    def step(self, bits):
        import torch
        bits = np.ascontiguousarray(np.asarray(bits).ravel(), dtype=np.uint8)
        Qm = self.Q_m
        if bits.size % Qm:
            raise capi.NRLDPCError("the number of bits must be a multiple of Q_m")
        h = self._handle()
        d_bits = torch.from_numpy(bits).cuda()
        sym = torch.empty((bits.size // Qm, 2), dtype=torch.float32, device="cuda")
        h.modulate_raw(d_bits, bits.size, Qm, sym, stream=torch.cuda.current_stream().cuda_stream)
        s = sym.cpu().numpy()
        return s[:, 0] + 1j * s[:, 1]


class NRDemodulator(_Modem):
    """llr = step(hDemod, rx) (NRDemodulator.m:72-96); Variance is tunable between steps (:94-96)."""

    def __init__(self, **kw):
        self.DecisionMethod = "Log-likelihood ratio"                            # NRDemodulator.m:5
        self.Variance = 1.0                                                     # :14
        super().__init__(**kw)

    def step(self, rx):
        import torch
        if self.DecisionMethod not in _METHOD:
            raise UnsupportedParameters("Unsupported decision method")
        rx = np.asarray(rx).ravel()
        s = np.ascontiguousarray(np.stack([rx.real, rx.imag], axis=1), dtype=np.float32)
        Qm = self.Q_m
        h = self._handle()
        d_sym = torch.from_numpy(s).cuda()
        out = torch.empty(rx.size * Qm, dtype=torch.float32, device="cuda")
        h.demodulate_raw(d_sym, rx.size, Qm, float(self.Variance), _METHOD[self.DecisionMethod], out,
                         stream=torch.cuda.current_stream().cuda_stream)
        return out.cpu().numpy().astype(np.float64)
