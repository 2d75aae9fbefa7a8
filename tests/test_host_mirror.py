"""Host-side mirror of NRLDPC.m getters / validation / error identifiers -- CPU only."""
import math

import numpy as np
import pytest

from ldpc_3gpp_matlab_b200 import nrldpc
from ldpc_3gpp_matlab_b200.capi import NRLDPCError, UnsupportedParameters


def test_getters_against_appendix_a3(golden_params):
    for g in golden_params:
        o = nrldpc.NRLDPC(A=g["A"], BG=g["BG"], G=g["G"], Q_m=2)
        assert o.transport_block_L == g["tb_L"]
        for k in ("B", "C", "K_prime", "K_b", "Z_c", "i_LS", "K", "N"):
            assert getattr(o, k) == g[k], (g["A"], k)
        assert o.E_r.tolist() == g["E_r"]
        assert o.N_cb == g["N"] and o.k_0 == 0
        o.validate_properties()


def test_getters_against_oracle_random(O):
    """Parameter distribution of the reference's only test, testbench.m:21-36."""
    rng = np.random.default_rng(0)
    n_ok = 0
    for _ in range(300):
        R = rng.choice([1 / 5, 1 / 3, 2 / 5, 1 / 2, 2 / 3, 3 / 4, 5 / 6, 8 / 9])
        A = int(math.ceil(100000 ** rng.random()))
        BG = 2 if (A <= 292 or (A <= 3824 and R <= 0.67) or R <= 0.25) else 1
        Qm = int(rng.choice([1, 2, 4, 6, 8])); NL = int(rng.integers(1, 5)); rv = int(rng.integers(0, 4))
        G = int(math.ceil(A / R / (Qm * NL))) * Qm * NL
        lbrm = int(rng.integers(0, 2)); tbs = int(A * rng.integers(1, 4))
        p = O.params(BG, A, G, Q_m=Qm, N_L=NL, rv_id=rv, I_LBRM=lbrm, TBS_LBRM=tbs)
        o = nrldpc.NRLDPC(A=A, BG=BG, G=G, Q_m=Qm, N_L=NL, rv_id=rv, I_LBRM=lbrm, TBS_LBRM=tbs)
        if p is None:
            with pytest.raises(UnsupportedParameters):
                o.validate_properties(); o.Z_c
            continue
        n_ok += 1
        for k in ("B", "C", "B_prime", "K_prime", "K_b", "Z_c", "i_LS", "K", "N", "N_cb", "k_0"):
            assert getattr(o, k) == getattr(p, k), (A, BG, k)
        assert o.E_r.tolist() == list(p.E_r[:p.C])
    assert n_ok > 200


def test_setter_validation_and_identifiers():
    for kw in (dict(BG=3), dict(A=-1), dict(rv_id=4), dict(G=-2), dict(Q_m=3), dict(N_L=5), dict(TBS_LBRM=-1)):
        with pytest.raises(UnsupportedParameters) as ei:
            nrldpc.NRLDPC(**kw)
        assert ei.value.identifier == "ldpc_3gpp_matlab:UnsupportedParameters"
    o = nrldpc.NRLDPC(A=100, BG=1, G=301, Q_m=2)
    with pytest.raises(UnsupportedParameters):
        o.validate_properties()
    with pytest.raises(NRLDPCError) as ei:
        nrldpc.NRLDPC(bogus=1)
    assert ei.value.identifier == "ldpc_3gpp_matlab:Error"


def test_matlab_round_and_G():
    assert nrldpc.matlab_round(4738.5) == 4739 and nrldpc.matlab_round(2.5) == 3 and nrldpc.matlab_round(-2.5) == -3
    assert nrldpc.matlab_round(8424 / (8 / 9) / 2) * 2 == 9478      # plot_BLER_vs_SNR.m:94


def test_crc_against_oracle(O):
    rng = np.random.default_rng(1)
    for kind in ("CRC16", "CRC24A", "CRC24B"):
        for n in (1, 7, 8, 20, 100, 1001, 8424):
            bits = rng.integers(0, 2, n, dtype=np.uint8)
            par = nrldpc.crc_bits(kind, bits)
            assert (par == O.crc(kind, bits)).all()
            assert not nrldpc.crc_bits(kind, np.concatenate([bits, par])).any()


def test_cbgti_and_lbrm():
    o = nrldpc.NRLDPC(A=20000, BG=1, G=60000, Q_m=2, CBGTI=[1])
    assert o.C == 3 and o.C_prime == 2 and o.E_r[1] == 0 and o.E_r.sum() == 60000
    o = nrldpc.NRLDPC(A=8424, BG=1, G=25272, Q_m=2, I_LBRM=1, TBS_LBRM=8424)
    assert o.N_ref == 12636 and o.N_cb == 12636
    assert nrldpc.NRLDPC(A=8424, BG=1, G=25272, Q_m=2, I_LBRM=1, TBS_LBRM=8424, rv_id=2).k_0 == (33 * 12636) // (66 * 384) * 384


def test_required_snr_interpolation_follows_interp1():
    """plot_SNR_vs_A.m:175: interp1(log10([prev_BLER, BLER]), [prev_EsN0, EsN0], log10(target_BLER))."""
    from ldpc_3gpp_matlab_b200.bler import interp_required_snr as f
    assert f(0.1, 0.001, 1.0, 1.5, 0.01) == pytest.approx(1.25)
    assert f(1.0, 0.005, -2.1, -2.0, 1e-2) == pytest.approx(-2.1 + 0.1 * math.log10(1e-2) / math.log10(0.005))
    assert f(0.03, 0.01, 0.0, 0.5, 0.01) == pytest.approx(0.5)            # the target is the last point itself
    assert math.isnan(f(1.0, 0.0, 0.5, 1.0, 0.01))                        # a point without errors: log10(0)
    assert math.isnan(f(None, 0.001, None, 1.0, 0.01))                    # no previous point
    assert math.isnan(f(0.5, 0.2, 0.0, 0.5, 0.01))                        # target outside the bracket
