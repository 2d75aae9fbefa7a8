"""Structure guards: tables, lifting, parameters (SURVEY.md Appendix A fixtures) -- CPU only."""
import hashlib

import numpy as np
import pytest

from conftest import ALL_Z


@pytest.mark.parametrize("bg", [1, 2])
def test_table_guards(O, golden_tables, bg):
    g = golden_tables[f"bg{bg}"]
    t = O.table(bg)
    assert t.shape == (g["edges"], 10)
    assert t[:, 0].max() + 1 == g["rows"] and t[:, 1].max() + 1 == g["cols"]
    assert int(t[:, 0].sum()) == g["sum_row"] and int(t[:, 1].sum()) == g["sum_col"]
    assert [int(v) for v in t[:, 2:].sum(0)] == g["sum_V"]
    assert [int(v) for v in t[:, 2:].max(0)] == g["max_V"]
    assert t[0].tolist() == g["first"] and t[-1].tolist() == g["last"]
    text = "\n".join(" ".join(str(int(v)) for v in row) for row in t)
    assert hashlib.sha256(text.encode()).hexdigest() == g["sha256"]
    assert np.bincount(t[:, 0]).tolist() == g["row_deg"]
    key = t[:, 0] * 100 + t[:, 1]
    assert (np.diff(key) > 0).all(), "edges must be sorted by (row, col) without duplicates"
    # extension part: row r >= 4 has exactly one entry in columns >= kcols+4, at column kcols+r, shift 0
    kc = 22 if bg == 1 else 10
    for r in range(4, g["rows"]):
        ext = t[(t[:, 0] == r) & (t[:, 1] >= kc + 4)]
        assert ext.shape[0] == 1 and ext[0, 1] == kc + r and not ext[0, 2:].any()


def test_lifting_sizes(O, golden_tables):
    zs = golden_tables["lifting_sizes"]
    assert zs == ALL_Z and len(zs) == 51
    for Z in range(1, 400):
        s = O.set_index(Z)
        assert (s >= 0) == (Z in zs)
    assert [O.set_index(z) for z in (2, 3, 5, 7, 9, 11, 13, 15, 384, 208, 52)] == [0, 1, 2, 3, 4, 5, 6, 7, 1, 6, 6]
    assert O.lifting_size(22, 8448) == 384 and O.lifting_size(10, 1957) == 208 and O.lifting_size(6, 36) == 6
    assert O.lifting_size(22, 8449) == -1


def test_params_against_appendix_a3(O, golden_params):
    for g in golden_params:
        p = O.params(g["BG"], g["A"], g["G"], Q_m=2)
        assert p is not None
        for k in ("tb_L", "B", "C", "K_prime", "K_b", "Z_c", "i_LS", "K", "N"):
            assert getattr(p, k) == g[k], (g["A"], k)
        assert p.K - p.K_prime == g["filler"]
        assert list(p.E_r[:p.C]) == g["E_r"]


def test_k0(O, golden_tables):
    for bg in (1, 2):
        for rv in range(4):
            p = O.params(bg, 1000, 3000, Q_m=2, rv_id=rv)
            assert p.k_0 == golden_tables["k0_num"][str(bg)][rv] * p.Z_c


def test_unsupported_parameters(O):
    assert O.params(3, 100, 200) is None
    assert O.params(1, 100, 201, Q_m=2) is None          # G % (Q_m*N_L)
    assert O.params(1, 100, 200, Q_m=3) is None


@pytest.mark.parametrize("bg", [1, 2])
def test_pcm_follows_get_pcm(O, bg):
    """get_pcm.m:8: block (r,c) = circshift(speye(Z), mod(V,Z), 2)."""
    for Z in (2, 7, 52):
        t = O.table(bg)
        ils = O.set_index(Z)
        rows, cols = O.pcm(bg, Z)
        H = np.zeros((t[:, 0].max() * Z + Z, t[:, 1].max() * Z + Z), np.uint8)
        H[rows, cols] = 1
        for e in range(t.shape[0]):
            blk = H[t[e, 0] * Z:(t[e, 0] + 1) * Z, t[e, 1] * Z:(t[e, 1] + 1) * Z]
            want = np.roll(np.eye(Z, dtype=np.uint8), int(t[e, 2 + ils]) % Z, axis=1)
            assert (blk == want).all()
        assert H.sum() == t.shape[0] * Z


@pytest.mark.parametrize("bg", [1, 2])
def test_encoder_codewords_satisfy_H_all_102(O, bg):
    rng = np.random.default_rng(bg)
    for Z in ALL_Z:
        info = rng.integers(0, 2, O.dims(bg, Z)["K"], dtype=np.uint8)
        cw = O.encode(bg, Z, info)
        assert (cw[:len(info)] == info).all()
        assert O.syndrome_weight(bg, Z, cw) == 0, (bg, Z)
        # independent check through the explicit matrix
        rows, cols = O.pcm(bg, Z)
        syn = np.zeros(rows.max() + 1, np.int64)
        np.add.at(syn, rows, cw[cols])
        assert not (syn & 1).any()


@pytest.mark.parametrize("bg", [1, 2])
def test_qc_encoder_equals_generic_gf2_solve(O, bg):
    """comm.LDPCEncoder contract (NRLDPCEncoder.m:49,158): parity is the unique GF(2) solution."""
    rng = np.random.default_rng(10 + bg)
    for Z in (2, 3, 5, 7, 9, 11, 13, 15, 16, 26):
        info = rng.integers(0, 2, (2, O.dims(bg, Z)["K"]), dtype=np.uint8)
        assert (O.encode(bg, Z, info, "qc") == O.encode(bg, Z, info, "gf2")).all()
