import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"

ALL_Z = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 44, 48, 52, 56, 60,
         64, 72, 80, 88, 96, 104, 112, 120, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288, 320, 352, 384]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure), built on demand with gcc."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def golden_tables():
    return json.loads((GOLDEN / "tables.json").read_text())


@pytest.fixture(scope="session")
def golden_params():
    return json.loads((GOLDEN / "params.json").read_text())


@pytest.fixture(scope="session")
def golden_decode():
    return np.load(GOLDEN / "decode_nms.npz")


@pytest.fixture(scope="session")
def golden_decode_bp():
    return np.load(GOLDEN / "decode_bp.npz")


def make_llr(O, bg, Z, B, E, esn0, rng, filler=0, k0=0):
    """Encode random info with the oracle, QPSK + AWGN, exact LLRs in the decoder's cw layout."""
    d = O.dims(bg, Z)
    info = rng.integers(0, 2, (B, d["K"]), dtype=np.uint8)
    if filler:
        info[:, d["K"] - filler:] = 0
    cw = O.encode(bg, Z, info)
    s2 = 10 ** (-esn0 / 10)
    y = (1 - 2.0 * cw) / np.sqrt(2) + rng.normal(0, np.sqrt(s2 / 2), cw.shape)
    llr = (2 * np.sqrt(2) * y / s2).astype(np.float32)
    llr[:, :2 * Z] = 0
    llr[:, 2 * Z + E:] = 0
    if filler:
        llr[:, d["K"] - filler:d["K"]] = np.inf
    return info, llr


def make_core_pass_llr(O, bg, Z, B, n_rows, rng, mag=8.0, wrong=1048576.0):
    """Noise-free codewords (LLR magnitude `mag`, cw layout incl. the punctured 2Z zeros) in three flavours by index mod 3:
    0 clean; 1 one EXTENSION parity variable (a degree-1 column of an active extension row) wrong with the largest
    magnitude -- every core check holds on the hard decisions and one extension check never does; 2 one information bit
    weakly wrong (the decoder corrects it).  Exercises the two-stage parity-check stop: a codeword that passes the
    core-row stage must still be rejected by the extension-row stage."""
    d = O.dims(bg, Z)
    kcols = d["K"] // Z
    info = rng.integers(0, 2, (B, d["K"]), dtype=np.uint8)
    cw = O.encode(bg, Z, info)
    llr = ((1.0 - 2.0 * cw) * mag).astype(np.float32)
    llr[:, :2 * Z] = 0
    kind = np.arange(B) % 3
    for i in range(B):
        if kind[i] == 1 and n_rows > 4:
            col = kcols + 4 + int(rng.integers(0, n_rows - 4))
            v = col * Z + int(rng.integers(0, Z))
            llr[i, v] = -np.sign(llr[i, v]) * wrong
        elif kind[i] == 2:
            v = 2 * Z + int(rng.integers(0, d["K"] - 2 * Z))
            llr[i, v] = -llr[i, v] / 4
    # columns beyond the active rows were "not transmitted"
    llr[:, (kcols + n_rows) * Z:] = 0
    return info, llr, kind
