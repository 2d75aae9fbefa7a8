"""The C-ABI library loads on a CPU-only box and exports exactly what include/nrldpc_b200.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "nrldpc_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nrldpc_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from ldpc_3gpp_matlab_b200 import capi
    lib = capi.load()
    names = _declared()
    assert sorted(capi.SYMBOLS) == names
    for n in names:
        assert getattr(lib, n) is not None
    assert b"sm_100a" in lib.nrldpc_version()


def test_no_cpu_fallback_and_error_codes():
    """Host-only entry points work without a GPU; create() on a GPU-less box fails loudly."""
    import torch
    from ldpc_3gpp_matlab_b200 import capi
    assert capi.set_index(384) == 1 and capi.lifting_size(22, 8448) == 384
    with pytest.raises(capi.UnsupportedParameters):
        capi.set_index(17)
    with pytest.raises(capi.UnsupportedParameters):
        capi.lifting_size(22, 9000)
    r, c, s = capi.base_graph(1, 1)
    assert len(r) == 316 and (r[0], c[0], s[0]) == (0, 0, 307)
    with pytest.raises(capi.UnsupportedParameters):
        capi.Handle(3, 384)
    with pytest.raises(capi.UnsupportedParameters):
        capi.Handle(1, 17)
    if not torch.cuda.is_available():
        with pytest.raises(capi.CudaError):
            capi.Handle(1, 384)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing in the package may import, include, link or load it."""
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#include\s+\"[^\"]*oracle)|libnrldpc_oracle|orc_[a-z_]+\s*\(", re.M)
    for p in (ROOT / "ldpc_3gpp_matlab_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".h", ".inc") or p.name == "Makefile":
            hits = [m.group(0) for m in bad.finditer(p.read_text()) if "orc_decode_nms)" not in m.group(0)]
            assert not hits, (p, hits)


def test_base_graph_matches_oracle_tables(O):
    """Product table (csrc/bg_tables.inc, tools/gen_tables.py) against the oracle's OWN table (oracle/orc_tables.h,
    oracle/gen_oracle_tables.py): two parsers, two layouts, nothing shared.  The content itself is pinned to the reference by
    sha256 in tests/test_tables.py and tests/test_twin.py."""
    from ldpc_3gpp_matlab_b200 import capi
    for bg in (1, 2):
        t = O.table(bg)
        for ils in range(8):
            r, c, s = capi.base_graph(bg, ils)
            assert (r == t[:, 0]).all() and (c == t[:, 1]).all() and (s == t[:, 2 + ils]).all()
