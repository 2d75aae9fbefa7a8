"""Independent numpy twin of the decoder oracles.  TEST INFRASTRUCTURE ONLY (same rules as nrldpc_oracle.c).

Purpose: oracle/nrldpc_oracle.c is the arbiter of every GPU parity test, and the reference ships no decoder
vectors (its arithmetic sits in the closed comm.LDPCDecoder, NRLDPCDecoder.m:120,265).  This module restates the
same decoders a second time with NO code, table or data structure in common with the C file or the product:

  * its own parse of Tables 5.3.2-2 / 5.3.2-3 from get_3gpp_base_graph.m:12-530 when the reference checkout is
    mounted (this container), else the committed copy of that parse, tests/golden/base_graph_rows.npz
    (written by `python -m oracle.twin --write-fixture`);
  * H built in MATRIX form by a literal transcription of get_pcm.m:7-9: for every table entry, the Z x Z block
    (r, c) of H is circshift(speye(Z), mod(V, Z), 2) -- dense numpy for small Z, scipy.sparse blocks otherwise;
  * decoders that work on H itself (neighbour lists come from H's non-zeros, not from the base graph):
      nms_layered   layered normalized min-sum, float32 (oracle A's definition)
      nms_layered_f16  the same in IEEE binary16 (oracle A16)
      bp_flooding   flooding sum-product, float64, parity-check stop (oracle B = MathWorks' documented
                    comm.LDPCDecoder algorithm)
tests/test_twin.py demands C oracle == twin bit for bit (hard decisions, iteration counts, parity flags, APP bit
patterns) for every lifting-size set, both base graphs, plus codewords at Z = 384.
"""
from __future__ import annotations

import math
import re
import sys
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
REF_TABLE = Path("/root/reference/get_3gpp_base_graph.m")
FIXTURE = _HERE.parent / "tests" / "golden" / "base_graph_rows.npz"

BG_SHAPE = {1: (46, 68, 22), 2: (42, 52, 10)}      # base rows, base columns, information columns (NRLDPC.m:414-454)
SET_LEADS = (2, 3, 5, 7, 9, 11, 13, 15)            # get_3gpp_valid_lifting_sizes.m:3-12: a * 2^j


def parse_reference_tables(path=REF_TABLE):
    """{1: int array [316, 10], 2: [197, 10]}: columns row, col, V(i_LS = 0..7), straight from the .m text."""
    text = Path(path).read_text()
    out = {}
    for g in (1, 2):
        start = text.index("table{%d}" % g)
        body = text[text.index("[", start) + 1:text.index("];", start)]
        out[g] = np.array([[int(t) for t in ln.split()] for ln in body.strip().splitlines()], dtype=np.int64)
    return out


_TABLES = None


def tables():
    global _TABLES
    if _TABLES is None:
        if REF_TABLE.exists():
            _TABLES = parse_reference_tables()
        else:
            z = np.load(FIXTURE)
            _TABLES = {1: z["bg1"].astype(np.int64), 2: z["bg2"].astype(np.int64)}
    return _TABLES


def set_index(Z):
    """get_3gpp_set_index.m:5-11 -- the set whose leading value a satisfies Z = a * 2^j."""
    for i, a in enumerate(SET_LEADS):
        z = a
        while z <= 384:
            if z == Z:
                return i
            z *= 2
    raise ValueError("Invalid lifting size")


def pcm(bg, Z):
    """get_pcm.m:1-11 on get_3gpp_base_graph(bg, i_LS): H as a dense uint8 matrix (small Z) or scipy CSR."""
    rows, cols, _ = BG_SHAPE[bg]
    tab = tables()[bg]
    ils = set_index(Z)
    if rows * Z * cols * Z <= 40_000_000:
        H = np.zeros((rows * Z, cols * Z), dtype=np.uint8)
        eye = np.eye(Z, dtype=np.uint8)
        for ent in tab:
            r, c, V = int(ent[0]), int(ent[1]), int(ent[2 + ils])
            H[r * Z:(r + 1) * Z, c * Z:(c + 1) * Z] = np.roll(eye, V % Z, axis=1)     # circshift(speye(Z), mod(V,Z), 2)
        return H
    import scipy.sparse as sp
    blocks = [[None] * cols for _ in range(rows)]      # every base row and column holds at least one entry
    for ent in tab:
        r, c, V = int(ent[0]), int(ent[1]), int(ent[2 + ils])
        s = V % Z
        perm = (np.arange(Z) + s) % Z                       # row i of the block has its one at column (i + s) mod Z
        blocks[r][c] = sp.csr_matrix((np.ones(Z, np.uint8), (np.arange(Z), perm)), shape=(Z, Z))
    return sp.bmat(blocks, format="csr", dtype=np.uint8)


class Graph:
    """Neighbour lists of H: per check the variables in ascending column order, per variable the checks in
    ascending row order.  Checks of one layer (Z consecutive rows of H) all have the same degree."""

    def __init__(self, bg, Z):
        self.bg, self.Z = bg, Z
        self.rows, self.cols, self.kcols = BG_SHAPE[bg]
        H = pcm(bg, Z)
        if isinstance(H, np.ndarray):
            rr, cc = np.nonzero(H)
        else:
            H = H.tocoo()
            order = np.lexsort((H.col, H.row))
            rr, cc = H.row[order], H.col[order]
        self.M, self.N = self.rows * Z, self.cols * Z
        deg = np.bincount(rr, minlength=self.M)
        self.check_ptr = np.concatenate(([0], np.cumsum(deg)))
        self.check_var = cc.astype(np.int64)                       # sorted by (row, col)
        self.check_deg = deg
        # per variable: positions (into check_var) of its edges, in ascending check order
        order = np.lexsort((rr, cc))
        vdeg = np.bincount(cc, minlength=self.N)
        self.var_ptr = np.concatenate(([0], np.cumsum(vdeg)))
        self.var_edge = order.astype(np.int64)
        self.var_deg = vdeg

    def layer(self, r):
        """[Z, deg] matrix of variable indices and the matching [Z, deg] edge positions of base row r."""
        Z = self.Z
        d = int(self.check_deg[r * Z])
        assert (self.check_deg[r * Z:(r + 1) * Z] == d).all()
        pos = self.check_ptr[r * Z:(r + 1) * Z, None] + np.arange(d)[None, :]
        return self.check_var[pos], pos

    def syndrome_ok(self, hard, n_rows):
        n = self.check_ptr[n_rows * self.Z]
        s = np.add.reduceat(hard[self.check_var[:n]].astype(np.int64), self.check_ptr[:n_rows * self.Z]) & 1
        return not s.any()


_GRAPHS = {}


def graph(bg, Z):
    if (bg, Z) not in _GRAPHS:
        _GRAPHS[(bg, Z)] = Graph(bg, Z)
    return _GRAPHS[(bg, Z)]


LLR_MAX = np.float32(1048576.0)


def _clamp_f32(llr):
    x = np.asarray(llr, dtype=np.float32).copy()
    x[np.isnan(x)] = LLR_MAX                                        # NaN marks filler (NRLDPCDecoder.m:224,264)
    x = np.minimum(np.maximum(x, -LLR_MAX), LLR_MAX)
    return (x + np.float32(0.0)).astype(np.float32)                 # -0 -> +0


def nms_layered(bg, Z, llr, max_iters, early_term=False, n_rows=0, alpha=0.75, deg1_shortcut=True):
    """Oracle A's definition on one codeword (llr: [cols*Z] in cw_tilde layout).  Returns hard[K], app, iters, ok.

    deg1_shortcut (oracle A revision 2): a variable that belongs to exactly ONE check row of the
    whole H (the extension parity columns) always enters that check with its channel value, t = llr, and its
    a-posteriori value llr + c is formed when read (hard decisions, soft output), never fed back.
    """
    g = graph(bg, Z)
    n_rows = n_rows or g.rows
    a32 = np.float32(alpha)
    app = _clamp_f32(llr)
    chan = app.copy()
    c2v = np.zeros(len(g.check_var), dtype=np.float32)
    single = g.var_deg == 1
    it, ok = 0, False
    while it < max_iters:
        for r in range(n_rows):
            var, pos = g.layer(r)
            if deg1_shortcut:
                one = single[var]
                t = np.where(one, chan[var], app[var] - c2v[pos]).astype(np.float32)
            else:
                t = (app[var] - c2v[pos]).astype(np.float32)
            mag = np.abs(t)
            arg = np.argmin(mag, axis=1)                             # first index wins ties
            m1 = mag[np.arange(Z), arg]
            rest = mag.copy()
            rest[np.arange(Z), arg] = np.inf
            m2 = rest.min(axis=1)
            neg = np.signbit(t)
            row_sign = np.logical_xor.reduce(neg, axis=1)
            use2 = np.arange(t.shape[1])[None, :] == arg[:, None]
            m = np.where(use2, (a32 * m2)[:, None], (a32 * m1)[:, None]).astype(np.float32)
            c = np.where(np.logical_xor(row_sign[:, None], neg), -m, m).astype(np.float32)
            c2v[pos] = c
            new = (t + c).astype(np.float32)
            app[var] = new                                           # degree-1 variables: llr + c, recomputed from llr each time
        it += 1
        if early_term or it == max_iters:
            ok = g.syndrome_ok(app < 0, n_rows)
            if early_term and ok:
                break
    hard = (app[:g.kcols * Z] < 0).astype(np.uint8)
    return hard, app, it, ok


def _h(x):
    """Round a float64 array once to binary16 (round to nearest even) and widen back: exact single rounding."""
    return np.asarray(x, dtype=np.float64).astype(np.float16).astype(np.float64)


def nms_layered_f16(bg, Z, llr, max_iters, early_term=False, n_rows=0, alpha=0.75, deg1_shortcut=True):
    """Oracle A16's definition (binary16 after every operation, minima capped at 2048 before the scaling)."""
    g = graph(bg, Z)
    n_rows = n_rows or g.rows
    ah = float(np.float16(np.float32(alpha)))
    x = np.asarray(llr, dtype=np.float32).copy()
    x[np.isnan(x)] = 2048.0
    x = np.minimum(np.maximum(x, np.float32(-2048.0)), np.float32(2048.0)) + np.float32(0.0)
    app = _h(x)
    chan = app.copy()
    c2v = np.zeros(len(g.check_var), dtype=np.float64)
    single = g.var_deg == 1
    it, ok = 0, False
    while it < max_iters:
        for r in range(n_rows):
            var, pos = g.layer(r)
            t = _h(app[var] - c2v[pos])
            if deg1_shortcut:
                t = np.where(single[var], chan[var], t)
            mag = np.abs(t)
            m1 = mag.min(axis=1)
            arg = np.argmin(mag, axis=1)
            rest = mag.copy()
            rest[np.arange(Z), arg] = np.inf
            m2 = rest.min(axis=1)
            neg = np.signbit(t)
            row_sign = np.logical_xor.reduce(neg, axis=1)
            m1s, m2s = _h(ah * np.minimum(m1, 2048.0)), _h(ah * np.minimum(m2, 2048.0))
            m = np.where(mag == m1[:, None], m2s[:, None], m1s[:, None])          # arg-min by value (ties: m2 == m1)
            c = np.where(np.logical_xor(row_sign[:, None], neg), -m, m)
            c2v[pos] = c
            app[var] = _h(t + c)
        it += 1
        if early_term or it == max_iters:
            ok = g.syndrome_ok(app < 0, n_rows)
            if early_term and ok:
                break
    hard = (app[:g.kcols * Z] < 0).astype(np.uint8)
    return hard, app.astype(np.float32), it, ok


_tanh = np.frompyfunc(math.tanh, 1, 1)      # libm, like the C restatement (numpy's SIMD tanh may differ by an ulp)
_atanh = np.frompyfunc(math.atanh, 1, 1)


def bp_flooding(bg, Z, llr, max_iters, early_term=True, n_rows=0):
    """Oracle B's definition on one codeword: flooding sum-product in float64 over H (MathWorks' documented
    comm.LDPCDecoder algorithm, stopping rule of NRLDPCDecoder.m:120).  Leave-one-out products by prefix / suffix
    products, atanh argument clipped to +-(1 - 2^-53), variable sums in ascending check order."""
    g = graph(bg, Z)
    n_rows = n_rows or g.rows
    L = np.asarray(llr, dtype=np.float64)
    nE = int(g.check_ptr[n_rows * Z])
    q = L[g.check_var[:nE]].copy()
    rmsg = np.zeros(nE)
    lim = 1.0 - 2.0 ** -53
    Q = L.copy()
    it, ok = 0, False
    # edges of each variable restricted to the active rows, ascending check order
    v_edges = [g.var_edge[g.var_ptr[v]:g.var_ptr[v + 1]] for v in range(g.N)]
    maxd = int(g.var_deg.max())
    pad = np.full((g.N, maxd), -1, dtype=np.int64)
    for v, e in enumerate(v_edges):
        e = e[e < nE]
        pad[v, :len(e)] = e
    while it < max_iters:
        for r in range(n_rows):
            _, pos = g.layer(r)
            th = _tanh(0.5 * q[pos]).astype(np.float64)
            d = th.shape[1]
            pre = np.ones((Z, d + 1))
            suf = np.ones((Z, d + 1))
            for k in range(d):
                pre[:, k + 1] = pre[:, k] * th[:, k]
            for k in range(d - 1, -1, -1):
                suf[:, k] = suf[:, k + 1] * th[:, k]
            x = np.clip(pre[:, :d] * suf[:, 1:], -lim, lim)
            rmsg[pos] = 2.0 * _atanh(x).astype(np.float64)
        Q = L.copy()
        for k in range(maxd):
            e = pad[:, k]
            has = e >= 0
            Q[has] = Q[has] + rmsg[e[has]]
        q = Q[g.check_var[:nE]] - rmsg
        it += 1
        ok = g.syndrome_ok(Q < 0, n_rows)
        if ok and early_term:
            break
    hard = (Q[:g.kcols * Z] < 0).astype(np.uint8)
    return hard, Q, it, ok


def encode(bg, Z, info):
    """Systematic encoding by solving H_p p = H_s s over GF(2) on H itself (comm.LDPCEncoder's contract,
    NRLDPCEncoder.m:49,158).  Dense elimination: small Z only."""
    g = graph(bg, Z)
    H = pcm(bg, Z)
    H = np.asarray(H if isinstance(H, np.ndarray) else H.todense(), dtype=np.uint8)
    K = g.kcols * Z
    rhs = (H[:, :K].astype(np.int64) @ np.asarray(info, dtype=np.int64)) & 1
    A = np.concatenate([H[:, K:], rhs[:, None].astype(np.uint8)], axis=1)
    M = A.shape[0]
    for c in range(M):
        piv = c + int(np.argmax(A[c:, c]))
        if not A[piv, c]:
            raise ValueError("parity part singular")
        if piv != c:
            A[[c, piv]] = A[[piv, c]]
        rows = np.nonzero(A[:, c])[0]
        rows = rows[rows != c]
        A[rows] ^= A[c]
    return np.concatenate([np.asarray(info, dtype=np.uint8), A[:, -1]])


if __name__ == "__main__":
    if "--write-fixture" in sys.argv:
        t = parse_reference_tables()
        np.savez_compressed(FIXTURE, bg1=t[1].astype(np.int16), bg2=t[2].astype(np.int16))
        print("wrote", FIXTURE, t[1].shape, t[2].shape)
