/*
 * nrldpc_oracle.c -- CPU ORACLE for the NR LDPC hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The shipped path
 * (ldpc_3gpp_matlab_b200/csrc) never links, imports or calls anything in oracle/.
 *
 * PARITY STATUS (read SURVEY.md section 8c):
 *   - structure (tables, lifting, parameters, rate matching, interleaving, encoding) restates
 *     reference source text and is pinned by algebra (H*c = 0 with an invertible parity part
 *     makes encoder output unique) and by the Appendix-A fixtures in tests/golden/.
 *   - decoding: PARITY UNPINNED.  The reference delegates the arithmetic to MathWorks'
 *     closed-source comm.LDPCDecoder (NRLDPCDecoder.m:120,265; toolbox version not pinned
 *     anywhere in the reference) and ships no decoder test or golden vector.  Two decoders
 *     are restated here:
 *       orc_decode_nms  -- oracle A: layered normalized min-sum, float32, the algorithm the
 *                          CUDA kernel implements; defined by us (no reference text exists).
 *       orc_decode_nms_f16 -- oracle A16: oracle A in IEEE binary16 arithmetic, the target of the
 *                          packed-half kernel variant.
 *       orc_decode_bp   -- oracle B: flooding sum-product, float64, stop when all parity
 *                          checks are satisfied: MathWorks' published algorithm for
 *                          comm.LDPCDecoder as configured at NRLDPCDecoder.m:120.
 *
 * All file:line citations are relative to /root/reference.
 * Build: see oracle/Makefile (gcc -O3 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "orc_tables.h"   /* the oracle's OWN table (oracle/gen_oracle_tables.py); nothing is shared with the product */

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Table 5.3.2-1 -- get_3gpp_valid_lifting_sizes.m:3-12, get_3gpp_set_index.m:5-11,
 * get_3gpp_lifting_size.m:5-16
 * ---------------------------------------------------------------------------------------- */
static const int orc_set_a[8] = {2, 3, 5, 7, 9, 11, 13, 15};
static const int orc_set_n[8] = {8, 8, 7, 6, 6, 6, 5, 5}; /* a * 2^j, j < n */

ORC_API int orc_set_index(int Z) {
    for (int s = 0; s < 8; ++s)
        for (int j = 0; j < orc_set_n[s]; ++j)
            if ((orc_set_a[s] << j) == Z) return s;
    return -1; /* 'Invalid lifting size.' get_3gpp_set_index.m:10 */
}

ORC_API int orc_lifting_size(int K_b, int K_prime) {
    int best = 0;
    for (int s = 0; s < 8; ++s)
        for (int j = 0; j < orc_set_n[s]; ++j) {
            int Z = orc_set_a[s] << j;
            if (K_b * Z >= K_prime && (best == 0 || Z < best)) best = Z;
        }
    return best ? best : -1; /* 'Invalid block length.' get_3gpp_lifting_size.m:15 */
}

static void bg_dims(int bg, int *rows, int *cols, int *kcols, int *edges) {
    if (bg == 1) { *rows = 46; *cols = 68; *kcols = 22; *edges = ORC_BG1_ENTRIES; }
    else         { *rows = 42; *cols = 52; *kcols = 10; *edges = ORC_BG2_ENTRIES; }
}
/* Accessors over the table rows {row, col, V(i_LS = 0..7)} (get_3gpp_base_graph.m:13-328, :333-529), unpacked once. */
static unsigned char orc_row_[2][ORC_BG1_ENTRIES], orc_col_[2][ORC_BG1_ENTRIES];
static unsigned short orc_shift_[2][8][ORC_BG1_ENTRIES];
static int orc_unpacked_ = 0;
static void orc_unpack(void) {
    if (__atomic_load_n(&orc_unpacked_, __ATOMIC_ACQUIRE)) return;
#pragma omp critical(orc_unpack_tables)
    if (!orc_unpacked_) {
        for (int g = 0; g < 2; ++g) {
            const int n = g == 0 ? ORC_BG1_ENTRIES : ORC_BG2_ENTRIES;
            for (int e = 0; e < n; ++e) {
                const short *t = g == 0 ? orc_bg1_entries[e] : orc_bg2_entries[e];
                orc_row_[g][e] = (unsigned char)t[0];
                orc_col_[g][e] = (unsigned char)t[1];
                for (int s = 0; s < 8; ++s) orc_shift_[g][s][e] = (unsigned short)t[2 + s];
            }
        }
        __atomic_store_n(&orc_unpacked_, 1, __ATOMIC_RELEASE);
    }
}
static const unsigned char *bg_row(int bg) { orc_unpack(); return orc_row_[bg == 1 ? 0 : 1]; }
static const unsigned char *bg_col(int bg) { orc_unpack(); return orc_col_[bg == 1 ? 0 : 1]; }
static const unsigned short *bg_shift(int bg, int ils) { orc_unpack(); return orc_shift_[bg == 1 ? 0 : 1][ils]; }

/* Raw table dump for the sha256 guards. out: edges x 10 ints. */
ORC_API int orc_table(int bg, int *out) {
    int R, C, Kc, E;
    if (bg != 1 && bg != 2) return -1;
    bg_dims(bg, &R, &C, &Kc, &E);
    for (int e = 0; e < E; ++e) {
        out[10 * e] = bg_row(bg)[e];
        out[10 * e + 1] = bg_col(bg)[e];
        for (int s = 0; s < 8; ++s) out[10 * e + 2 + s] = bg_shift(bg, s)[e];
    }
    return E;
}

/* ------------------------------------------------------------------------------------------
 * NRLDPC.m:297-543 -- every Dependent getter, evaluated once into a struct.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    /* inputs */
    int BG, A, G, Q_m, N_L, rv_id, I_LBRM, TBS_LBRM;
    /* derived */
    int tb_L, B, K_cb, C, cb_L, B_prime, K_prime, K_b, Z_c, i_LS, K, N, N_ref, N_cb, k_0;
    int E_r[64];
    int status; /* 0 ok, -1 UnsupportedParameters */
} orc_params_t;

ORC_API int orc_params(orc_params_t *p) {
    p->status = -1;
    if (p->BG < 1 || p->BG > 2 || p->A < 0 || p->G < 0) return -1;                 /* NRLDPC.m:240-275 */
    if (!(p->Q_m == 1 || p->Q_m == 2 || p->Q_m == 4 || p->Q_m == 6 || p->Q_m == 8)) return -1; /* :278-283 */
    if (p->N_L < 1 || p->N_L > 4 || p->rv_id < 0 || p->rv_id > 3) return -1;      /* :265-294 */
    p->tb_L = p->A > 3824 ? 24 : 16;                                              /* :297-303 */
    p->B = p->A + p->tb_L;                                                        /* :316-318 */
    p->K_cb = p->BG == 1 ? 8448 : 3840;                                           /* :321-331 */
    p->cb_L = p->B <= p->K_cb ? 0 : 24;                                           /* :347-364 */
    p->C = p->B <= p->K_cb ? 1 : (p->B + (p->K_cb - p->cb_L) - 1) / (p->K_cb - p->cb_L); /* :334-344 */
    if (p->C > 64) return -1;
    p->B_prime = p->B <= p->K_cb ? p->B : p->B + p->C * p->cb_L;                  /* :366-377 */
    if (p->B_prime % p->C != 0) return -1;                                        /* :552-554 */
    p->K_prime = p->B_prime / p->C;                                               /* :380-382 */
    if (p->BG == 1) p->K_b = 22;                                                  /* :385-406 */
    else p->K_b = p->K_prime > 640 ? 10 : p->K_prime > 560 ? 9 : p->K_prime > 192 ? 8 : 6;
    p->Z_c = orc_lifting_size(p->K_b, p->K_prime);                                /* :409-411 */
    if (p->Z_c < 0) return -1;
    p->K = p->Z_c * (p->BG == 1 ? 22 : 10);                                       /* :414-425 */
    p->i_LS = orc_set_index(p->Z_c);                                              /* :428-430 */
    p->N = p->Z_c * (p->BG == 1 ? 66 : 50);                                       /* :443-454 */
    p->N_ref = (int)floor((double)p->TBS_LBRM / ((double)p->C * (2.0 / 3.0)));    /* :457-460 */
    p->N_cb = p->I_LBRM == 0 ? p->N : (p->N < p->N_ref ? p->N : p->N_ref);        /* :463-469 */
    if (p->G % (p->Q_m * p->N_L) != 0) return -1;                                 /* :556-558 */
    {                                                                             /* :485-507, CBGTI = [] */
        int Cp = p->C, q = p->N_L * p->Q_m, j = 0;
        for (int r = 0; r < p->C; ++r) {
            if (j <= Cp - ((p->G / q) % Cp) - 1) p->E_r[r] = q * (p->G / (q * Cp));
            else p->E_r[r] = q * ((p->G + q * Cp - 1) / (q * Cp));
            ++j;
        }
    }
    {                                                                             /* :510-543 */
        static const int num1[4] = {0, 17, 33, 56}, num2[4] = {0, 13, 25, 43};
        int den = p->BG == 1 ? 66 : 50, num = (p->BG == 1 ? num1 : num2)[p->rv_id];
        p->k_0 = (int)(((long long)num * p->N_cb) / ((long long)den * p->Z_c)) * p->Z_c;
    }
    p->status = 0;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * get_pcm.m:1-11 -- lifted parity-check matrix as a list of ones.
 * Block (r,c) = circshift(speye(Z), mod(V,Z), 2): check r*Z+i has a one at column
 * c*Z + (i + V mod Z) mod Z.   rows_out/cols_out: edges*Z entries, check-major within edge.
 * ---------------------------------------------------------------------------------------- */
ORC_API long orc_pcm(int bg, int Z, int *rows_out, int *cols_out) {
    int R, C, Kc, E, ils = orc_set_index(Z);
    if ((bg != 1 && bg != 2) || ils < 0) return -1;
    bg_dims(bg, &R, &C, &Kc, &E);
    long n = 0;
    for (int e = 0; e < E; ++e) {
        int s = bg_shift(bg, ils)[e] % Z;
        for (int i = 0; i < Z; ++i, ++n) {
            rows_out[n] = bg_row(bg)[e] * Z + i;
            cols_out[n] = bg_col(bg)[e] * Z + (i + s) % Z;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Encoder oracle 1: the comm.LDPCEncoder contract (NRLDPCEncoder.m:49,158): systematic
 * cw = [c ; p] with H*cw = 0, found by generic Gauss-Jordan elimination over GF(2) on the
 * last M = rows(H) columns.  O(M^3/64): small Z only.
 * Returns 0, or -2 if the parity part is singular.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_encode_gf2(int bg, int Z, const uint8_t *info, uint8_t *cw) {
    int R, C, Kc, E, ils = orc_set_index(Z);
    if ((bg != 1 && bg != 2) || ils < 0) return -1;
    bg_dims(bg, &R, &C, &Kc, &E);
    const int M = R * Z, K = Kc * Z, W = (M + 1 + 63) / 64;
    uint64_t *a = (uint64_t *)calloc((size_t)M * W, 8); /* [Hp | syndrome] */
    if (!a) return -4;
    for (int e = 0; e < E; ++e) {
        int s = bg_shift(bg, ils)[e] % Z, r0 = bg_row(bg)[e] * Z, c0 = bg_col(bg)[e] * Z;
        for (int i = 0; i < Z; ++i) {
            int row = r0 + i, col = c0 + (i + s) % Z;
            if (col < K) { if (info[col] & 1) a[(size_t)row * W + M / 64] ^= 1ull << (M % 64); }
            else a[(size_t)row * W + (col - K) / 64] ^= 1ull << ((col - K) % 64);
        }
    }
    for (int c = 0; c < M; ++c) {
        int piv = -1;
        for (int r = c; r < M; ++r) if (a[(size_t)r * W + c / 64] >> (c % 64) & 1) { piv = r; break; }
        if (piv < 0) { free(a); return -2; }
        if (piv != c) for (int w = 0; w < W; ++w) { uint64_t t = a[(size_t)c * W + w]; a[(size_t)c * W + w] = a[(size_t)piv * W + w]; a[(size_t)piv * W + w] = t; }
        for (int r = 0; r < M; ++r)
            if (r != c && (a[(size_t)r * W + c / 64] >> (c % 64) & 1))
                for (int w = c / 64; w < W; ++w) a[(size_t)r * W + w] ^= a[(size_t)c * W + w];
    }
    for (int k = 0; k < K; ++k) cw[k] = info[k] & 1;
    for (int r = 0; r < M; ++r) cw[K + r] = (uint8_t)(a[(size_t)r * W + M / 64] >> (M % 64) & 1);
    free(a);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Encoder oracle 2: quasi-cyclic double-diagonal back-substitution (same unique answer).
 * The four core rows (0..3) over the four core parity columns Kc..Kc+3 hold a dual diagonal
 * plus three entries in column Kc (get_3gpp_base_graph.m, rows 0-3); summing the four rows
 * cancels the diagonal and leaves one rotated copy of p0.  Structure is read from the table,
 * not hard-coded.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_encode_qc(int bg, int Z, const uint8_t *info, uint8_t *cw) {
    int R, C, Kc, E, ils = orc_set_index(Z);
    if ((bg != 1 && bg != 2) || ils < 0) return -1;
    bg_dims(bg, &R, &C, &Kc, &E);
    const unsigned char *er = bg_row(bg), *ec = bg_col(bg);
    const unsigned short *es = bg_shift(bg, ils);
    const int K = Kc * Z;
    uint8_t *lam = (uint8_t *)calloc((size_t)4 * Z, 1), *p = cw + K;
    if (!lam) return -4;
    for (int k = 0; k < K; ++k) cw[k] = info[k] & 1;
    memset(p, 0, (size_t)R * Z);
    /* lambda_r = sum over systematic columns, r = 0..3 */
    for (int e = 0; e < E; ++e) {
        if (er[e] > 3 || ec[e] >= Kc) continue;
        int s = es[e] % Z;
        for (int i = 0; i < Z; ++i) lam[er[e] * Z + i] ^= cw[ec[e] * Z + (i + s) % Z];
    }
    /* shifts of column Kc in rows 0..3 (-1 = absent) */
    int s0[4] = {-1, -1, -1, -1};
    for (int e = 0; e < E; ++e) if (ec[e] == Kc && er[e] < 4) s0[er[e]] = es[e] % Z;
    /* the shift that appears once among the three entries survives the row sum */
    int vals[3], nv = 0, delta = 0;
    for (int r = 0; r < 4; ++r) if (s0[r] >= 0) vals[nv++] = s0[r];
    if (nv != 3) { free(lam); return -2; }
    delta = vals[0] == vals[1] ? vals[2] : (vals[0] == vals[2] ? vals[1] : vals[0]);
    for (int i = 0; i < Z; ++i)
        p[(i + delta) % Z] = lam[i] ^ lam[Z + i] ^ lam[2 * Z + i] ^ lam[3 * Z + i];
    /* forward substitution down the dual diagonal: row r gives p_{r+1} */
    for (int r = 0; r < 3; ++r)
        for (int i = 0; i < Z; ++i) {
            uint8_t v = lam[r * Z + i];
            if (s0[r] >= 0) v ^= p[(i + s0[r]) % Z];
            if (r > 0) v ^= p[r * Z + i];
            p[(r + 1) * Z + i] = v;
        }
    /* extension rows: p_r = sum of every other entry in row r (its own column is identity) */
    for (int e = 0; e < E; ++e) {
        int r = er[e];
        if (r < 4 || ec[e] == Kc + r) continue;
        int s = es[e] % Z;
        for (int i = 0; i < Z; ++i) p[r * Z + i] ^= cw[ec[e] * Z + (i + s) % Z];
    }
    free(lam);
    return 0;
}

/* H * cw over GF(2); returns the number of unsatisfied checks among the first n_rows base rows. */
ORC_API long orc_syndrome_weight(int bg, int Z, int n_rows, const uint8_t *cw) {
    int R, C, Kc, E, ils = orc_set_index(Z);
    if ((bg != 1 && bg != 2) || ils < 0) return -1;
    bg_dims(bg, &R, &C, &Kc, &E);
    if (n_rows <= 0 || n_rows > R) n_rows = R;
    uint8_t *syn = (uint8_t *)calloc((size_t)R * Z, 1);
    long w = 0;
    for (int e = 0; e < E; ++e) {
        int s = bg_shift(bg, ils)[e] % Z, r = bg_row(bg)[e], c = bg_col(bg)[e];
        if (r >= n_rows) continue;
        for (int i = 0; i < Z; ++i) syn[r * Z + i] ^= cw[c * Z + (i + s) % Z] & 1;
    }
    for (int i = 0; i < n_rows * Z; ++i) w += syn[i];
    free(syn);
    return w;
}

/* ------------------------------------------------------------------------------------------
 * Oracle A -- layered normalized min-sum, float32.  THE bit-exact target of the CUDA kernel.
 *
 * Boundary follows NRLDPCDecoder.m:262-266: input is the full cw_tilde layout, (cols*Z) LLRs,
 * positive => bit 0, 2Z leading zeros for the punctured columns, +inf for filler bits; output
 * is the hard decision of the first K positions (comm.LDPCDecoder's information-part output).
 *
 * Definition (ours -- the reference has no text for it):
 *   input clamp  x -> min(max(x,-LLR_MAX), LLR_MAX), NaN -> +LLR_MAX (NaN marks filler upstream,
 *                NRLDPCDecoder.m:264), LLR_MAX = 2^20: keeps inf-inf out of the recursion; -0 -> +0.
 *   schedule     base rows 0..n_rows-1 in order, all Z checks of a row independent.
 *   per check    t_e   = app[v_e] - c_e            (c_e = previous message, +0 in iteration 1)
 *                REVISION 2 (default): a variable that belongs to ONE check of the whole H -- the extension parity
 *                columns Kc+4 .. C-1, TS 38.212 tables -- has nothing but its channel value and this check's own
 *                message in its a-posteriori value, so in exact arithmetic app - c_e IS the channel value; revision 2
 *                takes t_e = (clamped) channel value for those edges in every iteration instead of re-deriving it
 *                through two float32 roundings.  Their a-posteriori value app = t_e + c_e' is still formed (hard
 *                decisions, syndrome, soft output) but never fed back.  Revision 1 (orc_set_nms_revision(1)) is the
 *                round-1 definition without this rule; tests/test_oracle.py compares the two.
 *                m1,m2 = two smallest |t_e| (strict '<' updates, first index wins ties)
 *                c_e'  = sgn * (e == argmin ? alpha*m2 : alpha*m1), sgn = XOR of sign BITS of the
 *                        other t's; every product/sum individually rounded (no FMA)
 *                app[v_e] = t_e + c_e'
 *   stop         after each iteration hard = (app < 0); if early_term and every check of the
 *                active rows is satisfied, stop.  iters_out = iterations executed.
 * ---------------------------------------------------------------------------------------- */
#define ORC_LLR_MAX 1048576.0f

static int orc_nms_revision_ = 2;
ORC_API void orc_set_nms_revision(int rev) { orc_nms_revision_ = rev == 1 ? 1 : 2; }
ORC_API int orc_get_nms_revision(void) { return orc_nms_revision_; }

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* vidx[e*Z + z] = variable index touched by check z of base-graph edge e (get_pcm.m:8) */
static int *build_vidx(int bg, int Z, int ils) {
    int R, C, Kc, E;
    bg_dims(bg, &R, &C, &Kc, &E);
    int *v = (int *)malloc(sizeof(int) * (size_t)E * Z);
    for (int e = 0; e < E; ++e)
        for (int z = 0; z < Z; ++z) v[(size_t)e * Z + z] = bg_col(bg)[e] * Z + (z + bg_shift(bg, ils)[e] % Z) % Z;
    return v;
}

/* First block column from which on every column of the base graph holds exactly one entry (22+4 = 26 for BG1, 10+4 = 14 for
 * BG2): counted from the table, not assumed. */
static int first_single_column(int bg) {
    int R, C, Kc, E, cnt[68] = {0};
    bg_dims(bg, &R, &C, &Kc, &E);
    for (int e = 0; e < E; ++e) ++cnt[bg_col(bg)[e]];
    int c = C;
    while (c > 0 && cnt[c - 1] == 1) --c;
    return c;
}

static int decode_nms_one(int bg, int Z, int ils, int n_rows, int max_iters, int early_term, float alpha,
                          const int *vidx, const float *llr, uint8_t *hard_info, float *app_out, uint8_t *parity_ok) {
    int R, C, Kc, E;
    bg_dims(bg, &R, &C, &Kc, &E);
    const unsigned char *er = bg_row(bg);
    int row_start[47];
    for (int r = 0, e = 0; r <= R; ++r) { while (e < E && er[e] < r) ++e; row_start[r] = e; }
    const int nV = C * Z;
    float *app = (float *)malloc(sizeof(float) * nV);
    float *chan = (float *)malloc(sizeof(float) * nV);
    float *c2v = (float *)calloc((size_t)E * Z, sizeof(float)); /* uncompressed; same values */
    for (int i = 0; i < nV; ++i) {
        float x = llr[i];
        app[i] = (x != x) ? ORC_LLR_MAX : (x > ORC_LLR_MAX ? ORC_LLR_MAX : (x < -ORC_LLR_MAX ? -ORC_LLR_MAX : x));
        app[i] += 0.0f; /* -0 -> +0: from here on no APP value is ever -0, so (app < 0) == sign bit */
        chan[i] = app[i];
    }
    /* column degrees over the WHOLE H: the degree-1 variables are the extension parity columns */
    const int single_from = orc_nms_revision_ >= 2 ? first_single_column(bg) * Z : nV;
    int it = 0, ok = 0;
    while (it < max_iters) {
        for (int r = 0; r < n_rows; ++r) {
            const int e0 = row_start[r], deg = row_start[r + 1] - e0;
            for (int z = 0; z < Z; ++z) {
                float t[32]; int v[32];
                float m1 = INFINITY, m2 = INFINITY; int arg = 0; uint32_t sgn = 0;
                for (int k = 0; k < deg; ++k) {
                    int e = e0 + k;
                    v[k] = vidx[(size_t)e * Z + z];
                    t[k] = v[k] >= single_from ? chan[v[k]] : app[v[k]] - c2v[(size_t)e * Z + z];
                    float a = fabsf(t[k]);
                    if (a < m1) { m2 = m1; m1 = a; arg = k; } else if (a < m2) { m2 = a; }
                    sgn ^= f2u(t[k]) & 0x80000000u;
                }
                const float m1s = alpha * m1, m2s = alpha * m2;
                for (int k = 0; k < deg; ++k) {
                    float mag = (k == arg) ? m2s : m1s;
                    float c = u2f(f2u(mag) ^ (sgn ^ (f2u(t[k]) & 0x80000000u)));
                    c2v[(size_t)(e0 + k) * Z + z] = c;
                    app[v[k]] = t[k] + c;
                }
            }
        }
        ++it;
        if (early_term || it == max_iters) {
            ok = 1;
            for (int r = 0; r < n_rows && ok; ++r)
                for (int z = 0; z < Z; ++z) {
                    int par = 0;
                    for (int e = row_start[r]; e < row_start[r + 1]; ++e)
                        par ^= app[vidx[(size_t)e * Z + z]] < 0.0f;
                    if (par) { ok = 0; break; }
                }
            if (early_term && ok) break;
        }
    }
    for (int k = 0; k < Kc * Z; ++k) hard_info[k] = app[k] < 0.0f;
    if (app_out) memcpy(app_out, app, sizeof(float) * nV);
    if (parity_ok) *parity_ok = (uint8_t)ok;
    free(app); free(chan); free(c2v);
    return it;
}

ORC_API int orc_decode_nms(int bg, int Z, int n_rows, int max_iters, int early_term, float alpha,
                           const float *llr, long batch, uint8_t *hard_info, float *app_out,
                           int32_t *iters_out, uint8_t *parity_ok, int n_threads) {
    int R, C, Kc, E, ils = orc_set_index(Z);
    if ((bg != 1 && bg != 2) || ils < 0 || max_iters < 1) return -1;
    bg_dims(bg, &R, &C, &Kc, &E);
    if (n_rows <= 0 || n_rows > R) n_rows = R;
    if (n_rows < 4) return -1;
    if (n_threads < 1) n_threads = 1;
    const long nV = (long)C * Z, K = (long)Kc * Z;
    int *vidx = build_vidx(bg, Z, ils);
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (long b = 0; b < batch; ++b) {
        int it = decode_nms_one(bg, Z, ils, n_rows, max_iters, early_term, alpha, vidx, llr + b * nV,
                                hard_info + b * K, app_out ? app_out + b * nV : NULL,
                                parity_ok ? parity_ok + b : NULL);
        if (iters_out) iters_out[b] = it;
    }
    free(vidx);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Oracle A16 -- the same layered normalized min-sum in IEEE binary16 arithmetic: the bit-exact
 * target of the packed-half kernel (nrldpc_cfg.llr_dtype = NRLDPC_F16X2).  Ours, like oracle A.
 *   input        x -> fp16(min(max(x,-2048),2048)), NaN -> +2048, -0 -> +0, round to nearest even
 *   per check    t_e = fp16(app - c_e);  m1, m2 = two smallest |t_e|;
 *                c_e' = sgn_e * fp16(alpha_h * min((|t_e| == m1 ? m2 : m1), 2048)), alpha_h = fp16(alpha);
 *                app = fp16(t_e + c_e')
 * Every operation is computed exactly in double (sums/products of two binary16 values are exact
 * there) and rounded once, which is what HADD2 / HMUL2 do.  Values are kept as bit patterns.
 * ---------------------------------------------------------------------------------------- */
static double h2d(uint16_t h) {
    const int s = h >> 15, e = (h >> 10) & 31, m = h & 1023;
    double v;
    if (e == 0) v = ldexp((double)m, -24);
    else if (e == 31) v = m ? NAN : INFINITY;
    else v = ldexp((double)(m + 1024), e - 25);
    return s ? -v : v;
}

static uint16_t d2h(double x) {
    uint64_t b; memcpy(&b, &x, 8);
    const uint16_t sign = (uint16_t)((b >> 48) & 0x8000u);
    const double ax = fabs(x);
    if (ax != ax) return (uint16_t)(sign | 0x7e00u);
    if (ax >= 65520.0) return (uint16_t)(sign | 0x7c00u);       /* >= halfway to 2^16: overflows to inf */
    if (ax == 0.0) return sign;
    int e = ilogb(ax);
    if (e < -14) e = -14;                                       /* subnormal range: fixed quantum 2^-24 */
    const double q = rint(ldexp(ax, 10 - e));                   /* RNE (default rounding mode), exact scaling */
    double v = ldexp(q, e - 10);
    if (v == 0.0) return sign;
    int ev = ilogb(v);
    if (ev < -14) return (uint16_t)(sign | (uint16_t)ldexp(v, 24));
    const int m = (int)ldexp(v, 10 - ev) - 1024;
    return (uint16_t)(sign | ((ev + 15) << 10) | m);
}

/* test hooks: the binary16 rounding / widening used by oracle A16 (checked against numpy.float16) */
ORC_API void orc_f16_round(const double *x, long n, uint16_t *out) { for (long i = 0; i < n; ++i) out[i] = d2h(x[i]); }
ORC_API void orc_f16_widen(const uint16_t *h, long n, double *out) { for (long i = 0; i < n; ++i) out[i] = h2d(h[i]); }

#define ORC_H_LLR_MAX 2048.0f
#define ORC_H_MSG_CAP 0x6800u /* 2048 */

static int decode_nms_f16_one(int bg, int Z, int ils, int n_rows, int max_iters, int early_term, float alpha,
                              const int *vidx, const float *llr, uint8_t *hard_info, float *app_out, uint8_t *parity_ok) {
    int R, C, Kc, E;
    bg_dims(bg, &R, &C, &Kc, &E);
    const unsigned char *er = bg_row(bg);
    int row_start[47];
    for (int r = 0, e = 0; r <= R; ++r) { while (e < E && er[e] < r) ++e; row_start[r] = e; }
    const int nV = C * Z;
    uint16_t *app = (uint16_t *)malloc(sizeof(uint16_t) * nV);
    uint16_t *chan = (uint16_t *)malloc(sizeof(uint16_t) * nV);
    uint16_t *c2v = (uint16_t *)calloc((size_t)E * Z, sizeof(uint16_t));
    const double alpha_h = h2d(d2h((double)alpha));
    for (int i = 0; i < nV; ++i) {
        float x = llr[i];
        x = (x != x) ? ORC_H_LLR_MAX : (x > ORC_H_LLR_MAX ? ORC_H_LLR_MAX : (x < -ORC_H_LLR_MAX ? -ORC_H_LLR_MAX : x));
        x += 0.0f;
        app[i] = d2h((double)x);
        chan[i] = app[i];
    }
    const int single_from = orc_nms_revision_ >= 2 ? first_single_column(bg) * Z : nV;   /* revision 2, see oracle A */
    int it = 0, ok = 0;
    while (it < max_iters) {
        for (int r = 0; r < n_rows; ++r) {
            const int e0 = row_start[r], deg = row_start[r + 1] - e0;
            for (int z = 0; z < Z; ++z) {
                uint16_t t[32]; int v[32];
                uint16_t m1 = 0x7c00u, m2 = 0x7c00u, sgn = 0; /* +inf; non-negative patterns order like integers */
                for (int k = 0; k < deg; ++k) {
                    const int e = e0 + k;
                    v[k] = vidx[(size_t)e * Z + z];
                    t[k] = v[k] >= single_from ? chan[v[k]] : d2h(h2d(app[v[k]]) - h2d(c2v[(size_t)e * Z + z]));
                    const uint16_t a = t[k] & 0x7fffu;
                    if (a < m1) { m2 = m1; m1 = a; } else if (a < m2) { m2 = a; }
                    sgn ^= t[k] & 0x8000u;
                }
                const uint16_t m1c = m1 < ORC_H_MSG_CAP ? m1 : ORC_H_MSG_CAP, m2c = m2 < ORC_H_MSG_CAP ? m2 : ORC_H_MSG_CAP;
                const uint16_t m1s = d2h(alpha_h * h2d(m1c)), m2s = d2h(alpha_h * h2d(m2c));
                for (int k = 0; k < deg; ++k) {
                    const uint16_t mag = ((t[k] & 0x7fffu) == m1) ? m2s : m1s;
                    const uint16_t c = (uint16_t)(mag ^ sgn ^ (t[k] & 0x8000u));
                    c2v[(size_t)(e0 + k) * Z + z] = c;
                    app[v[k]] = d2h(h2d(t[k]) + h2d(c));
                }
            }
        }
        ++it;
        if (early_term || it == max_iters) {
            ok = 1;
            for (int r = 0; r < n_rows && ok; ++r)
                for (int z = 0; z < Z; ++z) {
                    int par = 0;
                    for (int e = row_start[r]; e < row_start[r + 1]; ++e) par ^= h2d(app[vidx[(size_t)e * Z + z]]) < 0.0;
                    if (par) { ok = 0; break; }
                }
            if (early_term && ok) break;
        }
    }
    for (int k = 0; k < Kc * Z; ++k) hard_info[k] = h2d(app[k]) < 0.0;
    if (app_out) for (int i = 0; i < nV; ++i) app_out[i] = (float)h2d(app[i]);
    if (parity_ok) *parity_ok = (uint8_t)ok;
    free(app); free(chan); free(c2v);
    return it;
}

ORC_API int orc_decode_nms_f16(int bg, int Z, int n_rows, int max_iters, int early_term, float alpha,
                               const float *llr, long batch, uint8_t *hard_info, float *app_out,
                               int32_t *iters_out, uint8_t *parity_ok, int n_threads) {
    int R, C, Kc, E, ils = orc_set_index(Z);
    if ((bg != 1 && bg != 2) || ils < 0 || max_iters < 1) return -1;
    bg_dims(bg, &R, &C, &Kc, &E);
    if (n_rows <= 0 || n_rows > R) n_rows = R;
    if (n_rows < 4) return -1;
    if (n_threads < 1) n_threads = 1;
    const long nV = (long)C * Z, K = (long)Kc * Z;
    int *vidx = build_vidx(bg, Z, ils);
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (long b = 0; b < batch; ++b) {
        int it = decode_nms_f16_one(bg, Z, ils, n_rows, max_iters, early_term, alpha, vidx, llr + b * nV,
                                    hard_info + b * K, app_out ? app_out + b * nV : NULL,
                                    parity_ok ? parity_ok + b : NULL);
        if (iters_out) iters_out[b] = it;
    }
    free(vidx);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Oracle B -- flooding sum-product, float64, "Parity check satisfied" termination: the
 * algorithm MathWorks documents for comm.LDPCDecoder, which is what NRLDPCDecoder.m:120,265
 * actually runs.  q_ij initialised to L(c_i); per iteration all checks
 * r_ji = 2 atanh( prod_{i' != i} tanh(q_i'j / 2) ), then all variables Q_i = L(c_i) + sum_j r_ji,
 * q_ij = Q_i - r_ji; hard = (Q_i < 0); stop when H*hard = 0 or after max_iters.
 * Leave-one-out products are formed by prefix/suffix products (no division); the atanh
 * argument is clipped to +-(1 - 2^-53) so that +inf filler LLRs cannot produce inf - inf.
 * Uses the whole H (the reference always passes the full matrix), n_rows <= 0 => all rows.
 * ---------------------------------------------------------------------------------------- */
static int decode_bp_one(int bg, int Z, int ils, int n_rows, int max_iters, int early_term, const int *vidx,
                         const double *llr, uint8_t *hard_info, double *app_out, uint8_t *parity_ok) {
    int R, C, Kc, E;
    bg_dims(bg, &R, &C, &Kc, &E);
    const unsigned char *er = bg_row(bg);
    int row_start[47];
    for (int r = 0, e = 0; r <= R; ++r) { while (e < E && er[e] < r) ++e; row_start[r] = e; }
    const int nV = C * Z, Eact = row_start[n_rows];
    double *q = (double *)malloc(sizeof(double) * (size_t)Eact * Z);   /* v2c, [edge][check z] */
    double *rm = (double *)malloc(sizeof(double) * (size_t)Eact * Z);  /* c2v */
    double *Q = (double *)malloc(sizeof(double) * nV);
    const size_t nE = (size_t)Eact * Z;
    for (size_t i = 0; i < nE; ++i) q[i] = llr[vidx[i]];
    const double lim = 1.0 - ldexp(1.0, -53);
    int it = 0, ok = 0;
    while (it < max_iters) {
        for (int r = 0; r < n_rows; ++r) {
            const int e0 = row_start[r], deg = row_start[r + 1] - e0;
            for (int z = 0; z < Z; ++z) {
                double th[32], pre[33], suf[33];
                for (int k = 0; k < deg; ++k) th[k] = tanh(0.5 * q[(size_t)(e0 + k) * Z + z]);
                pre[0] = 1.0; for (int k = 0; k < deg; ++k) pre[k + 1] = pre[k] * th[k];
                suf[deg] = 1.0; for (int k = deg - 1; k >= 0; --k) suf[k] = suf[k + 1] * th[k];
                for (int k = 0; k < deg; ++k) {
                    double x = pre[k] * suf[k + 1];
                    x = x > lim ? lim : (x < -lim ? -lim : x);
                    rm[(size_t)(e0 + k) * Z + z] = 2.0 * atanh(x);
                }
            }
        }
        for (int i = 0; i < nV; ++i) Q[i] = llr[i];
        for (size_t i = 0; i < nE; ++i) Q[vidx[i]] += rm[i];
        for (size_t i = 0; i < nE; ++i) q[i] = Q[vidx[i]] - rm[i];
        ++it;
        ok = 1;
        for (int r = 0; r < n_rows && ok; ++r)
            for (int z = 0; z < Z; ++z) {
                int par = 0;
                for (int e = row_start[r]; e < row_start[r + 1]; ++e)
                    par ^= Q[vidx[(size_t)e * Z + z]] < 0.0;
                if (par) { ok = 0; break; }
            }
        if (ok && early_term) break;   /* early_term = 1 is the reference's setting (NRLDPCDecoder.m:120) */
    }
    if (it == 0) for (int i = 0; i < nV; ++i) Q[i] = llr[i];
    for (int k = 0; k < Kc * Z; ++k) hard_info[k] = Q[k] < 0.0;
    if (app_out) memcpy(app_out, Q, sizeof(double) * nV);
    if (parity_ok) *parity_ok = (uint8_t)ok;
    free(q); free(rm); free(Q);
    return it;
}

ORC_API int orc_decode_bp(int bg, int Z, int n_rows, int max_iters, const double *llr, long batch,
                          uint8_t *hard_info, int32_t *iters_out, uint8_t *parity_ok, int n_threads) {
    int R, C, Kc, E, ils = orc_set_index(Z);
    if ((bg != 1 && bg != 2) || ils < 0 || max_iters < 1) return -1;
    bg_dims(bg, &R, &C, &Kc, &E);
    if (n_rows <= 0 || n_rows > R) n_rows = R;
    if (n_threads < 1) n_threads = 1;
    const long nV = (long)C * Z, K = (long)Kc * Z;
    int *vidx = build_vidx(bg, Z, ils);
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (long b = 0; b < batch; ++b) {
        int it = decode_bp_one(bg, Z, ils, n_rows, max_iters, 1, vidx, llr + b * nV, hard_info + b * K, NULL,
                               parity_ok ? parity_ok + b : NULL);
        if (iters_out) iters_out[b] = it;
    }
    free(vidx);
    return 0;
}

/* Same with the termination rule selectable (early_term = 0: 'Maximum iteration count') and the a-posteriori
 * values returned: the checker of the device's NRLDPC_ALG_BP kernel (tests/test_gpu_parity.py). */
ORC_API int orc_decode_bp_ex(int bg, int Z, int n_rows, int max_iters, int early_term, const double *llr, long batch,
                             uint8_t *hard_info, double *app_out, int32_t *iters_out, uint8_t *parity_ok, int n_threads) {
    int R, C, Kc, E, ils = orc_set_index(Z);
    if ((bg != 1 && bg != 2) || ils < 0 || max_iters < 1) return -1;
    bg_dims(bg, &R, &C, &Kc, &E);
    if (n_rows <= 0 || n_rows > R) n_rows = R;
    if (n_threads < 1) n_threads = 1;
    const long nV = (long)C * Z, K = (long)Kc * Z;
    int *vidx = build_vidx(bg, Z, ils);
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (long b = 0; b < batch; ++b) {
        int it = decode_bp_one(bg, Z, ils, n_rows, max_iters, early_term, vidx, llr + b * nV, hard_info + b * K,
                               app_out ? app_out + b * nV : NULL, parity_ok ? parity_ok + b : NULL);
        if (iters_out) iters_out[b] = it;
    }
    free(vidx);
    return 0;
}

/* Same, float32 input (the GPU boundary's dtype) widened to double: used by bench.py --impl reference. */
ORC_API int orc_decode_bp_f32(int bg, int Z, int n_rows, int max_iters, const float *llr, long batch,
                              uint8_t *hard_info, int32_t *iters_out, uint8_t *parity_ok, int n_threads) {
    int R, C, Kc, E, ils = orc_set_index(Z);
    if ((bg != 1 && bg != 2) || ils < 0 || max_iters < 1) return -1;
    bg_dims(bg, &R, &C, &Kc, &E);
    if (n_rows <= 0 || n_rows > R) n_rows = R;
    if (n_threads < 1) n_threads = 1;
    const long nV = (long)C * Z, K = (long)Kc * Z;
    int *vidx = build_vidx(bg, Z, ils);
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (long b = 0; b < batch; ++b) {
        double *d = (double *)malloc(sizeof(double) * nV);
        for (long i = 0; i < nV; ++i) d[i] = llr[b * nV + i];
        int it = decode_bp_one(bg, Z, ils, n_rows, max_iters, 1, vidx, d, hard_info + b * K, NULL,
                               parity_ok ? parity_ok + b : NULL);
        if (iters_out) iters_out[b] = it;
        free(d);
    }
    free(vidx);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Rate matching, TX: NRLDPCEncoder.m:168-197 (bit_selection) and :200-225 (bit_interleaving).
 * d is the length-N buffer with filler marked 0xFF ("NaN", NRLDPCEncoder.m:155).
 * ---------------------------------------------------------------------------------------- */
#define ORC_FILL 0xFF

/* NRLDPCEncoder.m:149-163: cw (N+2Z, systematic) -> d (N), filler range [K'-2Z, K-2Z) -> NaN */
ORC_API void orc_cw_to_d(int Z, int K, int K_prime, int N, const uint8_t *cw, uint8_t *d) {
    for (int k = 0; k < N; ++k) d[k] = cw[k + 2 * Z];
    for (int k = (K_prime > 2 * Z ? K_prime : 2 * Z); k < K; ++k) d[k - 2 * Z] = ORC_FILL;
}

ORC_API void orc_bit_selection_tx(const uint8_t *d, int N_cb, int k_0, int E, uint8_t *e) {
    int k = 0, j = 0;
    while (k < E) {                                   /* :187-195 */
        uint8_t v = d[(k_0 + j) % N_cb];
        if (v != ORC_FILL) e[k++] = v;
        ++j;
    }
}

ORC_API void orc_interleave_tx(const uint8_t *e, int E, int Q_m, uint8_t *f) {
    for (int j = 0; j < E / Q_m; ++j)                 /* :219-223 */
        for (int i = 0; i < Q_m; ++i) f[i + j * Q_m] = e[i * (E / Q_m) + j];
}

/* RX: NRLDPCDecoder.m:172-197 (bit_interleaving) and :200-242 (bit_selection), float32 LLRs.
 * d_out: N values; filler positions hold NaN as in the reference (:224).  harq_buf (N_cb values,
 * may be NULL) is read-modify-written as at :236-239. */
ORC_API void orc_deinterleave_rx(const float *f, int E, int Q_m, float *e) {
    for (int j = 0; j < E / Q_m; ++j)                 /* :191-195 */
        for (int i = 0; i < Q_m; ++i) e[i * (E / Q_m) + j] = f[i + j * Q_m];
}

ORC_API void orc_bit_selection_rx(const float *e, int E, int N, int N_cb, int k_0, int Z, int K,
                                  int K_prime, float *harq_buf, float *d_out) {
    for (int i = 0; i < N; ++i) d_out[i] = 0.0f;      /* :223 */
    int f0 = K_prime - 2 * Z; if (f0 < 0) f0 = 0;     /* :224, 1-based max(K'-2Z+1,1):K-2Z */
    for (int i = f0; i < K - 2 * Z; ++i) d_out[i] = NAN;
    int k = 0, j = 0;
    while (k < E) {                                   /* :226-234 */
        int idx = (k_0 + j) % N_cb;
        if (!(d_out[idx] != d_out[idx])) { d_out[idx] = d_out[idx] + e[k]; ++k; }
        ++j;
    }
    if (harq_buf) {                                   /* :236-239 */
        for (int i = 0; i < N_cb; ++i) { d_out[i] = d_out[i] + harq_buf[i]; harq_buf[i] = d_out[i]; }
    }
}

/* NRLDPCDecoder.m:262-264: cw_tilde = [zeros(2Z); d_tilde], NaN -> +inf */
ORC_API void orc_d_to_cw_llr(const float *d, int N, int Z, float *cw_llr) {
    for (int i = 0; i < 2 * Z; ++i) cw_llr[i] = 0.0f;
    for (int i = 0; i < N; ++i) cw_llr[2 * Z + i] = (d[i] != d[i]) ? INFINITY : d[i];
}

/* ------------------------------------------------------------------------------------------
 * CRC -- get_3gpp_crc_polynomial.m:3-17 (CRC24A, CRC24B, CRC16), comm.CRCGenerator defaults:
 * zero initial state, no reflection, no final XOR, parity appended MSB first.
 * kind: 0 = CRC16, 1 = CRC24A, 2 = CRC24B.  bits: one bit per byte.  Returns L.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_crc(int kind, const uint8_t *bits, int n, uint8_t *parity) {
    static const uint32_t poly[3] = {0x1021u, 0x864CFBu, 0x800063u};
    const int L = kind == 0 ? 16 : 24;
    const uint32_t top = 1u << (L - 1), mask = (L == 16) ? 0xFFFFu : 0xFFFFFFu;
    uint32_t reg = 0;
    for (int i = 0; i < n; ++i) {
        uint32_t fb = ((reg & top) ? 1u : 0u) ^ (bits[i] & 1u);
        reg = (reg << 1) & mask;
        if (fb) reg ^= poly[kind];
    }
    for (int i = 0; i < L; ++i) parity[i] = (uint8_t)((reg >> (L - 1 - i)) & 1u);
    return L;
}

/* ------------------------------------------------------------------------------------------
 * QPSK map + exact-LLR demap: NRModulator.m:75 (TS 38.211 5.1.3: x = ((1-2b0) + j(1-2b1))/sqrt2),
 * NRDemodulator.m:5,78 with Variance = total complex noise variance (plot_BLER_vs_SNR.m:105-106):
 * LLR(b0) = 2*sqrt(2)*Re(y)/variance, LLR(b1) = 2*sqrt(2)*Im(y)/variance.
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_qpsk_mod(const uint8_t *bits, long n_sym, float *re, float *im) {
    const float a = 0.70710678118654752440f;
    for (long i = 0; i < n_sym; ++i) {
        re[i] = bits[2 * i] ? -a : a;
        im[i] = bits[2 * i + 1] ? -a : a;
    }
}
ORC_API void orc_qpsk_demod(const float *re, const float *im, long n_sym, float variance, float *llr) {
    const float g = 2.8284271247461900976f / variance;
    for (long i = 0; i < n_sym; ++i) { llr[2 * i] = g * re[i]; llr[2 * i + 1] = g * im[i]; }
}

/* ------------------------------------------------------------------------------------------
 * NRModulator.m:69-89 / NRDemodulator.m:72-96 for every Modulation value.
 * Constellations: TS 38.211 section 5.1, which is what the toolbox CustomSymbolMapping vectors at
 * NRModulator.m:73-81 encode (tools/make_golden_mod.py checks the two against each other and
 * tests/golden/constellations.json holds the reference's own vectors turned into points).
 *   BPSK   d = [(1-2b0) + j(1-2b0)]/sqrt2            QPSK  d = [(1-2b0) + j(1-2b1)]/sqrt2
 *   16QAM  d = {(1-2b0)[2-(1-2b2)] + j(1-2b1)[2-(1-2b3)]}/sqrt10
 *   64QAM  d = {(1-2b0)[4-(1-2b2)[2-(1-2b4)]] + j(1-2b1)[4-(1-2b3)[2-(1-2b5)]]}/sqrt42
 *   256QAM d = {(1-2b0)[8-(1-2b2)[4-(1-2b4)[2-(1-2b6)]]] + j(1-2b1)[8-(1-2b3)[4-(1-2b5)[2-(1-2b7)]]]}/sqrt170
 * Demodulation is the literal two-dimensional definition in double precision:
 *   exact   L(b_k) = log sum_{s: b_k=0} exp(-|r-s|^2/var) - log sum_{s: b_k=1} exp(-|r-s|^2/var)
 *   approx  L(b_k) = (min_{s: b_k=1} |r-s|^2 - min_{s: b_k=0} |r-s|^2)/var
 *   hard    bits of the nearest constellation point
 * ---------------------------------------------------------------------------------------- */
static void orc_point(int Qm, unsigned sym, float *re, float *im) {
    int b[8];
    for (int k = 0; k < Qm; ++k) b[k] = (sym >> (Qm - 1 - k)) & 1; /* first bit = MSB of the symbol integer */
    int x, y;
    float norm;
    switch (Qm) {
    case 1: x = 1 - 2 * b[0]; y = x; norm = 0.70710678118654752440f; break;
    case 2: x = 1 - 2 * b[0]; y = 1 - 2 * b[1]; norm = 0.70710678118654752440f; break;
    case 4: x = (1 - 2 * b[0]) * (2 - (1 - 2 * b[2])); y = (1 - 2 * b[1]) * (2 - (1 - 2 * b[3])); norm = 0.31622776601683793320f; break;
    case 6: x = (1 - 2 * b[0]) * (4 - (1 - 2 * b[2]) * (2 - (1 - 2 * b[4])));
            y = (1 - 2 * b[1]) * (4 - (1 - 2 * b[3]) * (2 - (1 - 2 * b[5]))); norm = 0.15430334996209191026f; break;
    default: x = (1 - 2 * b[0]) * (8 - (1 - 2 * b[2]) * (4 - (1 - 2 * b[4]) * (2 - (1 - 2 * b[6]))));
             y = (1 - 2 * b[1]) * (8 - (1 - 2 * b[3]) * (4 - (1 - 2 * b[5]) * (2 - (1 - 2 * b[7])))); norm = 0.07669649888473704465f; break;
    }
    *re = (float)x * norm;
    *im = (float)y * norm;
}

ORC_API int orc_modulate(const uint8_t *bits, long n_sym, int Qm, float *re, float *im) {
    if (!(Qm == 1 || Qm == 2 || Qm == 4 || Qm == 6 || Qm == 8)) return -1;
    for (long i = 0; i < n_sym; ++i) {
        unsigned s = 0;
        for (int k = 0; k < Qm; ++k) s = (s << 1) | (bits[i * Qm + k] & 1u);
        orc_point(Qm, s, re + i, im + i);
    }
    return 0;
}

ORC_API int orc_demodulate(const float *re, const float *im, long n_sym, int Qm, double variance, int method, double *out) {
    if (!(Qm == 1 || Qm == 2 || Qm == 4 || Qm == 6 || Qm == 8) || method < 0 || method > 2 || !(variance > 0)) return -1;
    const int M = 1 << Qm;
    float pr[256], pi[256];
    for (int s = 0; s < M; ++s) orc_point(Qm, (unsigned)s, pr + s, pi + s);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n_sym; ++i) {
        double d[256], dmin = INFINITY;
        int best = 0;
        for (int s = 0; s < M; ++s) {
            const double dx = (double)re[i] - (double)pr[s], dy = (double)im[i] - (double)pi[s];
            d[s] = dx * dx + dy * dy;
            if (d[s] < dmin) { dmin = d[s]; best = s; }
        }
        for (int k = 0; k < Qm; ++k) {
            const int bit = Qm - 1 - k;
            if (method == 2) { out[i * Qm + k] = (best >> bit) & 1; continue; }
            double m0 = INFINITY, m1 = INFINITY;
            for (int s = 0; s < M; ++s) { if ((s >> bit) & 1) { if (d[s] < m1) m1 = d[s]; } else if (d[s] < m0) m0 = d[s]; }
            if (method == 1) { out[i * Qm + k] = (m1 - m0) / variance; continue; }
            double s0 = 0, s1 = 0;
            for (int s = 0; s < M; ++s) { if ((s >> bit) & 1) s1 += exp(-(d[s] - m1) / variance); else s0 += exp(-(d[s] - m0) / variance); }
            out[i * Qm + k] = (m1 - m0) / variance + log(s0) - log(s1);
        }
    }
    return 0;
}
