// decode_kernel_h2.cuh -- packed-half variant of the layered normalized min-sum decoder
// (nrldpc_cfg.llr_dtype = NRLDPC_F16X2).
//
// Same mapping as decode_kernel.cuh (thread = check z of every layer, APP values resident in shared
// memory, compressed check-to-variable records in an L2-pinned scratch, layer loop unrolled per
// base graph), but every thread decodes TWO codewords at once: the a-posteriori LLRs of a codeword
// pair are interleaved as one 32-bit {fp16 A, fp16 B} word per variable, so one address
// computation, one LDS/STS and one packed HADD2 / HMNMX2 / HSET2 / LOP3 serve both codewords.
// The kernel is ALU-pipe bound, so halving the instructions per codeword is
// what doubles the throughput; HBM traffic is unchanged (the boundary stays float32).
//
// Arithmetic (bit-exact against oracle/nrldpc_oracle.c, orc_decode_nms_f16):
//   input   x -> fp16(min(max(x, -2048), 2048)) (round to nearest even; NaN / +inf filler -> +2048)
//   check   t_e = app - c_e (fp16, RN);  m1, m2 = two smallest |t_e|, each capped at 2048;
//           c_e' = sgn_e * fp16(alpha_h * (e is an arg-min ? m2 : m1)) with alpha_h = fp16(alpha);
//           app = t_e + c_e' (fp16, RN)
// The caps bound |app| by 2048 + 30 * 1536 < 65504, so no value can overflow to infinity.
#pragma once
#include <cuda_fp16.h>

#include "decode_kernel.cuh"

namespace nrldpc {

constexpr float kH2LlrMax = 2048.0f;
constexpr uint32_t kH2MsgCap = 0x68006800u;   // {2048, 2048} as packed fp16
constexpr uint32_t kH2Sign = 0x80008000u;

__device__ __forceinline__ __half2 as_h2(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ uint32_t as_u32(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }

__device__ __forceinline__ float clamp_llr_h2(float x) { return __fadd_rn(fmaxf(fminf(x, kH2LlrMax), -kH2LlrMax), 0.0f); }

// Packed-half twin of ext_app (decode_kernel.cuh): a-posteriori word {A, B} of the degree-1 parity variable of extension
// row `row` = channel values + the row's latest messages, rebuilt from the row's record (three-word records: every
// extension row has degree <= 11).  Arg-min index 15 in a half = "the degree-1 edge is the arg-min" for that codeword.
__device__ __forceinline__ uint32_t ext_app_h2(const uint32_t *my_rec, int row, uint64_t pol, uint32_t chan) {
    const uint4 rec = ld_rec(my_rec, row, pol);
    const uint32_t is_p = __heq2_mask(as_h2(rec.z & 0x000f000fu), as_h2(0x000f000fu));
    const uint32_t sel = bitselect(rec.x, rec.y, is_p);
    return as_u32(__hadd2(as_h2(chan), as_h2(sel ^ (chan & kH2Sign))));
}
// Hard decisions of the degree-1 parity variables of all active extension rows, both codewords of the pair (bit r of ma / mb
// = row r, codeword A / B), records fetched eight rows at a time: see ext_hard_mask in decode_kernel.cuh.
__device__ __noinline__ ulonglong2 ext_hard_mask_h2(const uint32_t *my_rec, const uint64_t pol, const int n_rows, const uint32_t p_addr,
                                                    const uint32_t col_bytes) {
    unsigned long long a = 0ull, b = 0ull;
    for (int r0 = 4; r0 < n_rows; r0 += 8) {
        uint32_t x[8], y[8], meta[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = min(r0 + i, n_rows - 1);
            x[i] = ld_word(my_rec + (r * 3 + 0) * kRecStride, pol);
            y[i] = ld_word(my_rec + (r * 3 + 1) * kRecStride, pol);
            meta[i] = ld_word(my_rec + (r * 3 + 2) * kRecStride, pol);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (r0 + i < n_rows) {
                const uint32_t chan = lds_u32(p_addr + (uint32_t)(r0 + i - 4) * col_bytes);
                const uint32_t is_p = __heq2_mask(as_h2(meta[i] & 0x000f000fu), as_h2(0x000f000fu));
                const uint32_t sel = bitselect(x[i], y[i], is_p);
                const uint32_t app = as_u32(__hadd2(as_h2(chan), as_h2(sel ^ (chan & kH2Sign))));
                a |= (unsigned long long)((app >> 15) & 1u) << (r0 + i);
                b |= (unsigned long long)(app >> 31) << (r0 + i);
            }
        }
    }
    return make_ulonglong2(a, b);
}

// last_row_parity of the pair kernel: the last active row's record is reloaded from the L2 scratch (holding it in registers
// through the layer loop, as the float32 kernel does, made ptxas spill inside the BG1 loop: 25 spill instructions)
struct ExtAppH2 {
    const uint32_t *my_rec; int row; uint64_t pol;
    __device__ __forceinline__ uint32_t operator()(const uint4, uint32_t chan) const { return ext_app_h2(my_rec, row, pol, chan); }
};

// One check row of degree DEG for check z of a codeword pair.  Record layout (uint4; three words
// when DEG <= 11, i.e. every layer but the four degree-19 ones of base graph 1):
//   x, y : alpha*min1, alpha*min2 of both codewords (packed fp16) with the row's sign product in their sign bits
//   z    : sign bits of the row's t values (message sign = row sign ^ sign(t_e)) on edges 0..min(NE,16)-1: edge e at bit 15-(n0-1-e) of each half
//          (NE = DEG, or DEG - 1 for an extension row: its degree-1 edge keeps no record, see decode_kernel.cuh);
//          DEG <= 11: also the arg-min edge index of each codeword in bits 0..3 / 16..19 (15: the degree-1 edge)
//   w    : DEG > 11 only: arg-min edge indices in bits 0..4 / 16..20 and, for DEG > 16, the sign bits of
//          edges 16..DEG-1 at the top of each half
// Arg-min indices are compared as fp16 bit patterns (HSET2 without flush-to-zero).
// syndrome with the base graph's shape known at compile time (see SyndromeRows in decode_kernel.cuh)
template <int BG, int R, int REND, bool FULL>
struct SyndromeRowsH2 {
    static __device__ __forceinline__ uint32_t run(const DecArgs &a, const Lane &l, const unsigned long long ma, const unsigned long long mb, uint32_t fail) {
        if (R >= 4 && R >= a.n_rows) return fail;
        constexpr int DEG = BgShape<BG>::deg(R);
        constexpr int NE = R >= 4 ? DEG - 1 : DEG;
        constexpr int E0 = BgShape<BG>::start(R);
        // degree-1 parity variable of an extension row: hard decisions of both codewords from ext_hard_mask_h2
        uint32_t par = R >= 4 ? (((uint32_t)((ma >> R) & 1ull) << 15) | ((uint32_t)((mb >> R) & 1ull) << 31)) : 0u;
#pragma unroll
        for (int e = 0; e < NE; ++e) par ^= lds_u32(edge_addr<FULL>(l, a.ed[E0 + e]));
        fail |= par;
        asm volatile("" : "+r"(fail));   // one row's loads are consumed before the next row's are issued (register pressure)
        return SyndromeRowsH2<BG, R + 1, REND, FULL>::run(a, l, ma, mb, fail);
    }
};
template <int BG, int REND, bool FULL>
struct SyndromeRowsH2<BG, REND, REND, FULL> {
    static __device__ __forceinline__ uint32_t run(const DecArgs &, const Lane &, const unsigned long long, const unsigned long long, uint32_t fail) { return fail; }
};

// Bit-sliced, two-stage syndrome of a codeword pair (see syndrome_bitsliced in decode_kernel.cuh): the hard
// decisions of codeword A (bit 15 of every word) and B (bit 31) are packed into hb[0][col][Z/32] and hb[1][col][Z/32].
// All pointers are shared-window byte addresses (explicit LDS / STS: a generic pointer costs an address-space
// resolution per access in an out-of-line routine).
__device__ __forceinline__ void pack_hard_bits_h2(uint32_t app_s, uint32_t hb_s, int Z, int col0, int col1, int n_cols_all, int z,
                                                  int ext_row0 = -1, const unsigned long long ma = 0ull, const unsigned long long mb = 0ull) {
    const uint32_t plane = (uint32_t)n_cols_all * (uint32_t)(Z >> 5) * 4u;
    uint32_t src = app_s + (uint32_t)(col0 * Z + z) * 4u;
    uint32_t dst = hb_s + (uint32_t)(col0 * (Z >> 5) + (z >> 5)) * 4u;
#pragma unroll 1
    for (int col = col0; col < col1; ++col, src += (uint32_t)Z * 4u, dst += (uint32_t)(Z >> 5) * 4u) {
        uint32_t ba, bb;
        if (ext_row0 >= 0) {   // degree-1 parity column (see pack_hard_bits)
            ba = (uint32_t)((ma >> (ext_row0 + (col - col0))) & 1ull);
            bb = (uint32_t)((mb >> (ext_row0 + (col - col0))) & 1ull);
        } else {
            const uint32_t x = lds_u32(src);
            ba = (x >> 15) & 1u;
            bb = x >> 31;
        }
        const uint32_t wa = __ballot_sync(0xffffffffu, ba);
        const uint32_t wb = __ballot_sync(0xffffffffu, bb);
        if ((z & 31) == 0) {
            sts_u32(dst, wa);
            sts_u32(dst + plane, wb);
        }
    }
}
// CTA-uniform result: bit 15 set if codeword A fails, bit 31 if B fails (the convention of the per-thread syndrome).
// live_a / live_b (CTA-uniform): the codeword is still being decoded -- the extension stage runs only if a live
// codeword passed the core stage.  Contains barriers: every thread of the CTA calls it.
template <int BG>
__device__ __noinline__ uint32_t syndrome_bitsliced_h2(uint32_t app_s, uint32_t hb_s, int Z, int n_rows, int z, const unsigned short *row_start,
                                                       bool live_a, bool live_b, const uint32_t *my_rec, const uint64_t pol) {
    using S = BgShape<BG>;
    const int W = Z >> 5, z0 = z & ~31, lane = z & 31;
    constexpr int kCore = S::kKcols + 4;
    const uint32_t hbB_s = hb_s + (uint32_t)(S::kCols * W) * 4u, sed_s = hb_s + (uint32_t)(S::kCols * W) * 8u;
    pack_hard_bits_h2(app_s, hb_s, Z, 0, kCore, S::kCols, z);
    __syncthreads();
    uint32_t fa = 0, fb = 0;
    constexpr unsigned long long kStarts = core_row_starts<BG>();
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
        const int e = (int)((kStarts >> (8 * r)) & 0xffu) + lane;
        uint32_t va = 0, vb = 0;
        if (e < (int)((kStarts >> (8 * r + 8)) & 0xffu)) {
            const uint32_t d = lds_u32(sed_s + (uint32_t)e * 4u);
            va = hb_window(hb_s, d, z0, Z, W);
            vb = hb_window(hbB_s, d, z0, Z, W);
        }
        fa |= __reduce_xor_sync(0xffffffffu, va);
        fb |= __reduce_xor_sync(0xffffffffu, vb);
    }
    // __syncthreads_or reduces a predicate, not a bit mask: one barrier per codeword
    uint32_t f = __syncthreads_or(fa != 0u) ? 0x00008000u : 0u;
    if (__syncthreads_or(fb != 0u)) f |= 0x80000000u;
    const bool need_ext = (live_a && !(f & 0x00008000u)) || (live_b && !(f & 0x80000000u));
    if (!need_ext || n_rows <= 4) return f;
    const ulonglong2 m = ext_hard_mask_h2(my_rec, pol, n_rows, app_s + (uint32_t)(kCore * Z + z) * 4u, (uint32_t)Z * 4u);
    pack_hard_bits_h2(app_s, hb_s, Z, kCore, min(S::kCols, S::kKcols + n_rows), S::kCols, z, 4, m.x, m.y);
    __syncthreads();
    fa = 0; fb = 0;
    for (int r = 4 + lane; r < n_rows; r += 32) {
        uint32_t xa = 0, xb = 0;
        for (int e = row_start[r]; e < row_start[r + 1]; ++e) {
            const uint32_t d = lds_u32(sed_s + (uint32_t)e * 4u);
            xa ^= hb_window(hb_s, d, z0, Z, W);
            xb ^= hb_window(hbB_s, d, z0, Z, W);
        }
        fa |= xa;
        fb |= xb;
    }
    if (__syncthreads_or(fa != 0u)) f |= 0x00008000u;
    if (__syncthreads_or(fb != 0u)) f |= 0x80000000u;
    return f;
}

// out of line, two stages: see syndrome_unrolled_core / _ext in decode_kernel.cuh
template <int BG, bool FULL>
__device__ __noinline__ uint32_t syndrome_unrolled_core_h2(const DecArgs &a, const Lane l) {
    return SyndromeRowsH2<BG, 0, 4, FULL>::run(a, l, 0ull, 0ull, 0u);
}
template <int BG, bool FULL>
__device__ __noinline__ uint32_t syndrome_unrolled_ext_h2(const DecArgs &a, const Lane l, const uint32_t *my_rec, const uint64_t pol) {
    const ulonglong2 m = ext_hard_mask_h2(my_rec, pol, a.n_rows, a.smem_base + l.slot_off + l.zoff + (uint32_t)((a.kcols + 4) * a.Z) * 4u, (uint32_t)a.Z * 4u);
    return SyndromeRowsH2<BG, 4, BgShape<BG>::kRows, FULL>::run(a, l, m.x, m.y, 0u);
}

template <int DEG>
struct RowStateH2 {
    uint32_t t[DEG];
    uint32_t addr[DEG];
    __half2 m1, m2;
    uint32_t sx, s0, s1;
};

// first half of a row update (see row_gather in decode_kernel.cuh)
template <int DEG, bool IDENT_LAST, bool ONE_CW>
__device__ __forceinline__ void row_gather_h2(const Lane &l, const uint2 *__restrict__ ed, const uint4 rec, RowStateH2<DEG> &s) {
    constexpr int NE = IDENT_LAST ? DEG - 1 : DEG;   // edges whose variable has other checks too (recorded edges)
    constexpr int N0 = NE < 16 ? NE : 16;
    __half2 m1 = as_h2(0u), m2 = as_h2(0u);
    uint32_t sx = 0, s0 = 0, s1 = 0;
    constexpr bool W3 = DEG <= 11;   // three-word record
    const uint32_t oargs = W3 ? (rec.z & 0x000f000fu) : (rec.w & 0x001f001fu);
#pragma unroll
    for (int e = 0; e < DEG; ++e) {
        const uint2 d = ed[e];
        const uint32_t a = edge_addr<ONE_CW>(l, d, IDENT_LAST && e == DEG - 1);
        s.addr[e] = a;
        const uint32_t x = lds_u32(a);
        __half2 tt = as_h2(x);                       // degree-1 variable: its channel value
        if (e < NE) {
            const uint32_t e2 = (uint32_t)e * 0x00010001u;
            const uint32_t is_arg = __heq2_mask(as_h2(oargs), as_h2(e2));
            const uint32_t mag = bitselect(rec.x, rec.y, is_arg);
            const uint32_t sw = e < 16 ? rec.z << (N0 - 1 - e) : rec.w << (NE - 1 - e);
            const uint32_t c = mag ^ (sw & kH2Sign);
            tt = __hsub2(as_h2(x), as_h2(c));
        }
        s.t[e] = as_u32(tt);
        const __half2 ab = __habs2(tt);
        if (e == 0) {
            m1 = ab;
        } else if (e == 1) {
            m2 = __hmax2(m1, ab);
            m1 = __hmin2(m1, ab);
        } else {
            m2 = __hmin2(m2, __hmax2(ab, m1));
            m1 = __hmin2(m1, ab);
        }
        sx ^= as_u32(tt);
        // (a shift on the ALU pipe: the FMA-pipe alternative mul.hi(s, 2^31) was measured 4 % slower)
        if (e < NE) {
            if (e < 16) s0 = bitselect(s0 >> 1, as_u32(tt), kH2Sign);
            else s1 = bitselect(s1 >> 1, as_u32(tt), kH2Sign);
        }
    }
    s.m1 = m1; s.m2 = m2; s.sx = sx; s.s0 = s0; s.s1 = s1;
}

// second half: new messages, APP write-back, the row's new record
template <int DEG, bool IDENT_LAST, bool PAR>
__device__ __forceinline__ uint4 row_scatter_h2_par(const RowStateH2<DEG> &s, const uint32_t alpha2, uint32_t &par, const bool live = true) {
    constexpr int NE = IDENT_LAST ? DEG - 1 : DEG;
    constexpr int N0 = NE < 16 ? NE : 16;
    constexpr int N1 = NE - N0;
    constexpr bool W3 = DEG <= 11;
    static_assert(!IDENT_LAST || W3, "extension rows have three-word records");
    const __half2 m1raw = s.m1;
    const __half2 m1 = __hmin2(s.m1, as_h2(kH2MsgCap));
    const __half2 m2 = __hmin2(s.m2, as_h2(kH2MsgCap));
    // multiply by alpha carrying the row's sign product (m >= 0: bit-identical to (alpha*m) | sg, also for m = 0)
    const uint32_t alpha_s = bitselect(alpha2, s.sx, kH2Sign);
    uint32_t m1ss = as_u32(__hmul2(as_h2(alpha_s), m1));
    uint32_t m2ss = as_u32(__hmul2(as_h2(alpha_s), m2));
    asm volatile("" : "+r"(m1ss), "+r"(m2ss));
    uint32_t args = IDENT_LAST ? 0x000f000fu : 0u;   // 15: no recorded edge attains the minimum (the degree-1 edge does)
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        // arg-min edges found by value (ties: min2 == min1, every tied edge gets the same message)
        const uint32_t is_min = __heq2_mask(__habs2(as_h2(s.t[e])), m1raw);
        const uint32_t sel = bitselect(m1ss, m2ss, is_min);
        args = bitselect(args, (uint32_t)e * 0x00010001u, is_min);
        const uint32_t c = sel ^ (s.t[e] & kH2Sign);
        const uint32_t app = as_u32(__hadd2(as_h2(s.t[e]), as_h2(c)));
        if (PAR) par ^= app;   // sign bits (15 / 31) = parities of this check on the hard decisions just written
        if (live) sts_u32(s.addr[e], app);
    }
    if (PAR && IDENT_LAST) {   // the degree-1 variable's a-posteriori value takes part in the check's parity only
        const uint32_t tp = s.t[DEG - 1];
        const uint32_t sel = bitselect(m1ss, m2ss, __heq2_mask(__habs2(as_h2(tp)), m1raw));
        par ^= as_u32(__hadd2(as_h2(tp), as_h2(sel ^ (tp & kH2Sign))));
    }
    constexpr uint32_t F0 = (((1u << N0) - 1u) << (16 - N0)) * 0x00010001u;
    constexpr uint32_t F1 = N1 > 0 ? (((1u << N1) - 1u) << (16 - N1)) * 0x00010001u : 0u;
    const uint32_t z = (s.s0 & F0) | (W3 ? args : 0u);
    const uint32_t w = W3 ? 0u : ((s.s1 & F1) | args);
    return make_uint4(m1ss, m2ss, z, w);
}
template <int DEG, bool IDENT_LAST>
__device__ __forceinline__ uint4 row_scatter_h2(const RowStateH2<DEG> &s, const uint32_t alpha2, const bool live = true) {
    uint32_t unused = 0;
    return row_scatter_h2_par<DEG, IDENT_LAST, false>(s, alpha2, unused, live);
}

template <int DEG, bool IDENT_LAST, bool ONE_CW>
__device__ __forceinline__ uint4 process_row_h2(const Lane &l, const uint2 *__restrict__ ed, const uint4 rec,
                                                const uint32_t alpha2, const bool live = true) {
    RowStateH2<DEG> s;
    row_gather_h2<DEG, IDENT_LAST, ONE_CW>(l, ed, rec, s);
    return row_scatter_h2<DEG, IDENT_LAST>(s, alpha2, live);
}

// ---- pieces of the pair kernel ---------------------------------------------------------------------
struct DecCtxH2 {
    Lane l;
    uint32_t *my_rec;
    uint64_t pol;
    uint4 cur, cur2;   // prefetched records of the next layer (and of its partner when the next layer is a row pair)
    bool done;   // this thread does no row work (inactive lane, or both codewords of its pair are finished)
    uint32_t last_fail;   // FULL, every base row active: bit 15 / 31 = codeword A / B has an unsatisfied check in the last layer
};

// Load one codeword pair (B may be absent: zeros) into its interleaved APP array.
__device__ __forceinline__ void load_pair(const float *__restrict__ rowA, const float *__restrict__ rowB, uint32_t *app,
                                          int ncw, int lane, int nlanes) {
    if (reinterpret_cast<uintptr_t>(app) & 15) {   // padded slot stride that is not a multiple of four words (tiny odd Z)
        for (int i = lane; i < ncw; i += nlanes)
            app[i] = as_u32(__floats2half2_rn(clamp_llr_h2(__ldcs(rowA + i)), clamp_llr_h2(rowB ? __ldcs(rowB + i) : 0.f)));
        return;
    }
    const float4 *a4 = reinterpret_cast<const float4 *>(rowA);
    const float4 *b4 = reinterpret_cast<const float4 *>(rowB);
    uint4 *dst = reinterpret_cast<uint4 *>(app);
#pragma unroll 4
    for (int i = lane; i < (ncw >> 2); i += nlanes) {
        const float4 va = __ldcs(a4 + i);
        const float4 vb = rowB ? __ldcs(b4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint4 o;
        o.x = as_u32(__floats2half2_rn(clamp_llr_h2(va.x), clamp_llr_h2(vb.x)));
        o.y = as_u32(__floats2half2_rn(clamp_llr_h2(va.y), clamp_llr_h2(vb.y)));
        o.z = as_u32(__floats2half2_rn(clamp_llr_h2(va.z), clamp_llr_h2(vb.z)));
        o.w = as_u32(__floats2half2_rn(clamp_llr_h2(va.w), clamp_llr_h2(vb.w)));
        dst[i] = o;
    }
}

// Outputs of ONE codeword (half 0 = A, 1 = B) of a pair, written by the pair's own lanes.
// Soft output: lane z handles position z of every block column (nlanes = Z), so for the degree-1 parity columns of the
// active extension rows it owns the row's record and rebuilds the a-posteriori value on the fly (ext_app_h2).
__device__ __forceinline__ void store_half(uint8_t *hard_base, float *soft_base, const int kcols, const int n_rows, const uint32_t *app,
                                           long long cw, int half, int ncw, int K,
                                           int lane, int nlanes, const uint32_t *my_rec, const uint64_t pol) {
    const int sh = half ? 31 : 15;
    uint8_t *hard = hard_base + cw * K;
    if ((K & 3) == 0 && (reinterpret_cast<uintptr_t>(app) & 15) == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(app);
        uint32_t *dst = reinterpret_cast<uint32_t *>(hard);
        for (int k = lane; k < (K >> 2); k += nlanes) {
            const uint4 v = src[k];
            dst[k] = ((v.x >> sh) & 1u) | (((v.y >> sh) & 1u) << 8) | (((v.z >> sh) & 1u) << 16) | (((v.w >> sh) & 1u) << 24);
        }
    } else {
        for (int k = lane; k < K; k += nlanes) hard[k] = (uint8_t)((app[k] >> sh) & 1u);
    }
    if (soft_base) {
        float *dst = soft_base + cw * ncw;
        const int ext0 = kcols + 4, ext1 = kcols + n_rows;
        for (int i = lane, col = 0; i < ncw; i += nlanes, ++col) {
            uint32_t w = app[i];
            if (col >= ext0 && col < ext1) w = ext_app_h2(my_rec, col - kcols, pol, w);
            const float2 f = __half22float2(as_h2(w));
            __stcs(dst + i, half ? f.y : f.x);
        }
    }
}

// syndrome of hard = (app < 0) over the active rows; bit 15 = codeword A fails, bit 31 = B fails
__device__ __forceinline__ uint32_t syndrome_fail_h2(const DecArgs &a, const DecCtxH2 &c) {
    uint32_t fail = 0;
    for (int r = 0; r < a.n_rows; ++r) {
        uint32_t par = 0;
        const int e1 = a.row_start[r + 1];
        for (int e = a.row_start[r]; e < e1; ++e) {
            uint32_t w = lds_u32(edge_addr<false>(c.l, a.ed[e]));
            if (r >= 4 && e == e1 - 1) w = ext_app_h2(c.my_rec, r, c.pol, w);   // (unbatched: this generic routine is not on any hot path)
            par ^= w;
        }
        fail |= par;
    }
    return fail & kH2Sign;
}

// Layer loop, unrolled per base graph.  As in decode_kernel.cuh: ld_from / ld_to gate the prefetch of the next
// layer's record, and two consecutive base rows that touch disjoint block columns (pair_first<BG>) run as one layer
// with one barrier.  Row pairs only occur among rows of degree <= 11, i.e. with three-word records.
template <int BG, int R, bool FULL, bool MASKED = false>
struct UnrolledRowsH2 {
    static __device__ __forceinline__ void run(const DecArgs &a, DecCtxH2 &c, const int ld_from, const int ld_to, const bool store_rec) {
        if (R >= 4 && R >= a.n_rows) return;   // n_rows >= 4 is validated by the host
        constexpr int DEG = BgShape<BG>::deg(R);
        constexpr int E0 = BgShape<BG>::start(R);
        constexpr bool kW4 = DEG > 11;                                                  // this layer has a 4th word
        constexpr bool PAIR = pair_first<BG>(R) && !kW4 && BgShape<BG>::deg(R + 1 < BgShape<BG>::kRows ? R + 1 : R) <= 11;
        uint32_t *w4 = c.my_rec + kRecSlots * 3 * kRecStride;                           // [layer 0..3][kRecStride]
        if (PAIR && R + 1 < a.n_rows) {
            constexpr int DEG2 = BgShape<BG>::deg(PAIR ? R + 1 : R);
            constexpr int E1 = BgShape<BG>::start(PAIR ? R + 1 : R);
            constexpr bool kLastPair = FULL && PAIR && R + 2 == BgShape<BG>::kRows;
            uint32_t par = 0;
            if (FULL || !c.done) {
                uint4 nxt = make_uint4(0u, 0u, 0u, 0u), nxt2 = nxt;
                if ((R + 1 >= ld_from && R + 1 < ld_to)) {
                    nxt = ld_rec(c.my_rec, R + 2, c.pol);   // no 4th word: layers >= 4 have degree <= 11
                    if (pair_first<BG>(R + 2)) nxt2 = ld_rec(c.my_rec, R + 3, c.pol);
                }
                RowStateH2<DEG> s0;
                RowStateH2<DEG2> s1;
                row_gather_h2<DEG, (R >= 4), FULL>(c.l, a.ed + E0, c.cur, s0);
                row_gather_h2<DEG2, (R >= 4), FULL>(c.l, a.ed + E1, c.cur2, s1);
                const uint4 rec0 = row_scatter_h2_par<DEG, (R >= 4), kLastPair>(s0, a.alpha_h2, par);
                const uint4 rec1 = row_scatter_h2_par<DEG2, (R >= 4), kLastPair>(s1, a.alpha_h2, par);
                if (store_rec) {
                    st_rec(c.my_rec, R, rec0, c.pol);
                    st_rec(c.my_rec, R + 1, rec1, c.pol);
                }
                c.cur = nxt;
                c.cur2 = nxt2;
            }
            // last layer of an iteration with every base row active: its hard decisions are final, the barrier doubles as
            // the CTA-wide OR of its parities (see decode_kernel.cuh); one reduction per codeword of the pair
            if (kLastPair) {
                const int fa = __syncthreads_or((int)((par >> 15) & 1u));
                const int fb = __syncthreads_or((int)(par >> 31));
                c.last_fail = (fa ? 0x00008000u : 0u) | (fb ? 0x80000000u : 0u);
            } else {
                __syncthreads();
            }
            UnrolledRowsH2<BG, (PAIR ? R + 2 : BgShape<BG>::kRows), FULL, MASKED>::run(a, c, ld_from, ld_to, store_rec);
        } else {
            if (FULL || !c.done) {
                constexpr bool kNextW4 = R + 1 < BgShape<BG>::kRows && BgShape<BG>::deg(R + 1 < BgShape<BG>::kRows ? R + 1 : R) > 11;
                uint4 nxt = make_uint4(0u, 0u, 0u, 0u), nxt2 = nxt;
                if ((R >= ld_from && R < ld_to)) {
                    nxt = ld_rec(c.my_rec, R + 1, c.pol);
                    if (kNextW4) nxt.w = ld_word(w4 + (R + 1) * kRecStride, c.pol);
                    if (!PAIR && pair_first<BG>(R + 1)) nxt2 = ld_rec(c.my_rec, R + 2, c.pol);
                }
                // layer 0's 4th word is not prefetched across the iteration boundary: fetch it on entry
                if (R == 0 && kW4 && (ld_from == 0)) c.cur.w = ld_word(w4, c.pol);
                const uint4 rec = process_row_h2<DEG, (R >= 4), FULL>(c.l, a.ed + E0, c.cur, a.alpha_h2);
                if (store_rec) {
                    st_rec(c.my_rec, R == 0 ? a.n_rows : R, rec, c.pol);
                    if (kW4) st_word(w4 + R * kRecStride, rec.w, c.pol);
                }
                c.cur = nxt;
                c.cur2 = nxt2;
            }
            __syncthreads();
            if (!PAIR) UnrolledRowsH2<BG, R + 1, FULL, MASKED>::run(a, c, ld_from, ld_to, store_rec);
        }
    }
};
template <int BG, bool FULL, bool MASKED>
struct UnrolledRowsH2<BG, BgShape<BG>::kRows, FULL, MASKED> {
    static __device__ __forceinline__ void run(const DecArgs &, DecCtxH2 &, int, int, bool) {}
};

// shared-window address of the packed hard decisions of both codewords (FULL kernels): behind the flags, the work slot
// and the (unused here) barrier slot
// one pair per CTA (FULL kernels): per-codeword load / store out of line, see load_group_ool in decode_kernel.cuh
__device__ __noinline__ void load_pair_ool(const float *__restrict__ rowA, const float *__restrict__ rowB, uint32_t *app, int ncw, int lane, int nlanes) {
    load_pair(rowA, rowB, app, ncw, lane, nlanes);
}
__device__ __noinline__ void store_half_ool(uint8_t *hard_base, float *soft_base, const int kcols, const int n_rows, const uint32_t *app, long long cw,
                                            int half, int ncw, int K, int lane, int nlanes, const uint32_t *my_rec, const uint64_t pol) {
    store_half(hard_base, soft_base, kcols, n_rows, app, cw, half, ncw, K, lane, nlanes, my_rec, pol);
}
template <bool FULL>
__device__ __forceinline__ void store_half_sel(const DecArgs &a, const uint32_t *app, long long cw, int half, int ncw, int K, int lane, int nlanes,
                                               const uint32_t *my_rec, const uint64_t pol) {
    if (FULL) store_half_ool(a.hard, a.soft, a.kcols, a.n_rows, app, cw, half, ncw, K, lane, nlanes, my_rec, pol);   // by value: no generic loads of the parameters out of line
    else store_half(a.hard, a.soft, a.kcols, a.n_rows, app, cw, half, ncw, K, lane, nlanes, my_rec, pol);
}

__device__ __forceinline__ uint32_t h2_hard_bits(int *s_flag, int cwpc) {
    return (((uint32_t)__cvta_generic_to_shared(s_flag + 4 * cwpc + 1) + 7u) & ~7u) + 8u;
}

// Here cwpc counts codeword PAIRS per CTA; FULL = one pair per CTA and every thread owns a check.
template <int BG, bool FULL, bool MASKED = false>
__global__ void __launch_bounds__(kDecThreads, kDecCtasPerSm) decode_nms_h2_kernel(const __grid_constant__ DecArgs a) {
    static_assert(FULL || !MASKED, "MASKED is a flavour of the one-pair (FULL) kernels");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Z = a.Z;
    const int ncw = a.ncols * Z;
    const int K = a.kcols * Z;
    uint32_t *app = reinterpret_cast<uint32_t *>(smem_raw);
    // layout behind the APP arrays as in decode_nms_kernel: flags [2 stages][2*cwpc], work-group slot, (unused here)
    // barrier slot, packed hard decisions of both codewords, lane-indexed edge table
    int *s_flag = reinterpret_cast<int *>(app + (size_t)a.cwpc * a.slot_stride);
    int &s_group = s_flag[4 * a.cwpc];
    if ((uint32_t)__cvta_generic_to_shared(smem_raw) != a.smem_base) __trap();
    const bool bitsliced = FULL && !MASKED && a.n_rows >= a.bitsliced_min_rows;
    if (bitsliced && (a.early_term || a.ok != nullptr))   // ordered by the barriers below
        fill_syndrome_edges(a, h2_hard_bits(s_flag, a.cwpc) + (uint32_t)(a.ncols * (Z >> 5)) * 8u);

    const int tid = threadIdx.x;
    const int slot = tid / Z;
    const int z = tid - slot * Z;
    const bool lane_ok = tid < a.cwpc * Z;
    const int per_group = 2 * a.cwpc;
    const long long n_groups = (a.batch + per_group - 1) / per_group;
    const bool want_ok = a.ok != nullptr;
    const bool keep_last = a.early_term || want_ok || a.soft != nullptr;   // see decode_nms_kernel

    DecCtxH2 c;
    c.l.zoff = (uint32_t)z * 4u;
    c.l.nZ4 = 0u - (uint32_t)Z * 4u;
    c.l.slot_off = FULL ? 0u : (uint32_t)(slot * a.slot_stride) * 4u;
    c.l.one = (uint32_t)a.one;
    c.my_rec = a.c2v + (size_t)(blockIdx.x / a.rec_group) * (kRecWords * kRecStride) + (blockIdx.x % a.rec_group) * blockDim.x + tid;
    c.pol = make_l2_policy(a.l2_pin);
    uint32_t *my_app = app + (size_t)(lane_ok ? slot : 0) * a.slot_stride;

    while (true) {
        __syncthreads();  // previous group's outputs are out of smem
        if (tid == 0) s_group = (int)((unsigned int)atomicAdd(a.work_counter, 1) - a.work_base);
        __syncthreads();
        const long long group = s_group;
        if (group >= n_groups) break;
        const long long cw0 = group * per_group;
        const int n_here = (int)min((long long)per_group, a.batch - cw0);  // codewords in this group

        const long long cwA = cw0 + 2 * slot, cwB = cwA + 1;
        const bool active = lane_ok && 2 * slot < n_here;
        const bool has_b = lane_ok && 2 * slot + 1 < n_here;
        if (active) {
            if (FULL) load_pair_ool(a.llr + cwA * ncw, has_b ? a.llr + cwB * ncw : nullptr, my_app, ncw, z, Z);
            else load_pair(a.llr + cwA * ncw, has_b ? a.llr + cwB * ncw : nullptr, my_app, ncw, z, Z);
        }
        for (int i = tid; i < 2 * per_group; i += blockDim.x) s_flag[i] = 0;
        __syncthreads();

        bool fin_a = !active, fin_b = !has_b;   // finished (converged and already written out, or absent)
        c.done = !active;
        c.cur = make_uint4(0u, 0u, 0u, 0u);
        c.cur2 = c.cur;
        int it_a = 0, it_b = 0, ok_a = 0, ok_b = 0;

        for (int it = 0; it < a.max_iters; ++it) {
            const bool first = it == 0, last = it + 1 == a.max_iters;
            c.last_fail = 0u;   // set by the last layer when every base row is active
            UnrolledRowsH2<BG, 0, FULL, MASKED>::run(a, c, first ? a.n_rows - 1 : 0, last ? a.n_rows - 1 : a.n_rows, !last || keep_last);
            if (!fin_a) it_a = it + 1;
            if (!fin_b) it_b = it + 1;
            if (a.early_term || (want_ok && last)) {
                uint32_t fu = 0;   // bit 15: codeword A fails, bit 31: B fails
                if (FULL && !MASKED && bitsliced) {
                    // codewords with an unsatisfied check in the last layer have not converged; the syndrome runs only if
                    // a live codeword of the pair is still undecided (it is exact for both)
                    if (a.n_rows < BgShape<BG>::kRows) {   // trimmed row count: re-read the last active row (see decode_kernel.cuh)
                        const uint32_t par = last_row_parity(a, c.l, make_uint4(0u, 0u, 0u, 0u), ExtAppH2{c.my_rec, a.n_rows - 1, c.pol});
                        const int fa = __syncthreads_or((int)((par >> 15) & 1u));
                        const int fb = __syncthreads_or((int)(par >> 31));
                        c.last_fail = (fa ? 0x00008000u : 0u) | (fb ? 0x80000000u : 0u);
                    }
                    const uint32_t lf = c.last_fail;
                    if ((!fin_a && !(lf & 0x00008000u)) || (!fin_b && !(lf & 0x80000000u)))
                        fu = syndrome_bitsliced_h2<BG>(a.smem_base, h2_hard_bits(s_flag, a.cwpc), Z, a.n_rows, tid, a.row_start, !fin_a, !fin_b,
                                                       c.my_rec, c.pol);
                    else
                        fu = 0x80008000u;
                } else {
                    const bool staged = a.n_rows >= a.staged_min_rows;
                    int *s_flag2 = s_flag + 2 * a.cwpc;
                    if (!c.done) {
                        uint32_t f = syndrome_unrolled_core_h2<BG, FULL>(a, c.l);
                        if (!staged) f |= syndrome_unrolled_ext_h2<BG, FULL>(a, c.l, c.my_rec, c.pol);
                        if (f & 0x00008000u) s_flag[2 * slot] = 1;
                        if (f & 0x80000000u) s_flag[2 * slot + 1] = 1;
                    }
                    if (staged) {
                        // extension rows only for pairs with a live codeword whose core checks all hold
                        __syncthreads();
                        if (!c.done && ((!fin_a && !s_flag[2 * slot]) || (!fin_b && !s_flag[2 * slot + 1]))) {
                            const uint32_t f = syndrome_unrolled_ext_h2<BG, FULL>(a, c.l, c.my_rec, c.pol);
                            if (f & 0x00008000u) s_flag2[2 * slot] = 1;
                            if (f & 0x80000000u) s_flag2[2 * slot + 1] = 1;
                        }
                    }
                    __syncthreads();
                    if (!c.done) fu = ((s_flag[2 * slot] | s_flag2[2 * slot]) ? 0x00008000u : 0u) | ((s_flag[2 * slot + 1] | s_flag2[2 * slot + 1]) ? 0x80000000u : 0u);
                }
                if (!fin_a) {
                    ok_a = (fu & 0x00008000u) ? 0 : 1;
                    if (ok_a && a.early_term) {   // converged: freeze this codeword's outputs now
                        store_half_sel<FULL>(a, my_app, cwA, 0, ncw, K, z, Z, c.my_rec, c.pol);
                        fin_a = true;
                    }
                }
                if (!fin_b) {
                    ok_b = (fu & 0x80000000u) ? 0 : 1;
                    if (ok_b && a.early_term) {
                        store_half_sel<FULL>(a, my_app, cwB, 1, ncw, K, z, Z, c.my_rec, c.pol);
                        fin_b = true;
                    }
                }
                c.done = fin_a && fin_b;
                const int all_done = __syncthreads_and(c.done ? 1 : 0);  // also orders the flag reset below
                for (int i = tid; i < 2 * per_group; i += blockDim.x) s_flag[i] = 0;
                if (a.early_term && all_done) break;
            }
        }
        __syncthreads();

        if (active && !fin_a) store_half_sel<FULL>(a, my_app, cwA, 0, ncw, K, z, Z, c.my_rec, c.pol);
        if (has_b && !fin_b) store_half_sel<FULL>(a, my_app, cwB, 1, ncw, K, z, Z, c.my_rec, c.pol);
        if (active && z == 0) {
            if (a.iters) a.iters[cwA] = it_a;
            if (a.ok) a.ok[cwA] = (uint8_t)ok_a;
            if (has_b) {
                if (a.iters) a.iters[cwB] = it_b;
                if (a.ok) a.ok[cwB] = (uint8_t)ok_b;
            }
        }
    }
}

}  // namespace nrldpc
