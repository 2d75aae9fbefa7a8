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

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t x;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(addr));
    return x;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ float clamp_llr_h2(float x) { return __fadd_rn(fmaxf(fminf(x, kH2LlrMax), -kH2LlrMax), 0.0f); }

// One check row of degree DEG for check z of a codeword pair.  Record layout (uint4; three words
// when DEG <= 11, i.e. every layer but the four degree-19 ones of base graph 1):
//   x, y : alpha*min1, alpha*min2 of both codewords (packed fp16) with the row's sign product in their sign bits
//   z    : sign bits of the row's t values (message sign = row sign ^ sign(t_e)) on edges 0..min(DEG,16)-1: edge e at bit 15-(n0-1-e) of each half;
//          DEG <= 11: also the arg-min edge index of each codeword in bits 0..3 / 16..19
//   w    : DEG > 11 only: arg-min edge indices in bits 0..4 / 16..20 and, for DEG > 16, the sign bits of
//          edges 16..DEG-1 at the top of each half
// Arg-min indices are compared as fp16 bit patterns (HSET2 without flush-to-zero).
// syndrome with the base graph's shape known at compile time (see SyndromeRows in decode_kernel.cuh)
template <int BG, int R, bool FULL>
struct SyndromeRowsH2 {
    static __device__ __forceinline__ uint32_t run(const DecArgs &a, const Lane &l, uint32_t fail) {
        if (R >= 4 && R >= a.n_rows) return fail;
        constexpr int DEG = BgShape<BG>::deg(R);
        constexpr int E0 = BgShape<BG>::start(R);
        uint32_t par = 0;
#pragma unroll
        for (int e = 0; e < DEG; ++e) par ^= lds_u32(edge_addr<FULL>(l, a.ed[E0 + e], (R >= 4) && e == DEG - 1));
        fail |= par;
        asm volatile("" : "+r"(fail));   // one row's loads are consumed before the next row's are issued (register pressure)
        return SyndromeRowsH2<BG, R + 1, FULL>::run(a, l, fail);
    }
};
template <int BG, bool FULL>
struct SyndromeRowsH2<BG, BgShape<BG>::kRows, FULL> {
    static __device__ __forceinline__ uint32_t run(const DecArgs &, const Lane &, uint32_t fail) { return fail; }
};

// Bit-sliced syndrome of a codeword pair (see pack_hard_bits / syndrome_bitsliced in decode_kernel.cuh): the hard
// decisions of codeword A (bit 15 of every word) and B (bit 31) are packed into hb[0][col][Z/32] and hb[1][col][Z/32].
__device__ __forceinline__ void pack_hard_bits_h2(const uint32_t *app, uint32_t *hb, int Z, int n_cols, int n_cols_all, int z) {
    const int W = Z >> 5, w = z >> 5;
    for (int col = 0; col < n_cols; ++col) {
        const uint32_t x = app[col * Z + z];
        const uint32_t wa = __ballot_sync(0xffffffffu, (x >> 15) & 1u);
        const uint32_t wb = __ballot_sync(0xffffffffu, x >> 31);
        if ((z & 31) == 0) {
            hb[col * W + w] = wa;
            hb[(n_cols_all + col) * W + w] = wb;
        }
    }
}
// returns bit 15 set if codeword A fails, bit 31 if B fails (the convention of syndrome_unrolled_h2)
__device__ __noinline__ uint32_t syndrome_bitsliced_h2(const DecArgs &a, const uint32_t *hb, int z) {
    const int Z = a.Z, W = Z >> 5, z0 = z & ~31;
    const uint32_t *hbB = hb + a.ncols * W;
    uint32_t fa = 0, fb = 0;
    for (int r = z & 31; r < a.n_rows; r += 32) {
        uint32_t xa = 0, xb = 0;
        for (int e = a.row_start[r]; e < a.row_start[r + 1]; ++e) {
            const uint2 d = a.ed[e];
            int p = z0 + (int)(d.x >> 2);
            if (p >= Z) p -= Z;
            const uint32_t cb = (d.y - a.smem_base) >> 7;
            const int i0 = p >> 5, i1 = i0 + 1 == W ? 0 : i0 + 1;
            xa ^= __funnelshift_r(hb[cb + i0], hb[cb + i1], p & 31);
            xb ^= __funnelshift_r(hbB[cb + i0], hbB[cb + i1], p & 31);
        }
        fa |= xa;
        fb |= xb;
    }
    return (fa ? 0x00008000u : 0u) | (fb ? 0x80000000u : 0u);
}

template <int BG, bool FULL>
__device__ __noinline__ uint32_t syndrome_unrolled_h2(const DecArgs &a, const Lane l) {   // out of line: see decode_kernel.cuh
    return SyndromeRowsH2<BG, 0, FULL>::run(a, l, 0u);
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
    constexpr int N0 = DEG < 16 ? DEG : 16;
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
        const uint32_t e2 = (uint32_t)e * 0x00010001u;
        const uint32_t is_arg = __heq2_mask(as_h2(oargs), as_h2(e2));
        const uint32_t mag = bitselect(rec.x, rec.y, is_arg);
        const uint32_t sw = e < 16 ? rec.z << (N0 - 1 - e) : rec.w << (DEG - 1 - e);
        const uint32_t c = mag ^ (sw & kH2Sign);
        const __half2 tt = __hsub2(as_h2(x), as_h2(c));
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
        if (e < 16) s0 = bitselect(s0 >> 1, as_u32(tt), kH2Sign);
        else s1 = bitselect(s1 >> 1, as_u32(tt), kH2Sign);
    }
    s.m1 = m1; s.m2 = m2; s.sx = sx; s.s0 = s0; s.s1 = s1;
}

// second half: new messages, APP write-back, the row's new record
template <int DEG>
__device__ __forceinline__ uint4 row_scatter_h2(const RowStateH2<DEG> &s, const uint32_t alpha2) {
    constexpr int N0 = DEG < 16 ? DEG : 16;
    constexpr int N1 = DEG - N0;
    constexpr bool W3 = DEG <= 11;
    const __half2 m1raw = s.m1;
    const __half2 m1 = __hmin2(s.m1, as_h2(kH2MsgCap));
    const __half2 m2 = __hmin2(s.m2, as_h2(kH2MsgCap));
    // multiply by alpha carrying the row's sign product (m >= 0: bit-identical to (alpha*m) | sg, also for m = 0)
    const uint32_t alpha_s = bitselect(alpha2, s.sx, kH2Sign);
    uint32_t m1ss = as_u32(__hmul2(as_h2(alpha_s), m1));
    uint32_t m2ss = as_u32(__hmul2(as_h2(alpha_s), m2));
    asm volatile("" : "+r"(m1ss), "+r"(m2ss));
    uint32_t args = 0;
#pragma unroll
    for (int e = 0; e < DEG; ++e) {
        // arg-min edges found by value (ties: min2 == min1, every tied edge gets the same message)
        const uint32_t is_min = __heq2_mask(__habs2(as_h2(s.t[e])), m1raw);
        const uint32_t sel = bitselect(m1ss, m2ss, is_min);
        args = bitselect(args, (uint32_t)e * 0x00010001u, is_min);
        const uint32_t c = sel ^ (s.t[e] & kH2Sign);
        sts_u32(s.addr[e], as_u32(__hadd2(as_h2(s.t[e]), as_h2(c))));
    }
    constexpr uint32_t F0 = (((1u << N0) - 1u) << (16 - N0)) * 0x00010001u;
    constexpr uint32_t F1 = N1 > 0 ? (((1u << N1) - 1u) << (16 - N1)) * 0x00010001u : 0u;
    const uint32_t z = (s.s0 & F0) | (W3 ? args : 0u);
    const uint32_t w = W3 ? 0u : ((s.s1 & F1) | args);
    return make_uint4(m1ss, m2ss, z, w);
}

template <int DEG, bool IDENT_LAST, bool ONE_CW>
__device__ __forceinline__ uint4 process_row_h2(const Lane &l, const uint2 *__restrict__ ed, const uint4 rec,
                                                const uint32_t alpha2) {
    RowStateH2<DEG> s;
    row_gather_h2<DEG, IDENT_LAST, ONE_CW>(l, ed, rec, s);
    return row_scatter_h2<DEG>(s, alpha2);
}

// ---- pieces of the pair kernel ---------------------------------------------------------------------
struct DecCtxH2 {
    Lane l;
    uint32_t *my_rec;
    uint64_t pol;
    uint4 cur, cur2;   // prefetched records of the next layer (and of its partner when the next layer is a row pair)
    bool done;   // this thread does no row work (inactive lane, or both codewords of its pair are finished)
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
__device__ __forceinline__ void store_half(const DecArgs &a, const uint32_t *app, long long cw, int half, int ncw, int K,
                                           int lane, int nlanes) {
    const int sh = half ? 31 : 15;
    uint8_t *hard = a.hard + cw * K;
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
    if (a.soft) {
        float *dst = a.soft + cw * ncw;
        for (int i = lane; i < ncw; i += nlanes) {
            const float2 f = __half22float2(as_h2(app[i]));
            __stcs(dst + i, half ? f.y : f.x);
        }
    }
}

// syndrome of hard = (app < 0) over the active rows; bit 15 = codeword A fails, bit 31 = B fails
__device__ __forceinline__ uint32_t syndrome_fail_h2(const DecArgs &a, const DecCtxH2 &c) {
    uint32_t fail = 0;
    for (int r = 0; r < a.n_rows; ++r) {
        uint32_t par = 0;
        for (int e = a.row_start[r]; e < a.row_start[r + 1]; ++e) par ^= lds_u32(edge_addr<false>(c.l, a.ed[e]));
        fail |= par;
    }
    return fail & kH2Sign;
}

// Layer loop, unrolled per base graph.  As in decode_kernel.cuh: ld_from / ld_to gate the prefetch of the next
// layer's record, and two consecutive base rows that touch disjoint block columns (pair_first<BG>) run as one layer
// with one barrier.  Row pairs only occur among rows of degree <= 11, i.e. with three-word records.
template <int BG, int R, bool FULL>
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
                const uint4 rec0 = row_scatter_h2<DEG>(s0, a.alpha_h2);
                const uint4 rec1 = row_scatter_h2<DEG2>(s1, a.alpha_h2);
                if (store_rec) {
                    st_rec(c.my_rec, R, rec0, c.pol);
                    st_rec(c.my_rec, R + 1, rec1, c.pol);
                }
                c.cur = nxt;
                c.cur2 = nxt2;
            }
            __syncthreads();
            UnrolledRowsH2<BG, (PAIR ? R + 2 : BgShape<BG>::kRows), FULL>::run(a, c, ld_from, ld_to, store_rec);
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
            if (!PAIR) UnrolledRowsH2<BG, R + 1, FULL>::run(a, c, ld_from, ld_to, store_rec);
        }
    }
};
template <int BG, bool FULL>
struct UnrolledRowsH2<BG, BgShape<BG>::kRows, FULL> {
    static __device__ __forceinline__ void run(const DecArgs &, DecCtxH2 &, int, int, bool) {}
};

// Here cwpc counts codeword PAIRS per CTA; FULL = one pair per CTA and every thread owns a check.
template <int BG, bool FULL>
__global__ void __launch_bounds__(kDecThreads, kDecCtasPerSm) decode_nms_h2_kernel(const __grid_constant__ DecArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Z = a.Z;
    const int ncw = a.ncols * Z;
    const int K = a.kcols * Z;
    uint32_t *app = reinterpret_cast<uint32_t *>(smem_raw);
    int *s_flag = reinterpret_cast<int *>(app + (size_t)a.cwpc * a.slot_stride);  // [2*cwpc] + work-group slot
    int &s_group = s_flag[2 * a.cwpc];
    // FULL kernels: packed hard decisions of both codewords, behind the work slot and the (unused here) barrier slot
    uint32_t *hb = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(s_flag + 2 * a.cwpc + 1) + 7) & ~(uintptr_t)7) + 2;
    if ((uint32_t)__cvta_generic_to_shared(smem_raw) != a.smem_base) __trap();

    const int tid = threadIdx.x;
    const int slot = tid / Z;
    const int z = tid - slot * Z;
    const bool lane_ok = tid < a.cwpc * Z;
    const int per_group = 2 * a.cwpc;
    const long long n_groups = (a.batch + per_group - 1) / per_group;
    const bool want_ok = a.ok != nullptr;

    DecCtxH2 c;
    c.l.zoff = (uint32_t)z * 4u;
    c.l.nZ4 = 0u - (uint32_t)Z * 4u;
    c.l.slot_off = FULL ? 0u : (uint32_t)(slot * a.slot_stride) * 4u;
    c.l.one = (uint32_t)a.one;
    c.my_rec = a.c2v + (size_t)blockIdx.x * (kRecWords * kRecStride) + tid;
    c.pol = make_l2_policy(a.l2_pin);
    uint32_t *my_app = app + (size_t)(lane_ok ? slot : 0) * a.slot_stride;

    while (true) {
        __syncthreads();  // previous group's outputs are out of smem
        if (tid == 0) s_group = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const long long group = s_group;
        if (group >= n_groups) break;
        const long long cw0 = group * per_group;
        const int n_here = (int)min((long long)per_group, a.batch - cw0);  // codewords in this group

        const long long cwA = cw0 + 2 * slot, cwB = cwA + 1;
        const bool active = lane_ok && 2 * slot < n_here;
        const bool has_b = lane_ok && 2 * slot + 1 < n_here;
        if (active) load_pair(a.llr + cwA * ncw, has_b ? a.llr + cwB * ncw : nullptr, my_app, ncw, z, Z);
        if (tid < per_group) s_flag[tid] = 0;
        __syncthreads();

        bool fin_a = !active, fin_b = !has_b;   // finished (converged and already written out, or absent)
        c.done = !active;
        c.cur = make_uint4(0u, 0u, 0u, 0u);
        c.cur2 = c.cur;
        int it_a = 0, it_b = 0, ok_a = 0, ok_b = 0;

        for (int it = 0; it < a.max_iters; ++it) {
            const bool first = it == 0, last = it + 1 == a.max_iters;
            UnrolledRowsH2<BG, 0, FULL>::run(a, c, first ? a.n_rows - 1 : 0, last ? a.n_rows - 1 : a.n_rows, !last);
            if (!fin_a) it_a = it + 1;
            if (!fin_b) it_b = it + 1;
            if (a.early_term || (want_ok && last)) {
                if (FULL && a.n_rows >= kBitslicedSyndromeMinRows) {
                    pack_hard_bits_h2(app, hb, Z, min(a.ncols, a.kcols + a.n_rows), a.ncols, tid);
                    __syncthreads();
                    const uint32_t f = syndrome_bitsliced_h2(a, hb, tid);
                    if (f & 0x00008000u) s_flag[0] = 1;
                    if (f & 0x80000000u) s_flag[1] = 1;
                } else if (!c.done) {
                    const uint32_t f = syndrome_unrolled_h2<BG, FULL>(a, c.l) & kH2Sign;
                    if (f & 0x00008000u) s_flag[2 * slot] = 1;
                    if (f & 0x80000000u) s_flag[2 * slot + 1] = 1;
                }
                __syncthreads();
                if (!fin_a) {
                    ok_a = s_flag[2 * slot] ? 0 : 1;
                    if (ok_a && a.early_term) {   // converged: freeze this codeword's outputs now
                        store_half(a, my_app, cwA, 0, ncw, K, z, Z);
                        fin_a = true;
                    }
                }
                if (!fin_b) {
                    ok_b = s_flag[2 * slot + 1] ? 0 : 1;
                    if (ok_b && a.early_term) {
                        store_half(a, my_app, cwB, 1, ncw, K, z, Z);
                        fin_b = true;
                    }
                }
                c.done = fin_a && fin_b;
                const int all_done = __syncthreads_and(c.done ? 1 : 0);  // also orders the flag reset below
                if (tid < per_group) s_flag[tid] = 0;
                if (a.early_term && all_done) break;
            }
        }
        __syncthreads();

        if (active && !fin_a) store_half(a, my_app, cwA, 0, ncw, K, z, Z);
        if (has_b && !fin_b) store_half(a, my_app, cwB, 1, ncw, K, z, Z);
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
