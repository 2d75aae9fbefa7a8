// decode_kernel.cuh -- layered normalized min-sum decoder for the QC-LDPC codes of TS 38.212.
//
// Replaces the arithmetic behind step(obj.hLDPCDecoder, cw_tilde) (NRLDPCDecoder.m:265) on the
// lifted matrix of get_pcm.m:8 (check r*Z+z touches variable c*Z + (z + V mod Z) mod Z).
//
// Mapping (sm_100a, no tensor cores -- sparse message passing):
//   * one CTA owns `cwpc` codewords at a time (1 for Z >= 193, floor(384/Z) for smaller Z) and is
//     persistent: it pulls codeword groups from a device work counter until the batch is done;
//   * thread (slot, z) owns check row z of every base row ("layer") of codeword `slot`: inside a
//     layer each block column appears at most once and every circulant is a permutation, so the Z
//     checks of a layer touch disjoint variables -> no atomics, one __syncthreads per layer;
//   * the a-posteriori LLRs (cols*Z floats per codeword, 104 KB at BG1/Z=384) live in shared
//     memory for all iterations; two CTAs fit per SM so one CTA's barriers/loads hide under the
//     other's arithmetic;
//   * check-to-variable messages are kept compressed (alpha*min1, alpha*min2, argmin, sign bits)
//     as one three-word record per check in a per-CTA global scratch that is pinned in L2
//     (evict_last cache policy), software-prefetched one layer ahead; the first iteration reads
//     nothing and the last writes nothing;
//   * consecutive base rows that touch disjoint block columns run as one layer (one barrier);
//   * 'Parity check satisfied' (the reference's stopping rule, NRLDPCDecoder.m:120): the last layer's
//     barrier doubles as a CTA-wide OR of its (final) checks, and only when they all hold does the
//     exact syndrome run, core rows first (bit-sliced for one codeword per CTA);
//   * the base graph's shape (row degrees, edge order, identity extension columns) is a
//     compile-time constant per base graph: the layer loop is fully unrolled, and the per-edge
//     offsets / wrap thresholds are read straight from the kernel-parameter constant bank as
//     instruction operands (no descriptor loads);
//   * arithmetic is float32 with every add/mul individually rounded (__fsub_rn/__fmul_rn/__fadd_rn:
//     no FMA contraction), signs handled as sign BITS, so results are bit-identical to the CPU
//     oracle (oracle/nrldpc_oracle.c, function orc_decode_nms).
//
// The kernel is bound by the SM's ALU pipe (integer / logic / min-max / select instructions), not
// by HBM, so the row update is written to minimise ALU-pipe instructions:
// address arithmetic is issued as IMAD (FMA pipe) around one unsigned min for the circulant wrap,
// the arg-min index is not tracked while scanning (the second pass finds it by value), sign bits
// are gathered with funnel shifts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nrldpc {

#ifndef NRLDPC_DEC_THREADS           // build-time experiment knobs (make EXTRA="-DNRLDPC_DEC_THREADS=768 -DNRLDPC_DEC_CTAS=1")
#define NRLDPC_DEC_THREADS 384
#define NRLDPC_DEC_CTAS 2
#endif
constexpr int kDecThreads = NRLDPC_DEC_THREADS;     // max threads per decode CTA (= largest lifting size)
constexpr int kDecCtasPerSm = NRLDPC_DEC_CTAS;
constexpr float kLlrMax = 1048576.0f;
constexpr int kMaxEdges = 316;
constexpr int kMaxRows = 46;
// c2v scratch of one CTA: 32-bit words [word index][kRecStride threads].  A layer's record is three words
// (word index 3*slot + j); layer r >= 1 lives in slot r, layer 0 in slot n_rows, so that "the next layer's
// record" is always the next slot, also across the iteration boundary.  The packed-half kernel keeps a
// fourth word for its degree-19 layers (0..3) behind the slots.  296 CTAs x 222,720 B = 66 MB: small enough
// to stay pinned in L2 (max persisting L2 on B200: 79 MB), which 16-byte records (85 MB) were measured not to be.
constexpr int kRecStride = kDecThreads;
constexpr int kRecSlots = kMaxRows + 1;
constexpr int kRecWords = kRecSlots * 3 + 4;

// Edge descriptor, read from the kernel-parameter constant bank:
//   x = shift*4               (bytes): lane z reads circulant position (z + shift) mod Z
//   y = smem_base + col*Z*4   (bytes): shared-window address of the block column inside the CTA's first
//                                      APP array (smem_base = shared address of the dynamic shared memory,
//                                      probed once per handle; the kernel traps if it ever differs)
struct DecArgs {
    const float *llr;        // [batch][ncw]
    uint8_t *hard;           // [batch][K]
    float *soft;             // [batch][ncw] or null
    int32_t *iters;          // [batch] or null
    uint8_t *ok;             // [batch] or null
    long long batch;
    int Z, ncols, kcols, n_rows, n_edges, max_iters, early_term, cwpc;
    int slot_stride;         // words between the APP arrays of two codewords (pairs) of a CTA: cols*Z + pad, see decode_slot_stride
    float alpha;
    int l2_pin;              // 1: c2v scratch accesses carry an L2 evict_last policy
    int bitsliced_min_rows;  // FULL kernels: bit-sliced syndrome from this many active rows on (else per-thread unrolled)
    int staged_min_rows;     // per-thread syndrome: core rows first, extension rows only if they hold, from this many rows on
    int one;                 // = 1 (see mad_u32)
    uint32_t smem_base;      // shared-window address of the kernel's dynamic shared memory
    uint32_t alpha_h2;       // {fp16(alpha), fp16(alpha)} for the packed-half kernel
    int spares;              // decode_nms_refill2_kernel: surplus codeword buffers (mailboxes) per CTA
    int rec_group;           // CTAs that share one [kRecWords][kRecStride] scratch block (CTA b uses thread columns (b % rec_group) * blockDim.x ...)
    uint32_t *c2v;           // [ceil(grid / rec_group)][kRecWords][kRecStride]; float32 record = {alpha*min1|sgn, alpha*min2|sgn, argmin | signbits << 5}
    int *work_counter;       // device ticket counter: never reset between launches, the host passes the value it has reached ...
    unsigned int work_base;  // ... so ticket - work_base is this launch's group (codeword) index: no memset node per launch
    unsigned short row_start[kMaxRows + 2];
    uint2 ed[kMaxEdges];
};

template <int BG> struct BgShape;
template <> struct BgShape<1> {
    static constexpr int kRows = NRLDPC_BG1_ROWS;
    static constexpr int kCols = 68, kKcols = 22;   // block columns, information columns (NRLDPC.m:414-454)
    static __host__ __device__ constexpr int deg(int r) { return nrldpc_bg1_deg[r]; }
    static __host__ __device__ constexpr int start(int r) { return nrldpc_bg1_start[r]; }
    static __host__ __device__ constexpr int col(int e) { return nrldpc_bg1_col[e]; }
};
template <> struct BgShape<2> {
    static constexpr int kRows = NRLDPC_BG2_ROWS;
    static constexpr int kCols = 52, kKcols = 10;
    static __host__ __device__ constexpr int deg(int r) { return nrldpc_bg2_deg[r]; }
    static __host__ __device__ constexpr int start(int r) { return nrldpc_bg2_start[r]; }
    static __host__ __device__ constexpr int col(int e) { return nrldpc_bg2_col[e]; }
};

// Row orthogonality of the TS 38.212 base graphs: from row 20 on (and for a few earlier rows) consecutive base rows
// touch DISJOINT block columns, so the two layers update disjoint variables and may run as one layer with a single
// barrier -- the result is bit-identical to running them one after the other.  Pairs are formed greedily in row
// order: BG1 (16,17), (20,21), (22,23) ... (44,45): 46 layers -> 32 barriers; BG2 (11,12), (17,18), (20,21) ... : 42 -> 29.
template <int BG>
__host__ __device__ constexpr bool rows_disjoint(int r) {   // rows r and r + 1
    if (r < 0 || r + 1 >= BgShape<BG>::kRows) return false;
    for (int i = BgShape<BG>::start(r); i < BgShape<BG>::start(r + 1); ++i)
        for (int j = BgShape<BG>::start(r + 1); j < BgShape<BG>::start(r + 2); ++j)
            if (BgShape<BG>::col(i) == BgShape<BG>::col(j)) return false;
    return true;
}
template <int BG>
__host__ __device__ constexpr bool pair_first(int r) {      // row r opens a pair (r, r + 1)
    bool first = false;
    for (int q = 0; q <= r; ++q) first = !first && rows_disjoint<BG>(q);   // first(q) = !first(q-1) && disjoint(q)
    return first;
}

__device__ __forceinline__ float clamp_llr(float x) {
    // NaN marks filler upstream (NRLDPCDecoder.m:224,264): fminf(NaN, M) = M.  The final + 0 turns -0 into +0:
    // no APP value is ever -0 afterwards, so a hard decision (app < 0) is exactly the sign bit.
    return __fadd_rn(fmaxf(fminf(x, kLlrMax), -kLlrMax), 0.0f);
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float x;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
    return x;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t x;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(addr));
    return x;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// c2v scratch accesses: L1-bypassing, with an L2 cache policy (evict_last pins the scratch in L2 so
// that the streamed LLR input cannot push dirty records out to HBM).
__device__ __forceinline__ uint64_t make_l2_policy(int pin) {
    uint64_t p;
    if (pin) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint32_t ld_word(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_word(uint32_t *p, uint32_t v, uint64_t pol) {
    asm volatile("st.global.cg.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
// three-word record of `slot` in this thread's column (p = word 0 of the column)
__device__ __forceinline__ uint4 ld_rec(const uint32_t *p, int slot, uint64_t pol) {
    uint4 v;
    v.x = ld_word(p + (slot * 3 + 0) * kRecStride, pol);
    v.y = ld_word(p + (slot * 3 + 1) * kRecStride, pol);
    v.z = ld_word(p + (slot * 3 + 2) * kRecStride, pol);
    v.w = 0u;
    return v;
}
__device__ __forceinline__ void st_rec(uint32_t *p, int slot, uint4 v, uint64_t pol) {
    st_word(p + (slot * 3 + 0) * kRecStride, v.x, pol);
    st_word(p + (slot * 3 + 1) * kRecStride, v.y, pol);
    st_word(p + (slot * 3 + 2) * kRecStride, v.z, pol);
}

// A-posteriori word of the degree-1 parity variable of extension row `row` for this thread's check: its channel value
// (what shared memory holds) plus the row's latest message to it, rebuilt from the record the row wrote in the last
// iteration (arg = 31: the degree-1 edge is the arg-min and receives alpha*min2, otherwise alpha*min1; sign = row sign
// product ^ own sign).  Records of rows >= 1 live in slot `row`.
__device__ __forceinline__ uint32_t ext_app(const uint32_t *my_rec, int row, uint64_t pol, uint32_t chan) {
    const uint4 rec = ld_rec(my_rec, row, pol);
    const uint32_t sel = ((rec.z & 31u) == 31u) ? rec.y : rec.x;
    return __float_as_uint(__fadd_rn(__uint_as_float(chan), __uint_as_float(sel ^ (chan & 0x80000000u))));
}

// The same for ALL active extension rows at once, as a bit mask (bit r = hard decision of row r's degree-1 parity variable,
// r = 4 .. n_rows-1).  The records are fetched eight rows at a time (24 loads in flight): taken one row after the other,
// each record is an L2 round trip, and 42 of them in sequence cost about 15 us per call -- several per cent of a codeword
// under the parity-check stop (measured: BASELINE config 3 with the stop ran SLOWER at its operating point than with the
// stop never taken).  p_addr: shared-window address of this thread's entry of row 4's parity column; col_bytes = 4*Z.
__device__ __noinline__ unsigned long long ext_hard_mask(const uint32_t *my_rec, const uint64_t pol, const int n_rows, const uint32_t p_addr,
                                                         const uint32_t col_bytes) {
    unsigned long long m = 0ull;
    for (int r0 = 4; r0 < n_rows; r0 += 8) {
        uint32_t x[8], y[8], meta[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = min(r0 + i, n_rows - 1);       // clamped: the surplus loads of the last chunk repeat its last row
            x[i] = ld_word(my_rec + (r * 3 + 0) * kRecStride, pol);
            y[i] = ld_word(my_rec + (r * 3 + 1) * kRecStride, pol);
            meta[i] = ld_word(my_rec + (r * 3 + 2) * kRecStride, pol);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (r0 + i < n_rows) {
                const uint32_t chan = lds_u32(p_addr + (uint32_t)(r0 + i - 4) * col_bytes);
                const uint32_t sel = ((meta[i] & 31u) == 31u) ? y[i] : x[i];
                const uint32_t app = __float_as_uint(__fadd_rn(__uint_as_float(chan), __uint_as_float(sel ^ (chan & 0x80000000u))));
                m |= (unsigned long long)(app >> 31) << (r0 + i);
            }
        }
    }
    return m;
}

// Integer multiply-add whose multiplier is the kernel parameter `one` (always 1): the compiler
// cannot fold it, so it is emitted as IMAD, which issues on the FMA pipe -- address arithmetic
// stays off the saturated ALU pipe.
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t one, uint32_t b) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
}

// (a & ~m) | (b & m) in one LOP3
__device__ __forceinline__ uint32_t bitselect(uint32_t a, uint32_t b, uint32_t m) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
}

// Per-thread addressing state.
//   zoff  = z*4                        lane position inside a circulant, bytes
//   nZ4   = -(Z*4)                     (two's complement)
//   slot_off = byte offset of this thread's codeword's APP array inside the CTA's APP storage (0 when FULL)
//   one   = 1, from the parameter bank
struct Lane {
    uint32_t zoff, nZ4, slot_off, one;
};

// Shared-memory address of the variable edge `d` connects to check z.
//   u = (z + shift)*4 wraps at Z*4: min(u, u - Z*4) as unsigned (u - Z*4 underflows when u < Z*4)
// ONE_CW: one codeword per CTA (slot_off = 0): the uniform d.y folds into the LDS/STS address operand.
template <bool ONE_CW>
__device__ __forceinline__ uint32_t edge_addr(const Lane &l, const uint2 d, const bool ident = false) {
    uint32_t m = l.zoff;
    if (!ident) {
        const uint32_t u = mad_u32(l.zoff, l.one, d.x);
        const uint32_t w = mad_u32(u, l.one, l.nZ4);
        m = min(u, w);
    }
    if (!ONE_CW) m = mad_u32(m, l.one, l.slot_off);
    return m + d.y;
}

// One check row of degree DEG for check z.  (om1, om2, ometa) is the compressed message record
// written for this check in the previous iteration: om1/om2 = alpha*min1, alpha*min2 with the row's
// sign product sg in their sign bit, ometa = arg-min index in bits 0..4 and the sign bits of the
// row's t values MSB-first above it (edge e at bit 5 + DEG-1-e): message e = (e == arg ? om2 : om1)
// with its sign bit flipped by sign(t_e), i.e. sign(c_e) = sg ^ sign(t_e).
// In the first iteration the record is all zeros (previous messages +0: x - (+0) = x exactly).
// IDENT_LAST: the row's last edge is an identity circulant onto a DEGREE-1 variable (extension parity column: shift 0
// for every lifting-size set and no other check touches it -- TS 38.212 tables).  Such a variable's
// a-posteriori value is its channel value plus this check's own message, so t = app - c IS the channel value: the edge
// enters the row with the value that sits in shared memory (never overwritten), no wrap, no message rebuild, no sign
// record, no write-back (oracle A revision 2).  Its a-posteriori value chan + c' is formed only where it is read: the
// last layer's parity filter, the syndrome's extension stage and the soft output (ext_app), from the row's record --
// arg = 31 in a record means "no edge with other checks attains the minimum", i.e. the degree-1 edge is the arg-min.
template <int DEG>
struct RowState {
    float t[DEG];
    uint32_t addr[DEG];
    float m1, m2;
    uint32_t sx, ts;
};

// first half of a row update: addresses, loads, t_e = app - c_e, the two minima and the sign bits
template <int DEG, bool IDENT_LAST, bool ONE_CW>
__device__ __forceinline__ void row_gather(const Lane &l, const uint2 *__restrict__ ed, const uint32_t om1,
                                           const uint32_t om2, const uint32_t ometa, RowState<DEG> &s) {
    constexpr int NE = IDENT_LAST ? DEG - 1 : DEG;   // edges whose variable has other checks too
    float m1 = 0.f, m2 = 0.f;
    uint32_t ts = 0, tp = 0;
    const uint32_t oarg = ometa & 31u;
#pragma unroll
    for (int e = 0; e < DEG; ++e) {
        const uint2 d = ed[e];
        const uint32_t a = edge_addr<ONE_CW>(l, d, IDENT_LAST && e == DEG - 1);
        s.addr[e] = a;
        const float x = lds_f32(a);
        float tt = x;                                  // degree-1 variable: its channel value
        if (e < NE) {
            const uint32_t mag = (oarg == (uint32_t)e) ? om2 : om1;
            // stored minimum (carries sg) with its sign bit flipped by sign(t_e) (meta bit 5 + NE-1-e -> bit 31): one LOP3
            const uint32_t c = mag ^ ((ometa << (26 - (NE - 1 - e))) & 0x80000000u);
            tt = __fsub_rn(x, __uint_as_float(c));
        }
        s.t[e] = tt;
        const float ab = fabsf(tt);
        if (e == 0) {
            m1 = ab;
        } else if (e == 1) {
            m2 = fmaxf(m1, ab);
            m1 = fminf(m1, ab);
        } else {
            m2 = fminf(m2, fmaxf(ab, m1));
            m1 = fminf(m1, ab);
        }
        if (e < NE) ts = __funnelshift_l(__float_as_uint(tt), ts, 1);  // ts = ts << 1 | signbit(tt)
        else tp = __float_as_uint(tt);
    }
    // the row's sign product = parity of the number of negative t's: one POPC over the collected sign bits per ROW instead of
    // one XOR per EDGE (the degree-1 edge's sign is not recorded in ts: it is folded in before the count)
    const uint32_t all = IDENT_LAST ? __funnelshift_l(tp, ts, 1) : ts;
    s.m1 = m1; s.m2 = m2; s.sx = (uint32_t)__popc(all) << 31; s.ts = ts;
}

// second half: new messages, APP write-back, the row's new record.  PAR: also XOR the new APP values into `par` -- its sign
// bit is this check's parity on the hard decisions just written (used for the last layer of an iteration, whose decisions
// are final: see UnrolledRows)
// TRACK >= 0 (extension rows of the multi-codeword kernels under the parity-check stop): the hard decision of the row's
// degree-1 parity variable is kept as bit TRACK of `ext_word`, so that the extension stage of the syndrome needs no record
// from the L2 scratch (with several codewords per CTA that stage runs in most passes, and every other slot waits for it)
template <int DEG, bool IDENT_LAST, bool PAR, int TRACK = -1>
__device__ __forceinline__ uint4 row_scatter_par(const RowState<DEG> &s, const float alpha, uint32_t &par, const bool live = true,
                                                 uint32_t *ext_word = nullptr) {
    constexpr int NE = IDENT_LAST ? DEG - 1 : DEG;
    // both candidate magnitudes with the row's sign product folded in: multiply by alpha carrying the sign
    // (m >= 0, so the product's sign bit is sg also when m = 0: bit-identical to (alpha*m) | sg)
    const float alpha_s = __uint_as_float(bitselect(__float_as_uint(alpha), s.sx, 0x80000000u));
    uint32_t m1ss = __float_as_uint(__fmul_rn(alpha_s, s.m1));
    uint32_t m2ss = __float_as_uint(__fmul_rn(alpha_s, s.m2));
    asm volatile("" : "+r"(m1ss), "+r"(m2ss));  // keep the sign folded per row, not re-derived per edge
    uint32_t arg = IDENT_LAST ? 31u : 0u;       // 31: none of the recorded edges attains the minimum (the degree-1 edge does)
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        // the arg-min edge is found by value: on a tie min2 == min1, so every tied edge gets the same message
        const bool is_min = fabsf(s.t[e]) == s.m1;
        const uint32_t sel = is_min ? m2ss : m1ss;
        arg = is_min ? (uint32_t)e : arg;
        const float c = __uint_as_float(sel ^ (__float_as_uint(s.t[e]) & 0x80000000u));
        const float app = __fadd_rn(s.t[e], c);
        if (PAR) par ^= __float_as_uint(app);
        if (live) sts_f32(s.addr[e], app);
    }
    if ((PAR || TRACK >= 0) && IDENT_LAST) {   // the degree-1 variable's a-posteriori value: for the check's parity / its hard decision only
        const float tp = s.t[DEG - 1];
        const uint32_t sel = fabsf(tp) == s.m1 ? m2ss : m1ss;
        const uint32_t app_p = __float_as_uint(__fadd_rn(tp, __uint_as_float(sel ^ (__float_as_uint(tp) & 0x80000000u))));
        if (PAR) par ^= app_p;
        if (TRACK >= 0) *ext_word = bitselect(*ext_word, app_p >> (31 - (TRACK >= 0 ? TRACK : 0)), 1u << (TRACK >= 0 ? TRACK : 0));
    }
    return make_uint4(m1ss, m2ss, arg | (s.ts << 5), 0u);
}
template <int DEG, bool IDENT_LAST>
__device__ __forceinline__ uint4 row_scatter(const RowState<DEG> &s, const float alpha, const bool live = true) {
    uint32_t unused = 0;
    return row_scatter_par<DEG, IDENT_LAST, false>(s, alpha, unused, live);
}

template <int DEG, bool IDENT_LAST, bool ONE_CW, int TRACK = -1>
__device__ __forceinline__ uint4 process_row(const Lane &l, const uint2 *__restrict__ ed, const uint32_t om1,
                                             const uint32_t om2, const uint32_t ometa, const float alpha, const bool live = true,
                                             uint32_t *ext_word = nullptr) {
    RowState<DEG> s;
    row_gather<DEG, IDENT_LAST, ONE_CW>(l, ed, om1, om2, ometa, s);
    uint32_t unused = 0;
    return row_scatter_par<DEG, IDENT_LAST, false, TRACK>(s, alpha, unused, live, ext_word);
}

// ---- pieces shared by all kernel variants --------------------------------------------------------
struct DecCtx {
    Lane l;
    uint32_t *my_rec;   // word 0 of this thread's record column
    uint64_t pol;
    uint4 cur, cur2; // prefetched records of the next layer (and of its partner when the next layer is a pair)
    bool done;       // this thread does no row work (inactive lane, or its codeword has converged)
    int last_fail;   // FULL kernels, every base row active: some check of the iteration's last layer is unsatisfied (CTA-uniform)
    uint32_t ext_lo, ext_hi;   // TRACK kernels: hard decisions of the degree-1 parity variables, bit (r - 4) of the pair = extension row r
    uint4 last_rec;  // FULL kernels, trimmed row count: the record the last ACTIVE row wrote in this iteration (kept in registers
                     // for last_row_parity: reloading it from the L2 scratch cost one L2 round trip per iteration)
};

// TMA staging of a codeword group: the group's rows are contiguous in HBM, so ONE bulk asynchronous copy
// (cp.async.bulk -> UBLKCP, completion counted on an mbarrier) brings all cols*Z circulant blocks into shared
// memory without passing through registers; the clamp (+-LLR_MAX, NaN filler, -0) is then applied in place.
__device__ __forceinline__ void load_group(const DecArgs &a, float *app, long long cw0, int n_here, int ncw,
                                           uint64_t *bar, uint32_t &parity) {
    if (a.slot_stride != ncw) {
        // padded slots (several codewords per CTA, see decode_slot_stride): the group is still one contiguous range in
        // HBM; it is read with coalesced loads and scattered to the padded slots with the clamp applied on the way
        const float *src = a.llr + cw0 * ncw;
        const int S = a.slot_stride;
        if ((S & 3) == 0) {          // 16-byte aligned slots: one bulk copy per codeword row, clamp applied in place
            const uint32_t row_bytes = (uint32_t)ncw * 4u;
            if (threadIdx.x == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, row_bytes * (uint32_t)n_here);
                for (int sl = 0; sl < n_here; ++sl) bulk_g2s_stream(app + (size_t)sl * S, src + (size_t)sl * ncw, row_bytes, bar);
            }
            mbar_wait(bar, parity);
            parity ^= 1u;
            const int ncw4 = ncw >> 2, n4 = n_here * ncw4;
            for (int i = threadIdx.x; i < n4; i += blockDim.x) {
                const int sl = i / ncw4;
                float4 *p = reinterpret_cast<float4 *>(app + (size_t)sl * S) + (i - sl * ncw4);
                float4 v = *p;
                v.x = clamp_llr(v.x); v.y = clamp_llr(v.y); v.z = clamp_llr(v.z); v.w = clamp_llr(v.w);
                *p = v;
            }
        } else {
            const int n = n_here * ncw;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int sl = i / ncw;
                app[(size_t)sl * S + (i - sl * ncw)] = clamp_llr(__ldcs(src + i));
            }
        }
        return;
    }
    const uint32_t bytes = (uint32_t)(n_here * ncw) * 4u;     // ncw*4 is a multiple of 16 for every (BG, Z)
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy accesses to app vs the async write
        mbar_expect_tx(bar, bytes);
        bulk_g2s_stream(app, a.llr + cw0 * ncw, bytes, bar);   // evict_first: must not displace the pinned c2v scratch
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    float4 *dst = reinterpret_cast<float4 *>(app);
    const int n4 = (n_here * ncw) >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 v = dst[i];
        v.x = clamp_llr(v.x); v.y = clamp_llr(v.y); v.z = clamp_llr(v.z); v.w = clamp_llr(v.w);
        dst[i] = v;
    }
}

// exact syndrome of hard = (app < 0) over the active rows r0 <= r < r1 ('Parity check satisfied', NRLDPCDecoder.m:120)
__device__ __forceinline__ uint32_t syndrome_fail(const DecArgs &a, const DecCtx &c, int r0, int r1) {
    uint32_t fail = 0;
    unsigned long long mask = 0ull;
    if (r1 > 4) mask = ext_hard_mask(c.my_rec, c.pol, a.n_rows, a.smem_base + c.l.slot_off + c.l.zoff + (uint32_t)((a.kcols + 4) * a.Z) * 4u, (uint32_t)a.Z * 4u);
    for (int r = r0; r < r1; ++r) {
        uint32_t par = 0;
        const int e1 = a.row_start[r + 1] - (r >= 4 ? 1 : 0);
        for (int e = a.row_start[r]; e < e1; ++e) par ^= lds_u32(edge_addr<false>(c.l, a.ed[e]));
        if (r >= 4) par ^= (uint32_t)((mask >> r) & 1ull) << 31;   // degree-1 parity variable
        fail |= par;
    }
    return fail;
}

// XOR of the a-posteriori words of check z of the LAST active base row (sign bit(s) = its parity); out of line so that the
// layer code's register allocation does not see it
// rec: the record that row wrote in this iteration (DecCtx::last_rec); EXT_APP: the float32 / packed-half routine that
// rebuilds the degree-1 parity variable's a-posteriori word from it
template <typename EXT_APP>
__device__ __noinline__ uint32_t last_row_parity(const DecArgs &a, const Lane l, const uint4 rec, EXT_APP ext) {
    uint32_t par = 0;
    const int r = a.n_rows - 1, e1 = a.row_start[a.n_rows];
    for (int e = a.row_start[r]; e < e1; ++e) {
        uint32_t w = lds_u32(edge_addr<false>(l, a.ed[e]));
        if (r >= 4 && e == e1 - 1) w = ext(rec, w);
        par ^= w;
    }
    return par;
}
__device__ __forceinline__ uint32_t ext_app_rec(const uint4 rec, uint32_t chan) {
    const uint32_t sel = ((rec.z & 31u) == 31u) ? rec.y : rec.x;
    return __float_as_uint(__fadd_rn(__uint_as_float(chan), __uint_as_float(sel ^ (chan & 0x80000000u))));
}
struct ExtAppF32 {
    __device__ __forceinline__ uint32_t operator()(const uint4 rec, uint32_t chan) const { return ext_app_rec(rec, chan); }
};

__device__ __forceinline__ void store_outputs(const DecArgs &a, const float *app, long long cw0, int n_here, int ncw, int K) {
    // hard decisions, four per 32-bit store (K = kcols*Z is a multiple of 4 for every even Z; odd Z stores bytes)
    if ((K & 3) == 0 && (a.slot_stride & 3) == 0) {
        const int K4 = K >> 2;
        for (int s = 0; s < n_here; ++s) {
            const float4 *src = reinterpret_cast<const float4 *>(app + (size_t)s * a.slot_stride);
            uint32_t *dst = reinterpret_cast<uint32_t *>(a.hard + (cw0 + s) * K);
            for (int k = threadIdx.x; k < K4; k += blockDim.x) {
                const float4 v = src[k];
                dst[k] = (__float_as_uint(v.x) >> 31) | ((__float_as_uint(v.y) >> 31) << 8) |
                         ((__float_as_uint(v.z) >> 31) << 16) | ((__float_as_uint(v.w) >> 31) << 24);
            }
        }
    } else {
        for (int s = 0; s < n_here; ++s) {
            const float *src = app + (size_t)s * a.slot_stride;
            uint8_t *dst = a.hard + (cw0 + s) * K;
            for (int k = threadIdx.x; k < K; k += blockDim.x) dst[k] = src[k] < 0.0f ? 1 : 0;
        }
    }
    if (a.soft && a.slot_stride == ncw) {
        float4 *dst = reinterpret_cast<float4 *>(a.soft + cw0 * ncw);
        const float4 *src = reinterpret_cast<const float4 *>(app);
        const int n4 = (n_here * ncw) >> 2;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) __stcs(dst + i, src[i]);
    } else if (a.soft) {
        float *dst = a.soft + cw0 * ncw;
        const int n = n_here * ncw;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int sl = i / ncw;
            __stcs(dst + i, app[(size_t)sl * a.slot_stride + (i - sl * ncw)]);
        }
    }
}

// One codeword per CTA (FULL kernels): the per-codeword load / store run OUT of line.  They execute once per codeword, and
// inlined into the kernel they weigh on the layer loop's register allocation.  Measured on one box, two rounds each
// (profiles/r02_v9_ool_helpers_ab.txt; inline -> out of line, ms per headline launch): float32 2.900 -> 2.890 fixed, 3.086 -> 3.074
// with the stop, config 4 1.136 -> 1.130; packed half 1.716 -> 1.675, 2.038 -> 2.012, 1.002 -> 0.99.  Everything they need is passed
// by value: with a reference to the kernel parameters every field access out of line is a generic load, and config 4 (short
// codewords) lost 1.3 %.  The multi-codeword kernels keep the inline versions: out of line BASELINE config 3 lost 9 %.
__device__ __noinline__ uint32_t load_one_ool(const float *src, float *app, const int ncw, uint64_t *bar, uint32_t parity) {
    const uint32_t bytes = (uint32_t)ncw * 4u;                 // ncw*4 is a multiple of 16 for every (BG, Z)
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy accesses to app vs the async write
        mbar_expect_tx(bar, bytes);
        bulk_g2s_stream(app, src, bytes, bar);                 // evict_first: must not displace the pinned c2v scratch
    }
    mbar_wait(bar, parity);
    float4 *dst = reinterpret_cast<float4 *>(app);
    for (int i = threadIdx.x; i < (ncw >> 2); i += blockDim.x) {
        float4 v = dst[i];
        v.x = clamp_llr(v.x); v.y = clamp_llr(v.y); v.z = clamp_llr(v.z); v.w = clamp_llr(v.w);
        dst[i] = v;
    }
    return parity ^ 1u;
}
__device__ __noinline__ void store_one_ool(const float *app, uint8_t *hard, float *soft, const int ncw, const int K) {
    if ((K & 3) == 0) {       // four hard decisions per 32-bit store (K = kcols*Z is a multiple of 4 for every even Z)
        const float4 *src = reinterpret_cast<const float4 *>(app);
        uint32_t *dst = reinterpret_cast<uint32_t *>(hard);
        for (int k = threadIdx.x; k < (K >> 2); k += blockDim.x) {
            const float4 v = src[k];
            dst[k] = (__float_as_uint(v.x) >> 31) | ((__float_as_uint(v.y) >> 31) << 8) |
                     ((__float_as_uint(v.z) >> 31) << 16) | ((__float_as_uint(v.w) >> 31) << 24);
        }
    } else {
        for (int k = threadIdx.x; k < K; k += blockDim.x) hard[k] = app[k] < 0.0f ? 1 : 0;
    }
    if (soft) {
        float4 *dst = reinterpret_cast<float4 *>(soft);
        const float4 *src = reinterpret_cast<const float4 *>(app);
        for (int i = threadIdx.x; i < (ncw >> 2); i += blockDim.x) __stcs(dst + i, src[i]);
    }
}

// The same syndrome with the base graph's shape known at compile time: edge operands come from the parameter bank
// as in the layer code (no descriptor loads, no loop control) -- about 5 instructions per edge instead of 12.  With
// 'Parity check satisfied' (the reference's only setting, NRLDPCDecoder.m:120) this runs after EVERY iteration.
template <int BG, int R, int REND, bool FULL>
struct SyndromeRows {
    // ext_mask: hard decisions of the degree-1 parity variables (ext_hard_mask), bit r for row r
    static __device__ __forceinline__ uint32_t run(const DecArgs &a, const Lane &l, const unsigned long long ext_mask, uint32_t fail) {
        if (R >= 4 && R >= a.n_rows) return fail;
        constexpr int DEG = BgShape<BG>::deg(R);
        constexpr int NE = R >= 4 ? DEG - 1 : DEG;
        constexpr int E0 = BgShape<BG>::start(R);
        uint32_t par = R >= 4 ? (uint32_t)((ext_mask >> R) & 1ull) << 31 : 0u;
#pragma unroll
        for (int e = 0; e < NE; ++e) par ^= lds_u32(edge_addr<FULL>(l, a.ed[E0 + e]));
        fail |= par;
        asm volatile("" : "+r"(fail));   // one row's loads are consumed before the next row's are issued (register pressure)
        return SyndromeRows<BG, R + 1, REND, FULL>::run(a, l, ext_mask, fail);
    }
};
template <int BG, int REND, bool FULL>
struct SyndromeRows<BG, REND, REND, FULL> {
    static __device__ __forceinline__ uint32_t run(const DecArgs &, const Lane &, const unsigned long long, uint32_t fail) { return fail; }
};

// Out of line on purpose: inlined into the decode kernel the 316 unrolled loads changed the register allocation of
// the layer code (552 bytes of spills, 3.10 -> 3.97 ms on the fixed-iteration headline that never runs it).
// Two stages: the four core rows (degree 19 / 8..10, they cover every information and core-parity column), then the
// extension rows.  A codeword that has not converged almost surely fails a core check, so with enough active rows the
// kernel runs the extension stage only for codewords whose core checks hold (DecArgs::staged_min_rows): the exact
// result of the full syndrome at a quarter of the loads in every iteration but a codeword's last.
template <int BG, bool FULL>
__device__ __noinline__ uint32_t syndrome_unrolled_core(const DecArgs &a, const Lane l) {
    return SyndromeRows<BG, 0, 4, FULL>::run(a, l, 0ull, 0u);
}
template <int BG, bool FULL>
__device__ __noinline__ uint32_t syndrome_unrolled_ext(const DecArgs &a, const Lane l, const uint32_t *my_rec, const uint64_t pol) {
    const unsigned long long mask = ext_hard_mask(my_rec, pol, a.n_rows, a.smem_base + l.slot_off + l.zoff + (uint32_t)((a.kcols + 4) * a.Z) * 4u, (uint32_t)a.Z * 4u);
    return SyndromeRows<BG, 4, BgShape<BG>::kRows, FULL>::run(a, l, mask, 0u);
}

// TRACK kernels: the parity variables' hard decisions are already in registers (DecCtx::ext_lo / ext_hi)
template <int BG, bool FULL>
__device__ __noinline__ uint32_t syndrome_unrolled_ext_bits(const DecArgs &a, const Lane l, const unsigned long long ext_mask) {
    return SyndromeRows<BG, 4, BgShape<BG>::kRows, FULL>::run(a, l, ext_mask, 0u);
}
struct ExtFromBit {   // last_row_parity with the parity variable's hard decision given (bit 31 of `word`)
    uint32_t word;
    __device__ __forceinline__ uint32_t operator()(const uint4, uint32_t) const { return word; }
};

// Bit-sliced syndrome for the FULL kernels (Z a multiple of 32, one codeword per CTA).  After an iteration every warp
// packs the hard decisions of its 32 variables of a block column into one word (ballot), giving hb[col][Z/32];
// the 32 checks z0..z0+31 of base row r then see, for an edge (col, shift), the 32 consecutive bits starting at
// (z0 + shift) mod Z of column col: two words and a funnel shift.  Warp w owns the checks 32w..32w+31 of EVERY row.
// Two stages (with 'Parity check satisfied', NRLDPCDecoder.m:120, this runs after every iteration, and all but a
// codeword's last run fail):
//   1. pack the kcols + 4 core columns; the four core rows, ONE EDGE PER LANE (degree <= 19) and a warp XOR reduction
//      (REDUX) per row; CTA-wide OR.  A codeword that has not converged almost surely stops here: ~200 instructions
//      per thread instead of the ~1300 of packing all columns and walking all rows;
//   2. only if every core check holds: pack the extension columns, walk the extension rows (lane l: rows 4+l, 36+l).
// The edge descriptors of this routine are lane-indexed, so they are read from a shared-memory copy (`sed`, filled
// once per CTA: shift | col*W << 16) -- lane-divergent reads of the kernel-parameter constant bank serialise.
// With few active rows the unrolled per-thread syndrome was cheaper than the unstaged bit-sliced one (BG1 Z=384, 5 rows:
// 25.4 against 21.6 Gb/s): DecArgs::bitsliced_min_rows.
// All pointers are shared-window byte addresses (explicit LDS / STS: a generic pointer costs an address-space
// resolution per access in an out-of-line routine).
// ext_row0 >= 0: column col0 + i is the degree-1 parity column of extension row ext_row0 + i -- shared memory holds its
// channel value, the hard decision is bit (ext_row0 + i) of ext_mask (ext_hard_mask; thread z owns check z of that row:
// identity circulant)
__device__ __forceinline__ void pack_hard_bits(uint32_t app_s, uint32_t hb_s, int Z, int col0, int col1, int z, int ext_row0 = -1,
                                               const unsigned long long ext_mask = 0ull) {
    uint32_t src = app_s + (uint32_t)(col0 * Z + z) * 4u;
    uint32_t dst = hb_s + (uint32_t)(col0 * (Z >> 5) + (z >> 5)) * 4u;
#pragma unroll 1   // code size: see syndrome_bitsliced
    for (int col = col0; col < col1; ++col, src += (uint32_t)Z * 4u, dst += (uint32_t)(Z >> 5) * 4u) {
        const uint32_t bit = ext_row0 >= 0 ? (uint32_t)((ext_mask >> (ext_row0 + (col - col0))) & 1ull) : lds_u32(src) >> 31;
        const uint32_t word = __ballot_sync(0xffffffffu, bit);
        if ((z & 31) == 0) sts_u32(dst, word);
    }
}
__device__ __forceinline__ void fill_syndrome_edges(const DecArgs &a, uint32_t sed_s) {
    for (int e = threadIdx.x; e < a.n_edges; e += blockDim.x) {
        const uint2 d = a.ed[e];
        // shift | byte offset of the column's words (col * Z * 4 / 32 = col * W * 4) << 16
        sts_u32(sed_s + (uint32_t)e * 4u, (d.x >> 2) | (((d.y - a.smem_base) >> 5) << 16));
    }
}
// the 32 hard decisions checks z0..z0+31 see through edge `desc`
__device__ __forceinline__ uint32_t hb_window(uint32_t hb_s, uint32_t desc, int z0, int Z, int W) {
    int p = z0 + (int)(desc & 0xffffu);
    if (p >= Z) p -= Z;
    const uint32_t col_s = hb_s + (desc >> 16);
    const int i0 = p >> 5, i1 = i0 + 1 == W ? 0 : i0 + 1;
    return __funnelshift_r(lds_u32(col_s + (uint32_t)i0 * 4u), lds_u32(col_s + (uint32_t)i1 * 4u), p & 31);
}
// first edges of the four core rows and their end, one byte each (BG1: 0 19 38 57 76; BG2: 0 8 18 26 36)
template <int BG>
__host__ __device__ constexpr unsigned long long core_row_starts() {
    unsigned long long v = 0;
    for (int r = 0; r <= 4; ++r) v |= (unsigned long long)BgShape<BG>::start(r) << (8 * r);
    return v;
}
// CTA-uniform result (non-zero: some active check fails).  Contains barriers: every thread of the CTA calls it.
// app_s: the codeword's APP array; hb_s: packed hard decisions [cols][W], the lane-indexed edge table lies 2*cols*W
// words behind it.  Scalars by value and the core rows' shape from BgShape: a reference to the kernel parameters would
// be a generic pointer in this out-of-line routine (LD.E per field); only the rare extension stage reads row_start.
template <int BG>
__device__ __noinline__ int syndrome_bitsliced(uint32_t app_s, uint32_t hb_s, int Z, int n_rows, int z, const unsigned short *row_start,
                                               const uint32_t *my_rec, const uint64_t pol) {
    using S = BgShape<BG>;
    const int W = Z >> 5, z0 = z & ~31, lane = z & 31;
    constexpr int kCore = S::kKcols + 4;
    const uint32_t sed_s = hb_s + (uint32_t)(S::kCols * W) * 8u;
    pack_hard_bits(app_s, hb_s, Z, 0, kCore, z);
    __syncthreads();
    uint32_t fail = 0;
    constexpr unsigned long long kStarts = core_row_starts<BG>();
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
        const int e = (int)((kStarts >> (8 * r)) & 0xffu) + lane;
        uint32_t v = 0;
        if (e < (int)((kStarts >> (8 * r + 8)) & 0xffu)) v = hb_window(hb_s, lds_u32(sed_s + (uint32_t)e * 4u), z0, Z, W);
        fail |= __reduce_xor_sync(0xffffffffu, v);
    }
    if (__syncthreads_or(fail != 0u)) return 1;
    if (n_rows <= 4) return 0;
    pack_hard_bits(app_s, hb_s, Z, kCore, min(S::kCols, S::kKcols + n_rows), z, 4,
                   ext_hard_mask(my_rec, pol, n_rows, app_s + (uint32_t)(kCore * Z + z) * 4u, (uint32_t)Z * 4u));
    __syncthreads();
    for (int r = 4 + lane; r < n_rows; r += 32) {
        uint32_t acc = 0;
        for (int e = row_start[r]; e < row_start[r + 1]; ++e) acc ^= hb_window(hb_s, lds_u32(sed_s + (uint32_t)e * 4u), z0, Z, W);
        fail |= acc;
    }
    return __syncthreads_or(fail != 0u);
}

// ---- one full iteration over the layers: looped (generic) --------------------------------------
__device__ __forceinline__ void iteration_looped(const DecArgs &a, DecCtx &c, const int it, const bool keep_last) {
    const bool store_rec = it + 1 < a.max_iters || keep_last;
    for (int r = 0; r < a.n_rows; ++r) {
        // software prefetch of the next layer's record
        const bool ld = (r + 1 == a.n_rows) ? store_rec : (it > 0);
        uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
        if (ld && !c.done) nxt = ld_rec(c.my_rec, r + 1, c.pol);
        if (!c.done) {
            const int e0 = a.row_start[r];
            const int deg = a.row_start[r + 1] - e0;
            const uint2 *ed = a.ed + e0;
            uint4 rec = make_uint4(0u, 0u, 0u, 0u);
            // extension rows (r >= 4) end in their degree-1 identity column (checked at nrldpc_create)
            switch (r >= 4 ? deg + 32 : deg) {
#define NRLDPC_ROW_CASE(D) case D: rec = process_row<D, false, false>(c.l, ed, c.cur.x, c.cur.y, c.cur.z, a.alpha); break; \
                           case D + 32: rec = process_row<D, true, false>(c.l, ed, c.cur.x, c.cur.y, c.cur.z, a.alpha); break;
                NRLDPC_ROW_CASE(3) NRLDPC_ROW_CASE(4) NRLDPC_ROW_CASE(5) NRLDPC_ROW_CASE(6)
                NRLDPC_ROW_CASE(7) NRLDPC_ROW_CASE(8) NRLDPC_ROW_CASE(9) NRLDPC_ROW_CASE(10)
#undef NRLDPC_ROW_CASE
                case 19: rec = process_row<19, false, false>(c.l, ed, c.cur.x, c.cur.y, c.cur.z, a.alpha); break;
                default: break;
            }
            if (store_rec) st_rec(c.my_rec, r == 0 ? a.n_rows : r, rec, c.pol);
        }
        c.cur = nxt;
        __syncthreads();
    }
}

// ---- one full iteration over the layers: fully unrolled for base graph BG -----------------------
// One instantiation serves every iteration: the unrolled layer code of BG1 is ~100 KB of SASS and
// has to stay resident in the SM's instruction cache (per-iteration specialisations were measured
// to thrash it).  ld_from / ld_to: the next layer's record is prefetched while processing a layer whose
// last row is R iff ld_from <= R < ld_to (first iteration: only across the iteration boundary; last: never
// across).  (Prefetching unconditionally over zeroed records was measured 0.5-1 % slower in this kernel and
// 7 % slower in the packed-half kernel.)
// FULL: every thread of the CTA runs the row code for the whole decode (one codeword per CTA): no per-thread activity
// test.  MASKED (FULL only): Z is not a multiple of 32.  The CTA is launched with exactly Z threads (its last warp is partly
// filled -- barriers and the barrier reductions count warps, so nothing else changes); the flavour only rules out the
// warp-ballot (bit-sliced) syndrome.  The CTA-uniform code thus serves every one-codeword CTA.
// MODE: 0 plain, 1 MASKED (FULL only), 2 TRACK (multi-codeword kernels under the stop: see row_scatter_par)
template <int BG, int R, bool FULL, int MODE = 0>
struct UnrolledRows {
    static constexpr bool MASKED = MODE == 1, TRACK = MODE == 2;
    static constexpr int kPos0 = (TRACK && R >= 4) ? ((R - 4) & 31) : -1;            // bit of row R in ext_lo / ext_hi
    static constexpr int kPos1 = (TRACK && R + 1 >= 4) ? ((R + 1 - 4) & 31) : -1;    // bit of row R + 1 (row pairs)
    static __device__ __forceinline__ void run(const DecArgs &a, DecCtx &c, const int ld_from, const int ld_to, const bool store_rec,
                                               const uint4 prev = make_uint4(0u, 0u, 0u, 0u)) {
        if (R >= 4 && R >= a.n_rows) {         // n_rows >= 4 is validated by the host
            if (FULL) c.last_rec = prev;       // the previous row was the last active one (trimmed row count)
            return;
        }
        constexpr int DEG = BgShape<BG>::deg(R);
        constexpr int E0 = BgShape<BG>::start(R);
        constexpr bool PAIR = pair_first<BG>(R);
        if (PAIR && R + 1 < a.n_rows) {
            // rows R and R+1 touch disjoint block columns: one layer, one barrier
            constexpr int DEG2 = BgShape<BG>::deg(PAIR ? R + 1 : R);
            constexpr int E1 = BgShape<BG>::start(PAIR ? R + 1 : R);
            constexpr bool kLastPair = FULL && PAIR && R + 2 == BgShape<BG>::kRows;
            uint32_t par = 0;
            uint4 rec_last = make_uint4(0u, 0u, 0u, 0u);
            if (FULL || !c.done) {
                uint4 nxt = make_uint4(0u, 0u, 0u, 0u), nxt2 = nxt;
                if ((R + 1 >= ld_from && R + 1 < ld_to)) {
                    nxt = ld_rec(c.my_rec, R + 2, c.pol);   // slot R+2 (slot n_rows holds layer 0)
                    if (pair_first<BG>(R + 2)) nxt2 = ld_rec(c.my_rec, R + 3, c.pol);
                }
                RowState<DEG> s0;
                RowState<DEG2> s1;
                row_gather<DEG, (R >= 4), FULL>(c.l, a.ed + E0, c.cur.x, c.cur.y, c.cur.z, s0);
                row_gather<DEG2, (R >= 4), FULL>(c.l, a.ed + E1, c.cur2.x, c.cur2.y, c.cur2.z, s1);
                const uint4 rec0 = row_scatter_par<DEG, (R >= 4), kLastPair, kPos0>(s0, a.alpha, par, true, (R - 4) < 32 ? &c.ext_lo : &c.ext_hi);
                const uint4 rec1 = row_scatter_par<DEG2, (R >= 4), kLastPair, kPos1>(s1, a.alpha, par, true, (R + 1 - 4) < 32 ? &c.ext_lo : &c.ext_hi);
                if (store_rec) {
                    st_rec(c.my_rec, R == 0 ? a.n_rows : R, rec0, c.pol);
                    st_rec(c.my_rec, R + 1, rec1, c.pol);
                }
                c.cur = nxt;
                c.cur2 = nxt2;
                if (FULL && R >= 4) rec_last = rec1;   // only an extension row's record is ever needed (rows 0..3 have no degree-1 variable)
            }
            // The hard decisions written by the LAST layer of an iteration are final, so an unsatisfied check there
            // proves that the codeword has not converged: with every base row active the layer's barrier doubles as the
            // CTA-wide OR of those parities and the kernel skips the syndrome after most iterations (exact either way).
            if (kLastPair) c.last_fail = __syncthreads_or((int)(par >> 31));
            else __syncthreads();
            UnrolledRows<BG, (PAIR ? R + 2 : BgShape<BG>::kRows), FULL, MODE>::run(a, c, ld_from, ld_to, store_rec, rec_last);
        } else {
            uint4 rec_last = make_uint4(0u, 0u, 0u, 0u);
            if (FULL || !c.done) {
                uint4 nxt = make_uint4(0u, 0u, 0u, 0u), nxt2 = nxt;
                if ((R >= ld_from && R < ld_to)) {
                    nxt = ld_rec(c.my_rec, R + 1, c.pol);   // slot R+1; slot n_rows holds layer 0
                    if (!PAIR && pair_first<BG>(R + 1)) nxt2 = ld_rec(c.my_rec, R + 2, c.pol);
                }
                const uint4 rec = process_row<DEG, (R >= 4), FULL, kPos0>(c.l, a.ed + E0, c.cur.x, c.cur.y, c.cur.z, a.alpha, true,
                                                                             (R - 4) < 32 ? &c.ext_lo : &c.ext_hi);
                if (store_rec) st_rec(c.my_rec, R == 0 ? a.n_rows : R, rec, c.pol);
                c.cur = nxt;
                c.cur2 = nxt2;
                if (FULL && R >= 4) rec_last = rec;
            }
            __syncthreads();
            // a pair opener running alone means R + 1 == n_rows: the iteration ends here
            if (!PAIR) UnrolledRows<BG, R + 1, FULL, MODE>::run(a, c, ld_from, ld_to, store_rec, rec_last);
            else if (FULL) c.last_rec = rec_last;
        }
    }
};
template <int BG, bool FULL, int MODE>
struct UnrolledRows<BG, BgShape<BG>::kRows, FULL, MODE> {
    static __device__ __forceinline__ void run(const DecArgs &, DecCtx &, int, int, bool, const uint4 = make_uint4(0u, 0u, 0u, 0u)) {}
};

template <int BG, bool FULL, int MODE>
__device__ __forceinline__ void iteration_unrolled(const DecArgs &a, DecCtx &c, const int it, const bool keep_last) {
    const bool first = it == 0, last = it + 1 == a.max_iters;
    c.last_fail = 0;   // set by the last layer when every base row is active
    // keep_last: the records of the final iteration are written too (they hold the messages to the degree-1 parity
    // variables, from which the syndrome and the soft output rebuild those variables' a-posteriori values)
    UnrolledRows<BG, 0, FULL, MODE>::run(a, c, first ? a.n_rows - 1 : 0, last ? a.n_rows - 1 : a.n_rows, !last || keep_last);
}

// BG = 0: generic looped variant; BG = 1 / 2: layer loop unrolled for that base graph.
template <int BG, bool FULL, int MODE = 0>
__global__ void __launch_bounds__(kDecThreads, kDecCtasPerSm) decode_nms_kernel(const __grid_constant__ DecArgs a) {
    constexpr bool MASKED = MODE == 1, TRACK = MODE == 2;
    static_assert(FULL || !MASKED, "MASKED is a flavour of the one-codeword (FULL) kernels");
    static_assert(!FULL || !TRACK, "TRACK is a flavour of the multi-codeword kernels");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Z = a.Z;
    const int ncw = a.ncols * Z;
    const int K = a.kcols * Z;
    float *app = reinterpret_cast<float *>(smem_raw);
    // behind the APP arrays: syndrome flags of the two stages [2][cwpc] (the packed-half kernel has [2][2*cwpc]), the
    // work-group slot, the TMA mbarrier, and for the FULL kernels the packed hard decisions [ncols][Z/32] (two planes
    // in the packed-half kernel) and the lane-indexed edge table of the bit-sliced syndrome (decode_smem_bytes)
    int *s_flag = reinterpret_cast<int *>(app + (size_t)a.cwpc * a.slot_stride);
    int &s_group = s_flag[4 * a.cwpc];
    uint64_t *bar = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(s_flag + 4 * a.cwpc + 1) + 7) & ~(uintptr_t)7);
    uint32_t bar_parity = 0;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the edge descriptors carry absolute shared addresses (DecArgs::ed): fail loudly if the window moved
    if ((uint32_t)__cvta_generic_to_shared(smem_raw) != a.smem_base) __trap();
    const bool bitsliced = FULL && !MASKED && a.n_rows >= a.bitsliced_min_rows;
    if (bitsliced && (a.early_term || a.ok != nullptr))   // ordered by the barriers below
        fill_syndrome_edges(a, (uint32_t)__cvta_generic_to_shared(bar + 1) + (uint32_t)(a.ncols * (Z >> 5)) * 8u);

    const int tid = threadIdx.x;
    const int slot = tid / Z;
    const int z = tid - slot * Z;
    const bool lane_ok = tid < a.cwpc * Z;
    const long long n_groups = (a.batch + a.cwpc - 1) / a.cwpc;
    const bool want_ok = a.ok != nullptr;
    // the final iteration's records are needed whenever somebody reads the degree-1 parity variables afterwards
    const bool keep_last = a.early_term || want_ok || a.soft != nullptr;

    DecCtx c;
    c.l.zoff = (uint32_t)z * 4u;
    c.l.nZ4 = 0u - (uint32_t)Z * 4u;
    c.l.slot_off = FULL ? 0u : (uint32_t)(slot * a.slot_stride) * 4u;
    c.l.one = (uint32_t)a.one;
    c.my_rec = a.c2v + (size_t)(blockIdx.x / a.rec_group) * (kRecWords * kRecStride) + (blockIdx.x % a.rec_group) * blockDim.x + tid;
    c.pol = make_l2_policy(a.l2_pin);

    while (true) {
        __syncthreads();  // previous group's outputs are out of smem
        if (tid == 0) s_group = (int)((unsigned int)atomicAdd(a.work_counter, 1) - a.work_base);
        __syncthreads();
        const long long group = s_group;
        if (group >= n_groups) break;
        const long long cw0 = group * a.cwpc;
        const int n_here = (int)min((long long)a.cwpc, a.batch - cw0);

        if (FULL) bar_parity = load_one_ool(a.llr + cw0 * ncw, app, ncw, bar, bar_parity);
        else load_group(a, app, cw0, n_here, ncw, bar, bar_parity);
        if (tid < 2 * a.cwpc) s_flag[tid] = 0;
        __syncthreads();

        const bool active = lane_ok && slot < n_here;
        c.done = !active;
        c.cur = make_uint4(0u, 0u, 0u, 0u);
        c.cur2 = c.cur;
        c.ext_lo = c.ext_hi = 0u;
        int my_iters = 0;
        int my_ok = 0;

        for (int it = 0; it < a.max_iters; ++it) {
            if (BG == 0) iteration_looped(a, c, it, keep_last);
            else iteration_unrolled<(BG == 0 ? 1 : BG), FULL, MODE>(a, c, it, keep_last);

            if (!c.done) my_iters = it + 1;
            const bool last = it + 1 == a.max_iters;
            if (a.early_term || (want_ok && last)) {
                if (FULL && !MASKED && bitsliced) {
                    // one codeword per CTA: the verdict is CTA-uniform, no flags
                    // trimmed row count: the last active row is only known at run time, so its Z checks (final as well) are
                    // re-read from shared memory -- a few loads per thread and one reducing barrier
                    if (a.n_rows < BgShape<(BG == 0 ? 1 : BG)>::kRows)
                        c.last_fail = __syncthreads_or((int)(last_row_parity(a, c.l, c.last_rec, ExtAppF32()) >> 31));
                    my_ok = c.last_fail ? 0   // an unsatisfied check in the last layer: not converged, no syndrome needed
                          : (syndrome_bitsliced<(BG == 0 ? 1 : BG)>(a.smem_base, (uint32_t)__cvta_generic_to_shared(bar + 1), Z, a.n_rows, tid, a.row_start,
                                                                     c.my_rec, c.pol) ? 0 : 1);
                    if (a.early_term && my_ok) break;
                } else {
                    constexpr int B = BG == 0 ? 1 : BG;
                    int *s_flag2 = s_flag + a.cwpc;
                    const bool staged = a.n_rows >= a.staged_min_rows;
                    // bit r = hard decision of extension row r's parity variable (TRACK: kept up to date by the row updates)
                    const unsigned long long ext_bits = TRACK ? ((((unsigned long long)c.ext_hi << 32) | c.ext_lo) << 4) : 0ull;
                    bool filtered = false;
                    if (TRACK) {
                        // per-slot last-layer filter: the decisions written by the last active row are final, so an unsatisfied
                        // check there proves the codeword has not converged and the syndrome is skipped for this slot.  The
                        // verdict goes to the SECOND flag (a failure like any other) and is read back before anybody writes
                        // that flag again (behind the next barrier): no thread reads a flag that a neighbour may be writing.
                        if (!c.done) {
                            const uint32_t pbit = (uint32_t)((ext_bits >> (a.n_rows - 1)) & 1ull) << 31;
                            if (last_row_parity(a, c.l, make_uint4(0u, 0u, 0u, 0u), ExtFromBit{pbit}) >> 31) s_flag2[slot] = 1;
                        }
                        __syncthreads();
                        filtered = s_flag2[slot] != 0;
                    }
                    if (!c.done && !filtered) {
                        uint32_t f = BG == 0 ? syndrome_fail(a, c, 0, 4) : syndrome_unrolled_core<B, FULL>(a, c.l);
                        if (!staged) f |= BG == 0 ? syndrome_fail(a, c, 4, a.n_rows)
                                        : TRACK ? syndrome_unrolled_ext_bits<B, FULL>(a, c.l, ext_bits) : syndrome_unrolled_ext<B, FULL>(a, c.l, c.my_rec, c.pol);
                        if (f >> 31) s_flag[slot] = 1;
                    }
                    if (staged) {
                        // extension rows only for codewords whose core checks all hold
                        __syncthreads();
                        if (!c.done && !filtered && !s_flag[slot]) {
                            const uint32_t f = BG == 0 ? syndrome_fail(a, c, 4, a.n_rows)
                                             : TRACK ? syndrome_unrolled_ext_bits<B, FULL>(a, c.l, ext_bits) : syndrome_unrolled_ext<B, FULL>(a, c.l, c.my_rec, c.pol);
                            if (f >> 31) s_flag2[slot] = 1;
                        }
                    }
                    __syncthreads();
                    if (!c.done) {
                        my_ok = (s_flag[slot] | s_flag2[slot]) ? 0 : 1;
                        if (my_ok && a.early_term) c.done = true;
                    }
                    const int all_done = __syncthreads_and(c.done ? 1 : 0);  // also orders the flag reset below
                    if (tid < 2 * a.cwpc) s_flag[tid] = 0;
                    if (a.early_term && all_done) break;
                }
            }
        }
        __syncthreads();

        if (a.soft != nullptr) {
            // soft output: the degree-1 parity variables of the active extension rows get their a-posteriori value
            // (channel value + the row's last message) written into shared memory now that decoding is over
            if (active) {
                const uint32_t base = a.smem_base + c.l.slot_off + c.l.zoff + (uint32_t)((a.kcols + 4) * Z) * 4u;
#pragma unroll 4   // independent records: several L2 round trips in flight
                for (int r = 4; r < a.n_rows; ++r) {
                    const uint32_t addr = base + (uint32_t)((r - 4) * Z) * 4u;
                    sts_u32(addr, ext_app(c.my_rec, r, c.pol, lds_u32(addr)));
                }
            }
            __syncthreads();
        }
        if (FULL) store_one_ool(app, a.hard + cw0 * K, a.soft ? a.soft + cw0 * ncw : nullptr, ncw, K);
        else store_outputs(a, app, cw0, n_here, ncw, K);
        if (active && z == 0) {
            if (a.iters) a.iters[cw0 + slot] = my_iters;
            if (a.ok) a.ok[cw0 + slot] = (uint8_t)my_ok;
        }
    }
}

// ---- multi-codeword CTAs under the parity-check stop: per-slot refill ---------------------------------------------
// In decode_nms_kernel a CTA's codewords form a group: a codeword that converges early idles until the slowest one of
// its group is done (BASELINE config 3 with the reference's stop, NRLDPCDecoder.m:120: 5.4 mean iterations in the time of
// about 8).  Here every slot (Z threads, one codeword) is its own pipeline:
//     RUN --converged or max_iters--> outputs written, next codeword fetched (device work counter, per codeword),
//         its LLRs requested with per-thread asynchronous copies (cp.async, 4 bytes: any Z, any alignment)
//     LOAD  the slot sits out ONE pass over the layers while the copies land (waiting for them inside the pass that
//           issued them would stall the whole CTA at the next barrier: measured 4.27 -> 6.22 ms in round 1)
//     RUN   at the end of that pass: wait_group, clamp in place (each thread clamps what it copied), iteration 0.
// The passes over the layers stay CTA-wide (one barrier per layer); slots differ only in which iteration they are in, so
// the record prefetch gate (first iteration: nothing to load) is per thread.  Arithmetic and outputs are those of
// decode_nms_kernel bit for bit (tests/test_gpu_parity.py); only the order in which codewords are started differs.
__device__ __forceinline__ void cp_async_4(uint32_t smem_addr, const float *gptr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int BG>
__global__ void __launch_bounds__(kDecThreads, kDecCtasPerSm) decode_nms_refill_kernel(const __grid_constant__ DecArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    enum { ST_RUN = 0, ST_LOAD = 1, ST_IDLE = 2 };
    const int Z = a.Z;
    const int ncw = a.ncols * Z;
    const int K = a.kcols * Z;
    float *app = reinterpret_cast<float *>(smem_raw);
    int *s_flag = reinterpret_cast<int *>(app + (size_t)a.cwpc * a.slot_stride);   // [cwpc] core stage, [cwpc] extension stage,
    int *s_flag2 = s_flag + a.cwpc;
    int *s_cw = s_flag + 2 * a.cwpc;                                                // [cwpc] next codeword of the slot
    if ((uint32_t)__cvta_generic_to_shared(smem_raw) != a.smem_base) __trap();

    const int tid = threadIdx.x;
    const int slot = tid / Z;
    const int z = tid - slot * Z;
    const bool lane_ok = tid < a.cwpc * Z;
    const int batch = (int)a.batch;              // the host uses this kernel for batches below 2^31 only

    DecCtx c;
    c.l.zoff = (uint32_t)z * 4u;
    c.l.nZ4 = 0u - (uint32_t)Z * 4u;
    c.l.slot_off = (uint32_t)(slot * a.slot_stride) * 4u;
    c.l.one = (uint32_t)a.one;
    c.my_rec = a.c2v + (size_t)(blockIdx.x / a.rec_group) * (kRecWords * kRecStride) + (blockIdx.x % a.rec_group) * blockDim.x + tid;
    c.pol = make_l2_policy(a.l2_pin);
    c.last_fail = 0;
    const uint32_t my_app_s = a.smem_base + c.l.slot_off + c.l.zoff;   // element z of block column 0 of this slot's APP array
    const uint32_t col_bytes = (uint32_t)Z * 4u;

    auto request = [&](int cw) {      // this thread's share of codeword cw: position z of every block column
        const float *src = a.llr + (long long)cw * ncw + z;
        uint32_t dst = my_app_s;
        for (int col = 0; col < a.ncols; ++col, src += Z, dst += col_bytes) cp_async_4(dst, src);
        cp_async_commit();
    };
    auto land = [&]() {               // copies done -> clamp in place (+-LLR_MAX, NaN filler, -0)
        cp_async_wait_all();
        uint32_t p = my_app_s;
        for (int col = 0; col < a.ncols; ++col, p += col_bytes) sts_f32(p, clamp_llr(lds_f32(p)));
    };

    int st = ST_IDLE, my_cw = -1, my_it = 0;
    if (lane_ok && z == 0) {
        const int n = (int)((unsigned int)atomicAdd(a.work_counter, 1) - a.work_base);
        s_cw[slot] = n < batch ? n : -1;
    }
    if (tid < 2 * a.cwpc) s_flag[tid] = 0;
    __syncthreads();
    if (lane_ok) {
        my_cw = s_cw[slot];
        if (my_cw >= 0) { request(my_cw); land(); st = ST_RUN; }   // the first codeword of a slot is waited for at once
    }
    c.cur = make_uint4(0u, 0u, 0u, 0u);
    c.cur2 = c.cur;
    c.ext_lo = c.ext_hi = 0u;
    if (__syncthreads_and(st == ST_IDLE)) return;

    const bool staged = a.n_rows >= a.staged_min_rows;
    while (true) {
        // one pass over the layers: iteration my_it of every running slot
        c.done = st != ST_RUN;
        UnrolledRows<BG, 0, false, 2>::run(a, c, my_it == 0 ? a.n_rows - 1 : 0, a.n_rows, true);

        const bool was_run = st == ST_RUN, was_load = st == ST_LOAD;
        const unsigned long long ext_bits = (((unsigned long long)c.ext_hi << 32) | c.ext_lo) << 4;   // see decode_nms_kernel (TRACK)
        if (was_run) {   // last-layer filter: verdict in the second flag, read back before that flag is written again (see decode_nms_kernel)
            const uint32_t pbit = (uint32_t)((ext_bits >> (a.n_rows - 1)) & 1ull) << 31;
            if (last_row_parity(a, c.l, make_uint4(0u, 0u, 0u, 0u), ExtFromBit{pbit}) >> 31) s_flag2[slot] = 1;
        }
        __syncthreads();
        const bool filtered = lane_ok && s_flag2[slot] != 0;
        if (was_run && !filtered) {
            uint32_t f = syndrome_unrolled_core<BG, false>(a, c.l);
            if (!staged) f |= syndrome_unrolled_ext_bits<BG, false>(a, c.l, ext_bits);
            if (f >> 31) s_flag[slot] = 1;
        }
        if (staged) {
            __syncthreads();
            if (was_run && !filtered && !s_flag[slot]) {
                const uint32_t f = syndrome_unrolled_ext_bits<BG, false>(a, c.l, ext_bits);
                if (f >> 31) s_flag2[slot] = 1;
            }
        }
        __syncthreads();
        bool fin = false;
        if (was_run) {
            ++my_it;
            const int ok = (s_flag[slot] | s_flag2[slot]) ? 0 : 1;
            fin = ok || my_it == a.max_iters;
            if (fin) {
                // outputs of this slot's codeword, written by its own Z threads
                const uint32_t base_s = a.smem_base + c.l.slot_off;
                uint8_t *hard = a.hard + (long long)my_cw * K;
                if ((K & 3) == 0 && (a.slot_stride & 3) == 0) {
                    for (int k4 = z; k4 < (K >> 2); k4 += Z) {
                        const uint32_t p = base_s + (uint32_t)k4 * 16u;
                        reinterpret_cast<uint32_t *>(hard)[k4] = (lds_u32(p) >> 31) | ((lds_u32(p + 4) >> 31) << 8) |
                                                                 ((lds_u32(p + 8) >> 31) << 16) | ((lds_u32(p + 12) >> 31) << 24);
                    }
                } else {
                    for (int k = z; k < K; k += Z) hard[k] = (uint8_t)(lds_u32(base_s + (uint32_t)k * 4u) >> 31);
                }
                if (a.soft != nullptr) {   // thread z owns position z of every column, and the records of its checks (ext_app)
                    float *soft = a.soft + (long long)my_cw * ncw + z;
                    uint32_t p = my_app_s;
                    for (int col = 0; col < a.ncols; ++col, p += col_bytes, soft += Z) {
                        uint32_t w = lds_u32(p);
                        if (col >= a.kcols + 4 && col < a.kcols + a.n_rows) w = ext_app(c.my_rec, col - a.kcols, c.pol, w);
                        __stcs(soft, __uint_as_float(w));
                    }
                }
                if (z == 0) {
                    if (a.iters) a.iters[my_cw] = my_it;
                    if (a.ok) a.ok[my_cw] = (uint8_t)ok;
                    const int n = (int)((unsigned int)atomicAdd(a.work_counter, 1) - a.work_base);
                    s_cw[slot] = n < batch ? n : -1;
                }
            }
        }
        __syncthreads();   // flags read, next codeword indices published, finished slots' APP arrays read out
        if (tid < 2 * a.cwpc) s_flag[tid] = 0;
        if (was_load) {    // requested during the previous pass: join with iteration 0
            land();
            st = ST_RUN; my_it = 0;
            c.cur = make_uint4(0u, 0u, 0u, 0u);
            c.cur2 = c.cur;
            c.ext_lo = c.ext_hi = 0u;
        }
        if (fin) {
            my_cw = s_cw[slot];
            if (my_cw >= 0) { request(my_cw); st = ST_LOAD; }
            else st = ST_IDLE;
        }
        if (__syncthreads_and(st == ST_IDLE)) break;   // also publishes the clamped values and the flag reset
    }
}

}  // namespace nrldpc
