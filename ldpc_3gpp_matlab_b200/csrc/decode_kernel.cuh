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
//     as one 16-byte record per check in a per-CTA global scratch that stays L2-resident,
//     software-prefetched one layer ahead; the first iteration reads nothing and the last writes
//     nothing;
//   * the base graph's shape (row degrees, edge order) is a compile-time constant per base graph:
//     the layer loop is fully unrolled, and the per-edge shift / column offsets are read straight
//     from the kernel-parameter constant bank as instruction operands (no descriptor loads);
//   * arithmetic is float32 with every add/mul individually rounded (__fsub_rn/__fmul_rn/__fadd_rn:
//     no FMA contraction), signs handled as sign BITS, so results are bit-identical to the CPU
//     oracle (oracle/nrldpc_oracle.c, function orc_decode_nms).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nrldpc {

constexpr int kDecThreads = 384;     // max threads per decode CTA (= largest lifting size)
constexpr int kDecCtasPerSm = 2;
constexpr float kLlrMax = 1048576.0f;
constexpr int kMaxEdges = 316;
constexpr int kMaxRows = 46;

// Edge descriptor, read from the kernel-parameter constant bank:
//   x = shift*4   (bytes): lane z reads circulant position (z + shift) mod Z
//   y = col*Z*4   (bytes): start of the block column inside one codeword's APP array
struct DecArgs {
    const float *llr;        // [batch][ncw]
    uint8_t *hard;           // [batch][K]
    float *soft;             // [batch][ncw] or null
    int32_t *iters;          // [batch] or null
    uint8_t *ok;             // [batch] or null
    long long batch;
    int Z, ncols, kcols, n_rows, n_edges, max_iters, early_term, cwpc;
    float alpha;
    uint4 *c2v;              // [grid][n_rows+1][blockDim] {alpha*min1, alpha*min2, argmin | signbits << 5, -}
    int *work_counter;
    unsigned short row_start[kMaxRows + 2];
    uint2 ed[kMaxEdges];
};

template <int BG> struct BgShape;
template <> struct BgShape<1> {
    static constexpr int kRows = NRLDPC_BG1_ROWS;
    static __host__ __device__ constexpr int deg(int r) { return nrldpc_bg1_deg[r]; }
    static __host__ __device__ constexpr int start(int r) { return nrldpc_bg1_start[r]; }
};
template <> struct BgShape<2> {
    static constexpr int kRows = NRLDPC_BG2_ROWS;
    static __host__ __device__ constexpr int deg(int r) { return nrldpc_bg2_deg[r]; }
    static __host__ __device__ constexpr int start(int r) { return nrldpc_bg2_start[r]; }
};

__device__ __forceinline__ float clamp_llr(float x) {
    // NaN marks filler upstream (NRLDPCDecoder.m:224,264): fminf(NaN, M) = M.
    return fmaxf(fminf(x, kLlrMax), -kLlrMax);
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float x;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
    return x;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// One check row of degree DEG for check z.  `base` = byte address (shared window) of this thread's
// codeword's APP array, zoff = z*4; (om1, om2, ometa) is the compressed message record written
// for this check in the previous iteration (zeros in the first).
// Sign bits of the messages are kept MSB-first: edge e sits at bit (DEG-1-e) of the sign field.
template <int DEG>
__device__ __forceinline__ void process_row(const uint32_t base, const uint2 *__restrict__ ed, const uint32_t zoff,
                                            const uint32_t Z4, const float om1, const float om2,
                                            const uint32_t ometa, const float alpha, float &nm1,
                                            float &nm2, uint32_t &nmeta) {
    float t[DEG];
    uint32_t addr[DEG];
    float m1 = __int_as_float(0x7f800000), m2 = __int_as_float(0x7f800000);
    uint32_t sx = 0, ts = 0;
    int arg = 0;
    const int oarg = (int)(ometa & 31u);
    const uint32_t osg = ometa >> 5;
#pragma unroll
    for (int e = 0; e < DEG; ++e) {
        const uint2 d = ed[e];
        const uint32_t u = zoff + d.x;                   // (z + shift)*4, wraps at Z*4:
        const uint32_t a = base + d.y + min(u, u - Z4);  // u - Z4 underflows to a huge value when u < Z4
        addr[e] = a;
        const float x = lds_f32(a);
        const float mag = (e == oarg) ? om2 : om1;
        const float c = __uint_as_float(__float_as_uint(mag) | ((osg << (31 - (DEG - 1 - e))) & 0x80000000u));
        const float tt = __fsub_rn(x, c);
        t[e] = tt;
        const float ab = fabsf(tt);
        arg = (ab < m1) ? e : arg;
        m2 = fminf(m2, fmaxf(ab, m1));
        m1 = fminf(m1, ab);
        sx ^= __float_as_uint(tt);
        ts = __funnelshift_l(__float_as_uint(tt), ts, 1);  // ts = ts << 1 | signbit(tt)
    }
    const uint32_t sg = sx & 0x80000000u;
    // both candidate magnitudes with the row's sign product folded in
    uint32_t m1ss = __float_as_uint(__fmul_rn(alpha, m1)) | sg;
    uint32_t m2ss = __float_as_uint(__fmul_rn(alpha, m2)) | sg;
    asm volatile("" : "+r"(m1ss), "+r"(m2ss));  // keep the sign folded per row, not re-derived per edge
#pragma unroll
    for (int e = 0; e < DEG; ++e) {
        const uint32_t sel = (e == arg) ? m2ss : m1ss;
        const float c = __uint_as_float(sel ^ (__float_as_uint(t[e]) & 0x80000000u));
        sts_f32(addr[e], __fadd_rn(t[e], c));
    }
    nm1 = __uint_as_float(m1ss & 0x7fffffffu);
    nm2 = __uint_as_float(m2ss & 0x7fffffffu);
    // sign of message e = sg ^ sign(t_e)
    const uint32_t csg = sg ? (ts ^ ((1u << DEG) - 1u)) : ts;
    nmeta = (uint32_t)arg | (csg << 5);
}

// ---- pieces shared by both kernel variants -------------------------------------------------------
struct DecCtx {
    uint32_t base, zoff, Z4;
    uint4 *my_rec;   // slot 0 of this thread's record column
    uint4 *rec0;     // slot n_rows: where layer 0's record lives
    uint32_t rec_stride;
    uint4 cur;
    bool done;
};

__device__ __forceinline__ void load_group(const DecArgs &a, float *app, long long cw0, int n_here, int ncw) {
    // the group's codewords are contiguous in HBM; ncw*4 bytes is a multiple of 16 for every (BG,Z)
    const float4 *src = reinterpret_cast<const float4 *>(a.llr + cw0 * ncw);
    float4 *dst = reinterpret_cast<float4 *>(app);
    const int n4 = (n_here * ncw) >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 v = __ldcs(src + i);
        v.x = clamp_llr(v.x); v.y = clamp_llr(v.y); v.z = clamp_llr(v.z); v.w = clamp_llr(v.w);
        dst[i] = v;
    }
}

// exact syndrome of hard = (app < 0) over the active rows ('Parity check satisfied', NRLDPCDecoder.m:120)
__device__ __forceinline__ int syndrome_fail(const DecArgs &a, const DecCtx &c) {
    int fail = 0;
    for (int r = 0; r < a.n_rows; ++r) {
        int par = 0;
        for (int e = a.row_start[r]; e < a.row_start[r + 1]; ++e) {
            const uint2 d = a.ed[e];
            const uint32_t u = c.zoff + d.x;
            par ^= (lds_f32(c.base + d.y + min(u, u - c.Z4)) < 0.0f) ? 1 : 0;
        }
        fail |= par;
    }
    return fail;
}

__device__ __forceinline__ void store_outputs(const DecArgs &a, const float *app, long long cw0, int n_here, int ncw, int K) {
    for (int s = 0; s < n_here; ++s) {
        const float *src = app + (size_t)s * ncw;
        uint8_t *dst = a.hard + (cw0 + s) * K;
        for (int k = threadIdx.x; k < K; k += blockDim.x) dst[k] = src[k] < 0.0f ? 1 : 0;
    }
    if (a.soft) {
        float4 *dst = reinterpret_cast<float4 *>(a.soft + cw0 * ncw);
        const float4 *src = reinterpret_cast<const float4 *>(app);
        const int n4 = (n_here * ncw) >> 2;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) __stcs(dst + i, src[i]);
    }
}

// ---- one full iteration over the layers: looped (generic) --------------------------------------
__device__ __forceinline__ void iteration_looped(const DecArgs &a, DecCtx &c, const int it) {
    const bool store_rec = it + 1 < a.max_iters;
    // Record slots: layer r >= 1 lives in slot r; layer 0 lives in slot n_rows, so that "the next
    // layer's record" is always the next slot, also across the iteration boundary.
    uint4 *rp = c.my_rec;
    for (int r = 0; r < a.n_rows; ++r) {
        // software prefetch of the next layer's record
        uint4 *np = rp + c.rec_stride;
        const bool ld = (r + 1 == a.n_rows) ? store_rec : (it > 0);
        uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
        if (ld && !c.done) nxt = __ldcg(np);
        if (!c.done) {
            const int e0 = a.row_start[r];
            const int deg = a.row_start[r + 1] - e0;
            const uint2 *ed = a.ed + e0;
            float n1, n2;
            uint32_t nmeta;
            switch (deg) {
#define NRLDPC_ROW_CASE(D) case D: process_row<D>(c.base, ed, c.zoff, c.Z4, __uint_as_float(c.cur.x), __uint_as_float(c.cur.y), c.cur.z, a.alpha, n1, n2, nmeta); break;
                NRLDPC_ROW_CASE(3) NRLDPC_ROW_CASE(4) NRLDPC_ROW_CASE(5) NRLDPC_ROW_CASE(6)
                NRLDPC_ROW_CASE(7) NRLDPC_ROW_CASE(8) NRLDPC_ROW_CASE(9) NRLDPC_ROW_CASE(10)
                NRLDPC_ROW_CASE(19)
#undef NRLDPC_ROW_CASE
                default: n1 = 0.f; n2 = 0.f; nmeta = 0; break;
            }
            if (store_rec) __stcg(r == 0 ? c.rec0 : rp, make_uint4(__float_as_uint(n1), __float_as_uint(n2), nmeta, 0u));
        }
        c.cur = nxt;
        rp = np;
        __syncthreads();
    }
}

// ---- one full iteration over the layers: fully unrolled for base graph BG -----------------------
template <int BG, int R>
struct UnrolledRows {
    static __device__ __forceinline__ void run(const DecArgs &a, DecCtx &c, const bool first, const bool store_rec) {
        if (R >= a.n_rows) return;
        constexpr int DEG = BgShape<BG>::deg(R);
        constexpr int E0 = BgShape<BG>::start(R);
        // slot R+1 (slot n_rows holds layer 0: see iteration_looped)
        const bool ld = (R + 1 == a.n_rows) ? store_rec : !first;
        uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
        if (ld && !c.done) nxt = __ldcg(c.my_rec + (size_t)(R + 1) * c.rec_stride);
        if (!c.done) {
            float n1, n2;
            uint32_t nmeta;
            process_row<DEG>(c.base, a.ed + E0, c.zoff, c.Z4, __uint_as_float(c.cur.x), __uint_as_float(c.cur.y),
                             c.cur.z, a.alpha, n1, n2, nmeta);
            if (store_rec)
                __stcg(c.my_rec + (size_t)(R == 0 ? a.n_rows : R) * c.rec_stride,
                       make_uint4(__float_as_uint(n1), __float_as_uint(n2), nmeta, 0u));
        }
        c.cur = nxt;
        __syncthreads();
        UnrolledRows<BG, R + 1>::run(a, c, first, store_rec);
    }
};
template <int BG>
struct UnrolledRows<BG, BgShape<BG>::kRows> {
    static __device__ __forceinline__ void run(const DecArgs &, DecCtx &, bool, bool) {}
};

// BG = 0: generic looped variant; BG = 1 / 2: layer loop unrolled for that base graph.
template <int BG>
__global__ void __launch_bounds__(kDecThreads, kDecCtasPerSm) decode_nms_kernel(const __grid_constant__ DecArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Z = a.Z;
    const int ncw = a.ncols * Z;
    const int K = a.kcols * Z;
    float *app = reinterpret_cast<float *>(smem_raw);
    int *s_flag = reinterpret_cast<int *>(app + (size_t)a.cwpc * ncw);
    __shared__ int s_group;

    const int tid = threadIdx.x;
    const int slot = tid / Z;
    const int z = tid - slot * Z;
    const bool lane_ok = tid < a.cwpc * Z;
    const long long n_groups = (a.batch + a.cwpc - 1) / a.cwpc;
    const bool want_ok = a.ok != nullptr;

    DecCtx c;
    c.Z4 = (uint32_t)Z * 4u;
    c.zoff = (uint32_t)z * 4u;
    c.base = (uint32_t)__cvta_generic_to_shared(app + (size_t)slot * ncw);
    c.rec_stride = blockDim.x;
    c.my_rec = a.c2v + (size_t)blockIdx.x * (a.n_rows + 1) * c.rec_stride + tid;
    c.rec0 = c.my_rec + (size_t)a.n_rows * c.rec_stride;

    while (true) {
        __syncthreads();  // previous group's outputs are out of smem
        if (tid == 0) s_group = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const long long group = s_group;
        if (group >= n_groups) break;
        const long long cw0 = group * a.cwpc;
        const int n_here = (int)min((long long)a.cwpc, a.batch - cw0);

        load_group(a, app, cw0, n_here, ncw);
        if (tid < a.cwpc) s_flag[tid] = 0;
        __syncthreads();

        const bool active = lane_ok && slot < n_here;
        c.done = !active;
        c.cur = make_uint4(0u, 0u, 0u, 0u);
        int my_iters = 0;
        int my_ok = 0;

        for (int it = 0; it < a.max_iters; ++it) {
            if (BG == 0) iteration_looped(a, c, it);
            else UnrolledRows<(BG == 0 ? 1 : BG), 0>::run(a, c, it == 0, it + 1 < a.max_iters);

            if (!c.done) my_iters = it + 1;
            const bool last = it + 1 == a.max_iters;
            if (a.early_term || (want_ok && last)) {
                if (!c.done && syndrome_fail(a, c)) s_flag[slot] = 1;
                __syncthreads();
                if (!c.done) {
                    my_ok = s_flag[slot] ? 0 : 1;
                    if (my_ok && a.early_term) c.done = true;
                }
                const int all_done = __syncthreads_and(c.done ? 1 : 0);  // also orders the flag reset below
                if (tid < a.cwpc) s_flag[tid] = 0;
                if (a.early_term && all_done) break;
            }
        }
        __syncthreads();

        store_outputs(a, app, cw0, n_here, ncw, K);
        if (active && z == 0) {
            if (a.iters) a.iters[cw0 + slot] = my_iters;
            if (a.ok) a.ok[cw0 + slot] = (uint8_t)my_ok;
        }
    }
}

}  // namespace nrldpc
