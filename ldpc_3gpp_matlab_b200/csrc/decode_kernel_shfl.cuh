// decode_kernel_shfl.cuh -- layered normalized min-sum with the check-node reduction done by WARP SHUFFLES:
// lane = (check, edge) instead of lane = check.
//
// This is the mapping BASELINE.json's north star names ("check-node min-and-sign reduction done with warp shuffles") and
// SURVEY.md section 7 proposes for small lifting sizes.  decode_nms_kernel gives every check of a layer to one thread,
// which scans the row's edges in registers; that fills the machine only when there are enough checks -- (codewords per
// CTA) x Z of them.  Here the W = 4 / 8 / 16 / 32 lanes of a warp segment (W = the row degree rounded up to a power of
// two) hold the edges of ONE check:
//     lane e   loads app[v_e], rebuilds the previous message from the check's record, t_e = app - c_e
//     segment  butterfly of (min1, min2) pairs over log2(W) levels of SHFL.BFLY; sign bits by one ballot, the row's sign
//              product by POPC of the segment's ballot bits; arg-min lane from a second ballot
//     lane e   new message, app[v_e] = t_e + c_e'; lane 0 of the segment writes the check's record
// and a CTA walks the (codeword, z) checks of a layer segment by segment.  Records and the a-posteriori values of the
// degree-1 parity variables live in shared memory (small Z: a codeword's whole state is at most 30 KB).
// Arithmetic is oracle A's (revision 2) operation for operation -- the two smallest magnitudes are order statistics, so the
// butterfly returns the same values as the sequential scan, and every add / multiply is individually rounded -- hence
// the outputs are bit-identical to decode_nms_kernel (tests/test_gpu_parity.py).
// Generic in the base graph (looped over rows, edge table in shared memory).  Selected with NRLDPC_DECODE_VARIANT=shfl for
// Z <= 32; measured against the default mapping in DESIGN.md section 8.
#pragma once
#include "decode_kernel.cuh"

namespace nrldpc {

struct ShflCtx {
    float *app;             // [cwpc][stride]
    uint32_t *rec;          // [n_rows][cmax][3]
    float *ext;             // [n_rows - 4][cmax]   a-posteriori value of the degree-1 parity variable of (row, check)
    const uint32_t *sed;    // [n_edges]  shift | (col * Z) << 16
    const int *done;        // [cwpc]     codeword slot finished (early termination) or empty
    int cmax, C, Z, stride;
    float alpha;
};

// One layer (base row r of degree deg, first edge e0) for the C checks of the CTA, W lanes per check.
template <int W>
__device__ __forceinline__ void shfl_row(const ShflCtx &s, const int r, const int e0, const int deg, const bool first_it) {
    const int lane = threadIdx.x & 31;
    const int e = lane & (W - 1), gbase = lane & ~(W - 1);
    const int G = blockDim.x / W, g = threadIdx.x / W;
    const bool act_e = e < deg, ident = r >= 4 && e == deg - 1;
    const uint32_t desc = act_e ? s.sed[e0 + e] : 0u;
    const int shift = (int)(desc & 0xffffu), colZ = (int)(desc >> 16);
    const uint32_t wmask = W == 32 ? 0xffffffffu : ((1u << W) - 1u);
    const int n_it = (s.C + G - 1) / G;            // the same for every lane: shuffles and ballots stay warp-wide
    for (int i = 0; i < n_it; ++i) {
        const int chk = g + i * G;
        const int slot = chk / s.Z, z = chk - slot * s.Z;
        const bool chk_ok = chk < s.C && !s.done[slot];      // uniform over the segment
        const bool act = act_e && chk_ok;
        int u = z + shift;
        if (u >= s.Z) u -= s.Z;
        const int idx = slot * s.stride + colZ + u;
        const float x = act ? s.app[idx] : 0.f;
        uint32_t *rp = s.rec + ((size_t)r * s.cmax + chk) * 3;
        uint32_t om1 = 0u, om2 = 0u, ometa = 0u;
        if (act && !first_it) { om1 = rp[0]; om2 = rp[1]; ometa = rp[2]; }
        // previous message of this edge: alpha*min2 for the arg-min edge, alpha*min1 otherwise, carrying the row's sign
        // product; its sign bit flipped by the recorded sign of t_e (bit 5 + e)
        const uint32_t mag = ((ometa & 31u) == (uint32_t)e) ? om2 : om1;
        const uint32_t c = mag ^ (((ometa >> (5 + e)) & 1u) << 31);
        const float t = ident ? x : __fsub_rn(x, __uint_as_float(c));   // degree-1 variable: its channel value (oracle A revision 2)
        float m1 = act ? fabsf(t) : __int_as_float(0x7f800000), m2 = __int_as_float(0x7f800000);
#pragma unroll
        for (int off = 1; off < W; off <<= 1) {
            const float o1 = __shfl_xor_sync(0xffffffffu, m1, off), o2 = __shfl_xor_sync(0xffffffffu, m2, off);
            m2 = fminf(fmaxf(m1, o1), fminf(m2, o2));
            m1 = fminf(m1, o1);
        }
        const uint32_t neg = (__ballot_sync(0xffffffffu, act && (__float_as_uint(t) >> 31)) >> gbase) & wmask;
        const bool is_min = act && fabsf(t) == m1;
        const uint32_t mins = (__ballot_sync(0xffffffffu, is_min && !ident) >> gbase) & wmask;
        const uint32_t arg = mins ? (uint32_t)(31 - __clz(mins)) : 31u;      // 31: only the degree-1 edge attains the minimum
        const uint32_t sx = ((uint32_t)__popc(neg) & 1u) << 31;
        const float alpha_s = __uint_as_float(__float_as_uint(s.alpha) | sx);
        const uint32_t m1ss = __float_as_uint(__fmul_rn(alpha_s, m1));
        const uint32_t m2ss = __float_as_uint(__fmul_rn(alpha_s, m2));
        const uint32_t sel = is_min ? m2ss : m1ss;
        const float app_new = __fadd_rn(t, __uint_as_float(sel ^ (__float_as_uint(t) & 0x80000000u)));
        if (act) {
            if (ident) s.ext[(size_t)(r - 4) * s.cmax + chk] = app_new;
            else s.app[idx] = app_new;
            if (e == 0) { rp[0] = m1ss; rp[1] = m2ss; rp[2] = arg | (neg << 5); }
        }
    }
}

// parities of the hard decisions for base row r: sets fail[slot] for every codeword with an unsatisfied check
template <int W>
__device__ __forceinline__ void shfl_syndrome_row(const ShflCtx &s, const int r, const int e0, const int deg, int *fail) {
    const int lane = threadIdx.x & 31;
    const int e = lane & (W - 1), gbase = lane & ~(W - 1);
    const int G = blockDim.x / W, g = threadIdx.x / W;
    const bool act_e = e < deg, ident = r >= 4 && e == deg - 1;
    const uint32_t desc = act_e ? s.sed[e0 + e] : 0u;
    const int shift = (int)(desc & 0xffffu), colZ = (int)(desc >> 16);
    const uint32_t wmask = W == 32 ? 0xffffffffu : ((1u << W) - 1u);
    const int n_it = (s.C + G - 1) / G;
    for (int i = 0; i < n_it; ++i) {
        const int chk = g + i * G;
        const int slot = chk / s.Z, z = chk - slot * s.Z;
        const bool act = act_e && chk < s.C && !s.done[slot];
        int u = z + shift;
        if (u >= s.Z) u -= s.Z;
        uint32_t w = 0u;
        if (act) w = __float_as_uint(ident ? s.ext[(size_t)(r - 4) * s.cmax + chk] : s.app[slot * s.stride + colZ + u]);
        const uint32_t bits = (__ballot_sync(0xffffffffu, act && (w >> 31)) >> gbase) & wmask;
        if (act && e == 0 && (__popc(bits) & 1)) fail[slot] = 1;
    }
}

#define NRLDPC_SHFL_DISPATCH(FN, deg, ...)          \
    do {                                            \
        if ((deg) <= 4) FN<4>(__VA_ARGS__);         \
        else if ((deg) <= 8) FN<8>(__VA_ARGS__);    \
        else if ((deg) <= 16) FN<16>(__VA_ARGS__);  \
        else FN<32>(__VA_ARGS__);                   \
    } while (0)

__global__ void __launch_bounds__(256) decode_nms_shfl_kernel(const __grid_constant__ DecArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Z = a.Z, ncw = a.ncols * Z, K = a.kcols * Z;
    const int cmax = a.cwpc * Z;
    const int rows_all = a.ncols - a.kcols;
    float *app = reinterpret_cast<float *>(smem_raw);
    uint32_t *rec = reinterpret_cast<uint32_t *>(app + (size_t)a.cwpc * ncw);
    float *ext = reinterpret_cast<float *>(rec + (size_t)rows_all * cmax * 3);
    uint32_t *sed = reinterpret_cast<uint32_t *>(ext + (size_t)(rows_all - 4) * cmax);
    int *s_done = reinterpret_cast<int *>(sed + a.n_edges);
    int *s_fail = s_done + a.cwpc;
    const int tid = threadIdx.x;
    for (int e = tid; e < a.n_edges; e += blockDim.x) {
        const uint2 d = a.ed[e];
        sed[e] = (d.x >> 2) | (((d.y - a.smem_base) >> 2) << 16);      // shift | col*Z << 16  (col*Z < 68*32)
    }
    ShflCtx s;
    s.app = app; s.rec = rec; s.ext = ext; s.sed = sed; s.done = s_done;
    s.cmax = cmax; s.Z = Z; s.stride = ncw; s.alpha = a.alpha;
    const long long n_groups = (a.batch + a.cwpc - 1) / a.cwpc;
    const bool want_ok = a.ok != nullptr;

    for (long long group = blockIdx.x; group < n_groups; group += gridDim.x) {
        const long long cw0 = group * a.cwpc;
        const int n_here = (int)min((long long)a.cwpc, a.batch - cw0);
        __syncthreads();   // previous group's outputs are out of shared memory
        const float *src = a.llr + cw0 * ncw;
        for (int i = tid; i < n_here * ncw; i += blockDim.x) app[i] = clamp_llr(__ldcs(src + i));
        // the degree-1 parity variables start from their channel values (needed if a row is never processed before a read)
        for (int i = tid; i < (rows_all - 4) * cmax; i += blockDim.x) {
            const int r = i / cmax, chk = i - r * cmax, slot = chk / Z, z = chk - slot * Z;
            ext[i] = slot < n_here ? clamp_llr(__ldcs(src + (size_t)slot * ncw + (size_t)(a.kcols + 4 + r) * Z + z)) : 0.f;
        }
        if (tid < a.cwpc) { s_done[tid] = tid < n_here ? 0 : 1; s_fail[tid] = 0; }
        s.C = n_here * Z;
        int my_iters = 0, my_ok = 0;       // meaningful for tid < n_here (thread tid keeps slot tid's counters)
        __syncthreads();

        for (int it = 0; it < a.max_iters; ++it) {
            for (int r = 0; r < a.n_rows; ++r) {
                const int e0 = a.row_start[r], deg = a.row_start[r + 1] - e0;
                NRLDPC_SHFL_DISPATCH(shfl_row, deg, s, r, e0, deg, it == 0);
                __syncthreads();
            }
            if (tid < n_here && !s_done[tid]) my_iters = it + 1;
            const bool last = it + 1 == a.max_iters;
            if (a.early_term || (want_ok && last)) {
                for (int r = 0; r < a.n_rows; ++r) {
                    const int e0 = a.row_start[r], deg = a.row_start[r + 1] - e0;
                    NRLDPC_SHFL_DISPATCH(shfl_syndrome_row, deg, s, r, e0, deg, s_fail);
                }
                __syncthreads();
                int all_done = 1;
                if (tid < a.cwpc) {
                    if (!s_done[tid]) {
                        my_ok = s_fail[tid] ? 0 : 1;
                        if (my_ok && a.early_term) s_done[tid] = 1;
                    }
                    s_fail[tid] = 0;
                    all_done = s_done[tid];
                }
                if (__syncthreads_and(all_done) && a.early_term) break;
            }
        }
        __syncthreads();

        // outputs
        for (int i = tid; i < n_here * K; i += blockDim.x) {
            const int sl = i / K, k = i - sl * K;
            a.hard[(cw0 + sl) * K + k] = (uint8_t)(__float_as_uint(app[sl * ncw + k]) >> 31);
        }
        if (a.soft != nullptr) {
            for (int i = tid; i < n_here * ncw; i += blockDim.x) {
                const int sl = i / ncw, n = i - sl * ncw;
                const int col = n / Z, z = n - col * Z;
                float v = app[i];
                if (col >= a.kcols + 4 && col < a.kcols + a.n_rows) v = ext[(size_t)(col - a.kcols - 4) * cmax + sl * Z + z];
                __stcs(a.soft + cw0 * ncw + i, v);
            }
        }
        if (tid < n_here) {
            if (a.iters) a.iters[cw0 + tid] = my_iters;
            if (a.ok) a.ok[cw0 + tid] = (uint8_t)my_ok;
        }
    }
}

}  // namespace nrldpc
