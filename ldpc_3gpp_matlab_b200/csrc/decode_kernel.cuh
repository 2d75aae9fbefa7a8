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
//   * check-to-variable messages are kept compressed (alpha*min1, alpha*min2, argmin, sign bits =
//     12 B per check) in a per-CTA global scratch that stays L2-resident (<= 63 MB for the whole
//     grid), software-prefetched one layer ahead; the first iteration reads nothing and the last
//     writes nothing;
//   * arithmetic is float32 with every add/mul individually rounded (__fsub_rn/__fmul_rn/__fadd_rn:
//     no FMA contraction), signs handled as sign BITS, so results are bit-identical to the CPU
//     oracle (oracle/nrldpc_oracle.c, orc_decode_nms).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nrldpc {

constexpr int kDecThreads = 384;     // max threads per decode CTA (= largest lifting size)
constexpr int kDecCtasPerSm = 2;
constexpr float kLlrMax = 1048576.0f;

struct DecArgs {
    const float *llr;        // [batch][ncw]
    uint8_t *hard;           // [batch][K]
    float *soft;             // [batch][ncw] or null
    int32_t *iters;          // [batch] or null
    uint8_t *ok;             // [batch] or null
    long long batch;
    int Z, ncols, kcols, n_rows, n_edges, max_iters, early_term, cwpc;
    float alpha;
    const uint32_t *edesc;   // [edges] (col*Z) << 16 | (shift mod Z)
    const int *row_start;    // [rows+1]
    float2 *c2v_mins;        // [grid][n_rows][blockDim]
    uint32_t *c2v_meta;      // [grid][n_rows][blockDim]  argmin | signbits << 5
    int *work_counter;
};

__device__ __forceinline__ float clamp_llr(float x) {
    // NaN marks filler upstream (NRLDPCDecoder.m:224,264): fminf(NaN, M) = M.
    return fmaxf(fminf(x, kLlrMax), -kLlrMax);
}

// One check row of degree DEG for check z.  (om1, om2, ometa) is the compressed message record
// written for this check in the previous iteration (zeros in the first).
template <int DEG>
__device__ __forceinline__ void process_row(float *__restrict__ app, const uint32_t *__restrict__ ed,
                                            const int z, const int Z, const float om1, const float om2,
                                            const uint32_t ometa, const float alpha, float &nm1,
                                            float &nm2, uint32_t &nmeta) {
    float t[DEG];
    int addr[DEG];
    float m1 = __int_as_float(0x7f800000), m2 = __int_as_float(0x7f800000);
    uint32_t sx = 0;
    int arg = 0;
    const int oarg = (int)(ometa & 31u);
    const uint32_t osg = ometa >> 5;
#pragma unroll
    for (int e = 0; e < DEG; ++e) {
        const uint32_t d = ed[e];
        int p = z + (int)(d & 0xffffu);
        p = (p >= Z) ? p - Z : p;
        const int a = (int)(d >> 16) + p;
        addr[e] = a;
        const float x = app[a];
        const float mag = (e == oarg) ? om2 : om1;
        const float c = __uint_as_float(__float_as_uint(mag) | (((osg >> e) & 1u) << 31));
        const float tt = __fsub_rn(x, c);
        t[e] = tt;
        const float ab = fabsf(tt);
        arg = (ab < m1) ? e : arg;
        m2 = fminf(m2, fmaxf(ab, m1));
        m1 = fminf(m1, ab);
        sx ^= __float_as_uint(tt);
    }
    const float m1s = __fmul_rn(alpha, m1), m2s = __fmul_rn(alpha, m2);
    const uint32_t sg = sx & 0x80000000u;
    uint32_t nsg = 0;
#pragma unroll
    for (int e = 0; e < DEG; ++e) {
        const float mag = (e == arg) ? m2s : m1s;
        const uint32_t cb = (__float_as_uint(t[e]) ^ sg) & 0x80000000u;
        const float c = __uint_as_float(__float_as_uint(mag) | cb);
        app[addr[e]] = __fadd_rn(t[e], c);
        nsg |= (cb >> 31) << e;
    }
    nm1 = m1s;
    nm2 = m2s;
    nmeta = (uint32_t)arg | (nsg << 5);
}

__global__ void __launch_bounds__(kDecThreads, kDecCtasPerSm) decode_nms_kernel(const DecArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Z = a.Z;
    const int ncw = a.ncols * Z;
    const int K = a.kcols * Z;
    float *app = reinterpret_cast<float *>(smem_raw);
    uint32_t *s_ed = reinterpret_cast<uint32_t *>(app + (size_t)a.cwpc * ncw);
    int *s_rs = reinterpret_cast<int *>(s_ed + a.n_edges);
    int *s_flag = s_rs + (a.n_rows + 1);
    __shared__ int s_group;

    const int tid = threadIdx.x;
    const int slot = tid / Z;
    const int z = tid - slot * Z;
    const bool lane_ok = tid < a.cwpc * Z;

    for (int i = tid; i < a.n_edges; i += blockDim.x) s_ed[i] = a.edesc[i];
    for (int i = tid; i <= a.n_rows; i += blockDim.x) s_rs[i] = a.row_start[i];

    const long long n_groups = (a.batch + a.cwpc - 1) / a.cwpc;
    const size_t rec_stride = blockDim.x;
    float2 *my_mins = a.c2v_mins + (size_t)blockIdx.x * a.n_rows * rec_stride + tid;
    uint32_t *my_meta = a.c2v_meta + (size_t)blockIdx.x * a.n_rows * rec_stride + tid;
    const bool want_ok = a.ok != nullptr;

    while (true) {
        __syncthreads();  // previous group's outputs are out of smem; tables are loaded
        if (tid == 0) s_group = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const long long group = s_group;
        if (group >= n_groups) break;
        const long long cw0 = group * a.cwpc;
        const int n_here = (int)min((long long)a.cwpc, a.batch - cw0);

        // ---- load + clamp: the group's codewords are contiguous in HBM -----------------------
        {
            const float4 *src = reinterpret_cast<const float4 *>(a.llr + cw0 * ncw);
            float4 *dst = reinterpret_cast<float4 *>(app);
            const int n4 = (n_here * ncw) >> 2;  // ncw*4 bytes is a multiple of 16 for every (BG,Z)
            for (int i = tid; i < n4; i += blockDim.x) {
                float4 v = __ldcs(src + i);
                v.x = clamp_llr(v.x); v.y = clamp_llr(v.y); v.z = clamp_llr(v.z); v.w = clamp_llr(v.w);
                dst[i] = v;
            }
            const int tail = (n_here * ncw) & 3;  // only if ncw is odd-sized (never for valid Z); kept for safety
            for (int i = (n4 << 2) + tid; i < (n4 << 2) + tail; i += blockDim.x)
                app[i] = clamp_llr(a.llr[cw0 * ncw + i]);
        }
        if (tid < a.cwpc) s_flag[tid] = 0;
        __syncthreads();

        const bool active = lane_ok && slot < n_here;
        float *my_app = app + (size_t)slot * ncw;
        bool done = !active;
        int my_iters = 0;
        int my_ok = 0;
        float2 cm = make_float2(0.f, 0.f);
        uint32_t cmeta = 0;

        for (int it = 0; it < a.max_iters; ++it) {
            const bool store_rec = it + 1 < a.max_iters;
            for (int r = 0; r < a.n_rows; ++r) {
                // software prefetch of the next layer's record (it wraps into the next iteration)
                int rn = r + 1, itn = it;
                if (rn == a.n_rows) { rn = 0; itn = it + 1; }
                float2 pm = make_float2(0.f, 0.f);
                uint32_t pmeta = 0;
                if (!done && itn > 0 && itn < a.max_iters) {
                    pm = __ldcg(my_mins + (size_t)rn * rec_stride);
                    pmeta = __ldcg(my_meta + (size_t)rn * rec_stride);
                }
                if (!done) {
                    const int e0 = s_rs[r];
                    const int deg = s_rs[r + 1] - e0;
                    const uint32_t *ed = s_ed + e0;
                    float n1, n2;
                    uint32_t nmeta;
                    switch (deg) {
#define NRLDPC_ROW_CASE(D) case D: process_row<D>(my_app, ed, z, Z, cm.x, cm.y, cmeta, a.alpha, n1, n2, nmeta); break;
                        NRLDPC_ROW_CASE(3) NRLDPC_ROW_CASE(4) NRLDPC_ROW_CASE(5) NRLDPC_ROW_CASE(6)
                        NRLDPC_ROW_CASE(7) NRLDPC_ROW_CASE(8) NRLDPC_ROW_CASE(9) NRLDPC_ROW_CASE(10)
                        NRLDPC_ROW_CASE(19)
#undef NRLDPC_ROW_CASE
                        default: n1 = 0.f; n2 = 0.f; nmeta = 0; break;
                    }
                    if (store_rec) {
                        __stcg(my_mins + (size_t)r * rec_stride, make_float2(n1, n2));
                        __stcg(my_meta + (size_t)r * rec_stride, nmeta);
                    }
                }
                cm = pm;
                cmeta = pmeta;
                __syncthreads();
            }
            if (!done) my_iters = it + 1;
            const bool last = it + 1 == a.max_iters;
            if (a.early_term || (want_ok && last)) {
                // exact syndrome of hard = (app < 0) over the active rows ('Parity check satisfied',
                // NRLDPCDecoder.m:120)
                if (!done) {
                    int fail = 0;
                    for (int r = 0; r < a.n_rows; ++r) {
                        int par = 0;
                        for (int e = s_rs[r]; e < s_rs[r + 1]; ++e) {
                            const uint32_t d = s_ed[e];
                            int p = z + (int)(d & 0xffffu);
                            p = (p >= Z) ? p - Z : p;
                            par ^= (my_app[(d >> 16) + p] < 0.0f) ? 1 : 0;
                        }
                        fail |= par;
                    }
                    if (fail) s_flag[slot] = 1;
                }
                __syncthreads();
                if (!done) {
                    my_ok = s_flag[slot] ? 0 : 1;
                    if (my_ok && a.early_term) done = true;
                }
                const int all_done = __syncthreads_and(done ? 1 : 0);  // also orders the flag reset below
                if (tid < a.cwpc) s_flag[tid] = 0;
                if (a.early_term && all_done) break;
            }
        }
        __syncthreads();

        // ---- outputs --------------------------------------------------------------------------
        for (int s = 0; s < n_here; ++s) {
            const float *src = app + (size_t)s * ncw;
            uint8_t *dst = a.hard + (cw0 + s) * K;
            for (int k = tid; k < K; k += blockDim.x) dst[k] = src[k] < 0.0f ? 1 : 0;
        }
        if (a.soft) {
            float4 *dst = reinterpret_cast<float4 *>(a.soft + cw0 * ncw);
            const float4 *src = reinterpret_cast<const float4 *>(app);
            const int n4 = (n_here * ncw) >> 2;
            for (int i = tid; i < n4; i += blockDim.x) __stcs(dst + i, src[i]);
        }
        if (active && z == 0) {
            if (a.iters) a.iters[cw0 + slot] = my_iters;
            if (a.ok) a.ok[cw0 + slot] = (uint8_t)my_ok;
        }
    }
}

}  // namespace nrldpc
