// decode_kernel_bp.cuh -- the REFERENCE's decoding algorithm on the device: flooding sum-product in float64
// with 'Parity check satisfied' termination, i.e. what comm.LDPCDecoder computes as configured at
// NRLDPCDecoder.m:120 and called at :265 (MathWorks' documented algorithm: q_ij <- L(c_i); every iteration all
// checks r_ji = 2 atanh(prod_{i' != i} tanh(q_i'j / 2)), then all variables Q_i = L(c_i) + sum_j r_ji,
// q_ij = Q_i - r_ji, hard = (Q_i < 0), stop when H * hard = 0).
//
// Selected with nrldpc_cfg.algorithm = NRLDPC_ALG_BP.  It exists so that a user of the reference can keep the
// reference's own algorithm (same BLER curve, same iteration counts) behind the same boundary; the layered
// normalized min-sum kernels (decode_kernel.cuh) remain the default and the fast path.
//
// Every floating-point operation is performed in the same order as in the CPU restatement it is tested
// against (oracle/nrldpc_oracle.c, decode_bp_one): leave-one-out products by prefix / suffix products, the
// atanh argument clipped to +-(1 - 2^-53), per-variable accumulation = channel LLR first, then the check
// messages in ascending base-row order.  The only difference is the tanh / atanh implementation (CUDA's
// double-precision math library here, the host libm there; both are accurate to about one unit in the last place).
//
// Mapping: one codeword per CTA at a time (persistent CTAs, device work counter).  The a-posteriori values Q
// (cols*Z doubles, 209 KB at BG1 / Z = 384) live in shared memory; the check-to-variable messages r (one double
// per edge of H) live in a per-CTA global scratch laid out [base edge][z], so the check phase (thread = check)
// and the variable phase (thread = variable, gathering the column's edges with the circulant shift) are both
// coalesced.  Flooding needs only two CTA barriers per iteration (no barrier between base rows).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace nrldpc {

// CTA width is a template parameter: 512 threads leave 128 registers per thread (the degree-19 rows keep 2 x 19
// doubles live), 1024 threads double the warps that hide the latency of the float64 math sequences and of the message
// loads at 64 registers.  Measured on 4096 x BG1 Z=384, 8 iterations: 102 ms at 512 threads, 73 ms at 1024 (default).
constexpr int kBpThreadsMax = 1024;

struct BpArgs {
    const void *llr;         // [batch][ncw] float32 or float64 (template parameter)
    uint8_t *hard;           // [batch][K]
    void *soft;              // [batch][ncw] same type as llr, or null
    int32_t *iters;          // [batch] or null
    uint8_t *ok;             // [batch] or null
    long long batch;
    int Z, ncols, kcols, n_rows, n_edges, max_iters, early_term;
    double *rmsg;            // [grid][n_edges_total][Z] check-to-variable messages
    long long rmsg_stride;   // doubles per CTA
    int *work_counter;
    // row-major (CSR) view: edges sorted by (row, col): row_start[r] .. row_start[r+1]; per edge shift and col*Z
    const int *row_start;    // [rows + 1]
    const int *e_shift;      // [edges]
    const int *e_colz;       // [edges] col * Z
    // column-major (CSC) view for the variable phase: for column c the edges col_edge[col_start[c] .. col_start[c+1])
    // in ascending edge (= ascending row) order
    const int *col_start;    // [cols + 1]
    const int *col_edge;     // [edges]
};

// The two float64 math sequences are kept out of line: inlined into the nine per-degree row bodies they made ~280 KB
// of code, and the kernel stalled on instruction fetch (ncu: stall_no_instruction 1.3 per issue).
#ifndef NRLDPC_BP_INLINE_MATH
#define NRLDPC_BP_INLINE_MATH 0
#endif
#if NRLDPC_BP_INLINE_MATH
#define NRLDPC_BP_MATH __device__ __forceinline__
#else
#define NRLDPC_BP_MATH __device__ __noinline__
#endif
NRLDPC_BP_MATH double bp_tanh_half(double q) { return tanh(__dmul_rn(0.5, q)); }
NRLDPC_BP_MATH double bp_two_atanh_clipped(double x) {
    const double lim = 1.0 - 1.1102230246251565e-16;   // 1 - 2^-53: +inf filler cannot produce inf - inf
    x = x > lim ? lim : (x < -lim ? -lim : x);
    return __dmul_rn(2.0, atanh(x));
}

// One check of degree DEG: reads Q (shared) and its own previous messages, writes its new messages in place.
template <int DEG>
__device__ __forceinline__ void bp_check(const double *__restrict__ Q, double *__restrict__ rm, const int *__restrict__ e_shift,
                                         const int *__restrict__ e_colz, const int e0, const int z, const int Z,
                                         const bool first) {
    double th[DEG], q[DEG];
#pragma unroll
    for (int k = 0; k < DEG; ++k) {
        int p = z + __ldg(e_shift + e0 + k);
        if (p >= Z) p -= Z;
        const double Qv = Q[__ldg(e_colz + e0 + k) + p];
        // q_ij = Q_i - r_ji; in the first iteration q_ij = L(c_i) (Q holds the channel values, no message yet: the
        // slot is read anyway -- it is this CTA's own scratch -- so that the load never waits behind a branch)
        const double r_old = rm[(size_t)(e0 + k) * Z + z];
        q[k] = first ? Qv : __dsub_rn(Qv, r_old);
    }
#pragma unroll
    for (int k = 0; k < DEG; ++k) th[k] = bp_tanh_half(q[k]);
    // suffix products suf[k] = th[k] * th[k+1] * ... (formed from the right, as the restatement does)
    double suf[DEG + 1];
    suf[DEG] = 1.0;
#pragma unroll
    for (int k = DEG - 1; k >= 0; --k) suf[k] = __dmul_rn(suf[k + 1], th[k]);
    double pre = 1.0;
#pragma unroll
    for (int k = 0; k < DEG; ++k) {
        rm[(size_t)(e0 + k) * Z + z] = bp_two_atanh_clipped(__dmul_rn(pre, suf[k + 1]));
        pre = __dmul_rn(pre, th[k]);
    }
}

template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) decode_bp_kernel(const BpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw_bp[];
    double *Q = reinterpret_cast<double *>(smem_raw_bp);
    __shared__ int s_group;
    const int Z = a.Z;
    const int ncw = a.ncols * Z;
    const int K = a.kcols * Z;
    const int n_checks = a.n_rows * Z;
    const int tid = threadIdx.x, nt = blockDim.x;
    double *rm = a.rmsg + (size_t)blockIdx.x * a.rmsg_stride;
    const T *llr_all = static_cast<const T *>(a.llr);
    T *soft_all = static_cast<T *>(a.soft);

    while (true) {
        __syncthreads();
        if (tid == 0) s_group = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const long long cw = s_group;
        if (cw >= a.batch) break;
        const T *llr = llr_all + cw * ncw;
        // NaN marks filler upstream (NRLDPCDecoder.m:224); the reference turns it into +inf before decoding (:264)
        for (int v = tid; v < ncw; v += nt) {
            const double x = (double)llr[v];
            Q[v] = (x != x) ? (double)INFINITY : x;
        }
        __syncthreads();

        int it = 0, ok = 0;
        while (it < a.max_iters) {
            // ---- check phase: all checks of all active rows, from the same Q (flooding)
            for (int c = tid; c < n_checks; c += nt) {
                const int r = c / Z, z = c - r * Z;
                const int e0 = __ldg(a.row_start + r);
                const int deg = __ldg(a.row_start + r + 1) - e0;
                const bool first = it == 0;
                switch (deg) {
#define NRLDPC_BP_CASE(D) case D: bp_check<D>(Q, rm, a.e_shift, a.e_colz, e0, z, Z, first); break;
                    NRLDPC_BP_CASE(3) NRLDPC_BP_CASE(4) NRLDPC_BP_CASE(5) NRLDPC_BP_CASE(6) NRLDPC_BP_CASE(7)
                    NRLDPC_BP_CASE(8) NRLDPC_BP_CASE(9) NRLDPC_BP_CASE(10) NRLDPC_BP_CASE(19)
#undef NRLDPC_BP_CASE
                    default: break;
                }
            }
            __syncthreads();
            // ---- variable phase: Q_i = L(c_i) + sum of the incoming messages in ascending row order
            for (int v = tid; v < ncw; v += nt) {
                const int col = v / Z, i = v - col * Z;
                const double x = (double)llr[v];
                double acc = (x != x) ? (double)INFINITY : x;
                const int c1 = __ldg(a.col_start + col + 1);
                for (int j = __ldg(a.col_start + col); j < c1; ++j) {
                    const int e = __ldg(a.col_edge + j);
                    if (e >= a.n_edges) break;   // edges of trimmed rows (the list is ascending)
                    int z = i - __ldg(a.e_shift + e);
                    if (z < 0) z += Z;
                    acc = __dadd_rn(acc, rm[(size_t)e * Z + z]);
                }
                Q[v] = acc;
            }
            __syncthreads();
            ++it;
            // ---- 'Parity check satisfied': exact syndrome of hard = (Q < 0) over the active rows
            const bool last = it == a.max_iters;
            if (a.early_term || (last && a.ok)) {
                int fail = 0;
                for (int c = tid; c < n_checks; c += nt) {
                    const int r = c / Z, z = c - r * Z;
                    int par = 0;
                    const int e1 = __ldg(a.row_start + r + 1);
                    for (int e = __ldg(a.row_start + r); e < e1; ++e) {
                        int p = z + __ldg(a.e_shift + e);
                        if (p >= Z) p -= Z;
                        par ^= Q[__ldg(a.e_colz + e) + p] < 0.0 ? 1 : 0;
                    }
                    fail |= par;
                }
                ok = __syncthreads_or(fail) ? 0 : 1;
                if (ok && a.early_term) break;
            }
        }

        uint8_t *hard = a.hard + cw * K;
        for (int k = tid; k < K; k += nt) hard[k] = Q[k] < 0.0 ? 1 : 0;
        if (soft_all) {
            T *soft = soft_all + cw * ncw;
            for (int v = tid; v < ncw; v += nt) soft[v] = (T)Q[v];
        }
        if (tid == 0) {
            if (a.iters) a.iters[cw] = it;
            if (a.ok) a.ok[cw] = (uint8_t)ok;
        }
    }
}

// float64 LLRs for the min-sum kernels (nrldpc_decode64 with the default algorithm): round to float32 on the device
__global__ void __launch_bounds__(256) narrow_f64_kernel(const double *__restrict__ in, float *__restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (float)in[i];
}

}  // namespace nrldpc
