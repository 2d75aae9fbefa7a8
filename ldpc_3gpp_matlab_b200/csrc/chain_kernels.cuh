// chain_kernels.cuh -- companion kernels either side of the decoder: QC encoder, rate matching,
// rate recovery (+HARQ combine, filler / puncture handling) and the QPSK/AWGN/LLR channel leg.
// All HBM-bound byte/float streaming: whole rows staged through shared memory with coalesced vector
// accesses on the HBM side, permutations applied on chip; no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nrldpc {

// ------------------------------------------------------------------------------------------------
// Encoder: replaces step(obj.hLDPCEncoder, c) (NRLDPCEncoder.m:158).  H = [A B 0; C D I] with B
// dual-diagonal, so parity is found by back-substitution instead of a generic GF(2) solve:
//   lambda_r = sum_{systematic cols} P^{s} c      (r = 0..3)
//   P^{delta} p0 = lambda_0 + lambda_1 + lambda_2 + lambda_3
//   p1, p2, p3 by substitution down the diagonal; extension parity p_r = row r times [c p0..p3].
// ------------------------------------------------------------------------------------------------
struct EncArgs {
    const uint8_t *info;      // [batch][K]
    uint8_t *cw;              // [batch][ncw]
    long long batch;
    int Z, ncols, kcols, n_rows, n_edges;
    const uint32_t *edesc;    // (col*Z) << 16 | shift
    const int *row_start;
    int s0[4];                // shift of column kcols in rows 0..3, -1 if absent
    int delta;                // surviving rotation of p0 in the sum of the four core rows
};

__device__ __forceinline__ int wrap_add(int z, int s, int Z) {
    int p = z + s;
    return p >= Z ? p - Z : p;
}

// Bit-sliced: a CTA encodes a slab of 32 codewords at once, codeword c living in bit c of one 32-bit word
// per codeword position (shared memory: cols*Z words), so every XOR of the back-substitution serves 32
// codewords and the byte-per-bit boundary costs one pack pass (uchar4 loads) and one unpack pass (uchar4
// stores, 128 contiguous bytes per warp and codeword row).  HBM-bound: K bytes in, cols*Z bytes out.
constexpr int kEncSlab = 32;

__global__ void __launch_bounds__(384) encode_kernel(const EncArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Z = a.Z, K = a.kcols * Z, ncw = a.ncols * Z;
    uint32_t *w = reinterpret_cast<uint32_t *>(smem_raw);     // [ncw] bit-sliced codeword positions
    uint32_t *s_ed = w + ncw;
    int *s_rs = reinterpret_cast<int *>(s_ed + a.n_edges);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < a.n_edges; i += nt) s_ed[i] = a.edesc[i];
    for (int i = tid; i <= a.n_rows; i += nt) s_rs[i] = a.row_start[i];
    const long long n_slabs = (a.batch + kEncSlab - 1) / kEncSlab;
    const bool vec4 = (K & 3) == 0;   // K = kcols*Z, ncw = cols*Z: multiples of 4 for every even Z (and ncw always)

    for (long long slab = blockIdx.x; slab < n_slabs; slab += gridDim.x) {
        __syncthreads();
        const long long cw0 = slab * kEncSlab;
        const int n_here = (int)min((long long)kEncSlab, a.batch - cw0);
        // ---- pack: info bytes of the slab's codewords -> bit-sliced words
        if (vec4) {
            for (int q = tid; q < (K >> 2); q += nt) {
                uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
                const uchar4 *src = reinterpret_cast<const uchar4 *>(a.info + cw0 * K) + q;
                for (int c = 0; c < n_here; ++c) {
                    const uchar4 v = src[(size_t)c * (K >> 2)];
                    w0 |= (uint32_t)(v.x & 1) << c; w1 |= (uint32_t)(v.y & 1) << c;
                    w2 |= (uint32_t)(v.z & 1) << c; w3 |= (uint32_t)(v.w & 1) << c;
                }
                reinterpret_cast<uint4 *>(w)[q] = make_uint4(w0, w1, w2, w3);
            }
        } else {
            for (int i = tid; i < K; i += nt) {
                uint32_t acc = 0;
                for (int c = 0; c < n_here; ++c) acc |= (uint32_t)(a.info[(cw0 + c) * K + i] & 1) << c;
                w[i] = acc;
            }
        }
        __syncthreads();
        // ---- core parity: lambda_r, p0, then substitution down the dual diagonal
        for (int z = tid; z < Z; z += nt) {
            uint32_t lam[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                uint32_t acc = 0;
                for (int e = s_rs[r]; e < s_rs[r + 1]; ++e) {
                    const uint32_t d = s_ed[e];
                    const int cb = (int)(d >> 16);
                    if (cb < K) acc ^= w[cb + wrap_add(z, (int)(d & 0xffffu), Z)];
                }
                lam[r] = acc;
            }
            w[K + wrap_add(z, a.delta, Z)] = lam[0] ^ lam[1] ^ lam[2] ^ lam[3];
            w[K + Z + z] = lam[0];       // stash lambda for the substitution after the barrier
            w[K + 2 * Z + z] = lam[1];
            w[K + 3 * Z + z] = lam[2];
        }
        __syncthreads();
        for (int z = tid; z < Z; z += nt) {
            uint32_t prev = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                uint32_t v = w[K + (r + 1) * Z + z];  // lambda_r (thread-private slot)
                if (a.s0[r] >= 0) v ^= w[K + wrap_add(z, a.s0[r], Z)];
                if (r > 0) v ^= prev;
                prev = v;
                w[K + (r + 1) * Z + z] = v;            // p_{r+1}[z]
            }
        }
        __syncthreads();
        // ---- extension rows: independent of each other
        for (int z = tid; z < Z; z += nt) {
            for (int r = 4; r < a.n_rows; ++r) {
                uint32_t acc = 0;
                const int e1 = s_rs[r + 1] - 1;  // last entry of an extension row is its own identity
                for (int e = s_rs[r]; e < e1; ++e) {
                    const uint32_t d = s_ed[e];
                    acc ^= w[(int)(d >> 16) + wrap_add(z, (int)(d & 0xffffu), Z)];
                }
                w[K + r * Z + z] = acc;
            }
        }
        __syncthreads();
        // ---- unpack: bit c of every word -> codeword row cw0 + c (four positions per 32-bit store)
        for (int q = tid; q < (ncw >> 2); q += nt) {
            const uint4 v = reinterpret_cast<const uint4 *>(w)[q];
            uint32_t *dst = reinterpret_cast<uint32_t *>(a.cw + cw0 * ncw) + q;
            for (int c = 0; c < n_here; ++c)
                dst[(size_t)c * (ncw >> 2)] = ((v.x >> c) & 1u) | (((v.y >> c) & 1u) << 8) | (((v.z >> c) & 1u) << 16) |
                                              (((v.w >> c) & 1u) << 24);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Rate matching geometry shared by TX and RX.  The reference walks the circular buffer with a
// scalar while-loop that skips filler (NRLDPCEncoder.m:187-195, NRLDPCDecoder.m:226-234); here the
// k-th selected position has a closed form.  In d-coordinates (d = cw without the 2Z punctured
// columns) the filler interval clipped to the circular buffer is [F0, F1); non-filler positions are
// ranked 0..Nnf-1 by rank(x) = x - clamp(x - F0, 0, F1 - F0) and unrank(m) = m < F0 ? m : m + (F1-F0).
// ------------------------------------------------------------------------------------------------
struct RmGeom {
    int E, Ncb, F0, F1, Nnf, rank_k0, Qm, EQ;  // EQ = E / Qm
    int Z2;                                    // 2Z
    int N, ncw;
    int F0u, F1u;                              // filler interval before clipping to Ncb
};

__device__ __forceinline__ int rm_unrank(const RmGeom &g, int m) { return m < g.F0 ? m : m + (g.F1 - g.F0); }
__device__ __forceinline__ int rm_rank(const RmGeom &g, int x) {
    int t = x - g.F0;
    t = t < 0 ? 0 : (t > g.F1 - g.F0 ? g.F1 - g.F0 : t);
    return x - t;
}

// TX: f[n] = e[(n % Qm) * E/Qm + n / Qm]  (NRLDPCEncoder.m:219-223), e[k] = d[unrank((rank(k0)+k) % Nnf)]
__global__ void __launch_bounds__(256) rate_match_kernel(const uint8_t *__restrict__ cw, uint8_t *__restrict__ f,
                                                         long long batch, RmGeom g) {
    const long long total = batch * g.E;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / g.E;
        const int n = (int)(i - b * g.E);
        const int k = (n % g.Qm) * g.EQ + n / g.Qm;
        const int m = (int)(((long long)g.rank_k0 + k) % g.Nnf);
        f[i] = cw[b * g.ncw + g.Z2 + rm_unrank(g, m)];
    }
}

// RX: one thread per entry of the decoder's input layout.  Soft-combining of wrapped repetitions
// is a gather in increasing k, i.e. the same addition order as the reference's serial loop
// (NRLDPCDecoder.m:230), followed by the HARQ accumulate (:236-239).
__global__ void __launch_bounds__(256) rate_recover_kernel(const float *__restrict__ f, float *__restrict__ harq,
                                                           float *__restrict__ out, long long batch, RmGeom g) {
    const long long total = batch * g.ncw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / g.ncw;
        const int c = (int)(i - b * g.ncw);
        const int n = c - g.Z2;
        float v = 0.0f;
        if (n >= 0) {
            if (n >= g.F0u && n < g.F1u) {
                v = __int_as_float(0x7f800000);  // filler: known 0 (NRLDPCDecoder.m:264)
            } else if (n < g.Ncb) {
                int first = rm_rank(g, n) - g.rank_k0;
                if (first < 0) first += g.Nnf;
                const float *fb = f + b * g.E;
                float acc = 0.0f;
                for (int k = first; k < g.E; k += g.Nnf) {
                    const int q = k / g.EQ;                    // NRLDPCDecoder.m:191-195
                    acc = __fadd_rn(acc, fb[q + (k - q * g.EQ) * g.Qm]);
                }
                if (harq) {
                    float *hb = harq + b * g.N + n;
                    acc = __fadd_rn(acc, *hb);
                    *hb = acc;
                }
                v = acc;
            }
        }
        out[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Shared-memory staged variants (one CTA per code block at a time): the HBM side of both kernels becomes
// plain coalesced vector loads / stores of whole rows; the permutation (circular buffer, filler gap,
// bit interleaver) is applied between shared memory and registers.  Used whenever a row fits in shared
// memory; the element-wise kernels above remain as the fallback for very long E.
// ------------------------------------------------------------------------------------------------
template <int QM>
__device__ __forceinline__ int rm_src(const RmGeom &g, const int n, const bool wraps) {
    // f[n] = e[(n % Qm) * E/Qm + n / Qm]; e[k] = d[unrank((rank(k0) + k) mod Nnf)]
    const int j = n / QM, i = n - j * QM;
    int m = g.rank_k0 + i * g.EQ + j;
    if (wraps) m %= g.Nnf;
    else if (m >= g.Nnf) m -= g.Nnf;
    return g.Z2 + rm_unrank(g, m);
}

template <int QM>
__global__ void __launch_bounds__(256) rate_match_staged_kernel(const uint8_t *__restrict__ cw, uint8_t *__restrict__ f,
                                                                long long batch, RmGeom g) {
    extern __shared__ __align__(16) unsigned char rm_smem[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const bool wraps = (long long)g.rank_k0 + g.E > 2LL * g.Nnf;   // more than one lap of the circular buffer
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        __syncthreads();
        const uint32_t *src = reinterpret_cast<const uint32_t *>(cw + b * g.ncw);   // ncw is a multiple of 4
        for (int i = tid; i < (g.ncw >> 2); i += nt) reinterpret_cast<uint32_t *>(rm_smem)[i] = src[i];
        __syncthreads();
        uint8_t *dst = f + b * g.E;
        const int head = (int)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);    // bytes before 4-byte alignment
        const int nq = g.E > head ? (g.E - head) >> 2 : 0;
        for (int q = tid; q < nq; q += nt) {
            const int n = head + 4 * q;
            const uint32_t v = (uint32_t)rm_smem[rm_src<QM>(g, n, wraps)] | ((uint32_t)rm_smem[rm_src<QM>(g, n + 1, wraps)] << 8) |
                               ((uint32_t)rm_smem[rm_src<QM>(g, n + 2, wraps)] << 16) | ((uint32_t)rm_smem[rm_src<QM>(g, n + 3, wraps)] << 24);
            *reinterpret_cast<uint32_t *>(dst + n) = v;
        }
        for (int n = tid; n < g.E; n += nt)
            if (n < head || n >= head + 4 * nq) dst[n] = rm_smem[rm_src<QM>(g, n, wraps)];
    }
}

template <int QM>
__global__ void __launch_bounds__(256) rate_recover_staged_kernel(const float *__restrict__ f, float *__restrict__ harq,
                                                                  float *__restrict__ out, long long batch, RmGeom g) {
    extern __shared__ __align__(16) unsigned char rm_smem[];
    float *e_s = reinterpret_cast<float *>(rm_smem);            // de-interleaved LLRs e[k], NRDecoder.m:191-195
    const int tid = threadIdx.x, nt = blockDim.x;
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        __syncthreads();
        const float *fb = f + b * g.E;
        for (int n = tid; n < g.E; n += nt) {
            const int j = n / QM, i = n - j * QM;
            e_s[i * g.EQ + j] = fb[n];
        }
        __syncthreads();
        float4 *ob = reinterpret_cast<float4 *>(out + b * g.ncw);   // ncw*4 bytes is a multiple of 16
        for (int q = tid; q < (g.ncw >> 2); q += nt) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int n = 4 * q + u - g.Z2;
                float r = 0.0f;
                if (n >= 0) {
                    if (n >= g.F0u && n < g.F1u) {
                        r = __int_as_float(0x7f800000);  // filler: known 0 (NRLDPCDecoder.m:264)
                    } else if (n < g.Ncb) {
                        int first = rm_rank(g, n) - g.rank_k0;
                        if (first < 0) first += g.Nnf;
                        float acc = 0.0f;
                        for (int k = first; k < g.E; k += g.Nnf) acc = __fadd_rn(acc, e_s[k]);   // same order as :230
                        if (harq) {
                            float *hb = harq + b * g.N + n;
                            acc = __fadd_rn(acc, *hb);
                            *hb = acc;
                        }
                        r = acc;
                    }
                }
                v[u] = r;
            }
            ob[q] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// TMA variant of rate recovery: each demodulator-LLR row (E floats) is brought into shared memory by ONE
// bulk asynchronous copy (cp.async.bulk, completion on an mbarrier), double-buffered so the copy of the
// next code block runs under the gather / soft-combine / store of the current one.  The bit de-interleaver
// is folded into the gather index (e[k] = f[k / (E/Qm) + (k mod E/Qm) * Qm], NRLDPCDecoder.m:191-195).
// Needs 16-byte aligned rows (E a multiple of 4); other shapes use the kernels above.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// streamed-once input: evict_first so that it does not displace L2-pinned state
__device__ __forceinline__ void bulk_g2s_stream(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
                 "r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src_gmem), "r"(bytes),
                 "r"((uint32_t)__cvta_generic_to_shared(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                 "r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src_gmem), "r"(bytes),
                 "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Gather table of rate recovery (round 2).  The map "entry c of the decoder's input layout <- received LLR(s)" depends only on
// the rate-matching geometry, not on the code block, so it is evaluated ONCE per geometry into a table in global memory (it stays
// in L2; every CTA reads it with 16-byte loads) instead of once per code block and entry (~120 instructions each):
//     code >= 0   the entry has exactly one source: f[code] (bit de-interleaver already folded in)
//     kRrZero     punctured prefix, or beyond the circular buffer: 0, no HARQ state
//     kRrFiller   filler: +inf (NRLDPCDecoder.m:264)
//     kRrEmpty    inside the circular buffer but not transmitted this time: 0 + the HARQ buffer
//     kRrMulti    wrapped repetitions (several sources, soft-combined in increasing k): the general loop below
// Same additions in the same order as before (0 + x, then + HARQ), hence bit-identical outputs.
// ------------------------------------------------------------------------------------------------
constexpr int kRrZero = -1, kRrFiller = -2, kRrEmpty = -3, kRrMulti = -4;

__global__ void __launch_bounds__(256) rr_table_kernel(int *__restrict__ table, RmGeom g, int QM, uint32_t magic_eq) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < g.ncw; c += gridDim.x * blockDim.x) {
        const int n = c - g.Z2;
        int code = kRrZero;
        if (n >= 0) {
            if (n >= g.F0u && n < g.F1u) {
                code = kRrFiller;
            } else if (n < g.Ncb) {
                int first = rm_rank(g, n) - g.rank_k0;
                if (first < 0) first += g.Nnf;
                if (first >= g.E) {
                    code = kRrEmpty;
                } else if (first + g.Nnf < g.E) {
                    code = kRrMulti;
                } else {
                    int i = (int)__umulhi((uint32_t)first, magic_eq);      // first / EQ, one below at most
                    int j = first - i * g.EQ;
                    if (j >= g.EQ) { j -= g.EQ; ++i; }
                    code = i + j * QM;
                }
            }
        }
        table[c] = code;
    }
}

// soft-combined repetitions of entry n (n in the circular buffer, not filler), e_f = the block's received LLRs in f order
__device__ __noinline__ float rr_multi(const RmGeom &g, const int n, const float *e_f, const uint32_t magic_eq, const int QM) {
    int first = rm_rank(g, n) - g.rank_k0;
    if (first < 0) first += g.Nnf;
    float acc = 0.0f;
    for (int k = first; k < g.E; k += g.Nnf) {
        int i = (int)__umulhi((uint32_t)k, magic_eq);
        int j = k - i * g.EQ;
        if (j >= g.EQ) { j -= g.EQ; ++i; }
        acc = __fadd_rn(acc, e_f[i + j * QM]);                             // same addition order as NRLDPCDecoder.m:230
    }
    return acc;
}

// one float4 of the decoder's input layout from the table (q = index of the float4 in the row)
__device__ __forceinline__ float4 rr_gather4(const RmGeom &g, const int *__restrict__ table, const int q, const float *e_f,
                                             float *__restrict__ harq_row, const uint32_t magic_eq, const int QM) {
    const int4 c4 = __ldg(reinterpret_cast<const int4 *>(table) + q);
    const int code[4] = {c4.x, c4.y, c4.z, c4.w};
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int cd = code[u];
        float r = 0.0f;
        if (cd == kRrFiller) {
            r = __int_as_float(0x7f800000);
        } else if (cd != kRrZero) {
            float acc = 0.0f;
            if (cd >= 0) acc = __fadd_rn(acc, e_f[cd]);
            else if (cd == kRrMulti) acc = rr_multi(g, 4 * q + u - g.Z2, e_f, magic_eq, QM);
            if (harq_row) {
                float *hb = harq_row + (4 * q + u - g.Z2);
                acc = __fadd_rn(acc, *hb);
                *hb = acc;
            }
            r = acc;
        }
        v[u] = r;
    }
    return make_float4(v[0], v[1], v[2], v[3]);
}

template <int QM>
__global__ void __launch_bounds__(512) rate_recover_tma_kernel(const float *__restrict__ f, float *__restrict__ harq,
                                                               float *__restrict__ out, long long batch, RmGeom g,
                                                               int n_buf, uint32_t magic_eq, const int *__restrict__ table) {
    extern __shared__ __align__(128) unsigned char rr_tma_smem[];
    __shared__ __align__(8) uint64_t bar[2];
    const int tid = threadIdx.x, nt = blockDim.x;
    const uint32_t row_bytes = (uint32_t)g.E * 4u;
    float *buf[2] = {reinterpret_cast<float *>(rr_tma_smem), reinterpret_cast<float *>(rr_tma_smem + (n_buf > 1 ? row_bytes : 0))};
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long b = blockIdx.x;
    if (tid == 0 && b < batch) {
        mbar_expect_tx(&bar[0], row_bytes);
        bulk_g2s_stream(buf[0], f + b * g.E, row_bytes, &bar[0]);
    }
    uint32_t phase[2] = {0u, 0u};
    for (int it = 0; b < batch; b += gridDim.x, ++it) {
        const int cur = n_buf > 1 ? (it & 1) : 0;
        const long long nb = b + gridDim.x;
        if (n_buf > 1 && tid == 0 && nb < batch) {          // prefetch the next row into the other buffer (free since the
            mbar_expect_tx(&bar[cur ^ 1], row_bytes);       // barrier at the end of the previous iteration)
            bulk_g2s_stream(buf[cur ^ 1], f + nb * g.E, row_bytes, &bar[cur ^ 1]);
        }
        mbar_wait(&bar[cur], phase[cur]);
        phase[cur] ^= 1u;
        const float *e_f = buf[cur];
        float4 *ob = reinterpret_cast<float4 *>(out + b * g.ncw);
        if (table != nullptr) {
            float *hrow = harq ? harq + b * g.N : nullptr;
            for (int q = tid; q < (g.ncw >> 2); q += nt) __stcs(ob + q, rr_gather4(g, table, q, e_f, hrow, magic_eq, QM));
        } else
        for (int q = tid; q < (g.ncw >> 2); q += nt) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int n = 4 * q + u - g.Z2;
                float r = 0.0f;
                if (n >= 0) {
                    if (n >= g.F0u && n < g.F1u) {
                        r = __int_as_float(0x7f800000);  // filler: known 0 (NRLDPCDecoder.m:264)
                    } else if (n < g.Ncb) {
                        int first = rm_rank(g, n) - g.rank_k0;
                        if (first < 0) first += g.Nnf;
                        float acc = 0.0f;
                        for (int k = first; k < g.E; k += g.Nnf) {
                            int i = (int)__umulhi((uint32_t)k, magic_eq);      // k / EQ, one below at most
                            int j = k - i * g.EQ;
                            if (j >= g.EQ) { j -= g.EQ; ++i; }
                            acc = __fadd_rn(acc, e_f[i + j * QM]);             // same addition order as :230
                        }
                        if (harq) {
                            float *hb = harq + b * g.N + n;
                            acc = __fadd_rn(acc, *hb);
                            *hb = acc;
                        }
                        r = acc;
                    }
                }
                v[u] = r;
            }
            __stcs(ob + q, make_float4(v[0], v[1], v[2], v[3]));
        }
        __syncthreads();   // every thread is done with buf[cur] before it is refilled
        if (n_buf == 1 && tid == 0 && nb < batch) {
            mbar_expect_tx(&bar[0], row_bytes);
            bulk_g2s_stream(buf[0], f + nb * g.E, row_bytes, &bar[0]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// QPSK + AWGN + exact LLR (plot_BLER_vs_SNR.m:130-132).  Counter-based Philox4x32-10 keyed by
// (seed, stream_id); counter = global symbol index / 2; Box-Muller turns the four 32-bit outputs
// into four N(0,1) samples (two complex noise samples = two QPSK symbols).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void box_muller(uint32_t u0, uint32_t u1, float &n0, float &n1) {
    const float a = ((float)(u0 >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
    const float b = ((float)(u1 >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * __logf(a));
    float s, c;
    __sincosf(6.283185307179586f * b, &s, &c);
    n0 = r * c;
    n1 = r * s;
}

__global__ void __launch_bounds__(256) qpsk_awgn_llr_kernel(const uint8_t *__restrict__ bits, float *__restrict__ llr,
                                                            long long n_quads, float sigma, float gain,
                                                            uint64_t seed, uint64_t stream_id) {
    // one thread per 4 bits (2 QPSK symbols); total bit count is a multiple of 4 or handled by the tail launch
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_quads;
         i += (long long)gridDim.x * blockDim.x) {
        uint32_t r[4];
        philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32),
                      (uint32_t)seed, (uint32_t)(seed >> 32), r);
        float n[4];
        box_muller(r[0], r[1], n[0], n[1]);
        box_muller(r[2], r[3], n[2], n[3]);
        const uchar4 b = reinterpret_cast<const uchar4 *>(bits)[i];
        const float amp = 0.70710678118654752440f;
        float4 o;
        o.x = gain * ((b.x ? -amp : amp) + sigma * n[0]);
        o.y = gain * ((b.y ? -amp : amp) + sigma * n[1]);
        o.z = gain * ((b.z ? -amp : amp) + sigma * n[2]);
        o.w = gain * ((b.w ? -amp : amp) + sigma * n[3]);
        reinterpret_cast<float4 *>(llr)[i] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// Channel leg for every modulation of NRModulator.m / NRDemodulator.m (BPSK, QPSK, 16/64/256-QAM).
// Constellations follow the TS 38.211 section 5.1 formulas, which the reference expresses as toolbox
// CustomSymbolMapping vectors (NRModulator.m:73-81); they are separable: the real part is a PAM level of
// the even bits b0,b2,.., the imaginary part of the odd bits b1,b3,.. (BPSK: both parts from b0).
//   level(c0..c_{m-1}) = (1-2c0) * [2^(m-1) - (1-2c1) * [2^(m-2) - ... [2 - (1-2c_{m-1})]]]   (odd integer)
//   amplitude = level * {1/sqrt2, 1/sqrt10, 1/sqrt42, 1/sqrt170}
// Demodulation (NRDemodulator.m:72-92, DecisionMethod): exact LLR = log sum_{s:b=0} exp(-|r-s|^2/var)
// - log sum_{s:b=1} exp(-|r-s|^2/var) (separable, so a log-sum-exp over the 2^m levels of one dimension),
// approximate LLR = max-log, hard decision = bit of the nearest point.  BPSK / QPSK use the closed
// forms 2*sqrt2*(x+y)/var and 2*sqrt2*x/var.
// ------------------------------------------------------------------------------------------------
constexpr int kDemodExact = 0, kDemodApprox = 1, kDemodHard = 2;

__host__ __device__ __forceinline__ int pam_level(const int m, const uint32_t code) {
    // code: c0 is the most significant of the m bits
    int r = 1;
    for (int j = m - 1; j >= 1; --j) r = (1 << (m - j)) - (1 - 2 * (int)((code >> (m - 1 - j)) & 1u)) * r;
    return (1 - 2 * (int)((code >> (m - 1)) & 1u)) * r;
}

__host__ __device__ __forceinline__ float qam_norm(const int Qm) {
    return Qm <= 2 ? 0.70710678118654752440f : Qm == 4 ? 0.31622776601683793320f
         : Qm == 6 ? 0.15430334996209191026f : 0.07669649888473704465f;
}

// bits (one per byte, b0 first) of one symbol -> (re, im)
__device__ __forceinline__ float2 map_symbol(const uint8_t *__restrict__ b, const int Qm) {
    const float norm = qam_norm(Qm);
    if (Qm == 1) {
        const float a = (b[0] & 1) ? -norm : norm;
        return make_float2(a, a);
    }
    const int m = Qm >> 1;
    uint32_t cr = 0, ci = 0;
    for (int j = 0; j < m; ++j) {
        cr = (cr << 1) | (b[2 * j] & 1u);
        ci = (ci << 1) | (b[2 * j + 1] & 1u);
    }
    return make_float2(__fmul_rn((float)pam_level(m, cr), norm), __fmul_rn((float)pam_level(m, ci), norm));
}

// LLRs (or hard bits) of the m bits carried by one dimension; out[j] belongs to c_j
template <int M>
__device__ __forceinline__ void pam_demod(const float x, const float inv_var, const float norm, const int method, float *out) {
    float d[1 << M];
#pragma unroll
    for (int l = 0; l < (1 << M); ++l) {
        const float diff = __fsub_rn(x, __fmul_rn((float)pam_level(M, (uint32_t)l), norm));
        d[l] = -__fmul_rn(__fmul_rn(diff, diff), inv_var);
    }
#pragma unroll
    for (int j = 0; j < M; ++j) {
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int l = 0; l < (1 << M); ++l) {
            if ((l >> (M - 1 - j)) & 1) m1 = fmaxf(m1, d[l]);
            else m0 = fmaxf(m0, d[l]);
        }
        float v;
        if (method == kDemodHard) {
            v = m1 > m0 ? 1.0f : 0.0f;
        } else if (method == kDemodApprox) {
            v = __fsub_rn(m0, m1);
        } else {
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
            for (int l = 0; l < (1 << M); ++l) {
                if ((l >> (M - 1 - j)) & 1) s1 += expf(d[l] - m1);
                else s0 += expf(d[l] - m0);
            }
            v = (m0 - m1) + (logf(s0) - logf(s1));
        }
        out[j] = v;
    }
}

// one received symbol -> Qm outputs in bit order b0..b_{Qm-1}
__device__ __forceinline__ void demod_symbol(const float2 r, const int Qm, const float inv_var, const int method, float *__restrict__ out) {
    const float norm = qam_norm(Qm);
    if (Qm <= 2) {
        const float gain = __fmul_rn(2.8284271247461900976f, inv_var);
        if (Qm == 1) {
            const float v = __fmul_rn(gain, __fadd_rn(r.x, r.y));
            out[0] = method == kDemodHard ? (v < 0.0f ? 1.0f : 0.0f) : v;
        } else {
            const float vx = __fmul_rn(gain, r.x), vy = __fmul_rn(gain, r.y);
            out[0] = method == kDemodHard ? (vx < 0.0f ? 1.0f : 0.0f) : vx;
            out[1] = method == kDemodHard ? (vy < 0.0f ? 1.0f : 0.0f) : vy;
        }
        return;
    }
    float lr[4], li[4];
    const int m = Qm >> 1;
    if (m == 2) { pam_demod<2>(r.x, inv_var, norm, method, lr); pam_demod<2>(r.y, inv_var, norm, method, li); }
    else if (m == 3) { pam_demod<3>(r.x, inv_var, norm, method, lr); pam_demod<3>(r.y, inv_var, norm, method, li); }
    else { pam_demod<4>(r.x, inv_var, norm, method, lr); pam_demod<4>(r.y, inv_var, norm, method, li); }
    for (int j = 0; j < m; ++j) {
        out[2 * j] = lr[j];
        out[2 * j + 1] = li[j];
    }
}

// complex AWGN for the symbol pair `pair` (symbols 2*pair, 2*pair+1): same keying as qpsk_awgn_llr_kernel
__device__ __forceinline__ void pair_noise(const long long pair, const uint64_t seed, const uint64_t stream_id, float n[4]) {
    uint32_t r[4];
    philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32),
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    box_muller(r[0], r[1], n[0], n[1]);
    box_muller(r[2], r[3], n[2], n[3]);
}

// NRModulator.step (NRModulator.m:87-89)
__global__ void __launch_bounds__(256) modulate_kernel(const uint8_t *__restrict__ bits, float2 *__restrict__ sym,
                                                       long long n_sym, int Qm) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_sym; i += (long long)gridDim.x * blockDim.x)
        sym[i] = map_symbol(bits + i * Qm, Qm);
}

// comm.AWGNChannel in SNR mode with unit signal power (plot_BLER_vs_SNR.m:50,105): sigma = sqrt(variance/2) per dimension
__global__ void __launch_bounds__(256) awgn_kernel(float2 *__restrict__ sym, long long n_sym, float sigma, uint64_t seed,
                                                   uint64_t stream_id) {
    const long long n_pairs = (n_sym + 1) >> 1;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_pairs; p += (long long)gridDim.x * blockDim.x) {
        float n[4];
        pair_noise(p, seed, stream_id, n);
        float2 a = sym[2 * p];
        a.x = __fadd_rn(a.x, __fmul_rn(sigma, n[0]));
        a.y = __fadd_rn(a.y, __fmul_rn(sigma, n[1]));
        sym[2 * p] = a;
        if (2 * p + 1 < n_sym) {
            float2 b = sym[2 * p + 1];
            b.x = __fadd_rn(b.x, __fmul_rn(sigma, n[2]));
            b.y = __fadd_rn(b.y, __fmul_rn(sigma, n[3]));
            sym[2 * p + 1] = b;
        }
    }
}

// NRDemodulator.step (NRDemodulator.m:90-92)
__global__ void __launch_bounds__(256) demodulate_kernel(const float2 *__restrict__ sym, float *__restrict__ out,
                                                         long long n_sym, int Qm, float inv_var, int method) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_sym; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
        demod_symbol(sym[i], Qm, inv_var, method, v);
        for (int j = 0; j < Qm; ++j) out[i * Qm + j] = v[j];
    }
}

// the three stages fused (no symbols in HBM): bits -> LLRs, bit-identical to modulate + awgn + demodulate
__global__ void __launch_bounds__(256) mod_awgn_demod_kernel(const uint8_t *__restrict__ bits, float *__restrict__ out,
                                                             long long n_sym, int Qm, float sigma, float inv_var, int method,
                                                             uint64_t seed, uint64_t stream_id) {
    const long long n_pairs = (n_sym + 1) >> 1;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_pairs; p += (long long)gridDim.x * blockDim.x) {
        float n[4];
        pair_noise(p, seed, stream_id, n);
        for (int h = 0; h < 2; ++h) {
            const long long i = 2 * p + h;
            if (i >= n_sym) break;
            float2 a = map_symbol(bits + i * Qm, Qm);
            a.x = __fadd_rn(a.x, __fmul_rn(sigma, n[2 * h]));
            a.y = __fadd_rn(a.y, __fmul_rn(sigma, n[2 * h + 1]));
            float v[8];
            demod_symbol(a, Qm, inv_var, method, v);
            for (int j = 0; j < Qm; ++j) out[i * Qm + j] = v[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// CRC attach / check for batches of blocks: comm.CRCGenerator / comm.CRCDetector with the polynomials of
// get_3gpp_crc_polynomial.m:3-17 (zero initial state, no reflection, no final XOR; bits MSB-first through an L-bit
// LFSR).  parity != NULL: the L parity bits of bits[0..n_bits)
// are written to parity (NRLDPCEncoder.m:80,114).  ok != NULL: 1 iff the remainder of all n_bits is zero,
// i.e. a block that already carries its parity passes (NRLDPCDecoder.m:300,336).
//
// One WARP per block: lane i runs the LFSR over its own contiguous chunk of ceil(n_bits/32) bits (chunks are aligned to
// the END of the block, so leading lanes may hold short or empty chunks), which yields r_i = chunk_i(x) * x^L mod P;
// the block's CRC is sum_i r_i * g^(31-i) mod P with g = x^chunk mod P, formed by a five-level Horner tree over the
// lanes (shuffles + carry-less modular multiplications).  The serial one-thread-per-block version it replaces took
// 675 us per call on 4096 blocks of 8448 bits (19 % of a BLER-loop batch); the CRC values are identical.
__device__ __forceinline__ uint32_t crc_step_bit(uint32_t reg, uint32_t bit, uint32_t poly, uint32_t mask, int L) {
    const uint32_t fb = ((reg >> (L - 1)) ^ bit) & 1u;
    reg = (reg << 1) & mask;
    return fb ? reg ^ poly : reg;
}
// a(x) * b(x) mod P over GF(2), operands and result below x^L
__device__ __forceinline__ uint32_t crc_mulmod(uint32_t a, uint32_t b, uint32_t poly, uint32_t mask, int L) {
    uint32_t res = 0;
    for (int k = L - 1; k >= 0; --k) {
        const uint32_t top = (res >> (L - 1)) & 1u;
        res = (res << 1) & mask;
        if (top) res ^= poly;
        if ((b >> k) & 1u) res ^= a;
    }
    return res;
}

// Round 2: the lane's chunk is consumed EIGHT bits per step -- word loads of four one-bit bytes are packed into a nibble
// by one multiplication ((w * 0x08040201) >> 24: bytes b0 b1 b2 b3 -> b0 b1 b2 b3 as a 4-bit number, MSB first), two nibbles
// index a 256-entry table of the LFSR's 8-step response (shared memory, built per CTA) -- about 1.6 instructions per bit instead
// of 6; heads / tails of a chunk that are not word aligned take the bit-serial step.  Identical CRC values.
__global__ void __launch_bounds__(128) crc_kernel(const uint8_t *__restrict__ bits, long long batch, int n_bits, long long stride,
                                                  uint32_t poly, int L, uint8_t *__restrict__ parity, long long parity_stride,
                                                  uint8_t *__restrict__ ok) {
    __shared__ uint32_t tab[256];
    const uint32_t mask = L == 32 ? 0xffffffffu : ((1u << L) - 1u);
    for (int t = threadIdx.x; t < 256; t += blockDim.x) {   // tab[t] = (t(x) * x^L) mod P: the register after shifting in byte t from zero
        uint32_t v = 0;
        for (int k = 7; k >= 0; --k) v = crc_step_bit(v, ((uint32_t)t >> k) & 1u, poly, mask, L);
        tab[t] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int chunk = (n_bits + 31) >> 5;                       // bits per lane
    // g = x^chunk mod P (the same for every block of the call)
    uint32_t g = 1u;
    for (int i = 0; i < chunk; ++i) g = crc_step_bit(g, 0u, poly, mask, L);
    for (long long b = warp; b < batch; b += n_warps) {
        const uint8_t *row = bits + b * stride;
        const int end = n_bits - (31 - lane) * chunk;           // this lane's chunk is [end - chunk, end) clipped at 0
        const int beg = end - chunk < 0 ? 0 : end - chunk;
        uint32_t v = 0;
        int i = beg;
        if ((reinterpret_cast<uintptr_t>(row) & 3) == 0) {
            for (; i < end && (i & 3); ++i) v = crc_step_bit(v, row[i], poly, mask, L);
            const uint32_t *w = reinterpret_cast<const uint32_t *>(row + i);
            for (; i + 8 <= end; i += 8, w += 2) {
                const uint32_t byte = ((((w[0] & 0x01010101u) * 0x08040201u) >> 24) << 4) | (((w[1] & 0x01010101u) * 0x08040201u) >> 24);
                v = ((v << 8) & mask) ^ tab[((v >> (L - 8)) ^ byte) & 0xffu];
            }
        }
        for (; i < end; ++i) v = crc_step_bit(v, row[i], poly, mask, L);
        // Horner tree: lane 31 ends up with sum_i r_i * g^(31 - i)
        uint32_t gp = g;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, v, d);
            if ((lane & (2 * d - 1)) == 2 * d - 1) v ^= crc_mulmod(up, gp, poly, mask, L);
            gp = crc_mulmod(gp, gp, poly, mask, L);
        }
        const uint32_t reg = __shfl_sync(0xffffffffu, v, 31);
        if (parity && lane < L) parity[b * parity_stride + lane] = (uint8_t)((reg >> (L - 1 - lane)) & 1u);
        if (ok && lane == 0) ok[b] = reg == 0 ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Uniform random bits, one per byte, for the information blocks of the Monte-Carlo loop (round(rand(A,1)),
// plot_BLER_vs_SNR.m:112): bits[r*stride + k], k < n_bits, from Philox4x32-10 keyed by (seed, stream id) with the counter =
// (row, index of the row's 128-bit group) -- independent of the launch geometry and of n_bits.  Nothing beyond n_bits of a row is touched (the
// CRC parity and the filler of a code block live there).  A thread owns 16 bits: one 16-byte store when the destination allows.
__global__ void __launch_bounds__(256) random_bits_kernel(uint8_t *__restrict__ bits, long long rows, int n_bits, long long stride,
                                                          uint64_t seed, uint64_t stream_id) {
    const int per_row = (n_bits + 15) >> 4;                     // 16-bit pieces per row
    const long long total = rows * per_row;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / per_row;
        const int piece = (int)(t - r * per_row);
        const long long ctr = (r << 20) | (long long)(piece >> 3);   // (row, 128-bit group of the row): a row's bits do not depend on n_bits (host: n_bits < 2^27)
        uint32_t x[4];
        philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), x);
        const uint32_t v = (x[(piece >> 1) & 3] >> ((piece & 1) * 16)) & 0xffffu;
        uint4 o;
        o.x = ((v & 0xfu) * 0x00204081u) & 0x01010101u;
        o.y = (((v >> 4) & 0xfu) * 0x00204081u) & 0x01010101u;
        o.z = (((v >> 8) & 0xfu) * 0x00204081u) & 0x01010101u;
        o.w = (((v >> 12) & 0xfu) * 0x00204081u) & 0x01010101u;
        uint8_t *dst = bits + r * stride + (long long)piece * 16;
        const int left = n_bits - piece * 16;
        if (left >= 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            *reinterpret_cast<uint4 *>(dst) = o;
        } else {
            const uint32_t wds[4] = {o.x, o.y, o.z, o.w};
            for (int k = 0; k < 16 && k < left; ++k) dst[k] = (uint8_t)((wds[k >> 2] >> ((k & 3) * 8)) & 1u);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// QPSK + AWGN + exact LLR fused into rate recovery ("fuse demap -> rate-recover"): the E received LLRs of
// a code block never go to HBM.  A CTA reads the block's E rate-matched BITS (coalesced uchar4 loads), forms the LLRs in
// shared memory with exactly the arithmetic and the Philox counters of qpsk_awgn_llr_kernel (counter = global bit
// quadruple index of the [batch][E] tensor), then gathers them into the decoder's input layout exactly as
// rate_recover_tma_kernel does: bit-identical to the two kernels in sequence, 0.83 GB of HBM traffic less per 4096
// headline blocks.  Requires E % 4 == 0 (the quadruples of a row are then the row's own).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) qpsk_awgn_rate_recover_kernel(const uint8_t *__restrict__ f_bits, float *__restrict__ harq,
                                                                     float *__restrict__ out, long long batch, RmGeom g, uint32_t magic_eq,
                                                                     float sigma, float gain, uint64_t seed, uint64_t stream_id,
                                                                     const int *__restrict__ table) {
    extern __shared__ __align__(16) unsigned char qrr_smem[];
    float *e_f = reinterpret_cast<float *>(qrr_smem);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int EQ4 = g.E >> 2;
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const long long quad0 = b * EQ4;
        const uchar4 *bits4 = reinterpret_cast<const uchar4 *>(f_bits) + quad0;
        for (int q = tid; q < EQ4; q += nt) {
            const long long i = quad0 + q;
            uint32_t r[4];
            philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), r);
            float n[4];
            box_muller(r[0], r[1], n[0], n[1]);
            box_muller(r[2], r[3], n[2], n[3]);
            const uchar4 bq = bits4[q];
            const float amp = 0.70710678118654752440f;
            float4 o;
            o.x = gain * ((bq.x ? -amp : amp) + sigma * n[0]);
            o.y = gain * ((bq.y ? -amp : amp) + sigma * n[1]);
            o.z = gain * ((bq.z ? -amp : amp) + sigma * n[2]);
            o.w = gain * ((bq.w ? -amp : amp) + sigma * n[3]);
            reinterpret_cast<float4 *>(e_f)[q] = o;
        }
        __syncthreads();
        float4 *ob = reinterpret_cast<float4 *>(out + b * g.ncw);
        if (table != nullptr) {
            float *hrow = harq ? harq + b * g.N : nullptr;
            for (int q = tid; q < (g.ncw >> 2); q += nt) __stcs(ob + q, rr_gather4(g, table, q, e_f, hrow, magic_eq, 2));
        } else
        for (int q = tid; q < (g.ncw >> 2); q += nt) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int n = 4 * q + u - g.Z2;
                float r = 0.0f;
                if (n >= 0) {
                    if (n >= g.F0u && n < g.F1u) {
                        r = __int_as_float(0x7f800000);  // filler: known 0 (NRLDPCDecoder.m:264)
                    } else if (n < g.Ncb) {
                        int first = rm_rank(g, n) - g.rank_k0;
                        if (first < 0) first += g.Nnf;
                        float acc = 0.0f;
                        for (int k = first; k < g.E; k += g.Nnf) {
                            int i = (int)__umulhi((uint32_t)k, magic_eq);      // k / EQ, one below at most
                            int j = k - i * g.EQ;
                            if (j >= g.EQ) { j -= g.EQ; ++i; }
                            acc = __fadd_rn(acc, e_f[i + j * 2]);               // Q_m = 2; same addition order as :230
                        }
                        if (harq) {
                            float *hb = harq + b * g.N + n;
                            acc = __fadd_rn(acc, *hb);
                            *hb = acc;
                        }
                        r = acc;
                    }
                }
                v[u] = r;
            }
            __stcs(ob + q, make_float4(v[0], v[1], v[2], v[3]));
        }
        __syncthreads();   // e_f is rewritten for the next block
    }
}

// ------------------------------------------------------------------------------------------------
// Block-error bookkeeping of the BLER loop (plot_BLER_vs_SNR.m:139-155 with NRLDPCDecoder.m:296-309,336-339) for one
// decoding attempt of B transport blocks of C code blocks: one warp per transport block.
//   ok      = TB CRC passed && every code block's CRC has passed (latched) && the A payload bits equal the sent ones
//             (a_hat is [] unless the CRCs pass; a block error is ~isequal(a, a_hat))
//   latch  |= ok                              (a block decoded by an earlier transmission stays decoded)
//   counters[3] += iterations of the C decodes of this attempt
//   finalize: counters[0] += 1, counters[1] += !latch, counters[2] += wrong bits among the first K' decoded bits of the C
//             blocks of a block in error
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bler_count_kernel(const uint8_t *__restrict__ hard, const uint8_t *__restrict__ info,
                                                         const uint8_t *__restrict__ tb_hat, long long tb_hat_stride,
                                                         const uint8_t *__restrict__ tb, long long tb_stride,
                                                         const uint8_t *__restrict__ tb_ok, const uint8_t *__restrict__ cb_passed,
                                                         const int32_t *__restrict__ iters, long long B, int C, int K, int Kp, int A,
                                                         uint8_t *__restrict__ latch, unsigned long long *__restrict__ counters,
                                                         int do_latch, int finalize) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long n_blocks = 0, n_err = 0, n_bit = 0, n_it = 0;
    for (long long b = warp; b < B; b += n_warps) {
        int latched = latch[b];
        if (do_latch) {
            int ok = tb_ok[b] != 0;
            if (cb_passed) for (int r = 0; r < C; ++r) ok &= cb_passed[b * C + r] != 0;
            const uint8_t *x = tb_hat + b * tb_hat_stride, *y = tb + b * tb_stride;
            uint32_t diff = 0;
            int i0 = 0;
            if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {   // 16 bytes per lane and load
                const int A16 = A & ~15;
                for (int i = lane * 16; i < A16; i += 512) {
                    const uint4 u = *reinterpret_cast<const uint4 *>(x + i), w = *reinterpret_cast<const uint4 *>(y + i);
                    diff |= (u.x ^ w.x) | (u.y ^ w.y) | (u.z ^ w.z) | (u.w ^ w.w);
                }
                i0 = A16;
            }
            for (int i = i0 + lane; i < A; i += 32) diff |= (uint32_t)(x[i] ^ y[i]);
            ok &= !__any_sync(0xffffffffu, diff != 0);
            latched |= ok;
            if (lane == 0) {
                latch[b] = (uint8_t)latched;
                for (int r = 0; r < C; ++r) n_it += (unsigned long long)iters[b * C + r];
            }
        }
        if (finalize) {
            if (lane == 0) { n_blocks += 1; n_err += latched ? 0 : 1; }
            if (!latched) {
                int wrong = 0;
                for (int r = 0; r < C; ++r) {
                    const uint8_t *x = hard + (b * C + r) * K, *y = info + (b * C + r) * K;
                    for (int i = lane; i < Kp; i += 32) wrong += (x[i] ^ y[i]) & 1;
                }
                wrong = __reduce_add_sync(0xffffffffu, wrong);
                if (lane == 0) n_bit += (unsigned long long)wrong;
            }
        }
    }
    if (lane == 0) {
        if (n_blocks) atomicAdd(counters + 0, n_blocks);
        if (n_err) atomicAdd(counters + 1, n_err);
        if (n_bit) atomicAdd(counters + 2, n_bit);
        if (n_it) atomicAdd(counters + 3, n_it);
    }
}

}  // namespace nrldpc
