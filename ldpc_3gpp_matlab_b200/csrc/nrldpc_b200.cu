// nrldpc_b200.cu -- host side of the C ABI declared in include/nrldpc_b200.h.
// Handle lifetime, table lifting, argument validation, launch configuration and the pipelined
// host<->device path.  No CPU compute fallback exists: every entry point either launches the
// sm_100a kernels or returns an error.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/nrldpc_b200.h"
#include "bg_tables.inc"
#include "chain_kernels.cuh"
#include "decode_kernel.cuh"
#include "decode_kernel_h2.cuh"
#include "decode_kernel_refill.cuh"
#include "decode_kernel_shfl.cuh"
#include "decode_kernel_bp.cuh"
#include "host_staging.h"

#define NRLDPC_EXPORT extern "C" __attribute__((visibility("default")))

namespace nrldpc {
// fp16 LLR transport (nrldpc_decode16): widen to the float32 layout the decode kernels read (exact conversion)
__global__ void __launch_bounds__(256) widen_f16_kernel(const uint2 *__restrict__ in, float4 *__restrict__ out, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const uint2 v = __ldcs(in + i);
        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&v.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
        out[i] = make_float4(a.x, a.y, b.x, b.y);
    }
}

// 8-bit LLR transport (nrldpc_decode8): llr = scale * q, q = 127 marks a filler / known-zero position (+inf)
__global__ void __launch_bounds__(256) widen_i8_kernel(const uint32_t *__restrict__ in, float4 *__restrict__ out, long long n4, float scale) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t v = __ldcs(in + i);
        float f[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int q = (int)(signed char)((v >> (8 * k)) & 0xffu);
            f[k] = q == 127 ? __int_as_float(0x7f800000) : __fmul_rn(scale, (float)q);
        }
        out[i] = make_float4(f[0], f[1], f[2], f[3]);
    }
}

// Shared-window address at which a kernel without static shared memory sees its dynamic shared memory.
__global__ void smem_base_probe(uint32_t *out) {
    extern __shared__ __align__(16) unsigned char probe_smem[];
    *out = (uint32_t)__cvta_generic_to_shared(probe_smem);
}
}  // namespace nrldpc

namespace {

constexpr int kNumPipe = 3;  // streams / staging sets for NRLDPC_MEM_HOST calls
enum { kInF32 = 0, kInF16 = 1, kInF64 = 2, kInI8 = 3 };   // element type of the LLR buffer handed to decode_impl

thread_local char g_create_error[256] = "";

struct PipeSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    float *llr = nullptr;       // device staging
    uint16_t *llr16 = nullptr;  // fp16 transport staging (nrldpc_decode16)
    double *llr64 = nullptr;    // float64 staging (nrldpc_decode64)
    double *soft64 = nullptr;
    double *rmsg = nullptr;     // sum-product check-to-variable messages (NRLDPC_ALG_BP)
    size_t rmsg_cap = 0;
    uint8_t *hard = nullptr;
    float *soft = nullptr;
    int32_t *iters = nullptr;
    uint8_t *ok = nullptr;
    uint8_t *bytes_in = nullptr, *bytes_out = nullptr;  // encode / rate-match staging
    float *f_in = nullptr;
    size_t cap_cw = 0;          // capacity in codewords of the decode staging
    size_t cap_bytes = 0;       // capacity of the generic staging buffers
    uint32_t *c2v = nullptr;    // decode scratch (one set per slot: kernels of different slots may overlap)
    int *counter = nullptr;     // ticket counter of the decode kernels (zeroed once; see DecArgs::work_base)
    unsigned int tickets = 0;   // value the min-sum launches enqueued so far will leave in `counter`
    bool counter_dirty = false; // a sum-product launch or a failed launch used the counter: zero it before the next min-sum launch
    size_t scratch_recs = 0;
    // pinned host ring of the staged host path (pageable and / or float64 caller buffers, see host_staging.h)
    unsigned char *h_in = nullptr;
    size_t h_in_cap = 0;
    uint8_t *h_hard = nullptr, *h_ok = nullptr;
    int32_t *h_iters = nullptr;
    size_t h_out_cw = 0;
    cudaEvent_t h2d_done = nullptr;
    int64_t pend_off = 0, pend_n = 0;   // chunk whose outputs still sit in the pinned buffers
};

}  // namespace

struct nrldpc_handle {
    nrldpc_cfg cfg;
    nrldpc_dims d;
    int device = 0;
    int num_sms = 0;
    char err[256];
    int64_t launches = 0;
    // device tables
    uint32_t *edesc = nullptr;
    int *row_start = nullptr;
    int h_row_start[48];
    int enc_s0[4];
    int enc_delta = 0;
    PipeSlot pipe[kNumPipe];
    int dec_variant = 1;             // NRLDPC_DECODE_VARIANT=loop selects the generic looped kernel, =shfl the lane-per-edge warp-shuffle kernel (Z <= 32)
    int shfl_cwpc = 0, shfl_threads = 256;   // NRLDPC_SHFL_CWPC / NRLDPC_SHFL_THREADS (0: 32 / Z codewords per CTA)
    int bitsliced_min_rows = 4;      // syndrome variants, see DecArgs (NRLDPC_BITSLICED_MIN_ROWS / NRLDPC_STAGED_MIN_ROWS: experiments)
    int staged_min_rows = 8;
    int l2_pin = 1;                  // NRLDPC_L2_PIN=0 drops the evict_last policy on the c2v scratch
    uint32_t smem_base = 0;          // shared-window address of dynamic shared memory (probed at create)
    nrldpc::DecArgs dec_args;
    cudaEvent_t dev_done = nullptr;  // last NRLDPC_MEM_DEVICE launch that used pipe[0]'s scratch
    float *dev_widen = nullptr;      // float32 copy of device fp16 / fp64 input (nrldpc_decode16/64, NRLDPC_MEM_DEVICE)
    size_t dev_widen_cw = 0;
    // sum-product kernel tables (decode_kernel_bp.cuh)
    int *bp_shift = nullptr, *bp_colz = nullptr, *bp_col_start = nullptr, *bp_col_edge = nullptr;
    // launch geometry looked up once per (kernel, CTA width, shared-memory size): the reference calls step() with ONE
    // codeword (NRLDPCDecoder.m:265), so no launcher may pay a runtime query per call
    struct OccEntry { const void *kern; int threads; size_t smem; int occ; };
    std::vector<OccEntry> occ_cache;
    int cwpc = 1;                    // codewords (pairs) per decode CTA, chosen at create (choose_decode_cwpc)
    int cwpc_override = 0;           // NRLDPC_CWPC (experiments)
    int occ_cap = 4;                 // resident decode CTAs per SM are capped here (NRLDPC_OCC_CAP)
    int shape_model = 1;             // NRLDPC_SHAPE_MODEL=0: the default rule only (ignore the measured shape table)
    int occ_cap_forced = 0;
    int grid_cap = 0;                // NRLDPC_GRID_CAP (experiments), read once at create
    int no_tma = 0;                  // NRLDPC_NO_TMA
    cudaStream_t last_dev_stream = nullptr;  // stream of the last NRLDPC_MEM_DEVICE decode (its scratch is shared)
    bool dev_used = false;
    nrldpc::HostPool *pool = nullptr;        // worker threads of the staged host path (created on first use)
    int host_threads = 0;
    long long l2_window = 0;                 // bytes of the persisting access-policy window over the c2v scratch (0: none; NRLDPC_L2_WINDOW=0/1)
    int zero_copy_max = 2;                   // host-memory decodes of up to this many codewords read / write pinned host memory directly (NRLDPC_ZERO_COPY_MAX, 0 = off)
    int refill = 1;                          // NRLDPC_REFILL: 0 = never refill slots, 1 = where measured to pay (default), 2 = refill kernel whenever possible,
                                             // 3 = prefetched refill (decode_kernel_refill.cuh) whenever possible
    struct RrTable { nrldpc::RmGeom g; int *d; };
    std::vector<RrTable> rr_tables;          // gather tables of rate recovery, one per geometry seen (chain_kernels.cuh)
    int rr_table_on = 1;                     // NRLDPC_RR_TABLE=0: evaluate the gather map per code block (the round-1 kernels)
    float i8_scale = 1.0f;                   // scale of the current nrldpc_decode8 call
    int refill_spares = 2;                   // NRLDPC_REFILL_SPARES: mailboxes per CTA of the prefetched-refill kernel
    int no_staging = 0;                      // NRLDPC_NO_STAGING=1: hand pageable buffers to cudaMemcpyAsync as they are (A/B)
    int bp_threads = 1024;           // CTA width of the sum-product kernel (NRLDPC_BP_THREADS=512 selects the 128-register build)
};

namespace {

int fail(nrldpc_handle *h, int code, const char *fmt, ...) {
    char *dst = h ? h->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 256, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(h, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(h, e_ == cudaErrorMemoryAllocation ? NRLDPC_ENOMEM : NRLDPC_ECUDA,        \
                        "%s: %s", #expr, cudaGetErrorString(e_));                                 \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the kernel FUNCTION (per device), not to a handle: two live
// handles that share an instantiation (decode_nms_kernel<1,false> for Z = 96 and Z = 8, encode_kernel for any Z) must
// not lower it under each other.  The attribute is only ever RAISED, process-wide, to the largest size any handle asked for.
int raise_max_smem(nrldpc_handle *h, const void *kern, size_t smem) {
    struct Key { int device; const void *kern; size_t smem; };
    static std::mutex mu;
    static std::vector<Key> seen;
    std::lock_guard<std::mutex> lock(mu);
    for (auto &k : seen)
        if (k.device == h->device && k.kern == kern) {
            if (k.smem >= smem) return 0;
            CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k.smem = smem;
            return 0;
        }
    CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    seen.push_back(Key{h->device, kern, smem});
    return 0;
}

// resident CTAs per SM of `kern` at this geometry (cached per handle), after making sure the launch is allowed
int cached_occupancy(nrldpc_handle *h, const void *kern, int threads, size_t smem, int *occ) {
    if (int rc = raise_max_smem(h, kern, smem)) return rc;
    for (const auto &e : h->occ_cache)
        if (e.kern == kern && e.threads == threads && e.smem == smem) { *occ = e.occ; return 0; }
    int q = 0;
    CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, threads, smem));
    q = std::max(1, q);
    h->occ_cache.push_back({kern, threads, smem, q});
    *occ = q;
    return 0;
}

// Entry points run on the handle's device and leave the caller's current device as they found it.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t enter(int dev) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) return e;
        if (prev == dev) return cudaSuccess;
        switched = true;
        return cudaSetDevice(dev);
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
#define ENTER_DEVICE(h) DeviceGuard guard_; CUDA_TRY(h, guard_.enter((h)->device))

const int kSetA[8] = {2, 3, 5, 7, 9, 11, 13, 15};
const int kSetN[8] = {8, 8, 7, 6, 6, 6, 5, 5};

struct BgView {
    int rows, cols, kcols, edges;
    const unsigned char *row, *col;
    const unsigned short (*shift)[NRLDPC_BG1_EDGES];
    const unsigned short (*shift2)[NRLDPC_BG2_EDGES];
    unsigned short sh(int ils, int e) const { return shift ? shift[ils][e] : shift2[ils][e]; }
};

BgView bg_view(int bg) {
    BgView v{};
    if (bg == 1) {
        v.rows = 46; v.cols = 68; v.kcols = 22; v.edges = NRLDPC_BG1_EDGES;
        v.row = nrldpc_bg1_row; v.col = nrldpc_bg1_col; v.shift = nrldpc_bg1_shift; v.shift2 = nullptr;
    } else {
        v.rows = 42; v.cols = 52; v.kcols = 10; v.edges = NRLDPC_BG2_EDGES;
        v.row = nrldpc_bg2_row; v.col = nrldpc_bg2_col; v.shift = nullptr; v.shift2 = nrldpc_bg2_shift;
    }
    return v;
}

// Words between the APP arrays of two codewords (float32) / codeword pairs (packed half) that share a CTA.
// Lane (slot, z) of a warp addresses slot*stride + col*Z + (z + shift) mod Z; with stride = cols*Z the slots of a warp
// collide in the shared-memory banks (BG1, Z = 8: stride 544 = 0 mod 32, four slots per warp -> 4-way conflicts; ncu:
// shared-memory wavefronts at 74 % of peak, ALU pipe at 59 %).  With stride = Z (mod 32) the banks of a warp follow the
// lane index through every circulant rotation, so whole slots never collide.
int decode_slot_stride(int cols, int Z, int cwpc) {
    const int ncw = cols * Z;
    if (cwpc == 1) return ncw;   // one codeword per CTA
    return ncw + (((Z - ncw) % 32) + 32) % 32;
}
int decode_threads_for(int cwpc, int Z) { return std::max(32, (cwpc * Z + 31) / 32 * 32); }

size_t decode_smem_for(const nrldpc_dims &d, int cwpc) {
    // APP arrays, syndrome flags of both stages (float32: [2][cwpc], packed half: [2][2*cwpc]) + work slot, mbarrier, and
    // (one codeword per CTA, Z a multiple of 32) the packed hard decisions of the bit-sliced syndrome -- two words per
    // (column, warp) cover the float32 and the packed-half kernel -- followed by its lane-indexed edge table
    const size_t hb = (cwpc == 1 && d.Z % 32 == 0) ? (size_t)d.cols * (d.Z / 32) * 8 + (size_t)d.edges * 4 : 0;
    return (size_t)cwpc * decode_slot_stride(d.cols, d.Z, cwpc) * 4 + (size_t)(4 * cwpc + 1) * 4 + 16 + hb;
}

// CTA shape of the decode kernels: codewords (float32) or codeword pairs (packed half) per CTA and the cap on resident
// CTAs per SM.  Default rule: as many codewords as fit in 384 threads (and in the shared memory of two CTAs per SM), at
// most 4 CTAs per SM.  Where the B200 scans (tools/gpu_shape_scan.py -> decode_shapes.inc; one with fixed iterations, one with
// the parity-check stop, whose best shapes differ) measured another shape at least 2-3 % faster, that shape is used.  What the scan showed (profiles/r02_shape_scan_*.jsonl):
//   * warps are bound to the SM's four schedulers by (warp index in the CTA) mod 4: Z = 224 (three 7-warp CTAs per SM)
//     runs at exactly 7/8 of the Z = 384 rate per lane, so CTAs of 4k warps are preferred when lanes are not wasted;
//   * many narrow one-codeword CTAs per SM lose on base graph 1 (BG1 Z = 128: 6 x 4 warps 8.4 Gb/s against 2 x 12 warps
//     10.5): the CTAs de-phase and each needs its own pass through the ~100 KB unrolled layer loop (instruction
//     fetch); on base graph 2 (66 KB loop) the same shape wins (Z = 128: 13.2 against 11.8 Gb/s).
// Every shape is bit-identical (the scan checks it; tests/test_gpu_parity.py runs forced shapes).
#include "decode_shapes.inc"
int lifting_index(int Z) {
    int idx = 0;
    for (int z = 2; z <= 384; ++z) {
        bool valid = false;
        for (int s = 0; s < 8 && !valid; ++s)
            for (int j = 0; j < kSetN[s]; ++j) valid = valid || (kSetA[s] << j) == z;
        if (!valid) continue;
        if (z == Z) return idx;
        ++idx;
    }
    return -1;
}
void choose_decode_shape(nrldpc_handle *h) {
    const nrldpc_dims &d = h->d;
    const int cmax = std::max(1, nrldpc::kDecThreads / d.Z);
    const int legacy = std::min(cmax, std::max(1, (216 * 1024 / nrldpc::kDecCtasPerSm / 4) / decode_slot_stride(d.cols, d.Z, 2)));
    h->cwpc = legacy;
    if (h->shape_model) {
        const int zi = lifting_index(d.Z);
        const unsigned char *e = nrldpc_decode_shape[h->cfg.early_term ? 1 : 0][h->cfg.llr_dtype == NRLDPC_F16X2 ? 1 : 0][d.bg - 1][zi < 0 ? 0 : zi];
        if (zi >= 0 && e[0] > 0 && e[0] <= cmax && decode_smem_for(d, e[0]) <= 227 * 1024) {
            h->cwpc = e[0];
            if (!h->occ_cap_forced) h->occ_cap = e[1];
        }
    }
    if (h->cwpc_override > 0) h->cwpc = std::min(h->cwpc_override, cmax);
}

int ensure_scratch(nrldpc_handle *h, PipeSlot &s, size_t recs) {
    if (!s.counter) {
        CUDA_TRY(h, cudaMalloc(&s.counter, sizeof(int)));
        CUDA_TRY(h, cudaMemset(s.counter, 0, sizeof(int)));
        s.tickets = 0; s.counter_dirty = false;
    }
    if (recs <= s.scratch_recs) return 0;
    if (s.c2v) cudaFree(s.c2v);
    s.c2v = nullptr; s.scratch_recs = 0;
    CUDA_TRY(h, cudaMalloc(&s.c2v, recs * sizeof(uint32_t)));
    s.scratch_recs = recs;
    return 0;
}

int launch_decode_shfl(nrldpc_handle *h, cudaStream_t stream, const float *llr, int64_t batch, int n_rows, uint8_t *hard, float *soft,
                       int32_t *iters, uint8_t *ok);

// Enqueue one decode launch on `stream` using slot `s`'s scratch.
int launch_decode(nrldpc_handle *h, PipeSlot &s, cudaStream_t stream, const float *llr, int64_t batch,
                  int n_rows, uint8_t *hard, float *soft, int32_t *iters, uint8_t *ok) {
    const int Z = h->d.Z;
    const bool h2 = h->cfg.llr_dtype == NRLDPC_F16X2;
    if (h->dec_variant == 2 && Z <= 32 && !h2) return launch_decode_shfl(h, stream, llr, batch, n_rows, hard, soft, iters, ok);
    // cwpc: codewords (float32) or codeword pairs (packed half) resident per CTA
    // one codeword (pair) per CTA: exactly Z threads, also when that leaves the last warp partly filled
    const int cwpc = h->cwpc, threads = cwpc == 1 ? Z : decode_threads_for(cwpc, Z);
    const int per_group = h2 ? 2 * cwpc : cwpc;
    const int64_t n_groups = (batch + per_group - 1) / per_group;
    size_t smem = decode_smem_for(h->d, cwpc);
    (void)n_rows;
    // variant 0: generic looped layers (float32 only); otherwise the layer loop is unrolled for the base graph.
    // FULL: one codeword (pair) per CTA, CTA-uniform code; MASKED: the same with a partially filled last warp (Z not a
    // multiple of 32: no warp-ballot syndrome)
    using Kern = void (*)(const nrldpc::DecArgs);
    const bool full = cwpc == 1, masked = full && (Z % 32) != 0;
    const bool bg1 = h->d.bg == 1;
    Kern kern;
    if (h2)
        kern = bg1 ? (masked ? (Kern)nrldpc::decode_nms_h2_kernel<1, true, true> : full ? (Kern)nrldpc::decode_nms_h2_kernel<1, true> : (Kern)nrldpc::decode_nms_h2_kernel<1, false>)
                   : (masked ? (Kern)nrldpc::decode_nms_h2_kernel<2, true, true> : full ? (Kern)nrldpc::decode_nms_h2_kernel<2, true> : (Kern)nrldpc::decode_nms_h2_kernel<2, false>);
    else if (h->dec_variant == 0)
        kern = (Kern)nrldpc::decode_nms_kernel<0, false, 0>;
    else {
        // multi-codeword CTAs under the parity-check stop track the parity variables' hard decisions in registers (mode 2)
        const bool track = !full && h->cfg.early_term;
        kern = bg1 ? (masked ? (Kern)nrldpc::decode_nms_kernel<1, true, 1> : full ? (Kern)nrldpc::decode_nms_kernel<1, true, 0>
                      : track ? (Kern)nrldpc::decode_nms_kernel<1, false, 2> : (Kern)nrldpc::decode_nms_kernel<1, false, 0>)
                   : (masked ? (Kern)nrldpc::decode_nms_kernel<2, true, 1> : full ? (Kern)nrldpc::decode_nms_kernel<2, true, 0>
                      : track ? (Kern)nrldpc::decode_nms_kernel<2, false, 2> : (Kern)nrldpc::decode_nms_kernel<2, false, 0>);
    }
    // multi-codeword CTAs under the parity-check stop: every slot is refilled on its own (decode_nms_refill_kernel)
    // (measured, profiles/r02_v4_refill_ab.txt: pays from about 24 codewords per CTA on -- Z <= 16 -- where a group waits long for
    // its slowest member; with fewer slots the pass a refilled slot sits out costs more than the idling it removes)
    const bool refill = !h2 && h->dec_variant != 0 && h->cfg.early_term && cwpc > 1 && batch < ((int64_t)1 << 31) - 65536 &&
                        (h->refill == 1 ? cwpc >= 24 : h->refill == 2);
    if (refill) kern = bg1 ? (Kern)nrldpc::decode_nms_refill_kernel<1> : (Kern)nrldpc::decode_nms_refill_kernel<2>;
    // prefetched refill: S slots + P mailboxes per CTA (bulk copies need 16-byte aligned buffers: Z a multiple of 4)
    const int stride_w = decode_slot_stride(h->d.cols, Z, cwpc);
    const int spares = h->refill_spares;
    const size_t smem2 = (size_t)(cwpc + spares) * stride_w * 4 + (size_t)(4 * cwpc + 3 * spares + 4) * 4 + 8 + (size_t)(1 + spares) * 8;
    const bool refill2 = !h2 && h->dec_variant != 0 && h->cfg.early_term && cwpc > 1 && batch < ((int64_t)1 << 31) - 65536 &&
                         (stride_w & 3) == 0 && smem2 <= 227 * 1024 && h->refill == 3;
    if (refill2) {
        kern = bg1 ? (Kern)nrldpc::decode_nms_refill2_kernel<1> : (Kern)nrldpc::decode_nms_refill2_kernel<2>;
        smem = smem2;
        s.counter_dirty = true;   // this kernel counts from zero (it draws an unpredictable number of tickets)
    }
    // persistent grid: every SM filled to its occupancy (2 CTAs of 384 threads at Z = 384, more for narrower CTAs);
    // attribute and occupancy are looked up once per (kernel, CTA width, shared-memory size)
    int occ = 1;
    if (int rc = cached_occupancy(h, reinterpret_cast<const void *>(kern), threads, smem, &occ)) return rc;
    occ = std::min(occ, h->occ_cap);
    int grid = (int)std::min<int64_t>(n_groups, (int64_t)h->num_sms * occ);
    if (h->grid_cap > 0) grid = std::max(1, std::min(grid, h->grid_cap));  // experiments only
    // c2v scratch: [kRecWords][kRecStride = 384 thread columns] blocks; narrow CTAs share a block side by side, so the
    // scratch stays about (threads resident on the device) x 145 words whatever the CTA width (it has to fit the L2)
    const int rec_group = std::max(1, nrldpc::kRecStride / threads);
    if (int rc = ensure_scratch(h, s, (size_t)((grid + rec_group - 1) / rec_group) * nrldpc::kRecWords * nrldpc::kRecStride)) return rc;
    // a launch that is being captured into a CUDA graph must be replayable: it zeroes the counter itself (memset node) and
    // counts from zero; the host's running ticket count is then void and the next ordinary launch starts afresh
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
    const bool capturing = cap != cudaStreamCaptureStatusNone;
    if (s.counter_dirty || capturing) {
        CUDA_TRY(h, cudaMemsetAsync(s.counter, 0, sizeof(int), stream));
        s.tickets = 0; s.counter_dirty = capturing;
    }
    if (refill2) s.counter_dirty = true;   // the next launch zeroes the counter again
    nrldpc::DecArgs &a = h->dec_args;  // tables were filled at create()
    a.llr = llr; a.hard = hard; a.soft = soft; a.iters = iters; a.ok = ok;
    a.batch = batch; a.Z = Z; a.ncols = h->d.cols; a.kcols = h->d.kcols; a.n_rows = n_rows;
    a.n_edges = h->h_row_start[n_rows]; a.max_iters = h->cfg.max_iters; a.early_term = h->cfg.early_term;
    a.slot_stride = decode_slot_stride(h->d.cols, Z, cwpc);
    a.cwpc = cwpc; a.spares = spares; a.rec_group = rec_group; a.alpha = h->cfg.alpha; a.l2_pin = h->l2_pin; a.one = 1;
    a.bitsliced_min_rows = h->bitsliced_min_rows; a.staged_min_rows = h->staged_min_rows;
    a.c2v = s.c2v; a.work_counter = s.counter; a.work_base = s.tickets;
    // tickets this launch draws: one per group (refill kernel: per codeword) plus the failing fetch that ends every CTA (slot)
    if (!capturing)
        s.tickets += refill ? (unsigned int)batch + (unsigned int)grid * (unsigned int)cwpc : (unsigned int)n_groups + (unsigned int)grid;
    const uint32_t ah = __half_as_ushort(__float2half_rn(h->cfg.alpha));
    a.alpha_h2 = ah | (ah << 16);
    if (a.smem_base != h->smem_base) {
        for (int e = 0; e < h->d.edges; ++e) a.ed[e].y += h->smem_base - a.smem_base;
        a.smem_base = h->smem_base;
    }
    if (h->l2_window > 0) {
        // the c2v scratch also gets a persisting access-policy window for this launch: lines of the window go to the L2's
        // persisting set-aside (sized at create), which the streamed LLR input cannot claim -- the evict_last hints
        // alone still let 3-5 % of the record stores leak out to HBM (profiles/r01_v7_dram_f16x2.csv)
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = s.c2v;
        attr[0].val.accessPolicyWindow.num_bytes = std::min<size_t>(s.scratch_recs * sizeof(uint32_t), (size_t)h->l2_window);
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr; cfg.numAttrs = 1;
        CUDA_TRY(h, cudaLaunchKernelEx(&cfg, kern, a));
    } else {
        kern<<<grid, threads, smem, stream>>>(a);
    }
    if (cudaError_t e_ = cudaGetLastError()) { s.counter_dirty = true; return fail(h, NRLDPC_ECUDA, "decode launch: %s", cudaGetErrorString(e_)); }
    h->launches += 1;
    return 0;
}

// Lane-per-edge warp-shuffle kernel (decode_kernel_shfl.cuh): experiment / small-Z latency path, float32, Z <= 32.
int launch_decode_shfl(nrldpc_handle *h, cudaStream_t stream, const float *llr, int64_t batch, int n_rows, uint8_t *hard, float *soft,
                       int32_t *iters, uint8_t *ok) {
    const int Z = h->d.Z, rows_all = h->d.cols - h->d.kcols;
    auto smem_for = [&](int c) {
        return (size_t)c * h->d.n_cw * 4 + (size_t)rows_all * c * Z * 12 + (size_t)(rows_all - 4) * c * Z * 4 + (size_t)h->d.edges * 4 + (size_t)c * 8 + 16;
    };
    int cwpc = h->shfl_cwpc > 0 ? h->shfl_cwpc : std::max(1, 32 / Z);
    cwpc = (int)std::max<int64_t>(1, std::min<int64_t>(cwpc, batch));
    while (cwpc > 1 && smem_for(cwpc) > 200 * 1024) --cwpc;
    const size_t smem = smem_for(cwpc);
    const int threads = h->shfl_threads;
    int occ = 1;
    if (int rc = cached_occupancy(h, reinterpret_cast<const void *>(nrldpc::decode_nms_shfl_kernel), threads, smem, &occ)) return rc;
    const int64_t n_groups = (batch + cwpc - 1) / cwpc;
    const int grid = (int)std::min<int64_t>(n_groups, (int64_t)h->num_sms * occ);
    nrldpc::DecArgs &a = h->dec_args;
    a.llr = llr; a.hard = hard; a.soft = soft; a.iters = iters; a.ok = ok;
    a.batch = batch; a.Z = Z; a.ncols = h->d.cols; a.kcols = h->d.kcols; a.n_rows = n_rows;
    a.n_edges = h->h_row_start[n_rows]; a.max_iters = h->cfg.max_iters; a.early_term = h->cfg.early_term;
    a.slot_stride = h->d.n_cw; a.cwpc = cwpc; a.alpha = h->cfg.alpha; a.one = 1;
    if (a.smem_base != h->smem_base) {
        for (int e = 0; e < h->d.edges; ++e) a.ed[e].y += h->smem_base - a.smem_base;
        a.smem_base = h->smem_base;
    }
    nrldpc::decode_nms_shfl_kernel<<<grid, threads, smem, stream>>>(a);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

// Enqueue one sum-product (reference algorithm) decode launch; T = element type of llr / soft.
template <typename T>
int launch_decode_bp(nrldpc_handle *h, PipeSlot &s, cudaStream_t stream, const T *llr, int64_t batch, int n_rows,
                     uint8_t *hard, T *soft, int32_t *iters, uint8_t *ok) {
    const int Z = h->d.Z;
    const size_t smem = (size_t)h->d.n_cw * sizeof(double);
    int width = h->bp_threads;
    const int threads = std::min(width, (n_rows * Z + 31) / 32 * 32);
    auto kern = width > 512 ? nrldpc::decode_bp_kernel<T, 1024> : nrldpc::decode_bp_kernel<T, 512>;
    int occ = 1;
    if (int rc = cached_occupancy(h, reinterpret_cast<const void *>(kern), threads, smem, &occ)) return rc;
    occ = std::min(occ, 4);
    const int grid = (int)std::min<int64_t>(batch, (int64_t)h->num_sms * occ);
    const size_t stride = (size_t)h->d.edges * Z;
    if (!s.counter) CUDA_TRY(h, cudaMalloc(&s.counter, sizeof(int)));
    s.counter_dirty = true;   // this kernel counts from zero (memset below)
    if ((size_t)grid * stride > s.rmsg_cap) {
        cudaFree(s.rmsg);
        s.rmsg = nullptr; s.rmsg_cap = 0;
        CUDA_TRY(h, cudaMalloc(&s.rmsg, (size_t)grid * stride * sizeof(double)));
        s.rmsg_cap = (size_t)grid * stride;
    }
    CUDA_TRY(h, cudaMemsetAsync(s.counter, 0, sizeof(int), stream));
    nrldpc::BpArgs a{};
    a.llr = llr; a.hard = hard; a.soft = soft; a.iters = iters; a.ok = ok; a.batch = batch;
    a.Z = Z; a.ncols = h->d.cols; a.kcols = h->d.kcols; a.n_rows = n_rows; a.n_edges = h->h_row_start[n_rows];
    a.max_iters = h->cfg.max_iters; a.early_term = h->cfg.early_term;
    a.rmsg = s.rmsg; a.rmsg_stride = (long long)stride; a.work_counter = s.counter;
    a.row_start = h->row_start; a.e_shift = h->bp_shift; a.e_colz = h->bp_colz;
    a.col_start = h->bp_col_start; a.col_edge = h->bp_col_edge;
    kern<<<grid, threads, smem, stream>>>(a);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

int ensure_pipe(nrldpc_handle *h) {
    for (auto &s : h->pipe) {
        if (!s.stream) CUDA_TRY(h, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        if (!s.done) CUDA_TRY(h, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        if (!s.h2d_done) CUDA_TRY(h, cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
    }
    return 0;
}

// true for ordinary (pageable, unregistered) host memory: the copy engine cannot read / write it directly, the driver
// would bounce it through its own small staging buffer on the calling thread
bool is_pageable(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

int ensure_host_ring(nrldpc_handle *h, PipeSlot &s, size_t in_bytes, size_t out_cw, bool want_iters, bool want_ok) {
    if (in_bytes > s.h_in_cap) {
        if (s.h_in) cudaFreeHost(s.h_in);
        s.h_in = nullptr; s.h_in_cap = 0;
        CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void **>(&s.h_in), in_bytes, cudaHostAllocDefault));
        s.h_in_cap = in_bytes;
    }
    if (out_cw > s.h_out_cw) {
        if (s.h_hard) cudaFreeHost(s.h_hard);
        if (s.h_ok) cudaFreeHost(s.h_ok);
        if (s.h_iters) cudaFreeHost(s.h_iters);
        s.h_hard = s.h_ok = nullptr; s.h_iters = nullptr; s.h_out_cw = 0;
        CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void **>(&s.h_hard), out_cw * h->d.K, cudaHostAllocDefault));
        CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void **>(&s.h_ok), out_cw, cudaHostAllocDefault));
        CUDA_TRY(h, cudaHostAlloc(reinterpret_cast<void **>(&s.h_iters), out_cw * sizeof(int32_t), cudaHostAllocDefault));
        s.h_out_cw = out_cw;
    }
    (void)want_iters; (void)want_ok;
    return 0;
}

int ensure_decode_staging(nrldpc_handle *h, PipeSlot &s, size_t cw, bool soft, int in_kind) {
    if (cw > s.cap_cw) {
        cudaFree(s.llr); cudaFree(s.hard); cudaFree(s.iters); cudaFree(s.ok); cudaFree(s.soft); cudaFree(s.llr16);
        cudaFree(s.llr64); cudaFree(s.soft64);
        s.llr = nullptr; s.hard = nullptr; s.iters = nullptr; s.ok = nullptr; s.soft = nullptr; s.llr16 = nullptr; s.cap_cw = 0;
        s.llr64 = nullptr; s.soft64 = nullptr;
        CUDA_TRY(h, cudaMalloc(&s.llr, cw * h->d.n_cw * sizeof(float)));
        CUDA_TRY(h, cudaMalloc(&s.hard, cw * h->d.K));
        CUDA_TRY(h, cudaMalloc(&s.iters, cw * sizeof(int32_t)));
        CUDA_TRY(h, cudaMalloc(&s.ok, cw));
        s.cap_cw = cw;
    }
    if (soft && in_kind != kInF64 && !s.soft) CUDA_TRY(h, cudaMalloc(&s.soft, s.cap_cw * h->d.n_cw * sizeof(float)));
    if ((in_kind == kInF16 || in_kind == kInI8) && !s.llr16) CUDA_TRY(h, cudaMalloc(&s.llr16, s.cap_cw * h->d.n_cw * sizeof(uint16_t)));
    if (in_kind == kInF64 && !s.llr64) CUDA_TRY(h, cudaMalloc(&s.llr64, s.cap_cw * h->d.n_cw * sizeof(double)));
    if (in_kind == kInF64 && soft && !s.soft64) CUDA_TRY(h, cudaMalloc(&s.soft64, s.cap_cw * h->d.n_cw * sizeof(double)));
    return 0;
}

int ensure_generic_staging(nrldpc_handle *h, PipeSlot &s, size_t bytes) {
    if (bytes <= s.cap_bytes) return 0;
    cudaFree(s.bytes_in); cudaFree(s.bytes_out); cudaFree(s.f_in);
    s.bytes_in = s.bytes_out = nullptr; s.f_in = nullptr; s.cap_bytes = 0;
    CUDA_TRY(h, cudaMalloc(&s.bytes_in, bytes));
    CUDA_TRY(h, cudaMalloc(&s.bytes_out, bytes));
    CUDA_TRY(h, cudaMalloc(&s.f_in, bytes));
    s.cap_bytes = bytes;
    return 0;
}

int make_geom(nrldpc_handle *h, const nrldpc_rm *rm, nrldpc::RmGeom *g) {
    if (!rm) return fail(h, NRLDPC_ESHAPE, "rate-matching geometry is NULL");
    const nrldpc_dims &d = h->d;
    if (!(rm->Q_m == 1 || rm->Q_m == 2 || rm->Q_m == 4 || rm->Q_m == 6 || rm->Q_m == 8))
        return fail(h, NRLDPC_EUNSUPPORTED, "Q_m (bits per symbol) is one of 1, 2, 4, 6, 8.");
    if (rm->E <= 0 || rm->E % rm->Q_m) return fail(h, NRLDPC_EUNSUPPORTED, "E must be a positive multiple of Q_m.");
    if (rm->N_cb <= 0 || rm->N_cb > d.N) return fail(h, NRLDPC_EUNSUPPORTED, "N_cb must be in (0, N].");
    if (rm->k_0 < 0 || rm->k_0 >= rm->N_cb) return fail(h, NRLDPC_EUNSUPPORTED, "k_0 must be in [0, N_cb).");
    if (rm->K_prime <= 0 || rm->K_prime > d.K) return fail(h, NRLDPC_EUNSUPPORTED, "K_prime must be in (0, K].");
    g->E = rm->E; g->Ncb = rm->N_cb; g->Qm = rm->Q_m; g->EQ = rm->E / rm->Q_m;
    g->Z2 = 2 * d.Z; g->N = d.N; g->ncw = d.n_cw;
    g->F0u = std::max(rm->K_prime - 2 * d.Z, 0);
    g->F1u = d.K - 2 * d.Z;
    if (g->F1u < g->F0u) g->F1u = g->F0u;
    g->F0 = std::min(g->F0u, rm->N_cb);
    g->F1 = std::min(g->F1u, rm->N_cb);
    g->Nnf = rm->N_cb - (g->F1 - g->F0);
    if (g->Nnf <= 0) return fail(h, NRLDPC_EUNSUPPORTED, "circular buffer holds only filler bits.");
    int t = rm->k_0 - g->F0;
    t = t < 0 ? 0 : (t > g->F1 - g->F0 ? g->F1 - g->F0 : t);
    g->rank_k0 = (rm->k_0 - t) % g->Nnf;
    return 0;
}

int grid_for(const nrldpc_handle *h, long long work_items, int threads) {
    long long blocks = (work_items + threads - 1) / threads;
    return (int)std::max<long long>(1, std::min<long long>(blocks, (long long)h->num_sms * 16));
}

// rate matching / recovery launchers: shared-memory staged kernels when a row fits, element-wise otherwise
constexpr size_t kStageSmemMax = 200 * 1024;

template <int QM>
int launch_rm_staged(nrldpc_handle *h, cudaStream_t st, const uint8_t *cw, uint8_t *f, int64_t n, const nrldpc::RmGeom &g) {
    const size_t smem = (size_t)g.ncw;
    if (int rc = raise_max_smem(h, reinterpret_cast<const void *>(nrldpc::rate_match_staged_kernel<QM>), smem)) return rc;
    const int grid = (int)std::min<int64_t>(n, (int64_t)h->num_sms * 8);
    nrldpc::rate_match_staged_kernel<QM><<<grid, 256, smem, st>>>(cw, f, n, g);
    return 0;
}

int launch_rate_match(nrldpc_handle *h, cudaStream_t st, const uint8_t *cw, uint8_t *f, int64_t n, const nrldpc::RmGeom &g) {
    const bool aligned = (reinterpret_cast<uintptr_t>(cw) & 3) == 0;
    int rc = 0;
    if (aligned && (size_t)g.ncw <= kStageSmemMax) {
        switch (g.Qm) {
            case 1: rc = launch_rm_staged<1>(h, st, cw, f, n, g); break;
            case 2: rc = launch_rm_staged<2>(h, st, cw, f, n, g); break;
            case 4: rc = launch_rm_staged<4>(h, st, cw, f, n, g); break;
            case 6: rc = launch_rm_staged<6>(h, st, cw, f, n, g); break;
            default: rc = launch_rm_staged<8>(h, st, cw, f, n, g); break;
        }
        if (rc) return rc;
    } else {
        nrldpc::rate_match_kernel<<<grid_for(h, n * g.E, 256), 256, 0, st>>>(cw, f, n, g);
    }
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

template <int QM>
int launch_rr_staged(nrldpc_handle *h, cudaStream_t st, const float *f, float *harq, float *out, int64_t n, const nrldpc::RmGeom &g) {
    const size_t smem = (size_t)g.E * sizeof(float);
    if (int rc = raise_max_smem(h, reinterpret_cast<const void *>(nrldpc::rate_recover_staged_kernel<QM>), smem)) return rc;
    const int grid = (int)std::min<int64_t>(n, (int64_t)h->num_sms * 8);
    nrldpc::rate_recover_staged_kernel<QM><<<grid, 256, smem, st>>>(f, harq, out, n, g);
    return 0;
}

// Gather table of this geometry (built once on an internal stream and finished before it is handed out, so that consumers on
// any stream may read it); nullptr when tables are switched off or too many geometries are alive: the kernels then evaluate
// the map themselves.
const int *rr_table_for(nrldpc_handle *h, const nrldpc::RmGeom &g, uint32_t magic) {
    if (!h->rr_table_on) return nullptr;
    for (const auto &t : h->rr_tables)
        if (memcmp(&t.g, &g, sizeof(g)) == 0) return t.d;
    if (h->rr_tables.size() >= 64) return nullptr;
    if (ensure_pipe(h)) return nullptr;
    int *d = nullptr;
    if (cudaMalloc(&d, (size_t)g.ncw * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    cudaStream_t st = h->pipe[0].stream;
    nrldpc::rr_table_kernel<<<grid_for(h, g.ncw, 256), 256, 0, st>>>(d, g, g.Qm, magic);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { cudaGetLastError(); cudaFree(d); return nullptr; }
    h->launches += 1;
    nrldpc_handle::RrTable t;
    memcpy(&t.g, &g, sizeof(g));
    t.d = d;
    h->rr_tables.push_back(t);
    return d;
}

template <int QM>
int launch_rr_tma(nrldpc_handle *h, cudaStream_t st, const float *f, float *harq, float *out, int64_t n, const nrldpc::RmGeom &g) {
    const size_t row = (size_t)g.E * sizeof(float);
    // double-buffer only while two CTAs still fit on an SM: measured on Cfg-H, two single-buffered CTAs per SM
    // (243 us) beat one double-buffered CTA (313 us)
    const int n_buf = 4 * row <= kStageSmemMax ? 2 : 1;
    const size_t smem = n_buf * row;
    int occ = 1;
    if (int rc = cached_occupancy(h, reinterpret_cast<const void *>(nrldpc::rate_recover_tma_kernel<QM>), 512, smem, &occ)) return rc;
    const int grid = (int)std::min<int64_t>(n, (int64_t)h->num_sms * std::max(1, occ));
    const uint32_t magic = g.EQ == 1 ? 0xffffffffu : (uint32_t)((1ull << 32) / (uint64_t)g.EQ);   // floor: quotient estimate is exact or one low
    nrldpc::rate_recover_tma_kernel<QM><<<grid, 512, smem, st>>>(f, harq, out, n, g, n_buf, magic, rr_table_for(h, g, magic));
    return 0;
}

int launch_rate_recover(nrldpc_handle *h, cudaStream_t st, const float *f, float *harq, float *out, int64_t n, const nrldpc::RmGeom &g) {
    const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    int rc = 0;
    const bool tma_ok = aligned && (g.E & 3) == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0 && g.EQ > 1 &&
                        (size_t)g.E * sizeof(float) <= kStageSmemMax && !h->no_tma;
    if (tma_ok) {
        switch (g.Qm) {
            case 1: rc = launch_rr_tma<1>(h, st, f, harq, out, n, g); break;
            case 2: rc = launch_rr_tma<2>(h, st, f, harq, out, n, g); break;
            case 4: rc = launch_rr_tma<4>(h, st, f, harq, out, n, g); break;
            case 6: rc = launch_rr_tma<6>(h, st, f, harq, out, n, g); break;
            default: rc = launch_rr_tma<8>(h, st, f, harq, out, n, g); break;
        }
        if (rc) return rc;
    } else if (aligned && (size_t)g.E * sizeof(float) <= kStageSmemMax) {
        switch (g.Qm) {
            case 1: rc = launch_rr_staged<1>(h, st, f, harq, out, n, g); break;
            case 2: rc = launch_rr_staged<2>(h, st, f, harq, out, n, g); break;
            case 4: rc = launch_rr_staged<4>(h, st, f, harq, out, n, g); break;
            case 6: rc = launch_rr_staged<6>(h, st, f, harq, out, n, g); break;
            default: rc = launch_rr_staged<8>(h, st, f, harq, out, n, g); break;
        }
        if (rc) return rc;
    } else {
        nrldpc::rate_recover_kernel<<<grid_for(h, n * g.ncw, 256), 256, 0, st>>>(f, harq, out, n, g);
    }
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

}  // namespace

// ================================================================================================
NRLDPC_EXPORT const char *nrldpc_version(void) { return "nrldpc_b200 0.1 (sm_100a)"; }

NRLDPC_EXPORT int nrldpc_set_index(int32_t Z) {
    for (int s = 0; s < 8; ++s)
        for (int j = 0; j < kSetN[s]; ++j)
            if ((kSetA[s] << j) == Z) return s;
    return NRLDPC_EUNSUPPORTED;
}

NRLDPC_EXPORT int nrldpc_lifting_size(int32_t K_b, int32_t K_prime) {
    int best = 0;
    for (int s = 0; s < 8; ++s)
        for (int j = 0; j < kSetN[s]; ++j) {
            const int Z = kSetA[s] << j;
            if ((long long)K_b * Z >= K_prime && (best == 0 || Z < best)) best = Z;
        }
    return best ? best : NRLDPC_EUNSUPPORTED;
}

NRLDPC_EXPORT int nrldpc_base_graph(int32_t bg, int32_t i_LS, int32_t *rows, int32_t *cols, int32_t *shifts) {
    if (bg < 1 || bg > 2 || i_LS < 0 || i_LS > 7) return NRLDPC_EUNSUPPORTED;
    const BgView v = bg_view(bg);
    for (int e = 0; e < v.edges; ++e) {
        if (rows) rows[e] = v.row[e];
        if (cols) cols[e] = v.col[e];
        if (shifts) shifts[e] = v.sh(i_LS, e);
    }
    return v.edges;
}

NRLDPC_EXPORT const char *nrldpc_last_error(const nrldpc_t *h) { return h ? h->err : g_create_error; }

NRLDPC_EXPORT int64_t nrldpc_launch_count(const nrldpc_t *h) { return h ? h->launches : 0; }

NRLDPC_EXPORT void *nrldpc_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
NRLDPC_EXPORT void nrldpc_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

NRLDPC_EXPORT int nrldpc_create(nrldpc_t **out, const nrldpc_cfg *cfg) {
    if (!out || !cfg) return fail(nullptr, NRLDPC_ESHAPE, "nrldpc_create: NULL argument");
    *out = nullptr;
    if (cfg->bg < 1 || cfg->bg > 2) return fail(nullptr, NRLDPC_EUNSUPPORTED, "bg selects the TS 38.212 base graph: 1 or 2.");
    const int ils = nrldpc_set_index(cfg->Z);
    if (ils < 0) return fail(nullptr, NRLDPC_EUNSUPPORTED, "Invalid lifting size.");
    if (cfg->max_iters < 1) return fail(nullptr, NRLDPC_EUNSUPPORTED, "MaximumIterationCount must be >= 1.");
    if (cfg->llr_dtype != NRLDPC_F32 && cfg->llr_dtype != NRLDPC_F16X2)
        return fail(nullptr, NRLDPC_EUNSUPPORTED, "llr_dtype must be NRLDPC_F32 or NRLDPC_F16X2.");
    if (cfg->algorithm != NRLDPC_ALG_NMS && cfg->algorithm != NRLDPC_ALG_BP)
        return fail(nullptr, NRLDPC_EUNSUPPORTED, "algorithm must be NRLDPC_ALG_NMS or NRLDPC_ALG_BP.");
    if (cfg->algorithm == NRLDPC_ALG_BP && cfg->llr_dtype != NRLDPC_F32)
        return fail(nullptr, NRLDPC_EUNSUPPORTED, "NRLDPC_ALG_BP computes in float64; llr_dtype must be NRLDPC_F32 (the default).");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(nullptr, NRLDPC_ECUDA, "no CUDA device available (%s); this library has no CPU path",
                    cudaGetErrorString(ce));
    nrldpc_handle *h = new (std::nothrow) nrldpc_handle();
    if (!h) return fail(nullptr, NRLDPC_ENOMEM, "out of host memory");
    h->err[0] = 0;
    h->cfg = *cfg;
    if (!(h->cfg.alpha > 0.0f)) h->cfg.alpha = 0.75f;
    int dev = cfg->device;
    if (dev < 0) cudaGetDevice(&dev);
    if (dev >= ndev) { delete h; return fail(nullptr, NRLDPC_EUNSUPPORTED, "device ordinal %d out of range", dev); }
    h->device = dev;
    DeviceGuard guard_;   // the caller's current device is restored on every return path
    ce = guard_.enter(dev);
    if (ce != cudaSuccess) { delete h; return fail(nullptr, NRLDPC_ECUDA, "cudaSetDevice(%d): %s", dev, cudaGetErrorString(ce)); }
    cudaDeviceProp prop{};
    ce = cudaGetDeviceProperties(&prop, dev);
    if (ce != cudaSuccess) { delete h; return fail(nullptr, NRLDPC_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(ce)); }
    if (prop.major != 10) {
        delete h;
        return fail(nullptr, NRLDPC_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
    }
    h->num_sms = prop.multiProcessorCount;
    {
        // persisting L2 set-aside for the c2v scratch (a device-wide limit: it is only ever raised, never lowered)
        const char *v = getenv("NRLDPC_L2_WINDOW");
        if (!(v && atoi(v) == 0) && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
            size_t cur = 0;
            cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
            if (cur < (size_t)prop.persistingL2CacheMaxSize) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, prop.persistingL2CacheMaxSize);
            cudaGetLastError();
            h->l2_window = std::min<long long>(prop.persistingL2CacheMaxSize, prop.accessPolicyMaxWindowSize);
        }
    }
    if (const char *v = getenv("NRLDPC_DECODE_VARIANT")) h->dec_variant = strcmp(v, "loop") == 0 ? 0 : strcmp(v, "shfl") == 0 ? 2 : 1;
    if (const char *v = getenv("NRLDPC_SHFL_CWPC")) h->shfl_cwpc = std::max(0, atoi(v));
    if (const char *v = getenv("NRLDPC_SHFL_THREADS")) h->shfl_threads = std::max(32, std::min(256, atoi(v) / 32 * 32));
    if (const char *v = getenv("NRLDPC_L2_PIN")) h->l2_pin = atoi(v) ? 1 : 0;
    if (const char *v = getenv("NRLDPC_BITSLICED_MIN_ROWS")) h->bitsliced_min_rows = std::max(4, atoi(v));
    if (const char *v = getenv("NRLDPC_STAGED_MIN_ROWS")) h->staged_min_rows = std::max(5, atoi(v));
    if (const char *v = getenv("NRLDPC_BP_THREADS")) h->bp_threads = atoi(v) > 512 ? 1024 : 512;
    if (const char *v = getenv("NRLDPC_GRID_CAP")) h->grid_cap = std::max(0, atoi(v));
    if (getenv("NRLDPC_NO_TMA")) h->no_tma = 1;
    if (getenv("NRLDPC_NO_STAGING")) h->no_staging = 1;
    if (const char *v = getenv("NRLDPC_ZERO_COPY_MAX")) h->zero_copy_max = std::max(0, std::min(64, atoi(v)));
    if (const char *v = getenv("NRLDPC_REFILL")) h->refill = std::max(0, std::min(3, atoi(v)));
    if (const char *v = getenv("NRLDPC_RR_TABLE")) h->rr_table_on = atoi(v) ? 1 : 0;
    if (const char *v = getenv("NRLDPC_REFILL_SPARES")) h->refill_spares = std::max(1, std::min(8, atoi(v)));
    if (const char *v = getenv("NRLDPC_CWPC")) h->cwpc_override = std::max(0, atoi(v));
    if (const char *v = getenv("NRLDPC_OCC_CAP")) { h->occ_cap = std::max(1, std::min(32, atoi(v))); h->occ_cap_forced = 1; }
    if (const char *v = getenv("NRLDPC_SHAPE_MODEL")) h->shape_model = atoi(v) ? 1 : 0;
    h->host_threads = nrldpc::default_host_threads();

    const BgView v = bg_view(cfg->bg);
    const int Z = cfg->Z;
    h->d = nrldpc_dims{cfg->bg, Z, ils, v.rows, v.cols, v.kcols, v.edges, v.kcols * Z, (v.cols - 2) * Z, v.cols * Z};
    choose_decode_shape(h);
    std::vector<uint32_t> ed(v.edges);
    for (int r = 0, e = 0; r <= v.rows; ++r) {
        while (e < v.edges && v.row[e] < r) ++e;
        h->h_row_start[r] = e;
    }
    memset(&h->dec_args, 0, sizeof(h->dec_args));
    bool bad_structure = false;
    for (int r = 0; r <= v.rows; ++r) h->dec_args.row_start[r] = (unsigned short)h->h_row_start[r];
    for (int e = 0; e < v.edges; ++e) {
        const int sft = v.sh(ils, e) % Z;
        ed[e] = ((uint32_t)(v.col[e] * Z) << 16) | (uint32_t)sft;
        h->dec_args.ed[e] = make_uint2((uint32_t)sft * 4u, (uint32_t)(v.col[e] * Z) * 4u);
        // every kernel relies on the structure of the extension part (TS 38.212 Tables 5.3.2-2/-3): row r >= 4 ends in an
        // identity circulant in column kcols + r
        if (v.row[e] >= 4 && (e + 1 == v.edges || v.row[e + 1] != v.row[e]) && (sft != 0 || v.col[e] != v.kcols + v.row[e]))
            bad_structure = true;
    }
    // ... and that column belongs to no other row (a degree-1 variable: DESIGN.md section 2, oracle A revision 2)
    for (int c = v.kcols + 4; c < v.cols; ++c) {
        int cnt = 0;
        for (int e = 0; e < v.edges; ++e) cnt += v.col[e] == c;
        if (cnt != 1) bad_structure = true;
    }
    if (bad_structure) { delete h; return fail(nullptr, NRLDPC_EUNSUPPORTED, "unexpected extension-parity structure in the base graph table"); }
    // encoder structure: shifts of the first core-parity column in rows 0..3
    int vals[3], nv = 0;
    for (int r = 0; r < 4; ++r) h->enc_s0[r] = -1;
    for (int e = 0; e < v.edges; ++e)
        if (v.col[e] == v.kcols && v.row[e] < 4) {
            h->enc_s0[v.row[e]] = v.sh(ils, e) % Z;
            if (nv < 3) vals[nv] = v.sh(ils, e) % Z;
            ++nv;
        }
    if (nv != 3) { delete h; return fail(nullptr, NRLDPC_EUNSUPPORTED, "unexpected core parity structure"); }
    h->enc_delta = vals[0] == vals[1] ? vals[2] : (vals[0] == vals[2] ? vals[1] : vals[0]);

    // tables of the sum-product kernel: per edge (shift, col*Z) and the column-major edge lists
    std::vector<int> bp_shift(v.edges), bp_colz(v.edges), bp_col_start(v.cols + 1, 0), bp_col_edge(v.edges);
    for (int e = 0; e < v.edges; ++e) {
        bp_shift[e] = v.sh(ils, e) % Z;
        bp_colz[e] = v.col[e] * Z;
        bp_col_start[v.col[e] + 1] += 1;
    }
    for (int c = 0; c < v.cols; ++c) bp_col_start[c + 1] += bp_col_start[c];
    {
        std::vector<int> fill(bp_col_start.begin(), bp_col_start.end() - 1);
        for (int e = 0; e < v.edges; ++e) bp_col_edge[fill[v.col[e]]++] = e;   // ascending e inside a column
    }

    int rc = 0;
    auto up = [&]() -> int {
        CUDA_TRY(h, cudaMalloc(&h->bp_shift, sizeof(int) * v.edges));
        CUDA_TRY(h, cudaMalloc(&h->bp_colz, sizeof(int) * v.edges));
        CUDA_TRY(h, cudaMalloc(&h->bp_col_start, sizeof(int) * (v.cols + 1)));
        CUDA_TRY(h, cudaMalloc(&h->bp_col_edge, sizeof(int) * v.edges));
        CUDA_TRY(h, cudaMemcpy(h->bp_shift, bp_shift.data(), sizeof(int) * v.edges, cudaMemcpyHostToDevice));
        CUDA_TRY(h, cudaMemcpy(h->bp_colz, bp_colz.data(), sizeof(int) * v.edges, cudaMemcpyHostToDevice));
        CUDA_TRY(h, cudaMemcpy(h->bp_col_start, bp_col_start.data(), sizeof(int) * (v.cols + 1), cudaMemcpyHostToDevice));
        CUDA_TRY(h, cudaMemcpy(h->bp_col_edge, bp_col_edge.data(), sizeof(int) * v.edges, cudaMemcpyHostToDevice));
        CUDA_TRY(h, cudaMalloc(&h->edesc, ed.size() * sizeof(uint32_t)));
        CUDA_TRY(h, cudaMalloc(&h->row_start, sizeof(int) * (v.rows + 1)));
        CUDA_TRY(h, cudaMemcpy(h->edesc, ed.data(), ed.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CUDA_TRY(h, cudaMemcpy(h->row_start, h->h_row_start, sizeof(int) * (v.rows + 1), cudaMemcpyHostToDevice));
        uint32_t *d_base = nullptr;
        CUDA_TRY(h, cudaMalloc(&d_base, sizeof(uint32_t)));
        nrldpc::smem_base_probe<<<1, 1, 16>>>(d_base);
        cudaError_t pe = cudaMemcpy(&h->smem_base, d_base, sizeof(uint32_t), cudaMemcpyDeviceToHost);
        cudaFree(d_base);
        CUDA_TRY(h, pe);
        return 0;
    };
    rc = up();
    if (rc) {
        snprintf(g_create_error, sizeof(g_create_error), "%s", h->err);
        nrldpc_destroy(h);
        return rc;
    }
    *out = h;
    return NRLDPC_OK;
}

NRLDPC_EXPORT void nrldpc_destroy(nrldpc_t *h) {
    if (!h) return;
    DeviceGuard guard_;
    guard_.enter(h->device);
    // wait for this handle's own work only (its pipeline streams and the last device-memory launch that used its
    // scratch), not for the whole device: other handles and the caller's streams keep running
    for (auto &s : h->pipe)
        if (s.stream) cudaStreamSynchronize(s.stream);
    if (h->dev_done && h->dev_used) cudaEventSynchronize(h->dev_done);
    for (auto &s : h->pipe) {
        cudaFree(s.llr); cudaFree(s.hard); cudaFree(s.soft); cudaFree(s.iters); cudaFree(s.ok); cudaFree(s.llr16);
        cudaFree(s.bytes_in); cudaFree(s.bytes_out); cudaFree(s.f_in);
        cudaFree(s.c2v); cudaFree(s.counter);
        cudaFree(s.llr64); cudaFree(s.soft64); cudaFree(s.rmsg);
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.h_hard) cudaFreeHost(s.h_hard);
        if (s.h_ok) cudaFreeHost(s.h_ok);
        if (s.h_iters) cudaFreeHost(s.h_iters);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    delete h->pool;
    if (h->dev_done) cudaEventDestroy(h->dev_done);
    cudaFree(h->dev_widen);
    for (auto &t : h->rr_tables) cudaFree(t.d);
    cudaFree(h->edesc);
    cudaFree(h->row_start);
    cudaFree(h->bp_shift); cudaFree(h->bp_colz); cudaFree(h->bp_col_start); cudaFree(h->bp_col_edge);
    delete h;
}

NRLDPC_EXPORT int nrldpc_synchronize(nrldpc_t *h) {
    if (!h) return NRLDPC_ESHAPE;
    ENTER_DEVICE(h);
    CUDA_TRY(h, cudaDeviceSynchronize());
    return 0;
}

NRLDPC_EXPORT int nrldpc_get_dims(const nrldpc_t *h, nrldpc_dims *out) {
    if (!h || !out) return NRLDPC_ESHAPE;
    *out = h->d;
    return 0;
}

// ------------------------------------------------------------------------------------------------
namespace {
int widen(nrldpc_handle *h, cudaStream_t st, const uint16_t *in, float *out, int64_t n_cw_total) {
    const long long n4 = (long long)n_cw_total * h->d.n_cw / 4;   // n_cw is a multiple of 4
    nrldpc::widen_f16_kernel<<<grid_for(h, n4, 256), 256, 0, st>>>(reinterpret_cast<const uint2 *>(in), reinterpret_cast<float4 *>(out), n4);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

int widen8(nrldpc_handle *h, cudaStream_t st, const int8_t *in, float *out, int64_t n_cw_total) {
    const long long n4 = (long long)n_cw_total * h->d.n_cw / 4;   // n_cw is a multiple of 4
    nrldpc::widen_i8_kernel<<<grid_for(h, n4, 256), 256, 0, st>>>(reinterpret_cast<const uint32_t *>(in), reinterpret_cast<float4 *>(out), n4, h->i8_scale);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

int narrow(nrldpc_handle *h, cudaStream_t st, const double *in, float *out, int64_t n_cw_total) {
    const long long n = (long long)n_cw_total * h->d.n_cw;
    nrldpc::narrow_f64_kernel<<<grid_for(h, n, 256), 256, 0, st>>>(in, out, n);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

// One decode launch (plus the conversion kernel the transport type needs) on device buffers.
//   in_kind = kInF32: llr float32, soft float32;  kInF16: llr binary16, soft float32;  kInF64: llr float64, soft float64.
//   f32_tmp: float32 scratch of batch*n_cw entries (needed when a conversion precedes the kernel)
int launch_any(nrldpc_handle *h, PipeSlot &s, cudaStream_t st, const void *llr, int in_kind, float *f32_tmp, int64_t batch,
               int n_rows, uint8_t *hard, void *soft, int32_t *iters, uint8_t *ok) {
    const bool bp = h->cfg.algorithm == NRLDPC_ALG_BP;
    if (in_kind == kInF64 && bp)   // the reference's arithmetic on the reference's own input type
        return launch_decode_bp<double>(h, s, st, static_cast<const double *>(llr), batch, n_rows, hard,
                                        static_cast<double *>(soft), iters, ok);
    const float *src = static_cast<const float *>(llr);
    if (in_kind == kInF16) {
        if (int rc = widen(h, st, static_cast<const uint16_t *>(llr), f32_tmp, batch)) return rc;
        src = f32_tmp;
    } else if (in_kind == kInI8) {
        if (int rc = widen8(h, st, static_cast<const int8_t *>(llr), f32_tmp, batch)) return rc;
        src = f32_tmp;
    } else if (in_kind == kInF64) {
        if (int rc = narrow(h, st, static_cast<const double *>(llr), f32_tmp, batch)) return rc;
        src = f32_tmp;
    }
    if (bp) return launch_decode_bp<float>(h, s, st, src, batch, n_rows, hard, static_cast<float *>(soft), iters, ok);
    return launch_decode(h, s, st, src, batch, n_rows, hard, static_cast<float *>(soft), iters, ok);
}

// llr: float32 / IEEE binary16 / float64 rows in cw layout (in_kind)
int decode_impl(nrldpc_t *h, const void *llr, int in_kind, int64_t batch, int32_t n_rows, uint8_t *info_hard,
                void *app_soft, int32_t *iters, uint8_t *parity_ok, int32_t mem, void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    if (batch < 0) return fail(h, NRLDPC_ESHAPE, "batch must be >= 0");
    if (batch == 0) return 0;
    if (!llr || !info_hard) return fail(h, NRLDPC_ESHAPE, "llr and info_hard must not be NULL");
    if (n_rows == 0) n_rows = h->d.rows;
    if (n_rows < 4 || n_rows > h->d.rows) return fail(h, NRLDPC_EUNSUPPORTED, "n_rows must be 0 or in [4, %d]", h->d.rows);
    const bool bp = h->cfg.algorithm == NRLDPC_ALG_BP;
    if (in_kind == kInF64 && !bp && app_soft)
        return fail(h, NRLDPC_EUNSUPPORTED, "nrldpc_decode64 returns app_soft only with NRLDPC_ALG_BP (the min-sum kernels compute in float32)");
    ENTER_DEVICE(h);
    const size_t in_elt = in_kind == kInF16 ? sizeof(uint16_t) : in_kind == kInF64 ? sizeof(double) : in_kind == kInI8 ? sizeof(int8_t) : sizeof(float);
    const size_t soft_elt = in_kind == kInF64 ? sizeof(double) : sizeof(float);
    const bool convert = in_kind == kInF16 || in_kind == kInI8 || (in_kind == kInF64 && !bp);
    if (mem == NRLDPC_MEM_DEVICE) {
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if ((reinterpret_cast<uintptr_t>(llr) & 15) || (app_soft && (reinterpret_cast<uintptr_t>(app_soft) & 15)))
            return fail(h, NRLDPC_ESHAPE, "device llr / app_soft pointers must be 16-byte aligned");
        if (reinterpret_cast<uintptr_t>(info_hard) & 3)
            return fail(h, NRLDPC_ESHAPE, "device info_hard pointer must be 4-byte aligned");
        if (!h->dev_done) CUDA_TRY(h, cudaEventCreateWithFlags(&h->dev_done, cudaEventDisableTiming));
        if (convert && (size_t)batch > h->dev_widen_cw) {
            CUDA_TRY(h, cudaStreamSynchronize(st));
            cudaFree(h->dev_widen);
            h->dev_widen = nullptr; h->dev_widen_cw = 0;
            CUDA_TRY(h, cudaMalloc(&h->dev_widen, (size_t)batch * h->d.n_cw * sizeof(float)));
            h->dev_widen_cw = (size_t)batch;
        }
        // device-mode launches share one scratch (c2v records, work counter, conversion buffer): a launch on another
        // stream than the previous one is ordered behind it
        // (a launch being captured into a CUDA graph takes no part in this: an event recorded during capture belongs to
        // the graph and cannot be waited on by ordinary work afterwards; whoever replays the graph orders the replays
        // against other use of the handle, nrldpc_b200.h)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
        const bool capturing = cap != cudaStreamCaptureStatusNone;
        if (!capturing && h->dev_used && st != h->last_dev_stream) CUDA_TRY(h, cudaStreamWaitEvent(st, h->dev_done, 0));
        if (int rc = launch_any(h, h->pipe[0], st, llr, in_kind, h->dev_widen, batch, n_rows, info_hard, app_soft, iters, parity_ok))
            return rc;
        if (capturing) return 0;
        CUDA_TRY(h, cudaEventRecord(h->dev_done, st));
        h->last_dev_stream = st;
        h->dev_used = true;
        return 0;
    }
    if (mem != NRLDPC_MEM_HOST) return fail(h, NRLDPC_ESHAPE, "mem must be NRLDPC_MEM_HOST or NRLDPC_MEM_DEVICE");

    // Host buffers: chunked 3-deep pipeline, H2D / kernel / D2H of different chunks overlap.  Chunks are one
    // persistent-grid wave of codewords (doubled while small) so the un-overlapped tail stays short.
    if (int rc = ensure_pipe(h)) return rc;
    if (h->dev_used) CUDA_TRY(h, cudaStreamWaitEvent(h->pipe[0].stream, h->dev_done, 0));

    // A few codewords (the reference calls step() with ONE code block, NRLDPCDecoder.m:257-266): no copy engine at all.
    // The kernel reads the LLRs straight from pinned host memory (its bulk copy crosses PCIe itself) and writes the
    // decisions straight back -- one launch and one synchronisation instead of H2D + launch + D2H + synchronisation.
    // Caller buffers that are pageable, float64 or misaligned pass through the pinned ring on the caller's thread.
    if (batch <= h->zero_copy_max && !bp && in_kind != kInF16 && in_kind != kInI8 && !app_soft) {
        PipeSlot &s = h->pipe[0];
        if (int rc = ensure_host_ring(h, s, (size_t)batch * h->d.n_cw * sizeof(float), (size_t)batch, true, true)) return rc;
        const float *src = static_cast<const float *>(llr);
        const size_t total = (size_t)batch * h->d.n_cw;
        if (in_kind == kInF64) {
            nrldpc::narrow_f64_to_f32(static_cast<const double *>(llr), reinterpret_cast<float *>(s.h_in), total);
            src = reinterpret_cast<const float *>(s.h_in);
        } else if (is_pageable(llr) || (reinterpret_cast<uintptr_t>(llr) & 15)) {
            memcpy(s.h_in, llr, total * sizeof(float));
            src = reinterpret_cast<const float *>(s.h_in);
        }
        const bool out_direct = !is_pageable(info_hard) && !(reinterpret_cast<uintptr_t>(info_hard) & 3) &&
                                (!iters || !is_pageable(iters)) && (!parity_ok || !is_pageable(parity_ok));
        uint8_t *d_hard = out_direct ? info_hard : s.h_hard;
        int32_t *d_iters = !iters ? nullptr : out_direct ? iters : s.h_iters;
        uint8_t *d_ok = !parity_ok ? nullptr : out_direct ? parity_ok : s.h_ok;
        if (int rc = launch_decode(h, s, s.stream, src, batch, n_rows, d_hard, nullptr, d_iters, d_ok)) return rc;
        CUDA_TRY(h, cudaStreamSynchronize(s.stream));
        if (!out_direct) {
            memcpy(info_hard, s.h_hard, (size_t)batch * h->d.K);
            if (iters) memcpy(iters, s.h_iters, (size_t)batch * sizeof(int32_t));
            if (parity_ok) memcpy(parity_ok, s.h_ok, (size_t)batch);
        }
        return 0;
    }

    const int cwpc = h->cwpc;
    const int64_t wave = bp ? (int64_t)h->num_sms
                            : (int64_t)h->num_sms * nrldpc::kDecCtasPerSm * cwpc *
                                  (h->cfg.llr_dtype == NRLDPC_F16X2 ? 2 : 1);  // codewords per full grid
    int64_t chunk = wave;
    while (chunk * 2 * h->d.n_cw * 4 <= (int64_t)40 << 20 && chunk * 2 * kNumPipe <= batch) chunk *= 2;
    chunk = std::min<int64_t>(chunk, batch);
    for (auto &s : h->pipe)
        if (int rc = ensure_decode_staging(h, s, (size_t)chunk, app_soft != nullptr,
                                           (in_kind == kInF64 && !bp && !h->no_staging) ? (int)kInF32 : in_kind))
            return rc;
    int k = 0;
    const unsigned char *llr_b = static_cast<const unsigned char *>(llr);
    unsigned char *soft_b = static_cast<unsigned char *>(app_soft);
    // Staged path (host_staging.h): float64 LLRs for the float32 kernels are NARROWED on host threads into a pinned ring
    // (PCIe then carries 4 bytes per LLR), pageable input of any type is copied there, and outputs bound for pageable
    // memory leave through pinned buffers -- so every cudaMemcpyAsync below is truly asynchronous and the caller's
    // thread prepares chunk i+1 while the copy engine and the kernel work on chunk i.
    const bool narrow_in = in_kind == kInF64 && !bp && !h->no_staging;
    const bool stage_in = narrow_in || (!h->no_staging && is_pageable(llr));
    const bool stage_out = !h->no_staging && (is_pageable(info_hard) || is_pageable(iters) || is_pageable(parity_ok));
    const size_t stage_elt = narrow_in ? sizeof(float) : in_elt;
    const int kind_dev = narrow_in ? (int)kInF32 : in_kind;        // element type that reaches the device
    if (stage_in || stage_out) {
        if (!h->pool) h->pool = new (std::nothrow) nrldpc::HostPool(h->host_threads);
        if (!h->pool) return fail(h, NRLDPC_ENOMEM, "out of host memory");
        for (auto &s : h->pipe) {
            if (int rc = ensure_host_ring(h, s, stage_in ? (size_t)chunk * h->d.n_cw * stage_elt : 0, stage_out ? (size_t)chunk : 0,
                                          iters != nullptr, parity_ok != nullptr))
                return rc;
            s.pend_n = 0;
        }
    }
    // outputs of a slot's previous chunk: pinned -> caller
    auto drain = [&](PipeSlot &s) -> int {
        if (!s.pend_n) return 0;
        CUDA_TRY(h, cudaEventSynchronize(s.done));
        const size_t nb = (size_t)s.pend_n * h->d.K;
        uint8_t *dst = info_hard + s.pend_off * h->d.K;
        const uint8_t *src = s.h_hard;
        h->pool->parallel_for(nb, (size_t)1 << 20, [&](size_t b, size_t e) { memcpy(dst + b, src + b, e - b); });
        if (iters) memcpy(iters + s.pend_off, s.h_iters, (size_t)s.pend_n * sizeof(int32_t));
        if (parity_ok) memcpy(parity_ok + s.pend_off, s.h_ok, (size_t)s.pend_n);
        s.pend_n = 0;
        return 0;
    };
    for (int64_t off = 0; off < batch; off += chunk, k = (k + 1) % kNumPipe) {
        PipeSlot &s = h->pipe[k];
        const int64_t n = std::min<int64_t>(chunk, batch - off);
        void *d_in = (kind_dev == kInF16 || kind_dev == kInI8) ? static_cast<void *>(s.llr16) : kind_dev == kInF64 ? static_cast<void *>(s.llr64)
                                                                                           : static_cast<void *>(s.llr);
        void *d_soft = !app_soft ? nullptr : in_kind == kInF64 ? static_cast<void *>(s.soft64) : static_cast<void *>(s.soft);
        const unsigned char *src = llr_b + (size_t)off * h->d.n_cw * in_elt;
        if (stage_out) { if (int rc = drain(s)) return rc; }
        if (stage_in) {
            CUDA_TRY(h, cudaEventSynchronize(s.h2d_done));   // the slot's previous chunk has left the pinned input buffer
            const size_t total = (size_t)n * h->d.n_cw;
            if (narrow_in) {
                const double *in = reinterpret_cast<const double *>(src);
                float *out = reinterpret_cast<float *>(s.h_in);
                h->pool->parallel_for(total, (size_t)1 << 16, [&](size_t b, size_t e) { nrldpc::narrow_f64_to_f32(in + b, out + b, e - b); });
            } else {
                unsigned char *out = s.h_in;
                h->pool->parallel_for(total * in_elt, (size_t)1 << 18, [&](size_t b, size_t e) { nrldpc::copy_stream(src + b, out + b, e - b); });
            }
            src = s.h_in;
        }
        CUDA_TRY(h, cudaMemcpyAsync(d_in, src, (size_t)n * h->d.n_cw * stage_elt, cudaMemcpyHostToDevice, s.stream));
        if (stage_in) CUDA_TRY(h, cudaEventRecord(s.h2d_done, s.stream));
        if (int rc = launch_any(h, s, s.stream, d_in, kind_dev, s.llr, n, n_rows, s.hard, d_soft, iters ? s.iters : nullptr,
                                parity_ok ? s.ok : nullptr))
            return rc;
        CUDA_TRY(h, cudaMemcpyAsync(stage_out ? s.h_hard : info_hard + off * h->d.K, s.hard, (size_t)n * h->d.K, cudaMemcpyDeviceToHost, s.stream));
        if (app_soft)
            CUDA_TRY(h, cudaMemcpyAsync(soft_b + (size_t)off * h->d.n_cw * soft_elt, d_soft, (size_t)n * h->d.n_cw * soft_elt,
                                        cudaMemcpyDeviceToHost, s.stream));
        if (iters) CUDA_TRY(h, cudaMemcpyAsync(stage_out ? s.h_iters : iters + off, s.iters, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
        if (parity_ok) CUDA_TRY(h, cudaMemcpyAsync(stage_out ? s.h_ok : parity_ok + off, s.ok, (size_t)n, cudaMemcpyDeviceToHost, s.stream));
        if (stage_out) {
            CUDA_TRY(h, cudaEventRecord(s.done, s.stream));
            s.pend_off = off; s.pend_n = n;
        }
    }
    if (stage_out)
        for (int i = 0; i < kNumPipe; ++i, k = (k + 1) % kNumPipe)   // oldest pending chunk first
            if (int rc = drain(h->pipe[k])) return rc;
    for (auto &s : h->pipe) CUDA_TRY(h, cudaStreamSynchronize(s.stream));
    return 0;
}
}  // namespace

NRLDPC_EXPORT int nrldpc_decode(nrldpc_t *h, const float *llr, int64_t batch, int32_t n_rows, uint8_t *info_hard,
                                float *app_soft, int32_t *iters, uint8_t *parity_ok, int32_t mem, void *stream) {
    return decode_impl(h, llr, kInF32, batch, n_rows, info_hard, app_soft, iters, parity_ok, mem, stream);
}

NRLDPC_EXPORT int nrldpc_decode16(nrldpc_t *h, const uint16_t *llr_f16, int64_t batch, int32_t n_rows, uint8_t *info_hard,
                                  float *app_soft, int32_t *iters, uint8_t *parity_ok, int32_t mem, void *stream) {
    return decode_impl(h, llr_f16, kInF16, batch, n_rows, info_hard, app_soft, iters, parity_ok, mem, stream);
}

NRLDPC_EXPORT int nrldpc_decode8(nrldpc_t *h, const int8_t *llr_q, float scale, int64_t batch, int32_t n_rows, uint8_t *info_hard,
                                 float *app_soft, int32_t *iters, uint8_t *parity_ok, int32_t mem, void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    if (!(scale > 0.0f) || !(scale < 1e30f)) return fail(h, NRLDPC_ESHAPE, "scale must be a positive finite number");
    h->i8_scale = scale;
    return decode_impl(h, llr_q, kInI8, batch, n_rows, info_hard, app_soft, iters, parity_ok, mem, stream);
}

NRLDPC_EXPORT int nrldpc_decode64(nrldpc_t *h, const double *llr_f64, int64_t batch, int32_t n_rows, uint8_t *info_hard,
                                  double *app_soft, int32_t *iters, uint8_t *parity_ok, int32_t mem, void *stream) {
    return decode_impl(h, llr_f64, kInF64, batch, n_rows, info_hard, app_soft, iters, parity_ok, mem, stream);
}

// ------------------------------------------------------------------------------------------------
namespace {
int launch_encode(nrldpc_handle *h, cudaStream_t st, const uint8_t *info, int64_t batch, uint8_t *cw) {
    nrldpc::EncArgs a{};
    a.info = info; a.cw = cw; a.batch = batch; a.Z = h->d.Z; a.ncols = h->d.cols; a.kcols = h->d.kcols;
    a.n_rows = h->d.rows; a.n_edges = h->d.edges; a.edesc = h->edesc; a.row_start = h->row_start;
    for (int r = 0; r < 4; ++r) a.s0[r] = h->enc_s0[r];
    a.delta = h->enc_delta;
    // one CTA per slab of 32 codewords; wide CTAs: the pack / unpack passes stride over all cols*Z positions
    const int threads = std::min(384, std::max(64, (h->d.n_cw / 4 + 31) / 32 * 32));
    const size_t smem = (size_t)h->d.n_cw * 4 + (size_t)h->d.edges * 4 + (h->d.rows + 1) * 4;
    const int64_t n_slabs = (batch + nrldpc::kEncSlab - 1) / nrldpc::kEncSlab;
    const int grid = (int)std::min<int64_t>(n_slabs, (int64_t)h->num_sms * 8);
    if (int rc = raise_max_smem(h, reinterpret_cast<const void *>(nrldpc::encode_kernel), smem)) return rc;
    nrldpc::encode_kernel<<<grid, threads, smem, st>>>(a);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}
}  // namespace

NRLDPC_EXPORT int nrldpc_encode(nrldpc_t *h, const uint8_t *info, int64_t batch, uint8_t *cw, int32_t mem, void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    if (batch < 0) return fail(h, NRLDPC_ESHAPE, "batch must be >= 0");
    if (batch == 0) return 0;
    if (!info || !cw) return fail(h, NRLDPC_ESHAPE, "info and cw must not be NULL");
    ENTER_DEVICE(h);
    if (mem == NRLDPC_MEM_DEVICE) {
        if ((reinterpret_cast<uintptr_t>(info) & 3) || (reinterpret_cast<uintptr_t>(cw) & 3))
            return fail(h, NRLDPC_ESHAPE, "device info / cw pointers must be 4-byte aligned");
        return launch_encode(h, static_cast<cudaStream_t>(stream), info, batch, cw);
    }
    if (mem != NRLDPC_MEM_HOST) return fail(h, NRLDPC_ESHAPE, "mem must be NRLDPC_MEM_HOST or NRLDPC_MEM_DEVICE");
    if (int rc = ensure_pipe(h)) return rc;
    PipeSlot &s = h->pipe[0];
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(batch, ((int64_t)64 << 20) / h->d.n_cw));
    if (int rc = ensure_generic_staging(h, s, (size_t)chunk * h->d.n_cw)) return rc;
    for (int64_t off = 0; off < batch; off += chunk) {
        const int64_t n = std::min<int64_t>(chunk, batch - off);
        CUDA_TRY(h, cudaMemcpyAsync(s.bytes_in, info + off * h->d.K, (size_t)n * h->d.K, cudaMemcpyHostToDevice, s.stream));
        if (int rc = launch_encode(h, s.stream, s.bytes_in, n, s.bytes_out)) return rc;
        CUDA_TRY(h, cudaMemcpyAsync(cw + off * h->d.n_cw, s.bytes_out, (size_t)n * h->d.n_cw, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(h, cudaStreamSynchronize(s.stream));
    }
    return 0;
}

NRLDPC_EXPORT int nrldpc_rate_match(nrldpc_t *h, const uint8_t *cw, int64_t batch, const nrldpc_rm *rm, uint8_t *f,
                                    int32_t mem, void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    nrldpc::RmGeom g{};
    if (int rc = make_geom(h, rm, &g)) return rc;
    if (batch < 0) return fail(h, NRLDPC_ESHAPE, "batch must be >= 0");
    if (batch == 0) return 0;
    if (!cw || !f) return fail(h, NRLDPC_ESHAPE, "cw and f must not be NULL");
    ENTER_DEVICE(h);
    if (mem == NRLDPC_MEM_DEVICE) {
        return launch_rate_match(h, static_cast<cudaStream_t>(stream), cw, f, batch, g);
    }
    if (mem != NRLDPC_MEM_HOST) return fail(h, NRLDPC_ESHAPE, "mem must be NRLDPC_MEM_HOST or NRLDPC_MEM_DEVICE");
    if (int rc = ensure_pipe(h)) return rc;
    PipeSlot &s = h->pipe[0];
    const int64_t per_cw = std::max<int64_t>(h->d.n_cw, g.E);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(batch, ((int64_t)64 << 20) / per_cw));
    if (int rc = ensure_generic_staging(h, s, (size_t)chunk * per_cw)) return rc;
    for (int64_t off = 0; off < batch; off += chunk) {
        const int64_t n = std::min<int64_t>(chunk, batch - off);
        CUDA_TRY(h, cudaMemcpyAsync(s.bytes_in, cw + off * h->d.n_cw, (size_t)n * h->d.n_cw, cudaMemcpyHostToDevice, s.stream));
        if (int rc = launch_rate_match(h, s.stream, s.bytes_in, s.bytes_out, n, g)) return rc;
        CUDA_TRY(h, cudaMemcpyAsync(f + off * g.E, s.bytes_out, (size_t)n * g.E, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(h, cudaStreamSynchronize(s.stream));
    }
    return 0;
}

NRLDPC_EXPORT int nrldpc_rate_recover(nrldpc_t *h, const float *f, int64_t batch, const nrldpc_rm *rm, float *harq,
                                      float *llr_cw, int32_t mem, void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    nrldpc::RmGeom g{};
    if (int rc = make_geom(h, rm, &g)) return rc;
    if (batch < 0) return fail(h, NRLDPC_ESHAPE, "batch must be >= 0");
    if (batch == 0) return 0;
    if (!f || !llr_cw) return fail(h, NRLDPC_ESHAPE, "f and llr_cw must not be NULL");
    ENTER_DEVICE(h);
    if (mem == NRLDPC_MEM_DEVICE) {
        return launch_rate_recover(h, static_cast<cudaStream_t>(stream), f, harq, llr_cw, batch, g);
    }
    if (mem != NRLDPC_MEM_HOST) return fail(h, NRLDPC_ESHAPE, "mem must be NRLDPC_MEM_HOST or NRLDPC_MEM_DEVICE");
    if (int rc = ensure_pipe(h)) return rc;
    PipeSlot &s = h->pipe[0];
    const int64_t per_cw = std::max<int64_t>(h->d.n_cw, g.E);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(batch, ((int64_t)16 << 20) / per_cw));
    if (int rc = ensure_generic_staging(h, s, (size_t)chunk * per_cw * sizeof(float))) return rc;
    float *d_f = s.f_in, *d_out = reinterpret_cast<float *>(s.bytes_out), *d_harq = reinterpret_cast<float *>(s.bytes_in);
    for (int64_t off = 0; off < batch; off += chunk) {
        const int64_t n = std::min<int64_t>(chunk, batch - off);
        CUDA_TRY(h, cudaMemcpyAsync(d_f, f + off * g.E, (size_t)n * g.E * sizeof(float), cudaMemcpyHostToDevice, s.stream));
        if (harq)
            CUDA_TRY(h, cudaMemcpyAsync(d_harq, harq + off * g.N, (size_t)n * g.N * sizeof(float), cudaMemcpyHostToDevice, s.stream));
        if (int rc = launch_rate_recover(h, s.stream, d_f, harq ? d_harq : nullptr, d_out, n, g)) return rc;
        CUDA_TRY(h, cudaMemcpyAsync(llr_cw + off * g.ncw, d_out, (size_t)n * g.ncw * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        if (harq)
            CUDA_TRY(h, cudaMemcpyAsync(harq + off * g.N, d_harq, (size_t)n * g.N * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(h, cudaStreamSynchronize(s.stream));
    }
    return 0;
}

NRLDPC_EXPORT int nrldpc_qpsk_awgn_llr(nrldpc_t *h, const uint8_t *f_bits, int64_t batch, int32_t E, float variance,
                                       uint64_t seed, uint64_t stream_id, float *f_llr, void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    if (batch < 0 || E <= 0) return fail(h, NRLDPC_ESHAPE, "batch must be >= 0 and E > 0");
    if (!(variance > 0.0f)) return fail(h, NRLDPC_EUNSUPPORTED, "variance must be positive");
    const long long total = (long long)batch * E;
    if (total % 4) return fail(h, NRLDPC_EUNSUPPORTED, "batch*E must be a multiple of 4 (two QPSK symbols per thread)");
    if (total == 0) return 0;
    if (!f_bits || !f_llr) return fail(h, NRLDPC_ESHAPE, "f_bits and f_llr must not be NULL");
    if ((reinterpret_cast<uintptr_t>(f_bits) & 3) || (reinterpret_cast<uintptr_t>(f_llr) & 15))
        return fail(h, NRLDPC_ESHAPE, "f_bits must be 4-byte and f_llr 16-byte aligned");
    ENTER_DEVICE(h);
    const long long quads = total / 4;
    nrldpc::qpsk_awgn_llr_kernel<<<grid_for(h, quads, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        f_bits, f_llr, quads, sqrtf(0.5f * variance), 2.8284271247461900976f / variance, seed, stream_id);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

NRLDPC_EXPORT int nrldpc_qpsk_awgn_rate_recover(nrldpc_t *h, const uint8_t *f_bits, int64_t batch, const nrldpc_rm *rm, float variance,
                                                uint64_t seed, uint64_t stream_id, float *harq, float *llr_cw, void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    nrldpc::RmGeom g{};
    if (int rc = make_geom(h, rm, &g)) return rc;
    if (batch < 0) return fail(h, NRLDPC_ESHAPE, "batch must be >= 0");
    if (!(variance > 0.0f)) return fail(h, NRLDPC_EUNSUPPORTED, "variance must be positive");
    if (g.Qm != 2) return fail(h, NRLDPC_EUNSUPPORTED, "the fused channel + rate-recovery kernel is QPSK only (Q_m = 2)");
    if (g.E % 4) return fail(h, NRLDPC_EUNSUPPORTED, "E must be a multiple of 4 (two QPSK symbols per Philox draw)");
    if ((size_t)g.E * sizeof(float) > kStageSmemMax) return fail(h, NRLDPC_EUNSUPPORTED, "E does not fit in shared memory; call the two stages");
    if (batch == 0) return 0;
    if (!f_bits || !llr_cw) return fail(h, NRLDPC_ESHAPE, "f_bits and llr_cw must not be NULL");
    if ((reinterpret_cast<uintptr_t>(f_bits) & 3) || (reinterpret_cast<uintptr_t>(llr_cw) & 15))
        return fail(h, NRLDPC_ESHAPE, "f_bits must be 4-byte and llr_cw 16-byte aligned");
    ENTER_DEVICE(h);
    const size_t smem = (size_t)g.E * sizeof(float);
    int occ = 1;
    if (int rc = cached_occupancy(h, reinterpret_cast<const void *>(nrldpc::qpsk_awgn_rate_recover_kernel), 512, smem, &occ)) return rc;
    const int grid = (int)std::min<int64_t>(batch, (int64_t)h->num_sms * occ);
    const uint32_t magic = g.EQ == 1 ? 0xffffffffu : (uint32_t)((1ull << 32) / (uint64_t)g.EQ);
    nrldpc::qpsk_awgn_rate_recover_kernel<<<grid, 512, smem, static_cast<cudaStream_t>(stream)>>>(
        f_bits, harq, llr_cw, batch, g, magic, sqrtf(0.5f * variance), 2.8284271247461900976f / variance, seed, stream_id,
        rr_table_for(h, g, magic));
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

NRLDPC_EXPORT int nrldpc_bler_count(nrldpc_t *h, const uint8_t *hard, const uint8_t *info, const uint8_t *tb_hat, int64_t tb_hat_stride,
                                    const uint8_t *tb, int64_t tb_stride, const uint8_t *tb_ok, const uint8_t *cb_passed,
                                    const int32_t *iters, int64_t n_tb, int32_t C, int32_t K_prime, int32_t A, uint8_t *latch,
                                    uint64_t *counters, int32_t do_latch, int32_t finalize, void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    if (n_tb < 0 || C < 1 || K_prime < 1 || K_prime > h->d.K || A < 0) return fail(h, NRLDPC_ESHAPE, "n_tb >= 0, C >= 1, 0 < K_prime <= K, A >= 0 required");
    if (n_tb == 0) return 0;
    if (!latch || !counters || (do_latch && (!tb_hat || !tb || !tb_ok || !iters)) || (finalize && (!hard || !info)))
        return fail(h, NRLDPC_ESHAPE, "nrldpc_bler_count: NULL buffer");
    ENTER_DEVICE(h);
    nrldpc::bler_count_kernel<<<grid_for(h, n_tb * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        hard, info, tb_hat, tb_hat_stride, tb, tb_stride, tb_ok, cb_passed, iters, n_tb, C, h->d.K, K_prime, A, latch,
        reinterpret_cast<unsigned long long *>(counters), do_latch, finalize);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

NRLDPC_EXPORT int nrldpc_random_bits(nrldpc_t *h, uint8_t *bits, int64_t rows, int32_t n_bits, int64_t stride, uint64_t seed,
                                     uint64_t stream_id, void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    if (rows < 0 || n_bits < 0 || stride < n_bits) return fail(h, NRLDPC_ESHAPE, "rows, n_bits >= 0 and stride >= n_bits required");
    if (n_bits >= (1 << 27) || rows >= ((int64_t)1 << 43)) return fail(h, NRLDPC_ESHAPE, "n_bits < 2^27 and rows < 2^43 required");
    if (rows == 0 || n_bits == 0) return 0;
    if (!bits) return fail(h, NRLDPC_ESHAPE, "bits must not be NULL");
    ENTER_DEVICE(h);
    const long long total = rows * (((long long)n_bits + 15) >> 4);
    nrldpc::random_bits_kernel<<<grid_for(h, total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(bits, rows, n_bits, stride, seed, stream_id);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Modulation / channel / demodulation for every NRModulator / NRDemodulator setting (device memory only)
namespace {
int check_modem(nrldpc_handle *h, int64_t n_bits, int32_t Q_m, const void *a, const void *b) {
    if (!(Q_m == 1 || Q_m == 2 || Q_m == 4 || Q_m == 6 || Q_m == 8))
        return fail(h, NRLDPC_EUNSUPPORTED, "Unsupported modulation");
    if (n_bits < 0 || n_bits % Q_m) return fail(h, NRLDPC_ESHAPE, "the number of bits must be a non-negative multiple of Q_m");
    if (n_bits && (!a || !b)) return fail(h, NRLDPC_ESHAPE, "buffers must not be NULL");
    return 0;
}
}  // namespace

NRLDPC_EXPORT int nrldpc_modulate(nrldpc_t *h, const uint8_t *bits, int64_t n_bits, int32_t Q_m, float *sym, void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    if (int rc = check_modem(h, n_bits, Q_m, bits, sym)) return rc;
    if (n_bits == 0) return 0;
    if (reinterpret_cast<uintptr_t>(sym) & 7) return fail(h, NRLDPC_ESHAPE, "sym must be 8-byte aligned");
    ENTER_DEVICE(h);
    const long long n_sym = n_bits / Q_m;
    nrldpc::modulate_kernel<<<grid_for(h, n_sym, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        bits, reinterpret_cast<float2 *>(sym), n_sym, Q_m);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

NRLDPC_EXPORT int nrldpc_awgn(nrldpc_t *h, float *sym, int64_t n_sym, float variance, uint64_t seed, uint64_t stream_id,
                              void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    if (n_sym < 0) return fail(h, NRLDPC_ESHAPE, "n_sym must be >= 0");
    if (!(variance >= 0.0f)) return fail(h, NRLDPC_EUNSUPPORTED, "variance must be non-negative");
    if (n_sym == 0) return 0;
    if (!sym || (reinterpret_cast<uintptr_t>(sym) & 7)) return fail(h, NRLDPC_ESHAPE, "sym must be a non-NULL 8-byte aligned pointer");
    ENTER_DEVICE(h);
    nrldpc::awgn_kernel<<<grid_for(h, (n_sym + 1) / 2, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<float2 *>(sym), n_sym, sqrtf(0.5f * variance), seed, stream_id);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

NRLDPC_EXPORT int nrldpc_demodulate(nrldpc_t *h, const float *sym, int64_t n_sym, int32_t Q_m, float variance, int32_t method,
                                    float *llr, void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    if (int rc = check_modem(h, n_sym < 0 ? -1 : n_sym * Q_m, Q_m, sym, llr)) return rc;
    if (method < 0 || method > 2) return fail(h, NRLDPC_EUNSUPPORTED, "Unsupported decision method");
    if (!(variance > 0.0f)) return fail(h, NRLDPC_EUNSUPPORTED, "variance must be positive");
    if (n_sym == 0) return 0;
    if (reinterpret_cast<uintptr_t>(sym) & 7) return fail(h, NRLDPC_ESHAPE, "sym must be 8-byte aligned");
    ENTER_DEVICE(h);
    nrldpc::demodulate_kernel<<<grid_for(h, n_sym, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2 *>(sym), llr, n_sym, Q_m, 1.0f / variance, method);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

NRLDPC_EXPORT int nrldpc_mod_awgn_llr(nrldpc_t *h, const uint8_t *bits, int64_t n_bits, int32_t Q_m, float variance,
                                      int32_t method, uint64_t seed, uint64_t stream_id, float *llr, void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    if (int rc = check_modem(h, n_bits, Q_m, bits, llr)) return rc;
    if (method < 0 || method > 2) return fail(h, NRLDPC_EUNSUPPORTED, "Unsupported decision method");
    if (!(variance > 0.0f)) return fail(h, NRLDPC_EUNSUPPORTED, "variance must be positive");
    if (n_bits == 0) return 0;
    ENTER_DEVICE(h);
    const long long n_sym = n_bits / Q_m;
    nrldpc::mod_awgn_demod_kernel<<<grid_for(h, (n_sym + 1) / 2, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        bits, llr, n_sym, Q_m, sqrtf(0.5f * variance), 1.0f / variance, method, seed, stream_id);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
NRLDPC_EXPORT int nrldpc_crc(nrldpc_t *h, const uint8_t *bits, int64_t batch, int32_t n_bits, int64_t stride, int32_t kind,
                             uint8_t *parity, int64_t parity_stride, uint8_t *ok, void *stream) {
    if (!h) return NRLDPC_ESHAPE;
    uint32_t poly; int L;
    switch (kind) {   // get_3gpp_crc_polynomial.m:3-17
        case NRLDPC_CRC16: poly = 0x1021u; L = 16; break;
        case NRLDPC_CRC24A: poly = 0x864CFBu; L = 24; break;
        case NRLDPC_CRC24B: poly = 0x800063u; L = 24; break;
        default: return fail(h, NRLDPC_EUNSUPPORTED, "Invalid CRC identifier.");
    }
    if (batch < 0 || n_bits < 0 || stride < n_bits) return fail(h, NRLDPC_ESHAPE, "batch, n_bits >= 0 and stride >= n_bits required");
    if (batch == 0) return 0;
    if (!bits || (!parity && !ok)) return fail(h, NRLDPC_ESHAPE, "bits and one of parity / ok must not be NULL");
    if (parity && parity_stride < L) return fail(h, NRLDPC_ESHAPE, "parity_stride must be at least the CRC length");
    ENTER_DEVICE(h);
    nrldpc::crc_kernel<<<grid_for(h, batch * 32, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(bits, batch, n_bits, stride, poly, L,
                                                                                             parity, parity_stride, ok);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return 0;
}
