// decode_kernel_refill.cuh -- multi-codeword CTAs under the parity-check stop with PREFETCHED slot refill.
//
// Under 'Parity check satisfied' (the reference's only setting, NRLDPCDecoder.m:120) the codewords of one CTA stop at
// different iterations.  decode_nms_kernel keeps a CTA's codewords together as a group: a codeword that converges early
// idles until the slowest one of its group is done (BASELINE config 3 with the stop: 5.4 mean iterations in the time of
// about 8).  decode_nms_refill_kernel (decode_kernel.cuh) refills every slot on its own but the slot sits out one whole
// pass over the layers while its next codeword's LLRs arrive.  Here the load is taken off the critical path:
//
//   * the CTA owns S codeword slots (Z threads each) and S + P codeword BUFFERS in shared memory; which buffer a slot
//     works on is a per-thread register (Lane::slot_off), so swapping buffers costs nothing inside the layer code;
//   * the P surplus buffers are MAILBOXES: each holds (or is receiving, by one cp.async.bulk / TMA copy that completes
//     on the mailbox's mbarrier) the LLRs of a codeword that no slot has started yet;
//   * at the end of a pass a slot whose codeword is finished (converged or max_iters) writes its outputs, then its
//     leader lane claims a mailbox (shared-memory ticket), waits for that mailbox's copy -- issued at least one whole
//     pass earlier, so the wait is a formality --, takes the mailbox's buffer and hands its old buffer to the mailbox
//     together with a fresh codeword index (device work counter) and a new bulk copy;
//   * the slot's threads clamp the adopted buffer in place (+-LLR_MAX, NaN filler, -0) and join the next pass with
//     iteration 0: no pass is sat out.  At most P slots are refilled per pass end; a slot that finds no mailbox (more
//     than P finished at once) retries at the next pass end.
//
// The passes over the layers stay CTA-wide (one barrier per layer, one instruction stream per CTA); slots differ only in
// the iteration they are in.  Arithmetic and outputs are those of decode_nms_kernel bit for bit
// (tests/test_gpu_parity.py); only the order in which codewords are started differs.
// The host zeroes the work counter before every launch of this kernel (the number of tickets it draws is not known in
// advance: leaders of different slots may both draw a ticket past the end of the batch).
#pragma once
#include "decode_kernel.cuh"

namespace nrldpc {

template <int BG>
__global__ void __launch_bounds__(kDecThreads, kDecCtasPerSm) decode_nms_refill2_kernel(const __grid_constant__ DecArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    enum { ST_RUN = 0, ST_WAIT = 1, ST_IDLE = 2 };
    const int Z = a.Z;
    const int ncw = a.ncols * Z;
    const int K = a.kcols * Z;
    const int S = a.cwpc, P = a.spares;
    float *app = reinterpret_cast<float *>(smem_raw);
    // behind the S + P buffers: flags of the two syndrome stages [2][S], buffer and codeword of every slot [2][S], buffer /
    // codeword / mbarrier parity of every mailbox [3][P], four scalars, the mbarriers [1 + P]
    int *s_flag = reinterpret_cast<int *>(app + (size_t)(S + P) * a.slot_stride);
    int *s_flag2 = s_flag + S;
    int *s_buf = s_flag + 2 * S;
    int *s_cw = s_flag + 3 * S;
    int *mb_buf = s_flag + 4 * S;
    int *mb_cw = mb_buf + P;
    int *mb_par = mb_cw + P;
    int *s_misc = mb_par + P;          // [0] claims of this pass end, [1] first mailbox to claim, [2] mailboxes holding a codeword, [3] batch exhausted
    uint64_t *bars = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(s_misc + 4) + 7) & ~(uintptr_t)7);   // [0] first fill, [1 + m] mailbox m
    if ((uint32_t)__cvta_generic_to_shared(smem_raw) != a.smem_base) __trap();

    const int tid = threadIdx.x;
    const int slot = tid / Z;
    const int z = tid - slot * Z;
    const bool lane_ok = tid < S * Z;
    const int batch = (int)a.batch;              // the host uses this kernel for batches below 2^31 only
    const uint32_t row_bytes = (uint32_t)ncw * 4u;     // a multiple of 16 for every (BG, Z); the host checks slot_stride % 4 == 0
    const uint32_t buf_bytes = (uint32_t)a.slot_stride * 4u;
    const uint32_t col_bytes = (uint32_t)Z * 4u;

    DecCtx c;
    c.l.zoff = (uint32_t)z * 4u;
    c.l.nZ4 = 0u - (uint32_t)Z * 4u;
    c.l.slot_off = (uint32_t)slot * buf_bytes;
    c.l.one = (uint32_t)a.one;
    c.my_rec = a.c2v + (size_t)(blockIdx.x / a.rec_group) * (kRecWords * kRecStride) + (blockIdx.x % a.rec_group) * blockDim.x + tid;
    c.pol = make_l2_policy(a.l2_pin);
    c.last_fail = 0;
    c.cur = make_uint4(0u, 0u, 0u, 0u);
    c.cur2 = c.cur;
    c.ext_lo = c.ext_hi = 0u;

    // ---- first fill: one ticket range for the S slots and the P mailboxes, one bulk copy per codeword
    if (tid == 0) {
        for (int i = 0; i <= P; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int n0 = (int)((unsigned int)atomicAdd(a.work_counter, S + P) - a.work_base);
        const int n_slots = max(0, min(S, batch - n0));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (n_slots > 0) mbar_expect_tx(&bars[0], row_bytes * (uint32_t)n_slots);
        for (int s = 0; s < S; ++s) {
            s_buf[s] = s;
            s_cw[s] = s < n_slots ? n0 + s : -1;
            if (s < n_slots) bulk_g2s_stream(app + (size_t)s * a.slot_stride, a.llr + (long long)(n0 + s) * ncw, row_bytes, &bars[0]);
        }
        int full = 0;
        for (int m = 0; m < P; ++m) {
            const int n = n0 + S + m;
            mb_buf[m] = S + m;
            mb_par[m] = 0;
            mb_cw[m] = n < batch ? n : -1;
            if (n < batch) {
                mbar_expect_tx(&bars[1 + m], row_bytes);
                bulk_g2s_stream(app + (size_t)(S + m) * a.slot_stride, a.llr + (long long)n * ncw, row_bytes, &bars[1 + m]);
                ++full;
            }
        }
        s_misc[0] = 0; s_misc[1] = 0; s_misc[2] = full; s_misc[3] = n0 + S + P >= batch ? 1 : 0;
    }
    if (tid < 2 * S) s_flag[tid] = 0;
    __syncthreads();

    auto clamp_mine = [&](uint32_t p) {   // this thread's share of a landed buffer: position z of every block column, clamped in place
        for (int col = 0; col < a.ncols; ++col, p += col_bytes) sts_f32(p, clamp_llr(lds_f32(p)));
    };

    int st = ST_IDLE, my_cw = -1, my_it = 0;
    if (s_cw[0] >= 0) mbar_wait(&bars[0], 0u);       // CTA-uniform: slot 0 got the first ticket of the range
    if (lane_ok) {
        my_cw = s_cw[slot];
        if (my_cw >= 0) { clamp_mine(a.smem_base + c.l.slot_off + c.l.zoff); st = ST_RUN; }
    }
    if (__syncthreads_and(st == ST_IDLE)) return;

    const bool staged = a.n_rows >= a.staged_min_rows;
    while (true) {
        // one pass over the layers: iteration my_it of every running slot
        c.done = st != ST_RUN;
        UnrolledRows<BG, 0, false, 2>::run(a, c, my_it == 0 ? a.n_rows - 1 : 0, a.n_rows, true);

        const bool was_run = st == ST_RUN;
        const unsigned long long ext_bits = (((unsigned long long)c.ext_hi << 32) | c.ext_lo) << 4;   // see decode_nms_kernel (TRACK)
        if (was_run) {   // last-layer filter: verdict in the second flag, read back before that flag is written again
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
        if (was_run) {
            ++my_it;
            const int ok = (s_flag[slot] | s_flag2[slot]) ? 0 : 1;
            if (ok || my_it == a.max_iters) {
                // outputs of this slot's codeword, written by its own Z threads
                const uint32_t base_s = a.smem_base + c.l.slot_off;
                uint8_t *hard = a.hard + (long long)my_cw * K;
                if ((K & 3) == 0) {
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
                    uint32_t p = base_s + c.l.zoff;
                    for (int col = 0; col < a.ncols; ++col, p += col_bytes, soft += Z) {
                        uint32_t w = lds_u32(p);
                        if (col >= a.kcols + 4 && col < a.kcols + a.n_rows) w = ext_app(c.my_rec, col - a.kcols, c.pol, w);
                        __stcs(soft, __uint_as_float(w));
                    }
                }
                if (z == 0) {
                    if (a.iters) a.iters[my_cw] = my_it;
                    if (a.ok) a.ok[my_cw] = (uint8_t)ok;
                }
                st = ST_WAIT;
            }
        }
        __syncthreads();   // flags read; finished slots' buffers read out; last pass end's claim counter consumed
        if (tid < 2 * S) s_flag[tid] = 0;
        // ---- leaders of finished slots claim a mailbox each (at most P per pass end)
        if (lane_ok && z == 0 && st == ST_WAIT) {
            int next = -2;                                  // -2: no mailbox this time, -1: batch exhausted, >= 0: codeword adopted
            const int k = atomicAdd(&s_misc[0], 1);
            if (k < P) {
                int m = *(volatile int *)&s_misc[1] + k;
                if (m >= P) m -= P;
                const int cw_new = mb_cw[m];
                if (cw_new >= 0) {
                    mbar_wait(&bars[1 + m], (uint32_t)mb_par[m]);
                    mb_par[m] ^= 1;
                    const int ob = s_buf[slot];
                    s_buf[slot] = mb_buf[m];
                    next = cw_new;
                    mb_buf[m] = ob;
                    int n = batch;
                    if (*(volatile int *)&s_misc[3] == 0) n = (int)((unsigned int)atomicAdd(a.work_counter, 1) - a.work_base);
                    if (n < batch) {
                        mb_cw[m] = n;
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slot's reads of its old buffer vs the async write
                        mbar_expect_tx(&bars[1 + m], row_bytes);
                        bulk_g2s_stream(app + (size_t)ob * a.slot_stride, a.llr + (long long)n * ncw, row_bytes, &bars[1 + m]);
                    } else {
                        mb_cw[m] = -1;
                        *(volatile int *)&s_misc[3] = 1;
                        atomicSub(&s_misc[2], 1);
                    }
                } else if (*(volatile int *)&s_misc[2] <= 0) {
                    next = -1;                              // an empty mailbox and none that holds a codeword: this slot is done
                }
            }
            s_cw[slot] = next;
        }
        __syncthreads();
        if (lane_ok && st == ST_WAIT) {
            const int v = s_cw[slot];
            if (v >= 0) {
                my_cw = v;
                c.l.slot_off = (uint32_t)s_buf[slot] * buf_bytes;
                clamp_mine(a.smem_base + c.l.slot_off + c.l.zoff);
                st = ST_RUN; my_it = 0;
                c.cur = make_uint4(0u, 0u, 0u, 0u);
                c.cur2 = c.cur;
                c.ext_lo = c.ext_hi = 0u;
            } else if (v == -1) {
                st = ST_IDLE;
            }
        }
        if (tid == 0) {
            int h = s_misc[1] + min(s_misc[0], P);
            if (h >= P) h -= P;
            s_misc[1] = h;
            s_misc[0] = 0;
        }
        if (__syncthreads_and(st == ST_IDLE)) break;   // also publishes the clamped values, the flag reset and the mailbox cursor
    }
}

}  // namespace nrldpc
