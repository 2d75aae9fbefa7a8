/*
 * nrldpc_b200.h -- C ABI of the B200-native 3GPP NR LDPC engine (libnrldpc_b200.so).
 *
 * This is the drop-in boundary for the reference's LDPC hot path.  Each entry point names the
 * reference interface it stands in for (file:line relative to robmaunder/ldpc-3gpp-matlab).
 * Plain C: opaque handle, raw pointers and sizes, integer status codes, no C++/torch types.
 *
 * Conventions shared by every call
 *   - LLR sign: positive => bit 0 (NRLDPCDecoder.m:264 maps known-zero filler bits to +inf).
 *   - "cw layout": the full lifted codeword of n_cw = cols*Z entries (68Z for BG1, 52Z for BG2),
 *     i.e. cw_tilde of NRLDPCDecoder.m:262 -- 2Z punctured systematic positions first.
 *   - bits are one per uint8_t (0/1); batches are row-major [batch][len].
 *   - mem: NRLDPC_MEM_HOST   buffers are host memory; the call copies in/out (pipelined over
 *                            internal streams) and returns when the outputs are valid
 *                            (MATLAB value semantics);
 *          NRLDPC_MEM_DEVICE buffers are device memory on the handle's GPU; work is enqueued
 *                            on `stream` (a cudaStream_t passed as void*, NULL = default
 *                            stream) and the call returns without synchronising.
 *   - status: 0 or a negative NRLDPC_E*; text via nrldpc_last_error().  Never aborts/throws.
 *     NRLDPC_EUNSUPPORTED <-> error('ldpc_3gpp_matlab:UnsupportedParameters',...) (callers catch
 *     and skip: plot_BLER_vs_SNR.m:172-176), NRLDPC_ESHAPE <-> 'ldpc_3gpp_matlab:Error'.
 *   - a handle is not thread-safe; distinct handles are independent (kernel attributes are only ever raised
 *     process-wide, so live handles of different (BG, Z) never invalidate each other's launches).
 *   - NRLDPC_MEM_DEVICE decodes of ONE handle share its scratch: launches on different streams are ordered
 *     behind each other by the library (an event wait), they do not overlap; use one handle per stream for overlap.
 *     A decode captured into a CUDA graph is replayable (it resets its own work counter) but takes no part in that
 *     ordering: do not replay the graph concurrently with other work on the same handle.
 *   - every entry point restores the caller's current CUDA device before it returns.
 */
#ifndef NRLDPC_B200_H
#define NRLDPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRLDPC_OK            0
#define NRLDPC_EUNSUPPORTED (-1)
#define NRLDPC_ESHAPE       (-2)
#define NRLDPC_ECUDA        (-3)
#define NRLDPC_ENOMEM       (-4)

#define NRLDPC_MEM_HOST   0
#define NRLDPC_MEM_DEVICE 1

#define NRLDPC_F32   0
#define NRLDPC_F16X2 1

/* Decoding algorithm (nrldpc_cfg.algorithm).
 *   NRLDPC_ALG_NMS  layered normalized min-sum (default; the fast path)
 *   NRLDPC_ALG_BP   the reference's own algorithm: flooding sum-product in float64 with the termination rule of
 *                   comm.LDPCDecoder as configured at NRLDPCDecoder.m:120.  Decisions and iteration counts equal the
 *                   CPU restatement of MathWorks' documented algorithm (oracle B) and its independent numpy twin;
 *                   equality with the closed toolbox itself is UNVERIFIED here (no MATLAB) -- matlab/make_golden_vectors.m
 *                   produces the vectors that close this, tests/test_matlab_golden.py consumes them when present */
#define NRLDPC_ALG_NMS 0
#define NRLDPC_ALG_BP  1

/* Input LLR magnitudes are clamped to this value on load (NaN and +inf filler -> +LLR_MAX). */
#define NRLDPC_LLR_MAX 1048576.0f

typedef struct nrldpc_handle nrldpc_t;

/* Replaces the name-value constructor comm.LDPCDecoder('ParityCheckMatrix',H,
 * 'MaximumIterationCount',iterations,'IterationTerminationCondition','Parity check satisfied')
 * at NRLDPCDecoder.m:120 and comm.LDPCEncoder('ParityCheckMatrix',H) at NRLDPCEncoder.m:49.
 * H is implied by (bg, Z): H = get_pcm(get_3gpp_base_graph(bg, i_LS(Z)), Z) (NRLDPC.m:433-440). */
typedef struct nrldpc_cfg {
    int32_t bg;         /* 1 or 2 (NRLDPC.m:240-245) */
    int32_t Z;          /* lifting size, one of the 51 of TS 38.212 Table 5.3.2-1 */
    int32_t max_iters;  /* MaximumIterationCount, >= 1 (NRLDPCDecoder.m:41: default 50) */
    int32_t early_term; /* 1 = 'Parity check satisfied' (NRLDPCDecoder.m:120), 0 = 'Maximum iteration count' */
    float   alpha;      /* min-sum normalisation; <= 0 selects the default 0.75 */
    int32_t device;     /* CUDA device ordinal, -1 = current device */
    int32_t llr_dtype;  /* decoder arithmetic: NRLDPC_F32 (default) or NRLDPC_F16X2 (two codewords per thread in
                           packed fp16, inputs clamped to +-2048; the buffers of nrldpc_decode stay float32) */
    int32_t algorithm;  /* NRLDPC_ALG_NMS (0, default) or NRLDPC_ALG_BP; NRLDPC_ALG_BP requires llr_dtype = NRLDPC_F32
                           and ignores alpha */
} nrldpc_cfg;

int  nrldpc_create(nrldpc_t **out, const nrldpc_cfg *cfg);
void nrldpc_destroy(nrldpc_t *h);                      /* release(obj) */
const char *nrldpc_last_error(const nrldpc_t *h);      /* h == NULL: last create() failure */
int  nrldpc_synchronize(nrldpc_t *h);                  /* waits for all work enqueued through h */

/* Geometry of the handle's code: K = kcols*Z, N = (cols-2)*Z (NRLDPC.m:414-454), n_cw = cols*Z. */
typedef struct nrldpc_dims {
    int32_t bg, Z, i_LS, rows, cols, kcols, edges, K, N, n_cw;
} nrldpc_dims;
int nrldpc_get_dims(const nrldpc_t *h, nrldpc_dims *out);

/* Table 5.3.2-1 helpers: get_3gpp_set_index.m:1-13 and get_3gpp_lifting_size.m:1-17.
 * Return NRLDPC_EUNSUPPORTED where the reference raises UnsupportedParameters. */
int nrldpc_set_index(int32_t Z);
int nrldpc_lifting_size(int32_t K_b, int32_t K_prime);

/* get_3gpp_base_graph.m:1-534 -- base-graph edges sorted by (row, col) with raw shifts of set
 * i_LS.  Any output may be NULL.  Returns the edge count (316 / 197) or a negative status. */
int nrldpc_base_graph(int32_t bg, int32_t i_LS, int32_t *rows, int32_t *cols, int32_t *shifts);

/* ---- decode: replaces step(obj.hLDPCDecoder, cw_tilde) at NRLDPCDecoder.m:265 ----------------
 * Layered normalized min-sum (float32, or packed fp16 with llr_dtype = NRLDPC_F16X2) -- or, with algorithm =
 * NRLDPC_ALG_BP, the reference's flooding sum-product in float64 -- over base rows 0..n_rows-1 (n_rows = 0 -> all
 * rows; 4 <= n_rows).  Rows whose parity bit was not transmitted carry zero LLR and contribute nothing,
 * so callers may trim them (DESIGN.md "active rows").
 *   llr       [batch][n_cw] float32, cw layout; +inf / NaN = filler; exact 0 = punctured / unsent
 *   info_hard [batch][K] uint8, hard decision (app < 0) of the information part (required; device
 *             pointers must be 4-byte aligned)
 *   app_soft  [batch][n_cw] float32 a-posteriori LLRs after the last iteration (nullable)
 *   iters     [batch] int32 iterations executed (nullable)
 *   parity_ok [batch] uint8, 1 if every check of the active rows is satisfied (nullable) */
int nrldpc_decode(nrldpc_t *h, const float *llr, int64_t batch, int32_t n_rows,
                  uint8_t *info_hard, float *app_soft, int32_t *iters, uint8_t *parity_ok,
                  int32_t mem, void *stream);

/* Same call with the LLRs transported as IEEE binary16 (half the host<->device bytes): the values are widened
 * exactly to float32 on the device and decoded as by nrldpc_decode.  With llr_dtype = NRLDPC_F16X2 the result
 * is bit-identical to nrldpc_decode on float32 LLRs that were rounded to binary16 (round to nearest even). */
int nrldpc_decode16(nrldpc_t *h, const uint16_t *llr_f16, int64_t batch, int32_t n_rows,
                    uint8_t *info_hard, float *app_soft, int32_t *iters, uint8_t *parity_ok,
                    int32_t mem, void *stream);

/* Same call with the LLRs transported as 8-bit integers (a quarter of the host<->device bytes of float32; what the
 * host-memory call is bound by is PCIe): llr = scale * q for q in [-128, 126], q = 127 marks a filler / known-zero
 * position (+inf, NRLDPCDecoder.m:264), q = 0 a punctured / unsent one.  The values are widened to float32 on the device
 * (one rounded multiplication) and decoded as by nrldpc_decode: the result is bit-identical to nrldpc_decode on the
 * float32 values scale * q.  Quantising the LLRs is the CALLER's decision and costs BLER (receivers use 6-8 bits);
 * nrldpc_decode on float32 stays the reference-facing call. */
int nrldpc_decode8(nrldpc_t *h, const int8_t *llr_q, float scale, int64_t batch, int32_t n_rows,
                   uint8_t *info_hard, float *app_soft, int32_t *iters, uint8_t *parity_ok,
                   int32_t mem, void *stream);

/* Same call on float64 buffers, the type the reference hands to step() (cw_tilde is double, NRLDPCDecoder.m:262).
 * With NRLDPC_ALG_BP the doubles are decoded as they are (float64 arithmetic, app_soft in float64); with the
 * default algorithm they are rounded to float32 on the device and app_soft must be NULL. */
int nrldpc_decode64(nrldpc_t *h, const double *llr_f64, int64_t batch, int32_t n_rows,
                    uint8_t *info_hard, double *app_soft, int32_t *iters, uint8_t *parity_ok,
                    int32_t mem, void *stream);

/* ---- encode: replaces step(obj.hLDPCEncoder, c) at NRLDPCEncoder.m:158 ------------------------
 *   info [batch][K] uint8 (filler positions must already be 0, NRLDPCEncoder.m:153)
 *   cw   [batch][n_cw] uint8 systematic codeword [info ; parity], H*cw = 0 */
int nrldpc_encode(nrldpc_t *h, const uint8_t *info, int64_t batch, uint8_t *cw,
                  int32_t mem, void *stream);

/* Per-code-block rate-matching geometry: the scalars NRLDPCEncoder.bit_selection /
 * NRLDPCDecoder.bit_selection read from the NRLDPC getters (NRLDPCDecoder.m:201-208). */
typedef struct nrldpc_rm {
    int32_t E;        /* E_r  (NRLDPC.m:485-507) */
    int32_t k_0;      /* NRLDPC.m:510-543 */
    int32_t N_cb;     /* NRLDPC.m:463-469 */
    int32_t K_prime;  /* NRLDPC.m:380-382; filler = d positions [max(K'-2Z,0), K-2Z) */
    int32_t Q_m;      /* 1,2,4,6,8 (NRLDPC.m:278-283); E % Q_m == 0 */
} nrldpc_rm;

/* ---- rate match: replaces NRLDPCEncoder.bit_selection + bit_interleaving
 * (NRLDPCEncoder.m:168-225) applied to d = cw(2Z+1:end) with filler skipped (:155,:190).
 *   cw [batch][n_cw] uint8 -> f [batch][E] uint8 */
int nrldpc_rate_match(nrldpc_t *h, const uint8_t *cw, int64_t batch, const nrldpc_rm *rm,
                      uint8_t *f, int32_t mem, void *stream);

/* ---- rate recover: replaces NRLDPCDecoder.bit_interleaving + bit_selection + the cw_tilde
 * assembly of LDPC_coding (NRLDPCDecoder.m:172-242 and :262-264) in one pass.
 *   f       [batch][E] float32 demodulator LLRs
 *   harq    [batch][N] float32 d_tilde_buffer (NRLDPCDecoder.m:236-239), read-modify-written over
 *           [0, N_cb); nullable (I_HARQ = 0)
 *   llr_cw  [batch][n_cw] float32 decoder input: 2Z zeros, soft-combined LLRs, +inf at filler */
int nrldpc_rate_recover(nrldpc_t *h, const float *f, int64_t batch, const nrldpc_rm *rm,
                        float *harq, float *llr_cw, int32_t mem, void *stream);

/* ---- channel leg of plot_BLER_vs_SNR.m:129-132 on device (QPSK only): Philox information bits
 * are NOT generated here; this maps f bits to TS 38.211 QPSK (NRModulator.m:75), adds complex
 * AWGN of total variance `variance` (comm.AWGNChannel, plot_BLER_vs_SNR.m:50,105) from a
 * counter-based generator keyed by (seed, stream_id, bit pair index), and demaps with the exact
 * LLR 2*sqrt(2)*y/variance (NRDemodulator.m:78; Variance as plot_BLER_vs_SNR.m:106).
 *   f_bits [batch][E] uint8 (E even) -> f_llr [batch][E] float32.  Device memory only. */
int nrldpc_qpsk_awgn_llr(nrldpc_t *h, const uint8_t *f_bits, int64_t batch, int32_t E,
                         float variance, uint64_t seed, uint64_t stream_id, float *f_llr,
                         void *stream);

/* The same channel leg fused into rate recovery: f_bits [batch][E] -> decoder input llr_cw [batch][n_cw] (and the HARQ
 * buffer), bit-identical to nrldpc_qpsk_awgn_llr followed by nrldpc_rate_recover with the same (seed, stream_id), without
 * the E received LLRs of every block going through HBM (plot_BLER_vs_SNR.m:130-133 up to NRLDPCDecoder's
 * bit_selection, :200-242).  QPSK (rm->Q_m = 2), E % 4 == 0, E floats must fit in shared memory.  Device memory only. */
int nrldpc_qpsk_awgn_rate_recover(nrldpc_t *h, const uint8_t *f_bits, int64_t batch, const nrldpc_rm *rm, float variance,
                                  uint64_t seed, uint64_t stream_id, float *harq, float *llr_cw, void *stream);

/* ---- channel leg for every modulation of the reference (device memory only) --------------------
 * Q_m = 1, 2, 4, 6, 8 <-> 'BPSK', 'QPSK', '16QAM', '64QAM', '256QAM' (NRModulator.m:47-63); bits are taken
 * Q_m at a time, first bit = b0 of TS 38.211 section 5.1; symbols are interleaved (re, im) float32 pairs.
 *   nrldpc_modulate    step(hMod, bits)   NRModulator.m:69-89   (the toolbox CustomSymbolMapping vectors at
 *                      :73-81 are the TS 38.211 maps, which is what is computed here)
 *   nrldpc_awgn        step(hChan, tx)    comm.AWGNChannel, plot_BLER_vs_SNR.m:50,105,131: adds complex noise of
 *                      total variance `variance` in place, counter-based generator keyed (seed, stream_id, index)
 *   nrldpc_demodulate  step(hDemod, rx)   NRDemodulator.m:72-96; method = NRLDPC_DEMOD_LLR ('Log-likelihood
 *                      ratio', exact), NRLDPC_DEMOD_APPROX ('Approximate log-likelihood ratio', max-log) or
 *                      NRLDPC_DEMOD_HARD ('Hard decision', outputs 0.0 / 1.0); `variance` is the Variance property
 *   nrldpc_mod_awgn_llr  the three fused, bit-identical to calling them in sequence with the same keys */
#define NRLDPC_DEMOD_LLR    0
#define NRLDPC_DEMOD_APPROX 1
#define NRLDPC_DEMOD_HARD   2
int nrldpc_modulate(nrldpc_t *h, const uint8_t *bits, int64_t n_bits, int32_t Q_m, float *sym, void *stream);
int nrldpc_awgn(nrldpc_t *h, float *sym, int64_t n_sym, float variance, uint64_t seed, uint64_t stream_id, void *stream);
int nrldpc_demodulate(nrldpc_t *h, const float *sym, int64_t n_sym, int32_t Q_m, float variance, int32_t method,
                      float *llr, void *stream);
int nrldpc_mod_awgn_llr(nrldpc_t *h, const uint8_t *bits, int64_t n_bits, int32_t Q_m, float variance, int32_t method,
                        uint64_t seed, uint64_t stream_id, float *llr, void *stream);

/* ---- CRC attach / check on device: comm.CRCGenerator / comm.CRCDetector with the polynomials of
 * get_3gpp_crc_polynomial.m:3-17 as used at NRLDPCEncoder.m:70-89,114 and NRLDPCDecoder.m:300,336.
 *   bits   [batch] rows of n_bits bits (one per byte), row b at bits + b*stride  (device memory)
 *   parity nullable: the L parity bits of each row are written at parity + b*parity_stride (may alias the
 *          tail of the same row buffer: parity = bits + n_bits, parity_stride = stride)
 *   ok     nullable [batch]: 1 iff the remainder of the n_bits bits is zero (row with its parity attached passes) */
#define NRLDPC_CRC16  0
#define NRLDPC_CRC24A 1
#define NRLDPC_CRC24B 2
int nrldpc_crc(nrldpc_t *h, const uint8_t *bits, int64_t batch, int32_t n_bits, int64_t stride, int32_t kind,
               uint8_t *parity, int64_t parity_stride, uint8_t *ok, void *stream);

/* ---- random information blocks of the Monte-Carlo loop on device: a = round(rand(A,1)) (plot_BLER_vs_SNR.m:112).
 * bits[r*stride + k] for r < rows, k < n_bits (one bit per byte, device memory; nothing beyond n_bits of a row is written)
 * from the counter-based generator keyed (seed, stream_id, index of the row's 128-bit group): the same bits whatever the launch
 * geometry, independent streams per (seed, stream_id) as the script asks of parallel runs (:23-27). */
int nrldpc_random_bits(nrldpc_t *h, uint8_t *bits, int64_t rows, int32_t n_bits, int64_t stride, uint64_t seed, uint64_t stream_id,
                       void *stream);

/* ---- block-error bookkeeping of the Monte-Carlo loop on device (plot_BLER_vs_SNR.m:139-155; a_hat = [] unless the CRCs
 * pass, NRLDPCDecoder.m:296-309,336-339; a block error is ~isequal(a, a_hat)), one decoding attempt of n_tb transport
 * blocks of C code blocks each.  All buffers are device memory.
 *   do_latch: ok = tb_ok[b] && all cb_passed[b*C..] (NULL: C = 1) && tb_hat[b][0:A] == tb[b][0:A]; latch[b] |= ok;
 *             counters[3] += iterations of the attempt's C decodes (iters [n_tb*C])
 *   finalize: counters[0] += 1; counters[1] += !latch[b]; counters[2] += wrong bits among the first K_prime bits of the C
 *             decoded blocks (hard vs info, both [n_tb*C][K]) of a transport block in error
 *   counters  [4] uint64: {blocks, block errors, bit errors, iterations}: what the ranks sum with one all-reduce */
int nrldpc_bler_count(nrldpc_t *h, const uint8_t *hard, const uint8_t *info, const uint8_t *tb_hat, int64_t tb_hat_stride,
                      const uint8_t *tb, int64_t tb_stride, const uint8_t *tb_ok, const uint8_t *cb_passed, const int32_t *iters,
                      int64_t n_tb, int32_t C, int32_t K_prime, int32_t A, uint8_t *latch, uint64_t *counters, int32_t do_latch,
                      int32_t finalize, void *stream);

/* Pinned host allocations for callers that want truly asynchronous NRLDPC_MEM_HOST transfers. */
void *nrldpc_host_alloc(uint64_t bytes);
void  nrldpc_host_free(void *p);

/* Launch bookkeeping for benchmarks: kernels launched through this handle since creation. */
int64_t nrldpc_launch_count(const nrldpc_t *h);

const char *nrldpc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NRLDPC_B200_H */
