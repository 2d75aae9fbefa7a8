// nrldpc_mex.cpp -- thin MEX gateway over the C ABI (include/nrldpc_b200.h).
//
// Status: compiled and EXECUTED in this repository against a stub of the MEX API (tests/stubs/mex.h + mex_shim.cpp,
// built by __graft_entry__.build()); tests/test_mex_gateway.py drives create / decode / encode / destroy through
// mexFunction on the GPU box and checks the column-major layout, the logical packing and both error identifiers.  It has
// not been run under a real MATLAB (none in the build image).  It is deliberately thin: every behaviour lives behind
// the C ABI.  Build (on a machine with MATLAB + CUDA):
//     mex -I../include nrldpc_mex.cpp -L../ldpc_3gpp_matlab_b200 -lnrldpc_b200
//
// Usage from MATLAB (see B200LDPCDecoder.m / B200LDPCEncoder.m):
//     h     = nrldpc_mex('create', BG, Z, max_iters, early_term, alpha, llr_dtype, algorithm);
//             algorithm: 0 = layered normalized min-sum (default), 1 = the reference's flooding sum-product in float64
//     c_hat = nrldpc_mex('decode', h, cw_tilde, n_rows);   % cw_tilde: (68Z or 52Z) x batch double, +inf = filler
//             cw_tilde may also be SINGLE (half the host bytes; min-sum modes) or INT8 with a scale as fifth argument:
//             c_hat = nrldpc_mex('decode', h, int8(q), n_rows, scale)  % llr = scale*q, q = 127 marks filler (a quarter of
//             the PCIe bytes of float32: nrldpc_decode8)
//     [c_hat, num_iters, parity_ok] = nrldpc_mex('decode', ...)   % as comm.LDPCDecoder's NumIterationsOutputPort /
//                                                                 % FinalParityChecksOutputPort (1 x batch each)
//     cw    = nrldpc_mex('encode', h, c);                  % c: K x batch, values 0/1
//     nrldpc_mex('destroy', h);
// Error codes are mapped onto the reference's identifiers: NRLDPC_EUNSUPPORTED ->
// 'ldpc_3gpp_matlab:UnsupportedParameters' (callers catch and skip, plot_BLER_vs_SNR.m:172-176),
// everything else -> 'ldpc_3gpp_matlab:Error'.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "mex.h"
#include "nrldpc_b200.h"

static void check(int rc, const nrldpc_t *h) {
    if (rc == NRLDPC_OK) return;
    const char *msg = nrldpc_last_error(h);
    if (rc == NRLDPC_EUNSUPPORTED) mexErrMsgIdAndTxt("ldpc_3gpp_matlab:UnsupportedParameters", "%s", msg);
    mexErrMsgIdAndTxt("ldpc_3gpp_matlab:Error", "%s (nrldpc rc=%d)", msg, rc);
}

static nrldpc_t *handle_of(const mxArray *a) {
    if (!mxIsUint64(a) || mxGetNumberOfElements(a) != 1) mexErrMsgIdAndTxt("ldpc_3gpp_matlab:Error", "bad handle");
    return reinterpret_cast<nrldpc_t *>(*static_cast<uint64_t *>(mxGetData(a)));
}

static void need(int nrhs, int n, const char *usage) {
    if (nrhs < n) mexErrMsgIdAndTxt("ldpc_3gpp_matlab:Error", "usage: nrldpc_mex(%s)", usage);
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("ldpc_3gpp_matlab:Error", "first argument must be a command string");
    char cmd[32];
    mxGetString(prhs[0], cmd, sizeof(cmd));

    if (!strcmp(cmd, "create")) {
        need(nrhs, 5, "'create', BG, Z, max_iters, early_term [, alpha, llr_dtype, algorithm]");
        nrldpc_cfg cfg{};
        cfg.bg = (int32_t)mxGetScalar(prhs[1]);
        cfg.Z = (int32_t)mxGetScalar(prhs[2]);
        cfg.max_iters = (int32_t)mxGetScalar(prhs[3]);
        cfg.early_term = (int32_t)mxGetScalar(prhs[4]);
        cfg.alpha = nrhs > 5 ? (float)mxGetScalar(prhs[5]) : 0.75f;
        cfg.device = -1;
        cfg.llr_dtype = nrhs > 6 ? (int32_t)mxGetScalar(prhs[6]) : NRLDPC_F32;   /* NRLDPC_F16X2 = packed-half decoder */
        cfg.algorithm = nrhs > 7 ? (int32_t)mxGetScalar(prhs[7]) : NRLDPC_ALG_NMS; /* NRLDPC_ALG_BP = comm.LDPCDecoder's own algorithm */
        nrldpc_t *h = nullptr;
        check(nrldpc_create(&h, &cfg), nullptr);
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *static_cast<uint64_t *>(mxGetData(plhs[0])) = reinterpret_cast<uint64_t>(h);
        return;
    }
    if (!strcmp(cmd, "destroy")) {
        need(nrhs, 2, "'destroy', h");
        nrldpc_destroy(handle_of(prhs[1]));
        return;
    }
    if (strcmp(cmd, "decode") && strcmp(cmd, "encode")) mexErrMsgIdAndTxt("ldpc_3gpp_matlab:Error", "unknown command '%s'", cmd);
    need(nrhs, 3, "'decode' | 'encode', h, data [, n_rows]");
    nrldpc_t *h = handle_of(prhs[1]);
    nrldpc_dims d;
    nrldpc_get_dims(h, &d);

    if (!strcmp(cmd, "decode")) {
        // MATLAB passes a column-major (n_cw x batch) double matrix: each column is one cw_tilde
        // (NRLDPCDecoder.m:262-264), i.e. exactly the row-major [batch][n_cw] layout of the C ABI.
        const mxArray *in = prhs[2];
        if (!(mxIsDouble(in) || mxIsSingle(in) || mxIsInt8(in)) || (int)mxGetM(in) != d.n_cw)
            mexErrMsgIdAndTxt("ldpc_3gpp_matlab:Error", "cw_tilde should be a double, single or int8 matrix with %d rows.", d.n_cw);
        const int64_t batch = (int64_t)mxGetN(in);
        const int n_rows = nrhs > 3 ? (int)mxGetScalar(prhs[3]) : 0;
        // the doubles go to the library as they are (nrldpc_decode64, ordinary pageable MATLAB memory): the sum-product
        // mode computes on them in float64 like comm.LDPCDecoder; for the min-sum mode the library narrows them to
        // float32 on host threads into a pinned staging ring (+inf stays +inf, NaN = filler too), so PCIe carries 4 bytes
        // per LLR -- no conversion loop on the MATLAB thread.  The decisions are written straight into the logical matrix
        // (mxLogical is one byte holding 0 / 1, which is what the library stores).
        plhs[0] = mxCreateLogicalMatrix(d.K, batch);  // comm.LDPCDecoder returns logical K x 1 (NRLDPCDecoder.m:265)
        static_assert(sizeof(mxLogical) == 1, "mxLogical is one byte");
        uint8_t *hard = reinterpret_cast<uint8_t *>(mxGetLogicals(plhs[0]));
        std::vector<int32_t> iters(nlhs > 1 ? (size_t)batch : 0);
        std::vector<uint8_t> ok(nlhs > 2 ? (size_t)batch : 0);
        int32_t *it_p = nlhs > 1 ? iters.data() : nullptr;
        uint8_t *ok_p = nlhs > 2 ? ok.data() : nullptr;
        if (mxIsDouble(in)) {
            check(nrldpc_decode64(h, mxGetPr(in), batch, n_rows, hard, nullptr, it_p, ok_p, NRLDPC_MEM_HOST, nullptr), h);
        } else if (mxIsSingle(in)) {       // single(cw_tilde): the C ABI's own type, copied into the pinned ring by host threads
            check(nrldpc_decode(h, static_cast<const float *>(mxGetData(in)), batch, n_rows, hard, nullptr, it_p, ok_p, NRLDPC_MEM_HOST, nullptr), h);
        } else {                           // int8 codes with a scale: llr = scale * q, 127 = filler
            if (nrhs < 5) mexErrMsgIdAndTxt("ldpc_3gpp_matlab:Error", "usage: nrldpc_mex('decode', h, int8(q), n_rows, scale)");
            check(nrldpc_decode8(h, static_cast<const int8_t *>(mxGetData(in)), (float)mxGetScalar(prhs[4]), batch, n_rows, hard, nullptr,
                                 it_p, ok_p, NRLDPC_MEM_HOST, nullptr), h);
        }
        if (nlhs > 1) {
            plhs[1] = mxCreateDoubleMatrix(1, batch, mxREAL);
            for (int64_t i = 0; i < batch; ++i) mxGetPr(plhs[1])[i] = iters[i];
        }
        if (nlhs > 2) {
            plhs[2] = mxCreateLogicalMatrix(1, batch);
            for (int64_t i = 0; i < batch; ++i) mxGetLogicals(plhs[2])[i] = ok[i] != 0;
        }
        return;
    }
    if (!strcmp(cmd, "encode")) {
        const mxArray *in = prhs[2];
        if ((int)mxGetM(in) != d.K) mexErrMsgIdAndTxt("ldpc_3gpp_matlab:Error", "c should have K=%d rows.", d.K);
        const int64_t batch = (int64_t)mxGetN(in);
        std::vector<uint8_t> info((size_t)batch * d.K), cw((size_t)batch * d.n_cw);
        if (mxIsDouble(in)) { const double *x = mxGetPr(in); for (size_t i = 0; i < info.size(); ++i) info[i] = x[i] != 0.0; }
        else if (mxIsLogical(in)) { const mxLogical *x = mxGetLogicals(in); for (size_t i = 0; i < info.size(); ++i) info[i] = x[i]; }
        else mexErrMsgIdAndTxt("ldpc_3gpp_matlab:Error", "c must be double or logical.");
        check(nrldpc_encode(h, info.data(), batch, cw.data(), NRLDPC_MEM_HOST, nullptr), h);
        plhs[0] = mxCreateDoubleMatrix(d.n_cw, batch, mxREAL);  // comm.LDPCEncoder keeps the input type (NRLDPCEncoder.m:158)
        double *o = mxGetPr(plhs[0]);
        for (size_t i = 0; i < cw.size(); ++i) o[i] = cw[i];
        return;
    }
}
