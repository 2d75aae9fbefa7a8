// mex_shim.cpp -- implementation of tests/stubs/mex.h plus a C entry point through which the Python tests call the
// gateway's mexFunction the way MATLAB would (prhs/plhs arrays of mxArray*).  TEST INFRASTRUCTURE ONLY.
// mexErrMsgIdAndTxt is "does not return" in MATLAB (it raises a MATLAB exception); here it throws a C++ exception that
// shim_mex catches and reports as (identifier, message), so the tests can check the identifiers the reference's callers
// catch (plot_BLER_vs_SNR.m:172-176: 'ldpc_3gpp_matlab:UnsupportedParameters').
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "mex.h"

struct mxArray_tag {
    mxClassID cls;
    size_t m, n, elt;
    void *data;
};

namespace {
struct MexError {
    std::string id, msg;
};
mxArray *make(size_t m, size_t n, mxClassID cls, size_t elt) {
    mxArray *a = new mxArray_tag{cls, m, n, elt, nullptr};
    a->data = calloc((m * n) > 0 ? m * n : 1, elt);
    return a;
}
}  // namespace

extern "C" {
bool mxIsChar(const mxArray *a) { return a && a->cls == mxCHAR_CLASS; }
bool mxIsDouble(const mxArray *a) { return a && a->cls == mxDOUBLE_CLASS; }
bool mxIsLogical(const mxArray *a) { return a && a->cls == mxLOGICAL_CLASS; }
bool mxIsUint64(const mxArray *a) { return a && a->cls == mxUINT64_CLASS; }
bool mxIsSingle(const mxArray *a) { return a && a->cls == mxSINGLE_CLASS; }
bool mxIsInt8(const mxArray *a) { return a && a->cls == mxINT8_CLASS; }
size_t mxGetM(const mxArray *a) { return a->m; }
size_t mxGetN(const mxArray *a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray *a) { return a->m * a->n; }
void *mxGetData(const mxArray *a) { return a->data; }
double *mxGetPr(const mxArray *a) { return a->cls == mxDOUBLE_CLASS ? static_cast<double *>(a->data) : nullptr; }
mxLogical *mxGetLogicals(const mxArray *a) { return a->cls == mxLOGICAL_CLASS ? static_cast<mxLogical *>(a->data) : nullptr; }
double mxGetScalar(const mxArray *a) {
    switch (a->cls) {
        case mxDOUBLE_CLASS: return *static_cast<double *>(a->data);
        case mxLOGICAL_CLASS: return *static_cast<mxLogical *>(a->data) ? 1.0 : 0.0;
        case mxUINT64_CLASS: return (double)*static_cast<uint64_t *>(a->data);
        case mxUINT8_CLASS: return (double)*static_cast<uint8_t *>(a->data);
        case mxSINGLE_CLASS: return (double)*static_cast<float *>(a->data);
        case mxINT8_CLASS: return (double)*static_cast<int8_t *>(a->data);
        default: return 0.0;
    }
}
int mxGetString(const mxArray *a, char *buf, mwSize buflen) {
    if (!mxIsChar(a) || !buflen) return 1;
    const size_t n = a->m * a->n;                       // stored here as one byte per character
    const size_t c = n < buflen - 1 ? n : buflen - 1;
    memcpy(buf, a->data, c);
    buf[c] = 0;
    return n >= buflen;
}
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity) {
    const size_t elt = cls == mxDOUBLE_CLASS || cls == mxUINT64_CLASS ? 8 : cls == mxSINGLE_CLASS ? 4 : 1;
    return make(m, n, cls, elt);
}
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity) { return make(m, n, mxDOUBLE_CLASS, 8); }
mxArray *mxCreateLogicalMatrix(mwSize m, mwSize n) { return make(m, n, mxLOGICAL_CLASS, sizeof(mxLogical)); }
void mxDestroyArray(mxArray *a) {
    if (!a) return;
    free(a->data);
    delete a;
}
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw MexError{id, buf};
}

// ---- what the Python tests call ---------------------------------------------------------------------------------
mxArray *shim_string(const char *s) {
    mxArray *a = make(1, strlen(s), mxCHAR_CLASS, 1);
    memcpy(a->data, s, strlen(s));
    return a;
}
mxArray *shim_double(const double *colmajor, size_t m, size_t n) {   // copies, like passing a MATLAB value
    mxArray *a = make(m, n, mxDOUBLE_CLASS, 8);
    if (colmajor) memcpy(a->data, colmajor, m * n * 8);
    return a;
}
mxArray *shim_logical(const uint8_t *colmajor, size_t m, size_t n) {
    mxArray *a = make(m, n, mxLOGICAL_CLASS, sizeof(mxLogical));
    for (size_t i = 0; i < m * n; ++i) static_cast<mxLogical *>(a->data)[i] = colmajor[i] != 0;
    return a;
}
mxArray *shim_uint8(const uint8_t *colmajor, size_t m, size_t n) {
    mxArray *a = make(m, n, mxUINT8_CLASS, 1);
    memcpy(a->data, colmajor, m * n);
    return a;
}
mxArray *shim_single(const float *colmajor, size_t m, size_t n) {
    mxArray *a = make(m, n, mxSINGLE_CLASS, 4);
    memcpy(a->data, colmajor, m * n * 4);
    return a;
}
mxArray *shim_int8(const int8_t *colmajor, size_t m, size_t n) {
    mxArray *a = make(m, n, mxINT8_CLASS, 1);
    memcpy(a->data, colmajor, m * n);
    return a;
}
int shim_class(const mxArray *a) { return (int)a->cls; }
size_t shim_elt(const mxArray *a) { return a->elt; }

// mexFunction under try/catch: 0 = returned normally, 1 = mexErrMsgIdAndTxt was raised (identifier / text copied out)
int shim_mex(int nlhs, mxArray **plhs, int nrhs, mxArray **prhs, char *err_id, char *err_msg, size_t cap) {
    try {
        mexFunction(nlhs, plhs, nrhs, const_cast<const mxArray **>(prhs));
        return 0;
    } catch (const MexError &e) {
        snprintf(err_id, cap, "%s", e.id.c_str());
        snprintf(err_msg, cap, "%s", e.msg.c_str());
        return 1;
    }
}
}
