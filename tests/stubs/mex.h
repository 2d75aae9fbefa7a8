/* mex.h -- STUB of the MATLAB MEX API, test infrastructure only.
 *
 * The build image has no MATLAB, so matlab/nrldpc_mex.cpp (the gateway behind the reference's
 * obj.hLDPCDecoder seam, NRLDPCDecoder.m:117-121,265) could never be compiled.  This header declares the small
 * subset of the documented MEX C API the gateway uses, with MATLAB's semantics (column-major storage, mxLogical =
 * one byte, mexErrMsgIdAndTxt does not return); tests/stubs/mex_shim.cpp implements it and exposes mexFunction to
 * the Python tests.  It is NOT MathWorks' header and only exists so that the gateway is compiled and executed.
 */
#ifndef NRLDPC_STUB_MEX_H
#define NRLDPC_STUB_MEX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mxArray_tag mxArray;
typedef bool mxLogical;
typedef size_t mwSize;
typedef enum { mxUNKNOWN_CLASS = 0, mxLOGICAL_CLASS = 3, mxCHAR_CLASS = 4, mxDOUBLE_CLASS = 6, mxSINGLE_CLASS = 7, mxINT8_CLASS = 8,
               mxUINT8_CLASS = 9, mxUINT64_CLASS = 15 } mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

bool mxIsChar(const mxArray *a);
bool mxIsDouble(const mxArray *a);
bool mxIsLogical(const mxArray *a);
bool mxIsSingle(const mxArray *a);
bool mxIsInt8(const mxArray *a);
bool mxIsUint64(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
void *mxGetData(const mxArray *a);
double *mxGetPr(const mxArray *a);
mxLogical *mxGetLogicals(const mxArray *a);
double mxGetScalar(const mxArray *a);
int mxGetString(const mxArray *a, char *buf, mwSize buflen);
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray *mxCreateLogicalMatrix(mwSize m, mwSize n);
void mxDestroyArray(mxArray *a);
#if defined(__GNUC__)
__attribute__((noreturn, format(printf, 2, 3)))
#endif
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

#ifdef __cplusplus
}
#endif
#endif
