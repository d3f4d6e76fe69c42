/*
 * oracle/stub/mex.h -- TEST INFRASTRUCTURE, not product code.
 *
 * A minimal stand-in for MATLAB's <mex.h> (R2018a interleaved-complex API) so that
 *   (a) the reference's two MEX sources compile UNMODIFIED, in place from
 *       /root/reference/mex/, into oracle/_ref/libsph_ref.so (the parity oracle), and
 *   (b) this repo's own MEX shims (sph-poiseuille-flow_b200/mex/) can be compiled and
 *       exercised on hosts that have no MATLAB.
 *
 * It declares exactly the 18 mx / mex entry points the reference uses:
 * mxGetDoubles mxGetScalar mxGetM mxGetN mxGetNumberOfElements mxIsDouble mxIsChar
 * mxGetString mxCreateDoubleMatrix mxCreateDoubleScalar mxCreateString mxDuplicateArray
 * mxDestroyArray mxMalloc mxCalloc mxRealloc mxFree mexErrMsgIdAndTxt
 * plus the types mxArray, mwSize and the mxREAL enumerator.  The implementation is
 * oracle/stub/mexhost.c.
 */
#ifndef SPHB_STUB_MEX_H
#define SPHB_STUB_MEX_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef size_t mwSize;
typedef ptrdiff_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

/* Opaque to MEX code, as in MATLAB.  Layout is private to mexhost.c. */
typedef struct mxArray_tag mxArray;

double *mxGetDoubles(const mxArray *a);
double mxGetScalar(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
int mxIsDouble(const mxArray *a);
int mxIsChar(const mxArray *a);
int mxGetString(const mxArray *a, char *buf, mwSize buflen);

mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity flag); /* zero-filled */
mxArray *mxCreateDoubleScalar(double v);
mxArray *mxCreateString(const char *s);
mxArray *mxDuplicateArray(const mxArray *a);
void mxDestroyArray(mxArray *a);

void *mxMalloc(size_t n);
void *mxCalloc(size_t n, size_t sz);
void *mxRealloc(void *p, size_t n);
void mxFree(void *p);

/* Does not return: unwinds (longjmp) to the enclosing mexhost_call(). */
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);

/* The gateway every MEX file defines. */
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

#ifdef __cplusplus
}
#endif
#endif
