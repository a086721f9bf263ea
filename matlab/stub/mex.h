/* STUB of the MEX C API -- NOT MathWorks' or GNU Octave's header.  This image has neither MATLAB nor
 * Octave (no mex.h / mkoctfile), so admm_b200_mex.c is compile-checked against these declarations
 * only (matlab/Makefile target `check`).  Build the real gateway with `mex` or `mkoctfile --mex`,
 * which supply the real header. */
#ifndef ADMM_B200_STUB_MEX_H
#define ADMM_B200_STUB_MEX_H
#include <stddef.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxUNKNOWN_CLASS = 0, mxDOUBLE_CLASS = 6, mxUINT8_CLASS = 9, mxUINT64_CLASS = 13 } mxClassID;
#ifdef __cplusplus
extern "C" {
#endif
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);
int mexPrintf(const char* fmt, ...);
void mexLock(void);
int mexAtExit(void (*fn)(void));
double* mxGetPr(const mxArray* a);
void* mxGetData(const mxArray* a);
double mxGetScalar(const mxArray* a);
size_t mxGetM(const mxArray* a);
size_t mxGetN(const mxArray* a);
size_t mxGetNumberOfElements(const mxArray* a);
int mxIsDouble(const mxArray* a);
int mxIsComplex(const mxArray* a);
int mxIsSparse(const mxArray* a);
int mxIsChar(const mxArray* a);
int mxIsStruct(const mxArray* a);
int mxIsEmpty(const mxArray* a);
char* mxArrayToString(const mxArray* a);
void mxFree(void* p);
void* mxMalloc(size_t n);
mxClassID mxGetClassID(const mxArray* a);
mxArray* mxGetField(const mxArray* s, mwIndex i, const char* name);
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray* mxCreateDoubleScalar(double v);
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char** names);
void mxSetField(mxArray* s, mwIndex i, const char* name, mxArray* v);
int mxAddField(mxArray* s, const char* name);
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
