/* Minimal stand-in for MATLAB's mex.h: just enough declarations to type-check matlab/bds_mex.c
 * in an image without MATLAB (tests/test_abi.py).  Not a MATLAB header. */
#ifndef STUB_MEX_H
#define STUB_MEX_H
#include <stddef.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxINT8_CLASS = 8, mxINT32_CLASS = 12, mxDOUBLE_CLASS = 6 } mxClassID;
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);
void mexLock(void);
int mexAtExit(void (*fn)(void));
int mxIsInt8(const mxArray*);
double mxGetScalar(const mxArray*);
double* mxGetDoubles(const mxArray*);
int8_t* mxGetInt8s(const mxArray*);
int mxIsComplex(const mxArray*);
void* mxGetComplexInt8s(const mxArray*);
void* mxGetData(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
size_t mxGetM(const mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
void mxDestroyArray(mxArray*);
void* mxMalloc(size_t);
void* mxCalloc(size_t, size_t);
void mxFree(void*);
#endif
