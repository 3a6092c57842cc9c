/*
 * blas.h — BLAS-3 surface of the B200 Cholesky library (device-pointer `cu*` and host-pointer
 * `cuMultiGPU*` families).
 *
 * Drop-in for the GPU half of reference include/blas.h: same enum values (blas.h:24-27), same
 * `xerbla` hook (blas.h:37-43), same handle life-cycle (blas.h:156-158) and the same signatures
 * for cu{S,D,C,Z}{gemm2,syrk|herk,trmm2,trsm} (blas.h:160-253) and cuMultiGPU{S,D,C,Z}{gemm,
 * syrk|herk,trmm,trsm} (blas.h:255-360).  Column-major, leading dimensions in ELEMENTS, complex
 * scalars are C99 complex passed by value.
 *
 * Not exported: the reference's un-prefixed CPU routines (sgemm ... ztrsm, blas.h:45-153).  This
 * library has no CPU path; those names live only in the test oracle (oracle/).
 * Every cu* routine is asynchronous on the CUstream argument.
 */
#ifndef CHOL_B200_BLAS_H
#define CHOL_B200_BLAS_H

#include <stddef.h>
#include <stdbool.h>
#include <cuda.h>

#include "cumultigpu.h"

/* Complex scalars are passed BY VALUE as C99 complex (blas.h:79-87).  For C++ translation units
 * the layout- and ABI-compatible CUDA vector structs stand in (a struct of two floats/doubles is
 * classified exactly like the C99 complex type by the x86-64 and AArch64 calling conventions). */
#ifdef __cplusplus
#include <cuComplex.h>
typedef cuFloatComplex chol_cfloat;
typedef cuDoubleComplex chol_cdouble;
extern "C" {
#else
#include <complex.h>
typedef float complex chol_cfloat;
typedef double complex chol_cdouble;
#endif

/* BLAS enums - castable to char for compatibility with Fortran BLAS/LAPACK (blas.h:24-27). */
typedef enum { CBlasNoTrans = 'N', CBlasTrans = 'T', CBlasConjTrans = 'C' } CBlasTranspose;
typedef enum { CBlasLower = 'L', CBlasUpper = 'U' } CBlasUplo;
typedef enum { CBlasLeft = 'L', CBlasRight = 'R' } CBlasSide;
typedef enum { CBlasNonUnit = 'N', CBlasUnit = 'U' } CBlasDiag;

/* Argument-error hook: called as xerbla(function name, 1-based argument position) before a
 * routine returns CUDA_ERROR_INVALID_VALUE; may be NULL (blas.h:37-43, blas/xerbla.c:5-9). */
typedef void (*xerbla_t)(const char *, long);
extern xerbla_t xerbla;
#define XERBLA(info)                        \
  do {                                      \
    if (xerbla != NULL) xerbla(__func__, info); \
  } while (false)

/* ---- single-GPU handle (blas/handle.c:84-107): adopts the calling thread's current context,
 *      or retains device 0's primary context if there is none ---- */
typedef struct __cublashandle_st *CUBLAShandle;
CUresult cuBLASCreate(CUBLAShandle *);
CUresult cuBLASDestroy(CUBLAShandle);

/* ---- s: C_tri = alpha*op(A)*op(A)^T + beta*C_tri  (blas/ssyrk.c) ---- */
CUresult cuSsyrk(CUBLAShandle, CBlasUplo, CBlasTranspose, size_t n, size_t k,
               float alpha, CUdeviceptr A, size_t lda, float beta, CUdeviceptr C, size_t ldc, CUstream);
/* D = alpha*op(A)*op(B) + beta*C, out of place (blas/sgemm.c cuSgemm2); cuSgemm is the in-place form */
CUresult cuSgemm2(CUBLAShandle, CBlasTranspose, CBlasTranspose, size_t m, size_t n, size_t k,
                float alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb,
                float beta, CUdeviceptr C, size_t ldc, CUdeviceptr D, size_t ldd, CUstream);
#define cuSgemm(h, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, stream) \
  cuSgemm2(h, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, C, ldc, stream)
/* X = alpha*op(A)*B (Left) or alpha*B*op(A) (Right), A triangular, out of place (blas/strmm.c cuStrmm2) */
CUresult cuStrmm2(CUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
                float alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUdeviceptr X, size_t ldx, CUstream);
/* in-place form (defined but not declared in the reference: blas/strmm.c cuStrmm) */
CUresult cuStrmm(CUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
               float alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUstream);
/* B = alpha*op(A)^-1*B (Left) or alpha*B*op(A)^-1 (Right) (blas/strsm.c cuStrsm) */
CUresult cuStrsm(CUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
               float alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUstream);

/* ---- d: C_tri = alpha*op(A)*op(A)^T + beta*C_tri  (blas/dsyrk.c) ---- */
CUresult cuDsyrk(CUBLAShandle, CBlasUplo, CBlasTranspose, size_t n, size_t k,
               double alpha, CUdeviceptr A, size_t lda, double beta, CUdeviceptr C, size_t ldc, CUstream);
/* D = alpha*op(A)*op(B) + beta*C, out of place (blas/dgemm.c cuDgemm2); cuDgemm is the in-place form */
CUresult cuDgemm2(CUBLAShandle, CBlasTranspose, CBlasTranspose, size_t m, size_t n, size_t k,
                double alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb,
                double beta, CUdeviceptr C, size_t ldc, CUdeviceptr D, size_t ldd, CUstream);
#define cuDgemm(h, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, stream) \
  cuDgemm2(h, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, C, ldc, stream)
/* X = alpha*op(A)*B (Left) or alpha*B*op(A) (Right), A triangular, out of place (blas/dtrmm.c cuDtrmm2) */
CUresult cuDtrmm2(CUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
                double alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUdeviceptr X, size_t ldx, CUstream);
/* in-place form (defined but not declared in the reference: blas/dtrmm.c cuDtrmm) */
CUresult cuDtrmm(CUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
               double alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUstream);
/* B = alpha*op(A)^-1*B (Left) or alpha*B*op(A)^-1 (Right) (blas/dtrsm.c cuDtrsm) */
CUresult cuDtrsm(CUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
               double alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUstream);

/* ---- c: C_tri = alpha*op(A)*op(A)^H + beta*C_tri  (blas/cherk.c) ---- */
CUresult cuCherk(CUBLAShandle, CBlasUplo, CBlasTranspose, size_t n, size_t k,
               float alpha, CUdeviceptr A, size_t lda, float beta, CUdeviceptr C, size_t ldc, CUstream);
/* D = alpha*op(A)*op(B) + beta*C, out of place (blas/cgemm.c cuCgemm2); cuCgemm is the in-place form */
CUresult cuCgemm2(CUBLAShandle, CBlasTranspose, CBlasTranspose, size_t m, size_t n, size_t k,
                chol_cfloat alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb,
                chol_cfloat beta, CUdeviceptr C, size_t ldc, CUdeviceptr D, size_t ldd, CUstream);
#define cuCgemm(h, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, stream) \
  cuCgemm2(h, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, C, ldc, stream)
/* X = alpha*op(A)*B (Left) or alpha*B*op(A) (Right), A triangular, out of place (blas/ctrmm.c cuCtrmm2) */
CUresult cuCtrmm2(CUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
                chol_cfloat alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUdeviceptr X, size_t ldx, CUstream);
/* in-place form (defined but not declared in the reference: blas/ctrmm.c cuCtrmm) */
CUresult cuCtrmm(CUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
               chol_cfloat alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUstream);
/* B = alpha*op(A)^-1*B (Left) or alpha*B*op(A)^-1 (Right) (blas/ctrsm.c cuCtrsm) */
CUresult cuCtrsm(CUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
               chol_cfloat alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUstream);

/* ---- z: C_tri = alpha*op(A)*op(A)^H + beta*C_tri  (blas/zherk.c) ---- */
CUresult cuZherk(CUBLAShandle, CBlasUplo, CBlasTranspose, size_t n, size_t k,
               double alpha, CUdeviceptr A, size_t lda, double beta, CUdeviceptr C, size_t ldc, CUstream);
/* D = alpha*op(A)*op(B) + beta*C, out of place (blas/zgemm.c cuZgemm2); cuZgemm is the in-place form */
CUresult cuZgemm2(CUBLAShandle, CBlasTranspose, CBlasTranspose, size_t m, size_t n, size_t k,
                chol_cdouble alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb,
                chol_cdouble beta, CUdeviceptr C, size_t ldc, CUdeviceptr D, size_t ldd, CUstream);
#define cuZgemm(h, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, stream) \
  cuZgemm2(h, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, C, ldc, stream)
/* X = alpha*op(A)*B (Left) or alpha*B*op(A) (Right), A triangular, out of place (blas/ztrmm.c cuZtrmm2) */
CUresult cuZtrmm2(CUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
                chol_cdouble alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUdeviceptr X, size_t ldx, CUstream);
/* in-place form (defined but not declared in the reference: blas/ztrmm.c cuZtrmm) */
CUresult cuZtrmm(CUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
               chol_cdouble alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUstream);
/* B = alpha*op(A)^-1*B (Left) or alpha*B*op(A)^-1 (Right) (blas/ztrsm.c cuZtrsm) */
CUresult cuZtrsm(CUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
               chol_cdouble alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUstream);

/* ---- multi-GPU handle: operands in HOST memory, work on every GPU of the CUmultiGPU ---- */
typedef struct __cumultigpublashandle_st *CUmultiGPUBLAShandle;
CUresult cuMultiGPUBLASCreate(CUmultiGPUBLAShandle *, CUmultiGPU);
CUresult cuMultiGPUBLASDestroy(CUmultiGPUBLAShandle);
CUresult cuMultiGPUBLASSynchronize(CUmultiGPUBLAShandle);

CUresult cuMultiGPUSsyrk(CUmultiGPUBLAShandle, CBlasUplo, CBlasTranspose, size_t n, size_t k,
                       float alpha, const float *A, size_t lda, float beta, float *C, size_t ldc);
CUresult cuMultiGPUSgemm(CUmultiGPUBLAShandle, CBlasTranspose, CBlasTranspose, size_t m, size_t n, size_t k,
                       float alpha, const float *A, size_t lda, const float *B, size_t ldb,
                       float beta, float *C, size_t ldc);
CUresult cuMultiGPUStrmm(CUmultiGPUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
                       float alpha, const float *A, size_t lda, float *B, size_t ldb);
CUresult cuMultiGPUStrsm(CUmultiGPUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
                       float alpha, const float *A, size_t lda, float *B, size_t ldb);

CUresult cuMultiGPUDsyrk(CUmultiGPUBLAShandle, CBlasUplo, CBlasTranspose, size_t n, size_t k,
                       double alpha, const double *A, size_t lda, double beta, double *C, size_t ldc);
CUresult cuMultiGPUDgemm(CUmultiGPUBLAShandle, CBlasTranspose, CBlasTranspose, size_t m, size_t n, size_t k,
                       double alpha, const double *A, size_t lda, const double *B, size_t ldb,
                       double beta, double *C, size_t ldc);
CUresult cuMultiGPUDtrmm(CUmultiGPUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
                       double alpha, const double *A, size_t lda, double *B, size_t ldb);
CUresult cuMultiGPUDtrsm(CUmultiGPUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
                       double alpha, const double *A, size_t lda, double *B, size_t ldb);

CUresult cuMultiGPUCherk(CUmultiGPUBLAShandle, CBlasUplo, CBlasTranspose, size_t n, size_t k,
                       float alpha, const chol_cfloat *A, size_t lda, float beta, chol_cfloat *C, size_t ldc);
CUresult cuMultiGPUCgemm(CUmultiGPUBLAShandle, CBlasTranspose, CBlasTranspose, size_t m, size_t n, size_t k,
                       chol_cfloat alpha, const chol_cfloat *A, size_t lda, const chol_cfloat *B, size_t ldb,
                       chol_cfloat beta, chol_cfloat *C, size_t ldc);
CUresult cuMultiGPUCtrmm(CUmultiGPUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
                       chol_cfloat alpha, const chol_cfloat *A, size_t lda, chol_cfloat *B, size_t ldb);
CUresult cuMultiGPUCtrsm(CUmultiGPUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
                       chol_cfloat alpha, const chol_cfloat *A, size_t lda, chol_cfloat *B, size_t ldb);

CUresult cuMultiGPUZherk(CUmultiGPUBLAShandle, CBlasUplo, CBlasTranspose, size_t n, size_t k,
                       double alpha, const chol_cdouble *A, size_t lda, double beta, chol_cdouble *C, size_t ldc);
CUresult cuMultiGPUZgemm(CUmultiGPUBLAShandle, CBlasTranspose, CBlasTranspose, size_t m, size_t n, size_t k,
                       chol_cdouble alpha, const chol_cdouble *A, size_t lda, const chol_cdouble *B, size_t ldb,
                       chol_cdouble beta, chol_cdouble *C, size_t ldc);
CUresult cuMultiGPUZtrmm(CUmultiGPUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
                       chol_cdouble alpha, const chol_cdouble *A, size_t lda, chol_cdouble *B, size_t ldb);
CUresult cuMultiGPUZtrsm(CUmultiGPUBLAShandle, CBlasSide, CBlasUplo, CBlasTranspose, CBlasDiag, size_t m, size_t n,
                       chol_cdouble alpha, const chol_cdouble *A, size_t lda, chol_cdouble *B, size_t ldb);

#ifdef __cplusplus
}
#endif
#endif
