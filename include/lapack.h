/*
 * lapack.h — POTRF / TRTRI / LAUUM / POTRI / logdet of the B200 Cholesky library.
 *
 * Drop-in for the GPU half of reference include/lapack.h: cu{S,D,C,Z}{potrf,trtri,trti22,lauum,
 * potri,logdet} (lapack.h:64-122, :165-169) on device pointers and cuMultiGPU{S,D,C,Z}{potrf,
 * trtri,lauum,potri} (lapack.h:124-159) on host pointers, with the reference's handle life-cycle
 * (lapack/handle.c:77-108).  The un-prefixed CPU routines (lapack.h:16-62, :161-164) are NOT
 * exported - this library has no CPU path; they exist only in the test oracle (oracle/).
 *
 * Semantics kept from the reference:
 *   - column-major, lda in elements, only the `uplo` triangle is read or written;
 *   - *info < 0: argument -info is invalid (XERBLA is called, CUDA_ERROR_INVALID_VALUE returned);
 *     *info > 0: potrf - the leading minor of that order is not positive definite
 *     (lapack/dpotrf.c:70-73); trtri/potri - that diagonal entry is exactly zero (dtrtri.c:49-52);
 *   - POTRI takes the CHOLESKY FACTOR as input, not A (lapack/dpotri.c:19-22).
 * Semantics tightened (documented in DESIGN.md): the routines return only after all device work
 * has completed (the reference leaves the last kernels in flight, lapack/dpotrf.c:283-295), and a
 * numerical failure returns CUDA_SUCCESS with *info > 0 for all four precisions (the reference is
 * inconsistent: cuDpotrf SUCCESS, cuZpotrf INVALID_VALUE).
 */
#ifndef CHOL_B200_LAPACK_H
#define CHOL_B200_LAPACK_H

#include "blas.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct __culapackhandle_st *CULAPACKhandle;
CUresult cuLAPACKCreate(CULAPACKhandle *);
CUresult cuLAPACKDestroy(CULAPACKhandle);

/* ---- s ---- */
/* A = L*L^H (Lower) or U^H*U (Upper), in place (lapack/spotrf.c cuSpotrf) */
CUresult cuSpotrf(CULAPACKhandle, CBlasUplo, size_t n, CUdeviceptr A, size_t lda, long *info);
/* A := A^-1 for triangular A, in place (lapack/strtri.c cuStrtri) */
CUresult cuStrtri(CULAPACKhandle, CBlasUplo, CBlasDiag, size_t n, CUdeviceptr A, size_t lda, long *info);
/* unblocked out-of-place inverse of one small block (n <= 128; 64 for z): B_tri := A_tri^-1; the
 * info argument is a DEVICE pointer to a 32-bit int which the kernel sets to 0 on entry and to the
 * 1-based column of a zero pivot, as the reference's kernel does (lapack.h:91-98, lapack/dtrtri.cu:23-29);
 * it need not be initialised */
CUresult cuStrti22(CULAPACKhandle, CBlasUplo, CBlasDiag, size_t n, CUdeviceptr A, size_t lda,
                 CUdeviceptr B, size_t ldb, CUdeviceptr info, CUstream);
#define cuStrti2(handle, uplo, diag, n, A, lda, info) cuStrti22(handle, uplo, diag, n, A, lda, A, lda, info, 0)
/* A := U*U^H (Upper) or L^H*L (Lower), in place (lapack/slauum.c cuSlauum) */
CUresult cuSlauum(CULAPACKhandle, CBlasUplo, size_t n, CUdeviceptr A, size_t lda, long *info);
/* inverse of an SPD/HPD matrix from its Cholesky factor: trtri(NonUnit) then lauum (lapack/spotri.c) */
CUresult cuSpotri(CULAPACKhandle, CBlasUplo, size_t n, CUdeviceptr A, size_t lda, long *info);
/* *result = 2*sum_i log(Re x[i*incx]); x = diag of the factor with incx = lda+1 (lapack/slogdet.c).
 * result is a HOST pointer; it is valid once `stream` has been synchronised. */
CUresult cuSlogdet(CULAPACKhandle, CUdeviceptr x, size_t incx, size_t n, float *result, CUstream);

/* ---- d ---- */
/* A = L*L^H (Lower) or U^H*U (Upper), in place (lapack/dpotrf.c cuDpotrf) */
CUresult cuDpotrf(CULAPACKhandle, CBlasUplo, size_t n, CUdeviceptr A, size_t lda, long *info);
/* A := A^-1 for triangular A, in place (lapack/dtrtri.c cuDtrtri) */
CUresult cuDtrtri(CULAPACKhandle, CBlasUplo, CBlasDiag, size_t n, CUdeviceptr A, size_t lda, long *info);
/* unblocked out-of-place inverse of one small block (n <= 128; 64 for z): B_tri := A_tri^-1; the
 * info argument is a DEVICE pointer to a 32-bit int which the kernel sets to 0 on entry and to the
 * 1-based column of a zero pivot, as the reference's kernel does (lapack.h:91-98, lapack/dtrtri.cu:23-29);
 * it need not be initialised */
CUresult cuDtrti22(CULAPACKhandle, CBlasUplo, CBlasDiag, size_t n, CUdeviceptr A, size_t lda,
                 CUdeviceptr B, size_t ldb, CUdeviceptr info, CUstream);
#define cuDtrti2(handle, uplo, diag, n, A, lda, info) cuDtrti22(handle, uplo, diag, n, A, lda, A, lda, info, 0)
/* A := U*U^H (Upper) or L^H*L (Lower), in place (lapack/dlauum.c cuDlauum) */
CUresult cuDlauum(CULAPACKhandle, CBlasUplo, size_t n, CUdeviceptr A, size_t lda, long *info);
/* inverse of an SPD/HPD matrix from its Cholesky factor: trtri(NonUnit) then lauum (lapack/dpotri.c) */
CUresult cuDpotri(CULAPACKhandle, CBlasUplo, size_t n, CUdeviceptr A, size_t lda, long *info);
/* *result = 2*sum_i log(Re x[i*incx]); x = diag of the factor with incx = lda+1 (lapack/dlogdet.c).
 * result is a HOST pointer; it is valid once `stream` has been synchronised. */
CUresult cuDlogdet(CULAPACKhandle, CUdeviceptr x, size_t incx, size_t n, double *result, CUstream);

/* ---- c ---- */
/* A = L*L^H (Lower) or U^H*U (Upper), in place (lapack/cpotrf.c cuCpotrf) */
CUresult cuCpotrf(CULAPACKhandle, CBlasUplo, size_t n, CUdeviceptr A, size_t lda, long *info);
/* A := A^-1 for triangular A, in place (lapack/ctrtri.c cuCtrtri) */
CUresult cuCtrtri(CULAPACKhandle, CBlasUplo, CBlasDiag, size_t n, CUdeviceptr A, size_t lda, long *info);
/* unblocked out-of-place inverse of one small block (n <= 128; 64 for z): B_tri := A_tri^-1; the
 * info argument is a DEVICE pointer to a 32-bit int which the kernel sets to 0 on entry and to the
 * 1-based column of a zero pivot, as the reference's kernel does (lapack.h:91-98, lapack/dtrtri.cu:23-29);
 * it need not be initialised */
CUresult cuCtrti22(CULAPACKhandle, CBlasUplo, CBlasDiag, size_t n, CUdeviceptr A, size_t lda,
                 CUdeviceptr B, size_t ldb, CUdeviceptr info, CUstream);
#define cuCtrti2(handle, uplo, diag, n, A, lda, info) cuCtrti22(handle, uplo, diag, n, A, lda, A, lda, info, 0)
/* A := U*U^H (Upper) or L^H*L (Lower), in place (lapack/clauum.c cuClauum) */
CUresult cuClauum(CULAPACKhandle, CBlasUplo, size_t n, CUdeviceptr A, size_t lda, long *info);
/* inverse of an SPD/HPD matrix from its Cholesky factor: trtri(NonUnit) then lauum (lapack/cpotri.c) */
CUresult cuCpotri(CULAPACKhandle, CBlasUplo, size_t n, CUdeviceptr A, size_t lda, long *info);
/* *result = 2*sum_i log(Re x[i*incx]); x = diag of the factor with incx = lda+1 (lapack/clogdet.c).
 * result is a HOST pointer; it is valid once `stream` has been synchronised. */
CUresult cuClogdet(CULAPACKhandle, CUdeviceptr x, size_t incx, size_t n, float *result, CUstream);

/* ---- z ---- */
/* A = L*L^H (Lower) or U^H*U (Upper), in place (lapack/zpotrf.c cuZpotrf) */
CUresult cuZpotrf(CULAPACKhandle, CBlasUplo, size_t n, CUdeviceptr A, size_t lda, long *info);
/* A := A^-1 for triangular A, in place (lapack/ztrtri.c cuZtrtri) */
CUresult cuZtrtri(CULAPACKhandle, CBlasUplo, CBlasDiag, size_t n, CUdeviceptr A, size_t lda, long *info);
/* unblocked out-of-place inverse of one small block (n <= 128; 64 for z): B_tri := A_tri^-1; the
 * info argument is a DEVICE pointer to a 32-bit int which the kernel sets to 0 on entry and to the
 * 1-based column of a zero pivot, as the reference's kernel does (lapack.h:91-98, lapack/dtrtri.cu:23-29);
 * it need not be initialised */
CUresult cuZtrti22(CULAPACKhandle, CBlasUplo, CBlasDiag, size_t n, CUdeviceptr A, size_t lda,
                 CUdeviceptr B, size_t ldb, CUdeviceptr info, CUstream);
#define cuZtrti2(handle, uplo, diag, n, A, lda, info) cuZtrti22(handle, uplo, diag, n, A, lda, A, lda, info, 0)
/* A := U*U^H (Upper) or L^H*L (Lower), in place (lapack/zlauum.c cuZlauum) */
CUresult cuZlauum(CULAPACKhandle, CBlasUplo, size_t n, CUdeviceptr A, size_t lda, long *info);
/* inverse of an SPD/HPD matrix from its Cholesky factor: trtri(NonUnit) then lauum (lapack/zpotri.c) */
CUresult cuZpotri(CULAPACKhandle, CBlasUplo, size_t n, CUdeviceptr A, size_t lda, long *info);
/* *result = 2*sum_i log(Re x[i*incx]); x = diag of the factor with incx = lda+1 (lapack/zlogdet.c).
 * result is a HOST pointer; it is valid once `stream` has been synchronised. */
CUresult cuZlogdet(CULAPACKhandle, CUdeviceptr x, size_t incx, size_t n, double *result, CUstream);

/* ---- multi-GPU: matrices in HOST memory, factorisation on every GPU of the CUmultiGPU ---- */
typedef struct __cumultigpulapackhandle_st *CUmultiGPULAPACKhandle;
CUresult cuMultiGPULAPACKCreate(CUmultiGPULAPACKhandle *, CUmultiGPU);
CUresult cuMultiGPULAPACKDestroy(CUmultiGPULAPACKhandle);
CUresult cuMultiGPULAPACKSynchronize(CUmultiGPULAPACKhandle);

CUresult cuMultiGPUSpotrf(CUmultiGPULAPACKhandle, CBlasUplo, size_t n, float *A, size_t lda, long *info);
CUresult cuMultiGPUStrtri(CUmultiGPULAPACKhandle, CBlasUplo, CBlasDiag, size_t n, float *A, size_t lda, long *info);
CUresult cuMultiGPUSlauum(CUmultiGPULAPACKhandle, CBlasUplo, size_t n, float *A, size_t lda, long *info);
CUresult cuMultiGPUSpotri(CUmultiGPULAPACKhandle, CBlasUplo, size_t n, float *A, size_t lda, long *info);

CUresult cuMultiGPUDpotrf(CUmultiGPULAPACKhandle, CBlasUplo, size_t n, double *A, size_t lda, long *info);
CUresult cuMultiGPUDtrtri(CUmultiGPULAPACKhandle, CBlasUplo, CBlasDiag, size_t n, double *A, size_t lda, long *info);
CUresult cuMultiGPUDlauum(CUmultiGPULAPACKhandle, CBlasUplo, size_t n, double *A, size_t lda, long *info);
CUresult cuMultiGPUDpotri(CUmultiGPULAPACKhandle, CBlasUplo, size_t n, double *A, size_t lda, long *info);

CUresult cuMultiGPUCpotrf(CUmultiGPULAPACKhandle, CBlasUplo, size_t n, chol_cfloat *A, size_t lda, long *info);
CUresult cuMultiGPUCtrtri(CUmultiGPULAPACKhandle, CBlasUplo, CBlasDiag, size_t n, chol_cfloat *A, size_t lda, long *info);
CUresult cuMultiGPUClauum(CUmultiGPULAPACKhandle, CBlasUplo, size_t n, chol_cfloat *A, size_t lda, long *info);
CUresult cuMultiGPUCpotri(CUmultiGPULAPACKhandle, CBlasUplo, size_t n, chol_cfloat *A, size_t lda, long *info);

CUresult cuMultiGPUZpotrf(CUmultiGPULAPACKhandle, CBlasUplo, size_t n, chol_cdouble *A, size_t lda, long *info);
CUresult cuMultiGPUZtrtri(CUmultiGPULAPACKhandle, CBlasUplo, CBlasDiag, size_t n, chol_cdouble *A, size_t lda, long *info);
CUresult cuMultiGPUZlauum(CUmultiGPULAPACKhandle, CBlasUplo, size_t n, chol_cdouble *A, size_t lda, long *info);
CUresult cuMultiGPUZpotri(CUmultiGPULAPACKhandle, CBlasUplo, size_t n, chol_cdouble *A, size_t lda, long *info);

#ifdef __cplusplus
}
#endif
#endif
