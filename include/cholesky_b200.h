/*
 * cholesky_b200.h — entry points of libcholesky_b200.so that the reference does NOT have.
 *
 * blas.h / lapack.h / cumultigpu.h / error.h are the drop-in surface (same names and argument
 * meaning as the reference's headers of the same name).  Everything below is additional and
 * optional: a caller of the reference API never needs it.
 *
 *   cholDist*        one-process-per-GPU form of the distributed Cholesky.  The reference only has
 *                    the one-process / one-thread-per-GPU model (multigpu/multigpu.c:405-538) with
 *                    the matrix in host memory; under a launcher such as torchrun every rank owns
 *                    one GPU and keeps its share of the matrix RESIDENT in device memory.
 *   cuLAPACK*        tuning and measurement hooks on a CULAPACKhandle.
 *   chol_probe_fp64  the FP64 tensor-core peak of the device, measured (roofline denominator).
 *   chol_launch_count  number of kernels this library has launched (bench.py's gpu_launches).
 */
#ifndef CHOL_B200_EXT_H
#define CHOL_B200_EXT_H

#include "lapack.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------
 * Distributed POTRF, one process per GPU.
 *
 * Layout: 1-D block-cyclic by block column of width nb = cholDistBlockSize(...).  Global block
 * column c lives on rank c % nranks as that rank's local block column c / nranks; a rank's block
 * columns are packed side by side in ONE column-major device array `Aloc` with n rows, leading
 * dimension ldl >= n and cholDistLocalCols(n, nb, rank, nranks) columns.  Only the lower triangle
 * (global row >= global column) is read or written; an upper factorisation is the lower one of
 * A^H.  Per step the owner factors the panel and ncclBroadcast sends it to the other ranks.
 * ------------------------------------------------------------------------------------------- */
typedef struct chol_rank *cholDist;

/* 128-byte NCCL unique id: create on rank 0, ship to the other ranks with the launcher's own
 * means (bench.py uses a torch.distributed broadcast). */
CUresult cholDistGetUniqueId(void *id128);
/* Binds the calling thread's current CUDA context (or device 0's primary context if none).
 * nranks == 1 needs neither an id nor NCCL.  Collective over the ranks when nranks > 1. */
CUresult cholDistCreate(cholDist *out, int rank, int nranks, const void *id128);
CUresult cholDistDestroy(cholDist);
/* nb = 0 restores the default; nb must be a multiple of 64.  Must be the same on every rank. */
CUresult cholDistSetBlockSize(cholDist, size_t nb);
/* block size a factorisation of order n in precision prec ('s','d','c','z') will use */
size_t cholDistBlockSize(cholDist, int prec, size_t n);
/* number of columns of an n x n matrix that `rank` holds */
size_t cholDistLocalCols(size_t n, size_t nb, int rank, int nranks);
/* the rank's CULAPACKhandle (streams, workspace, profiling hooks); owned by the cholDist */
CULAPACKhandle cholDistLapackHandle(cholDist);

/* Factor the distributed matrix in place (lower).  Collective.  *info = 0, or the 1-based global
 * column of the first non-positive pivot THIS rank met (combine with a min over the ranks that
 * report > 0), or -4 if ldl < n.  Returns after this rank's device work has completed. */
CUresult cholDistSpotrf(cholDist, size_t n, CUdeviceptr Aloc, size_t ldl, long *info);
CUresult cholDistDpotrf(cholDist, size_t n, CUdeviceptr Aloc, size_t ldl, long *info);
CUresult cholDistCpotrf(cholDist, size_t n, CUdeviceptr Aloc, size_t ldl, long *info);
CUresult cholDistZpotrf(cholDist, size_t n, CUdeviceptr Aloc, size_t ldl, long *info);

/* ---------------------------------------------------------------------------------------------
 * POTRS: solve A X = B for n x nrhs right-hand sides from the Cholesky factor cu*potrf left in A
 * (the natural consumer of the factor; the reference stops at POTRI).  B is overwritten by X.
 * Device pointers, column-major.  *info = -5 (lda < n) or -7 (ldb < n), else 0.  Returns after the
 * device work has completed.
 * ------------------------------------------------------------------------------------------- */
CUresult cuSpotrs(CULAPACKhandle, CBlasUplo, size_t n, size_t nrhs, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, long *info);
CUresult cuDpotrs(CULAPACKhandle, CBlasUplo, size_t n, size_t nrhs, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, long *info);
CUresult cuCpotrs(CULAPACKhandle, CBlasUplo, size_t n, size_t nrhs, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, long *info);
CUresult cuZpotrs(CULAPACKhandle, CBlasUplo, size_t n, size_t nrhs, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, long *info);

/* ---------------------------------------------------------------------------------------------
 * Fused log-determinant: the diagonal-block kernel of POTRF adds 2 log L(i,i) while it factors, so log det A of the
 * matrix the last cu{S,D,C,Z}potrf on this handle factored is available without another pass (the reference needs
 * cu*logdet on the factor's diagonal, lapack/dlogdet.c:37-74).  CUDA_ERROR_INVALID_VALUE if that call failed (info > 0)
 * or none has run.  With cholDist*potrf every rank holds the share of the diagonal blocks it factored: sum over ranks.
 * ------------------------------------------------------------------------------------------- */
CUresult cuLAPACKPotrfLogdet(CULAPACKhandle, double *logdet);

/* ---------------------------------------------------------------------------------------------
 * Tuning / measurement hooks
 * ------------------------------------------------------------------------------------------- */
/* outer block size of the LAPACK drivers (k of the trailing update); 0 = built-in default
 * (per precision and n, measured on B200), otherwise a multiple of 64 */
CUresult cuLAPACKSetBlockSize(CULAPACKhandle, size_t nb);
/* enable != 0: time every trailing-update launch of the following POTRF calls with CUDA events on
 * the stream the kernel runs on, and reset the totals */
CUresult cuLAPACKProfile(CULAPACKhandle, int enable);
/* totals since cuLAPACKProfile(h, 1): device milliseconds, algorithmic flops, launches */
CUresult cuLAPACKProfileRead(CULAPACKhandle, double *ms, double *flops, int *launches);
/* per-launch (ms, flops) of the first min(cap, n) timed launches; returns n */
int cuLAPACKProfileHistory(CULAPACKhandle, float *ms, double *flops, int cap);

/* FP64 throughput of the current device in TFLOP/s from a register-only kernel that runs `iters`
 * rounds of the GEMM engine's own instruction mix.  kind 0: DMMA (mma.sync.m8n8k4.f64),
 * 1: DFMA, 2: both interleaved, 3: DMMA with one shared operand pair, 10/20: DMMA with 1/2 CTAs
 * per SM.  Synchronous.  Returns a cudaError_t value (0 = success). */
int chol_probe_fp64(int kind, int iters, double *tflops, float *ms);
/* kernels launched by this library so far; reset != 0 zeroes the counter afterwards */
unsigned long long chol_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif
