/*
 * blas3.c — the BLAS-3 surface on device pointers: cu{S,D,C,Z}{gemm2, syrk|herk, trmm2, trmm, trsm}.
 *
 * Same argument checks, XERBLA positions and early-outs as the reference launchers
 * (blas/dgemm.c:148-200, dsyrk.c:163-209, zherk.c:190-238, dtrmm.c:374-465, dtrsm.c:218-265), but
 * every routine is expressed on ONE tile engine (csrc/gemm_dmma.cu for d/z, csrc/gemm_ffma.cu
 * for s/c) through a descriptor:
 *   SYRK/HERK = GEMM whose grid holds only the tiles of one triangle,
 *   TRMM2     = GEMM whose loader masks the triangular operand and skips its zero k-range,
 *   TRSM      = blocked substitution: invert 128x128 diagonal blocks in shared memory
 *               (csrc/panel_kernels.cu trti2) and apply them / update the rest with GEMMs
 *               (the reference's idea of turning the panel solve into a multiply,
 *               lapack/spotrf.c:253,:382, applied to every TRSM variant).
 * All calls are asynchronous on the caller's stream.
 */
#include <complex.h>

#include "chol_internal.h"

static const double ONE[2] = {1.0, 0.0};
static const double ZERO[2] = {0.0, 0.0};
static const double MONE[2] = {-1.0, 0.0};

CUresult chol_gemm_ex(int prec, int opA, int opB, int triA, int unitA, int triB, int unitB, int cuplo, int herm,
                      size_t m, size_t n, size_t k, const double alpha[2], CUdeviceptr A, size_t lda,
                      CUdeviceptr B, size_t ldb, const double beta[2], CUdeviceptr C, size_t ldc, CUdeviceptr D,
                      size_t ldd, CUstream stream) {
  chol_gemm_desc d;
  d.prec = prec;
  d.opA = opA; d.opB = opB;
  d.triA = triA; d.triB = triB;
  d.unitA = unitA; d.unitB = unitB;
  d.cuplo = cuplo;
  d.herm = herm;
  d.m = m; d.n = n; d.k = k;
  d.alpha[0] = alpha[0]; d.alpha[1] = alpha[1];
  d.beta[0] = beta[0]; d.beta[1] = beta[1];
  d.A = (const void *)A; d.lda = lda;
  d.B = (const void *)B; d.ldb = ldb;
  d.C = (const void *)C; d.ldc = ldc;
  d.D = (void *)D; d.ldd = ldd;
  d.stream = (void *)stream;
  d.force_generic = 0;
  d.cyc_nb = d.cyc_p = d.cyc_off = 0;
  RT_CHECK(chol_gemm_launch(&d));
  return CUDA_SUCCESS;
}

CUresult chol_gemm(int prec, int opA, int opB, size_t m, size_t n, size_t k, const double alpha[2], CUdeviceptr A,
                   size_t lda, CUdeviceptr B, size_t ldb, const double beta[2], CUdeviceptr C, size_t ldc,
                   CUdeviceptr D, size_t ldd, CUstream stream) {
  return chol_gemm_ex(prec, opA, opB, 0, 0, 0, 0, 0, 0, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, D, ldd, stream);
}

/* C_tri = alpha*A*A^H + beta*C_tri (trans N) or alpha*A^H*A + beta*C_tri (trans T/C) */
CUresult chol_rankk(int prec, int uplo, int trans, size_t n, size_t k, double alpha, CUdeviceptr A, size_t lda,
                    double beta, CUdeviceptr C, size_t ldc, CUstream stream) {
  const double a[2] = {alpha, 0.0}, b[2] = {beta, 0.0};
  const int h = chol_is_complex(prec) ? 'C' : 'T';
  if (trans == 'N')
    return chol_gemm_ex(prec, 'N', h, 0, 0, 0, 0, uplo, chol_is_complex(prec), n, n, k, a, A, lda, A, lda, b, C, ldc,
                        C, ldc, stream);
  return chol_gemm_ex(prec, h, 'N', 0, 0, 0, 0, uplo, chol_is_complex(prec), n, n, k, a, A, lda, A, lda, b, C, ldc, C,
                      ldc, stream);
}

CUresult chol_trmm2(int prec, int side, int uplo, int trans, int diag, size_t m, size_t n, const double alpha[2],
                    CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUdeviceptr X, size_t ldx, CUstream stream) {
  if (side == 'L')
    return chol_gemm_ex(prec, trans, 'N', uplo, diag == 'U', 0, 0, 0, 0, m, n, m, alpha, A, lda, B, ldb, ZERO, 0, 0, X,
                        ldx, stream);
  return chol_gemm_ex(prec, 'N', trans, 0, 0, uplo, diag == 'U', 0, 0, m, n, n, alpha, B, ldb, A, lda, ZERO, 0, 0, X,
                      ldx, stream);
}

static CUresult trsm_body(CUBLAShandle h, int prec, int side, int uplo, int trans, int diag, size_t m, size_t n,
                          const double alpha[2], CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUstream stream) {
  const size_t es = chol_esize(prec);
  const size_t tb = (size_t)chol_diag_max(prec);
  const size_t na = (side == 'L') ? m : n;
  const bool eff_lower = (uplo == 'L') == (trans == 'N'); /* shape of T = op(A) */
  const bool forward = (side == 'L') == eff_lower;
  const size_t lds = (side == 'L') ? tb : chol_roundup(m, 2);
  const size_t s_elems = (side == 'L') ? tb * n : lds * tb;
  CUdeviceptr scr;
  CU_ERROR_CHECK(chol_blas_scratch(h, (tb * tb + s_elems) * es + 256, &scr));
  const CUdeviceptr Ainv = scr, S = scr + (CUdeviceptr)(chol_roundup(tb * tb * es, 256));
  const size_t nblk = (na + tb - 1) / tb;
  for (size_t b = 0; b < nblk; b++) {
    const size_t blk = forward ? b : nblk - 1 - b;
    const size_t i = blk * tb, ib = chol_min(tb, na - i);
    const bool first = (b == 0);
    const double *a_diag = first ? alpha : ONE;
    const double *b_upd = first ? alpha : ONE;
    /* rows/cols still to be solved after this block */
    const size_t r0 = forward ? i + ib : 0, rn = forward ? na - i - ib : i;
    /* no info word: a zero pivot gives Inf/NaN in B exactly as a substitution TRSM would (blas/dtrsm.c), and one
     * singular block cannot turn the later blocks' inverses into no-ops */
    RT_CHECK(chol_trti2_launch(prec, uplo, diag, ib, (const void *)chol_at(A, lda, i, i, es), lda, (void *)Ainv, tb,
                               1 | CHOL_TRTI2_NO_INFO, (long *)h->dinfo, 0, (void *)stream));
    if (side == 'L') {
      const CUdeviceptr Bi = chol_at(B, ldb, i, 0, es);
      CU_ERROR_CHECK(chol_gemm(prec, trans, 'N', ib, n, ib, a_diag, Ainv, tb, Bi, ldb, ZERO, 0, 0, S, lds, stream));
      RT_CHECK(chol_copy_launch(prec, 'N', 0, ib, n, (const void *)S, lds, (void *)Bi, ldb, (void *)stream));
      if (rn) {
        const CUdeviceptr Brem = chol_at(B, ldb, r0, 0, es);
        const CUdeviceptr Asub = (trans == 'N') ? chol_at(A, lda, r0, i, es) : chol_at(A, lda, i, r0, es);
        CU_ERROR_CHECK(chol_gemm(prec, trans, 'N', rn, n, ib, MONE, Asub, lda, Bi, ldb, b_upd, Brem, ldb, Brem, ldb, stream));
      }
    } else {
      const CUdeviceptr Bj = chol_at(B, ldb, 0, i, es);
      CU_ERROR_CHECK(chol_gemm(prec, 'N', trans, m, ib, ib, a_diag, Bj, ldb, Ainv, tb, ZERO, 0, 0, S, lds, stream));
      RT_CHECK(chol_copy_launch(prec, 'N', 0, m, ib, (const void *)S, lds, (void *)Bj, ldb, (void *)stream));
      if (rn) {
        const CUdeviceptr Brem = chol_at(B, ldb, 0, r0, es);
        const CUdeviceptr Asub = (trans == 'N') ? chol_at(A, lda, i, r0, es) : chol_at(A, lda, r0, i, es);
        CU_ERROR_CHECK(chol_gemm(prec, 'N', trans, m, rn, ib, MONE, Bj, ldb, Asub, lda, b_upd, Brem, ldb, Brem, ldb, stream));
      }
    }
  }
  return CUDA_SUCCESS;
}

CUresult chol_trsm(CUBLAShandle h, int prec, int side, int uplo, int trans, int diag, size_t m, size_t n,
                   const double alpha[2], CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUstream stream) {
  if (m == 0 || n == 0) return CUDA_SUCCESS;
  if (alpha[0] == 0.0 && alpha[1] == 0.0) {
    RT_CHECK(chol_fill_launch(prec, 0, m, n, 0.0, 0.0, (void *)B, ldb, (void *)stream));
    return CUDA_SUCCESS;
  }
  CU_ERROR_CHECK(chol_blas_scratch_acquire(h, stream));
  const CUresult e = trsm_body(h, prec, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, stream);
  CU_ERROR_CHECK(chol_blas_scratch_release(h, stream)); /* also after a failure: whatever was queued may use the scratch */
  return e;
}

static CUresult chol_trmm_inplace(CUBLAShandle h, int prec, int side, int uplo, int trans, int diag, size_t m, size_t n,
                                  const double alpha[2], CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb,
                                  CUstream stream) {
  /* reference blas/dtrmm.c:428-465: allocate X, trmm2 into it, copy back */
  if (m == 0 || n == 0) return CUDA_SUCCESS;
  const size_t es = chol_esize(prec), ldx = chol_roundup(m, 2);
  CUdeviceptr X;
  CU_ERROR_CHECK(chol_blas_scratch_acquire(h, stream));
  CU_ERROR_CHECK(chol_blas_scratch(h, ldx * n * es, &X));
  CU_ERROR_CHECK(chol_trmm2(prec, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, X, ldx, stream));
  RT_CHECK(chol_copy_launch(prec, 'N', 0, m, n, (const void *)X, ldx, (void *)B, ldb, (void *)stream));
  CU_ERROR_CHECK(chol_blas_scratch_release(h, stream));
  return CUDA_SUCCESS;
}

/* ---------------------------------------------------------------------------------------------
 * typed entry points.  SC(x) turns the by-value scalar into (re, im).
 * ------------------------------------------------------------------------------------------- */
#define CHECK_ARG(cond, pos)                \
  do {                                      \
    if (cond) {                             \
      XERBLA(pos);                          \
      return CUDA_ERROR_INVALID_VALUE;      \
    }                                       \
  } while (false)

/* run `call` inside the handle's context, as the reference launchers do (blas/dgemm.c:173-197);
 * gemm2 / syrk|herk / trmm2 keep no per-handle state and also accept a NULL handle (caller's context) */
#define IN_BLAS_CONTEXT(handle, call)                          \
  do {                                                         \
    if ((handle) == NULL) return (call);                       \
    CU_ERROR_CHECK(cuCtxPushCurrent((handle)->context));       \
    const CUresult body__ = (call);                            \
    CUcontext popped__;                                        \
    const CUresult pop__ = cuCtxPopCurrent(&popped__);         \
    if (body__ != CUDA_SUCCESS) return body__;                 \
    CU_ERROR_CHECK(pop__);                                     \
    return CUDA_SUCCESS;                                       \
  } while (false)

#define GEN_BLAS3(U, p, SCALAR, REAL, RE, IM, RANKK, ISCPLX)                                                          \
  CUresult cu##U##gemm2(CUBLAShandle handle, CBlasTranspose transA, CBlasTranspose transB, size_t m, size_t n,         \
                        size_t k, SCALAR alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, SCALAR beta,     \
                        CUdeviceptr C, size_t ldc, CUdeviceptr D, size_t ldd, CUstream stream) {                       \
    const size_t nRowA = (transA == CBlasNoTrans) ? m : k, nRowB = (transB == CBlasNoTrans) ? k : n;                   \
    CHECK_ARG(lda < nRowA, 8);                                                                                         \
    CHECK_ARG(ldb < nRowB, 10);                                                                                        \
    CHECK_ARG(ldc < m, 13);                                                                                            \
    CHECK_ARG(ldd < m, 15);                                                                                            \
    const double a[2] = {RE(alpha), IM(alpha)}, b[2] = {RE(beta), IM(beta)};                                           \
    if (m == 0 || n == 0 || (C == D && ((a[0] == 0.0 && a[1] == 0.0) || k == 0) && b[0] == 1.0 && b[1] == 0.0))        \
      return CUDA_SUCCESS;                                                                                             \
    IN_BLAS_CONTEXT(handle, chol_gemm(p, transA, transB, m, n, k, a, A, lda, B, ldb, b, C, ldc, D, ldd, stream));      \
  }                                                                                                                    \
  CUresult cu##U##RANKK(CUBLAShandle handle, CBlasUplo uplo, CBlasTranspose trans, size_t n, size_t k, REAL alpha,     \
                        CUdeviceptr A, size_t lda, REAL beta, CUdeviceptr C, size_t ldc, CUstream stream) {            \
    const size_t nRowA = (trans == CBlasNoTrans) ? n : k;                                                              \
    CHECK_ARG(ISCPLX && trans == CBlasTrans, 2);                                                                       \
    CHECK_ARG(lda < nRowA, 7);                                                                                         \
    CHECK_ARG(ldc < n, 10);                                                                                            \
    if (n == 0 || ((alpha == 0 || k == 0) && beta == 1)) return CUDA_SUCCESS;                                          \
    IN_BLAS_CONTEXT(handle, chol_rankk(p, uplo, trans, n, k, (double)alpha, A, lda, (double)beta, C, ldc, stream));    \
  }                                                                                                                    \
  CUresult cu##U##trmm2(CUBLAShandle handle, CBlasSide side, CBlasUplo uplo, CBlasTranspose trans, CBlasDiag diag,     \
                        size_t m, size_t n, SCALAR alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb,        \
                        CUdeviceptr X, size_t ldx, CUstream stream) {                                                  \
    const size_t nRowA = (side == CBlasLeft) ? m : n;                                                                  \
    CHECK_ARG(lda < nRowA, 9);                                                                                         \
    CHECK_ARG(ldb < m, 11);                                                                                            \
    CHECK_ARG(ldx < m, 13);                                                                                            \
    if (m == 0 || n == 0) return CUDA_SUCCESS;                                                                         \
    const double a[2] = {RE(alpha), IM(alpha)};                                                                        \
    IN_BLAS_CONTEXT(handle, chol_trmm2(p, side, uplo, trans, diag, m, n, a, A, lda, B, ldb, X, ldx, stream));          \
  }                                                                                                                    \
  CUresult cu##U##trmm(CUBLAShandle handle, CBlasSide side, CBlasUplo uplo, CBlasTranspose trans, CBlasDiag diag,      \
                       size_t m, size_t n, SCALAR alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb,         \
                       CUstream stream) {                                                                              \
    const size_t nRowA = (side == CBlasLeft) ? m : n;                                                                  \
    CHECK_ARG(lda < nRowA, 9);                                                                                         \
    CHECK_ARG(ldb < m, 11);                                                                                            \
    if (handle == NULL) return CUDA_ERROR_INVALID_VALUE; /* needs the handle's scratch */                              \
    const double a[2] = {RE(alpha), IM(alpha)};                                                                        \
    IN_BLAS_CONTEXT(handle, chol_trmm_inplace(handle, p, side, uplo, trans, diag, m, n, a, A, lda, B, ldb, stream));   \
  }                                                                                                                    \
  CUresult cu##U##trsm(CUBLAShandle handle, CBlasSide side, CBlasUplo uplo, CBlasTranspose trans, CBlasDiag diag,      \
                       size_t m, size_t n, SCALAR alpha, CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb,         \
                       CUstream stream) {                                                                              \
    const size_t nRowA = (side == CBlasLeft) ? m : n;                                                                  \
    CHECK_ARG(lda < nRowA, 9);                                                                                         \
    CHECK_ARG(ldb < m, 11);                                                                                            \
    if (handle == NULL) return CUDA_ERROR_INVALID_VALUE; /* needs the handle's scratch */                              \
    const double a[2] = {RE(alpha), IM(alpha)};                                                                        \
    IN_BLAS_CONTEXT(handle, chol_trsm(handle, p, side, uplo, trans, diag, m, n, a, A, lda, B, ldb, stream));           \
  }

#define RE_R(x) ((double)(x))
#define IM_R(x) (0.0)
#define RE_C(x) ((double)creal(x))
#define IM_C(x) ((double)cimag(x))

GEN_BLAS3(S, 's', float, float, RE_R, IM_R, syrk, 0)
GEN_BLAS3(D, 'd', double, double, RE_R, IM_R, syrk, 0)
GEN_BLAS3(C, 'c', float complex, float, RE_C, IM_C, herk, 1)
GEN_BLAS3(Z, 'z', double complex, double, RE_C, IM_C, herk, 1)
