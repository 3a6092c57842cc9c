/*
 * multigpu_blas.c — the rest of the host-pointer cuMultiGPU* surface:
 *   cuMultiGPU{S,D,C,Z}{gemm, syrk|herk, trmm, trsm}   (reference blas/dgemm.c:432-566,
 *       dsyrk.c:211-308, dtrmm.c:467-586, dtrsm.c:267-389)
 *   cuMultiGPU{S,D,C,Z}{trtri, lauum, potri}            (reference lapack/dtrtri.c:470-530,
 *       dlauum.c:294-352, dpotri.c:48-69)
 *
 * The reference cuts C into MBxNB tiles, deals them round-robin and re-uploads A/B k-panels for
 * every tile.  Here each GPU gets ONE contiguous slab of the independent dimension (columns of
 * C/B, or rows for right-sided triangular ops), uploads its operands once, runs the single-GPU
 * routine on resident data and downloads its slab: PCIe traffic is the minimum (each operand
 * element crosses once per GPU that needs it) and no host BLAS is involved (the reference computes
 * diagonal blocks on the CPU, dsyrk.c:287-288, dtrsm.c:267-389).
 * TRTRI/LAUUM/POTRI on host pointers run staged on the first GPU of the CUmultiGPU for now; their
 * block-cyclic distribution is the next row of SURVEY.md §8(f).
 */
#define _GNU_SOURCE
#include <complex.h>
#include <stdlib.h>
#include <string.h>

#include "chol_internal.h"

CUmultiGPU chol_mgpu_of_lapack(CUmultiGPULAPACKhandle h);
CUmultiGPU chol_mgpu_of_blas(CUmultiGPUBLAShandle h);
CULAPACKhandle chol_mgpu_rank_lapack(CUmultiGPU m, int i);
CUresult chol_mgpu_run_on_all(CUmultiGPU m, CUresult (*fn)(const void *), const void *args, size_t stride, size_t size);

enum { OP_GEMM, OP_RANKK, OP_TRMM, OP_TRSM, OP_TRTRI, OP_LAUUM, OP_POTRI };

typedef struct {
  int op, prec, rank, nranks;
  CULAPACKhandle lap;
  int tA, tB, side, uplo, diag;
  size_t m, n, k;
  double alpha[2], beta[2];
  const char *A; size_t lda;
  const char *B; size_t ldb;
  char *C; size_t ldc;
  long info;
} mb_args;

static CUresult h2d(const char *host, size_t ldh, CUdeviceptr dev, size_t ldd, size_t rows, size_t cols, size_t es, CUstream st) {
  if (rows == 0 || cols == 0) return CUDA_SUCCESS;
  CUDA_MEMCPY2D c;
  memset(&c, 0, sizeof c);
  c.srcMemoryType = CU_MEMORYTYPE_HOST; c.srcHost = host; c.srcPitch = ldh * es;
  c.dstMemoryType = CU_MEMORYTYPE_DEVICE; c.dstDevice = dev; c.dstPitch = ldd * es;
  c.WidthInBytes = rows * es; c.Height = cols;
  return cuMemcpy2DAsync(&c, st);
}
static CUresult d2h(char *host, size_t ldh, CUdeviceptr dev, size_t ldd, size_t rows, size_t cols, size_t es, CUstream st) {
  if (rows == 0 || cols == 0) return CUDA_SUCCESS;
  CUDA_MEMCPY2D c;
  memset(&c, 0, sizeof c);
  c.srcMemoryType = CU_MEMORYTYPE_DEVICE; c.srcDevice = dev; c.srcPitch = ldd * es;
  c.dstMemoryType = CU_MEMORYTYPE_HOST; c.dstHost = host; c.dstPitch = ldh * es;
  c.WidthInBytes = rows * es; c.Height = cols;
  return cuMemcpy2DAsync(&c, st);
}

/* this rank's slab [lo, hi) of a dimension of size n, in multiples of 64 */
static void slab(size_t n, int rank, int P, size_t *lo, size_t *hi) {
  const size_t per = chol_roundup((n + (size_t)P - 1) / (size_t)P, 64);
  *lo = chol_min(n, per * (size_t)rank);
  *hi = chol_min(n, per * (size_t)(rank + 1));
}

static CUresult mb_task(const void *argp) {
  mb_args *a = *(mb_args *const *)argp;
  const int prec = a->prec;
  const size_t es = chol_esize(prec);
  CULAPACKhandle lap = a->lap;
  CUBLAShandle blas = lap->blas_handle;
  CUstream st = lap->s_update;
  CUdeviceptr dA = 0, dB = 0, dC = 0;
  CUresult e = CUDA_SUCCESS;
#define TRY(x) do { if (e == CUDA_SUCCESS) e = (x); } while (0)
  switch (a->op) {
    case OP_GEMM: {
      size_t lo, hi;
      slab(a->n, a->rank, a->nranks, &lo, &hi);
      const size_t ns = hi - lo;
      if (ns == 0) break;
      const size_t ar = a->tA == 'N' ? a->m : a->k, ac = a->tA == 'N' ? a->k : a->m;
      const size_t br = a->tB == 'N' ? a->k : ns, bc = a->tB == 'N' ? ns : a->k;
      const size_t lA = chol_roundup(ar, 2), lB = chol_roundup(br, 2), lC = chol_roundup(a->m, 2);
      const int need_ab = !(a->alpha[0] == 0.0 && a->alpha[1] == 0.0) && a->k > 0;
      TRY(cuMemAlloc(&dA, (lA * ac + 1) * es));
      TRY(cuMemAlloc(&dB, (lB * bc + 1) * es));
      TRY(cuMemAlloc(&dC, (lC * ns + 1) * es));
      if (need_ab) {
        TRY(h2d(a->A, a->lda, dA, lA, ar, ac, es, st));
        TRY(h2d(a->tB == 'N' ? a->B + lo * a->ldb * es : a->B + lo * es, a->ldb, dB, lB, br, bc, es, st));
      }
      if (!(a->beta[0] == 0.0 && a->beta[1] == 0.0)) TRY(h2d(a->C + lo * a->ldc * es, a->ldc, dC, lC, a->m, ns, es, st));
      TRY(chol_gemm(prec, a->tA, a->tB, a->m, ns, need_ab ? a->k : 0, a->alpha, dA, lA, dB, lB, a->beta, dC, lC, dC, lC, st));
      TRY(d2h(a->C + lo * a->ldc * es, a->ldc, dC, lC, a->m, ns, es, st));
      break;
    }
    case OP_RANKK: { /* column slab [lo,hi) of the triangle: rectangle part + diagonal square */
      size_t lo, hi;
      slab(a->n, a->rank, a->nranks, &lo, &hi);
      const size_t ns = hi - lo;
      if (ns == 0) break;
      const size_t n = a->n, k = a->k;
      const size_t ar = a->tA == 'N' ? n : k, ac = a->tA == 'N' ? k : n;
      const size_t lA = chol_roundup(ar, 2), lC = chol_roundup(n, 2);
      const int herm = chol_is_complex(prec), hT = herm ? 'C' : 'T';
      const size_t r0 = a->uplo == 'L' ? lo : 0, r1 = a->uplo == 'L' ? n : hi; /* rows of the slab that matter */
      TRY(cuMemAlloc(&dA, (lA * ac + 1) * es));
      TRY(cuMemAlloc(&dC, (lC * ns + 1) * es));
      TRY(h2d(a->A, a->lda, dA, lA, ar, ac, es, st));
      TRY(h2d(a->C + (lo * a->ldc + r0) * es, a->ldc, chol_at(dC, lC, r0, 0, es), lC, r1 - r0, ns, es, st));
      /* op(A) rows i: trans N -> A(i, :), trans T/C -> A(:, i)^H */
#define AROW(i) (a->tA == 'N' ? chol_at(dA, lA, (i), 0, es) : chol_at(dA, lA, 0, (i), es))
      const int oa = a->tA == 'N' ? 'N' : hT, ob = a->tA == 'N' ? hT : 'N';
      const size_t rect_r0 = a->uplo == 'L' ? hi : 0, rect_rows = a->uplo == 'L' ? n - hi : lo;
      if (rect_rows)
        TRY(chol_gemm_ex(prec, oa, ob, 0, 0, 0, 0, 0, 0, rect_rows, ns, k, a->alpha, AROW(rect_r0), lA, AROW(lo), lA, a->beta,
                         chol_at(dC, lC, rect_r0, 0, es), lC, chol_at(dC, lC, rect_r0, 0, es), lC, st));
      TRY(chol_gemm_ex(prec, oa, ob, 0, 0, 0, 0, a->uplo, herm, ns, ns, k, a->alpha, AROW(lo), lA, AROW(lo), lA, a->beta,
                       chol_at(dC, lC, lo, 0, es), lC, chol_at(dC, lC, lo, 0, es), lC, st));
#undef AROW
      /* download: rectangle + the diagonal square's own triangle (column by column keeps the
       * caller's other triangle untouched) */
      if (rect_rows) TRY(d2h(a->C + (lo * a->ldc + rect_r0) * es, a->ldc, chol_at(dC, lC, rect_r0, 0, es), lC, rect_rows, ns, es, st));
      for (size_t c = 0; c < ns && e == CUDA_SUCCESS; c++) {
        const size_t rr0 = a->uplo == 'L' ? lo + c : lo, cnt = a->uplo == 'L' ? ns - c : c + 1;
        e = cuMemcpyDtoHAsync(a->C + ((lo + c) * a->ldc + rr0) * es, chol_at(dC, lC, rr0, c, es), cnt * es, st);
      }
      break;
    }
    case OP_TRMM:
    case OP_TRSM: {
      /* Left: columns of B are independent; Right: rows of B are */
      const int left = a->side == 'L';
      size_t lo, hi;
      slab(left ? a->n : a->m, a->rank, a->nranks, &lo, &hi);
      const size_t cnt = hi - lo;
      if (cnt == 0) break;
      const size_t na = left ? a->m : a->n, bm = left ? a->m : cnt, bn = left ? cnt : a->n;
      const size_t lA = chol_roundup(na, 2), lB = chol_roundup(bm, 2);
      char *Bh = a->C + (left ? lo * a->ldc * es : lo * es);
      TRY(cuMemAlloc(&dA, (lA * na + 1) * es));
      TRY(cuMemAlloc(&dB, (lB * bn + 1) * es));
      TRY(h2d(a->A, a->lda, dA, lA, na, na, es, st));
      TRY(h2d(Bh, a->ldc, dB, lB, bm, bn, es, st));
      if (a->op == OP_TRSM) {
        TRY(chol_trsm(blas, prec, a->side, a->uplo, a->tA, a->diag, bm, bn, a->alpha, dA, lA, dB, lB, st));
        TRY(d2h(Bh, a->ldc, dB, lB, bm, bn, es, st));
      } else {
        TRY(cuMemAlloc(&dC, (lB * bn + 1) * es));
        TRY(chol_trmm2(prec, a->side, a->uplo, a->tA, a->diag, bm, bn, a->alpha, dA, lA, dB, lB, dC, lB, st));
        TRY(d2h(Bh, a->ldc, dC, lB, bm, bn, es, st));
      }
      break;
    }
    case OP_TRTRI:
    case OP_LAUUM:
    case OP_POTRI: {
      if (a->rank != 0) break;
      const size_t n = a->n, lA = chol_roundup(n, 16);
      TRY(cuMemAlloc(&dA, lA * n * es));
      TRY(h2d(a->A, a->lda, dA, lA, n, n, es, st));
      TRY(cuStreamSynchronize(st));
      if (a->op == OP_TRTRI) TRY(chol_trtri(lap, prec, a->uplo, a->diag, n, dA, lA, &a->info));
      else if (a->op == OP_LAUUM) TRY(chol_lauum(lap, prec, a->uplo, n, dA, lA, &a->info));
      else TRY(chol_potri(lap, prec, a->uplo, n, dA, lA, &a->info));
      TRY(d2h(a->C, a->ldc, dA, lA, n, n, es, st)); /* the other triangle returns unchanged */
      break;
    }
  }
#undef TRY
  const CUresult es2 = cuStreamSynchronize(st);
  if (e == CUDA_SUCCESS) e = es2;
  if (dA) cuMemFree(dA);
  if (dB) cuMemFree(dB);
  if (dC) cuMemFree(dC);
  return e;
}

static CUresult mb_run(CUmultiGPU m, mb_args *proto) {
  const int P = cuMultiGPUGetContextCount(m);
  mb_args *args = calloc((size_t)P, sizeof *args);
  mb_args **ptrs = calloc((size_t)P, sizeof *ptrs);
  if (args == NULL || ptrs == NULL) return CUDA_ERROR_OUT_OF_MEMORY;
  for (int i = 0; i < P; i++) {
    args[i] = *proto;
    args[i].rank = i;
    args[i].nranks = P;
    args[i].lap = chol_mgpu_rank_lapack(m, i);
    ptrs[i] = &args[i];
  }
  const CUresult e = chol_mgpu_run_on_all(m, mb_task, ptrs, sizeof(*ptrs), sizeof(*ptrs));
  proto->info = args[0].info;
  free(args);
  free(ptrs);
  return e;
}

#define CHECK_ARG(cond, pos)                \
  do {                                      \
    if (cond) {                             \
      XERBLA(pos);                          \
      return CUDA_ERROR_INVALID_VALUE;      \
    }                                       \
  } while (false)
#define LAPACK_ARG(cond, pos)               \
  do {                                      \
    if (cond) {                             \
      *info = -(pos);                       \
      XERBLA(pos);                          \
      return CUDA_ERROR_INVALID_VALUE;      \
    }                                       \
  } while (false)

#define RE_R(x) ((double)(x))
#define IM_R(x) (0.0)
#define RE_C(x) ((double)creal(x))
#define IM_C(x) ((double)cimag(x))

#define GEN_MGPU(U, p, SCALAR, REAL, RE, IM, RANKK, ISCPLX)                                                           \
  CUresult cuMultiGPU##U##gemm(CUmultiGPUBLAShandle h, CBlasTranspose tA, CBlasTranspose tB, size_t m, size_t n,       \
                               size_t k, SCALAR alpha, const SCALAR *A, size_t lda, const SCALAR *B, size_t ldb,       \
                               SCALAR beta, SCALAR *C, size_t ldc) {                                                   \
    const size_t nRowA = (tA == CBlasNoTrans) ? m : k, nRowB = (tB == CBlasNoTrans) ? k : n;                           \
    CHECK_ARG(lda < nRowA, 8);                                                                                         \
    CHECK_ARG(ldb < nRowB, 10);                                                                                        \
    CHECK_ARG(ldc < m, 13);                                                                                            \
    mb_args a;                                                                                                         \
    memset(&a, 0, sizeof a);                                                                                           \
    a.op = OP_GEMM; a.prec = p; a.tA = tA; a.tB = tB; a.m = m; a.n = n; a.k = k;                                       \
    a.alpha[0] = RE(alpha); a.alpha[1] = IM(alpha); a.beta[0] = RE(beta); a.beta[1] = IM(beta);                        \
    if (m == 0 || n == 0 || (((a.alpha[0] == 0.0 && a.alpha[1] == 0.0) || k == 0) && a.beta[0] == 1.0 && a.beta[1] == 0.0)) \
      return CUDA_SUCCESS;                                                                                             \
    a.A = (const char *)A; a.lda = lda; a.B = (const char *)B; a.ldb = ldb; a.C = (char *)C; a.ldc = ldc;              \
    return mb_run(chol_mgpu_of_blas(h), &a);                                                                           \
  }                                                                                                                    \
  CUresult cuMultiGPU##U##RANKK(CUmultiGPUBLAShandle h, CBlasUplo uplo, CBlasTranspose trans, size_t n, size_t k,      \
                                REAL alpha, const SCALAR *A, size_t lda, REAL beta, SCALAR *C, size_t ldc) {           \
    const size_t nRowA = (trans == CBlasNoTrans) ? n : k;                                                              \
    CHECK_ARG(ISCPLX && trans == CBlasTrans, 2);                                                                       \
    CHECK_ARG(lda < nRowA, 7);                                                                                         \
    CHECK_ARG(ldc < n, 10);                                                                                            \
    if (n == 0 || ((alpha == 0 || k == 0) && beta == 1)) return CUDA_SUCCESS;                                          \
    mb_args a;                                                                                                         \
    memset(&a, 0, sizeof a);                                                                                           \
    a.op = OP_RANKK; a.prec = p; a.uplo = uplo; a.tA = trans; a.n = n; a.k = k;                                        \
    a.alpha[0] = (double)alpha; a.beta[0] = (double)beta;                                                              \
    a.A = (const char *)A; a.lda = lda; a.C = (char *)C; a.ldc = ldc;                                                  \
    return mb_run(chol_mgpu_of_blas(h), &a);                                                                           \
  }                                                                                                                    \
  CUresult cuMultiGPU##U##trmm(CUmultiGPUBLAShandle h, CBlasSide side, CBlasUplo uplo, CBlasTranspose trans,           \
                               CBlasDiag diag, size_t m, size_t n, SCALAR alpha, const SCALAR *A, size_t lda,          \
                               SCALAR *B, size_t ldb) {                                                                \
    const size_t nRowA = (side == CBlasLeft) ? m : n;                                                                  \
    CHECK_ARG(lda < nRowA, 9);                                                                                         \
    CHECK_ARG(ldb < m, 11);                                                                                            \
    if (m == 0 || n == 0) return CUDA_SUCCESS;                                                                         \
    mb_args a;                                                                                                         \
    memset(&a, 0, sizeof a);                                                                                           \
    a.op = OP_TRMM; a.prec = p; a.side = side; a.uplo = uplo; a.tA = trans; a.diag = diag; a.m = m; a.n = n;           \
    a.alpha[0] = RE(alpha); a.alpha[1] = IM(alpha);                                                                    \
    a.A = (const char *)A; a.lda = lda; a.C = (char *)B; a.ldc = ldb;                                                  \
    return mb_run(chol_mgpu_of_blas(h), &a);                                                                           \
  }                                                                                                                    \
  CUresult cuMultiGPU##U##trsm(CUmultiGPUBLAShandle h, CBlasSide side, CBlasUplo uplo, CBlasTranspose trans,           \
                               CBlasDiag diag, size_t m, size_t n, SCALAR alpha, const SCALAR *A, size_t lda,          \
                               SCALAR *B, size_t ldb) {                                                                \
    const size_t nRowA = (side == CBlasLeft) ? m : n;                                                                  \
    CHECK_ARG(lda < nRowA, 9);                                                                                         \
    CHECK_ARG(ldb < m, 11);                                                                                            \
    if (m == 0 || n == 0) return CUDA_SUCCESS;                                                                         \
    mb_args a;                                                                                                         \
    memset(&a, 0, sizeof a);                                                                                           \
    a.op = OP_TRSM; a.prec = p; a.side = side; a.uplo = uplo; a.tA = trans; a.diag = diag; a.m = m; a.n = n;           \
    a.alpha[0] = RE(alpha); a.alpha[1] = IM(alpha);                                                                    \
    a.A = (const char *)A; a.lda = lda; a.C = (char *)B; a.ldc = ldb;                                                  \
    return mb_run(chol_mgpu_of_blas(h), &a);                                                                           \
  }                                                                                                                    \
  CUresult cuMultiGPU##U##trtri(CUmultiGPULAPACKhandle h, CBlasUplo uplo, CBlasDiag diag, size_t n, SCALAR *A,         \
                                size_t lda, long *info) {                                                              \
    *info = 0;                                                                                                         \
    LAPACK_ARG(lda < n, 5);                                                                                            \
    if (n == 0) return CUDA_SUCCESS;                                                                                   \
    mb_args a;                                                                                                         \
    memset(&a, 0, sizeof a);                                                                                           \
    a.op = OP_TRTRI; a.prec = p; a.uplo = uplo; a.diag = diag; a.n = n;                                                \
    a.A = (const char *)A; a.lda = lda; a.C = (char *)A; a.ldc = lda;                                                  \
    const CUresult e = mb_run(chol_mgpu_of_lapack(h), &a);                                                             \
    *info = a.info;                                                                                                    \
    return e;                                                                                                          \
  }                                                                                                                    \
  CUresult cuMultiGPU##U##lauum(CUmultiGPULAPACKhandle h, CBlasUplo uplo, size_t n, SCALAR *A, size_t lda,             \
                                long *info) {                                                                          \
    *info = 0;                                                                                                         \
    LAPACK_ARG(lda < n, 4);                                                                                            \
    if (n == 0) return CUDA_SUCCESS;                                                                                   \
    mb_args a;                                                                                                         \
    memset(&a, 0, sizeof a);                                                                                           \
    a.op = OP_LAUUM; a.prec = p; a.uplo = uplo; a.n = n;                                                               \
    a.A = (const char *)A; a.lda = lda; a.C = (char *)A; a.ldc = lda;                                                  \
    const CUresult e = mb_run(chol_mgpu_of_lapack(h), &a);                                                             \
    *info = a.info;                                                                                                    \
    return e;                                                                                                          \
  }                                                                                                                    \
  CUresult cuMultiGPU##U##potri(CUmultiGPULAPACKhandle h, CBlasUplo uplo, size_t n, SCALAR *A, size_t lda,             \
                                long *info) {                                                                          \
    *info = 0;                                                                                                         \
    LAPACK_ARG(lda < n, 4);                                                                                            \
    if (n == 0) return CUDA_SUCCESS;                                                                                   \
    mb_args a;                                                                                                         \
    memset(&a, 0, sizeof a);                                                                                           \
    a.op = OP_POTRI; a.prec = p; a.uplo = uplo; a.n = n;                                                               \
    a.A = (const char *)A; a.lda = lda; a.C = (char *)A; a.ldc = lda;                                                  \
    const CUresult e = mb_run(chol_mgpu_of_lapack(h), &a);                                                             \
    *info = a.info;                                                                                                    \
    return e;                                                                                                          \
  }

GEN_MGPU(S, 's', float, float, RE_R, IM_R, syrk, 0)
GEN_MGPU(D, 'd', double, double, RE_R, IM_R, syrk, 0)
GEN_MGPU(C, 'c', float complex, float, RE_C, IM_C, herk, 1)
GEN_MGPU(Z, 'z', double complex, double, RE_C, IM_C, herk, 1)
