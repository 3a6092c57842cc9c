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
 * TRTRI/LAUUM/POTRI on host pointers: with one GPU staged on it; with more, the distributed block-cyclic sweeps of
 * host/multigpu.c.  An exchange-free split by output slab is kept further down
 * (CHOL_MGPU_INVERSE=slab; PCIe-bound: 2 B200, N = 32768: 0.91-0.99 s against 0.63 s distributed).
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
      /* download: rectangle + the diagonal square.  The whole square (both triangles) was uploaded and the
       * cuplo-masked GEMM never writes the other triangle, so one 2-D copy returns the caller's other triangle
       * bit-identical (the argument the POTRF gather already uses) - no per-column copies. */
      if (rect_rows) TRY(d2h(a->C + (lo * a->ldc + rect_r0) * es, a->ldc, chol_at(dC, lC, rect_r0, 0, es), lC, rect_rows, ns, es, st));
      TRY(d2h(a->C + (lo * a->ldc + lo) * es, a->ldc, chol_at(dC, lC, lo, 0, es), lC, ns, ns, es, st));
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
  if (m == NULL) return CUDA_ERROR_INVALID_VALUE;
  const int P = cuMultiGPUGetContextCount(m);
  mb_args *args = calloc((size_t)P, sizeof *args);
  mb_args **ptrs = calloc((size_t)P, sizeof *ptrs);
  if (args == NULL || ptrs == NULL) {
    free(args);
    free(ptrs);
    return CUDA_ERROR_OUT_OF_MEMORY;
  }
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

/* =============================================================================================
 * Slab-distributed TRTRI / LAUUM / POTRI on host pointers (CHOL_MGPU_INVERSE=slab; verified on 2 B200,
 * profiles/r02_call8_2gpu_summary.txt; the default with more than one GPU is the distributed form).
 *
 * Both operations split by output slab with no GPU-to-GPU exchange:
 *   TRTRI lower: columns J = [lo,hi) of inv(L) depend on T = L[lo:n, lo:n] only:
 *                  X11 = inv(T11)            (blocked TRTRI of the w x w diagonal block)
 *                  X21 = -inv(T22) (T21 X11) (TRMM2 then TRSM)
 *         upper: columns J of inv(U) depend on T = U[0:hi, 0:hi]:  X22 = inv(T22), X12 = -inv(T11) (T12 X22)
 *   LAUUM lower: rows I = [lo,hi) of L^H L are (L[lo:n, I])^H L[lo:n, 0:hi)   (the block-row step of
 *                chol_lauum with ib = hi-lo), upper: columns I of U U^H are U[0:hi, lo:n] (U[I, lo:n])^H.
 * Every slab READS parts of the matrix that other slabs WRITE, so all uploads finish (phase 1)
 * before any result is downloaded (phase 2).  Slab boundaries equalise the work (cubic / quadratic
 * cost profiles), in multiples of 64.  POTRI runs the two operations back to back through the
 * host matrix.  CHOL_MGPU_INVERSE_SLABS=k gives every GPU k slabs (tests: several slabs on one GPU).
 * ===========================================================================================*/
typedef struct {
  int op, prec, uplo, diag;
  CULAPACKhandle lap;
  size_t n, lda, lo, hi;
  char *A;
  CUdeviceptr dT, dY, dD;
  long info;
} inv_slab;

typedef struct {
  inv_slab *slabs;
  int nslabs, rank, nranks, phase;
} inv_job;

static const double ONE[2] = {1.0, 0.0};
static const double ZERO[2] = {0.0, 0.0};
static const double MONE[2] = {-1.0, 0.0};

/* cuMultiGPU*{trtri,lauum,potri} with more than one GPU: the distributed block-cyclic sweeps of host/multigpu.c (one NCCL
 * exchange per block step, like POTRF).  With one GPU the plain single-GPU routine on a staged copy does the same work.
 * CHOL_MGPU_INVERSE=slab selects the exchange-free split by output slab further down (2 B200, N = 32768: 0.99 s against
 * 1.18 s staged - PCIe-bound, profiles/r02_call8_2gpu_summary.txt), =staged the one-GPU form, =dist this default. */
CUresult chol_mgpu_inverse_dist(CUmultiGPU m, int op, int prec, int uplo, int diag, size_t n, void *A, size_t lda, long *info);
static int inverse_mode(CUmultiGPU m) { /* 0 staged, 1 slab, 2 distributed */
  const char *e = getenv("CHOL_MGPU_INVERSE");
  if (e != NULL && strcmp(e, "slab") == 0) return 1;
  if (e != NULL && strcmp(e, "staged") == 0) return 0;
  if (e != NULL && strcmp(e, "dist") == 0) return 2;
  return (m != NULL && cuMultiGPUGetContextCount(m) > 1) ? 2 : 0;
}
int chol_inverse_slabs_enabled(CUmultiGPU m) { return inverse_mode(m) == 1; }

/* cost profile over the split dimension: 0 -> (n-x)^2, 1 -> x^2, 2 -> x (n-x) */
void chol_inverse_slab_bounds(int profile, size_t n, int nslabs, size_t *bounds /* nslabs + 1 */) {
  const size_t g = 64, nblk = (n + g - 1) / g;
  double total = 0.0;
  for (size_t b = 0; b < nblk; b++) {
    const double x = (double)chol_min(n, b * g + g / 2), len = (double)(chol_min(n, (b + 1) * g) - b * g);
    const double d = profile == 0 ? ((double)n - x) * ((double)n - x) : (profile == 1 ? x * x : x * ((double)n - x));
    total += (d + 1.0) * len;
  }
  bounds[0] = 0;
  int r = 1;
  double acc = 0.0;
  for (size_t b = 0; b < nblk && r < nslabs; b++) {
    const double x = (double)chol_min(n, b * g + g / 2), len = (double)(chol_min(n, (b + 1) * g) - b * g);
    const double d = profile == 0 ? ((double)n - x) * ((double)n - x) : (profile == 1 ? x * x : x * ((double)n - x));
    acc += (d + 1.0) * len;
    while (r < nslabs && acc >= total * (double)r / (double)nslabs) bounds[r++] = chol_min(n, (b + 1) * g);
  }
  while (r <= nslabs) bounds[r++] = n;
}

static CUresult inv_upload(inv_slab *a, CUstream st) {
  const size_t es = chol_esize(a->prec), n = a->n, lo = a->lo, hi = a->hi, w = hi - lo;
  if (w == 0) return CUDA_SUCCESS;
  if (a->op == OP_TRTRI) {
    /* T = the uplo triangle of A[o:o+s, o:o+s], uploaded as block-column trapezoids */
    const size_t o = a->uplo == 'L' ? lo : 0, s = a->uplo == 'L' ? n - lo : hi, ldt = chol_roundup(s, 16), bw = 2048;
    CU_ERROR_CHECK(cuMemAlloc(&a->dT, ldt * s * es));
    for (size_t c0 = 0; c0 < s; c0 += bw) {
      const size_t cb = chol_min(bw, s - c0);
      const size_t r0 = a->uplo == 'L' ? c0 : 0, rows = a->uplo == 'L' ? s - c0 : c0 + cb;
      CU_ERROR_CHECK(h2d(a->A + ((o + c0) * a->lda + o + r0) * es, a->lda, chol_at(a->dT, ldt, r0, c0, es), ldt, rows, cb, es, st));
    }
    /* the slab's own w x w diagonal block with BOTH triangles: its result goes back as one 2-D copy of the whole
     * square (the in-place inverse never writes the other triangle, so the caller gets those bits back unchanged) */
    const size_t d0 = a->uplo == 'L' ? 0 : lo;
    CU_ERROR_CHECK(h2d(a->A + (lo * a->lda + lo) * es, a->lda, chol_at(a->dT, ldt, d0, d0, es), ldt, w, w, es, st));
    const size_t r = s - w;
    if (r) CU_ERROR_CHECK(cuMemAlloc(&a->dY, chol_roundup(r, 16) * w * es));
  } else { /* LAUUM: lower R = A[lo:n, 0:hi) (k x hi), upper R = A[0:hi, lo:n) (hi x k) */
    const size_t k = n - lo;
    const size_t rows = a->uplo == 'L' ? k : hi, cols = a->uplo == 'L' ? hi : k, ldr = chol_roundup(rows, 16);
    CU_ERROR_CHECK(cuMemAlloc(&a->dT, ldr * cols * es));
    CU_ERROR_CHECK(h2d(a->uplo == 'L' ? a->A + lo * es : a->A + lo * a->lda * es, a->lda, a->dT, ldr, rows, cols, es, st));
    if (lo) CU_ERROR_CHECK(cuMemAlloc(&a->dY, chol_roundup(a->uplo == 'L' ? w : lo, 16) * (a->uplo == 'L' ? lo : w) * es));
    CU_ERROR_CHECK(cuMemAlloc(&a->dD, chol_roundup(w, 16) * w * es));
    /* the product block starts as the caller's square: the masked HERK writes its own triangle only, and the whole
     * square is downloaded with one copy */
    CU_ERROR_CHECK(h2d(a->A + (lo * a->lda + lo) * es, a->lda, a->dD, chol_roundup(w, 16), w, w, es, st));
  }
  return CUDA_SUCCESS;
}

/* the w x w diagonal block of a slab -> the caller's matrix.  The device copy holds the caller's other triangle
 * (uploaded with the slab, never written by the triangle-masked kernels), so ONE 2-D copy of the whole square
 * returns the result and leaves the other triangle bit-identical - instead of one small copy per column. */
static CUresult inv_download_triangle(const inv_slab *a, CUdeviceptr D, size_t ldd, CUstream st) {
  const size_t es = chol_esize(a->prec), lo = a->lo, w = a->hi - a->lo;
  return d2h(a->A + (lo * a->lda + lo) * es, a->lda, D, ldd, w, w, es, st);
}

static CUresult inv_compute(inv_slab *a, CUstream st) {
  const int prec = a->prec, uplo = a->uplo;
  const size_t es = chol_esize(prec), n = a->n, lo = a->lo, hi = a->hi, w = hi - lo;
  if (w == 0) return CUDA_SUCCESS;
  if (a->op == OP_TRTRI) {
    const size_t s = uplo == 'L' ? n - lo : hi, ldt = chol_roundup(s, 16), r = s - w, ldy = chol_roundup(r ? r : 1, 16);
    const size_t d0 = uplo == 'L' ? 0 : lo;  /* position of the w x w diagonal block inside T */
    const size_t t0 = uplo == 'L' ? w : 0;   /* position of the other (r x r) triangle */
    const CUdeviceptr Tww = chol_at(a->dT, ldt, d0, d0, es), Trr = chol_at(a->dT, ldt, t0, t0, es);
    const CUdeviceptr Trw = uplo == 'L' ? chol_at(a->dT, ldt, w, 0, es) : chol_at(a->dT, ldt, 0, lo, es); /* r x w */
    long info = 0;
    CU_ERROR_CHECK(chol_trtri(a->lap, prec, uplo, a->diag, w, Tww, ldt, &info));
    if (info > 0) {
      a->info = (long)lo + info;
      return CUDA_SUCCESS; /* this slab is left untouched, like the block the reference stops in */
    }
    if (r) {
      CU_ERROR_CHECK(chol_trmm2(prec, 'R', uplo, 'N', a->diag, r, w, ONE, Tww, ldt, Trw, ldt, a->dY, ldy, st));
      CU_ERROR_CHECK(chol_trsm(a->lap->blas_handle, prec, 'L', uplo, 'N', a->diag, r, w, MONE, Trr, ldt, a->dY, ldy, st));
      CU_ERROR_CHECK(d2h(a->A + (lo * a->lda + (uplo == 'L' ? hi : 0)) * es, a->lda, a->dY, ldy, r, w, es, st));
    }
    CU_ERROR_CHECK(inv_download_triangle(a, Tww, ldt, st));
  } else {
    const int herm = chol_is_complex(prec);
    const size_t k = n - lo, ldd = chol_roundup(w, 16);
    if (uplo == 'L') {
      const size_t ldr = chol_roundup(k, 16), ldy = chol_roundup(w, 16);
      const CUdeviceptr Aii = chol_at(a->dT, ldr, 0, lo, es);
      if (lo) {
        CU_ERROR_CHECK(chol_gemm_ex(prec, 'C', 'N', 'L', 0, 0, 0, 0, 0, w, lo, k, ONE, Aii, ldr, a->dT, ldr, ZERO, 0, 0, a->dY, ldy, st));
        CU_ERROR_CHECK(d2h(a->A + lo * es, a->lda, a->dY, ldy, w, lo, es, st));
      }
      CU_ERROR_CHECK(chol_gemm_ex(prec, 'C', 'N', 'L', 0, 'L', 0, 'L', herm, w, w, k, ONE, Aii, ldr, Aii, ldr, ZERO, 0, 0, a->dD, ldd, st));
    } else {
      const size_t ldr = chol_roundup(hi, 16), ldy = chol_roundup(lo ? lo : 1, 16);
      const CUdeviceptr Aii = chol_at(a->dT, ldr, lo, 0, es);
      if (lo) {
        CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'C', 0, 0, 'U', 0, 0, 0, lo, w, k, ONE, a->dT, ldr, Aii, ldr, ZERO, 0, 0, a->dY, ldy, st));
        CU_ERROR_CHECK(d2h(a->A + lo * a->lda * es, a->lda, a->dY, ldy, lo, w, es, st));
      }
      CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'C', 'U', 0, 'U', 0, 'U', herm, w, w, k, ONE, Aii, ldr, Aii, ldr, ZERO, 0, 0, a->dD, ldd, st));
    }
    CU_ERROR_CHECK(inv_download_triangle(a, a->dD, ldd, st));
  }
  return CUDA_SUCCESS;
}

/* phase 0: upload, 1: compute + download, 2: release; a rank handles slabs rank, rank+P, ... */
static CUresult inv_task(const void *argp) {
  const inv_job *job = *(inv_job *const *)argp;
  CUresult first = CUDA_SUCCESS;
  for (int v = job->rank; v < job->nslabs; v += job->nranks) {
    inv_slab *a = &job->slabs[v];
    CUstream st = a->lap->s_update;
    CUresult e = CUDA_SUCCESS;
    if (job->phase == 0) e = inv_upload(a, st);
    else if (job->phase == 1) e = inv_compute(a, st);
    else {
      if (a->dT) cuMemFree(a->dT);
      if (a->dY) cuMemFree(a->dY);
      if (a->dD) cuMemFree(a->dD);
      a->dT = a->dY = a->dD = 0;
    }
    if (job->phase != 2) {
      const CUresult es2 = cuStreamSynchronize(st);
      if (e == CUDA_SUCCESS) e = es2;
    }
    if (first == CUDA_SUCCESS) first = e;
  }
  return first;
}

static CUresult inv_run_op(CUmultiGPU m, int op, int prec, int uplo, int diag, size_t n, char *A, size_t lda, long *info) {
  const int P = cuMultiGPUGetContextCount(m);
  const char *ev = getenv("CHOL_MGPU_INVERSE_SLABS");
  const int per = (ev != NULL && atoi(ev) > 0) ? atoi(ev) : 1, V = P * per;
  inv_slab *slabs = calloc((size_t)V, sizeof *slabs);
  inv_job *jobs = calloc((size_t)P, sizeof *jobs);
  inv_job **ptrs = calloc((size_t)P, sizeof *ptrs);
  size_t *bounds = calloc((size_t)V + 1, sizeof *bounds);
  if (!slabs || !jobs || !ptrs || !bounds) {
    free(slabs); free(jobs); free(ptrs); free(bounds);
    return CUDA_ERROR_OUT_OF_MEMORY;
  }
  chol_inverse_slab_bounds(op == OP_LAUUM ? 2 : (uplo == 'L' ? 0 : 1), n, V, bounds);
  for (int v = 0; v < V; v++) {
    slabs[v].op = op; slabs[v].prec = prec; slabs[v].uplo = uplo; slabs[v].diag = diag;
    slabs[v].lap = chol_mgpu_rank_lapack(m, v % P);
    slabs[v].n = n; slabs[v].lda = lda; slabs[v].A = A;
    slabs[v].lo = bounds[v]; slabs[v].hi = bounds[v + 1];
  }
  CUresult e = CUDA_SUCCESS;
  for (int phase = 0; phase < 3; phase++) {
    if (phase == 1 && e != CUDA_SUCCESS) continue; /* a failed upload: skip the compute, still release */
    for (int i = 0; i < P; i++) {
      jobs[i].slabs = slabs; jobs[i].nslabs = V; jobs[i].rank = i; jobs[i].nranks = P; jobs[i].phase = phase;
      ptrs[i] = &jobs[i];
    }
    const CUresult pe = chol_mgpu_run_on_all(m, inv_task, ptrs, sizeof(*ptrs), sizeof(*ptrs));
    if (e == CUDA_SUCCESS) e = pe;
  }
  /* the reference reports the first zero pivot its sweep meets: bottom-up for lower, top-down for upper */
  long found = 0;
  for (int v = 0; v < V; v++)
    if (slabs[v].info > 0 && (found == 0 || (uplo == 'L' ? slabs[v].info > found : slabs[v].info < found))) found = slabs[v].info;
  *info = found;
  free(slabs); free(jobs); free(ptrs); free(bounds);
  return e;
}

CUresult chol_mgpu_inverse_slabs(CUmultiGPU m, int op, int prec, int uplo, int diag, size_t n, void *A, size_t lda, long *info) {
  if (m == NULL) return CUDA_ERROR_INVALID_VALUE;
  *info = 0;
  if (op == OP_POTRI) {
    CU_ERROR_CHECK(inv_run_op(m, OP_TRTRI, prec, uplo, 'N', n, (char *)A, lda, info));
    if (*info != 0) return CUDA_SUCCESS;
    return inv_run_op(m, OP_LAUUM, prec, uplo, 'N', n, (char *)A, lda, info);
  }
  return inv_run_op(m, op, prec, uplo, diag, n, (char *)A, lda, info);
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
    if (inverse_mode(chol_mgpu_of_lapack(h)) == 2)                                                                    \
      return chol_mgpu_inverse_dist(chol_mgpu_of_lapack(h), 0, p, uplo, diag, n, A, lda, info);                      \
    if (chol_inverse_slabs_enabled(chol_mgpu_of_lapack(h)))                                                                                  \
      return chol_mgpu_inverse_slabs(chol_mgpu_of_lapack(h), OP_TRTRI, p, uplo, diag, n, A, lda, info);                \
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
    if (inverse_mode(chol_mgpu_of_lapack(h)) == 2)                                                                    \
      return chol_mgpu_inverse_dist(chol_mgpu_of_lapack(h), 1, p, uplo, 'N', n, A, lda, info);                      \
    if (chol_inverse_slabs_enabled(chol_mgpu_of_lapack(h)))                                                                                  \
      return chol_mgpu_inverse_slabs(chol_mgpu_of_lapack(h), OP_LAUUM, p, uplo, 'N', n, A, lda, info);                 \
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
    if (inverse_mode(chol_mgpu_of_lapack(h)) == 2)                                                                    \
      return chol_mgpu_inverse_dist(chol_mgpu_of_lapack(h), 2, p, uplo, 'N', n, A, lda, info);                      \
    if (chol_inverse_slabs_enabled(chol_mgpu_of_lapack(h)))                                                                                  \
      return chol_mgpu_inverse_slabs(chol_mgpu_of_lapack(h), OP_POTRI, p, uplo, 'N', n, A, lda, info);                 \
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
