/*
 * A plain C caller written against the reference's API (cf. reference test/lapack/cudpotrf.c:47-139):
 * own CUDA context (cuCtxCreate, not a framework's primary context), pitched device matrix,
 * cuLAPACKCreate / cuDpotrf / cuDlogdet / cuDpotri / cuLAPACKDestroy.  Links against
 * libcholesky_b200.so only.  Prints "PASSED" and returns 0 on success.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "lapack.h"
#include "error.h"

#define CHECK(call)                                                                   \
  do {                                                                                \
    CUresult e__ = (call);                                                            \
    if (e__ != CUDA_SUCCESS) {                                                        \
      fprintf(stderr, "%s failed: %s\n", #call, cuCholGetErrorString(e__));           \
      return 2;                                                                       \
    }                                                                                 \
  } while (0)

int main(int argc, char **argv) {
  const size_t n = argc > 1 ? (size_t)atol(argv[1]) : 300, lda = (n + 1u) & ~(size_t)1u; /* as the reference's tests */
  double *A = malloc(lda * n * sizeof *A), *L = malloc(lda * n * sizeof *L);
  srand(0);
  /* SPD: diagonally dominant symmetric matrix */
  for (size_t j = 0; j < n; j++)
    for (size_t i = j; i < n; i++) {
      const double v = (i == j) ? (double)n + 1.0 : (double)rand() / RAND_MAX - 0.5;
      A[j * lda + i] = v;
      A[i * lda + j] = v;
    }
  CUdevice dev;
  CUcontext ctx;
  CHECK(cuInit(0));
  CHECK(cuDeviceGet(&dev, 0));
  CHECK(cuCtxCreate(&ctx, 0, dev));
  CULAPACKhandle h;
  CHECK(cuLAPACKCreate(&h));
  CUdeviceptr dA;
  CHECK(cuMemAlloc(&dA, lda * n * sizeof(double)));
  CHECK(cuMemcpyHtoD(dA, A, lda * n * sizeof(double)));
  long info = -1;
  CHECK(cuDpotrf(h, CBlasLower, n, dA, lda, &info));
  if (info != 0) { fprintf(stderr, "info = %ld\n", info); return 3; }
  CHECK(cuMemcpyDtoH(L, dA, lda * n * sizeof(double)));
  /* residual max |A - L L^T| over the lower triangle, and logdet against the host sum */
  double err = 0.0, ld_ref = 0.0;
  for (size_t j = 0; j < n; j++) {
    ld_ref += 2.0 * log(L[j * lda + j]);
    for (size_t i = j; i < n; i++) {
      double s = 0.0;
      for (size_t k = 0; k <= j; k++) s += L[k * lda + i] * L[k * lda + j];
      const double d = fabs(s - A[j * lda + i]);
      if (d > err) err = d;
    }
  }
  double ld = 0.0;
  CHECK(cuDlogdet(h, dA, lda + 1, n, &ld, 0));
  CHECK(cuCtxSynchronize());
  CHECK(cuDpotri(h, CBlasLower, n, dA, lda, &info));
  if (info != 0) { fprintf(stderr, "potri info = %ld\n", info); return 3; }
  CHECK(cuMemcpyDtoH(L, dA, lda * n * sizeof(double)));
  /* max |A * Ainv - I| with the symmetric completion of the lower triangle */
  double err2 = 0.0;
  for (size_t i = 0; i < n; i += 37)
    for (size_t j = 0; j < n; j += 41) {
      double s = 0.0;
      for (size_t k = 0; k < n; k++) s += A[k * lda + i] * ((k >= j) ? L[j * lda + k] : L[k * lda + j]);
      const double d = fabs(s - (i == j ? 1.0 : 0.0));
      if (d > err2) err2 = d;
    }
  /* argument error convention: lda < n -> info = -4, CUDA_ERROR_INVALID_VALUE, xerbla hook called */
  xerbla = NULL;
  const CUresult bad = cuDpotrf(h, CBlasLower, n, dA, n - 1, &info);
  CHECK(cuMemFree(dA));
  CHECK(cuLAPACKDestroy(h));
  CHECK(cuCtxDestroy(ctx));
  printf("n=%zu  max|A-LL^T| = %.3e  |logdet - ref| = %.3e  max|A*Ainv-I| = %.3e  bad-lda: rc=%d info=%ld\n", n, err,
         fabs(ld - ld_ref), err2, (int)bad, info);
  const int ok = err < 1e-11 * n && fabs(ld - ld_ref) < 1e-9 && err2 < 1e-10 && bad == CUDA_ERROR_INVALID_VALUE && info == -4;
  printf(ok ? "PASSED\n" : "FAILED\n");
  free(A);
  free(L);
  return ok ? 0 : 1;
}
