/* mgpu_potri_chain.c — a plain-C caller of the host-pointer multi-GPU API the way the reference's tests drive it
 * (test/lapack/cumultigpudpotrf.c:47-60, cumultigpudpotri.c): cuMultiGPUDpotrf -> cuMultiGPUDpotri on every visible
 * device, both triangles, then cuMultiGPUDgemm forms A * inv(A) from the symmetrised result and the program checks it
 * against the identity.  TEST PROGRAM: built and run by tests/test_host_drivers_mock.py against the mock back end
 * (1 and 3 mock devices) and by scripts/mock_tsan.sh under ThreadSanitizer; it links against nothing but the library.
 *     ./mgpu_potri_chain [n]                                                                                         */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "blas.h"
#include "cumultigpu.h"
#include "lapack.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    CUresult e_ = (call);                                                        \
    if (e_ != CUDA_SUCCESS) {                                                    \
      printf("FAILED: %s -> %d (line %d)\n", #call, (int)e_, __LINE__);          \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

int main(int argc, char **argv) {
  const size_t n = argc > 1 ? (size_t)atol(argv[1]) : 700, lda = n + 3;
  int ndev = 0;
  CHECK(cuInit(0));
  CHECK(cuDeviceGetCount(&ndev));
  if (ndev > 16) ndev = 16;
  CUdevice devs[16];
  for (int i = 0; i < ndev; i++) CHECK(cuDeviceGet(&devs[i], i));
  CUmultiGPU m;
  CUmultiGPULAPACKhandle lh;
  CUmultiGPUBLAShandle bh;
  CHECK(cuMultiGPUCreate(&m, devs, ndev));
  CHECK(cuMultiGPULAPACKCreate(&lh, m));
  CHECK(cuMultiGPUBLASCreate(&bh, m));
  double *A0 = malloc(lda * n * sizeof *A0), *A = malloc(lda * n * sizeof *A), *P = malloc(lda * n * sizeof *P);
  srand(0);
  for (size_t j = 0; j < n; j++)
    for (size_t i = j; i < n; i++) {
      const double v = (i == j) ? (double)n + 1.0 : (double)rand() / RAND_MAX - 0.5;
      A0[j * lda + i] = A0[i * lda + j] = v;
    }
  double worst_all = 0.0;
  for (int pass = 0; pass < 2; pass++) {
    const CBlasUplo uplo = pass ? CBlasUpper : CBlasLower;
    for (size_t i = 0; i < lda * n; i++) A[i] = A0[i];
    long info = -99;
    CHECK(cuMultiGPUDpotrf(lh, uplo, n, A, lda, &info));
    if (info != 0) { printf("FAILED: potrf info %ld\n", info); return 1; }
    CHECK(cuMultiGPUDpotri(lh, uplo, n, A, lda, &info));
    if (info != 0) { printf("FAILED: potri info %ld\n", info); return 1; }
    /* the other triangle still holds the input: check, then mirror the inverse into it */
    for (size_t j = 0; j < n; j++)
      for (size_t i = j + 1; i < n; i++) {
        double *stored = pass ? &A[i * lda + j] : &A[j * lda + i], *other = pass ? &A[j * lda + i] : &A[i * lda + j];
        if (*other != A0[i * lda + j]) { printf("FAILED: other triangle changed at (%zu, %zu)\n", i, j); return 1; }
        *other = *stored;
      }
    CHECK(cuMultiGPUDgemm(bh, CBlasNoTrans, CBlasNoTrans, n, n, n, 1.0, A0, lda, A, lda, 0.0, P, lda));
    CHECK(cuMultiGPUBLASSynchronize(bh));
    double worst = 0.0;
    for (size_t j = 0; j < n; j++)
      for (size_t i = 0; i < n; i++) {
        const double d = fabs(P[j * lda + i] - (i == j ? 1.0 : 0.0));
        if (d > worst) worst = d;
      }
    printf("%s: max |A inv(A) - I| = %.3e on %d device(s)\n", pass ? "upper" : "lower", worst, ndev);
    if (worst > worst_all) worst_all = worst;
  }
  CHECK(cuMultiGPUBLASDestroy(bh));
  CHECK(cuMultiGPULAPACKDestroy(lh));
  CHECK(cuMultiGPUDestroy(m));
  free(A0);
  free(A);
  free(P);
  if (!(worst_all <= 32.0 * (double)n * 2.220446049250313e-16)) { printf("FAILED: %.3e\n", worst_all); return 1; }
  printf("PASSED\n");
  return 0;
}
