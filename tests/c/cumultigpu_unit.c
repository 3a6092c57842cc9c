/*
 * Unit test of the CUmultiGPU / CUtask runtime from a plain C caller, after the reference's own two
 * unit tests (test/multigpu/cumultigpu.c:24-55, test/multigpu/cutask.c): context count, a task that
 * runs on the worker thread WITH the device's context current, an out-of-range worker index ->
 * CUDA_ERROR_INVALID_VALUE, results plumbed back through cuTaskDestroy; plus what the reference's
 * runtime documents but does not test: a failing background task is latched and surfaces once at
 * cuMultiGPUSynchronize (multigpu.c:139-147), and tasks on one worker run in submission order.
 * Then the multi-GPU LAPACK handle on every visible device factors a small host matrix.
 * Links against libcholesky_b200.so only.  Prints "PASSED" and returns 0 on success.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "cumultigpu.h"
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
#define EXPECT(cond)                                                \
  do {                                                              \
    if (!(cond)) {                                                  \
      fprintf(stderr, "line %d: expected %s\n", __LINE__, #cond);   \
      return 3;                                                     \
    }                                                               \
  } while (0)

static CUresult whoami(const void *args) {
  (void)args;
  CUdevice d;
  const CUresult e = cuCtxGetDevice(&d); /* fails unless the worker has a current context */
  if (e != CUDA_SUCCESS) return e;
  char name[256];
  size_t bytes;
  if (cuDeviceGetName(name, 256, d) == CUDA_SUCCESS && cuDeviceTotalMem(&bytes, d) == CUDA_SUCCESS)
    printf("worker on %s, %zu MB\n", name, bytes >> 20);
  return CUDA_SUCCESS;
}

static int order_log[8], order_n = 0;
static CUresult record_order(const void *args) {
  order_log[order_n++] = *(const int *)args; /* one worker thread: no race */
  return CUDA_SUCCESS;
}

static CUresult always_fails(const void *args) {
  (void)args;
  return CUDA_ERROR_NOT_READY;
}

static void quiet(const char *a, const char *b, const char *c, int d, int e, const char *(*f)(int)) {
  (void)a; (void)b; (void)c; (void)d; (void)e; (void)f;
}

int main(void) {
  CHECK(cuInit(0));
  int ndev = 0;
  CHECK(cuDeviceGetCount(&ndev));
  EXPECT(ndev >= 1);
  if (ndev > 8) ndev = 8;
  CUdevice devs[8];
  for (int i = 0; i < ndev; i++) CHECK(cuDeviceGet(&devs[i], i));

  /* ---- one worker: the reference's cumultigpu.c ---- */
  CUmultiGPU m;
  CHECK(cuMultiGPUCreate(&m, devs, 1));
  EXPECT(cuMultiGPUGetContextCount(m) == 1);
  CUtask t;
  CHECK(cuTaskCreate(&t, whoami, NULL, 0));
  EXPECT(cuMultiGPURunTask(m, 1, t) == CUDA_ERROR_INVALID_VALUE);
  EXPECT(cuMultiGPURunTask(m, -1, t) == CUDA_ERROR_INVALID_VALUE);
  EXPECT(cuMultiGPURunTask(m, 0, t) == CUDA_SUCCESS);
  CHECK(cuMultiGPUSynchronize(m));
  CUresult r = CUDA_ERROR_UNKNOWN;
  CHECK(cuTaskDestroy(t, &r));
  EXPECT(r == CUDA_SUCCESS);

  /* submission order on one worker, arguments copied at creation */
  CUtask q[4];
  for (int i = 0; i < 4; i++) {
    int v = 10 + i;
    CHECK(cuTaskCreate(&q[i], record_order, &v, sizeof v));
    v = -1;
    CHECK(cuMultiGPURunTask(m, 0, q[i]));
  }
  for (int i = 0; i < 4; i++) CHECK(cuTaskDestroy(q[i], &r));
  EXPECT(order_n == 4 && order_log[0] == 10 && order_log[1] == 11 && order_log[2] == 12 && order_log[3] == 13);

  /* a failing task: its own result, and latched for the next synchronize, once */
  errorHandler = quiet;
  CHECK(cuTaskCreate(&t, always_fails, NULL, 0));
  CHECK(cuMultiGPURunTask(m, 0, t));
  CHECK(cuTaskDestroy(t, &r));
  EXPECT(r == CUDA_ERROR_NOT_READY);
  EXPECT(cuMultiGPUSynchronize(m) == CUDA_ERROR_NOT_READY);
  EXPECT(cuMultiGPUSynchronize(m) == CUDA_SUCCESS);
  CHECK(cuMultiGPUDestroy(m));

  /* ---- every visible device: host-pointer POTRF through the multi-GPU LAPACK handle ---- */
  CHECK(cuMultiGPUCreate(&m, devs, ndev));
  EXPECT(cuMultiGPUGetContextCount(m) == ndev);
  CUmultiGPULAPACKhandle h;
  CHECK(cuMultiGPULAPACKCreate(&h, m));
  const size_t n = 777, lda = 780;
  double *A = malloc(lda * n * sizeof *A), *A0 = malloc(lda * n * sizeof *A0);
  srand(0);
  for (size_t j = 0; j < n; j++)
    for (size_t i = j; i < n; i++) {
      const double v = (i == j) ? (double)n + 1.0 : (double)rand() / RAND_MAX - 0.5;
      A[j * lda + i] = A[i * lda + j] = v;
    }
  for (size_t i = 0; i < lda * n; i++) A0[i] = A[i];
  long info = -99;
  CHECK(cuMultiGPUDpotrf(h, CBlasLower, n, A, lda, &info));
  EXPECT(info == 0);
  double worst = 0.0;
  for (size_t j = 0; j < n; j += 7)       /* sampled entries of L L^T - A */
    for (size_t i = j; i < n; i += 5) {
      double s = 0.0;
      for (size_t k = 0; k <= j; k++) s += A[k * lda + i] * A[k * lda + j];
      const double d = fabs(s - A0[j * lda + i]);
      if (d > worst) worst = d;
    }
  printf("max sampled |L L^T - A| = %.3e\n", worst);
  EXPECT(worst <= 1e-10);
  for (size_t j = 1; j < n; j++)           /* the strictly upper triangle is untouched */
    for (size_t i = 0; i < j; i++) EXPECT(A[j * lda + i] == A0[j * lda + i]);
  CHECK(cuMultiGPULAPACKSynchronize(h));
  CHECK(cuMultiGPULAPACKDestroy(h));
  CHECK(cuMultiGPUDestroy(m));
  free(A);
  free(A0);
  printf("PASSED\n");
  return 0;
}
