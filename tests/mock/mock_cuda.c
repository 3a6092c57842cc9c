/*
 * mock_cuda.c — TEST INFRASTRUCTURE ONLY.  A stand-in for the few CUDA driver entry points the host
 * code (cuda-cholesky_b200/host) calls, so that the C drivers - blocked POTRF / TRTRI / LAUUM /
 * POTRI / POTRS, TRSM / TRMM, the CUmultiGPU runtime, the distributed factorisation with its
 * look-ahead, the slab-distributed inverse - can be exercised on a machine without a GPU
 * (tests/test_host_drivers_mock.py).  "Device" memory is host memory, every stream operation runs
 * at once, events do nothing: this checks algebra, indexing and control flow, NOT stream ordering.
 * MOCK_CUDA_DEVICES (default 1) sets the number of devices.  Linked only into the mock test library.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <cuda.h>

#define MAXDEV 16
struct CUctx_st { int device; };
struct CUstream_st { int dummy; };
struct CUevent_st { int dummy; };
static struct CUctx_st ctxs[MAXDEV];
static __thread CUcontext stack[16];
static __thread int top = 0; /* stack[top-1] is current */

static int ndev(void) {
  const char *e = getenv("MOCK_CUDA_DEVICES");
  const int n = e ? atoi(e) : 1;
  return n < 1 ? 1 : (n > MAXDEV ? MAXDEV : n);
}

CUresult cuInit(unsigned int flags) { (void)flags; for (int i = 0; i < MAXDEV; i++) ctxs[i].device = i; return CUDA_SUCCESS; }
CUresult cuGetErrorString(CUresult e, const char **s) { (void)e; *s = "mock CUDA error"; return CUDA_SUCCESS; }
CUresult cuDeviceGet(CUdevice *d, int ordinal) { if (ordinal < 0 || ordinal >= ndev()) return CUDA_ERROR_INVALID_DEVICE; *d = ordinal; return CUDA_SUCCESS; }
CUresult cuDeviceGetCount(int *n) { *n = ndev(); return CUDA_SUCCESS; }
CUresult cuDevicePrimaryCtxRetain(CUcontext *c, CUdevice d) { if (d < 0 || d >= ndev()) return CUDA_ERROR_INVALID_DEVICE; ctxs[d].device = d; *c = &ctxs[d]; return CUDA_SUCCESS; }
CUresult cuDevicePrimaryCtxRelease(CUdevice d) { (void)d; return CUDA_SUCCESS; }
CUresult cuCtxGetCurrent(CUcontext *c) { *c = top ? stack[top - 1] : NULL; return CUDA_SUCCESS; }
CUresult cuCtxSetCurrent(CUcontext c) { if (top == 0) top = 1; stack[top - 1] = c; return CUDA_SUCCESS; }
CUresult cuCtxPushCurrent(CUcontext c) { if (top >= 16) return CUDA_ERROR_INVALID_VALUE; stack[top++] = c; return CUDA_SUCCESS; }
CUresult cuCtxPopCurrent(CUcontext *c) { if (top == 0) return CUDA_ERROR_INVALID_CONTEXT; *c = stack[--top]; return CUDA_SUCCESS; }
CUresult cuCtxGetDevice(CUdevice *d) { if (top == 0 || stack[top - 1] == NULL) return CUDA_ERROR_INVALID_CONTEXT; *d = stack[top - 1]->device; return CUDA_SUCCESS; }
CUresult cuCtxSynchronize(void) { return CUDA_SUCCESS; }
CUresult cuCtxGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -5; return CUDA_SUCCESS; }

/* the test harness asks how deep the calling thread's context stack is (push / pop balance) */
int mock_cuda_ctx_depth(void) { return top; }

CUresult cuStreamCreate(CUstream *s, unsigned int f) { (void)f; *s = calloc(1, sizeof **s); return *s ? CUDA_SUCCESS : CUDA_ERROR_OUT_OF_MEMORY; }
CUresult cuStreamCreateWithPriority(CUstream *s, unsigned int f, int p) { (void)p; return cuStreamCreate(s, f); }
CUresult cuStreamDestroy(CUstream s) { free(s); return CUDA_SUCCESS; }
CUresult cuStreamSynchronize(CUstream s) { (void)s; return CUDA_SUCCESS; }
CUresult cuStreamWaitEvent(CUstream s, CUevent e, unsigned int f) { (void)s; (void)e; (void)f; return CUDA_SUCCESS; }
CUresult cuEventCreate(CUevent *e, unsigned int f) { (void)f; *e = calloc(1, sizeof **e); return *e ? CUDA_SUCCESS : CUDA_ERROR_OUT_OF_MEMORY; }
CUresult cuEventDestroy(CUevent e) { free(e); return CUDA_SUCCESS; }
CUresult cuEventRecord(CUevent e, CUstream s) { (void)e; (void)s; return CUDA_SUCCESS; }
CUresult cuEventElapsedTime(float *ms, CUevent a, CUevent b) { (void)a; (void)b; *ms = 1.0f; return CUDA_SUCCESS; }

/* ---- memory: "device" = host.  MOCK_CUDA_MEM_LIMIT (bytes) makes larger device allocations fail. ---- */
static pthread_mutex_t reg_lock = PTHREAD_MUTEX_INITIALIZER;
static struct { void *p; size_t n; } pinned[256];
static long live_device_allocs = 0;
long mock_cuda_live_allocs(void) { return live_device_allocs; }

CUresult cuMemAlloc(CUdeviceptr *d, size_t bytes) {
  const char *lim = getenv("MOCK_CUDA_MEM_LIMIT");
  if (lim && bytes > (size_t)atoll(lim)) return CUDA_ERROR_OUT_OF_MEMORY;
  void *p = NULL;
  if (posix_memalign(&p, 256, bytes ? bytes : 1) != 0) return CUDA_ERROR_OUT_OF_MEMORY;
  memset(p, 0xA5, bytes); /* uninitialised device memory is not zero */
  *d = (CUdeviceptr)(uintptr_t)p;
  __sync_fetch_and_add(&live_device_allocs, 1);
  return CUDA_SUCCESS;
}
CUresult cuMemFree(CUdeviceptr d) { free((void *)(uintptr_t)d); __sync_fetch_and_sub(&live_device_allocs, 1); return CUDA_SUCCESS; }
CUresult cuMemAllocHost(void **p, size_t bytes) {
  if (posix_memalign(p, 256, bytes ? bytes : 1) != 0) return CUDA_ERROR_OUT_OF_MEMORY;
  pthread_mutex_lock(&reg_lock);
  for (int i = 0; i < 256; i++)
    if (pinned[i].p == NULL) { pinned[i].p = *p; pinned[i].n = bytes; break; }
  pthread_mutex_unlock(&reg_lock);
  return CUDA_SUCCESS;
}
CUresult cuMemFreeHost(void *p) {
  pthread_mutex_lock(&reg_lock);
  for (int i = 0; i < 256; i++)
    if (pinned[i].p == p) pinned[i].p = NULL;
  pthread_mutex_unlock(&reg_lock);
  free(p);
  return CUDA_SUCCESS;
}
/* tests mark a caller-owned buffer as page-locked to steer the host-pointer API onto its overlapped paths */
void mock_cuda_register_pinned(void *p, size_t bytes) {
  pthread_mutex_lock(&reg_lock);
  for (int i = 0; i < 256; i++)
    if (pinned[i].p == NULL) { pinned[i].p = p; pinned[i].n = bytes; break; }
  pthread_mutex_unlock(&reg_lock);
}
void mock_cuda_unregister_pinned(void *p) {
  pthread_mutex_lock(&reg_lock);
  for (int i = 0; i < 256; i++)
    if (pinned[i].p == p) pinned[i].p = NULL;
  pthread_mutex_unlock(&reg_lock);
}
CUresult cuPointerGetAttribute(void *data, CUpointer_attribute attr, CUdeviceptr ptr) {
  if (attr != CU_POINTER_ATTRIBUTE_MEMORY_TYPE) return CUDA_ERROR_INVALID_VALUE;
  const char *q = (const char *)(uintptr_t)ptr;
  CUresult r = CUDA_ERROR_INVALID_VALUE; /* what the driver answers for pageable host memory */
  pthread_mutex_lock(&reg_lock);
  for (int i = 0; i < 256; i++)
    if (pinned[i].p && q >= (const char *)pinned[i].p && q < (const char *)pinned[i].p + pinned[i].n) {
      *(unsigned int *)data = CU_MEMORYTYPE_HOST;
      r = CUDA_SUCCESS;
    }
  pthread_mutex_unlock(&reg_lock);
  return r;
}
CUresult cuMemsetD8(CUdeviceptr d, unsigned char v, size_t n) { memset((void *)(uintptr_t)d, v, n); return CUDA_SUCCESS; }
CUresult cuMemsetD8Async(CUdeviceptr d, unsigned char v, size_t n, CUstream s) { (void)s; return cuMemsetD8(d, v, n); }
CUresult cuMemcpyDtoHAsync(void *h, CUdeviceptr d, size_t n, CUstream s) { (void)s; memcpy(h, (const void *)(uintptr_t)d, n); return CUDA_SUCCESS; }
CUresult cuMemcpy2DAsync(const CUDA_MEMCPY2D *c, CUstream s) {
  (void)s;
  const char *src = c->srcMemoryType == CU_MEMORYTYPE_HOST ? (const char *)c->srcHost : (const char *)(uintptr_t)c->srcDevice;
  char *dst = c->dstMemoryType == CU_MEMORYTYPE_HOST ? (char *)c->dstHost : (char *)(uintptr_t)c->dstDevice;
  if (c->WidthInBytes > c->srcPitch || c->WidthInBytes > c->dstPitch) {
    if (c->Height > 1) return CUDA_ERROR_INVALID_VALUE; /* the driver rejects a width larger than the pitch */
  }
  for (size_t r = 0; r < c->Height; r++)
    memcpy(dst + (c->dstY + r) * c->dstPitch + c->dstXInBytes, src + (c->srcY + r) * c->srcPitch + c->srcXInBytes, c->WidthInBytes);
  return CUDA_SUCCESS;
}

/* ---- used by the C test programs (tests/c), not by the library ---- */
CUresult cuDeviceGetName(char *name, int len, CUdevice d) { (void)d; strncpy(name, "mock device", (size_t)len); if (len > 0) name[len - 1] = 0; return CUDA_SUCCESS; }
CUresult cuDeviceTotalMem(size_t *bytes, CUdevice d) { (void)d; *bytes = (size_t)180 << 30; return CUDA_SUCCESS; }
CUresult cuCtxCreate(CUcontext *c, unsigned int flags, CUdevice d) {
  (void)flags;
  if (d < 0 || d >= ndev()) return CUDA_ERROR_INVALID_DEVICE;
  ctxs[d].device = d;
  *c = &ctxs[d];
  return cuCtxPushCurrent(*c);
}
CUresult cuCtxDestroy(CUcontext c) { (void)c; if (top > 0) --top; return CUDA_SUCCESS; }
CUresult cuMemcpyHtoD(CUdeviceptr d, const void *h, size_t n) { memcpy((void *)(uintptr_t)d, h, n); return CUDA_SUCCESS; }
CUresult cuMemcpyDtoH(void *h, CUdeviceptr d, size_t n) { memcpy(h, (const void *)(uintptr_t)d, n); return CUDA_SUCCESS; }
CUresult cuMemcpy2D(const CUDA_MEMCPY2D *c) { return cuMemcpy2DAsync(c, NULL); }
