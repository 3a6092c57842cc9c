/*
 * mock_cuda.c — TEST INFRASTRUCTURE ONLY.  A stand-in for the few CUDA driver entry points the host
 * code (cuda-cholesky_b200/host) calls, so that the C drivers - blocked POTRF / TRTRI / LAUUM /
 * POTRI / POTRS, TRSM / TRMM, the CUmultiGPU runtime, the distributed factorisation with its
 * look-ahead, the slab-distributed inverse - can be exercised on a machine without a GPU
 * (tests/test_host_drivers_mock.py).  "Device" memory is host memory.
 *
 * Streams are REAL asynchronous queues: every CUstream owns a worker thread that runs the
 * operations queued on it in order (copies, memsets, the mock kernels of mock_kernels.cpp, the
 * mock NCCL broadcast), cuEventRecord / cuStreamWaitEvent carry the CUDA semantics (a wait refers
 * to the most recent record at the time of the call; waiting on a never-recorded event is a
 * no-op), the NULL stream synchronises with everything.  So a missing dependency between the
 * panel, update, upload and download streams is a real race here too; MOCK_CUDA_JITTER=1 makes
 * every stream stall for a random few hundred microseconds before each operation to shake such
 * races out, MOCK_CUDA_SLOW_H2D_US=n makes every host-to-device copy take n microseconds.
 * MOCK_CUDA_SYNC=1 falls back to running every operation at once.
 * MOCK_CUDA_DEVICES (default 1) sets the number of devices.  Linked only into the mock test library.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <cuda.h>

#define MAXDEV 16
struct CUctx_st { int device; };
static struct CUctx_st ctxs[MAXDEV] = {{0}, {1}, {2}, {3}, {4}, {5}, {6}, {7}, {8}, {9}, {10}, {11}, {12}, {13}, {14}, {15}};
static __thread CUcontext stack[16];
static __thread int top = 0; /* stack[top-1] is current */

static int ndev(void) {
  const char *e = getenv("MOCK_CUDA_DEVICES");
  const int n = e ? atoi(e) : 1;
  return n < 1 ? 1 : (n > MAXDEV ? MAXDEV : n);
}

CUresult cuInit(unsigned int flags) { (void)flags; return CUDA_SUCCESS; } /* callable from every worker thread at once, like the real one */
CUresult cuGetErrorString(CUresult e, const char **s) { (void)e; *s = "mock CUDA error"; return CUDA_SUCCESS; }
CUresult cuDeviceGet(CUdevice *d, int ordinal) { if (ordinal < 0 || ordinal >= ndev()) return CUDA_ERROR_INVALID_DEVICE; *d = ordinal; return CUDA_SUCCESS; }
CUresult cuDeviceGetCount(int *n) { *n = ndev(); return CUDA_SUCCESS; }
CUresult cuDevicePrimaryCtxRetain(CUcontext *c, CUdevice d) { if (d < 0 || d >= ndev()) return CUDA_ERROR_INVALID_DEVICE; *c = &ctxs[d]; return CUDA_SUCCESS; }
CUresult cuDevicePrimaryCtxRelease(CUdevice d) { (void)d; return CUDA_SUCCESS; }
CUresult cuCtxGetCurrent(CUcontext *c) { *c = top ? stack[top - 1] : NULL; return CUDA_SUCCESS; }
CUresult cuCtxSetCurrent(CUcontext c) { if (top == 0) top = 1; stack[top - 1] = c; return CUDA_SUCCESS; }
CUresult cuCtxPushCurrent(CUcontext c) { if (top >= 16) return CUDA_ERROR_INVALID_VALUE; stack[top++] = c; return CUDA_SUCCESS; }
CUresult cuCtxPopCurrent(CUcontext *c) { if (top == 0) return CUDA_ERROR_INVALID_CONTEXT; *c = stack[--top]; return CUDA_SUCCESS; }
CUresult cuCtxGetDevice(CUdevice *d) { if (top == 0 || stack[top - 1] == NULL) return CUDA_ERROR_INVALID_CONTEXT; *d = stack[top - 1]->device; return CUDA_SUCCESS; }
CUresult cuCtxGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -5; return CUDA_SUCCESS; }

/* the test harness asks how deep the calling thread's context stack is (push / pop balance) */
int mock_cuda_ctx_depth(void) { return top; }

/* ---- streams: a worker thread and a FIFO of operations each ---- */
struct op { void (*fn)(void *); void *arg; struct op *next; };
struct CUstream_st {
  pthread_t thread;
  pthread_mutex_t mu;
  pthread_cond_t cv;      /* work available / queue drained */
  struct op *head, *tail;
  int busy, stop;
  unsigned rng;
  struct CUstream_st *next_stream; /* registry, for device-wide synchronisation */
};
struct CUevent_st { pthread_mutex_t mu; pthread_cond_t cv; unsigned long long recorded, completed; };
static pthread_mutex_t streams_lock = PTHREAD_MUTEX_INITIALIZER;
static struct CUstream_st *all_streams = NULL;

static int sync_mode(void) { const char *e = getenv("MOCK_CUDA_SYNC"); return e && atoi(e); }
static int jitter(void) { const char *e = getenv("MOCK_CUDA_JITTER"); return e && atoi(e); }

static void *stream_main(void *p) {
  struct CUstream_st *s = p;
  for (;;) {
    pthread_mutex_lock(&s->mu);
    while (s->head == NULL && !s->stop) pthread_cond_wait(&s->cv, &s->mu);
    struct op *o = s->head;
    if (o == NULL) { pthread_mutex_unlock(&s->mu); break; }
    s->head = o->next;
    if (s->head == NULL) s->tail = NULL;
    s->busy = 1;
    pthread_mutex_unlock(&s->mu);
    if (jitter()) {
      s->rng = s->rng * 1664525u + 1013904223u;
      if ((s->rng >> 16) % 3 == 0) { struct timespec ts = {0, (long)((s->rng >> 8) % 400) * 1000}; nanosleep(&ts, NULL); }
    }
    o->fn(o->arg);
    free(o);
    pthread_mutex_lock(&s->mu);
    s->busy = 0;
    pthread_cond_broadcast(&s->cv);
    pthread_mutex_unlock(&s->mu);
  }
  return NULL;
}

static void drain(struct CUstream_st *s) {
  pthread_mutex_lock(&s->mu);
  while (s->head != NULL || s->busy) pthread_cond_wait(&s->cv, &s->mu);
  pthread_mutex_unlock(&s->mu);
}
static void drain_all(void) {
  pthread_mutex_lock(&streams_lock);
  /* the registry only grows or shrinks under this lock; draining while holding it is fine because
   * stream workers never take it */
  for (struct CUstream_st *s = all_streams; s; s = s->next_stream) drain(s);
  pthread_mutex_unlock(&streams_lock);
}

/* queue fn(arg) on the stream (NULL stream, or MOCK_CUDA_SYNC: run now, after everything queued anywhere).
 * Exported for mock_kernels.cpp and mock_nccl.c. */
void mock_stream_enqueue(CUstream s, void (*fn)(void *), void *arg) {
  if (s == NULL || sync_mode()) {
    if (s == NULL) drain_all(); else drain(s);
    fn(arg);
    return;
  }
  struct op *o = malloc(sizeof *o);
  o->fn = fn; o->arg = arg; o->next = NULL;
  pthread_mutex_lock(&s->mu);
  if (s->tail) s->tail->next = o; else s->head = o;
  s->tail = o;
  pthread_cond_broadcast(&s->cv);
  pthread_mutex_unlock(&s->mu);
}

CUresult cuStreamCreate(CUstream *out, unsigned int f) {
  (void)f;
  struct CUstream_st *s = calloc(1, sizeof *s);
  if (s == NULL) return CUDA_ERROR_OUT_OF_MEMORY;
  pthread_mutex_init(&s->mu, NULL);
  pthread_cond_init(&s->cv, NULL);
  s->rng = (unsigned)(uintptr_t)s;
  if (pthread_create(&s->thread, NULL, stream_main, s) != 0) { free(s); return CUDA_ERROR_OPERATING_SYSTEM; }
  pthread_mutex_lock(&streams_lock);
  s->next_stream = all_streams;
  all_streams = s;
  pthread_mutex_unlock(&streams_lock);
  *out = s;
  return CUDA_SUCCESS;
}
CUresult cuStreamCreateWithPriority(CUstream *s, unsigned int f, int p) { (void)p; return cuStreamCreate(s, f); }
CUresult cuStreamDestroy(CUstream s) {
  drain(s);
  pthread_mutex_lock(&streams_lock);
  for (struct CUstream_st **pp = &all_streams; *pp; pp = &(*pp)->next_stream)
    if (*pp == s) { *pp = s->next_stream; break; }
  pthread_mutex_unlock(&streams_lock);
  pthread_mutex_lock(&s->mu);
  s->stop = 1;
  pthread_cond_broadcast(&s->cv);
  pthread_mutex_unlock(&s->mu);
  pthread_join(s->thread, NULL);
  pthread_mutex_destroy(&s->mu);
  pthread_cond_destroy(&s->cv);
  free(s);
  return CUDA_SUCCESS;
}
CUresult cuStreamSynchronize(CUstream s) { if (s == NULL) drain_all(); else drain(s); return CUDA_SUCCESS; }
CUresult cuCtxSynchronize(void) { drain_all(); return CUDA_SUCCESS; }

CUresult cuEventCreate(CUevent *out, unsigned int f) {
  (void)f;
  struct CUevent_st *e = calloc(1, sizeof *e);
  if (e == NULL) return CUDA_ERROR_OUT_OF_MEMORY;
  pthread_mutex_init(&e->mu, NULL);
  pthread_cond_init(&e->cv, NULL);
  *out = e;
  return CUDA_SUCCESS;
}
CUresult cuEventDestroy(CUevent e) { pthread_mutex_destroy(&e->mu); pthread_cond_destroy(&e->cv); free(e); return CUDA_SUCCESS; }
struct evop { struct CUevent_st *e; unsigned long long seq; };
static void op_record(void *p) {
  struct evop *o = p;
  pthread_mutex_lock(&o->e->mu);
  if (o->seq > o->e->completed) o->e->completed = o->seq;
  pthread_cond_broadcast(&o->e->cv);
  pthread_mutex_unlock(&o->e->mu);
  free(o);
}
static void op_wait(void *p) {
  struct evop *o = p;
  pthread_mutex_lock(&o->e->mu);
  while (o->e->completed < o->seq) pthread_cond_wait(&o->e->cv, &o->e->mu);
  pthread_mutex_unlock(&o->e->mu);
  free(o);
}
CUresult cuEventRecord(CUevent e, CUstream s) {
  struct evop *o = malloc(sizeof *o);
  pthread_mutex_lock(&e->mu);
  o->e = e;
  o->seq = ++e->recorded;
  pthread_mutex_unlock(&e->mu);
  mock_stream_enqueue(s, op_record, o);
  return CUDA_SUCCESS;
}
CUresult cuStreamWaitEvent(CUstream s, CUevent e, unsigned int f) {
  (void)f;
  pthread_mutex_lock(&e->mu);
  const unsigned long long target = e->recorded; /* the most recent record at the time of this call */
  pthread_mutex_unlock(&e->mu);
  if (target == 0) return CUDA_SUCCESS; /* never recorded: no-op, as in CUDA */
  struct evop *o = malloc(sizeof *o);
  o->e = e;
  o->seq = target;
  mock_stream_enqueue(s, op_wait, o);
  return CUDA_SUCCESS;
}
CUresult cuEventSynchronize(CUevent e) {
  pthread_mutex_lock(&e->mu);
  const unsigned long long target = e->recorded;
  while (e->completed < target) pthread_cond_wait(&e->cv, &e->mu);
  pthread_mutex_unlock(&e->mu);
  return CUDA_SUCCESS;
}
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
CUresult cuMemFree(CUdeviceptr d) { drain_all(); /* cuMemFree synchronises the device */ free((void *)(uintptr_t)d); __sync_fetch_and_sub(&live_device_allocs, 1); return CUDA_SUCCESS; }
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
struct memop { int kind; void *dst; const void *src; size_t n; unsigned char v; CUDA_MEMCPY2D c2; };
static void copy2d_now(const CUDA_MEMCPY2D *c) {
  const char *src = c->srcMemoryType == CU_MEMORYTYPE_HOST ? (const char *)c->srcHost : (const char *)(uintptr_t)c->srcDevice;
  char *dst = c->dstMemoryType == CU_MEMORYTYPE_HOST ? (char *)c->dstHost : (char *)(uintptr_t)c->dstDevice;
  for (size_t r = 0; r < c->Height; r++)
    memcpy(dst + (c->dstY + r) * c->dstPitch + c->dstXInBytes, src + (c->srcY + r) * c->srcPitch + c->srcXInBytes, c->WidthInBytes);
}
static void op_mem(void *p) {
  struct memop *o = p;
  if (o->kind == 0) memset(o->dst, o->v, o->n);
  else if (o->kind == 1) memcpy(o->dst, o->src, o->n);
  else {
    /* MOCK_CUDA_SLOW_H2D_US: a host -> device copy takes that long, as if it crossed PCIe - anything that
     * touches the destination without waiting for the copy then works on stale data */
    const char *e = getenv("MOCK_CUDA_SLOW_H2D_US");
    if (e && o->c2.srcMemoryType == CU_MEMORYTYPE_HOST && o->c2.dstMemoryType == CU_MEMORYTYPE_DEVICE) {
      const long us = atol(e);
      struct timespec ts = {us / 1000000, (us % 1000000) * 1000};
      nanosleep(&ts, NULL);
    }
    copy2d_now(&o->c2);
  }
  free(o);
}
CUresult cuMemsetD8(CUdeviceptr d, unsigned char v, size_t n) { drain_all(); memset((void *)(uintptr_t)d, v, n); return CUDA_SUCCESS; }
CUresult cuMemsetD8Async(CUdeviceptr d, unsigned char v, size_t n, CUstream s) {
  struct memop *o = calloc(1, sizeof *o);
  o->kind = 0; o->dst = (void *)(uintptr_t)d; o->v = v; o->n = n;
  mock_stream_enqueue(s, op_mem, o);
  return CUDA_SUCCESS;
}
/* host -> device, asynchronous.  From page-locked memory the source is read when the copy EXECUTES (overwriting the
 * buffer before then corrupts the copy, as on the device); from pageable memory the driver stages the data before it
 * returns, so the source is snapshotted now. */
static void op_h2d_staged(void *p) {
  struct memop *o = p;
  memcpy(o->dst, o->src, o->n);
  free((void *)o->src);
  free(o);
}
CUresult cuMemcpyHtoDAsync(CUdeviceptr d, const void *h, size_t n, CUstream s) {
  struct memop *o = calloc(1, sizeof *o);
  o->dst = (void *)(uintptr_t)d; o->n = n;
  int is_pinned = 0;
  pthread_mutex_lock(&reg_lock);
  for (int i = 0; i < 256; i++)
    if (pinned[i].p && (const char *)h >= (const char *)pinned[i].p && (const char *)h < (const char *)pinned[i].p + pinned[i].n) is_pinned = 1;
  pthread_mutex_unlock(&reg_lock);
  if (is_pinned) {
    o->kind = 1; o->src = h;
    mock_stream_enqueue(s, op_mem, o);
  } else {
    void *copy = malloc(n ? n : 1);
    memcpy(copy, h, n);
    o->src = copy;
    mock_stream_enqueue(s, op_h2d_staged, o);
  }
  return CUDA_SUCCESS;
}
CUresult cuMemcpyDtoHAsync(void *h, CUdeviceptr d, size_t n, CUstream s) {
  struct memop *o = calloc(1, sizeof *o);
  o->kind = 1; o->dst = h; o->src = (const void *)(uintptr_t)d; o->n = n;
  mock_stream_enqueue(s, op_mem, o);
  return CUDA_SUCCESS;
}
CUresult cuMemcpy2DAsync(const CUDA_MEMCPY2D *c, CUstream s) {
  if ((c->WidthInBytes > c->srcPitch || c->WidthInBytes > c->dstPitch) && c->Height > 1)
    return CUDA_ERROR_INVALID_VALUE; /* the driver rejects a width larger than the pitch */
  struct memop *o = calloc(1, sizeof *o);
  o->kind = 2; o->c2 = *c;
  mock_stream_enqueue(s, op_mem, o);
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
CUresult cuMemcpyHtoD(CUdeviceptr d, const void *h, size_t n) { drain_all(); memcpy((void *)(uintptr_t)d, h, n); return CUDA_SUCCESS; }
CUresult cuMemcpyDtoH(void *h, CUdeviceptr d, size_t n) { drain_all(); memcpy(h, (const void *)(uintptr_t)d, n); return CUDA_SUCCESS; }
CUresult cuMemcpy2D(const CUDA_MEMCPY2D *c) { drain_all(); copy2d_now(c); return CUDA_SUCCESS; }
