/*
 * multigpu.c — the multi-GPU runtime and the distributed Cholesky.
 *
 * Replaces reference multigpu/multigpu.c (pthread-per-GPU ring-buffer scheduler, :13-538) and the
 * cuMultiGPU* drivers built on it (lapack/dpotrf.c:298-366, blas/dgemm.c:211-566), which deal C
 * tiles round-robin to the GPUs and stage EVERY tile over PCIe in both directions each step.
 *
 * Here:
 *   - the API keeps the reference's signatures: CUtask, CUmultiGPU (include/cumultigpu.h) and
 *     cuMultiGPU{S,D,C,Z}potrf (include/lapack.h) with the matrix in HOST memory;
 *   - the matrix is scattered ONCE into a 1-D block-cyclic layout (block column c lives on GPU
 *     c mod P, packed side by side, resident in HBM), factored there, and gathered once;
 *   - per step the owner factors the diagonal block and solves the panel into a contiguous buffer
 *     W; W is broadcast GPU-to-GPU over NVLink with ncclBroadcast on the high-priority panel
 *     stream; every GPU then updates all the block columns it owns with ONE block-cyclic-aware
 *     launch of the DMMA tile engine (csrc/gemm_common.cuh ColMap);
 *   - depth-1 look-ahead: the owner of the next block column updates that column first, factors
 *     it and starts its broadcast while all GPUs are still busy with the current trailing update;
 *   - the same SPMD core serves (a) one worker thread per GPU inside one process (cuMultiGPU*, the
 *     reference's model) and (b) one process per GPU (cholDist*, used under torchrun: the NCCL
 *     unique id is exchanged by the launcher).
 * Upper-triangular inputs are transposed on the way in and out (U^H is a lower factor), so the
 * distributed core only ever sees the lower form.
 */
#define _GNU_SOURCE
#include <complex.h>
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "chol_internal.h"

static const double ONE[2] = {1.0, 0.0};
static const double MONE[2] = {-1.0, 0.0};
static const double ZERO2[2] = {0.0, 0.0};

/* =============================================================================================
 * NCCL, bound at run time (dlopen by soname: inside a PyTorch process this resolves to the copy
 * torch already loaded).  Only what the panel broadcast needs.
 * ===========================================================================================*/
typedef struct { char internal[128]; } chol_nccl_id;
typedef void *chol_nccl_comm;
static struct {
  int loaded;
  int (*GetUniqueId)(chol_nccl_id *);
  int (*CommInitRank)(chol_nccl_comm *, int, chol_nccl_id, int);
  int (*CommInitRankConfig)(chol_nccl_comm *, int, chol_nccl_id, int, void *);
  int (*CommInitAll)(chol_nccl_comm *, int, const int *);
  int (*CommDestroy)(chol_nccl_comm);
  int (*Broadcast)(const void *, void *, size_t, int, int, chol_nccl_comm, CUstream);
  int (*GroupStart)(void);
  int (*GroupEnd)(void);
  const char *(*GetErrorString)(int);
} nccl;
static pthread_mutex_t nccl_lock = PTHREAD_MUTEX_INITIALIZER;
#define CHOL_NCCL_UINT8 1

/* The panel broadcast shares the GPU with the trailing update: a handful of NCCL CTAs is enough
 * for <= 134 MB per step over NVSwitch, and every SM NCCL does not hold keeps running DMMA tiles
 * (measured on 8 B200: 8 CTAs -> +3 % whole-job throughput over NCCL's default).  CHOL_NCCL_MAX_CTAS
 * overrides.  ncclConfig_t layout of NCCL 2.27 (nccl.h:70-88); newer libraries accept older versions. */
typedef struct {
  size_t size;
  unsigned int magic, version;
  int blocking, cgaClusterSize, minCTAs, maxCTAs;
  const char *netName;
  int splitShare, trafficClass;
  const char *commName;
  int collnetEnable, CTAPolicy, shrinkShare, nvlsCTAs;
} chol_nccl_config;
#define CHOL_NCCL_UNDEF_INT (-2147483647 - 1)
static int nccl_max_ctas(void) {
  const char *e = getenv("CHOL_NCCL_MAX_CTAS");
  const int v = e ? atoi(e) : 8;
  return v > 0 ? v : 8;
}

static const char *nccl_strerror(int e) { return nccl.GetErrorString ? nccl.GetErrorString(e) : "NCCL error"; }

static CUresult nccl_load(void) {
  pthread_mutex_lock(&nccl_lock);
  if (!nccl.loaded) {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == NULL) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h != NULL) {
      nccl.GetUniqueId = (int (*)(chol_nccl_id *))dlsym(h, "ncclGetUniqueId");
      nccl.CommInitRank = (int (*)(chol_nccl_comm *, int, chol_nccl_id, int))dlsym(h, "ncclCommInitRank");
      nccl.CommInitRankConfig = (int (*)(chol_nccl_comm *, int, chol_nccl_id, int, void *))dlsym(h, "ncclCommInitRankConfig");
      nccl.CommInitAll = (int (*)(chol_nccl_comm *, int, const int *))dlsym(h, "ncclCommInitAll");
      nccl.CommDestroy = (int (*)(chol_nccl_comm))dlsym(h, "ncclCommDestroy");
      nccl.Broadcast = (int (*)(const void *, void *, size_t, int, int, chol_nccl_comm, CUstream))dlsym(h, "ncclBroadcast");
      nccl.GroupStart = (int (*)(void))dlsym(h, "ncclGroupStart");
      nccl.GroupEnd = (int (*)(void))dlsym(h, "ncclGroupEnd");
      nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
      nccl.loaded = nccl.GetUniqueId && nccl.CommInitRank && nccl.CommInitAll && nccl.CommDestroy && nccl.Broadcast &&
                    nccl.GroupStart && nccl.GroupEnd;
    }
  }
  pthread_mutex_unlock(&nccl_lock);
  if (!nccl.loaded) {
    if (errorHandler != NULL)
      errorHandler("dlopen(libnccl.so.2)", __func__, __FILE__, __LINE__, -1, (const char *(*)(int))nccl_strerror);
    return CUDA_ERROR_NOT_SUPPORTED;
  }
  return CUDA_SUCCESS;
}
#define NCCL_CHECK(call)                                                                                       \
  do {                                                                                                         \
    int __n__ = (call);                                                                                        \
    if (__n__ != 0) {                                                                                          \
      if (errorHandler != NULL) errorHandler(STRING(call), __func__, __FILE__, __LINE__, __n__, nccl_strerror); \
      return CUDA_ERROR_UNKNOWN;                                                                               \
    }                                                                                                          \
  } while (false)

/* =============================================================================================
 * CUtask  (reference multigpu/multigpu.c:297-400)
 * ===========================================================================================*/
struct __cutask_st {
  CUresult (*function)(const void *);
  void *args;
  CUresult result;
  bool complete;
  pthread_mutex_t mutex;
  pthread_cond_t cond;
  struct __cutask_st *next; /* queue link */
};

CUresult cuTaskCreate(CUtask *task, CUresult (*function)(const void *), const void *args, size_t size) {
  if (task == NULL || function == NULL || (args == NULL && size > 0)) return CUDA_ERROR_INVALID_VALUE;
  struct __cutask_st *t = malloc(sizeof *t);
  if (t == NULL) return CUDA_ERROR_OUT_OF_MEMORY;
  t->function = function;
  t->args = NULL;
  if (size > 0) {
    if ((t->args = malloc(size)) == NULL) {
      free(t);
      return CUDA_ERROR_OUT_OF_MEMORY;
    }
    memcpy(t->args, args, size);
  }
  t->result = CUDA_SUCCESS;
  t->complete = false;
  t->next = NULL;
  pthread_mutex_init(&t->mutex, NULL);
  pthread_cond_init(&t->cond, NULL);
  *task = t;
  return CUDA_SUCCESS;
}

/* Once `complete` is visible the owner may already be inside cuTaskDestroy: the task is not
 * touched after the unlock. */
CUresult cuTaskExecute(CUtask t) {
  if (t == NULL) return CUDA_ERROR_INVALID_VALUE;
  const CUresult r = t->function(t->args);
  ERROR_CHECK(pthread_mutex_lock(&t->mutex));
  t->result = r;
  t->complete = true;
  pthread_cond_broadcast(&t->cond);
  ERROR_CHECK(pthread_mutex_unlock(&t->mutex));
  return CUDA_SUCCESS;
}

CUresult cuTaskDestroy(CUtask t, CUresult *result) {
  if (t == NULL) return CUDA_ERROR_INVALID_VALUE;
  ERROR_CHECK(pthread_mutex_lock(&t->mutex));
  while (!t->complete) pthread_cond_wait(&t->cond, &t->mutex);
  ERROR_CHECK(pthread_mutex_unlock(&t->mutex));
  if (result != NULL) *result = t->result;
  pthread_mutex_destroy(&t->mutex);
  pthread_cond_destroy(&t->cond);
  free(t->args);
  free(t);
  return CUDA_SUCCESS;
}

/* =============================================================================================
 * CUmultiGPU: one worker thread per device, FIFO of tasks  (reference multigpu.c:128-292, :405-538)
 * ===========================================================================================*/
struct chol_rank; /* per-GPU state of the distributed solver, below */

struct worker {
  pthread_t thread;
  pthread_mutex_t mutex;
  pthread_cond_t cond;
  struct __cutask_st *head, *tail;
  bool stop;
  CUdevice device;
  CUcontext context;
  CUresult error; /* first failing task result is latched (reference multigpu.c:139-147) */
};

struct __cumultigpu_st {
  int n;
  struct worker *w;
  struct chol_rank *ranks; /* created on first multi-GPU LAPACK call */
  chol_nccl_comm *comms;
};

static void *worker_main(void *arg) {
  struct worker *w = arg;
  cuCtxSetCurrent(w->context);
  for (;;) {
    pthread_mutex_lock(&w->mutex);
    while (w->head == NULL && !w->stop) pthread_cond_wait(&w->cond, &w->mutex);
    struct __cutask_st *t = w->head;
    if (t != NULL) {
      w->head = t->next;
      if (w->head == NULL) w->tail = NULL;
    }
    const bool stop = (t == NULL) && w->stop;
    pthread_mutex_unlock(&w->mutex);
    if (stop) break;
    /* latch the first failure BEFORE completion is published: whoever waits for a later task on
     * this worker (cuMultiGPUSynchronize) then reads it through that task's mutex */
    struct __cutask_st *const task = t;
    const CUresult r = task->function(task->args);
    if (r != CUDA_SUCCESS && w->error == CUDA_SUCCESS) w->error = r;
    pthread_mutex_lock(&task->mutex);
    task->result = r;
    task->complete = true;
    pthread_cond_broadcast(&task->cond);
    pthread_mutex_unlock(&task->mutex); /* `task` may be freed by its owner from here on */
  }
  return NULL;
}

int cuMultiGPUGetContextCount(CUmultiGPU m) { return m == NULL ? 0 : m->n; }

CUresult cuMultiGPUCreate(CUmultiGPU *mGPU, CUdevice *devices, int n) {
  if (mGPU == NULL || devices == NULL || n <= 0) return CUDA_ERROR_INVALID_VALUE;
  CU_ERROR_CHECK(cuInit(0));
  struct __cumultigpu_st *m = calloc(1, sizeof *m);
  if (m == NULL) return CUDA_ERROR_OUT_OF_MEMORY;
  m->w = calloc((size_t)n, sizeof *m->w);
  if (m->w == NULL) {
    free(m);
    return CUDA_ERROR_OUT_OF_MEMORY;
  }
  for (int i = 0; i < n; i++) {
    struct worker *w = &m->w[i];
    w->device = devices[i];
    CUresult e = cuDevicePrimaryCtxRetain(&w->context, w->device);
    if (e == CUDA_SUCCESS) {
      pthread_mutex_init(&w->mutex, NULL);
      pthread_cond_init(&w->cond, NULL);
      if (pthread_create(&w->thread, NULL, worker_main, w) != 0) {
        e = CUDA_ERROR_OPERATING_SYSTEM;
        pthread_mutex_destroy(&w->mutex);
        pthread_cond_destroy(&w->cond);
        cuDevicePrimaryCtxRelease(w->device);
      }
    }
    if (e != CUDA_SUCCESS) {
      m->n = i; /* workers 0..i-1 are running: tear those down */
      cuMultiGPUDestroy(m);
      CU_ERROR_CHECK(e);
    }
  }
  m->n = n;
  *mGPU = m;
  return CUDA_SUCCESS;
}

CUresult cuMultiGPURunTask(CUmultiGPU m, int i, CUtask task) {
  if (m == NULL || task == NULL || i < 0 || i >= m->n) return CUDA_ERROR_INVALID_VALUE;
  struct worker *w = &m->w[i];
  ERROR_CHECK(pthread_mutex_lock(&w->mutex));
  task->next = NULL;
  if (w->tail) w->tail->next = task; else w->head = task;
  w->tail = task;
  pthread_cond_signal(&w->cond);
  ERROR_CHECK(pthread_mutex_unlock(&w->mutex));
  return CUDA_SUCCESS;
}

static CUresult sync_task(const void *args) {
  (void)args;
  return cuCtxSynchronize();
}

/* run `fn(args_i)` on every worker and wait; returns the first failure */
static CUresult run_on_all(CUmultiGPU m, CUresult (*fn)(const void *), const void *args, size_t stride, size_t size) {
  CUtask *tasks = calloc((size_t)m->n, sizeof *tasks);
  if (tasks == NULL) return CUDA_ERROR_OUT_OF_MEMORY;
  CUresult first = CUDA_SUCCESS;
  int launched = 0;
  for (int i = 0; i < m->n && first == CUDA_SUCCESS; i++) {
    first = cuTaskCreate(&tasks[i], fn, args ? (const char *)args + (size_t)i * stride : NULL, size);
    if (first == CUDA_SUCCESS) first = cuMultiGPURunTask(m, i, tasks[i]);
    if (first == CUDA_SUCCESS) launched++;
  }
  for (int i = 0; i < launched; i++) {
    CUresult r = CUDA_SUCCESS;
    cuTaskDestroy(tasks[i], &r);
    if (first == CUDA_SUCCESS && r != CUDA_SUCCESS) first = r;
  }
  free(tasks);
  return first;
}

CUresult cuMultiGPUSynchronize(CUmultiGPU m) {
  if (m == NULL) return CUDA_ERROR_INVALID_VALUE;
  CUresult r = run_on_all(m, sync_task, NULL, 0, 0);
  for (int i = 0; i < m->n && r == CUDA_SUCCESS; i++) {
    r = m->w[i].error;
    m->w[i].error = CUDA_SUCCESS;
  }
  return r;
}

/* =============================================================================================
 * The SPMD distributed POTRF (lower form, 1-D block-cyclic block columns)
 * ===========================================================================================*/
typedef struct chol_rank {
  int rank, nranks;
  CUcontext context;
  chol_nccl_comm comm; /* NULL when nranks == 1 */
  CULAPACKhandle lap;
  CUdeviceptr local;   /* single-process mode: the rank's packed block columns */
  size_t local_bytes;
} chol_rank;

/* number of columns rank r owns of an n-column matrix with block size nb over P ranks */
static size_t local_cols(size_t n, size_t nb, int r, int P) {
  const size_t nblk = (n + nb - 1) / nb;
  size_t cols = 0;
  for (size_t c = (size_t)r; c < nblk; c += (size_t)P) cols += chol_min(nb, n - c * nb);
  return cols;
}

/* D (m x n packed local columns, global triangle column of local column c: gcol(c) with block-cyclic parameters
 * cnb, cP, coff) += alpha * W[0:m, :] * W[gcol(c), :]^H on the global lower triangle (m >= gcol(c)) */
static CUresult gemm_cyc_ex(int prec, int herm, double alpha, size_t m, size_t n, size_t k, CUdeviceptr W, size_t ldw, CUdeviceptr D,
                            size_t ldd, int cnb, int cP, int coff, CUstream st) {
  chol_gemm_desc d;
  memset(&d, 0, sizeof d);
  d.prec = prec;
  d.opA = 'N';
  d.opB = 'C';
  d.cuplo = 'L';
  d.herm = herm;
  d.m = m; d.n = n; d.k = k;
  d.alpha[0] = alpha;
  d.beta[0] = 1.0;
  d.A = (const void *)W; d.lda = ldw;
  d.B = (const void *)W; d.ldb = ldw;
  d.C = (const void *)D; d.ldc = ldd;
  d.D = (void *)D; d.ldd = ldd;
  d.stream = (void *)st;
  d.cyc_nb = cnb; d.cyc_p = cP; d.cyc_off = coff;
  RT_CHECK(chol_gemm_launch(&d));
  return CUDA_SUCCESS;
}

static CUresult gemm_cyc(int prec, int herm, size_t m, size_t n, size_t k, CUdeviceptr W, size_t ldw, CUdeviceptr D,
                         size_t ldd, int cnb, int cP, CUstream st) {
  chol_gemm_desc d;
  memset(&d, 0, sizeof d);
  d.prec = prec;
  d.opA = 'N';
  d.opB = 'C';
  d.cuplo = 'L';
  d.herm = herm;
  d.m = m; d.n = n; d.k = k;
  d.alpha[0] = -1.0;
  d.beta[0] = 1.0;
  d.A = (const void *)W; d.lda = ldw;
  d.B = (const void *)W; d.ldb = ldw;
  d.C = (const void *)D; d.ldc = ldd;
  d.D = (void *)D; d.ldd = ldd;
  d.stream = (void *)st;
  d.cyc_nb = cnb; d.cyc_p = cP; d.cyc_off = 0;
  RT_CHECK(chol_gemm_launch(&d));
  return CUDA_SUCCESS;
}

/* A_loc: this rank's block columns (global block c = lc*P + rank at local columns lc*nb...),
 * n rows, leading dimension ldl.  Factor in place (lower).  *info = local failure (global column,
 * 1-based) or 0. */
/* block size of the distributed layout: with P ranks the per-rank update shrinks by P while the
 * serial panel chain does not, so fewer columns per rank want a smaller block (measured at
 * N = 65536 on B200: P = 2, 4 -> 512 beats 256 by 5-7 %; P = 8 -> 256 beats 512 by 5 %) */
static size_t dist_nb(const chol_rank *R, int prec, size_t n) {
  if (R->lap->nb) return R->lap->nb;
  const size_t nb = chol_default_nb(prec, n);
  if (R->nranks <= 1) return nb;
  const size_t cap = (n / (size_t)R->nranks >= 16384) ? 512 : 256;
  return nb > cap ? cap : nb;
}

static CUresult dist_carve(chol_rank *R, int prec, size_t n, chol_ws *w) {
  return chol_carve_nb(R->lap, prec, n, dist_nb(R, prec, n), w);
}

/* optional: stream every finished block column of a rank straight back to the caller's host matrix
 * while the factorisation continues (lower form only) */
typedef struct {
  char *hostA;
  size_t lda;
  CUstream s_copy;
  CUevent ev[8];
} dist_gather;

/* optional: the rank's block columns are still arriving from the host while the first steps run
 * (upload stream, in local block-column order; one event per group of gblk local block columns).
 * The first split_steps trailing updates are launched group by group, each behind its group's
 * event; from step split_steps on everything has arrived and the update is one launch again. */
typedef struct {
  CUevent *ev;
  size_t ngroups, gblk, split_steps;
} dist_stage;

static CUresult copy2d(int h2d, void *host, size_t ldh, CUdeviceptr dev, size_t ldd, size_t rows, size_t cols, size_t es,
                       CUstream st);

static CUresult dist_potrf(chol_rank *R, int prec, size_t n, CUdeviceptr Aloc, size_t ldl, long *info, dist_gather *ga,
                           const dist_stage *sg) {
  CULAPACKhandle h = R->lap;
  const int P = R->nranks, me = R->rank;
  chol_ws w;
  CU_ERROR_CHECK(dist_carve(R, prec, n, &w));
  const size_t es = w.es, nb = w.nb, nblk = (n + nb - 1) / nb;
  const int herm = chol_is_complex(prec);
  long *dinfo = (long *)h->dinfo;
  CUstream sp = h->s_panel, su = h->s_update, sl = h->s_look;
  const bool split = chol_lookahead_split();
  CUevent look_pending = NULL; /* lower part of the block column this rank factors next, still being updated on sl */
  const size_t my_cols = local_cols(n, nb, me, P);
  CU_ERROR_CHECK(cuMemsetD8Async(h->dinfo, 0, 16, sp)); /* info and this rank's share of the fused log-determinant */
  h->last_logdet_valid = false;
  h->prof_n = 0;

  for (size_t c = 0; c < nblk; c++) {
    const size_t j = c * nb, jb = chol_min(nb, n - j), rest = n - j - jb;
    const int owner = (int)(c % (size_t)P);
    const CUdeviceptr W = w.W[c & 1];
    const size_t ldw = chol_roundup(rest ? rest : 1, 16);
    /* W's buffer was last read by the trailing update of step c-2 */
    if (c >= 2) CU_ERROR_CHECK(cuStreamWaitEvent(sp, h->ev_rest[(c - 2) % CHOL_EVENT_RING], 0));
    const bool staged = sg != NULL && c < sg->split_steps;
    if (sg != NULL && c == sg->split_steps) { /* one upload stream: its last event implies the others */
      CU_ERROR_CHECK(cuStreamWaitEvent(sp, sg->ev[sg->ngroups - 1], 0));
      CU_ERROR_CHECK(cuStreamWaitEvent(su, sg->ev[sg->ngroups - 1], 0));
    }
    if (me == owner) {
      const size_t lc = c / (size_t)P;
      const CUdeviceptr Ajj = chol_at(Aloc, ldl, j, lc * nb, es);
      if (staged) CU_ERROR_CHECK(cuStreamWaitEvent(sp, sg->ev[lc / sg->gblk], 0));
      if (!rest) {
        CU_ERROR_CHECK(chol_potrf_diag(prec, 'L', jb, Ajj, ldl, dinfo, (long)j, &w, sp));
        if (ga != NULL) {
          CUevent e = ga->ev[lc & 7];
          CU_ERROR_CHECK(cuEventRecord(e, sp));
          CU_ERROR_CHECK(cuStreamWaitEvent(ga->s_copy, e, 0));
          CU_ERROR_CHECK(copy2d(0, ga->hostA + (j * ga->lda + j) * es, ga->lda, Ajj, ldl, jb, jb, es, ga->s_copy));
        }
      }
      if (rest) {
        const CUdeviceptr Pn = chol_at(Aloc, ldl, j + jb, lc * nb, es);
        CU_ERROR_CHECK(chol_panel_factor_solve_lower(h, prec, jb, rest, Ajj, Pn, ldl, W, ldw, dinfo, (long)j, &w, sp, look_pending));
        look_pending = NULL;
      }
    }
    if (rest == 0) break;
    if (P > 1) /* the factored panel travels GPU-to-GPU over NVLink */
      NCCL_CHECK(nccl.Broadcast((const void *)W, (void *)W, ldw * jb * es, CHOL_NCCL_UINT8, owner, R->comm, sp));
    if (me == owner) { /* write-back into the matrix is off the critical path: after the broadcast has been queued */
      const CUdeviceptr Pn = chol_at(Aloc, ldl, j + jb, (c / (size_t)P) * nb, es);
      RT_CHECK(chol_copy_launch(prec, 'N', 0, rest, jb, (const void *)W, ldw, (void *)Pn, ldl, (void *)sp));
      if (ga != NULL) { /* block column c is final: D2H on the copy stream, overlapped with the remaining steps */
        CUevent e = ga->ev[(c / (size_t)P) & 7];
        CU_ERROR_CHECK(cuEventRecord(e, sp));
        CU_ERROR_CHECK(cuStreamWaitEvent(ga->s_copy, e, 0));
        CU_ERROR_CHECK(copy2d(0, ga->hostA + (j * ga->lda + j) * es, ga->lda, chol_at(Aloc, ldl, j, (c / (size_t)P) * nb, es), ldl,
                              n - j, jb, es, ga->s_copy));
      }
    }
    CU_ERROR_CHECK(cuEventRecord(h->ev_panel[c % CHOL_EVENT_RING], sp));

    /* trailing update of every owned block column with global index >= c+2, one launch */
    const size_t lc2 = (c + 2 > (size_t)me) ? (c + 2 - (size_t)me + (size_t)P - 1) / (size_t)P : 0;
    const size_t col2 = lc2 * nb;
    if (col2 < my_cols) {
      const size_t g_first = (lc2 * (size_t)P + (size_t)me) * nb; /* global column of local column col2 */
      const size_t m2 = n - g_first, n2 = my_cols - col2;
      const CUdeviceptr W2 = chol_at(W, ldw, g_first - (j + jb), 0, es);
      const CUdeviceptr D = chol_at(Aloc, ldl, g_first, col2, es);
      CU_ERROR_CHECK(cuStreamWaitEvent(su, h->ev_panel[c % CHOL_EVENT_RING], 0));
      const bool prof = !staged && h->profile && h->prof_n < h->prof_cap;
      if (prof) CU_ERROR_CHECK(cuEventRecord(h->prof_e0[h->prof_n], su));
      if (!staged) {
        CU_ERROR_CHECK(gemm_cyc(prec, herm, m2, n2, jb, W2, ldw, D, ldl, P > 1 ? (int)nb : 0, P, su));
      } else { /* one launch per upload group, each as soon as its columns are on the device */
        const size_t nlocb = (my_cols + nb - 1) / nb;
        for (size_t lcs = lc2; lcs < nlocb;) {
          const size_t g = lcs / sg->gblk, lce = chol_min((g + 1) * sg->gblk, nlocb);
          const size_t c0 = lcs * nb, c1 = chol_min(lce * nb, my_cols);
          const size_t gf = (lcs * (size_t)P + (size_t)me) * nb;
          CU_ERROR_CHECK(cuStreamWaitEvent(su, sg->ev[g], 0));
          CU_ERROR_CHECK(gemm_cyc(prec, herm, n - gf, c1 - c0, jb, chol_at(W, ldw, gf - (j + jb), 0, es), ldw,
                                  chol_at(Aloc, ldl, gf, c0, es), ldl, P > 1 ? (int)nb : 0, P, su));
          lcs = lce;
        }
      }
      if (prof) {
        CU_ERROR_CHECK(cuEventRecord(h->prof_e1[h->prof_n], su));
        /* algorithmic flops of this rank's share: sum over owned block columns of rows*width*k */
        double fl = 0.0;
        for (size_t lc = lc2; lc * nb < my_cols; lc++) {
          const size_t g = (lc * (size_t)P + (size_t)me) * nb, wd = chol_min(nb, n - g);
          fl += ((double)(n - g) * (double)wd - 0.5 * (double)wd * (double)(wd - 1)) * 2.0 * (double)jb;
        }
        h->prof_flops[h->prof_n++] = fl * (herm ? 4.0 : 1.0);
      }
    }
    CU_ERROR_CHECK(cuEventRecord(h->ev_rest[c % CHOL_EVENT_RING], su));

    /* look-ahead: the owner of block column c+1 brings it up to date on the panel stream */
    if ((int)((c + 1) % (size_t)P) == me) {
      const size_t lc1 = (c + 1) / (size_t)P, jb2 = chol_min(nb, rest);
      const CUdeviceptr A22 = chol_at(Aloc, ldl, j + jb, lc1 * nb, es);
      if (c >= 1) CU_ERROR_CHECK(cuStreamWaitEvent(sp, h->ev_rest[(c - 1) % CHOL_EVENT_RING], 0));
      if (staged) CU_ERROR_CHECK(cuStreamWaitEvent(sp, sg->ev[lc1 / sg->gblk], 0));
      const size_t below = rest - jb2;
      if (split && below) { /* diagonal block here, the rows below it on the look-ahead stream (see chol_lookahead_split) */
        CU_ERROR_CHECK(cuStreamWaitEvent(sl, h->ev_panel[c % CHOL_EVENT_RING], 0)); /* W has arrived / been solved */
        if (c >= 1) CU_ERROR_CHECK(cuStreamWaitEvent(sl, h->ev_rest[(c - 1) % CHOL_EVENT_RING], 0));
        if (staged) CU_ERROR_CHECK(cuStreamWaitEvent(sl, sg->ev[lc1 / sg->gblk], 0));
        CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'C', 0, 0, 0, 0, 'L', herm, jb2, jb2, jb, MONE, W, ldw, W, ldw, ONE, A22, ldl, A22, ldl, sp));
        const CUdeviceptr A32 = chol_at(A22, ldl, jb2, 0, es);
        CU_ERROR_CHECK(chol_gemm(prec, 'N', 'C', below, jb2, jb, MONE, chol_at(W, ldw, jb2, 0, es), ldw, W, ldw, ONE, A32, ldl, A32, ldl, sl));
        look_pending = h->ev_look[c % CHOL_EVENT_RING];
        CU_ERROR_CHECK(cuEventRecord(look_pending, sl));
      } else {
        CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'C', 0, 0, 0, 0, 'L', herm, rest, jb2, jb, MONE, W, ldw, W, ldw, ONE, A22, ldl,
                                    A22, ldl, sp));
      }
    }
  }
  CU_ERROR_CHECK(cuMemcpyDtoHAsync(h->hinfo, h->dinfo, 16, sp));
  CU_ERROR_CHECK(cuStreamSynchronize(sp));
  CU_ERROR_CHECK(cuStreamSynchronize(su));
  CU_ERROR_CHECK(cuStreamSynchronize(sl));
  if (ga != NULL) CU_ERROR_CHECK(cuStreamSynchronize(ga->s_copy));
  *info = *h->hinfo;
  h->last_logdet = ((const double *)h->hinfo)[1]; /* the diagonal blocks THIS rank factored: sum over ranks = log det A */
  h->last_logdet_valid = (*info == 0);
  for (int i = 0; i < h->prof_n; i++) {
    float ms = 0.f;
    CU_ERROR_CHECK(cuEventElapsedTime(&ms, h->prof_e0[i], h->prof_e1[i]));
    h->prof_total_ms += ms;
    h->prof_total_flops += h->prof_flops[i];
    h->prof_launches++;
    if (h->prof_hist_n < h->prof_cap) {
      h->prof_hist_ms[h->prof_hist_n] = ms;
      h->prof_hist_flops[h->prof_hist_n++] = h->prof_flops[i];
    }
  }
  h->prof_n = 0;
  return CUDA_SUCCESS;
}

/* =============================================================================================
 * Distributed TRTRI / LAUUM / POTRI on the same 1-D block-cyclic layout (reference entry
 * points lapack/dtrtri.c:470-530, dlauum.c:294-352, dpotri.c:48-69).  Lower form; one exchange per block step.
 *
 * TRTRI (in place, X = inv(L)), sweep over block columns k = 0, 1, ...  State before step k: the first k block columns
 * hold  [ X00 ; M ]  with X00 = inv(L00) and M = -L_B0 X00 ("still to be multiplied by inv(L_BR) from the left"),
 * block columns >= k hold the untouched L.  Step k with L_BR = [L11 0; L21 L22], M = [M1; M2]:
 *      X11 = inv(L11)                      owner of column k, then BROADCAST of the panel [X11; L21]
 *      X10 = X11 M1                        every rank, on its own columns < k  (row block k becomes final)
 *      M2 -= L21 X10                       every rank, on its own columns < k  (the n^3/3 flops, one GEMM per rank)
 *      L21 <- -L21 X11, L11 <- X11         owner of column k
 * The panels are columns of the ORIGINAL factor (steps < k never write block columns >= k), so the broadcasts do not
 * wait for any update: they run ahead on the panel stream, double-buffered, while the update stream computes.
 *
 * LAUUM (in place, C = L^H L, lower), sweep over block ROWS k = 0, 1, ...:  C[0:k+1, 0:k+1] += L[k, 0:k+1]^H L[k, 0:k+1].
 * Every rank contributes its columns' piece of row block k (conjugate-transposed and packed), the pieces are broadcast
 * in one NCCL group and scattered to their block-cyclic rows of a tall panel TT = L[k, 0:k+1]^H ((k+1) nb x nb); then
 * C(i, c) += TT(i, :) TT(gcol(c), :)^H for the rank's own columns c - the same block-cyclic-aware tile launch as the
 * POTRF trailing update, with alpha = +1.  Row block k of the matrix is consumed in step k and turns from L into C.
 * ===========================================================================================*/
static size_t cols_upto_block(size_t n, size_t nb, int r, int P, long kmax) { /* columns of rank r with global block <= kmax */
  size_t cols = 0;
  for (long c = r; c <= kmax && (size_t)c * nb < n; c += P) cols += chol_min(nb, n - (size_t)c * nb);
  return cols;
}

static CUresult dist_trtri(chol_rank *R, int prec, int diag, size_t n, CUdeviceptr Aloc, size_t ldl, CUdeviceptr tmp) {
  CULAPACKhandle h = R->lap;
  const int P = R->nranks, me = R->rank;
  chol_ws w;
  CU_ERROR_CHECK(dist_carve(R, prec, n, &w));
  const size_t es = w.es, nb = w.nb, nblk = (n + nb - 1) / nb;
  const int unit = diag == 'U';
  long *dinfo = (long *)h->dinfo;
  CUstream sc = h->s_panel, su = h->s_update;
  CU_ERROR_CHECK(cuMemsetD8Async(h->dinfo, 0, sizeof(long), sc)); /* zero pivots were excluded on the host: a sink */
  for (size_t k = 0; k < nblk; k++) {
    const size_t j = k * nb, jb = chol_min(nb, n - j), rest = n - j - jb;
    const int owner = (int)(k % (size_t)P);
    const CUdeviceptr W = w.W[k & 1];
    const size_t ldp = chol_roundup(jb + rest, 16);
    const size_t lc = k / (size_t)P;
    const CUdeviceptr Ajj = chol_at(Aloc, ldl, j, lc * nb, es), A21 = chol_at(Aloc, ldl, j + jb, lc * nb, es);
    /* ---- panel stream: [X11; L21] into W, broadcast ---- */
    if (k >= 2) CU_ERROR_CHECK(cuStreamWaitEvent(sc, h->ev_rest[(k - 2) % CHOL_EVENT_RING], 0)); /* W's last reader */
    if (me == owner) {
      CU_ERROR_CHECK(chol_trtri_small(prec, 'L', diag, jb, Ajj, ldl, W, ldp, dinfo, (long)j, &w, sc));
      if (rest) RT_CHECK(chol_copy_launch(prec, 'N', 0, rest, jb, (const void *)A21, ldl, (void *)chol_at(W, ldp, jb, 0, es), ldp, (void *)sc));
    }
    if (P > 1) NCCL_CHECK(nccl.Broadcast((const void *)W, (void *)W, ldp * jb * es, CHOL_NCCL_UINT8, owner, R->comm, sc));
    CU_ERROR_CHECK(cuEventRecord(h->ev_panel[k % CHOL_EVENT_RING], sc));
    /* ---- update stream ---- */
    CU_ERROR_CHECK(cuStreamWaitEvent(su, h->ev_panel[k % CHOL_EVENT_RING], 0));
    const size_t cb = k ? cols_upto_block(n, nb, me, P, (long)k - 1) : 0; /* own columns left of block column k */
    if (cb) {
      const CUdeviceptr M1 = chol_at(Aloc, ldl, j, 0, es);
      CU_ERROR_CHECK(chol_gemm(prec, 'N', 'N', jb, cb, jb, ONE, W, ldp, M1, ldl, ZERO2, 0, 0, tmp, jb, su));
      RT_CHECK(chol_copy_launch(prec, 'N', 0, jb, cb, (const void *)tmp, jb, (void *)M1, ldl, (void *)su));
      if (rest) {
        const CUdeviceptr M2 = chol_at(Aloc, ldl, j + jb, 0, es);
        CU_ERROR_CHECK(chol_gemm(prec, 'N', 'N', rest, cb, jb, MONE, chol_at(W, ldp, jb, 0, es), ldp, tmp, jb, ONE, M2, ldl, M2, ldl, su));
      }
    }
    if (me == owner) {
      if (rest)
        CU_ERROR_CHECK(chol_gemm(prec, 'N', 'N', rest, jb, jb, MONE, chol_at(W, ldp, jb, 0, es), ldp, W, ldp, ZERO2, 0, 0, A21, ldl, su));
      /* a unit-diagonal inverse leaves the caller's diagonal untouched (lapack/dtrtri.c:56-57) */
      RT_CHECK(chol_copy_launch(prec, 'N', unit ? 'l' : 'L', jb, jb, (const void *)W, ldp, (void *)Ajj, ldl, (void *)su));
    }
    CU_ERROR_CHECK(cuEventRecord(h->ev_rest[k % CHOL_EVENT_RING], su));
  }
  CU_ERROR_CHECK(cuStreamSynchronize(sc));
  CU_ERROR_CHECK(cuStreamSynchronize(su));
  return CUDA_SUCCESS;
}

static CUresult dist_lauum(chol_rank *R, int prec, size_t n, CUdeviceptr Aloc, size_t ldl) {
  CULAPACKhandle h = R->lap;
  const int P = R->nranks, me = R->rank;
  chol_ws w;
  CU_ERROR_CHECK(dist_carve(R, prec, n, &w));
  const size_t es = w.es, nb = w.nb, nblk = (n + nb - 1) / nb;
  const int herm = chol_is_complex(prec);
  CUstream su = h->s_update;
  const CUdeviceptr TT = w.W[0], S = w.W[1]; /* TT: (k+1) nb x nb in global row order; S: the P packed pieces, one after the other */
  const size_t ldt = w.ldw;
  for (size_t k = 0; k < nblk; k++) {
    const size_t j = k * nb, jb = chol_min(nb, n - j); /* row block k: rows j .. j+jb-1, columns 0 .. j+jb-1 */
    const int owner = (int)(k % (size_t)P);
    /* pieces: rank r holds cols_r columns of row block k (its block columns <= k); piece r = their conjugate transpose,
     * cols_r x jb, packed with leading dimension cols_r at element offset off_r of S */
    size_t off = 0, my_off = 0, my_cols = 0;
    for (int r = 0; r < P; r++) {
      const size_t cr = cols_upto_block(n, nb, r, P, (long)k);
      if (r == me) { my_off = off; my_cols = cr; }
      off += cr * jb;
    }
    const size_t cb = k ? cols_upto_block(n, nb, me, P, (long)k - 1) : 0; /* own columns left of the diagonal block */
    const CUdeviceptr Rk = chol_at(Aloc, ldl, j, 0, es);                  /* row block k of the own columns */
    if (my_cols) {
      const CUdeviceptr Sme = S + (CUdeviceptr)(my_off * es);
      if (cb) RT_CHECK(chol_copy_launch(prec, 'C', 0, jb, cb, (const void *)Rk, ldl, (void *)Sme, my_cols, (void *)su));
      if (me == owner) { /* the diagonal block: only its lower triangle is the factor (the other triangle is the caller's) */
        const CUdeviceptr Sd = Sme + (CUdeviceptr)(cb * es);
        RT_CHECK(chol_fill_launch(prec, 0, jb, jb, 0.0, 0.0, (void *)Sd, my_cols, (void *)su));
        RT_CHECK(chol_copy_launch(prec, 'C', 'L', jb, jb, (const void *)chol_at(Rk, ldl, 0, cb, es), ldl, (void *)Sd, my_cols, (void *)su));
      }
    }
    if (P > 1) {
      NCCL_CHECK(nccl.GroupStart());
      size_t o = 0;
      for (int r = 0; r < P; r++) {
        const size_t cr = cols_upto_block(n, nb, r, P, (long)k);
        if (cr) {
          const CUdeviceptr Sr = S + (CUdeviceptr)(o * es);
          NCCL_CHECK(nccl.Broadcast((const void *)Sr, (void *)Sr, cr * jb * es, CHOL_NCCL_UINT8, r, R->comm, su));
        }
        o += cr * jb;
      }
      NCCL_CHECK(nccl.GroupEnd());
    }
    { /* scatter every piece to its block-cyclic rows of TT */
      size_t o = 0;
      for (int r = 0; r < P; r++) {
        const size_t cr = cols_upto_block(n, nb, r, P, (long)k);
        if (cr)
          RT_CHECK(chol_scatter_rows_launch(prec, cr, jb, (const void *)(S + (CUdeviceptr)(o * es)), cr, (void *)TT, ldt, (int)nb, P, r, (void *)su));
        o += cr * jb;
      }
    }
    if (cb) {
      /* row block k of the own columns has been read: it becomes an accumulator of C */
      RT_CHECK(chol_fill_launch(prec, 0, jb, cb, 0.0, 0.0, (void *)Rk, ldl, (void *)su));
      CU_ERROR_CHECK(gemm_cyc_ex(prec, herm, 1.0, j + jb, cb, jb, TT, ldt, Aloc, ldl, (int)nb, P, me, su));
    }
    if (me == owner) { /* C_kk = L_kk^H L_kk replaces the factor's diagonal block (lower triangle only) */
      const CUdeviceptr TTk = chol_at(TT, ldt, j, 0, es), Ckk = chol_at(Rk, ldl, 0, cb, es);
      CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'C', 0, 0, 0, 0, 'L', herm, jb, jb, jb, ONE, TTk, ldt, TTk, ldt, ZERO2, 0, 0, Ckk, ldl, su));
    }
  }
  CU_ERROR_CHECK(cuStreamSynchronize(su));
  return CUDA_SUCCESS;
}

/* =============================================================================================
 * one-process-per-GPU entry points (C-ABI extension used under torchrun)
 * ===========================================================================================*/

/* 128-byte NCCL unique id, to be created on rank 0 and sent to the other ranks by the launcher */
CUresult cholDistGetUniqueId(void *id128) {
  CU_ERROR_CHECK(nccl_load());
  chol_nccl_id id;
  NCCL_CHECK(nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return CUDA_SUCCESS;
}

static CUresult dist_comm_init(chol_rank *R, const void *id128) {
  CU_ERROR_CHECK(nccl_load());
  chol_nccl_id id;
  memcpy(&id, id128, sizeof id);
  if (nccl.CommInitRankConfig != NULL) {
    chol_nccl_config cfg = {sizeof(chol_nccl_config), 0xcafebeef, 22703, CHOL_NCCL_UNDEF_INT, CHOL_NCCL_UNDEF_INT,
                            CHOL_NCCL_UNDEF_INT, nccl_max_ctas(), NULL, CHOL_NCCL_UNDEF_INT, CHOL_NCCL_UNDEF_INT, NULL,
                            CHOL_NCCL_UNDEF_INT, CHOL_NCCL_UNDEF_INT, CHOL_NCCL_UNDEF_INT, CHOL_NCCL_UNDEF_INT};
    NCCL_CHECK(nccl.CommInitRankConfig(&R->comm, R->nranks, id, R->rank, &cfg));
  } else
    NCCL_CHECK(nccl.CommInitRank(&R->comm, R->nranks, id, R->rank));
  return CUDA_SUCCESS;
}

/* binds the calling thread's current context; nranks == 1 needs neither an id nor NCCL */
CUresult cholDistCreate(cholDist *out, int rank, int nranks, const void *id128) {
  if (out == NULL || nranks < 1 || rank < 0 || rank >= nranks) return CUDA_ERROR_INVALID_VALUE;
  chol_rank *R = calloc(1, sizeof *R);
  if (R == NULL) return CUDA_ERROR_OUT_OF_MEMORY;
  R->rank = rank;
  R->nranks = nranks;
  CUresult e = cuLAPACKCreate(&R->lap);
  if (e != CUDA_SUCCESS) {
    free(R);
    return e;
  }
  R->context = R->lap->context;
  if (nranks > 1) {
    e = (id128 == NULL) ? CUDA_ERROR_INVALID_VALUE : dist_comm_init(R, id128);
    if (e != CUDA_SUCCESS) {
      cuLAPACKDestroy(R->lap);
      free(R);
      return e;
    }
  }
  *out = R;
  return CUDA_SUCCESS;
}

CUresult cholDistDestroy(cholDist R) {
  if (R == NULL) return CUDA_ERROR_INVALID_VALUE;
  if (R->comm) nccl.CommDestroy(R->comm);
  if (R->lap) cuLAPACKDestroy(R->lap);
  free(R);
  return CUDA_SUCCESS;
}

CUresult cholDistSetBlockSize(cholDist R, size_t nb) { return R ? cuLAPACKSetBlockSize(R->lap, nb) : CUDA_ERROR_INVALID_VALUE; }
CULAPACKhandle cholDistLapackHandle(cholDist R) { return R ? R->lap : NULL; }
size_t cholDistBlockSize(cholDist R, int prec, size_t n) { return R ? dist_nb(R, prec, n) : chol_default_nb(prec, n); }
/* columns of an n x n matrix held by `rank` (they are packed side by side, block c = lc*P + rank) */
size_t cholDistLocalCols(size_t n, size_t nb, int rank, int nranks) { return local_cols(n, nb, rank, nranks); }

#define GEN_DIST(U, p)                                                                                      \
  /* local part of A = L*L^H in the block-cyclic lower layout; info = first failure seen by THIS rank */    \
  CUresult cholDist##U##potrf(cholDist R, size_t n, CUdeviceptr Aloc, size_t ldl, long *info) {             \
    *info = 0;                                                                                              \
    if (R == NULL) return CUDA_ERROR_INVALID_VALUE;                                                         \
    if (ldl < n) {                                                                                          \
      *info = -4;                                                                                           \
      XERBLA(4);                                                                                            \
      return CUDA_ERROR_INVALID_VALUE;                                                                      \
    }                                                                                                       \
    if (n == 0) return CUDA_SUCCESS;                                                                        \
    CU_ERROR_CHECK(cuCtxPushCurrent(R->context));                                                           \
    const CUresult e = dist_potrf(R, p, n, Aloc, ldl, info, NULL, NULL);                                    \
    CUcontext dummy;                                                                                        \
    cuCtxPopCurrent(&dummy);                                                                                \
    return e;                                                                                               \
  }
GEN_DIST(S, 's')
GEN_DIST(D, 'd')
GEN_DIST(C, 'c')
GEN_DIST(Z, 'z')

/* =============================================================================================
 * cuMultiGPU LAPACK handle and the host-pointer POTRF (reference lapack/handle.c:90-108,
 * lapack/dpotrf.c:298-366)
 * ===========================================================================================*/
struct __cumultigpulapackhandle_st {
  CUmultiGPU mgpu;
};
struct __cumultigpublashandle_st {
  CUmultiGPU mgpu;
};

static CUresult rank_init_task(const void *args) {
  chol_rank *R = *(chol_rank *const *)args;
  CUresult e = cuLAPACKCreate(&R->lap); /* adopts the worker's context */
  if (e == CUDA_SUCCESS) R->context = R->lap->context;
  return e;
}

/* NCCL reads NCCL_MAX_CTAS when a communicator is created; set it for the duration of
 * ncclCommInitAll only (an explicit setting by the user wins) and put the environment back */
static int nccl_comm_init_all_capped(chol_nccl_comm *comms, int P, const int *devs) {
  const bool user_set = getenv("NCCL_MAX_CTAS") != NULL;
  if (!user_set) {
    char buf[16];
    snprintf(buf, sizeof buf, "%d", nccl_max_ctas());
    setenv("NCCL_MAX_CTAS", buf, 0);
  }
  const int rc = nccl.CommInitAll(comms, P, devs);
  if (!user_set) unsetenv("NCCL_MAX_CTAS");
  return rc;
}

static void release_ranks(CUmultiGPU m);

static CUresult ensure_ranks(CUmultiGPU m) {
  if (m->ranks != NULL) return CUDA_SUCCESS;
  const int P = m->n;
  chol_rank *ranks = calloc((size_t)P, sizeof *ranks);
  chol_nccl_comm *comms = calloc((size_t)P, sizeof *comms);
  chol_rank **ptrs = malloc(sizeof(*ptrs) * (size_t)P);
  int *devs = malloc(sizeof(int) * (size_t)P);
  CUresult e = (ranks && comms && ptrs && devs) ? CUDA_SUCCESS : CUDA_ERROR_OUT_OF_MEMORY;
  if (e == CUDA_SUCCESS && P > 1) {
    e = nccl_load();
    if (e == CUDA_SUCCESS) {
      for (int i = 0; i < P; i++) devs[i] = (int)m->w[i].device;
      const int rc = nccl_comm_init_all_capped(comms, P, devs);
      if (rc != 0) {
        if (errorHandler != NULL) errorHandler("ncclCommInitAll", __func__, __FILE__, __LINE__, rc, nccl_strerror);
        e = CUDA_ERROR_UNKNOWN;
      }
    }
  }
  if (e == CUDA_SUCCESS) {
    for (int i = 0; i < P; i++) {
      ranks[i].rank = i;
      ranks[i].nranks = P;
      ranks[i].comm = comms[i];
      ptrs[i] = &ranks[i];
    }
    m->ranks = ranks;
    m->comms = comms;
    e = run_on_all(m, rank_init_task, ptrs, sizeof(*ptrs), sizeof(*ptrs));
    if (e != CUDA_SUCCESS) release_ranks(m); /* a half-initialised set must not look ready */
  } else {
    free(ranks);
    free(comms);
  }
  free(ptrs);
  free(devs);
  return e;
}

static CUresult rank_free_task(const void *args) {
  chol_rank *R = *(chol_rank *const *)args;
  if (R->local) cuMemFree(R->local);
  if (R->lap) cuLAPACKDestroy(R->lap);
  R->local = 0;
  R->lap = NULL;
  return CUDA_SUCCESS;
}

/* per-GPU solver state: device buffers and LAPACK handles go on their own workers, then the comms */
static void release_ranks(CUmultiGPU m) {
  if (m->ranks == NULL) return;
  chol_rank **ptrs = malloc(sizeof(*ptrs) * (size_t)m->n);
  if (ptrs != NULL) {
    for (int i = 0; i < m->n; i++) ptrs[i] = &m->ranks[i];
    run_on_all(m, rank_free_task, ptrs, sizeof(*ptrs), sizeof(*ptrs));
    free(ptrs);
  }
  for (int i = 0; i < m->n; i++)
    if (m->comms[i]) nccl.CommDestroy(m->comms[i]);
  free(m->ranks);
  free(m->comms);
  m->ranks = NULL;
  m->comms = NULL;
}

CUresult cuMultiGPUDestroy(CUmultiGPU m) {
  if (m == NULL) return CUDA_ERROR_INVALID_VALUE;
  release_ranks(m);
  for (int i = 0; i < m->n; i++) {
    struct worker *w = &m->w[i];
    pthread_mutex_lock(&w->mutex);
    w->stop = true;
    pthread_cond_signal(&w->cond);
    pthread_mutex_unlock(&w->mutex);
    pthread_join(w->thread, NULL);
    pthread_mutex_destroy(&w->mutex);
    pthread_cond_destroy(&w->cond);
    cuDevicePrimaryCtxRelease(w->device);
  }
  free(m->w);
  free(m);
  return CUDA_SUCCESS;
}

CUresult cuMultiGPULAPACKCreate(CUmultiGPULAPACKhandle *handle, CUmultiGPU mGPU) {
  if (handle == NULL || mGPU == NULL) return CUDA_ERROR_INVALID_VALUE;
  struct __cumultigpulapackhandle_st *h = calloc(1, sizeof *h);
  if (h == NULL) return CUDA_ERROR_OUT_OF_MEMORY;
  h->mgpu = mGPU;
  const CUresult e = ensure_ranks(mGPU);
  if (e != CUDA_SUCCESS) {
    free(h);
    return e;
  }
  *handle = h;
  return CUDA_SUCCESS;
}
CUresult cuMultiGPULAPACKDestroy(CUmultiGPULAPACKhandle h) {
  if (h == NULL) return CUDA_ERROR_INVALID_VALUE;
  free(h);
  return CUDA_SUCCESS;
}
CUresult cuMultiGPULAPACKSynchronize(CUmultiGPULAPACKhandle h) {
  return h == NULL ? CUDA_ERROR_INVALID_VALUE : cuMultiGPUSynchronize(h->mgpu);
}
CUresult cuMultiGPUBLASCreate(CUmultiGPUBLAShandle *handle, CUmultiGPU mGPU) {
  if (handle == NULL || mGPU == NULL) return CUDA_ERROR_INVALID_VALUE;
  struct __cumultigpublashandle_st *h = calloc(1, sizeof *h);
  if (h == NULL) return CUDA_ERROR_OUT_OF_MEMORY;
  h->mgpu = mGPU;
  const CUresult e = ensure_ranks(mGPU);
  if (e != CUDA_SUCCESS) {
    free(h);
    return e;
  }
  *handle = h;
  return CUDA_SUCCESS;
}
CUresult cuMultiGPUBLASDestroy(CUmultiGPUBLAShandle h) {
  if (h == NULL) return CUDA_ERROR_INVALID_VALUE;
  free(h);
  return CUDA_SUCCESS;
}
CUresult cuMultiGPUBLASSynchronize(CUmultiGPUBLAShandle h) {
  return h == NULL ? CUDA_ERROR_INVALID_VALUE : cuMultiGPUSynchronize(h->mgpu);
}
CUmultiGPU chol_mgpu_of_lapack(CUmultiGPULAPACKhandle h) { return h ? h->mgpu : NULL; }
CUmultiGPU chol_mgpu_of_blas(CUmultiGPUBLAShandle h) { return h ? h->mgpu : NULL; }
int chol_mgpu_device(CUmultiGPU m, int i) { return (int)m->w[i].device; }
CULAPACKhandle chol_mgpu_rank_lapack(CUmultiGPU m, int i) { return m->ranks ? m->ranks[i].lap : NULL; }
CUresult chol_mgpu_run_on_all(CUmultiGPU m, CUresult (*fn)(const void *), const void *args, size_t stride, size_t size) {
  return run_on_all(m, fn, args, stride, size);
}

/* ---- scatter / factor / gather, executed by every worker ---- */
typedef struct {
  chol_rank *R;
  int prec, uplo;
  size_t n, lda;
  void *hostA;
  long info;
} potrf_args;

static CUresult copy2d(int h2d, void *host, size_t ldh, CUdeviceptr dev, size_t ldd, size_t rows, size_t cols, size_t es,
                       CUstream st) {
  CUDA_MEMCPY2D c;
  memset(&c, 0, sizeof c);
  if (h2d) {
    c.srcMemoryType = CU_MEMORYTYPE_HOST; c.srcHost = host; c.srcPitch = ldh * es;
    c.dstMemoryType = CU_MEMORYTYPE_DEVICE; c.dstDevice = dev; c.dstPitch = ldd * es;
  } else {
    c.srcMemoryType = CU_MEMORYTYPE_DEVICE; c.srcDevice = dev; c.srcPitch = ldd * es;
    c.dstMemoryType = CU_MEMORYTYPE_HOST; c.dstHost = host; c.dstPitch = ldh * es;
  }
  c.WidthInBytes = rows * es;
  c.Height = cols;
  return cuMemcpy2DAsync(&c, st);
}

/* phase 1 (every rank, before anyone starts the collective factorisation): workspace and the
 * rank's resident block columns.  An allocation failure on one GPU then fails the call on all of
 * them instead of leaving the other ranks waiting in a broadcast. */
static CUresult rank_potrf_prepare_task(const void *argp) {
  potrf_args *a = *(potrf_args *const *)argp;
  chol_rank *R = a->R;
  const size_t n = a->n, es = chol_esize(a->prec);
  chol_ws w;
  CU_ERROR_CHECK(dist_carve(R, a->prec, n, &w));
  const size_t my_cols = local_cols(n, w.nb, R->rank, R->nranks);
  const size_t need = chol_roundup(n, 16) * (my_cols ? my_cols : 1) * es;
  if (need > R->local_bytes) {
    if (R->local) CU_ERROR_CHECK(cuMemFree(R->local));
    R->local = 0;
    R->local_bytes = 0;
    CU_ERROR_CHECK(cuMemAlloc(&R->local, need));
    R->local_bytes = need;
  }
  return CUDA_SUCCESS;
}

/* start factoring while the block columns are still being uploaded (lower, page-locked host memory).
 * Measured on B200 (profiles/r02_staged_ab.txt): e2e dpotrf N = 65536 30.36 -> 31.65 TFLOP/s.
 * CHOL_MGPU_H2D_OVERLAP=0 turns it off for A/B runs. */
static bool h2d_overlap_enabled(void) {
  const char *e = getenv("CHOL_MGPU_H2D_OVERLAP");
  return e == NULL || atoi(e) != 0;
}

/* number of leading steps whose update must be split by upload group: those that can start before
 * the upload (bytes at ~50 GB/s of PCIe, CHOL_H2D_GBPS overrides) has finished, at the update rate
 * the engine sustains (~30 real TFLOP/s FP64, ~35 FP32), with a 1.5x safety margin. */
static size_t staged_split_steps(int prec, size_t n, size_t nb, int P, double upload_bytes) {
  const char *e = getenv("CHOL_H2D_GBPS");
  const double gbps = (e != NULL && atof(e) > 0.0) ? atof(e) : 50.0;
  const double t_up = 1.5 * upload_bytes / (gbps * 1e9);
  const double rate = (prec == 'd' || prec == 'z') ? 30e12 : 35e12, cplx = chol_is_complex(prec) ? 4.0 : 1.0;
  const size_t nblk = (n + nb - 1) / nb;
  double t = 0.0;
  size_t c = 0;
  for (; c < nblk && t < t_up; c++) {
    const double rest = (double)(n - chol_min(n, (c + 1) * nb));
    t += cplx * rest * rest * (double)nb / (double)P / rate;
  }
  return c;
}

/* phase 2: scatter / factor / gather */
static CUresult rank_potrf_task(const void *argp) {
  potrf_args *a = *(potrf_args *const *)argp;
  chol_rank *R = a->R;
  const int P = R->nranks, me = R->rank, prec = a->prec;
  const size_t n = a->n, es = chol_esize(prec), lda = a->lda;
  CULAPACKhandle h = R->lap;
  chol_ws w;
  CU_ERROR_CHECK(dist_carve(R, prec, n, &w));
  const size_t nb = w.nb, nblk = (n + nb - 1) / nb, my_cols = local_cols(n, nb, me, P);
  const size_t ldl = chol_roundup(n, 16);
  CUstream st = h->s_update;
  char *hostA = a->hostA;
  unsigned int memtype = 0; /* overlap only with page-locked host memory: pageable copies would block this thread */
  const bool pinned = cuPointerGetAttribute(&memtype, CU_POINTER_ATTRIBUTE_MEMORY_TYPE, (CUdeviceptr)(size_t)hostA) == CUDA_SUCCESS &&
                      memtype == CU_MEMORYTYPE_HOST;
  const size_t nlocb = (my_cols + nb - 1) / nb;

  if (a->uplo == 'L' && pinned && h2d_overlap_enabled() && nlocb >= 2) {
    /* ---- upload, factorisation and download all overlap ---- */
    enum { MAXG = 16 }; /* upload groups: the factorisation starts when the first one (1/16 of the rank's block columns) has arrived */
    CUevent up_ev[MAXG] = {0};
    dist_stage sg;
    dist_gather ga;
    memset(&ga, 0, sizeof ga);
    sg.ev = up_ev;
    sg.gblk = (nlocb + MAXG - 1) / MAXG;
    sg.ngroups = (nlocb + sg.gblk - 1) / sg.gblk;
    ga.hostA = hostA;
    ga.lda = lda;
    CUstream s_up = NULL;
    CUresult e = cuStreamCreate(&s_up, CU_STREAM_NON_BLOCKING);
    if (e == CUDA_SUCCESS) e = cuStreamCreate(&ga.s_copy, CU_STREAM_NON_BLOCKING);
    for (size_t g = 0; g < sg.ngroups && e == CUDA_SUCCESS; g++) e = cuEventCreate(&up_ev[g], CU_EVENT_DISABLE_TIMING);
    for (int i = 0; i < 8 && e == CUDA_SUCCESS; i++) e = cuEventCreate(&ga.ev[i], CU_EVENT_DISABLE_TIMING);
    double bytes = 0.0;
    for (size_t c = (size_t)me, lc = 0; c < nblk && e == CUDA_SUCCESS; c += (size_t)P, lc++) {
      const size_t j = c * nb, jb = chol_min(nb, n - j), rows = n - j;
      e = copy2d(1, hostA + (j * lda + j) * es, lda, chol_at(R->local, ldl, j, lc * nb, es), ldl, rows, jb, es, s_up);
      bytes += (double)rows * (double)jb * (double)es;
      if (e == CUDA_SUCCESS && ((lc + 1) % sg.gblk == 0 || lc + 1 == nlocb)) e = cuEventRecord(up_ev[lc / sg.gblk], s_up);
    }
    if (e == CUDA_SUCCESS) {
      sg.split_steps = staged_split_steps(prec, n, nb, P, bytes);
      e = dist_potrf(R, prec, n, R->local, ldl, &a->info, &ga, &sg);
    }
    if (s_up) cuStreamSynchronize(s_up);
    if (ga.s_copy) cuStreamSynchronize(ga.s_copy);
    for (size_t g = 0; g < sg.ngroups; g++)
      if (up_ev[g]) cuEventDestroy(up_ev[g]);
    for (int i = 0; i < 8; i++)
      if (ga.ev[i]) cuEventDestroy(ga.ev[i]);
    if (s_up) cuStreamDestroy(s_up);
    if (ga.s_copy) cuStreamDestroy(ga.s_copy);
    CU_ERROR_CHECK(e);
    return CUDA_SUCCESS;
  }

  /* ---- scatter: block column c (lower) or block row c (upper, transposed on the device) ---- */
  for (size_t c = (size_t)me, lc = 0; c < nblk; c += (size_t)P, lc++) {
    const size_t j = c * nb, jb = chol_min(nb, n - j), rows = n - j;
    const CUdeviceptr dst = chol_at(R->local, ldl, j, lc * nb, es);
    if (a->uplo == 'L') {
      CU_ERROR_CHECK(copy2d(1, hostA + (j * lda + j) * es, lda, dst, ldl, rows, jb, es, st));
    } else { /* U[j:j+jb, j:n] -> staging (jb x rows) -> conj-transpose into the local column */
      const size_t lds = chol_roundup(jb, 2);
      CU_ERROR_CHECK(copy2d(1, hostA + (j * lda + j) * es, lda, w.W[lc & 1], lds, jb, rows, es, st));
      RT_CHECK(chol_copy_launch(prec, 'C', 0, jb, rows, (const void *)w.W[lc & 1], lds, (void *)dst, ldl, (void *)st));
    }
  }
  CU_ERROR_CHECK(cuStreamSynchronize(st));
  /* ---- factor (lower: finished block columns stream back to the host while later steps run) ---- */
  if (a->uplo == 'L' && pinned) {
    dist_gather ga;
    memset(&ga, 0, sizeof ga);
    ga.hostA = hostA;
    ga.lda = lda;
    CUresult e = cuStreamCreate(&ga.s_copy, CU_STREAM_NON_BLOCKING);
    for (int i = 0; i < 8 && e == CUDA_SUCCESS; i++) e = cuEventCreate(&ga.ev[i], CU_EVENT_DISABLE_TIMING);
    if (e == CUDA_SUCCESS) e = dist_potrf(R, prec, n, R->local, ldl, &a->info, &ga, NULL);
    for (int i = 0; i < 8; i++)
      if (ga.ev[i]) cuEventDestroy(ga.ev[i]);
    if (ga.s_copy) cuStreamDestroy(ga.s_copy);
    CU_ERROR_CHECK(e);
    return CUDA_SUCCESS;
  }
  CU_ERROR_CHECK(dist_potrf(R, prec, n, R->local, ldl, &a->info, NULL, NULL));
  /* ---- gather: whole block columns / rows travel back (upper: through a transposing staging
   * buffer).  The part of each diagonal block that lies in the caller's OTHER triangle was uploaded
   * with the block, is never written by the factorisation and returns bit-identical. */
  for (size_t c = (size_t)me, lc = 0; c < nblk; c += (size_t)P, lc++) {
    const size_t j = c * nb, jb = chol_min(nb, n - j), rows = n - j;
    const CUdeviceptr src = chol_at(R->local, ldl, j, lc * nb, es);
    if (a->uplo == 'L') {
      CU_ERROR_CHECK(copy2d(0, hostA + (j * lda + j) * es, lda, src, ldl, rows, jb, es, st));
    } else {
      const size_t lds = chol_roundup(jb, 2);
      RT_CHECK(chol_copy_launch(prec, 'C', 0, rows, jb, (const void *)src, ldl, (void *)w.W[lc & 1], lds, (void *)st));
      CU_ERROR_CHECK(copy2d(0, hostA + (j * lda + j) * es, lda, w.W[lc & 1], lds, jb, rows, es, st));
    }
  }
  CU_ERROR_CHECK(cuStreamSynchronize(st));
  return CUDA_SUCCESS;
}

static CUresult mgpu_potrf(CUmultiGPULAPACKhandle handle, int prec, int uplo, size_t n, void *A, size_t lda, long *info) {
  if (handle == NULL || A == NULL) return CUDA_ERROR_INVALID_VALUE;
  CUmultiGPU m = handle->mgpu;
  const int P = m->n;
  potrf_args *args = calloc((size_t)P, sizeof *args);
  potrf_args **ptrs = calloc((size_t)P, sizeof *ptrs);
  if (args == NULL || ptrs == NULL) {
    free(args);
    free(ptrs);
    return CUDA_ERROR_OUT_OF_MEMORY;
  }
  for (int i = 0; i < P; i++) {
    args[i].R = &m->ranks[i];
    args[i].prec = prec;
    args[i].uplo = uplo;
    args[i].n = n;
    args[i].lda = lda;
    args[i].hostA = A;
    ptrs[i] = &args[i];
  }
  CUresult e = run_on_all(m, rank_potrf_prepare_task, ptrs, sizeof(*ptrs), sizeof(*ptrs));
  if (e == CUDA_SUCCESS) e = run_on_all(m, rank_potrf_task, ptrs, sizeof(*ptrs), sizeof(*ptrs));
  long first = 0;
  for (int i = 0; i < P; i++)
    if (args[i].info > 0 && (first == 0 || args[i].info < first)) first = args[i].info;
  *info = first;
  free(args);
  free(ptrs);
  return e;
}

/* =============================================================================================
 * host-pointer TRTRI / LAUUM / POTRI over all GPUs: scatter into the block-cyclic layout, the distributed sweeps
 * above, gather.  Upper inputs travel as U^H (a lower factor): inv(U) = inv(L)^H and U U^H = L^H L.
 * ===========================================================================================*/
typedef struct {
  chol_rank *R;
  int op, prec, uplo, diag; /* op: 0 trtri, 1 lauum, 2 potri */
  size_t n, lda;
  void *hostA;
} inverse_args;

static CUresult rank_inverse_task(const void *argp) {
  inverse_args *a = *(inverse_args *const *)argp;
  chol_rank *R = a->R;
  const int P = R->nranks, me = R->rank, prec = a->prec;
  const size_t n = a->n, es = chol_esize(prec), lda = a->lda;
  chol_ws w;
  CU_ERROR_CHECK(dist_carve(R, prec, n, &w));
  const size_t nb = w.nb, nblk = (n + nb - 1) / nb, my_cols = local_cols(n, nb, me, P), ldl = chol_roundup(n, 16);
  CUstream st = R->lap->s_update;
  char *hostA = a->hostA;
  CUdeviceptr tmp = 0;
  if (a->op != 1) CU_ERROR_CHECK(cuMemAlloc(&tmp, nb * (my_cols ? my_cols : 1) * es));
  CUresult e = CUDA_SUCCESS;
  for (size_t c = (size_t)me, lc = 0; c < nblk && e == CUDA_SUCCESS; c += (size_t)P, lc++) {
    const size_t j = c * nb, jb = chol_min(nb, n - j), rows = n - j;
    const CUdeviceptr dst = chol_at(R->local, ldl, j, lc * nb, es);
    if (a->uplo == 'L') {
      e = copy2d(1, hostA + (j * lda + j) * es, lda, dst, ldl, rows, jb, es, st);
    } else { /* U[j:j+jb, j:n] -> staging (jb x rows) -> conjugate transpose into the local column */
      const size_t lds = chol_roundup(jb, 2);
      e = copy2d(1, hostA + (j * lda + j) * es, lda, w.W[lc & 1], lds, jb, rows, es, st);
      if (e == CUDA_SUCCESS && chol_copy_launch(prec, 'C', 0, jb, rows, (const void *)w.W[lc & 1], lds, (void *)dst, ldl, (void *)st) != 0)
        e = CUDA_ERROR_LAUNCH_FAILED;
    }
  }
  if (e == CUDA_SUCCESS) e = cuStreamSynchronize(st);
  if (e == CUDA_SUCCESS && a->op != 1) e = dist_trtri(R, prec, a->diag, n, R->local, ldl, tmp);
  if (e == CUDA_SUCCESS && a->op != 0) e = dist_lauum(R, prec, n, R->local, ldl);
  for (size_t c = (size_t)me, lc = 0; c < nblk && e == CUDA_SUCCESS; c += (size_t)P, lc++) {
    const size_t j = c * nb, jb = chol_min(nb, n - j), rows = n - j;
    const CUdeviceptr src = chol_at(R->local, ldl, j, lc * nb, es);
    if (a->uplo == 'L') {
      e = copy2d(0, hostA + (j * lda + j) * es, lda, src, ldl, rows, jb, es, st);
    } else {
      const size_t lds = chol_roundup(jb, 2);
      if (chol_copy_launch(prec, 'C', 0, rows, jb, (const void *)src, ldl, (void *)w.W[lc & 1], lds, (void *)st) != 0) e = CUDA_ERROR_LAUNCH_FAILED;
      if (e == CUDA_SUCCESS) e = copy2d(0, hostA + (j * lda + j) * es, lda, w.W[lc & 1], lds, jb, rows, es, st);
    }
  }
  const CUresult es2 = cuStreamSynchronize(st);
  if (e == CUDA_SUCCESS) e = es2;
  if (tmp) cuMemFree(tmp);
  CU_ERROR_CHECK(e);
  return CUDA_SUCCESS;
}

/* a zero on the diagonal of a non-unit triangular matrix: the reference reports the first one its sweep meets -
 * bottom-up for lower, top-down for upper (lapack/dtrtri.c:49-52, :73-77) - and stops; here nothing is computed */
static long host_zero_pivot(int prec, int uplo, size_t n, const void *A, size_t lda) {
  long found = 0;
  for (size_t i = 0; i < n; i++) {
    const size_t at = i * lda + i;
    bool zero;
    switch (prec) {
      case 's': zero = ((const float *)A)[at] == 0.0f; break;
      case 'd': zero = ((const double *)A)[at] == 0.0; break;
      case 'c': zero = ((const float *)A)[2 * at] == 0.0f && ((const float *)A)[2 * at + 1] == 0.0f; break;
      default: zero = ((const double *)A)[2 * at] == 0.0 && ((const double *)A)[2 * at + 1] == 0.0; break;
    }
    if (zero) {
      if (uplo != 'L') return (long)i + 1;
      found = (long)i + 1;
    }
  }
  return found;
}

/* op: 0 trtri, 1 lauum, 2 potri */
CUresult chol_mgpu_inverse_dist(CUmultiGPU m, int op, int prec, int uplo, int diag, size_t n, void *A, size_t lda, long *info) {
  if (m == NULL || A == NULL) return CUDA_ERROR_INVALID_VALUE;
  *info = 0;
  if (op != 1 && diag != 'U') {
    *info = host_zero_pivot(prec, uplo, n, A, lda);
    if (*info != 0) return CUDA_SUCCESS;
  }
  CU_ERROR_CHECK(ensure_ranks(m));
  const int P = m->n;
  inverse_args *args = calloc((size_t)P, sizeof *args);
  inverse_args **ptrs = calloc((size_t)P, sizeof *ptrs);
  potrf_args *pargs = calloc((size_t)P, sizeof *pargs);
  potrf_args **pptrs = calloc((size_t)P, sizeof *pptrs);
  if (args == NULL || ptrs == NULL || pargs == NULL || pptrs == NULL) {
    free(args); free(ptrs); free(pargs); free(pptrs);
    return CUDA_ERROR_OUT_OF_MEMORY;
  }
  for (int i = 0; i < P; i++) {
    args[i].R = &m->ranks[i]; args[i].op = op; args[i].prec = prec; args[i].uplo = uplo; args[i].diag = diag;
    args[i].n = n; args[i].lda = lda; args[i].hostA = A;
    ptrs[i] = &args[i];
    pargs[i].R = &m->ranks[i]; pargs[i].prec = prec; pargs[i].n = n;
    pptrs[i] = &pargs[i];
  }
  /* phase 1: every rank's resident block columns (an allocation failure fails the call on all ranks before anyone
   * waits in a broadcast), phase 2: the sweeps */
  CUresult e = run_on_all(m, rank_potrf_prepare_task, pptrs, sizeof(*pptrs), sizeof(*pptrs));
  if (e == CUDA_SUCCESS) e = run_on_all(m, rank_inverse_task, ptrs, sizeof(*ptrs), sizeof(*ptrs));
  free(args); free(ptrs); free(pargs); free(pptrs);
  return e;
}

#define GEN_MGPU_POTRF(U, p, SCALAR)                                                                              \
  CUresult cuMultiGPU##U##potrf(CUmultiGPULAPACKhandle handle, CBlasUplo uplo, size_t n, SCALAR *A, size_t lda,   \
                                long *info) {                                                                     \
    *info = 0;                                                                                                    \
    if (lda < n) {                                                                                                \
      *info = -4;                                                                                                 \
      XERBLA(4);                                                                                                  \
      return CUDA_ERROR_INVALID_VALUE;                                                                            \
    }                                                                                                             \
    if (n == 0) return CUDA_SUCCESS;                                                                              \
    return mgpu_potrf(handle, p, uplo, n, A, lda, info);                                                          \
  }
GEN_MGPU_POTRF(S, 's', float)
GEN_MGPU_POTRF(D, 'd', double)
GEN_MGPU_POTRF(C, 'c', float complex)
GEN_MGPU_POTRF(Z, 'z', double complex)
