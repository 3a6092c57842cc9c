/*
 * mock_nccl.c — TEST INFRASTRUCTURE ONLY: a libnccl.so.2 stand-in for the single-process,
 * one-thread-per-GPU form (ncclCommInitAll) used by the cuMultiGPU* drivers and for the one-rank-per-GPU
 * form (ncclCommInitRank; the ranks are threads of the test process), so that the
 * distributed factorisation's control flow (ownership, panel broadcast order, look-ahead, info
 * combination) can run on a machine without GPUs (tests/test_host_drivers_mock.py).  A broadcast is queued on
 * the caller's (mock) stream like a kernel; when it runs it is a rendezvous of the group's ranks: the
 * root publishes its buffer, the others copy from it.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { char internal[128]; } ncclUniqueId;
struct group { int n; pthread_barrier_t bar; const void *src; int refs; };
struct comm { struct group *g; int rank; };
typedef struct comm *ncclComm_t;

static int next_id = 1;
const char *ncclGetErrorString(int e) { return e == 0 ? "no error" : "mock NCCL error"; }
int ncclGetUniqueId(ncclUniqueId *id) {
  memset(id, 0, sizeof *id);
  const int k = __sync_fetch_and_add(&next_id, 1);
  memcpy(id->internal, &k, sizeof k);
  return 0;
}
int ncclGroupStart(void) { return 0; }
int ncclGroupEnd(void) { return 0; }

int ncclCommInitAll(ncclComm_t *comms, int n, const int *devs) {
  (void)devs;
  struct group *g = calloc(1, sizeof *g);
  if (g == NULL) return 1;
  g->n = n;
  g->refs = n;
  pthread_barrier_init(&g->bar, NULL, (unsigned)n);
  for (int i = 0; i < n; i++) {
    comms[i] = calloc(1, sizeof **comms);
    comms[i]->g = g;
    comms[i]->rank = i;
  }
  return 0;
}
/* ncclCommInitRank: in the tests the "processes" of a one-rank-per-GPU job are threads of the test process; the
 * n callers that present the same unique id meet here and share one group */
static pthread_mutex_t reg_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t reg_cv = PTHREAD_COND_INITIALIZER;
static struct { int key; struct group *g; int joined; } pending[16];

int ncclCommInitRank(ncclComm_t *c, int n, ncclUniqueId id, int rank) {
  int key;
  memcpy(&key, id.internal, sizeof key);
  if (key == 0 || n < 1 || rank < 0 || rank >= n) return 4;
  pthread_mutex_lock(&reg_mu);
  int slot = -1;
  for (int i = 0; i < 16; i++)
    if (pending[i].key == key) slot = i;
  if (slot < 0) {
    for (int i = 0; i < 16 && slot < 0; i++)
      if (pending[i].key == 0) slot = i;
    if (slot < 0) { pthread_mutex_unlock(&reg_mu); return 1; }
    struct group *g = calloc(1, sizeof *g);
    g->n = n;
    g->refs = n;
    pthread_barrier_init(&g->bar, NULL, (unsigned)n);
    pending[slot].key = key;
    pending[slot].g = g;
    pending[slot].joined = 0;
  }
  struct group *g = pending[slot].g;
  if (++pending[slot].joined == n) {
    pending[slot].key = 0; /* complete: the slot can be reused */
    pthread_cond_broadcast(&reg_cv);
  } else {
    while (pending[slot].g == g && pending[slot].key == key) pthread_cond_wait(&reg_cv, &reg_mu);
  }
  pthread_mutex_unlock(&reg_mu);
  *c = calloc(1, sizeof **c);
  (*c)->g = g;
  (*c)->rank = rank;
  return 0;
}
int ncclCommInitRankConfig(ncclComm_t *c, int n, ncclUniqueId id, int rank, void *cfg) { (void)cfg; return ncclCommInitRank(c, n, id, rank); }

int ncclCommDestroy(ncclComm_t c) {
  if (__sync_sub_and_fetch(&c->g->refs, 1) == 0) {
    pthread_barrier_destroy(&c->g->bar);
    free(c->g);
  }
  free(c);
  return 0;
}

void mock_stream_enqueue(void *stream, void (*fn)(void *), void *arg); /* mock_cuda.c, in the mock library */

struct bcast { const void *send; void *recv; size_t count; int root; struct comm *c; };
static void do_bcast(void *p) {
  struct bcast *b = p;
  struct group *g = b->c->g;
  if (b->c->rank == b->root) g->src = b->send;
  pthread_barrier_wait(&g->bar);
  if (b->c->rank != b->root && b->recv != g->src) memcpy(b->recv, g->src, b->count);
  pthread_barrier_wait(&g->bar);
  free(b);
}

int ncclBroadcast(const void *send, void *recv, size_t count, int dtype, int root, ncclComm_t c, void *stream) {
  if (dtype != 1 /* ncclUint8: the library broadcasts bytes */) return 4;
  struct bcast *b = malloc(sizeof *b);
  b->send = send; b->recv = recv; b->count = count; b->root = root; b->c = c;
  mock_stream_enqueue(stream, do_bcast, b);
  return 0;
}
