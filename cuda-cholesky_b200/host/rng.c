/*
 * rng.c — the cuRng64* entry points (include/rng.h; reference include/rng.h:326-413, which rng/curng64.c left
 * unimplemented): dSFMT-19937 on the device, the same stream as the reference's CPU generator.
 *
 * A request for n values starting at stream word `pos` covers the refills r_first .. r_last of the generator.
 * They are cut into segments of L = 2^level refills on an ABSOLUTE grid (segment j = refills jL .. jL + L - 1), so
 * the polynomials x^(jL) mod p that take the seed state to a segment's start do not depend on the seed or on the
 * request and are cached per process (host/rng_poly.c).  Per call:
 *     rows a_(jL) for the segments touched            -> device                (host cache; uploaded if not resident)
 *     start states  T^(jL) s0 = sum_i a_(jL)[i] K_i   dsfmt_jump_kernel        (K = the seed's first D states)
 *     one warp per segment generates its refills and stores the words inside the request   dsfmt_generate_kernel
 * L is chosen near sqrt(number of refills): the jump costs ~0.5 us per segment, a segment ~0.5 us per refill.
 */
#include <stdlib.h>
#include <string.h>

#include "chol_internal.h"
#include "rng.h"
#include "rng_internal.h"

struct __curng64_type_st {
  const char *name;
  int id;
};
static struct __curng64_type_st dsfmt_19937_type = {"Double precision SIMD-oriented Fast Mersenne Twister 2^19937 (device)", 1};
const CUrng64_t cu_dsfmt_19937_t = &dsfmt_19937_type;

#define RNG_MIN_LEVEL 4      /* segments of at least 16 refills */
#define RNG_LEVEL_BIAS 0     /* default shift of the segment level against "L^2 >= refills" (measured) */
#define RNG_MAX_SEGMENTS 4096

struct __curng64_st {
  CUcontext context;
  CUdevice device;
  bool primary_retained;
  CUstream stream;
  bool seeded;
  uint64_t s0[CHOL_DSFMT_WORDS][2];
  unsigned long long pos; /* 128-bit words of the stream consumed */
  int D;                  /* rows of K = degree of the polynomial K was built for */
  unsigned generation;
  CUdeviceptr K;
  size_t K_rows;
  CUdeviceptr states;
  size_t states_cap; /* segments */
  CUdeviceptr dpolys;
  size_t dpolys_cap; /* bytes */
  uint32_t *hpolys;  /* page-locked staging of the polynomial rows */
  size_t hpolys_cap;
  CUevent upload_done;
  bool upload_pending;
  int res_level; /* rows resident in dpolys: level, first index, count, generation */
  uint64_t res_j0;
  size_t res_count;
  unsigned res_generation;
  int level_bias;    /* tuning hook: added to the default segment level (cuRng64SetSegmentBias) */
  bool profile;      /* time the two kernels of every Get with events on the generator's stream */
  CUevent prof_ev[3];
  bool prof_valid;
  int last_level;
  size_t last_nseg;
};

#define RNG_ENTER(r)                                   \
  if ((r) == NULL) return CUDA_ERROR_INVALID_VALUE;    \
  CU_ERROR_CHECK(cuCtxPushCurrent((r)->context))
/* leave the context on every path: evaluates `expr` (a CUresult), pops, returns */
static CUresult rng_leave(CUresult e) {
  CUcontext dummy;
  CUresult p = cuCtxPopCurrent(&dummy);
  return e != CUDA_SUCCESS ? e : p;
}

CUresult cuRng64Create(CUrng64 *out, const CUrng64_t type) {
  if (out == NULL || type != cu_dsfmt_19937_t) return CUDA_ERROR_INVALID_VALUE;
  CU_ERROR_CHECK(cuInit(0));
  struct __curng64_st *r = calloc(1, sizeof *r);
  if (r == NULL) return CUDA_ERROR_OUT_OF_MEMORY;
  CUresult e = cuCtxGetCurrent(&r->context);
  if (e == CUDA_SUCCESS && r->context == NULL) {
    e = cuDeviceGet(&r->device, 0);
    if (e == CUDA_SUCCESS) e = cuDevicePrimaryCtxRetain(&r->context, r->device);
    if (e == CUDA_SUCCESS) e = cuCtxSetCurrent(r->context);
    r->primary_retained = (e == CUDA_SUCCESS);
  } else if (e == CUDA_SUCCESS) {
    e = cuCtxGetDevice(&r->device);
  }
  if (e == CUDA_SUCCESS) e = cuEventCreate(&r->upload_done, CU_EVENT_DISABLE_TIMING);
  if (e != CUDA_SUCCESS) {
    if (r->primary_retained) cuDevicePrimaryCtxRelease(r->device);
    free(r);
    CU_ERROR_CHECK(e);
  }
  r->res_level = -1;
  *out = r;
  return CUDA_SUCCESS;
}

CUresult cuRng64Destroy(CUrng64 r) {
  RNG_ENTER(r);
  cuStreamSynchronize(r->stream);
  if (r->K) cuMemFree(r->K);
  if (r->states) cuMemFree(r->states);
  if (r->dpolys) cuMemFree(r->dpolys);
  if (r->hpolys) cuMemFreeHost(r->hpolys);
  if (r->upload_done) cuEventDestroy(r->upload_done);
  for (int i = 0; i < 3; i++)
    if (r->prof_ev[i]) cuEventDestroy(r->prof_ev[i]);
  CUcontext dummy;
  CU_ERROR_CHECK(cuCtxPopCurrent(&dummy));
  if (r->primary_retained) cuDevicePrimaryCtxRelease(r->device);
  free(r);
  return CUDA_SUCCESS;
}

/* (re)build the Krylov block K_i = T^i s0, i < D, for the current polynomial: row 0 is uploaded, rows 1 .. D-1 are the
 * states after refills 0 .. D-2, written by one warp of the generation kernel in snapshot mode */
static CUresult rng_build_krylov(CUrng64 r) {
  int D;
  unsigned gen;
  if (chol_dsfmt_poly_ensure((const uint64_t(*)[2])r->s0, &D, &gen) != 0) return CUDA_ERROR_UNKNOWN;
  if ((size_t)D > r->K_rows) {
    if (r->K) {
      CU_ERROR_CHECK(cuStreamSynchronize(r->stream));
      CU_ERROR_CHECK(cuMemFree(r->K));
      r->K = 0;
      r->K_rows = 0;
    }
    CU_ERROR_CHECK(cuMemAlloc(&r->K, (size_t)D * CHOL_DSFMT_STATE_BYTES));
    r->K_rows = (size_t)D;
  }
  CU_ERROR_CHECK(cuMemcpyHtoDAsync(r->K, r->s0, CHOL_DSFMT_STATE_BYTES, r->stream));
  RT_CHECK(chol_dsfmt_generate_launch((const void *)r->K, 0ull, (unsigned)(D - 1), 1, 0ull, 0ull, NULL, 0ull, 0, (void *)r->K,
                                      (void *)r->stream));
  r->D = D;
  r->generation = gen;
  return CUDA_SUCCESS;
}

static CUresult rng_set(CUrng64 r, uint64_t seed) {
  chol_dsfmt_seed_state(seed, r->s0);
  r->pos = 0;
  r->seeded = false;
  CUresult e = rng_build_krylov(r);
  if (e == CUDA_SUCCESS) r->seeded = true;
  return e;
}

CUresult cuRng64Set(const CUrng64 r, uint64_t seed) {
  RNG_ENTER(r);
  return rng_leave(rng_set(r, seed));
}

static size_t words_of(size_t n) { return n / 2 + (n & 1); }

static CUresult rng_reserve(CUrng64 r, size_t nseg, size_t poly_u32) {
  if (nseg > r->states_cap) {
    if (r->states) {
      CU_ERROR_CHECK(cuStreamSynchronize(r->stream));
      CU_ERROR_CHECK(cuMemFree(r->states));
      r->states = 0;
      r->states_cap = 0;
    }
    const size_t cap = chol_roundup(nseg, 64);
    CU_ERROR_CHECK(cuMemAlloc(&r->states, cap * CHOL_DSFMT_STATE_BYTES));
    r->states_cap = cap;
  }
  const size_t bytes = nseg * poly_u32 * sizeof(uint32_t);
  if (bytes > r->dpolys_cap) {
    const size_t cap = chol_roundup(bytes, (size_t)1 << 20);
    if (r->dpolys) {
      CU_ERROR_CHECK(cuStreamSynchronize(r->stream));
      CU_ERROR_CHECK(cuMemFree(r->dpolys));
      r->dpolys = 0;
      r->dpolys_cap = 0;
      r->res_level = -1;
    }
    CU_ERROR_CHECK(cuMemAlloc(&r->dpolys, cap));
    r->dpolys_cap = cap;
    if (r->hpolys) {
      if (r->upload_pending) CU_ERROR_CHECK(cuEventSynchronize(r->upload_done));
      r->upload_pending = false;
      CU_ERROR_CHECK(cuMemFreeHost(r->hpolys));
      r->hpolys = NULL;
      r->hpolys_cap = 0;
    }
    CU_ERROR_CHECK(cuMemAllocHost((void **)&r->hpolys, cap));
    r->hpolys_cap = cap;
  }
  return CUDA_SUCCESS;
}

static CUresult rng_get(CUrng64 r, int kind, CUdeviceptr x, size_t n) {
  if (!r->seeded) return CUDA_ERROR_NOT_INITIALIZED; /* the CPU generator has no default seed either */
  if (n == 0) return CUDA_SUCCESS;
  if (x == 0 || (x & 7) != 0) return CUDA_ERROR_INVALID_VALUE;
  const unsigned long long w0 = r->pos, w1 = r->pos + words_of(n);
  const unsigned long long r_first = w0 / CHOL_DSFMT_N, r_last = (w1 - 1) / CHOL_DSFMT_N;
  const unsigned long long R = r_last - r_first + 1;
  /* segment length L = 2^level refills: the jump costs per SEGMENT, a segment is serial in its refills, so L ~ sqrt(R)
   * balances them (RNG_LEVEL_BIAS shifts the balance point by the measured cost ratio) */
  int level = RNG_MIN_LEVEL;
  while ((1ull << (2 * level)) < R) level++;
  level += RNG_LEVEL_BIAS + r->level_bias;
  if (level < RNG_MIN_LEVEL) level = RNG_MIN_LEVEL;
  if (level > 40) level = 40;
  while (((r_last >> level) - (r_first >> level) + 1) > RNG_MAX_SEGMENTS) level++;
  const uint64_t j0 = r_first >> level, j1 = r_last >> level;
  const size_t nseg = (size_t)(j1 - j0 + 1);
  for (int attempt = 0;; attempt++) {
    const size_t poly_u32 = 2 * (((size_t)r->D + 63) / 64);
    CU_ERROR_CHECK(rng_reserve(r, nseg, poly_u32));
    const bool resident = r->res_level == level && r->res_generation == r->generation && j0 >= r->res_j0 &&
                          j0 + nseg <= r->res_j0 + r->res_count;
    size_t row0 = 0;
    if (resident) {
      row0 = (size_t)(j0 - r->res_j0);
    } else {
      if (r->upload_pending) CU_ERROR_CHECK(cuEventSynchronize(r->upload_done)); /* hpolys is being read by a copy */
      r->upload_pending = false;
      if (chol_dsfmt_poly_rows(r->generation, level, j0, nseg, r->hpolys, poly_u32) != 0) {
        /* the process-wide polynomial was extended for another generator's seed: rebuild K for it and retry once */
        if (attempt > 0) return CUDA_ERROR_UNKNOWN;
        CU_ERROR_CHECK(rng_build_krylov(r));
        continue;
      }
      CU_ERROR_CHECK(cuMemcpyHtoDAsync(r->dpolys, r->hpolys, nseg * poly_u32 * sizeof(uint32_t), r->stream));
      CU_ERROR_CHECK(cuEventRecord(r->upload_done, r->stream));
      r->upload_pending = true;
      r->res_level = level;
      r->res_j0 = j0;
      r->res_count = nseg;
      r->res_generation = r->generation;
    }
    const CUdeviceptr rows = r->dpolys + row0 * poly_u32 * sizeof(uint32_t);
    if (r->profile) CU_ERROR_CHECK(cuEventRecord(r->prof_ev[0], r->stream));
    RT_CHECK(chol_dsfmt_jump_launch((const void *)r->K, (const void *)rows, (int)poly_u32, r->D, (void *)r->states, (int)nseg,
                                    (void *)r->stream));
    if (r->profile) CU_ERROR_CHECK(cuEventRecord(r->prof_ev[1], r->stream));
    RT_CHECK(chol_dsfmt_generate_launch((const void *)r->states, (unsigned long long)(j0 << level), 1u << level, (int)nseg, w0, w1,
                                        (void *)x, (unsigned long long)n, kind, NULL, (void *)r->stream));
    if (r->profile) {
      CU_ERROR_CHECK(cuEventRecord(r->prof_ev[2], r->stream));
      r->prof_valid = true;
    }
    r->last_level = level;
    r->last_nseg = nseg;
    break;
  }
  r->pos = w1;
  return CUDA_SUCCESS;
}

/* kinds as in csrc/rng_dsfmt.cu */
CUresult cuRng64Get(const CUrng64 r, CUdeviceptr x, size_t n) {
  RNG_ENTER(r);
  return rng_leave(rng_get(r, 0, x, n));
}
CUresult cuRng64GetCloseOpen(const CUrng64 r, CUdeviceptr x, size_t n) {
  RNG_ENTER(r);
  return rng_leave(rng_get(r, 1, x, n));
}
CUresult cuRng64GetOpenClose(const CUrng64 r, CUdeviceptr x, size_t n) {
  RNG_ENTER(r);
  return rng_leave(rng_get(r, 2, x, n));
}
CUresult cuRng64GetOpenOpen(const CUrng64 r, CUdeviceptr x, size_t n) {
  RNG_ENTER(r);
  return rng_leave(rng_get(r, 3, x, n));
}
CUresult cuRng64GetCloseClose(const CUrng64 r, CUdeviceptr x, size_t n) {
  RNG_ENTER(r);
  return rng_leave(rng_get(r, 4, x, n));
}

static CUresult rng_set_stream(CUrng64 r, CUstream s) {
  if (s != r->stream) CU_ERROR_CHECK(cuStreamSynchronize(r->stream));
  r->stream = s;
  return CUDA_SUCCESS;
}
CUresult cuRng64SetStream(CUrng64 r, CUstream s) {
  RNG_ENTER(r);
  return rng_leave(rng_set_stream(r, s));
}

CUresult cuRng64Discard(CUrng64 r, size_t n) {
  if (r == NULL) return CUDA_ERROR_INVALID_VALUE;
  if (!r->seeded) return CUDA_ERROR_NOT_INITIALIZED;
  r->pos += words_of(n);
  return CUDA_SUCCESS;
}

unsigned long long cuRng64Position(CUrng64 r) { return r == NULL ? 0ull : r->pos; }

CUresult cuRng64SetSegmentBias(CUrng64 r, int bias) {
  if (r == NULL || bias < -8 || bias > 8) return CUDA_ERROR_INVALID_VALUE;
  r->level_bias = bias;
  return CUDA_SUCCESS;
}

static CUresult rng_profile(CUrng64 r, int enable) {
  if (enable && r->prof_ev[0] == NULL)
    for (int i = 0; i < 3; i++) CU_ERROR_CHECK(cuEventCreate(&r->prof_ev[i], CU_EVENT_DEFAULT));
  r->profile = enable != 0;
  r->prof_valid = false;
  return CUDA_SUCCESS;
}
CUresult cuRng64Profile(CUrng64 r, int enable) {
  RNG_ENTER(r);
  return rng_leave(rng_profile(r, enable));
}

static CUresult rng_profile_read(CUrng64 r, float *jump_ms, float *generate_ms, int *segments, int *segment_refills) {
  if (!r->prof_valid) return CUDA_ERROR_NOT_READY;
  CU_ERROR_CHECK(cuEventSynchronize(r->prof_ev[2]));
  if (jump_ms) CU_ERROR_CHECK(cuEventElapsedTime(jump_ms, r->prof_ev[0], r->prof_ev[1]));
  if (generate_ms) CU_ERROR_CHECK(cuEventElapsedTime(generate_ms, r->prof_ev[1], r->prof_ev[2]));
  if (segments) *segments = (int)r->last_nseg;
  if (segment_refills) *segment_refills = 1 << r->last_level;
  return CUDA_SUCCESS;
}
CUresult cuRng64ProfileRead(CUrng64 r, float *jump_ms, float *generate_ms, int *segments, int *segment_refills) {
  RNG_ENTER(r);
  return rng_leave(rng_profile_read(r, jump_ms, generate_ms, segments, segment_refills));
}
