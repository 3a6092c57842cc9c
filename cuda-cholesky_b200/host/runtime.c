/*
 * runtime.c — error hooks and handle life-cycle.
 *   xerbla / errorHandler defaults : reference blas/xerbla.c:5-9, multigpu/error.c:4-11
 *   cuBLASCreate / cuBLASDestroy   : reference blas/handle.c:84-107
 *   cuLAPACKCreate / cuLAPACKDestroy : reference lapack/handle.c:77-88
 */
#include <stdio.h>
#include <stdlib.h>

#include "chol_internal.h"

static void default_xerbla(const char *function, long info) {
  fprintf(stderr, "On entry to %s parameter %ld had an invalid value\n", function, info);
}
xerbla_t xerbla = &default_xerbla;

static void default_error_handler(const char *call, const char *function, const char *file, int line, int error,
                                  const char *(*strerror_fn)(int)) {
  fprintf(stderr, "Error:\n\t%s returned %d (%s)\n\tin %s (%s:%d)\n", call, error,
          strerror_fn ? strerror_fn(error) : "?", function, file, line);
}
errorHandler_t errorHandler = &default_error_handler;

const char *cuCholGetErrorString(CUresult e) {
  const char *s = NULL;
  if (cuGetErrorString(e, &s) == CUDA_SUCCESS && s != NULL) return s;
  return "unknown CUDA driver error";
}

static const char *rt_strerror(int e) {
  (void)e;
  return "CUDA runtime error reported by a kernel launcher";
}

CUresult chol_rt_error(int rt, const char *call, const char *func, const char *file, int line) {
  if (errorHandler != NULL) errorHandler(call, func, file, line, rt, rt_strerror);
  return rt == 1 ? CUDA_ERROR_INVALID_VALUE : (rt == 2 ? CUDA_ERROR_OUT_OF_MEMORY : CUDA_ERROR_LAUNCH_FAILED);
}

/* ------------------------------------------------------------------------------------------ */
CUresult cuBLASCreate(CUBLAShandle *handle) {
  if (handle == NULL) return CUDA_ERROR_INVALID_VALUE;
  CU_ERROR_CHECK(cuInit(0));
  struct __cublashandle_st *h = calloc(1, sizeof *h);
  if (h == NULL) return CUDA_ERROR_OUT_OF_MEMORY;
  /* adopt the calling thread's current context (works with a framework's primary context);
   * if there is none, retain device 0's primary context (the reference creates one, handle.c:6-14) */
  CUresult e = cuCtxGetCurrent(&h->context);
  if (e == CUDA_SUCCESS && h->context == NULL) {
    e = cuDeviceGet(&h->device, 0);
    if (e == CUDA_SUCCESS) e = cuDevicePrimaryCtxRetain(&h->context, h->device);
    if (e == CUDA_SUCCESS) e = cuCtxSetCurrent(h->context);
    h->primary_retained = (e == CUDA_SUCCESS);
  } else if (e == CUDA_SUCCESS) {
    e = cuCtxGetDevice(&h->device);
  }
  if (e == CUDA_SUCCESS) e = cuMemAlloc(&h->dinfo, sizeof(long));
  if (e == CUDA_SUCCESS) e = cuEventCreate(&h->scratch_done, CU_EVENT_DISABLE_TIMING);
  if (e != CUDA_SUCCESS) {
    if (h->dinfo) cuMemFree(h->dinfo);
    if (h->primary_retained) cuDevicePrimaryCtxRelease(h->device);
    free(h);
    CU_ERROR_CHECK(e);
  }
  *handle = h;
  return CUDA_SUCCESS;
}

CUresult cuBLASDestroy(CUBLAShandle h) {
  if (h == NULL) return CUDA_ERROR_INVALID_VALUE;
  CU_ERROR_CHECK(cuCtxPushCurrent(h->context));
  if (h->scratch) cuMemFree(h->scratch);
  if (h->dinfo) cuMemFree(h->dinfo);
  if (h->scratch_done) cuEventDestroy(h->scratch_done);
  CUcontext dummy;
  CU_ERROR_CHECK(cuCtxPopCurrent(&dummy));
  if (h->primary_retained) cuDevicePrimaryCtxRelease(h->device);
  free(h);
  return CUDA_SUCCESS;
}

CUresult chol_blas_scratch(CUBLAShandle h, size_t bytes, CUdeviceptr *out) {
  if (bytes > h->scratch_bytes) {
    if (h->scratch) {
      CU_ERROR_CHECK(cuCtxSynchronize());
      CU_ERROR_CHECK(cuMemFree(h->scratch));
      h->scratch = 0;
      h->scratch_bytes = 0;
    }
    bytes = chol_roundup(bytes, (size_t)1 << 20);
    CU_ERROR_CHECK(cuMemAlloc(&h->scratch, bytes));
    h->scratch_bytes = bytes;
  }
  *out = h->scratch;
  return CUDA_SUCCESS;
}

CUresult chol_blas_scratch_acquire(CUBLAShandle h, CUstream stream) {
  if (h->scratch_in_use && h->scratch_stream != stream) CU_ERROR_CHECK(cuStreamWaitEvent(stream, h->scratch_done, 0));
  return CUDA_SUCCESS;
}

CUresult chol_blas_scratch_release(CUBLAShandle h, CUstream stream) {
  CU_ERROR_CHECK(cuEventRecord(h->scratch_done, stream));
  h->scratch_stream = stream;
  h->scratch_in_use = true;
  return CUDA_SUCCESS;
}

/* ------------------------------------------------------------------------------------------ */
CUresult cuLAPACKCreate(CULAPACKhandle *handle) {
  if (handle == NULL) return CUDA_ERROR_INVALID_VALUE;
  struct __culapackhandle_st *h = calloc(1, sizeof *h);
  if (h == NULL) return CUDA_ERROR_OUT_OF_MEMORY;
  CUresult e = cuBLASCreate(&h->blas_handle);
  if (e != CUDA_SUCCESS) {
    free(h);
    return e;
  }
  h->context = h->blas_handle->context;
  int lo = 0, hi = 0;
  e = cuCtxGetStreamPriorityRange(&lo, &hi); /* hi = numerically smallest = greatest priority */
  if (e == CUDA_SUCCESS) e = cuStreamCreateWithPriority(&h->s_panel, CU_STREAM_DEFAULT, hi);
  if (e == CUDA_SUCCESS) e = cuStreamCreateWithPriority(&h->s_update, CU_STREAM_DEFAULT, lo);
  if (e == CUDA_SUCCESS) e = cuStreamCreateWithPriority(&h->s_look, CU_STREAM_DEFAULT, hi);
  for (int i = 0; i < CHOL_EVENT_RING && e == CUDA_SUCCESS; i++) {
    e = cuEventCreate(&h->ev_panel[i], CU_EVENT_DISABLE_TIMING);
    if (e == CUDA_SUCCESS) e = cuEventCreate(&h->ev_rest[i], CU_EVENT_DISABLE_TIMING);
    if (e == CUDA_SUCCESS) e = cuEventCreate(&h->ev_look[i], CU_EVENT_DISABLE_TIMING);
  }
  for (int i = 0; i < CHOL_DIAG_EVENTS && e == CUDA_SUCCESS; i++) {
    e = cuEventCreate(&h->ev_dx[i], CU_EVENT_DISABLE_TIMING);
    if (e == CUDA_SUCCESS) e = cuEventCreate(&h->ev_dcol[i], CU_EVENT_DISABLE_TIMING);
  }
  if (e == CUDA_SUCCESS) e = cuEventCreate(&h->ev_solve, CU_EVENT_DISABLE_TIMING);
  if (e == CUDA_SUCCESS) e = cuMemAlloc(&h->dinfo, 16);
  if (e == CUDA_SUCCESS) e = cuMemsetD8(h->dinfo, 0, 16);
  if (e == CUDA_SUCCESS) e = cuMemAllocHost((void **)&h->hinfo, 16);
  if (e == CUDA_SUCCESS) e = cuMemAlloc(&h->logdet_scr, chol_logdet_scratch_bytes() + 64);
  if (e == CUDA_SUCCESS) e = cuMemsetD8(h->logdet_scr, 0, chol_logdet_scratch_bytes() + 64);
  if (e == CUDA_SUCCESS) e = cuMemAllocHost(&h->hlogdet, 16);
  if (e == CUDA_SUCCESS) e = cuEventCreate(&h->logdet_done, CU_EVENT_DISABLE_TIMING);
  if (e != CUDA_SUCCESS) {
    cuLAPACKDestroy(h);
    CU_ERROR_CHECK(e);
  }
  *handle = h;
  return CUDA_SUCCESS;
}

CUresult cuLAPACKDestroy(CULAPACKhandle h) {
  if (h == NULL) return CUDA_ERROR_INVALID_VALUE;
  if (h->context) {
    CU_ERROR_CHECK(cuCtxPushCurrent(h->context));
    if (h->s_panel) cuStreamDestroy(h->s_panel);
    if (h->s_update) cuStreamDestroy(h->s_update);
    if (h->s_look) cuStreamDestroy(h->s_look);
    for (int i = 0; i < CHOL_EVENT_RING; i++) {
      if (h->ev_panel[i]) cuEventDestroy(h->ev_panel[i]);
      if (h->ev_rest[i]) cuEventDestroy(h->ev_rest[i]);
      if (h->ev_look[i]) cuEventDestroy(h->ev_look[i]);
    }
    for (int i = 0; i < CHOL_DIAG_EVENTS; i++) {
      if (h->ev_dx[i]) cuEventDestroy(h->ev_dx[i]);
      if (h->ev_dcol[i]) cuEventDestroy(h->ev_dcol[i]);
    }
    if (h->ev_solve) cuEventDestroy(h->ev_solve);
    if (h->ws) cuMemFree(h->ws);
    if (h->dinfo) cuMemFree(h->dinfo);
    if (h->hinfo) cuMemFreeHost(h->hinfo);
    if (h->logdet_scr) cuMemFree(h->logdet_scr);
    if (h->hlogdet) cuMemFreeHost(h->hlogdet);
    if (h->logdet_done) cuEventDestroy(h->logdet_done);
    for (int i = 0; i < h->prof_cap; i++) {
      cuEventDestroy(h->prof_e0[i]);
      cuEventDestroy(h->prof_e1[i]);
    }
    free(h->prof_e0);
    free(h->prof_e1);
    free(h->prof_flops);
    free(h->prof_hist_ms);
    free(h->prof_hist_flops);
    CUcontext dummy;
    CU_ERROR_CHECK(cuCtxPopCurrent(&dummy));
  }
  if (h->blas_handle) cuBLASDestroy(h->blas_handle);
  free(h);
  return CUDA_SUCCESS;
}

CUresult chol_lapack_workspace(CULAPACKhandle h, size_t bytes, CUdeviceptr *out) {
  if (bytes > h->ws_bytes) {
    if (h->ws) {
      CU_ERROR_CHECK(cuCtxSynchronize());
      CU_ERROR_CHECK(cuMemFree(h->ws));
      h->ws = 0;
      h->ws_bytes = 0;
    }
    bytes = chol_roundup(bytes, (size_t)1 << 20);
    CU_ERROR_CHECK(cuMemAlloc(&h->ws, bytes));
    h->ws_bytes = bytes;
  }
  *out = h->ws;
  return CUDA_SUCCESS;
}

/* measurement hook: time every trailing-update launch of the next LAPACK calls with CUDA events */
CUresult cuLAPACKProfile(CULAPACKhandle h, int enable) {
  if (h == NULL) return CUDA_ERROR_INVALID_VALUE;
  if (enable && h->prof_cap == 0) {
    const int cap = 2048;
    if (h->prof_e0 == NULL) h->prof_e0 = calloc((size_t)cap, sizeof(CUevent));
    if (h->prof_e1 == NULL) h->prof_e1 = calloc((size_t)cap, sizeof(CUevent));
    if (h->prof_flops == NULL) h->prof_flops = calloc((size_t)cap, sizeof(double));
    if (h->prof_hist_ms == NULL) h->prof_hist_ms = calloc((size_t)cap, sizeof(float));
    if (h->prof_hist_flops == NULL) h->prof_hist_flops = calloc((size_t)cap, sizeof(double));
    if (!h->prof_e0 || !h->prof_e1 || !h->prof_flops || !h->prof_hist_ms || !h->prof_hist_flops)
      return CUDA_ERROR_OUT_OF_MEMORY; /* whatever was allocated is released by cuLAPACKDestroy */
    CU_ERROR_CHECK(cuCtxPushCurrent(h->context));
    CUresult e = CUDA_SUCCESS;
    int made = 0;
    for (; made < cap && e == CUDA_SUCCESS; made++) {
      e = cuEventCreate(&h->prof_e0[made], CU_EVENT_DEFAULT);
      if (e == CUDA_SUCCESS) e = cuEventCreate(&h->prof_e1[made], CU_EVENT_DEFAULT);
      if (e != CUDA_SUCCESS && h->prof_e0[made]) cuEventDestroy(h->prof_e0[made]);
      if (e != CUDA_SUCCESS) break;
    }
    if (e != CUDA_SUCCESS) { /* roll back: profiling stays off, nothing leaks */
      for (int i = 0; i < made; i++) {
        cuEventDestroy(h->prof_e0[i]);
        cuEventDestroy(h->prof_e1[i]);
      }
    }
    CUcontext dummy;
    CU_ERROR_CHECK(cuCtxPopCurrent(&dummy));
    CU_ERROR_CHECK(e);
    h->prof_cap = cap;
  }
  h->profile = enable;
  h->prof_total_ms = h->prof_total_flops = 0.0;
  h->prof_launches = 0;
  h->prof_hist_n = 0;
  return CUDA_SUCCESS;
}

/* per-launch history (ms, algorithmic flops) of the first min(cap, n) timed launches; returns n */
int cuLAPACKProfileHistory(CULAPACKhandle h, float *ms, double *flops, int cap) {
  if (h == NULL) return 0;
  for (int i = 0; i < h->prof_hist_n && i < cap; i++) {
    ms[i] = h->prof_hist_ms[i];
    flops[i] = h->prof_hist_flops[i];
  }
  return h->prof_hist_n;
}

/* totals since cuLAPACKProfile(h, 1): device milliseconds, algorithmic flops, launches */
CUresult cuLAPACKProfileRead(CULAPACKhandle h, double *ms, double *flops, int *launches) {
  if (h == NULL) return CUDA_ERROR_INVALID_VALUE;
  if (ms) *ms = h->prof_total_ms;
  if (flops) *flops = h->prof_total_flops;
  if (launches) *launches = h->prof_launches;
  return CUDA_SUCCESS;
}

/* tuning / test hook: outer block size of the LAPACK drivers (0 restores the default) */
CUresult cuLAPACKSetBlockSize(CULAPACKhandle h, size_t nb) {
  if (h == NULL || (nb != 0 && nb % 64 != 0)) return CUDA_ERROR_INVALID_VALUE;
  h->nb = nb;
  return CUDA_SUCCESS;
}
