/*
 * chol_internal.h — private declarations shared by the C host code (handles, workspaces, the
 * precision-generic drivers).  Host code is C (north_star: "Host code stays in C and calls CUDA
 * through a thin C-ABI layer"); the CUDA side is reached only through csrc/chol_kernels.h.
 */
#ifndef CHOL_INTERNAL_H
#define CHOL_INTERNAL_H

#include <stdbool.h>
#include <stddef.h>
#include <cuda.h>

#include "blas.h"
#include "lapack.h"
#include "error.h"
#include "cholesky_b200.h"
#include "../csrc/chol_kernels.h"

/* reference blas/handle.h:4-11 keeps a context and 16 lazily loaded CUmodules; kernels here are
 * linked in, so the handle keeps the context and a growable device scratch instead. */
struct __cublashandle_st {
  CUcontext context;
  CUdevice device;
  bool primary_retained; /* cuBLASCreate had to retain a primary context (no current context) */
  CUdeviceptr scratch;   /* TRSM/TRMM temporaries */
  size_t scratch_bytes;
  CUdeviceptr dinfo;     /* device long, sink for the diagonal-inverse kernels used by TRSM */
  /* scratch and dinfo are shared by every TRSM/TRMM call on this handle: calls on different
   * streams are ordered against each other through this event (same stream: stream order) */
  CUevent scratch_done;
  CUstream scratch_stream;
  bool scratch_in_use;
};

#define CHOL_EVENT_RING 4
#define CHOL_DIAG_EVENTS 16 /* sub-blocks of one diagonal block whose completion the pipelined panel solve can wait for */

/* reference lapack/handle.h:4-11 */
struct __culapackhandle_st {
  CUBLAShandle blas_handle;
  CUcontext context;
  CUstream s_panel;  /* high priority: diagonal block, panel solve, look-ahead update */
  CUstream s_update; /* trailing-matrix update */
  CUstream s_look;   /* second high-priority stream: full-height GEMMs of the pipelined panel solve; lower part of the
                      * look-ahead column when the look-ahead is split (CHOL_LOOKAHEAD_SPLIT) */
  CUevent ev_panel[CHOL_EVENT_RING];
  CUevent ev_rest[CHOL_EVENT_RING];
  CUevent ev_look[CHOL_EVENT_RING];
  CUevent ev_dx[CHOL_DIAG_EVENTS], ev_dcol[CHOL_DIAG_EVENTS], ev_solve; /* pipelined panel solve (lapack.c panel_factor_solve) */
  CUdeviceptr ws; /* workspace, cached across calls (the reference allocates per call, dpotrf.c:200) */
  size_t ws_bytes;
  CUdeviceptr dinfo;      /* device: long info, then double = the log-determinant accumulated by the diagonal-block kernels */
  long *hinfo;            /* pinned host mirror (16 bytes) */
  double last_logdet;     /* of the last POTRF on this handle (cuLAPACKPotrfLogdet) */
  bool last_logdet_valid;
  CUdeviceptr logdet_scr; /* reduction scratch + result; shared by every logdet call on the handle: calls on
                           * different streams are ordered through logdet_done (same stream: stream order) */
  CUevent logdet_done;
  CUstream logdet_stream;
  bool logdet_in_use;
  void *hlogdet;          /* pinned */
  size_t nb;              /* outer block size (0 = default per precision) */
  /* optional per-launch timing of the dominant kernel (the trailing-matrix update): CUDA events
   * recorded on the stream the kernel is launched on; read back by bench.py for the roofline */
  int profile;
  int prof_n, prof_cap;
  CUevent *prof_e0, *prof_e1;
  double *prof_flops;
  double prof_total_ms, prof_total_flops;
  int prof_launches;
  float *prof_hist_ms;    /* per-launch history since cuLAPACKProfile(h, 1) (capacity prof_cap) */
  double *prof_hist_flops;
  int prof_hist_n;
};

static inline size_t chol_esize(int prec) { return prec == 's' ? 4 : (prec == 'z' ? 16 : 8); }
static inline bool chol_is_complex(int prec) { return prec == 'c' || prec == 'z'; }
static inline size_t chol_min(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t chol_roundup(size_t a, size_t b) { return (a + b - 1) / b * b; }
/* address of element (i, j) of a column-major device matrix */
static inline CUdeviceptr chol_at(CUdeviceptr A, size_t lda, size_t i, size_t j, size_t es) {
  return A + (CUdeviceptr)((j * lda + i) * es);
}

/* cudaError_t (as int, from the launchers) -> CUresult, through the error hook */
CUresult chol_rt_error(int cuda_rt_error, const char *call, const char *func, const char *file, int line);
#define RT_CHECK(call)                                                                  \
  do {                                                                                  \
    int __rt__ = (call);                                                                \
    if (__rt__ != 0) return chol_rt_error(__rt__, STRING(call), __func__, __FILE__, __LINE__); \
  } while (false)

CUresult chol_blas_scratch(CUBLAShandle h, size_t bytes, CUdeviceptr *out);
/* bracket a call that uses the handle's scratch on `stream` */
CUresult chol_blas_scratch_acquire(CUBLAShandle h, CUstream stream);
CUresult chol_blas_scratch_release(CUBLAShandle h, CUstream stream);
CUresult chol_lapack_workspace(CULAPACKhandle h, size_t bytes, CUdeviceptr *out);

/* ---- precision-generic BLAS-3 cores (blas3.c).  alpha/beta as (re, im) ---- */
CUresult chol_gemm(int prec, int opA, int opB, size_t m, size_t n, size_t k, const double alpha[2], CUdeviceptr A,
                   size_t lda, CUdeviceptr B, size_t ldb, const double beta[2], CUdeviceptr C, size_t ldc,
                   CUdeviceptr D, size_t ldd, CUstream stream);
/* general descriptor form: triangular masks, triangle-only output, Hermitian diagonal */
CUresult chol_gemm_ex(int prec, int opA, int opB, int triA, int unitA, int triB, int unitB, int cuplo, int herm,
                      size_t m, size_t n, size_t k, const double alpha[2], CUdeviceptr A, size_t lda,
                      CUdeviceptr B, size_t ldb, const double beta[2], CUdeviceptr C, size_t ldc, CUdeviceptr D,
                      size_t ldd, CUstream stream);
CUresult chol_rankk(int prec, int uplo, int trans, size_t n, size_t k, double alpha, CUdeviceptr A, size_t lda,
                    double beta, CUdeviceptr C, size_t ldc, CUstream stream);
CUresult chol_trmm2(int prec, int side, int uplo, int trans, int diag, size_t m, size_t n, const double alpha[2],
                    CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUdeviceptr X, size_t ldx, CUstream stream);
CUresult chol_trsm(CUBLAShandle h, int prec, int side, int uplo, int trans, int diag, size_t m, size_t n,
                   const double alpha[2], CUdeviceptr A, size_t lda, CUdeviceptr B, size_t ldb, CUstream stream);

/* workspace layout of one LAPACK call */
typedef struct chol_ws {
  size_t nb, db, ldw, es;
  CUdeviceptr W[2]; /* panel buffers, ldw x nb */
  CUdeviceptr Xd;   /* nb x nb : inverse of a diagonal block (TRTRI), product block (LAUUM) */
  CUdeviceptr T1;   /* nb x nb : temporary of the blocked small inverse */
  CUdeviceptr logacc; /* device double the diagonal-block kernels add 2 sum log L(c,c) to (0 = none) */
  size_t kp;        /* s, c with the tensor-core update: K pitch of the TF32 planes (0 = FFMA path) */
  CUdeviceptr Whi[2], Wlo[2]; /* ldw x kp floats (c: 2 ldw x kp, real embedding), K-major: tf32 hi / lo planes of W[0], W[1] */
} chol_ws;

size_t chol_default_nb(int prec, size_t n);
CUresult chol_carve(CULAPACKhandle h, int prec, size_t n, chol_ws *w);
CUresult chol_carve_nb(CULAPACKhandle h, int prec, size_t n, size_t nb, chol_ws *w);
/* Cholesky of a jb x jb (jb <= nb) diagonal block (potf2 + substitution + rank-db updates, no inverse);
 * ev_fact[c] / ev_col[c] (optional): diagonal sub-block c factored / sub-block column (row) c final */
CUresult chol_potrf_diag(int prec, int uplo, size_t jb, CUdeviceptr A, size_t lda, long *dinfo, long base,
                         const chol_ws *w, CUstream st);
CUresult chol_diag_factor_ev(int prec, int uplo, size_t jb, CUdeviceptr A, size_t lda, long *dinfo, long base,
                             const chol_ws *w, CUstream st, CUevent *ev_fact, CUevent *ev_col);
/* POTRF panel step: factor the jb x jb diagonal block at Ajj, solve the panel next to it into W (rest x jb).
 * h != NULL: the solve is pipelined with the diagonal chain on the handle's second high-priority stream */
CUresult chol_panel_factor_solve_lower(CULAPACKhandle h, int prec, size_t jb, size_t rest, CUdeviceptr Ajj, CUdeviceptr P, size_t lda,
                                       CUdeviceptr W, size_t ldw, long *dinfo, long base, const chol_ws *w, CUstream st,
                                       CUevent before_solve);
bool chol_panel_pipeline(void);
bool chol_lookahead_split(void);
bool chol_s_tensor(void);
bool chol_trtri_lookahead(void);
/* ... upper: the jb x rest block row P right of the diagonal block, W = P^H * inv(U11) (rest x jb) */
CUresult chol_panel_factor_solve_upper(CULAPACKhandle h, int prec, size_t jb, size_t rest, CUdeviceptr Ajj, CUdeviceptr P, size_t lda,
                                       CUdeviceptr W, size_t ldw, long *dinfo, long base, const chol_ws *w, CUstream st,
                                       CUevent before_solve);
CUresult chol_trtri_small(int prec, int uplo, int diag, size_t jb, CUdeviceptr A, size_t lda, CUdeviceptr X,
                          size_t ldx, long *dinfo, long base, const chol_ws *w, CUstream st);

/* ---- precision-generic LAPACK cores (lapack.c) ---- */
CUresult chol_potrf(CULAPACKhandle h, int prec, int uplo, size_t n, CUdeviceptr A, size_t lda, long *info);
CUresult chol_trtri(CULAPACKhandle h, int prec, int uplo, int diag, size_t n, CUdeviceptr A, size_t lda, long *info);
CUresult chol_lauum(CULAPACKhandle h, int prec, int uplo, size_t n, CUdeviceptr A, size_t lda, long *info);
CUresult chol_potri(CULAPACKhandle h, int prec, int uplo, size_t n, CUdeviceptr A, size_t lda, long *info);
CUresult chol_potrs(CULAPACKhandle h, int prec, int uplo, size_t n, size_t nrhs, CUdeviceptr A, size_t lda, CUdeviceptr B,
                    size_t ldb);
CUresult chol_logdet(CULAPACKhandle h, int prec, CUdeviceptr x, size_t incx, size_t n, void *result, CUstream stream);

#endif
