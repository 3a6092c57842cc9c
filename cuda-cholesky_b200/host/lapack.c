/*
 * lapack.c — POTRF / TRTRI / LAUUM / POTRI / logdet on device pointers, all four precisions.
 *
 * What the reference does (lapack/dpotrf.c:172-296, dtrtri.c:326-468, dlauum.c:154-292): a
 * left-looking loop whose diagonal block makes a D2H -> CPU -> H2D round trip every step.
 * What this file does instead (B200-first):
 *   POTRF  right-looking, block size nb (256 ... 1024 by precision and n, chol_default_nb): per step
 *            [panel stream, high priority]  diagonal block sub-block by sub-block: potf2 in registers -> the rows
 *            below it by substitution (trsm_rows, rows spread over the SMs) -> rank-128 update; no inverse
 *            [second high-priority stream]  panel solve into a contiguous buffer W, sub-block by sub-block
 *            (full-height tensor-core GEMM + substitution), pipelined with the diagonal chain
 *            [update stream]                trailing-matrix HERK/SYRK  A22 -= W*W^H  on tensor-core tiles, then the
 *            write-back of W into the matrix; the next block column is updated first (look-ahead, split in
 *            its diagonal block and the rows below)
 *          The panel chain of step j+1 overlaps the trailing update of step j (depth-1
 *          look-ahead); nothing runs on the CPU and nothing crosses PCIe.
 *          Upper is the same algorithm on W = U12^H, so the trailing update is the identical
 *          TMA-fed "NT" kernel for both triangles.
 *   TRTRI  block columns (the reference's sweep order, dtrtri.c:119-163), each step =
 *          small out-of-place inverse + two GEMMs (the "two-TRMM" form of lapack/strtri.c:408-424)
 *   LAUUM  per block row one masked GEMM + one masked HERK over the trapezoid L[i:,i] / U[i,i:]
 *   POTRI  = TRTRI(NonUnit) then LAUUM (lapack/dpotri.c:19-22); input is the Cholesky factor
 *   POTRS  (addition) two blocked TRSMs on the factor
 *   logdet one single-pass reduction kernel (lapack/dlogdet.c:37-74)
 * info is accumulated on the device (csrc/panel_kernels.cu) and read once at the end.
 * The stream schedule has A/B switches (CHOL_PANEL_PIPELINE, CHOL_LOOKAHEAD_SPLIT, CHOL_TRTRI_LOOKAHEAD, all on by
 * default and measured on hardware, profiles/r02_*); cu*potrf also leaves log det A for cuLAPACKPotrfLogdet.
 */
#include <complex.h>
#include <stdlib.h>

#include "chol_internal.h"

static const double ONE[2] = {1.0, 0.0};
static const double ZERO[2] = {0.0, 0.0};
static const double MONE[2] = {-1.0, 0.0};

/* an environment switch, read once per process; the cached value is published with atomics because the
 * multi-GPU worker threads ask concurrently (every reader would compute the same value anyway) */
static bool chol_env_switch(int *state, const char *name, int dflt) {
  int v = __atomic_load_n(state, __ATOMIC_RELAXED);
  if (v < 0) {
    const char *e = getenv(name);
    v = (e != NULL) ? (atoi(e) != 0) : dflt;
    __atomic_store_n(state, v, __ATOMIC_RELAXED);
  }
  return v != 0;
}
static int chol_lookahead_split_state = -1, chol_trtri_lookahead_state = -1, chol_s_tensor_state = -1, chol_panel_pipeline_state = -1;
/* s- and c-precision POTRF: trailing and look-ahead updates on the tcgen05 tensor cores as a 3xTF32 split (default;
 * CHOL_S_TENSOR=0 keeps them on the FP32 FFMA engine for A/B runs and for the plain-FP32 tolerance) */
bool chol_s_tensor(void) { return chol_env_switch(&chol_s_tensor_state, "CHOL_S_TENSOR", 1); }

/* outer block size: k of the trailing update.  Larger k amortises the C-tile epilogue (measured on
 * B200, d: 512 beats 256 from N ~ 32768 up), smaller k shortens the serial panel chain. */
size_t chol_default_nb(int prec, size_t n) {
  /* measured on B200 with the substitution-based panel chain (scripts/nb_sweep.py, profiles/r02_call10_summary.txt):
   * d  N = 8192: nb 256 / 512 -> 18.7 / 20.1 TFLOP/s;  16384: 512 / 768 / 1024 -> 28.9 / 29.3 / 29.3;  32768: 512 / 1024 -> 32.1 / 33.0
   * z  N = 16384: 256 / 512 -> 29.8 / 30.0;  s 16384, 32768: 512 best;  c 16384: 256 / 512 -> 118 / 105 */
  if (prec == 'z') return n >= 12288 ? 512 : (n >= 8192 ? 256 : 128);
  if (prec == 'd') return n >= 24576 ? 1024 : (n >= 12288 ? 768 : (n >= 6144 ? 512 : 256));
  if (prec == 'c') return n >= 24576 ? 512 : 256; /* N = 32768: 512 / 256 -> 150 / 146 */
  return n >= 12288 ? 512 : 256;
}

CUresult chol_carve(CULAPACKhandle h, int prec, size_t n, chol_ws *w) {
  return chol_carve_nb(h, prec, n, h->nb ? h->nb : chol_default_nb(prec, n), w);
}

/* carve the handle's cached workspace for an n x n problem blocked by nb */
CUresult chol_carve_nb(CULAPACKhandle h, int prec, size_t n, size_t nb, chol_ws *w) {
  w->es = chol_esize(prec);
  w->nb = nb;
  w->logacc = h->dinfo + sizeof(long);
  w->db = (size_t)chol_diag_max(prec);
  if (w->nb < w->db) w->db = w->nb;
  w->ldw = chol_roundup(n > w->nb ? n : w->nb, 16);
  const size_t szW = chol_roundup(w->ldw * w->nb * w->es, 256), szX = chol_roundup(w->nb * w->nb * w->es, 256);
  /* s, c: TF32 (hi, lo) planes of the two panel buffers for the tensor-core trailing update (csrc/gemm_tf32x3.cu);
   * c uses the real embedding of the complex product: two plane rows per matrix row, 2 nb columns */
  const size_t cx = prec == 'c' ? 2 : 1;
  w->kp = ((prec == 's' || prec == 'c') && chol_s_tensor()) ? chol_tf32_kpitch(cx * w->nb) : 0;
  const size_t szS = chol_roundup(cx * w->ldw * w->kp * sizeof(float), 1024);
  CUdeviceptr base;
  CU_ERROR_CHECK(chol_lapack_workspace(h, 2 * szW + 2 * szX + 4 * szS + 2048, &base));
  w->W[0] = base;
  w->W[1] = base + szW;
  w->Xd = base + 2 * szW;
  w->T1 = w->Xd + szX;
  const CUdeviceptr planes = chol_roundup(w->T1 + szX, 1024);
  for (int i = 0; i < 2; i++) {
    w->Whi[i] = w->kp ? planes + (CUdeviceptr)(2 * i) * szS : 0;
    w->Wlo[i] = w->kp ? planes + (CUdeviceptr)(2 * i + 1) * szS : 0;
  }
  return CUDA_SUCCESS;
}

/* `producer` = the stream on which the info-writing kernels ran */
static CUresult finish_info(CULAPACKhandle h, CUstream producer, long *info) {
  CU_ERROR_CHECK(cuMemcpyDtoHAsync(h->hinfo, h->dinfo, 16, producer)); /* info and the fused log-determinant */
  CU_ERROR_CHECK(cuStreamSynchronize(h->s_panel));
  CU_ERROR_CHECK(cuStreamSynchronize(h->s_update));
  CU_ERROR_CHECK(cuStreamSynchronize(h->s_look));
  *info = *h->hinfo;
  return CUDA_SUCCESS;
}

/* Cholesky of a jb x jb (jb <= nb) diagonal block on one stream, blocked by db (128; z 64): per sub-block
 *   potf2 (one CTA, registers)  ->  the rows below / columns right of it solved by SUBSTITUTION, in place
 *   (chol_trsm_rows_launch, rows spread over the SMs)  ->  rank-db update of the rest of the block.
 * No inverse is formed (rounds 1-2 inverted every sub-block and multiplied: an 80 us fused kernel per 128 columns
 * on the serial chain, and explicit inverses lose digits on ill-conditioned blocks that substitution keeps).
 * ev_fact / ev_col (both or neither; at least nblk events each): ev_fact[c] is recorded on st when diagonal
 * sub-block c is factored, ev_col[c] when the factor's sub-block column (lower) / row (upper) c is final -
 * what a panel solve running on another stream waits for. */
CUresult chol_diag_factor_ev(int prec, int uplo, size_t jb, CUdeviceptr A, size_t lda, long *dinfo, long base,
                             const chol_ws *w, CUstream st, CUevent *ev_fact, CUevent *ev_col) {
  const size_t es = w->es, db = w->db, nblk = (jb + db - 1) / db;
  const int herm = chol_is_complex(prec);
  for (size_t c = 0; c < nblk; c++) {
    const size_t i = c * db, ib = chol_min(db, jb - i), r = jb - i - ib;
    const CUdeviceptr Aii = chol_at(A, lda, i, i, es);
    RT_CHECK(chol_potf2_logdet_launch(prec, uplo, ib, (void *)Aii, lda, dinfo, base + (long)i, (double *)w->logacc, (void *)st));
    if (ev_fact) CU_ERROR_CHECK(cuEventRecord(ev_fact[c], st));
    if (r == 0) break;
    const CUdeviceptr A22 = chol_at(A, lda, i + ib, i + ib, es);
    if (uplo == 'L') { /* L21 = A21 * L11^-H, then A22 -= L21 * L21^H */
      const CUdeviceptr A21 = chol_at(A, lda, i + ib, i, es);
      RT_CHECK(chol_trsm_rows_launch(prec, 'L', ib, (const void *)Aii, lda, r, 0, (const void *)A21, lda, 0, (void *)A21, lda,
                                     dinfo, (void *)st));
      if (ev_col) CU_ERROR_CHECK(cuEventRecord(ev_col[c], st));
      CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'C', 0, 0, 0, 0, 'L', herm, r, r, ib, MONE, A21, lda, A21, lda, ONE, A22, lda, A22,
                                  lda, st));
    } else { /* U12 = U11^-H * A12, i.e. U12^H = A12^H * U11^-1 (rows = columns of A12), then A22 -= U12^H * U12 */
      const CUdeviceptr A12 = chol_at(A, lda, i, i + ib, es);
      RT_CHECK(chol_trsm_rows_launch(prec, 'U', ib, (const void *)Aii, lda, r, 1, (const void *)A12, lda, 1, (void *)A12, lda,
                                     dinfo, (void *)st));
      if (ev_col) CU_ERROR_CHECK(cuEventRecord(ev_col[c], st));
      CU_ERROR_CHECK(chol_gemm_ex(prec, 'C', 'N', 0, 0, 0, 0, 'U', herm, r, r, ib, MONE, A12, lda, A12, lda, ONE, A22, lda, A22,
                                  lda, st));
    }
  }
  return CUDA_SUCCESS;
}

CUresult chol_potrf_diag(int prec, int uplo, size_t jb, CUdeviceptr A, size_t lda, long *dinfo, long base,
                         const chol_ws *w, CUstream st) {
  return chol_diag_factor_ev(prec, uplo, jb, A, lda, dinfo, base, w, st, NULL, NULL);
}

/* X (jb x jb, ldx, other triangle zero) = inverse of the triangular jb x jb block A, jb <= nb */
CUresult chol_trtri_small(int prec, int uplo, int diag, size_t jb, CUdeviceptr A, size_t lda, CUdeviceptr X,
                          size_t ldx, long *dinfo, long base, const chol_ws *w, CUstream st) {
  const size_t es = w->es, db = w->db, nblk = (jb + db - 1) / db;
  if (nblk > 1) RT_CHECK(chol_fill_launch(prec, 0, jb, jb, 0.0, 0.0, (void *)X, ldx, (void *)st));
  for (size_t b = 0; b < nblk; b++) {
    const size_t c = (uplo == 'L') ? nblk - 1 - b : b; /* the reference's sweep direction */
    const size_t i = c * db, ib = chol_min(db, jb - i);
    const CUdeviceptr Xcc = chol_at(X, ldx, i, i, es);
    RT_CHECK(chol_trti2_launch(prec, uplo, diag, ib, (const void *)chol_at(A, lda, i, i, es), lda, (void *)Xcc, ldx, 1,
                               dinfo, base + (long)i, (void *)st));
    if (uplo == 'L') {
      const size_t r = jb - i - ib;
      if (r == 0) continue;
      /* X21 = -X22 * (L21 * X11) */
      CU_ERROR_CHECK(chol_gemm(prec, 'N', 'N', r, ib, ib, ONE, chol_at(A, lda, i + ib, i, es), lda, Xcc, ldx, ZERO, 0, 0,
                               w->T1, w->nb, st));
      CU_ERROR_CHECK(chol_gemm(prec, 'N', 'N', r, ib, r, MONE, chol_at(X, ldx, i + ib, i + ib, es), ldx, w->T1, w->nb,
                               ZERO, 0, 0, chol_at(X, ldx, i + ib, i, es), ldx, st));
    } else {
      if (i == 0) continue;
      /* X12 = -X11 * (U12 * X22) for the block column above the diagonal block */
      CU_ERROR_CHECK(chol_gemm(prec, 'N', 'N', i, ib, ib, ONE, chol_at(A, lda, 0, i, es), lda, Xcc, ldx, ZERO, 0, 0,
                               w->T1, w->nb, st));
      CU_ERROR_CHECK(chol_gemm(prec, 'N', 'N', i, ib, i, MONE, X, ldx, w->T1, w->nb, ZERO, 0, 0,
                               chol_at(X, ldx, 0, i, es), ldx, st));
    }
  }
  return CUDA_SUCCESS;
}

/* CHOL_LOOKAHEAD_SPLIT=1: the look-ahead update of the next block column is split - its jb2 x jb2 diagonal block
 * stays on the panel stream, so the next diagonal block can be factored at once, while the rows below it (the
 * bulk of the column) are updated on a third stream at the same time; the next panel solve waits for that part.
 * Measured neutral on one GPU (dpotrf N = 16384 28.30 -> 28.41 TFLOP/s, profiles/r02_call2_*.txt), so off. */
bool chol_lookahead_split(void) { return chol_env_switch(&chol_lookahead_split_state, "CHOL_LOOKAHEAD_SPLIT", 1); }

bool chol_trtri_lookahead(void) { return chol_env_switch(&chol_trtri_lookahead_state, "CHOL_TRTRI_LOOKAHEAD", 1); }

/* Panel step of POTRF: factor the jb x jb diagonal block at Ajj and solve the panel next to it into W (rest x jb).
 *   lower: P = the rest x jb block below Ajj,      W_c = (P_c   - W[:, 0:c] * L[c, 0:c]^H) * L_cc^-H
 *   upper: P = the jb x rest block right of Ajj,   W_c = (P_c^H - W[:, 0:c] * U[0:c, c])   * U_cc^-1
 * by block forward substitution over the db-wide sub-blocks c: one full-height GEMM (tensor cores) and one
 * substitution with the db x db triangle (chol_trsm_rows_launch) per sub-block; nothing is inverted.
 * With a handle the two halves are PIPELINED: the diagonal chain (chol_diag_factor_ev) runs on `st`, the
 * full-height work on the handle's second high-priority stream, sub-block c of the solve starting as soon as
 * sub-block c of the chain is done (events), so only the last sub-block's solve is left on the critical path.
 * `st` has seen everything when the function returns.  before_solve (may be NULL): an event that must have
 * completed before P is read - the lower part of a split look-ahead runs on another stream while the diagonal
 * block is factored here.  CHOL_PANEL_PIPELINE=0 keeps everything on `st` (A/B runs). */
bool chol_panel_pipeline(void) { return chol_env_switch(&chol_panel_pipeline_state, "CHOL_PANEL_PIPELINE", 1); }

static CUresult panel_factor_solve(CULAPACKhandle h, int prec, int uplo, size_t jb, size_t rest, CUdeviceptr Ajj, CUdeviceptr P,
                                   size_t lda, CUdeviceptr W, size_t ldw, long *dinfo, long base, const chol_ws *w, CUstream st,
                                   CUevent before_solve) {
  const size_t es = w->es, db = w->db, nblk = (jb + db - 1) / db;
  const bool lower = uplo == 'L';
  const bool pipe = h != NULL && nblk <= CHOL_DIAG_EVENTS && chol_panel_pipeline();
  CUstream ss = pipe ? h->s_look : st;
  if (pipe) { /* the solve stream starts from what `st` has seen (P's look-ahead update may have run there) */
    CU_ERROR_CHECK(cuEventRecord(h->ev_solve, st));
    CU_ERROR_CHECK(cuStreamWaitEvent(ss, h->ev_solve, 0));
    if (before_solve) CU_ERROR_CHECK(cuStreamWaitEvent(ss, before_solve, 0));
  }
  CU_ERROR_CHECK(chol_diag_factor_ev(prec, uplo, jb, Ajj, lda, dinfo, base, w, st, pipe ? h->ev_dx : NULL, pipe ? h->ev_dcol : NULL));
  if (!pipe && before_solve) CU_ERROR_CHECK(cuStreamWaitEvent(st, before_solve, 0));
  for (size_t c = 0; c < nblk; c++) {
    const size_t i = c * db, ib = chol_min(db, jb - i);
    const CUdeviceptr Wc = chol_at(W, ldw, 0, i, es), Lcc = chol_at(Ajj, lda, i, i, es);
    const CUdeviceptr Pc = lower ? chol_at(P, lda, 0, i, es) : chol_at(P, lda, i, 0, es);
    if (c == 0) { /* straight from the matrix (upper: read as its conjugate transpose) */
      if (pipe) CU_ERROR_CHECK(cuStreamWaitEvent(ss, h->ev_dx[0], 0));
      RT_CHECK(chol_trsm_rows_launch(prec, uplo, ib, (const void *)Lcc, lda, rest, lower ? 0 : 1, (const void *)Pc, lda, 0, (void *)Wc,
                                     ldw, dinfo, (void *)ss));
      continue;
    }
    if (pipe) CU_ERROR_CHECK(cuStreamWaitEvent(ss, h->ev_dcol[c - 1], 0)); /* L[c, 0:i] / U[0:i, c] is final */
    if (lower) { /* W_c = P_c - W[:, 0:i] * L[i:i+ib, 0:i]^H  (C = P_c in the matrix, D = W_c) */
      CU_ERROR_CHECK(chol_gemm(prec, 'N', 'C', rest, ib, i, MONE, W, ldw, chol_at(Ajj, lda, i, 0, es), lda, ONE, Pc, lda, Wc, ldw, ss));
    } else {     /* W_c = P_c^H, then W_c -= W[:, 0:i] * U[0:i, i:i+ib] */
      RT_CHECK(chol_copy_launch(prec, 'C', 0, ib, rest, (const void *)Pc, lda, (void *)Wc, ldw, (void *)ss));
      CU_ERROR_CHECK(chol_gemm(prec, 'N', 'N', rest, ib, i, MONE, W, ldw, chol_at(Ajj, lda, 0, i, es), lda, ONE, Wc, ldw, Wc, ldw, ss));
    }
    if (pipe) CU_ERROR_CHECK(cuStreamWaitEvent(ss, h->ev_dx[c], 0)); /* diagonal sub-block c is factored */
    RT_CHECK(chol_trsm_rows_launch(prec, uplo, ib, (const void *)Lcc, lda, rest, 0, (const void *)Wc, ldw, 0, (void *)Wc, ldw, dinfo,
                                   (void *)ss));
  }
  if (pipe) {
    CU_ERROR_CHECK(cuEventRecord(h->ev_solve, ss));
    CU_ERROR_CHECK(cuStreamWaitEvent(st, h->ev_solve, 0));
  }
  return CUDA_SUCCESS;
}

CUresult chol_panel_factor_solve_lower(CULAPACKhandle h, int prec, size_t jb, size_t rest, CUdeviceptr Ajj, CUdeviceptr P, size_t lda,
                                       CUdeviceptr W, size_t ldw, long *dinfo, long base, const chol_ws *w, CUstream st,
                                       CUevent before_solve) {
  return panel_factor_solve(h, prec, 'L', jb, rest, Ajj, P, lda, W, ldw, dinfo, base, w, st, before_solve);
}
CUresult chol_panel_factor_solve_upper(CULAPACKhandle h, int prec, size_t jb, size_t rest, CUdeviceptr Ajj, CUdeviceptr P, size_t lda,
                                       CUdeviceptr W, size_t ldw, long *dinfo, long base, const chol_ws *w, CUstream st,
                                       CUevent before_solve) {
  return panel_factor_solve(h, prec, 'U', jb, rest, Ajj, P, lda, W, ldw, dinfo, base, w, st, before_solve);
}

/* ------------------------------------------------------------------------------------------ */
static CUresult potrf_body(CULAPACKhandle h, int prec, int uplo, size_t n, CUdeviceptr A, size_t lda, long *info) {
  chol_ws w;
  CU_ERROR_CHECK(chol_carve(h, prec, n, &w));
  const size_t es = w.es, nb = w.nb;
  const int herm = chol_is_complex(prec);
  long *dinfo = (long *)h->dinfo;
  CUstream sp = h->s_panel, su = h->s_update, sl = h->s_look;
  CU_ERROR_CHECK(cuMemsetD8Async(h->dinfo, 0, 16, sp)); /* info and the log-determinant accumulator */
  h->last_logdet_valid = false;
  h->prof_n = 0;
  const bool split = chol_lookahead_split();
  CUevent look_pending = NULL; /* the lower part of this step's block column is still being updated on sl */

  size_t step = 0;
  for (size_t j = 0; j < n; j += nb, step++) {
    const size_t jb = chol_min(nb, n - j), rest = n - j - jb;
    const CUdeviceptr Ajj = chol_at(A, lda, j, j, es);
    const CUdeviceptr W = w.W[step & 1];
    /* --- panel stream: diagonal block, its inverse, panel solve into W, write-back --- */
    if (rest == 0) {
      CU_ERROR_CHECK(chol_potrf_diag(prec, uplo, jb, Ajj, lda, dinfo, (long)j, &w, sp));
      break;
    }
    const CUevent before_solve = look_pending;
    look_pending = NULL;
    /* The solved panel goes back into the matrix OFF the serial chain: nothing in the factorisation reads it there
     * (the updates read W), so the write-back rides on the update stream behind this step's trailing update; W is
     * not reused before that has finished (ev_rest).  Without a trailing update it stays on the panel stream. */
    const size_t jb2 = chol_min(nb, rest), rest2 = rest - jb2;
    const CUdeviceptr P = uplo == 'L' ? chol_at(A, lda, j + jb, j, es) /* rest x jb */ : chol_at(A, lda, j, j + jb, es) /* jb x rest */;
    const int wb_trans = uplo == 'L' ? 'N' : 'C'; /* upper: W = P^H * U11^-1 */
    if (uplo == 'L')
      CU_ERROR_CHECK(chol_panel_factor_solve_lower(h, prec, jb, rest, Ajj, P, lda, W, w.ldw, dinfo, (long)j, &w, sp, before_solve));
    else
      CU_ERROR_CHECK(chol_panel_factor_solve_upper(h, prec, jb, rest, Ajj, P, lda, W, w.ldw, dinfo, (long)j, &w, sp, before_solve));
    if (!rest2) RT_CHECK(chol_copy_launch(prec, wb_trans, 0, rest, jb, (const void *)W, w.ldw, (void *)P, lda, (void *)sp));
    /* s, c: the updates of this step run on the tensor cores - split W once into its TF32 (hi, lo) planes */
    const bool tens = w.kp != 0 && rest >= 512 && jb >= 64;
    const CUdeviceptr Whi = w.Whi[step & 1], Wlo = w.Wlo[step & 1];
    const size_t prow = (herm ? 2 : 1) * w.kp * sizeof(float); /* bytes of planes per matrix row */
    if (tens && !herm) RT_CHECK(chol_tf32_split_launch(0, rest, jb, (const void *)W, w.ldw, (void *)Whi, (void *)Wlo, w.kp, (void *)sp));
    if (tens && herm) RT_CHECK(chol_tf32_split_c_launch(rest, jb, (const void *)W, w.ldw, (void *)Whi, (void *)Wlo, w.kp, (void *)sp));
    CU_ERROR_CHECK(cuEventRecord(h->ev_panel[step % CHOL_EVENT_RING], sp));

    /* --- trailing update  A22 -= W W^H, split into the next block column/row (look-ahead, panel
     *     stream) and the rest (update stream) --- */
    if (rest2) {
      CU_ERROR_CHECK(cuStreamWaitEvent(su, h->ev_panel[step % CHOL_EVENT_RING], 0));
      const CUdeviceptr W2 = chol_at(W, w.ldw, jb2, 0, es);
      const CUdeviceptr A33 = chol_at(A, lda, j + jb + jb2, j + jb + jb2, es);
      const bool prof = h->profile && h->prof_n < h->prof_cap;
      if (prof) CU_ERROR_CHECK(cuEventRecord(h->prof_e0[h->prof_n], su));
      if (tens) {
        const CUdeviceptr H2 = Whi + (CUdeviceptr)(jb2 * prow), L2 = Wlo + (CUdeviceptr)(jb2 * prow);
        if (!herm)
          RT_CHECK(chol_tf32x3_gemm_launch(rest2, rest2, jb, -1.0, (const void *)H2, (const void *)L2, (const void *)H2, (const void *)L2,
                                           w.kp, 1.0, (const void *)A33, lda, (void *)A33, lda, uplo, (void *)su));
        else
          RT_CHECK(chol_tf32x3_herk_launch(rest2, rest2, jb, -1.0, (const void *)H2, (const void *)L2, (const void *)H2, (const void *)L2,
                                           w.kp, 1.0, (const void *)A33, lda, (void *)A33, lda, uplo, 1, (void *)su));
      } else
        CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'C', 0, 0, 0, 0, uplo, herm, rest2, rest2, jb, MONE, W2, w.ldw, W2, w.ldw, ONE,
                                    A33, lda, A33, lda, su));
      if (prof) {
        CU_ERROR_CHECK(cuEventRecord(h->prof_e1[h->prof_n], su));
        h->prof_flops[h->prof_n++] = (double)rest2 * (double)(rest2 + 1) * (double)jb * (herm ? 4.0 : 1.0);
      }
      RT_CHECK(chol_copy_launch(prec, wb_trans, 0, rest, jb, (const void *)W, w.ldw, (void *)P, lda, (void *)su)); /* write-back */
      CU_ERROR_CHECK(cuEventRecord(h->ev_rest[step % CHOL_EVENT_RING], su));
    }
    /* the look-ahead block was last written by the previous step's trailing update */
    if (step > 0) CU_ERROR_CHECK(cuStreamWaitEvent(sp, h->ev_rest[(step - 1) % CHOL_EVENT_RING], 0));
    {
      const CUdeviceptr A22 = chol_at(A, lda, j + jb, j + jb, es);
      const size_t below = rest - jb2; /* rows (columns, upper) of the look-ahead column outside its diagonal block */
      if (split && below) {
        /* diagonal block here; the rest on the look-ahead stream, which needs W (panel event) and the previous
         * trailing update of this column */
        CU_ERROR_CHECK(cuStreamWaitEvent(sl, h->ev_panel[step % CHOL_EVENT_RING], 0));
        if (step > 0) CU_ERROR_CHECK(cuStreamWaitEvent(sl, h->ev_rest[(step - 1) % CHOL_EVENT_RING], 0));
        const CUdeviceptr Wb = chol_at(W, w.ldw, jb2, 0, es);
        if (uplo == 'L') {
          CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'C', 0, 0, 0, 0, 'L', herm, jb2, jb2, jb, MONE, W, w.ldw, W, w.ldw, ONE, A22, lda, A22,
                                      lda, sp));
          const CUdeviceptr A32 = chol_at(A22, lda, jb2, 0, es);
          CU_ERROR_CHECK(chol_gemm(prec, 'N', 'C', below, jb2, jb, MONE, Wb, w.ldw, W, w.ldw, ONE, A32, lda, A32, lda, sl));
        } else {
          CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'C', 0, 0, 0, 0, 'U', herm, jb2, jb2, jb, MONE, W, w.ldw, W, w.ldw, ONE, A22, lda, A22,
                                      lda, sp));
          const CUdeviceptr A23 = chol_at(A22, lda, 0, jb2, es);
          CU_ERROR_CHECK(chol_gemm(prec, 'N', 'C', jb2, below, jb, MONE, W, w.ldw, Wb, w.ldw, ONE, A23, lda, A23, lda, sl));
        }
        look_pending = h->ev_look[step % CHOL_EVENT_RING];
        CU_ERROR_CHECK(cuEventRecord(look_pending, sl));
      } else if (tens) { /* the same block column / block row on the tensor cores */
        const size_t lm = uplo == 'L' ? rest : jb2, ln = uplo == 'L' ? jb2 : rest;
        if (!herm)
          RT_CHECK(chol_tf32x3_gemm_launch(lm, ln, jb, -1.0, (const void *)Whi, (const void *)Wlo, (const void *)Whi, (const void *)Wlo,
                                           w.kp, 1.0, (const void *)A22, lda, (void *)A22, lda, uplo, (void *)sp));
        else
          RT_CHECK(chol_tf32x3_herk_launch(lm, ln, jb, -1.0, (const void *)Whi, (const void *)Wlo, (const void *)Whi, (const void *)Wlo,
                                           w.kp, 1.0, (const void *)A22, lda, (void *)A22, lda, uplo, 1, (void *)sp));
      } else if (uplo == 'L') /* rest x jb2 block column, lower part */
        CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'C', 0, 0, 0, 0, 'L', herm, rest, jb2, jb, MONE, W, w.ldw, W, w.ldw, ONE, A22,
                                    lda, A22, lda, sp));
      else /* jb2 x rest block row, upper part */
        CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'C', 0, 0, 0, 0, 'U', herm, jb2, rest, jb, MONE, W, w.ldw, W, w.ldw, ONE, A22,
                                    lda, A22, lda, sp));
    }
  }
  CU_ERROR_CHECK(finish_info(h, sp, info));
  h->last_logdet = ((const double *)h->hinfo)[1];
  h->last_logdet_valid = (*info == 0);
  for (int i = 0; i < h->prof_n; i++) { /* the streams are idle: all events have completed */
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

/* ------------------------------------------------------------------------------------------ */
/* TRTRI's own block size (its per-block small inverse grows with nb^2 launches; measured in round 1, nb_sweep) */
static size_t trtri_default_nb(int prec, size_t n) {
  if (prec == 'z') return n >= 8192 ? 256 : 128;
  if (prec == 'd') return n >= 49152 ? 1024 : (n >= 12288 ? 512 : 256);
  return n >= 16384 ? 512 : 256;
}

static CUresult trtri_body(CULAPACKhandle h, int prec, int uplo, int diag, size_t n, CUdeviceptr A, size_t lda, long *info) {
  chol_ws w;
  CU_ERROR_CHECK(chol_carve_nb(h, prec, n, h->nb ? h->nb : trtri_default_nb(prec, n), &w));
  const size_t es = w.es, nb = w.nb;
  long *dinfo = (long *)h->dinfo;
  CUstream st = h->s_update;
  const int unit = (diag == 'U');
  const int strict = (uplo == 'L') ? 'l' : 'u';
  const size_t nblk = (n + nb - 1) / nb;
  /* CHOL_TRTRI_LOOKAHEAD=1 (written after the round-1 GPU budget was spent, off until measured): the small inverse
   * of a diagonal block depends on nothing but that block, so the NEXT block's inverse runs on the panel stream
   * while this block's two large GEMMs run on the update stream (two X buffers; the info-writing kernels stay
   * on one stream, in sweep order, so `info` is unchanged). */
  const bool la = chol_trtri_lookahead() && nblk > 1;
  CUstream sp = la ? h->s_panel : st;
  const CUdeviceptr Xb[2] = {w.Xd, la ? w.W[1] : w.Xd};
  CU_ERROR_CHECK(cuMemsetD8Async(h->dinfo, 0, sizeof(long), sp));
  for (size_t b = 0; b < nblk; b++) {
    const size_t blk = (uplo == 'L') ? nblk - 1 - b : b; /* reference sweep: lapack/dtrtri.c:119,:137 */
    const size_t j = blk * nb, jb = chol_min(nb, n - j);
    const CUdeviceptr Ajj = chol_at(A, lda, j, j, es);
    const CUdeviceptr Xd = Xb[b & 1];
    if (!la) {
      CU_ERROR_CHECK(chol_trtri_small(prec, uplo, diag, jb, Ajj, lda, Xd, nb, dinfo, (long)j, &w, st));
    } else {
      if (b == 0) {
        CU_ERROR_CHECK(chol_trtri_small(prec, uplo, diag, jb, Ajj, lda, Xd, nb, dinfo, (long)j, &w, sp));
        CU_ERROR_CHECK(cuEventRecord(h->ev_panel[0], sp));
      }
      if (b + 1 < nblk) { /* the next block's small inverse, into the other X buffer (last used by block b-1) */
        const size_t blk1 = (uplo == 'L') ? nblk - 2 - b : b + 1, j1 = blk1 * nb, jb1 = chol_min(nb, n - j1);
        if (b >= 1) CU_ERROR_CHECK(cuStreamWaitEvent(sp, h->ev_rest[(b - 1) % CHOL_EVENT_RING], 0));
        CU_ERROR_CHECK(chol_trtri_small(prec, uplo, diag, jb1, chol_at(A, lda, j1, j1, es), lda, Xb[(b + 1) & 1], nb, dinfo, (long)j1, &w,
                                        sp));
        CU_ERROR_CHECK(cuEventRecord(h->ev_panel[(b + 1) % CHOL_EVENT_RING], sp));
      }
      CU_ERROR_CHECK(cuStreamWaitEvent(st, h->ev_panel[b % CHOL_EVENT_RING], 0));
    }
    if (uplo == 'L') {
      const size_t rest = n - j - jb;
      if (rest) { /* A21 = -inv(A22) * (A21 * inv(A11)) ; inv(A22) is already in place */
        const CUdeviceptr A21 = chol_at(A, lda, j + jb, j, es), A22 = chol_at(A, lda, j + jb, j + jb, es);
        CU_ERROR_CHECK(chol_gemm(prec, 'N', 'N', rest, jb, jb, ONE, A21, lda, Xd, nb, ZERO, 0, 0, w.W[0], w.ldw, st));
        CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'N', 'L', unit, 0, 0, 0, 0, rest, jb, rest, MONE, A22, lda, w.W[0], w.ldw,
                                    ZERO, 0, 0, A21, lda, st));
      }
    } else if (j) { /* A12 = -inv(A11) * (A12 * inv(A22)) */
      const CUdeviceptr A12 = chol_at(A, lda, 0, j, es);
      CU_ERROR_CHECK(chol_gemm(prec, 'N', 'N', j, jb, jb, ONE, A12, lda, Xd, nb, ZERO, 0, 0, w.W[0], w.ldw, st));
      CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'N', 'U', unit, 0, 0, 0, 0, j, jb, j, MONE, A, lda, w.W[0], w.ldw, ZERO, 0, 0,
                                  A12, lda, st));
    }
    /* a unit-diagonal inverse leaves the caller's diagonal untouched (lapack/dtrtri.c:56-57) */
    RT_CHECK(chol_copy_launch(prec, 'N', unit ? strict : uplo, jb, jb, (const void *)Xd, nb, (void *)Ajj, lda, (void *)st));
    if (la) CU_ERROR_CHECK(cuEventRecord(h->ev_rest[b % CHOL_EVENT_RING], st));
  }
  CU_ERROR_CHECK(finish_info(h, sp, info));
  return CUDA_SUCCESS;
}

/* ------------------------------------------------------------------------------------------ */
static CUresult lauum_body(CULAPACKhandle h, int prec, int uplo, size_t n, CUdeviceptr A, size_t lda, long *info) {
  chol_ws w;
  /* LAUUM's GEMMs are ib x i with k = n-i: few, fat tiles; a taller block row keeps all SMs busy
   * (measured d, N = 32768: nb 256 / 512 / 1024 -> 15.6 / 22.2 / 28.8 TFLOP/s) */
  size_t lnb = h->nb ? h->nb : chol_default_nb(prec, n);
  if (!h->nb && (prec == 'd' || prec == 's') && n >= 16384) lnb = 1024;
  CU_ERROR_CHECK(chol_carve_nb(h, prec, n, lnb, &w));
  const size_t es = w.es, nb = w.nb;
  const int herm = chol_is_complex(prec);
  CUstream st = h->s_update;
  *info = 0;
  for (size_t i = 0; i < n; i += nb) {
    const size_t ib = chol_min(nb, n - i), k = n - i;
    const CUdeviceptr Aii = chol_at(A, lda, i, i, es);
    if (uplo == 'L') {
      /* row block i of L^H L = (L[i:, i-block])^H * L[i:, 0:i+ib]; the trapezoid's top block is
       * lower triangular (mask), products go to scratch because the output aliases the operands */
      if (i)
        CU_ERROR_CHECK(chol_gemm_ex(prec, 'C', 'N', 'L', 0, 0, 0, 0, 0, ib, i, k, ONE, Aii, lda, chol_at(A, lda, i, 0, es), lda,
                                    ZERO, 0, 0, w.W[0], nb, st));
      CU_ERROR_CHECK(chol_gemm_ex(prec, 'C', 'N', 'L', 0, 'L', 0, 'L', herm, ib, ib, k, ONE, Aii, lda, Aii, lda, ZERO, 0, 0,
                                  w.Xd, nb, st));
      if (i) RT_CHECK(chol_copy_launch(prec, 'N', 0, ib, i, (const void *)w.W[0], nb, (void *)chol_at(A, lda, i, 0, es), lda, (void *)st));
    } else {
      /* column block i of U U^H = U[0:i+ib, i:] * (U[i-block, i:])^H */
      if (i)
        CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'C', 0, 0, 'U', 0, 0, 0, i, ib, k, ONE, chol_at(A, lda, 0, i, es), lda, Aii, lda,
                                    ZERO, 0, 0, w.W[0], w.ldw, st));
      CU_ERROR_CHECK(chol_gemm_ex(prec, 'N', 'C', 'U', 0, 'U', 0, 'U', herm, ib, ib, k, ONE, Aii, lda, Aii, lda, ZERO, 0, 0,
                                  w.Xd, nb, st));
      if (i) RT_CHECK(chol_copy_launch(prec, 'N', 0, i, ib, (const void *)w.W[0], w.ldw, (void *)chol_at(A, lda, 0, i, es), lda, (void *)st));
    }
    RT_CHECK(chol_copy_launch(prec, 'N', uplo, ib, ib, (const void *)w.Xd, nb, (void *)Aii, lda, (void *)st));
  }
  CU_ERROR_CHECK(cuStreamSynchronize(st));
  return CUDA_SUCCESS;
}

/* The drivers run inside the handle's context (the reference pushes it per call, lapack/dpotrf.c:181);
 * the pop happens on every path, so a failing call leaves the caller's context stack as it found it. */
#define IN_HANDLE_CONTEXT(h, call)                   \
  do {                                               \
    CU_ERROR_CHECK(cuCtxPushCurrent((h)->context));  \
    const CUresult body__ = (call);                  \
    CUcontext popped__;                              \
    const CUresult pop__ = cuCtxPopCurrent(&popped__); \
    if (body__ != CUDA_SUCCESS) return body__; /* already reported where it happened */ \
    CU_ERROR_CHECK(pop__);                           \
    return CUDA_SUCCESS;                             \
  } while (false)

CUresult chol_potrf(CULAPACKhandle h, int prec, int uplo, size_t n, CUdeviceptr A, size_t lda, long *info) {
  IN_HANDLE_CONTEXT(h, potrf_body(h, prec, uplo, n, A, lda, info));
}
CUresult chol_trtri(CULAPACKhandle h, int prec, int uplo, int diag, size_t n, CUdeviceptr A, size_t lda, long *info) {
  IN_HANDLE_CONTEXT(h, trtri_body(h, prec, uplo, diag, n, A, lda, info));
}
CUresult chol_lauum(CULAPACKhandle h, int prec, int uplo, size_t n, CUdeviceptr A, size_t lda, long *info) {
  IN_HANDLE_CONTEXT(h, lauum_body(h, prec, uplo, n, A, lda, info));
}

CUresult chol_potri(CULAPACKhandle h, int prec, int uplo, size_t n, CUdeviceptr A, size_t lda, long *info) {
  CU_ERROR_CHECK(chol_trtri(h, prec, uplo, 'N', n, A, lda, info));
  if (*info != 0) return CUDA_SUCCESS;
  return chol_lauum(h, prec, uplo, n, A, lda, info);
}

/* Solve A X = B from the Cholesky factor (the reference stops at POTRI):
 * lower  L Y = B then L^H X = Y,  upper  U^H Y = B then U X = Y - two blocked TRSMs on the update stream. */
static CUresult potrs_body(CULAPACKhandle h, int prec, int uplo, size_t n, size_t nrhs, CUdeviceptr A, size_t lda,
                           CUdeviceptr B, size_t ldb) {
  CUstream st = h->s_update;
  const int first = (uplo == 'L') ? 'N' : 'C', second = (uplo == 'L') ? 'C' : 'N';
  CU_ERROR_CHECK(chol_trsm(h->blas_handle, prec, 'L', uplo, first, 'N', n, nrhs, ONE, A, lda, B, ldb, st));
  CU_ERROR_CHECK(chol_trsm(h->blas_handle, prec, 'L', uplo, second, 'N', n, nrhs, ONE, A, lda, B, ldb, st));
  CU_ERROR_CHECK(cuStreamSynchronize(st));
  return CUDA_SUCCESS;
}
CUresult chol_potrs(CULAPACKhandle h, int prec, int uplo, size_t n, size_t nrhs, CUdeviceptr A, size_t lda, CUdeviceptr B,
                    size_t ldb) {
  IN_HANDLE_CONTEXT(h, potrs_body(h, prec, uplo, n, nrhs, A, lda, B, ldb));
}

/* log det A = 2 sum log L(i,i) of the last cu*potrf on this handle, accumulated by the diagonal-block kernels while they
 * factored; INVALID_VALUE when that factorisation failed or none has run */
CUresult cuLAPACKPotrfLogdet(CULAPACKhandle h, double *logdet) {
  if (h == NULL || logdet == NULL || !h->last_logdet_valid) return CUDA_ERROR_INVALID_VALUE;
  *logdet = h->last_logdet;
  return CUDA_SUCCESS;
}

static CUresult logdet_body(CULAPACKhandle h, int prec, CUdeviceptr x, size_t incx, size_t n, void *result, CUstream stream) {
  const size_t rs = (prec == 's' || prec == 'c') ? 4 : 8;
  const CUdeviceptr dres = h->logdet_scr + (CUdeviceptr)chol_roundup(chol_logdet_scratch_bytes(), 16);
  /* partials, ticket and result slot are one per handle (the reference allocates its temporary per call,
   * lapack/dlogdet.c:37-74): a call on another stream first waits for the previous call's read-back */
  if (h->logdet_in_use && h->logdet_stream != stream) CU_ERROR_CHECK(cuStreamWaitEvent(stream, h->logdet_done, 0));
  RT_CHECK(chol_logdet_launch(prec, (const void *)x, incx, n, (void *)dres, (void *)h->logdet_scr, (void *)stream));
  CU_ERROR_CHECK(cuMemcpyDtoHAsync(result, dres, rs, stream));
  CU_ERROR_CHECK(cuEventRecord(h->logdet_done, stream));
  h->logdet_stream = stream;
  h->logdet_in_use = true;
  return CUDA_SUCCESS;
}
CUresult chol_logdet(CULAPACKhandle h, int prec, CUdeviceptr x, size_t incx, size_t n, void *result, CUstream stream) {
  IN_HANDLE_CONTEXT(h, logdet_body(h, prec, x, incx, n, result, stream));
}

/* ---------------------------------------------------------------------------------------------
 * typed entry points (include/lapack.h).  Argument checks follow the reference:
 * potrf/lauum/potri lda -> info -4 (dpotrf.c:174-179), trtri lda -> -5 (dtrtri.c:330-335).
 * ------------------------------------------------------------------------------------------- */
#define LAPACK_ARG(cond, pos)               \
  do {                                      \
    if (cond) {                             \
      *info = -(pos);                       \
      XERBLA(pos);                          \
      return CUDA_ERROR_INVALID_VALUE;      \
    }                                       \
  } while (false)

#define GEN_LAPACK(U, p, REAL)                                                                                       \
  CUresult cu##U##potrf(CULAPACKhandle h, CBlasUplo uplo, size_t n, CUdeviceptr A, size_t lda, long *info) {         \
    *info = 0;                                                                                                       \
    if (h == NULL) return CUDA_ERROR_INVALID_VALUE;                                                                  \
    LAPACK_ARG(lda < n, 4);                                                                                          \
    if (n == 0) return CUDA_SUCCESS;                                                                                 \
    return chol_potrf(h, p, uplo, n, A, lda, info);                                                                  \
  }                                                                                                                  \
  CUresult cu##U##trtri(CULAPACKhandle h, CBlasUplo uplo, CBlasDiag diag, size_t n, CUdeviceptr A, size_t lda,       \
                        long *info) {                                                                                \
    *info = 0;                                                                                                       \
    if (h == NULL) return CUDA_ERROR_INVALID_VALUE;                                                                  \
    LAPACK_ARG(lda < n, 5);                                                                                          \
    if (n == 0) return CUDA_SUCCESS;                                                                                 \
    return chol_trtri(h, p, uplo, diag, n, A, lda, info);                                                            \
  }                                                                                                                  \
  CUresult cu##U##trti22(CULAPACKhandle h, CBlasUplo uplo, CBlasDiag diag, size_t n, CUdeviceptr A, size_t lda,      \
                         CUdeviceptr B, size_t ldb, CUdeviceptr info, CUstream stream) {                             \
    (void)h;                                                                                                         \
    if (n > (size_t)chol_diag_max(p) || lda < n || ldb < n) return CUDA_ERROR_INVALID_VALUE;                         \
    /* info: device int, zeroed by the kernel itself on entry (the reference kernel's contract, dtrtri.cu:23-29) */  \
    RT_CHECK(chol_trti2_launch(p, uplo, diag, n, (const void *)A, lda, (void *)B, ldb, CHOL_TRTI2_INFO_INT32,        \
                               (long *)info, 0, (void *)stream));                                                    \
    return CUDA_SUCCESS;                                                                                             \
  }                                                                                                                  \
  CUresult cu##U##lauum(CULAPACKhandle h, CBlasUplo uplo, size_t n, CUdeviceptr A, size_t lda, long *info) {         \
    *info = 0;                                                                                                       \
    if (h == NULL) return CUDA_ERROR_INVALID_VALUE;                                                                  \
    LAPACK_ARG(lda < n, 4);                                                                                          \
    if (n == 0) return CUDA_SUCCESS;                                                                                 \
    return chol_lauum(h, p, uplo, n, A, lda, info);                                                                  \
  }                                                                                                                  \
  CUresult cu##U##potri(CULAPACKhandle h, CBlasUplo uplo, size_t n, CUdeviceptr A, size_t lda, long *info) {         \
    *info = 0;                                                                                                       \
    if (h == NULL) return CUDA_ERROR_INVALID_VALUE;                                                                  \
    LAPACK_ARG(lda < n, 4);                                                                                          \
    if (n == 0) return CUDA_SUCCESS;                                                                                 \
    return chol_potri(h, p, uplo, n, A, lda, info);                                                                  \
  }                                                                                                                  \
  CUresult cu##U##potrs(CULAPACKhandle h, CBlasUplo uplo, size_t n, size_t nrhs, CUdeviceptr A, size_t lda,          \
                        CUdeviceptr B, size_t ldb, long *info) {                                                     \
    *info = 0;                                                                                                       \
    if (h == NULL) return CUDA_ERROR_INVALID_VALUE;                                                                  \
    LAPACK_ARG(lda < n, 5);                                                                                          \
    LAPACK_ARG(ldb < n, 7);                                                                                          \
    if (n == 0 || nrhs == 0) return CUDA_SUCCESS;                                                                    \
    return chol_potrs(h, p, uplo, n, nrhs, A, lda, B, ldb);                                                          \
  }                                                                                                                  \
  CUresult cu##U##logdet(CULAPACKhandle h, CUdeviceptr x, size_t incx, size_t n, REAL *result, CUstream stream) {    \
    if (h == NULL || result == NULL) return CUDA_ERROR_INVALID_VALUE;                                                \
    if (n == 0) {                                                                                                    \
      *result = 0;                                                                                                   \
      return CUDA_SUCCESS;                                                                                           \
    }                                                                                                                \
    return chol_logdet(h, p, x, incx, n, result, stream);                                                            \
  }

GEN_LAPACK(S, 's', float)
GEN_LAPACK(D, 'd', double)
GEN_LAPACK(C, 'c', float)
GEN_LAPACK(Z, 'z', double)
