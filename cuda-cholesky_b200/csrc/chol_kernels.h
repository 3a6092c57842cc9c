/*
 * chol_kernels.h — the thin C-ABI between the C host code (the .c files under ../host) and the
 * hand-written sm_100a kernels (the .cu files in this directory).  Plain C: pointers, sizes and
 * chars only.  Every launcher is asynchronous on `stream` (a CUstream / cudaStream_t) and
 * returns a cudaError_t value as int (0 = success).
 *
 * Conventions (same as the reference's BLAS, include/blas.h:24-27): column-major, leading
 * dimensions in elements, enum arguments are the reference's char values.
 */
#ifndef CHOL_KERNELS_H
#define CHOL_KERNELS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- the GEMM engine: D = alpha * op(A) * op(B) + beta * C  (+ triangle / Hermitian options) ----
 * One descriptor serves GEMM, SYRK/HERK (cuplo), TRMM (triA/triB masks) and the inverse-based
 * panel solves.  prec 'd'/'z' run on FP64 DMMA tensor-core tiles, 's'/'c' on FP32 FFMA tiles.  */
typedef struct chol_gemm_desc {
  int prec;             /* 's','d','c','z' */
  int opA, opB;         /* 'N','T','C': how the stored operand enters the product */
  int triA, triB;       /* 0, 'L' or 'U': stored operand is triangular; the other triangle is
                           never read and counts as zero */
  int unitA, unitB;     /* nonzero: triangular operand has an implicit unit diagonal */
  int cuplo;            /* 0 = all of D; 'L'/'U' = only that triangle of C is read and of D written */
  int herm;             /* nonzero: force Im(diag(D)) = 0 (HERK, blas/zherk.c:108) */
  size_t m, n, k;       /* D is m x n, inner dimension k */
  double alpha[2];      /* (re, im); real precisions use [0] */
  double beta[2];
  const void *A; size_t lda;
  const void *B; size_t ldb;
  const void *C; size_t ldc;   /* may equal D */
  void *D; size_t ldd;
  void *stream;
  int force_generic;    /* test hook: never take the TMA-fed path */
  /* block-cyclic column map (multi-GPU trailing update; 0,0,0 = off): D's packed local column n
   * stands for global triangle column  g(n) = ((n / cyc_nb) * cyc_p + cyc_off) * cyc_nb + n % cyc_nb;
   * the B operand row of column n is g(n) and cuplo 'L' keeps m >= g(n).  Requires op(B) != 'N'. */
  int cyc_nb, cyc_p, cyc_off;
} chol_gemm_desc;

int chol_gemm_launch(const chol_gemm_desc *d);
/* 1 if the descriptor would take the TMA (cp.async.bulk) operand path, 0 for the generic loader */
int chol_gemm_uses_tma(const chol_gemm_desc *d);

/* ---- s-precision tensor-core engine (csrc/gemm_tf32x3.cu): D = alpha * A * B^T + beta * C as a 3xTF32 split
 * on tcgen05 with FP32 accumulation in TMEM.  Operands are given as TF32 (hi, lo) planes, K-major with
 * pitch kp = chol_tf32_kpitch(k) floats: plane[r * kp + c] for r < rows, zero for k <= c < kp. ---- */
size_t chol_tf32_kpitch(size_t k);
/* split the stored FP32 operand into its planes.  trans == 0: element (r, c) at A[c * lda + r] (rows x k,
 * column-major); trans != 0: at A[r * lda + c] */
int chol_tf32_split_launch(int trans, size_t rows, size_t k, const void *A, size_t lda, void *hi, void *lo, size_t kp,
                           void *stream);
/* cuplo 0 / 'L' / 'U': all of D, or only that triangle of C read and of D written (C may equal D) */
int chol_tf32x3_gemm_launch(size_t m, size_t n, size_t k, double alpha, const void *Ahi, const void *Alo, const void *Bhi,
                            const void *Blo, size_t kp, double beta, const void *C, size_t ldc, void *D, size_t ldd, int cuplo,
                            void *stream);

/* c-precision: the same engine on the real embedding of the complex product.  chol_tf32_split_c_launch writes, for a
 * complex operand W (rows x k, column-major), planes of 2*rows rows and pitch kp = chol_tf32_kpitch(2*k) floats:
 * plane row 2i = (Re, Im) pairs of W's row i, plane row 2i+1 = (Im, -Re) pairs.  chol_tf32x3_herk_launch computes
 * D (m x n complex) = alpha * A * B^H + beta * C (real alpha, beta; cuplo on complex indices; herm != 0: Im(diag) = 0)
 * where Ahi/Alo point at the plane row of A's first row (2 * i0 * kp floats into the planes) and Bhi/Blo likewise. */
int chol_tf32_split_c_launch(size_t rows, size_t k, const void *W, size_t ldw, void *hi, void *lo, size_t kp, void *stream);
int chol_tf32x3_herk_launch(size_t m, size_t n, size_t k, double alpha, const void *Ahi, const void *Alo, const void *Bhi,
                            const void *Blo, size_t kp, double beta, const void *C, size_t ldc, void *D, size_t ldd, int cuplo,
                            int herm, void *stream);

/* ---- diagonal-block kernels: one CTA, block staged in shared memory (n <= CHOL_DIAG_MAX) ---- */
#define CHOL_DIAG_MAX 128
int chol_diag_max(int prec); /* largest n the one-CTA diagonal kernels accept: 128 (s,d,c), 64 (z) */
/* Cholesky of an n x n diagonal block in place.  *dinfo (device long) is set to
 * info_base + failing column (1-based) if it is still 0 and a pivot is <= 0 or NaN; when
 * *dinfo != 0 on entry the kernel does nothing (lapack/dpotrf.c:70-73 semantics). */
int chol_potf2_launch(int prec, int uplo, size_t n, void *A, size_t lda, long *dinfo, long info_base, void *stream);
/* the same, and on success *logacc (device double, may be NULL) += 2 * sum_c log Re L(c,c): the log-determinant of the
 * factored matrix accumulates over the diagonal-block launches of one factorisation (they are serialised on one stream) */
int chol_potf2_logdet_launch(int prec, int uplo, size_t n, void *A, size_t lda, long *dinfo, long info_base, double *logacc,
                             void *stream);
/* X (m x n) = B * L^-H by substitution, L = the n x n lower-triangular factor of a diagonal block (uplo 'L': stored
 * as such; uplo 'U': stored as U = L^H, so this is X = B * U^-1).  Any number of rows m; rows are independent, the
 * launch spreads them over the SMs.  transB / transX != 0: the operand is given / wanted as its conjugate transpose
 * (n x m: element (row, j) at ptr[j + row * ld], conjugated).  X may equal B.  dinfo (may be NULL): if *dinfo != 0
 * on entry the launch does nothing (the factorisation has already failed).  No inverse is formed anywhere. */
int chol_trsm_rows_launch(int prec, int uplo, size_t n, const void *L, size_t ldl, size_t m, int transB, const void *B,
                          size_t ldb, int transX, void *X, size_t ldx, const long *dinfo, void *stream);
/* inverse of an n x n triangular block, out of place (B may equal A).  Only the uplo triangle of
 * B is written; if (zero_other & 1) the opposite triangle of B is set to zero as well.  Zero pivot
 * -> *dinfo = info_base + column (1-based) when *dinfo == 0 (lapack/dtrtri.c:49-52); a nonzero
 * *dinfo on entry makes the launch a no-op (sticky, for the LAPACK drivers).  Two more protocols,
 * selected by bits of zero_other:
 *   CHOL_TRTI2_INFO_INT32   dinfo points to a 32-bit int that the kernel zeroes itself on entry and
 *                           sets on a zero pivot - the reference's device-kernel contract
 *                           (lapack/dtrtri.cu:23-29), used by cu*trti22;
 *   CHOL_TRTI2_NO_INFO      dinfo is not touched; a zero pivot yields Inf/NaN as a BLAS TRSM does. */
#define CHOL_TRTI2_INFO_INT32 2
#define CHOL_TRTI2_NO_INFO 4
int chol_trti2_launch(int prec, int uplo, int diag, size_t n, const void *A, size_t lda, void *B, size_t ldb,
                      int zero_other, long *dinfo, long info_base, void *stream);
/* A := U*U^H (upper) or L^H*L (lower) of an n x n triangular block in place (lapack/dlauum.c:43-69) */
int chol_lauu2_launch(int prec, int uplo, size_t n, void *A, size_t lda, void *stream);

/* ---- data movement helpers ---- */
/* B(m x n) := A(m x n), or B(n x m) := A^T / A^H when trans is 'T' / 'C'.  uplo 0 = all;
 * 'L'/'U' = only that triangle of the SOURCE is copied (to the corresponding destination slot),
 * 'l'/'u' = the same without the diagonal. */
int chol_copy_launch(int prec, int trans, int uplo, size_t m, size_t n, const void *A, size_t lda, void *B,
                     size_t ldb, void *stream);
/* B(g(r), c) := A(r, c), r < m, c < n, with the block-cyclic row map g(r) = ((r / cyc_nb) * cyc_p + cyc_off) * cyc_nb + r % cyc_nb
 * (a rank's packed piece of a row block goes to its global positions; distributed LAUUM) */
int chol_scatter_rows_launch(int prec, size_t m, size_t n, const void *A, size_t lda, void *B, size_t ldb, int cyc_nb, int cyc_p,
                             int cyc_off, void *stream);
/* A(m x n) := value on the chosen part: uplo 0 all, 'L' strictly-lower, 'U' strictly-upper */
int chol_fill_launch(int prec, int uplo, size_t m, size_t n, double re, double im, void *A, size_t lda, void *stream);

/* ---- logdet: *dresult (device REAL) = 2 * sum_i log(Re x[i*incx]); one kernel, no host sync ---- */
int chol_logdet_launch(int prec, const void *x, size_t incx, size_t n, void *dresult, void *dscratch, void *stream);
size_t chol_logdet_scratch_bytes(void);

/* ---- device dSFMT-19937 (csrc/rng_dsfmt.cu, host/rng.c).  A state is 192 128-bit words (191 + the lung). ----
 * jump: states[g] (g < nseg, 3072 bytes each) = XOR over the set bits i < D of polys[g] (poly_u32 32-bit words per row)
 * of K[i] (the seed's first D states): the state after any number of refills as a GF(2) combination. */
int chol_dsfmt_jump_launch(const void *K, const void *polys, int poly_u32, int D, void *states, int nseg, void *stream);
/* generate: one warp per segment g, from states[g] = the state before refill refill0 + g * seg_refills, produces that
 * segment's refills (refill r = stream words 191 r .. 191 r + 190) and stores the words inside [w_begin, w_end):
 * word t -> elements 2 (t - w_begin), +1 of out (element index < n_out), converted per kind (0 raw 64-bit, 1 [0,1),
 * 2 (0,1], 3 (0,1), 4 the reference's getCloseClose).  snap != NULL (nseg = 1): the raw state after refill r goes to
 * snap[(r + 1) * 192 words] instead (the Krylov block), nothing to out. */
int chol_dsfmt_generate_launch(const void *states, unsigned long long refill0, unsigned seg_refills, int nseg,
                               unsigned long long w_begin, unsigned long long w_end, void *out, unsigned long long n_out,
                               int kind, void *snap, void *stream);

/* number of kernels launched by the library so far (optionally resetting the counter) */
unsigned long long chol_launch_count(int reset);

/* ---- measurement probes (used by bench.py / profiles, not by the library path) ---- */
/* kind 0: DMMA m8n8k4 only; 1: DFMA only; 2: both interleaved.  Returns achieved TFLOP/s. */
int chol_probe_fp64(int kind, int iters, double *tflops, float *ms);
/* dependent-issue latency in cycles: kind 0 DFMA, 1 DMUL, 2 rsqrt(double), 3 64-bit shuffle, 4 shared-memory load */
int chol_probe_latency(int kind, double *cycles_per_op);
/* test hook: later chol_potf2_launch calls write 64 clock64() phase timestamps to this device buffer (NULL = off) */
void chol_potf2_set_debug(long long *device_buffer_64);

#ifdef __cplusplus
}
#endif
#endif
