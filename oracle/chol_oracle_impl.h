/*
 * chol_oracle_impl.h — TEST INFRASTRUCTURE ONLY.  Type-generic body of the CPU oracle.
 *
 * A plain-C restatement (not a copy) of the reference's CPU path for the Cholesky / SPD-inverse
 * hot path.  It is compiled four times (chol_oracle_{s,d,c,z}.c) with
 *     SCALAR   element type          REAL   its real type
 *     X(name)  symbol builder: oracle_<p><name>, p in {s,d,c,z}
 *     IS_COMPLEX 0/1
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the resulting library.  The product library (libcholesky_b200.so) never links or calls it.
 *
 * Parity pinning: tests/test_oracle_pins.py checks every routine here against
 * oracle/_ref/libref_oracle.so (the reference's own sources compiled unmodified, see
 * oracle/Makefile) when that library is present, and against tests/golden/ fixtures that were
 * generated from it (tests/golden/make_golden.py) when it is not.
 *
 * Each routine cites the reference file:line whose algorithm it follows.  All matrices are
 * column-major, leading dimensions in elements, enums are the reference's char values
 * ('N','T','C' / 'L','U' / 'L','R' / 'N','U'; include/blas.h:24-27).
 */
#include <stddef.h>
#include <math.h>
#include <complex.h>

#ifndef SCALAR
#error "include from chol_oracle_{s,d,c,z}.c"
#endif

#if IS_COMPLEX
#define CONJ(x) (sizeof(REAL) == sizeof(float) ? (SCALAR)conjf((float complex)(x)) : (SCALAR)conj((double complex)(x)))
#define RE(x) ((REAL)creal((double complex)(x)))
#else
#define CONJ(x) (x)
#define RE(x) (x)
#endif

#define EL(M, ld, i, j) ((M)[(size_t)(j) * (ld) + (i)])

static inline size_t X(min_)(size_t a, size_t b) { return a < b ? a : b; }

/* element (i,j) of op(A) */
static inline SCALAR X(op_)(const SCALAR *A, size_t lda, int trans, size_t i, size_t j) {
  if (trans == 'N') return EL(A, lda, i, j);
  if (trans == 'T') return EL(A, lda, j, i);
  return CONJ(EL(A, lda, j, i));
}

/* ------------------------------------------------------------------------------------------
 * GEMM   C := alpha*op(A)*op(B) + beta*C            reference: blas/dgemm.c:34-146 (zgemm.c:34)
 * Same column-parallel structure: a column of C is scaled by beta, then accumulated either as
 * axpys over columns of A (A not transposed) or as dot products (A transposed).
 * Returns the reference's XERBLA position (8/10/13) or 0.
 * ------------------------------------------------------------------------------------------ */
int X(gemm)(int transA, int transB, size_t m, size_t n, size_t k, SCALAR alpha,
            const SCALAR *A, size_t lda, const SCALAR *B, size_t ldb, SCALAR beta, SCALAR *C,
            size_t ldc) {
  const size_t nRowA = (transA == 'N') ? m : k, nRowB = (transB == 'N') ? k : n;
  if (lda < nRowA) return 8;
  if (ldb < nRowB) return 10;
  if (ldc < m) return 13;
  if (m == 0 || n == 0 || ((alpha == 0 || k == 0) && beta == 1)) return 0;

#pragma omp parallel for schedule(static)
  for (size_t j = 0; j < n; j++) {
    SCALAR *c = &EL(C, ldc, 0, j);
    if (beta == 0)
      for (size_t i = 0; i < m; i++) c[i] = 0;
    else if (beta != 1)
      for (size_t i = 0; i < m; i++) c[i] *= beta;
    if (alpha == 0 || k == 0) continue;
    if (transA == 'N') {
      for (size_t l = 0; l < k; l++) {
        const SCALAR t = alpha * X(op_)(B, ldb, transB, l, j);
        const SCALAR *a = &EL(A, lda, 0, l);
        for (size_t i = 0; i < m; i++) c[i] += t * a[i];
      }
    } else {
      for (size_t i = 0; i < m; i++) {
        const SCALAR *a = &EL(A, lda, 0, i);
        SCALAR t = 0;
        if (transA == 'T')
          for (size_t l = 0; l < k; l++) t += a[l] * X(op_)(B, ldb, transB, l, j);
        else
          for (size_t l = 0; l < k; l++) t += CONJ(a[l]) * X(op_)(B, ldb, transB, l, j);
        c[i] += alpha * t;
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * SYRK (real)  C_tri := alpha*A*A^T + beta*C_tri          reference: blas/dsyrk.c:34-161
 * HERK (cplx)  C_tri := alpha*A*A^H + beta*C_tri          reference: blas/zherk.c:34-186
 *   alpha,beta real; the imaginary part of the diagonal is forced to zero (zherk.c:108,:132);
 *   trans=='T' is argument error 2 for HERK (zherk.c:42-43).
 * Only the `uplo` triangle of C is referenced.
 * ------------------------------------------------------------------------------------------ */
#if IS_COMPLEX
int X(herk)
#else
int X(syrk)
#endif
(int uplo, int trans, size_t n, size_t k, REAL alpha, const SCALAR *A, size_t lda, REAL beta,
 SCALAR *C, size_t ldc) {
  const size_t nRowA = (trans == 'N') ? n : k;
#if IS_COMPLEX
  if (trans == 'T') return 2;
#endif
  if (lda < nRowA) return 7;
  if (ldc < n) return 10;
  if (n == 0 || ((alpha == 0 || k == 0) && beta == 1)) return 0;

#pragma omp parallel for schedule(dynamic, 8)
  for (size_t j = 0; j < n; j++) {
    const size_t i0 = (uplo == 'U') ? 0 : j, i1 = (uplo == 'U') ? j + 1 : n;
    SCALAR *c = &EL(C, ldc, 0, j);
    if (beta == 0)
      for (size_t i = i0; i < i1; i++) c[i] = 0;
    else if (beta != 1)
      for (size_t i = i0; i < i1; i++) c[i] *= beta;
    if (alpha != 0 && k != 0) {
      if (trans == 'N') {
        for (size_t l = 0; l < k; l++) {
          const SCALAR t = alpha * CONJ(EL(A, lda, j, l));
          const SCALAR *a = &EL(A, lda, 0, l);
          for (size_t i = i0; i < i1; i++) c[i] += t * a[i];
        }
      } else {
        const SCALAR *aj = &EL(A, lda, 0, j);
        for (size_t i = i0; i < i1; i++) {
          const SCALAR *ai = &EL(A, lda, 0, i);
          SCALAR t = 0;
          for (size_t l = 0; l < k; l++) t += CONJ(ai[l]) * aj[l];
          c[i] += alpha * t;
        }
      }
    }
#if IS_COMPLEX
    c[j] = RE(c[j]);
#endif
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * TRSM   B := alpha*op(A)^-1*B  (Left)   or   B := alpha*B*op(A)^-1  (Right)
 * reference: blas/dtrsm.c:34-216 (ztrsm.c).  Restated through the shape of T = op(A):
 * T is lower-triangular iff (uplo=='L') == (trans=='N').  Left = forward/backward substitution
 * per column of B (columns are independent); Right = column sweep over B.
 * ------------------------------------------------------------------------------------------ */
int X(trsm)(int side, int uplo, int trans, int diag, size_t m, size_t n, SCALAR alpha,
            const SCALAR *A, size_t lda, SCALAR *B, size_t ldb) {
  const size_t nRowA = (side == 'L') ? m : n;
  if (lda < nRowA) return 9;
  if (ldb < m) return 11;
  if (m == 0 || n == 0) return 0;
  if (alpha == 0) {
    for (size_t j = 0; j < n; j++)
      for (size_t i = 0; i < m; i++) EL(B, ldb, i, j) = 0;
    return 0;
  }
  const int lower = (uplo == 'L') == (trans == 'N');
  if (side == 'L') {
#pragma omp parallel for schedule(static)
    for (size_t j = 0; j < n; j++) {
      SCALAR *b = &EL(B, ldb, 0, j);
      if (alpha != 1)
        for (size_t i = 0; i < m; i++) b[i] *= alpha;
      if (lower) {
        for (size_t i = 0; i < m; i++) {
          SCALAR t = b[i];
          if (trans == 'N') { /* column-oriented elimination keeps unit stride on A */
            if (diag == 'N') t /= EL(A, lda, i, i);
            b[i] = t;
            for (size_t l = i + 1; l < m; l++) b[l] -= t * EL(A, lda, l, i);
          } else {
            for (size_t l = 0; l < i; l++) t -= X(op_)(A, lda, trans, i, l) * b[l];
            if (diag == 'N') t /= X(op_)(A, lda, trans, i, i);
            b[i] = t;
          }
        }
      } else {
        for (size_t ii = m; ii-- > 0;) {
          SCALAR t = b[ii];
          if (trans == 'N') {
            if (diag == 'N') t /= EL(A, lda, ii, ii);
            b[ii] = t;
            for (size_t l = 0; l < ii; l++) b[l] -= t * EL(A, lda, l, ii);
          } else {
            for (size_t l = ii + 1; l < m; l++) t -= X(op_)(A, lda, trans, ii, l) * b[l];
            if (diag == 'N') t /= X(op_)(A, lda, trans, ii, ii);
            b[ii] = t;
          }
        }
      }
    }
  } else {
    /* X*T = alpha*B.  T upper: sweep j upwards, X[:,j] = (B[:,j] - sum_{l<j} X[:,l] T(l,j)) / T(j,j)
     *                 T lower: sweep j downwards.                                              */
    for (size_t jj = 0; jj < n; jj++) {
      const size_t j = lower ? n - 1 - jj : jj;
      SCALAR *b = &EL(B, ldb, 0, j);
      if (alpha != 1)
        for (size_t i = 0; i < m; i++) b[i] *= alpha;
      const size_t l0 = lower ? j + 1 : 0, l1 = lower ? n : j;
      for (size_t l = l0; l < l1; l++) {
        const SCALAR t = X(op_)(A, lda, trans, l, j);
        if (t == 0) continue;
        const SCALAR *x = &EL(B, ldb, 0, l);
        /* columns l were already divided by alpha-scaled pivots: undo nothing, B[:,l] is X[:,l] */
        for (size_t i = 0; i < m; i++) b[i] -= t * x[i];
      }
      if (diag == 'N') {
        const SCALAR d = X(op_)(A, lda, trans, j, j);
        for (size_t i = 0; i < m; i++) b[i] /= d;
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * TRMM2  X := alpha*op(A)*B (Left) or alpha*B*op(A) (Right), out of place
 * reference: blas/dtrmm.c:206-372.   TRMM (in place, B==X): blas/dtrmm.c:44-204.
 * One body serves both: the sweep order is chosen so that X may alias B.
 * ------------------------------------------------------------------------------------------ */
int X(trmm2)(int side, int uplo, int trans, int diag, size_t m, size_t n, SCALAR alpha,
             const SCALAR *A, size_t lda, const SCALAR *B, size_t ldb, SCALAR *Xo, size_t ldx) {
  const size_t nRowA = (side == 'L') ? m : n;
  if (lda < nRowA) return 9;
  if (ldb < m) return 11;
  if (ldx < m) return 13;
  if (m == 0 || n == 0) return 0;
  if (alpha == 0) {
    for (size_t j = 0; j < n; j++)
      for (size_t i = 0; i < m; i++) EL(Xo, ldx, i, j) = 0;
    return 0;
  }
  const int lower = (uplo == 'L') == (trans == 'N');
  if (side == 'L') {
    /* x_i = sum_l T(i,l) b_l ; lower: l<=i -> go i downwards so b_l (l<i) are still inputs */
#pragma omp parallel for schedule(static)
    for (size_t j = 0; j < n; j++) {
      const SCALAR *b = &EL(B, ldb, 0, j);
      SCALAR *x = &EL(Xo, ldx, 0, j);
      for (size_t ii = 0; ii < m; ii++) {
        const size_t i = lower ? m - 1 - ii : ii;
        SCALAR t = (diag == 'N') ? X(op_)(A, lda, trans, i, i) * b[i] : b[i];
        const size_t l0 = lower ? 0 : i + 1, l1 = lower ? i : m;
        for (size_t l = l0; l < l1; l++) t += X(op_)(A, lda, trans, i, l) * b[l];
        x[i] = alpha * t;
      }
    }
  } else {
    /* X[:,j] = sum_l B[:,l] T(l,j) ; T upper: l<=j -> go j downwards */
    for (size_t jj = 0; jj < n; jj++) {
      const size_t j = lower ? jj : n - 1 - jj;
      SCALAR *x = &EL(Xo, ldx, 0, j);
      const SCALAR *bj = &EL(B, ldb, 0, j);
      const SCALAR d = (diag == 'N') ? alpha * X(op_)(A, lda, trans, j, j) : alpha;
      for (size_t i = 0; i < m; i++) x[i] = d * bj[i];
      const size_t l0 = lower ? j + 1 : 0, l1 = lower ? n : j;
      for (size_t l = l0; l < l1; l++) {
        const SCALAR t = alpha * X(op_)(A, lda, trans, l, j);
        if (t == 0) continue;
        const SCALAR *bl = &EL(B, ldb, 0, l);
        for (size_t i = 0; i < m; i++) x[i] += t * bl[i];
      }
    }
  }
  return 0;
}

int X(trmm)(int side, int uplo, int trans, int diag, size_t m, size_t n, SCALAR alpha,
            const SCALAR *A, size_t lda, SCALAR *B, size_t ldb) {
  return X(trmm2)(side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, B, ldb);
}

/* ------------------------------------------------------------------------------------------
 * POTF2  unblocked Cholesky of an n x n block.   reference: lapack/dpotrf.c:34-81, zpotrf.c:35-83
 * Upper: row-by-row dot products; on a non-positive pivot the reduced pivot IS stored before
 * returning (dpotrf.c:46-49).  Lower: right-to-left column updates; pivot not stored on failure
 * (dpotrf.c:69-73).  info = 1-based failing column.
 * ------------------------------------------------------------------------------------------ */
static void X(potf2_)(int uplo, size_t n, SCALAR *A, size_t lda, long *info) {
  if (uplo == 'U') {
    for (size_t i = 0; i < n; i++) {
      REAL s = 0;
      for (size_t k = 0; k < i; k++) s += RE(EL(A, lda, k, i) * CONJ(EL(A, lda, k, i)));
      REAL aii = RE(EL(A, lda, i, i)) - s;
      if (aii <= 0 || isnan(aii)) {
        EL(A, lda, i, i) = aii;
        *info = (long)i + 1;
        return;
      }
      aii = (REAL)sqrt((double)aii);
      EL(A, lda, i, i) = aii;
      for (size_t j = i + 1; j < n; j++) {
        SCALAR t = 0;
        for (size_t k = 0; k < i; k++) t += EL(A, lda, k, j) * CONJ(EL(A, lda, k, i));
        EL(A, lda, i, j) = (EL(A, lda, i, j) - t) / aii;
      }
    }
  } else {
    for (size_t j = 0; j < n; j++) {
      for (size_t k = 0; k < j; k++) {
        const SCALAR t = CONJ(EL(A, lda, j, k));
        for (size_t i = j; i < n; i++) EL(A, lda, i, j) -= t * EL(A, lda, i, k);
      }
      REAL ajj = RE(EL(A, lda, j, j));
      if (ajj <= 0 || isnan(ajj)) {
        *info = (long)j + 1;
        return;
      }
      ajj = (REAL)sqrt((double)ajj);
      EL(A, lda, j, j) = ajj;
      for (size_t i = j + 1; i < n; i++) EL(A, lda, i, j) /= ajj;
    }
  }
}

#if IS_COMPLEX
#define RANKK X(herk)
#define HTRANS 'C'
#else
#define RANKK X(syrk)
#define HTRANS 'T'
#endif

/* ------------------------------------------------------------------------------------------
 * POTRF  blocked, LEFT-looking, nb = 16 (upper) / 32 (lower).  reference: lapack/dpotrf.c:83-146
 * per block column j:  rank-k update of the diagonal block with everything to its left/above,
 * potf2, GEMM update of the panel, TRSM with the fresh diagonal block.
 * info<0: argument (lda -> -4); info>0: first non-positive-definite column (1-based).
 * ------------------------------------------------------------------------------------------ */
void X(potrf)(int uplo, size_t n, SCALAR *A, size_t lda, long *info) {
  *info = 0;
  if (lda < n) { *info = -4; return; }
  if (n == 0) return;
  const size_t nb = (uplo == 'U') ? 16 : 32;
  if (nb > n) { X(potf2_)(uplo, n, A, lda, info); return; }

  for (size_t j = 0; j < n; j += nb) {
    const size_t jb = X(min_)(nb, n - j), rest = n - j - jb;
    if (uplo == 'U') {
      RANKK('U', HTRANS, jb, j, (REAL)-1, &EL(A, lda, 0, j), lda, (REAL)1, &EL(A, lda, j, j), lda);
      X(potf2_)('U', jb, &EL(A, lda, j, j), lda, info);
      if (*info != 0) { *info += (long)j; return; }
      if (rest) {
        X(gemm)(HTRANS, 'N', jb, rest, j, (SCALAR)-1, &EL(A, lda, 0, j), lda, &EL(A, lda, 0, j + jb),
                lda, (SCALAR)1, &EL(A, lda, j, j + jb), lda);
        X(trsm)('L', 'U', HTRANS, 'N', jb, rest, (SCALAR)1, &EL(A, lda, j, j), lda,
                &EL(A, lda, j, j + jb), lda);
      }
    } else {
      RANKK('L', 'N', jb, j, (REAL)-1, &EL(A, lda, j, 0), lda, (REAL)1, &EL(A, lda, j, j), lda);
      X(potf2_)('L', jb, &EL(A, lda, j, j), lda, info);
      if (*info != 0) { *info += (long)j; return; }
      if (rest) {
        X(gemm)('N', HTRANS, rest, jb, j, (SCALAR)-1, &EL(A, lda, j + jb, 0), lda, &EL(A, lda, j, 0),
                lda, (SCALAR)1, &EL(A, lda, j + jb, j), lda);
        X(trsm)('R', 'L', HTRANS, 'N', rest, jb, (SCALAR)1, &EL(A, lda, j, j), lda,
                &EL(A, lda, j + jb, j), lda);
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * TRTI2  unblocked triangular inverse, in place.      reference: lapack/dtrtri.c:43-96
 * Upper: columns left to right; column j of the inverse = -(U00^-1 * u_j) / u_jj where U00^-1
 * is already in place.  Lower: columns right to left.  A zero pivot gives info = j+1.
 * ------------------------------------------------------------------------------------------ */
static void X(trti2_)(int uplo, int diag, size_t n, SCALAR *A, size_t lda, long *info) {
  for (size_t jj = 0; jj < n; jj++) {
    const size_t j = (uplo == 'U') ? jj : n - 1 - jj;
    SCALAR ajj;
    if (diag == 'N') {
      if (EL(A, lda, j, j) == 0) { *info = (long)j + 1; return; }
      EL(A, lda, j, j) = (SCALAR)1 / EL(A, lda, j, j);
      ajj = -EL(A, lda, j, j);
    } else
      ajj = -1;
    if (uplo == 'U') {
      /* x := U00inv * x for x = A[0:j, j]  (upper-triangular matrix-vector product, ascending) */
      for (size_t k = 0; k < j; k++) {
        const SCALAR t = EL(A, lda, k, j);
        if (diag == 'N') EL(A, lda, k, j) *= EL(A, lda, k, k);
        for (size_t i = 0; i < k; i++) EL(A, lda, i, j) += t * EL(A, lda, i, k);
      }
      for (size_t i = 0; i < j; i++) EL(A, lda, i, j) *= ajj;
    } else {
      for (size_t i = n - 1; i > j; i--) {
        const SCALAR t = EL(A, lda, i, j);
        if (diag == 'N') EL(A, lda, i, j) *= EL(A, lda, i, i);
        for (size_t k = i + 1; k < n; k++) EL(A, lda, k, j) += t * EL(A, lda, k, i);
      }
      for (size_t i = j + 1; i < n; i++) EL(A, lda, i, j) *= ajj;
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * TRTRI  blocked triangular inverse in place.          reference: lapack/dtrtri.c:98-164
 * nb = 64 (d,c,z), 32/64 upper/lower (s: lapack/strtri.c:113).  Upper: block columns ascending,
 * TRMM(Left,Upper) with the already-inverted leading block, TRSM(Right, alpha=-1) with the
 * not-yet-inverted diagonal block, then trti2 of the diagonal block.  Lower: mirror, descending.
 * info<0: argument (lda -> -5); info>0: zero pivot (1-based).
 * ------------------------------------------------------------------------------------------ */
void X(trtri)(int uplo, int diag, size_t n, SCALAR *A, size_t lda, long *info) {
  *info = 0;
  if (lda < n) { *info = -5; return; }
  if (n == 0) return;
  const size_t nb = (sizeof(SCALAR) == sizeof(float)) ? ((uplo == 'U') ? 32 : 64) : 64;
  if (n < nb) { X(trti2_)(uplo, diag, n, A, lda, info); return; }

  if (uplo == 'U') {
    for (size_t j = 0; j < n; j += nb) {
      const size_t jb = X(min_)(nb, n - j);
      X(trmm)('L', 'U', 'N', diag, j, jb, (SCALAR)1, A, lda, &EL(A, lda, 0, j), lda);
      X(trsm)('R', 'U', 'N', diag, j, jb, (SCALAR)-1, &EL(A, lda, j, j), lda, &EL(A, lda, 0, j), lda);
      X(trti2_)('U', diag, jb, &EL(A, lda, j, j), lda, info);
      if (*info != 0) { *info += (long)j; return; }
    }
  } else {
    size_t j = ((n + nb - 1) / nb) * nb;
    do {
      j -= nb;
      const size_t jb = X(min_)(nb, n - j), rest = n - j - jb;
      if (rest) {
        X(trmm)('L', 'L', 'N', diag, rest, jb, (SCALAR)1, &EL(A, lda, j + jb, j + jb), lda,
                &EL(A, lda, j + jb, j), lda);
        X(trsm)('R', 'L', 'N', diag, rest, jb, (SCALAR)-1, &EL(A, lda, j, j), lda,
                &EL(A, lda, j + jb, j), lda);
      }
      X(trti2_)('L', diag, jb, &EL(A, lda, j, j), lda, info);
      if (*info != 0) { *info += (long)j; return; }
    } while (j > 0);
  }
}

/* TRTRI2  out of place: B_tri := inverse(A_tri).  reference: lapack/dtrtri.c:226-299 (blocked
 * with trmm2); restated as "copy the triangle, invert in place" - same arithmetic per entry,
 * the other triangle of B is not touched, A is not modified.  info: lda -> -5, ldb -> -7. */
void X(trtri2)(int uplo, int diag, size_t n, const SCALAR *A, size_t lda, SCALAR *B, size_t ldb,
               long *info) {
  *info = 0;
  if (lda < n) *info = -5;
  if (ldb < n) *info = -7;
  if (*info != 0 || n == 0) return;
  for (size_t j = 0; j < n; j++) {
    const size_t i0 = (uplo == 'U') ? 0 : j, i1 = (uplo == 'U') ? j + 1 : n;
    for (size_t i = i0; i < i1; i++) EL(B, ldb, i, j) = EL(A, lda, i, j);
  }
  X(trtri)(uplo, diag, n, B, ldb, info);
}

/* ------------------------------------------------------------------------------------------
 * LAUU2 / LAUUM   A := U*U^H (upper) or L^H*L (lower).  reference: lapack/dlauum.c:43-69, :71-129
 * nb = 16 upper / 32 lower.  Per diagonal block i: TRMM of the off-diagonal strip with the
 * diagonal block, lauu2 of the diagonal block, then GEMM + rank-k with the strip beyond it.
 * ------------------------------------------------------------------------------------------ */
static void X(lauu2_)(int uplo, size_t n, SCALAR *A, size_t lda) {
  if (uplo == 'U') {
    for (size_t j = 0; j < n; j++) {
      const SCALAR ajj = CONJ(EL(A, lda, j, j));
      for (size_t i = 0; i <= j; i++) EL(A, lda, i, j) *= ajj;
      for (size_t k = j + 1; k < n; k++) {
        const SCALAR t = CONJ(EL(A, lda, j, k));
        for (size_t i = 0; i <= j; i++) EL(A, lda, i, j) += t * EL(A, lda, i, k);
      }
    }
  } else {
    for (size_t j = 0; j < n; j++)
      for (size_t i = j; i < n; i++) {
        SCALAR t = EL(A, lda, i, j) * CONJ(EL(A, lda, i, i));
        for (size_t k = i + 1; k < n; k++) t += CONJ(EL(A, lda, k, i)) * EL(A, lda, k, j);
        EL(A, lda, i, j) = t;
      }
  }
}

void X(lauum)(int uplo, size_t n, SCALAR *A, size_t lda, long *info) {
  *info = 0;
  if (lda < n) { *info = -4; return; }
  if (n == 0) return;
  const size_t nb = (uplo == 'U') ? 16 : 32;
  if (nb > n) { X(lauu2_)(uplo, n, A, lda); return; }
  for (size_t i = 0; i < n; i += nb) {
    const size_t ib = X(min_)(nb, n - i), rest = n - i - ib;
    if (uplo == 'U') {
      X(trmm)('R', 'U', HTRANS, 'N', i, ib, (SCALAR)1, &EL(A, lda, i, i), lda, &EL(A, lda, 0, i), lda);
      X(lauu2_)('U', ib, &EL(A, lda, i, i), lda);
      if (rest) {
        X(gemm)('N', HTRANS, i, ib, rest, (SCALAR)1, &EL(A, lda, 0, i + ib), lda,
                &EL(A, lda, i, i + ib), lda, (SCALAR)1, &EL(A, lda, 0, i), lda);
        RANKK('U', 'N', ib, rest, (REAL)1, &EL(A, lda, i, i + ib), lda, (REAL)1, &EL(A, lda, i, i), lda);
      }
    } else {
      X(trmm)('L', 'L', HTRANS, 'N', ib, i, (SCALAR)1, &EL(A, lda, i, i), lda, &EL(A, lda, i, 0), lda);
      X(lauu2_)('L', ib, &EL(A, lda, i, i), lda);
      if (rest) {
        X(gemm)(HTRANS, 'N', ib, i, rest, (SCALAR)1, &EL(A, lda, i + ib, i), lda,
                &EL(A, lda, i + ib, 0), lda, (SCALAR)1, &EL(A, lda, i, 0), lda);
        RANKK('L', HTRANS, ib, rest, (REAL)1, &EL(A, lda, i + ib, i), lda, (REAL)1, &EL(A, lda, i, i), lda);
      }
    }
  }
}

/* POTRI  inverse of an SPD matrix GIVEN ITS CHOLESKY FACTOR: trtri(NonUnit) then lauum.
 * reference: lapack/dpotri.c:4-23 */
void X(potri)(int uplo, size_t n, SCALAR *A, size_t lda, long *info) {
  *info = 0;
  if (lda < n) { *info = -4; return; }
  if (n == 0) return;
  X(trtri)(uplo, 'N', n, A, lda, info);
  if (*info != 0) return;
  X(lauum)(uplo, n, A, lda, info);
}

/* LOGDET  2 * sum_i log(Re x[i*incx]).  reference: lapack/dlogdet.c:10-25; the complex
 * versions use the SINGLE precision log of the real part even for double complex
 * (lapack/zlogdet.c:18,22) - mirrored here so the pin against _ref is exact. */
REAL X(logdet)(const SCALAR *x, size_t incx, size_t n) {
  REAL total = 0;
  for (size_t i = 0; i < n; i++) {
#if IS_COMPLEX
    total += logf((float)RE(x[i * incx]));
#else
    total += (sizeof(REAL) == sizeof(float)) ? (REAL)logf((float)x[i * incx]) : (REAL)log((double)x[i * incx]);
#endif
  }
  return (REAL)2 * total;
}

#undef RANKK
#undef HTRANS
#undef CONJ
#undef RE
#undef EL
