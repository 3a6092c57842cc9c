/* chol_oracle.h — TEST INFRASTRUCTURE ONLY: prototypes of the CPU oracle (libchol_oracle.so).
 * oracle_<p><routine> restates the reference's CPU <p><routine> (include/lapack.h, include/blas.h);
 * enums are passed as their char values.  BLAS-level routines return the reference's XERBLA
 * argument position (0 = ok) instead of calling a hook. */
#ifndef CHOL_ORACLE_H
#define CHOL_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
void oracle_dpotrf(int uplo, size_t n, double *A, size_t lda, long *info);
void oracle_dtrtri(int uplo, int diag, size_t n, double *A, size_t lda, long *info);
void oracle_dlauum(int uplo, size_t n, double *A, size_t lda, long *info);
void oracle_dpotri(int uplo, size_t n, double *A, size_t lda, long *info);
double oracle_dlogdet(const double *x, size_t incx, size_t n);
int oracle_dgemm(int transA, int transB, size_t m, size_t n, size_t k, double alpha, const double *A,
                 size_t lda, const double *B, size_t ldb, double beta, double *C, size_t ldc);
/* ... s/c/z variants follow the same pattern; tests bind them through ctypes by name. */
size_t oracle_dsfmt_state_size(void);
void oracle_dsfmt_seed(void *state, uint64_t seed);
void oracle_dsfmt_fill_close_open(void *state, double *out, size_t n);
#ifdef __cplusplus
}
#endif
#endif
