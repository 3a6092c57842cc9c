// mock_kernels.cpp — TEST INFRASTRUCTURE ONLY.  The launcher ABI of csrc/chol_kernels.h implemented
// as plain CPU loops that follow the documented semantics of each launcher (not the kernels' code):
// with tests/mock/mock_cuda.c this lets the host drivers run without a GPU
// (tests/test_host_drivers_mock.py).  A "launch" is queued on its stream (mock_cuda.c runs every
// stream on its own thread, in order), so launches are asynchronous as on the device.
#include <cmath>
#include <complex>
#include <cstring>
#include <functional>
#include <vector>

#include <cstdint>
#include "../../cuda-cholesky_b200/csrc/chol_kernels.h"

extern "C" void mock_stream_enqueue(void *stream, void (*fn)(void *), void *arg);  // mock_cuda.c

namespace {

unsigned long long launches = 0;
void run_fn(void *p) {
  auto *f = static_cast<std::function<void()> *>(p);
  (*f)();
  delete f;
}
template <typename F> int enqueue(void *stream, F f) {
  mock_stream_enqueue(stream, run_fn, new std::function<void()>(std::move(f)));
  return 0;
}
template <typename T> struct RealOf { using type = T; };
template <typename R> struct RealOf<std::complex<R>> { using type = R; };
template <typename T> T cj(T x) { return x; }
template <typename R> std::complex<R> cj(std::complex<R> x) { return std::conj(x); }
template <typename T> double re(T x) { return (double)std::real(x); }
template <typename T> T from_re(double r) { return T((typename RealOf<T>::type)r); }

long long gcol(const chol_gemm_desc *d, long long n) {
  return d->cyc_nb ? ((n / d->cyc_nb) * d->cyc_p + d->cyc_off) * d->cyc_nb + n % d->cyc_nb : n;
}

template <typename T>
int gemm_t(const chol_gemm_desc *d) {
  constexpr bool CPLX = !std::is_same<typename RealOf<T>::type, T>::value;
  using W = typename std::conditional<CPLX, std::complex<double>, double>::type;  // accumulate in double
  const T *A = (const T *)d->A, *B = (const T *)d->B, *C = (const T *)d->C;
  T *D = (T *)d->D;
  const size_t M = d->m, N = d->n, K = d->k;
  W alpha, beta;
  if constexpr (CPLX) { alpha = W(d->alpha[0], d->alpha[1]); beta = W(d->beta[0], d->beta[1]); }
  else { alpha = d->alpha[0]; beta = d->beta[0]; }
  const bool alpha_zero = alpha == W(0), beta_zero = beta == W(0);
  if (d->cyc_nb && (d->opB == 'N' || d->cuplo == 'U' || d->triA || d->triB)) return 1;
  // op(A) as a dense M x K array and op(B) as K x N (N = local columns; the B row of local column n is
  // its GLOBAL column), masks and implicit unit diagonals applied: then the product is a plain loop nest
  std::vector<W> Aop(M * K), Bop(K * N);
  for (size_t k = 0; k < K && !alpha_zero; ++k)
    for (size_t m = 0; m < M; ++m) {
      const size_t r = d->opA == 'N' ? m : k, c = d->opA == 'N' ? k : m;
      W v;
      if ((d->triA == 'L' && r < c) || (d->triA == 'U' && r > c)) v = W(0);
      else if (d->triA && d->unitA && r == c) v = W(1);
      else v = W(A[r + c * d->lda]);
      if constexpr (CPLX) { if (d->opA == 'C') v = std::conj(v); }
      Aop[m + k * M] = v;
    }
  for (size_t n = 0; n < N && !alpha_zero; ++n) {
    const size_t ng = (size_t)gcol(d, (long long)n);
    for (size_t k = 0; k < K; ++k) {
      const size_t r = d->opB == 'N' ? k : ng, c = d->opB == 'N' ? ng : k;
      W v;
      if ((d->triB == 'L' && r < c) || (d->triB == 'U' && r > c)) v = W(0);
      else if (d->triB && d->unitB && r == c) v = W(1);
      else v = W(B[r + c * d->ldb]);
      if constexpr (CPLX) { if (d->opB == 'C') v = std::conj(v); }
      Bop[k + n * K] = v;
    }
  }
  // every output element depends on C(m,n) alone, so C == D is fine; operands may alias D (the drivers
  // route such products through scratch, but a spec implementation should not care): finish the whole
  // product before the first store
  std::vector<W> acc(M * N, W(0));
  if (!alpha_zero)
    for (size_t n = 0; n < N; ++n)
      for (size_t k = 0; k < K; ++k) {
        const W b = Bop[k + n * K];
        if (b == W(0)) continue;
        W *an = &acc[n * M];
        const W *ak = &Aop[k * M];
        for (size_t m = 0; m < M; ++m) an[m] += ak[m] * b;
      }
  for (size_t n = 0; n < N; ++n) {
    const size_t ng = (size_t)gcol(d, (long long)n);
    for (size_t m = 0; m < M; ++m) {
      if ((d->cuplo == 'L' && m < ng) || (d->cuplo == 'U' && m > ng)) continue;
      W o = alpha * acc[m + n * M];
      if (!beta_zero) o += beta * W(C[m + n * d->ldc]);
      if constexpr (CPLX) {
        if (d->herm && m == ng) o = W(o.real(), 0.0);
        D[m + n * d->ldd] = T((typename RealOf<T>::type)o.real(), (typename RealOf<T>::type)o.imag());
      } else {
        D[m + n * d->ldd] = (T)o;
      }
    }
  }
  return 0;
}

// lower-form view of a triangle stored as uplo: L(i,c) for i >= c
template <typename T> T getL(const T *A, size_t lda, bool upper, size_t i, size_t c) { return upper ? cj(A[c + i * lda]) : A[i + c * lda]; }
template <typename T> void putL(T *A, size_t lda, bool upper, size_t i, size_t c, T v) { if (upper) A[c + i * lda] = cj(v); else A[i + c * lda] = v; }

template <typename T>
int potf2_t(bool upper, size_t n, T *A, size_t lda, long *dinfo, long base, T *X, size_t ldx) {
  if (*dinfo != 0) return 0;
  std::vector<T> L(n * n, T(0));
  for (size_t c = 0; c < n; ++c)
    for (size_t i = c; i < n; ++i) L[i + c * n] = getL(A, lda, upper, i, c);
  for (size_t c = 0; c < n; ++c) {
    const double dd = re(L[c + c * n]);
    if (!(dd > 0)) {
      *dinfo = base + (long)c + 1;
      A[c + c * lda] = from_re<T>(dd);
      return 0;  // the columns finished so far have been written, nothing else is touched
    }
    const double r = std::sqrt(dd);
    L[c + c * n] = from_re<T>(r);
    putL(A, lda, upper, c, c, L[c + c * n]);
    for (size_t i = c + 1; i < n; ++i) {
      L[i + c * n] = L[i + c * n] / from_re<T>(r);
      putL(A, lda, upper, i, c, L[i + c * n]);
    }
    for (size_t k = c + 1; k < n; ++k)
      for (size_t i = k; i < n; ++i) L[i + k * n] -= L[i + c * n] * cj(L[k + c * n]);
  }
  if (X != nullptr) {  // X := inverse of the factor, uplo triangle, other triangle zero
    for (size_t j = 0; j < n; ++j) {
      std::vector<T> x(n, T(0));
      x[j] = T(1) / L[j + j * n];
      for (size_t i = j + 1; i < n; ++i) {
        T s = 0;
        for (size_t k = j; k < i; ++k) s += L[i + k * n] * x[k];
        x[i] = -s / L[i + i * n];
      }
      for (size_t i = 0; i < n; ++i) {
        const T v = i >= j ? x[i] : T(0);
        if (!upper) X[i + j * ldx] = v; else X[j + i * ldx] = cj(v);
      }
    }
  }
  return 0;
}

template <typename T>
int trti2_t(bool upper, bool unit, size_t n, const T *A, size_t lda, T *B, size_t ldb, int zero_other, long *dinfo, long base) {
  const int info_mode = zero_other & 6;  // csrc/chol_kernels.h: 0 sticky long, 2 self-zeroed int, 4 none (Inf on a zero pivot)
  zero_other &= 1;
  if (info_mode == 0 && *dinfo != 0) return 0;
  if (info_mode == 2) *reinterpret_cast<int *>(dinfo) = 0;
  std::vector<T> L(n * n, T(0));  // lower form WITHOUT conjugation: inv(U) = (inv(U^T))^T
  for (size_t c = 0; c < n; ++c)
    for (size_t i = c; i < n; ++i) L[i + c * n] = upper ? A[c + i * lda] : A[i + c * lda];
  if (!unit && info_mode != 4) {
    long lo = -1, hi = -1;
    for (size_t i = 0; i < n; ++i)
      if (L[i + i * n] == T(0)) { if (lo < 0) lo = (long)i; hi = (long)i; }
    if (hi >= 0) {  // the reference's sweep direction
      const long v = base + (upper ? lo : hi) + 1;
      if (info_mode == 2) *reinterpret_cast<int *>(dinfo) = (int)v; else *dinfo = v;
      return 0;
    }
  }
  for (size_t j = 0; j < n; ++j) {
    std::vector<T> x(n, T(0));
    x[j] = unit ? T(1) : T(1) / L[j + j * n];
    for (size_t i = j + 1; i < n; ++i) {
      T s = 0;
      for (size_t k = j; k < i; ++k) s += L[i + k * n] * x[k];
      x[i] = unit ? -s : -s / L[i + i * n];
    }
    for (size_t i = 0; i < n; ++i) {
      if (i >= j) { if (!upper) B[i + j * ldb] = x[i]; else B[j + i * ldb] = x[i]; }
      else if (zero_other) { if (!upper) B[i + j * ldb] = T(0); else B[j + i * ldb] = T(0); }
    }
  }
  return 0;
}

template <typename T>
int lauu2_t(bool upper, size_t n, T *A, size_t lda) {
  std::vector<T> L(n * n, T(0));  // lower form with conjugation: upper U -> L = U^H, U U^H = L^H L
  for (size_t c = 0; c < n; ++c)
    for (size_t i = c; i < n; ++i) L[i + c * n] = getL(A, lda, upper, i, c);
  for (size_t b = 0; b < n; ++b)
    for (size_t a = b; a < n; ++a) {
      T acc = 0;
      for (size_t k = a; k < n; ++k) acc += cj(L[k + a * n]) * L[k + b * n];
      if (!upper) A[a + b * lda] = acc; else A[b + a * lda] = cj(acc);
    }
  return 0;
}

// X = B * L^-H by substitution from the documented semantics (csrc/chol_kernels.h): element (row, j) of B / X sits
// at ptr[j * ld + row], or - transposed forms - conjugated at ptr[row * ld + j]
template <typename T>
int trsm_rows_t(bool upper, size_t n, const T *L, size_t ldl, size_t m, int transB, const T *B, size_t ldb, int transX, T *X,
                       size_t ldx, const long *dinfo) {
  if (dinfo != nullptr && *dinfo != 0) return 0;
  std::vector<T> x(n);
  for (size_t r = 0; r < m; ++r) {
    for (size_t j = 0; j < n; ++j) x[j] = transB ? cj(B[r * ldb + j]) : B[j * ldb + r];
    for (size_t k = 0; k < n; ++k) {
      T s = x[k];
      for (size_t j = 0; j < k; ++j) s -= x[j] * cj(getL(L, ldl, upper, k, j));  // conj(L(k,j)), j < k
      x[k] = s / getL(L, ldl, upper, k, k);
    }
    for (size_t j = 0; j < n; ++j) { if (transX) X[r * ldx + j] = cj(x[j]); else X[j * ldx + r] = x[j]; }
  }
  return 0;
}

bool keep(int uplo, size_t r, size_t c) {
  return uplo == 0 || (uplo == 'L' && r >= c) || (uplo == 'U' && r <= c) || (uplo == 'l' && r > c) || (uplo == 'u' && r < c);
}
template <typename T>
int copy_t(int trans, int uplo, size_t m, size_t n, const T *A, size_t lda, T *B, size_t ldb) {
  for (size_t c = 0; c < n; ++c)
    for (size_t r = 0; r < m; ++r) {
      if (!keep(uplo, r, c)) continue;
      if (trans == 'N') B[r + c * ldb] = A[r + c * lda];
      else B[c + r * ldb] = trans == 'C' ? cj(A[r + c * lda]) : A[r + c * lda];
    }
  return 0;
}
template <typename T>
int fill_t(int uplo, size_t m, size_t n, T v, T *A, size_t lda) {
  for (size_t c = 0; c < n; ++c)
    for (size_t r = 0; r < m; ++r)
      if (uplo == 0 || (uplo == 'L' && r > c) || (uplo == 'U' && r < c)) A[r + c * lda] = v;
  return 0;
}

}  // namespace

#define DISPATCH(prec, CALL)                                           \
  switch (prec) {                                                      \
    case 's': { using T = float; CALL; break; }                        \
    case 'd': { using T = double; CALL; break; }                       \
    case 'c': { using T = std::complex<float>; CALL; break; }          \
    case 'z': { using T = std::complex<double>; CALL; break; }         \
  }
static bool known(int prec) { return prec == 's' || prec == 'd' || prec == 'c' || prec == 'z'; }

extern "C" {
unsigned long long chol_launch_count(int reset) { const unsigned long long v = launches; if (reset) launches = 0; return v; }
int chol_diag_max(int prec) { return prec == 'z' ? 64 : 128; }
int chol_gemm_uses_tma(const chol_gemm_desc *) { return 0; }
int chol_probe_fp64(int, int, double *tflops, float *ms) { *tflops = 1.0; if (ms) *ms = 1.0f; return 0; }
size_t chol_logdet_scratch_bytes(void) { return 4096; }

int chol_gemm_launch(const chol_gemm_desc *dp) {
  __sync_fetch_and_add(&launches, 1);
  if (dp->m == 0 || dp->n == 0) return 0;
  if (!known(dp->prec)) return 1;
  if (dp->cyc_nb && (dp->opB == 'N' || dp->cuplo == 'U' || dp->triA || dp->triB)) return 1;
  const chol_gemm_desc d = *dp;
  return enqueue(d.stream, [d] { DISPATCH(d.prec, gemm_t<T>(&d)) });
}
int chol_potf2_logdet_launch(int prec, int uplo, size_t n, void *A, size_t lda, long *dinfo, long base, double *logacc, void *st) {
  __sync_fetch_and_add(&launches, 1);
  if (n == 0) return 0;
  if (!known(prec) || n > (size_t)chol_diag_max(prec)) return 1;
  return enqueue(st, [=] {
    if (*dinfo != 0) return;
    DISPATCH(prec, potf2_t<T>(uplo == 'U', n, (T *)A, lda, dinfo, base, (T *)nullptr, 0))
    if (*dinfo == 0 && logacc != nullptr) {
      double s = 0.0;
      for (size_t i = 0; i < n; ++i) {
        switch (prec) {
          case 's': s += std::log((double)((const float *)A)[i * lda + i]); break;
          case 'd': s += std::log(((const double *)A)[i * lda + i]); break;
          case 'c': s += std::log((double)((const std::complex<float> *)A)[i * lda + i].real()); break;
          default: s += std::log(((const std::complex<double> *)A)[i * lda + i].real()); break;
        }
      }
      *logacc += 2.0 * s;
    }
  });
}
int chol_potf2_launch(int prec, int uplo, size_t n, void *A, size_t lda, long *dinfo, long base, void *st) {
  __sync_fetch_and_add(&launches, 1);
  if (n == 0) return 0;
  if (!known(prec) || n > (size_t)chol_diag_max(prec)) return 1;
  return enqueue(st, [=] { DISPATCH(prec, potf2_t<T>(uplo == 'U', n, (T *)A, lda, dinfo, base, (T *)nullptr, 0)) });
}
int chol_trsm_rows_launch(int prec, int uplo, size_t n, const void *L, size_t ldl, size_t m, int transB, const void *B, size_t ldb,
                          int transX, void *X, size_t ldx, const long *dinfo, void *st) {
  __sync_fetch_and_add(&launches, 1);
  if (n == 0 || m == 0) return 0;
  if (!known(prec) || n > (size_t)chol_diag_max(prec)) return 1;
  return enqueue(st, [=] { DISPATCH(prec, trsm_rows_t<T>(uplo == 'U', n, (const T *)L, ldl, m, transB, (const T *)B, ldb, transX, (T *)X, ldx, dinfo)) });
}
int chol_trti2_launch(int prec, int uplo, int diag, size_t n, const void *A, size_t lda, void *B, size_t ldb, int zero_other, long *dinfo,
                      long base, void *st) {
  __sync_fetch_and_add(&launches, 1);
  if (n == 0) return 0;
  if (!known(prec) || n > (size_t)chol_diag_max(prec)) return 1;
  return enqueue(st, [=] { DISPATCH(prec, trti2_t<T>(uplo == 'U', diag == 'U', n, (const T *)A, lda, (T *)B, ldb, zero_other, dinfo, base)) });
}
int chol_lauu2_launch(int prec, int uplo, size_t n, void *A, size_t lda, void *st) {
  __sync_fetch_and_add(&launches, 1);
  if (n == 0) return 0;
  if (!known(prec) || n > (size_t)chol_diag_max(prec)) return 1;
  return enqueue(st, [=] { DISPATCH(prec, lauu2_t<T>(uplo == 'U', n, (T *)A, lda)) });
}
int chol_copy_launch(int prec, int trans, int uplo, size_t m, size_t n, const void *A, size_t lda, void *B, size_t ldb, void *st) {
  __sync_fetch_and_add(&launches, 1);
  if (m == 0 || n == 0) return 0;
  if (!known(prec)) return 1;
  return enqueue(st, [=] { DISPATCH(prec, copy_t<T>(trans, uplo, m, n, (const T *)A, lda, (T *)B, ldb)) });
}
int chol_scatter_rows_launch(int prec, size_t m, size_t n, const void *A, size_t lda, void *B, size_t ldb, int cnb, int cp, int coff, void *st) {
  __sync_fetch_and_add(&launches, 1);
  if (m == 0 || n == 0) return 0;
  if (!known(prec) || cnb <= 0 || cp <= 0) return 1;
  const size_t es = prec == 's' ? 4 : (prec == 'z' ? 16 : 8);
  return enqueue(st, [=] {
    for (size_t c = 0; c < n; ++c)
      for (size_t r = 0; r < m; ++r) {
        const size_t gr = ((r / (size_t)cnb) * (size_t)cp + (size_t)coff) * (size_t)cnb + r % (size_t)cnb;
        std::memcpy((char *)B + (c * ldb + gr) * es, (const char *)A + (c * lda + r) * es, es);
      }
  });
}
int chol_fill_launch(int prec, int uplo, size_t m, size_t n, double re_, double im_, void *A, size_t lda, void *st) {
  __sync_fetch_and_add(&launches, 1);
  if (m == 0 || n == 0) return 0;
  if (!known(prec)) return 1;
  return enqueue(st, [=] {
    switch (prec) {
      case 's': fill_t<float>(uplo, m, n, (float)re_, (float *)A, lda); break;
      case 'd': fill_t<double>(uplo, m, n, re_, (double *)A, lda); break;
      case 'c': fill_t<std::complex<float>>(uplo, m, n, std::complex<float>((float)re_, (float)im_), (std::complex<float> *)A, lda); break;
      case 'z': fill_t<std::complex<double>>(uplo, m, n, std::complex<double>(re_, im_), (std::complex<double> *)A, lda); break;
    }
  });
}
// the s-precision tensor-core engine (csrc/gemm_tf32x3.cu) from its documented semantics: the planes carry
// hi = x rounded to 11 significant bits, lo = x - hi; the product is hi*hi + hi*lo + lo*hi in FP32
size_t chol_tf32_kpitch(size_t k) { return (k + 31) / 32 * 32; }
static float mock_tf32(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xffffe000u;  // round to nearest (ties away), 10 explicit mantissa bits
  std::memcpy(&x, &u, 4);
  return x;
}
int chol_tf32_split_launch(int trans, size_t rows, size_t k, const void *A, size_t lda, void *hi, void *lo, size_t kp, void *st) {
  __sync_fetch_and_add(&launches, 1);
  return enqueue(st, [=] {
    const float *a = (const float *)A;
    float *h = (float *)hi, *l = (float *)lo;
    for (size_t r = 0; r < rows; ++r)
      for (size_t c = 0; c < kp; ++c) {
        const float x = c < k ? (trans ? a[r * lda + c] : a[c * lda + r]) : 0.f;
        h[r * kp + c] = mock_tf32(x);
        l[r * kp + c] = mock_tf32(x - h[r * kp + c]);
      }
  });
}
int chol_tf32x3_gemm_launch(size_t m, size_t n, size_t k, double alpha, const void *Ahi, const void *Alo, const void *Bhi, const void *Blo,
                            size_t kp, double beta, const void *C, size_t ldc, void *D, size_t ldd, int cuplo, void *st) {
  __sync_fetch_and_add(&launches, 1);
  return enqueue(st, [=] {
    const float *ah = (const float *)Ahi, *al = (const float *)Alo, *bh = (const float *)Bhi, *bl = (const float *)Blo;
    const float *c = (const float *)C;
    float *d = (float *)D;
    for (size_t j = 0; j < n; ++j)
      for (size_t i = 0; i < m; ++i) {
        if ((cuplo == 'L' && i < j) || (cuplo == 'U' && i > j)) continue;
        float acc = 0.f;
        for (size_t q = 0; q < k; ++q)
          acc += al[i * kp + q] * bh[j * kp + q] + ah[i * kp + q] * bl[j * kp + q] + ah[i * kp + q] * bh[j * kp + q];
        d[j * ldd + i] = (float)alpha * acc + (beta != 0.0 ? (float)beta * c[j * ldc + i] : 0.f);
      }
  });
}
// c: planes of the real embedding, from the documented layout (csrc/chol_kernels.h): plane row 2i = (Re, Im) pairs of
// W's row i, plane row 2i+1 = (Im, -Re) pairs; pitch kp floats
int chol_tf32_split_c_launch(size_t rows, size_t k, const void *W, size_t ldw, void *hi, void *lo, size_t kp, void *st) {
  __sync_fetch_and_add(&launches, 1);
  return enqueue(st, [=] {
    const std::complex<float> *w = (const std::complex<float> *)W;
    float *h = (float *)hi, *l = (float *)lo;
    for (size_t r = 0; r < rows; ++r)
      for (size_t c = 0; 2 * c < kp; ++c) {
        const std::complex<float> x = c < k ? w[c * ldw + r] : std::complex<float>(0.f, 0.f);
        const float v[2][2] = {{x.real(), x.imag()}, {x.imag(), -x.real()}};
        for (int o = 0; o < 2; ++o)
          for (int q = 0; q < 2; ++q) {
            const size_t at = (2 * r + o) * kp + 2 * c + q;
            h[at] = mock_tf32(v[o][q]);
            l[at] = mock_tf32(v[o][q] - h[at]);
          }
      }
  });
}
int chol_tf32x3_herk_launch(size_t m, size_t n, size_t k, double alpha, const void *Ahi, const void *Alo, const void *Bhi, const void *Blo,
                            size_t kp, double beta, const void *C, size_t ldc, void *D, size_t ldd, int cuplo, int herm, void *st) {
  __sync_fetch_and_add(&launches, 1);
  return enqueue(st, [=] {
    const float *ah = (const float *)Ahi, *al = (const float *)Alo, *bh = (const float *)Bhi, *bl = (const float *)Blo;
    const std::complex<float> *c = (const std::complex<float> *)C;
    std::complex<float> *d = (std::complex<float> *)D;
    for (size_t j = 0; j < n; ++j)
      for (size_t i = 0; i < m; ++i) {
        if ((cuplo == 'L' && i < j) || (cuplo == 'U' && i > j)) continue;
        float acc[2] = {0.f, 0.f};
        for (int o = 0; o < 2; ++o)
          for (size_t q = 0; q < 2 * k; ++q) {
            const size_t ia = (2 * i + o) * kp + q, ib = (2 * j) * kp + q;
            acc[o] += al[ia] * bh[ib] + ah[ia] * bl[ib] + ah[ia] * bh[ib];
          }
        std::complex<float> v = (float)alpha * std::complex<float>(acc[0], acc[1]);
        if (beta != 0.0) v += (float)beta * c[j * ldc + i];
        if (herm && i == j) v = std::complex<float>(v.real(), 0.f);
        d[j * ldd + i] = v;
      }
  });
}
int chol_logdet_launch(int prec, const void *x, size_t incx, size_t n, void *dresult, void *, void *st) {
  __sync_fetch_and_add(&launches, 1);
  if (!known(prec)) return 1;
  return enqueue(st, [=] {
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) {
      switch (prec) {
        case 's': s += (double)std::log(((const float *)x)[i * incx]); break;
        case 'd': s += std::log(((const double *)x)[i * incx]); break;
        case 'c': s += (double)std::log(((const std::complex<float> *)x)[i * incx].real()); break;
        case 'z': s += std::log(((const std::complex<double> *)x)[i * incx].real()); break;
      }
    }
    if (prec == 's' || prec == 'c') *(float *)dresult = (float)(2.0 * s); else *(double *)dresult = 2.0 * s;
  });
}

// ---- device dSFMT-19937 launchers (csrc/chol_kernels.h) as plain loops written from their documented semantics ----
// a state = 192 words of 128 bits (two uint64_t each), word 191 = the lung; one refill = the 191 steps of the generator
static void mock_refill(uint64_t (*st)[2]) {
  uint64_t l0 = st[191][0], l1 = st[191][1];
  for (int i = 0; i < 191; ++i) {
    const uint64_t *b = st[(i + 117) % 191];
    const uint64_t t0 = st[i][0], t1 = st[i][1];
    const uint64_t n0 = (t0 << 19) ^ (l1 >> 32) ^ (l1 << 32) ^ b[0], n1 = (t1 << 19) ^ (l0 >> 32) ^ (l0 << 32) ^ b[1];
    l0 = n0; l1 = n1;
    st[i][0] = (l0 >> 12) ^ (l0 & 0x000ffafffffffb3fULL) ^ t0;
    st[i][1] = (l1 >> 12) ^ (l1 & 0x000ffdfffc90fffdULL) ^ t1;
  }
  st[191][0] = l0; st[191][1] = l1;
}
static uint64_t mock_convert(uint64_t w, int kind) {
  double d, r;
  std::memcpy(&d, &w, 8);
  if (kind == 0) return w;
  if (kind == 1) r = d - 1.0;
  else if (kind == 2) r = 2.0 - d;
  else if (kind == 3) { const uint64_t v = w | 1u; std::memcpy(&d, &v, 8); r = d - 1.0; }
  else { volatile double p = (double)(w >> 10) * (1.0 / 9007199254740991.0); r = p; }
  uint64_t o;
  std::memcpy(&o, &r, 8);
  return o;
}
int chol_dsfmt_jump_launch(const void *K, const void *polys, int poly_u32, int D, void *states, int nseg, void *st) {
  __sync_fetch_and_add(&launches, 1);
  if (nseg <= 0 || D <= 0 || poly_u32 * 32 < D) return 1;
  return enqueue(st, [=] {
    const uint32_t *k = (const uint32_t *)K, *p = (const uint32_t *)polys;
    uint32_t *out = (uint32_t *)states;
    for (int g = 0; g < nseg; ++g) {
      uint32_t acc[768] = {0};
      for (int i = 0; i < D; ++i)
        if ((p[(size_t)g * poly_u32 + i / 32] >> (i % 32)) & 1u)
          for (int w = 0; w < 768; ++w) acc[w] ^= k[(size_t)i * 768 + w];
      std::memcpy(out + (size_t)g * 768, acc, sizeof acc);
    }
  });
}
int chol_dsfmt_generate_launch(const void *states, unsigned long long refill0, unsigned seg_refills, int nseg, unsigned long long w_begin,
                               unsigned long long w_end, void *out, unsigned long long n_out, int kind, void *snap, void *st) {
  __sync_fetch_and_add(&launches, 1);
  if (nseg <= 0 || seg_refills == 0 || kind < 0 || kind > 4 || (snap != nullptr && nseg != 1)) return 1;
  return enqueue(st, [=] {
    for (int g = 0; g < nseg; ++g) {
      uint64_t cur[192][2];
      std::memcpy(cur, (const char *)states + (size_t)g * sizeof cur, sizeof cur);
      for (unsigned k = 0; k < seg_refills; ++k) {
        const unsigned long long r = refill0 + (unsigned long long)g * seg_refills + k;
        if (snap == nullptr && r * 191 >= w_end) break;
        mock_refill(cur);
        if (snap != nullptr) {
          std::memcpy((char *)snap + (size_t)(r + 1) * sizeof cur, cur, sizeof cur);
          continue;
        }
        for (int i = 0; i < 191; ++i) {
          const unsigned long long t = r * 191 + i;
          if (t < w_begin || t >= w_end) continue;
          const unsigned long long e = 2 * (t - w_begin);
          uint64_t *o = (uint64_t *)out;
          if (e < n_out) o[e] = mock_convert(cur[i][0], kind);
          if (e + 1 < n_out) o[e + 1] = mock_convert(cur[i][1], kind);
        }
      }
    }
  });
}
}
