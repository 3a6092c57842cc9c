// emu_panel.cpp — the diagonal-block kernels of csrc/panel_kernels.cu executed on the CPU
// (tests/emu/cuda_emu.h): potf2 (in-kernel look-ahead), trsm_rows, trti2, lauu2 against plain
// loops, for lower / upper storage, full and ragged block sizes, and a non-positive-definite block.
// Prints one line per case and "EMU PASSED" at the end; exit code 1 on the first mismatch.
#define CUDA_EMU_IMPLEMENTATION
#include "cuda_emu.h"

#include "../../cuda-cholesky_b200/csrc/panel_kernels.cu"

#include <complex>
#include <random>

using namespace chol;

template <typename T> struct Host;  // device element type -> std type
template <> struct Host<double> { using type = double; };
template <> struct Host<float> { using type = float; };
template <> struct Host<double2> { using type = std::complex<double>; };
template <> struct Host<float2> { using type = std::complex<float>; };
template <typename H> static H cj(H x) { return x; }
template <typename R> static std::complex<R> cj(std::complex<R> x) { return std::conj(x); }
template <typename H> static double ab(H x) { return std::abs(x); }

template <typename T, typename H>
static void spd(int n, int lda, std::vector<H> &A, unsigned seed) {  // Hermitian positive definite, both triangles filled
  std::mt19937 g(seed);
  std::uniform_real_distribution<double> u(-0.5, 0.5);
  std::vector<H> B((size_t)n * n);
  for (auto &x : B) {
    if constexpr (ScalarOf<T>::cplx) x = H((typename H::value_type)u(g), (typename H::value_type)u(g));
    else x = (H)u(g);
  }
  A.assign((size_t)lda * n, H(77));  // padding rows keep a sentinel
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      H s = (i == j) ? H(n) : H(0);
      for (int k = 0; k < n; ++k) s += B[i + (size_t)k * n] * cj(B[j + (size_t)k * n]);
      A[i + (size_t)j * lda] = s;
    }
}

template <typename H>
static int ref_potrf_lower(int n, std::vector<H> &L, int ld) {  // in place, lower of a full Hermitian matrix; returns info
  for (int c = 0; c < n; ++c) {
    double d = std::real(L[c + (size_t)c * ld]);
    if (!(d > 0)) return c + 1;
    const double r = std::sqrt(d);
    L[c + (size_t)c * ld] = H(r);
    for (int i = c + 1; i < n; ++i) L[i + (size_t)c * ld] = L[i + (size_t)c * ld] / H(r);
    for (int k = c + 1; k < n; ++k)
      for (int i = k; i < n; ++i) L[i + (size_t)k * ld] -= L[i + (size_t)c * ld] * cj(L[k + (size_t)c * ld]);
  }
  return 0;
}

static int failures = 0;
static void check(bool ok, const char *what, double err) {
  std::printf("%-72s %s (%.2e)\n", what, ok ? "ok" : "MISMATCH", err);
  if (!ok) ++failures;
}

template <typename T, int NMAX>
static void run_potf2(int upper, int n, int lda, std::vector<typename Host<T>::type> &A, long *info, long base, double *logacc = nullptr) {
  const size_t sm = (size_t)NMAX * (NMAX + 1) * sizeof(T);
  T *dA = reinterpret_cast<T *>(A.data());
  emu::launch(dim3(1), dim3(PANEL_THREADS), sm, [&] { potf2_kernel<T, NMAX / 32, NMAX / 16>(upper, n, dA, lda, info, base, logacc); });
}

// potf2 (register-resident, in-kernel look-ahead): factor vs plain loops, other triangle and padding rows untouched
template <typename T, int NMAX>
static void case_potf2(const char *name, int n, double tol) {
  using H = typename Host<T>::type;
  const int lda = n + 3;
  for (int upper = 0; upper < 2; ++upper) {
    std::vector<H> A0, A1;
    spd<T, H>(n, lda, A0, 11u + n);
    A1 = A0;
    long i1 = 0;
    double logacc = 1.5;  // the fused logdet ADDS to what is there
    if (getenv("EMU_TRACE")) std::fprintf(stderr, "[%s n=%d upper=%d] potf2...\n", name, n, upper);
    run_potf2<T, NMAX>(upper, n, lda, A1, &i1, 100, &logacc);
    std::vector<H> L = A0;
    const int ri = ref_potrf_lower(n, L, lda);
    double want_ld = 1.5;
    for (int c = 0; c < n; ++c) want_ld += 2.0 * std::log((double)std::real(L[c + (size_t)c * lda]));
    if (std::fabs(logacc - want_ld) > 1e-4 * std::fabs(want_ld) + 1e-4) { check(false, "fused logdet", std::fabs(logacc - want_ld)); }
    double e_ref = 0, e_other = 0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        const bool in_tri = upper ? (i <= j) : (i >= j);
        const H a1 = A1[i + (size_t)j * lda];
        if (in_tri) e_ref = std::max(e_ref, ab(a1 - (upper ? cj(L[j + (size_t)i * lda]) : L[i + (size_t)j * lda])));
        else e_other = std::max(e_other, ab(a1 - A0[i + (size_t)j * lda]));
      }
    char what[160];
    std::snprintf(what, sizeof what, "%s potf2 %s n=%d: info, factor vs loops, other triangle untouched", name, upper ? "upper" : "lower", n);
    check(ri == 0 && i1 == 0 && e_ref <= tol && e_other == 0.0, what, e_ref);
    for (int j = 0; j < n; ++j)  // padding rows never touched
      for (int i = n; i < lda; ++i)
        if (A1[i + (size_t)j * lda] != H(77)) { check(false, "padding rows of A overwritten", 0); return; }
  }
}

template <typename T, int NMAX>
static void case_not_pd(const char *name, int n, int bad, double tol = 1e-10) {
  using H = typename Host<T>::type;
  const int lda = n;
  for (int upper = 0; upper < 2; ++upper) {
    std::vector<H> A0, A1;
    spd<T, H>(n, lda, A0, 5);
    A0[bad + (size_t)bad * lda] = H(-3);
    A1 = A0;
    long i1 = 0;
    if (getenv("EMU_TRACE")) std::fprintf(stderr, "[not PD %s n=%d upper=%d]...\n", name, n, upper);
    run_potf2<T, NMAX>(upper, n, lda, A1, &i1, 1000);
    // the columns finished before the failure are the reference's
    std::vector<H> L = A0;
    ref_potrf_lower(n, L, lda);  // stops at the bad column; columns < bad are final
    double e = 0;
    for (int j = 0; j < bad; ++j)
      for (int i = j; i < n; ++i) e = std::max(e, ab(A1[upper ? j + (size_t)i * lda : i + (size_t)j * lda] - (upper ? cj(L[i + (size_t)j * lda]) : L[i + (size_t)j * lda])));
    char what[160];
    std::snprintf(what, sizeof what, "%s potf2 %s n=%d not PD at column %d: info, finished columns", name, upper ? "upper" : "lower", n, bad);
    check(i1 == 1000 + bad + 1 && e <= tol, what, e);
    // a kernel that finds *dinfo != 0 on entry must not touch anything
    std::vector<H> A3 = A0;
    long i3 = 42;
    run_potf2<T, NMAX>(upper, n, lda, A3, &i3, 0);
    bool same = (i3 == 42);
    for (size_t k = 0; k < A3.size(); ++k) same = same && (A3[k] == A0[k]);
    check(same, "  ... and a block entered with info != 0 is left alone", 0);
  }
}

// trsm_rows: X = B * L^-H against plain substitution; lower / upper storage of the triangle, B and X plain or as
// their conjugate transposes, in place and out of place, row counts that do not fill the last warp
template <typename T, int NMAX, int G, int RR>
static void case_trsm_rows(const char *name, int n, int m, int nwarps, double tol) {
  using H = typename Host<T>::type;
  using Real = typename ScalarOf<T>::Real;
  const int ldl = n + 2;
  std::mt19937 g(99u + n + m);
  std::uniform_real_distribution<double> u(-0.5, 0.5);
  auto rnd = [&]() -> H { if constexpr (ScalarOf<T>::cplx) return H((typename H::value_type)u(g), (typename H::value_type)u(g)); else return (H)u(g); };
  std::vector<H> Lf((size_t)n * n, H(0));  // lower factor, diagonal real positive
  for (int j = 0; j < n; ++j)
    for (int i = j; i < n; ++i) Lf[i + (size_t)j * n] = (i == j) ? H(2.0 + 0.01 * i) : rnd();
  for (int upper = 0; upper < 2; ++upper)
    for (int tb = 0; tb < 2; ++tb)
      for (int tx = 0; tx < 2; ++tx)
        for (int inplace = 0; inplace < 2; ++inplace) {
          if (inplace && tb != tx) continue;
          std::vector<H> Ls((size_t)ldl * n, H(33));  // the other triangle holds a sentinel that must never be read
          for (int j = 0; j < n; ++j)
            for (int i = j; i < n; ++i) {
              if (!upper) Ls[i + (size_t)j * ldl] = Lf[i + (size_t)j * n]; else Ls[j + (size_t)i * ldl] = cj(Lf[i + (size_t)j * n]);
            }
          const int ldb = (tb ? n : m) + 1, ldx = (tx ? n : m) + 3;
          std::vector<H> Bm((size_t)m * n), Bs((size_t)ldb * (tb ? m : n), H(5)), Xs((size_t)ldx * (tx ? m : n), H(7));
          for (auto &v : Bm) v = rnd();
          for (int r = 0; r < m; ++r)
            for (int j = 0; j < n; ++j) {
              if (tb) Bs[j + (size_t)r * ldb] = cj(Bm[r + (size_t)j * m]); else Bs[r + (size_t)j * ldb] = Bm[r + (size_t)j * m];
            }
          std::vector<H> want((size_t)m * n);
          for (int r = 0; r < m; ++r)
            for (int k = 0; k < n; ++k) {
              H sacc = Bm[r + (size_t)k * m];
              for (int j = 0; j < k; ++j) sacc -= want[r + (size_t)j * m] * cj(Lf[k + (size_t)j * n]);
              want[r + (size_t)k * m] = sacc / Lf[k + (size_t)k * n];
            }
          const size_t sm = (size_t)NMAX * (NMAX + 1) * sizeof(T) + NMAX * sizeof(Real);
          const int cw = nwarps > 1 ? nwarps - 1 : 1;  // computing warps; a CTA's last warp only helps staging
          const int rows_per_cta = cw * (32 / G) * RR, grid = (m + rows_per_cta - 1) / rows_per_cta;
          const T *dL = reinterpret_cast<const T *>(Ls.data());
          std::vector<H> &Out = inplace ? Bs : Xs;
          const int ldo = inplace ? ldb : ldx;
          const T *dB = reinterpret_cast<const T *>(Bs.data());
          T *dX = reinterpret_cast<T *>(Out.data());
          long zero = 0;
          emu::launch(dim3(grid), dim3(256), sm, [&] {  // the kernel is written for 256 threads: all of them stage, cw warps own rows
            trsm_rows_kernel<T, NMAX, G, RR>(upper, n, dL, ldl, m, cw, tb, dB, ldb, tx, dX, ldo, &zero);
          });
          double e = 0;
          for (int r = 0; r < m; ++r)
            for (int j = 0; j < n; ++j) {
              const H got = tx ? cj(Out[j + (size_t)r * ldo]) : Out[r + (size_t)j * ldo];
              e = std::max(e, ab(got - want[r + (size_t)j * m]));
            }
          bool pad_ok = true;  // padding rows of the output keep their sentinel
          const H sent = inplace ? H(5) : H(7);
          const int orow = tx ? n : m, ocol = tx ? m : n;
          for (int c2 = 0; c2 < ocol; ++c2)
            for (int r2 = orow; r2 < ldo; ++r2) pad_ok = pad_ok && Out[r2 + (size_t)c2 * ldo] == sent;
          char what[200];
          std::snprintf(what, sizeof what, "%s trsm_rows<G=%d,RR=%d> %s n=%d m=%d transB=%d transX=%d %s", name, G, RR, upper ? "upper" : "lower", n,
                        m, tb, tx, inplace ? "in place" : "out of place");
          check(e <= tol && pad_ok, what, e);
        }
  // a launch that finds *dinfo != 0 leaves X alone
  std::vector<H> Ls((size_t)n * n, H(1)), Bs((size_t)m * n, H(2)), Xs((size_t)m * n, H(9));
  long bad = 3;
  const size_t sm = (size_t)NMAX * (NMAX + 1) * sizeof(T) + NMAX * sizeof(Real);
  emu::launch(dim3(1), dim3(256), sm, [&] {
    trsm_rows_kernel<T, NMAX, G, RR>(0, n, reinterpret_cast<const T *>(Ls.data()), n, m, nwarps, 0, reinterpret_cast<const T *>(Bs.data()), m, 0,
                                     reinterpret_cast<T *>(Xs.data()), m, &bad);
  });
  bool same = true;
  for (auto &v : Xs) same = same && v == H(9);
  check(same, "  ... trsm_rows entered with info != 0 leaves X alone", 0);
}

// ---- the unfused kernels: potf2 (the last diagonal block of a factorisation), trti2, lauu2 ----
template <typename T, int NMAX>
static void case_unfused(const char *name, int n, double tol) {
  using H = typename Host<T>::type;
  const int lda = n + 1;
  for (int upper = 0; upper < 2; ++upper) {
    std::vector<H> A0, A;
    spd<T, H>(n, lda, A0, 23u + n);
    A = A0;
    long info = 0;
    T *dA = reinterpret_cast<T *>(A.data());
    emu::launch(dim3(1), dim3(PANEL_THREADS), (size_t)NMAX * (NMAX + 1) * sizeof(T), [&] { potf2_kernel<T, NMAX / 32, NMAX / 16>(upper, n, dA, lda, &info, 7, nullptr); });
    std::vector<H> L = A0;
    ref_potrf_lower(n, L, lda);
    double e = 0, e_other = 0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        const bool in_tri = upper ? (i <= j) : (i >= j);
        if (in_tri) e = std::max(e, ab(A[i + (size_t)j * lda] - (upper ? cj(L[j + (size_t)i * lda]) : L[i + (size_t)j * lda])));
        else e_other = std::max(e_other, ab(A[i + (size_t)j * lda] - A0[i + (size_t)j * lda]));
      }
    char what[160];
    std::snprintf(what, sizeof what, "%s potf2 %s n=%d vs loops, other triangle untouched", name, upper ? "upper" : "lower", n);
    check(info == 0 && e <= tol && e_other == 0.0, what, e);

    // trti2 of that factor, out of place with the other triangle zeroed: X F = I
    std::vector<H> X((size_t)lda * n, H(55));
    T *dX = reinterpret_cast<T *>(X.data());
    long tinfo = 0;
    const size_t sm = ((size_t)(n | 1) * n + n) * sizeof(T);
    emu::launch(dim3(1), dim3(PANEL_THREADS), sm, [&] {
      trti2_kernel<T, NMAX / 32, NMAX / 16>(upper, 0, n, dA, lda, dX, lda, 1, &tinfo, 0);
    });
    {  // the pruned-chain variant must give the same bits
      std::vector<H> X2((size_t)lda * n, H(55));
      T *dX2 = reinterpret_cast<T *>(X2.data());
      long t2 = 0;
      emu::launch(dim3(1), dim3(PANEL_THREADS), sm, [&] {
        trti2_kernel<T, NMAX / 32, NMAX / 16, true>(upper, 0, n, dA, lda, dX2, lda, 1, &t2, 0);
      });
      char w2[120];
      std::snprintf(w2, sizeof w2, "%s trti2 %s n=%d: V2 == measured variant", name, upper ? "upper" : "lower", n);
      check(t2 == tinfo && X2 == X, w2, 0);
    }
    double e_inv = 0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        H sum = 0;
        for (int k = 0; k < n; ++k) {
          const bool fin = upper ? (k <= j) : (k >= j);
          if (fin) sum += X[i + (size_t)k * lda] * A[k + (size_t)j * lda];  // X's other triangle is zero
        }
        e_inv = std::max(e_inv, ab(sum - H(i == j ? 1 : 0)));
      }
    std::snprintf(what, sizeof what, "%s trti2 %s n=%d: X F = I", name, upper ? "upper" : "lower", n);
    check(tinfo == 0 && e_inv <= 50 * tol, what, e_inv);

    // lauu2 of the factor: lower L^H L, upper U U^H, triangle only
    std::vector<H> P = A;
    T *dP = reinterpret_cast<T *>(P.data());
    emu::launch(dim3(1), dim3(PANEL_THREADS), (size_t)(n | 1) * n * sizeof(T), [&] { lauu2_kernel<T>(upper, n, dP, lda); });
    double e_l = 0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        const bool in_tri = upper ? (i <= j) : (i >= j);
        if (!in_tri) { e_l = std::max(e_l, ab(P[i + (size_t)j * lda] - A[i + (size_t)j * lda])); continue; }
        H sum = 0;
        for (int k = 0; k < n; ++k) {
          if (!upper) { if (k >= i && k >= j) sum += cj(A[k + (size_t)i * lda]) * A[k + (size_t)j * lda]; }  // (L^H L)(i,j)
          else if (k >= i && k >= j) sum += A[i + (size_t)k * lda] * cj(A[j + (size_t)k * lda]);              // (U U^H)(i,j)
        }
        e_l = std::max(e_l, ab(P[i + (size_t)j * lda] - sum));
      }
    std::snprintf(what, sizeof what, "%s lauu2 %s n=%d vs loops", name, upper ? "upper" : "lower", n);
    check(e_l <= 200 * tol * n, what, e_l);
  }
  // potf2: a non-positive pivot
  std::vector<H> A0;
  spd<T, H>(n, n, A0, 3);
  const int bad = n / 2;
  A0[bad + (size_t)bad * n] = H(-1);
  long info = 0;
  T *dA = reinterpret_cast<T *>(A0.data());
  emu::launch(dim3(1), dim3(PANEL_THREADS), (size_t)NMAX * (NMAX + 1) * sizeof(T), [&] { potf2_kernel<T, NMAX / 32, NMAX / 16>(0, n, dA, n, &info, 500, nullptr); });
  char what[160];
  std::snprintf(what, sizeof what, "%s potf2 n=%d not PD at column %d: info", name, n, bad);
  check(info == 500 + bad + 1, what, 0);
}

// ---- data movement and the logdet reduction ----
template <typename T>
static void case_copy_fill_logdet(const char *name) {
  using H = typename Host<T>::type;
  const int m = 70, n = 45, lda = 73, ldb = 81;
  std::mt19937 g(17);
  std::uniform_real_distribution<double> u(0.5, 1.5);
  std::vector<H> A((size_t)lda * n);
  for (auto &x : A) { if constexpr (ScalarOf<T>::cplx) x = H((typename H::value_type)u(g), (typename H::value_type)u(g)); else x = (H)u(g); }
  const T *dA = reinterpret_cast<const T *>(A.data());
  for (int trans = 0; trans < 3; ++trans)     // 0 N, 1 T, 2 C
    for (int uplo = 0; uplo < 5; ++uplo) {    // 0 all, 1 lower, 2 upper, 3 strictly lower, 4 strictly upper (of the SOURCE)
      std::vector<H> B((size_t)ldb * std::max(m, n), H(9)), W = B;
      T *dB = reinterpret_cast<T *>(B.data());
      emu::launch(dim3((m + 31) / 32, (n + 31) / 32), dim3(256), 0, [&] { copy_kernel<T>(trans, uplo, m, n, dA, lda, dB, ldb); });
      for (int c = 0; c < n; ++c)
        for (int r = 0; r < m; ++r) {
          const bool keep = uplo == 0 || (uplo == 1 && r >= c) || (uplo == 2 && r <= c) || (uplo == 3 && r > c) || (uplo == 4 && r < c);
          if (!keep) continue;
          const H v = A[r + (size_t)c * lda];
          if (trans == 0) W[r + (size_t)c * ldb] = v; else W[c + (size_t)r * ldb] = trans == 2 ? cj(v) : v;
        }
      char what[120];
      std::snprintf(what, sizeof what, "%s copy trans=%d uplo=%d", name, trans, uplo);
      check(B == W, what, 0);
    }
  for (int uplo = 0; uplo < 3; ++uplo) {  // 0 all, 1 strictly lower, 2 strictly upper
    std::vector<H> B((size_t)lda * n, H(3)), W = B;
    T *dB = reinterpret_cast<T *>(B.data());
    const H val = ScalarOf<T>::cplx ? H(cj(H(2)) ) : H(2);
    T tv;
    std::memcpy(&tv, &val, sizeof tv);
    emu::launch(dim3((m + 31) / 32, (n + 31) / 32), dim3(256), 0, [&] { fill_kernel<T>(uplo, m, n, tv, dB, lda); });
    for (int c = 0; c < n; ++c)
      for (int r = 0; r < m; ++r)
        if (uplo == 0 || (uplo == 1 && r > c) || (uplo == 2 && r < c)) W[r + (size_t)c * lda] = val;
    char what[120];
    std::snprintf(what, sizeof what, "%s fill uplo=%d", name, uplo);
    check(B == W, what, 0);
  }
  // logdet: 2 * sum log(Re x[i*incx]), several blocks, the last one to finish adds the partials
  using Real = typename ScalarOf<T>::Real;
  const long long cnt = 1500, incx = 3;
  std::vector<H> x((size_t)cnt * incx);
  for (auto &v : x) { if constexpr (ScalarOf<T>::cplx) v = H((typename H::value_type)u(g), (typename H::value_type)u(g)); else v = (H)u(g); }
  double want = 0;
  for (long long i = 0; i < cnt; ++i) want += std::log((double)std::real(x[(size_t)(i * incx)]));
  want *= 2;
  LogdetScratch scr;
  std::memset(&scr, 0, sizeof scr);
  Real res = -1;
  const int blocks = (int)((cnt + LOGDET_THREADS - 1) / LOGDET_THREADS);
  for (int rep = 0; rep < 2; ++rep) {  // twice: the ticket must have been reset by the first launch
    res = -1;
    emu::launch(dim3(blocks), dim3(LOGDET_THREADS), 0, [&] { logdet_kernel<T>(reinterpret_cast<const T *>(x.data()), incx, cnt, &res, &scr); });
    char what[120];
    std::snprintf(what, sizeof what, "%s logdet, %d blocks, launch %d", name, blocks, rep);
    check(std::fabs((double)res - want) <= (sizeof(Real) == 4 ? 2e-4 : 1e-12) * std::fabs(want), what, std::fabs((double)res - want));
  }
}

// scatter_rows: a rank's packed piece of a row block to its block-cyclic rows (distributed LAUUM)
template <typename T>
static void case_scatter_rows(const char *name) {
  using H = typename Host<T>::type;
  const int m = 300, n = 37, lda = 301, cnb = 128, cp = 3, coff = 2;
  const int rows_out = ((m - 1) / cnb * cp + coff) * cnb + (m - 1) % cnb + 1, ldb = rows_out + 2;
  std::vector<H> A((size_t)lda * n), B((size_t)ldb * n, H(9)), W = B;
  for (size_t k = 0; k < A.size(); ++k) A[k] = H((double)(k % 1013) + 0.25);
  emu::launch(dim3((m + 31) / 32, (n + 31) / 32), dim3(256), 0, [&] {
    scatter_rows_kernel<T>(m, n, reinterpret_cast<const T *>(A.data()), lda, reinterpret_cast<T *>(B.data()), ldb, cnb, cp, coff);
  });
  for (int c = 0; c < n; ++c)
    for (int r = 0; r < m; ++r) W[(size_t)c * ldb + ((r / cnb) * cp + coff) * cnb + r % cnb] = A[(size_t)c * lda + r];
  char what[120];
  std::snprintf(what, sizeof what, "%s scatter_rows to block-cyclic rows (everything else untouched)", name);
  check(B == W, what, 0);
}

// ./emu_panel sizes PREC [STEP]: every block order the one-CTA kernels accept (1 .. 128, z: 64), not just the handful of the
// default run - potf2 (+ a not-PD column drawn per size), trti2 / lauu2, and trsm_rows in both shapes with a row count drawn per size
template <typename T, int NMAX, int GW>
static void every_size(const char *name, double tol, int step) {
  std::mt19937 g(4242);
  for (int n = 1; n <= NMAX; n += step) {
    case_potf2<T, NMAX>(name, n, tol);
    case_not_pd<T, NMAX>(name, n, (int)(g() % (unsigned)n), std::max(tol, 1e-10));
    case_unfused<T, NMAX>(name, n, tol);
    case_trsm_rows<T, NMAX, 32, 1>(name, n, 1 + (int)(g() % 9u), 1 + (int)(g() % 4u), tol);
    case_trsm_rows<T, NMAX, GW, 4>(name, n, 1 + (int)(g() % 70u), 1 + (int)(g() % 3u), tol);
  }
}

int main(int argc, char **argv) {
  std::setvbuf(stdout, nullptr, _IOLBF, 0);
  if (argc > 2 && std::string(argv[1]) == "sizes") {
    const int step = argc > 3 ? std::max(1, atoi(argv[3])) : 1;
    switch (argv[2][0]) {
      case 'd': every_size<double, 128, 8>("d", 1e-12, step); break;
      case 'z': every_size<double2, 64, 8>("z", 1e-12, step); break;
      case 's': every_size<float, 128, 8>("s", 2e-4, step); break;
      default: every_size<float2, 128, 16>("c", 2e-4, step); break;
    }
    if (failures) { std::printf("EMU FAILED: %d mismatches\n", failures); return 1; }
    std::printf("EMU PASSED\n");
    return 0;
  }
  case_potf2<double, 128>("d", 128, 1e-12);
  case_potf2<double, 128>("d", 100, 1e-12);
  case_potf2<double, 128>("d", 37, 1e-12);
  case_potf2<double, 128>("d", 17, 1e-12);
  case_potf2<double, 128>("d", 16, 1e-12);
  case_potf2<double, 128>("d", 1, 1e-12);
  case_potf2<double2, 64>("z", 64, 1e-12);
  case_potf2<double2, 64>("z", 45, 1e-12);
  case_potf2<float, 128>("s", 97, 2e-4);
  case_potf2<float2, 128>("c", 80, 2e-4);
  case_not_pd<double, 128>("d", 128, 70);
  case_not_pd<double, 128>("d", 128, 127);
  case_not_pd<double, 128>("d", 128, 16);
  case_not_pd<double, 128>("d", 50, 0);
  case_not_pd<double, 128>("d", 50, 1);
  case_not_pd<double2, 64>("z", 64, 33);
  case_trsm_rows<double, 128, 32, 1>("d", 128, 5, 2, 1e-12);
  case_trsm_rows<double, 128, 32, 1>("d", 77, 3, 4, 1e-12);
  case_trsm_rows<double, 128, 8, 4>("d", 128, 37, 2, 1e-12);
  case_trsm_rows<double, 128, 8, 4>("d", 50, 16, 1, 1e-12);
  case_trsm_rows<double2, 64, 32, 1>("z", 64, 3, 2, 1e-12);
  case_trsm_rows<double2, 64, 8, 4>("z", 41, 21, 2, 1e-12);
  case_trsm_rows<float, 128, 8, 4>("s", 100, 20, 1, 2e-4);
  case_trsm_rows<float2, 128, 16, 4>("c", 128, 11, 2, 2e-4);
  case_trsm_rows<float2, 128, 32, 1>("c", 90, 2, 2, 2e-4);
  case_unfused<double, 128>("d", 128, 1e-12);
  case_unfused<double, 128>("d", 70, 1e-12);
  case_unfused<double2, 64>("z", 50, 1e-12);
  case_unfused<float, 128>("s", 90, 2e-4);
  case_scatter_rows<double>("d");
  case_scatter_rows<double2>("z");
  case_scatter_rows<float>("s");
  case_copy_fill_logdet<double>("d");
  case_copy_fill_logdet<double2>("z");
  case_copy_fill_logdet<float>("s");
  if (failures) { std::printf("EMU FAILED: %d mismatches\n", failures); return 1; }
  std::printf("EMU PASSED\n");
  return 0;
}
