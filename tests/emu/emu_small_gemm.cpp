// emu_small_gemm.cpp — csrc/gemm_small.cu (the small-tile FP64 GEMM of the panel chain) executed on
// the CPU (tests/emu/cuda_emu.h) against plain loops: every op(A)/op(B), full and triangle-only
// output, ragged sizes, C != D, alpha = 0 and beta = 0.  Prints "EMU PASSED" / exit code 1.
#define CUDA_EMU_IMPLEMENTATION
#include "cuda_emu.h"

#include "../../cuda-cholesky_b200/csrc/gemm_small.cu"

#include <random>
#include <string>

using namespace chol;

static int failures = 0;

static void run_case(int M, int N, int K, int opA, int opB, int cuplo, double alpha, double beta, bool inplace) {
  std::mt19937 g(7u + M + 3 * N + 5 * K);
  std::uniform_real_distribution<double> u(-1, 1);
  const int ar = opA ? K : M, ac = opA ? M : K, br = opB ? N : K, bc = opB ? K : N;
  const int lda = ar + 1, ldb = br + 2, ldc = M + 3, ldd = inplace ? ldc : M + 1;
  std::vector<double> A((size_t)lda * ac), B((size_t)ldb * bc), C((size_t)ldc * N), D((size_t)ldd * N, 99.0);
  for (auto &x : A) x = u(g);
  for (auto &x : B) x = u(g);
  for (auto &x : C) x = u(g);
  std::vector<double> C0 = C;
  SmallArgs a;
  a.A = A.data(); a.B = B.data(); a.C = C.data(); a.D = inplace ? C.data() : D.data();
  a.lda = lda; a.ldb = ldb; a.ldc = ldc; a.ldd = ldd;
  a.M = M; a.N = N; a.K = (alpha == 0.0) ? 0 : K;
  a.alpha = alpha; a.beta = beta; a.beta_zero = (beta == 0.0);
  a.opA = opA; a.opB = opB; a.cuplo = cuplo;
  dim3 grid((M + S_BM - 1) / S_BM, (N + S_BN - 1) / S_BN);
  emu::launch(grid, dim3(S_THREADS), 0, [&] { dgemm_small_kernel(a); });
  const double *out = inplace ? C.data() : D.data();
  double err = 0;
  for (int n = 0; n < N; ++n)
    for (int m = 0; m < M; ++m) {
      const bool touched = cuplo == 0 || (cuplo == 1 && m >= n) || (cuplo == 2 && m <= n);
      double want;
      if (touched) {
        double s = 0;
        for (int k = 0; k < K; ++k) {
          const double av = opA ? A[k + (size_t)m * lda] : A[m + (size_t)k * lda];
          const double bv = opB ? B[n + (size_t)k * ldb] : B[k + (size_t)n * ldb];
          s += av * bv;
        }
        want = alpha * s + (beta == 0.0 ? 0.0 : beta * C0[m + (size_t)n * ldc]);
      } else {
        want = inplace ? C0[m + (size_t)n * ldc] : 99.0;  // never written
      }
      err = std::max(err, std::fabs(out[m + (size_t)n * ldd] - want));
    }
  // the padding rows of the output are never written
  for (int n = 0; n < N; ++n)
    for (int m = M; m < ldd; ++m) {
      const double pad = inplace ? C0[m + (size_t)n * ldc] : 99.0;
      if (out[m + (size_t)n * ldd] != pad) err = 1e30;
    }
  const bool ok = err <= 1e-12 * (K + 1);
  if (!ok) {
    ++failures;
    std::printf("MISMATCH M=%d N=%d K=%d opA=%d opB=%d cuplo=%d alpha=%g beta=%g inplace=%d err=%.2e\n", M, N, K, opA, opB, cuplo, alpha, beta,
                (int)inplace, err);
  }
}

int main(int argc, char **argv) {
  std::setvbuf(stdout, nullptr, _IOLBF, 0);
  int cases = 0;
  if (argc > 3 && std::string(argv[1]) == "fuzz") {  // ./emu_small_gemm fuzz SEED COUNT: random shapes / ops / triangles / scalars
    std::mt19937 g((unsigned)atoi(argv[2]));
    auto ri = [&](int lo, int hi) { return lo + (int)(g() % (unsigned)(hi - lo + 1)); };
    const double alphas[4] = {-1.0, 0.7, 1.0, 0.0}, betas[4] = {1.0, -0.3, 0.0, 2.0};
    for (int i = 0, n = atoi(argv[3]); i < n; ++i, ++cases)
      run_case(ri(1, 130), ri(1, 130), ri(0, 9) == 0 ? 0 : ri(1, 150), ri(0, 1), ri(0, 1), ri(0, 2), alphas[ri(0, 3)], betas[ri(0, 3)], ri(0, 1) != 0);
    std::printf("fuzz: %d cases, %d mismatches\n", cases, failures);
    if (failures) { std::printf("EMU FAILED: %d mismatches\n", failures); return 1; }
    std::printf("EMU PASSED\n");
    return 0;
  }
  const int shapes[][3] = {{32, 32, 16}, {37, 50, 23}, {64, 33, 40}, {1, 1, 1}, {96, 96, 128}, {5, 70, 3}};
  for (auto &sh : shapes)
    for (int opA = 0; opA < 2; ++opA)
      for (int opB = 0; opB < 2; ++opB)
        for (int cuplo = 0; cuplo < 3; ++cuplo) {
          run_case(sh[0], sh[1], sh[2], opA, opB, cuplo, -1.0, 1.0, true);     // the panel chain's own form: C -= A B, in place
          run_case(sh[0], sh[1], sh[2], opA, opB, cuplo, 0.7, -0.3, false);    // C != D
          cases += 2;
        }
  run_case(40, 40, 20, 0, 1, 0, 1.0, 0.0, false);   // beta = 0: C is never read
  run_case(40, 40, 20, 0, 1, 1, 0.0, 2.0, true);    // alpha = 0: D = beta C
  run_case(33, 31, 0, 0, 0, 0, 1.0, 0.5, true);     // k = 0
  cases += 3;
  std::printf("%d cases\n", cases);
  if (failures) { std::printf("EMU FAILED: %d mismatches\n", failures); return 1; }
  std::printf("EMU PASSED\n");
  return 0;
}
