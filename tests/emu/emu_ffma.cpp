// emu_ffma.cpp — the FP32 GEMM engine (csrc/gemm_ffma.cu: k-tile of 16, hoisted loader for the hot shape)
// executed on the CPU (tests/emu/cuda_emu.h; cp.async
// groups complete as late as wait_group allows) against plain loops: s and c; op(A)/op(B) incl.
// conjugation; triangular operand masks with unit diagonals and the k-range skip; triangle-only and
// Hermitian output; ragged edges; aligned (16-byte fast path) and unaligned operands.
#define CUDA_EMU_IMPLEMENTATION
#include "cuda_emu.h"

#include "../../cuda-cholesky_b200/csrc/gemm_ffma.cu"

#include <complex>
#include <random>
#include <string>

using namespace chol;

template <typename T> struct Host;
template <> struct Host<float> { using type = float; };
template <> struct Host<float2> { using type = std::complex<float>; };
static float cj(float x) { return x; }
static std::complex<float> cj(std::complex<float> x) { return std::conj(x); }

static int failures = 0, cases = 0;

struct Case {
  int M, N, K, opA, opB, triA, triB, unitA, unitB, cuplo, herm;  // op: 0 N, 1 T, 2 C; tri: 0, 1 lower, 2 upper
  bool aligned;
};

template <typename T>
static void run_case(const Case &c, bool deep) {
  using H = typename Host<T>::type;
  constexpr bool CPLX = ScalarOf<T>::cplx;
  std::mt19937 g(1234u + 7 * c.M + 11 * c.N + 13 * c.K + c.opA + 3 * c.opB);
  std::uniform_real_distribution<float> u(-1, 1);
  auto rnd = [&] { if constexpr (CPLX) return H(u(g), u(g)); else return H(u(g)); };
  const int ar = c.opA ? c.K : c.M, ac = c.opA ? c.M : c.K, br = c.opB ? c.N : c.K, bc = c.opB ? c.K : c.N;
  const int pad = c.aligned ? 0 : 1;
  const int lda = ((ar + 3) / 4) * 4 + pad, ldb = ((br + 3) / 4) * 4 + pad, ldc = c.M + 5;
  // 16-byte aligned bases (+ one element when the case wants the slow path)
  std::vector<H> Abuf((size_t)lda * ac + 8), Bbuf((size_t)ldb * bc + 8), C((size_t)ldc * c.N), C0;
  H *A = Abuf.data() + pad, *B = Bbuf.data() + pad;
  for (auto &x : Abuf) x = rnd();
  for (auto &x : Bbuf) x = rnd();
  for (auto &x : C) x = rnd();
  C0 = C;
  const H alpha = CPLX ? H(0.7f) + H(rnd()) * 0.25f : H(0.7f), beta = CPLX ? H(-0.3f) + H(rnd()) * 0.25f : H(-0.3f);

  FArgs<T> a;
  std::memset(&a, 0, sizeof a);
  a.A = reinterpret_cast<const T *>(A); a.B = reinterpret_cast<const T *>(B);
  a.C = reinterpret_cast<const T *>(C.data()); a.D = reinterpret_cast<T *>(C.data());
  a.lda = lda; a.ldb = ldb; a.ldc = ldc; a.ldd = ldc;
  a.M = c.M; a.N = c.N; a.K = c.K;
  std::memcpy(&a.alpha, &alpha, sizeof(T)); std::memcpy(&a.beta, &beta, sizeof(T));
  a.beta_zero = 0;
  a.opA = c.opA; a.opB = c.opB;
  if (!CPLX) { if (a.opA == 2) a.opA = 1; if (a.opB == 2) a.opB = 1; }
  a.triA = c.triA; a.triB = c.triB; a.unitA = c.unitA; a.unitB = c.unitB;
  a.cuplo = c.cuplo; a.herm = CPLX ? c.herm : 0;
  using Cfg = FCfg<T>;
  a.NI = (a.M + Cfg::BM - 1) / Cfg::BM; a.NJ = (a.N + Cfg::BN - 1) / Cfg::BN;
  a.vecA = (((uintptr_t)A & 15) == 0) && (lda % Cfg::VW == 0);
  a.vecB = (((uintptr_t)B & 15) == 0) && (ldb % Cfg::VW == 0);
  a.conjA_regs = CPLX && c.opA == 2; a.conjB_regs = CPLX && c.opB == 2;
  a.cm.cnb = a.cm.cP = a.cm.coff = 0;
  const long long tiles = total_tiles<Cfg::BM, Cfg::BN>(8, a.cuplo, a.NI, a.NJ, a.cm);
  (void)deep;
  const size_t bytes = 2 * 4 * 16 * (size_t)Cfg::LDS * sizeof(T);
  emu::launch(dim3((unsigned)tiles), dim3(256), bytes, [&] { gemm_ffma_kernel<T>(a); });

  auto opa = [&](int m, int k) -> H {  // op(A)(m,k) with the stored-triangle mask and unit diagonal
    const int r = c.opA ? k : m, col = c.opA ? m : k;
    if ((c.triA == 1 && r < col) || (c.triA == 2 && r > col)) return H(0);
    H v = (c.unitA && c.triA && r == col) ? H(1) : A[r + (size_t)col * lda];
    return c.opA == 2 ? cj(v) : v;
  };
  auto opb = [&](int k, int n) -> H {
    const int r = c.opB ? n : k, col = c.opB ? k : n;
    if ((c.triB == 1 && r < col) || (c.triB == 2 && r > col)) return H(0);
    H v = (c.unitB && c.triB && r == col) ? H(1) : B[r + (size_t)col * ldb];
    return c.opB == 2 ? cj(v) : v;
  };
  double err = 0;
  for (int n = 0; n < c.N; ++n)
    for (int m = 0; m < c.M; ++m) {
      const bool touched = c.cuplo == 0 || (c.cuplo == 1 && m >= n) || (c.cuplo == 2 && m <= n);
      H want = C0[m + (size_t)n * ldc];
      if (touched) {
        std::complex<double> s = 0;
        for (int k = 0; k < c.K; ++k) s += std::complex<double>(opa(m, k)) * std::complex<double>(opb(k, n));
        std::complex<double> o = std::complex<double>(alpha) * s + std::complex<double>(beta) * std::complex<double>(C0[m + (size_t)n * ldc]);
        if (CPLX && c.herm && m == n) o = std::complex<double>(o.real(), 0.0);
        if constexpr (CPLX) want = H((float)o.real(), (float)o.imag()); else want = (float)o.real();
      }
      err = std::max(err, (double)std::abs(C[m + (size_t)n * ldc] - want));
    }
  for (int n = 0; n < c.N; ++n)
    for (int m = c.M; m < ldc; ++m)
      if (C[m + (size_t)n * ldc] != C0[m + (size_t)n * ldc]) err = 1e30;  // padding rows untouched
  ++cases;
  if (!(err <= 2e-5 * (c.K + 4))) {
    ++failures;
    std::printf("MISMATCH %s %s M=%d N=%d K=%d op=%d,%d tri=%d,%d unit=%d,%d cuplo=%d herm=%d aligned=%d err=%.2e\n", CPLX ? "c" : "s",
                deep ? "deep" : "measured", c.M, c.N, c.K, c.opA, c.opB, c.triA, c.triB, c.unitA, c.unitB, c.cuplo, c.herm, (int)c.aligned, err);
  }
}

template <typename T>
static void sweep(bool deep) {
  constexpr bool CPLX = ScalarOf<T>::cplx;
  const int hT = CPLX ? 2 : 1;
  const int big = CPLX ? 150 : 280;  // > 2 tiles in each direction, ragged
  // the POTRF shapes: NT update (interior fast path + edges), triangle-only, Hermitian diagonal
  for (bool al : {true, false}) {
    run_case<T>({big, big, 70, 0, hT, 0, 0, 0, 0, 1, 1, al}, deep);
    run_case<T>({big, big, 64, 0, hT, 0, 0, 0, 0, 2, 1, al}, deep);
    run_case<T>({big, 100, 48, 0, hT, 0, 0, 0, 0, 0, 0, al}, deep);
  }
  // every op pair, plain
  for (int oa : {0, 1, hT})
    for (int ob : {0, 1, hT}) run_case<T>({90, 77, 37, oa, ob, 0, 0, 0, 0, 0, 0, true}, deep);
  // triangular operands (TRMM forms): A triangular on the left (k = m), B triangular on the right (k = n)
  for (int tri : {1, 2})
    for (int unit : {0, 1})
      for (int op : {0, 1, hT}) {
        run_case<T>({big, 60, big, op, 0, tri, 0, unit, 0, 0, 0, true}, deep);
        run_case<T>({60, big, big, 0, op, 0, tri, 0, unit, 0, 0, false}, deep);
      }
  // LAUUM-like: both operands masked, triangle-only Hermitian output
  run_case<T>({100, 100, 130, hT, 0, 1, 1, 0, 0, 1, 1, true}, deep);
  run_case<T>({100, 100, 130, 0, hT, 2, 2, 0, 0, 2, 1, true}, deep);
  // degenerate sizes
  run_case<T>({1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, true}, deep);
  run_case<T>({5, 3, 0, 0, hT, 0, 0, 0, 0, 0, 0, true}, deep);
}

// ./emu_ffma fuzz SEED COUNT: random descriptors inside the engine's contract (shapes up to a few tiles, every op pair,
// aligned / unaligned operands, triangular operands on either side, triangle-only / Hermitian output)
template <typename T>
static void fuzz_one(std::mt19937 &g) {
  constexpr bool CPLX = ScalarOf<T>::cplx;
  auto ri = [&](int lo, int hi) { return lo + (int)(g() % (unsigned)(hi - lo + 1)); };
  const int hT = CPLX ? 2 : 1;
  const int ops[3] = {0, 1, hT};
  Case c{};
  c.M = ri(0, 4) == 0 ? ri(1, 30) : ri(1, CPLX ? 170 : 300);
  c.N = ri(0, 4) == 0 ? ri(1, 30) : ri(1, CPLX ? 170 : 300);
  c.K = ri(0, 9) == 0 ? 0 : ri(1, 90);
  c.opA = ops[ri(0, 2)];
  c.opB = ops[ri(0, 2)];
  c.aligned = ri(0, 1) != 0;
  const int kind = ri(0, 7);
  if (kind <= 2) {  // rank-k update family
    c.opA = 0;
    c.opB = hT;
    c.cuplo = ri(1, 2);
    c.herm = ri(0, 1);
    if (ri(0, 1)) c.N = c.M;
  } else if (kind == 3) {
    c.K = c.M;
    c.triA = ri(1, 2);
    c.unitA = ri(0, 1);
  } else if (kind == 4) {
    c.K = c.N;
    c.triB = ri(1, 2);
    c.unitB = ri(0, 1);
  } else if (kind == 5) {  // both masked, triangle-only output (LAUUM forms)
    c.M = c.N = ri(1, 140);
    c.K = c.M;
    c.triA = c.triB = ri(1, 2);
    c.cuplo = ri(1, 2);
    c.herm = ri(0, 1);
  }
  run_case<T>(c, true);
}

int main(int argc, char **argv) {
  std::setvbuf(stdout, nullptr, _IOLBF, 0);
  if (argc > 3 && std::string(argv[1]) == "fuzz") {
    std::mt19937 g((unsigned)atoi(argv[2]));
    for (int i = 0, n = atoi(argv[3]); i < n; ++i) {
      if (g() & 1u) fuzz_one<float>(g); else fuzz_one<float2>(g);
    }
    std::printf("fuzz: %d cases, %d mismatches\n", cases, failures);
    if (failures) { std::printf("EMU FAILED: %d mismatches\n", failures); return 1; }
    std::printf("EMU PASSED\n");
    return 0;
  }
  sweep<float>(true);
  sweep<float2>(true);
  std::printf("deep-k kernels: %d cases so far, %d mismatches\n", cases, failures);
  if (failures) { std::printf("EMU FAILED: %d mismatches\n", failures); return 1; }
  std::printf("EMU PASSED\n");
  return 0;
}
