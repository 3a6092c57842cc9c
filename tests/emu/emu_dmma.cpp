// emu_dmma.cpp — the FP64 tensor-core GEMM engine (csrc/gemm_dmma.cu) executed on the CPU
// (tests/emu/cuda_emu.h + ptx_emu.h): warp-specialised producer / consumer warps talking through
// emulated mbarriers, TMA bulk rows, cp.async gathers with barrier arrival, and the m8n8k4 MMA as a
// warp exchange.  Every schedule of the engine - classic one-tile CTAs (with the TMA-staged C
// epilogue), classic persistent CTAs (the ring rolls across tiles), the staggered two-group kernel -
// against plain loops, d and z, TMA-eligible and generic descriptors, triangular masks, triangle-only /
// Hermitian output, the block-cyclic column map.  A protocol deadlock blocks here as it would spin on
// the GPU (the test runs under a timeout).
#define CUDA_EMU_IMPLEMENTATION
#include "cuda_emu.h"

#include "../../cuda-cholesky_b200/csrc/gemm_dmma.cu"

#include <chrono>
#include <complex>
#include <random>
#include <string>

using namespace chol;

static int failures = 0, cases = 0;
// fuzz mode (./emu_dmma fuzz SEED COUNT): extra degrees of freedom the fixed sweep does not have
static bool g_fuzz = false, g_ldc_odd = false, g_beta_zero = false;
static std::complex<double> cj(std::complex<double> x) { return std::conj(x); }
static double cj(double x) { return x; }

enum Sched { CLASSIC, PERSISTENT, STAGGERED, STAGGERED_RINGC };

struct Case {
  int M, N, K;
  char opA, opB, triA, triB;  // 'N','T','C'; 0,'L','U'
  int unitA, unitB;
  char cuplo;
  int herm;
  bool aligned;
  int cyc_nb, cyc_p, cyc_off;
  int stage_c;
};

template <bool CPLX>
static void run_case(const Case &c, Sched sched) {  // (sched by value: fuzz mode may downgrade it)
  using H = typename std::conditional<CPLX, std::complex<double>, double>::type;
  using Cfg = GemmCfg<CPLX>;
  std::mt19937 g(99u + 3 * c.M + 5 * c.N + 7 * c.K + c.opA + c.opB);
  std::uniform_real_distribution<double> u(-1, 1);
  auto rnd = [&] { if constexpr (CPLX) return H(u(g), u(g)); else return H(u(g)); };
  // with a block-cyclic map the B operand has one row per GLOBAL column: size it generously
  const int nglob = c.cyc_nb ? ((c.N + c.cyc_nb - 1) / c.cyc_nb * c.cyc_p + c.cyc_off + 1) * c.cyc_nb : c.N;
  const bool an = c.opA == 'N', bn = c.opB == 'N';
  const int ar = an ? c.M : c.K, ac = an ? c.K : c.M, br = bn ? c.K : nglob, bc = bn ? nglob : c.K;
  const int pad = c.aligned ? 0 : 1;
  const int lda = ((ar + 1) / 2) * 2 + pad, ldb = ((br + 1) / 2) * 2 + pad, ldc = ((c.M + 1) / 2) * 2 + 2 * pad + 2 + (g_ldc_odd ? 1 : 0);
  std::vector<H> Abuf((size_t)lda * ac + 4), Bbuf((size_t)ldb * bc + 4), C((size_t)ldc * c.N + 2), C0;
  H *A = Abuf.data() + pad, *B = Bbuf.data() + pad;
  for (auto &x : Abuf) x = rnd();
  for (auto &x : Bbuf) x = rnd();
  for (auto &x : C) x = rnd();
  C0 = C;
  chol_gemm_desc d;
  std::memset(&d, 0, sizeof d);
  d.prec = CPLX ? 'z' : 'd';
  d.opA = c.opA; d.opB = c.opB; d.triA = c.triA; d.triB = c.triB; d.unitA = c.unitA; d.unitB = c.unitB;
  d.cuplo = c.cuplo; d.herm = c.herm;
  d.m = c.M; d.n = c.N; d.k = c.K;
  d.alpha[0] = -0.75; d.alpha[1] = CPLX ? 0.25 : 0.0;
  d.beta[0] = 1.25; d.beta[1] = CPLX ? -0.5 : 0.0;
  if (g_beta_zero) d.beta[0] = d.beta[1] = 0.0;
  d.A = A; d.lda = lda; d.B = B; d.ldb = ldb; d.C = C.data(); d.ldc = ldc; d.D = C.data(); d.ldd = ldc;
  d.cyc_nb = c.cyc_nb; d.cyc_p = c.cyc_p; d.cyc_off = c.cyc_off;
  GemmArgs<CPLX> a;
  long long tiles = 0;
  bool tma = false;
  if (fill_args<CPLX>(&d, a, tiles, tma) != 0 || tiles == 0) { std::printf("fill_args rejected a case\n"); ++failures; return; }
  a.stage_c = c.stage_c;
  if (g_fuzz) {  // a drawn schedule the descriptor cannot take falls back the way the launcher does
    if (sched == STAGGERED_RINGC && !tma) sched = STAGGERED;
    if ((sched == STAGGERED || sched == STAGGERED_RINGC) && a.K == 0) sched = CLASSIC;
  }
  const H alpha = [&] { if constexpr (CPLX) return H(d.alpha[0], d.alpha[1]); else return H(d.alpha[0]); }();
  const H beta = [&] { if constexpr (CPLX) return H(d.beta[0], d.beta[1]); else return H(d.beta[0]); }();

  auto go = [&](auto kern_classic, auto kern_pp) {
    if (sched == STAGGERED_RINGC) {
      if (!tma) { std::printf("ring-staged case is not TMA-eligible\n"); ++failures; return; }
      a.tps = 2;
      a.stage_c = 2;
      emu::launch(dim3((unsigned)((tiles + 2 * a.tps - 1) / (2 * a.tps))), dim3(Cfg::PP_THREADS), Cfg::PP_SMEM_BYTES,
                  [&] { gemm_dmma_ppc_kernel<CPLX>(a); });
    } else if (sched == STAGGERED) {
      a.tps = 2;
      emu::launch(dim3((unsigned)((tiles + 2 * a.tps - 1) / (2 * a.tps))), dim3(Cfg::PP_THREADS), Cfg::PP_SMEM_BYTES, [&] { kern_pp(a); });
    } else {
      const unsigned grid = sched == PERSISTENT ? (unsigned)std::max<long long>(1, tiles / 3) : (unsigned)tiles;
      emu::launch(dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, [&] { kern_classic(a); });
    }
  };
  if (tma) go(gemm_dmma_kernel<CPLX, 0>, gemm_dmma_pp_kernel<CPLX, 0>);
  else go(gemm_dmma_kernel<CPLX, 1>, gemm_dmma_pp_kernel<CPLX, 1>);

  auto gcolh = [&](int n) { return c.cyc_nb ? ((n / c.cyc_nb) * c.cyc_p + c.cyc_off) * c.cyc_nb + n % c.cyc_nb : n; };
  auto opa = [&](int m, int k) -> H {
    const int r = an ? m : k, col = an ? k : m;
    if ((c.triA == 'L' && r < col) || (c.triA == 'U' && r > col)) return H(0);
    H v = (c.unitA && c.triA && r == col) ? H(1) : A[r + (size_t)col * lda];
    return (CPLX && c.opA == 'C') ? cj(v) : v;
  };
  auto opb = [&](int k, int ng) -> H {  // ng = global column
    const int r = bn ? k : ng, col = bn ? ng : k;
    if ((c.triB == 'L' && r < col) || (c.triB == 'U' && r > col)) return H(0);
    H v = (c.unitB && c.triB && r == col) ? H(1) : B[r + (size_t)col * ldb];
    return (CPLX && c.opB == 'C') ? cj(v) : v;
  };
  double err = 0;
  for (int n = 0; n < c.N; ++n)
    for (int m = 0; m < c.M; ++m) {
      const int ng = gcolh(n);
      const bool touched = c.cuplo == 0 || (c.cuplo == 'L' && m >= ng) || (c.cuplo == 'U' && m <= ng);
      H want = C0[m + (size_t)n * ldc];
      if (touched) {
        H s = 0;
        for (int k = 0; k < c.K; ++k) s += opa(m, k) * opb(k, ng);
        want = alpha * s + beta * C0[m + (size_t)n * ldc];
        if constexpr (CPLX) { if (c.herm && m == ng) want = H(want.real(), 0.0); }
      }
      err = std::max(err, (double)std::abs(C[m + (size_t)n * ldc] - want));
    }
  for (int n = 0; n < c.N; ++n)
    for (int m = c.M; m < ldc; ++m)
      if (C[m + (size_t)n * ldc] != C0[m + (size_t)n * ldc]) err = 1e30;
  ++cases;
  const bool ok = err <= 1e-13 * (c.K + 8);
  std::printf("%s %-10s %s M=%d N=%d K=%d op=%c%c tri=%c%c cuplo=%c herm=%d cyc=%d/%d stage_c=%d tiles=%lld: %s (%.1e)\n", CPLX ? "z" : "d",
              sched == CLASSIC ? "classic" : (sched == PERSISTENT ? "persistent" : (sched == STAGGERED ? "staggered" : "stag+ringC")), tma ? "TMA    " : "generic", c.M, c.N, c.K, c.opA,
              c.opB, c.triA ? c.triA : '-', c.triB ? c.triB : '-', c.cuplo ? c.cuplo : '-', c.herm, c.cyc_nb, c.cyc_p, c.stage_c, tiles,
              ok ? "ok" : "MISMATCH", err);
  if (!ok) ++failures;
}

template <bool CPLX>
static void sweep() {
  const char hT = CPLX ? 'C' : 'T';
  const int BM = GemmCfg<CPLX>::BM;
  // the trailing update of POTRF: NT, aligned (TMA path), lower / upper triangle, Hermitian diagonal
  for (Sched s : {CLASSIC, PERSISTENT, STAGGERED}) {
    run_case<CPLX>({3 * BM, 192, 48, 'N', hT, 0, 0, 0, 0, 'L', 1, true, 0, 0, 0, 1}, s);
    run_case<CPLX>({2 * BM + 10, 150, 32, 'N', hT, 0, 0, 0, 0, 0, 0, true, 0, 0, 0, 1}, s);  // ragged edges
  }
  // the staggered kernel with the C tile staged through the ring: interior, diagonal-crossing and ragged tiles,
  // both triangles, the block-cyclic map, short and long k loops
  run_case<CPLX>({3 * BM, 192, 48, 'N', hT, 0, 0, 0, 0, 'L', 1, true, 0, 0, 0, 1}, STAGGERED_RINGC);
  run_case<CPLX>({3 * BM, 192, 16, 'N', hT, 0, 0, 0, 0, 'U', 1, true, 0, 0, 0, 1}, STAGGERED_RINGC);
  run_case<CPLX>({2 * BM + 10, 150, 32, 'N', hT, 0, 0, 0, 0, 0, 0, true, 0, 0, 0, 1}, STAGGERED_RINGC);
  run_case<CPLX>({4 * BM, 256, 80, 'N', hT, 0, 0, 0, 0, 0, 0, true, 0, 0, 0, 1}, STAGGERED_RINGC);
  run_case<CPLX>({5 * BM, 128, 32, 'N', hT, 0, 0, 0, 0, 'L', 1, true, 64, 2, 1, 1}, STAGGERED_RINGC);
  run_case<CPLX>({3 * BM, 192, 48, 'N', hT, 0, 0, 0, 0, 'U', 1, true, 0, 0, 0, 1}, CLASSIC);
  run_case<CPLX>({2 * BM, 128, 64, 'N', hT, 0, 0, 0, 0, 0, 0, true, 0, 0, 0, 0}, CLASSIC);   // C read from global memory
  // block-cyclic column map of the distributed update (P = 2, this rank's offset 1), lower
  run_case<CPLX>({5 * BM, 128, 32, 'N', hT, 0, 0, 0, 0, 'L', 1, true, 64, 2, 1, 1}, CLASSIC);
  run_case<CPLX>({5 * BM, 128, 32, 'N', hT, 0, 0, 0, 0, 'L', 1, true, 64, 2, 0, 1}, STAGGERED);
  // generic loader: other ops, unaligned operands, K not a multiple of the stage depth
  for (Sched s : {CLASSIC, STAGGERED}) {
    run_case<CPLX>({BM + 20, 100, 37, 'N', 'N', 0, 0, 0, 0, 0, 0, true, 0, 0, 0, 1}, s);
    run_case<CPLX>({BM + 20, 100, 37, hT, 'N', 0, 0, 0, 0, 0, 0, false, 0, 0, 0, 1}, s);
    run_case<CPLX>({BM + 20, 100, 40, 'N', hT, 0, 0, 0, 0, 'L', 1, false, 0, 0, 0, 1}, s);
  }
  // triangular operands (TRMM / LAUUM / TRTRI forms) with the structural-zero k-skip and unit diagonals
  run_case<CPLX>({2 * BM + 5, 70, 2 * BM + 5, 'N', 'N', 'L', 0, 1, 0, 0, 0, true, 0, 0, 0, 1}, CLASSIC);
  run_case<CPLX>({2 * BM + 5, 70, 2 * BM + 5, hT, 'N', 'U', 0, 0, 0, 0, 0, true, 0, 0, 0, 1}, STAGGERED);
  run_case<CPLX>({70, 150, 150, 'N', hT, 0, 'L', 0, 1, 0, 0, true, 0, 0, 0, 1}, CLASSIC);
  run_case<CPLX>({100, 100, 140, hT, 'N', 'L', 'L', 0, 0, 'L', 1, true, 0, 0, 0, 1}, CLASSIC);
  // degenerate: K = 0 (D = beta C), one element
  run_case<CPLX>({40, 30, 0, 'N', hT, 0, 0, 0, 0, 0, 0, true, 0, 0, 0, 1}, CLASSIC);
  run_case<CPLX>({1, 1, 1, 'N', 'N', 0, 0, 0, 0, 0, 0, true, 0, 0, 0, 1}, CLASSIC);
}

// random descriptors inside the launcher's contract (fill_args): shapes up to a few tiles, every op pair, aligned and
// unaligned operands, odd ldc (scalar epilogue), beta = 0, triangular operands, triangle-only / Hermitian output, the
// block-cyclic column map, all four schedules
template <bool CPLX>
static void fuzz_one(std::mt19937 &g) {
  auto ri = [&](int lo, int hi) { return lo + (int)(g() % (unsigned)(hi - lo + 1)); };
  const int BM = GemmCfg<CPLX>::BM;
  const char hT = CPLX ? 'C' : 'T';
  const char ops[3] = {'N', 'T', hT};
  Case c{};
  c.M = ri(0, 4) == 0 ? ri(1, 40) : ri(1, 3 * BM + 20);
  c.N = ri(0, 4) == 0 ? ri(1, 20) : ri(1, 210);
  c.K = ri(0, 9) == 0 ? 0 : ri(1, 100);
  c.opA = ops[ri(0, 2)];
  c.opB = ops[ri(0, 2)];
  c.aligned = ri(0, 2) != 0;
  c.stage_c = ri(0, 3) == 0 ? 0 : 1;
  const int kind = ri(0, 9);
  if (kind <= 3) {  // the POTRF / SYRK family: NT, triangle-only output
    c.opA = 'N';
    c.opB = hT;
    c.cuplo = "LU"[ri(0, 1)];
    c.herm = ri(0, 1);
    if (ri(0, 1)) c.N = c.M;
    if (ri(0, 3) != 0) {  // mostly the TMA-fed form (tma_eligible): aligned, even m / n, k a multiple of the k-tile
      c.aligned = true;
      c.M += c.M & 1;
      c.N += c.N & 1;
      c.K = GemmCfg<CPLX>::BK * ri(1, 6);
    }
    if (c.cuplo == 'L' && ri(0, 2) == 0) {  // block-cyclic map
      c.cyc_nb = 64 * ri(1, 2);
      c.cyc_p = ri(1, 3);
      c.cyc_off = ri(0, c.cyc_p - 1);
    }
  } else if (kind == 4) {  // triangular A (TRMM left): op(A) is M x K, square
    c.K = c.M;
    c.triA = "LU"[ri(0, 1)];
    c.unitA = ri(0, 1);
  } else if (kind == 5) {  // triangular B (TRMM right): op(B) is K x N, square
    c.K = c.N;
    c.triB = "LU"[ri(0, 1)];
    c.unitB = ri(0, 1);
  } else if (kind == 6) {  // plain GEMM with a block-cyclic map
    if (c.opB == 'N') c.opB = hT;
    c.cyc_nb = 64;
    c.cyc_p = ri(2, 3);
    c.cyc_off = ri(0, c.cyc_p - 1);
  }
  g_ldc_odd = ri(0, 3) == 0;
  g_beta_zero = ri(0, 5) == 0;
  run_case<CPLX>(c, (Sched)ri(0, 3));
}

int main(int argc, char **argv) {
  std::setvbuf(stdout, nullptr, _IOLBF, 0);
  if (const char *w = getenv("EMU_WATCHDOG")) {  // deadlock report: who waits on which mbarrier, then give up
    const int secs = atoi(w);
    std::thread([secs] {
      std::this_thread::sleep_for(std::chrono::seconds(secs));
      emu_mbar::dump("still running - blocked waits and barrier states:");
      std::_Exit(3);
    }).detach();
  }
  if (argc > 1 && std::string(argv[1]) == "staggered-only") {
    run_case<false>({3 * 128, 192, 48, 'N', 'T', 0, 0, 0, 0, 'L', 1, true, 0, 0, 0, 1}, STAGGERED);
    return failures;
  }
  if (argc > 3 && std::string(argv[1]) == "fuzz") {
    g_fuzz = true;
    std::mt19937 g((unsigned)atoi(argv[2]));
    for (int i = 0, n = atoi(argv[3]); i < n; ++i) {
      if (g() & 1u) fuzz_one<false>(g); else fuzz_one<true>(g);
    }
    std::printf("%d cases\n", cases);
    if (failures) { std::printf("EMU FAILED: %d mismatches\n", failures); return 1; }
    std::printf("EMU PASSED\n");
    return 0;
  }
  sweep<false>();
  sweep<true>();
  std::printf("%d cases\n", cases);
  if (failures) { std::printf("EMU FAILED: %d mismatches\n", failures); return 1; }
  std::printf("EMU PASSED\n");
  return 0;
}
