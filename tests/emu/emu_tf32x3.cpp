// emu_tf32x3.cpp — the tcgen05 3xTF32 engine (csrc/gemm_tf32x3.cu) executed on the CPU (tests/emu/cuda_emu.h +
// tcgen05_emu.h): the warp-specialised persistent kernel - TMA producer thread, MMA-issuing thread, TMEM allocation
// warp, eight epilogue warps, two shared-memory stages and two TMEM accumulators handed around through mbarriers -
// and the operand-split kernels, against plain loops in double precision.  s (real) and c (the real embedding of the
// complex product: interleaved rows, Hermitian diagonal), all of D / lower / upper, ragged edges in m, n and k,
// several tiles per CTA (both accumulators and both stages wrap), beta = 0.  The launchers are not part of the
// emulation: run_s / run_c restate their argument set-up (csrc/gemm_tf32x3.cu: tf32x3_launch, make_map).
// A protocol deadlock blocks here as it would spin on the GPU (watchdog: EMU_WATCHDOG=seconds).
//     ./emu_tf32x3            the fixed sweep          ./emu_tf32x3 fuzz SEED COUNT    random descriptors
#define CUDA_EMU_IMPLEMENTATION
#include "cuda_emu.h"

#include "../../cuda-cholesky_b200/csrc/gemm_tf32x3.cu"

#include <chrono>
#include <complex>
#include <random>
#include <string>

using namespace chol;

static int failures = 0, cases = 0;

static size_t kpitch(size_t k) { return (k + TBK - 1) / TBK * TBK; }  // chol_tf32_kpitch

// 16-byte aligned float buffer (the planes are TMA sources)
struct Planes {
  std::vector<float> store;
  float *p;
  explicit Planes(size_t n) : store(n + 8, 7.f) {
    uintptr_t a = reinterpret_cast<uintptr_t>(store.data());
    p = reinterpret_cast<float *>((a + 15) & ~uintptr_t(15));
  }
};

static void fill_targs(TArgs &p, size_t m, size_t n, size_t k, float alpha, float beta, const float *C, size_t ldc, float *D, size_t ldd,
                       int cuplo, int rowshift, int herm) {  // tf32x3_launch
  p.M = (int)m;
  p.N = (int)n;
  p.nk = (int)((k + TBK - 1) / TBK);
  p.NI = (int)((m + TBM - 1) / TBM);
  p.NJ = (int)((n + TBN - 1) / TBN);
  p.cuplo = cuplo;
  p.rs = rowshift;
  p.herm = herm;
  p.ntiles = t_count_tiles(p.cuplo, p.rs, p.NI, p.NJ);
  p.alpha = alpha;
  p.beta = beta;
  p.C = C;
  p.ldc = (long long)ldc;
  p.D = D;
  p.ldd = (long long)ldd;
}
static CUtensorMap make_map(const float *base, size_t rows, size_t kp, size_t row_pitch, unsigned box_rows) {  // make_map
  CUtensorMap m;
  if (emu_tc::encode_tiled_2d_f32_sw128(&m, base, kp, rows, row_pitch * 4, TBK, box_rows)) {
    std::printf("tensor map rejected\n");
    std::exit(2);
  }
  return m;
}
static void launch_engine(const TArgs &p, const float *Ahi, const float *Alo, size_t pitchA, size_t rowsA, const float *Bhi, const float *Blo,
                          size_t pitchB, size_t rowsB, size_t kp, int ctas) {
  if (p.ntiles == 0 || p.nk == 0) return;
  const CUtensorMap mAhi = make_map(Ahi, rowsA, kp, pitchA, TBM), mAlo = make_map(Alo, rowsA, kp, pitchA, TBM);
  const CUtensorMap mBhi = make_map(Bhi, rowsB, kp, pitchB, TBN), mBlo = make_map(Blo, rowsB, kp, pitchB, TBN);
  const int grid = p.ntiles < ctas ? p.ntiles : ctas;
  emu::launch(dim3((unsigned)grid), dim3(T_THREADS), T_SMEM_BYTES, [&] { gemm_tf32x3_kernel(mAhi, mAlo, mBhi, mBlo, p); });
}

// ---- s: D = alpha A B^T + beta C ----
struct CaseS { int m, n, k, cuplo; float beta; int ctas, transA; };
static void run_s(const CaseS &c) {
  std::mt19937 g(77u + 3 * c.m + 5 * c.n + 7 * c.k + c.cuplo);
  std::uniform_real_distribution<float> u(-1, 1);
  const int lda = c.transA ? c.k + 3 : c.m + 1, ldb = c.n + 2, ldc = c.m + 5;
  std::vector<float> A((size_t)lda * (c.transA ? c.m : c.k) + 1), B((size_t)ldb * c.k + 1), C((size_t)ldc * c.n), C0;
  for (auto &x : A) x = u(g);
  for (auto &x : B) x = u(g);
  for (auto &x : C) x = u(g);
  C0 = C;
  auto a_at = [&](int r, int kk) { return c.transA ? A[(size_t)r * lda + kk] : A[(size_t)kk * lda + r]; };
  const size_t kp = kpitch(c.k);
  Planes Ah((size_t)c.m * kp), Al((size_t)c.m * kp), Bh((size_t)c.n * kp), Bl((size_t)c.n * kp);
  // chol_tf32_split_launch
  emu::launch(dim3((unsigned)((c.m + 31) / 32), (unsigned)(kp / 32)), dim3(256), 0,
              [&] { tf32_split_kernel(c.transA, c.m, c.k, (int)kp, A.data(), (long long)lda, Ah.p, Al.p); });
  emu::launch(dim3((unsigned)((c.n + 31) / 32), (unsigned)(kp / 32)), dim3(256), 0,
              [&] { tf32_split_kernel(0, c.n, c.k, (int)kp, B.data(), (long long)ldb, Bh.p, Bl.p); });
  // the planes: hi + lo reproduces the operand to 2^-22, the pad columns are zero
  double split_err = 0;
  for (int r = 0; r < c.m; ++r)
    for (size_t kk = 0; kk < kp; ++kk) {
      const double want = (int)kk < c.k ? a_at(r, (int)kk) : 0.0;
      split_err = std::max(split_err, std::abs((double)Ah.p[r * kp + kk] + (double)Al.p[r * kp + kk] - want));
    }
  const float alpha = -0.75f;
  TArgs p;
  fill_targs(p, c.m, c.n, c.k, alpha, c.beta, C.data(), ldc, C.data(), ldc, c.cuplo, 0, 0);
  launch_engine(p, Ah.p, Al.p, kp, c.m, Bh.p, Bl.p, kp, c.n, kp, c.ctas);
  double err = 0;
  for (int j = 0; j < c.n; ++j)
    for (int i = 0; i < c.m; ++i) {
      const bool touched = c.cuplo == 0 || (c.cuplo == 1 && i >= j) || (c.cuplo == 2 && i <= j);
      double want = C0[i + (size_t)j * ldc];
      if (touched) {
        double s = 0;
        for (int kk = 0; kk < c.k; ++kk) s += (double)a_at(i, kk) * (double)B[(size_t)kk * ldb + j];
        want = (double)alpha * s + (double)c.beta * C0[i + (size_t)j * ldc];
      }
      const double d = std::abs((double)C[i + (size_t)j * ldc] - want);
      err = std::max(err, std::isnan(d) ? 1e30 : d);
    }
  for (int j = 0; j < c.n; ++j)
    for (int i = c.m; i < ldc; ++i)
      if (C[i + (size_t)j * ldc] != C0[i + (size_t)j * ldc]) err = 1e30;  // padding rows untouched
  ++cases;
  const bool ok = err <= 1e-6 * (c.k + 8) && split_err <= 3e-7;
  std::printf("s m=%d n=%d k=%d cuplo=%d beta=%g ctas=%d transA=%d tiles=%d nk=%d: %s (%.1e, split %.1e)\n", c.m, c.n, c.k, c.cuplo, c.beta, c.ctas,
              c.transA, p.ntiles, p.nk, ok ? "ok" : "MISMATCH", err, split_err);
  if (!ok) ++failures;
}

// ---- c: D = alpha A B^H + beta C on interleaved complex storage, A = W[i0 : i0 + m], B = W[j0 : j0 + n] ----
struct CaseC { int rows, k, i0, m, j0, n, cuplo, herm; float beta; int ctas; };
static void run_c(const CaseC &c) {
  using Z = std::complex<float>;
  std::mt19937 g(91u + 3 * c.m + 5 * c.n + 7 * c.k + c.cuplo + c.i0);
  std::uniform_real_distribution<float> u(-1, 1);
  const int ldw = c.rows + 1, ldc = c.m + 3;
  std::vector<Z> W((size_t)ldw * c.k + 1), C((size_t)ldc * c.n), C0;
  for (auto &x : W) x = Z(u(g), u(g));
  for (auto &x : C) x = Z(u(g), u(g));
  C0 = C;
  const size_t kp = kpitch(2 * (size_t)c.k);
  Planes Hi((size_t)2 * c.rows * kp), Lo((size_t)2 * c.rows * kp);
  // chol_tf32_split_c_launch
  emu::launch(dim3((unsigned)((c.rows + 31) / 32), (unsigned)((kp / 2 + 31) / 32)), dim3(256), 0, [&] {
    tf32_split_c_kernel(c.rows, c.k, (int)kp, reinterpret_cast<const float2 *>(W.data()), (long long)ldw, Hi.p, Lo.p);
  });
  const float alpha = -1.f;
  TArgs p;  // chol_tf32x3_herk_launch: tf32x3_launch(2 m, n, 2 k, ..., pitchA = kp, pitchB = 2 kp, 2 ldc, rowshift 1, herm)
  fill_targs(p, 2 * (size_t)c.m, c.n, 2 * (size_t)c.k, alpha, c.beta, reinterpret_cast<const float *>(C.data()), 2 * (size_t)ldc,
             reinterpret_cast<float *>(C.data()), 2 * (size_t)ldc, c.cuplo, 1, c.herm);
  launch_engine(p, Hi.p + (size_t)2 * c.i0 * kp, Lo.p + (size_t)2 * c.i0 * kp, kp, 2 * (size_t)c.m, Hi.p + (size_t)2 * c.j0 * kp,
                Lo.p + (size_t)2 * c.j0 * kp, 2 * kp, c.n, kp, c.ctas);
  double err = 0;
  for (int j = 0; j < c.n; ++j)
    for (int i = 0; i < c.m; ++i) {
      const bool touched = c.cuplo == 0 || (c.cuplo == 1 && i >= j) || (c.cuplo == 2 && i <= j);
      std::complex<double> want = C0[i + (size_t)j * ldc];
      if (touched) {
        std::complex<double> s = 0;
        for (int kk = 0; kk < c.k; ++kk)
          s += std::complex<double>(W[(size_t)kk * ldw + c.i0 + i]) * std::conj(std::complex<double>(W[(size_t)kk * ldw + c.j0 + j]));
        want = (double)alpha * s + (double)c.beta * std::complex<double>(C0[i + (size_t)j * ldc]);
        if (c.herm && i == j) want = std::complex<double>(want.real(), 0.0);
      }
      const double d = std::abs(std::complex<double>(C[i + (size_t)j * ldc]) - want);
      err = std::max(err, std::isnan(d) ? 1e30 : d);
    }
  for (int j = 0; j < c.n; ++j)
    for (int i = c.m; i < ldc; ++i)
      if (C[i + (size_t)j * ldc] != C0[i + (size_t)j * ldc]) err = 1e30;
  ++cases;
  const bool ok = err <= 2e-6 * (2 * c.k + 8);
  std::printf("c rows=%d k=%d A=[%d,+%d) B=[%d,+%d) cuplo=%d herm=%d beta=%g ctas=%d tiles=%d nk=%d: %s (%.1e)\n", c.rows, c.k, c.i0, c.m, c.j0, c.n,
              c.cuplo, c.herm, c.beta, c.ctas, p.ntiles, p.nk, ok ? "ok" : "MISMATCH", err);
  if (!ok) ++failures;
}

static void sweep() {
  // the POTRF trailing update: A = B, triangle only; one CTA walks several tiles (both accumulators and stages wrap)
  run_s({300, 300, 64, 1, 1.f, 1, 0});
  run_s({300, 300, 64, 2, 1.f, 2, 0});
  run_s({521, 521, 96, 1, 1.f, 3, 0});
  run_s({265, 265, 256, 1, 1.f, 2, 0});   // long k loop: the two stages wrap four times per tile
  // all of D, ragged in every dimension, k not a multiple of the k-tile, beta = 0, transposed operand storage
  run_s({130, 270, 37, 0, 1.f, 2, 0});
  run_s({257, 100, 33, 0, 0.f, 1, 0});
  run_s({100, 513, 8, 0, -0.5f, 3, 1});
  run_s({1, 1, 1, 0, 1.f, 1, 0});
  run_s({128, 256, 32, 0, 1.f, 1, 0});    // exactly one tile, one k step
  run_s({129, 257, 32, 2, 1.f, 2, 1});
  // c: HERK of a row range of the panel (A = B: the update of A33), Hermitian diagonal, both triangles
  run_c({300, 40, 20, 280, 20, 280, 1, 1, 1.f, 2});
  run_c({300, 40, 20, 280, 20, 280, 2, 1, 1.f, 1});
  run_c({200, 70, 0, 200, 0, 200, 1, 1, 1.f, 3});
  // c: the look-ahead column (A below B), all of D, ragged
  run_c({330, 24, 64, 266, 0, 64, 0, 0, 1.f, 2});
  run_c({130, 17, 1, 129, 0, 70, 0, 0, 0.f, 1});
  run_c({65, 5, 0, 65, 0, 65, 1, 1, 1.f, 1});
}

static void fuzz(unsigned seed, int count) {
  std::mt19937 g(seed);
  auto ri = [&](int lo, int hi) { return lo + (int)(g() % (unsigned)(hi - lo + 1)); };
  const float betas[4] = {1.f, 1.f, 0.f, -0.5f};
  for (int it = 0; it < count; ++it) {
    if (ri(0, 2) != 0) {
      CaseS c{};
      c.m = ri(0, 4) == 0 ? ri(1, 40) : ri(1, 540);
      c.n = ri(0, 4) == 0 ? ri(1, 40) : ri(1, 540);
      c.k = ri(0, 3) == 0 ? ri(1, 31) : ri(1, 200);
      c.cuplo = ri(0, 2);
      if (c.cuplo && ri(0, 2)) c.n = c.m;
      c.beta = betas[ri(0, 3)];
      c.ctas = ri(1, 3);
      c.transA = ri(0, 3) == 0;
      run_s(c);
    } else {
      CaseC c{};
      c.k = ri(0, 3) == 0 ? ri(1, 15) : ri(1, 100);
      c.m = ri(1, 270);
      c.n = ri(1, 270);
      c.cuplo = ri(0, 2);
      c.herm = 0;
      c.i0 = ri(0, 70);
      c.j0 = ri(0, 70);
      if (c.cuplo) {  // triangle-only output: the diagonal of D is the diagonal of W W^H
        c.j0 = c.i0;
        c.herm = ri(0, 1);
        if (ri(0, 1)) c.n = c.m;
      }
      c.rows = std::max(c.i0 + c.m, c.j0 + c.n) + ri(0, 5);
      c.beta = betas[ri(0, 3)];
      c.ctas = ri(1, 3);
      run_c(c);
    }
  }
}

int main(int argc, char **argv) {
  std::setvbuf(stdout, nullptr, _IOLBF, 0);
  if (const char *w = getenv("EMU_WATCHDOG")) {
    const int secs = atoi(w);
    std::thread([secs] {
      std::this_thread::sleep_for(std::chrono::seconds(secs));
      emu_mbar::dump("still running - blocked waits and barrier states:");
      std::_Exit(3);
    }).detach();
  }
  if (argc > 3 && std::string(argv[1]) == "fuzz") fuzz((unsigned)atoi(argv[2]), atoi(argv[3]));
  else sweep();
  std::printf("%d cases\n", cases);
  if (failures) { std::printf("EMU FAILED: %d mismatches\n", failures); return 1; }
  std::printf("EMU PASSED\n");
  return 0;
}
