// gemm_small.cu — small-tile FP64 GEMM for the latency-bound GEMMs of the panel chain (measured on B200:
// dpotrf N = 16384 25.7 -> 26.9 TFLOP/s when it was introduced, profiles/r02_staged_ab.txt).
//
// Why: the diagonal-block recursion of POTRF (host/lapack.c chol_diag_factor_ev) issues GEMMs
// such as 384 x 128 x 128 or 256 x 256 x 128.  On the 128 x 64 DMMA tiles they are 6-16 CTAs: 6-16 of
// 148 SMs, each needing >= 16k cycles of its FP64 pipe per tile (profiles/r01_static_analysis.md §3),
// ~25 us per launch on the serial panel chain.  The plain DFMA rate of a B200 SM is close to its
// DMMA rate, so for these shapes nothing is lost by leaving the tensor-core path: 32 x 32 tiles of
// register-blocked DFMA spread one launch over 50-150 SMs and cut the per-launch time to the
// launch + pipeline-fill floor.
//
// Scope (everything else stays on the DMMA engine): prec 'd', no triangular operand masks, no
// block-cyclic map; any op(A)/op(B), any alignment, alpha/beta, C != D, cuplo (triangle-only
// output, diagonal-crossing tiles masked per element).
//   CTA 128 threads = 16 (m) x 8 (n), thread tile 2 x 4, CTA tile 32 x 32, k-chunks of 32 staged in
//   shared memory as [k][m] / [k][n] (pitch 34: 16-byte aligned rows), next chunk prefetched into
//   registers while the current one is multiplied.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "chol_kernels.h"
#include "launch_count.h"

namespace chol {

struct SmallArgs {
  const double *A, *B, *C;
  double *D;
  long long lda, ldb, ldc, ldd;
  int M, N, K;
  double alpha, beta;
  int opA, opB;  // 0: N, 1: T (C = T for real data)
  int cuplo;     // 0 all, 1 lower, 2 upper
  int beta_zero;
};

constexpr int S_BM = 32, S_BN = 32, S_BK = 32, S_LD = 34, S_THREADS = 128;  // k-chunks of 32: the panel-chain shapes (k = 128) are four chunks
constexpr int S_LPT = S_BM * S_BK / S_THREADS;  // operand elements per thread and chunk

__global__ void __launch_bounds__(S_THREADS) dgemm_small_kernel(const __grid_constant__ SmallArgs p) {
  __shared__ __align__(16) double As[S_BK * S_LD];
  __shared__ __align__(16) double Bs[S_BK * S_LD];
  const int m0 = blockIdx.x * S_BM, n0 = blockIdx.y * S_BN;
  // triangle-only output: tiles entirely outside the triangle have nothing to do
  if (p.cuplo == 1 && m0 + S_BM - 1 < n0) return;
  if (p.cuplo == 2 && m0 > n0 + S_BN - 1) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;  // rows 2*tx, 2*tx+1; columns 4*ty .. 4*ty+3

  // loader coordinates: S_BM * S_BK elements per operand chunk, S_LPT per thread; the fast index follows the
  // stored (contiguous) direction so that global reads coalesce
  int a_mm[S_LPT], a_kk[S_LPT], b_nn[S_LPT], b_kk[S_LPT];
#pragma unroll
  for (int r = 0; r < S_LPT; ++r) {
    const int idx = tid + S_THREADS * r;
    if (p.opA == 0) { a_mm[r] = idx % S_BM; a_kk[r] = idx / S_BM; } else { a_kk[r] = idx % S_BK; a_mm[r] = idx / S_BK; }
    if (p.opB != 0) { b_nn[r] = idx % S_BN; b_kk[r] = idx / S_BN; } else { b_kk[r] = idx % S_BK; b_nn[r] = idx / S_BK; }
  }
  auto fetch = [&](int k0, double (&ra)[S_LPT], double (&rb)[S_LPT]) {
#pragma unroll
    for (int r = 0; r < S_LPT; ++r) {
      const long long m = m0 + a_mm[r], n = n0 + b_nn[r], ka = k0 + a_kk[r], kb = k0 + b_kk[r];
      // op(A)(m, k): N -> A[k*lda + m], T -> A[m*lda + k];  op(B)(k, n): N -> B[n*ldb + k], T -> B[k*ldb + n]
      ra[r] = (m < p.M && ka < p.K) ? (p.opA == 0 ? p.A[ka * p.lda + m] : p.A[m * p.lda + ka]) : 0.0;
      rb[r] = (n < p.N && kb < p.K) ? (p.opB == 0 ? p.B[n * p.ldb + kb] : p.B[kb * p.ldb + n]) : 0.0;
    }
  };

  double acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  // the C tile is requested before the k loop: in the panel chain these launches are a few microseconds long and a
  // load issued after the loop would add a full memory round trip to every one of them
  double cpre[2][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long n = n0 + 4 * ty + j, m = m0 + 2 * tx + i;
      const bool keep = n < p.N && m < p.M && !(p.cuplo == 1 && m < n) && !(p.cuplo == 2 && m > n);
      cpre[i][j] = (keep && !p.beta_zero) ? p.C[n * p.ldc + m] : 0.0;
    }
  double ra[S_LPT], rb[S_LPT];
  const int nk = (p.K + S_BK - 1) / S_BK;
  if (nk > 0) fetch(0, ra, rb);
  for (int kt = 0; kt < nk; ++kt) {
#pragma unroll
    for (int r = 0; r < S_LPT; ++r) {
      As[a_kk[r] * S_LD + a_mm[r]] = ra[r];
      Bs[b_kk[r] * S_LD + b_nn[r]] = rb[r];
    }
    __syncthreads();
    if (kt + 1 < nk) fetch((kt + 1) * S_BK, ra, rb);  // in flight while this chunk is multiplied
#pragma unroll
    for (int kk = 0; kk < S_BK; ++kk) {
      const double2 a = *reinterpret_cast<const double2 *>(&As[kk * S_LD + 2 * tx]);
      const double2 b0 = *reinterpret_cast<const double2 *>(&Bs[kk * S_LD + 4 * ty]);
      const double2 b1 = *reinterpret_cast<const double2 *>(&Bs[kk * S_LD + 4 * ty + 2]);
      acc[0][0] = fma(a.x, b0.x, acc[0][0]); acc[0][1] = fma(a.x, b0.y, acc[0][1]);
      acc[0][2] = fma(a.x, b1.x, acc[0][2]); acc[0][3] = fma(a.x, b1.y, acc[0][3]);
      acc[1][0] = fma(a.y, b0.x, acc[1][0]); acc[1][1] = fma(a.y, b0.y, acc[1][1]);
      acc[1][2] = fma(a.y, b1.x, acc[1][2]); acc[1][3] = fma(a.y, b1.y, acc[1][3]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long n = n0 + 4 * ty + j;
    if (n >= p.N) continue;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long m = m0 + 2 * tx + i;
      if (m >= p.M) continue;
      if (p.cuplo == 1 && m < n) continue;
      if (p.cuplo == 2 && m > n) continue;
      double o = p.alpha * acc[i][j];
      if (!p.beta_zero) o += p.beta * cpre[i][j];
      p.D[n * p.ldd + m] = o;
    }
  }
}

}  // namespace chol

#ifndef CHOL_CPU_EMU  // the launchers are not part of the CPU emulation (tests/emu)
// 1 if the descriptor is one the small-tile kernel takes (and the switch is on)
extern "C" int chol_gemm_small_eligible(const chol_gemm_desc *d) {
  if (d->prec != 'd') return 0;
  if (d->triA || d->triB || d->cyc_nb || d->force_generic) return 0;
  if (d->m == 0 || d->n == 0) return 0;
  // "small" = the 128 x 64 DMMA grid would leave most SMs idle, and the problem is not k-heavy
  const size_t dmma_tiles = ((d->m + 127) / 128) * ((d->n + 63) / 64);
  return dmma_tiles <= 48 && d->k <= 1024 && d->m <= 65535u * 32u && d->n <= 65535u * 32u;
}

extern "C" int chol_gemm_small_launch(const chol_gemm_desc *d) {
  using namespace chol;
  SmallArgs a;
  a.A = (const double *)d->A; a.B = (const double *)d->B; a.C = (const double *)d->C; a.D = (double *)d->D;
  a.lda = (long long)d->lda; a.ldb = (long long)d->ldb; a.ldc = (long long)d->ldc; a.ldd = (long long)d->ldd;
  a.M = (int)d->m; a.N = (int)d->n; a.K = (d->alpha[0] == 0.0) ? 0 : (int)d->k;  // alpha = 0: A and B are never read
  a.alpha = d->alpha[0];
  a.beta = d->beta[0];
  a.beta_zero = (d->beta[0] == 0.0);
  a.opA = (d->opA == 'N') ? 0 : 1;
  a.opB = (d->opB == 'N') ? 0 : 1;
  a.cuplo = d->cuplo == 'L' ? 1 : (d->cuplo == 'U' ? 2 : 0);
  dim3 grid((unsigned)((d->m + S_BM - 1) / S_BM), (unsigned)((d->n + S_BN - 1) / S_BN));
  dgemm_small_kernel<<<grid, S_THREADS, 0, (cudaStream_t)d->stream>>>(a);
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}
#endif  // CHOL_CPU_EMU
