// gemm_ffma.cu — the FP32 GEMM engine (s and c) for sm_100a: FFMA register tiles.
//
// Same descriptor, scheduler and operand semantics as the FP64 DMMA engine (gemm_dmma.cu); it is
// what cuSgemm2/cuCgemm2, cuSsyrk/cuCherk, cuStrmm2/cuCtrmm2 and the s/c LAPACK drivers run on
// (reference: blas/sgemm.c:148, cgemm.c:226, ssyrk.c:163, cherk.c:190).  north_star allows "FP32
// FFMA or a 3xTF32 split" for s/c; this is the FFMA choice: exact FP32 products, FP32 accumulate,
// so the s/c tolerance is the plain FP32 one.
//   * CTA tile 128x128 (s) / 64x64 (c), BK = 16, 256 threads, 8x8 (s) / 4x4-complex (c) register
//     tile per thread, operands in shared memory as [k][m] / [k][n], 128-bit shared loads;
//   * 4-stage cp.async (LDGSTS) pipeline straight into dynamic shared memory (66 KB, 2 CTAs/SM), one
//     __syncthreads per k-tile; for the hot shape (interior NT tile, aligned operands) the loader is
//     hoisted out of the k loop: fixed shared slots and running source pointers, 4 cp.async + 2 pointer
//     bumps per k-tile and thread.  (Round 1 shipped BK = 8: FFMA held ~73 % of the issue slots; BK = 16
//     measured 37.3 -> 43.6 TFLOP/s on spotrf N = 32768, profiles/r02_staged_ab.txt.)
//   * the trailing updates of spotrf run on the tensor cores instead (gemm_tf32x3.cu);
//   * any transpose / conjugate / triangular mask handled by the loader, k-ranges that are
//     structurally zero are skipped, triangular (cuplo) grids launch only the needed tiles.
#include <stdlib.h>
#include <atomic>

#include "gemm_ffma.cuh"

namespace chol {

template <typename T>
__global__ void __launch_bounds__(256, 2) gemm_ffma_kernel(const __grid_constant__ FArgs<T> p) {
  using Cfg = FCfg<T>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = 16, LDS = Cfg::LDS, VW = Cfg::VW, TM = Cfg::TM, TN = Cfg::TN;
  constexpr int ST = 4;  // cp.async stages
  CHOL_DYN_SMEM(f_dyn_smem);
  typedef T Stage[BK * LDS];
  Stage *As = reinterpret_cast<Stage *>(f_dyn_smem), *Bs = As + ST;

  int bi, bj;
  if (!tile_coords<BM, BN>(8, p.cuplo, p.NI, p.NJ, p.cm, (long long)blockIdx.x, bi, bj)) return;
  const int m0 = bi * BM, n0 = bj * BN;
  const int g0 = (int)gcol(p.cm, n0);
  int kb = 0, ke = p.K;
  if (p.triA) {
    const bool lower = (p.triA == 1) == (p.opA == 0);
    if (lower) ke = min(ke, m0 + BM); else kb = max(kb, m0);
  }
  if (p.triB) {
    const bool kgen = (p.triB == 1) == (p.opB == 0);
    if (kgen) kb = max(kb, g0); else ke = min(ke, g0 + BN);
  }
  kb = (kb / BK) * BK;
  const int nk = ke > kb ? (ke - kb + BK - 1) / BK : 0;

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int rowsA = min(BM, p.M - m0), rowsB = min(BN, p.N - n0);
  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = zero_of<T>();

  // Hoisted loader for the hot shape (interior tile, op(A) = N, op(B) = T/C, no masks, 16-byte
  // aligned): every thread owns two fixed 16-byte slots per operand and k-tile; its source pointers
  // advance by a constant stride, so a k-tile costs 4 cp.async + 2 pointer bumps instead of the
  // generic loader's predicates and 64-bit address arithmetic.
  constexpr int VPR = BM / VW;                 // 16-byte vectors per k-row (BM == BN): 32 for s and c
  constexpr int KPT = 256 / VPR;               // k-rows covered by one pass of the 256 threads: 8
  static_assert(BM == BN && 256 % VPR == 0 && BK == 2 * KPT, "hoisted loader layout: two passes per k-tile");
  const bool hoist = rowsA == BM && rowsB == BN && p.opA == 0 && p.opB != 0 && !p.triA && !p.triB && p.vecA && p.vecB;
  const int rv = tid % VPR, kk0 = tid / VPR;
  const int doff = kk0 * LDS + rv * VW;        // slot of the first pass; the second pass is KPT k-rows further
  const T *pa = p.A + (long long)(kb + kk0) * p.lda + m0 + rv * VW;  // running source pointers (one per operand)
  const T *pb = p.B + (long long)(kb + kk0) * p.ldb + g0 + rv * VW;
  const long long halfA = (long long)KPT * p.lda, halfB = (long long)KPT * p.ldb;

  auto issue = [&](int kt) {  // called with kt = 0, 1, 2, ... in order
    if (kt < nk) {
      const int s = kt % ST, k0 = kb + kt * BK;
      if (hoist && k0 + BK <= ke) {
        f_cp_async<16>(&As[s][doff], pa);
        f_cp_async<16>(&As[s][doff + KPT * LDS], pa + halfA);
        f_cp_async<16>(&Bs[s][doff], pb);
        f_cp_async<16>(&Bs[s][doff + KPT * LDS], pb + halfB);
      } else {
        f_fill<T, BM, BK, LDS>(As[s], p.A, p.lda, p.opA, p.triA, p.unitA, false, m0, p.M, rowsA, k0, ke, p.vecA, tid);
        f_fill<T, BN, BK, LDS>(Bs[s], p.B, p.ldb, p.opB, p.triB, p.unitB, true, g0, p.N, rowsB, k0, ke, p.vecB, tid);
      }
      pa += 2 * halfA;
      pb += 2 * halfB;
    }
    CHOL_CP_ASYNC_COMMIT();
  };
#pragma unroll
  for (int kt = 0; kt < ST - 1; ++kt) issue(kt);

  for (int kt = 0; kt < nk; ++kt) {
    CHOL_CP_ASYNC_WAIT(ST - 2);
    __syncthreads();          // stage kt has landed for everyone; stage kt-1 is no longer being read
    issue(kt + ST - 1);       // refill the slot that stage kt-1 used
    const int cur = kt % ST;
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[TM], b[TN];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 va = *reinterpret_cast<const float4 *>(&As[cur][kk * LDS + h * (BM / 2) + tx * VW]);
        const float4 vb = *reinterpret_cast<const float4 *>(&Bs[cur][kk * LDS + h * (BN / 2) + ty * VW]);
        if constexpr (Cfg::CPLX) {
          a[h * 2] = make_float2(va.x, va.y); a[h * 2 + 1] = make_float2(va.z, va.w);
          b[h * 2] = make_float2(vb.x, vb.y); b[h * 2 + 1] = make_float2(vb.z, vb.w);
        } else {
          a[h * 4] = va.x; a[h * 4 + 1] = va.y; a[h * 4 + 2] = va.z; a[h * 4 + 3] = va.w;
          b[h * 4] = vb.x; b[h * 4 + 1] = vb.y; b[h * 4 + 2] = vb.z; b[h * 4 + 3] = vb.w;
        }
      }
      if constexpr (Cfg::CPLX) {  // conjugation of raw cp.async copies happens here
        if (p.conjA_regs) {
#pragma unroll
          for (int i = 0; i < TM; ++i) a[i].y = -a[i].y;
        }
        if (p.conjB_regs) {
#pragma unroll
          for (int j = 0; j < TN; ++j) b[j].y = -b[j].y;
        }
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) fma_acc(acc[i][j], a[i], b[j]);
    }
  }

  // epilogue: thread owns rows m0 + h*(BM/2) + tx*VW + v, columns n0 + h2*(BN/2) + ty*VW + w
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const long long n = n0 + (j / VW) * (BN / 2) + ty * VW + (j % VW);
    if (n >= p.N) continue;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const long long m = m0 + (i / VW) * (BM / 2) + tx * VW + (i % VW);
      if (m >= p.M) continue;
      if (p.cuplo == 1 && m < n - n0 + g0) continue;
      if (p.cuplo == 2 && m > n - n0 + g0) continue;
      T o = mul(p.alpha, acc[i][j]);
      if (!p.beta_zero) fma_acc(o, p.beta, p.C[n * p.ldc + m]);
      if constexpr (Cfg::CPLX) { if (p.herm && m == n - n0 + g0) o.y = 0.f; }
      p.D[n * p.ldd + m] = o;
    }
  }
}

#ifndef CHOL_CPU_EMU  // the launchers are not part of the CPU emulation (tests/emu)
template <typename T>
static int launch_f(const chol_gemm_desc *d) {
  using Cfg = FCfg<T>;
  using Real = typename ScalarOf<T>::Real;
  if (d->m == 0 || d->n == 0) return 0;
  FArgs<T> a;
  a.A = (const T *)d->A; a.B = (const T *)d->B; a.C = (const T *)d->C; a.D = (T *)d->D;
  a.lda = (long long)d->lda; a.ldb = (long long)d->ldb; a.ldc = (long long)d->ldc; a.ldd = (long long)d->ldd;
  a.M = (int)d->m; a.N = (int)d->n; a.K = (int)d->k;
  const bool alpha_zero = d->alpha[0] == 0.0 && (!Cfg::CPLX || d->alpha[1] == 0.0);
  if (alpha_zero) a.K = 0;
  a.alpha = mk(T(), (Real)d->alpha[0], (Real)d->alpha[1]);
  a.beta = mk(T(), (Real)d->beta[0], (Real)d->beta[1]);
  a.beta_zero = d->beta[0] == 0.0 && (!Cfg::CPLX || d->beta[1] == 0.0);
  a.opA = op_code(d->opA); a.opB = op_code(d->opB);
  if (!Cfg::CPLX) { if (a.opA == 2) a.opA = 1; if (a.opB == 2) a.opB = 1; }
  a.triA = tri_code(d->triA); a.triB = tri_code(d->triB);
  a.unitA = d->unitA; a.unitB = d->unitB;
  a.cuplo = tri_code(d->cuplo);
  a.herm = Cfg::CPLX ? d->herm : 0;
  a.NI = (a.M + Cfg::BM - 1) / Cfg::BM;
  a.NJ = (a.N + Cfg::BN - 1) / Cfg::BN;
  a.vec_ok = 0;
  a.vecA = (((uintptr_t)d->A & 15) == 0) && (d->lda % Cfg::VW == 0);
  a.vecB = (((uintptr_t)d->B & 15) == 0) && (d->ldb % Cfg::VW == 0);
  a.conjA_regs = Cfg::CPLX && d->opA == 'C';
  a.conjB_regs = Cfg::CPLX && d->opB == 'C';
  a.cm.cnb = d->cyc_nb; a.cm.cP = d->cyc_p; a.cm.coff = d->cyc_off;
  if (a.cm.cnb && (a.cm.cnb % Cfg::BN || a.cuplo == 2 || d->triA || d->triB || a.cm.cP < 1)) return (int)cudaErrorInvalidValue;
  const long long tiles = total_tiles<Cfg::BM, Cfg::BN>(8, a.cuplo, a.NI, a.NJ, a.cm);
  if (tiles <= 0) return 0;
  if (tiles > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  constexpr int bytes = 2 * 4 * 16 * Cfg::LDS * (int)sizeof(T);
  static std::atomic<int> attr_set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63].load(std::memory_order_acquire)) {
    const cudaError_t e = cudaFuncSetAttribute(gemm_ffma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev & 63].store(1, std::memory_order_release);
  }
  gemm_ffma_kernel<T><<<(unsigned)tiles, 256, bytes, (cudaStream_t)d->stream>>>(a);
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}

#endif  // CHOL_CPU_EMU

}  // namespace chol

#ifndef CHOL_CPU_EMU
int chol_gemm_fp32_launch(const chol_gemm_desc *d) {
  if (d->prec == 's') return chol::launch_f<float>(d);
  if (d->prec == 'c') return chol::launch_f<float2>(d);
  return (int)cudaErrorInvalidValue;
}
#endif  // CHOL_CPU_EMU
