// gemm_ffma.cu — the FP32 GEMM engine (s and c) for sm_100a: FFMA register tiles.
//
// Same descriptor, scheduler and operand semantics as the FP64 DMMA engine (gemm_dmma.cu); it is
// what cuSgemm2/cuCgemm2, cuSsyrk/cuCherk, cuStrmm2/cuCtrmm2 and the s/c LAPACK drivers run on
// (reference: blas/sgemm.c:148, cgemm.c:226, ssyrk.c:163, cherk.c:190).  north_star allows "FP32
// FFMA or a 3xTF32 split" for s/c; this is the FFMA choice: exact FP32 products, FP32 accumulate,
// so the s/c tolerance is the plain FP32 one.
//   * CTA tile 128x128 (s) / 64x64 (c), BK = 8, 256 threads, 8x8 (s) / 4x4-complex (c) register
//     tile per thread, operands in shared memory as [k][m] / [k][n], 128-bit shared loads;
//   * 4-stage cp.async (LDGSTS) pipeline straight into shared memory, one __syncthreads per k-tile;
//   * any transpose / conjugate / triangular mask handled by the loader, k-ranges that are
//     structurally zero are skipped, triangular (cuplo) grids launch only the needed tiles.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>

#include "chol_kernels.h"
#include "launch_count.h"
#include "gemm_common.cuh"

namespace chol {

template <typename T>
struct FCfg {
  static constexpr bool CPLX = ScalarOf<T>::cplx;
  static constexpr int VW = 16 / sizeof(T);  // elements per 128-bit access: 4 (float) / 2 (float2)
  static constexpr int TM = 2 * VW, TN = 2 * VW;
  static constexpr int BM = 16 * TM, BN = 16 * TN;  // 128 / 64
  static constexpr int LDS = BM + VW;  // padded pitch (elements); keeps 16-byte alignment
  static constexpr int THREADS = 256;
  static constexpr int STAGES = 4;     // cp.async stages
  // bytes of the two operand rings for a k-tile depth of BK
  static constexpr int smem_bytes(int BK) { return 2 * STAGES * BK * LDS * (int)sizeof(T); }
};

template <typename T>
struct FArgs {
  const T *A, *B, *C;
  T *D;
  long long lda, ldb, ldc, ldd;
  int M, N, K;
  T alpha, beta;
  int opA, opB, triA, triB, unitA, unitB, cuplo, herm, NI, NJ, vec_ok, beta_zero;
  int vecA, vecB;            // 16-byte cp.async along the stored rows is legal
  int conjA_regs, conjB_regs;  // complex: conjugate fragments in registers (fast path copies raw data)
  ColMap cm;
};

// true if the stored-triangle mask cannot touch the tile rows [r0, r0+rows) x k [k0, k0+bk)
__device__ __forceinline__ bool f_tri_clear(int tri, bool along_rows, int r0, int rows, int k0, int bk) {
  if (!tri) return true;
  const int row_min = along_rows ? r0 : k0, row_max = along_rows ? r0 + rows - 1 : k0 + bk - 1;
  const int col_min = along_rows ? k0 : r0, col_max = along_rows ? k0 + bk - 1 : r0 + rows - 1;
  return tri == 1 ? (row_min > col_max) : (row_max < col_min);
}
template <int BYTES>
__device__ __forceinline__ void f_cp_async(void *dst, const void *src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(src), "n"(BYTES) : "memory");
}

// One ROWS x BK operand tile -> shared [k][row] (pitch LDS).  Interior, unmasked tiles go through
// cp.async (16-byte vectors along the stored rows when aligned, element-wise otherwise); edge and
// diagonal-crossing tiles take the per-element path with masks.  256 threads cooperate.
template <typename T, int ROWS, int BK, int LDS>
__device__ __forceinline__ void f_fill(T *dst, const T *__restrict__ src, long long ld, int op, int tri, int unit, bool is_b,
                                       int r0, int rlimit, int rows_valid, int k0, int ke, bool vec16, int tid) {
  // logical element (row r0+rr, k0+kk); stored coordinates: A: op==0 -> (row, k) else (k, row); B: op!=0 -> (row, k) else (k, row)
  const bool along_rows = is_b ? (op != 0) : (op == 0);
  const bool conj = false;  // conjugation is applied to the register fragments (conjA_regs / conjB_regs)
  const bool fast = (rows_valid == ROWS) && (k0 + BK <= ke) && f_tri_clear(tri, along_rows, r0, ROWS, k0, BK);
  constexpr int VW = 16 / sizeof(T);
  if (fast) {
    if (along_rows) {
      if (vec16) {
        for (int v = tid; v < ROWS / VW * BK; v += 256) {
          const int rv = v % (ROWS / VW), kk = v / (ROWS / VW);
          f_cp_async<16>(dst + kk * LDS + rv * VW, src + (long long)(k0 + kk) * ld + r0 + rv * VW);
        }
      } else {
        for (int e = tid; e < ROWS * BK; e += 256) {
          const int rr = e % ROWS, kk = e / ROWS;
          f_cp_async<sizeof(T)>(dst + kk * LDS + rr, src + (long long)(k0 + kk) * ld + r0 + rr);
        }
      }
    } else {
      for (int e = tid; e < ROWS * BK; e += 256) {
        const int kk = e % BK, rr = e / BK;
        f_cp_async<sizeof(T)>(dst + kk * LDS + rr, src + (long long)(r0 + rr) * ld + k0 + kk);
      }
    }
    return;
  }
  for (int e = tid; e < ROWS * BK; e += 256) {
    int rr, kk;
    if (along_rows) { rr = e % ROWS; kk = e / ROWS; } else { kk = e % BK; rr = e / BK; }
    const long long r = r0 + rr, k = k0 + kk;
    const bool in = (rr < rows_valid) && (k < ke);
    (void)rlimit;
    dst[kk * LDS + rr] = along_rows ? fetch_stored<T>(src, ld, in, r, k, tri, unit, conj)
                                    : fetch_stored<T>(src, ld, in, k, r, tri, unit, conj);
  }
}

// BK = 8: the measured default (static shared memory, 33 KB).  BK = 16 (CHOL_FFMA_BK=16, dynamic
// shared memory, 66 KB): the per-k-tile loader bookkeeping (about 100 integer instructions per
// thread, SASS) and the barrier are amortised over twice as many FFMAs - 512 -> 1024 per thread.
template <typename T, int BK>
__global__ void __launch_bounds__(256, 2) gemm_ffma_kernel(const __grid_constant__ FArgs<T> p) {
  using Cfg = FCfg<T>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, LDS = Cfg::LDS, VW = Cfg::VW, TM = Cfg::TM, TN = Cfg::TN;
  constexpr int ST = Cfg::STAGES;
  typedef T Stage[BK * LDS];
  Stage *As, *Bs;
  if constexpr (BK == 8) {
    __shared__ __align__(16) T sA[ST][BK * LDS];
    __shared__ __align__(16) T sB[ST][BK * LDS];
    As = sA;
    Bs = sB;
  } else {
    extern __shared__ __align__(16) unsigned char f_dyn_smem[];
    As = reinterpret_cast<Stage *>(f_dyn_smem);
    Bs = As + ST;
  }

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

  auto issue = [&](int kt) {
    if (kt < nk) {
      const int s = kt % ST, k0 = kb + kt * BK;
      f_fill<T, BM, BK, LDS>(As[s], p.A, p.lda, p.opA, p.triA, p.unitA, false, m0, p.M, rowsA, k0, ke, p.vecA, tid);
      f_fill<T, BN, BK, LDS>(Bs[s], p.B, p.ldb, p.opB, p.triB, p.unitB, true, g0, p.N, rowsB, k0, ke, p.vecB, tid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int kt = 0; kt < ST - 1; ++kt) issue(kt);

  for (int kt = 0; kt < nk; ++kt) {
    asm volatile("cp.async.wait_group %0;" ::"n"(ST - 2) : "memory");
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
  // CHOL_FFMA_BK=16 selects the deeper k-tile (written after the round-1 GPU budget was spent: off
  // until it has been verified and measured on hardware)
  static const int bk = [] { const char *e = getenv("CHOL_FFMA_BK"); return (e && atoi(e) == 16) ? 16 : 8; }();
  if (bk == 16) {
    constexpr int bytes = Cfg::smem_bytes(16);
    static std::atomic<int> attr_set[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63].load(std::memory_order_acquire)) {
      const cudaError_t e = cudaFuncSetAttribute(gemm_ffma_kernel<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (e != cudaSuccess) return (int)e;
      attr_set[dev & 63].store(1, std::memory_order_release);
    }
    gemm_ffma_kernel<T, 16><<<(unsigned)tiles, 256, bytes, (cudaStream_t)d->stream>>>(a);
  } else {
    gemm_ffma_kernel<T, 8><<<(unsigned)tiles, 256, 0, (cudaStream_t)d->stream>>>(a);
  }
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}

}  // namespace chol

int chol_gemm_fp32_launch(const chol_gemm_desc *d) {
  if (d->prec == 's') return chol::launch_f<float>(d);
  if (d->prec == 'c') return chol::launch_f<float2>(d);
  return (int)cudaErrorInvalidValue;
}
