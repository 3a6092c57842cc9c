// gemm_dmma.cu — the FP64 tensor-core GEMM engine (d and z) for sm_100a.
//
// One kernel family serves every dense contraction on the Cholesky / SPD-inverse path:
//   GEMM  (reference blas/dgemm.c:148 cuDgemm2, zgemm.c:226)      D = alpha op(A) op(B) + beta C
//   SYRK / HERK (blas/dsyrk.c:163, zherk.c:190)                   cuplo: only one triangle of C/D
//   TRMM2 (blas/dtrmm.c:374)                                      triA/triB: triangular operand masks
//   inverse-based panel solves of POTRF/TRTRI (lapack/spotrf.c:253,:382 idea)
//
// Design (B200-first, not a port of blas/dgemm.cu's 64x8 SIMT tiles):
//   * math: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4, the native FP64 tensor-core op on sm_100a);
//     z uses 4 real DMMAs per complex MAC on split (re, im) accumulators.
//   * CTA tile 128x64 (d) / 64x64 (z), 4 consumer warps (2x2), warp tile 64x32 (d) / 32x32 (z),
//     accumulators in registers (64 doubles/thread); 2 CTAs per SM so that one CTA's C-tile
//     epilogue (HBM latency bound) overlaps the other's DMMA main loop.
//   * operands staged in shared memory as [k][m] / [k][n] rows with a +4 (d) / +2 (z) element pad:
//     conflict-free for the 128-bit fragment loads of the DMMA lane layout.
//   * 4-stage mbarrier ring, warp-specialised: warps 4-5 are producers (MODE 0 needs only warp 4).
//       MODE 0 (hot path: op(A)=N, op(B)=T/C, aligned): every k-row of the A/B tile is one TMA
//              bulk copy (cp.async.bulk, SASS UBLKCP) issued by one lane, completion counted as
//              transaction bytes on the stage's mbarrier;
//       MODE 1 (any transpose / conjugate / triangular mask / odd alignment): the producer warp
//              gathers elements through registers into the same shared layout.
//   * triangular tile scheduler: for cuplo the grid holds only tiles that touch the triangle
//     (the reference launches the full square and lets half the blocks return, blas/dsyrk.cu:68-75);
//     tiles are enumerated in 8-row-block super-rows so that concurrently resident CTAs share the
//     same B rows and an L2-resident set of A rows.
//   * triangular operands skip the k-range that is structurally zero.
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "chol_kernels.h"
#include "launch_count.h"
#include "gemm_common.cuh"
#include "ptx_sm100.cuh"

namespace chol {

template <bool CPLX>
struct GemmCfg {
  static constexpr int BM = CPLX ? 64 : 128;
  static constexpr int BN = 64;
  static constexpr int BK = CPLX ? 8 : 16;
  static constexpr int LDA = CPLX ? 66 : 132;  // shared row pitch, elements
  static constexpr int LDB = CPLX ? 66 : 68;
  static constexpr int STAGES = 4;
  static constexpr int A_STAGE = BK * LDA;
  static constexpr int B_STAGE = BK * LDB;
  static constexpr int ESZ = CPLX ? 16 : 8;
  static constexpr int R = BM / BN;  // tile aspect, used by the triangular scheduler
  static constexpr int GROUP = 8;    // row-blocks per super-row
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) * ESZ + 2 * STAGES * 8;
  static constexpr int THREADS = 192;  // 4 consumer warps + 2 producer warps (register file is allocated in warp pairs)
};

template <bool CPLX>
struct GemmArgs {
  using T = typename std::conditional<CPLX, double2, double>::type;
  const T *A, *B, *C;
  T *D;
  long long lda, ldb, ldc, ldd;
  int M, N, K;
  T alpha, beta;
  int opA, opB;      // 0 N, 1 T, 2 C
  int triA, triB;    // 0 none, 1 lower, 2 upper (of the STORED operand)
  int unitA, unitB;
  int cuplo;         // 0 all, 1 lower, 2 upper
  int herm;
  int NI, NJ;        // tile counts
  int vec_ok;        // C/D allow 16-byte accesses
  int beta_zero;
  int conjB_regs;    // MODE 0, z: conjugate B fragments in registers (raw TMA copy cannot)
  ColMap cm;         // block-cyclic column map of D's columns (cnb = 0: identity)
};

template <bool CPLX, int MODE>
__global__ void __launch_bounds__(GemmCfg<CPLX>::THREADS, 2) gemm_dmma_kernel(const GemmArgs<CPLX> p) {
  using Cfg = GemmCfg<CPLX>;
  using T = typename GemmArgs<CPLX>::T;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, LDA = Cfg::LDA, LDB = Cfg::LDB, ST = Cfg::STAGES;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  T *As = reinterpret_cast<T *>(smem_raw);
  T *Bs = As + ST * Cfg::A_STAGE;
  uint64_t *full = reinterpret_cast<uint64_t *>(Bs + ST * Cfg::B_STAGE);
  uint64_t *empty = full + ST;

  int bi, bj;
  if (!tile_coords<BM, BN, Cfg::GROUP>(p.cuplo, p.NI, p.NJ, p.cm, (long long)blockIdx.x, bi, bj)) return;
  const int m0 = bi * BM, n0 = bj * BN;
  const int g0 = (int)gcol(p.cm, n0);  // triangle coordinate / B row of the tile's first column

  // structurally-zero k ranges of triangular operands
  int kb = 0, ke = p.K;
  if (p.triA) {
    const bool lower = (p.triA == 1) == (p.opA == 0);  // op(A)[m][k] != 0 only for k <= m
    if (lower) ke = min(ke, m0 + BM); else kb = max(kb, m0);
  }
  if (p.triB) {
    const bool kgen = (p.triB == 1) == (p.opB == 0);  // op(B)[k][n] != 0 only for k >= n
    if (kgen) kb = max(kb, g0); else ke = min(ke, g0 + BN);
  }
  kb = (kb / BK) * BK;
  const int nk = ke > kb ? (ke - kb + BK - 1) / BK : 0;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full[s], MODE == 0 ? 1 : 2);
      mbar_init(&empty[s], 4);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp >= 4) {
    // ======================= producer warps =======================
    if (MODE == 0 && warp == 5) return;
    const int pw = warp - 4;
    const int rowsA = min(BM, p.M - m0), rowsB = min(BN, p.N - n0);
    for (int kt = 0; kt < nk; ++kt) {
      const int s = kt % ST;
      mbar_wait(&empty[s], ((kt / ST) & 1) ^ 1);
      T *as = As + s * Cfg::A_STAGE, *bs = Bs + s * Cfg::B_STAGE;
      const int k0 = kb + kt * BK;
      if constexpr (MODE == 0) {
        if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)(BK * (rowsA + rowsB) * Cfg::ESZ));
        __syncwarp();
        if (lane < BK)
          tma_bulk_g2s(as + lane * LDA, p.A + (long long)(k0 + lane) * p.lda + m0, rowsA * Cfg::ESZ, &full[s]);
        else if (lane < 2 * BK)
          tma_bulk_g2s(bs + (lane - BK) * LDB, p.B + (long long)(k0 + lane - BK) * p.ldb + g0, rowsB * Cfg::ESZ, &full[s]);
      } else {
        constexpr int U = 8;  // loads in flight per lane
        // A tile: logical op(A)[m0+mm][k0+kk] -> as[kk*LDA + mm]
        for (int base = 0; base < BM * BK; base += 64 * U) {
          T v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int idx = base + u * 64 + pw * 32 + lane;
            int mm, kk;
            if (p.opA == 0) { mm = idx % BM; kk = idx / BM; } else { kk = idx % BK; mm = idx / BK; }
            const long long m = m0 + mm, k = k0 + kk;
            const bool in = (m < p.M) && (k < ke);
            if (p.opA == 0) v[u] = fetch_stored<T>(p.A, p.lda, in, m, k, p.triA, p.unitA, false);
            else            v[u] = fetch_stored<T>(p.A, p.lda, in, k, m, p.triA, p.unitA, p.opA == 2);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int idx = base + u * 64 + pw * 32 + lane;
            int mm, kk;
            if (p.opA == 0) { mm = idx % BM; kk = idx / BM; } else { kk = idx % BK; mm = idx / BK; }
            as[kk * LDA + mm] = v[u];
          }
        }
        // B tile: logical op(B)[k0+kk][n0+nn] -> bs[kk*LDB + nn]
        for (int base = 0; base < BN * BK; base += 64 * U) {
          T v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int idx = base + u * 64 + pw * 32 + lane;
            int nn, kk;
            if (p.opB != 0) { nn = idx % BN; kk = idx / BN; } else { kk = idx % BK; nn = idx / BK; }
            const long long n = g0 + nn, k = k0 + kk;
            const bool in = (n0 + nn < p.N) && (k < ke);
            if (p.opB == 0) v[u] = fetch_stored<T>(p.B, p.ldb, in, k, n, p.triB, p.unitB, false);
            else            v[u] = fetch_stored<T>(p.B, p.ldb, in, n, k, p.triB, p.unitB, p.opB == 2);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int idx = base + u * 64 + pw * 32 + lane;
            int nn, kk;
            if (p.opB != 0) { nn = idx % BN; kk = idx / BN; } else { kk = idx % BK; nn = idx / BK; }
            bs[kk * LDB + nn] = v[u];
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
    }
    return;
  }

  // ======================= consumer warps =======================
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;

  if constexpr (!CPLX) {
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < nk; ++kt) {
      const int s = kt % ST;
      mbar_wait(&full[s], (kt / ST) & 1);
      const double *as = As + s * Cfg::A_STAGE + wm * 64 + 2 * g;
      const double *bs = Bs + s * Cfg::B_STAGE + wn * 32 + 2 * g;
#pragma unroll
      for (int ks = 0; ks < BK / 4; ++ks) {
        const int k = ks * 4 + t;
        double2 a[4], b[2];
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = *reinterpret_cast<const double2 *>(as + k * LDA + 16 * q);
#pragma unroll
        for (int q = 0; q < 2; ++q) b[q] = *reinterpret_cast<const double2 *>(bs + k * LDB + 16 * q);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dmma_884(acc[i][j][0], acc[i][j][1], (i & 1) ? a[i >> 1].y : a[i >> 1].x, (j & 1) ? b[j >> 1].y : b[j >> 1].x);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }

    // ---- epilogue: D = alpha*acc + beta*C.  Thread owns rows m, m+1 (m = ..+16q+2g) of columns
    //      n = n0 + wn*32 + 16*(j>>1) + 2*(2t+c) + (j&1)
    const double alpha = p.alpha, beta = p.beta;
    const bool interior = (m0 + BM <= p.M) && (n0 + BN <= p.N) &&
                          (p.cuplo == 0 || (p.cuplo == 1 && m0 >= g0 + BN - 1) || (p.cuplo == 2 && m0 + BM - 1 <= g0));
    if (interior && p.vec_ok) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long long m = m0 + wm * 64 + 16 * q + 2 * g;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {  // 4 columns (16-byte row pairs) per batch
          double2 cv[4];
          if (!p.beta_zero) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = jh * 2 + (e >> 1), c = e & 1;
              const long long n = n0 + wn * 32 + 16 * (j >> 1) + 2 * (2 * t + c) + (j & 1);
              cv[e] = *reinterpret_cast<const double2 *>(p.C + n * p.ldc + m);
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = jh * 2 + (e >> 1), c = e & 1;
            const long long n = n0 + wn * 32 + 16 * (j >> 1) + 2 * (2 * t + c) + (j & 1);
            double2 o;
            o.x = alpha * acc[2 * q][j][c];
            o.y = alpha * acc[2 * q + 1][j][c];
            if (!p.beta_zero) {
              o.x += beta * cv[e].x;
              o.y += beta * cv[e].y;
            }
            *reinterpret_cast<double2 *>(p.D + n * p.ldd + m) = o;
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const long long m = m0 + wm * 64 + 16 * (i >> 1) + 2 * g + (i & 1);
            const long long n = n0 + wn * 32 + 16 * (j >> 1) + 2 * (2 * t + c) + (j & 1);
            if (m >= p.M || n >= p.N) continue;
            if (p.cuplo == 1 && m < n - n0 + g0) continue;
            if (p.cuplo == 2 && m > n - n0 + g0) continue;
            double o = alpha * acc[i][j][c];
            if (!p.beta_zero) o += beta * p.C[n * p.ldc + m];
            p.D[n * p.ldd + m] = o;
          }
    }
  } else {
    // ---------------- complex: split accumulators, 4 DMMAs per complex MAC ----------------
    double accr[4][4][2], acci[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) accr[i][j][0] = accr[i][j][1] = acci[i][j][0] = acci[i][j][1] = 0.0;

    const bool conjb = (MODE == 0) && p.conjB_regs;
    for (int kt = 0; kt < nk; ++kt) {
      const int s = kt % ST;
      mbar_wait(&full[s], (kt / ST) & 1);
      const double2 *as = As + s * Cfg::A_STAGE + wm * 32 + g;
      const double2 *bs = Bs + s * Cfg::B_STAGE + wn * 32 + g;
#pragma unroll
      for (int ks = 0; ks < BK / 4; ++ks) {
        const int k = ks * 4 + t;
        double2 a[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = as[k * LDA + 8 * q];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double2 b = bs[k * LDB + 8 * j];
          if (conjb) b.y = -b.y;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            dmma_884(accr[i][j][0], accr[i][j][1], a[i].x, b.x);
            dmma_884(acci[i][j][0], acci[i][j][1], a[i].x, b.y);
            dmma_884(acci[i][j][0], acci[i][j][1], a[i].y, b.x);
            dmma_884(accr[i][j][0], accr[i][j][1], -a[i].y, b.y);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }

    const double2 alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const long long m = m0 + wm * 32 + 8 * i + g;
          const long long n = n0 + wn * 32 + 8 * j + 2 * t + c;
          if (m >= p.M || n >= p.N) continue;
          if (p.cuplo == 1 && m < n - n0 + g0) continue;
          if (p.cuplo == 2 && m > n - n0 + g0) continue;
          const double xr = accr[i][j][c], xi = acci[i][j][c];
          double2 o = make_double2(alpha.x * xr - alpha.y * xi, alpha.x * xi + alpha.y * xr);
          if (!p.beta_zero) {
            const double2 cv = p.C[n * p.ldc + m];
            o.x += beta.x * cv.x - beta.y * cv.y;
            o.y += beta.x * cv.y + beta.y * cv.x;
          }
          if (p.herm && m == n - n0 + g0) o.y = 0.0;
          p.D[n * p.ldd + m] = o;
        }
  }
}

template <bool CPLX>
static bool tma_eligible(const chol_gemm_desc *d) {
  using Cfg = GemmCfg<CPLX>;
  if (d->force_generic) return false;
  if (d->opA != 'N' || d->opB == 'N') return false;
  if (!CPLX && d->opB != 'T' && d->opB != 'C') return false;
  if (CPLX && d->opB != 'T' && d->opB != 'C') return false;
  if (d->triA || d->triB) return false;
  if (d->k == 0 || d->k % Cfg::BK) return false;
  if (((uintptr_t)d->A | (uintptr_t)d->B) & 15) return false;
  if (!CPLX) {  // 8-byte elements: rows must start and end on 16-byte boundaries
    if ((d->lda | d->ldb | d->m | d->n) & 1) return false;
  }
  return true;
}

template <bool CPLX>
static int launch_t(const chol_gemm_desc *d) {
  using Cfg = GemmCfg<CPLX>;
  using T = typename GemmArgs<CPLX>::T;
  if (d->m == 0 || d->n == 0) return 0;
  GemmArgs<CPLX> a;
  a.A = (const T *)d->A; a.B = (const T *)d->B; a.C = (const T *)d->C; a.D = (T *)d->D;
  a.lda = (long long)d->lda; a.ldb = (long long)d->ldb; a.ldc = (long long)d->ldc; a.ldd = (long long)d->ldd;
  a.M = (int)d->m; a.N = (int)d->n; a.K = (int)d->k;
  const bool alpha_zero = d->alpha[0] == 0.0 && (!CPLX || d->alpha[1] == 0.0);
  if (alpha_zero) a.K = 0;  // D = beta*C only; A and B are never read (BLAS semantics)
  if constexpr (CPLX) {
    a.alpha = make_double2(d->alpha[0], d->alpha[1]);
    a.beta = make_double2(d->beta[0], d->beta[1]);
    a.beta_zero = (d->beta[0] == 0.0 && d->beta[1] == 0.0);
  } else {
    a.alpha = d->alpha[0];
    a.beta = d->beta[0];
    a.beta_zero = (d->beta[0] == 0.0);
  }
  a.opA = op_code(d->opA); a.opB = op_code(d->opB);
  if (!CPLX) { if (a.opA == 2) a.opA = 1; if (a.opB == 2) a.opB = 1; }
  a.triA = tri_code(d->triA); a.triB = tri_code(d->triB);
  a.unitA = d->unitA; a.unitB = d->unitB;
  a.cuplo = tri_code(d->cuplo);
  a.herm = CPLX ? d->herm : 0;
  a.NI = (a.M + Cfg::BM - 1) / Cfg::BM;
  a.NJ = (a.N + Cfg::BN - 1) / Cfg::BN;
  a.vec_ok = CPLX ? 1 : ((((uintptr_t)d->D & 15) == 0) && (d->ldd % 2 == 0) &&
                         (a.beta_zero || ((((uintptr_t)d->C & 15) == 0) && (d->ldc % 2 == 0))));
  a.conjB_regs = CPLX && d->opB == 'C';
  a.cm.cnb = d->cyc_nb; a.cm.cP = d->cyc_p; a.cm.coff = d->cyc_off;
  if (a.cm.cnb && (a.cm.cnb % Cfg::BN || a.cuplo == 2 || d->triA || d->triB || a.cm.cP < 1)) return (int)cudaErrorInvalidValue;
  const bool tma = !alpha_zero && tma_eligible<CPLX>(d);
  const long long tiles = total_tiles<Cfg::BM, Cfg::BN, Cfg::GROUP>(a.cuplo, a.NI, a.NJ, a.cm);
  if (tiles <= 0) return 0;
  if (tiles > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)d->stream;
  static bool attr_set[64][2][2];  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = tma ? gemm_dmma_kernel<CPLX, 0> : gemm_dmma_kernel<CPLX, 1>;
  if (!attr_set[dev & 63][CPLX][tma]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev & 63][CPLX][tma] = true;
  }
  kern<<<(unsigned)tiles, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(a);
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}

}  // namespace chol

int chol_gemm_fp32_launch(const chol_gemm_desc *d);  // gemm_ffma.cu

extern "C" int chol_gemm_launch(const chol_gemm_desc *d) {
  switch (d->prec) {
    case 'd': return chol::launch_t<false>(d);
    case 'z': return chol::launch_t<true>(d);
    case 's':
    case 'c': return chol_gemm_fp32_launch(d);
    default: return (int)cudaErrorInvalidValue;
  }
}

// test hook (host only, no GPU needed): enumerate the tile schedule of the d-engine (128x64 tiles)
// exactly as the kernel's CTAs decode it.  Writes (bi, bj) pairs of the VALID tiles, returns the
// grid size; *nvalid receives the number of valid tiles.
extern "C" long long chol_test_tile_schedule(int cuplo, int m, int n, int cyc_nb, int cyc_p, int cyc_off, int *out,
                                             long long cap, long long *nvalid) {
  using Cfg = chol::GemmCfg<false>;
  chol::ColMap cm{cyc_nb, cyc_p, cyc_off};
  const int cu = chol::tri_code(cuplo);
  const int NI = (m + Cfg::BM - 1) / Cfg::BM, NJ = (n + Cfg::BN - 1) / Cfg::BN;
  const long long tiles = chol::total_tiles<Cfg::BM, Cfg::BN, Cfg::GROUP>(cu, NI, NJ, cm);
  long long v = 0;
  for (long long t = 0; t < tiles; ++t) {
    int bi, bj;
    if (chol::tile_coords<Cfg::BM, Cfg::BN, Cfg::GROUP>(cu, NI, NJ, cm, t, bi, bj)) {
      if (v < cap) { out[2 * v] = bi; out[2 * v + 1] = bj; }
      ++v;
    }
  }
  *nvalid = v;
  return tiles;
}

extern "C" int chol_gemm_uses_tma(const chol_gemm_desc *d) {
  const bool alpha_zero = d->alpha[0] == 0.0 && ((d->prec != 'z' && d->prec != 'c') || d->alpha[1] == 0.0);
  if (alpha_zero) return 0;
  if (d->prec == 'd') return chol::tma_eligible<false>(d) ? 1 : 0;
  if (d->prec == 'z') return chol::tma_eligible<true>(d) ? 1 : 0;
  return 0;
}
