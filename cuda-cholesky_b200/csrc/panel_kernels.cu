// panel_kernels.cu — the latency-bound pieces of the path, hand-written for sm_100a:
//   potf2  : Cholesky of a diagonal block        (reference spec: lapack/dpotrf.c:34-81, kernel lapack/dpotrf.cu:23-148)
//   trti2  : inverse of a triangular block       (lapack/dtrtri.c:43-96, lapack/dtrtri.cu:23)
//   lauu2  : U*U^H / L^H*L of a triangular block (lapack/dlauum.c:43-69, lapack/dlauum.cu:23)
//   copy / transpose / fill helpers, the single-pass logdet reduction (lapack/dlogdet.cu:53-118)
//   and the FP64 peak probes used by bench.py.
//
// The reference runs these on the CPU (its device versions are dead code: 1 block x 32 threads,
// SURVEY.md §2.3).  Here the whole block is staged once in shared memory (up to 128x128 doubles,
// 129 KB of the 227 KB a B200 CTA can own), factored by one 512-thread CTA with one barrier per
// column, and results stream straight back to HBM - nothing crosses PCIe, nothing runs on the host.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "chol_kernels.h"
#include "emu_compat.cuh"
#include "launch_count.h"
#ifndef CHOL_CPU_EMU
#include "ptx_sm100.cuh"
#endif
#include "scalar_ops.cuh"

namespace chol {

constexpr int PANEL_THREADS = 512;

template <typename T>
__device__ __forceinline__ void load_lower_form(T *S, int LD, int upper, bool conj_upper, int n, const T *A, long long lda) {
  // S[i + k*LD] (i >= k) = A(i,k) for lower storage, op(A(k,i)) for upper storage
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int r = idx % n, c = idx / n;  // coalesced along r
    if (!upper) {
      if (r >= c) S[r + c * LD] = A[(long long)c * lda + r];
    } else if (r <= c) {
      T v = A[(long long)c * lda + r];
      S[c + r * LD] = conj_upper ? conj_of(v) : v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// potf2: the block lives in REGISTERS, distributed 2-D cyclically over the CTA's 16 warps x 32 lanes
// (thread (lane, w) owns rows lane+32a, columns w+16b): the active window stays load-balanced as it
// shrinks.  Right-looking, one __syncthreads per column: the warp that owns column c takes the
// pivot by shuffle, scales its column, publishes it through a double-buffered shared-memory vector
// and streams it to HBM; every thread then applies the rank-1 update to its own registers.
template <typename T, int NA, int NB>
__global__ void __launch_bounds__(PANEL_THREADS) potf2_kernel(int upper, int n, T *A, long long lda, long *dinfo, long info_base) {
  using Real = typename ScalarOf<T>::Real;
  __shared__ T col[2][32 * NA];
  __shared__ int sfail;
  if (*dinfo != 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) sfail = 0;
  T reg[NA][NB];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int i = lane + 32 * a, k = warp + 16 * b;
      T v = zero_of<T>();
      if (i < n && k < n && i >= k) v = upper ? conj_of(A[(long long)i * lda + k]) : A[(long long)k * lda + i];
      reg[a][b] = v;
    }
  __syncthreads();
#pragma unroll
  for (int bo = 0; bo < NB; ++bo) {
    constexpr int dummy = 0; (void)dummy;
    const int ao = bo >> 1;  // rows 16bo..16bo+15 live in register slice a = (16bo + cw) / 32 = bo / 2
    for (int cw = 0; cw < 16; ++cw) {
      const int c = 16 * bo + cw;
      if (c >= n) break;
      T *cb = col[c & 1];
      if (warp == cw) {  // this warp owns column c (register slice b = bo)
        const int lo = c & 31;
        Real d;
        d = __shfl_sync(0xffffffffu, re_of(reg[ao][bo]), lo);
        if (!(d > Real(0))) {  // <= 0 or NaN: not positive definite (lapack/dpotrf.c:70-73)
          if (lane == 0) {
            sfail = c + 1;
            *dinfo = info_base + c + 1;
            A[(long long)c * lda + c] = from_real<T>(d);
          }
        } else {
          const Real r = sqrt(d), rinv = Real(1) / r;
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            const int i = lane + 32 * a;
            if (i >= c && i < n) {
              const T x = (i == c) ? from_real<T>(r) : rscale(reg[a][bo], rinv);
              cb[i] = x;
              if (!upper) A[(long long)c * lda + i] = x; else A[(long long)i * lda + c] = conj_of(x);
            }
          }
        }
      }
      __syncthreads();
      if (sfail == c + 1) return;  // (see the note on sfail in potf2_inv_kernel)
      // rank-1 update of the trailing window: columns k > c, rows i >= k
      T li[NA];
#pragma unroll
      for (int a = 0; a < NA; ++a) li[a] = cb[lane + 32 * a];
#pragma unroll
      for (int b = bo; b < NB; ++b) {
        const int k = warp + 16 * b;
        if (k > c && k < n) {
          const T f = neg(conj_of(cb[k]));
#pragma unroll
          for (int a = 0; a < NA; ++a)
            if (lane + 32 * a >= k) fma_acc(reg[a][b], li[a], f);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// trti2: X = L^-1 by forward substitution, the triangle staged in shared memory.  Every warp carries
// EIGHT columns at once (columns warp+16c): the residuals live in registers (lane = row mod 32), the
// current unknown is broadcast by shuffle, and the eight independent chains hide each other's latency
// while sharing the loads of L's column k.
// V2 (same switch as the fused kernel's V2, CHOL_POTF2_V2): chains that cannot have started in the current
// 32-row slice are pruned at compile time, as in potf2_inv_kernel's phase 2.
template <typename T, int NQ, int NC, bool V2 = false>
__global__ void __launch_bounds__(PANEL_THREADS) trti2_kernel(int upper, int unit, int n, const T *A, long long lda, T *B,
                                                              long long ldb, int zero_other, long *dinfo, long info_base) {
  CHOL_DYN_SMEM(smem_raw);
  T *L = reinterpret_cast<T *>(smem_raw);
  const int LD = n | 1;
  T *dinv = L + (size_t)n * LD;
  __shared__ int fail_lo, fail_hi;
  // info protocol (bits 1-2 of the flags word, csrc/chol_kernels.h):
  //   0  LAPACK drivers: *dinfo is a sticky 64-bit word - skip if already set, first failure wins
  //   2  the reference's own device kernel contract (lapack/dtrtri.cu:23-29): a 32-bit int that the
  //      kernel itself sets to 0 on entry, no early-out
  //   4  BLAS use (TRSM diagonal blocks): no info word at all; a zero pivot gives Inf like the BLAS
  const int info_mode = zero_other & 6;
  zero_other &= 1;
  if (info_mode == 0 && *dinfo != 0) return;
  if (threadIdx.x == 0) {
    fail_lo = 0x7fffffff; fail_hi = -1;
    if (info_mode == 2) *reinterpret_cast<int *>(dinfo) = 0;
  }
  load_lower_form(L, LD, upper, false, n, A, lda);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (unit) dinv[i] = one_of<T>();
    else {
      const T d = L[i + i * LD];
      if (is_zero(d) && info_mode != 4) { atomicMin(&fail_lo, i); atomicMax(&fail_hi, i); dinv[i] = zero_of<T>(); }
      else dinv[i] = recip(d);
    }
  }
  __syncthreads();
  if (fail_hi >= 0) {
    // the reference sweeps columns upwards for Upper and downwards for Lower and reports the first
    // zero pivot it meets (lapack/dtrtri.c:49-52, :73-77)
    if (threadIdx.x == 0) {
      const long v = info_base + (upper ? fail_lo : fail_hi) + 1;
      if (info_mode == 2) *reinterpret_cast<int *>(dinfo) = (int)v; else *dinfo = v;
    }
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // NC columns per warp (16 warps x NC = nmax), NQ = nmax / 32 row slices
  T r[NC][NQ];           // residual; entry (lane, q) turns into x_row once row 32q+lane is solved
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int q = 0; q < NQ; ++q) r[c][q] = (lane + 32 * q == warp + 16 * c) ? one_of<T>() : zero_of<T>();
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    if (32 * q >= n) continue;
    for (int kk = 0; kk < 32; ++kk) {
      const int k = 32 * q + kk;
      if (k >= n) break;
      if (k < warp) continue;  // none of this warp's columns has started yet
      const T dk = dinv[k];
      T lk[NQ];
#pragma unroll
      for (int q2 = 0; q2 < NQ; ++q2) {
        const int i = lane + 32 * q2;
        lk[q2] = (q2 >= q && i > k && i < n) ? L[i + k * LD] : zero_of<T>();
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {  // branch-free: the NC chains interleave
        if (V2 && 16 * c > 32 * q + 31) continue;  // compile-time: column warp+16c lies below this whole row slice
        const bool active = (k >= warp + 16 * c);
        T rk;
        if constexpr (ScalarOf<T>::cplx) { rk.x = __shfl_sync(0xffffffffu, r[c][q].x, kk); rk.y = __shfl_sync(0xffffffffu, r[c][q].y, kk); }
        else rk = __shfl_sync(0xffffffffu, r[c][q], kk);
        const T xk = active ? mul(rk, dk) : zero_of<T>();
        if (lane == kk && active) r[c][q] = xk;
        const T nxk = neg(xk);
#pragma unroll
        for (int q2 = 0; q2 < NQ; ++q2)
          if (q2 >= q) fma_acc(r[c][q2], lk[q2], nxk);  // rows <= k have lk = 0
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int j = warp + 16 * c;
    if (j >= n) continue;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int i = lane + 32 * q;
      if (i < n) {
        if (i >= j) { if (!upper) B[(long long)j * ldb + i] = r[c][q]; else B[(long long)i * ldb + j] = r[c][q]; }
        else if (zero_other) { if (!upper) B[(long long)j * ldb + i] = zero_of<T>(); else B[(long long)i * ldb + j] = zero_of<T>(); }
      }
    }
  }
}

// L(:, 0:ncols) from its shared-memory copy (Ls[i + c*LD], zeros above the diagonal) to the caller's
// triangle, coalesced for both storages: lower A(i,c) = L(i,c) - a warp per column, lanes along rows;
// upper A(c,i) = conj(L(i,c)) - a warp per ROW i, lanes along the columns c <= i (LD is odd: the
// shared reads are conflict-free either way).
template <typename T, int NA, int LD>
__device__ __forceinline__ void store_factor_columns(int upper, int n, int ncols, const T *Ls, T *A, long long lda, int warp,
                                                     int lane) {
  if (!upper) {
    for (int c = warp; c < ncols; c += PANEL_THREADS / 32)
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        const int i = lane + 32 * a;
        if (i >= c && i < n) A[(long long)c * lda + i] = Ls[i + c * LD];
      }
  } else {
    for (int i = warp; i < n; i += PANEL_THREADS / 32)
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        const int c = lane + 32 * a;
        if (c <= i && c < ncols) A[(long long)i * lda + c] = conj_of(Ls[i + c * LD]);
      }
  }
}

// ------------------------------------------------------------------------------------------------
// potf2 + inverse, fused (the POTRF panel chain needs both the factor L of a diagonal block and
// L^-1, lapack/spotrf.c:253): phase 1 is the register-resident right-looking factorisation above,
// except that every finished column is published into a persistent shared-memory copy of L (and
// 1/L_cc is kept); phase 2 inverts that copy by forward substitution, eight columns per warp as
// independent, branch-free chains.  One launch, one pass over HBM, no division after the rsqrt.
//
// V2 (CHOL_POTF2_V2=1; written after the round-1 GPU budget was spent, off until verified and measured):
// the SASS of the owner warp's per-column section - the serial critical path, executed 128 times -
// is dominated by the global stores of the finished column and their address arithmetic, and by
// per-row branches.  V2 keeps only pivot -> rsqrt -> scale -> shared store before the barrier
// (branch-free; rows beyond n hold zeros), writes L to HBM once after phase 1 from the shared copy
// (coalesced for both triangles, overlapping phase 2), and prunes at compile time the phase-2 chains
// that cannot have started in the current 32-row slice (column warp+16c >= 16c > 32q+31).
template <typename T, int NA, int NB, int NQ, int NC, bool V2>
__global__ void __launch_bounds__(PANEL_THREADS) potf2_inv_kernel(int upper, int n, T *A, long long lda, T *X, long long ldx,
                                                                  long *dinfo, long info_base) {
  using Real = typename ScalarOf<T>::Real;
  CHOL_DYN_SMEM(smem_raw);
  constexpr int NMAX = 32 * NA, LD = NMAX + 1;
  T *Ls = reinterpret_cast<T *>(smem_raw);          // L(i,c) at Ls[i + c*LD]
  Real *dinv = reinterpret_cast<Real *>(Ls + (size_t)NMAX * LD);
  __shared__ int sfail;
  if (*dinfo != 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) sfail = 0;
  T reg[NA][NB];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int i = lane + 32 * a, k = warp + 16 * b;
      T v = zero_of<T>();
      if (i < n && k < n && i >= k) v = upper ? conj_of(A[(long long)i * lda + k]) : A[(long long)k * lda + i];
      reg[a][b] = v;
    }
  __syncthreads();
  // ---------------- phase 1: factor ----------------
#pragma unroll
  for (int bo = 0; bo < NB; ++bo) {
    const int ao = bo >> 1;
    for (int cw = 0; cw < 16; ++cw) {
      const int c = 16 * bo + cw;
      if (c >= n) break;
      T *cb = Ls + c * LD;
      if (warp == cw) {
        const Real d = __shfl_sync(0xffffffffu, re_of(reg[ao][bo]), c & 31);
        if (!(d > Real(0))) {  // <= 0 or NaN: not positive definite (lapack/dpotrf.c:70-73)
          if (lane == 0) {
            sfail = c + 1;
            *dinfo = info_base + c + 1;
            A[(long long)c * lda + c] = from_real<T>(d);
          }
        } else {
          const Real rinv = rsqrt(d), r = d * rinv;
          if (V2 || lane == 0) dinv[c] = rinv;  // V2: every lane stores the same value - no divergent block (ptxas duplicated the rsqrt for it)
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            const int i = lane + 32 * a;
            if constexpr (V2) {  // branch-free; registers of rows >= n are zero, the shared column has NMAX rows
              const T x = (i > c) ? rscale(reg[a][bo], rinv) : ((i == c) ? from_real<T>(r) : zero_of<T>());
              cb[i] = x;
            } else if (i < n) {
              const T x = (i > c) ? rscale(reg[a][bo], rinv) : ((i == c) ? from_real<T>(r) : zero_of<T>());
              cb[i] = x;  // rows above the diagonal are stored as zeros: phase 2 reads a full column
              if (i >= c) { if (!upper) A[(long long)c * lda + i] = x; else A[(long long)i * lda + c] = conj_of(x); }
            }
          }
        }
      }
      __syncthreads();
      // sfail holds the failing column + 1.  The owner of column c+1 may already have run ahead through its
      // update, found ITS pivot bad and written sfail while a slow warp is only now testing the flag for
      // column c: a plain "nonzero" test would let that warp leave one barrier early and hang the block
      // (found by the CPU emulation, tests/emu).  Comparing with c + 1 makes every thread leave after the
      // same barrier; nobody can be more than one column ahead.
      if (sfail == c + 1) {
        if constexpr (V2) store_factor_columns<T, NA, LD>(upper, n, c, Ls, A, lda, warp, lane);  // columns finished so far
        return;
      }
      T li[NA], f[NB];
#pragma unroll
      for (int a = 0; a < NA; ++a) li[a] = cb[lane + 32 * a];
#pragma unroll
      for (int b = bo; b < NB; ++b) {
        const int k = warp + 16 * b;
        f[b] = (k > c && k < n) ? neg(conj_of(cb[k])) : zero_of<T>();
      }
#pragma unroll
      for (int b = bo; b < NB; ++b)
#pragma unroll
        for (int a = 0; a < NA; ++a)
          if (32 * a + 31 >= 16 * b) fma_acc(reg[a][b], li[a], (lane + 32 * a >= warp + 16 * b) ? f[b] : zero_of<T>());
    }
  }
  __syncthreads();
  if constexpr (V2) store_factor_columns<T, NA, LD>(upper, n, n, Ls, A, lda, warp, lane);  // fire and forget: overlaps phase 2
  // ---------------- phase 2: X = L^-1 ----------------
  T r[NC][NQ];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int q = 0; q < NQ; ++q) r[c][q] = (lane + 32 * q == warp + 16 * c) ? one_of<T>() : zero_of<T>();
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    if (32 * q >= n) continue;
    for (int kk = 0; kk < 32; ++kk) {
      const int k = 32 * q + kk;
      if (k >= n) break;
      if (k < warp) continue;  // none of this warp's columns has started yet (warp-uniform)
      const Real dk = dinv[k];
      T lk[NQ];
#pragma unroll
      for (int q2 = 0; q2 < NQ; ++q2) {
        const int i = lane + 32 * q2;
        lk[q2] = (q2 >= q && i > k && i < n) ? Ls[i + k * LD] : zero_of<T>();
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {  // branch-free: the NC chains interleave
        if (V2 && 16 * c > 32 * q + 31) continue;  // compile-time: column warp+16c lies below this whole row slice
        const bool active = (k >= warp + 16 * c);
        T rk;
        if constexpr (ScalarOf<T>::cplx) { rk.x = __shfl_sync(0xffffffffu, r[c][q].x, kk); rk.y = __shfl_sync(0xffffffffu, r[c][q].y, kk); }
        else rk = __shfl_sync(0xffffffffu, r[c][q], kk);
        const T xk = active ? rscale(rk, dk) : zero_of<T>();
        if (lane == kk && active) r[c][q] = xk;
        const T nxk = neg(xk);
#pragma unroll
        for (int q2 = 0; q2 < NQ; ++q2)
          if (q2 >= q) fma_acc(r[c][q2], lk[q2], nxk);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int j = warp + 16 * c;
    if (j >= n) continue;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int i = lane + 32 * q;
      if (i < n) {  // the other triangle of X is zeroed: the panel solve uses X as a dense operand
        const T v = (i >= j) ? r[c][q] : zero_of<T>();
        if (!upper) X[(long long)j * ldx + i] = v; else X[(long long)i * ldx + j] = conj_of(v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(PANEL_THREADS) lauu2_kernel(int upper, int n, T *A, long long lda) {
  CHOL_DYN_SMEM(smem_raw);
  T *L = reinterpret_cast<T *>(smem_raw);
  const int LD = n | 1;
  load_lower_form(L, LD, upper, true, n, A, lda);  // upper: L = U^H, so U*U^H = L^H*L
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // R(a,b) = sum_{k>=a} conj(L(k,a)) * L(k,b),  a >= b
  for (int b = warp; b < n; b += nwarps)
    for (int a = b + lane; a < n; a += 32) {
      T acc = zero_of<T>();
      for (int k = a; k < n; ++k) fma_acc(acc, conj_of(L[k + a * LD]), L[k + b * LD]);
      if (!upper) A[(long long)b * lda + a] = acc; else A[(long long)a * lda + b] = conj_of(acc);
    }
}

// ------------------------------------------------------------------------------------------------
// uplo filter on SOURCE coordinates: 0 all, 1 lower, 2 upper, 3 strictly lower, 4 strictly upper
__device__ __forceinline__ bool tri_keep(int uplo, int r, int c) {
  return uplo == 0 || (uplo == 1 && r >= c) || (uplo == 2 && r <= c) || (uplo == 3 && r > c) || (uplo == 4 && r < c);
}

template <typename T>
__global__ void __launch_bounds__(256) copy_kernel(int trans, int uplo, int m, int n, const T *A, long long lda, T *B, long long ldb) {
  __shared__ T tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (trans == 0) {
    for (int cc = ty; cc < 32; cc += 8) {
      const int r = r0 + tx, c = c0 + cc;
      if (r < m && c < n && tri_keep(uplo, r, c))
        B[(long long)c * ldb + r] = A[(long long)c * lda + r];
    }
    return;
  }
  for (int cc = ty; cc < 32; cc += 8) {
    const int r = r0 + tx, c = c0 + cc;
    if (r < m && c < n) tile[cc][tx] = A[(long long)c * lda + r];
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int r = r0 + rr, c = c0 + tx;  // source coordinates; destination is B(c, r)
    if (r < m && c < n && tri_keep(uplo, r, c)) {
      T v = tile[tx][rr];
      B[(long long)r * ldb + c] = (trans == 2) ? conj_of(v) : v;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(int uplo, int m, int n, T value, T *A, long long lda) {
  const int r = blockIdx.x * 32 + (threadIdx.x & 31);
  for (int cc = threadIdx.x >> 5; cc < 32; cc += 8) {
    const int c = blockIdx.y * 32 + cc;
    if (r < m && c < n && (uplo == 0 || (uplo == 1 && r > c) || (uplo == 2 && r < c))) A[(long long)c * lda + r] = value;
  }
}

// ------------------------------------------------------------------------------------------------
// logdet: 2*sum log(Re x[i*incx]).  Single launch; every block reduces a strided slice, the last
// block to finish (atomic ticket, as in the reference's threadfence reduction) sums the partials in
// a FIXED order, so the result is deterministic.  d/z use the double-precision log (the
// reference's z path uses logf, lapack/zlogdet.c:18 - see DESIGN.md), s/c use logf.
constexpr int LOGDET_BLOCKS = 296;
constexpr int LOGDET_THREADS = 256;
struct LogdetScratch { double partial[LOGDET_BLOCKS]; unsigned int ticket; };

template <typename T>
__global__ void __launch_bounds__(LOGDET_THREADS) logdet_kernel(const T *x, long long incx, long long n,
                                                                typename ScalarOf<T>::Real *result, LogdetScratch *scr) {
  using Real = typename ScalarOf<T>::Real;
  __shared__ double wsum[LOGDET_THREADS / 32];
  __shared__ bool last;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const Real v = re_of(x[i * incx]);
    if constexpr (sizeof(Real) == 4) acc += (double)logf(v); else acc += log(v);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < LOGDET_THREADS / 32; ++w) s += wsum[w];
    scr->partial[blockIdx.x] = s;
    __threadfence();
    const unsigned int t = atomicAdd(&scr->ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double s = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) s += ((volatile double *)scr->partial)[b];
    *result = (Real)(2.0 * s);
    scr->ticket = 0;  // ready for the next call
  }
}

#ifndef CHOL_CPU_EMU  // everything below launches kernels or uses inline PTX: not part of the CPU emulation
// ------------------------------------------------------------------------------------------------
// FP64 peak probes (registers only, no memory traffic):
//   kind 0 = DMMA, the GEMM engine's own instruction mix: 8 A-fragments x 4 B-fragments -> 32
//            independent m8n8k4 accumulators per warp
//   kind 1 = DFMA only (the FP64 vector pipe), kind 2 = both interleaved in the same warps,
//   kind 3 = DMMA with a single shared operand pair (an upper bound on operand-port effects)
template <int KIND>
__global__ void __launch_bounds__(128, 2) probe_fp64_kernel(double *out, int iters) {
  double acc[8][4][2], a[8], b[4], f[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + (threadIdx.x + i) * 1e-9;
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = 1.0 - (threadIdx.x + j) * 1e-9;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = i;
  for (int it = 0; it < iters; ++it) {
    if (KIND == 0 || KIND == 2 || KIND == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], KIND == 3 ? a[0] : a[i], KIND == 3 ? b[0] : b[j]);
    }
    if (KIND == 1 || KIND == 2) {
#pragma unroll
      for (int r = 0; r < 16; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a[i & 7], b[i & 3]);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += f[i];
  if (s == 12345.678) out[0] = s;
}

template <typename T> static size_t panel_smem(size_t n, size_t extra) { return ((n | 1) * n + extra) * sizeof(T); }

template <typename K> static int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return 0;
}

template <typename T>
static int potf2_t(int uplo, size_t n, void *A, size_t lda, long *dinfo, long base, cudaStream_t st) {
  constexpr int NMAX = sizeof(T) == 16 ? 64 : 128;  // chol_diag_max
  potf2_kernel<T, NMAX / 32, NMAX / 16><<<1, PANEL_THREADS, 0, st>>>(uplo == 'U', (int)n, (T *)A, (long long)lda, dinfo, base);
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}
template <typename T>
static int potf2_inv_t(int uplo, size_t n, void *A, size_t lda, void *X, size_t ldx, long *dinfo, long base, cudaStream_t st) {
  constexpr int NMAX = sizeof(T) == 16 ? 64 : 128;
  using Real = typename ScalarOf<T>::Real;
  const size_t sm = (size_t)NMAX * (NMAX + 1) * sizeof(T) + NMAX * sizeof(Real);
  // V2 = short serial section (d: 109.6 -> 82.9 us per 128-block on B200, profiles/r02_staged_ab.txt); the V = false
  // instantiation only exists in the CPU emulation, which checks that both give the same bits
  auto kern = potf2_inv_kernel<T, NMAX / 32, NMAX / 16, NMAX / 32, NMAX / 16, true>;
  int e = set_smem(kern, sm); if (e) return e;
  kern<<<1, PANEL_THREADS, sm, st>>>(uplo == 'U', (int)n, (T *)A, (long long)lda, (T *)X, (long long)ldx, dinfo, base);
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}
template <typename T>
static int trti2_t(int uplo, int diag, size_t n, const void *A, size_t lda, void *B, size_t ldb, int zo, long *dinfo, long base, cudaStream_t st) {
  const size_t sm = panel_smem<T>(n, n);
  constexpr int NMAX = sizeof(T) == 16 ? 64 : 128;
  auto kern = trti2_kernel<T, NMAX / 32, NMAX / 16, true>;
  int e = set_smem(kern, sm); if (e) return e;
  kern<<<1, PANEL_THREADS, sm, st>>>(uplo == 'U', diag == 'U', (int)n, (const T *)A, (long long)lda, (T *)B, (long long)ldb, zo, dinfo, base);
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}
template <typename T>
static int lauu2_t(int uplo, size_t n, void *A, size_t lda, cudaStream_t st) {
  const size_t sm = panel_smem<T>(n, 0);
  int e = set_smem(lauu2_kernel<T>, sm); if (e) return e;
  lauu2_kernel<T><<<1, PANEL_THREADS, sm, st>>>(uplo == 'U', (int)n, (T *)A, (long long)lda);
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}
template <typename T>
static int copy_t(int trans, int uplo, size_t m, size_t n, const void *A, size_t lda, void *B, size_t ldb, cudaStream_t st) {
  dim3 grid((unsigned)((m + 31) / 32), (unsigned)((n + 31) / 32));
  const int tc = trans == 'N' ? 0 : (trans == 'T' ? 1 : 2);
  const int uc = uplo == 'L' ? 1 : (uplo == 'U' ? 2 : (uplo == 'l' ? 3 : (uplo == 'u' ? 4 : 0)));
  copy_kernel<T><<<grid, 256, 0, st>>>(tc, uc, (int)m, (int)n, (const T *)A, (long long)lda, (T *)B, (long long)ldb);
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}
template <typename T>
static int fill_t(int uplo, size_t m, size_t n, T v, void *A, size_t lda, cudaStream_t st) {
  dim3 grid((unsigned)((m + 31) / 32), (unsigned)((n + 31) / 32));
  const int uc = uplo == 'L' ? 1 : (uplo == 'U' ? 2 : 0);
  fill_kernel<T><<<grid, 256, 0, st>>>(uc, (int)m, (int)n, v, (T *)A, (long long)lda);
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}
template <typename T>
static int logdet_t(const void *x, size_t incx, size_t n, void *res, void *scr, cudaStream_t st) {
  long long want = ((long long)n + LOGDET_THREADS - 1) / LOGDET_THREADS;
  int blocks = (int)(want < 1 ? 1 : (want > LOGDET_BLOCKS ? LOGDET_BLOCKS : want));
  logdet_kernel<T><<<blocks, LOGDET_THREADS, 0, st>>>((const T *)x, (long long)incx, (long long)n,
                                                    (typename ScalarOf<T>::Real *)res, (LogdetScratch *)scr);
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}

}  // namespace chol

using namespace chol;

#define DISPATCH(prec, CALL)                        \
  switch (prec) {                                   \
    case 's': { using T = float;   return CALL; }   \
    case 'd': { using T = double;  return CALL; }   \
    case 'c': { using T = float2;  return CALL; }   \
    case 'z': { using T = double2; return CALL; }   \
    default: return (int)cudaErrorInvalidValue;     \
  }

extern "C" { unsigned long long chol_launch_counter = 0; }
extern "C" unsigned long long chol_launch_count(int reset) {
  const unsigned long long v = chol_launch_counter;
  if (reset) chol_launch_counter = 0;
  return v;
}

extern "C" int chol_diag_max(int prec) { return prec == 'z' ? 64 : 128; }

extern "C" int chol_potf2_launch(int prec, int uplo, size_t n, void *A, size_t lda, long *dinfo, long info_base, void *stream) {
  if (n == 0) return 0;
  if (n > (size_t)chol_diag_max(prec)) return (int)cudaErrorInvalidValue;
  DISPATCH(prec, potf2_t<T>(uplo, n, A, lda, dinfo, info_base, (cudaStream_t)stream));
}
extern "C" int chol_potf2_inv_launch(int prec, int uplo, size_t n, void *A, size_t lda, void *X, size_t ldx, long *dinfo,
                                     long info_base, void *stream) {
  if (n == 0) return 0;
  if (n > (size_t)chol_diag_max(prec)) return (int)cudaErrorInvalidValue;
  DISPATCH(prec, potf2_inv_t<T>(uplo, n, A, lda, X, ldx, dinfo, info_base, (cudaStream_t)stream));
}
extern "C" int chol_trti2_launch(int prec, int uplo, int diag, size_t n, const void *A, size_t lda, void *B, size_t ldb,
                                 int zero_other, long *dinfo, long info_base, void *stream) {
  if (n == 0) return 0;
  if (n > (size_t)chol_diag_max(prec)) return (int)cudaErrorInvalidValue;
  DISPATCH(prec, trti2_t<T>(uplo, diag, n, A, lda, B, ldb, zero_other, dinfo, info_base, (cudaStream_t)stream));
}
extern "C" int chol_lauu2_launch(int prec, int uplo, size_t n, void *A, size_t lda, void *stream) {
  if (n == 0) return 0;
  if (n > (size_t)chol_diag_max(prec)) return (int)cudaErrorInvalidValue;
  DISPATCH(prec, lauu2_t<T>(uplo, n, A, lda, (cudaStream_t)stream));
}
extern "C" int chol_copy_launch(int prec, int trans, int uplo, size_t m, size_t n, const void *A, size_t lda, void *B,
                                size_t ldb, void *stream) {
  if (m == 0 || n == 0) return 0;
  DISPATCH(prec, copy_t<T>(trans, uplo, m, n, A, lda, B, ldb, (cudaStream_t)stream));
}
extern "C" int chol_fill_launch(int prec, int uplo, size_t m, size_t n, double re, double im, void *A, size_t lda, void *stream) {
  if (m == 0 || n == 0) return 0;
  DISPATCH(prec, fill_t<T>(uplo, m, n, mk(T(), (typename ScalarOf<T>::Real)re, (typename ScalarOf<T>::Real)im), A, lda, (cudaStream_t)stream));
}
extern "C" size_t chol_logdet_scratch_bytes(void) { return sizeof(LogdetScratch); }
extern "C" int chol_logdet_launch(int prec, const void *x, size_t incx, size_t n, void *dresult, void *dscratch, void *stream) {
  DISPATCH(prec, logdet_t<T>(x, incx, n, dresult, dscratch, (cudaStream_t)stream));
}

extern "C" int chol_probe_fp64(int kind, int iters, double *tflops, float *ms_out) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double *out = nullptr;
  cudaMalloc(&out, 8);
  const int per_sm = kind >= 10 ? kind / 10 : 2;  // kind 10/20/..: DMMA with 1/2/.. CTAs (4 warps each) per SM
  if (kind >= 10) kind = 0;
  const int blocks = sms * per_sm, threads = 128;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {  // first pass warms up
    cudaEventRecord(e0);
    if (kind == 0) probe_fp64_kernel<0><<<blocks, threads>>>(out, iters);
    else if (kind == 1) probe_fp64_kernel<1><<<blocks, threads>>>(out, iters);
    else if (kind == 2) probe_fp64_kernel<2><<<blocks, threads>>>(out, iters);
    else probe_fp64_kernel<3><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warps = (double)blocks * threads / 32.0;
  double flops = 0.0;
  if (kind == 0 || kind == 2 || kind == 3) flops += warps * iters * 32.0 * (8 * 8 * 4 * 2.0);
  if (kind == 1 || kind == 2) flops += warps * iters * 256.0 * 32 * 2.0;
  *tflops = flops / (ms * 1e-3) / 1e12;
  if (ms_out) *ms_out = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  return (int)cudaGetLastError();
}
#else
}  // namespace chol (the emulation build stops before the launchers)
#endif  // CHOL_CPU_EMU
