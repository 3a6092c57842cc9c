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

// Stage the triangle of an n x n block (n <= 128) in shared memory in LOWER form: S[i + k*LD] (i >= k) = A(i,k) for
// lower storage, op(A(k,i)) for upper storage.  Lanes run along the contiguous direction of the storage, and every
// thread issues a batch of independent loads before the first store: a one-CTA kernel has nothing else to hide
// the HBM latency with (the round-1 loop, one load -> one store per trip, cost ~20 round trips of ~0.6 us each).
template <typename T>
__device__ __forceinline__ void load_lower_form(T *S, int LD, int upper, bool conj_upper, int n, const T *A, long long lda) {
  constexpr int U = 8;
  const int total = 128 * n;  // idx = c * 128 + r: r fast (rows of the stored column), c = stored column
  for (int base = threadIdx.x; base < total; base += U * blockDim.x) {
    T v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * blockDim.x, r = idx & 127, c = idx >> 7;
      v[u] = zero_of<T>();
      if (idx < total && r < n && (upper ? r <= c : r >= c)) v[u] = A[(long long)c * lda + r];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * blockDim.x, r = idx & 127, c = idx >> 7;
      if (idx < total && r < n) {
        if (!upper) { if (r >= c) S[r + c * LD] = v[u]; }
        else if (r <= c) S[c + r * LD] = conj_upper ? conj_of(v[u]) : v[u];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
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
template <typename T>
__device__ __forceinline__ T shfl_elem(T v, int src) {
  if constexpr (ScalarOf<T>::cplx) {
    T r;
    r.x = __shfl_sync(0xffffffffu, v.x, src);
    r.y = __shfl_sync(0xffffffffu, v.y, src);
    return r;
  } else {
    return __shfl_sync(0xffffffffu, v, src);
  }
}

// potf2: Cholesky of one diagonal block (n <= 128; z: 64) by ONE CTA.  The trailing part of the block lives in
// REGISTERS: thread (lane, w) of the 16 warps owns rows lane + 32a (a < NA) of the CW columns CW*w .. CW*w + CW-1
// ("panel w").  Blocked right-looking over the 16 panels:
//   * the owning warp factors its panel BY ITSELF in shared memory (potf2_panel): the serial chain of a column is
//     broadcast load -> rsqrt -> scale -> FMA -> store with no CTA barrier in it (measured on B200: DFMA 8,
//     shuffle 30, rsqrt 62, shared load 29 cycles; r02 call 5), in a few dozen instructions of rolled loops;
//   * the finished panel is thereby published, ONE __syncthreads per panel (16 per block, not 128), and every warp
//     right of it applies the rank-CW update to its own panel from shared memory;
//   * look-ahead: the owner of panel p+1 factors and publishes it right after its own update, while the other
//     warps are still busy with theirs; warps left of the panel have retired and only meet the barriers.
// Entries above the diagonal are never read as matrix data (the published columns have zeros there), so the update
// needs no per-element mask.  History: the round-1/2 kernels did one barrier per COLUMN with a masked, branch-free
// rank-1 update on all 16 warps - issue-bound at 600-850 cycles per column (80 us fused with the inverse, 45-51 us
// alone); fully unrolled register-only variants of this kernel ran to 12k instructions (instruction-fetch bound).
// One warp factors its CW-column panel (columns c0 .. c0+CW-1, rows >= c0) held in reg[a][b] and publishes it in Ls.
// The serial recurrence of a Cholesky column is SCALAR - d(c+1) needs only rsqrt(d(c)) and one entry of column c - so
// every lane factors the panel's CW x CW diagonal block redundantly in its own registers (broadcast loads, then
// rsqrt -> multiply -> FMA per column: ~80 cycles, no shuffle, no shared-memory round trip, no branch), and then
// solves its own rows against that block by substitution, also in registers.  Measured alternatives (r02 calls
// 5, 6): a barrier per column ~600 cycles per column, a shared-memory panel with per-column branches ~450.
template <typename T, int NA, int CW, int LD>
__device__ __forceinline__ void potf2_panel(T (&reg)[NA][CW], T *Ls, int c0, int n, int lane, int *sfail, T *A, long long lda,
                                            long *dinfo, long info_base) {
  using Real = typename ScalarOf<T>::Real;
  // the panel enters shared memory as it is (rows above the diagonal block hold finite garbage that is never used)
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int b = 0; b < CW; ++b) Ls[lane + 32 * a + (c0 + b) * LD] = reg[a][b];
  __syncwarp();
  T D[CW][CW];  // lower triangle of the diagonal block, the same in every lane
#pragma unroll
  for (int s = 0; s < CW; ++s)
#pragma unroll
    for (int r = s; r < CW; ++r) D[r][s] = Ls[(c0 + r) + (c0 + s) * LD];
  Real rinv[CW];
  int bad = 0;  // first failing column of the panel + 1 (uniform over the warp)
  Real dbad = Real(0);
#pragma unroll
  for (int s = 0; s < CW; ++s) {
    const Real d = re_of(D[s][s]);
    const bool live = c0 + s < n;  // columns beyond n are zero: keep them zero instead of dividing by it
    if (live && !(d > Real(0)) && bad == 0) { bad = s + 1; dbad = d; }  // <= 0 or NaN: not positive definite (lapack/dpotrf.c:70-73)
    rinv[s] = live ? rsqrt(d) : Real(0);
    D[s][s] = from_real<T>(d * rinv[s]);
#pragma unroll
    for (int r = s + 1; r < CW; ++r) D[r][s] = rscale(D[r][s], rinv[s]);
#pragma unroll
    for (int s2 = s + 1; s2 < CW; ++s2) {
      const T m = neg(conj_of(D[s2][s]));
#pragma unroll
      for (int r = s2; r < CW; ++r) fma_acc(D[r][s2], D[r][s], m);
    }
  }
  if (bad != 0 && lane == 0 && *sfail == 0) {  // the columns after it hold garbage that nobody uses: the CTA leaves at the next barrier
    const int c = c0 + bad - 1;
    *sfail = c + 1;
    *dinfo = info_base + c + 1;
    A[(long long)c * lda + c] = from_real<T>(dbad);  // the reference leaves the updated pivot in place (dpotrf.c:45-49)
  }
  // every lane's rows: x(i, :) = a(i, :) * Ldd^-H by substitution (right-looking over the CW columns); for a row
  // INSIDE the diagonal block the same recurrence reproduces the factor's row, beyond its diagonal it is masked.
  // The rows are read back from the shared copy (the register copy is dead since the dump: the diagonal block
  // above needs the registers).
  __syncwarp();
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    const int i = lane + 32 * a;
    T x[CW];
#pragma unroll
    for (int s = 0; s < CW; ++s) x[s] = Ls[i + (c0 + s) * LD];
#pragma unroll
    for (int s = 0; s < CW; ++s) {
      x[s] = rscale(x[s], rinv[s]);
#pragma unroll
      for (int s2 = s + 1; s2 < CW; ++s2) fma_acc(x[s2], x[s], neg(conj_of(D[s2][s])));
    }
#pragma unroll
    for (int s = 0; s < CW; ++s) {
      T v = (i >= c0 + s) ? x[s] : zero_of<T>();
      if constexpr (ScalarOf<T>::cplx) { if (i == c0 + s) v.y = 0; }  // the diagonal is real
      Ls[i + (c0 + s) * LD] = v;
    }
  }
}

template <typename T, int NA, int CW, bool DBG = false>
__global__ void __launch_bounds__(PANEL_THREADS) potf2_kernel(int upper, int n, T *A, long long lda, long *dinfo, long info_base,
                                                              double *logacc, long long *dbg = nullptr) {
  CHOL_DYN_SMEM(smem_raw);
  constexpr int NMAX = 32 * NA, LD = NMAX + 1;
  static_assert(16 * CW == NMAX, "16 warps x CW columns");
  T *Ls = reinterpret_cast<T *>(smem_raw);  // L(i,c) at Ls[i + c*LD], zeros above the diagonal and beyond n
  __shared__ int sfail;                      // first failing column + 1
  // (the "already failed" test of *dinfo is taken AFTER the staging loads have been issued: a dependent global load
  // in front of them costs a full memory round trip on the serial chain of the factorisation)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) sfail = 0;
#ifndef CHOL_CPU_EMU
#define CHOL_DBG(slot) do { if constexpr (DBG) { if (lane == 0) dbg[slot] = clock64(); } } while (false)  /* phase timestamps (scripts/potf2_ab.py) */
#else
#define CHOL_DBG(slot) ((void)0)
#endif
  if (warp == 0) CHOL_DBG(0);
  // The block enters through the shared buffer, so that the HBM reads are coalesced for both storages (lanes along
  // the stored columns; upper: transposed on the way in).  Every thread issues ALL its loads before the first store -
  // a one-CTA kernel has nothing else to hide the HBM latency with - and writes zeros for everything outside the
  // triangle / beyond n, so the buffer is fully defined.
  {
    constexpr int PER = NMAX * NMAX / PANEL_THREADS;
    T v[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int idx = threadIdx.x + u * PANEL_THREADS, r = idx % NMAX, c = idx / NMAX;  // r fast: rows of the stored column c
      v[u] = zero_of<T>();
      if (r < n && c < n && (upper ? r <= c : r >= c)) v[u] = A[(long long)c * lda + r];
    }
    if (*dinfo != 0) return;  // an earlier block has failed: nothing is touched (uniform over the CTA)
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int idx = threadIdx.x + u * PANEL_THREADS, r = idx % NMAX, c = idx / NMAX;
      if (!upper) Ls[r + c * LD] = v[u];                 // (r, c): the triangle, zeros above it
      else if (r <= c) Ls[c + r * LD] = conj_of(v[u]);   // L = U^H: element (c, r) of the lower form
      else Ls[c + r * LD] = zero_of<T>();                // r > c: slot (c, r) lies above the diagonal of the lower form
    }
  }
  __syncthreads();
  T reg[NA][CW];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int b = 0; b < CW; ++b) reg[a][b] = Ls[lane + 32 * a + (CW * warp + b) * LD];  // zeros above the diagonal / beyond n
  const int np = (n + CW - 1) / CW;  // panels (uniform: n is a kernel argument)
  // Round p starts with panel p published.  A failure in it (or before) was flagged BEFORE the barrier that published
  // it; a flag raised meanwhile by the owner of panel p+1 lies beyond CW*p + CW and is ignored until the next round,
  // so every thread leaves after the same barrier.
#define CHOL_POTF2_LEAVE_IF_FAILED(p)                                                                              \
  do {                                                                                                             \
    const int fail = sfail;                                                                                        \
    if (fail != 0 && fail <= CW * (p) + CW) {                                                                      \
      store_factor_columns<T, NA, LD>(upper, n, fail - 1, Ls, A, lda, warp, lane); /* the columns before it */     \
      return;                                                                                                      \
    }                                                                                                              \
  } while (false)
  // rank-CW update of this warp's panel with panel p: a(i,k) -= sum_j L(i,cj) conj(L(k,cj)), cj = CW*p + j
  const int amin = (CW * warp) >> 5;  // slices above it lie entirely above this warp's diagonal block
  auto update = [&](int p) {
#pragma unroll
    for (int j = 0; j < CW; ++j) {
      const T *col = Ls + (CW * p + j) * LD;
      T li[NA], f[CW];
#pragma unroll
      for (int a = 0; a < NA; ++a) li[a] = col[lane + 32 * a];
#pragma unroll
      for (int b = 0; b < CW; ++b) f[b] = neg(conj_of(col[CW * warp + b]));
#pragma unroll
      for (int a = 0; a < NA; ++a)
        if (a >= amin) {
#pragma unroll
          for (int b = 0; b < CW; ++b) fma_acc(reg[a][b], li[a], f[b]);
        }
    }
  };
  int p = 0;
  if (warp == 0) {  // panel 0 needs no update
    CHOL_DBG(1);
    potf2_panel<T, NA, CW, LD>(reg, Ls, 0, n, lane, &sfail, A, lda, dinfo, info_base);
    CHOL_DBG(2);
    __syncthreads();  // panel 0 is published
  } else {
    __syncthreads();
    // rounds in which this warp updates its panel and waits ...
#pragma unroll 1
    for (; p + 1 < warp && p < np; ++p) {
      CHOL_POTF2_LEAVE_IF_FAILED(p);
      update(p);
      __syncthreads();
    }
    // ... the round in which it is the look-ahead owner: last update, then it factors and publishes its own panel
    // while the warps right of it are still busy with their updates (the registers are dead from there on, which is
    // what lets the panel routine keep its diagonal block in registers) ...
    if (p < np && p == warp - 1) {
      CHOL_POTF2_LEAVE_IF_FAILED(p);
      CHOL_DBG(4 + 3 * p);
      update(p);
      CHOL_DBG(5 + 3 * p);
      if (CW * warp < n) potf2_panel<T, NA, CW, LD>(reg, Ls, CW * warp, n, lane, &sfail, A, lda, dinfo, info_base);
      CHOL_DBG(6 + 3 * p);
      __syncthreads();
      ++p;
    }
  }
  // ... and the rounds after it has retired: only the barriers
#pragma unroll 1
  for (; p < np; ++p) {
    CHOL_POTF2_LEAVE_IF_FAILED(p);
    __syncthreads();
  }
#undef CHOL_POTF2_LEAVE_IF_FAILED
  if (warp == 0) CHOL_DBG(60);
  {
    const int fail = sfail;  // a failure in the last panel
    store_factor_columns<T, NA, LD>(upper, n, fail != 0 ? fail - 1 : n, Ls, A, lda, warp, lane);
    // fused logdet ("the diagonal is in shared memory anyway"): *logacc += 2 sum log L(c,c), off
    // the serial chain, summed in a fixed order; the diagonal-block launches of one factorisation are serialised on
    // their stream, so a plain read-modify-write by one thread is enough
    if (logacc != nullptr && fail == 0) {
      __shared__ double part[PANEL_THREADS / 32];
      const int c = CW * warp + lane;
      double v = (lane < CW && c < n) ? log((double)re_of(Ls[c + c * LD])) : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) part[warp] = v;
      __syncthreads();
      if (threadIdx.x == 0) {
        double sum = 0.0;
        for (int q = 0; q < PANEL_THREADS / 32; ++q) sum += part[q];
        *logacc += 2.0 * sum;
      }
    }
  }
  if (warp == 0) CHOL_DBG(61);
#undef CHOL_DBG
}

// ------------------------------------------------------------------------------------------------
// trsm_rows: X = B * L^-H for an n x n lower-triangular block L (n <= 128; z: 64) and ANY number of rows m -
// the panel solve of POTRF by substitution (what the reference's dtrsm does, blas/dtrsm.c:34-216; lower:
// W = P L^-H, upper: W = P^H U^-1 with U = L^H), without forming an inverse.  The rows of B are independent
// recurrences  x_k = (b_k - sum_{j<k} x_j conj(L_kj)) / L_kk,  so the kernel scales across SMs by rows while each
// row's chain stays in registers:
//   * L^H is staged once per CTA in shared memory (row k = the coefficients conj(L_jk), j > k, contiguous in j),
//     the reciprocals of its diagonal next to it;
//   * a row (RR rows) is owned by a GROUP of G lanes: lane g of the group holds columns g, g+G, g+2G, ... of
//     the right-hand side in registers; step k: the owner's b_k is broadcast by shuffle, scaled to x_k, and
//     every lane applies  b_j -= x_k conj(L_jk)  to the columns j > k it holds (right-looking), so a step is
//     one shuffle + one multiply + a few independent FMAs deep;
//   * G = 32, RR = 1 (one row per warp, 4 columns per lane) keeps the chain short for the small panels of the
//     diagonal-block recursion; G = 8, RR = 4 (16 rows per warp, every coefficient load feeds 4 FMAs) is the
//     throughput shape for the tall panel.
// transB / transX: the operand is given / wanted as its conjugate transpose (n x m), i.e. element (row, j) sits
// at ptr[j + row * ld] conjugated - upper storage, where the panel is the block ROW right of the diagonal block.
template <typename T, int NMAX, int G, int RR>
__global__ void __launch_bounds__(256) trsm_rows_kernel(int upper, int n, const T *L, long long ldl, long long m, int cwarps, int transB,
                                                        const T *B, long long ldb, int transX, T *X, long long ldx,
                                                        const long *dinfo) {
  using Real = typename ScalarOf<T>::Real;
  constexpr int LD = NMAX + 1, NT = NMAX / G, NG = 32 / G;
  CHOL_DYN_SMEM(smem_raw);
  T *Ms = reinterpret_cast<T *>(smem_raw);                         // Ms[k * LD + j] = conj(L(j,k)) / L(k,k), j > k
  Real *dinv = reinterpret_cast<Real *>(Ms + (size_t)NMAX * LD);   // 1 / L(k,k)
  // With y_k = b_k - sum_{j<k} x_j conj(L_kj) and x_k = y_k / L_kk the recurrence reads
  //     y_k = b_k - sum_{j<k} y_j * M[j][k],   M[j][k] = conj(L_kj) / L_jj,
  // so the serial chain of a step is one shuffle and one FMA (no multiply); x = y / diag(L) at the very end.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane % NG, lg = lane / NG;  // group within the warp, lane within the group
  // rows of the warp: row0 + rr*NG + g (consecutive lanes of one lg read consecutive rows of a column)
  const long long row0 = ((long long)blockIdx.x * cwarps + warp) * (NG * RR);
  const bool owns_rows = warp < cwarps && row0 < m;  // helper warps only stage
  T b[RR][NT];
  auto load_rows = [&]() {
#pragma unroll
    for (int rr = 0; rr < RR; ++rr) {
      const long long row = row0 + rr * NG + g;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int j = lg + G * t;
        T v = zero_of<T>();
        if (owns_rows && row < m && j < n) v = transB ? conj_of(B[(long long)row * ldb + j]) : B[(long long)j * ldb + row];
        b[rr][t] = v;
      }
    }
  };
  // one-row-per-warp shape: the few right-hand side values are requested first and arrive while the triangle is staged
  // (the tall shape needs its registers for the staging batch)
  constexpr bool EARLY_ROWS = RR * NT <= 8;
  if constexpr (EARLY_ROWS) load_rows();
  // all 256 threads of the CTA stage (also the warps that own no rows), ALL loads of a thread in flight at once;
  // lanes along the contiguous direction of the storage: lower -> j (rows of L), upper -> k (rows of U).  The
  // reciprocals of the diagonal are computed while those loads are in flight.
  {
    constexpr int PER = NMAX * NMAX / 256;
    T v[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int idx = threadIdx.x + u * 256, fast = idx % NMAX, slow = idx / NMAX;
      const int j = upper ? slow : fast, k = upper ? fast : slow;
      v[u] = zero_of<T>();
      if (j < n && j > k) v[u] = upper ? L[(long long)j * ldl + k] : conj_of(L[(long long)k * ldl + j]);
    }
    if (threadIdx.x < NMAX) dinv[threadIdx.x] = ((int)threadIdx.x < n) ? Real(1) / re_of(L[(long long)threadIdx.x * ldl + threadIdx.x]) : Real(0);
    // the factorisation has already failed: leave X alone (tested here, behind the loads, not in front of them)
    if (dinfo != nullptr && *dinfo != 0) return;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int idx = threadIdx.x + u * 256, fast = idx % NMAX, slow = idx / NMAX;
      const int j = upper ? slow : fast, k = upper ? fast : slow;
      Ms[k * LD + j] = rscale(v[u], dinv[k]);
    }
  }
  __syncthreads();
  if (!owns_rows) return;
  if constexpr (!EARLY_ROWS) load_rows();
#pragma unroll
  for (int kb = 0; kb < NT; ++kb) {
    const int kend = min(G, n - G * kb);  // <= 0 beyond the block: nothing to do (keeps the kb loop fully unrollable)
    // the steps of one register slice stay a LOOP: fully unrolled, the 128 steps are ~10k straight-line instructions
    // that every warp executes once - instruction-fetch bound (875 cycles per step measured, r02 call 5); two steps per
    // trip let the coefficient loads of one step overlap the FMAs of the other
#pragma unroll 2
    for (int kk = 0; kk < kend; ++kk) {
      const int k = G * kb + kk;
      const int src = g + NG * kk;
      T yk[RR];  // -y_k
#pragma unroll
      for (int rr = 0; rr < RR; ++rr) {
        T bk;
        if constexpr (ScalarOf<T>::cplx) { bk.x = __shfl_sync(0xffffffffu, b[rr][kb].x, src); bk.y = __shfl_sync(0xffffffffu, b[rr][kb].y, src); }
        else bk = __shfl_sync(0xffffffffu, b[rr][kb], src);
        yk[rr] = neg(bk);
      }
      const T *mrow = Ms + k * LD + lg;
      if (lg > kk) {  // slice kb: only the columns right of k (the owner's and the finished ones keep their y)
        const T c0 = mrow[G * kb];
#pragma unroll
        for (int rr = 0; rr < RR; ++rr) fma_acc(b[rr][kb], yk[rr], c0);
      }
#pragma unroll
      for (int t = kb + 1; t < NT; ++t) {
        const T c1 = mrow[G * t];
#pragma unroll
        for (int rr = 0; rr < RR; ++rr) fma_acc(b[rr][t], yk[rr], c1);
      }
    }
  }
#pragma unroll
  for (int rr = 0; rr < RR; ++rr) {
    const long long row = row0 + rr * NG + g;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int j = lg + G * t;
      if (row < m && j < n) {
        const T x = rscale(b[rr][t], dinv[j]);
        if (transX) X[(long long)row * ldx + j] = conj_of(x); else X[(long long)j * ldx + row] = x;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// trti2: X = L^-1 by forward substitution, the triangle staged in shared memory.  Every warp carries
// EIGHT columns at once (columns warp+16c): the residuals live in registers (lane = row mod 32), the
// current unknown is broadcast by shuffle, and the eight independent chains hide each other's latency
// while sharing the loads of L's column k.
// Chains that cannot have started in the current
// 32-row slice are pruned at compile time.
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

// B(g(r), c) := A(r, c) for an m x n block: the rows of a rank's packed piece go to their block-cyclic positions
// g(r) = ((r / cnb) * cp + coff) * cnb + r % cnb  (the distributed LAUUM assembles a row block of the factor this way)
template <typename T>
__global__ void __launch_bounds__(256) scatter_rows_kernel(int m, int n, const T *A, long long lda, T *B, long long ldb, int cnb, int cp,
                                                           int coff) {
  const int r = blockIdx.x * 32 + (threadIdx.x & 31);
  if (r >= m) return;
  const long long gr = ((long long)(r / cnb) * cp + coff) * cnb + r % cnb;
  for (int cc = threadIdx.x >> 5; cc < 32; cc += 8) {
    const int c = blockIdx.y * 32 + cc;
    if (c < n) B[(long long)c * ldb + gr] = A[(long long)c * lda + r];
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

// dependent-issue latencies of the FP64 pipe (one warp, clock64 around a chain of N dependent operations):
//   kind 0 DFMA, 1 DMUL, 2 rsqrt(double) (MUFU.RSQ64H + Newton steps), 3 64-bit shuffle, 4 shared-memory load (pointer chase)
__global__ void probe_latency_kernel(int kind, int n, double seed, long long *cycles, double *sink) {
  __shared__ int chase[32];
  if (threadIdx.x < 32) chase[threadIdx.x] = (threadIdx.x + 1) & 31;
  __syncwarp();
  double x = seed, y = 1.0 + 1e-9 * threadIdx.x;
  int p = threadIdx.x & 31;
  const long long t0 = clock64();
  if (kind == 0) for (int i = 0; i < n; ++i) x = fma(x, y, seed);
  else if (kind == 1) for (int i = 0; i < n; ++i) x = x * y;
  else if (kind == 2) for (int i = 0; i < n; ++i) x = rsqrt(x) + seed;
  else if (kind == 3) for (int i = 0; i < n; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  else for (int i = 0; i < n; ++i) p = chase[p];
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  if (x == 12345.678 || p == 77) *sink = x;
}

template <typename T> static size_t panel_smem(size_t n, size_t extra) { return ((n | 1) * n + extra) * sizeof(T); }

template <typename K> static int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return 0;
}

static long long *g_potf2_dbg = nullptr;  // test hook: 64 phase timestamps of the next potf2 launches (chol_potf2_set_debug)
template <typename T>
static int potf2_t(int uplo, size_t n, void *A, size_t lda, long *dinfo, long base, double *logacc, cudaStream_t st) {
  constexpr int NMAX = sizeof(T) == 16 ? 64 : 128;  // chol_diag_max
  const size_t sm = (size_t)NMAX * (NMAX + 1) * sizeof(T);
  if (g_potf2_dbg != nullptr) {  // instrumented instantiation (test hook)
    auto kd = potf2_kernel<T, NMAX / 32, NMAX / 16, true>;
    int ed = set_smem(kd, sm); if (ed) return ed;
    kd<<<1, PANEL_THREADS, sm, st>>>(uplo == 'U', (int)n, (T *)A, (long long)lda, dinfo, base, logacc, g_potf2_dbg);
    return (int)cudaGetLastError();
  }
  auto kern = potf2_kernel<T, NMAX / 32, NMAX / 16>;  // 16 warps x NMAX/16 columns
  int e = set_smem(kern, sm); if (e) return e;
  kern<<<1, PANEL_THREADS, sm, st>>>(uplo == 'U', (int)n, (T *)A, (long long)lda, dinfo, base, logacc, nullptr);
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}
// rows up to which the one-row-per-warp shape is used (8 warps per CTA, one CTA per SM: a single wave)
constexpr size_t TRSM_ROWS_SMALL = 148 * 8;
template <typename T>
static int trsm_rows_t(int uplo, size_t n, const void *L, size_t ldl, size_t m, int transB, const void *B, size_t ldb, int transX,
                       void *X, size_t ldx, const long *dinfo, cudaStream_t st) {
  constexpr int NMAX = sizeof(T) == 16 ? 64 : 128;
  using Real = typename ScalarOf<T>::Real;
  const size_t sm = (size_t)NMAX * (NMAX + 1) * sizeof(T) + NMAX * sizeof(Real);
  // 256 threads per CTA in every shape: all of them stage the triangle; `cw` warps of the CTA own rows
  if (m <= TRSM_ROWS_SMALL) {
    auto kern = trsm_rows_kernel<T, NMAX, 32, 1>;
    int e = set_smem(kern, sm); if (e) return e;
    const unsigned cw = 8, grid = (unsigned)((m + cw - 1) / cw);
    kern<<<grid, 256, sm, st>>>(uplo == 'U', (int)n, (const T *)L, (long long)ldl, (long long)m, (int)cw, transB, (const T *)B,
                                (long long)ldb, transX, (T *)X, (long long)ldx, dinfo);
  } else {
    // 16 register slices per lane at most (c, n = 128: groups of 16 lanes, so that the step loop unrolls fully)
    constexpr int GT = (ScalarOf<T>::cplx && NMAX == 128) ? 16 : 8, RPW = (32 / GT) * 4;  // rows per warp
    auto kern = trsm_rows_kernel<T, NMAX, GT, 4>;
    int e = set_smem(kern, sm); if (e) return e;
    const size_t warps = (m + RPW - 1) / RPW;
    unsigned cw = 1;
    while (cw < 8 && (warps + cw - 1) / cw > 148) cw *= 2;  // fill the SMs before stacking warps on one
    const unsigned grid = (unsigned)((warps + cw - 1) / cw);
    kern<<<grid, 256, sm, st>>>(uplo == 'U', (int)n, (const T *)L, (long long)ldl, (long long)m, (int)cw, transB, (const T *)B,
                                (long long)ldb, transX, (T *)X, (long long)ldx, dinfo);
  }
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
static int scatter_rows_t(size_t m, size_t n, const void *A, size_t lda, void *B, size_t ldb, int cnb, int cp, int coff, cudaStream_t st) {
  dim3 grid((unsigned)((m + 31) / 32), (unsigned)((n + 31) / 32));
  scatter_rows_kernel<T><<<grid, 256, 0, st>>>((int)m, (int)n, (const T *)A, (long long)lda, (T *)B, (long long)ldb, cnb, cp, coff);
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

extern "C" int chol_potf2_logdet_launch(int prec, int uplo, size_t n, void *A, size_t lda, long *dinfo, long info_base, double *logacc,
                                        void *stream) {
  if (n == 0) return 0;
  if (n > (size_t)chol_diag_max(prec)) return (int)cudaErrorInvalidValue;
  DISPATCH(prec, potf2_t<T>(uplo, n, A, lda, dinfo, info_base, logacc, (cudaStream_t)stream));
}
extern "C" int chol_potf2_launch(int prec, int uplo, size_t n, void *A, size_t lda, long *dinfo, long info_base, void *stream) {
  return chol_potf2_logdet_launch(prec, uplo, n, A, lda, dinfo, info_base, nullptr, stream);
}
extern "C" int chol_trsm_rows_launch(int prec, int uplo, size_t n, const void *L, size_t ldl, size_t m, int transB, const void *B,
                                     size_t ldb, int transX, void *X, size_t ldx, const long *dinfo, void *stream) {
  if (n == 0 || m == 0) return 0;
  if (n > (size_t)chol_diag_max(prec)) return (int)cudaErrorInvalidValue;
  DISPATCH(prec, trsm_rows_t<T>(uplo, n, L, ldl, m, transB, B, ldb, transX, X, ldx, dinfo, (cudaStream_t)stream));
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
extern "C" int chol_scatter_rows_launch(int prec, size_t m, size_t n, const void *A, size_t lda, void *B, size_t ldb, int cyc_nb, int cyc_p,
                                        int cyc_off, void *stream) {
  if (m == 0 || n == 0) return 0;
  if (cyc_nb <= 0 || cyc_p <= 0) return (int)cudaErrorInvalidValue;
  DISPATCH(prec, scatter_rows_t<T>(m, n, A, lda, B, ldb, cyc_nb, cyc_p, cyc_off, (cudaStream_t)stream));
}
extern "C" int chol_fill_launch(int prec, int uplo, size_t m, size_t n, double re, double im, void *A, size_t lda, void *stream) {
  if (m == 0 || n == 0) return 0;
  DISPATCH(prec, fill_t<T>(uplo, m, n, mk(T(), (typename ScalarOf<T>::Real)re, (typename ScalarOf<T>::Real)im), A, lda, (cudaStream_t)stream));
}
extern "C" size_t chol_logdet_scratch_bytes(void) { return sizeof(LogdetScratch); }
extern "C" int chol_logdet_launch(int prec, const void *x, size_t incx, size_t n, void *dresult, void *dscratch, void *stream) {
  DISPATCH(prec, logdet_t<T>(x, incx, n, dresult, dscratch, (cudaStream_t)stream));
}

extern "C" void chol_potf2_set_debug(long long *device_buffer_64) { g_potf2_dbg = device_buffer_64; }

extern "C" int chol_probe_latency(int kind, double *cycles_per_op) {
  long long *dc = nullptr;
  double *sink = nullptr;
  cudaMalloc(&dc, 8);
  cudaMalloc(&sink, 8);
  const int n = 4096;
  long long c = 0;
  for (int rep = 0; rep < 2; ++rep) {
    probe_latency_kernel<<<1, 32>>>(kind, n, 1.0000001, dc, sink);
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  }
  *cycles_per_op = (double)c / n;
  cudaFree(dc);
  cudaFree(sink);
  return (int)cudaGetLastError();
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
