// gemm_common.cuh - pieces shared by the FP64 DMMA engine (gemm_dmma.cu) and the FP32 FFMA engine
// (gemm_ffma.cu): the triangular super-row tile scheduler and the generic operand fetch.
#pragma once
#include <cuda_runtime.h>
#include "scalar_ops.cuh"

namespace chol {

// ---- tile scheduler ---------------------------------------------------------------------------
// Tiles are enumerated in super-rows of GROUP row-blocks; inside a super-row the column-tile index
// runs slowest and the row-block fastest, so CTAs that are resident together share the same B rows
// and a small, L2-resident set of A rows.  For triangle-only outputs (cuplo) the grid holds just
// the column tiles that can touch the triangle.
//
// Block-cyclic column map (multi-GPU): a rank's packed local columns n map to global triangle
// coordinates  gcol(n) = ((n / cnb) * cP + coff) * cnb + n % cnb  (cnb = 0: identity).  The
// lower-triangle test is then  m >= gcol(n)  and the B operand row of column n is gcol(n).
struct ColMap {
  int cnb, cP, coff;
};
__host__ __device__ inline long long gcol(const ColMap &cm, long long n) {
  return cm.cnb ? ((n / cm.cnb) * cm.cP + cm.coff) * cm.cnb + n % cm.cnb : n;
}
// number of column tiles (width BN, BN | cnb) whose first column has gcol <= m_max
template <int BN>
__host__ __device__ inline int cols_upto(const ColMap &cm, long long m_max) {
  if (m_max < 0) return 0;
  if (!cm.cnb) return (int)((unsigned)m_max / (unsigned)BN) + 1;  // rows fit 32 bits
  const long long q = m_max / cm.cnb;  // global block column that holds m_max
  if (q < cm.coff) return 0;
  const long long lc = (q - cm.coff) / cm.cP;  // last local block whose global index is <= q
  const int tpb = cm.cnb / BN;
  int cnt = (int)lc * tpb;  // local blocks 0..lc-1 are entirely reachable
  if (lc * cm.cP + cm.coff == q) cnt += (int)((m_max - q * cm.cnb) / BN) + 1;
  else cnt += tpb;
  return cnt;
}
template <int BM, int BN>
__host__ __device__ inline int sr_ncols(int GROUP, int cuplo, int NJ, const ColMap &cm, int s) {
  if (cuplo == 1) {
    const int c = cols_upto<BN>(cm, (long long)GROUP * (s + 1) * BM - 1);
    return c < NJ ? c : NJ;
  }
  if (cuplo == 2) {
    const int b = (BM / BN) * GROUP * s;
    return b < NJ ? NJ - b : 0;
  }
  return NJ;
}
template <int BM, int BN>
__host__ inline long long total_tiles(int GROUP, int cuplo, int NI, int NJ, const ColMap &cm) {
  long long t = 0;
  for (int s = 0; s * GROUP < NI; ++s) t += (long long)GROUP * sr_ncols<BM, BN>(GROUP, cuplo, NJ, cm, s);
  return t;
}
template <int BM, int BN>
__host__ __device__ inline bool tile_coords(int GROUP, int cuplo, int NI, int NJ, const ColMap &cm, long long t, int &bi, int &bj) {
  int s = 0;
  if (cuplo == 0) {
    s = (int)(t / ((long long)GROUP * NJ));
    t -= (long long)s * GROUP * NJ;
  } else {
    for (;;) {
      long long c = (long long)GROUP * sr_ncols<BM, BN>(GROUP, cuplo, NJ, cm, s);
      if (t < c) break;
      t -= c;
      ++s;
    }
  }
  const int jbase = (cuplo == 2) ? (BM / BN) * GROUP * s : 0;
  bj = jbase + (int)(t / GROUP);
  bi = GROUP * s + (int)(t % GROUP);
  if (bi >= NI || bj >= NJ) return false;
  if (cuplo == 1 && bj >= cols_upto<BN>(cm, (long long)(bi + 1) * BM - 1)) return false;  // above the diagonal
  if (cuplo == 2 && bj < (BM / BN) * bi) return false;                                     // below the diagonal
  return true;
}

// (GROUP = row-blocks per super-row is a run-time choice of the launcher: 16 keeps 2048 rows of A in L2
// while B streams by once - fewer DRAM re-reads on big grids; short, wide problems use 8 so that a
// super-row is not mostly empty.)
// Host-built prefix table of the super-row tile counts (passed in the kernel arguments): the CTA's
// tile decode becomes a short scan over warp-uniform constants instead of a chain of divisions.
constexpr int SCHED_MAX_SR = 160;
struct SchedTable {
  int nsr;                      // number of super-rows; 0 = table not usable, decode arithmetically
  int start[SCHED_MAX_SR + 1];  // first tile id of super-row s; start[nsr] = grid size
};
template <int BM, int BN>
__host__ inline void build_sched_table(SchedTable &tb, int GROUP, int cuplo, int NI, int NJ, const ColMap &cm) {
  const int nsr = (NI + GROUP - 1) / GROUP;
  tb.nsr = 0;
  if (nsr > SCHED_MAX_SR) return;
  long long t = 0;
  for (int s = 0; s < nsr; ++s) {
    tb.start[s] = (int)t;
    t += (long long)GROUP * sr_ncols<BM, BN>(GROUP, cuplo, NJ, cm, s);
    if (t > 0x7fffffffLL) return;
  }
  tb.start[nsr] = (int)t;
  tb.nsr = nsr;
}
template <int BM, int BN>
__device__ __forceinline__ bool tile_coords_tb(const SchedTable &tb, int GROUP, int cuplo, int NI, int NJ, const ColMap &cm,
                                               long long t, int &bi, int &bj) {
  if (tb.nsr == 0) return tile_coords<BM, BN>(GROUP, cuplo, NI, NJ, cm, t, bi, bj);
  const int ti = (int)t;
  int s = 0;
  // start[] is increasing: scan in strides of 8, then by 1 (at most ~28 uniform constant loads)
  while (s + 8 <= tb.nsr && tb.start[s + 8] <= ti) s += 8;
  while (s + 1 < tb.nsr && tb.start[s + 1] <= ti) ++s;
  const int local = ti - tb.start[s];
  const int jbase = (cuplo == 2) ? (BM / BN) * GROUP * s : 0;
  bj = jbase + local / GROUP;
  bi = GROUP * s + local % GROUP;
  if (bi >= NI || bj >= NJ) return false;
  if (cuplo == 1 && bj >= cols_upto<BN>(cm, (long long)(bi + 1) * BM - 1)) return false;
  if (cuplo == 2 && bj < (BM / BN) * bi) return false;
  return true;
}

// element (r, c) of a stored operand: bounds, triangular mask, implicit unit diagonal, conjugation
template <typename T>
__device__ __forceinline__ T fetch_stored(const T *__restrict__ P, long long ld, bool in, long long r, long long c,
                                          int tri, int unit, bool conj) {
  if (!in) return zero_of<T>();
  if (tri == 1 && r < c) return zero_of<T>();
  if (tri == 2 && r > c) return zero_of<T>();
  if (unit && tri && r == c) return one_of<T>();
  const T v = P[c * ld + r];
  return conj ? conj_of(v) : v;
}


static inline int op_code(int c) { return c == 'N' ? 0 : (c == 'T' ? 1 : 2); }
static inline int tri_code(int c) { return c == 'L' ? 1 : (c == 'U' ? 2 : 0); }

}  // namespace chol
