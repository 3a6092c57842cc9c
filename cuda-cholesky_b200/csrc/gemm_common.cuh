// gemm_common.cuh - pieces shared by the FP64 DMMA engine (gemm_dmma.cu) and the FP32 FFMA engine
// (gemm_ffma.cu): the triangular super-row tile scheduler and the generic operand fetch.
#pragma once
#include <cuda_runtime.h>
#include "scalar_ops.cuh"

namespace chol {

// ---- tile scheduler ---------------------------------------------------------------------------
template <int R, int GROUP>
__host__ __device__ inline int sr_ncols(int cuplo, int NJ, int s) {
  if (cuplo == 1) {
    int c = R * GROUP * (s + 1);
    return c < NJ ? c : NJ;
  }
  if (cuplo == 2) {
    int b = R * GROUP * s;
    return b < NJ ? NJ - b : 0;
  }
  return NJ;
}
template <int R, int GROUP>
__host__ inline long long total_tiles(int cuplo, int NI, int NJ) {
  long long t = 0;
  for (int s = 0; s * GROUP < NI; ++s) t += (long long)GROUP * sr_ncols<R, GROUP>(cuplo, NJ, s);
  return t;
}
template <int R, int GROUP>
__device__ inline bool tile_coords(int cuplo, int NI, int NJ, long long t, int &bi, int &bj) {
  int s = 0;
  if (cuplo == 0) {
    s = (int)(t / ((long long)GROUP * NJ));
    t -= (long long)s * GROUP * NJ;
  } else {
    for (;;) {
      long long c = (long long)GROUP * sr_ncols<R, GROUP>(cuplo, NJ, s);
      if (t < c) break;
      t -= c;
      ++s;
    }
  }
  const int jbase = (cuplo == 2) ? R * GROUP * s : 0;
  bj = jbase + (int)(t / GROUP);
  bi = GROUP * s + (int)(t % GROUP);
  if (bi >= NI || bj >= NJ) return false;
  if (cuplo == 1 && bj > R * bi + R - 1) return false;  // tile entirely above the diagonal
  if (cuplo == 2 && bj < R * bi) return false;          // tile entirely below the diagonal
  return true;
}

// ---- generic operand fetch (MODE 1) -----------------------------------------------------------
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
