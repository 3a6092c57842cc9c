// gemm_ffma.cuh — configuration, argument block and operand loader of the FP32 engine, shared by
// gemm_ffma.cu (the measured kernel) and gemm_ffma_deep.cu (its deeper-k-tile variant).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "chol_kernels.h"
#include "emu_compat.cuh"
#include "launch_count.h"
#include "gemm_common.cuh"

namespace chol {

template <typename T>
struct FCfg {
  static constexpr bool CPLX = ScalarOf<T>::cplx;
  static constexpr int VW = 16 / sizeof(T);  // elements per 128-bit access: 4 (float) / 2 (float2)
  static constexpr int TM = 2 * VW, TN = 2 * VW;
  static constexpr int BM = 16 * TM, BN = 16 * TN;  // 128 / 64
  static constexpr int BK = 8;
  static constexpr int LDS = BM + VW;  // padded pitch (elements); keeps 16-byte alignment
  static constexpr int THREADS = 256;
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
#ifdef CHOL_CPU_EMU
  emu::cp_async(dst, src, BYTES);
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(src), "n"(BYTES) : "memory");
#endif
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

}  // namespace chol
