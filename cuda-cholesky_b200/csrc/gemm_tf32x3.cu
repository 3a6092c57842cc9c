// gemm_tf32x3.cu — the s-precision trailing update on Blackwell's 5th-generation tensor cores.
//
//   D = alpha * A * B^T + beta * C        A: m x k, B: n x k, FP32 in, FP32 out
//
// north_star allows "FP32 FFMA or a 3xTF32 split with a stated tolerance" for s/c; this is the split
// (it replaces the SIMT kernel class of the reference's blas/sgemm.cu:29 for the contraction that
// dominates spotrf, lapack/spotrf.c:253-261).  Every FP32 operand x is written once, by
// tf32_split_kernel, as  hi = tf32(x), lo = tf32(x - hi)  (round-to-nearest, 11 + 11 significant
// bits) into K-major planes, and the product is accumulated in FP32 as
//        hi*hi + hi*lo + lo*hi          (the lo*lo term, <= 2^-22 relative, is dropped)
// by `tcgen05.mma.kind::tf32` into TMEM.  One CTA per SM, persistent over the triangle's tiles:
//   warp 0     TMA producer: four `cp.async.bulk.tensor.2d` boxes per stage (A hi/lo 128 x 32,
//              B hi/lo 256 x 32, 128-byte swizzle), completion counted on the stage's mbarrier
//   warp 1     MMA issuer: one thread, 12 `tcgen05.mma` (128 x 256 x 8) per stage, `tcgen05.commit`
//              releases the stage to the producer and, after the last k step, the accumulator to
//              the epilogue
//   warp 2     TMEM allocation (512 columns = two 128 x 256 FP32 accumulators, double-buffered so
//              the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 4-11 epilogue: `tcgen05.ld` 32 lanes x 32 columns -> registers -> masked
//              read-modify-write of the caller's column-major C (lanes run along rows: every
//              column access of a warp is one 128-byte line)
// Stated tolerance of the split: each product carries a relative error <= 2^-21 (vs 2^-24 for an
// FP32 FMA); tests use eps_s/c(tensor) = 2^-21 in the same bounds as the FP32 engine.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "chol_kernels.h"
#include "launch_count.h"

namespace chol {

constexpr int TBM = 128, TBN = 256, TBK = 32;   // CTA tile; TBK * 4 B = 128 B = one swizzle atom
constexpr int TST = 2;                          // smem stages
constexpr int T_A_BYTES = TBM * TBK * 4;        // 16 KB per plane
constexpr int T_B_BYTES = TBN * TBK * 4;        // 32 KB per plane
constexpr int T_STAGE_BYTES = 2 * T_A_BYTES + 2 * T_B_BYTES;  // 96 KB
constexpr int T_SMEM_BYTES = TST * T_STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int T_THREADS = 384;                  // 12 warps
constexpr int T_EPI_THREADS = 256;

struct TArgs {
  int M, N, nk;      // nk = k steps of TBK
  int NI, NJ;        // row / column tiles
  int cuplo;         // 0 all, 1 lower, 2 upper
  int rs;            // row shift of the triangle test: D's row r stands for matrix row r >> rs (1 = interleaved complex rows)
  int herm;          // rs == 1: force Im(diag) = 0 (HERK)
  int ntiles;
  float alpha, beta;
  const float *C;
  long long ldc;
  float *D;
  long long ldd;
};

// ---- PTX wrappers -------------------------------------------------------------------------------
// (tests/emu/tcgen05_emu.h supplies CPU stand-ins with the same names when the kernel's source is run by the
// emulation harness tests/emu/emu_tf32x3.cpp; under nvcc the code below is what is compiled, unchanged)
#ifdef CHOL_CPU_EMU
}  // namespace chol
#include "tcgen05_emu.h"
namespace chol {
#else
#define T_DYN_SMEM(name) extern __shared__ uint8_t name[]
#define T_FENCE_MBAR_INIT() asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory")
#define T_TMEM_ALLOC(slot, cols)                                                                                                    \
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(slot)), "r"(cols) : "memory");       \
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory")
#define T_TMEM_DEALLOC(base, cols) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory")
__device__ __forceinline__ uint32_t s_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void t_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void t_mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void t_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t_mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "TW_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra TW_DONE;\n"
      "bra TW_LOOP;\n"
      "TW_DONE:\n"
      "}" ::"r"(s_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(s_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(s_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D(tmem) (+)= A(smem) * B(smem)^T, 128 x TBN x 8, TF32 in, FP32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on `bar` when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  // the wait names the registers so that no use of them can be scheduled above it
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]) :: "memory");
}

#endif  // CHOL_CPU_EMU

// K-major operand tile, rows of 128 bytes, 128-byte swizzle (what the TMA box writes): 8-row groups
// are 1024 bytes apart (SBO), the leading offset is unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) /*sm100 descriptor version*/ |
         (2ull << 61) /*SWIZZLE_128B*/;
}
// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = TBN
constexpr uint32_t T_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);

// tile t of the (triangle's) column-major tile list -> (bi, bj)
__device__ __forceinline__ void t_tile_coords(const TArgs &p, int t, int &bi, int &bj) {
  if (p.cuplo == 0) {
    bj = t / p.NI;
    bi = t - bj * p.NI;
    return;
  }
  const int R = (TBN << p.rs) / TBM;  // row tiles per column tile
  for (bj = 0;; ++bj) {
    const int lo = (p.cuplo == 1) ? min(R * bj, p.NI) : 0;
    const int hi = (p.cuplo == 1) ? p.NI : min(p.NI, R * bj + R);
    const int cnt = hi - lo;
    if (t < cnt) {
      bi = lo + t;
      return;
    }
    t -= cnt;
  }
}
static int t_count_tiles(int cuplo, int rs, int NI, int NJ) {
  if (cuplo == 0) return NI * NJ;
  const int R = (TBN << rs) / TBM;
  int tot = 0;
  for (int bj = 0; bj < NJ; ++bj) {
    const int lo = (cuplo == 1) ? (R * bj < NI ? R * bj : NI) : 0;
    const int hi = (cuplo == 1) ? NI : (NI < R * bj + R ? NI : R * bj + R);
    tot += hi - lo;
  }
  return tot;
}

__global__ void __launch_bounds__(T_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
                   const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo,
                   const __grid_constant__ TArgs p) {
  T_DYN_SMEM(t_smem_raw);
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(t_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TST * T_STAGE_BYTES);
  uint64_t *full = bars, *empty = bars + TST, *acc_full = bars + 2 * TST, *acc_empty = bars + 2 * TST + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * TST + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TST; ++s) {
      t_mbar_init(&full[s], 1);
      t_mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      t_mbar_init(&acc_full[b], 1);
      t_mbar_init(&acc_empty[b], T_EPI_THREADS);
    }
    T_FENCE_MBAR_INIT();
  }
  if (warp == 2) {
    T_TMEM_ALLOC(tmem_slot, 512u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t g = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        int bi, bj;
        t_tile_coords(p, t, bi, bj);
        const int m0 = bi * TBM, n0 = bj * TBN;
        for (int ks = 0; ks < p.nk; ++ks, ++g) {
          const uint32_t s = g % TST, ph = (g / TST) & 1u;
          t_mbar_wait(&empty[s], ph ^ 1u);
          uint8_t *st = smem + s * T_STAGE_BYTES;
          t_mbar_expect_tx(&full[s], T_STAGE_BYTES);
          tma_load_2d(st, &mapAhi, ks * TBK, m0, &full[s]);
          tma_load_2d(st + T_A_BYTES, &mapAlo, ks * TBK, m0, &full[s]);
          tma_load_2d(st + 2 * T_A_BYTES, &mapBhi, ks * TBK, n0, &full[s]);
          tma_load_2d(st + 2 * T_A_BYTES + T_B_BYTES, &mapBlo, ks * TBK, n0, &full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t g = 0, it = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
        const uint32_t ab = it & 1u, aph = (it >> 1) & 1u;
        t_mbar_wait(&acc_empty[ab], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * TBN;
        for (int ks = 0; ks < p.nk; ++ks, ++g) {
          const uint32_t s = g % TST, ph = (g / TST) & 1u;
          t_mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = s_u32(smem + s * T_STAGE_BYTES);
          const uint64_t dAhi = smem_desc_sw128(sa), dAlo = smem_desc_sw128(sa + T_A_BYTES);
          const uint64_t dBhi = smem_desc_sw128(sa + 2 * T_A_BYTES), dBlo = smem_desc_sw128(sa + 2 * T_A_BYTES + T_B_BYTES);
#pragma unroll
          for (int kk = 0; kk < TBK / 8; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 32 >> 4);  // 8 tf32 = 32 bytes along K inside the swizzle atom
            umma_tf32(d_tmem, dAlo + adv, dBhi + adv, T_IDESC, (ks | kk) != 0);
            umma_tf32(d_tmem, dAhi + adv, dBlo + adv, T_IDESC, 1u);
            umma_tf32(d_tmem, dAhi + adv, dBhi + adv, T_IDESC, 1u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[ab]);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp & 3, half = (warp - 4) >> 2;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
      int bi, bj;
      t_tile_coords(p, t, bi, bj);
      const uint32_t ab = it & 1u, aph = (it >> 1) & 1u;
      t_mbar_wait(&acc_full[ab], aph);
      tc_fence_after();
      const int rw0 = bi * TBM + 32 * q;       // first row of this warp
      const long long row = rw0 + lane;
#pragma unroll 1
      for (int c = 0; c < (TBN / 2) / 32; ++c) {
        const int col0 = bj * TBN + half * (TBN / 2) + c * 32;
        // warp-uniform: anything of this 32 x 32 patch inside the matrix and the triangle?
        bool any = rw0 < p.M && col0 < p.N;
        if (p.cuplo == 1) any = any && (((rw0 + 31) >> p.rs) >= col0);
        if (p.cuplo == 2) any = any && ((rw0 >> p.rs) <= col0 + 31);
        if (!any) continue;
        float acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + ab * TBN + half * (TBN / 2) + c * 32, acc);
        float cv[32];
        const long long mrow = row >> p.rs;  // the matrix row this (real) row belongs to
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const long long col = col0 + j;
          bool keep = row < p.M && col < p.N;
          if (p.cuplo == 1) keep = keep && mrow >= col;
          if (p.cuplo == 2) keep = keep && mrow <= col;
          cv[j] = (keep && p.beta != 0.f) ? p.C[col * p.ldc + row] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const long long col = col0 + j;
          bool keep = row < p.M && col < p.N;
          if (p.cuplo == 1) keep = keep && mrow >= col;
          if (p.cuplo == 2) keep = keep && mrow <= col;
          if (keep) {
            float o = p.alpha * acc[j] + p.beta * cv[j];
            if (p.herm && mrow == col && (row & 1)) o = 0.f;  // Im of a Hermitian diagonal (blas/cherk.c: the diagonal is real)
            p.D[col * p.ldd + row] = o;
          }
        }
      }
      tc_fence_before();
      t_mbar_arrive(&acc_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    T_TMEM_DEALLOC(tmem_base, 512u);
  }
}

// ---- operand split: FP32 -> (hi, lo) TF32 planes, K-major ---------------------------------------
// hi[r * kp + c], lo[...] for r < rows, c < kp (zero for c >= k).  trans == 0: source element (r, c) at
// A[c * lda + r] (the stored operand is rows x k, column-major - the "N" operand of an NT product);
// trans != 0: at A[r * lda + c].
#ifndef CHOL_CPU_EMU  // (emulation: tests/emu/tcgen05_emu.h)
__device__ __forceinline__ void tf32_split(float x, float &hi, float &lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
  lo = __uint_as_float(l);
}
#endif
__global__ void __launch_bounds__(256) tf32_split_kernel(int trans, int rows, int k, int kp, const float *__restrict__ A, long long lda,
                                                         float *__restrict__ hi, float *__restrict__ lo) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (!trans) {
    for (int cc = ty; cc < 32; cc += 8) {
      const int r = r0 + tx, c = c0 + cc;
      tile[cc][tx] = (r < rows && c < k) ? A[(long long)c * lda + r] : 0.f;
    }
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) {
      const int r = r0 + rr, c = c0 + tx;
      if (r < rows && c < kp) {
        float h, l;
        tf32_split(tile[tx][rr], h, l);
        hi[(long long)r * kp + c] = h;
        lo[(long long)r * kp + c] = l;
      }
    }
  } else {
    for (int rr = ty; rr < 32; rr += 8) {
      const int r = r0 + rr, c = c0 + tx;
      if (r < rows && c < kp) {
        float h, l;
        tf32_split(c < k ? A[(long long)r * lda + c] : 0.f, h, l);
        hi[(long long)r * kp + c] = h;
        lo[(long long)r * kp + c] = l;
      }
    }
  }
}

// complex operand W (rows x k, column-major, interleaved re/im) -> the REAL planes of the embedding
//   Re(W W^H)(i,j) = [a_i | b_i] . [a_j | b_j],   Im(W W^H)(i,j) = [b_i | -a_i] . [a_j | b_j]        (W = a + i b)
// plane row 2i   = (a_i0, b_i0, a_i1, b_i1, ...)   - also the B operand's row i (read with twice the row pitch)
// plane row 2i+1 = (b_i0, -a_i0, b_i1, -a_i1, ...)
// so that ONE real product  D'(2m x n) = A'(2m x 2k) * B'(n x 2k)^T  writes the interleaved complex matrix in place.
__global__ void __launch_bounds__(256) tf32_split_c_kernel(int rows, int k, int kp, const float2 *__restrict__ W, long long ldw,
                                                           float *__restrict__ hi, float *__restrict__ lo) {
  __shared__ float2 tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;  // c in complex columns
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int cc = ty; cc < 32; cc += 8) {
    const int r = r0 + tx, c = c0 + cc;
    tile[cc][tx] = (r < rows && c < k) ? W[(long long)c * ldw + r] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int r = r0 + rr, c = c0 + tx;
    if (r < rows && 2 * c < kp) {
      const float2 w = tile[tx][rr];
      float ah, al, bh, bl;
      tf32_split(w.x, ah, al);
      tf32_split(w.y, bh, bl);
      const long long e = (long long)(2 * r) * kp + 2 * c, o = e + kp;
      *reinterpret_cast<float2 *>(hi + e) = make_float2(ah, bh);
      *reinterpret_cast<float2 *>(lo + e) = make_float2(al, bl);
      *reinterpret_cast<float2 *>(hi + o) = make_float2(bh, -ah);
      *reinterpret_cast<float2 *>(lo + o) = make_float2(bl, -al);
    }
  }
}

#ifndef CHOL_CPU_EMU  // the launchers are not part of the CPU emulation (tests/emu restates their argument set-up)
static int make_map(CUtensorMap *map, const void *base, size_t rows, size_t kp, size_t row_pitch, unsigned box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)row_pitch * 4};
  cuuint32_t box[2] = {(cuuint32_t)TBK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), gdim, gstr, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

#endif  // CHOL_CPU_EMU

}  // namespace chol

#ifndef CHOL_CPU_EMU
using namespace chol;

extern "C" size_t chol_tf32_kpitch(size_t k) { return (k + TBK - 1) / TBK * TBK; }

extern "C" int chol_tf32_split_launch(int trans, size_t rows, size_t k, const void *A, size_t lda, void *hi, void *lo, size_t kp,
                                      void *stream) {
  if (rows == 0 || kp == 0) return 0;
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)(kp / 32));
  CHOL_COUNT_LAUNCH();
  tf32_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(trans, (int)rows, (int)k, (int)kp, (const float *)A, (long long)lda,
                                                             (float *)hi, (float *)lo);
  return (int)cudaGetLastError();
}

extern "C" int chol_tf32_split_c_launch(size_t rows, size_t k, const void *W, size_t ldw, void *hi, void *lo, size_t kp, void *stream) {
  if (rows == 0 || kp == 0) return 0;
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((kp / 2 + 31) / 32));
  CHOL_COUNT_LAUNCH();
  tf32_split_c_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((int)rows, (int)k, (int)kp, (const float2 *)W, (long long)ldw, (float *)hi,
                                                               (float *)lo);
  return (int)cudaGetLastError();
}

// general form: separate row pitches of the A and B planes, interleaved-complex rows (rowshift = 1) and the Hermitian diagonal
static int tf32x3_launch(size_t m, size_t n, size_t k, double alpha, const void *Ahi, const void *Alo, size_t pitchA, const void *Bhi,
                         const void *Blo, size_t pitchB, size_t kp, double beta, const void *C, size_t ldc, void *D, size_t ldd,
                         int cuplo, int rowshift, int herm, void *stream) {
  if (m == 0 || n == 0) return 0;
  int sm_count = 0, dev = 0;  // per call: the library serves several devices / contexts from one process
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  TArgs p;
  p.M = (int)m;
  p.N = (int)n;
  p.nk = (int)((k + TBK - 1) / TBK);
  p.NI = (int)((m + TBM - 1) / TBM);
  p.NJ = (int)((n + TBN - 1) / TBN);
  p.cuplo = cuplo == 'L' ? 1 : (cuplo == 'U' ? 2 : 0);
  p.rs = rowshift;
  p.herm = herm;
  p.ntiles = t_count_tiles(p.cuplo, p.rs, p.NI, p.NJ);
  p.alpha = (float)alpha;
  p.beta = (float)beta;
  p.C = (const float *)C;
  p.ldc = (long long)ldc;
  p.D = (float *)D;
  p.ldd = (long long)ldd;
  if (p.ntiles == 0 || p.nk == 0) return 0;
  CUtensorMap mAhi, mAlo, mBhi, mBlo;
  if (make_map(&mAhi, Ahi, m, kp, pitchA, TBM) || make_map(&mAlo, Alo, m, kp, pitchA, TBM) || make_map(&mBhi, Bhi, n, kp, pitchB, TBN) ||
      make_map(&mBlo, Blo, n, kp, pitchB, TBN))
    return (int)cudaErrorInvalidValue;
  const int grid = p.ntiles < sm_count ? p.ntiles : sm_count;
  CHOL_COUNT_LAUNCH();
  gemm_tf32x3_kernel<<<grid, T_THREADS, T_SMEM_BYTES, (cudaStream_t)stream>>>(mAhi, mAlo, mBhi, mBlo, p);
  return (int)cudaGetLastError();
}

extern "C" int chol_tf32x3_gemm_launch(size_t m, size_t n, size_t k, double alpha, const void *Ahi, const void *Alo, const void *Bhi,
                                       const void *Blo, size_t kp, double beta, const void *C, size_t ldc, void *D, size_t ldd,
                                       int cuplo, void *stream) {
  return tf32x3_launch(m, n, k, alpha, Ahi, Alo, kp, Bhi, Blo, kp, kp, beta, C, ldc, D, ldd, cuplo, 0, 0, stream);
}

// D (m x n complex, interleaved, ldd in complex elements) = alpha * A * B^H + beta * C with real alpha, beta; A and B are
// row ranges of planes written by chol_tf32_split_c_launch (Ahi/Alo point at plane row 2*i0 of A's first row i0, Bhi/Blo
// at plane row 2*j0 of B's first row j0); cuplo on complex indices; herm != 0 zeroes Im of the diagonal (HERK).
extern "C" int chol_tf32x3_herk_launch(size_t m, size_t n, size_t k, double alpha, const void *Ahi, const void *Alo, const void *Bhi,
                                       const void *Blo, size_t kp, double beta, const void *C, size_t ldc, void *D, size_t ldd,
                                       int cuplo, int herm, void *stream) {
  return tf32x3_launch(2 * m, n, 2 * k, alpha, Ahi, Alo, kp, Bhi, Blo, 2 * kp, kp, beta, C, 2 * ldc, D, 2 * ldd, cuplo, 1, herm, stream);
}
#endif  // CHOL_CPU_EMU
