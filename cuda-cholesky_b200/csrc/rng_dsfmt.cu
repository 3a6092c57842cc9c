// rng_dsfmt.cu — dSFMT-19937 on the device, bit for bit the stream of the reference's CPU generator
// (rng/dsfmt.c:92-125 `generate`, constants rng/dsfmt_19937.c:1-15), produced in parallel.
// the input generator of the BASELINE configs ("A = B*B^T + N*I, dSFMT seed").
//
// The generator is a linear recurrence over GF(2) on a state of 191 128-bit words plus a 128-bit "lung":
//     lung <- P(lung) ^ (s[i] << 19) ^ s[(i + 117) mod 191]          (P reverses the four 32-bit words)
//     s[i] <- (lung >> 12) ^ (lung & MSK) ^ s[i]                       i = 0 .. 190 = one refill = 382 doubles
// Two facts make it parallel:
//  * inside a refill, step i needs s[(i+117) mod 191], which is either untouched so far (i < 74) or was
//    produced 74 steps earlier - so up to 74 consecutive steps have all their inputs, and the only serial
//    link is the lung.  P is an involution, so over an even number of steps the lung chain is a plain XOR
//    prefix: a warp takes 64 steps at once, lane l owns steps 2l and 2l+1, folds them into
//    c_l = P(x_2l) ^ x_2l+1 and one 5-step warp scan gives every lane its incoming lung.  A refill is three
//    such rounds (64 + 64 + 63 steps).
//  * across refills, the state after r refills is T^r s0 for a fixed linear map T.  With the minimal
//    polynomial p of T on s0 (host: Berlekamp-Massey, host/rng_poly.c), T^r s0 = sum_i a_r[i] T^i s0 with
//    a_r = x^r mod p: the states at ANY set of refill counts are GF(2) combinations of the first
//    D = deg p states (the "Krylov block" K, written once per seed by dsfmt_generate_kernel in snapshot mode).
//    dsfmt_jump_kernel forms those combinations for all segments of a request in one launch; then one
//    warp per segment generates its refills from there (dsfmt_generate_kernel) straight into the output.
// Output conversions follow rng/dsfmt.c:148-285 literally (getOpenOpen / OpenClose / CloseOpen / CloseClose);
// pairs of doubles come from one 128-bit word and an odd tail drops the high half, as there.
//
// Bound: HBM write bandwidth (8 B per double, nothing is read but the 3 KB start state per segment);
// per 1 KB of output a warp issues 24 shuffles and ~110 other instructions.
#include <cuda_runtime.h>
#include <stdint.h>

#include "chol_kernels.h"
#include "emu_compat.cuh"
#include "launch_count.h"

namespace chol {

constexpr int DS_N = 191;            // 128-bit words of the state (rng/dsfmt.c:4 with MEXP 19937)
constexpr int DS_POS1 = 117;         // rng/dsfmt_19937.c:3
constexpr int DS_SL1 = 19;           // rng/dsfmt_19937.c:4
constexpr int DS_SR = 12;            // rng/dsfmt.c:7
constexpr unsigned long long DS_MSK1 = 0x000ffafffffffb3fULL;  // rng/dsfmt_19937.c:5-6
constexpr unsigned long long DS_MSK2 = 0x000ffdfffc90fffdULL;
constexpr int DS_STATE_WORDS = DS_N + 1;        // with the lung
constexpr int DS_STATE_U32 = DS_STATE_WORDS * 4;  // 768
constexpr int GEN_WARPS = 4;                    // segments per CTA of the generation kernel
constexpr int JUMP_THREADS = 256;               // 768 / 256 = 3 state columns per thread
constexpr int JUMP_GT = 8;                      // segments per CTA of the jump kernel

struct __align__(16) W128 {
  unsigned long long lo, hi;
};

__device__ __forceinline__ unsigned long long rot32(unsigned long long v) { return (v >> 32) | (v << 32); }
// the lung shuffle _mm_shuffle_epi32(lung, 0x1b) of rng/dsfmt.c:101: 32-bit words (w0,w1,w2,w3) -> (w3,w2,w1,w0)
__device__ __forceinline__ W128 perm(W128 a) { return W128{rot32(a.hi), rot32(a.lo)}; }
__device__ __forceinline__ W128 operator^(W128 a, W128 b) { return W128{a.lo ^ b.lo, a.hi ^ b.hi}; }
// x_i = (s_i << SL1) ^ s_(i+POS1)
__device__ __forceinline__ W128 mix_in(W128 s, W128 b) { return W128{(s.lo << DS_SL1) ^ b.lo, (s.hi << DS_SL1) ^ b.hi}; }
// new s_i from the new lung
__device__ __forceinline__ W128 mix_out(W128 lung, W128 s) {
  return W128{(lung.lo >> DS_SR) ^ (lung.lo & DS_MSK1) ^ s.lo, (lung.hi >> DS_SR) ^ (lung.hi & DS_MSK2) ^ s.hi};
}
__device__ __forceinline__ W128 shfl_up_w(W128 v, int d) {
  W128 r;
  r.lo = __shfl_up_sync(0xffffffffu, v.lo, d);
  r.hi = __shfl_up_sync(0xffffffffu, v.hi, d);
  return r;
}
__device__ __forceinline__ W128 shfl_w(W128 v, int src) {
  W128 r;
  r.lo = __shfl_sync(0xffffffffu, v.lo, src);
  r.hi = __shfl_sync(0xffffffffu, v.hi, src);
  return r;
}

// One refill of the warp's state st[0..191] (st[191] = lung) in shared memory: three rounds of up to 64 steps.
// On return st holds the new state; the caller must __syncwarp() before other lanes' words are read.
__device__ __forceinline__ void dsfmt_refill_warp(W128 *st, int lane) {
  W128 lung = st[DS_N];  // every lane keeps the running lung
#pragma unroll 1
  for (int c0 = 0; c0 < DS_N; c0 += 64) {
    const int i0 = c0 + 2 * lane, i1 = i0 + 1;
    const bool has0 = i0 < DS_N, has1 = i1 < DS_N;  // only the last lane of the last round lacks its second step
    W128 s0 = W128{0, 0}, s1 = s0, x0 = s0, x1 = s0;
    if (has0) {
      s0 = st[i0];
      int b = i0 + DS_POS1;
      if (b >= DS_N) b -= DS_N;
      x0 = mix_in(s0, st[b]);
    }
    if (has1) {
      s1 = st[i1];
      int b = i1 + DS_POS1;
      if (b >= DS_N) b -= DS_N;
      x1 = mix_in(s1, st[b]);
    }
    // two steps leave the lung as lung ^ P(x0) ^ x1; a lane without a second step is the last of the refill and its
    // aggregate is not used by anyone
    const W128 c = perm(x0) ^ x1;
    W128 incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const W128 up = shfl_up_w(incl, d);
      if (lane >= d) incl = incl ^ up;
    }
    const W128 lin = lung ^ incl ^ c;  // exclusive prefix: the lung entering step i0
    const W128 y0 = perm(lin) ^ x0;    // lung after step i0
    const W128 y1 = perm(y0) ^ x1;     // lung after step i1
    if (has0) st[i0] = mix_out(y0, s0);
    if (has1) st[i1] = mix_out(y1, s1);
    lung = shfl_w(has1 ? y1 : y0, 31);
    __syncwarp();  // the next round reads words this round wrote (s[i-74]) from other lanes
  }
  if (lane == 0) st[DS_N] = lung;
}

// output kinds (include/rng.h): what is stored for a 64-bit half `w` of a state word, a double in [1, 2)
enum { DS_RAW = 0, DS_CLOSE_OPEN = 1, DS_OPEN_CLOSE = 2, DS_OPEN_OPEN = 3, DS_CLOSE_CLOSE = 4 };

template <int KIND>
__device__ __forceinline__ unsigned long long ds_convert(unsigned long long w) {
  if (KIND == DS_RAW) return w;
  double d;
  if (KIND == DS_CLOSE_OPEN) d = __longlong_as_double((long long)w) - 1.0;                 // rng/dsfmt.c:231
  else if (KIND == DS_OPEN_CLOSE) d = 2.0 - __longlong_as_double((long long)w);            // rng/dsfmt.c:196
  else if (KIND == DS_OPEN_OPEN) d = __longlong_as_double((long long)(w | 1ULL)) - 1.0;    // rng/dsfmt.c:160-161
  else d = __dmul_rn((double)(w >> 10), 1.0 / 9007199254740991.0);                         // rng/dsfmt.c:266 (shift by 10, as there)
  return (unsigned long long)__double_as_longlong(d);
}

// One warp per segment.  Segment g starts from states[g] = T^(refill0 + g * seg_refills) s0 and produces refills
// refill0 + g * seg_refills + {0 .. seg_refills - 1}; refill r holds stream words 191 r .. 191 r + 190.  Words inside
// [w_begin, w_end) are converted and stored: word t -> elements 2 (t - w_begin), 2 (t - w_begin) + 1 of `out`
// (the last one only if it is < n_out).  snap != NULL (Krylov mode, one segment): the raw state after refill r is
// stored at snap[(r + 1) * 192 ...] instead, nothing goes to `out`.
template <int KIND>
__global__ void __launch_bounds__(GEN_WARPS * 32)
dsfmt_generate_kernel(const W128 *__restrict__ states, unsigned long long refill0, unsigned seg_refills, int nseg,
                      unsigned long long w_begin, unsigned long long w_end, unsigned long long *__restrict__ out,
                      unsigned long long n_out, W128 *__restrict__ snap) {
  __shared__ W128 smem[GEN_WARPS][DS_STATE_WORDS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int g = blockIdx.x * GEN_WARPS + wid;
  if (g >= nseg) return;  // whole warps leave; no block-wide barrier below
  W128 *st = smem[wid];
  for (int i = lane; i < DS_STATE_WORDS; i += 32) st[i] = states[(size_t)g * DS_STATE_WORDS + i];
  __syncwarp();
  const unsigned long long r_begin = refill0 + (unsigned long long)g * seg_refills;
  const bool aligned16 = (((size_t)out) & 15) == 0;
  for (unsigned k = 0; k < seg_refills; ++k) {
    const unsigned long long r = r_begin + k;
    const unsigned long long t0 = r * DS_N;  // first stream word of this refill
    if (snap == nullptr && t0 >= w_end) break;
    dsfmt_refill_warp(st, lane);
    __syncwarp();
    if (snap != nullptr) {
      W128 *dst = snap + (size_t)(r + 1) * DS_STATE_WORDS;
      for (int i = lane; i < DS_STATE_WORDS; i += 32) dst[i] = st[i];
      __syncwarp();  // the next refill overwrites st
      continue;
    }
    if (t0 + DS_N <= w_begin) continue;  // before the window: the state only moves on
#pragma unroll 1
    for (int i = lane; i < DS_N; i += 32) {
      const unsigned long long t = t0 + i;
      if (t < w_begin || t >= w_end) continue;
      const W128 v = st[i];
      const unsigned long long e = 2 * (t - w_begin);
      const unsigned long long a = ds_convert<KIND>(v.lo), b = ds_convert<KIND>(v.hi);
      if (e + 1 < n_out) {
        if (aligned16) {
          *reinterpret_cast<ulonglong2 *>(out + e) = make_ulonglong2(a, b);
        } else {
          out[e] = a;
          out[e + 1] = b;
        }
      } else if (e < n_out) {
        out[e] = a;  // odd tail: the high half of the last word is dropped (rng/dsfmt.c:245-251)
      }
    }
    __syncwarp();  // stores read st; the next refill overwrites it
  }
}

// states[g][:] ^= XOR over i in this CTA's slice of [0, D) with bit i of polys[g] set of K[i][:]
// (K: D states of 768 32-bit words; polys: nseg rows of poly_u32 words; states zeroed by the caller).
// grid.x = segment tiles of JUMP_GT, grid.y = slices of the i range; a thread owns 3 of the 768 columns.
__global__ void __launch_bounds__(JUMP_THREADS)
dsfmt_jump_kernel(const uint32_t *__restrict__ K, const uint32_t *__restrict__ polys, int poly_u32, int D,
                  uint32_t *__restrict__ states, int nseg) {
  __shared__ uint32_t bits[JUMP_GT][32];  // one 1024-bit stretch of every segment's polynomial at a time
  const int t = threadIdx.x;
  const int g0 = blockIdx.x * JUMP_GT;
  const int words = (D + 31) / 32;
  const int per = (words + gridDim.y - 1) / gridDim.y;
  const int w_lo = blockIdx.y * per, w_hi = min(words, w_lo + per);
  uint32_t acc[JUMP_GT][3];
#pragma unroll
  for (int g = 0; g < JUMP_GT; ++g) acc[g][0] = acc[g][1] = acc[g][2] = 0u;
  for (int wb = w_lo; wb < w_hi; wb += 32) {
    __syncthreads();
    {
      const int g = t >> 5, w = wb + (t & 31);  // 8 x 32 = 256 threads
      bits[g][t & 31] = (g0 + g < nseg && w < w_hi && w < poly_u32) ? polys[(size_t)(g0 + g) * poly_u32 + w] : 0u;
    }
    __syncthreads();
    const int wn = min(32, w_hi - wb);
    for (int wi = 0; wi < wn; ++wi) {
      uint32_t a[JUMP_GT];
      uint32_t any = 0u;
#pragma unroll
      for (int g = 0; g < JUMP_GT; ++g) {
        a[g] = bits[g][wi];
        any |= a[g];
      }
      if (any == 0u) continue;  // uniform across the CTA
      const int ibase = (wb + wi) * 32;
      const int nb = min(32, D - ibase);
#pragma unroll 4
      for (int ib = 0; ib < nb; ++ib) {
        const uint32_t *row = K + (size_t)(ibase + ib) * DS_STATE_U32;
        const uint32_t k0 = row[t], k1 = row[t + JUMP_THREADS], k2 = row[t + 2 * JUMP_THREADS];
#pragma unroll
        for (int g = 0; g < JUMP_GT; ++g) {
          const uint32_t m = 0u - ((a[g] >> ib) & 1u);
          acc[g][0] ^= k0 & m;
          acc[g][1] ^= k1 & m;
          acc[g][2] ^= k2 & m;
        }
      }
    }
  }
#pragma unroll
  for (int g = 0; g < JUMP_GT; ++g) {
    if (g0 + g >= nseg) break;
    uint32_t *dst = states + (size_t)(g0 + g) * DS_STATE_U32;
    if (gridDim.y == 1) {
      dst[t] = acc[g][0];
      dst[t + JUMP_THREADS] = acc[g][1];
      dst[t + 2 * JUMP_THREADS] = acc[g][2];
    } else {
      if (acc[g][0]) atomicXor(dst + t, acc[g][0]);
      if (acc[g][1]) atomicXor(dst + t + JUMP_THREADS, acc[g][1]);
      if (acc[g][2]) atomicXor(dst + t + 2 * JUMP_THREADS, acc[g][2]);
    }
  }
}

}  // namespace chol

#ifndef CHOL_CPU_EMU  // the launchers are not part of the CPU emulation (tests/emu)
using namespace chol;

extern "C" int chol_dsfmt_generate_launch(const void *states, unsigned long long refill0, unsigned seg_refills, int nseg,
                                          unsigned long long w_begin, unsigned long long w_end, void *out,
                                          unsigned long long n_out, int kind, void *snap, void *stream) {
  if (nseg <= 0 || seg_refills == 0) return (int)cudaErrorInvalidValue;
  if (snap != nullptr && nseg != 1) return (int)cudaErrorInvalidValue;
  const dim3 grid((unsigned)((nseg + GEN_WARPS - 1) / GEN_WARPS));
  cudaStream_t st = (cudaStream_t)stream;
#define DS_LAUNCH(KIND)                                                                                                  \
  dsfmt_generate_kernel<KIND><<<grid, GEN_WARPS * 32, 0, st>>>((const W128 *)states, refill0, seg_refills, nseg, w_begin, \
                                                              w_end, (unsigned long long *)out, n_out, (W128 *)snap)
  switch (kind) {
    case DS_RAW: DS_LAUNCH(DS_RAW); break;
    case DS_CLOSE_OPEN: DS_LAUNCH(DS_CLOSE_OPEN); break;
    case DS_OPEN_CLOSE: DS_LAUNCH(DS_OPEN_CLOSE); break;
    case DS_OPEN_OPEN: DS_LAUNCH(DS_OPEN_OPEN); break;
    case DS_CLOSE_CLOSE: DS_LAUNCH(DS_CLOSE_CLOSE); break;
    default: return (int)cudaErrorInvalidValue;
  }
#undef DS_LAUNCH
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}

extern "C" int chol_dsfmt_jump_launch(const void *K, const void *polys, int poly_u32, int D, void *states, int nseg,
                                      void *stream) {
  if (nseg <= 0 || D <= 0 || poly_u32 * 32 < D) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned gx = (unsigned)((nseg + JUMP_GT - 1) / JUMP_GT);
  // enough CTAs to fill the 148 SMs twice when there are few segments; slices are whole 1024-bit stretches
  const unsigned stretches = (unsigned)((D + 1023) / 1024);
  unsigned gy = (2 * 148 + gx - 1) / gx;
  if (gy > stretches) gy = stretches;
  if (gy < 1) gy = 1;
  if (gy > 1) {
    cudaError_t e = cudaMemsetAsync(states, 0, (size_t)nseg * DS_STATE_U32 * sizeof(uint32_t), st);
    if (e != cudaSuccess) return (int)e;
  }
  dsfmt_jump_kernel<<<dim3(gx, gy), JUMP_THREADS, 0, st>>>((const uint32_t *)K, (const uint32_t *)polys, poly_u32, D,
                                                          (uint32_t *)states, nseg);
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}
#endif  // CHOL_CPU_EMU
