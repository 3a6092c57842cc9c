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
constexpr int JUMP_THREADS = DS_STATE_WORDS;    // one 128-bit column of the state per thread (6 warps)
constexpr int JUMP_GT = 16;                     // segments per CTA of the jump kernel
constexpr int JUMP_BATCH = 8;                   // K rows a thread holds in registers (and as many in flight)

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

// what happens to word i of a refill the moment its owner lane has it in registers
template <int KIND>
struct EmitWindow {  // converted and stored if the stream word lies inside the request
  unsigned long long t0, w_begin, w_end, n_out;
  unsigned long long *out;
  bool aligned16;
  __device__ __forceinline__ void operator()(int i, W128 v) const {
    const unsigned long long t = t0 + (unsigned)i;
    if (t < w_begin || t >= w_end) return;
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
};
struct EmitSnapshot {  // the raw state (Krylov block)
  W128 *dst;
  __device__ __forceinline__ void operator()(int i, W128 v) const { dst[i] = v; }
};

// One refill of the warp's state st[0..190] in shared memory, three rounds of up to 64 steps; `lung` lives in registers
// (every lane holds it).  Each new word goes to st and to `emit` straight from the registers of the lane that made it:
// a lane's two words are 32 contiguous bytes of the stream, a round is 1 KB, so the stores fill whole sectors and stay
// off the critical path (the next round does not wait for them).  Ends with a __syncwarp(): st is consistent on return.
template <typename Emit>
__device__ __forceinline__ void dsfmt_refill_warp(W128 *st, W128 &lung, int lane, const Emit &emit) {
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
    lung = shfl_w(has1 ? y1 : y0, 31);
    if (has0) {
      const W128 r0 = mix_out(y0, s0);
      st[i0] = r0;
      emit(i0, r0);
    }
    if (has1) {
      const W128 r1 = mix_out(y1, s1);
      st[i1] = r1;
      emit(i1, r1);
    }
    __syncwarp();  // the next round reads words this round wrote (s[i-74]) from other lanes
  }
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
  __shared__ W128 smem[GEN_WARPS][DS_N];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int g = blockIdx.x * GEN_WARPS + wid;
  if (g >= nseg) return;  // whole warps leave; no block-wide barrier below
  W128 *st = smem[wid];
  const W128 *src = states + (size_t)g * DS_STATE_WORDS;
  for (int i = lane; i < DS_N; i += 32) st[i] = src[i];
  W128 lung = src[DS_N];
  __syncwarp();
  const unsigned long long r_begin = refill0 + (unsigned long long)g * seg_refills;
  const bool aligned16 = (((size_t)out) & 15) == 0;
  for (unsigned k = 0; k < seg_refills; ++k) {
    const unsigned long long r = r_begin + k;
    const unsigned long long t0 = r * DS_N;  // first stream word of this refill
    if (snap != nullptr) {
      W128 *dst = snap + (size_t)(r + 1) * DS_STATE_WORDS;
      dsfmt_refill_warp(st, lung, lane, EmitSnapshot{dst});
      if (lane == 0) dst[DS_N] = lung;
    } else {
      if (t0 >= w_end) break;
      dsfmt_refill_warp(st, lung, lane, EmitWindow<KIND>{t0, w_begin, w_end, n_out, out, aligned16});
    }
  }
}

// states[g][:] ^= XOR over the rows i of this CTA's slice of [0, D) with bit i of polys[g] set of K[i][:]
// (K: D states of 192 128-bit words; polys: nseg rows of poly_u32 32-bit words; states zeroed by the caller when
// gridDim.y > 1).  grid.x = tiles of JUMP_GT segments, grid.y = slices of the row range (whole 32-bit words of the
// polynomials).  A thread owns one 128-bit column of the state for all JUMP_GT segments (64 accumulator registers) and
// streams the K rows through registers eight at a time, the next eight already in flight (the table lives in L2; one
// row is read once per CTA, i.e. once per JUMP_GT segments).
__device__ __forceinline__ void jump_load_rows(uint4 (&k)[JUMP_BATCH], const uint4 *__restrict__ K, int row0, int D, int t) {
#pragma unroll
  for (int j = 0; j < JUMP_BATCH; ++j)
    k[j] = (row0 + j < D) ? K[(size_t)(row0 + j) * DS_STATE_WORDS + t] : make_uint4(0u, 0u, 0u, 0u);
}

__global__ void __launch_bounds__(JUMP_THREADS)
dsfmt_jump_kernel(const uint4 *__restrict__ K, const uint32_t *__restrict__ polys, int poly_u32, int D,
                  uint32_t *__restrict__ states, int nseg) {
  __shared__ uint32_t bits[JUMP_GT][32];  // one 1024-row stretch of every segment's polynomial at a time
  const int t = threadIdx.x;
  const int g0 = blockIdx.x * JUMP_GT;
  const int words = (D + 31) / 32;
  const int per = (words + gridDim.y - 1) / gridDim.y;
  const int w_lo = blockIdx.y * per, w_hi = min(words, w_lo + per);
  uint4 acc[JUMP_GT];
#pragma unroll
  for (int g = 0; g < JUMP_GT; ++g) acc[g] = make_uint4(0u, 0u, 0u, 0u);
  const int r_lo = w_lo * 32, r_hi = min(D, w_hi * 32);
  uint4 kc[JUMP_BATCH], kn[JUMP_BATCH];
  if (r_lo < r_hi) jump_load_rows(kc, K, r_lo, D, t);
  for (int rb = r_lo; rb < r_hi; rb += JUMP_BATCH) {
    if (rb + JUMP_BATCH < r_hi) jump_load_rows(kn, K, rb + JUMP_BATCH, D, t);  // in flight while this batch is applied
    const int w = rb >> 5;
    if (((w - w_lo) & 31) == 0 && (rb & 31) == 0) {  // a new 1024-row stretch: stage its polynomial words
      __syncthreads();
      for (int e = t; e < JUMP_GT * 32; e += JUMP_THREADS) {
        const int g = e >> 5, ww = w + (e & 31);
        bits[g][e & 31] = (g0 + g < nseg && ww < w_hi && ww < poly_u32) ? polys[(size_t)(g0 + g) * poly_u32 + ww] : 0u;
      }
      __syncthreads();
    }
    const int wi = (w - w_lo) & 31, sh = rb & 31;
#pragma unroll
    for (int g = 0; g < JUMP_GT; ++g) {
      const uint32_t m = bits[g][wi] >> sh;  // bit j = row rb + j (rows beyond D have zero bits: deg a < D)
#pragma unroll
      for (int j = 0; j < JUMP_BATCH; ++j)
        if ((m >> j) & 1u) {
          acc[g].x ^= kc[j].x;
          acc[g].y ^= kc[j].y;
          acc[g].z ^= kc[j].z;
          acc[g].w ^= kc[j].w;
        }
    }
#pragma unroll
    for (int j = 0; j < JUMP_BATCH; ++j) kc[j] = kn[j];
  }
#pragma unroll
  for (int g = 0; g < JUMP_GT; ++g) {
    if (g0 + g >= nseg) break;
    uint32_t *dst = states + (size_t)(g0 + g) * DS_STATE_U32 + 4 * t;
    if (gridDim.y == 1) {
      *reinterpret_cast<uint4 *>(dst) = acc[g];
    } else {
      if (acc[g].x) atomicXor(dst + 0, acc[g].x);
      if (acc[g].y) atomicXor(dst + 1, acc[g].y);
      if (acc[g].z) atomicXor(dst + 2, acc[g].z);
      if (acc[g].w) atomicXor(dst + 3, acc[g].w);
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
  dsfmt_jump_kernel<<<dim3(gx, gy), JUMP_THREADS, 0, st>>>((const uint4 *)K, (const uint32_t *)polys, poly_u32, D,
                                                          (uint32_t *)states, nseg);
  CHOL_COUNT_LAUNCH();
  return (int)cudaGetLastError();
}
#endif  // CHOL_CPU_EMU
