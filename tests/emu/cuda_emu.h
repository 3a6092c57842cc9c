// cuda_emu.h — run a CUDA kernel's SOURCE on the CPU, one OS thread per CUDA thread, one block at a
// time: enough of the execution model (threadIdx/blockIdx, static and dynamic shared memory,
// __syncthreads, warp shuffles, a few atomics) to check the LOGIC of the kernels in
// cuda-cholesky_b200/csrc without a GPU.  TEST INFRASTRUCTURE ONLY (tests/test_kernel_emulation.py);
// it says nothing about races that only real hardware scheduling exposes, and nothing about speed.
//
// How: the .cu file is #included into a host C++ program after this header.  Under plain g++ the
// CUDA headers define __global__/__device__/... as nothing; __shared__ is pre-defined here as
// `static` (one block runs at a time, so function-local statics ARE the block's shared memory) and
// the kernels' dynamic shared memory comes from CHOL_DYN_SMEM (csrc/emu_compat.cuh).
#pragma once
#define CHOL_CPU_EMU 1
#define __shared__ static
#include <cuda_runtime.h>
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

#include <pthread.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace emu {

struct Barrier {  // reusable, blocking (the box has far fewer cores than a block has threads)
  pthread_barrier_t b;
  explicit Barrier(unsigned n) { pthread_barrier_init(&b, nullptr, n); }
  ~Barrier() { pthread_barrier_destroy(&b); }
  void wait() { pthread_barrier_wait(&b); }
};

struct Warp {
  Barrier bar{32};
  alignas(16) unsigned char slot[32][16];
};

struct Block {
  unsigned nthreads;
  Barrier bar;
  std::vector<Warp> warps;
  std::vector<unsigned char> dyn;
  explicit Block(unsigned n, size_t dyn_bytes) : nthreads(n), bar(n), warps((n + 31) / 32), dyn(dyn_bytes + 256) {}
};

inline Block *&current() { static Block *b = nullptr; return b; }

// cp.async (LDGSTS) groups, completed as LATE as the program allows: a copy lands only when a
// wait_group forces its group - the adversarial end of what the hardware may do, so code that
// reads a stage before waiting for it reads stale data here.
struct Copy { void *dst; const void *src; int bytes; };
inline std::vector<Copy> &cp_open() { thread_local std::vector<Copy> v; return v; }
inline std::vector<std::vector<Copy>> &cp_groups() { thread_local std::vector<std::vector<Copy>> v; return v; }
inline void cp_async(void *dst, const void *src, int bytes) { cp_open().push_back(Copy{dst, src, bytes}); }
inline void cp_async_commit() { cp_groups().push_back(cp_open()); cp_open().clear(); }
inline void cp_async_wait(int keep) {
  auto &g = cp_groups();
  while ((int)g.size() > keep) {
    for (const Copy &c : g.front()) std::memcpy(c.dst, c.src, (size_t)c.bytes);
    g.erase(g.begin());
  }
}
inline unsigned char *dyn_smem() {  // 128-byte aligned start of the block's dynamic shared memory
  auto p = reinterpret_cast<uintptr_t>(current()->dyn.data());
  return reinterpret_cast<unsigned char *>((p + 127) & ~uintptr_t(127));
}

}  // namespace emu

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
#ifdef CUDA_EMU_IMPLEMENTATION
thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
#endif

// EMU_JITTER=1: every thread stalls for a random short while now and then (after barriers, around waits), so that
// the relative progress of warps and lanes varies from run to run far more than the OS scheduler alone would do
namespace emu {
inline bool jitter_on() { static const bool on = [] { const char *e = std::getenv("EMU_JITTER"); return e && std::atoi(e); }(); return on; }
inline void jitter() {
  if (!jitter_on()) return;
  thread_local unsigned rng = 2463534242u ^ (unsigned)(uintptr_t)&rng;
  rng ^= rng << 13; rng ^= rng >> 17; rng ^= rng << 5;
  if ((rng & 7u) == 0) std::this_thread::sleep_for(std::chrono::microseconds((rng >> 8) % 300));
}
}  // namespace emu
inline void __syncthreads() { emu::current()->bar.wait(); emu::jitter(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::current()->warps[threadIdx.x >> 5].bar.wait(); emu::jitter(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

// all 32 lanes of the warp must call (full-mask use only, as in the kernels)
template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {  // width: power of two; lanes exchange inside their sub-group
  static_assert(sizeof(T) <= 16, "shuffle payload");
  emu::Warp &w = emu::current()->warps[threadIdx.x >> 5];
  const int lane = (int)(threadIdx.x & 31);
  std::memcpy(w.slot[lane], &v, sizeof(T));
  w.bar.wait();
  T r;
  std::memcpy(&r, w.slot[(lane & ~(width - 1)) | (src & (width - 1))], sizeof(T));
  w.bar.wait();
  emu::jitter();  // lanes leave a shuffle at different times: code that needs a __syncwarp() and lacks it shows
  return r;
}
template <typename T>
inline T __shfl_xor_sync(unsigned m, T v, int lanemask) { return __shfl_sync(m, v, (int)((threadIdx.x & 31) ^ lanemask)); }
// lanes below `delta` get their own value back, as on the device
template <typename T>
inline T __shfl_up_sync(unsigned m, T v, unsigned delta, int width = 32) {
  const int sub = (int)(threadIdx.x & 31) & (width - 1);
  return __shfl_sync(m, v, sub >= (int)delta ? sub - (int)delta : sub, width);
}

inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float rsqrt(float x) { return 1.0f / std::sqrt(x); }

template <typename T> inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T atomicXor(T *p, T v) { return __atomic_fetch_xor(p, v, __ATOMIC_SEQ_CST); }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline int atomicMin(int *p, int v) { int o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
inline int atomicMax(int *p, int v) { int o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
using ::fma;
using ::fmaf;
using ::sqrt;
using ::log;
using ::logf;
template <typename T> inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> inline T max(T a, T b) { return a > b ? a : b; }

namespace emu {

// run `body()` (a call of the kernel function with its arguments) as a grid of blocks of threads
inline void launch(dim3 grid, dim3 block, size_t dyn_bytes, const std::function<void()> &body) {
  const unsigned nt = block.x * block.y * block.z;
  if (nt % 32 != 0) { std::fprintf(stderr, "emu: block size must be a multiple of 32\n"); std::abort(); }
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        Block blk(nt, dyn_bytes);
        current() = &blk;
        std::vector<std::thread> th;
        th.reserve(nt);
        for (unsigned t = 0; t < nt; ++t)
          th.emplace_back([&, t] {
            threadIdx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            blockIdx = uint3{bx, by, bz};
            blockDim = block;
            gridDim = grid;
            body();
          });
        for (auto &x : th) x.join();
        current() = nullptr;
      }
}

}  // namespace emu
