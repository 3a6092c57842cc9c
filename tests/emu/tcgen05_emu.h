// tcgen05_emu.h — CPU stand-ins for the inline-PTX primitives of csrc/gemm_tf32x3.cu (included from there when
// CHOL_CPU_EMU is defined): tensor-map TMA boxes with the 128-byte swizzle, tcgen05.mma kind::tf32 reading its operands
// through shared-memory matrix descriptors, TMEM, tcgen05.commit, tcgen05.ld, plus the mbarrier calls mapped onto
// ptx_emu.h.  TEST INFRASTRUCTURE ONLY (tests/emu/emu_tf32x3.cpp).
//
// What is modelled, and from where (so that this is a check of the kernel's encodings, not a mirror of them):
//  * descriptor bit fields as CUTLASS documents them (cute/arch/mma_sm100_desc.hpp: SmemDescriptor - start address
//    [0,14) in 16-byte units, leading byte offset [16,30), stride byte offset [32,46), version [46,48), base offset
//    [49,52), layout type [61,64) with SWIZZLE_128B = 2; InstrDescriptor - c_format [4,6) F32 = 1, a_format [7,10) /
//    b_format [10,13) TF32 = 2, negate bits 13 / 14, a_major bit 15, b_major bit 16 (0 = K-major), n_dim [17,23) = N >> 3,
//    m_dim [24,29) = M >> 4); anything else than the one supported combination aborts;
//  * the 128-byte swizzle as a function of the SHARED ADDRESS (16-byte chunk index, address bits [4,7), XORed with
//    bits [7,10)), applied by the TMA write and by the MMA read alike: a tile that is not 1024-byte aligned, or a
//    K advance that leaves the 128-byte row, reads garbage here as on the device;
//  * K-major SWIZZLE_128B operand: row r of the operand tile lives at start + (r / 8) * SBO + (r % 8) * 128;
//  * TMA: box elements outside the tensor are filled with zeros, the barrier still gets the whole box's bytes; the box
//    lands as late as the program allows - when a thread waits on its barrier (ptx_emu.h, emu_mbar::async_copy);
//  * tcgen05.mma is ASYNCHRONOUS: the emulation queues it and executes the queue at the issuing thread's next
//    tcgen05.commit (the latest moment the program allows), so an epilogue that reads TMEM, or a producer that refills
//    a stage, before the commit's barrier has fired sees stale data; TMEM starts out as NaN;
//  * tcgen05.ld.32x32b: warp w of a warpgroup may only touch TMEM lanes 32 (w % 4) .. + 31 (aborts otherwise);
//    EMU_TC_SLOW_LD_MS makes the loads slow (see tmem_ld32) to put the epilogue behind the issuer.
// Not modelled: timing, bank conflicts, the tf32 datapath's internal accumulation order (plain FP32 sums here).
#pragma once
#include <cuda.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <limits>
#include <thread>
#include <vector>

#include "ptx_emu.h"

namespace chol {

// ---- shared-memory addresses: offsets into the block's dynamic shared memory, keeping the low 10 bits of the real
// address (so that "1024-byte aligned" means the same thing here and for the kernel's own alignment code) ----
namespace emu_tc {
inline uintptr_t smem_origin() { return reinterpret_cast<uintptr_t>(::emu::current()->dyn.data()) & ~uintptr_t(1023); }
inline unsigned char *smem_ptr(uint32_t a) { return reinterpret_cast<unsigned char *>(smem_origin() + a); }
inline uint32_t swz128(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }
[[noreturn]] inline void die(const char *what) {
  std::fprintf(stderr, "tcgen05 emulation: %s\n", what);
  std::abort();
}
// what the harness's stand-in for cuTensorMapEncodeTiled writes into the opaque 128 bytes of a CUtensorMap
struct Map {
  uint64_t magic;
  const float *base;
  uint64_t dim0, dim1;     // elements along K (contiguous) and rows
  uint64_t stride1_bytes;  // row pitch
  uint32_t box0, box1;
};
static_assert(sizeof(Map) <= sizeof(CUtensorMap), "fits the opaque tensor map");
constexpr uint64_t MAP_MAGIC = 0x7463656d75314d41ull;
inline int encode_tiled_2d_f32_sw128(CUtensorMap *out, const void *base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes, uint32_t box0,
                                     uint32_t box1) {
  // the driver's own argument rules for this form (cuTensorMapEncodeTiled): 16-byte aligned base and pitch, inner box
  // extent * element size <= the swizzle span (128 bytes), box extents <= 256
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || (stride1_bytes & 15u) || box0 * 4u > 128u || box0 == 0 || box1 == 0 || box1 > 256u ||
      dim0 == 0 || dim1 == 0 || stride1_bytes < dim0 * 4u)
    return 1;
  Map m{MAP_MAGIC, static_cast<const float *>(base), dim0, dim1, stride1_bytes, box0, box1};
  std::memset(out, 0, sizeof *out);
  std::memcpy(out, &m, sizeof m);
  return 0;
}

constexpr int TMEM_LANES = 128, TMEM_COLS = 512;
inline float (*tmem())[TMEM_COLS] {
  static float t[TMEM_LANES][TMEM_COLS];
  return t;
}
struct Mma { uint32_t d; uint64_t a, b; uint32_t idesc, acc; };
inline std::vector<Mma> &queue() { thread_local std::vector<Mma> q; return q; }

inline float tf32_operand(float x) {  // the tensor core reads 19 bits of an FP32 operand
  uint32_t u;
  std::memcpy(&u, &x, 4);
  u &= ~0x1fffu;
  std::memcpy(&x, &u, 4);
  return x;
}
struct Desc { uint32_t start, sbo; };
inline Desc decode_smem_desc(uint64_t d) {
  const uint32_t version = (uint32_t)(d >> 46) & 3u, base_offset = (uint32_t)(d >> 49) & 7u, lbo_mode = (uint32_t)(d >> 52) & 1u,
                 layout = (uint32_t)(d >> 61) & 7u;
  if (version != 1) die("shared-memory descriptor: version field must be 1 on sm_100");
  if (layout != 2) die("shared-memory descriptor: only SWIZZLE_128B (layout type 2) is modelled");
  if (base_offset != 0 || lbo_mode != 0) die("shared-memory descriptor: base offset / LBO mode must be 0");
  return Desc{(uint32_t)(d & 0x3FFFu) << 4, (uint32_t)((d >> 32) & 0x3FFFu) << 4};
}
inline float operand(const Desc &t, int row, int k) {  // K-major, rows of 128 bytes, 8-row groups SBO apart
  const uint32_t logical = t.start + (uint32_t)(row >> 3) * t.sbo + (uint32_t)(row & 7) * 128u + (uint32_t)k * 4u;
  float v;
  std::memcpy(&v, smem_ptr(swz128(logical)), 4);
  return tf32_operand(v);
}
inline void execute(const Mma &m) {
  const uint32_t i = m.idesc;
  if (((i >> 4) & 3u) != 1u) die("instruction descriptor: accumulator format must be F32");
  if (((i >> 7) & 7u) != 2u || ((i >> 10) & 7u) != 2u) die("instruction descriptor: A and B formats must be TF32 for kind::tf32");
  if ((i & 0xFu) || ((i >> 13) & 0xFu)) die("instruction descriptor: sparse / saturate / negate / MN-major bits are not modelled");
  const int N = (int)((i >> 17) & 0x3Fu) << 3, M = (int)((i >> 24) & 0x1Fu) << 4;
  if (M != 128) die("instruction descriptor: M must be 128 (cta_group::1, one accumulator row per TMEM lane)");
  if (N < 8 || N > 256 || (N & 15)) die("instruction descriptor: N must be a multiple of 16 up to 256 for M = 128");
  const uint32_t lane0 = m.d >> 16, col0 = m.d & 0xFFFFu;
  if (lane0 != 0 || col0 + (uint32_t)N > (uint32_t)TMEM_COLS) die("tcgen05.mma: accumulator outside the allocated TMEM columns");
  const Desc a = decode_smem_desc(m.a), b = decode_smem_desc(m.b);
  float av[128][8];
  for (int r = 0; r < M; ++r)
    for (int k = 0; k < 8; ++k) av[r][k] = operand(a, r, k);
  for (int c = 0; c < N; ++c) {
    float bv[8];
    for (int k = 0; k < 8; ++k) bv[k] = operand(b, c, k);
    for (int r = 0; r < M; ++r) {
      float s = m.acc ? tmem()[r][col0 + c] : 0.f;
      for (int k = 0; k < 8; ++k) s += av[r][k] * bv[k];
      tmem()[r][col0 + c] = s;
    }
  }
}
}  // namespace emu_tc

inline uint32_t s_u32(const void *p) { return (uint32_t)(reinterpret_cast<uintptr_t>(p) - emu_tc::smem_origin()); }

inline void t_mbar_init(uint64_t *bar, uint32_t count) { mbar_init(bar, count); }
inline void t_mbar_arrive(uint64_t *bar) { mbar_arrive(bar); }
inline void t_mbar_expect_tx(uint64_t *bar, uint32_t bytes) { mbar_arrive_expect_tx(bar, bytes); }
inline void t_mbar_wait(uint64_t *bar, uint32_t parity) { mbar_wait(bar, parity); }

inline void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  emu_tc::Map m;
  std::memcpy(&m, map, sizeof m);
  if (m.magic != emu_tc::MAP_MAGIC) emu_tc::die("tensor map was not made by the emulation's encoder");
  const uint32_t d0 = s_u32(dst);
  if (d0 & 127u) emu_tc::die("TMA destination must be 128-byte aligned");
  ::emu::jitter();
  const uintptr_t origin = emu_tc::smem_origin();  // (the copy may be executed by another thread of the block: same block, same origin)
  emu_mbar::async_copy(bar, (long long)m.box0 * m.box1 * 4, [=] {  // lands when somebody waits on `bar` (ptx_emu.h)
    for (uint32_t r = 0; r < m.box1; ++r)
      for (uint32_t c = 0; c < m.box0; ++c) {
        const long long gr = (long long)c1 + r, gc = (long long)c0 + c;
        float v = 0.f;  // out-of-bounds box elements are zero-filled
        if (gr >= 0 && gc >= 0 && (uint64_t)gr < m.dim1 && (uint64_t)gc < m.dim0)
          v = *reinterpret_cast<const float *>(reinterpret_cast<const unsigned char *>(m.base) + (uint64_t)gr * m.stride1_bytes + (uint64_t)gc * 4u);
        std::memcpy(reinterpret_cast<unsigned char *>(origin + emu_tc::swz128(d0 + r * (m.box0 * 4u) + c * 4u)), &v, 4);
      }
  });
}

inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  emu_tc::queue().push_back(emu_tc::Mma{d_tmem, adesc, bdesc, idesc, accumulate});
}
inline void umma_commit(uint64_t *bar) {
  ::emu::jitter();
  for (const emu_tc::Mma &m : emu_tc::queue()) emu_tc::execute(m);
  emu_tc::queue().clear();
  emu_mbar::arrive(bar, 1, 0);
}
inline void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  const uint32_t lane0 = taddr >> 16, col0 = taddr & 0xFFFFu, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane0 != 32u * (warp & 3u)) emu_tc::die("tcgen05.ld: a warp may only read the TMEM lanes of its own quarter (32 * (warp % 4))");
  if (col0 + 32u > (uint32_t)emu_tc::TMEM_COLS) emu_tc::die("tcgen05.ld: columns outside the allocation");
  // EMU_TC_SLOW_LD_MS=n: every tcgen05.ld takes n ms, so the epilogue falls a whole tile behind the MMA issuer (whose
  // emulated MMAs are slow) - an accumulator that the issuer reuses without waiting for the epilogue's release is then
  // overwritten while it is being read, which the plain run's timing never produces
  static const int slow_ms = [] { const char *e = std::getenv("EMU_TC_SLOW_LD_MS"); return e ? std::atoi(e) : 0; }();
  if (slow_ms > 0) std::this_thread::sleep_for(std::chrono::milliseconds(slow_ms));
  ::emu::jitter();
  for (int j = 0; j < 32; ++j) v[j] = emu_tc::tmem()[lane0 + lane][col0 + j];
}

// the in-kernel statements that are macros in the product source
#define T_DYN_SMEM(name) uint8_t *name = ::emu::dyn_smem()
#define T_FENCE_MBAR_INIT() ((void)0)
#define T_TMEM_ALLOC(slot, cols)                                                                            \
  do {                                                                                                      \
    if ((cols) != 512u) ::chol::emu_tc::die("the emulation models one 512-column allocation");              \
    if ((threadIdx.x & 31) == 0) {                                                                          \
      for (int r_ = 0; r_ < ::chol::emu_tc::TMEM_LANES; ++r_)                                               \
        for (int c_ = 0; c_ < ::chol::emu_tc::TMEM_COLS; ++c_)                                              \
          ::chol::emu_tc::tmem()[r_][c_] = std::numeric_limits<float>::quiet_NaN();                         \
      *(slot) = 0u;                                                                                         \
    }                                                                                                       \
    __syncwarp();                                                                                           \
  } while (0)
#define T_TMEM_DEALLOC(base, cols) ((void)(base))

// cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 explicit mantissa bits
inline void tf32_split(float x, float &hi, float &lo) {
  auto rna = [](float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & ~0x1fffu;
    std::memcpy(&f, &u, 4);
    return f;
  };
  hi = rna(x);
  lo = rna(x - hi);
}

}  // namespace chol
