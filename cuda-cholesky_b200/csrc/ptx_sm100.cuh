// ptx_sm100.cuh — the few inline-PTX primitives the kernels use (sm_100a).
//   * mbarrier producer/consumer handshakes
//   * cp.async.bulk (TMA bulk engine, SASS UBLKCP) global -> shared row copies
//   * FP64 tensor-core MMA  mma.sync.m8n8k4.f64  (SASS DMMA.8x8x4 - the only native FP64 MMA
//     shape on sm_100a: m16n8k{4,8,16}.f64 are expanded by ptxas into 2/4/8 of these)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#ifdef CHOL_CPU_EMU
#include "ptx_emu.h"  // tests/emu: CPU stand-ins with the same names and signatures
#else

namespace chol {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// make mbarrier.init visible to the async proxy (TMA) before first use
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t *bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// TMA bulk copy of `bytes` (multiple of 16, 16B-aligned both sides) global -> this CTA's shared
// memory; completion is signalled as transaction bytes on `bar`.
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// L2 prefetch of `bytes` (multiple of 16) through the TMA engine - no register or smem destination
__device__ __forceinline__ void l2_prefetch_bulk(const void *src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

// one element (8 or 16 bytes) global -> shared, asynchronously (SASS LDGSTS)
template <int BYTES>
__device__ __forceinline__ void cp_async_elem(void *dst_smem, const void *src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "n"(BYTES) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (count it in the init)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D(8x8) += A(8x4) * B(4x8).  lane = 4*g + t:  a = A[g][t], b = B[t][g], d0/d1 = D[g][2t], D[g][2t+1]
__device__ __forceinline__ void dmma_884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

}  // namespace chol

#endif  // CHOL_CPU_EMU
