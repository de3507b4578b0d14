// Bulk (TMA) staging of read-only tables into shared memory: one elected thread issues
// cp.async.bulk global -> shared copies that complete on an mbarrier; the CTA waits once.
#pragma once
#include <stdint.h>

namespace mst {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// stage the robot image at smem[0 .. rbytes) and the environment image right behind it;
// every thread of the CTA must call this (it contains the CTA barrier)
__device__ __forceinline__ void stage_meshes(unsigned char* smem, const void* robot_img, size_t rbytes,
                                             const void* env_img, size_t ebytes, unsigned long long* bar) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (unsigned)(rbytes + ebytes));
    if (rbytes) bulk_g2s(smem, robot_img, (unsigned)rbytes, bar);
    if (ebytes) bulk_g2s(smem + rbytes, env_img, (unsigned)ebytes, bar);
  }
  mbar_wait(bar, 0);
}

}  // namespace mst
