// climt_b200 -- bulk asynchronous copies global -> shared memory (the 1-D form of Blackwell's TMA engine: cp.async.bulk,
// SASS UBLKCP) completing on an mbarrier.  Device code only.  Used by the taumol kernels to stage a band's table block.
#pragma once
#include <cstdint>

namespace cb {
namespace bulk {

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // make the initialised barrier visible to the async proxy
}
// the issuing thread's arrival + the number of bytes the copies will deliver
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// bytes: multiple of 16; dst and src 16-byte aligned
__device__ __forceinline__ void copy_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
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
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}

}  // namespace bulk

// Block-wide barrier that lanes of one warp may reach from DIFFERENT program locations (lanes whose column is past the end of a
// ragged chunk wait in a dummy loop while their neighbours are inside the transfer function).  __syncthreads() is bar.sync =
// barrier.sync.aligned, which is undefined when the lanes of a warp diverge (r02: hung the first ragged launch); the unaligned
// form counts arriving THREADS.  Barrier 1 (0 is __syncthreads'); nthreads must be a multiple of 32.
__device__ __forceinline__ void barrier_unaligned(int nthreads) {
  asm volatile("barrier.sync 1, %0;" ::"r"(nthreads) : "memory");
}
}  // namespace cb
