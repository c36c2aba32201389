/**
 * @file tma.hxx
 * @brief Thin inline-PTX wrappers for the Blackwell/Hopper async-copy engine:
 * mbarrier + 1-D bulk copy global -> shared (`cp.async.bulk`, SASS `UBLKCP`).
 *
 * Used to stage row-offset windows (the arrays the merge-path / upper-bound
 * searches run over) and nonzero streams into shared memory without tying up
 * registers or the LSU. No reference counterpart: the reference stages with
 * per-thread loads (reference schedule/merge_path_flat.hxx:309-316).
 *
 * Rules of the instruction (PTX ISA, cp.async.bulk): source and destination
 * 16-byte aligned, size a multiple of 16 bytes; completion is signalled as
 * `complete_tx` bytes on an mbarrier in shared memory.
 */
#pragma once

#include <cstdint>

namespace loops {
namespace tma {

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 900)
#define LOOPS_HAS_BULK_COPY 1
#else
#define LOOPS_HAS_BULK_COPY 0
#endif

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

/// Initialise an mbarrier expecting `arrivals` arrive operations per phase and
/// make the initialisation visible to the async proxy.
__device__ __forceinline__ void barrier_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)),
               "r"(arrivals)
               : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

/// One arrival that also announces `bytes` of pending async-copy traffic.
__device__ __forceinline__ void barrier_arrive_expect_tx(uint64_t* bar,
                                                         uint32_t bytes) {
  asm volatile(
      "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
          smem_addr(bar)),
      "r"(bytes)
      : "memory");
}

/// Plain arrival (consumer releasing a stage).
__device__ __forceinline__ void barrier_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar))
               : "memory");
}

/// Block until the phase with the given parity has completed.
__device__ __forceinline__ void barrier_wait(uint64_t* bar, uint32_t parity) {
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

/// Same, but each failed probe lets the hardware park the thread (suspend-time
/// hint, in ns) instead of spinning: waiting warps stop competing for issue
/// slots with the warps that still have work.
__device__ __forceinline__ void barrier_wait_suspend(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITS_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONES_%=;\n"
      "bra WAITS_%=;\n"
      "DONES_%=:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}

/// 1-D bulk copy global -> shared::cta, completion on `bar`.
/// `bytes` % 16 == 0, both addresses 16-byte aligned, bytes > 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst,
                                         const void* gmem_src,
                                         uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}

/// Same with an L2 eviction-priority hint (createpolicy result).
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst,
                                              const void* gmem_src,
                                              uint32_t bytes,
                                              uint64_t* bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_addr(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
      : "memory");
}

__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;"
               : "=l"(p));
  return p;
}

__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;"
               : "=l"(p));
  return p;
}

/// Order prior generic-proxy shared-memory accesses before later async-proxy
/// ones (needed before re-filling a buffer threads have just read/written).
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace tma
}  // namespace loops
