/*
 * plg_async.cuh - thin inline-PTX wrappers for the sm_100a asynchronous-copy machinery used
 * by the streaming kernels: mbarrier (full/empty pipeline barriers) and the TMA unit's 1-D
 * bulk copies global<->shared (cp.async.bulk, SASS UBLKCP).  Everything here operates on the
 * executing CTA's own shared memory (no clusters).
 */
#ifndef PLG_ASYNC_CUH_
#define PLG_ASYNC_CUH_

#include <cstdint>

namespace plg_async {

__device__ __forceinline__ uint32_t smem_addr(const void * p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}

/* make barrier initialisation visible to the async proxy before the first bulk copy */
__device__ __forceinline__ void fence_barrier_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

/* generic-proxy writes to shared memory -> visible to subsequent async-proxy (bulk) reads */
__device__ __forceinline__ void fence_proxy_async_smem()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t * bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t * bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t * bar, uint32_t parity)
{
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

/* non-blocking probe (try_wait may suspend the warp for a system-dependent time before it
 * reports failure: wrong tool for a loop that polls several barriers) */
__device__ __forceinline__ bool mbar_test_wait(uint64_t * bar, uint32_t parity)
{
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity)
{
  while (!mbar_try_wait(bar, parity)) { }
}

/* TMA 1-D bulk copy global -> shared; completion (bytes) is signalled on `bar`.
 * dst, src 16-byte aligned, bytes a multiple of 16. */
__device__ __forceinline__ void bulk_g2s(void * dst_smem, const void * src_gmem, uint32_t bytes,
                                         uint64_t * bar)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_addr(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}

/* TMA 1-D bulk copy shared -> global, tracked by the issuing thread's bulk async-group */
__device__ __forceinline__ void bulk_s2g(void * dst_gmem, const void * src_smem, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_addr(src_smem)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void bulk_commit()
{
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

/* wait until at most N of this thread's bulk groups still have to READ their source */
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int N>
__device__ __forceinline__ void bulk_wait()
{
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

} // namespace plg_async

#endif
