// async_copy.cuh -- bulk asynchronous global -> shared copies and the mbarriers they complete on (sm_90+ PTX:
// cp.async.bulk / mbarrier; SASS UBLKCP + SYNCS on sm_100a).
//
// The gather kernels stage a cell tile's neighbourhood as a handful of CONTIGUOUS ranges of entry-ordered arrays
// (tile.cuh, ops_csr.cu): one elected thread arms an mbarrier with the byte count and issues one bulk copy per
// range; the copy engine fills shared memory while the CTA is still working on the previous tile, and the whole
// CTA later waits on the barrier's phase.  No thread ever touches the staged data on its way in.
//
// Rules the callers keep: source, destination and size are multiples of 16 bytes; a buffer that threads have read
// or written through ordinary loads / stores is handed to the copy engine only after a CTA barrier, and the issuing
// thread fences the generic proxy against the async proxy first (async_proxy_fence).
//
// tests/cpusim compiles this header for the host: there a bulk copy is an immediate memcpy by the issuing thread
// and a barrier wait is a CTA barrier (every wait in the library is CTA-uniform).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

namespace sphb {

typedef unsigned long long mbar_t;

#ifndef SPHB_CPUSIM
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(mbar_t *bar, unsigned arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
}
// makes freshly initialised barriers visible to the async proxy (follow with a CTA barrier)
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// one arrival that also announces `bytes` of asynchronous copies to come
__device__ __forceinline__ void mbar_arrive_expect_tx(mbar_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

// global -> shared, completes `bytes` on the barrier
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, mbar_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy ones
__device__ __forceinline__ void async_proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// waits until the phase with the given parity has completed (CTA-uniform in this library)
__device__ __forceinline__ void mbar_wait(mbar_t *bar, unsigned parity)
{
    unsigned done = 0;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
#else
inline void mbar_init(mbar_t *bar, unsigned) { *bar = 0ull; }
inline void mbar_init_fence() {}
inline void mbar_arrive_expect_tx(mbar_t *, unsigned) {}
inline void bulk_copy_g2s(void *dst, const void *src, unsigned bytes, mbar_t *) { std::memcpy(dst, src, bytes); }
inline void async_proxy_fence() {}
inline void mbar_wait(mbar_t *, unsigned) { __syncthreads(); }
#endif

}  // namespace sphb
