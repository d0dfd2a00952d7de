// Asynchronous copies used by the pipelined step kernels to move data into shared memory without
// staging through registers: Ampere-style cp.async (SASS LDGSTS) for scattered gathers, and 1-D bulk
// copies (cp.async.bulk, the TMA engine, SASS UBLKCP) completing on mbarriers for contiguous streams.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace vbdx {

__device__ __forceinline__ uint32_t SmemAddr(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void CpAsync4(uint32_t dstSmem, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dstSmem), "l"(src) : "memory");
}

__device__ __forceinline__ void CpAsync16(uint32_t dstSmem, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dstSmem), "l"(src) : "memory");
}

__device__ __forceinline__ void CpAsyncWaitAll()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}

__device__ __forceinline__ void CpAsyncCommit()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int kPending>
__device__ __forceinline__ void CpAsyncWaitGroup()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

__device__ __forceinline__ void MbarInit(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void MbarWait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// same, for lanes of one warp that wait on *different* barriers (divergent exit)
__device__ __forceinline__ void MbarWaitDivergent(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_D:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_LOOP_D;\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void MbarArrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void MbarArriveExpectTx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void BulkLoad(uint32_t dstSmem, const void* srcGmem, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dstSmem),
        "l"(srcGmem), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}

// asks the L2 to fetch a contiguous range (no destination, no completion): one instruction of one lane
__device__ __forceinline__ void BulkPrefetchL2(const void* srcGmem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(srcGmem), "r"(bytes) : "memory");
}

}  // namespace vbdx
