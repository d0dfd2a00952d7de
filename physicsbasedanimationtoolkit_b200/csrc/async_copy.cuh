// Ampere-style asynchronous copies (cp.async / SASS LDGSTS) used by the pipelined step kernels to
// move static data and gathered positions into shared memory without staging through registers.
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

}  // namespace vbdx
