// Error type, CUDA error check and the plain device allocation used by the host driver.
#pragma once

#include "../../include/vbdx.h"

#include <cstdint>
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

namespace vbdx {

struct Error : std::runtime_error {
    vbdx_status status;
    Error(vbdx_status s, std::string const& what) : std::runtime_error(what), status(s) {}
};

#define VBDX_CUDA(call)                                                                          \
    do                                                                                           \
    {                                                                                            \
        cudaError_t const err__ = (call);                                                        \
        if (err__ != cudaSuccess)                                                                \
            throw ::vbdx::Error(                                                                 \
                err__ == cudaErrorMemoryAllocation ? VBDX_OUT_OF_MEMORY : VBDX_CUDA_ERROR,       \
                std::string(#call) + ": " + cudaGetErrorString(err__));                          \
    } while (0)

static void Require(bool cond, char const* what)
{
    if (!cond)
        throw Error(VBDX_INVALID_ARGUMENT, what);
}

// plain device allocation with byte accounting
template <class T>
struct DevBuf {
    T* p     = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(DevBuf const&)            = delete;
    DevBuf& operator=(DevBuf const&) = delete;
    ~DevBuf() { Free(); }
    void Alloc(size_t count, int64_t* accounting = nullptr)
    {
        Free();
        n = count;
        if (count == 0)
            return;
        VBDX_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)));
        if (accounting)
            *accounting += static_cast<int64_t>(count * sizeof(T));
    }
    void Free()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void Upload(T const* src, size_t count, cudaStream_t s)
    {
        VBDX_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void Download(T* dst, size_t count, cudaStream_t s) const
    {
        VBDX_CUDA(cudaMemcpyAsync(dst, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
};

static inline int Blocks(int64_t n, int threads) { return static_cast<int>((n + threads - 1) / threads); }

}  // namespace vbdx
