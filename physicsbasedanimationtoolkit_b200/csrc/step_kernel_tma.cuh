// Warp-specialised variant of the persistent VBD step kernel (sm_100a).
//
// One CTA per SM.  The last warp is a *producer*: a single elected lane streams the CTA's
// incidence-record blocks (1.5 KB each, static rest data) from HBM into a shared-memory ring with
// 1-D bulk asynchronous copies (cp.async.bulk, i.e. the TMA engine; SASS UBLKCP) that signal
// per-slot "full" mbarriers.  Because the records never change, the producer runs ahead of the
// colour barriers: while the consumers wait for the other SMs at the end of colour c, the
// records of colour c+1 are already landing in shared memory, so the HBM stream is continuous.
// The remaining warps are *consumers*: each owns a static subset of the CTA's warp tiles, waits
// on the slot's mbarrier, gathers the (mutable) neighbour positions from L2, accumulates the
// closed-form Stable Neo-Hookean block, reduces over the lanes that share a vertex, and does the
// fused damping / inertia / Newton / Chebyshev epilogue exactly as step_kernel.cuh (same
// arithmetic, same summation order).
#pragma once

#include "step_kernel.cuh"

namespace vbdx {

constexpr int kTmaThreads       = 768;                     // 23 consumer warps + 1 producer warp
constexpr int kTmaConsumerWarps = kTmaThreads / 32 - 1;

struct TmaParams {
    StepParams base;
    const uint32_t* __restrict__ ctaBlockBegin;  // [nColors][gridDim.x + 1] first record block per CTA
    uint32_t ringSlots;                          // R: ring capacity in blocks
};

__device__ __forceinline__ uint32_t SmemAddr(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
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

// Record source of the consumers: this lane's 48-byte record out of the shared-memory ring slot that the
// producer filled; the slot is handed back as soon as the warp has copied its records to registers.
struct RingRecords {
    unsigned char const* smem;
    uint32_t full, empty, R, slot, fill, lane;
    __device__ __forceinline__ void Fetch(float4& c0, float4& c1, float4& c2)
    {
        MbarWait(full + 8 * slot, fill & 1u);
        unsigned char const* blk = smem + slot * kBlockBytes + lane * 16;
        c0 = *reinterpret_cast<float4 const*>(blk);
        c1 = *reinterpret_cast<float4 const*>(blk + 512);
        c2 = *reinterpret_cast<float4 const*>(blk + 1024);
        __syncwarp();
        if (lane == 0)
            MbarArrive(empty + 8 * slot);
        if (++slot == R)
        {
            slot = 0;
            ++fill;
        }
    }
};

// grid barrier among the consumer threads of all CTAs (the producer warp never joins)
__device__ __forceinline__ void ConsumerGridBarrier(unsigned int* counter, unsigned int& target)
{
    asm volatile("bar.sync 1, %0;" ::"n"(kTmaConsumerWarps * 32) : "memory");
    if (threadIdx.x == 0)
    {
        target += gridDim.x;
        AddRelease(counter, 1u);
        while (LoadAcquire(counter) < target)
        {
        }
        __threadfence();
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kTmaConsumerWarps * 32) : "memory");
}

template <bool kChebyshev, bool kDamping>
__global__ void __launch_bounds__(kTmaThreads, 1) StepKernelTma(const __grid_constant__ TmaParams tp)
{
    extern __shared__ __align__(128) unsigned char smem[];
    StepParams const& p  = tp.base;
    uint32_t const R     = tp.ringSlots;
    uint32_t const ring  = SmemAddr(smem);
    uint32_t const full  = ring + R * kBlockBytes;
    uint32_t const empty = full + R * 8;
    float4* const stage  = reinterpret_cast<float4*>(smem + R * (kBlockBytes + 16)) + (threadIdx.x >> 5) * p.stageEntries;
    uint32_t const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;

    if (threadIdx.x == 0)
    {
        for (uint32_t s = 0; s < R; ++s)
        {
            MbarInit(full + 8 * s, 1);
            MbarInit(empty + 8 * s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    uint32_t const* blkBegin = tp.ctaBlockBegin + blockIdx.x;
    uint32_t const stride    = gridDim.x + 1;

    if (warp == kTmaConsumerWarps)
    {
        // ------------------------------ producer ------------------------------
        if (lane == 0)
        {
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            uint32_t slot = 0, fill = 0;
            for (int s = 0; s < p.substeps; ++s)
                for (int k = 0; k < p.iterations; ++k)
                    for (int c = 0; c < p.nColors; ++c)
                    {
                        uint32_t const b0 = __ldg(blkBegin + c * stride), b1 = __ldg(blkBegin + c * stride + 1);
                        for (uint32_t b = b0; b < b1; ++b)
                        {
                            if (fill > 0)
                                MbarWait(empty + 8 * slot, (fill & 1u) ^ 1u);
                            MbarArriveExpectTx(full + 8 * slot, kBlockBytes);
                            BulkLoad(ring + slot * kBlockBytes, p.records + static_cast<size_t>(b) * kBlockFloat4,
                                     kBlockBytes, full + 8 * slot, policy);
                            if (++slot == R)
                            {
                                slot = 0;
                                ++fill;
                            }
                        }
                    }
        }
        return;
    }

    // ------------------------------ consumers ------------------------------
    unsigned int target      = 0;
    uint32_t const ctid      = blockIdx.x * (kTmaConsumerWarps * 32) + threadIdx.x;
    uint32_t const cstride   = gridDim.x * (kTmaConsumerWarps * 32);
    uint32_t streamBase      = 0;  // blocks of this CTA's stream before the current colour

    for (int s = 0; s < p.substeps; ++s)
    {
        for (uint32_t i = ctid; i < static_cast<uint32_t>(p.nVerts); i += cstride)
        {
            float4 const x4 = __ldcg(p.pos + p.pOff + i);
            float4 v4       = __ldcg(p.vel + i);
            float3 const vprev = make_float3(v4.x, v4.y, v4.z);
            if (s > 0)
            {
                float4 const xt4 = __ldcg(p.xt + i);
                v4.x = (x4.x - xt4.x) / p.sdt;
                v4.y = (x4.y - xt4.y) / p.sdt;
                v4.z = (x4.z - xt4.z) / p.sdt;
                p.vel[i] = v4;
            }
            float3 vtm1 = make_float3(v4.x, v4.y, v4.z);
            if (p.vtm1 != nullptr)
            {
                if (s > 0)
                    vtm1 = vprev;
                else
                {
                    float4 const q = __ldcg(p.vtm1 + i);
                    vtm1           = make_float3(q.x, q.y, q.z);
                }
            }
            float4 const a4 = __ldg(p.aext + i);
            float4 xm       = __ldcg(p.xtildeM + i);
            xm.x            = x4.x + p.sdt * v4.x + p.sdt2 * a4.x;
            xm.y            = x4.y + p.sdt * v4.y + p.sdt2 * a4.y;
            xm.z            = x4.z + p.sdt * v4.z + p.sdt2 * a4.z;
            p.xtildeM[i]    = xm;
            p.xt[i]         = x4;
            float3 const x0 = InitialPosition(
                make_float3(x4.x, x4.y, x4.z), vtm1, make_float3(v4.x, v4.y, v4.z),
                make_float3(a4.x, a4.y, a4.z), p.sdt, p.sdt2, p.strategy);
            float4 const o = make_float4(x0.x, x0.y, x0.z, 0.f);
            p.pos[i]       = o;
            if constexpr (kChebyshev)
                p.pos[p.pOff + i] = o;
        }
        ConsumerGridBarrier(p.barrier, target);

        for (int k = 0; k < p.iterations; ++k)
        {
            float const omega = kChebyshev ? __ldg(p.omega + k) : 1.f;
            for (int c = 0; c < p.nColors; ++c)
            {
                uint32_t const* range = p.ctaTileRange + static_cast<size_t>(c) * stride + blockIdx.x;
                uint32_t const tBegin = __ldg(range), tEnd = __ldg(range + 1);
                uint32_t const b0 = __ldg(blkBegin + c * stride), b1 = __ldg(blkBegin + c * stride + 1);
                for (uint32_t T = tBegin + warp; T < tEnd; T += kTmaConsumerWarps)
                {
                    uint4 const td    = __ldg(p.tiles + T);
                    uint32_t const n0 = streamBase + (td.x - b0);
                    RingRecords src{smem, full, empty, R, n0 % R, n0 / R, lane};
                    ProcessTile<kChebyshev, kDamping>(p, td, stage, src, k, omega, lane);
                }
                streamBase += b1 - b0;
                ConsumerGridBarrier(p.barrier, target);
            }
        }
    }
    for (uint32_t i = ctid; i < static_cast<uint32_t>(p.nVerts); i += cstride)
    {
        float4 const x4  = __ldcg(p.pos + p.pOff + i);
        float4 const xt4 = __ldcg(p.xt + i);
        float4 v4        = __ldcg(p.vel + i);
        if (p.vtm1 != nullptr)
            p.vtm1[i] = v4;
        v4.x     = (x4.x - xt4.x) / p.sdt;
        v4.y     = (x4.y - xt4.y) / p.sdt;
        v4.z     = (x4.z - xt4.z) / p.sdt;
        p.vel[i] = v4;
    }
}

}  // namespace vbdx
