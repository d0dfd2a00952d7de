// Warp-specialised variant of the persistent VBD step kernel (sm_100a).
//
// One CTA per SM.  The last warp is a *producer*: a single elected lane streams the CTA's
// incidence-record blocks (2 KB each, static rest data) from HBM into a shared-memory ring with
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
constexpr int kBlockBytes       = kBlockFloat4 * 16;        // 2048

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
    float4 const* __restrict__ posQ = p.pos;
    float4 const* __restrict__ posP = p.pos + p.pOff;

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
                    uint4 const td        = __ldg(p.tiles + T);
                    uint32_t const lw     = td.z & 0xffu;
                    uint32_t const iters  = (td.z >> 8) & 0xffffu;
                    uint32_t const nverts = td.z >> 24;
                    uint32_t const grp    = lane >> lw;
                    bool const valid      = grp < nverts;
                    uint32_t const vi     = td.y + (valid ? grp : 0u);
                    float4 const xi       = LoadPos(posP + vi);
                    uint32_t const n0     = streamBase + (td.x - b0);
                    uint32_t slot         = n0 % R;
                    uint32_t fill         = n0 / R;

                    float h00 = 0.f, h01 = 0.f, h02 = 0.f, h11 = 0.f, h12 = 0.f, h22 = 0.f, hd = 0.f;
                    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
                    // software pipeline: the neighbour gathers of block t+1 are in flight while block t is computed
                    MbarWait(full + 8 * slot, fill & 1u);
                    float4 c0 = *reinterpret_cast<float4 const*>(smem + slot * kBlockBytes + lane * 16);
                    float4 q1, q2, q3;
                    {
                        uint32_t const j1 = __float_as_uint(c0.x), j2 = __float_as_uint(c0.y), j3 = __float_as_uint(c0.z);
                        q1 = LoadPos(posQ + (j1 & ~kPrevFlag) + ((j1 & kPrevFlag) ? p.pOff : 0u));
                        q2 = LoadPos(posQ + (j2 & ~kPrevFlag) + ((j2 & kPrevFlag) ? p.pOff : 0u));
                        q3 = LoadPos(posQ + (j3 & ~kPrevFlag) + ((j3 & kPrevFlag) ? p.pOff : 0u));
                    }
#pragma unroll 1
                    for (uint32_t t = 0; t < iters; ++t)
                    {
                        unsigned char const* blk = smem + slot * kBlockBytes + lane * 16;
                        float4 const c1 = *reinterpret_cast<float4 const*>(blk + 512);
                        float4 const c2 = *reinterpret_cast<float4 const*>(blk + 1024);
                        float4 const c3 = *reinterpret_cast<float4 const*>(blk + 1536);
                        float const a0  = c0.w;
                        float4 const p1 = q1, p2 = q2, p3 = q3;
                        __syncwarp();
                        if (lane == 0)
                            MbarArrive(empty + 8 * slot);  // slot may be refilled
                        if (++slot == R)
                        {
                            slot = 0;
                            ++fill;
                        }
                        if (t + 1 < iters)
                        {
                            MbarWait(full + 8 * slot, fill & 1u);
                            c0 = *reinterpret_cast<float4 const*>(smem + slot * kBlockBytes + lane * 16);
                            uint32_t const j1 = __float_as_uint(c0.x), j2 = __float_as_uint(c0.y), j3 = __float_as_uint(c0.z);
                            q1 = LoadPos(posQ + (j1 & ~kPrevFlag) + ((j1 & kPrevFlag) ? p.pOff : 0u));
                            q2 = LoadPos(posQ + (j2 & ~kPrevFlag) + ((j2 & kPrevFlag) ? p.pOff : 0u));
                            q3 = LoadPos(posQ + (j3 & ~kPrevFlag) + ((j3 & kPrevFlag) ? p.pOff : 0u));
                        }
                        float const a1 = c1.x, a2 = c1.y;
                        float const bb0 = c1.z, bb1 = c1.w, bb2 = c2.x;
                        float const e0 = c2.y, e1 = c2.z, e2 = c2.w;
                        float const wmu = c3.x, wlam = c3.y, alpha = c3.z, gh2 = c3.w;
                        float const d1x = p1.x - xi.x, d1y = p1.y - xi.y, d1z = p1.z - xi.z;
                        float const d2x = p2.x - xi.x, d2y = p2.y - xi.y, d2z = p2.z - xi.z;
                        float const d3x = p3.x - xi.x, d3y = p3.y - xi.y, d3z = p3.z - xi.z;
                        float const F00 = d1x * a0 + d2x * bb0 + d3x * e0;
                        float const F01 = d1x * a1 + d2x * bb1 + d3x * e1;
                        float const F02 = d1x * a2 + d2x * bb2 + d3x * e2;
                        float const F10 = d1y * a0 + d2y * bb0 + d3y * e0;
                        float const F11 = d1y * a1 + d2y * bb1 + d3y * e1;
                        float const F12 = d1y * a2 + d2y * bb2 + d3y * e2;
                        float const F20 = d1z * a0 + d2z * bb0 + d3z * e0;
                        float const F21 = d1z * a1 + d2z * bb1 + d3z * e1;
                        float const F22 = d1z * a2 + d2z * bb2 + d3z * e2;
                        float const C00 = F11 * F22 - F12 * F21;
                        float const C01 = F12 * F20 - F10 * F22;
                        float const C02 = F10 * F21 - F11 * F20;
                        float const C10 = F02 * F21 - F01 * F22;
                        float const C11 = F00 * F22 - F02 * F20;
                        float const C12 = F01 * F20 - F00 * F21;
                        float const C20 = F01 * F12 - F02 * F11;
                        float const C21 = F02 * F10 - F00 * F12;
                        float const C22 = F00 * F11 - F01 * F10;
                        float const J   = F00 * C00 + F01 * C01 + F02 * C02;
                        float const u0 = -(a0 + bb0 + e0), u1 = -(a1 + bb1 + e1), u2 = -(a2 + bb2 + e2);
                        float const Fq0 = F00 * u0 + F01 * u1 + F02 * u2;
                        float const Fq1 = F10 * u0 + F11 * u1 + F12 * u2;
                        float const Fq2 = F20 * u0 + F21 * u1 + F22 * u2;
                        float const Cq0 = C00 * u0 + C01 * u1 + C02 * u2;
                        float const Cq1 = C10 * u0 + C11 * u1 + C12 * u2;
                        float const Cq2 = C20 * u0 + C21 * u1 + C22 * u2;
                        float const sJ  = wlam * (J - alpha);
                        g0 += wmu * Fq0 + sJ * Cq0;
                        g1 += wmu * Fq1 + sJ * Cq1;
                        g2 += wmu * Fq2 + sJ * Cq2;
                        float const t0 = wlam * Cq0, t1 = wlam * Cq1, t2 = wlam * Cq2;
                        h00 += t0 * Cq0;
                        h01 += t0 * Cq1;
                        h02 += t0 * Cq2;
                        h11 += t1 * Cq1;
                        h12 += t1 * Cq2;
                        h22 += t2 * Cq2;
                        hd += wmu * gh2;
                    }
                    for (uint32_t o = (1u << lw) >> 1; o > 0; o >>= 1)
                    {
                        h00 += __shfl_xor_sync(0xffffffffu, h00, o);
                        h01 += __shfl_xor_sync(0xffffffffu, h01, o);
                        h02 += __shfl_xor_sync(0xffffffffu, h02, o);
                        h11 += __shfl_xor_sync(0xffffffffu, h11, o);
                        h12 += __shfl_xor_sync(0xffffffffu, h12, o);
                        h22 += __shfl_xor_sync(0xffffffffu, h22, o);
                        hd += __shfl_xor_sync(0xffffffffu, hd, o);
                        g0 += __shfl_xor_sync(0xffffffffu, g0, o);
                        g1 += __shfl_xor_sync(0xffffffffu, g1, o);
                        g2 += __shfl_xor_sync(0xffffffffu, g2, o);
                    }
                    if (valid && (lane & ((1u << lw) - 1u)) == 0u)
                    {
                        h00 += hd;
                        h11 += hd;
                        h22 += hd;
                        float x = xi.x, y = xi.y, z = xi.z;
                        if constexpr (kDamping)
                        {
                            float4 const xt = __ldcg(p.xt + vi);
                            float const D   = p.dampD;
                            float const ex = x - xt.x, ey = y - xt.y, ez = z - xt.z;
                            g0 += D * (h00 * ex + h01 * ey + h02 * ez);
                            g1 += D * (h01 * ex + h11 * ey + h12 * ez);
                            g2 += D * (h02 * ex + h12 * ey + h22 * ez);
                            float const sc = 1.f + D;
                            h00 *= sc, h01 *= sc, h02 *= sc, h11 *= sc, h12 *= sc, h22 *= sc;
                        }
                        float4 const xm = __ldcg(p.xtildeM + vi);
                        float const K   = xm.w / p.sdt2;
                        h00 += K, h11 += K, h22 += K;
                        g0 += K * (x - xm.x);
                        g1 += K * (y - xm.y);
                        g2 += K * (z - xm.z);
                        float const i00 = h11 * h22 - h12 * h12;
                        float const i01 = h02 * h12 - h01 * h22;
                        float const i02 = h01 * h12 - h02 * h11;
                        float const det = h00 * i00 + h01 * i01 + h02 * i02;
                        if (fabsf(det) > p.detHZero)
                        {
                            float const i11 = h00 * h22 - h02 * h02;
                            float const i12 = h01 * h02 - h00 * h12;
                            float const i22 = h00 * h11 - h01 * h01;
                            float const r   = 1.f / det;
                            x -= r * (i00 * g0 + i01 * g1 + i02 * g2);
                            y -= r * (i01 * g0 + i11 * g1 + i12 * g2);
                            z -= r * (i02 * g0 + i12 * g1 + i22 * g2);
                        }
                        float4 const raw = make_float4(x, y, z, 0.f);
                        if constexpr (kChebyshev)
                        {
                            float4 out = raw;
                            if (k > 1)
                            {
                                float4 const h2 = __ldcg(p.hist + vi);
                                out.x = omega * (x - h2.x) + h2.x;
                                out.y = omega * (y - h2.y) + h2.y;
                                out.z = omega * (z - h2.z) + h2.z;
                            }
                            p.hist[vi]         = make_float4(xi.x, xi.y, xi.z, 0.f);
                            p.pos[vi]          = raw;
                            p.pos[p.pOff + vi] = out;
                        }
                        else
                        {
                            p.pos[vi] = raw;
                        }
                    }
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
