// Warp-specialised variant of the persistent VBD step kernel (sm_100a).
//
// One CTA per SM.  The last warp is a *producer*: its lanes stream the CTA's
// incidence-record blocks (1 KB each, static rest data) from HBM into a shared-memory ring with
// 1-D bulk asynchronous copies (cp.async.bulk, i.e. the TMA engine; SASS UBLKCP) that signal
// per-slot "full" mbarriers.  Because the records never change, the producer runs ahead of the
// colour barriers: while the consumers wait for the other SMs at the end of colour c, the
// records of colour c+1 are already landing in shared memory, so the HBM stream is continuous.
//
// The remaining warps are *consumers*.  Each owns a static subset of the CTA's warp tiles and
// runs a two-deep software pipeline over its own tile sequence (which crosses colour and
// iteration boundaries): the descriptor of tile i+3 and the ring ids of tile i+2 are fetched with
// cp.async (LDGSTS) into per-warp shared memory while tile i is processed, and so are the positions
// of tile i+1 when it belongs to the same colour.  When a colour barrier releases, everything static
// a warp needs is therefore already on chip and the only exposed latency is the gather of the
// (mutable) positions of its first tile from L2.
// Arithmetic and summation order are those of step_kernel.cuh (shared ProcessTile).
#pragma once

#include "async_copy.cuh"
#include "step_kernel.cuh"

namespace vbdx {

constexpr int kTmaMaxThreads   = 640;  // <= 102 registers per thread
constexpr int kProducerWarps   = 1;    // one producer warp; each of its lanes issues every 32nd block of the stream

struct TmaParams {
    StepParams base;
    const uint32_t* __restrict__ ctaBlockBegin;  // [nColors][gridDim.x + 1] first record block per CTA
    uint32_t ringSlots;                          // R: ring capacity in blocks
};

__device__ __forceinline__ uint32_t LoadAcquireShared(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ void StoreReleaseShared(uint32_t addr, uint32_t v)
{
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Record source of the consumers: this lane's 32-byte record out of the shared-memory ring slot
// that the producer filled; the slot is handed back as soon as the warp has copied its records to
// registers.  mbarrier phases only tell adjacent fills of a slot apart, and consumers do not
// consume in stream order, so a consumer first makes sure the producer has *issued* block n
// (monotonic `produced` counter) before it waits on the slot's parity.
struct RingRecords {
    unsigned char const* smem;
    uint32_t full, empty, produced, R, slot, fill, n, lane;
    __device__ __forceinline__ void Fetch(float4& c0, float4& c1)
    {
        while (LoadAcquireShared(produced) <= n)
        {
        }
        MbarWait(full + 8 * slot, fill & 1u);
        unsigned char const* blk = smem + slot * kBlockBytes + lane * 16;
        c0 = *reinterpret_cast<float4 const*>(blk);
        c1 = *reinterpret_cast<float4 const*>(blk + 512);
        __syncwarp();
        if (lane == 0)
            MbarArrive(empty + 8 * slot);
        ++n;
        if (++slot == R)
        {
            slot = 0;
            ++fill;
        }
    }
};

// grid barrier among the consumer threads of all CTAs (the producer warp never joins)
__device__ __forceinline__ void ConsumerGridBarrier(unsigned int* counter, unsigned int& target, uint32_t nConsumerThreads,
                                                    unsigned long long* trace = nullptr)
{
    asm volatile("bar.sync 1, %0;" ::"r"(nConsumerThreads) : "memory");
    if (threadIdx.x == 0)
    {
        if (trace)
            trace[2] = GlobalTimer();
        target += gridDim.x;
        AddRelease(counter, 1u);
        while (LoadAcquire(counter) < target)
        {
        }
        __threadfence();
        if (trace)
            trace[3] = GlobalTimer();
    }
    asm volatile("bar.sync 1, %0;" ::"r"(nConsumerThreads) : "memory");
}

// position of a consumer warp in its own tile sequence
struct TileCursor {
    int k, c;       // iteration within the substep sequence (s * iterations + k), colour
    uint32_t T;     // tile index
    bool valid;
};

// shared-memory footprint of the TMA kernel (host and device must agree)
__host__ __device__ inline size_t TmaSmemBytes(uint32_t R, uint32_t nColors, uint32_t nConsumerWarps, uint32_t stageEntries)
{
    size_t b = static_cast<size_t>(R) * (kBlockBytes + 16);          // ring + full/empty barriers
    b += 16;                                                          // produced counter (+pad)
    b += static_cast<size_t>(nColors + 1) * 16;                       // range table
    b += static_cast<size_t>(nConsumerWarps) * (4 * 16);              // tile-descriptor ring per warp
    b += static_cast<size_t>(nConsumerWarps) * stageEntries * (4 * 4 + 2 * 16);  // ring of 4 id lists + 2 position buffers
    return b;
}

template <bool kChebyshev, bool kDamping>
__global__ void __launch_bounds__(kTmaMaxThreads, 1) StepKernelTma(const __grid_constant__ TmaParams tp)
{
    extern __shared__ __align__(128) unsigned char smem[];
    StepParams const& p  = tp.base;
    uint32_t const R     = tp.ringSlots;
    uint32_t const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t const NC   = (blockDim.x >> 5) - kProducerWarps;  // consumer warps
    uint32_t const SE   = p.stageEntries;
    uint32_t const nC   = static_cast<uint32_t>(p.nColors);

    // carve shared memory
    uint32_t const ring     = SmemAddr(smem);
    uint32_t const full     = ring + R * kBlockBytes;
    uint32_t const empty    = full + R * 8;
    uint32_t const produced = empty + R * 8;
    unsigned char* cur      = smem + static_cast<size_t>(R) * (kBlockBytes + 16) + 16;
    uint4* const rangeTab   = reinterpret_cast<uint4*>(cur);  // per colour: {tBegin, tEnd, b0, blocks before this colour in one sweep}
    cur += static_cast<size_t>(nC + 1) * 16;
    uint4* const tdBuf = reinterpret_cast<uint4*>(cur) + warp * 4;
    cur += static_cast<size_t>(NC) * 64;
    float4* const stage = reinterpret_cast<float4*>(cur) + static_cast<size_t>(warp) * 2 * SE;
    cur += static_cast<size_t>(NC) * 2 * SE * 16;
    uint32_t* const idsBuf = reinterpret_cast<uint32_t*>(cur) + static_cast<size_t>(warp) * 4 * SE;

    uint32_t const stride    = gridDim.x + 1;
    uint32_t const* blkBegin = tp.ctaBlockBegin + blockIdx.x;

    if (threadIdx.x == 0)
    {
        for (uint32_t s = 0; s < R; ++s)
        {
            MbarInit(full + 8 * s, 1);
            MbarInit(empty + 8 * s, 1);
        }
        *reinterpret_cast<volatile uint32_t*>(smem + static_cast<size_t>(R) * (kBlockBytes + 16)) = 0u;
        uint32_t before = 0;
        for (uint32_t c = 0; c < nC; ++c)
        {
            uint32_t const* range = p.ctaTileRange + static_cast<size_t>(c) * stride + blockIdx.x;
            uint32_t const b0 = blkBegin[c * stride], b1 = blkBegin[c * stride + 1];
            rangeTab[c] = make_uint4(range[0], range[1], b0, before);
            before += b1 - b0;
        }
        rangeTab[nC] = make_uint4(0, 0, 0, before);  // .w = record blocks of this CTA per sweep
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    uint32_t const blocksPerSweep = rangeTab[nC].w;
    int const totalSweeps         = p.substeps * p.iterations;

    if (warp >= NC)
    {
        // ------------------------------ producer warp ------------------------------
        // Lane l issues blocks l, l+L, l+2L, ... of this CTA's stream (L = min(32, R) lanes), so L bulk
        // copies are in flight per pass of the loop and the per-block cost of the mbarrier handshake is
        // amortised over the warp.  Fills of one slot are issued in stream order by construction.
        uint32_t const L     = R < 32u ? R : 32u;
        uint32_t const total = static_cast<uint32_t>(totalSweeps) * blocksPerSweep;
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        uint32_t c = 0, off = lane;  // position inside the sweep: colour and block offset within it
        uint32_t slot = lane % R, fill = lane / R;
        for (uint32_t base = 0; base < total; base += L)
        {
            uint32_t const n = base + lane;
            if (lane < L && n < total)
            {
                for (;;)
                {
                    uint32_t const inColour = (c + 1 < nC ? rangeTab[c + 1].w : blocksPerSweep) - rangeTab[c].w;
                    if (off < inColour)
                        break;
                    off -= inColour;
                    if (++c == nC)
                        c = 0;
                }
                if (fill > 0)
                    MbarWaitDivergent(empty + 8 * slot, (fill & 1u) ^ 1u);
                MbarArriveExpectTx(full + 8 * slot, kBlockBytes);
                BulkLoad(ring + slot * kBlockBytes, p.records + static_cast<size_t>(rangeTab[c].z + off) * kBlockFloat4,
                         kBlockBytes, full + 8 * slot, policy);
                off += L;
                slot += L;
                while (slot >= R)
                {
                    slot -= R;
                    ++fill;
                }
            }
            __syncwarp();
            if (lane == 0)
                StoreReleaseShared(produced, base + L < total ? base + L : total);
        }
        return;
    }

    // ------------------------------ consumers ------------------------------
    uint32_t const nConsumerThreads = NC * 32;
    unsigned int target    = 0;
    uint32_t const ctid    = blockIdx.x * nConsumerThreads + threadIdx.x;
    uint32_t const cstride = gridDim.x * nConsumerThreads;

    // this warp's tile sequence over all sweeps; `Advance` steps to its next tile
    auto First = [&](TileCursor& t) {
        t.k = 0;
        t.c = -1;
        t.T = 0;
        t.valid = totalSweeps > 0;
    };
    auto Advance = [&](TileCursor& t) {
        if (!t.valid)
            return;
        if (t.c >= 0)
        {
            t.T += NC;
            if (t.T < rangeTab[t.c].y)
                return;
        }
        for (;;)
        {
            if (++t.c == static_cast<int>(nC))
            {
                t.c = 0;
                if (++t.k == totalSweeps)
                {
                    t.valid = false;
                    return;
                }
            }
            t.T = rangeTab[t.c].x + warp;
            if (t.T < rangeTab[t.c].y)
                return;
            if (blocksPerSweep == 0 && t.c == static_cast<int>(nC) - 1)
            {
                t.valid = false;  // this CTA owns no tiles at all
                return;
            }
        }
    };
    // a warp with no tiles in any colour must not spin forever in Advance
    bool warpHasTiles = false;
    for (uint32_t c = 0; c < nC; ++c)
        warpHasTiles |= rangeTab[c].x + warp < rangeTab[c].y;

    // Software pipeline over this warp's tile sequence (tile i = the one being processed):
    //   descriptor of tile i+3, ring ids of tile i+2 and -- when no colour barrier lies in between --
    //   the position gather of tile i+1 are in flight (cp.async) while tile i is computed.
    TileCursor c1, c2, c3;
    First(c1);
    c1.valid &= warpHasTiles;
    Advance(c1);  // tile 0
    TileCursor const c0 = c1;
    Advance(c1);  // tile 1
    c2 = c1;
    Advance(c2);  // tile 2
    c3 = c2;
    Advance(c3);  // tile 3
    uint32_t seq = 0;  // index of the tile being processed in this warp's sequence
    auto IssueTd = [&](TileCursor const& t, uint32_t s) {
        if (t.valid && lane == 0)
            CpAsync16(SmemAddr(tdBuf + (s & 3u)), p.tiles + t.T);
    };
    auto IssueIds = [&](uint32_t s) {
        // ring ids of sequence tile s, whose descriptor is already in tdBuf
        uint4 const td        = tdBuf[s & 3u];
        uint32_t const chunks = TileChunks(td.z);
        uint32_t const dst    = SmemAddr(idsBuf + (s & 3u) * SE + lane);
        for (uint32_t j = 0; j < chunks; ++j)
            CpAsync4(dst + 128 * j, p.ringIds + td.w + 32 * j + lane);
    };
    auto IssueGather = [&](uint32_t s) {
        // positions of sequence tile s (own vertices + 1-rings): descriptor and ids already in shared memory
        uint4 const td        = tdBuf[s & 3u];
        uint32_t const chunks = TileChunks(td.z);
        uint32_t const* ids   = idsBuf + (s & 3u) * SE + lane;
        uint32_t const dst    = SmemAddr(stage + (s & 1u) * SE + lane);
        for (uint32_t j = 0; j < chunks; ++j)
        {
            uint32_t const id = ids[32 * j];
            CpAsync16(dst + 512 * j, p.pos + (id & ~kPrevFlag) + ((id & kPrevFlag) ? p.pOff : 0u));
        }
    };
    // prologue: descriptors of tiles 0..2, then the ids of tiles 0 and 1
    IssueTd(c0, 0);
    IssueTd(c1, 1);
    IssueTd(c2, 2);
    CpAsyncWaitAll();
    __syncwarp();
    if (c0.valid)
        IssueIds(0);
    if (c1.valid)
        IssueIds(1);
    bool gathered = false;  // has the gather of tile `seq` been issued already

    for (int s = 0; s < p.substeps; ++s)
    {
        if (!p.skipPreStep)
            for (uint32_t i = ctid; i < p.ghostBegin; i += cstride)
                PreStepVertex<kChebyshev>(p, i, s);
        ConsumerGridBarrier(p.barrier, target, nConsumerThreads);

        for (int k = 0; k < p.iterations; ++k)
        {
            float const omega        = kChebyshev ? __ldg(p.omega + p.iterBegin + k) : 1.f;
            uint32_t const sweepBase = static_cast<uint32_t>(s * p.iterations + k) * blocksPerSweep;
            for (uint32_t c = 0; c < nC; ++c)
            {
                uint4 const rt = rangeTab[c];
                unsigned long long* tr = nullptr;
                if (p.trace != nullptr && k == p.traceIteration)
                {
                    tr = p.trace + (static_cast<size_t>(c) * gridDim.x + blockIdx.x) * kTraceStamps;
                    if (threadIdx.x == 0)
                        tr[0] = GlobalTimer();
                }
                for (uint32_t T = rt.x + warp; T < rt.y; T += NC)
                {
                    unsigned long long* tr0 = (tr && warp == 0 && T == rt.x) ? tr : nullptr;
                    // everything requested one tile ago has landed: descriptor i+2, ids i+1, gather i (if issued)
                    CpAsyncWaitAll();
                    __syncwarp();
                    uint4 const td = tdBuf[seq & 3u];
                    if (tr0 && lane == 0)
                        tr0[4] = GlobalTimer();
                    bool const wasGathered = gathered;
                    if (!wasGathered)
                    {
                        IssueGather(seq);  // first tile after a barrier: cannot be requested earlier
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    }
                    if (c2.valid)
                        IssueIds(seq + 2);
                    IssueTd(c3, seq + 3);
                    // the next tile's positions may be gathered now iff it belongs to this same colour sweep
                    gathered = c1.valid && c1.k == s * p.iterations + k && c1.c == static_cast<int>(c);
                    if (gathered)
                        IssueGather(seq + 1);
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    c1 = c2;
                    c2 = c3;
                    Advance(c3);
                    if (!wasGathered)
                    {
                        asm volatile("cp.async.wait_group 1;" ::: "memory");
                        __syncwarp();
                    }
                    uint32_t const n0 = sweepBase + rt.w + (td.x - rt.z);
                    RingRecords src{smem, full, empty, produced, R, n0 % R, n0 / R, n0, lane};
                    ProcessTile<kChebyshev, kDamping, false>(p, td, stage + (seq & 1u) * SE, src, static_cast<int>(c), p.iterBegin + k, omega, lane, tr0);
                    if (tr0 && lane == 0)
                        tr0[7] = GlobalTimer();
                    ++seq;
                }
                if (tr && threadIdx.x == 0)
                    tr[1] = GlobalTimer();
                ConsumerGridBarrier(p.barrier, target, nConsumerThreads, tr);
            }
        }
    }
    if (!p.skipPostStep)
        for (uint32_t i = ctid; i < p.ghostBegin; i += cstride)
            PostStepVertex(p, i);
}

}  // namespace vbdx
