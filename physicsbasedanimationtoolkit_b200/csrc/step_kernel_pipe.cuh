// Pipelined variant of the persistent VBD step kernel (sm_100a): every warp runs a private
// software pipeline over its own tile sequence with asynchronous copies (cp.async / LDGSTS).
//
// All data a tile needs except the vertex positions is static, so it is requested long before it
// is used -- across colour barriers, iterations and substeps:
//   * the descriptor of tile i+2 and
//   * the ring ids of tile i+2 and the incidence records (the HBM stream) of tile i+1
// travel into per-warp shared memory while tile i is computed and while the warp waits at the
// grid barrier.  Positions are mutable, but only the colour being swept changes: the positions of
// tile i+1 that cannot change any more (all of them if it is in the same colour, all but its
// "late" chunks if it is in the next colour) are gathered ahead as well.  After a barrier the only
// exposed memory latency is the L2 gather of the few neighbours of the colour just swept.
// Tiles are dealt round-robin over all warps of the grid (heaviest first), exactly like the direct
// kernel; arithmetic and summation order are the shared ProcessTile.
#pragma once

#include "async_copy.cuh"
#include "step_kernel.cuh"

namespace vbdx {

struct PipeParams {
    StepParams base;
    uint32_t maxIters;  // record buffer capacity per warp, in blocks
    int clusterBarrier; // the grid is ONE thread-block cluster: colours are separated by the hardware cluster barrier
    int dataflow;       // no barrier between colours: every position carries the number of its write in .w, a tile checks
                        // the values it gathered and polls the few that are not there yet (see AwaitTags)
    // the lean barrier-free kernel (step_kernel_flow.cuh) walks a schedule the host laid out (BuildFlowSchedule)
    uint32_t flowActiveWarps;                    // warps that have tiles (the sweep-lag bound of the contact path counts them)
    const uint32_t* __restrict__ flowWarpBegin;  // [grid warps + 1]: range of warp w (w = warp-in-CTA * gridDim.x + CTA) in flowTiles
    const uint4* __restrict__ flowTiles;         // tile descriptors in warp-major order; .w = first entry of the tile in flowIds
    const uint32_t* __restrict__ flowIds;        // pre-decoded ring entries, [group of 4 chunks][lane][4]
};
// per-warp shared memory besides the staging, id and record buffers: 4 tile descriptors, 2 mbarriers, 4 sweep numbers
constexpr uint32_t kPipeWarpFixed = 4 * 16 + 16 + 16;

constexpr int kPipeMaxThreads = 544;  // compiled for up to 16 compute warps + 1 barrier warp per CTA (<= 120 registers)

// shared-memory footprint (host and device must agree)
__host__ __device__ inline size_t PipeSmemBytes(uint32_t nColors, uint32_t warps, uint32_t stageEntries, uint32_t maxIters)
{
    size_t b = 2 * ((static_cast<size_t>(nColors) + 1 + 3) / 4) * 16;  // colour -> first tile table, colour -> first warp of the deal
    b += static_cast<size_t>(warps) * (kPipeWarpFixed + 2 * stageEntries * 4 + stageEntries * 16 + static_cast<size_t>(maxIters) * kBlockBytes);
    return b;
}

// ... of the lean barrier-free kernel (step_kernel_flow.cuh): no id buffers (the ring entries live in registers)
// (+ 512 B: the substep's start positions of the tile's vertices, one slot per lane -- read by the damping / contact epilogue)
__host__ __device__ inline size_t FlowWarpBytes(uint32_t stageEntries, uint32_t maxIters)
{
    return kPipeWarpFixed + 512 + static_cast<size_t>(stageEntries) * 16 + static_cast<size_t>(maxIters) * kBlockBytes;
}
__host__ __device__ inline size_t FlowSmemBytes(uint32_t warps, uint32_t stageEntries, uint32_t maxIters)
{
    return static_cast<size_t>(warps) * FlowWarpBytes(stageEntries, maxIters);
}

// records of the current tile, already in this warp's shared-memory buffer
struct SmemRecords {
    float4 const* rec;
    __device__ __forceinline__ void Fetch(float4& c0, float4& c1)
    {
        c0 = rec[0];
        c1 = rec[32];
        rec += kBlockFloat4;
    }
    __device__ __forceinline__ void Rewind(uint32_t blocks) { rec -= static_cast<size_t>(blocks) * kBlockFloat4; }
};

// Warp-specialised grid barrier.  The compute warps *arrive* (non-blocking), wait only until the
// barrier warp's fence has gone through the load/store path (a fence queued behind a burst of gather
// requests takes microseconds), do their prefetch work in the barrier's shadow, and then wait for the
// release.  The barrier warp does the fence + atomic arrival and polls, undisturbed by that work.
__device__ __forceinline__ void NamedArrive(int id, uint32_t count)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void NamedSync(int id, uint32_t count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
constexpr int kBarArrived = 1, kBarReleased = 2, kBarFenced = 3;

__device__ __forceinline__ unsigned int LoadAcquireSys(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int LoadRelaxedSys(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void StoreReleaseSys(unsigned int* p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void StoreReleaseGpu(unsigned int* p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ float4 LoadPosSys(const float4* q)
{
    unsigned long long lo, hi;
    asm volatile("{\n .reg .b128 t;\n ld.relaxed.sys.global.b128 t, [%2];\n mov.b128 {%0, %1}, t;\n}\n" : "=l"(lo), "=l"(hi) : "l"(q) : "memory");
    return UnpackB128(lo, hi);
}

// One barrier of the barrier warp.  Single GPU: arrive on the counter, poll it.
// Domain decomposition (p.world > 1): halo data synchronises itself (every ghost value carries the tag of its
// write, readers wait for the tag they need), so the colour barrier stays local.  What remains between GPUs is
// a bound on how far a GPU may run ahead, so that it never overwrites a ghost copy a slower peer still reads:
// once the local barrier is complete CTA 0 publishes this GPU's epoch in its neighbours' flag arrays (a relaxed
// store over NVLink, nobody waits for it now), and every CTA waits until its neighbours have finished epoch e - lag
// (lag = 2 inside a substep -- a copy is rewritten nColors + 1 phases after its last reader at the earliest --
// and 0 at the end of a substep, whose pre-step rewrites every ghost).  A peer that never shows up raises
// distError instead of hanging the GPU.
// first half: this CTA's arrival (release: everything its threads wrote is ordered before the arrival)
__device__ __forceinline__ void BarrierSignal(StepParams const& p, unsigned int& target, unsigned int& epoch, uint32_t lane,
                                              unsigned long long* trace)
{
    ++epoch;
    if (lane == 0)
    {
        if (trace)
            trace[2] = GlobalTimer();
        target += gridDim.x;
        AddRelease(p.barrier, 1u);
        if (trace)
            trace[1] = GlobalTimer();
    }
    __syncwarp();
}

// second half: wait for every CTA of this GPU and -- domain decomposition -- for the neighbours' epochs
__device__ __forceinline__ void BarrierAwait(StepParams const& p, unsigned int const& target, unsigned int const& epoch, uint32_t lane,
                                             unsigned int lag, unsigned long long* trace)
{
    unsigned int const e = p.epochBase + epoch;
    if (lane == 0)
    {
        // The neighbours' epochs: one relaxed look before the local wait (e - lag is old news in steady state; the
        // acquire of the local poll below orders what follows) ...
        unsigned int pending = 0;
        if (p.world > 1)
            for (int r = 0; r < p.world; ++r)
                if (((p.peerMask >> r) & 1u) && static_cast<int>(LoadRelaxedSys(p.myFlags + r) - (e - lag)) < 0)
                    pending |= 1u << r;
        while (LoadAcquire(p.barrier) < target)
        {
        }
        // ... every CTA of this GPU has finished the phase: CTA 0 tells the neighbours (flags live in the reader's
        // memory, written over NVLink; nobody waits for this one now) ...
        if (p.world > 1 && blockIdx.x == 0)
            for (int r = 0; r < p.world; ++r)
                if ((p.peerMask >> r) & 1u)
                    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p.peerFlags[r] + p.rank), "r"(e) : "memory");
        // ... and only then a blocking wait for the neighbours that were behind (always the case at the end of a
        // substep, where lag = 0: publishing first is what keeps two GPUs from waiting for each other)
        if (pending != 0u)
        {
            unsigned long long const t0 = GlobalTimer();
            for (int r = 0; r < p.world; ++r)
                if ((pending >> r) & 1u)
                    while (static_cast<int>(LoadAcquireSys(p.myFlags + r) - (e - lag)) < 0)
                        if (GlobalTimer() - t0 > p.distTimeoutNs || LoadAcquire(p.distError) != 0u)  // a peer is missing
                        {
                            atomicExch(p.distError, 1u);
                            break;
                        }
            if (blockIdx.x == 0)
            {
                atomicAdd(p.distStats + 2, 1u);
                atomicAdd(p.distStats + 3, static_cast<unsigned int>(GlobalTimer() - t0));
            }
        }
        if (trace)
            trace[3] = GlobalTimer();
    }
    __syncwarp();
}

__device__ __forceinline__ void BarrierWarpStep(StepParams const& p, unsigned int& target, unsigned int& epoch, uint32_t lane,
                                                unsigned int lag, unsigned long long* trace)
{
    NamedSync(kBarArrived, blockDim.x);  // every compute thread of this CTA is done with the phase
    BarrierSignal(p, target, epoch, lane, trace);
    NamedArrive(kBarFenced, blockDim.x);
    BarrierAwait(p, target, epoch, lane, lag, trace);
    NamedArrive(kBarReleased, blockDim.x);
}

// Hardware barrier of one thread-block cluster, split form: arrive with release (non-blocking), wait with acquire, both
// at cluster scope -- every position a thread wrote before its arrive is visible to every thread of the cluster after
// its wait.  Used instead of the grid barrier when the whole (small) problem is swept by ONE cluster
// (PipeParams::clusterBarrier, VBDX_KERNEL_CLUSTER): ~0.3 us against ~2 us of fence + atomic + poll round trips through L2.
__device__ __forceinline__ void ClusterArrive()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void ClusterWait()
{
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <bool kChebyshev, bool kDamping, bool kStvk = false, bool kDataflow = false>
__global__ void __launch_bounds__(kPipeMaxThreads, 1) StepKernelPipe(const __grid_constant__ PipeParams pp)
{
    extern __shared__ __align__(128) unsigned char smem[];
    StepParams const& p = pp.base;
    uint32_t const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t const nWarps = (blockDim.x >> 5) - 1;  // compute warps; the last warp only runs the grid barrier
    uint32_t const SE   = p.stageEntries;
    uint32_t const nC   = static_cast<uint32_t>(p.nColors);
    uint32_t const gwarp = warp * gridDim.x + blockIdx.x, gWarps = nWarps * gridDim.x;

    uint32_t* const colorTab = reinterpret_cast<uint32_t*>(smem);
    unsigned char* mine      = smem + 2 * ((static_cast<size_t>(nC) + 1 + 3) / 4) * 16 +
                          static_cast<size_t>(warp) * (kPipeWarpFixed + 2 * SE * 4 + SE * 16 + static_cast<size_t>(pp.maxIters) * kBlockBytes);
    float4* const recBuf   = reinterpret_cast<float4*>(mine);
    float4* const stage    = recBuf + static_cast<size_t>(pp.maxIters) * kBlockFloat4;
    uint4* const tdBuf     = reinterpret_cast<uint4*>(stage + SE);
    uint32_t const barRec  = SmemAddr(tdBuf + 4);      // completion of the bulk copy of a tile's records
    uint32_t const barIds  = barRec + 8;                // ... of a tile's ring ids
    uint32_t* const idsBuf = reinterpret_cast<uint32_t*>(tdBuf + 6);

    for (uint32_t c = threadIdx.x; c <= nC; c += blockDim.x)
        colorTab[c] = p.colorTileBegin[c];
    if (lane == 0 && warp < nWarps)
    {
        MbarInit(barRec, 1);
        MbarInit(barIds, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    uint64_t streamPolicy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(streamPolicy));
    uint32_t recFills = 0, idsFills = 0;    // bulk copies issued so far on each barrier
    uint32_t recWaits = 0, idsWaits = 0;    // ... and waited for
    int const totalSweeps = p.substeps * p.iterations;

    if (warp == nWarps)
    {
        // ------------------------------ barrier warp ------------------------------
        unsigned int target = 0, epoch = 0;
        unsigned int const lagIn = nC >= 2u ? 2u : 0u;
        if (pp.clusterBarrier)
        {
            // every thread of the cluster takes part in the hardware barrier: this warp just keeps step
            int const phases = p.substeps * (1 + p.iterations * static_cast<int>(nC));
            for (int ph = 0; ph < phases; ++ph)
            {
                ClusterArrive();
                ClusterWait();
            }
            return;
        }
        if constexpr (kDataflow)
        {
            // barriers only around the pre-step pass: after it, and after the last sweep of the substep
            for (int s = 0; s < p.substeps; ++s)
            {
                BarrierWarpStep(p, target, epoch, lane, 0u, nullptr);
                if (p.iterations > 0)
                    BarrierWarpStep(p, target, epoch, lane, 0u, nullptr);
            }
        }
        else
        {
            for (int s = 0; s < p.substeps; ++s)
            {
                BarrierWarpStep(p, target, epoch, lane, p.iterations > 0 ? lagIn : 0u, nullptr);  // after the pre-step pass
                for (int k = 0; k < p.iterations; ++k)
                    for (uint32_t c = 0; c < nC; ++c)
                    {
                        unsigned long long* tr = nullptr;
                        if (p.trace != nullptr && k == p.traceIteration)
                            tr = p.trace + (static_cast<size_t>(c) * gridDim.x + blockIdx.x) * kTraceStamps;
                        bool const lastOfSubstep = k + 1 == p.iterations && c + 1 == nC;
                        BarrierWarpStep(p, target, epoch, lane, lastOfSubstep ? 0u : lagIn, tr);
                    }
            }
        }
        return;
    }

    // this warp's tile sequence over all sweeps
    struct Cursor {
        int k, c;
        uint32_t T;
        bool valid;
    };
    bool warpHasTiles = false;
    for (uint32_t c = 0; c < nC; ++c)
        warpHasTiles |= colorTab[c] + gwarp < colorTab[c + 1];
    auto Advance = [&](Cursor& t) {
        if (!t.valid)
            return;
        if (t.c >= 0)
        {
            t.T += gWarps;
            if (t.T < colorTab[t.c + 1])
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
            t.T = colorTab[t.c] + gwarp;
            if (t.T < colorTab[t.c + 1])
                return;
        }
    };
    auto IssueTd = [&](Cursor const& t, uint32_t s) {
        if (t.valid && lane == 0)
            CpAsync16(SmemAddr(tdBuf + (s & 3u)), p.tiles + t.T);
    };
    // ring ids of sequence tile s (descriptor already in tdBuf): one bulk copy by the TMA engine
    auto WaitIds = [&]() {
        while (idsWaits < idsFills)
            MbarWait(barIds, idsWaits++ & 1u);
    };
    auto IssueIds = [&](uint32_t s) {
        WaitIds();  // at most one copy in flight per barrier
        uint4 const td       = tdBuf[s & 3u];
        uint32_t const bytes = TileChunks(td.z) * 128u;
        if (lane == 0)
        {
            MbarArriveExpectTx(barIds, bytes);
            BulkLoad(SmemAddr(idsBuf + (s & 1u) * SE), p.ringIds + td.w, bytes, barIds, streamPolicy);
        }
        ++idsFills;
    };
    // incidence records (the HBM stream) of sequence tile s into this warp's record buffer, likewise
    auto WaitRecords = [&]() {
        while (recWaits < recFills)
            MbarWait(barRec, recWaits++ & 1u);
    };
    auto IssueRecords = [&](uint32_t s) {
        WaitRecords();
        uint4 const td       = tdBuf[s & 3u];
        uint32_t const bytes = TileIters(td.z) * kBlockBytes;
        if (lane == 0)
        {
            MbarArriveExpectTx(barRec, bytes);
            BulkLoad(SmemAddr(recBuf), p.records + static_cast<size_t>(td.x) * kBlockFloat4, bytes, barRec, streamPolicy);
        }
        ++recFills;
    };
    // positions of ring chunks [from, to) of sequence tile s (descriptor and ids already in shared memory)
    uint32_t const ghostFrom = p.world > 1 ? p.ghostBegin : 0xffffffffu;  // ids from here on are written by peers
    // tagLow = tag of the sweep before the tile's own (what its higher-colour ghosts must carry; lower-colour
    // ghosts carry tagLow + 1): selects which copy of a ghost is read
    auto IssueGather = [&](uint32_t s, uint32_t from, uint32_t to, uint32_t tagLow) {
        if (from >= to)
            return;
        WaitIds();
        uint32_t const* ids = idsBuf + (s & 1u) * SE + lane;
        uint32_t const dst  = SmemAddr(stage + lane);
        for (uint32_t j = from; j < to; ++j)
        {
            uint32_t const id   = ids[32 * j];
            uint32_t const base = id & ~kPrevFlag;
            bool const prev     = (id & kPrevFlag) != 0u;
            uint32_t const at   = base >= ghostFrom ? GhostIndex(p, base, prev, tagLow + (prev ? 0u : 1u)) : base + (prev ? p.pOff : 0u);
            CpAsync16(dst + 512 * j, p.pos + at);
        }
    };
    // Domain decomposition: the staged ghosts of tile s must carry the tags of the writes this sweep reads; one
    // that has not crossed NVLink yet is polled until it has.
    auto AwaitGhosts = [&](uint32_t s, uint32_t chunks, uint32_t tagLow) {
        uint32_t const* ids = idsBuf + (s & 1u) * SE + lane;
        for (uint32_t j = 0; j < chunks; ++j)
        {
            uint32_t const id   = ids[32 * j];
            uint32_t const base = id & ~kPrevFlag;
            if (base < ghostFrom)
                continue;
            bool const prev       = (id & kPrevFlag) != 0u;
            uint32_t const expect = tagLow + (prev ? 0u : 1u);
            float4 q              = stage[32 * j + lane];
            if (__float_as_uint(q.w) != expect)
            {
                float4 const* src           = p.pos + GhostIndex(p, base, prev, expect);
                unsigned long long const t0 = GlobalTimer();
                for (uint32_t polls = 1;; ++polls)
                {
                    q = LoadPosSys(src);
                    if (__float_as_uint(q.w) == expect)
                        break;
                    // the owner is missing? (looked at rarely: the poll itself must stay one load per round trip)
                    if ((polls & 255u) == 0u && (GlobalTimer() - t0 > p.distTimeoutNs || LoadAcquire(p.distError) != 0u))
                    {
                        atomicExch(p.distError, 1u);
                        break;
                    }
                }
                stage[32 * j + lane] = q;
                atomicAdd(p.distStats, 1u);
                atomicAdd(p.distStats + 1, static_cast<unsigned int>(GlobalTimer() - t0));
            }
        }
        __syncwarp();
    };

    // Barrier-free sweep (pp.dataflow): the staged positions of tile s must carry the numbers of the writes this sweep
    // reads -- tagLow + 1 (this iteration) for neighbours of a lower colour, tagLow (the previous iteration, or the
    // pre-step) for neighbours of a higher colour and for the tile's own start values.  A value that is not there yet
    // is polled.  Dependencies are exactly the 1-rings, every warp walks its tiles in (iteration, colour) order and all
    // warps are co-resident, so the earliest unfinished tile can always proceed; and a vertex cannot be overwritten
    // before its readers have read it, because its next update needs THEIR results first.
    bool dfDead = false;  // a dependency timed out (reported to the host): stop waiting, just finish
    auto AwaitTags = [&](uint32_t s, uint32_t chunks, uint32_t tagLow, uint32_t vbase) {
        uint32_t const* ids = idsBuf + (s & 1u) * SE + lane;
        if (!dfDead && LoadAcquire(p.distError) != 0u)
            dfDead = true;
        if (dfDead)
            return;
        for (uint32_t j = 0; j < chunks; ++j)
        {
            uint32_t const id   = ids[32 * j];
            uint32_t const base = id & ~kPrevFlag;
            bool const prev     = (id & kPrevFlag) != 0u;
            if (base >= p.activeEnd || (base == vbase && !prev))  // constrained vertices never change; padding entries
                continue;
            uint32_t const expect = tagLow + (prev ? 0u : 1u);
            float4 q              = stage[32 * j + lane];
            if (__float_as_uint(q.w) != expect)
            {
                float4 const* src           = p.pos + base + (prev ? p.pOff : 0u);
                unsigned long long const t0 = GlobalTimer();
                for (uint32_t polls = 1;; ++polls)
                {
                    q = LoadPosGpu(src);
                    if (__float_as_uint(q.w) == expect)
                        break;
                    if ((polls & 1023u) == 0u && (GlobalTimer() - t0 > p.distTimeoutNs || LoadAcquire(p.distError) != 0u))
                    {
                        // a bug, not a slow peer: never hang the GPU; leave what was being waited for to the host
                        if (atomicCAS(p.distError, 0u, 2u) == 0u)
                        {
                            p.distError[1] = base | (prev ? kPrevFlag : 0u);
                            p.distError[2] = expect;
                            p.distError[3] = __float_as_uint(q.w);
                            p.distError[4] = vbase;
                            p.distError[5] = tagLow - p.tagBase;
                        }
                        dfDead = true;
                        break;
                    }
                }
                stage[32 * j + lane] = q;
            }
        }
        __syncwarp();
    };

    Cursor c1{0, -1, 0, totalSweeps > 0 && warpHasTiles}, c2;
    Advance(c1);  // tile 0
    Cursor const c0 = c1;
    Advance(c1);  // tile 1
    c2 = c1;
    Advance(c2);  // tile 2
    IssueTd(c0, 0);
    IssueTd(c1, 1);
    CpAsyncWaitAll();
    __syncwarp();
    if (c0.valid)
    {
        IssueIds(0);
        IssueRecords(0);
    }
    if (c1.valid)
        IssueIds(1);
    uint32_t seq      = 0;
    uint32_t gathered = 0;  // ring chunks of tile `seq` whose positions have already been requested
    bool deferred     = false;  // static data of tile `seq` still has to be requested
    Cursor cUp        = c0;     // cursor of tile `seq`

    uint32_t const gtid    = blockIdx.x * (nWarps * 32) + threadIdx.x;
    uint32_t const gstride = gridDim.x * (nWarps * 32);
    for (int s = 0; s < p.substeps; ++s)
    {
        if (!p.skipPreStep)
            for (uint32_t i = gtid; i < p.ghostBegin; i += gstride)
                PreStepVertex<kChebyshev>(p, i, s);
        if (pp.clusterBarrier)
        {
            __syncwarp();
            ClusterArrive();
            ClusterWait();
        }
        else
        {
            NamedArrive(kBarArrived, blockDim.x);
            NamedSync(kBarFenced, blockDim.x);
            NamedSync(kBarReleased, blockDim.x);
        }

        for (int k = 0; k < p.iterations; ++k)
        {
            float const omega = kChebyshev ? __ldg(p.omega + p.iterBegin + k) : 1.f;
            // tags (domain decomposition): the pre-step wrote tagBase + s (iterations + 1), iteration k writes that + k + 1
            uint32_t const tagLow = p.tagBase + static_cast<uint32_t>(s * (p.iterations + 1) + k);
            for (uint32_t c = 0; c < nC; ++c)
            {
                unsigned long long* tr = nullptr;
                if (p.trace != nullptr && k == p.traceIteration)
                {
                    tr = p.trace + (static_cast<size_t>(c) * gridDim.x + blockIdx.x) * kTraceStamps;
                    if (threadIdx.x == 0)
                        tr[0] = GlobalTimer();
                }
                uint32_t const tBegin = colorTab[c], tEnd = colorTab[c + 1];
                for (uint32_t T = tBegin + gwarp; T < tEnd; T += gWarps)
                {
                    unsigned long long* tr0 = (tr && warp == 0 && T == tBegin + gwarp) ? tr : nullptr;
                    // requested one tile ago: descriptor i+1, and ids + records of this tile
                    CpAsyncWaitAll();
                    __syncwarp();
                    uint4 const td = tdBuf[seq & 3u];
                    if (tr0 && lane == 0)
                        tr0[4] = GlobalTimer();
                    uint32_t const chunks = TileChunks(td.z);
                    if (gathered < chunks)
                        IssueGather(seq, gathered, chunks, tagLow);  // what could not be requested before the barrier
                    CpAsyncCommit();
                    IssueTd(c2, seq + 2);
                    CpAsyncCommit();
                    CpAsyncWaitGroup<1>();  // positions have landed; the descriptor may still be in flight
                    __syncwarp();
                    if (ghostFrom != 0xffffffffu && TileReadsGhosts(td.z))
                        AwaitGhosts(seq, chunks, tagLow);
                    if constexpr (kDataflow)
                        AwaitTags(seq, chunks, tagLow, td.y);
                    // Next tile in this same colour sweep (multi-round colours): all its positions are final, so
                    // everything it needs is requested as soon as this tile's buffers are free.  Otherwise the
                    // requests are issued in the shadow of the grid barrier (after GridArrive below).
                    bool const nextValid = c1.valid;
                    // (barrier-free mode: every next tile is requested at once; what was not final yet is re-read above)
                    bool const sameSweep = nextValid && (kDataflow || (c1.k == s * p.iterations + k && c1.c == static_cast<int>(c)));
                    uint32_t const seqNext = seq + 1;
                    uint32_t nextGathered  = 0;
                    auto prefetchNext = [&]() {
                        if (!sameSweep)
                            return;
                        __syncwarp();  // this tile's records, ids and staged positions are consumed
                        CpAsyncWaitAll();  // descriptors i+1 (requested a tile ago) and i+2
                        __syncwarp();
                        IssueRecords(seqNext);
                        nextGathered = TileChunks(tdBuf[seqNext & 3u].z);
                        // (the next tile may belong to a later sweep in barrier-free mode: its own write numbers select the ghost copies)
                        IssueGather(seqNext, 0, nextGathered, kDataflow ? p.tagBase + static_cast<uint32_t>(c1.k + c1.k / p.iterations) : tagLow);
                        if (c2.valid)
                            IssueIds(seq + 2);
                        CpAsyncCommit();
                    };
                    deferred = nextValid && !sameSweep;
                    WaitRecords();
                    SmemRecords src{recBuf + lane};
                    ProcessTile<kChebyshev, kDamping, false, SmemRecords, decltype(prefetchNext), kStvk>(p, td, stage, src, static_cast<int>(c), p.iterBegin + k, omega, lane, tr0, prefetchNext,
                                                                                                          tagLow + 1u);
                    if (tr0 && lane == 0)
                        tr0[7] = GlobalTimer();
                    gathered = nextGathered;
                    cUp = c1;  // cursor of the upcoming tile
                    c1 = c2;
                    Advance(c2);
                    ++seq;
                }
                if (kDataflow && !(k + 1 == p.iterations && c + 1 == nC))
                    continue;  // no barrier between colours
                if (pp.clusterBarrier)
                {
                    __syncwarp();
                    ClusterArrive();
                }
                else
                {
                    NamedArrive(kBarArrived, blockDim.x);
                    NamedSync(kBarFenced, blockDim.x);  // the barrier warp's fence is through
                }
                if (deferred)
                {
                    // In the shadow of the barrier: records and ids of the upcoming tile, plus the positions
                    // that cannot change any more.  If that tile belongs to the immediately following colour
                    // of this substep these are its "early" chunks (everything but neighbours of the colour
                    // just swept); if it lies further ahead, nothing.
                    int const sweep = s * p.iterations + k;
                    bool const nextPhase = (cUp.k == sweep && cUp.c == static_cast<int>(c) + 1) ||
                                           (cUp.k == sweep + 1 && cUp.c == 0 && c + 1 == nC && (sweep + 1) % p.iterations != 0);
                    CpAsyncWaitAll();  // descriptors of tiles seq and seq+1
                    __syncwarp();
                    bool const stamp = tr && threadIdx.x == 0;
                    if (stamp)
                        tr[8] = GlobalTimer();
                    IssueRecords(seq);
                    if (stamp)
                        tr[9] = GlobalTimer();
                    gathered = nextPhase ? TileEarlyChunks(tdBuf[seq & 3u].z) : 0u;
                    IssueGather(seq, 0, gathered, p.tagBase + static_cast<uint32_t>(cUp.k + cUp.k / p.iterations));
                    if (stamp)
                        tr[10] = GlobalTimer();
                    if (c1.valid)
                        IssueIds(seq + 1);
                    CpAsyncCommit();
                    if (stamp)
                        tr[11] = GlobalTimer();
                    deferred = false;
                }
                if (pp.clusterBarrier)
                {
                    __syncwarp();
                    ClusterWait();
                }
                else
                    NamedSync(kBarReleased, blockDim.x);
            }
        }
    }
    if (!p.skipPostStep)
        for (uint32_t i = gtid; i < p.ghostBegin; i += gstride)
            PostStepVertex(p, i);
}

}  // namespace vbdx
