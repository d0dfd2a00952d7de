// The barrier-free VBD step kernel for sm_100a, written around ONE observation (profiles/r02*): a warp executes its tiles
// one after the other at about one instruction per 8 cycles, on a mesh of a million tets it owns about one tile per colour,
// and all 16 warps of an SM do the same thing at the same time -- so a step lasts (iterations x colours) x (instructions per
// tile) x 8 cycles, and the SM's issue slots are half full.  Everything that is not arithmetic was therefore cut down to a
// few instructions per tile or moved to the host:
//
//   * the tile sequence of every warp is laid out by the host (FlowSchedule: descriptors in warp-major order), the kernel
//     just walks it; descriptors arrive by cp.async three tiles ahead in a 4-slot shared-memory ring;
//   * ring entries are pre-decoded by the host (FlowIds): final index into the position buffers, "previous iterate",
//     "never changes" and "ghost" bits, four chunks of a lane packed in one 16-byte word -- a tile's ids live in eight
//     registers (two LDG.128 issued a tile ahead), no id buffers, no mbarrier for them;
//   * positions are requested (cp.async) right AFTER the tile's store -- the store is what other warps wait for --, the
//     staged values are checked against the write number they must carry and only stale ones are re-read, all of a lane's
//     re-reads in flight together;
//   * incidence records arrive by one TMA bulk copy per tile, requested by one lane as soon as the previous tile's
//     accumulation is done;
//   * no barrier warp (16 warps = 128 registers per thread, no spills); the two grid barriers per substep run on warp 0.
//
// Arithmetic, summation order and protocol (write numbers in .w; ghost copies by parity under domain
// decomposition) are those of StepKernelPipe<.., dataflow>: results are bit-identical to it and to the barrier
// kernels (ProcessTile is shared).
//
// Replaces (from scratch) the reference's per-colour launch pair gpu/impl/vbd/Kernels.cuh:148-233 +
// gpu/impl/vbd/Integrator.cu:303-327; arithmetic per sim/vbd/Integrator.cpp:98-136 (see step_kernel.cuh).
#pragma once

#include "step_kernel_pipe.cuh"

namespace vbdx {

constexpr int kFlowMaxThreads = 512;  // 16 warps: a register allocation of 128 per thread (17 warps round up to 20: 96)

// FlowIds entry (host: BuildFlowSchedule): where the value lives and what to expect of it
constexpr uint32_t kFlowPrev    = 0x80000000u;  // previous-iterate buffer P: must carry the previous sweep's write number
constexpr uint32_t kFlowStatic  = 0x40000000u;  // never changes (constrained vertex, padding): not checked
constexpr uint32_t kFlowGhost   = 0x20000000u;  // written by a peer GPU; index = ghost number (two copies by parity of the write)
constexpr uint32_t kFlowIndex   = 0x1fffffffu;

// kDist: domain decomposition (ghost entries exist); kStvk: St. Venant-Kirchhoff records (two blocks per incident tet)
template <bool kChebyshev, bool kDamping, bool kDist, bool kStvk = false>
__global__ void __launch_bounds__(kFlowMaxThreads, 1) StepKernelFlow(const __grid_constant__ PipeParams pp)
{
    extern __shared__ __align__(128) unsigned char smem[];
    StepParams const& p = pp.base;
    uint32_t const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t const nWarps = blockDim.x >> 5;  // every warp computes; warp 0 also runs the two grid barriers of a substep
    uint32_t const SE     = p.stageEntries;
    uint32_t const nC     = static_cast<uint32_t>(p.nColors);
    uint32_t const gwarp  = warp * gridDim.x + blockIdx.x;

    // per warp (FlowWarpBytes): record buffer | staged positions | 4 descriptors | mbarrier | 4 sweep numbers | 32 x xt
    unsigned char* mine   = smem + static_cast<size_t>(warp) * FlowWarpBytes(SE, pp.maxIters);
    float4* const recBuf  = reinterpret_cast<float4*>(mine);
    float4* const stage   = recBuf + static_cast<size_t>(pp.maxIters) * kBlockFloat4;
    uint4* const tdRing   = reinterpret_cast<uint4*>(stage + SE);  // descriptors of tiles seq .. seq+3 of this warp's sequence
    uint32_t const barRec = SmemAddr(tdRing + 4);                   // completion of the bulk copy of a tile's records
    int* const kkRing     = reinterpret_cast<int*>(tdRing + 5);     // their sweep numbers within the substep (-1: no tile)
    float4* const xtStage = reinterpret_cast<float4*>(tdRing + 6);  // xt of the tile's vertices, per lane (damping / contact)

    if (lane == 0)
    {
        MbarInit(barRec, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // grid barriers exist only around the pre-step pass (after it, and after the last sweep of the substep)
    unsigned int target = 0, epoch = 0;
    auto GridSync = [&]() {
        __syncthreads();
        if (warp == 0)
        {
            BarrierSignal(p, target, epoch, lane, nullptr);
            BarrierAwait(p, target, epoch, lane, 0u, nullptr);
        }
        __syncthreads();
    };

    uint64_t streamPolicy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(streamPolicy));
    uint32_t recWaits = 0;  // completed waits on barRec (phase parity)
    uint32_t seq      = 0;  // tiles processed so far: selects the ring slot
    int const I       = p.iterations;

    // this warp's tiles of one sweep, in the order it runs them (colour-major), laid out by the host
    uint32_t const myBegin = __ldg(pp.flowWarpBegin + gwarp), myCount = __ldg(pp.flowWarpBegin + gwarp + 1) - myBegin;
    uint4 const* const myTiles = pp.flowTiles + myBegin;
    struct Gen {
        uint32_t j;  // next tile of the sweep
        int kk;      // sweep within the substep
    };
    // request the descriptor of the sequence's next tile into ring slot `slot` (cp.async: it arrives with the next gather)
    auto Fetch = [&](Gen& g, uint32_t slot) {
        bool const any = myCount != 0u && g.kk < I;
        if (lane == 0)
        {
            if (any)
                CpAsync16(SmemAddr(tdRing + slot), myTiles + g.j);
            kkRing[slot] = any ? g.kk : -1;
        }
        if (any && ++g.j == myCount)
        {
            g.j = 0;
            ++g.kk;
        }
    };
    auto IssueRecords = [&](uint4 const td) {
        if (lane == 0)
        {
            uint32_t const bytes = TileIters(td.z) * kBlockBytes;
            MbarArriveExpectTx(barRec, bytes);
            BulkLoad(SmemAddr(recBuf), p.records + static_cast<size_t>(td.x) * kBlockFloat4, bytes, barRec, streamPolicy);
        }
    };
    // a tile's pre-decoded ring entries: chunks 4g .. 4g+3 of this lane are one 16-byte word
    auto LoadIds = [&](uint4 const td, uint32_t g) {
        uint4 v;
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(reinterpret_cast<uint4 const*>(pp.flowIds + td.w) + g * 32u + lane));
        return v;
    };
    // index into p.pos of the value an entry names (a ghost has one copy per parity of the write number it must carry)
    auto Index = [&](uint32_t e, uint32_t tagLow) {
        uint32_t at = e & kFlowIndex;
        if constexpr (kDist)
        {
            uint32_t const prev = e >> 31;
            uint32_t const odd  = (tagLow + 1u - prev) & 1u;
            uint32_t const off  = (prev && p.pOff != 0u) ? (odd ? p.nGhost : p.pOff) : 0u;
            at                  = (e & kFlowGhost) ? (odd ? p.ghostExt : p.ghostBegin) + off + at : at;
        }
        return at;
    };
    auto Gather = [&](uint4 const td, uint4 const a, uint4 const b, uint32_t tagLow) {
        uint32_t const dst    = SmemAddr(stage + lane);
        uint32_t const chunks = TileChunks(td.z);
        if constexpr (kDamping)
        {
            // the epilogue's xt (constant during the sweeps): staged with the ring instead of an L2 round trip on the tile's
            // critical path or four registers held across the accumulation loop
            uint32_t const g = lane >> TileLog2W(td.z);
            CpAsync16(SmemAddr(xtStage + lane), p.xt + td.y + (g < TileVerts(td.z) ? g : 0u));
        }
        uint32_t const e[8]   = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (uint32_t u = 0; u < 8; ++u)
            if (u < chunks)
                CpAsync16(dst + 512u * u, p.pos + Index(e[u], tagLow));
        for (uint32_t j0 = 8; j0 < chunks; j0 += 4)  // rings of more than 256 entries (rare): ids straight from global memory
        {
            uint4 const c        = LoadIds(td, j0 >> 2);
            uint32_t const f[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u)
                if (j0 + u < chunks)
                    CpAsync16(dst + 512u * (j0 + u), p.pos + Index(f[u], tagLow));
        }
    };
    // The staged values must carry the numbers of the writes this sweep reads: tagLow + 1 (this sweep) for neighbours of
    // a lower colour, tagLow (the previous sweep, or the pre-step) for neighbours of a higher colour and the tile's own
    // start values.  What is not there yet is re-read, all of a lane's stale entries in flight together.
    bool dead = false;  // a wait timed out (reported to the host): stop waiting, just finish
    auto EntryOf = [&](uint4 const td, uint4 const a, uint4 const b, uint32_t j) {
        uint32_t const e[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t r          = e[0];
#pragma unroll
        for (uint32_t u = 1; u < 8; ++u)
            r = j == u ? e[u] : r;
        if (j >= 8u)
            r = __ldg(pp.flowIds + td.w + (j >> 2) * 128u + lane * 4u + (j & 3u));
        return r;
    };
    auto Await = [&](uint4 const td, uint4 const a, uint4 const b, uint32_t tagLow) {
        if (dead)
            return;
        uint32_t const chunks = TileChunks(td.z);
        uint32_t const e[8]   = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t stale        = 0u;
#pragma unroll
        for (uint32_t u = 0; u < 8; ++u)
            if (u < chunks && !(e[u] & kFlowStatic) && __float_as_uint(stage[32u * u + lane].w) != tagLow + 1u - (e[u] >> 31))
                stale |= 1u << u;
        for (uint32_t j = 8; j < chunks; ++j)
        {
            uint32_t const f = EntryOf(td, a, b, j);
            if (!(f & kFlowStatic) && __float_as_uint(stage[32u * j + lane].w) != tagLow + 1u - (f >> 31))
                stale |= 1u << j;
        }
        if (__any_sync(0xffffffffu, stale != 0u))
        {
            unsigned long long t0 = 0;
            bool countedLate      = false;
            for (uint32_t polls = 1; stale != 0u; ++polls)
            {
                uint32_t m = stale;
                uint32_t jj[4], ee[4];
                float4 qq[4];
                int n = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (m != 0u)
                    {
                        jj[u] = static_cast<uint32_t>(__ffs(static_cast<int>(m)) - 1);
                        m &= m - 1u;
                        ee[u]             = EntryOf(td, a, b, jj[u]);
                        float4 const* src = p.pos + Index(ee[u], tagLow);
                        qq[u]             = (kDist && (ee[u] & kFlowGhost)) ? LoadPosSys(src) : LoadPosGpu(src);
                        n                 = u + 1;
                    }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (u < n)
                    {
                        if (__float_as_uint(qq[u].w) == tagLow + 1u - (ee[u] >> 31))
                        {
                            stage[32u * jj[u] + lane] = qq[u];
                            stale &= ~(1u << jj[u]);
                        }
                        else if (kDist && !countedLate && (ee[u] & kFlowGhost))
                        {
                            atomicAdd(p.distStats, 1u);  // a halo value that had not crossed NVLink yet (diagnostics)
                            countedLate = true;
                        }
                    }
                if ((polls & 1023u) == 0u)
                {
                    if (t0 == 0)
                        t0 = GlobalTimer();
                    if (GlobalTimer() - t0 > p.distTimeoutNs || LoadAcquire(p.distError) != 0u)
                    {
                        // a bug or a dead peer, never a hang: leave what was being waited for to the host
                        if (stale != 0u)
                        {
                            uint32_t const j = static_cast<uint32_t>(__ffs(static_cast<int>(stale)) - 1);
                            uint32_t const f = EntryOf(td, a, b, j);
                            bool const ghost = kDist && (f & kFlowGhost);
                            if (atomicCAS(p.distError, 0u, ghost ? 1u : 2u) == 0u && !ghost)
                            {
                                p.distError[1] = f;
                                p.distError[2] = tagLow + 1u - (f >> 31);
                                p.distError[3] = __float_as_uint(stage[32u * j + lane].w);
                                p.distError[4] = td.y;
                                p.distError[5] = tagLow - p.tagBase;
                            }
                        }
                        stale = 0u;
                        dead  = true;
                    }
                }
            }
            dead = __any_sync(0xffffffffu, dead);
        }
        __syncwarp();
    };

    uint32_t const gtid    = blockIdx.x * (nWarps * 32) + threadIdx.x;
    uint32_t const gstride = gridDim.x * (nWarps * 32);
    float omega            = 1.f;
    int omegaOf            = -1;
    // Contact (p.hist4 != nullptr; launched per substep, after PreStepKernel and the active-set update): readers of the
    // write history (step_kernel.cuh, HistSlot) must find the slot they need un-overwritten, so no warp starts sweep k
    // before every warp that has tiles finished sweep k - 2.  One counter per sweep of the launch (p.sweepDone[k] = warps that
    // have finished sweep k; a single running total would let fast warps vouch for slow ones -- the CPU model check,
    // tests/test_dataflow_protocol.py, found exactly that), one relaxed add per warp and sweep; the counter a warp needs is
    // looked at when it finishes a sweep, a sweep ahead of its use, and in steady state never waited for.
    bool const lagBound     = p.hist4 != nullptr;
    unsigned int sweepsSeen = 0;
    int sweepOf             = 0;  // the sweep the tile about to run belongs to, as far as the bound was checked
    for (int s = 0; s < p.substeps; ++s)
    {
        if (!p.skipPreStep)
            for (uint32_t i = gtid; i < p.ghostBegin; i += gstride)
                PreStepVertex<kChebyshev>(p, i, s);
        // this substep's tile sequence; the first tile's static data is requested in the shadow of the barrier
        Gen g{0u, 0};
        __syncwarp();  // the previous substep's readers of the ring are done
        Fetch(g, seq & 3u);
        Fetch(g, (seq + 1u) & 3u);
        Fetch(g, (seq + 2u) & 3u);
        CpAsyncWaitAll();
        __syncwarp();
        uint4 idA = make_uint4(0u, 0u, 0u, 0u), idB = idA;  // ring entries of the tile about to run (chunks 0-3, 4-7)
        if (kkRing[seq & 3u] >= 0)
        {
            uint4 const td = tdRing[seq & 3u];
            IssueRecords(td);
            idA = LoadIds(td, 0);
            idB = LoadIds(td, 1);  // (a tile of <= 4 chunks has no second word: what follows it is read and ignored)
        }
        if (!p.skipPreStep)
            GridSync();
        // the pre-step wrote tagBase + s (I + 1); sweep kk of this substep writes that + kk + 1
        uint32_t const tagSub = p.tagBase + static_cast<uint32_t>(s) * static_cast<uint32_t>(I + 1);
        if (kkRing[seq & 3u] >= 0)
            Gather(tdRing[seq & 3u], idA, idB, tagSub + static_cast<uint32_t>(kkRing[seq & 3u]));
        uint32_t slot = 0;  // diagnostics: index of the tile within its sweep
        int slotOf    = -1;
        for (;;)
        {
            int const k0 = kkRing[seq & 3u];
            if (k0 < 0)
                break;
            uint32_t const tagLow = tagSub + static_cast<uint32_t>(k0);
            if (lagBound && k0 != sweepOf)
            {
                sweepOf = k0;
                if (k0 >= 2 && !dead)
                {
                    unsigned int const need = pp.flowActiveWarps;
                    unsigned long long t0   = 0;
                    for (uint32_t polls = 1; sweepsSeen < need; ++polls)
                    {
                        sweepsSeen = LoadRelaxedGpu(p.sweepDone + (k0 - 2));
                        if ((polls & 255u) == 0u)
                        {
                            if (t0 == 0)
                                t0 = GlobalTimer();
                            if (GlobalTimer() - t0 > p.distTimeoutNs || LoadAcquire(p.distError) != 0u)
                            {
                                atomicCAS(p.distError, 0u, 2u);
                                dead = true;
                                break;
                            }
                        }
                    }
                    dead = __any_sync(0xffffffffu, dead);
                }
            }
            if (kChebyshev && omegaOf != k0)
            {
                omega   = __ldg(p.omega + k0);
                omegaOf = k0;
            }
            unsigned long long* tr = nullptr;
            if (p.trace != nullptr)
            {
                slot   = slotOf == k0 ? slot + 1u : 0u;
                slotOf = k0;
                if (k0 == p.traceIteration && warp == 0 && slot < nC)
                    tr = p.trace + (static_cast<size_t>(slot) * gridDim.x + blockIdx.x) * kTraceStamps;
                if (tr && lane == 0)
                    tr[0] = tr[4] = GlobalTimer();
            }
            CpAsyncWaitAll();  // this tile's positions, and the descriptor of tile seq + 2
            __syncwarp();
            uint4 const td0 = tdRing[seq & 3u];
            Await(td0, idA, idB, tagLow);
            if (tr && lane == 0)
                tr[8] = GlobalTimer();  // every value this tile reads is there
            // the next tile's ring entries: two loads now, used after this tile's store
            // (unconditional, straight into the registers the gather reads: a conditional assignment would make the compiler
            // copy the loaded words right here, i.e. wait for them; without a next tile this tile's words are simply re-read)
            int const k1    = kkRing[(seq + 1u) & 3u];
            uint4 const td1 = tdRing[(k1 >= 0 ? seq + 1u : seq) & 3u];
            idA             = LoadIds(td1, 0);
            idB             = LoadIds(td1, 1);
            // two tiles ahead: have the L2 fetch the records and ring entries from HBM (one lane, two instructions), so that
            // the loads requested a tile ahead find them there.  (Here, not after the store: the ring entries of tile seq + 2
            // are loaded at this point of the NEXT tile and need the whole tile as lead time -- measured, 0.92 vs 0.99 ms.)
            if (lane == 0 && kkRing[(seq + 2u) & 3u] >= 0)
            {
                uint4 const td2 = tdRing[(seq + 2u) & 3u];
                BulkPrefetchL2(p.records + static_cast<size_t>(td2.x) * kBlockFloat4, TileIters(td2.z) * kBlockBytes);
                BulkPrefetchL2(pp.flowIds + td2.w, ((TileChunks(td2.z) + 3u) / 4u) * 512u);
            }
            MbarWait(barRec, recWaits++ & 1u);
            SmemRecords src{recBuf + lane};
            auto afterAccumulate = [&]() {
                __syncwarp();  // every lane is done with the record buffer
                if (k1 >= 0)
                    IssueRecords(td1);
            };
            ProcessTile<kChebyshev, kDamping, false, SmemRecords, decltype(afterAccumulate), kStvk>(
                p, td0, stage, src, 0, k0, omega, lane, tr, afterAccumulate, tagLow + 1u, kDamping ? xtStage : nullptr);
            __syncwarp();  // ... and with the staged positions
            if (lagBound && k1 != k0)
            {
                // this warp's last tile of sweep k0 (its history reads are done: their values were used)
                if (lane == 0)
                    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p.sweepDone + k0) : "memory");
                sweepsSeen = k1 >= 2 ? LoadRelaxedGpu(p.sweepDone + (k1 - 2)) : 0u;  // what the next sweep will ask for
            }
            if (k1 >= 0)
                Gather(td1, idA, idB, tagSub + static_cast<uint32_t>(k1));
            Fetch(g, (seq + 3u) & 3u);
            if (tr && lane == 0)
                tr[7] = GlobalTimer();
            ++seq;
        }
        GridSync();
    }
    for (uint32_t i = gtid; i < p.ghostBegin; i += gstride)
        PostStepVertex(p, i);
}

}  // namespace vbdx
