// The persistent VBD step kernel for sm_100a: one cooperative launch runs every substep,
// every iteration and every colour of a time step, with software grid barriers between
// colours.  Replaces, from scratch, the reference's per-colour launch pair
// (gpu/impl/vbd/Kernels.cuh:148-233 VbdIteration + gpu/impl/vbd/Integrator.cu:317-325 copy-back),
// its three thrust pre/post passes (gpu/impl/vbd/Integrator.cu:190-248,329-347) and the
// per-iteration Chebyshev pass (gpu/impl/vbd/ChebyshevIntegrator.cu:38-59).
//
// Arithmetic follows the reference CPU path (the parity oracle):
//   InertialTarget / InitialPositionsForSolve   sim/vbd/Kernels.h:29-94
//   SolveVertex                                 sim/vbd/Integrator.cpp:98-136
//   AddDamping / AddInertiaDerivatives / IntegratePositions   sim/vbd/Kernels.h:179-191,310-339
//   ChebyshevUpdate                             sim/vbd/Kernels.h:104-119, ChebyshevIntegrator.cpp:15-32
// with the Stable Neo-Hookean vertex block in closed form (DESIGN.md "Math"): no deformation
// gradient is ever formed, only the edge vector to one neighbour and the opposite face's normal.
#pragma once

#include "contact.cuh"
#include "vbdx_internal.h"

#include <cuda_runtime.h>

namespace vbdx {

constexpr int kTraceStamps = 12;  // %globaltimer stamps per (colour, CTA) of the diagnostics trace

struct StepParams {
    // static topology
    const float4* __restrict__ records;        // [nBlocks][2][32] float4
    const uint4* __restrict__ tiles;           // TileDesc
    const uint32_t* __restrict__ ctaTileRange; // [nColors][gridDim.x + 1] (TMA kernel)
    const uint32_t* __restrict__ colorTileBegin; // [nColors + 1] (direct kernel)
    const uint32_t* __restrict__ ringIds;      // ring lists of all tiles
    uint32_t stageEntries;                     // per-warp shared-memory staging capacity (float4 entries)
    int nColors;
    int nVerts;   // all internal vertices (swept first, Dirichlet last)
    // state (internal vertex order, float4 per vertex)
    float4* pos;           // current-iterate buffer Q at [0,nVerts); previous-iterate buffer P at [pOff, pOff+nVerts)
    uint32_t pOff;         // 0 (P aliases Q) or nVerts (Chebyshev)
    float4* hist;          // Chebyshev: blended iterate of two iterations ago
    float4* xtildeM;       // inertial target xyz, mass in w
    float4* xt;            // positions at the start of the substep
    float4* vel;           // velocities
    float4* vtm1;          // previous velocities (only with VBDX_FLAG_ADAPTIVE_VBD_GPU_HISTORY), else null
    const float4* __restrict__ aext;
    const float* __restrict__ omega;  // Chebyshev weights per iteration
    // scalars
    float sdt, sdt2;
    float dampD;       // kD / sdt
    float detHZero;
    int strategy;
    int iterations, substeps;
    unsigned int* barrier;  // zeroed before launch
    unsigned int* sweepDone;  // barrier-free sweeps with contact: per sweep of the launch, the warps that have finished it (zeroed before launch)
    unsigned int* nonFinite;  // sentinel: owned vertices whose final position is NaN/Inf (counted by the velocity update; zeroed per step)
    // vertex-triangle contact (fc == nullptr: disabled)
    const int32_t* __restrict__ fc;      // 8 triangle ids per internal vertex, -1 terminated
    const int4* __restrict__ triF;       // collision triangles (internal vertex ids)
    const float* __restrict__ XVA;       // vertex areas (internal order)
    const float* __restrict__ FA;        // triangle areas
    float4* snap;                        // 2 x nVerts: positions as they were when the iteration started
    float4* hist4;                       // barrier-free sweeps with contact (null otherwise): the last 4 writes of every vertex, see HistSlot
    const uint32_t* __restrict__ colorVertexBegin;  // nColors + 1: internal id range of every colour
    float muC, muF, epsv;
    int skipPreStep;                     // the pre-step pass was done by PreStepKernel (contact path) or by an earlier partial launch
    int skipPostStep;                    // partial launch: more iterations of this substep follow
    int iterBegin;                       // partial launch: index (within the substep's solve) of this launch's first iteration
    uint32_t activeEnd;                  // internal ids >= activeEnd are never swept (Dirichlet vertices, then ghosts)
    int lineSearch;                      // 0: accept the full Newton step (the reference, sim/vbd/Kernels.h:329-339); 1: guarded step
    // multi-GPU domain decomposition (world == 1: single GPU)
    uint32_t ghostBegin;                 // internal ids >= ghostBegin are ghosts: written by their owner GPU only
    const uint32_t* __restrict__ sendPtr;  // per internal vertex < ghostBegin: range into sendDst (null: nothing to send)
    const uint32_t* __restrict__ sendDst;  // peer rank << 28 | internal slot of the ghost on that peer
    float4* peerPos[8];                  // position buffers of the peers (CUDA IPC mappings over NVLink)
    uint32_t peerPOff[8];                // their previous-iterate offsets
    unsigned int* peerFlags[8];          // their flag arrays: slot [rank] is written by this GPU
    unsigned int* myFlags;               // slot r: last epoch GPU r finished;  slot 8: release epoch for this GPU's CTAs
    unsigned int* distError;             // set when a peer did not show up in time
    unsigned int* distStats;             // [0] ghosts that had to be polled, [1] ns spent polling them, [2] barriers that waited for a peer's epoch, [3] ns
    unsigned int epochBase;              // epochs used by earlier launches
    unsigned long long distTimeoutNs;    // how long to wait for a peer before giving up
    // Ghost values carry the number of their write in .w ("tag": the pre-step of substep s writes tag
    // tagBase + s (iterations + 1), iteration k writes that + k + 1), and every ghost exists twice: the copy for
    // even tags sits in the ordinary slot, the copy for odd tags at [ghostExt + g] (Q) / [ghostExt + nGhost + g] (P).
    unsigned int tagBase;                // tags used by earlier launches
    uint32_t ghostExt, nGhost;
    uint32_t peerGhostExt[8], peerGhostBegin[8], peerNGhost[8];
    int rank, world;
    unsigned int peerMask;               // ranks this GPU exchanges halo data with (bit r); only they are synchronised with
    unsigned long long* trace;  // optional [nColors][gridDim.x][kTraceStamps] timestamps of one iteration (diagnostics)
    int traceIteration;
};

__device__ __forceinline__ unsigned int AddReleaseReturn(unsigned int* p, unsigned int v)
{
    unsigned int old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ unsigned int AddAcqRelReturn(unsigned int* p, unsigned int v)
{
    unsigned int old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ unsigned long long GlobalTimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned int LoadAcquire(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned int LoadRelaxedGpu(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void AddRelease(unsigned int* p, unsigned int v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// All CTAs of the (cooperatively launched, fully resident) grid meet here.  The release/acquire
// pair at gpu scope orders every position written before the barrier against every weak load
// after it, and invalidates this SM's L1 so those loads may use the default cached path.
__device__ __forceinline__ void GridBarrier(unsigned int* counter, unsigned int& target, unsigned long long* trace = nullptr)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        if (trace)
            trace[2] = GlobalTimer();  // every warp of this CTA has finished the phase
        target += gridDim.x;
        AddRelease(counter, 1u);
        while (LoadAcquire(counter) < target)
        {
        }
        if (trace)
            trace[3] = GlobalTimer();  // barrier released
    }
    __syncthreads();
}

// Split form of the grid barrier: between GridArrive and GridWait a warp may do any work that
// neither writes positions nor reads positions the colour just swept could still change.
__device__ __forceinline__ void GridArrive(unsigned int* counter, unsigned int& target, unsigned long long* trace = nullptr)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        if (trace)
            trace[2] = GlobalTimer();
        target += gridDim.x;
        AddRelease(counter, 1u);
        if (trace)
            trace[1] = GlobalTimer();  // fence + arrival issued
    }
    // the fence of the signalling thread travels through the same load/store path as everybody's requests:
    // hold the other warps' shadow work back until it is through
    __syncthreads();
}

__device__ __forceinline__ void GridWait(unsigned int* counter, unsigned int const& target, unsigned long long* trace = nullptr)
{
    if (threadIdx.x == 0)
    {
        while (LoadAcquire(counter) < target)
        {
        }
        if (trace)
            trace[3] = GlobalTimer();
    }
    __syncthreads();
}

__device__ __forceinline__ float4 LoadPos(const float4* p)
{
    return __ldcg(p);  // L2-coherent load: positions change between colours
}

// A position and the number of its write travel as ONE 16-byte datum (x, y, z, tag).  The barrier-free sweep and the halo
// exchange compare only the tag of what they read, so the 16 bytes must be written and read as a unit.  PTX models a
// vector access (st.v4 / ld.v4) as four scalar accesses; a .b128 access is a single access, hence these helpers: every
// tagged position is stored with st.relaxed.{gpu,sys}.b128 and every poll re-reads with ld.relaxed.{gpu,sys}.b128
// (SASS: STG.E.128.STRONG / LDG.E.128.STRONG).  The first read of a tile goes through cp.async 16 (one 16-byte request
// to the L2, which holds the line's only coherent copy: .cg bypasses the L1); its tag is checked the same way.
__device__ __forceinline__ void StorePosGpu(float4* p, float4 v)
{
    asm volatile("{\n .reg .b128 t;\n mov.b128 t, {%1, %2};\n st.relaxed.gpu.global.b128 [%0], t;\n}\n" ::"l"(p),
                 "l"(static_cast<unsigned long long>(__float_as_uint(v.x)) | (static_cast<unsigned long long>(__float_as_uint(v.y)) << 32)),
                 "l"(static_cast<unsigned long long>(__float_as_uint(v.z)) | (static_cast<unsigned long long>(__float_as_uint(v.w)) << 32))
                 : "memory");
}
__device__ __forceinline__ void StorePosSys(float4* p, float4 v)
{
    asm volatile("{\n .reg .b128 t;\n mov.b128 t, {%1, %2};\n st.relaxed.sys.global.b128 [%0], t;\n}\n" ::"l"(p),
                 "l"(static_cast<unsigned long long>(__float_as_uint(v.x)) | (static_cast<unsigned long long>(__float_as_uint(v.y)) << 32)),
                 "l"(static_cast<unsigned long long>(__float_as_uint(v.z)) | (static_cast<unsigned long long>(__float_as_uint(v.w)) << 32))
                 : "memory");
}

// InitialPositionsForSolve (sim/vbd/Kernels.h:45-94).  Roundings are spelled out (no implicit
// multiply-add contraction) so that every kernel that inlines this code produces the same bits.
__device__ __forceinline__ float3 InitialPosition(
    float3 xt,
    float3 vtm1,
    float3 vt,
    float3 a,
    float dt,
    float dt2,
    int strategy)
{
    if (strategy == 0)
        return xt;
    float3 x = make_float3(fmaf(dt, vt.x, xt.x), fmaf(dt, vt.y, xt.y), fmaf(dt, vt.z, xt.z));
    if (strategy == 1)
        return x;
    float atilde = 1.f;
    if (strategy >= 3)
    {
        float const an2 = fmaf(a.z, a.z, fmaf(a.y, a.y, __fmul_rn(a.x, a.x)));
        atilde          = 0.f;
        if (an2 != 0.f)
        {
            if (strategy == 3)
            {
                float const d = fmaf(__fdiv_rn(__fsub_rn(vt.z, vtm1.z), dt), a.z,
                                     fmaf(__fdiv_rn(__fsub_rn(vt.y, vtm1.y), dt), a.y,
                                          __fmul_rn(__fdiv_rn(__fsub_rn(vt.x, vtm1.x), dt), a.x)));
                atilde = fminf(fmaxf(__fdiv_rn(d, an2), 0.f), 1.f);
            }
            else
            {
                float const nrm =
                    __fadd_rn(__fsqrt_rn(fmaf(vt.z, vt.z, fmaf(vt.y, vt.y, __fmul_rn(vt.x, vt.x)))), 1.17549435e-38f);
                float const d = fmaf(__fdiv_rn(vt.z, nrm), a.z, fmaf(__fdiv_rn(vt.y, nrm), a.y, __fmul_rn(__fdiv_rn(vt.x, nrm), a.x)));
                atilde        = fminf(fabsf(__fdiv_rn(d, an2)), 1.f);
            }
        }
    }
    float const s = __fmul_rn(dt2, atilde);
    return make_float3(fmaf(s, a.x, x.x), fmaf(s, a.y, x.y), fmaf(s, a.z, x.z));
}

// Domain decomposition: push the new position of an owned vertex into the ghost slots of the peers that hold
// it: plain 16-byte stores into peer memory over NVLink, the write's tag travelling in .w of the same store, so
// that a reader can tell from the datum itself whether it has arrived (no fence, no flag on the critical path).
__device__ __forceinline__ void SendToPeers(StepParams const& p, uint32_t vi, float4 raw, float4 blended, uint32_t tag)
{
    if (p.sendPtr == nullptr)
        return;
    uint32_t const b = __ldg(p.sendPtr + vi), e = __ldg(p.sendPtr + vi + 1);
    raw.w = blended.w = __uint_as_float(tag);
    bool const odd = (tag & 1u) != 0u;
    for (uint32_t k = b; k < e; ++k)
    {
        uint32_t const dst = __ldg(p.sendDst + k);
        uint32_t const r = dst >> 28, slot = dst & 0x0fffffffu;
        uint32_t const q = odd ? p.peerGhostExt[r] + (slot - p.peerGhostBegin[r]) : slot;
        StorePosSys(p.peerPos[r] + q, raw);
        if (p.peerPOff[r] != 0u)
            StorePosSys(p.peerPos[r] + q + (odd ? p.peerNGhost[r] : p.peerPOff[r]), blended);
    }
}

// index (into pos) of the copy of ghost `base` that holds the write with tag `tag`; prev = previous-iterate buffer
__device__ __forceinline__ uint32_t GhostIndex(StepParams const& p, uint32_t base, bool prev, uint32_t tag)
{
    return (tag & 1u) ? p.ghostExt + (prev && p.pOff != 0u ? p.nGhost : 0u) + (base - p.ghostBegin) : base + (prev ? p.pOff : 0u);
}

// polls of a tagged position: one 16-byte access (see StorePosGpu, step_kernel.cuh)
__device__ __forceinline__ float4 UnpackB128(unsigned long long lo, unsigned long long hi)
{
    return make_float4(__uint_as_float(static_cast<unsigned>(lo)), __uint_as_float(static_cast<unsigned>(lo >> 32)),
                       __uint_as_float(static_cast<unsigned>(hi)), __uint_as_float(static_cast<unsigned>(hi >> 32)));
}
__device__ __forceinline__ float4 LoadPosGpu(const float4* q)
{
    unsigned long long lo, hi;
    asm volatile("{\n .reg .b128 t;\n ld.relaxed.gpu.global.b128 t, [%2];\n mov.b128 {%0, %1}, t;\n}\n" : "=l"(lo), "=l"(hi) : "l"(q) : "memory");
    return UnpackB128(lo, hi);
}

// Contact reads go beyond the 1-rings, to vertices of OTHER bodies whose tiles share no dependency with the reader's: in a
// barrier-free sweep such a vertex may be a sweep behind or ahead.  So with contact every write also goes to a history of
// the vertex' last four writes, slot = write number mod 4, separately the raw sweep result and (Chebyshev) the blended
// iterate: a reader takes the write it needs (this sweep's raw result of a lower colour; the previous sweep's iterate of
// its own or a higher colour -- what the reference's per-colour write buffer exposes, gpu/impl/vbd/Kernels.cuh:203-223),
// waiting for it if it is not there yet.  That no slot is overwritten while somebody may still read it is the kernel's
// part: no warp starts sweep k before every warp has finished sweep k - 2 (StepKernelFlow).
__device__ __forceinline__ float4* HistSlot(StepParams const& p, uint32_t tag, bool blended)
{
    uint32_t const per = p.pOff != 0u ? 2u : 1u;  // Chebyshev keeps both
    return p.hist4 + static_cast<size_t>((tag & 3u) * per + ((blended && per == 2u) ? 1u : 0u)) * static_cast<size_t>(p.nVerts);
}
__device__ __noinline__ float4 AwaitHist(StepParams const& p, float4 const* src, uint32_t want)
{
    float4 q              = LoadPosGpu(src);
    unsigned long long t0 = 0;
    for (uint32_t polls = 1; __float_as_uint(q.w) != want; ++polls)
    {
        if ((polls & 255u) == 0u)
        {
            if (t0 == 0)
                t0 = GlobalTimer();
            if (GlobalTimer() - t0 > p.distTimeoutNs || LoadAcquire(p.distError) != 0u)
            {
                if (atomicCAS(p.distError, 0u, 2u) == 0u)
                {
                    p.distError[1] = static_cast<uint32_t>(src - p.hist4) % static_cast<uint32_t>(p.nVerts);
                    p.distError[2] = want;
                    p.distError[3] = __float_as_uint(q.w);
                    p.distError[4] = 0xffffffffu;  // a contact read
                    p.distError[5] = want - p.tagBase;
                }
                break;
            }
        }
        q = LoadPosGpu(src);
    }
    return q;
}

// Per-vertex pre-step, fused with the velocity update of the previous substep
// (sim/vbd/Integrator.cpp:31-35,39; sim/vbd/Kernels.h:29-94):
//   v = (x - xt)/h [s > 0];  xt = x;  xtilde = xt + h v + h^2 a;  x = initial guess
template <bool kChebyshev>
__device__ __forceinline__ void PreStepVertex(StepParams const& p, uint32_t i, int s)
{
    float4 const x4    = __ldcg(p.pos + p.pOff + i);
    float4 v4          = __ldcg(p.vel + i);
    float3 const vprev = make_float3(v4.x, v4.y, v4.z);
    if (s > 0)
    {
        float4 const xt4 = __ldcg(p.xt + i);
        v4.x = __fdiv_rn(__fsub_rn(x4.x, xt4.x), p.sdt);
        v4.y = __fdiv_rn(__fsub_rn(x4.y, xt4.y), p.sdt);
        v4.z = __fdiv_rn(__fsub_rn(x4.z, xt4.z), p.sdt);
        p.vel[i] = v4;
    }
    // "previous velocity" of InitialPositionsForSolve: the CPU reference passes vt == v
    // (sim/vbd/Integrator.cpp:32,61-68); the GPU reference keeps a real v(t-1)
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
    xm.x            = fmaf(p.sdt2, a4.x, fmaf(p.sdt, v4.x, x4.x));
    xm.y            = fmaf(p.sdt2, a4.y, fmaf(p.sdt, v4.y, x4.y));
    xm.z            = fmaf(p.sdt2, a4.z, fmaf(p.sdt, v4.z, x4.z));
    p.xtildeM[i]    = xm;
    p.xt[i]         = x4;
    float3 const x0 = InitialPosition(
        make_float3(x4.x, x4.y, x4.z), vtm1, make_float3(v4.x, v4.y, v4.z), make_float3(a4.x, a4.y, a4.z), p.sdt,
        p.sdt2, p.strategy);
    // .w carries the number of the write (pre-step of substep s; sweep k adds k + 1): the halo exchange of the domain
    // decomposition and the barrier-free sweep (step_kernel_pipe.cuh, PipeParams::dataflow) read it back
    uint32_t const tag = p.tagBase + static_cast<uint32_t>(s) * static_cast<uint32_t>(p.iterations + 1);
    float4 const o = make_float4(x0.x, x0.y, x0.z, __uint_as_float(tag));
    StorePosGpu(p.pos + i, o);
    if constexpr (kChebyshev)
        StorePosGpu(p.pos + p.pOff + i, o);
    if (p.snap != nullptr)
        p.snap[i] = o;
    if (p.hist4 != nullptr)
        StorePosGpu(HistSlot(p, tag, true) + i, o);
    SendToPeers(p, i, o, o, tag);
}

// velocity update of the last substep (sim/vbd/Integrator.cpp:39); with the GPU-history flag also
// v(t-1) <- v like the reference's UpdateBdfState (gpu/impl/vbd/Integrator.cu:329-347)
__device__ __forceinline__ void PostStepVertex(StepParams const& p, uint32_t i)
{
    float4 const x4  = __ldcg(p.pos + p.pOff + i);
    float4 const xt4 = __ldcg(p.xt + i);
    float4 v4        = __ldcg(p.vel + i);
    if (p.vtm1 != nullptr)
        p.vtm1[i] = v4;
    v4.x     = __fdiv_rn(__fsub_rn(x4.x, xt4.x), p.sdt);
    v4.y     = __fdiv_rn(__fsub_rn(x4.y, xt4.y), p.sdt);
    v4.z     = __fdiv_rn(__fsub_rn(x4.z, xt4.z), p.sdt);
    p.vel[i] = v4;
    if (!isfinite(x4.x + x4.y + x4.z) && p.nonFinite != nullptr)
        atomicAdd(p.nonFinite, 1u);
}

template <bool kChebyshev>
__global__ void PreStepKernel(const __grid_constant__ StepParams p)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p.ghostBegin)
        PreStepVertex<kChebyshev>(p, i, 0);
}

// Record source of the direct kernel: each warp reads its blocks straight from global memory
// (streaming loads, evict-first).
struct DirectRecords {
    float4 const* rec;
    __device__ __forceinline__ void Fetch(float4& c0, float4& c1)
    {
        c0 = __ldcs(rec);
        c1 = __ldcs(rec + 32);
        rec += kBlockFloat4;
    }
};

// One warp tile: stage the 1-rings in shared memory, accumulate the elastic blocks of all
// incident tets, reduce over the lanes that share a vertex, solve, write back.
//   RecordSource::Fetch(c0,c1) yields this lane's 32-byte record of the next block.
struct NoHook {
    __device__ __forceinline__ void operator()() const {}
};

template <bool kChebyshev, bool kDamping, bool kStageInside, class RecordSource, class AfterAccumulate = NoHook, bool kStvk = false>
__device__ __forceinline__ void ProcessTile(
    StepParams const& p,
    uint4 const td,
    float4* __restrict__ stage,
    RecordSource& src,
    int color,
    int k,
    float omega,
    uint32_t lane,
    unsigned long long* trace = nullptr,
    AfterAccumulate afterAccumulate = AfterAccumulate{},  // runs once the tile's records have been consumed
    uint32_t sendTag = 0u,                                // domain decomposition: tag of this sweep's writes
    float4 const* xtStaged = nullptr)                     // damping / contact: xt of this lane's vertex, staged with the tile's gather
{
    float4 const* __restrict__ posQ = p.pos;
    uint32_t const lw         = TileLog2W(td.z);
    uint32_t const nverts     = TileVerts(td.z);
    uint32_t const ringChunks = TileChunks(td.z);
    uint32_t const iters      = TileIters(td.z);
    uint32_t const grp        = lane >> lw;
    bool const valid          = grp < nverts;
    uint32_t const vi         = td.y + (valid ? grp : 0u);

    // first record block: issue its loads before anything else so they overlap the ring gather
    float4 n0, n1;
    src.Fetch(n0, n1);
    // The tile's list = its own vertices (start values, from the previous-iterate buffer P) followed
    // by their 1-rings.  Gather it once: ringChunks independent scattered loads per lane.
    if constexpr (kStageInside)
    {
        uint32_t const* ids = p.ringIds + td.w + lane;
        __syncwarp();  // the previous tile's readers are done with the staging area
        for (uint32_t j = 0; j < ringChunks; ++j)
        {
            uint32_t const id = __ldg(ids + 32 * j);
            stage[32 * j + lane] = LoadPos(posQ + (id & ~kPrevFlag) + ((id & kPrevFlag) ? p.pOff : 0u));
        }
        __syncwarp();
    }
    float4 const xi = stage[valid ? grp : 0u];  // own position; nobody writes vertex vi during this colour
    if (trace && lane == 0)
        trace[5] = GlobalTimer();  // 1-rings staged
    // epilogue operands, requested now so that they are in registers when the solve needs them
    float4 const xm = __ldcg(p.xtildeM + vi);
    float4 h2       = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (kChebyshev)
        if (k > 1)
            h2 = __ldcg(p.hist + vi);
    // damping / contact: the substep's start position and the head of the vertex' contact list (-1: no contacts), likewise
    [[maybe_unused]] float4 xtv = make_float4(0.f, 0.f, 0.f, 0.f);
    [[maybe_unused]] int fc0    = -1;
    if constexpr (kDamping)
    {
        if (xtStaged == nullptr)
            xtv = __ldcg(p.xt + vi);
        if (p.fc != nullptr)
            fc0 = __ldcg(p.fc + static_cast<size_t>(vi) * kMaxContacts);
    }

    float h00 = 0.f, h01 = 0.f, h02 = 0.f, h11 = 0.f, h12 = 0.f, h22 = 0.f, hd = 0.f;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if constexpr (kStvk)
    {
        // St. Venant-Kirchhoff, psi = mu tr(E^2) + lambda/2 tr(E)^2, E = (F^T F - I)/2
        // (physics/SaintVenantKirchhoffEnergy.h), contracted to the vertex block like sim/vbd/Kernels.h:137-171:
        //   F = d_a (x) g_a + d_b (x) g_b + d_c (x) g_c        (d_n = x_n - x_i,  g_n = grad N_n,  sum_n g_n = 0)
        //   S = 2 mu E + lambda tr(E) I,   grad_i = w F S g_i,
        //   H_i = w [ (g_i.S g_i) I + (mu + lambda) (F g_i)(F g_i)^T + mu |g_i|^2 F F^T ]
        // An incidence record is two blocks (vbdx_internal.h).
#pragma unroll 1
        for (uint32_t t = 0; t < iters; t += 2)
        {
            float4 const c0 = n0, c1 = n1;
            float4 c2, c3;
            src.Fetch(c2, c3);
            if (t + 2 < iters)
                src.Fetch(n0, n1);
            uint32_t const idx = __float_as_uint(c0.x);
            float4 const p1 = stage[idx & 1023u];
            float4 const p2 = stage[(idx >> 10) & 1023u];
            float4 const p3 = stage[(idx >> 20) & 1023u];
            float const da[3] = {p1.x - xi.x, p1.y - xi.y, p1.z - xi.z};
            float const db[3] = {p2.x - xi.x, p2.y - xi.y, p2.z - xi.z};
            float const dc[3] = {p3.x - xi.x, p3.y - xi.y, p3.z - xi.z};
            float const ga[3] = {c0.y, c0.z, c0.w}, gb[3] = {c1.x, c1.y, c1.z}, gc[3] = {c2.x, c2.y, c2.z};
            float const wmu = c1.w, wlam = c2.w;
            float const gi[3] = {-(ga[0] + gb[0] + gc[0]), -(ga[1] + gb[1] + gc[1]), -(ga[2] + gb[2] + gc[2])};
            float F[3][3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    F[r][c] = da[r] * ga[c] + db[r] * gb[c] + dc[r] * gc[c];
            // E (symmetric) and S' = 2 wmu E + wlam tr(E) I
            float const e00 = 0.5f * (F[0][0] * F[0][0] + F[1][0] * F[1][0] + F[2][0] * F[2][0] - 1.f);
            float const e11 = 0.5f * (F[0][1] * F[0][1] + F[1][1] * F[1][1] + F[2][1] * F[2][1] - 1.f);
            float const e22 = 0.5f * (F[0][2] * F[0][2] + F[1][2] * F[1][2] + F[2][2] * F[2][2] - 1.f);
            float const e01 = 0.5f * (F[0][0] * F[0][1] + F[1][0] * F[1][1] + F[2][0] * F[2][1]);
            float const e02 = 0.5f * (F[0][0] * F[0][2] + F[1][0] * F[1][2] + F[2][0] * F[2][2]);
            float const e12 = 0.5f * (F[0][1] * F[0][2] + F[1][1] * F[1][2] + F[2][1] * F[2][2]);
            float const ltr = wlam * (e00 + e11 + e22), m2 = 2.f * wmu;
            float const s00 = m2 * e00 + ltr, s11 = m2 * e11 + ltr, s22 = m2 * e22 + ltr;
            float const s01 = m2 * e01, s02 = m2 * e02, s12 = m2 * e12;
            float const sg0 = s00 * gi[0] + s01 * gi[1] + s02 * gi[2];
            float const sg1 = s01 * gi[0] + s11 * gi[1] + s12 * gi[2];
            float const sg2 = s02 * gi[0] + s12 * gi[1] + s22 * gi[2];
            g0 += F[0][0] * sg0 + F[0][1] * sg1 + F[0][2] * sg2;
            g1 += F[1][0] * sg0 + F[1][1] * sg1 + F[1][2] * sg2;
            g2 += F[2][0] * sg0 + F[2][1] * sg1 + F[2][2] * sg2;
            hd += gi[0] * sg0 + gi[1] * sg1 + gi[2] * sg2;
            float const f0 = F[0][0] * gi[0] + F[0][1] * gi[1] + F[0][2] * gi[2];
            float const f1 = F[1][0] * gi[0] + F[1][1] * gi[1] + F[1][2] * gi[2];
            float const f2 = F[2][0] * gi[0] + F[2][1] * gi[1] + F[2][2] * gi[2];
            float const ml = wmu + wlam, mg = wmu * (gi[0] * gi[0] + gi[1] * gi[1] + gi[2] * gi[2]);
            h00 += ml * f0 * f0 + mg * (F[0][0] * F[0][0] + F[0][1] * F[0][1] + F[0][2] * F[0][2]);
            h01 += ml * f0 * f1 + mg * (F[0][0] * F[1][0] + F[0][1] * F[1][1] + F[0][2] * F[1][2]);
            h02 += ml * f0 * f2 + mg * (F[0][0] * F[2][0] + F[0][1] * F[2][1] + F[0][2] * F[2][2]);
            h11 += ml * f1 * f1 + mg * (F[1][0] * F[1][0] + F[1][1] * F[1][1] + F[1][2] * F[1][2]);
            h12 += ml * f1 * f2 + mg * (F[1][0] * F[2][0] + F[1][1] * F[2][1] + F[1][2] * F[2][2]);
            h22 += ml * f2 * f2 + mg * (F[2][0] * F[2][0] + F[2][1] * F[2][1] + F[2][2] * F[2][2]);
            (void)c3;
        }
    }
    else
#pragma unroll 1
    for (uint32_t t = 0; t < iters; ++t)
    {
        float4 const c0 = n0, c1 = n1;
        if (t + 1 < iters)
            src.Fetch(n0, n1);  // next block in flight while this one is computed
        uint32_t const idx = __float_as_uint(c0.x);
        float4 const p1 = stage[idx & 1023u];
        float4 const p2 = stage[(idx >> 10) & 1023u];
        float4 const p3 = stage[(idx >> 20) & 1023u];
        // With a, b, c the other vertices of the tet:  d = x_a - x_i,  e1 = x_b - x_a,  e2 = x_c - x_a.
        //   F grad N_i           = d (u_a+u_b+u_c) + e1 u_b + e2 u_c
        //   dJ/dx_i = cof(F) grad N_i = -detG S,   S = e1 x e2   (the opposite face's area normal)
        //   J = det F            = detG (d . S)
        float const dx = p1.x - xi.x, dy = p1.y - xi.y, dz = p1.z - xi.z;
        float const e1x = p2.x - p1.x, e1y = p2.y - p1.y, e1z = p2.z - p1.z;
        float const e2x = p3.x - p1.x, e2y = p3.y - p1.y, e2z = p3.z - p1.z;
        float const Sx = e1y * e2z - e1z * e2y;
        float const Sy = e1z * e2x - e1x * e2z;
        float const Sz = e1x * e2y - e1y * e2x;
        float const detD = dx * Sx + dy * Sy + dz * Sz;
        float const beta = c1.x;
        float const cS   = c1.y - beta * detD;  // = -wg lambda detG (J - alpha)
        // gradient: wg mu F grad N_i + wg lambda (J - alpha) dJ/dx_i
        g0 += dx * c0.y + e1x * c0.z + e2x * c0.w + cS * Sx;
        g1 += dy * c0.y + e1y * c0.z + e2y * c0.w + cS * Sy;
        g2 += dz * c0.y + e1z * c0.z + e2z * c0.w + cS * Sz;
        // Hessian: wg mu |grad N_i|^2 I + wg lambda (dJ/dx_i)(dJ/dx_i)^T
        float const t0 = beta * Sx, t1 = beta * Sy, t2 = beta * Sz;
        h00 += t0 * Sx;
        h01 += t0 * Sy;
        h02 += t0 * Sz;
        h11 += t1 * Sy;
        h12 += t1 * Sz;
        h22 += t2 * Sz;
        hd += c1.z;
    }
    bool const secondPass = kStvk && !kDamping && p.lineSearch != 0;  // warp-uniform
    if (!secondPass)
        afterAccumulate();
    if (trace && lane == 0)
        trace[6] = GlobalTimer();  // incident tets accumulated
    // butterfly over the w lanes that share a vertex (fixed order => deterministic)
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
    bool const leader = valid && (lane & ((1u << lw) - 1u)) == 0u;
    float nx = xi.x, ny = xi.y, nz = xi.z;  // the vertex' new position (leader lane)
    if (leader)
    {
        h00 = __fadd_rn(h00, hd);
        h11 = __fadd_rn(h11, hd);
        h22 = __fadd_rn(h22, hd);
        float x = xi.x, y = xi.y, z = xi.z;
        if constexpr (kDamping)
        {
            float4 const xt = xtStaged != nullptr ? xtStaged[lane] : xtv;
            float const D   = p.dampD;
            float const ex = x - xt.x, ey = y - xt.y, ez = z - xt.z;
            g0 = fmaf(D, fmaf(h02, ez, fmaf(h01, ey, __fmul_rn(h00, ex))), g0);
            g1 = fmaf(D, fmaf(h12, ez, fmaf(h11, ey, __fmul_rn(h01, ex))), g1);
            g2 = fmaf(D, fmaf(h22, ez, fmaf(h12, ey, __fmul_rn(h02, ex))), g2);
            // explicit roundings from here to the Newton step: the statements around the (runtime-optional)
            // contact block must not be contracted differently in the different kernels that inline this code
            float const sc = 1.f + D;
            h00 = __fmul_rn(h00, sc), h01 = __fmul_rn(h01, sc), h02 = __fmul_rn(h02, sc);
            h11 = __fmul_rn(h11, sc), h12 = __fmul_rn(h12, sc), h22 = __fmul_rn(h22, sc);
        }
        if constexpr (kDamping)
        {
            // vertex-triangle contact (gpu/impl/vbd/Kernels.cuh:203-223): area-scaled penalty over <= 8 triangles.
            // Triangle vertices are read as the reference's per-colour write buffer makes them visible: colours
            // already swept in this iteration -> current values, later colours -> previous iterate, the colour
            // being swept -> the values it had when the iteration started (snapshot).
            if (fc0 >= 0)
            {
                int f[kMaxContacts];
                int nContacts  = 0;
                float const kC = ContactPenaltyScale(p.fc + static_cast<size_t>(vi) * kMaxContacts, p.FA, __ldg(p.XVA + vi), p.muC, f, nContacts);
                if (nContacts > 0)
                {
                    uint32_t cb = 0u, ce = 0u;
                    if (p.hist4 == nullptr)
                        cb = __ldg(p.colorVertexBegin + color), ce = __ldg(p.colorVertexBegin + color + 1);
                    else
                        for (int c = 0; c < p.nColors; ++c)  // (the barrier-free kernel does not know a tile's colour)
                            if (__ldg(p.colorVertexBegin + c + 1) <= vi)
                                cb = __ldg(p.colorVertexBegin + c + 1);
                    float4 const* snapK = p.snap + static_cast<size_t>(k & 1) * p.nVerts;
                    float3 const xtv3   = F3(xtStaged != nullptr ? xtStaged[lane] : xtv);
                    float gC[3] = {0.f, 0.f, 0.f}, HC[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    for (int c = 0; c < nContacts; ++c)
                    {
                        int4 const tri   = __ldg(p.triF + f[c]);
                        int const id[3]  = {tri.x, tri.y, tri.z};
                        float3 xf[3], xtf[3];
                        for (int a = 0; a < 3; ++a)
                        {
                            uint32_t const j = static_cast<uint32_t>(id[a]);
                            float4 q;
                            if (p.hist4 == nullptr)
                                q = j < cb ? __ldcg(p.pos + j) : j >= ce ? __ldcg(p.pos + p.pOff + j) : __ldcg(snapK + j);
                            else if (j >= p.activeEnd)
                                q = __ldcg(p.pos + p.pOff + j);  // never swept
                            else
                                q = AwaitHist(p, HistSlot(p, j < cb ? sendTag : sendTag - 1u, j >= cb) + j, j < cb ? sendTag : sendTag - 1u);
                            xf[a]            = F3(q);
                            xtf[a]           = F3(__ldcg(p.xt + j));
                        }
                        AccumulateVertexTriangleContact(xtv3, make_float3(x, y, z), xtf, xf, p.sdt, __fmul_rn(kC, __ldg(p.FA + f[c])), p.muF,
                                                        p.epsv, gC, HC);
                    }
                    g0 = __fadd_rn(g0, gC[0]), g1 = __fadd_rn(g1, gC[1]), g2 = __fadd_rn(g2, gC[2]);
                    h00 = __fadd_rn(h00, HC[0]), h01 = __fadd_rn(h01, HC[1]), h02 = __fadd_rn(h02, HC[2]);
                    h11 = __fadd_rn(h11, HC[3]), h12 = __fadd_rn(h12, HC[4]), h22 = __fadd_rn(h22, HC[5]);
                }
            }
        }
        float const K   = xm.w / p.sdt2;
        h00 = __fadd_rn(h00, K), h11 = __fadd_rn(h11, K), h22 = __fadd_rn(h22, K);
        g0 = fmaf(K, x - xm.x, g0);
        g1 = fmaf(K, y - xm.y, g1);
        g2 = fmaf(K, z - xm.z, g2);
        // Newton step with the explicit cofactor inverse of the symmetric 3x3
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
            x = fmaf(-r, fmaf(i02, g2, fmaf(i01, g1, __fmul_rn(i00, g0))), x);
            y = fmaf(-r, fmaf(i12, g2, fmaf(i11, g1, __fmul_rn(i01, g0))), y);
            z = fmaf(-r, fmaf(i22, g2, fmaf(i12, g1, __fmul_rn(i02, g0))), z);
        }
        if constexpr (kStvk)
        if (p.lineSearch != 0)
        {
            // Guard, part 1: the step must be a finite descent direction of the local objective whose gradient and
            // Hessian were just accumulated.  Compiled for St. Venant-Kirchhoff only: with the Stable Neo-Hookean
            // energy det F is affine in ONE vertex, so the local objective is exactly quadratic with the positive
            // definite Hessian used above and the Newton step IS its minimiser -- there is nothing to guard.  StVK's
            // full Hessian (the reference's energies return it unprojected) turns indefinite under compression.
            // Fallback: steepest descent scaled by a bound on the spectral radius of H (Gershgorin).
            float const dx = x - xi.x, dy = y - xi.y, dz = z - xi.z;
            float const slope = g0 * dx + g1 * dy + g2 * dz;
            if (!(slope < 0.f) || !isfinite(slope))
            {
                float const bound = fmaxf(fabsf(h00) + fabsf(h01) + fabsf(h02),
                                          fmaxf(fabsf(h01) + fabsf(h11) + fabsf(h12), fabsf(h02) + fabsf(h12) + fabsf(h22)));
                float const sc = bound > 0.f ? 1.f / bound : 0.f;
                x = fmaf(-sc, g0, xi.x), y = fmaf(-sc, g1, xi.y), z = fmaf(-sc, g2, xi.z);
            }
        }
        nx = x, ny = y, nz = z;
    }
    if constexpr (kStvk && !kDamping)
    {
        if (secondPass)
        {
            // Guard, part 2 (St. Venant-Kirchhoff, whose local objective is quartic in x_i): Armijo backtracking on the
            // TRUE local objective  1/2 K |x - xtilde|^2 + sum_e w psi(F_e(x))  over t in {1, 1/2, 1/4}, all trial
            // points evaluated in one more pass over the tile's records (still in shared memory).  A vertex for which
            // no trial point passes keeps its position.
            uint32_t const head = lane & ~((1u << lw) - 1u);
            float const ddx = __shfl_sync(0xffffffffu, nx - xi.x, head);
            float const ddy = __shfl_sync(0xffffffffu, ny - xi.y, head);
            float const ddz = __shfl_sync(0xffffffffu, nz - xi.z, head);
            float de1 = 0.f, de2 = 0.f, de4 = 0.f;  // psi(t) - psi(0) summed over this lane's tets, t = 1, 1/2, 1/4
            src.Rewind(iters);
            float4 m0, m1;
            src.Fetch(m0, m1);
#pragma unroll 1
            for (uint32_t t = 0; t < iters; t += 2)
            {
                float4 const c0 = m0, c1 = m1;
                float4 c2, c3;
                src.Fetch(c2, c3);
                if (t + 2 < iters)
                    src.Fetch(m0, m1);
                uint32_t const idx = __float_as_uint(c0.x);
                float4 const p1 = stage[idx & 1023u];
                float4 const p2 = stage[(idx >> 10) & 1023u];
                float4 const p3 = stage[(idx >> 20) & 1023u];
                float const da[3] = {p1.x - xi.x, p1.y - xi.y, p1.z - xi.z};
                float const db[3] = {p2.x - xi.x, p2.y - xi.y, p2.z - xi.z};
                float const dc[3] = {p3.x - xi.x, p3.y - xi.y, p3.z - xi.z};
                float const ga[3] = {c0.y, c0.z, c0.w}, gb[3] = {c1.x, c1.y, c1.z}, gc[3] = {c2.x, c2.y, c2.z};
                float const wmu = c1.w, wlam = c2.w;
                float const gi[3] = {-(ga[0] + gb[0] + gc[0]), -(ga[1] + gb[1] + gc[1]), -(ga[2] + gb[2] + gc[2])};
                float const dd[3] = {ddx, ddy, ddz};
                float F[3][3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        F[r][c] = da[r] * ga[c] + db[r] * gb[c] + dc[r] * gc[c];
                auto psi = [&](float tt) {
                    // F(t) = F + t d (x) grad N_i
                    float G[3][3];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            G[r][c] = fmaf(tt * dd[r], gi[c], F[r][c]);
                    float const e00 = 0.5f * (G[0][0] * G[0][0] + G[1][0] * G[1][0] + G[2][0] * G[2][0] - 1.f);
                    float const e11 = 0.5f * (G[0][1] * G[0][1] + G[1][1] * G[1][1] + G[2][1] * G[2][1] - 1.f);
                    float const e22 = 0.5f * (G[0][2] * G[0][2] + G[1][2] * G[1][2] + G[2][2] * G[2][2] - 1.f);
                    float const e01 = 0.5f * (G[0][0] * G[0][1] + G[1][0] * G[1][1] + G[2][0] * G[2][1]);
                    float const e02 = 0.5f * (G[0][0] * G[0][2] + G[1][0] * G[1][2] + G[2][0] * G[2][2]);
                    float const e12 = 0.5f * (G[0][1] * G[0][2] + G[1][1] * G[1][2] + G[2][1] * G[2][2]);
                    float const tr = e00 + e11 + e22;
                    return wmu * (e00 * e00 + e11 * e11 + e22 * e22 + 2.f * (e01 * e01 + e02 * e02 + e12 * e12)) + 0.5f * wlam * tr * tr;
                };
                float const psi0 = psi(0.f);
                de1 += psi(1.f) - psi0;
                de2 += psi(0.5f) - psi0;
                de4 += psi(0.25f) - psi0;
                (void)c3;
            }
            afterAccumulate();
            for (uint32_t o = (1u << lw) >> 1; o > 0; o >>= 1)
            {
                de1 += __shfl_xor_sync(0xffffffffu, de1, o);
                de2 += __shfl_xor_sync(0xffffffffu, de2, o);
                de4 += __shfl_xor_sync(0xffffffffu, de4, o);
            }
            if (leader)
            {
                // g0..g2 hold the gradient of the same objective at x_i (elastic + inertia; this variant carries no
                // damping or contact);  inertia difference: K t d.(x - xtilde) + 1/2 K t^2 |d|^2
                float const K     = xm.w / p.sdt2;
                float const slope = g0 * ddx + g1 * ddy + g2 * ddz;
                float const lin   = K * (ddx * (xi.x - xm.x) + ddy * (xi.y - xm.y) + ddz * (xi.z - xm.z));
                float const quad  = 0.5f * K * (ddx * ddx + ddy * ddy + ddz * ddz);
                float const c1a   = 1e-4f;
                float t = 0.f;
                if (de1 + lin + quad <= c1a * slope)
                    t = 1.f;
                else if (de2 + 0.5f * lin + 0.25f * quad <= c1a * 0.5f * slope)
                    t = 0.5f;
                else if (de4 + 0.25f * lin + 0.0625f * quad <= c1a * 0.25f * slope)
                    t = 0.25f;
                if (t != 1.f)
                    nx = fmaf(t, ddx, xi.x), ny = fmaf(t, ddy, xi.y), nz = fmaf(t, ddz, xi.z);
            }
        }
    }
    if (leader)
    {
        float const x = nx, y = ny, z = nz;
        float4 const raw = make_float4(x, y, z, __uint_as_float(sendTag));
        if constexpr (kChebyshev)
        {
            // Q <- raw sweep result (read by higher colours in this iteration);
            // P <- blended iterate (read by lower colours in the next iteration and as this
            // vertex' own start); hist <- previous blended iterate.
            float4 out = raw;
            if (k > 1)
            {
                out.x = fmaf(omega, x - h2.x, h2.x);
                out.y = fmaf(omega, y - h2.y, h2.y);
                out.z = fmaf(omega, z - h2.z, h2.z);
            }
            p.hist[vi] = make_float4(xi.x, xi.y, xi.z, 0.f);
            StorePosGpu(p.pos + vi, raw);
            StorePosGpu(p.pos + p.pOff + vi, out);
            SendToPeers(p, vi, raw, out, sendTag);
            if (p.snap != nullptr)
                p.snap[static_cast<size_t>((k + 1) & 1) * p.nVerts + vi] = out;  // what iteration k+1 starts from
            if (p.hist4 != nullptr)
            {
                StorePosGpu(HistSlot(p, sendTag, false) + vi, raw);
                StorePosGpu(HistSlot(p, sendTag, true) + vi, out);
            }
        }
        else
        {
            StorePosGpu(p.pos + vi, raw);
            SendToPeers(p, vi, raw, raw, sendTag);
            if (p.snap != nullptr)
                p.snap[static_cast<size_t>((k + 1) & 1) * p.nVerts + vi] = raw;
            if (p.hist4 != nullptr)
                StorePosGpu(HistSlot(p, sendTag, false) + vi, raw);
        }
    }
    if (trace && lane == 0)
        trace[11] = GlobalTimer();  // new positions stored
}

// Sweep of one colour: every warp walks its share of the colour's tiles.
template <bool kChebyshev, bool kDamping>
__device__ __forceinline__ void SweepColor(StepParams const& p, int color, int k, float omega, float4* stage, unsigned long long* trace)
{
    // Tiles of a colour are sorted heaviest first and dealt round-robin over every warp of the
    // persistent grid, so each warp gets one heavy tile before anyone gets a second, lighter one.
    uint32_t const tBegin = __ldg(p.colorTileBegin + color), tEnd = __ldg(p.colorTileBegin + color + 1);
    uint32_t const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    uint32_t const gwarp = warp * gridDim.x + blockIdx.x, gWarps = nWarps * gridDim.x;
    for (uint32_t T = tBegin + gwarp; T < tEnd; T += gWarps)
    {
        uint4 const td = __ldg(p.tiles + T);
        unsigned long long* tr = (trace && warp == 0 && T == tBegin + gwarp) ? trace : nullptr;
        if (tr && lane == 0)
            tr[4] = GlobalTimer() + (td.x & 0u);  // tile descriptor arrived
        DirectRecords src{p.records + static_cast<size_t>(td.x) * kBlockFloat4 + lane};
        ProcessTile<kChebyshev, kDamping, true>(p, td, stage, src, color, k, omega, lane, tr);
        if (tr && lane == 0)
            tr[7] = GlobalTimer();  // first tile of warp 0 finished
    }
}

template <bool kChebyshev, bool kDamping>
__global__ void __launch_bounds__(256, 3) StepKernel(const __grid_constant__ StepParams p)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float4* const stage = reinterpret_cast<float4*>(smemRaw) + (threadIdx.x >> 5) * p.stageEntries;
    unsigned int target = 0;
    uint32_t const gtid    = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t const gstride = gridDim.x * blockDim.x;
    for (int s = 0; s < p.substeps; ++s)
    {
        if (!p.skipPreStep)
            for (uint32_t i = gtid; i < p.ghostBegin; i += gstride)
                PreStepVertex<kChebyshev>(p, i, s);
        GridBarrier(p.barrier, target);
        for (int k = 0; k < p.iterations; ++k)
        {
            float const omega = kChebyshev ? __ldg(p.omega + p.iterBegin + k) : 1.f;
            for (int c = 0; c < p.nColors; ++c)
            {
                unsigned long long* tr = nullptr;
                if (p.trace != nullptr && k == p.traceIteration)
                {
                    tr = p.trace + (static_cast<size_t>(c) * gridDim.x + blockIdx.x) * kTraceStamps;
                    if (threadIdx.x == 0)
                        tr[0] = GlobalTimer();
                }
                SweepColor<kChebyshev, kDamping>(p, c, p.iterBegin + k, omega, stage, tr);
                GridBarrier(p.barrier, target, tr);
            }
        }
    }
    if (!p.skipPostStep)
        for (uint32_t i = gtid; i < p.ghostBegin; i += gstride)
            PostStepVertex(p, i);
}

}  // namespace vbdx
