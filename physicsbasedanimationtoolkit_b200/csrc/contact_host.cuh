// Host driver of the contact path: owns the device buffers of the LBVH and of the vertex-triangle active
// set and sequences the kernels of lbvh.cuh / contact.cuh.  Mirrors the reference's
// gpu/impl/geometry/Bvh.{cu,cuh} (Build / ConstructBoxes) and
// gpu/impl/contact/VertexTriangleMixedCcdDcd.{cu,cuh} (Initialize / Update / FinalizeActiveSet).
#pragma once

#include "contact.cuh"
#include "device_buffer.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

namespace vbdx {

// A fixed sequence of small launches recorded once as a CUDA graph and replayed: the active-set prologue is ~20 kernels
// of a few microseconds each, whose launch gaps and host cost exceed their work.  `key` = everything the recorded launches
// were given by value (pointers, dt, ...): a different key records anew.  VBDX_CONTACT_GRAPH=0 launches directly.
struct ReplayedLaunches {
    cudaGraphExec_t exec = nullptr;
    int64_t kernels      = 0;
    std::vector<uint64_t> key;
    ~ReplayedLaunches()
    {
        if (exec != nullptr)
            cudaGraphExecDestroy(exec);
    }
    ReplayedLaunches()                                   = default;
    ReplayedLaunches(ReplayedLaunches const&)            = delete;
    ReplayedLaunches& operator=(ReplayedLaunches const&) = delete;

    static bool Enabled()
    {
        static bool const on = [] {
            char const* e = std::getenv("VBDX_CONTACT_GRAPH");
            return e == nullptr || std::atoi(e) != 0;
        }();
        return on;
    }
    template <class Body>
    void Run(std::vector<uint64_t> const& k, cudaStream_t s, int64_t* launches, Body&& body)
    {
        cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
        bool const legacy = s == nullptr || s == cudaStreamLegacy;  // cannot be captured
        if (!Enabled() || legacy || cudaStreamIsCapturing(s, &status) != cudaSuccess || status != cudaStreamCaptureStatusNone)
        {
            body(launches);
            return;
        }
        if (exec == nullptr || k != key)
        {
            if (exec != nullptr)
                cudaGraphExecDestroy(exec), exec = nullptr;
            int64_t counted = 0;
            VBDX_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
            cudaGraph_t graph = nullptr;
            try
            {
                body(&counted);
            }
            catch (...)
            {
                cudaStreamEndCapture(s, &graph);
                if (graph != nullptr)
                    cudaGraphDestroy(graph);
                throw;
            }
            VBDX_CUDA(cudaStreamEndCapture(s, &graph));
            cudaError_t const e = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess)
            {
                exec = nullptr;
                VBDX_CUDA(e);
            }
            kernels = counted;
            key     = k;
        }
        VBDX_CUDA(cudaGraphLaunch(exec, s));
        *launches += kernels;
    }
};

inline uint64_t KeyOf(void const* p) { return static_cast<uint64_t>(reinterpret_cast<uintptr_t>(p)); }
inline uint64_t KeyOf(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, sizeof(u));
    return u;
}

// workspace of the one-launch radix sort (lbvh.cuh): zeroed barrier counter + chunk sums, the SM count as the CTA limit
inline void InitSortSync(DevBuf<uint32_t>& aux, RadixSortSync& sync, int64_t* bytes)
{
    aux.Alloc(1 + kSortMaxCtas, bytes);
    VBDX_CUDA(cudaMemset(aux.p, 0, (1 + kSortMaxCtas) * sizeof(uint32_t)));
    int dev = 0, sms = 0;
    VBDX_CUDA(cudaGetDevice(&dev));
    VBDX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    sync.aux = aux.p, sync.maxCtas = sms;
}

struct DeviceBvh {
    uint32_t n = 0;
    DevBuf<uint32_t> codes, inds, codesTmp, indsTmp, counts, visits, sortAux;
    RadixSortSync sortSync;
    DevBuf<int32_t> child0, child1, parent, right0, right1;
    DevBuf<float4> nodeLo, nodeHi;
    DevBuf<int2> up;

    void Alloc(uint32_t nLeaves, int64_t* bytes)
    {
        n = nLeaves;
        codes.Alloc(n, bytes), inds.Alloc(n, bytes), codesTmp.Alloc(n, bytes), indsTmp.Alloc(n, bytes);
        counts.Alloc(RadixSortCountsSize(n), bytes);
        InitSortSync(sortAux, sortSync, bytes);
        size_t const ni = n > 1 ? n - 1 : 1;
        visits.Alloc(ni, bytes);
        VBDX_CUDA(cudaMemset(visits.p, 0, ni * sizeof(uint32_t)));  // arrival counters: never reset afterwards (BvhRefit)
        up.Alloc(2 * static_cast<size_t>(n), bytes);
        child0.Alloc(ni, bytes), child1.Alloc(ni, bytes), right0.Alloc(ni, bytes), right1.Alloc(ni, bytes);
        parent.Alloc(2 * static_cast<size_t>(n), bytes);
        nodeLo.Alloc(2 * static_cast<size_t>(n), bytes), nodeHi.Alloc(2 * static_cast<size_t>(n), bytes);
    }
    BvhView View()
    {
        BvhView v{};
        v.n = n, v.codes = codes.p, v.inds = inds.p;
        v.child[0] = child0.p, v.child[1] = child1.p, v.parent = parent.p;
        v.rightmost[0] = right0.p, v.rightmost[1] = right1.p;
        v.nodeLo = nodeLo.p, v.nodeHi = nodeHi.p, v.visits = visits.p, v.up = up.p;
        return v;
    }
    // internal boxes from per-primitive boxes, keeping the topology (Bvh::ConstructBoxes)
    void Refit(const float4* primLo, const float4* primHi, cudaStream_t s, int64_t* launches)
    {
        BvhRefit<<<Blocks(n, 256), 256, 0, s>>>(View(), primLo, primHi);
        *launches += 1;
    }
    // Bvh::Build: Morton codes of the box centroids, stable sort, hierarchy, boxes
    void Build(const float4* primLo, const float4* primHi, const WorldBox* world, cudaStream_t s, int64_t* launches)
    {
        MortonOfBoxes<<<Blocks(n, 256), 256, 0, s>>>(primLo, primHi, n, world, codes.p, inds.p);
        RadixSortPairs(codes.p, inds.p, codesTmp.p, indsTmp.p, n, counts.p, sortSync, s, launches);
        VBDX_CUDA(cudaMemsetAsync(parent.p, 0xff, 2 * static_cast<size_t>(n) * sizeof(int32_t), s));
        VBDX_CUDA(cudaMemsetAsync(up.p, 0xff, 2 * static_cast<size_t>(n) * sizeof(int2), s));
        if (n > 1)
            BvhHierarchy<<<Blocks(n - 1, 256), 256, 0, s>>>(View());
        *launches += 2;
        Refit(primLo, primHi, s, launches);
    }
};

struct ContactState {
    bool enabled = false;
    ContactMesh mesh{};
    uint32_t nCV = 0, nF = 0;
    int updateFrequency = 1;
    bool hasWorldBox    = false;
    float eps           = FLT_EPSILON;  // nearest-neighbour tie tolerance (VertexTriangleMixedCcdDcd.cuh: eps)
    // static (internal ids)
    DevBuf<int32_t> B, V;
    DevBuf<int4> F;
    DevBuf<float> XVA, FA;
    DevBuf<uint32_t> colorVertexBegin;
    // contact lists consumed by the sweep, and the iteration-start snapshots
    DevBuf<int32_t> fc;
    DevBuf<float4> snap;
    // detector
    DeviceBvh bvh;
    DevBuf<uint32_t> ids, idsTmp, codes, codesTmp, counts, scanScratch, flags, offsets, nActive, sortAux;
    RadixSortSync sortSync;
    DevBuf<float4> ptLo, ptHi, triLo, triHi;
    DevBuf<uint8_t> active;
    DevBuf<float> dupper;
    DevBuf<int32_t> av, nn;
    DevBuf<WorldBox> world;
    DevBuf<int> bounds;

    void Alloc(int64_t nV, int64_t* bytes, cudaStream_t s)
    {
        fc.Alloc(static_cast<size_t>(nV) * kMaxContacts, bytes);
        snap.Alloc(2 * static_cast<size_t>(nV), bytes);
        bvh.Alloc(nF, bytes);
        ids.Alloc(nCV, bytes), idsTmp.Alloc(nCV, bytes), codes.Alloc(nCV, bytes), codesTmp.Alloc(nCV, bytes);
        counts.Alloc(RadixSortCountsSize(nCV), bytes);
        InitSortSync(sortAux, sortSync, bytes);
        scanScratch.Alloc((static_cast<size_t>(nCV) + 1) / kScanTile + 2, bytes);
        flags.Alloc(nCV + 1, bytes), offsets.Alloc(nCV + 1, bytes), nActive.Alloc(1, bytes);
        ptLo.Alloc(nCV, bytes), ptHi.Alloc(nCV, bytes), triLo.Alloc(nF, bytes), triHi.Alloc(nF, bytes);
        active.Alloc(nCV, bytes), dupper.Alloc(nCV, bytes), av.Alloc(nCV, bytes);
        nn.Alloc(static_cast<size_t>(nCV) * kMaxContacts, bytes);
        world.Alloc(1, bytes), bounds.Alloc(6, bytes);
        // initial state of the detector (VertexTriangleMixedCcdDcd ctor, :21-50): ids = identity,
        // nothing active, av = -1, dupper = max; contact lists empty (gpu/impl/vbd/Integrator.cu:61)
        std::vector<uint32_t> iota(nCV);
        for (uint32_t k = 0; k < nCV; ++k)
            iota[k] = k;
        ids.Upload(iota.data(), nCV, s);
        std::vector<float> big(nCV, FLT_MAX);
        dupper.Upload(big.data(), nCV, s);
        VBDX_CUDA(cudaMemsetAsync(active.p, 0, nCV, s));
        VBDX_CUDA(cudaMemsetAsync(av.p, 0xff, nCV * sizeof(int32_t), s));
        VBDX_CUDA(cudaMemsetAsync(nn.p, 0xff, static_cast<size_t>(nCV) * kMaxContacts * sizeof(int32_t), s));
        VBDX_CUDA(cudaMemsetAsync(fc.p, 0xff, static_cast<size_t>(nV) * kMaxContacts * sizeof(int32_t), s));
        VBDX_CUDA(cudaMemsetAsync(nActive.p, 0, sizeof(uint32_t), s));
        VBDX_CUDA(cudaStreamSynchronize(s));  // the host vectors go out of scope
    }

    void SetWorldBox(const float lo[3], const float hi[3], cudaStream_t s)
    {
        WorldBox w;
        for (int d = 0; d < 3; ++d)
        {
            w.lo[d]  = lo[d];
            w.ext[d] = hi[d] - lo[d];
        }
        VBDX_CUDA(cudaMemcpyAsync(world.p, &w, sizeof(w), cudaMemcpyHostToDevice, s));
        VBDX_CUDA(cudaStreamSynchronize(s));
        hasWorldBox = true;
    }

    // VertexTriangleMixedCcdDcd::InitializeActiveSet with the predictor of gpu/impl/vbd/Integrator.cu:163-188
    ReplayedLaunches replayInit, replayNearest[2];
    void InitializeActiveSet(const float4* x, const float4* vel, const float4* aext, int64_t nV, float dt, cudaStream_t s, int64_t* launches)
    {
        replayInit.Run({KeyOf(x), KeyOf(vel), KeyOf(aext), static_cast<uint64_t>(nV), KeyOf(dt), hasWorldBox ? 1u : 0u}, s, launches,
                       [&](int64_t* n) { InitializeActiveSetNow(x, vel, aext, nV, dt, s, n); });
    }
    void NearestPass(const float4* x, int mode, cudaStream_t s, int64_t* launches)
    {
        replayNearest[mode != 0].Run({KeyOf(x), KeyOf(eps)}, s, launches, [&](int64_t* n) { NearestPassNow(x, mode, s, n); });
    }
    void InitializeActiveSetNow(const float4* x, const float4* vel, const float4* aext, int64_t nV, float dt, cudaStream_t s, int64_t* launches)
    {
        if (!hasWorldBox)
        {
            // the reference needs SetSceneBoundingBox from the caller; without it use the scene's bounds
            SceneBoundsReset<<<1, 32, 0, s>>>(bounds.p);
            SceneBoundsReduce<<<std::min(Blocks(nV, 256), 296), 256, 0, s>>>(x, static_cast<uint32_t>(nV), bounds.p);
            SceneBoundsFinish<<<1, 32, 0, s>>>(bounds.p, world.p);
            *launches += 3;
        }
        // swept vertices: boxes in the current order -> codes -> sort ids -> boxes in sorted order
        SweptPointBoxes<<<Blocks(nCV, 256), 256, 0, s>>>(mesh, ids.p, x, vel, aext, dt, ptLo.p, ptHi.p);
        MortonOfBoxes<<<Blocks(nCV, 256), 256, 0, s>>>(ptLo.p, ptHi.p, nCV, world.p, codes.p, nullptr);
        RadixSortPairs(codes.p, ids.p, codesTmp.p, idsTmp.p, nCV, counts.p, sortSync, s, launches);
        SweptPointBoxes<<<Blocks(nCV, 256), 256, 0, s>>>(mesh, ids.p, x, vel, aext, dt, ptLo.p, ptHi.p);
        // swept triangles and their BVH
        TriangleBoxes<<<Blocks(nF, 256), 256, 0, s>>>(mesh, x, vel, aext, dt, triLo.p, triHi.p);
        *launches += 4;
        bvh.Build(triLo.p, triHi.p, world.p, s, launches);
        // overlaps -> active flags, warm-start radii; compaction in sorted order
        MarkActive<<<Blocks(nCV, 128), 128, 0, s>>>(mesh, bvh.View(), ids.p, ptLo.p, ptHi.p, triLo.p, triHi.p, active.p, dupper.p);
        ActiveFlags<<<Blocks(nCV + 1, 256), 256, 0, s>>>(ids.p, active.p, nCV, flags.p);
        ExclusiveScanU32(flags.p, offsets.p, nCV + 1, scanScratch.p, s);
        CompactActive<<<Blocks(nCV, 256), 256, 0, s>>>(ids.p, flags.p, offsets.p, nCV, av.p, nActive.p);
        *launches += 6;
    }

    // refit with current positions, then k-NN: mode 0 = UpdateActiveSet (+ contact lists), 1 = FinalizeActiveSet
    void NearestPassNow(const float4* x, int mode, cudaStream_t s, int64_t* launches)
    {
        TriangleBoxes<<<Blocks(nF, 256), 256, 0, s>>>(mesh, x, nullptr, nullptr, 0.f, triLo.p, triHi.p);
        bvh.Refit(triLo.p, triHi.p, s, launches);
        if (mode == 0)
        {
            FillI32<<<Blocks(static_cast<int64_t>(nCV) * kMaxContacts, 256), 256, 0, s>>>(nn.p, -1, static_cast<size_t>(nCV) * kMaxContacts);
            ++*launches;
        }
        NearestTriangles<<<std::min(Blocks(static_cast<int64_t>(nCV) * 32, 128), 2368), 128, 0, s>>>(mesh, bvh.View(), av.p, nActive.p, x, dupper.p, eps, mode, nn.p, fc.p, active.p);  // a warp per vertex
        *launches += 2;
    }
};

}  // namespace vbdx
