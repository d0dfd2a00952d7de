// C-ABI layer and host driver of the B200 VBD integrator (include/vbdx.h).
// Host code is C++20; every device computation is a hand-written sm_100a kernel from
// setup_kernels.cuh / step_kernel.cuh.  No Thrust/CUB dispatch, no CPU fallback.
#include "../../include/vbdx.h"

#include "device_buffer.cuh"
#include "anderson.cuh"
#include "accelerators.cuh"
#include "broyden.cuh"
#include "contact_host.cuh"
#include "geometry_api.cuh"
#include "diagnostics.cuh"
#include "setup_kernels.cuh"
#include "step_kernel.cuh"
#include "step_kernel_pipe.cuh"
#include "step_kernel_flow.cuh"
#include "step_kernel_tma.cuh"
#include "vbdx_internal.h"

#include <nvtx3/nvToolsExt.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace vbdx {

// NVTX ranges named like the reference's Tracy zones (PBAT_PROFILE_CUDA_NAMED_SCOPE in gpu/impl/vbd/Integrator.cu:84,93,165,...):
// a profiler timeline (nsys, ncu --nvtx) shows the same phase names on both implementations.  Host-side ranges around the
// enqueue of each phase; the header-only NVTX v3 costs a pointer check when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(char const* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(NvtxRange const&)            = delete;
    NvtxRange& operator=(NvtxRange const&) = delete;
};

using StepKernelFn = void (*)(StepParams);
using TmaKernelFn  = void (*)(TmaParams);
using PipeKernelFn = void (*)(PipeParams);

constexpr int kClusterCtas = 8;  // portable maximum cluster size

struct Integrator {
    int device = 0, smCount = 0;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    cudaEvent_t evBegin = nullptr, evEnd = nullptr;
    int64_t nV = 0, nT = 0;
    int64_t deviceBytes = 0, kernelLaunches = 0;
    double lastStepMs = 0;
    bool stepTimed    = false;

    Plan plan;
    int gridBlocks = 0, blockThreads = 256;
    int variant = VBDX_KERNEL_DIRECT;
    bool dataflow    = true;   // barrier-free sweeps where they apply (PipeParams::dataflow); VBDX_DATAFLOW=0 turns them off
    bool flowKernel  = true;   // ... run by the lean kernel of step_kernel_flow.cuh; VBDX_FLOW=0 selects StepKernelPipe<.., dataflow>
    bool usedDataflow  = false;
    unsigned int dfTag = 1;    // write numbers used by earlier launches (single GPU)
    bool clusterMode = false;  // the pipelined kernel launched as ONE thread-block cluster (VBDX_KERNEL_CLUSTER)
    bool clusterOnly = false;  // ... because the caller asked for it (kernel_variant): whole steps too; chosen by default, it serves partial launches only
    uint32_t ringSlots = 0, maxTileIters = 1;
    size_t smemBytes   = 0;
    int64_t nRecordSlots = 0;

    // parameters
    int strategy = VBDX_INIT_ADAPTIVE_PBAT, acceleration = VBDX_ACCEL_NONE, omegaMode = 0;
    double kD = 0, detHZero = 1e-7, rho = 1;
    int flags = 0;
    std::vector<float> omegaHost;
    int omegaIterations = -1;

    // setup products kept for introspection (caller numbering)
    DevBuf<int32_t> dE;
    DevBuf<uint32_t> dPtr, dAdj;
    DevBuf<double> dJinv, dVol, dMass, dLame;  // dLame empty = the reference's default material everywhere
    double mu0 = 0, lambda0 = 0;
    // Anderson acceleration (anderson.cuh)
    int window = 5;
    double nesterovL = 1.0;
    int nesterovStart = 3;
    double trEta = 0.2, trTau = 2.0;
    bool trCurved = true;
    int material = VBDX_MATERIAL_STABLE_NEO_HOOKEAN;
    int lineSearch = 0;  // vbdx_set_line_search_guard
    std::vector<int64_t> batchOffsets;  // vbdx_create_batch: first vertex of every scene, then nV
    DevBuf<float4> dAndVec;   // xkm1, Gkm1, Fkm1, Fk, DF[m], DG[m]
    DevBuf<double> dAndSmall; // gram m*m, scratch 2m, alpha m
    // objective / gradient evaluation (diagnostics.cuh)
    DevBuf<double> dObjX, dObjXt, dObjGrad, dObjVal;
    DevBuf<int32_t> dColor;
    DevBuf<uint8_t> dIsDbc;
    std::vector<int64_t> colors;

    // static sweep data
    DevBuf<float4> dRecords;
    DevBuf<TileDesc> dTiles;
    DevBuf<uint32_t> dCtaRange, dCtaBlockBegin, dRingIds, dColorTileBegin;
    uint32_t stageEntries = 32;
    DevBuf<int32_t> dNew2Old, dOld2New;
    // schedule of the lean barrier-free kernel (step_kernel_flow.cuh; BuildFlowSchedule)
    DevBuf<uint32_t> dFlowWarpBegin, dFlowIds;
    DevBuf<uint4> dFlowTiles;
    int flowWarps = 16, flowGridBlocks = 0;  // launch shape of the lean kernel (no barrier warp, no id buffers)
    size_t flowSmemBytes = 0;
    uint32_t flowActiveWarps = 0;  // warps of the lean kernel's grid that have tiles
    uint32_t flowBoundarySms = 0, flowBoundaryWarps = 0;  // domain decomposition: CTAs (and their warps in use) that run the tiles next to another GPU
    DevBuf<float4> dHist4;         // contact under barrier-free sweeps: every vertex' last four writes (step_kernel.cuh, HistSlot)
    DevBuf<unsigned int> dSweepDone;  // ... and per sweep of a launch the warps that have finished it (the sweep-lag bound)
    void BuildFlowSchedule();

    // state
    DevBuf<float4> dPos, dHist, dXtildeM, dXt, dVel, dVtm1, dAext;
    DevBuf<float> dOmega;
    DevBuf<unsigned int> dBarrier;
    DevBuf<double> dStaging;  // 3 nV doubles
    DevBuf<unsigned long long> dTrace;
    int traceIteration = -1;
    ContactState contact;
    // domain decomposition
    int distRank = 0, distWorld = 1;
    unsigned int distEpoch = 0, peerMask = 0;
    unsigned int distTag = 2;         // tags of ghost writes used so far (0 = never written)
    int64_t nGhost = 0;
    uint32_t peerGhostExt[8] = {}, peerGhostBegin[8] = {}, peerNGhost[8] = {};
    DevBuf<unsigned int> dDistFlags;  // [0..7] peers' epochs, [8] local release, [9] error
    DevBuf<uint32_t> dSendPtr, dSendDst;
    float4* peerPos[8]        = {};
    uint32_t peerPOff[8]      = {};
    unsigned int* peerFlags[8] = {};
    void* ipcOpened[16]        = {};
    double muC = 1e6, muF = 0.3, epsv = 1e-3;

    ~Integrator()
    {
        for (void* q : ipcOpened)
            if (q)
                cudaIpcCloseMemHandle(q);
        if (evBegin)
            cudaEventDestroy(evBegin);
        if (evEnd)
            cudaEventDestroy(evEnd);
        if (ownStream)
            cudaStreamDestroy(ownStream);
    }

    StepKernelFn Kernel() const
    {
        bool const cheb = acceleration == VBDX_ACCEL_CHEBYSHEV;
        bool const damp = kD != 0.0 || contact.enabled;  // the "extras" variants carry damping and contact
        if (cheb)
            return damp ? StepKernel<true, true> : StepKernel<true, false>;
        return damp ? StepKernel<false, true> : StepKernel<false, false>;
    }

    // dataflowSweep = the round-1 kernel's barrier-free instantiation (VBDX_FLOW=0; it has no damping / contact form)
    PipeKernelFn KernelPipe(bool dataflowSweep = false) const
    {
        bool const cheb = acceleration == VBDX_ACCEL_CHEBYSHEV;
        bool const damp = kD != 0.0 || contact.enabled;  // the "extras" variants carry damping and contact
        if (material == VBDX_MATERIAL_STVK)
        {
            if (dataflowSweep && !damp)
                return cheb ? StepKernelPipe<true, false, true, true> : StepKernelPipe<false, false, true, true>;
            if (cheb)
                return damp ? StepKernelPipe<true, true, true> : StepKernelPipe<true, false, true>;
            return damp ? StepKernelPipe<false, true, true> : StepKernelPipe<false, false, true>;
        }
        if (dataflowSweep && !damp)
            return cheb ? StepKernelPipe<true, false, false, true> : StepKernelPipe<false, false, false, true>;
        if (cheb)
            return damp ? StepKernelPipe<true, true> : StepKernelPipe<true, false>;
        return damp ? StepKernelPipe<false, true> : StepKernelPipe<false, false>;
    }

    // the lean barrier-free kernel (step_kernel_flow.cuh): Stable Neo-Hookean, with or without Chebyshev / Rayleigh damping
    PipeKernelFn KernelFlow() const
    {
        bool const cheb = acceleration == VBDX_ACCEL_CHEBYSHEV;
        bool const damp = kD != 0.0 || contact.enabled;  // the damping variants carry the contact term
        if (material == VBDX_MATERIAL_STVK)  // (single GPU; domain decomposition keeps the pipelined kernel for StVK)
        {
            if (cheb)
                return damp ? StepKernelFlow<true, true, false, true> : StepKernelFlow<true, false, false, true>;
            return damp ? StepKernelFlow<false, true, false, true> : StepKernelFlow<false, false, false, true>;
        }
        if (distWorld > 1 || nGhost > 0)
        {
            if (cheb)
                return damp ? StepKernelFlow<true, true, true> : StepKernelFlow<true, false, true>;
            return damp ? StepKernelFlow<false, true, true> : StepKernelFlow<false, false, true>;
        }
        if (cheb)
            return damp ? StepKernelFlow<true, true, false> : StepKernelFlow<true, false, false>;
        return damp ? StepKernelFlow<false, true, false> : StepKernelFlow<false, false, false>;
    }
    bool UseFlow(int iterations) const
    {
        return dataflow && flowKernel && dFlowTiles.p != nullptr && variant == VBDX_KERNEL_PIPELINED && !clusterOnly &&
               (!contact.enabled || dHist4.p != nullptr) &&
               (material == VBDX_MATERIAL_STABLE_NEO_HOOKEAN || (distWorld == 1 && nGhost == 0)) && iterations > 0;
    }

    TmaKernelFn KernelTma() const
    {
        bool const cheb = acceleration == VBDX_ACCEL_CHEBYSHEV;
        bool const damp = kD != 0.0 || contact.enabled;  // the "extras" variants carry damping and contact
        if (cheb)
            return damp ? StepKernelTma<true, true> : StepKernelTma<true, false>;
        return damp ? StepKernelTma<false, true> : StepKernelTma<false, false>;
    }

    void Create(vbdx_data_desc const& d);
    void Step(double dt, int iterations, int substeps, bool sync);
    void PrepareOmega(int iterations);
    StepParams MakeParams(double sdt, int iterations, int substeps);
    void LaunchStepKernel(StepParams const& q);
    void RunStep(StepParams const& p, double dt, int iterations, int substeps, bool sync);
    void LaunchPreStep(StepParams const& q);
    void CheckAsyncErrors();
    bool WillSweepBarrierFree(int iterations) const
    {
        return UseFlow(iterations) ||
               (dataflow && variant == VBDX_KERNEL_PIPELINED && !clusterMode && !contact.enabled && kD == 0.0 && iterations > 0);
    }
    void AndersonStep(StepParams const& p, double dt, int iterations, int substeps);
    void BroydenStep(StepParams const& p, double dt, int iterations, int substeps);
    void NesterovStep(StepParams const& p, double dt, int iterations, int substeps);
    void TrustRegionStep(StepParams const& p, double dt, int iterations, int substeps);
    double ObjectiveOfState(double sdt);
    // contact hooks shared by the windowed accelerators (same sequence as RunStep's contact branch)
    void ContactBeginStep(StepParams const& p, double dt)
    {
        if (contact.enabled)
            contact.InitializeActiveSet(dPos.p + p.pOff, dVel.p, dAext.p, nV, static_cast<float>(dt), stream, &kernelLaunches);
    }
    void ContactAfterPreStep(StepParams const& p, int s)
    {
        if (contact.enabled && s % contact.updateFrequency == 0)
            contact.NearestPass(dPos.p + p.pOff, 0, stream, &kernelLaunches);
    }
    void ContactEndStep(StepParams const& p)
    {
        if (contact.enabled)
            contact.NearestPass(dPos.p + p.pOff, 1, stream, &kernelLaunches);
    }
    float4* SnapNext(int k) { return contact.enabled ? contact.snap.p + static_cast<size_t>((k + 1) & 1) * nV : nullptr; }
    void StepPartial(double sdt, int kBegin, int kEnd, int totalIterations, int flags);
    void Objective(const double* xk, const double* xtilde, double dt, double* f, double* grad);
    template <class T>
    void SetVertexField(T const* src, int64_t n, float4* dst0, float4* dst1, bool rows = false, bool sync = true);
    template <class T>
    void GetVertexField(float4 const* src, T* dst, int64_t n, bool rows = false, bool sync = true);
};

void Integrator::Create(vbdx_data_desc const& d)
{
    Require(d.abi_version == VBDX_ABI_VERSION && d.struct_size == sizeof(vbdx_data_desc),
            "vbdx_data_desc: ABI version / struct size mismatch");
    Require(d.nV > 0 && d.nT > 0 && d.X && d.E, "need a volume mesh: X (3 x nV) and E (4 x nT)");
    Require(d.nV < (int64_t(1) << 31) - 1 && d.nT < (int64_t(1) << 29), "mesh too large for 32-bit device indices");
    Require(d.acceleration >= VBDX_ACCEL_NONE && d.acceleration <= VBDX_ACCEL_TRUST_REGION, "unknown acceleration strategy");
    if (d.acceleration == VBDX_ACCEL_NESTEROV)
    {
        // sim/vbd/Data.cpp:284-293
        Require(d.nesterov_L > 0, "Expected L > 0");
        Require(d.nesterov_start >= 0, "Expected start >= 0");
        Require(d.nGhosts == 0, "Nesterov acceleration is not combined with domain decomposition");
        nesterovL = d.nesterov_L, nesterovStart = d.nesterov_start;
    }
    if (d.acceleration == VBDX_ACCEL_TRUST_REGION)
    {
        // sim/vbd/Data.cpp:294-303
        Require(d.tr_eta >= 0, "Expected eta >= 0");
        Require(d.tr_tau > 1, "Expected tau > 1");
        Require(d.nGhosts == 0, "trust-region acceleration is not combined with domain decomposition");
        if (d.nF > 0 && d.nCV > 0)
            throw Error(VBDX_UNSUPPORTED, "the trust-region accelerator's objective carries no contact term here: use it without a collision mesh");
        trEta = d.tr_eta, trTau = d.tr_tau, trCurved = d.tr_curved != 0;
    }
    if (d.acceleration == VBDX_ACCEL_ANDERSON || d.acceleration == VBDX_ACCEL_BROYDEN)
    {
        // sim/vbd/Data.cpp:277-283 (window >= 1); the device solver keeps the window's Gram matrix in registers
        Require(d.window_size >= 1, "Expected window size >= 1");
        if (d.window_size > kMaxAndersonWindow)
            throw Error(VBDX_UNSUPPORTED, "Anderson/Broyden windows larger than 16 are not supported");
        Require(d.nGhosts == 0, "Anderson/Broyden acceleration is not combined with domain decomposition yet");
        window = d.window_size;
    }
    Require(d.material == VBDX_MATERIAL_STABLE_NEO_HOOKEAN || d.material == VBDX_MATERIAL_STVK, "unknown material");
    if (d.material == VBDX_MATERIAL_STVK && d.kernel_variant != VBDX_KERNEL_DEFAULT && d.kernel_variant != VBDX_KERNEL_PIPELINED &&
        d.kernel_variant != VBDX_KERNEL_CLUSTER)
        throw Error(VBDX_UNSUPPORTED, "St. Venant-Kirchhoff runs on the pipelined step kernel only");
    material = d.material;
    if (d.acceleration == VBDX_ACCEL_CHEBYSHEV)
        Require(d.rho > 0 && d.rho < 1, "Expected 0 < rho < 1");  // sim/vbd/Data.cpp:272-276
    Require(d.strategy >= 0 && d.strategy <= VBDX_INIT_ADAPTIVE_PBAT, "unknown initialization strategy");
    Require(d.nDbc >= 0 && (d.nDbc == 0 || d.dbc), "dbc pointer missing");
    nV = d.nV, nT = d.nT;
    strategy = d.strategy, acceleration = d.acceleration, omegaMode = d.omega_mode;
    kD = d.kD, detHZero = d.detHZero, rho = d.rho, flags = d.flags;

    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    {
        cudaGetLastError();
        throw Error(VBDX_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
    }
    if (d.device >= 0)
    {
        Require(d.device < count, "device ordinal out of range");
        VBDX_CUDA(cudaSetDevice(d.device));
    }
    VBDX_CUDA(cudaGetDevice(&device));
    cudaDeviceProp prop;
    VBDX_CUDA(cudaGetDeviceProperties(&prop, device));
    smCount = prop.multiProcessorCount;
    if (!prop.cooperativeLaunch)
        throw Error(VBDX_UNSUPPORTED, "device does not support cooperative launches");
    VBDX_CUDA(cudaStreamCreateWithFlags(&ownStream, cudaStreamNonBlocking));
    stream = ownStream;
    VBDX_CUDA(cudaEventCreate(&evBegin));
    VBDX_CUDA(cudaEventCreate(&evEnd));

    // ---- host-side validation and narrowing of the connectivity
    std::vector<int32_t> E32(static_cast<size_t>(4 * nT));
    for (int64_t k = 0; k < 4 * nT; ++k)
    {
        int64_t const v = d.E[k];
        Require(v >= 0 && v < nV, "element index out of range");
        E32[k] = static_cast<int32_t>(v);
    }
    std::vector<uint8_t> isDbc(static_cast<size_t>(nV), 0);
    for (int64_t k = 0; k < d.nDbc; ++k)
    {
        Require(d.dbc[k] >= 0 && d.dbc[k] < nV, "Dirichlet vertex index out of range");
        isDbc[d.dbc[k]] = 1;
    }
    Require(d.nGhosts >= 0 && (d.nGhosts == 0 || d.ghosts), "ghosts pointer missing");
    for (int64_t k = 0; k < d.nGhosts; ++k)
    {
        Require(d.ghosts[k] >= 0 && d.ghosts[k] < nV, "ghost vertex index out of range");
        Require(isDbc[d.ghosts[k]] == 0, "a ghost vertex must not be a Dirichlet vertex or listed twice (constrained vertices never change: keep them as plain Dirichlet vertices on every rank)");
        isDbc[d.ghosts[k]] = 2;  // never swept, and never touched by the pre-step: the owner GPU writes it
    }
    if (d.colors)
    {
        colors.assign(d.colors, d.colors + nV);
        for (int64_t c : colors)
            Require(c >= 0 && c < (int64_t(1) << 18), "vertex colours must be in [0, 2^18)");
    }
    else
        GreedyColorMesh(nV, nT, d.E, d.ordering, d.selection, colors);
    std::vector<int32_t> color32(colors.begin(), colors.end());

    // ---- device: CSR, element data
    dE.Alloc(4 * nT, &deviceBytes);
    dE.Upload(E32.data(), E32.size(), stream);
    DevBuf<double> dX;
    dX.Alloc(3 * nV);
    dX.Upload(d.X, 3 * nV, stream);
    dColor.Alloc(nV, &deviceBytes);
    dColor.Upload(color32.data(), nV, stream);
    dIsDbc.Alloc(nV, &deviceBytes);
    dIsDbc.Upload(isDbc.data(), nV, stream);

    dPtr.Alloc(nV + 1, &deviceBytes);
    dAdj.Alloc(4 * nT, &deviceBytes);
    DevBuf<uint32_t> dCursor, dScratch, dErr;
    dCursor.Alloc(nV + 1);
    dScratch.Alloc((nV + 1 + kScanTile - 1) / kScanTile + 1);
    dErr.Alloc(1);
    VBDX_CUDA(cudaMemsetAsync(dCursor.p, 0, (nV + 1) * sizeof(uint32_t), stream));
    VBDX_CUDA(cudaMemsetAsync(dErr.p, 0, sizeof(uint32_t), stream));
    CountIncidences<<<Blocks(4 * nT, 256), 256, 0, stream>>>(dE.p, 4 * nT, dCursor.p);
    ExclusiveScanU32(dCursor.p, dPtr.p, nV + 1, dScratch.p, stream);
    VBDX_CUDA(cudaMemsetAsync(dCursor.p, 0, (nV + 1) * sizeof(uint32_t), stream));
    FillIncidences<<<Blocks(4 * nT, 256), 256, 0, stream>>>(dE.p, 4 * nT, dPtr.p, dCursor.p, dAdj.p);
    SortRows<<<Blocks(nV, 128), 128, 0, stream>>>(dPtr.p, dAdj.p, nV);
    dJinv.Alloc(9 * nT, &deviceBytes);
    dVol.Alloc(nT, &deviceBytes);
    ElementQuantities<<<Blocks(nT, 128), 128, 0, stream>>>(dX.p, dE.p, nT, dColor.p, dIsDbc.p, dJinv.p, dVol.p, dErr.p);
    kernelLaunches += 7;

    DevBuf<double> dRho;
    // default material: Y = 1e6, nu = 0.45 (sim/vbd/Data.cpp:199-205, physics/HyperElasticity.cpp:6-11)
    double const Y = 1e6, nu = 0.45;
    double const muDefault = Y / (2. * (1. + nu)), lamDefault = (Y * nu) / ((1. + nu) * (1. - 2. * nu));
    if (d.lame)
    {
        for (int64_t e = 0; e < nT; ++e)
            Require(d.lame[2 * e + 1] != 0.0, "lambda must be non-zero (alpha = 1 + mu/lambda)");
        dLame.Alloc(2 * nT, &deviceBytes);
        dLame.Upload(d.lame, 2 * nT, stream);
    }
    mu0 = muDefault, lambda0 = lamDefault;
    dMass.Alloc(nV, &deviceBytes);
    if (d.m)
        dMass.Upload(d.m, nV, stream);
    else
    {
        if (d.rhoe)
        {
            dRho.Alloc(nT);
            dRho.Upload(d.rhoe, nT, stream);
        }
        VertexMass<<<Blocks(nV, 128), 128, 0, stream>>>(dPtr.p, dAdj.p, dVol.p, dRho.p, 1e3, dMass.p, nV);
        ++kernelLaunches;
    }

    std::vector<uint32_t> ptrHost(static_cast<size_t>(nV + 1)), adjHost(static_cast<size_t>(4 * nT));
    uint32_t err = 0;
    dPtr.Download(ptrHost.data(), nV + 1, stream);
    dAdj.Download(adjHost.data(), 4 * nT, stream);
    dErr.Download(&err, 1, stream);
    VBDX_CUDA(cudaStreamSynchronize(stream));
    if (err & 1u)
        throw Error(VBDX_INVALID_ARGUMENT, "inverted or degenerate tetrahedron in the rest mesh (det J <= 1e-10)");
    if (err & 2u)
        throw Error(VBDX_INVALID_ARGUMENT, "invalid colouring: two swept vertices of one tetrahedron share a colour");

    // ---- host plan (tiles, ring lists), then the launch shape, then the tile -> CTA partition
    bool const cheb0    = acceleration == VBDX_ACCEL_CHEBYSHEV;
    int const tileIters = d.tile_iters > 0 ? d.tile_iters : 8;
    try
    {
        BuildPlan(nV, E32.data(), ptrHost.data(), adjHost.data(), colors.data(), isDbc.data(), d.X, tileIters,
                  (flags & VBDX_FLAG_NATURAL_VERTEX_ORDER) != 0, material == VBDX_MATERIAL_STVK ? 2 : 1, d.n_colors, plan);
    }
    catch (std::length_error const& e)
    {
        throw Error(VBDX_UNSUPPORTED, e.what());
    }
    nRecordSlots = plan.nBlocks * 32;
    stageEntries = static_cast<uint32_t>(std::max(32, plan.maxRingPerTile));
    // the damping variant can be switched on later (SetRayleighDampingCoefficient), so size the
    // persistent grid for the least-resident variant of this acceleration mode
    variant = d.kernel_variant;
    Require(variant >= VBDX_KERNEL_DEFAULT && variant <= VBDX_KERNEL_CLUSTER, "unknown kernel variant");
    if (variant == VBDX_KERNEL_CLUSTER)
        Require(d.nGhosts == 0, "the cluster kernel variant is single-GPU only");
    // (every kernel is allowed the device's maximum of dynamic shared memory: the attribute belongs to the FUNCTION, not to
    // a handle -- a second handle with smaller tiles must not lower it under the first one's launches)
    int maxOptin = 0;
    VBDX_CUDA(cudaDeviceGetAttribute(&maxOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    maxTileIters = 1;
    for (TileDesc const& t : plan.tiles)
        maxTileIters = std::max(maxTileIters, TileIters(t.meta));
    // pipelined kernel: compute warps + one barrier warp per CTA.  Default: one CTA of 16 compute warps per SM
    // (fewer arrivals at the grid barrier) when its tile buffers fit in shared memory, else 8 compute warps.
    // (as many compute warps as the per-warp tile buffers leave room for: a mesh whose largest tile stages 288 instead of
    // 256 vertices must get 15 warps, not half the machine)
    int pipeWarps = d.consumer_warps > 0 ? std::min(d.consumer_warps, kPipeMaxThreads / 32 - 1) : 16;
    if (d.consumer_warps <= 0)
        while (pipeWarps > 1 && PipeSmemBytes(plan.nColors, pipeWarps, stageEntries, maxTileIters) > static_cast<size_t>(maxOptin))
            --pipeWarps;
    int const pipeThreads = pipeWarps * 32 + 32;
    flowWarps             = d.consumer_warps > 0 ? std::min(d.consumer_warps, kFlowMaxThreads / 32) : kFlowMaxThreads / 32;
    while (flowWarps > 1 && FlowSmemBytes(flowWarps, stageEntries, maxTileIters) > static_cast<size_t>(maxOptin))
        --flowWarps;
    flowSmemBytes = FlowSmemBytes(flowWarps, stageEntries, maxTileIters);
    size_t const pipeSmem = PipeSmemBytes(plan.nColors, pipeWarps, stageEntries, maxTileIters);
    if (variant == VBDX_KERNEL_DEFAULT)
    {
        variant = pipeSmem <= static_cast<size_t>(maxOptin) || material == VBDX_MATERIAL_STVK ? VBDX_KERNEL_PIPELINED : VBDX_KERNEL_DIRECT;
        // Tiny meshes: when every colour leaves at least half the warps of ONE thread-block cluster idle, a phase of the
        // BARRIER sweep is pure latency; running the same kernel as one cluster replaces the grid barrier (fence + atomic +
        // poll through L2) by the hardware cluster barrier.  Measured: 4-5 % per step on a 10 k-tet mesh (0.507 vs 0.528 ms),
        // a loss from ~20 k tets on (8 SMs of gather bandwidth instead of 148), hence the tight threshold.  Since round 2
        // this shape serves the launches that keep barriers (partial launches of the accelerators and traces); whole steps
        // run the lean barrier-free kernel, which is faster at every size (one cube: 0.19 vs 0.37 ms; 10 k tets: 0.31 vs 0.51).
        uint32_t maxColorTiles = 0;
        for (int32_t c = 0; c < plan.nColors; ++c)
            maxColorTiles = std::max(maxColorTiles, plan.colorTileBegin[c + 1] - plan.colorTileBegin[c]);
        if (variant == VBDX_KERNEL_PIPELINED && d.nGhosts == 0 && 2 * maxColorTiles <= static_cast<uint32_t>(kClusterCtas * pipeWarps))
            variant = VBDX_KERNEL_CLUSTER;
    }
    // the cluster variant IS the pipelined kernel, launched as one cluster and told to use the cluster barrier
    if (char const* e = std::getenv("VBDX_DATAFLOW"))
        dataflow = std::atoi(e) != 0;
    if (char const* e = std::getenv("VBDX_FLOW"))
        flowKernel = std::atoi(e) != 0;
    clusterOnly = d.kernel_variant == VBDX_KERNEL_CLUSTER;
    clusterMode = variant == VBDX_KERNEL_CLUSTER;
    if (clusterMode)
    {
        if (pipeSmem > static_cast<size_t>(maxOptin))
            throw Error(VBDX_UNSUPPORTED, "per-warp tile buffers do not fit in shared memory (lower tile_iters)");
        variant = VBDX_KERNEL_PIPELINED;
    }
    int perSm = 1 << 30;
    if (variant == VBDX_KERNEL_PIPELINED)
    {
        blockThreads = pipeThreads;
        smemBytes    = pipeSmem;
        if (smemBytes > static_cast<size_t>(maxOptin))
            throw Error(VBDX_UNSUPPORTED, "per-warp tile buffers do not fit in shared memory (lower tile_iters)");
        bool const stvk = material == VBDX_MATERIAL_STVK;
        for (PipeKernelFn fn :
             {stvk ? (cheb0 ? StepKernelPipe<true, false, true> : StepKernelPipe<false, false, true>)
                   : (cheb0 ? StepKernelPipe<true, false> : StepKernelPipe<false, false>),
              stvk ? (cheb0 ? StepKernelPipe<true, true, true> : StepKernelPipe<false, true, true>)
                   : (cheb0 ? StepKernelPipe<true, true> : StepKernelPipe<false, true>),
              stvk ? (cheb0 ? StepKernelPipe<true, false, true, true> : StepKernelPipe<false, false, true, true>)
                   : (cheb0 ? StepKernelPipe<true, false, false, true> : StepKernelPipe<false, false, false, true>)})
        {
            VBDX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, maxOptin));
            int n = 0;
            VBDX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, blockThreads, smemBytes));
            perSm = std::min(perSm, n);
        }
        // the lean barrier-free kernel has its own launch shape: no barrier warp, no id buffers
        if (flowSmemBytes <= static_cast<size_t>(maxOptin) && !(stvk && d.nGhosts > 0))
        {
            int flowPerSm = 1 << 30;
            std::vector<PipeKernelFn> fns;
            if (stvk)
                fns = {cheb0 ? StepKernelFlow<true, false, false, true> : StepKernelFlow<false, false, false, true>,
                       cheb0 ? StepKernelFlow<true, true, false, true> : StepKernelFlow<false, true, false, true>};
            else
                fns = {cheb0 ? StepKernelFlow<true, false, false> : StepKernelFlow<false, false, false>,
                       cheb0 ? StepKernelFlow<true, true, false> : StepKernelFlow<false, true, false>,
                       cheb0 ? StepKernelFlow<true, false, true> : StepKernelFlow<false, false, true>,
                       cheb0 ? StepKernelFlow<true, true, true> : StepKernelFlow<false, true, true>};
            for (PipeKernelFn fn : fns)
            {
                VBDX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, maxOptin));
                int n = 0;
                VBDX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, flowWarps * 32, flowSmemBytes));
                flowPerSm = std::min(flowPerSm, n);
            }
            flowGridBlocks = std::max(0, flowPerSm) * smCount;
        }
    }
    else
    if (variant == VBDX_KERNEL_TMA)
    {
        int const consumers = d.consumer_warps > 0 ? std::min(d.consumer_warps, kTmaMaxThreads / 32 - kProducerWarps) : 15;
        blockThreads        = (consumers + kProducerWarps) * 32;
        size_t const fixed  = TmaSmemBytes(0, plan.nColors, consumers, stageEntries);
        if (fixed + 2 * (kBlockBytes + 16) > static_cast<size_t>(maxOptin))
            throw Error(VBDX_UNSUPPORTED, "1-ring staging does not fit in shared memory (vertex valence too high)");
        uint32_t const maxSlots = static_cast<uint32_t>((maxOptin - fixed) / (kBlockBytes + 16));
        ringSlots = d.ring_slots > 0 ? static_cast<uint32_t>(d.ring_slots) : maxSlots;
        ringSlots = std::min(ringSlots, maxSlots);
        if (ringSlots < 2u)
            throw Error(VBDX_UNSUPPORTED, "not enough shared memory for the record ring");
        smemBytes = TmaSmemBytes(ringSlots, plan.nColors, consumers, stageEntries);
        for (TmaKernelFn fn : {cheb0 ? StepKernelTma<true, false> : StepKernelTma<false, false>,
                               cheb0 ? StepKernelTma<true, true> : StepKernelTma<false, true>})
        {
            VBDX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, maxOptin));
            int n = 0;
            VBDX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, blockThreads, smemBytes));
            perSm = std::min(perSm, n);
        }
        perSm = std::min(perSm, 1);  // the ring is sized for one CTA per SM
    }
    else
    {
        blockThreads = 256;
        smemBytes    = static_cast<size_t>(blockThreads / 32) * stageEntries * sizeof(float4);
        if (smemBytes > static_cast<size_t>(maxOptin))
            throw Error(VBDX_UNSUPPORTED, "1-ring staging does not fit in shared memory (vertex valence too high)");
        for (StepKernelFn fn : {cheb0 ? StepKernel<true, false> : StepKernel<false, false>,
                                cheb0 ? StepKernel<true, true> : StepKernel<false, true>})
        {
            VBDX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, maxOptin));
            int n = 0;
            VBDX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, blockThreads, smemBytes));
            perSm = std::min(perSm, n);
        }
    }
    if (perSm < 1)
        throw Error(VBDX_CUDA_ERROR, "step kernel does not fit on an SM");
    gridBlocks = clusterMode ? kClusterCtas : perSm * smCount;
    if (char const* e = std::getenv("VBDX_GRID_BLOCKS"); e && !clusterMode)  // tuning: fewer CTAs = cheaper grid barrier on small meshes
        gridBlocks = std::max(1, std::min(std::atoi(e), gridBlocks));
    PartitionTiles(plan, gridBlocks);
    if (char const* e = std::getenv("VBDX_GRID_BLOCKS"); e && flowGridBlocks > 0)
        flowGridBlocks = std::max(1, std::min(std::atoi(e), flowGridBlocks));
    nGhost = static_cast<int64_t>(nV) - plan.ghostBegin;  // (the schedule treats the tiles next to another GPU specially)
    if (variant == VBDX_KERNEL_PIPELINED && !clusterOnly && flowGridBlocks > 0)
        BuildFlowSchedule();

    dTiles.Alloc(plan.tiles.size() + 1, &deviceBytes);
    dTiles.Upload(plan.tiles.data(), plan.tiles.size(), stream);
    dCtaRange.Alloc(plan.ctaTileRange.size() + 1, &deviceBytes);
    dCtaRange.Upload(plan.ctaTileRange.data(), plan.ctaTileRange.size(), stream);
    dColorTileBegin.Alloc(plan.colorTileBegin.size() + 1, &deviceBytes);
    dColorTileBegin.Upload(plan.colorTileBegin.data(), plan.colorTileBegin.size(), stream);
    dCtaBlockBegin.Alloc(plan.ctaBlockBegin.size() + 1, &deviceBytes);
    dCtaBlockBegin.Upload(plan.ctaBlockBegin.data(), plan.ctaBlockBegin.size(), stream);
    dNew2Old.Alloc(nV, &deviceBytes);
    dNew2Old.Upload(plan.new2old.data(), nV, stream);
    dOld2New.Alloc(nV, &deviceBytes);
    dOld2New.Upload(plan.old2new.data(), nV, stream);
    dRingIds.Alloc(plan.ringIds.size() + 32, &deviceBytes);
    dRingIds.Upload(plan.ringIds.data(), plan.ringIds.size(), stream);
    dRecords.Alloc(static_cast<size_t>(plan.nBlocks) * kBlockFloat4 + 1, &deviceBytes);
    if (!plan.tiles.empty())
    {
        DevBuf<uint32_t> dRecIdx;
        dRecIdx.Alloc(plan.recIdx.size() + 1);
        dRecIdx.Upload(plan.recIdx.data(), plan.recIdx.size(), stream);
        int const nTiles = static_cast<int>(plan.tiles.size());
        FillRecords<<<Blocks(static_cast<int64_t>(nTiles) * 32, 256), 256, 0, stream>>>(
            dTiles.p, nTiles, dNew2Old.p, dOld2New.p, dPtr.p, dAdj.p, dE.p, dJinv.p, dVol.p, dLame.p,
            muDefault, lamDefault, dRecIdx.p, material == VBDX_MATERIAL_STVK ? 1 : 0, dRecords.p);
        ++kernelLaunches;
        VBDX_CUDA(cudaStreamSynchronize(stream));  // dRecIdx is released at the end of this scope
    }

    // ---- state
    bool const cheb = acceleration == VBDX_ACCEL_CHEBYSHEV;
    // Q (and P under Chebyshev), then the odd-tag copies of the ghosts (domain decomposition, step_kernel.cuh)
    nGhost = static_cast<int64_t>(nV) - plan.ghostBegin;
    dPos.Alloc(static_cast<size_t>(nV) * (cheb ? 2 : 1) + 2 * static_cast<size_t>(nGhost), &deviceBytes);
    if (cheb)
        dHist.Alloc(nV, &deviceBytes);
    dXtildeM.Alloc(nV, &deviceBytes);
    dXt.Alloc(nV, &deviceBytes);
    dVel.Alloc(nV, &deviceBytes);
    dAext.Alloc(nV, &deviceBytes);
    if (flags & VBDX_FLAG_ADAPTIVE_VBD_GPU_HISTORY)
        dVtm1.Alloc(nV, &deviceBytes);
    dBarrier.Alloc(3, &deviceBytes);  // [0] grid barrier counter, [1] spare, [2] non-finite sentinel
    VBDX_CUDA(cudaMemsetAsync(dBarrier.p, 0, 3 * sizeof(unsigned int), stream));
    dDistFlags.Alloc(16, &deviceBytes);
    VBDX_CUDA(cudaMemsetAsync(dDistFlags.p, 0, 16 * sizeof(unsigned int), stream));
    dStaging.Alloc(3 * nV, &deviceBytes);
    DevBuf<double> dV0, dA0;
    if (d.v)
    {
        dV0.Alloc(3 * nV);
        dV0.Upload(d.v, 3 * nV, stream);
    }
    if (d.aext)
    {
        dA0.Alloc(3 * nV);
        dA0.Upload(d.aext, 3 * nV, stream);
    }
    InitState<<<Blocks(nV, 256), 256, 0, stream>>>(
        nV, dNew2Old.p, dX.p, dV0.p, dA0.p, dMass.p, dIsDbc.p, cheb ? static_cast<uint32_t>(nV) : 0u, dPos.p,
        dHist.p, dXtildeM.p, dXt.p, dVel.p, dVtm1.p, dAext.p);
    ++kernelLaunches;
    VBDX_CUDA(cudaStreamSynchronize(stream));
    VBDX_CUDA(cudaGetLastError());

    // ---- vertex-triangle contact (Data::V, F, B; sim/vbd/Data.cpp:34-54 for the areas)
    muC = d.muC, muF = d.muF, epsv = d.epsv;
    if (d.nF > 0 && d.nCV > 0)
    {
        Require(d.F != nullptr && d.V != nullptr, "collision mesh pointers missing");
        Require(d.active_set_update_frequency >= 1, "active set update frequency must be >= 1");
        ContactState& cs = contact;
        cs.nCV = static_cast<uint32_t>(d.nCV), cs.nF = static_cast<uint32_t>(d.nF);
        cs.updateFrequency = d.active_set_update_frequency;
        std::vector<int32_t> Bh(nV), Vh(d.nCV);
        std::vector<int4> Fh(d.nF);
        std::vector<float> XVAh(nV, 0.f), FAh(d.nF);
        std::vector<double> xva(nV, 0.0);
        for (int64_t i = 0; i < nV; ++i)
            Bh[plan.old2new[i]] = d.B ? static_cast<int32_t>(d.B[i]) : 1;  // default body map: all ones (sim/vbd/Data.cpp:27-30)
        for (int64_t k = 0; k < d.nCV; ++k)
        {
            Require(d.V[k] >= 0 && d.V[k] < nV, "collision vertex index out of range");
            Vh[k] = plan.old2new[d.V[k]];
        }
        for (int64_t f = 0; f < d.nF; ++f)
        {
            int64_t const a = d.F[3 * f], b = d.F[3 * f + 1], c = d.F[3 * f + 2];
            Require(a >= 0 && a < nV && b >= 0 && b < nV && c >= 0 && c < nV, "collision triangle index out of range");
            Fh[f] = make_int4(plan.old2new[a], plan.old2new[b], plan.old2new[c], Bh[plan.old2new[a]]);  // .w: the triangle's body
            double ab[3], ac[3];
            for (int k = 0; k < 3; ++k)
            {
                ab[k] = d.X[3 * b + k] - d.X[3 * a + k];
                ac[k] = d.X[3 * c + k] - d.X[3 * a + k];
            }
            double const nx = ab[1] * ac[2] - ab[2] * ac[1], ny = ab[2] * ac[0] - ab[0] * ac[2], nz = ab[0] * ac[1] - ab[1] * ac[0];
            double const dbl = std::sqrt(nx * nx + ny * ny + nz * nz);
            xva[a] += dbl / 6, xva[b] += dbl / 6, xva[c] += dbl / 6;
            FAh[f] = static_cast<float>(dbl / 2);
        }
        for (int64_t i = 0; i < nV; ++i)
            XVAh[plan.old2new[i]] = static_cast<float>(xva[i]);
        // internal id range of every colour (swept vertices are numbered colour-major)
        std::vector<uint32_t> cvb(plan.nColors + 1, 0);
        for (int32_t c = 0; c < plan.nColors; ++c)
        {
            uint32_t const t0 = plan.colorTileBegin[c];
            cvb[c]            = t0 < plan.tiles.size() ? plan.tiles[t0].vbase : static_cast<uint32_t>(plan.nActive);
        }
        cvb[plan.nColors] = static_cast<uint32_t>(plan.nActive);
        for (int32_t c = plan.nColors - 1; c >= 0; --c)  // empty colours inherit the next begin
            if (plan.colorTileBegin[c] == plan.colorTileBegin[c + 1])
                cvb[c] = cvb[c + 1];
        cs.B.Alloc(nV, &deviceBytes), cs.V.Alloc(d.nCV, &deviceBytes), cs.F.Alloc(d.nF, &deviceBytes);
        cs.XVA.Alloc(nV, &deviceBytes), cs.FA.Alloc(d.nF, &deviceBytes), cs.colorVertexBegin.Alloc(cvb.size(), &deviceBytes);
        cs.B.Upload(Bh.data(), nV, stream), cs.V.Upload(Vh.data(), d.nCV, stream), cs.F.Upload(Fh.data(), d.nF, stream);
        cs.XVA.Upload(XVAh.data(), nV, stream), cs.FA.Upload(FAh.data(), d.nF, stream);
        cs.colorVertexBegin.Upload(cvb.data(), cvb.size(), stream);
        cs.mesh = ContactMesh{cs.B.p, cs.V.p, cs.F.p, cs.nCV, cs.nF};
        cs.Alloc(nV, &deviceBytes, stream);
        // snapshot buffer 0 = initial positions
        VBDX_CUDA(cudaMemcpyAsync(cs.snap.p, dPos.p, nV * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        if (dFlowTiles.p != nullptr && distWorld == 1)
        {
            size_t const n = 4 * (acceleration == VBDX_ACCEL_CHEBYSHEV ? 2 : 1) * static_cast<size_t>(nV);
            dHist4.Alloc(n, &deviceBytes);
            VBDX_CUDA(cudaMemsetAsync(dHist4.p, 0, n * sizeof(float4), stream));  // write number 0: never asked for
        }
        VBDX_CUDA(cudaStreamSynchronize(stream));
        cs.enabled = true;
    }
}

// What the lean barrier-free kernel walks (step_kernel_flow.cuh): per warp of the persistent grid the tiles it runs in one
// sweep, in order (a colour's tiles are dealt round-robin over the warps, heaviest first, like the other kernels do on the
// fly), and every tile's ring entries pre-decoded: final index into the position buffers plus previous-iterate / never-changes
// / ghost bits, the four chunks of a lane packed into one 16-byte word.
void Integrator::BuildFlowSchedule()
{
    uint32_t const gWarps = static_cast<uint32_t>(flowGridBlocks) * static_cast<uint32_t>(flowWarps);
    size_t const nTiles   = plan.tiles.size();
    bool const cheb       = acceleration == VBDX_ACCEL_CHEBYSHEV;
    uint32_t const pOff   = cheb ? static_cast<uint32_t>(nV) : 0u;
    Require(2 * static_cast<int64_t>(nV) < (int64_t(1) << 29), "mesh too large for the packed ring entries of the barrier-free kernel");
    std::vector<uint32_t> idsStart(nTiles + 1, 0);
    for (size_t t = 0; t < nTiles; ++t)
        idsStart[t + 1] = idsStart[t] + ((TileChunks(plan.tiles[t].meta) + 3u) / 4u) * 128u;
    std::vector<uint32_t> ids(idsStart[nTiles], kFlowStatic);
    uint32_t const activeEnd = static_cast<uint32_t>(plan.nActive), ghostBegin = static_cast<uint32_t>(plan.ghostBegin);
    for (size_t t = 0; t < nTiles; ++t)
    {
        TileDesc const& td    = plan.tiles[t];
        uint32_t const chunks = TileChunks(td.meta);
        for (uint32_t j = 0; j < ((chunks + 3u) / 4u) * 4u; ++j)
            for (uint32_t lane = 0; lane < 32; ++lane)
            {
                uint32_t e = kFlowStatic | td.vbase;  // chunks beyond the tile's: never loaded
                if (j < chunks)
                {
                    uint32_t const raw  = plan.ringIds[td.ringStart + 32u * j + lane];
                    uint32_t const base = raw & ~kPrevFlag;
                    bool const prev     = (raw & kPrevFlag) != 0u;
                    if (raw == td.vbase)  // padding: the tile's first vertex without the flag
                        e = kFlowStatic | base;
                    else if (base >= ghostBegin)
                        e = kFlowGhost | (prev ? kFlowPrev : 0u) | (base - ghostBegin);
                    else
                        e = (prev ? kFlowPrev : 0u) | (base >= activeEnd ? kFlowStatic : 0u) | (base + (prev ? pOff : 0u));
                }
                ids[idsStart[t] + (j >> 2) * 128u + lane * 4u + (j & 3u)] = e;
            }
    }
    // Deal order of a colour's tiles: the planner's (heaviest first).  Optionally (domain decomposition) the tiles next to
    // another GPU's vertices first; with about one tile per warp and colour the order within a colour decides little, and
    // it measured slower.
    std::vector<uint32_t> order(nTiles);
    bool boundaryFirst = false;  // (measured on 2 GPUs: 1.045 ms/step with it, 1.021 without -- off unless VBDX_BOUNDARY_FIRST=1)
    if (char const* e = std::getenv("VBDX_BOUNDARY_FIRST"))
        boundaryFirst = nGhost > 0 && std::atoi(e) != 0;
    for (int32_t c = 0; c < plan.nColors; ++c)
    {
        uint32_t const tb = plan.colorTileBegin[c], te = plan.colorTileBegin[c + 1];
        uint32_t k = tb;
        for (int pass = 0; pass < 2; ++pass)
            for (uint32_t t = tb; t < te; ++t)
                if ((boundaryFirst && TileReadsGhosts(plan.tiles[t].meta)) == (pass == 0))
                    order[k++] = t;
        if (!boundaryFirst)
            for (uint32_t t = tb; t < te; ++t)
                order[t] = t;
    }
    // Which warp runs which tile (home = warp-in-CTA * CTAs + CTA): a colour's tiles round-robin over all warps of the grid.
    // Experiment, opt-in (VBDX_BOUNDARY_WARPS=n, domain decomposition, at most one tile per warp and colour): the tiles next
    // to another GPU go to a few CTAs of their own with only n of their warps in use.  A tile's colour-to-colour chain
    // there includes the NVLink hop of the ghosts it waits for (~0.5 us more than a local dependency), and a tile runs
    // faster on an SM that runs fewer of them.  Only the placement changes: same tiles, same arithmetic, same bits.
    // Measured on 2 GPUs: 1.021 ms/step without, 1.003-1.009 with n = 11-12, 1.04 with n = 9 or 16 -- the grid has only 8 %
    // more warps than a colour has tiles, so the boundary SMs cannot be made light enough; off by default.
    uint32_t const G = static_cast<uint32_t>(flowGridBlocks), W = static_cast<uint32_t>(flowWarps);
    uint32_t boundarySms = 0, boundaryWarps = 0;
    {
        char const* e      = std::getenv("VBDX_BOUNDARY_WARPS");  // unset / 0: off; n: warps in use per boundary SM; -1: the fewest that fit, from 4
        int const wanted   = e ? std::atoi(e) : 0;
        uint32_t nbMax = 0, niMax = 0;
        for (int32_t c = 0; c < plan.nColors; ++c)
        {
            uint32_t nb = 0;
            for (uint32_t t = plan.colorTileBegin[c]; t < plan.colorTileBegin[c + 1]; ++t)
                nb += TileReadsGhosts(plan.tiles[t].meta) ? 1u : 0u;
            nbMax = std::max(nbMax, nb);
            niMax = std::max(niMax, plan.colorTileBegin[c + 1] - plan.colorTileBegin[c] - nb);
        }
        if (nGhost > 0 && nbMax > 0 && wanted != 0)
            for (uint32_t wb = wanted > 0 ? static_cast<uint32_t>(wanted) : 4u; wb <= (wanted > 0 ? static_cast<uint32_t>(wanted) : W) && wb <= W; ++wb)
            {
                uint32_t const sb = (nbMax + wb - 1) / wb;
                if (sb < G && static_cast<uint64_t>(niMax) <= static_cast<uint64_t>(G - sb) * W)
                {
                    boundarySms = sb, boundaryWarps = wb;
                    break;
                }
            }
    }
    std::vector<uint32_t> home(nTiles);
    for (int32_t c = 0; c < plan.nColors; ++c)
    {
        uint32_t ib = 0, ii = 0;
        for (uint32_t t = plan.colorTileBegin[c]; t < plan.colorTileBegin[c + 1]; ++t)
        {
            if (boundarySms == 0)
                home[t] = (t - plan.colorTileBegin[c]) % gWarps;
            else if (TileReadsGhosts(plan.tiles[order[t]].meta))
                home[t] = (ib / boundarySms) * G + ib % boundarySms, ++ib;
            else
                home[t] = (ii / (G - boundarySms)) * G + boundarySms + ii % (G - boundarySms), ++ii;
        }
    }
    flowBoundarySms = boundarySms, flowBoundaryWarps = boundaryWarps;
    std::vector<uint32_t> wBegin(static_cast<size_t>(gWarps) + 1, 0);
    for (size_t t = 0; t < nTiles; ++t)
        ++wBegin[home[t] + 1];
    for (uint32_t w = 0; w < gWarps; ++w)
        wBegin[w + 1] += wBegin[w];
    std::vector<uint4> seqTiles(nTiles);
    std::vector<uint32_t> cursor(wBegin.begin(), wBegin.end() - 1);
    for (int32_t c = 0; c < plan.nColors; ++c)
        for (uint32_t t = plan.colorTileBegin[c]; t < plan.colorTileBegin[c + 1]; ++t)
        {
            TileDesc const& td = plan.tiles[order[t]];
            seqTiles[cursor[home[t]]++] = make_uint4(td.blockStart, td.vbase, td.meta, idsStart[order[t]]);
        }
    flowActiveWarps = 0;
    for (uint32_t w = 0; w < gWarps; ++w)
        flowActiveWarps += wBegin[w + 1] > wBegin[w] ? 1u : 0u;
    dFlowWarpBegin.Alloc(wBegin.size(), &deviceBytes);
    dFlowWarpBegin.Upload(wBegin.data(), wBegin.size(), stream);
    dFlowTiles.Alloc(nTiles + 1, &deviceBytes);
    if (nTiles)
        dFlowTiles.Upload(seqTiles.data(), nTiles, stream);
    dFlowIds.Alloc(ids.size() + 128, &deviceBytes);
    if (!ids.empty())
        dFlowIds.Upload(ids.data(), ids.size(), stream);
    VBDX_CUDA(cudaStreamSynchronize(stream));  // the host vectors go out of scope
}

void Integrator::Step(double dt, int iterations, int substeps, bool sync)
{
    Require(dt > 0 && iterations >= 0 && substeps >= 1, "Step: need dt > 0, iterations >= 0, substeps >= 1");
    VBDX_CUDA(cudaSetDevice(device));
    PrepareOmega(iterations);
    StepParams p = MakeParams(dt / static_cast<double>(substeps), iterations, substeps);
    RunStep(p, dt, iterations, substeps, sync);
}

// ChebyshevOmega for every iteration of a solve (sim/vbd/Kernels.h:96-102), uploaded once per iteration count
void Integrator::PrepareOmega(int iterations)
{
    bool const cheb = acceleration == VBDX_ACCEL_CHEBYSHEV;
    if (cheb && omegaIterations != iterations)
    {
        // ChebyshevOmega, evaluated in double like the reference's CPU integrator and rounded once
        omegaHost.assign(std::max(iterations, 1), 1.f);
        double const rho2 = rho * rho;
        double omega      = 0;
        for (int k = 0; k < iterations; ++k)
        {
            if (k == 0)
                omega = 1.0;
            else if (k == 1)
                omega = 2.0 / (2.0 - rho2);
            else
                omega = (omegaMode == VBDX_OMEGA_REFERENCE) ? (4.0 / 4.0 - rho2 * omega) : 4.0 / (4.0 - rho2 * omega);
            omegaHost[k] = static_cast<float>(omega);
        }
        if (dOmega.n < omegaHost.size())
            dOmega.Alloc(omegaHost.size(), &deviceBytes);
        dOmega.Upload(omegaHost.data(), omegaHost.size(), stream);
        VBDX_CUDA(cudaStreamSynchronize(stream));  // omegaHost must outlive the copy
        omegaIterations = iterations;
    }
}

// Everything a step kernel launch needs, for `substeps` substeps of length sdt with `iterations` sweeps each.
StepParams Integrator::MakeParams(double sdt, int iterations, int substeps)
{
    bool const cheb = acceleration == VBDX_ACCEL_CHEBYSHEV;
    StepParams p{};
    p.lineSearch   = lineSearch;
    p.records      = dRecords.p;
    p.tiles        = reinterpret_cast<uint4 const*>(dTiles.p);
    p.ctaTileRange = dCtaRange.p;
    p.colorTileBegin = dColorTileBegin.p;
    p.ringIds      = dRingIds.p;
    p.stageEntries = stageEntries;
    p.nColors      = plan.nColors;
    p.nVerts       = static_cast<int>(nV);
    p.pos          = dPos.p;
    p.pOff         = cheb ? static_cast<uint32_t>(nV) : 0u;
    p.hist         = dHist.p;
    p.xtildeM      = dXtildeM.p;
    p.xt           = dXt.p;
    p.vel          = dVel.p;
    p.vtm1         = dVtm1.p;
    p.aext         = dAext.p;
    p.omega        = dOmega.p;
    p.sdt          = static_cast<float>(sdt);
    p.sdt2         = static_cast<float>(sdt * sdt);
    p.dampD        = static_cast<float>(kD / sdt);
    p.detHZero     = static_cast<float>(detHZero);
    p.strategy     = strategy;
    p.iterations   = iterations;
    p.substeps     = substeps;
    p.barrier      = dBarrier.p;
    p.nonFinite    = dBarrier.p + 2;
    p.ghostBegin   = static_cast<uint32_t>(plan.ghostBegin);
    p.activeEnd    = static_cast<uint32_t>(plan.nActive);
    p.rank = distRank, p.world = distWorld;
    if (distWorld == 1)
    {
        // write numbers carried in .w of every position (barrier-free sweeps); a dependency that never arrives is a bug:
        // give up after VBDX_DATAFLOW_TIMEOUT_S instead of hanging the GPU
        p.tagBase   = dfTag;
        dfTag += static_cast<unsigned int>(substeps) * (static_cast<unsigned int>(iterations) + 1u);
        p.distError = dDistFlags.p + 9;
        const char* t   = std::getenv("VBDX_DATAFLOW_TIMEOUT_S");
        double const ts = t ? std::atof(t) : 2.0;  // a dependency is microseconds away
        p.distTimeoutNs = static_cast<unsigned long long>((ts > 0 ? ts : 2.0) * 1e9);
    }
    if (distWorld > 1)
    {
        p.sendPtr = dSendPtr.p, p.sendDst = dSendDst.p;
        for (int r = 0; r < 8; ++r)
            p.peerPos[r] = peerPos[r], p.peerPOff[r] = peerPOff[r], p.peerFlags[r] = peerFlags[r];
        p.myFlags   = dDistFlags.p;
        p.distError = dDistFlags.p + 9;
        p.distStats = dDistFlags.p + 10;
        p.epochBase = distEpoch;
        p.peerMask  = peerMask;
        p.tagBase   = distTag;
        {
            // processes reach their first step at different times; VBDX_DIST_TIMEOUT_S bounds the wait for a dead peer
            const char* t   = std::getenv("VBDX_DIST_TIMEOUT_S");
            double const ts = t ? std::atof(t) : 30.0;
            p.distTimeoutNs = static_cast<unsigned long long>((ts > 0 ? ts : 30.0) * 1e9);
        }
        p.ghostExt  = static_cast<uint32_t>(nV) * (cheb ? 2u : 1u);
        p.nGhost    = static_cast<uint32_t>(nGhost);
        for (int r = 0; r < 8; ++r)
            p.peerGhostExt[r] = peerGhostExt[r], p.peerGhostBegin[r] = peerGhostBegin[r], p.peerNGhost[r] = peerNGhost[r];
        distTag += static_cast<unsigned int>(substeps) * (static_cast<unsigned int>(iterations) + 1u);
        // barriers per substep: after the pre-step and after every colour -- or, with the barrier-free sweep, after the
        // pre-step and after the last sweep only (every rank takes the same decision: same settings, same environment)
        distEpoch += static_cast<unsigned int>(substeps) *
                     (WillSweepBarrierFree(iterations) ? 2u : 1u + static_cast<unsigned int>(iterations) * static_cast<unsigned int>(plan.nColors));
    }
    p.trace        = traceIteration >= 0 ? dTrace.p : nullptr;
    p.traceIteration = traceIteration;
    if (contact.enabled)
    {
        p.fc               = contact.fc.p;
        p.triF             = contact.F.p;
        p.XVA              = contact.XVA.p;
        p.FA               = contact.FA.p;
        p.snap             = contact.snap.p;
        p.colorVertexBegin = contact.colorVertexBegin.p;
        p.muC = static_cast<float>(muC), p.muF = static_cast<float>(muF), p.epsv = static_cast<float>(epsv);
    }
    return p;
}

void Integrator::LaunchStepKernel(StepParams const& q)
{
    {
        VBDX_CUDA(cudaMemsetAsync(dBarrier.p, 0, 2 * sizeof(unsigned int), stream));
        bool const wholeStepBarrierFree = variant == VBDX_KERNEL_PIPELINED && UseFlow(q.iterations) && !q.skipPostStep && q.iterBegin == 0 &&
                                          (contact.enabled ? q.hist4 != nullptr : !q.skipPreStep);
        if (variant == VBDX_KERNEL_PIPELINED && (!clusterMode || wholeStepBarrierFree))
        {
            PipeParams pp{};
            pp.base      = q;
            pp.maxIters  = maxTileIters;
            // barrier-free sweeps: whole steps of the base / Chebyshev solve; partial launches (traces, windowed accelerators)
            // keep the colour barriers.  With contact the step is one launch per substep behind PreStepKernel and the
            // active-set update: barrier-free where RunStep set up the write history (q.hist4).
            pp.dataflow = WillSweepBarrierFree(q.iterations) && !q.skipPostStep && q.iterBegin == 0 &&
                          (contact.enabled ? q.hist4 != nullptr : !q.skipPreStep);
            pp.flowWarpBegin = dFlowWarpBegin.p, pp.flowTiles = dFlowTiles.p, pp.flowIds = dFlowIds.p;
            pp.flowActiveWarps = flowActiveWarps;
            usedDataflow |= pp.dataflow != 0;
            void* args[] = {&pp};
            bool const flow = pp.dataflow != 0 && UseFlow(q.iterations);
            if (flow)
                VBDX_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void const*>(KernelFlow()), dim3(flowGridBlocks), dim3(flowWarps * 32), args,
                                                      flowSmemBytes, stream));
            else
                VBDX_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void const*>(KernelPipe(pp.dataflow != 0)), dim3(gridBlocks),
                                                      dim3(blockThreads), args, smemBytes, stream));
        }
        else if (variant == VBDX_KERNEL_PIPELINED)
        {
            // one thread-block cluster = the whole grid; no cooperative launch needed (a cluster is co-scheduled)
            PipeParams pp{};
            pp.base           = q;
            pp.maxIters       = maxTileIters;
            pp.clusterBarrier = 1;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim          = dim3(gridBlocks);
            cfg.blockDim         = dim3(blockThreads);
            cfg.dynamicSmemBytes = smemBytes;
            cfg.stream           = stream;
            cudaLaunchAttribute attr{};
            attr.id               = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = static_cast<unsigned>(gridBlocks), attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
            cfg.attrs    = &attr;
            cfg.numAttrs = 1;
            void* args[] = {&pp};
            VBDX_CUDA(cudaLaunchKernelExC(&cfg, reinterpret_cast<void const*>(KernelPipe()), args));
        }
        else if (variant == VBDX_KERNEL_TMA)
        {
            TmaParams tp{};
            tp.base          = q;
            tp.ctaBlockBegin = dCtaBlockBegin.p;
            tp.ringSlots     = ringSlots;
            void* args[]     = {&tp};
            VBDX_CUDA(cudaLaunchCooperativeKernel(
                reinterpret_cast<void const*>(KernelTma()), dim3(gridBlocks), dim3(blockThreads), args, smemBytes, stream));
        }
        else
        {
            StepParams qq = q;
            void* args[]  = {&qq};
            VBDX_CUDA(cudaLaunchCooperativeKernel(
                reinterpret_cast<void const*>(Kernel()), dim3(gridBlocks), dim3(blockThreads), args, smemBytes, stream));
        }
        ++kernelLaunches;
    }
}

void Integrator::RunStep(StepParams const& p, double dt, int iterations, int substeps, bool sync)
{
    NvtxRange const zone("pbat.gpu.impl.vbd.Integrator.Step");
    bool const cheb = acceleration == VBDX_ACCEL_CHEBYSHEV;
    auto launchStep = [&](StepParams const& q) {
        NvtxRange const z("pbat.gpu.impl.vbd.Integrator.Solve");
        LaunchStepKernel(q);
    };
    VBDX_CUDA(cudaEventRecord(evBegin, stream));
    VBDX_CUDA(cudaMemsetAsync(dBarrier.p + 2, 0, sizeof(unsigned int), stream));  // the step's non-finite sentinel
    if (acceleration == VBDX_ACCEL_ANDERSON)
        AndersonStep(p, dt, iterations, substeps);
    else if (acceleration == VBDX_ACCEL_BROYDEN)
        BroydenStep(p, dt, iterations, substeps);
    else if (acceleration == VBDX_ACCEL_NESTEROV)
        NesterovStep(p, dt, iterations, substeps);
    else if (acceleration == VBDX_ACCEL_TRUST_REGION && !trCurved)
        TrustRegionStep(p, dt, iterations, substeps);
    // (the trust region's curved path: the reference's SolveCurvedTrustRegionConstraint is a stub that returns 0
    // (TrustRegionIntegrator.cu:700-703), which empties the admissible step interval: no accelerated step is ever tried,
    // every iteration falls back to the plain sweep -- the iterates ARE the base solve's, so that is what runs)
    else if (!contact.enabled)
        launchStep(p);
    else
    {
        // gpu/impl/vbd/Integrator.cu:82-103: active set from the full-step predictor, then per substep
        // inertial targets + initial guess, (every `frequency` substeps) nearest triangles from the initial
        // guess, the solve, the velocity update; finally prune vertices that left the surface
        float4 const* xFinal = dPos.p + p.pOff;
        {
            NvtxRange const z("pbat.gpu.impl.vbd.Integrator.InitializeActiveSet");
            contact.InitializeActiveSet(xFinal, dVel.p, dAext.p, nV, static_cast<float>(dt), stream, &kernelLaunches);
        }
        StepParams q  = p;
        q.substeps    = 1;
        q.skipPreStep = 1;
        if (UseFlow(iterations))
        {
            q.hist4 = dHist4.p, q.snap = nullptr;  // barrier-free sweeps: contact reads go to the write history
            if (dSweepDone.n < static_cast<size_t>(iterations))
                dSweepDone.Alloc(static_cast<size_t>(iterations) + 64, &deviceBytes);
            q.sweepDone = dSweepDone.p;
        }
        for (int s = 0; s < substeps; ++s)
        {
            q.tagBase = p.tagBase + static_cast<unsigned int>(s) * (static_cast<unsigned int>(iterations) + 1u);  // write numbers of this substep
            {
                NvtxRange const z("pbat.gpu.impl.vbd.Integrator.ComputeInertialTargets");  // + InitializeBcdSolution: one fused kernel
                if (cheb)
                    PreStepKernel<true><<<Blocks(nV, 256), 256, 0, stream>>>(q);
                else
                    PreStepKernel<false><<<Blocks(nV, 256), 256, 0, stream>>>(q);
                ++kernelLaunches;
            }
            if (s % contact.updateFrequency == 0)
            {
                NvtxRange const z("pbat.gpu.impl.vbd.Integrator.UpdateActiveSet");
                contact.NearestPass(xFinal, 0, stream, &kernelLaunches);
            }
            if (q.sweepDone != nullptr)
                VBDX_CUDA(cudaMemsetAsync(q.sweepDone, 0, static_cast<size_t>(iterations) * sizeof(unsigned int), stream));
            launchStep(q);
        }
        contact.NearestPass(xFinal, 1, stream, &kernelLaunches);
    }
    VBDX_CUDA(cudaEventRecord(evEnd, stream));
    stepTimed = true;
    if (sync)
    {
        VBDX_CUDA(cudaStreamSynchronize(stream));
        CheckAsyncErrors();
        float ms = 0;
        VBDX_CUDA(cudaEventElapsedTime(&ms, evBegin, evEnd));
        lastStepMs = ms;
    }
}

// Called with the stream idle (vbdx_step, vbdx_synchronize): did a dependency or halo wait of the last step(s) time out?
// Such a step finished with stale inputs: its result is discarded (positions go back to the start of the step's last
// substep; velocities are undefined until the caller sets them), the error word is cleared so that later launches wait
// again, and the failure is reported.  A barrier-free sweep that failed makes the handle sweep with barriers from then on.
void Integrator::CheckAsyncErrors()
{
    if (!(distWorld > 1 || usedDataflow))
        return;
    unsigned int dbg[6] = {0, 0, 0, 0, 0, 0};
    VBDX_CUDA(cudaMemcpy(dbg, dDistFlags.p + 9, sizeof(dbg), cudaMemcpyDeviceToHost));
    unsigned int const e = dbg[0];
    if (e == 0u)
        return;
    VBDX_CUDA(cudaMemset(dDistFlags.p + 9, 0, sizeof(dbg)));
    size_t const owned = static_cast<size_t>(plan.ghostBegin);
    VBDX_CUDA(cudaMemcpy(dPos.p, dXt.p, owned * sizeof(float4), cudaMemcpyDeviceToDevice));
    if (acceleration == VBDX_ACCEL_CHEBYSHEV)
        VBDX_CUDA(cudaMemcpy(dPos.p + nV, dXt.p, owned * sizeof(float4), cudaMemcpyDeviceToDevice));
    if (e == 2u)
    {
        dataflow = false;  // this handle sweeps with barriers from now on
        throw Error(VBDX_CUDA_ERROR, "internal error: a barrier-free sweep waited for a vertex update that never came (entry " +
                                         std::to_string(dbg[1] & 0x1fffffffu) + ((dbg[1] >> 31) ? " [previous-iterate buffer]" : "") + ", expected write " +
                                         std::to_string(dbg[2]) + ", found " + std::to_string(dbg[3]) + ", tile of vertex " + std::to_string(dbg[4]) +
                                         ", sweep " + std::to_string(dbg[5]) + ", write base " + std::to_string(dfTag) +
                                         "); the step was discarded (positions restored, set the velocities before continuing); VBDX_DATAFLOW=0 selects the barrier sweep");
    }
    throw Error(VBDX_CUDA_ERROR, "domain decomposition: a peer GPU did not deliver its halo or reach the barrier in time (VBDX_DIST_TIMEOUT_S, default 30 s); "
                                 "the step was discarded on this rank (positions restored, set the velocities before continuing)");
}

void Integrator::LaunchPreStep(StepParams const& q)
{
    if (acceleration == VBDX_ACCEL_CHEBYSHEV)
        PreStepKernel<true><<<Blocks(nV, 256), 256, 0, stream>>>(q);
    else
        PreStepKernel<false><<<Blocks(nV, 256), 256, 0, stream>>>(q);
    ++kernelLaunches;
}

// AndersonIntegrator::Solve inside Integrator::Step (sim/vbd/AndersonIntegrator.cpp:24-58): the sweeps are one-iteration
// launches of the persistent step kernel, the window lives in anderson.cuh's kernels; nothing returns to the host.
void Integrator::AndersonStep(StepParams const& p, double dt, int iterations, int substeps)
{
    NvtxRange const zone("pbat.gpu.impl.vbd.AndersonIntegrator.Solve");
    int const m = window;
    if (dAndVec.n == 0)
    {
        dAndVec.Alloc(static_cast<size_t>(nV) * (4 + 2 * m), &deviceBytes);
        dAndSmall.Alloc(static_cast<size_t>(m) * m + 3 * m, &deviceBytes);
    }
    AndersonView a{};
    a.n = nV, a.m = m, a.pos = dPos.p;
    a.xkm1 = dAndVec.p, a.Gkm1 = a.xkm1 + nV, a.Fkm1 = a.Gkm1 + nV, a.Fk = a.Fkm1 + nV, a.DF = a.Fk + nV;
    a.DG      = a.DF + static_cast<size_t>(m) * nV;
    a.gram    = dAndSmall.p;
    a.scratch = a.gram + m * m;
    a.alpha   = a.scratch + 2 * m;
    StepParams q   = p;
    q.substeps     = 1;
    q.skipPreStep  = 1;
    q.skipPostStep = 1;
    q.iterations   = 1;
    int const grid = Blocks(nV, 256);
    ContactBeginStep(p, dt);
    for (int s = 0; s < substeps; ++s)
    {
        LaunchPreStep(q);
        ContactAfterPreStep(p, s);
        {
            // the reference sweeps once before its loop whatever `iterations` is (AndersonIntegrator.cpp:30-33)
            VBDX_CUDA(cudaMemsetAsync(dAndSmall.p, 0, dAndSmall.n * sizeof(double), stream));
            VBDX_CUDA(cudaMemcpyAsync(a.xkm1, dPos.p, nV * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
            q.iterBegin = 0;
            LaunchStepKernel(q);
            AndersonFirst<<<grid, 256, 0, stream>>>(a);
            ++kernelLaunches;
            // the first sweep's result is the second sweep's start: x^{k-1} = x
            if (iterations > 1)
                VBDX_CUDA(cudaMemcpyAsync(a.xkm1, dPos.p, nV * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        }
        for (int k = 1; k < iterations; ++k)
        {
            q.iterBegin = k;
            LaunchStepKernel(q);
            int const dkl = (k - 1) % m, mk = std::min(m, k);
            AndersonWindow<kMaxAndersonWindow><<<std::min(grid, 4 * 148), 256, 0, stream>>>(a, dkl, mk);
            AndersonSolveSmall<<<1, 32, 0, stream>>>(a, dkl, mk, 1e-10);
            AndersonApply<<<grid, 256, 0, stream>>>(a, mk, plan.nActive, SnapNext(k));
            kernelLaunches += 3;
        }
        StepParams post = q;
        post.iterations = 0, post.skipPostStep = 0;
        LaunchStepKernel(post);  // velocity update only
    }
    ContactEndStep(p);
}

// BroydenIntegrator::Solve inside Integrator::Step (sim/vbd/BroydenIntegrator.cpp:41-77), same launch structure as
// AndersonStep; the window buffers are shared with it (dAndVec: xkm1, fkm1, fk, unused, X[m], GF[m]).
void Integrator::BroydenStep(StepParams const& p, double dt, int iterations, int substeps)
{
    int const m = window;
    if (dAndVec.n == 0)
    {
        dAndVec.Alloc(static_cast<size_t>(nV) * (4 + 2 * m), &deviceBytes);
        dAndSmall.Alloc(static_cast<size_t>(m) * m + 3 * m, &deviceBytes);
    }
    BroydenView a{};
    a.n = nV, a.m = m, a.pos = dPos.p;
    a.xkm1 = dAndVec.p, a.fkm1 = a.xkm1 + nV, a.fk = a.fkm1 + nV, a.X = a.fk + 2 * static_cast<size_t>(nV);
    a.GF      = a.X + static_cast<size_t>(m) * nV;
    a.gram    = dAndSmall.p;
    a.scratch = a.gram + m * m;
    a.gamma   = a.scratch + 2 * m;
    StepParams q   = p;
    q.substeps     = 1;
    q.skipPreStep  = 1;
    q.skipPostStep = 1;
    q.iterations   = 1;
    int const grid = Blocks(nV, 256);
    ContactBeginStep(p, dt);
    for (int s = 0; s < substeps; ++s)
    {
        LaunchPreStep(q);
        ContactAfterPreStep(p, s);
        VBDX_CUDA(cudaMemsetAsync(dAndSmall.p, 0, dAndSmall.n * sizeof(double), stream));
        VBDX_CUDA(cudaMemcpyAsync(a.xkm1, dPos.p, nV * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        // the reference sweeps once before its loop whatever `iterations` is (BroydenIntegrator.cpp:47-49)
        q.iterBegin = 0;
        LaunchStepKernel(q);
        BroydenFirst<<<grid, 256, 0, stream>>>(a);
        ++kernelLaunches;
        for (int k = 1; k < iterations; ++k)
        {
            int const col = (k - 1) % m, mk = std::min(m, k);
            BroydenBeforeSweep<<<grid, 256, 0, stream>>>(a, col);
            q.iterBegin = k;
            LaunchStepKernel(q);
            BroydenWindow<kMaxAndersonWindow><<<std::min(grid, 4 * 148), 256, 0, stream>>>(a, col, mk);
            BroydenSolveSmall<<<1, 32, 0, stream>>>(a, col, mk, std::max(1, m - k), 1e-10);
            BroydenApply<<<grid, 256, 0, stream>>>(a, mk, plan.nActive, SnapNext(k));
            kernelLaunches += 4;
        }
        StepParams post = q;
        post.iterations = 0, post.skipPostStep = 0;
        LaunchStepKernel(post);  // velocity update only
    }
    ContactEndStep(p);
}


// NesterovIntegrator::Solve inside Integrator::Step (sim/vbd/NesterovIntegrator.cpp:18-44), literally: x^{k-1} is captured
// once before the loop, the sweep starts from x, and from iteration start + 1 on x <- y^k - (x_swept - x^{k-1}) / L.
void Integrator::NesterovStep(StepParams const& p, double dt, int iterations, int substeps)
{
    if (dAndVec.n == 0)
        dAndVec.Alloc(static_cast<size_t>(nV) * 2, &deviceBytes);
    float4* const xkm1 = dAndVec.p;
    float4* const yk   = xkm1 + nV;
    StepParams q   = p;
    q.substeps     = 1;
    q.skipPreStep  = 1;
    q.skipPostStep = 1;
    q.iterations   = 1;
    int const grid = Blocks(nV, 256);
    ContactBeginStep(p, dt);
    for (int s = 0; s < substeps; ++s)
    {
        LaunchPreStep(q);
        ContactAfterPreStep(p, s);
        VBDX_CUDA(cudaMemcpyAsync(xkm1, dPos.p, nV * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        double const alpha = 1.0 / nesterovL;
        double lambda = 0.0, beta = 0.0;
        for (int k = 0; k < iterations; ++k)
        {
            bool const on = nesterovStart < k;
            if (on)
            {
                NesterovExtrapolate<<<grid, 256, 0, stream>>>(nV, dPos.p, xkm1, yk, static_cast<float>(beta));
                ++kernelLaunches;
            }
            q.iterBegin = k;
            LaunchStepKernel(q);
            if (on)
            {
                NesterovCorrect<<<grid, 256, 0, stream>>>(nV, dPos.p, xkm1, yk, static_cast<float>(alpha), plan.nActive, SnapNext(k));
                ++kernelLaunches;
                double const lambdak = lambda;
                lambda               = (1.0 + std::sqrt(1.0 + 4.0 * lambda * lambda)) / 2.0;
                beta                 = (lambdak - 1.0) / lambda;
            }
        }
        StepParams post = q;
        post.iterations = 0, post.skipPostStep = 0;
        LaunchStepKernel(post);  // velocity update only
    }
    ContactEndStep(p);
}

// K/2 + dt^2 U of the CURRENT device state (TrustRegionIntegrator::ObjectiveFunction, TrustRegionIntegrator.cu:265-364,
// without the contact term): positions and inertial targets go to the double-precision diagnostics kernels.
double Integrator::ObjectiveOfState(double sdt)
{
    if (dObjX.n == 0)
    {
        dObjX.Alloc(3 * nV, &deviceBytes), dObjXt.Alloc(3 * nV, &deviceBytes), dObjGrad.Alloc(3 * nV, &deviceBytes);
        dObjVal.Alloc(1, &deviceBytes);
    }
    GatherToCaller<double><<<Blocks(nV, 256), 256, 0, stream>>>(nV, dOld2New.p, dPos.p, dObjX.p, 1, 3);
    GatherToCaller<double><<<Blocks(nV, 256), 256, 0, stream>>>(nV, dOld2New.p, dXtildeM.p, dObjXt.p, 1, 3);
    VBDX_CUDA(cudaMemsetAsync(dObjVal.p, 0, sizeof(double), stream));
    ObjectiveKinetic<<<std::min(Blocks(nV, 256), 1184), 256, 0, stream>>>(dObjX.p, dObjXt.p, dMass.p, nV, dObjVal.p, nullptr);
    ObjectiveElastic<<<std::min(Blocks(nT, 256), 1184), 256, 0, stream>>>(dObjX.p, dE.p, dJinv.p, dVol.p, dLame.p, mu0, lambda0, nT, sdt * sdt,
                                                                            material == VBDX_MATERIAL_STVK ? 1 : 0, dObjVal.p, nullptr);
    kernelLaunches += 4;
    double f = 0;
    dObjVal.Download(&f, 1, stream);
    VBDX_CUDA(cudaStreamSynchronize(stream));
    return f;
}

// TrustRegionIntegrator::SolveWithLinearAcceleratedPath inside Integrator::Step (gpu/impl/vbd/TrustRegionIntegrator.cu:47-166).
// Like the reference, the scalar logic runs on the host and needs the objective value after every sweep (one small
// read-back per evaluation); iterates, step sizes and path updates stay on the device.
void Integrator::TrustRegionStep(StepParams const& p, double dt, int iterations, int substeps)
{
    NvtxRange const zone("pbat.gpu.impl.vbd.TrustRegionIntegrator.SolveWithLinearAcceleratedPath");
    if (dAndVec.n == 0)
    {
        dAndVec.Alloc(static_cast<size_t>(nV) * 2, &deviceBytes);
        dAndSmall.Alloc(1, &deviceBytes);
    }
    float4* const xkm1 = dAndVec.p;
    float4* const xkm2 = xkm1 + nV;
    double const sdt   = dt / substeps;
    StepParams q   = p;
    q.substeps     = 1;
    q.skipPreStep  = 1;
    q.skipPostStep = 1;
    q.iterations   = 1;
    int const grid = Blocks(nV, 256);
    double const eta = static_cast<float>(trEta), tau = static_cast<float>(trTau);
    double const zero = 1e-6f, fltMin = 1.17549435e-38, fltMax = 3.40282347e+38;
    for (int s = 0; s < substeps; ++s)
    {
        LaunchPreStep(q);
        double fk = ObjectiveOfState(sdt), fkm1 = 0, fkm2 = 0, tk = -1.0, tkm1 = 0, tkm2 = 0, R2 = 0.0;
        // (the reference's xkm1 / xkm2 start as whatever the constructor left there; they are overwritten before use)
        auto updateIterates = [&]() {
            VBDX_CUDA(cudaMemcpyAsync(xkm2, xkm1, nV * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
            VBDX_CUDA(cudaMemcpyAsync(xkm1, dPos.p, nV * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
            fkm2 = fkm1, fkm1 = fk;
            tkm2 = tkm1, tkm1 = tk;
        };
        for (int k = 0; k < iterations; ++k)
        {
            q.iterBegin = k;
            if (k < 2)
            {
                updateIterates();
                LaunchStepKernel(q);
                fk = ObjectiveOfState(sdt);
                tk = static_cast<double>(k);
                continue;
            }
            // ConstructModel: parabola through (t, f) of the last three iterates, time translated so that t_k = 0
            tkm2 -= tk, tkm1 -= tk, tk = 0.0;
            double const d2 = fkm2 - fk, d1 = fkm1 - fk, det = tkm2 * tkm2 * tkm1 - tkm1 * tkm1 * tkm2;
            double const a2 = (d2 * tkm1 - d1 * tkm2) / det, a1 = (tkm2 * tkm2 * d1 - tkm1 * tkm1 * d2) / det, a0 = fk;
            updateIterates();
            LaunchStepKernel(q);
            VBDX_CUDA(cudaMemsetAsync(dAndSmall.p, 0, sizeof(double), stream));
            SquaredStepSize<<<std::min(grid, 1184), 256, 0, stream>>>(nV, dPos.p, xkm1, dAndSmall.p);
            ++kernelLaunches;
            double dx2 = 0;
            dAndSmall.Download(&dx2, 1, stream);
            VBDX_CUDA(cudaStreamSynchronize(stream));
            if (R2 < dx2 + zero)
                R2 = tau * tau * dx2;
            double const lower = 1.0, upper = std::sqrt(R2 / dx2);
            double tstar;
            if (std::fabs(a2) > fltMin)
                tstar = a2 > fltMin ? -a1 / (2.0 * a2) : fltMax;
            else if (std::fabs(a1) > fltMin)
                tstar = a1 > 0 ? -fltMax : fltMax;
            else
                tstar = 0.0;
            double const t     = tstar < lower ? lower : (upper < tstar ? upper : tstar);
            bool const tryStep = !(std::fabs(t - lower) < zero);
            bool accepted      = false;
            if (tryStep)
            {
                ScaleStep<<<grid, 256, 0, stream>>>(nV, dPos.p, xkm1, static_cast<float>(t), plan.nActive, nullptr);
                ++kernelLaunches;
                fk               = ObjectiveOfState(sdt);
                double const rho = (fkm1 - fk) / (fkm1 - (a2 * t * t + a1 * t + a0));
                accepted         = rho > eta;
            }
            if (accepted)
            {
                if ((upper - t) < zero)
                    R2 *= tau * tau;
                tk = t;
            }
            else
            {
                R2 /= tau * tau;
                if (tryStep)
                {
                    ScaleStep<<<grid, 256, 0, stream>>>(nV, dPos.p, xkm1, static_cast<float>(1.0 / t), plan.nActive, nullptr);
                    ++kernelLaunches;
                }
                fk = ObjectiveOfState(sdt);
                tk = tkm1 + 1;
            }
        }
        StepParams post = q;
        post.iterations = 0, post.skipPostStep = 0;
        LaunchStepKernel(post);  // velocity update only
    }
}

// A slice of one substep: [pre-step] [iterations kBegin .. kEnd of a solve of totalIterations] [velocity update].
// What TraceNextStep / TracedStep need to look at every iterate (sim/vbd/Integrator.cpp:47-52,73-75,202-235).
void Integrator::StepPartial(double sdt, int kBegin, int kEnd, int totalIterations, int flags)
{
    Require(sdt > 0 && kBegin >= 0 && kEnd >= kBegin && totalIterations >= kEnd, "StepPartial: bad iteration range");
    Require(!contact.enabled && distWorld == 1, "partial steps are not available with contact or domain decomposition");
    Require(acceleration == VBDX_ACCEL_NONE || acceleration == VBDX_ACCEL_CHEBYSHEV, "partial steps need the base or Chebyshev solve");
    VBDX_CUDA(cudaSetDevice(device));
    PrepareOmega(totalIterations);
    StepParams q   = MakeParams(sdt, kEnd - kBegin, 1);
    q.iterBegin    = kBegin;
    q.skipPreStep  = (flags & VBDX_PARTIAL_PRE_STEP) ? 0 : 1;
    q.skipPostStep = (flags & VBDX_PARTIAL_POST_STEP) ? 0 : 1;
    LaunchStepKernel(q);
    VBDX_CUDA(cudaStreamSynchronize(stream));
}

// f and (optionally) grad f at caller-supplied xk, xtilde (3 x nV column-major doubles, caller order)
void Integrator::Objective(const double* xk, const double* xtilde, double dt, double* f, double* grad)
{
    Require(xk && xtilde && (f || grad) && dt > 0, "Objective: need xk, xtilde, dt > 0 and an output");
    VBDX_CUDA(cudaSetDevice(device));
    if (dObjX.n == 0)
    {
        dObjX.Alloc(3 * nV, &deviceBytes), dObjXt.Alloc(3 * nV, &deviceBytes), dObjGrad.Alloc(3 * nV, &deviceBytes);
        dObjVal.Alloc(1, &deviceBytes);
    }
    dObjX.Upload(xk, 3 * nV, stream);
    dObjXt.Upload(xtilde, 3 * nV, stream);
    VBDX_CUDA(cudaMemsetAsync(dObjVal.p, 0, sizeof(double), stream));
    double* g = grad ? dObjGrad.p : nullptr;
    ObjectiveKinetic<<<std::min(Blocks(nV, 256), 1184), 256, 0, stream>>>(dObjX.p, dObjXt.p, dMass.p, nV, dObjVal.p, g);
    ObjectiveElastic<<<std::min(Blocks(nT, 256), 1184), 256, 0, stream>>>(dObjX.p, dE.p, dJinv.p, dVol.p, dLame.p, mu0, lambda0, nT, dt * dt,
                                                                            material == VBDX_MATERIAL_STVK ? 1 : 0, dObjVal.p, g);
    kernelLaunches += 2;
    if (f)
        dObjVal.Download(f, 1, stream);
    if (grad)
        dObjGrad.Download(grad, 3 * nV, stream);
    VBDX_CUDA(cudaStreamSynchronize(stream));
}

template <class T>
void Integrator::SetVertexField(T const* src, int64_t n, float4* dst0, float4* dst1, bool rows, bool sync)
{
    Require(src != nullptr && n == nV, "expected a 3 x nV array");  // gpu/impl/common/Eigen.cuh:28-35
    VBDX_CUDA(cudaSetDevice(device));
    T* staging = reinterpret_cast<T*>(dStaging.p);
    VBDX_CUDA(cudaMemcpyAsync(staging, src, 3 * nV * sizeof(T), cudaMemcpyHostToDevice, stream));
    ScatterFromCaller<T><<<Blocks(nV, 256), 256, 0, stream>>>(nV, dOld2New.p, staging, rows ? nV : 1, rows ? 1 : 3, dst0, dst1,
                                                                 distWorld > 1 ? plan.ghostBegin : nV);
    ++kernelLaunches;
    if (sync)
        VBDX_CUDA(cudaStreamSynchronize(stream));
}

template <class T>
void Integrator::GetVertexField(float4 const* src, T* dst, int64_t n, bool rows, bool sync)
{
    Require(dst != nullptr && n == nV, "expected a 3 x nV array");
    VBDX_CUDA(cudaSetDevice(device));
    T* staging = reinterpret_cast<T*>(dStaging.p);
    GatherToCaller<T><<<Blocks(nV, 256), 256, 0, stream>>>(nV, dOld2New.p, src, staging, rows ? nV : 1, rows ? 1 : 3);
    ++kernelLaunches;
    VBDX_CUDA(cudaMemcpyAsync(dst, staging, 3 * nV * sizeof(T), cudaMemcpyDeviceToHost, stream));
    if (sync)
        VBDX_CUDA(cudaStreamSynchronize(stream));
}

}  // namespace vbdx

// ============================================================================================
// C ABI
// ============================================================================================
using vbdx::Integrator;

struct vbdx_integrator {
    Integrator impl;
};

static thread_local std::string gLastError;

template <class F>
static vbdx_status Guard(F&& f)
{
    try
    {
        f();
        return VBDX_OK;
    }
    catch (vbdx::Error const& e)
    {
        gLastError = e.what();
        return e.status;
    }
    catch (std::bad_alloc const&)
    {
        gLastError = "host allocation failed";
        return VBDX_OUT_OF_MEMORY;
    }
    catch (std::exception const& e)
    {
        gLastError = e.what();
        return VBDX_CUDA_ERROR;
    }
}

static vbdx_status NeedHandle(vbdx_integrator* h)
{
    if (h)
        return VBDX_OK;
    gLastError = "null integrator handle";
    return VBDX_INVALID_ARGUMENT;
}

extern "C" {

const char* vbdx_last_error(void) { return gLastError.c_str(); }
int32_t vbdx_abi_version(void) { return VBDX_ABI_VERSION; }

int32_t vbdx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void vbdx_data_desc_init(vbdx_data_desc* d)
{
    std::memset(d, 0, sizeof(*d));
    d->abi_version  = VBDX_ABI_VERSION;
    d->struct_size  = sizeof(vbdx_data_desc);
    d->window_size  = 5;  // sim/vbd/Data.h:237
    d->nesterov_L = 1.0, d->nesterov_start = 3;  // sim/vbd/Data.h:238-239
    d->tr_eta = 0.2, d->tr_tau = 2.0, d->tr_curved = 1;  // sim/vbd/Data.h:241-243
    d->ordering     = VBDX_ORDER_LARGEST_DEGREE;  // sim/vbd/Data.h:211-216
    d->selection    = VBDX_SELECT_LEAST_USED;
    d->strategy     = VBDX_INIT_ADAPTIVE_PBAT;    // sim/vbd/Data.h:222-223
    d->acceleration = VBDX_ACCEL_NONE;
    d->omega_mode   = VBDX_OMEGA_REFERENCE;
    d->material     = VBDX_MATERIAL_STABLE_NEO_HOOKEAN;
    d->kD           = 0.0;
    d->detHZero     = 1e-7;
    d->rho          = 1.0;
    d->muC          = 1e6;
    d->muF          = 0.3;
    d->epsv         = 1e-3;
    d->active_set_update_frequency = 1;
    d->device       = -1;
}

vbdx_status vbdx_create(const vbdx_data_desc* desc, vbdx_integrator** out)
{
    if (!desc || !out)
    {
        gLastError = "vbdx_create: null argument";
        return VBDX_INVALID_ARGUMENT;
    }
    *out = nullptr;
    std::unique_ptr<vbdx_integrator> h;
    vbdx_status const st = Guard([&] {
        h = std::make_unique<vbdx_integrator>();
        h->impl.Create(*desc);
    });
    if (st == VBDX_OK)
        *out = h.release();
    return st;
}

// Independent scenes stepped as one problem (SURVEY.md 8e second row, BASELINE configs[4]): the scenes are
// concatenated into one disconnected mesh, so that one persistent launch per step sweeps every scene's colour c together.
// Colours are computed (or taken) per scene, which makes a scene inside a batch evolve bit-identically to the scene alone.
vbdx_status vbdx_create_batch(const vbdx_data_desc* descs, int32_t n, vbdx_integrator** out)
{
    if (!descs || !out || n < 1)
    {
        gLastError = "vbdx_create_batch: null argument or no scenes";
        return VBDX_INVALID_ARGUMENT;
    }
    *out = nullptr;
    std::unique_ptr<vbdx_integrator> h;
    vbdx_status const st = Guard([&] {
        using vbdx::Require;
        vbdx_data_desc const& d0 = descs[0];
        std::vector<int64_t> vOff(static_cast<size_t>(n) + 1, 0), tOff(static_cast<size_t>(n) + 1, 0);
        bool anyV = false, anyA = false, anyM = false, allM = true, anyRho = false, anyLame = false;
        int64_t nDbc = 0;
        for (int32_t s = 0; s < n; ++s)
        {
            vbdx_data_desc const& d = descs[s];
            Require(d.abi_version == VBDX_ABI_VERSION && d.struct_size == sizeof(vbdx_data_desc), "descriptor version / size mismatch (call vbdx_data_desc_init)");
            Require(d.nV > 0 && d.nT > 0 && d.X && d.E, "every scene needs a mesh");
            Require(d.nF == 0 && d.nCV == 0, "contact is not available in a batch (scenes share coordinates)");
            Require(d.nGhosts == 0, "ghost vertices are not available in a batch");
            Require(d.strategy == d0.strategy && d.acceleration == d0.acceleration && d.omega_mode == d0.omega_mode &&
                        d.material == d0.material && d.kD == d0.kD && d.detHZero == d0.detHZero && d.rho == d0.rho &&
                        d.flags == d0.flags && d.window_size == d0.window_size,
                    "the scenes of a batch must share strategy, acceleration, material and solver scalars");
            vOff[s + 1] = vOff[s] + d.nV;
            tOff[s + 1] = tOff[s] + d.nT;
            anyV |= d.v != nullptr, anyA |= d.aext != nullptr, anyM |= d.m != nullptr, allM &= d.m != nullptr;
            anyRho |= d.rhoe != nullptr, anyLame |= d.lame != nullptr;
            Require(d.nDbc >= 0 && (d.nDbc == 0 || d.dbc), "dbc pointer missing");
            nDbc += d.nDbc;
        }
        Require(!anyM || allM, "either every scene of a batch passes lumped masses or none does");
        int64_t const nV = vOff[n], nT = tOff[n];
        std::vector<double> X(3 * nV), v(anyV ? 3 * nV : 0, 0.0), a(anyA ? 3 * nV : 0), m(anyM ? nV : 0), rhoe(anyRho ? nT : 0, 1e3), lame(anyLame ? 2 * nT : 0);
        std::vector<int64_t> E(4 * nT), dbc(nDbc), colors(nV);
        double const Y = 1e6, nu = 0.45;  // sim/vbd/Data.cpp:199-205
        double const mu = Y / (2. * (1. + nu)), lam = (Y * nu) / ((1. + nu) * (1. - 2. * nu));
        int64_t kd = 0;
        for (int32_t s = 0; s < n; ++s)
        {
            vbdx_data_desc const& d = descs[s];
            std::copy(d.X, d.X + 3 * d.nV, X.begin() + 3 * vOff[s]);
            for (int64_t k = 0; k < 4 * d.nT; ++k)
            {
                Require(d.E[k] >= 0 && d.E[k] < d.nV, "element index out of range");
                E[4 * tOff[s] + k] = d.E[k] + vOff[s];
            }
            if (d.v)
                std::copy(d.v, d.v + 3 * d.nV, v.begin() + 3 * vOff[s]);
            if (anyA)
                for (int64_t i = 0; i < d.nV; ++i)
                    for (int c = 0; c < 3; ++c)
                        a[3 * (vOff[s] + i) + c] = d.aext ? d.aext[3 * i + c] : (c == 2 ? -9.81 : 0.0);  // sim/vbd/Data.cpp:191-195
            if (d.m)
                std::copy(d.m, d.m + d.nV, m.begin() + vOff[s]);
            if (d.rhoe)
                std::copy(d.rhoe, d.rhoe + d.nT, rhoe.begin() + tOff[s]);
            if (anyLame)
                for (int64_t e = 0; e < d.nT; ++e)
                {
                    lame[2 * (tOff[s] + e)]     = d.lame ? d.lame[2 * e] : mu;
                    lame[2 * (tOff[s] + e) + 1] = d.lame ? d.lame[2 * e + 1] : lam;
                }
            for (int64_t k = 0; k < d.nDbc; ++k)
            {
                Require(d.dbc[k] >= 0 && d.dbc[k] < d.nV, "Dirichlet vertex index out of range");
                dbc[kd++] = d.dbc[k] + vOff[s];
            }
            if (d.colors)
                std::copy(d.colors, d.colors + d.nV, colors.begin() + vOff[s]);
            else
            {
                std::vector<int64_t> c;
                vbdx::GreedyColorMesh(d.nV, d.nT, d.E, d.ordering, d.selection, c);
                std::copy(c.begin(), c.end(), colors.begin() + vOff[s]);
            }
        }
        vbdx_data_desc big = d0;
        big.nV = nV, big.nT = nT, big.X = X.data(), big.E = E.data();
        big.v = anyV ? v.data() : nullptr, big.aext = anyA ? a.data() : nullptr, big.m = anyM ? m.data() : nullptr;
        big.rhoe = anyRho ? rhoe.data() : nullptr, big.lame = anyLame ? lame.data() : nullptr;
        big.dbc = nDbc ? dbc.data() : nullptr, big.nDbc = nDbc, big.colors = colors.data();
        h = std::make_unique<vbdx_integrator>();
        h->impl.Create(big);
        h->impl.batchOffsets = vOff;
    });
    if (st == VBDX_OK)
        *out = h.release();
    return st;
}

vbdx_status vbdx_batch_offsets(vbdx_integrator* h, int32_t* n_scenes, int64_t* vertex_offsets)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] {
        vbdx::Require(!h->impl.batchOffsets.empty(), "not a batch handle (vbdx_create_batch)");
        if (n_scenes)
            *n_scenes = static_cast<int32_t>(h->impl.batchOffsets.size()) - 1;
        if (vertex_offsets)
            std::copy(h->impl.batchOffsets.begin(), h->impl.batchOffsets.end(), vertex_offsets);
    });
}

vbdx_status vbdx_destroy(vbdx_integrator* h)
{
    if (h)
    {
        cudaSetDevice(h->impl.device);
        cudaStreamSynchronize(h->impl.stream);
        delete h;
    }
    return VBDX_OK;
}

vbdx_status vbdx_step(vbdx_integrator* h, double dt, int32_t iterations, int32_t substeps)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] { h->impl.Step(dt, iterations, substeps, true); });
}

vbdx_status vbdx_step_async(vbdx_integrator* h, double dt, int32_t iterations, int32_t substeps)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] { h->impl.Step(dt, iterations, substeps, false); });
}

vbdx_status vbdx_synchronize(vbdx_integrator* h)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] {
        VBDX_CUDA(cudaSetDevice(h->impl.device));
        VBDX_CUDA(cudaStreamSynchronize(h->impl.stream));
        h->impl.CheckAsyncErrors();  // a time-out inside an asynchronous step surfaces here
        if (h->impl.stepTimed)
        {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, h->impl.evBegin, h->impl.evEnd) == cudaSuccess)
                h->impl.lastStepMs = ms;
        }
    });
}

#define VBDX_SETTER(name, T, dst0, dst1)                                                         \
    vbdx_status name(vbdx_integrator* h, const T* a, int64_t nV)                                 \
    {                                                                                            \
        if (vbdx_status s = NeedHandle(h))                                                       \
            return s;                                                                            \
        return Guard([&] { h->impl.SetVertexField<T>(a, nV, dst0, dst1); });                     \
    }
#define VBDX_GETTER(name, T, src)                                                                \
    vbdx_status name(vbdx_integrator* h, T* a, int64_t nV)                                       \
    {                                                                                            \
        if (vbdx_status s = NeedHandle(h))                                                       \
            return s;                                                                            \
        return Guard([&] { h->impl.GetVertexField<T>(src, a, nV); });                            \
    }
// positions live in Q (and in P when the Chebyshev double buffer exists); the step result is P
#define VBDX_POS_Q (h->impl.dPos.p)
#define VBDX_POS_P (h->impl.dHist.p ? h->impl.dPos.p + h->impl.nV : nullptr)
VBDX_SETTER(vbdx_set_positions_f32, float, VBDX_POS_Q, VBDX_POS_P)
VBDX_SETTER(vbdx_set_positions_f64, double, VBDX_POS_Q, VBDX_POS_P)
VBDX_SETTER(vbdx_set_velocities_f32, float, h->impl.dVel.p, nullptr)
VBDX_SETTER(vbdx_set_velocities_f64, double, h->impl.dVel.p, nullptr)
VBDX_SETTER(vbdx_set_external_acceleration_f32, float, h->impl.dAext.p, nullptr)
VBDX_SETTER(vbdx_set_external_acceleration_f64, double, h->impl.dAext.p, nullptr)
VBDX_GETTER(vbdx_get_positions_f32, float, (VBDX_POS_P ? VBDX_POS_P : VBDX_POS_Q))
VBDX_GETTER(vbdx_get_positions_f64, double, (VBDX_POS_P ? VBDX_POS_P : VBDX_POS_Q))
VBDX_GETTER(vbdx_get_velocities_f32, float, h->impl.dVel.p)
VBDX_GETTER(vbdx_get_velocities_f64, double, h->impl.dVel.p)

static vbdx_status SetVertexFieldImpl(vbdx_integrator* h, int32_t field, int32_t dtype, int32_t layout, const void* src, int64_t nV, bool sync)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] {
        vbdx::Require(field >= VBDX_FIELD_POSITIONS && field <= VBDX_FIELD_EXTERNAL_ACCELERATION, "unknown vertex field");
        vbdx::Require(dtype == VBDX_F32 || dtype == VBDX_F64, "unknown dtype");
        vbdx::Require(layout == VBDX_LAYOUT_COLUMNS || layout == VBDX_LAYOUT_ROWS, "unknown layout");
        float4* dst0 = field == VBDX_FIELD_POSITIONS ? VBDX_POS_Q : field == VBDX_FIELD_VELOCITIES ? h->impl.dVel.p : h->impl.dAext.p;
        float4* dst1 = field == VBDX_FIELD_POSITIONS ? VBDX_POS_P : nullptr;
        bool const rows = layout == VBDX_LAYOUT_ROWS;
        if (dtype == VBDX_F32)
            h->impl.SetVertexField<float>(static_cast<const float*>(src), nV, dst0, dst1, rows, sync);
        else
            h->impl.SetVertexField<double>(static_cast<const double*>(src), nV, dst0, dst1, rows, sync);
    });
}
vbdx_status vbdx_set_vertex_field(vbdx_integrator* h, int32_t field, int32_t dtype, int32_t layout, const void* src, int64_t nV)
{
    return SetVertexFieldImpl(h, field, dtype, layout, src, nV, true);
}
vbdx_status vbdx_set_vertex_field_async(vbdx_integrator* h, int32_t field, int32_t dtype, int32_t layout, const void* src, int64_t nV)
{
    return SetVertexFieldImpl(h, field, dtype, layout, src, nV, false);
}

static vbdx_status GetVertexFieldImpl(vbdx_integrator* h, int32_t field, int32_t dtype, int32_t layout, void* dst, int64_t nV, bool sync)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] {
        vbdx::Require(field == VBDX_FIELD_POSITIONS || field == VBDX_FIELD_VELOCITIES || field == VBDX_FIELD_INERTIAL_TARGET ||
                          field == VBDX_FIELD_PREVIOUS_POSITIONS,
                      "this field cannot be read back");
        vbdx::Require(dtype == VBDX_F32 || dtype == VBDX_F64, "unknown dtype");
        vbdx::Require(layout == VBDX_LAYOUT_COLUMNS || layout == VBDX_LAYOUT_ROWS, "unknown layout");
        float4 const* src = field == VBDX_FIELD_POSITIONS            ? (VBDX_POS_P ? VBDX_POS_P : VBDX_POS_Q)
                            : field == VBDX_FIELD_VELOCITIES         ? h->impl.dVel.p
                            : field == VBDX_FIELD_INERTIAL_TARGET    ? h->impl.dXtildeM.p
                                                                     : h->impl.dXt.p;
        bool const rows   = layout == VBDX_LAYOUT_ROWS;
        if (dtype == VBDX_F32)
            h->impl.GetVertexField<float>(src, static_cast<float*>(dst), nV, rows, sync);
        else
            h->impl.GetVertexField<double>(src, static_cast<double*>(dst), nV, rows, sync);
    });
}
vbdx_status vbdx_get_vertex_field(vbdx_integrator* h, int32_t field, int32_t dtype, int32_t layout, void* dst, int64_t nV)
{
    return GetVertexFieldImpl(h, field, dtype, layout, dst, nV, true);
}
vbdx_status vbdx_get_vertex_field_async(vbdx_integrator* h, int32_t field, int32_t dtype, int32_t layout, void* dst, int64_t nV)
{
    return GetVertexFieldImpl(h, field, dtype, layout, dst, nV, false);
}

vbdx_status vbdx_host_alloc(void** out, int64_t bytes)
{
    if (out == nullptr || bytes <= 0)
    {
        gLastError = "vbdx_host_alloc: null output or non-positive size";
        return VBDX_INVALID_ARGUMENT;
    }
    return Guard([&] { VBDX_CUDA(cudaHostAlloc(out, static_cast<size_t>(bytes), cudaHostAllocPortable)); });
}

vbdx_status vbdx_host_free(void* p)
{
    return Guard([&] {
        if (p)
            VBDX_CUDA(cudaFreeHost(p));
    });
}

vbdx_status vbdx_step_partial(vbdx_integrator* h, double sdt, int32_t k_begin, int32_t k_end, int32_t total_iterations, int32_t flags)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] { h->impl.StepPartial(sdt, k_begin, k_end, total_iterations, flags); });
}

vbdx_status vbdx_objective(vbdx_integrator* h, const double* xk, const double* xtilde, double dt, double* f, double* grad)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] { h->impl.Objective(xk, xtilde, dt, f, grad); });
}

vbdx_status vbdx_set_detH_zero(vbdx_integrator* h, double zero)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    h->impl.detHZero = zero;
    return VBDX_OK;
}

vbdx_status vbdx_set_rayleigh_damping(vbdx_integrator* h, double kD)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    h->impl.kD = kD;
    return VBDX_OK;
}

vbdx_status vbdx_set_initialization_strategy(vbdx_integrator* h, int32_t strategy)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    if (strategy < 0 || strategy > VBDX_INIT_ADAPTIVE_PBAT)
    {
        gLastError = "unknown initialization strategy";
        return VBDX_INVALID_ARGUMENT;
    }
    h->impl.strategy = strategy;
    return VBDX_OK;
}

vbdx_status vbdx_set_line_search_guard(vbdx_integrator* h, int32_t enabled)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    h->impl.lineSearch = enabled != 0 ? 1 : 0;
    return VBDX_OK;
}

vbdx_status vbdx_set_block_size(vbdx_integrator* h, int32_t block_size)
{
    (void)block_size;  // the sweep is warp-tiled; the reference's knob has no equivalent here
    return NeedHandle(h);
}

vbdx_status vbdx_set_scene_bounding_box(vbdx_integrator* h, const float min3[3], const float max3[3])
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] {
        vbdx::Require(min3 && max3 && max3[0] > min3[0] && max3[1] > min3[1] && max3[2] > min3[2], "scene bounding box must have positive extent");
        if (h->impl.contact.enabled)
        {
            VBDX_CUDA(cudaSetDevice(h->impl.device));
            h->impl.contact.SetWorldBox(min3, max3, h->impl.stream);
        }
    });
}

vbdx_status vbdx_get_internal_ids(vbdx_integrator* h, int64_t* old2new)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    for (int64_t i = 0; i < h->impl.nV; ++i)
        old2new[i] = h->impl.plan.old2new[i];
    return VBDX_OK;
}

vbdx_status vbdx_dist_ipc_handles(vbdx_integrator* h, void* out128)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
        auto& I = h->impl;
        VBDX_CUDA(cudaSetDevice(I.device));
        cudaIpcMemHandle_t hp, hf;
        VBDX_CUDA(cudaIpcGetMemHandle(&hp, I.dPos.p));
        VBDX_CUDA(cudaIpcGetMemHandle(&hf, I.dDistFlags.p));
        std::memcpy(out128, &hp, 64);
        std::memcpy(static_cast<char*>(out128) + 64, &hf, 64);
    });
}

vbdx_status vbdx_dist_connect(vbdx_integrator* h, int32_t rank, int32_t world, const void* all_handles, const int64_t* peer_nverts,
                              const int64_t* peer_nghosts, int64_t nSend, const int64_t* send_local, const int64_t* send_peer, const int64_t* send_remote,
                              uint32_t recv_mask)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] {
        auto& I = h->impl;
        vbdx::Require(world >= 1 && world <= 8 && rank >= 0 && rank < world, "domain decomposition supports 1..8 GPUs of one node");
        vbdx::Require(I.variant == VBDX_KERNEL_PIPELINED, "domain decomposition needs the pipelined kernel variant");
        vbdx::Require(!I.contact.enabled, "contact is not supported together with domain decomposition yet");
        VBDX_CUDA(cudaSetDevice(I.device));
        bool const cheb = I.acceleration == VBDX_ACCEL_CHEBYSHEV;
        for (int r = 0; r < world; ++r)
        {
            if (r == rank)
                continue;
            cudaIpcMemHandle_t hp, hf;
            std::memcpy(&hp, static_cast<const char*>(all_handles) + 128 * r, 64);
            std::memcpy(&hf, static_cast<const char*>(all_handles) + 128 * r + 64, 64);
            void *pp = nullptr, *pf = nullptr;
            VBDX_CUDA(cudaIpcOpenMemHandle(&pp, hp, cudaIpcMemLazyEnablePeerAccess));
            VBDX_CUDA(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
            I.ipcOpened[2 * r] = pp, I.ipcOpened[2 * r + 1] = pf;
            I.peerPos[r]   = static_cast<float4*>(pp);
            I.peerFlags[r] = static_cast<unsigned int*>(pf);
            I.peerPOff[r]  = cheb ? static_cast<uint32_t>(peer_nverts[r]) : 0u;
            vbdx::Require(peer_nghosts[r] >= 0 && peer_nghosts[r] <= peer_nverts[r], "bad peer ghost count");
            I.peerNGhost[r]     = static_cast<uint32_t>(peer_nghosts[r]);
            I.peerGhostBegin[r] = static_cast<uint32_t>(peer_nverts[r] - peer_nghosts[r]);  // ghosts are last in internal order
            I.peerGhostExt[r]   = static_cast<uint32_t>(peer_nverts[r]) * (cheb ? 2u : 1u);
        }
        // per-vertex send lists in internal order
        std::vector<uint32_t> ptr(I.nV + 1, 0), dst(static_cast<size_t>(nSend));
        for (int64_t k = 0; k < nSend; ++k)
        {
            vbdx::Require(send_local[k] >= 0 && send_local[k] < I.nV && send_peer[k] >= 0 && send_peer[k] < world && send_peer[k] != rank &&
                              send_remote[k] >= I.peerGhostBegin[send_peer[k]] &&
                              send_remote[k] < I.peerGhostBegin[send_peer[k]] + I.peerNGhost[send_peer[k]] && send_remote[k] < (int64_t(1) << 28),
                          "bad send list entry (the remote slot must be a ghost of that peer)");
            int32_t const vi = I.plan.old2new[send_local[k]];
            vbdx::Require(vi < I.plan.ghostBegin, "a ghost vertex cannot be sent");
            ++ptr[vi + 1];
        }
        for (int64_t i = 0; i < I.nV; ++i)
            ptr[i + 1] += ptr[i];
        std::vector<uint32_t> cur(ptr.begin(), ptr.end() - 1);
        for (int64_t k = 0; k < nSend; ++k)
        {
            int32_t const vi = I.plan.old2new[send_local[k]];
            dst[cur[vi]++]   = (static_cast<uint32_t>(send_peer[k]) << 28) | static_cast<uint32_t>(send_remote[k]);
        }
        I.dSendPtr.Alloc(ptr.size(), &I.deviceBytes);
        I.dSendDst.Alloc(dst.size() + 1, &I.deviceBytes);
        I.dSendPtr.Upload(ptr.data(), ptr.size(), I.stream);
        if (!dst.empty())
            I.dSendDst.Upload(dst.data(), dst.size(), I.stream);
        VBDX_CUDA(cudaStreamSynchronize(I.stream));
        I.distRank = rank, I.distWorld = world;
        I.peerMask = recv_mask;
        for (int64_t k = 0; k < nSend; ++k)
            I.peerMask |= 1u << send_peer[k];
    });
}

vbdx_status vbdx_dist_stats(vbdx_integrator* h, uint32_t out4[4], int32_t reset)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] {
        auto& I = h->impl;
        VBDX_CUDA(cudaSetDevice(I.device));
        VBDX_CUDA(cudaStreamSynchronize(I.stream));
        VBDX_CUDA(cudaMemcpy(out4, I.dDistFlags.p + 10, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (reset)
            VBDX_CUDA(cudaMemset(I.dDistFlags.p + 10, 0, 4 * sizeof(uint32_t)));
    });
}

vbdx_status vbdx_get_contact_state(vbdx_integrator* h, int32_t* active, int32_t* nn, int64_t* nActive)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] {
        auto& I  = h->impl;
        auto& cs = I.contact;
        vbdx::Require(cs.enabled, "the integrator was created without a collision mesh");
        VBDX_CUDA(cudaSetDevice(I.device));
        std::vector<uint8_t> a(cs.nCV);
        cs.active.Download(a.data(), cs.nCV, I.stream);
        if (nn)
            cs.nn.Download(nn, static_cast<size_t>(cs.nCV) * vbdx::kMaxContacts, I.stream);
        uint32_t na = 0;
        cs.nActive.Download(&na, 1, I.stream);
        VBDX_CUDA(cudaStreamSynchronize(I.stream));
        if (active)
            for (uint32_t k = 0; k < cs.nCV; ++k)
                active[k] = a[k];
        if (nActive)
            *nActive = na;
    });
}

vbdx_status vbdx_debug_bvh_build(int64_t n, const float* lo, const float* hi, const float wmin[3], const float wmax[3], int32_t* child,
                                 int32_t* parent, int32_t* rightmost, int32_t* inds, uint32_t* codes, float* nodeLo, float* nodeHi)
{
    return Guard([&] {
        vbdx::Require(n >= 1 && lo && hi && wmin && wmax, "vbdx_debug_bvh_build: bad arguments");
        if (vbdx_device_count() == 0)
            throw vbdx::Error(VBDX_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
        cudaStream_t s = nullptr;
        int64_t bytes = 0, launches = 0;
        vbdx::DeviceBvh bvh;
        bvh.Alloc(static_cast<uint32_t>(n), &bytes);
        std::vector<float4> l(n), u(n);
        for (int64_t i = 0; i < n; ++i)
        {
            l[i] = make_float4(lo[3 * i], lo[3 * i + 1], lo[3 * i + 2], 0.f);
            u[i] = make_float4(hi[3 * i], hi[3 * i + 1], hi[3 * i + 2], 0.f);
        }
        vbdx::DevBuf<float4> dl, du;
        vbdx::DevBuf<vbdx::WorldBox> dw;
        dl.Alloc(n), du.Alloc(n), dw.Alloc(1);
        dl.Upload(l.data(), n, s), du.Upload(u.data(), n, s);
        vbdx::WorldBox w;
        for (int d = 0; d < 3; ++d)
            w.lo[d] = wmin[d], w.ext[d] = wmax[d] - wmin[d];
        dw.Upload(&w, 1, s);
        bvh.Build(dl.p, du.p, dw.p, s, &launches);
        VBDX_CUDA(cudaStreamSynchronize(s));
        size_t const ni = static_cast<size_t>(n - 1);
        if (child && ni)
        {
            bvh.child0.Download(child, ni, s);
            bvh.child1.Download(child + ni, ni, s);
        }
        if (rightmost && ni)
        {
            bvh.right0.Download(rightmost, ni, s);
            bvh.right1.Download(rightmost + ni, ni, s);
        }
        if (parent)
            bvh.parent.Download(parent, 2 * n - 1, s);
        if (inds)
            bvh.inds.Download(reinterpret_cast<uint32_t*>(inds), n, s);
        if (codes)
            bvh.codes.Download(codes, n, s);
        std::vector<float4> nl(2 * n - 1), nh(2 * n - 1);
        bvh.nodeLo.Download(nl.data(), 2 * n - 1, s);
        bvh.nodeHi.Download(nh.data(), 2 * n - 1, s);
        VBDX_CUDA(cudaStreamSynchronize(s));
        for (int64_t k = 0; k < 2 * n - 1; ++k)
        {
            if (nodeLo)
                nodeLo[3 * k] = nl[k].x, nodeLo[3 * k + 1] = nl[k].y, nodeLo[3 * k + 2] = nl[k].z;
            if (nodeHi)
                nodeHi[3 * k] = nh[k].x, nodeHi[3 * k + 1] = nh[k].y, nodeHi[3 * k + 2] = nh[k].z;
        }
    });
}

// ---------------------------------------------------------------------------------------------------------------
// stand-alone LBVH and vertex-triangle detector (geometry_api.cuh)
// ---------------------------------------------------------------------------------------------------------------
struct vbdx_bvh {
    vbdx::BvhHandle impl;
};
struct vbdx_contact {
    vbdx::ContactHandle impl;
};

static void NeedDevice()
{
    if (vbdx_device_count() == 0)
        throw vbdx::Error(VBDX_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
}

vbdx_status vbdx_bvh_create(int64_t max_boxes, vbdx_bvh** out)
{
    if (!out)
        return VBDX_INVALID_ARGUMENT;
    *out = nullptr;
    std::unique_ptr<vbdx_bvh> h;
    vbdx_status const st = Guard([&] {
        vbdx::Require(max_boxes >= 1 && max_boxes < (int64_t(1) << 30), "vbdx_bvh_create: max_boxes out of range");
        NeedDevice();
        h = std::make_unique<vbdx_bvh>();
        auto& b = h->impl;
        VBDX_CUDA(cudaGetDevice(&b.device));
        b.capacity = max_boxes;
        b.bvh.Alloc(static_cast<uint32_t>(max_boxes), &b.bytes);
        b.lo.Alloc(max_boxes, &b.bytes), b.hi.Alloc(max_boxes, &b.bytes), b.world.Alloc(1, &b.bytes);
    });
    if (st == VBDX_OK)
        *out = h.release();
    return st;
}

vbdx_status vbdx_bvh_destroy(vbdx_bvh* h)
{
    delete h;
    return VBDX_OK;
}

vbdx_status vbdx_bvh_build(vbdx_bvh* h, int64_t n, const float* lo, const float* hi, const float wmin[3], const float wmax[3])
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] {
        auto& b = h->impl;
        vbdx::Require(n >= 1 && n <= b.capacity && lo && hi && wmin && wmax, "vbdx_bvh_build: bad arguments (n must be in [1, max_boxes])");
        VBDX_CUDA(cudaSetDevice(b.device));
        cudaStream_t s = nullptr;
        vbdx::UploadXyz(b.lo, lo, n, s);
        vbdx::UploadXyz(b.hi, hi, n, s);
        vbdx::WorldBox w;
        for (int d = 0; d < 3; ++d)
            w.lo[d] = wmin[d], w.ext[d] = wmax[d] - wmin[d];
        b.world.Upload(&w, 1, s);
        b.n     = n;
        b.bvh.n = static_cast<uint32_t>(n);
        b.bvh.Build(b.lo.p, b.hi.p, b.world.p, s, &b.launches);
        VBDX_CUDA(cudaStreamSynchronize(s));
    });
}

vbdx_status vbdx_bvh_get(vbdx_bvh* h, int32_t* child, int32_t* parent, int32_t* rightmost, int32_t* inds, uint32_t* codes, float* node_lo,
                         float* node_hi, int32_t* visits)
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] {
        auto& b = h->impl;
        vbdx::Require(b.n >= 1, "vbdx_bvh_get: build() has not been called");
        VBDX_CUDA(cudaSetDevice(b.device));
        cudaStream_t s  = nullptr;
        int64_t const n = b.n;
        size_t const ni = static_cast<size_t>(n - 1);
        if (child && ni)
            b.bvh.child0.Download(child, ni, s), b.bvh.child1.Download(child + ni, ni, s);
        if (rightmost && ni)
            b.bvh.right0.Download(rightmost, ni, s), b.bvh.right1.Download(rightmost + ni, ni, s);
        if (parent)
            b.bvh.parent.Download(parent, 2 * n - 1, s);
        if (inds)
            b.bvh.inds.Download(reinterpret_cast<uint32_t*>(inds), n, s);
        if (codes)
            b.bvh.codes.Download(codes, n, s);
        std::vector<uint32_t> arrivals(visits && ni ? ni : 0);
        if (visits && ni)
            b.bvh.visits.Download(arrivals.data(), ni, s);
        std::vector<float4> nl(2 * n - 1), nh(2 * n - 1);
        b.bvh.nodeLo.Download(nl.data(), 2 * n - 1, s);
        b.bvh.nodeHi.Download(nh.data(), 2 * n - 1, s);
        VBDX_CUDA(cudaStreamSynchronize(s));
        // the reference resets its visit counters per box computation (2 = both children arrived, gpu/impl/geometry/Bvh.cu:209);
        // the counters here run on, two arrivals per computation: report the last computation's
        for (size_t k = 0; k < arrivals.size(); ++k)
            visits[k] = arrivals[k] == 0u ? 0 : 2 - static_cast<int32_t>(arrivals[k] & 1u);
        for (int64_t k = 0; k < 2 * n - 1; ++k)
        {
            if (node_lo)
                node_lo[3 * k] = nl[k].x, node_lo[3 * k + 1] = nl[k].y, node_lo[3 * k + 2] = nl[k].z;
            if (node_hi)
                node_hi[3 * k] = nh[k].x, node_hi[3 * k + 1] = nh[k].y, node_hi[3 * k + 2] = nh[k].z;
        }
    });
}

vbdx_status vbdx_bvh_detect_overlaps(vbdx_bvh* h, const int32_t* set, int64_t max_overlaps, int32_t* pairs, int64_t* n_found)
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] {
        auto& b = h->impl;
        vbdx::Require(b.n >= 1 && max_overlaps >= 0 && (pairs || max_overlaps == 0) && n_found, "vbdx_bvh_detect_overlaps: bad arguments");
        VBDX_CUDA(cudaSetDevice(b.device));
        cudaStream_t s = nullptr;
        vbdx::DevBuf<int32_t> dSet, dPairs;
        vbdx::DevBuf<unsigned long long> dCount;
        if (set)
        {
            dSet.Alloc(b.n);
            dSet.Upload(set, b.n, s);
        }
        dPairs.Alloc(2 * static_cast<size_t>(std::max<int64_t>(max_overlaps, 1)));
        dCount.Alloc(1);
        VBDX_CUDA(cudaMemsetAsync(dCount.p, 0, sizeof(unsigned long long), s));
        vbdx::BvhSelfOverlaps<<<vbdx::Blocks(b.n, 128), 128, 0, s>>>(b.bvh.View(), dSet.p, dPairs.p, dCount.p, static_cast<unsigned long long>(max_overlaps));
        ++b.launches;
        unsigned long long found = 0;
        dCount.Download(&found, 1, s);
        VBDX_CUDA(cudaStreamSynchronize(s));
        *n_found = static_cast<int64_t>(found);
        int64_t const keep = std::min<int64_t>(static_cast<int64_t>(found), max_overlaps);
        if (keep > 0)
        {
            dPairs.Download(pairs, 2 * static_cast<size_t>(keep), s);
            VBDX_CUDA(cudaStreamSynchronize(s));
        }
    });
}

vbdx_status vbdx_bvh_nearest_triangles(vbdx_bvh* h, int64_t nQ, const float* X, int64_t nP, const float* V, int64_t nF, const int32_t* F, int32_t* out)
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] {
        auto& b = h->impl;
        vbdx::Require(b.n >= 1 && nF == b.n, "vbdx_bvh_nearest_triangles: the tree must have been built over these triangles' boxes");
        vbdx::Require(nQ >= 0 && nP >= 1 && X && V && F && out, "vbdx_bvh_nearest_triangles: bad arguments");
        if (nQ == 0)
            return;
        VBDX_CUDA(cudaSetDevice(b.device));
        cudaStream_t s = nullptr;
        vbdx::DevBuf<float4> dX, dV;
        vbdx::DevBuf<int4> dF;
        vbdx::DevBuf<int32_t> dOut;
        dX.Alloc(nQ), dV.Alloc(nP), dF.Alloc(nF), dOut.Alloc(nQ);
        vbdx::UploadXyz(dX, X, nQ, s);
        vbdx::UploadXyz(dV, V, nP, s);
        std::vector<int4> Fh(static_cast<size_t>(nF));
        for (int64_t f = 0; f < nF; ++f)
        {
            for (int k = 0; k < 3; ++k)
                vbdx::Require(F[3 * f + k] >= 0 && F[3 * f + k] < nP, "triangle index out of range");
            Fh[f] = make_int4(F[3 * f], F[3 * f + 1], F[3 * f + 2], 0);
        }
        dF.Upload(Fh.data(), nF, s);
        vbdx::BvhNearestTriangle<<<vbdx::Blocks(nQ, 128), 128, 0, s>>>(b.bvh.View(), dX.p, nQ, dV.p, dF.p, dOut.p);
        ++b.launches;
        dOut.Download(out, nQ, s);
        VBDX_CUDA(cudaStreamSynchronize(s));
    });
}

vbdx_status vbdx_contact_create(int64_t nV, const int64_t* B, const int64_t* V, int64_t nCV, const int64_t* F, int64_t nF, vbdx_contact** out)
{
    if (!out)
        return VBDX_INVALID_ARGUMENT;
    *out = nullptr;
    std::unique_ptr<vbdx_contact> h;
    vbdx_status const st = Guard([&] {
        vbdx::Require(nV >= 1 && nCV >= 1 && nF >= 1 && V && F, "vbdx_contact_create: a collision mesh is needed");
        NeedDevice();
        h = std::make_unique<vbdx_contact>();
        auto& c = h->impl;
        VBDX_CUDA(cudaGetDevice(&c.device));
        c.nV = nV;
        auto& cs = c.cs;
        cs.nCV = static_cast<uint32_t>(nCV), cs.nF = static_cast<uint32_t>(nF);
        std::vector<int32_t> Bh(nV), Vh(nCV);
        std::vector<int4> Fh(nF);
        for (int64_t i = 0; i < nV; ++i)
            Bh[i] = B ? static_cast<int32_t>(B[i]) : 1;
        for (int64_t k = 0; k < nCV; ++k)
        {
            vbdx::Require(V[k] >= 0 && V[k] < nV, "collision vertex index out of range");
            Vh[k] = static_cast<int32_t>(V[k]);
        }
        for (int64_t f = 0; f < nF; ++f)
        {
            for (int k = 0; k < 3; ++k)
                vbdx::Require(F[3 * f + k] >= 0 && F[3 * f + k] < nV, "collision triangle index out of range");
            Fh[f] = make_int4(static_cast<int>(F[3 * f]), static_cast<int>(F[3 * f + 1]), static_cast<int>(F[3 * f + 2]), Bh[F[3 * f]]);
        }
        cudaStream_t s = nullptr;
        cs.B.Alloc(nV, &c.bytes), cs.V.Alloc(nCV, &c.bytes), cs.F.Alloc(nF, &c.bytes);
        cs.B.Upload(Bh.data(), nV, s), cs.V.Upload(Vh.data(), nCV, s), cs.F.Upload(Fh.data(), nF, s);
        cs.mesh = vbdx::ContactMesh{cs.B.p, cs.V.p, cs.F.p, cs.nCV, cs.nF};
        cs.Alloc(nV, &c.bytes, s);
        cs.enabled = true;
        c.xa.Alloc(nV, &c.bytes), c.xb.Alloc(nV, &c.bytes), c.zero.Alloc(nV, &c.bytes);
        VBDX_CUDA(cudaMemsetAsync(c.zero.p, 0, nV * sizeof(float4), s));
        VBDX_CUDA(cudaStreamSynchronize(s));
    });
    if (st == VBDX_OK)
        *out = h.release();
    return st;
}

vbdx_status vbdx_contact_destroy(vbdx_contact* h)
{
    delete h;
    return VBDX_OK;
}

vbdx_status vbdx_contact_initialize_active_set(vbdx_contact* h, const float* xt, const float* xtp1, const float wmin[3], const float wmax[3])
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] {
        auto& c = h->impl;
        vbdx::Require(xt && xtp1 && wmin && wmax, "vbdx_contact_initialize_active_set: null argument");
        VBDX_CUDA(cudaSetDevice(c.device));
        cudaStream_t s = nullptr;
        // the detector's predictor is x + dt v + dt^2 a (gpu/impl/vbd/Integrator.cu:163-188): v = xtp1 - xt, dt = 1, a = 0
        vbdx::UploadXyz(c.xa, xt, c.nV, s);
        vbdx::UploadXyz(c.xb, xtp1, c.nV, s, xt);
        c.cs.SetWorldBox(wmin, wmax, s);
        c.cs.InitializeActiveSet(c.xa.p, c.xb.p, c.zero.p, c.nV, 1.f, s, &c.launches);
        VBDX_CUDA(cudaStreamSynchronize(s));
    });
}

static vbdx_status ContactNearest(vbdx_contact* h, const float* x, int mode)
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] {
        auto& c = h->impl;
        vbdx::Require(x != nullptr, "positions missing");
        VBDX_CUDA(cudaSetDevice(c.device));
        cudaStream_t s = nullptr;
        vbdx::UploadXyz(c.xa, x, c.nV, s);
        c.cs.eps = c.eps;
        c.cs.NearestPass(c.xa.p, mode, s, &c.launches);
        VBDX_CUDA(cudaStreamSynchronize(s));
    });
}

vbdx_status vbdx_contact_update_active_set(vbdx_contact* h, const float* x) { return ContactNearest(h, x, 0); }
vbdx_status vbdx_contact_finalize_active_set(vbdx_contact* h, const float* x) { return ContactNearest(h, x, 1); }

vbdx_status vbdx_contact_set_eps(vbdx_contact* h, float eps)
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    h->impl.eps = eps;
    return VBDX_OK;
}

vbdx_status vbdx_contact_get(vbdx_contact* h, int32_t* active_mask, int32_t* nn, int32_t* av, int64_t* n_active)
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] {
        auto& c  = h->impl;
        auto& cs = c.cs;
        VBDX_CUDA(cudaSetDevice(c.device));
        cudaStream_t s = nullptr;
        std::vector<uint8_t> a(cs.nCV);
        cs.active.Download(a.data(), cs.nCV, s);
        if (nn)
            cs.nn.Download(nn, static_cast<size_t>(cs.nCV) * vbdx::kMaxContacts, s);
        if (av)
            cs.av.Download(av, cs.nCV, s);
        uint32_t na = 0;
        cs.nActive.Download(&na, 1, s);
        VBDX_CUDA(cudaStreamSynchronize(s));
        if (active_mask)
            for (uint32_t k = 0; k < cs.nCV; ++k)
                active_mask[k] = a[k];
        if (n_active)
            *n_active = na;
    });
}

// Test hooks: the sweep's contact term and penalty scaling on caller-supplied inputs, so that the GPU tests can put
// csrc/contact.cuh next to the reference's own functions (compiled by the test infrastructure).  Layouts: include/vbdx.h.
}  // extern "C"
namespace vbdx {
__global__ void DebugContactPairs(int n, const float* in, float* out)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float* q = in + 28 * i;
    float3 const xtv = make_float3(q[0], q[1], q[2]), xv = make_float3(q[3], q[4], q[5]);
    float3 xtf[3], xf[3];
    for (int c = 0; c < 3; ++c)
    {
        xtf[c] = make_float3(q[6 + 3 * c], q[7 + 3 * c], q[8 + 3 * c]);
        xf[c]  = make_float3(q[15 + 3 * c], q[16 + 3 * c], q[17 + 3 * c]);
    }
    float g[3] = {0.f, 0.f, 0.f}, H[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    AccumulateVertexTriangleContact(xtv, xv, xtf, xf, q[24], q[25], q[26], q[27], g, H);
    float* o = out + 13 * i;
    o[0]     = 0.f;  // the sweep needs no energy
    o[1] = g[0], o[2] = g[1], o[3] = g[2];
    o[4] = H[0], o[5] = H[1], o[6] = H[2];
    o[7] = H[1], o[8] = H[3], o[9] = H[4];
    o[10] = H[2], o[11] = H[4], o[12] = H[5];
}
__global__ void DebugContactPenalties(int nVerts, const int* fc, const float* XVA, const float* FA, float muC, int* nContacts, float* penalty)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nVerts)
        return;
    int f[kMaxContacts];
    int n          = 0;
    float const kC = ContactPenaltyScale(fc + static_cast<size_t>(i) * kMaxContacts, FA, XVA[i], muC, f, n);
    nContacts[i]   = n;
    for (int c = 0; c < kMaxContacts; ++c)
        penalty[kMaxContacts * i + c] = c < n ? kC * FA[f[c]] : 0.f;
}
}  // namespace vbdx
extern "C" {

vbdx_status vbdx_debug_contact_pairs(int32_t n, const float* in28, float* out13)
{
    return Guard([&] {
        vbdx::Require(n >= 1 && in28 && out13, "vbdx_debug_contact_pairs: bad arguments");
        NeedDevice();
        vbdx::DevBuf<float> di, dout;
        di.Alloc(static_cast<size_t>(28) * n), dout.Alloc(static_cast<size_t>(13) * n);
        di.Upload(in28, static_cast<size_t>(28) * n, nullptr);
        vbdx::DebugContactPairs<<<vbdx::Blocks(n, 128), 128>>>(n, di.p, dout.p);
        dout.Download(out13, static_cast<size_t>(13) * n, nullptr);
        VBDX_CUDA(cudaDeviceSynchronize());
    });
}

vbdx_status vbdx_debug_contact_penalties(int32_t nVerts, int32_t nTris, const int32_t* fc, const float* XVA, const float* FA, float muC,
                                         int32_t* nContacts, float* penalty)
{
    return Guard([&] {
        vbdx::Require(nVerts >= 1 && nTris >= 1 && fc && XVA && FA && nContacts && penalty, "vbdx_debug_contact_penalties: bad arguments");
        NeedDevice();
        vbdx::DevBuf<int32_t> dfc, dn;
        vbdx::DevBuf<float> dx, df, dp;
        dfc.Alloc(static_cast<size_t>(8) * nVerts), dn.Alloc(nVerts), dx.Alloc(nVerts), df.Alloc(nTris), dp.Alloc(static_cast<size_t>(8) * nVerts);
        dfc.Upload(fc, static_cast<size_t>(8) * nVerts, nullptr), dx.Upload(XVA, nVerts, nullptr), df.Upload(FA, nTris, nullptr);
        vbdx::DebugContactPenalties<<<vbdx::Blocks(nVerts, 128), 128>>>(nVerts, dfc.p, dx.p, df.p, muC, dn.p, dp.p);
        dn.Download(nContacts, nVerts, nullptr), dp.Download(penalty, static_cast<size_t>(8) * nVerts, nullptr);
        VBDX_CUDA(cudaDeviceSynchronize());
    });
}

// Host-only: the planner's output for a mesh (no device needed).  Lets the CPU test-suite model-check the protocol of the
// barrier-free sweep against the very ring lists, flags and padding the kernels consume.
struct vbdx_plan {
    vbdx::Plan plan;
};

vbdx_status vbdx_debug_plan_create(int64_t nV, int64_t nT, const int64_t* E, const int64_t* colors, const uint8_t* is_constrained,
                                   const double* X, int32_t tile_iters, vbdx_plan** out)
{
    if (!out)
        return VBDX_INVALID_ARGUMENT;
    *out = nullptr;
    std::unique_ptr<vbdx_plan> h;
    vbdx_status const st = Guard([&] {
        vbdx::Require(nV >= 1 && nT >= 1 && E && colors && is_constrained && X, "vbdx_debug_plan_create: bad arguments");
        std::vector<int32_t> E32(static_cast<size_t>(4 * nT));
        std::vector<uint32_t> ptr(static_cast<size_t>(nV) + 1, 0), adj(static_cast<size_t>(4 * nT));
        for (int64_t k = 0; k < 4 * nT; ++k)
        {
            vbdx::Require(E[k] >= 0 && E[k] < nV, "element index out of range");
            E32[k] = static_cast<int32_t>(E[k]);
            ++ptr[E[k] + 1];
        }
        for (int64_t i = 0; i < nV; ++i)
            ptr[i + 1] += ptr[i];
        std::vector<uint32_t> cursor(ptr.begin(), ptr.end() - 1);
        for (int64_t k = 0; k < 4 * nT; ++k)  // ascending k per vertex = ascending element id (sim/vbd/Data.cpp:223-226)
            adj[cursor[E32[k]]++] = static_cast<uint32_t>(k);
        h = std::make_unique<vbdx_plan>();
        try
        {
            vbdx::BuildPlan(nV, E32.data(), ptr.data(), adj.data(), colors, is_constrained, X, tile_iters > 0 ? tile_iters : 8, false, 1, 0, h->plan);
        }
        catch (std::length_error const& e)
        {
            throw vbdx::Error(VBDX_UNSUPPORTED, e.what());
        }
    });
    if (st == VBDX_OK)
        *out = h.release();
    return st;
}

/* what: 0 sizes {nTiles, nRingIds, nColors, nActive, ghostBegin} (int64 x 5), 1 tiles (uint32 x 4 per tile), 2 ring ids, 3 colour tile
 * begins (nColors + 1), 4 new2old (int32 x nV) */
vbdx_status vbdx_debug_plan_get(vbdx_plan* h, int32_t what, void* out)
{
    if (!h || !out)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] {
        vbdx::Plan const& p = h->plan;
        switch (what)
        {
            case 0: {
                int64_t* o = static_cast<int64_t*>(out);
                o[0] = static_cast<int64_t>(p.tiles.size()), o[1] = static_cast<int64_t>(p.ringIds.size()), o[2] = p.nColors, o[3] = p.nActive, o[4] = p.ghostBegin;
                break;
            }
            case 1: std::memcpy(out, p.tiles.data(), p.tiles.size() * sizeof(vbdx::TileDesc)); break;
            case 2: std::memcpy(out, p.ringIds.data(), p.ringIds.size() * sizeof(uint32_t)); break;
            case 3: std::memcpy(out, p.colorTileBegin.data(), p.colorTileBegin.size() * sizeof(uint32_t)); break;
            case 4: std::memcpy(out, p.new2old.data(), p.new2old.size() * sizeof(int32_t)); break;
            default: throw vbdx::Error(VBDX_INVALID_ARGUMENT, "vbdx_debug_plan_get: unknown item");
        }
    });
}

vbdx_status vbdx_debug_plan_destroy(vbdx_plan* h)
{
    delete h;
    return VBDX_OK;
}

vbdx_status vbdx_set_stream(vbdx_integrator* h, void* cuda_stream)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] {
        VBDX_CUDA(cudaStreamSynchronize(h->impl.stream));
        h->impl.stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->impl.ownStream;
    });
}

vbdx_status vbdx_debug_trace(vbdx_integrator* h, int32_t iteration, unsigned long long* out, int64_t capacity)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] {
        auto& I = h->impl;
        VBDX_CUDA(cudaSetDevice(I.device));
        size_t const n = static_cast<size_t>(I.plan.nColors) * I.gridBlocks * vbdx::kTraceStamps;
        if (out == nullptr)
        {
            // arm: the next steps record the timestamps of `iteration` (direct kernel only)
            if (I.dTrace.n < n)
                I.dTrace.Alloc(n, &I.deviceBytes);
            VBDX_CUDA(cudaMemsetAsync(I.dTrace.p, 0, n * sizeof(unsigned long long), I.stream));
            I.traceIteration = iteration;
            return;
        }
        vbdx::Require(capacity >= static_cast<int64_t>(n) && I.dTrace.p != nullptr, "trace buffer too small or tracing not armed");
        I.dTrace.Download(out, n, I.stream);
        VBDX_CUDA(cudaStreamSynchronize(I.stream));
        I.traceIteration = -1;
    });
}

vbdx_status vbdx_get_info(vbdx_integrator* h, vbdx_info* out)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    auto const& I        = h->impl;
    out->nV              = I.nV;
    out->nT              = I.nT;
    out->nActiveVertices = I.plan.nActive;
    out->nIncidences     = I.plan.nIncidences;
    out->nRecordSlots    = I.nRecordSlots;
    out->nColors         = I.plan.nColors;
    out->nTiles          = static_cast<int32_t>(I.plan.tiles.size());
    // launch shape of the kernel that runs whole steps: the lean barrier-free one where it applies
    bool const flowShape = I.dFlowTiles.p != nullptr && I.dataflow && I.flowKernel && (!I.contact.enabled || I.dHist4.p != nullptr);
    out->gridBlocks      = flowShape ? I.flowGridBlocks : I.gridBlocks;
    out->blockThreads    = flowShape ? I.flowWarps * 32 : I.blockThreads;
    out->device          = I.device;
    out->smCount         = I.smCount;
    out->deviceBytes     = I.deviceBytes;
    out->kernelLaunches  = I.kernelLaunches;
    out->lastStepMs      = I.lastStepMs;
    out->nRingEntries    = I.plan.nRingEntries;
    out->nGhosts         = I.nGhost;
    out->nonFiniteVertices = 0;
    if (I.dBarrier.p != nullptr)
    {
        unsigned int bad = 0;
        cudaSetDevice(I.device);
        if (cudaStreamSynchronize(I.stream) == cudaSuccess && cudaMemcpy(&bad, I.dBarrier.p + 2, sizeof(bad), cudaMemcpyDeviceToHost) == cudaSuccess)
            out->nonFiniteVertices = bad;
        else
            cudaGetLastError();
    }
    return VBDX_OK;
}

vbdx_status vbdx_get_adjacency(vbdx_integrator* h, int64_t* GVGp, int64_t* GVGe, int64_t* GVGilocal)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] {
        auto& I = h->impl;
        VBDX_CUDA(cudaSetDevice(I.device));
        std::vector<uint32_t> ptr(I.nV + 1), adj(4 * I.nT);
        I.dPtr.Download(ptr.data(), ptr.size(), I.stream);
        I.dAdj.Download(adj.data(), adj.size(), I.stream);
        VBDX_CUDA(cudaStreamSynchronize(I.stream));
        if (GVGp)
            for (size_t k = 0; k < ptr.size(); ++k)
                GVGp[k] = ptr[k];
        for (size_t k = 0; k < adj.size(); ++k)
        {
            if (GVGe)
                GVGe[k] = adj[k] >> 2;
            if (GVGilocal)
                GVGilocal[k] = adj[k] & 3u;
        }
    });
}

vbdx_status vbdx_get_element_data(vbdx_integrator* h, double* GP, double* wg, double* m)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    return Guard([&] {
        auto& I = h->impl;
        VBDX_CUDA(cudaSetDevice(I.device));
        if (GP)
        {
            std::vector<double> ji(9 * I.nT);
            I.dJinv.Download(ji.data(), ji.size(), I.stream);
            VBDX_CUDA(cudaStreamSynchronize(I.stream));
            // Data::GP layout: 4 x 3nT column-major, GP[12e + 4d + i] (sim/vbd/Data.h:191-192)
            for (int64_t e = 0; e < I.nT; ++e)
                for (int d = 0; d < 3; ++d)
                {
                    double const* J = &ji[9 * e];
                    GP[12 * e + 4 * d + 0] = -(J[d] + J[3 + d] + J[6 + d]);
                    for (int i = 1; i < 4; ++i)
                        GP[12 * e + 4 * d + i] = J[3 * (i - 1) + d];
                }
        }
        if (wg)
            I.dVol.Download(wg, I.nT, I.stream);
        if (m)
            I.dMass.Download(m, I.nV, I.stream);
        VBDX_CUDA(cudaStreamSynchronize(I.stream));
    });
}

vbdx_status vbdx_get_colors(vbdx_integrator* h, int64_t* colors)
{
    if (vbdx_status s = NeedHandle(h))
        return s;
    if (colors)
        std::memcpy(colors, h->impl.colors.data(), h->impl.colors.size() * sizeof(int64_t));
    return VBDX_OK;
}

vbdx_status vbdx_greedy_color(int64_t nV, int64_t nT, const int64_t* E, int32_t ordering, int32_t selection, int64_t* colors_out)
{
    return Guard([&] {
        vbdx::Require(nV > 0 && nT >= 0 && E && colors_out, "vbdx_greedy_color: bad arguments");
        for (int64_t k = 0; k < 4 * nT; ++k)
            vbdx::Require(E[k] >= 0 && E[k] < nV, "element index out of range");
        std::vector<int64_t> colors;
        vbdx::GreedyColorMesh(nV, nT, E, ordering, selection, colors);
        std::memcpy(colors_out, colors.data(), colors.size() * sizeof(int64_t));
    });
}

}  // extern "C"

// ============================================================================================
// greedy colouring on the device (coloring.cuh)
// ============================================================================================
#include "coloring.cuh"

extern "C" vbdx_status vbdx_greedy_color_device(int64_t nV, int64_t nT, const int64_t* E, int32_t ordering, int32_t selection, int32_t device,
                                                int64_t* colors_out, int32_t* rounds_out)
{
    using namespace vbdx;
    return Guard([&] {
        Require(nV > 0 && nT >= 0 && E && colors_out, "vbdx_greedy_color_device: bad arguments");
        Require(ordering >= 0 && ordering <= 2, "unknown ordering strategy");
        if (selection != 1)
            throw Error(VBDX_UNSUPPORTED,
                        "device colouring reproduces the sequential result for the FirstAvailable selection only: LeastUsed picks by a global running "
                        "count of vertices per colour and is inherently sequential (use vbdx_greedy_color)");
        Require(nV < (int64_t(1) << 31) && 4 * nT < (int64_t(1) << 32), "mesh too large");
        std::vector<int32_t> E32(4 * static_cast<size_t>(nT));
        for (int64_t k = 0; k < 4 * nT; ++k)
        {
            Require(E[k] >= 0 && E[k] < nV, "element index out of range");
            E32[k] = static_cast<int32_t>(E[k]);
        }
        VBDX_CUDA(cudaSetDevice(device));
        cudaStream_t s = nullptr;
        DevBuf<int32_t> dE, dColors;
        DevBuf<uint32_t> dPtr, dCursor, dAdj, dDeg, dScratch, dSmall;
        int64_t const nEntries = 4 * nT;
        dE.Alloc(std::max<size_t>(E32.size(), 1)), dColors.Alloc(nV), dPtr.Alloc(nV + 1), dCursor.Alloc(nV + 1), dAdj.Alloc(std::max<int64_t>(nEntries, 1));
        dDeg.Alloc(nV), dScratch.Alloc((nV + 1) / kScanTile + 2), dSmall.Alloc(4);  // [0] max degree, [1] overflow, [2] coloured, [3] too many colours
        if (nEntries)
            dE.Upload(E32.data(), E32.size(), s);
        VBDX_CUDA(cudaMemsetAsync(dCursor.p, 0, (nV + 1) * sizeof(uint32_t), s));
        VBDX_CUDA(cudaMemsetAsync(dSmall.p, 0, 4 * sizeof(uint32_t), s));
        VBDX_CUDA(cudaMemsetAsync(dColors.p, 0xff, nV * sizeof(int32_t), s));
        if (nEntries)
            CountIncidences<<<Blocks(nEntries, 256), 256, 0, s>>>(dE.p, nEntries, dCursor.p);
        ExclusiveScanU32(dCursor.p, dPtr.p, nV + 1, dScratch.p, s);
        VBDX_CUDA(cudaMemsetAsync(dCursor.p, 0, (nV + 1) * sizeof(uint32_t), s));
        if (nEntries)
            FillIncidences<<<Blocks(nEntries, 256), 256, 0, s>>>(dE.p, nEntries, dPtr.p, dCursor.p, dAdj.p);
        ColorDegrees<<<Blocks(nV, 128), 128, 0, s>>>(nV, dE.p, dPtr.p, dAdj.p, dDeg.p, dSmall.p, dSmall.p + 1);
        uint32_t small[4] = {0, 0, 0, 0};
        int rounds        = 0;
        for (;;)
        {
            for (int r = 0; r < 8; ++r, ++rounds)  // a few rounds per look at the counter
                ColorRound<<<Blocks(nV, 128), 128, 0, s>>>(nV, dE.p, dPtr.p, dAdj.p, dDeg.p, dSmall.p, ordering, dColors.p, dSmall.p + 2, dSmall.p + 3);
            VBDX_CUDA(cudaMemcpyAsync(small, dSmall.p, sizeof(small), cudaMemcpyDeviceToHost, s));
            VBDX_CUDA(cudaStreamSynchronize(s));
            if (small[1] != 0u)
                throw Error(VBDX_UNSUPPORTED, "a vertex has more distinct neighbours than the device colouring handles (use vbdx_greedy_color)");
            if (small[3] != 0u)
                throw Error(VBDX_UNSUPPORTED, "more colours than the device colouring handles (use vbdx_greedy_color)");
            if (static_cast<int64_t>(small[2]) >= nV)
                break;
            Require(rounds <= nV + 8, "internal error: device colouring does not progress");
        }
        std::vector<int32_t> c32(nV);
        dColors.Download(c32.data(), nV, s);
        VBDX_CUDA(cudaStreamSynchronize(s));
        for (int64_t i = 0; i < nV; ++i)
            colors_out[i] = c32[i];
        if (rounds_out)
            *rounds_out = rounds;
    });
}

// ============================================================================================
// XPBD (xpbd.cuh): host driver and C ABI
// ============================================================================================
#include "xpbd.cuh"

namespace vbdx {

__global__ void XpbdSetXyz(int64_t n, const double* src, float4* dst, int keepW)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        dst[i] = make_float4(static_cast<float>(src[3 * i]), static_cast<float>(src[3 * i + 1]), static_cast<float>(src[3 * i + 2]),
                             keepW ? dst[i].w : 0.f);
}
__global__ void XpbdGetXyz(int64_t n, const float4* src, double* dst)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        dst[3 * i] = src[i].x, dst[3 * i + 1] = src[i].y, dst[3 * i + 2] = src[i].z;
}
__global__ void XpbdSetW(int64_t n, const float* w, float4* dst)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        dst[i].w = w[i];
}

struct XpbdIntegrator {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t evBegin = nullptr, evEnd = nullptr;
    int64_t nV = 0, nT = 0, nCV = 0, nF = 0;
    int64_t deviceBytes = 0, kernelLaunches = 0;
    double lastStepMs = 0;
    int gridBlocks = 0, nPartitions = 0;
    double muS = 0.3, muD = 0.2;
    std::vector<int32_t> slotOfTet;  // caller tet id -> constraint slot (partition-major order)
    DevBuf<float4> dX, dXt, dVel, dAext, dRec, dXb;
    DevBuf<int4> dTetIds;
    DevBuf<float2> dBeta, dLambda;
    DevBuf<uint32_t> dItemBegin, dPartBegin;
    DevBuf<float> dMuV, dAlphaC, dBetaC, dLambdaC;
    DevBuf<unsigned int> dBarrier;
    DevBuf<double> dStaging;
    std::vector<float> alphaHost;  // 2 per slot (compliances can be replaced: SetCompliance)
    ContactState contact;

    ~XpbdIntegrator()
    {
        if (evBegin)
            cudaEventDestroy(evBegin);
        if (evEnd)
            cudaEventDestroy(evEnd);
        if (stream)
            cudaStreamDestroy(stream);
    }

    void Create(vbdx_xpbd_desc const& d);
    XpbdParams MakeParams(double sdt, int iterations, int substeps);
    void Step(double dt, int iterations, int substeps);
    void UploadRecords(std::vector<float4> const& rec) { dRec.Upload(rec.data(), rec.size(), stream); }
    std::vector<float4> recHost;
};

void XpbdIntegrator::Create(vbdx_xpbd_desc const& d)
{
    Require(d.abi_version == VBDX_ABI_VERSION && d.struct_size == sizeof(vbdx_xpbd_desc), "vbdx_xpbd_desc: ABI version / struct size mismatch");
    Require(d.nV > 0 && d.nT > 0 && d.X && d.T, "need a volume mesh: X (3 x nV) and T (4 x nT)");
    Require(d.nV < (int64_t(1) << 31) - 1 && d.nT < (int64_t(1) << 30), "mesh too large for 32-bit device indices");
    nV = d.nV, nT = d.nT;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    {
        cudaGetLastError();
        throw Error(VBDX_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
    }
    if (d.device >= 0)
    {
        Require(d.device < count, "device ordinal out of range");
        VBDX_CUDA(cudaSetDevice(d.device));
    }
    VBDX_CUDA(cudaGetDevice(&device));
    VBDX_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    VBDX_CUDA(cudaEventCreate(&evBegin));
    VBDX_CUDA(cudaEventCreate(&evEnd));
    for (int64_t k = 0; k < 4 * nT; ++k)
        Require(d.T[k] >= 0 && d.T[k] < nV, "element index out of range");
    // ---- partitions: plain (every constraint its own work item) or clustered (sim/xpbd/Data.h:95-103)
    Require(d.Pptr && d.Padj && d.nPartitions >= 1, "XPBD needs constraint partitions (Data::WithPartitions)");
    bool const clustered = d.SGptr != nullptr && d.nClusterPartitions > 0;
    if (clustered)
        Require(d.SGadj && d.Cptr && d.Cadj, "cluster partitions need SGptr, SGadj, Cptr and Cadj");
    std::vector<uint32_t> partBegin, itemBegin;
    std::vector<int64_t> order;  // constraint (tet) of every slot
    if (!clustered)
    {
        nPartitions = d.nPartitions;
        for (int32_t q = 0; q <= d.nPartitions; ++q)
            partBegin.push_back(static_cast<uint32_t>(d.Pptr[q]));
        for (int64_t k = 0; k < d.Pptr[d.nPartitions]; ++k)
        {
            itemBegin.push_back(static_cast<uint32_t>(order.size()));
            order.push_back(d.Padj[k]);
        }
    }
    else
    {
        nPartitions = d.nClusterPartitions;
        for (int32_t q = 0; q < d.nClusterPartitions; ++q)
        {
            partBegin.push_back(static_cast<uint32_t>(itemBegin.size()));
            for (int64_t k = d.SGptr[q]; k < d.SGptr[q + 1]; ++k)
            {
                int64_t const cl = d.SGadj[k];
                itemBegin.push_back(static_cast<uint32_t>(order.size()));
                for (int64_t j = d.Cptr[cl]; j < d.Cptr[cl + 1]; ++j)
                    order.push_back(d.Cadj[j]);
            }
        }
        partBegin.push_back(static_cast<uint32_t>(itemBegin.size()));
    }
    itemBegin.push_back(static_cast<uint32_t>(order.size()));
    Require(static_cast<int64_t>(order.size()) == nT, "the partitions must list every element exactly once");
    slotOfTet.assign(nT, -1);
    for (size_t sidx = 0; sidx < order.size(); ++sidx)
    {
        Require(order[sidx] >= 0 && order[sidx] < nT && slotOfTet[order[sidx]] < 0, "the partitions must list every element exactly once");
        slotOfTet[order[sidx]] = static_cast<int32_t>(sidx);
    }
    {
        // no two constraints of a partition that run concurrently may share a vertex
        std::vector<int32_t> owner(nV, -1), ownerPart(nV, -1);
        for (int32_t q = 0; q < nPartitions; ++q)
            for (uint32_t it = partBegin[q]; it < partBegin[q + 1]; ++it)
                for (uint32_t c = itemBegin[it]; c < itemBegin[it + 1]; ++c)
                    for (int a = 0; a < 4; ++a)
                    {
                        int64_t const vtx = d.T[4 * order[c] + a];
                        Require(!(ownerPart[vtx] == q && owner[vtx] != static_cast<int32_t>(it)),
                                "invalid partitioning: two constraints of one partition share a vertex");
                        ownerPart[vtx] = q, owner[vtx] = static_cast<int32_t>(it);
                    }
    }
    // ---- per-constraint data (sim/xpbd/Data.cpp:131-153): DmInv, compliance 1 / (lame * volume), gamma = 1 + mu / lambda
    double const Y = 1e6, nu = 0.45;
    double const muDefault = Y / (2. * (1. + nu)), lamDefault = (Y * nu) / ((1. + nu) * (1. - 2. * nu));
    std::vector<int4> ids(nT);
    recHost.assign(3 * static_cast<size_t>(nT), make_float4(0, 0, 0, 0));
    std::vector<float2> beta(nT, make_float2(0.f, 0.f));
    for (int64_t sidx = 0; sidx < nT; ++sidx)
    {
        int64_t const t = order[sidx];
        int64_t const* e = d.T + 4 * t;
        ids[sidx]        = make_int4(static_cast<int>(e[0]), static_cast<int>(e[1]), static_cast<int>(e[2]), static_cast<int>(e[3]));
        double Ds[3][3];
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r)
                Ds[r][c] = d.X[3 * e[c + 1] + r] - d.X[3 * e[0] + r];
        double const det = Ds[0][0] * (Ds[1][1] * Ds[2][2] - Ds[1][2] * Ds[2][1]) - Ds[0][1] * (Ds[1][0] * Ds[2][2] - Ds[1][2] * Ds[2][0]) +
                           Ds[0][2] * (Ds[1][0] * Ds[2][1] - Ds[1][1] * Ds[2][0]);
        Require(det > 1e-300, "inverted or degenerate tetrahedron in the rest mesh");
        double inv[3][3];
        inv[0][0] = (Ds[1][1] * Ds[2][2] - Ds[1][2] * Ds[2][1]) / det, inv[0][1] = (Ds[0][2] * Ds[2][1] - Ds[0][1] * Ds[2][2]) / det;
        inv[0][2] = (Ds[0][1] * Ds[1][2] - Ds[0][2] * Ds[1][1]) / det, inv[1][0] = (Ds[1][2] * Ds[2][0] - Ds[1][0] * Ds[2][2]) / det;
        inv[1][1] = (Ds[0][0] * Ds[2][2] - Ds[0][2] * Ds[2][0]) / det, inv[1][2] = (Ds[0][2] * Ds[1][0] - Ds[0][0] * Ds[1][2]) / det;
        inv[2][0] = (Ds[1][0] * Ds[2][1] - Ds[1][1] * Ds[2][0]) / det, inv[2][1] = (Ds[0][1] * Ds[2][0] - Ds[0][0] * Ds[2][1]) / det;
        inv[2][2] = (Ds[0][0] * Ds[1][1] - Ds[0][1] * Ds[1][0]) / det;
        double const mu = d.lame ? d.lame[2 * t] : muDefault, lam = d.lame ? d.lame[2 * t + 1] : lamDefault, vol = det / 6.0;
        double const aD = d.alphaSNH ? d.alphaSNH[2 * t] : 1.0 / (mu * vol), aH = d.alphaSNH ? d.alphaSNH[2 * t + 1] : 1.0 / (lam * vol);
        // record = DmInv column c in xyz, w = gamma / alpha_D / alpha_H
        for (int c = 0; c < 3; ++c)
            recHost[3 * sidx + c] = make_float4(static_cast<float>(inv[0][c]), static_cast<float>(inv[1][c]), static_cast<float>(inv[2][c]),
                                                static_cast<float>(c == 0 ? 1.0 + mu / lam : c == 1 ? aD : aH));
        if (d.betaSNH)
            beta[sidx] = make_float2(static_cast<float>(d.betaSNH[2 * t]), static_cast<float>(d.betaSNH[2 * t + 1]));
    }
    // ---- particles (sim/xpbd/Data.cpp:104-126): minv default 1e-3, Dirichlet vertices get minv = 0 and v = a = 0
    std::vector<float4> x(nV), v(nV, make_float4(0, 0, 0, 0)), a(nV, make_float4(0.f, 0.f, -9.81f, 0.f));
    for (int64_t i = 0; i < nV; ++i)
    {
        x[i] = make_float4(static_cast<float>(d.X[3 * i]), static_cast<float>(d.X[3 * i + 1]), static_cast<float>(d.X[3 * i + 2]),
                           static_cast<float>(d.minv ? d.minv[i] : 1e-3));
        if (d.v)
            v[i] = make_float4(static_cast<float>(d.v[3 * i]), static_cast<float>(d.v[3 * i + 1]), static_cast<float>(d.v[3 * i + 2]), 0.f);
        if (d.aext)
            a[i] = make_float4(static_cast<float>(d.aext[3 * i]), static_cast<float>(d.aext[3 * i + 1]), static_cast<float>(d.aext[3 * i + 2]), 0.f);
    }
    for (int64_t k = 0; k < d.nDbc; ++k)
    {
        Require(d.dbc && d.dbc[k] >= 0 && d.dbc[k] < nV, "Dirichlet vertex index out of range");
        x[d.dbc[k]].w = 0.f;
        v[d.dbc[k]] = a[d.dbc[k]] = make_float4(0, 0, 0, 0);
    }
    dX.Alloc(nV, &deviceBytes), dXt.Alloc(nV, &deviceBytes), dVel.Alloc(nV, &deviceBytes), dAext.Alloc(nV, &deviceBytes);
    dX.Upload(x.data(), nV, stream), dXt.Upload(x.data(), nV, stream), dVel.Upload(v.data(), nV, stream), dAext.Upload(a.data(), nV, stream);
    dTetIds.Alloc(nT, &deviceBytes), dRec.Alloc(3 * static_cast<size_t>(nT), &deviceBytes), dBeta.Alloc(nT, &deviceBytes), dLambda.Alloc(nT, &deviceBytes);
    dTetIds.Upload(ids.data(), nT, stream), dRec.Upload(recHost.data(), recHost.size(), stream), dBeta.Upload(beta.data(), nT, stream);
    VBDX_CUDA(cudaMemsetAsync(dLambda.p, 0, nT * sizeof(float2), stream));
    dItemBegin.Alloc(itemBegin.size(), &deviceBytes), dPartBegin.Alloc(partBegin.size(), &deviceBytes);
    dItemBegin.Upload(itemBegin.data(), itemBegin.size(), stream), dPartBegin.Upload(partBegin.data(), partBegin.size(), stream);
    dBarrier.Alloc(1, &deviceBytes);
    dStaging.Alloc(3 * nV, &deviceBytes);
    VBDX_CUDA(cudaStreamSynchronize(stream));
    muS = d.muS, muD = d.muD;
    // ---- contact
    if (d.nF > 0 && d.nCV > 0)
    {
        Require(d.F != nullptr && d.V != nullptr, "collision mesh pointers missing");
        Require(d.active_set_update_frequency >= 1, "active set update frequency must be >= 1");
        nCV = d.nCV, nF = d.nF;
        ContactState& cs = contact;
        cs.nCV = static_cast<uint32_t>(nCV), cs.nF = static_cast<uint32_t>(nF);
        cs.updateFrequency = d.active_set_update_frequency;
        std::vector<int32_t> Bh(nV), Vh(nCV);
        std::vector<int4> Fh(nF);
        for (int64_t i = 0; i < nV; ++i)
            Bh[i] = d.BV ? static_cast<int32_t>(d.BV[i]) : 0;  // sim/xpbd/Data.cpp:121-124: one body by default
        for (int64_t k = 0; k < nCV; ++k)
        {
            Require(d.V[k] >= 0 && d.V[k] < nV, "collision vertex index out of range");
            Vh[k] = static_cast<int32_t>(d.V[k]);
        }
        for (int64_t f = 0; f < nF; ++f)
        {
            for (int k = 0; k < 3; ++k)
                Require(d.F[3 * f + k] >= 0 && d.F[3 * f + k] < nV, "collision triangle index out of range");
            Fh[f] = make_int4(static_cast<int>(d.F[3 * f]), static_cast<int>(d.F[3 * f + 1]), static_cast<int>(d.F[3 * f + 2]), Bh[d.F[3 * f]]);
        }
        cs.B.Alloc(nV, &deviceBytes), cs.V.Alloc(nCV, &deviceBytes), cs.F.Alloc(nF, &deviceBytes);
        cs.B.Upload(Bh.data(), nV, stream), cs.V.Upload(Vh.data(), nCV, stream), cs.F.Upload(Fh.data(), nF, stream);
        cs.mesh = ContactMesh{cs.B.p, cs.V.p, cs.F.p, cs.nCV, cs.nF};
        cs.Alloc(nV, &deviceBytes, stream);
        cs.enabled = true;
        std::vector<float> muV(nCV, 1.f), aC(nCV, 0.f), bC(nCV, 0.f);
        for (int64_t k = 0; k < nCV; ++k)
        {
            if (d.muV)
                muV[k] = static_cast<float>(d.muV[k]);
            if (d.alphaC)
                aC[k] = static_cast<float>(d.alphaC[k]);
            if (d.betaC)
                bC[k] = static_cast<float>(d.betaC[k]);
        }
        dMuV.Alloc(nCV, &deviceBytes), dAlphaC.Alloc(nCV, &deviceBytes), dBetaC.Alloc(nCV, &deviceBytes), dLambdaC.Alloc(nCV, &deviceBytes);
        dXb.Alloc(nCV, &deviceBytes);
        dMuV.Upload(muV.data(), nCV, stream), dAlphaC.Upload(aC.data(), nCV, stream), dBetaC.Upload(bC.data(), nCV, stream);
        VBDX_CUDA(cudaMemsetAsync(dLambdaC.p, 0, nCV * sizeof(float), stream));
        VBDX_CUDA(cudaStreamSynchronize(stream));
    }
    int perSm = 0;
    VBDX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, XpbdSolveKernel, 256, 0));
    cudaDeviceProp prop;
    VBDX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (perSm < 1 || !prop.cooperativeLaunch)
        throw Error(VBDX_CUDA_ERROR, "the XPBD solve kernel does not fit on an SM");
    // no more CTAs than the largest partition can feed
    uint32_t maxItems = 1;
    for (int32_t q = 0; q < nPartitions; ++q)
        maxItems = std::max(maxItems, partBegin[q + 1] - partBegin[q]);
    int64_t const want = std::max<int64_t>(Blocks(std::max<int64_t>(maxItems, nCV), 256), 1);
    gridBlocks         = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(perSm) * prop.multiProcessorCount, std::max<int64_t>(want, prop.multiProcessorCount)));
}

XpbdParams XpbdIntegrator::MakeParams(double sdt, int iterations, int substeps)
{
    XpbdParams p{};
    p.x = dX.p, p.xt = dXt.p, p.vel = dVel.p, p.aext = dAext.p;
    p.tetIds = dTetIds.p, p.rec = dRec.p, p.beta = dBeta.p, p.lambda = dLambda.p;
    p.itemBegin = dItemBegin.p, p.partBegin = dPartBegin.p;
    p.nPartitions = nPartitions, p.nV = static_cast<int>(nV), p.nT = static_cast<int>(nT);
    p.sdt = static_cast<float>(sdt), p.sdt2 = static_cast<float>(sdt * sdt);
    p.iterations = iterations, p.substeps = substeps;
    p.barrier = dBarrier.p;
    if (contact.enabled)
    {
        p.nCV = static_cast<int>(nCV);
        p.nActive = contact.nActive.p, p.av = contact.av.p, p.V = contact.V.p, p.triF = contact.F.p, p.nn = contact.nn.p;
        p.xb = dXb.p, p.muV = dMuV.p, p.alphaC = dAlphaC.p, p.betaC = dBetaC.p, p.lambdaC = dLambdaC.p;
        p.muS = static_cast<float>(muS), p.muD = static_cast<float>(muD);
    }
    return p;
}

void XpbdIntegrator::Step(double dt, int iterations, int substeps)
{
    Require(dt > 0 && iterations >= 0 && substeps >= 1, "Step: need dt > 0, iterations >= 0, substeps >= 1");
    NvtxRange const zone("pbat.gpu.impl.xpbd.Integrator.Step");
    VBDX_CUDA(cudaSetDevice(device));
    double const sdt = dt / substeps;
    VBDX_CUDA(cudaEventRecord(evBegin, stream));
    auto solve = [&](XpbdParams& p) {
        VBDX_CUDA(cudaMemsetAsync(dBarrier.p, 0, sizeof(unsigned int), stream));
        void* args[] = {&p};
        VBDX_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void const*>(XpbdSolveKernel), dim3(gridBlocks), dim3(256), args, 0, stream));
        ++kernelLaunches;
    };
    if (!contact.enabled)
    {
        XpbdParams p = MakeParams(sdt, iterations, substeps);
        solve(p);
    }
    else
    {
        // active set from the full-step predictor x + dt v + dt^2 a (gpu/impl/xpbd/Integrator.cu:96-117)
        contact.InitializeActiveSet(dX.p, dVel.p, dAext.p, nV, static_cast<float>(dt), stream, &kernelLaunches);
        XpbdParams p  = MakeParams(sdt, iterations, 1);
        p.skipPreStep = 1;
        int const n   = static_cast<int>(std::max(std::max(nV, nT), nCV));
        for (int s = 0; s < substeps; ++s)
        {
            XpbdPreStepKernel<<<Blocks(n, 256), 256, 0, stream>>>(p);
            ++kernelLaunches;
            if (s % contact.updateFrequency == 0)
                contact.NearestPass(dX.p, 0, stream, &kernelLaunches);
            solve(p);
        }
        contact.NearestPass(dX.p, 1, stream, &kernelLaunches);
    }
    VBDX_CUDA(cudaEventRecord(evEnd, stream));
    VBDX_CUDA(cudaStreamSynchronize(stream));
    float ms = 0;
    VBDX_CUDA(cudaEventElapsedTime(&ms, evBegin, evEnd));
    lastStepMs = ms;
}

}  // namespace vbdx

struct vbdx_xpbd {
    vbdx::XpbdIntegrator impl;
};

extern "C" {

void vbdx_xpbd_desc_init(vbdx_xpbd_desc* d)
{
    std::memset(d, 0, sizeof(*d));
    d->abi_version = VBDX_ABI_VERSION;
    d->struct_size = sizeof(vbdx_xpbd_desc);
    d->muS = 0.3, d->muD = 0.2;  // sim/xpbd/Data.h:85-86
    d->active_set_update_frequency = 1;
    d->device = -1;
}

vbdx_status vbdx_xpbd_create(const vbdx_xpbd_desc* desc, vbdx_xpbd** out)
{
    if (!desc || !out)
    {
        gLastError = "vbdx_xpbd_create: null argument";
        return VBDX_INVALID_ARGUMENT;
    }
    *out = nullptr;
    std::unique_ptr<vbdx_xpbd> h;
    vbdx_status const st = Guard([&] {
        h = std::make_unique<vbdx_xpbd>();
        h->impl.Create(*desc);
    });
    if (st == VBDX_OK)
        *out = h.release();
    return st;
}

vbdx_status vbdx_xpbd_destroy(vbdx_xpbd* h)
{
    if (h)
    {
        cudaSetDevice(h->impl.device);
        cudaStreamSynchronize(h->impl.stream);
        delete h;
    }
    return VBDX_OK;
}

vbdx_status vbdx_xpbd_step(vbdx_xpbd* h, double dt, int32_t iterations, int32_t substeps)
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] { h->impl.Step(dt, iterations, substeps); });
}

static vbdx_status XpbdSet(vbdx_xpbd* h, const double* a, int64_t nV, float4* dst, int keepW)
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] {
        auto& I = h->impl;
        vbdx::Require(a != nullptr && nV == I.nV, "expected a 3 x nV array");
        VBDX_CUDA(cudaSetDevice(I.device));
        VBDX_CUDA(cudaMemcpyAsync(I.dStaging.p, a, 3 * nV * sizeof(double), cudaMemcpyHostToDevice, I.stream));
        vbdx::XpbdSetXyz<<<vbdx::Blocks(nV, 256), 256, 0, I.stream>>>(nV, I.dStaging.p, dst, keepW);
        ++I.kernelLaunches;
        VBDX_CUDA(cudaStreamSynchronize(I.stream));
    });
}
static vbdx_status XpbdGet(vbdx_xpbd* h, double* a, int64_t nV, const float4* src)
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] {
        auto& I = h->impl;
        vbdx::Require(a != nullptr && nV == I.nV, "expected a 3 x nV array");
        VBDX_CUDA(cudaSetDevice(I.device));
        vbdx::XpbdGetXyz<<<vbdx::Blocks(nV, 256), 256, 0, I.stream>>>(nV, src, I.dStaging.p);
        ++I.kernelLaunches;
        VBDX_CUDA(cudaMemcpyAsync(a, I.dStaging.p, 3 * nV * sizeof(double), cudaMemcpyDeviceToHost, I.stream));
        VBDX_CUDA(cudaStreamSynchronize(I.stream));
    });
}
vbdx_status vbdx_xpbd_set_positions(vbdx_xpbd* h, const double* x, int64_t nV) { return XpbdSet(h, x, nV, h ? h->impl.dX.p : nullptr, 1); }
vbdx_status vbdx_xpbd_set_velocities(vbdx_xpbd* h, const double* v, int64_t nV) { return XpbdSet(h, v, nV, h ? h->impl.dVel.p : nullptr, 0); }
vbdx_status vbdx_xpbd_set_external_acceleration(vbdx_xpbd* h, const double* a, int64_t nV) { return XpbdSet(h, a, nV, h ? h->impl.dAext.p : nullptr, 0); }
vbdx_status vbdx_xpbd_get_positions(vbdx_xpbd* h, double* x, int64_t nV) { return XpbdGet(h, x, nV, h ? h->impl.dX.p : nullptr); }
vbdx_status vbdx_xpbd_get_velocities(vbdx_xpbd* h, double* v, int64_t nV) { return XpbdGet(h, v, nV, h ? h->impl.dVel.p : nullptr); }

vbdx_status vbdx_xpbd_set_compliance(vbdx_xpbd* h, int32_t constraint, const double* alpha, int64_t n)
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] {
        auto& I = h->impl;
        VBDX_CUDA(cudaSetDevice(I.device));
        if (constraint == VBDX_XPBD_STABLE_NEO_HOOKEAN)
        {
            vbdx::Require(alpha != nullptr && n == 2 * I.nT, "expected 2 compliances per element");
            for (int64_t t = 0; t < I.nT; ++t)
            {
                int64_t const s = I.slotOfTet[t];
                I.recHost[3 * s + 1].w = static_cast<float>(alpha[2 * t]);
                I.recHost[3 * s + 2].w = static_cast<float>(alpha[2 * t + 1]);
            }
            I.dRec.Upload(I.recHost.data(), I.recHost.size(), I.stream);
        }
        else if (constraint == VBDX_XPBD_COLLISION)
        {
            vbdx::Require(alpha != nullptr && n == I.nCV && I.nCV > 0, "expected one compliance per collision vertex");
            std::vector<float> a(alpha, alpha + n);
            I.dAlphaC.Upload(a.data(), n, I.stream);
        }
        else
            throw vbdx::Error(VBDX_INVALID_ARGUMENT, "unknown constraint type");
        VBDX_CUDA(cudaStreamSynchronize(I.stream));
    });
}

vbdx_status vbdx_xpbd_set_friction_coefficients(vbdx_xpbd* h, double muS, double muD)
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    h->impl.muS = muS, h->impl.muD = muD;
    return VBDX_OK;
}

vbdx_status vbdx_xpbd_set_scene_bounding_box(vbdx_xpbd* h, const float min3[3], const float max3[3])
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] {
        vbdx::Require(min3 && max3 && max3[0] > min3[0] && max3[1] > min3[1] && max3[2] > min3[2], "scene bounding box must have positive extent");
        if (h->impl.contact.enabled)
        {
            VBDX_CUDA(cudaSetDevice(h->impl.device));
            h->impl.contact.SetWorldBox(min3, max3, h->impl.stream);
        }
    });
}

vbdx_status vbdx_xpbd_get_info(vbdx_xpbd* h, int64_t* out8)
{
    if (!h || !out8)
        return VBDX_INVALID_ARGUMENT;
    auto const& I = h->impl;
    out8[0] = I.nV, out8[1] = I.nT, out8[2] = I.nPartitions, out8[3] = I.gridBlocks, out8[4] = I.kernelLaunches, out8[5] = I.deviceBytes;
    out8[6] = static_cast<int64_t>(I.lastStepMs * 1e6), out8[7] = I.nCV;
    return VBDX_OK;
}

vbdx_status vbdx_xpbd_get_contact_state(vbdx_xpbd* h, int32_t* active, int32_t* nn, int64_t* nActive)
{
    if (!h)
        return VBDX_INVALID_ARGUMENT;
    return Guard([&] {
        auto& I  = h->impl;
        auto& cs = I.contact;
        vbdx::Require(cs.enabled, "the integrator was created without a collision mesh");
        VBDX_CUDA(cudaSetDevice(I.device));
        std::vector<uint8_t> a(cs.nCV);
        cs.active.Download(a.data(), cs.nCV, I.stream);
        if (nn)
            cs.nn.Download(nn, static_cast<size_t>(cs.nCV) * vbdx::kMaxContacts, I.stream);
        uint32_t na = 0;
        cs.nActive.Download(&na, 1, I.stream);
        VBDX_CUDA(cudaStreamSynchronize(I.stream));
        if (active)
            for (uint32_t k = 0; k < cs.nCV; ++k)
                active[k] = a[k];
        if (nActive)
            *nActive = na;
    });
}

vbdx_status vbdx_graph_greedy_color(int64_t n, const int64_t* ptr, const int64_t* adj, int32_t ordering, int32_t selection, int64_t* colors_out)
{
    return Guard([&] {
        vbdx::Require(n >= 0 && ptr && colors_out && (adj || ptr[n] == 0), "vbdx_graph_greedy_color: bad arguments");
        std::vector<int32_t> adj32(static_cast<size_t>(ptr[n]));
        for (int64_t k = 0; k < ptr[n]; ++k)
        {
            vbdx::Require(adj[k] >= 0 && adj[k] < n, "adjacency entry out of range");
            adj32[k] = static_cast<int32_t>(adj[k]);
        }
        std::vector<int64_t> colors;
        vbdx::GreedyColorGraph(n, ptr, adj32.data(), ordering, selection, colors);
        std::memcpy(colors_out, colors.data(), colors.size() * sizeof(int64_t));
    });
}

}  // extern "C"
