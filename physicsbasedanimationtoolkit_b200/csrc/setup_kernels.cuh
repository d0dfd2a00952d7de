// Device-side problem construction for the B200 VBD integrator: the vertex->tet CSR
// (count -> scan -> fill -> per-row sort), element rest data in double precision, lumped masses,
// and the packed incidence records the sweep streams.  Replaces, on the device, the parts of
// pbat::sim::vbd::Data::Construct that the reference does on the host with Eigen
// (sim/vbd/Data.cpp:210-226: fem::ShapeFunctionGradients, MeshQuadratureWeights, lumped mass,
// graph::MeshAdjacencyMatrix + transpose).
#pragma once

#include "vbdx_internal.h"

#include <cuda_runtime.h>

namespace vbdx {

// ------------------------------------------------------------------------------------------
// exclusive scan of uint32 (three small kernels; setup-only, not a hot path)
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems   = 8;
constexpr int kScanTile    = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t BlockExclusiveScan(uint32_t v, uint32_t* smem /* >= 33 */, uint32_t& total)
{
    uint32_t const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t const n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= static_cast<uint32_t>(o))
            inc += n;
    }
    if (lane == 31)
        smem[warp] = inc;
    __syncthreads();
    if (warp == 0)
    {
        uint32_t const nW = blockDim.x >> 5;
        uint32_t w        = lane < nW ? smem[lane] : 0u;
        uint32_t winc     = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t const n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= static_cast<uint32_t>(o))
                winc += n;
        }
        if (lane < nW)
            smem[lane] = winc - w;
        if (lane == 31)
            smem[32] = winc;
    }
    __syncthreads();
    total               = smem[32];
    uint32_t const excl = smem[warp] + inc - v;
    __syncthreads();
    return excl;
}

__global__ void ScanTileSums(const uint32_t* in, uint32_t* tileSums, int64_t n)
{
    __shared__ uint32_t smem[33];
    int64_t const base = static_cast<int64_t>(blockIdx.x) * kScanTile + threadIdx.x * kScanItems;
    uint32_t s         = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n)
            s += in[base + k];
    uint32_t total;
    BlockExclusiveScan(s, smem, total);
    if (threadIdx.x == 0)
        tileSums[blockIdx.x] = total;
}

__global__ void ScanSingleBlock(uint32_t* data, int64_t n)
{
    __shared__ uint32_t smem[33];
    uint32_t carry = 0;
    for (int64_t base = 0; base < n; base += blockDim.x)
    {
        int64_t const i  = base + threadIdx.x;
        uint32_t const v = i < n ? data[i] : 0u;
        uint32_t total;
        uint32_t const e = BlockExclusiveScan(v, smem, total);
        if (i < n)
            data[i] = carry + e;
        carry += total;
    }
}

__global__ void ScanApply(const uint32_t* in, const uint32_t* tileOffsets, uint32_t* out, int64_t n)
{
    __shared__ uint32_t smem[33];
    int64_t const base = static_cast<int64_t>(blockIdx.x) * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
    {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        s += v[k];
    }
    uint32_t total;
    uint32_t run = BlockExclusiveScan(s, smem, total) + tileOffsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
    {
        if (base + k < n)
            out[base + k] = run;
        run += v[k];
    }
}

// out[i] = sum_{j<i} in[j] for i in [0,n); in and out may alias.  scratch >= ceil(n/kScanTile) uint32.
inline void ExclusiveScanU32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* scratch, cudaStream_t s)
{
    if (n <= 0)
        return;
    int const tiles = static_cast<int>((n + kScanTile - 1) / kScanTile);
    ScanTileSums<<<tiles, kScanThreads, 0, s>>>(in, scratch, n);
    ScanSingleBlock<<<1, 1024, 0, s>>>(scratch, tiles);
    ScanApply<<<tiles, kScanThreads, 0, s>>>(in, scratch, out, n);
}

// ------------------------------------------------------------------------------------------
// vertex -> tet CSR
// ------------------------------------------------------------------------------------------
__global__ void CountIncidences(const int32_t* E, int64_t nEntries, uint32_t* deg)
{
    int64_t const k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k < nEntries)
        atomicAdd(&deg[E[k]], 1u);
}

__global__ void FillIncidences(const int32_t* E, int64_t nEntries, const uint32_t* ptr, uint32_t* cursor, uint32_t* adj)
{
    int64_t const k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k < nEntries)
    {
        int32_t const v     = E[k];
        uint32_t const slot = ptr[v] + atomicAdd(&cursor[v], 1u);
        adj[slot]           = static_cast<uint32_t>(k);  // k = 4*e + ilocal: tet id and local index packed
    }
}

// Rows are short (tens of entries): one thread per vertex, insertion sort => ascending element id,
// the order the reference's transposed adjacency has (sim/vbd/Data.cpp:223-226).
__global__ void SortRows(const uint32_t* ptr, uint32_t* adj, int64_t nV)
{
    int64_t const v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= nV)
        return;
    uint32_t const b = ptr[v], e = ptr[v + 1];
    for (uint32_t i = b + 1; i < e; ++i)
    {
        uint32_t const key = adj[i];
        uint32_t j         = i;
        while (j > b && adj[j - 1] > key)
        {
            adj[j] = adj[j - 1];
            --j;
        }
        adj[j] = key;
    }
}

// ------------------------------------------------------------------------------------------
// element rest data (double): Jinv rows = shape-function gradients of local vertices 1..3
// (fem/ShapeFunctions.h:267-297 with GN of fem/Tetrahedron.h:75-97), vol = det(J)/6
// (fem/MeshQuadrature.h:75-88).  errFlag bit 0: inverted/degenerate tet (fem/Jacobian.h:68-80);
// bit 1: two swept vertices of one tet share a colour.
// ------------------------------------------------------------------------------------------
__global__ void ElementQuantities(
    const double* X,
    const int32_t* E,
    int64_t nT,
    const int32_t* color,
    const uint8_t* isDbc,
    double* Jinv,
    double* vol,
    uint32_t* errFlag)
{
    int64_t const e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= nT)
        return;
    int32_t t[4];
    for (int a = 0; a < 4; ++a)
        t[a] = E[4 * e + a];
    double J[3][3];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r)
            J[r][c] = X[3 * static_cast<int64_t>(t[c + 1]) + r] - X[3 * static_cast<int64_t>(t[0]) + r];
    double const c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    double const c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    double const c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    double const det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    if (!(det > 1e-10))
        atomicOr(errFlag, 1u);
    double const r = 1.0 / det;
    double* o      = Jinv + 9 * e;
    o[0] = c00 * r;
    o[1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
    o[2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
    o[3] = c01 * r;
    o[4] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
    o[5] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
    o[6] = c02 * r;
    o[7] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
    o[8] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
    vol[e] = det / 6.0;
    for (int a = 0; a < 4; ++a)
        for (int b = a + 1; b < 4; ++b)
            if (color[t[a]] == color[t[b]] && !isDbc[t[a]] && !isDbc[t[b]])
                atomicOr(errFlag, 2u);
}

// lumped mass m_i = sum_e rho_e V_e / 4 over the (sorted) CSR row: deterministic order
__global__ void VertexMass(const uint32_t* ptr, const uint32_t* adj, const double* vol, const double* rhoe, double rhoDefault, double* m, int64_t nV)
{
    int64_t const v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= nV)
        return;
    double s = 0;
    for (uint32_t k = ptr[v]; k < ptr[v + 1]; ++k)
    {
        uint32_t const e = adj[k] >> 2;
        s += (rhoe ? rhoe[e] : rhoDefault) * vol[e] / 4.0;
    }
    m[v] = s;
}

// ------------------------------------------------------------------------------------------
// incidence records: one warp per tile (layout: vbdx_internal.h)
// ------------------------------------------------------------------------------------------
__global__ void FillRecords(
    const TileDesc* tiles,
    int nTiles,
    const int32_t* new2old,
    const int32_t* old2new,
    const uint32_t* ptr,
    const uint32_t* adj,
    const int32_t* E,
    const double* Jinv,
    const double* vol,
    const double* lame,  // 2 x nT or null
    double muDefault,
    double lamDefault,
    const uint32_t* recIdx,  // packed local ring indices per record slot (host planner)
    int stvk,                // 0: Stable Neo-Hookean records, 1: St. Venant-Kirchhoff (two blocks per incident tet)
    float4* records)
{
    int const T = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (T >= nTiles)
        return;
    uint32_t const lane   = threadIdx.x & 31u;
    TileDesc const td     = tiles[T];
    uint32_t const lw     = TileLog2W(td.meta);
    uint32_t const nverts = TileVerts(td.meta);
    uint32_t const bpi    = stvk ? 2u : 1u;
    uint32_t const iters  = TileIters(td.meta) / bpi;
    uint32_t const w      = 1u << lw;
    uint32_t const grp    = lane >> lw;
    uint32_t const sub    = lane & (w - 1u);
    bool const valid      = grp < nverts;
    uint32_t const vi     = td.vbase + (valid ? grp : 0u);
    int32_t const vo      = new2old[vi];
    uint32_t const rowB = ptr[vo], degv = ptr[vo + 1] - rowB;
    for (uint32_t t = 0; t < iters; ++t)
    {
        uint32_t const k = t * w + sub;
        uint32_t const idx = recIdx[static_cast<size_t>(td.blockStart + t * bpi) * 32 + lane];
        float rec[6]     = {0, 0, 0, 0, 0, 0};
        float sv[11]     = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // StVK: grad N_a, grad N_b, grad N_c, wg mu, wg lambda
        if (valid && k < degv)
        {
            uint32_t const packed = adj[rowB + k];
            int64_t const e       = packed >> 2;
            int const il          = packed & 3u;
            double const* Ji      = Jinv + 9 * e;
            double g[4][3];  // shape-function gradients of the four local vertices
            for (int d = 0; d < 3; ++d)
            {
                g[0][d] = -(Ji[d] + Ji[3 + d] + Ji[6 + d]);
                for (int a = 1; a < 4; ++a)
                    g[a][d] = Ji[3 * (a - 1) + d];
            }
            double const* q = g[il];
            double u[3];
            double const* o[3];
            int n = 0;
            for (int a = 0; a < 4; ++a)
            {
                if (a == il)
                    continue;
                o[n] = g[a];
                u[n] = g[a][0] * q[0] + g[a][1] * q[1] + g[a][2] * q[2];
                ++n;
            }
            double const detG = o[0][0] * (o[1][1] * o[2][2] - o[1][2] * o[2][1]) -
                                o[0][1] * (o[1][0] * o[2][2] - o[1][2] * o[2][0]) +
                                o[0][2] * (o[1][0] * o[2][1] - o[1][1] * o[2][0]);
            double const mu = lame ? lame[2 * e] : muDefault, lam = lame ? lame[2 * e + 1] : lamDefault;
            double const wmu = vol[e] * mu, wlam = vol[e] * lam, alpha = 1.0 + mu / lam;
            rec[0] = static_cast<float>(wmu * (u[0] + u[1] + u[2]));
            rec[1] = static_cast<float>(wmu * u[1]);
            rec[2] = static_cast<float>(wmu * u[2]);
            rec[3] = static_cast<float>(wlam * detG * detG);
            rec[4] = static_cast<float>(wlam * detG * alpha);
            rec[5] = static_cast<float>(wmu * (q[0] * q[0] + q[1] * q[1] + q[2] * q[2]));
            for (int a = 0; a < 3; ++a)
                for (int d = 0; d < 3; ++d)
                    sv[3 * a + d] = static_cast<float>(o[a][d]);
            sv[9] = static_cast<float>(wmu), sv[10] = static_cast<float>(wlam);
        }
        float4* out = records + static_cast<size_t>(td.blockStart + t * bpi) * kBlockFloat4 + lane;
        if (!stvk)
        {
            out[0]  = make_float4(__uint_as_float(idx), rec[0], rec[1], rec[2]);
            out[32] = make_float4(rec[3], rec[4], rec[5], 0.f);
        }
        else
        {
            out[0]                 = make_float4(__uint_as_float(idx), sv[0], sv[1], sv[2]);
            out[32]                = make_float4(sv[3], sv[4], sv[5], sv[9]);
            out[kBlockFloat4]      = make_float4(sv[6], sv[7], sv[8], sv[10]);
            out[kBlockFloat4 + 32] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// ------------------------------------------------------------------------------------------
// state initialisation and caller <-> internal order transfers
// ------------------------------------------------------------------------------------------
__global__ void InitState(
    int64_t nV,
    const int32_t* new2old,
    const double* X,
    const double* v,      // may be null
    const double* aext,   // may be null => (0,0,-9.81)
    const double* m,
    const uint8_t* isDbc,
    uint32_t pOff,
    float4* pos,
    float4* hist,
    float4* xtildeM,
    float4* xt,
    float4* vel,
    float4* vtm1,
    float4* acc)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nV)
        return;
    int64_t const o  = new2old[i];
    float4 const x4  = make_float4(static_cast<float>(X[3 * o]), static_cast<float>(X[3 * o + 1]), static_cast<float>(X[3 * o + 2]), 0.f);
    bool const fixed = isDbc[o] != 0;
    float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f), a4 = make_float4(0.f, 0.f, fixed ? 0.f : -9.81f, 0.f);
    if (v && !fixed)
        v4 = make_float4(static_cast<float>(v[3 * o]), static_cast<float>(v[3 * o + 1]), static_cast<float>(v[3 * o + 2]), 0.f);
    if (aext && !fixed)
        a4 = make_float4(static_cast<float>(aext[3 * o]), static_cast<float>(aext[3 * o + 1]), static_cast<float>(aext[3 * o + 2]), 0.f);
    pos[i] = x4;
    if (pOff)
        pos[pOff + i] = x4;
    if (hist)
        hist[i] = x4;
    xtildeM[i] = make_float4(x4.x, x4.y, x4.z, static_cast<float>(m[o]));
    xt[i]      = x4;
    vel[i]     = v4;
    if (vtm1)
        vtm1[i] = v4;
    acc[i] = a4;
}

// Caller order <-> internal order.  Component d of caller vertex o sits at src[d * sd + o * si]:
// (sd, si) = (1, 3) for a column-major 3 x nV matrix (Eigen, xyz interleaved), (nV, 1) for a row-major one.
template <class T>
__global__ void ScatterFromCaller(int64_t nV, const int32_t* old2new, const T* src, int64_t sd, int64_t si, float4* dst0, float4* dst1,
                                  int64_t ghostBegin)
{
    int64_t const o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (o >= nV)
        return;
    int64_t const i = old2new[o];
    if (i >= ghostBegin)
        return;  // ghosts belong to the GPU that owns the vertex: it may be writing them right now
    T const* a      = src + o * si;
    float4 const q  = make_float4(static_cast<float>(a[0]), static_cast<float>(a[sd]), static_cast<float>(a[2 * sd]), 0.f);
    float const w   = dst0[i].w;
    dst0[i]         = make_float4(q.x, q.y, q.z, w);
    if (dst1)
        dst1[i] = q;
}

template <class T>
__global__ void GatherToCaller(int64_t nV, const int32_t* old2new, const float4* src, T* dst, int64_t sd, int64_t si)
{
    int64_t const o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (o >= nV)
        return;
    float4 const q = src[old2new[o]];
    T* a           = dst + o * si;
    a[0]           = static_cast<T>(q.x);
    a[sd]          = static_cast<T>(q.y);
    a[2 * sd]      = static_cast<T>(q.z);
}

}  // namespace vbdx
