// Linear BVH (Karras 2012) for sm_100a: 30-bit Morton codes, a hand-written stable LSD radix sort,
// hierarchy generation, bottom-up boxes, and the two traversals the VBD contact path needs (box overlap
// and k-nearest-neighbour branch and bound).  Written from scratch; behaviourally it follows the
// reference's
//   geometry/Morton.h:35-80, gpu/impl/geometry/Morton.cu:23-44      (codes of box centroids)
//   gpu/impl/geometry/Bvh.cu:139-157                                 (stable sort by code, then by index)
//   gpu/impl/geometry/Bvh.cu:24-116                                  (range / split / children / rightmost)
//   gpu/impl/geometry/Bvh.cu:175-233                                 (internal boxes, atomic visit counters)
//   gpu/impl/geometry/Bvh.cuh:347-474,476-581                        (nearest neighbours, range search)
// so that the tree topology equals the reference's golden arrays (gpu/impl/geometry/Bvh.cu:367-410).
// Node numbering is the reference's: internal nodes 0..n-2 (root 0), leaves n-1..2n-2 in sorted order.
#pragma once

#include "setup_kernels.cuh"

#include <algorithm>
#include <cfloat>
#include <cuda_runtime.h>

namespace vbdx {

// ------------------------------------------------------------------------------------------
// stable LSD radix sort of (uint32 key, uint32 value) pairs, 8 bits per pass, ONE cooperative launch
//   one warp owns a contiguous segment of kSortSegment items and walks it in chunks of 32, so that
//   ranks within a digit follow item order (stability) without any cross-warp bookkeeping;
//   per pass: per-segment digit counts -> grid barrier -> exclusive scan of the (digit-major) count table, every CTA
//   its own contiguous chunk -> grid barrier -> scatter (offset = scanned count + sum of the chunks before) -> grid
//   barrier.  The collision meshes of the contact path are 1e5 keys: 20 small launches per sort cost more in launch
//   gaps than in work (profiles/r02g_config3_launches.csv), twelve grid barriers of <= 148 CTAs do not.
//   Everything another CTA wrote during the launch is read with ld.global.cg (the L1 is not coherent).
// ------------------------------------------------------------------------------------------
constexpr int kSortSegment     = 256;   // 8 chunks per warp: short dependent chains, n/256 warps in flight
constexpr int kSortWarpsPerCta = 8;
constexpr int kSortMaxCtas     = 512;   // chunk sums of the scan live in shared memory

// per call site: grid barrier counter [0] (zeroed before every launch: the launch may be a replayed graph node) + chunk sums [1 ..]
struct RadixSortSync {
    unsigned int* aux = nullptr;  // 1 + kSortMaxCtas words
    int maxCtas       = 0;        // co-resident CTAs the launch may use (<= SM count: a cheap barrier)
};

__device__ __forceinline__ void SortGridSync(unsigned int* counter, unsigned int& target)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        target += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int v;
        do
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        while (static_cast<int>(v - target) < 0);
    }
    __syncthreads();
}

// exclusive scan of one value per thread over the CTA (kSortWarpsPerCta warps); total = sum over the CTA
__device__ __forceinline__ uint32_t SortBlockScan(uint32_t v, uint32_t* warpSums, uint32_t& total)
{
    uint32_t const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (uint32_t o = 1; o < 32; o <<= 1)
    {
        uint32_t const t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
            inc += t;
    }
    __syncthreads();  // the previous round's readers of warpSums are done
    if (lane == 31)
        warpSums[warp] = inc;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (uint32_t w = 0; w < kSortWarpsPerCta; ++w)
    {
        uint32_t const sum = warpSums[w];
        before += w < warp ? sum : 0u;
        all += sum;
    }
    total = all;
    return inc - v + before;
}

__global__ void __launch_bounds__(kSortWarpsPerCta * 32)
    RadixSortFused(uint32_t* keys, uint32_t* vals, uint32_t* keysTmp, uint32_t* valsTmp, uint32_t n, uint32_t nSeg, uint32_t* counts,
                   unsigned int* aux, unsigned int target)
{
    __shared__ uint32_t tab[kSortWarpsPerCta][256];
    __shared__ uint32_t warpSums[kSortWarpsPerCta];
    __shared__ uint32_t chunkBefore[kSortMaxCtas];
    uint32_t const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t const L     = 256u * nSeg;                                  // the count table, digit-major
    uint32_t const chunk = ((L + gridDim.x - 1) / gridDim.x + 3u) & ~3u;  // per CTA, a multiple of 4
    uint32_t *kin = keys, *vin = vals, *kout = keysTmp, *vout = valsTmp;
    for (int pass = 0; pass < 4; ++pass)
    {
        int const shift = 8 * pass;
        // digit counts of every segment
        for (uint32_t seg = blockIdx.x * kSortWarpsPerCta + warp; seg < nSeg; seg += gridDim.x * kSortWarpsPerCta)
        {
            for (uint32_t d = lane; d < 256; d += 32)
                tab[warp][d] = 0;
            __syncwarp();
            uint32_t const begin = seg * kSortSegment;
            uint32_t const end   = min(begin + static_cast<uint32_t>(kSortSegment), n);
            for (uint32_t base = begin; base < end; base += 32)
            {
                uint32_t const i    = base + lane;
                bool const in       = i < end;
                uint32_t const mask = __ballot_sync(0xffffffffu, in);
                if (in)
                {
                    uint32_t const d     = (__ldcg(kin + i) >> shift) & 255u;
                    uint32_t const peers = __match_any_sync(mask, d);
                    if (lane == static_cast<uint32_t>(__ffs(peers) - 1))
                        tab[warp][d] += __popc(peers);
                }
                __syncwarp();
            }
            for (uint32_t d = lane; d < 256; d += 32)
                counts[d * nSeg + seg] = tab[warp][d];
            __syncwarp();
        }
        SortGridSync(aux, target);
        // exclusive scan of this CTA's chunk of the table, in place; the chunk's sum goes to aux[1 + CTA]
        {
            uint32_t const cb = min(blockIdx.x * chunk, L), ce = min(cb + chunk, L);
            uint32_t carry    = 0;
            for (uint32_t base = cb; base < ce; base += kSortWarpsPerCta * 32 * 4)
            {
                uint32_t const i = base + threadIdx.x * 4u;
                uint32_t v[4];
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)
                    v[j] = i + j < ce ? __ldcg(counts + i + j) : 0u;
                uint32_t total;
                uint32_t run = carry + SortBlockScan(v[0] + v[1] + v[2] + v[3], warpSums, total);
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)
                {
                    if (i + j < ce)
                        counts[i + j] = run;
                    run += v[j];
                }
                carry += total;
            }
            if (threadIdx.x == 0)
                aux[1 + blockIdx.x] = carry;
        }
        SortGridSync(aux, target);
        // what the chunks before a chunk add up to
        {
            uint32_t carry = 0;
            for (uint32_t base = 0; base < gridDim.x; base += kSortWarpsPerCta * 32)
            {
                uint32_t const c = base + threadIdx.x;
                uint32_t total;
                uint32_t const ex = SortBlockScan(c < gridDim.x ? __ldcg(aux + 1 + c) : 0u, warpSums, total);
                if (c < gridDim.x)
                    chunkBefore[c] = carry + ex;
                carry += total;
            }
            __syncthreads();
        }
        // scatter
        for (uint32_t seg = blockIdx.x * kSortWarpsPerCta + warp; seg < nSeg; seg += gridDim.x * kSortWarpsPerCta)
        {
            for (uint32_t d = lane; d < 256; d += 32)
            {
                uint32_t const at = d * nSeg + seg;
                tab[warp][d]      = __ldcg(counts + at) + chunkBefore[at / chunk];
            }
            __syncwarp();
            uint32_t const begin = seg * kSortSegment;
            uint32_t const end   = min(begin + static_cast<uint32_t>(kSortSegment), n);
            for (uint32_t base = begin; base < end; base += 32)
            {
                uint32_t const i    = base + lane;
                bool const in       = i < end;
                uint32_t const mask = __ballot_sync(0xffffffffu, in);
                uint32_t key = 0, val = 0, d = 0, peers = 0, dst = 0;
                if (in)
                {
                    key   = __ldcg(kin + i);
                    val   = __ldcg(vin + i);
                    d     = (key >> shift) & 255u;
                    peers = __match_any_sync(mask, d);
                    dst   = tab[warp][d] + __popc(peers & ((1u << lane) - 1u));
                }
                __syncwarp();
                if (in && lane == static_cast<uint32_t>(__ffs(peers) - 1))
                    tab[warp][d] += __popc(peers);
                __syncwarp();
                if (in)
                {
                    kout[dst] = key;
                    vout[dst] = val;
                }
            }
            __syncwarp();
        }
        SortGridSync(aux, target);
        uint32_t* t = kin;
        kin         = kout;
        kout        = t;
        t           = vin;
        vin         = vout;
        vout        = t;
    }
}

// Sorts in place (4 passes ping-pong through the temporaries).  counts: 256 * ceil(n / kSortSegment) words.
inline void RadixSortPairs(
    uint32_t* keys,
    uint32_t* vals,
    uint32_t* keysTmp,
    uint32_t* valsTmp,
    uint32_t n,
    uint32_t* counts,
    RadixSortSync& sync,
    cudaStream_t s,
    int64_t* launches = nullptr)
{
    if (n < 2)
        return;
    uint32_t nSeg   = (n + kSortSegment - 1) / kSortSegment;
    int const need  = static_cast<int>((nSeg + kSortWarpsPerCta - 1) / kSortWarpsPerCta);
    int const ctas  = std::max(1, std::min({need, sync.maxCtas, kSortMaxCtas}));
    unsigned int zero = 0;
    void* args[]      = {&keys, &vals, &keysTmp, &valsTmp, &n, &nSeg, &counts, &sync.aux, &zero};
    VBDX_CUDA(cudaMemsetAsync(sync.aux, 0, sizeof(unsigned int), s));
    VBDX_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void const*>(RadixSortFused), dim3(ctas), dim3(kSortWarpsPerCta * 32), args, 0, s));
    if (launches)
        *launches += 1;
}

inline size_t RadixSortCountsSize(uint32_t n)
{
    return 256 * static_cast<size_t>((n + kSortSegment - 1) / kSortSegment) + 1;
}

// ------------------------------------------------------------------------------------------
// Morton codes (geometry/Morton.h:35-80)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ExpandBits10Dev(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__device__ __forceinline__ uint32_t Morton3D(float x, float y, float z)
{
    uint32_t const xx = ExpandBits10Dev(static_cast<uint32_t>(fminf(fmaxf(x * 1024.f, 0.f), 1023.f)));
    uint32_t const yy = ExpandBits10Dev(static_cast<uint32_t>(fminf(fmaxf(y * 1024.f, 0.f), 1023.f)));
    uint32_t const zz = ExpandBits10Dev(static_cast<uint32_t>(fminf(fmaxf(z * 1024.f, 0.f), 1023.f)));
    return xx * 4 + yy * 2 + zz;
}

struct WorldBox {
    float lo[3], ext[3];  // minimum corner and extent (gpu/impl/geometry/Morton.cu:32-40)
};

// code of the centroid of box i; also resets ids to the identity (the reference re-sequences `inds`
// before every BVH sort, gpu/impl/geometry/Bvh.cu:144)
__global__ void MortonOfBoxes(const float4* lo, const float4* hi, uint32_t n, const WorldBox* world, uint32_t* codes, uint32_t* ids)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    WorldBox const w = *world;
    float4 const a = lo[i], b = hi[i];
    codes[i] = Morton3D((0.5f * (a.x + b.x) - w.lo[0]) / w.ext[0], (0.5f * (a.y + b.y) - w.lo[1]) / w.ext[1],
                        (0.5f * (a.z + b.z) - w.lo[2]) / w.ext[2]);
    if (ids)
        ids[i] = i;
}

// ------------------------------------------------------------------------------------------
// the tree
// ------------------------------------------------------------------------------------------
struct BvhView {
    uint32_t n;            // leaves
    const uint32_t* codes; // sorted Morton codes
    const uint32_t* inds;  // leaf -> primitive
    int32_t* child[2];     // n-1
    int32_t* parent;       // 2n-1
    int32_t* rightmost[2]; // n-1
    float4* nodeLo;        // 2n-1: internal boxes then leaf boxes (sorted order)
    float4* nodeHi;
    uint32_t* visits;      // n-1: arrivals at the node, two per box computation (never reset: an odd count = one child is done)
    int2* up;              // 2n-1: (parent, sibling) of every node, (-1, -1) for the root
};

__device__ __forceinline__ int BvhDelta(const uint32_t* codes, int n, int i, int j)
{
    if (j < 0 || j >= n)
        return -1;
    uint32_t const a = codes[i], b = codes[j];
    if (a == b)
        return 32 + __clz(i ^ j);  // duplicate codes: fall back to the leaf index
    return __clz(a ^ b);
}

__global__ void BvhHierarchy(BvhView t)
{
    int const in = blockIdx.x * blockDim.x + threadIdx.x;
    int const n  = static_cast<int>(t.n);
    if (in >= n - 1)
        return;
    // direction and extent of the node's key range
    int const d    = (BvhDelta(t.codes, n, in, in + 1) - BvhDelta(t.codes, n, in, in - 1)) > 0 ? 1 : -1;
    int const dmin = BvhDelta(t.codes, n, in, in - d);
    int lmax       = 2;
    while (BvhDelta(t.codes, n, in, in + lmax * d) > dmin)
        lmax <<= 1;
    int l = 0;
    do
    {
        lmax >>= 1;
        if (BvhDelta(t.codes, n, in, in + (l + lmax) * d) > dmin)
            l += lmax;
    } while (lmax > 1);
    int const j = in + l * d;
    // split position
    int const dnode = BvhDelta(t.codes, n, in, j);
    int s = 0, len = l;
    do
    {
        len = (len + 1) >> 1;
        if (BvhDelta(t.codes, n, in, in + (s + len) * d) > dnode)
            s += len;
    } while (len > 1);
    int const gamma = in + s * d + min(d, 0);
    int const lo = min(in, j), hi = max(in, j);
    int const leafBegin = n - 1;
    int const lc        = (lo == gamma) ? leafBegin + gamma : gamma;
    int const rc        = (hi == gamma + 1) ? leafBegin + gamma + 1 : gamma + 1;
    t.child[0][in]      = lc;
    t.child[1][in]      = rc;
    t.parent[lc]        = in;
    t.parent[rc]        = in;
    t.up[lc]            = make_int2(in, rc);
    t.up[rc]            = make_int2(in, lc);
    t.rightmost[0][in]  = leafBegin + gamma;
    t.rightmost[1][in]  = leafBegin + hi;
}

// Boxes of all nodes from the per-primitive boxes, one launch (Bvh::ConstructBoxes, gpu/impl/geometry/Bvh.cu:175-233): a
// thread per leaf writes the leaf's box and climbs; at every node the second thread to arrive merges the box it carries
// in registers with its sibling's and goes on.  Per level the dependent chain is one ordered atomic (acq_rel: the box
// this thread wrote below is released with it, the sibling's -- written by the thread that arrived first -- acquired)
// and one box load; (parent, sibling) of the next level are requested before the atomic.  The arrival counters are
// never reset: two arrivals per box computation, so an odd count before the add means "the other child is done".
__global__ void BvhRefit(BvhView t, const float4* primLo, const float4* primHi)
{
    uint32_t const k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= t.n)
        return;
    uint32_t const f = t.inds[k];
    int cur          = static_cast<int>(t.n - 1 + k);
    float4 lo = primLo[f], hi = primHi[f];
    t.nodeLo[cur] = lo;
    t.nodeHi[cur] = hi;
    int2 u        = t.up[cur];
    while (u.x >= 0)
    {
        int2 const next = t.up[u.x];
        unsigned int old;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(t.visits + u.x), "r"(1u) : "memory");
        if ((old & 1u) == 0u)
            break;
        float4 const sl = __ldcg(t.nodeLo + u.y), sh = __ldcg(t.nodeHi + u.y);
        // (.w: an integer range carried with the box -- the bodies of the triangles below the node on the contact path)
        lo = make_float4(fminf(lo.x, sl.x), fminf(lo.y, sl.y), fminf(lo.z, sl.z), __int_as_float(min(__float_as_int(lo.w), __float_as_int(sl.w))));
        hi = make_float4(fmaxf(hi.x, sh.x), fmaxf(hi.y, sh.y), fmaxf(hi.z, sh.z), __int_as_float(max(__float_as_int(hi.w), __float_as_int(sh.w))));
        cur           = u.x;
        t.nodeLo[cur] = lo;
        t.nodeHi[cur] = hi;
        u             = next;
    }
}

// ------------------------------------------------------------------------------------------
// traversals (per-thread explicit stack, right child first like the reference's push order)
// ------------------------------------------------------------------------------------------
constexpr int kBvhStack = 64;

// geometry/OverlapQueries.h:608-616
__device__ __forceinline__ bool BoxesOverlap(float4 al, float4 ah, float4 bl, float4 bh)
{
    return (al.x <= bh.x && ah.x >= bl.x) && (al.y <= bh.y && ah.y >= bl.y) && (al.z <= bh.z && ah.z >= bl.z);
}

// calls f(leafSlot, primitive) for every leaf whose box overlaps [qlo, qhi]  (Bvh.cuh:282-345)
// skipOnly >= 0: nodes whose .w range (see BvhRefit) is exactly [skipOnly, skipOnly] are not descended into
template <class F>
__device__ __forceinline__ void BvhForEachOverlap(BvhView const& t, float4 qlo, float4 qhi, F f, int skipOnly = -1)
{
    int stack[kBvhStack];
    int top        = 0;
    stack[top++]   = 0;
    int const leaf0 = static_cast<int>(t.n) - 1;
    do
    {
        int const node = stack[--top];
        float4 const nl = t.nodeLo[node], nh = t.nodeHi[node];
        if (!BoxesOverlap(nl, nh, qlo, qhi))
            continue;
        if (skipOnly >= 0 && __float_as_int(nl.w) == skipOnly && __float_as_int(nh.w) == skipOnly)
            continue;
        if (node >= leaf0)
            f(node - leaf0, t.inds[node - leaf0]);
        else if (top + 2 <= kBvhStack)
        {
            stack[top++] = t.child[0][node];
            stack[top++] = t.child[1][node];
        }
    } while (top > 0);
}

// squared distance point <-> box (geometry/DistanceQueries.h:199-210)
__device__ __forceinline__ float PointBoxDistance2(float3 p, float4 lo, float4 hi)
{
    float const cx = fminf(fmaxf(p.x, lo.x), hi.x), cy = fminf(fmaxf(p.y, lo.y), hi.y), cz = fminf(fmaxf(p.z, lo.z), hi.z);
    float const dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
    return dx * dx + dy * dy + dz * dz;
}

// Depth-first branch and bound (Bvh.cuh:347-474): keeps the nearest primitive and the ones whose squared
// distance lies within +-eps of the running minimum, at most kMaxNN of them, in discovery order.
// dist(primitive) returns the squared distance (FLT_MAX to reject).  Returns the number found.
template <int kMaxNN, class FDist>
__device__ __forceinline__ int BvhNearest(BvhView const& t, float3 q, float dmin, float eps, FDist dist, int* out, float& dminOut)
{
    int stack[kBvhStack];
    int top         = 0;
    stack[top++]    = 0;
    int count       = 0;
    int const leaf0 = static_cast<int>(t.n) - 1;
    do
    {
        int const node  = stack[--top];
        float const lo  = dmin - eps, hi = dmin + eps;
        float const db  = PointBoxDistance2(q, t.nodeLo[node], t.nodeHi[node]);
        if (!(db <= hi))
            continue;
        if (node < leaf0)
        {
            if (top + 2 <= kBvhStack)
            {
                stack[top++] = t.child[0][node];
                stack[top++] = t.child[1][node];
            }
        }
        else
        {
            int const prim = static_cast<int>(t.inds[node - leaf0]);
            float const d  = dist(prim);
            if (d < lo)
            {
                count  = 0;
                out[count++] = prim;
                dmin   = d;
            }
            else if (d <= hi && count < kMaxNN)
                out[count++] = prim;
        }
    } while (top > 0);
    dminOut = dmin;
    return count;
}

}  // namespace vbdx
