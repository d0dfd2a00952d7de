// Greedy vertex colouring on the device for the selection strategy that parallelises.
//
// The reference colours sequentially (graph/Color.h:45-135): vertices in natural order or in stable order of their degree in
// the primal graph, each taking a colour none of its already coloured neighbours has.  With the FirstAvailable selection
// that colour is the smallest one missing among the neighbours that come EARLIER in the order -- it depends on nothing
// else, so every vertex can be coloured as soon as those neighbours are (Jones-Plassmann with the reference's order as the
// priority), and the result is the sequential one, vertex for vertex.  The LeastUsed selection (the reference's default)
// picks by how many vertices each colour has so far, a global running count: that one is inherently sequential and stays
// on the host (csrc/plan.cpp, GreedyColorGraph).
//
// Works on the vertex -> tet incidence lists (no vertex-vertex graph is built): a vertex' neighbours are the other three
// vertices of its incident tets, enumerated with repetitions; only the degree needs them distinct.
#pragma once

#include "setup_kernels.cuh"

namespace vbdx {

constexpr int kColorMaxNeighbours = 384;  // distinct neighbours a vertex may have here (beyond: the host colours)
constexpr int kColorMaskWords     = 8;    // up to 256 colours

// degree in the primal graph = number of distinct vertices sharing a tet with v
__global__ void ColorDegrees(int64_t nV, const int32_t* E, const uint32_t* ptr, const uint32_t* adj, uint32_t* deg, uint32_t* maxDeg, uint32_t* overflow)
{
    int64_t const v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= nV)
        return;
    int32_t seen[kColorMaxNeighbours];
    int n = 0;
    for (uint32_t k = ptr[v]; k < ptr[v + 1]; ++k)
    {
        uint32_t const e = adj[k] >> 2;
        for (int a = 0; a < 4; ++a)
        {
            int32_t const u = E[4 * static_cast<size_t>(e) + a];
            if (u == v)
                continue;
            bool dup = false;
            for (int j = 0; j < n && !dup; ++j)
                dup = seen[j] == u;
            if (!dup)
            {
                if (n == kColorMaxNeighbours)
                {
                    atomicExch(overflow, 1u);
                    return;
                }
                seen[n++] = u;
            }
        }
    }
    deg[v] = static_cast<uint32_t>(n);
    atomicMax(maxDeg, static_cast<uint32_t>(n));
}

// position of a vertex in the reference's visiting order, as a sortable key: (degree slot, index) -- stable counting sort
__device__ __forceinline__ uint64_t ColorKey(uint32_t v, uint32_t deg, uint32_t maxDeg, int ordering)
{
    uint32_t const slot = ordering == 0 ? 0u : ordering == 2 ? maxDeg - deg : deg;  // Natural / LargestDegree / SmallestDegree
    return (static_cast<uint64_t>(slot) << 32) | v;
}

// One round: every uncoloured vertex whose earlier neighbours are all coloured takes the smallest colour they leave.
__global__ void ColorRound(int64_t nV, const int32_t* E, const uint32_t* ptr, const uint32_t* adj, const uint32_t* deg, const uint32_t* maxDegPtr,
                           int ordering, int32_t* colors, uint32_t* nColored, uint32_t* tooManyColors)
{
    int64_t const v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= nV || __ldcg(colors + v) >= 0)
        return;
    uint32_t const maxDeg = *maxDegPtr;
    uint64_t const mine   = ColorKey(static_cast<uint32_t>(v), deg[v], maxDeg, ordering);
    uint32_t mask[kColorMaskWords];
    for (int w = 0; w < kColorMaskWords; ++w)
        mask[w] = 0u;
    for (uint32_t k = ptr[v]; k < ptr[v + 1]; ++k)
    {
        uint32_t const e = adj[k] >> 2;
        for (int a = 0; a < 4; ++a)
        {
            int32_t const u = E[4 * static_cast<size_t>(e) + a];
            if (u == v || ColorKey(static_cast<uint32_t>(u), deg[u], maxDeg, ordering) > mine)
                continue;  // (a later vertex: not coloured when the sequential algorithm reaches v)
            int32_t const c = __ldcg(colors + u);
            if (c < 0)
                return;  // an earlier neighbour is not coloured yet: next round
            if (c < 32 * kColorMaskWords)
                mask[c >> 5] |= 1u << (c & 31);
        }
    }
    int color = -1;
    for (int w = 0; w < kColorMaskWords && color < 0; ++w)
        if (mask[w] != 0xffffffffu)
            color = 32 * w + __ffs(static_cast<int>(~mask[w])) - 1;
    if (color < 0)
    {
        atomicExch(tooManyColors, 1u);
        color = 32 * kColorMaskWords;  // (reported as an error by the host)
    }
    colors[v] = color;
    atomicAdd(nColored, 1u);
}

}  // namespace vbdx
