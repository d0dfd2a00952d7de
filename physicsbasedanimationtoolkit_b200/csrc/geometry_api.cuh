// Stand-alone handles over the LBVH (lbvh.cuh) and the vertex-triangle detector (contact.cuh, contact_host.cuh): what
// pbat::gpu::geometry::Bvh (gpu/geometry/Bvh.h, bindings/pypbat/gpu/geometry/Bvh.cpp:19-109) and
// pbat::gpu::contact::VertexTriangleMixedCcdDcd (gpu/contact/VertexTriangleMixedCcdDcd.h,
// bindings/pypbat/gpu/contact/VertexTriangleMixedCcdDcd.cpp:18-82) expose on their own, outside the integrator
// (SURVEY.md section 8f rank 4).  Included by vbdx.cu; C entry points at the bottom of that file.
#pragma once

#include "contact_host.cuh"

namespace vbdx {

// Self-overlaps of the leaf boxes (Bvh::DetectOverlaps, gpu/impl/geometry/Bvh.cuh:282-345): one thread per leaf in
// sorted order; a subtree is skipped when its right-most leaf is not to the right of the query leaf, so that every pair
// is found exactly once.  Pairs are reported as (min, max) of the primitive indices; `set` (optional) keeps only pairs
// from different sets.  count may exceed maxPairs: the caller sees how many were dropped.
__global__ void BvhSelfOverlaps(BvhView t, const int32_t* set, int32_t* pairs, unsigned long long* count, unsigned long long maxPairs)
{
    int const s = blockIdx.x * blockDim.x + threadIdx.x;
    int const n = static_cast<int>(t.n);
    if (s >= n)
        return;
    int const leaf0 = n - 1;
    float4 const qlo = t.nodeLo[leaf0 + s], qhi = t.nodeHi[leaf0 + s];
    int const pi = static_cast<int>(t.inds[s]);
    if (n == 1)
        return;
    int stack[kBvhStack];
    int top      = 0;
    stack[top++] = 0;
    do
    {
        int const node = stack[--top];
        if (!BoxesOverlap(t.nodeLo[node], t.nodeHi[node], qlo, qhi))
            continue;
        if (node >= leaf0)
        {
            int const s2 = node - leaf0;
            if (s2 <= s)
                continue;
            int const pj = static_cast<int>(t.inds[s2]);
            if (set != nullptr && set[pi] == set[pj])
                continue;
            unsigned long long const k = atomicAdd(count, 1ull);
            if (k < maxPairs)
            {
                pairs[2 * k]     = min(pi, pj);
                pairs[2 * k + 1] = max(pi, pj);
            }
        }
        else if (top + 2 <= kBvhStack)
        {
            if (t.rightmost[0][node] - leaf0 > s)
                stack[top++] = t.child[0][node];
            if (t.rightmost[1][node] - leaf0 > s)
                stack[top++] = t.child[1][node];
        }
    } while (top > 0);
}

// nearest triangle of every query point (Bvh::PointTriangleNearestNeighbors): branch and bound over the triangle boxes
__global__ void BvhNearestTriangle(BvhView t, const float4* X, int64_t nQ, const float4* V, const int4* F, int32_t* out)
{
    int64_t const q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= nQ)
        return;
    float3 const p = F3(X[q]);
    int best[1]    = {-1};
    float dmin;
    int const found = BvhNearest<1>(
        t, p, FLT_MAX, 0.f,
        [&](int f) {
            int4 const tri = F[f];
            return PointTriangleDistance2(p, F3(V[tri.x]), F3(V[tri.y]), F3(V[tri.z]));
        },
        best, dmin);
    out[q] = found > 0 ? best[0] : -1;
}

struct BvhHandle {
    DeviceBvh bvh;
    DevBuf<float4> lo, hi;
    DevBuf<WorldBox> world;
    int64_t capacity = 0, n = 0, bytes = 0, launches = 0;
    int device = 0;
};

struct ContactHandle {
    ContactState cs;
    DevBuf<float4> xa, xb, zero;
    int64_t nV = 0, bytes = 0, launches = 0;
    int device = 0;
    float eps = FLT_EPSILON;
};

inline void UploadXyz(DevBuf<float4>& dst, const float* src, int64_t n, cudaStream_t s, const float* minus = nullptr)
{
    std::vector<float4> h(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i)
        h[i] = minus ? make_float4(src[3 * i] - minus[3 * i], src[3 * i + 1] - minus[3 * i + 1], src[3 * i + 2] - minus[3 * i + 2], 0.f)
                     : make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
    dst.Upload(h.data(), n, s);
    VBDX_CUDA(cudaStreamSynchronize(s));
}

}  // namespace vbdx
