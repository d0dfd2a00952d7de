// Vertex-triangle contact for the B200 VBD integrator: active-set management over a linear BVH and
// the per-vertex contact energy.  Written from scratch; the behaviour it reproduces is the reference's
// GPU path (the reference's CPU integrator has no contact):
//   gpu/impl/contact/VertexTriangleMixedCcdDcd.cu:51-223   InitializeActiveSet / UpdateActiveSet / FinalizeActiveSet
//   gpu/impl/vbd/Integrator.cu:82-103,163-188,250-273      where Step calls them and fills the contact lists fc
//   gpu/impl/vbd/Kernels.cuh:80-114,203-223                area-scaled penalty, loop over <= 8 contacts
//   sim/vbd/Kernels.h:223-302                              normal penalty + IPC-style smoothed friction
// Deliberate, documented deviations (DESIGN.md "Contact"): the warm-start radius `dupper` of a vertex is
// the maximum over the triangles whose swept box it overlaps (the reference keeps whichever its traversal
// visited last), and vertices are sorted by a stable sort (the reference's sort of the query points is
// unstable): both only affect implementation-defined orderings in the reference.
#pragma once

#include "lbvh.cuh"

namespace vbdx {

constexpr int kMaxContacts = 8;  // gpu/impl/contact/VertexTriangleMixedCcdDcd.cuh kMaxNeighbours

// ------------------------------------------------------------------------------------------
// geometry helpers (fp32 restatements)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 F3(float4 a) { return make_float3(a.x, a.y, a.z); }
__device__ __forceinline__ float3 Sub(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 Add(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 Mul(float s, float3 a) { return make_float3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float Dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 Cross(float3 a, float3 b)
{
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// squared distance point <-> triangle, Ericson's region walk (geometry/ClosestPointQueries.h:299-393,
// geometry/DistanceQueries.h:212-221)
__device__ __forceinline__ float PointTriangleDistance2(float3 P, float3 A, float3 B, float3 C)
{
    float3 const AB = Sub(B, A), AC = Sub(C, A), AP = Sub(P, A);
    float const d1 = Dot(AB, AP), d2 = Dot(AC, AP);
    float u, v, w;
    float3 const BP = Sub(P, B);
    float const d3 = Dot(AB, BP), d4 = Dot(AC, BP);
    float3 const CP = Sub(P, C);
    float const d5 = Dot(AB, CP), d6 = Dot(AC, CP);
    float const vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
    if (d1 <= 0.f && d2 <= 0.f)
        u = 1.f, v = 0.f, w = 0.f;
    else if (d3 >= 0.f && d4 <= d3)
        u = 0.f, v = 1.f, w = 0.f;
    else if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f)
    {
        v = d1 / (d1 - d3);
        u = 1.f - v, w = 0.f;
    }
    else if (d6 >= 0.f && d5 <= d6)
        u = 0.f, v = 0.f, w = 1.f;
    else if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f)
    {
        w = d2 / (d2 - d6);
        u = 1.f - w, v = 0.f;
    }
    else if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f)
    {
        w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        u = 0.f, v = 1.f - w;
    }
    else
    {
        float const denom = 1.f / (va + vb + vc);
        v = vb * denom, w = vc * denom;
        u = 1.f - v - w;
    }
    float3 const Q = Add(Add(Mul(u, A), Mul(v, B)), Mul(w, C));
    float3 const D = Sub(P, Q);
    return Dot(D, D);
}

// The same operations with every rounding spelled out: the contact term below is inlined into several kernels (barrier,
// barrier-free, XPBD-free paths ...), and left to itself the compiler contracts a * b + c into an fma differently from one
// inlining context to the next -- one ulp that flips a borderline inside-the-triangle decision and makes two kernels that
// must agree bit for bit diverge (seen on BASELINE configs[2], where stacked identical grids project vertex onto vertex).
__device__ __forceinline__ float3 SubR(float3 a, float3 b) { return make_float3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ float3 MulR(float s, float3 a) { return make_float3(__fmul_rn(s, a.x), __fmul_rn(s, a.y), __fmul_rn(s, a.z)); }
__device__ __forceinline__ float DotR(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, __fmul_rn(a.x, b.x))); }
__device__ __forceinline__ float3 CrossR(float3 a, float3 b)
{
    return make_float3(fmaf(a.y, b.z, -__fmul_rn(a.z, b.y)), fmaf(a.z, b.x, -__fmul_rn(a.x, b.z)), fmaf(a.x, b.y, -__fmul_rn(a.y, b.x)));
}
// u a + v b + w c
__device__ __forceinline__ float3 Blend(float u, float3 a, float v, float3 b, float w, float3 c)
{
    return make_float3(fmaf(w, c.x, fmaf(v, b.x, __fmul_rn(u, a.x))), fmaf(w, c.y, fmaf(v, b.y, __fmul_rn(u, a.y))),
                       fmaf(w, c.z, fmaf(v, b.z, __fmul_rn(u, a.z))));
}

// Contact energy derivatives of one (vertex, triangle) pair (sim/vbd/Kernels.h:223-302).
// g += dE/dx_v, H (symmetric, 6 entries h00 h01 h02 h11 h12 h22) += d2E/dx_v2
__device__ __forceinline__ void AccumulateVertexTriangleContact(
    float3 xtv, float3 xv, float3 const xtf[3], float3 const xf[3], float dt, float k, float muF, float epsv,
    float g[3], float H[6])
{
    float3 T0 = SubR(xf[1], xf[0]), T1 = SubR(xf[2], xf[0]);
    float3 n             = CrossR(T0, T1);
    float const dblarea  = __fsqrt_rn(DotR(n, n));
    if (dblarea <= 1e-8f)
        return;
    n               = MulR(__fdiv_rn(1.f, dblarea), n);
    float3 const xc = SubR(xv, MulR(DotR(n, SubR(xv, xf[0])), n));
    // barycentric coordinates of the projection (geometry/IntersectionQueries.h:45-67)
    float3 const AP = SubR(xc, xf[0]);
    float const d00 = DotR(T0, T0), d01 = DotR(T0, T1), d11 = DotR(T1, T1), d20 = DotR(AP, T0), d21 = DotR(AP, T1);
    float const denom = fmaf(d00, d11, -__fmul_rn(d01, d01));
    float const bv = __fdiv_rn(fmaf(d11, d20, -__fmul_rn(d01, d21)), denom), bw = __fdiv_rn(fmaf(d00, d21, -__fmul_rn(d01, d20)), denom);
    float const bu = __fsub_rn(__fsub_rn(1.f, bv), bw);
    bool const inside = bu >= 0.f && bu <= 1.f && bv >= 0.f && bv <= 1.f && bw >= 0.f && bw <= 1.f;
    if (!inside)
        return;
    float3 const xb    = Blend(bu, xf[0], bv, xf[1], bw, xf[2]);
    float const d      = fminf(0.f, DotR(SubR(xv, xb), n));
    float const lambda = __fmul_rn(k, d);
    g[0] = fmaf(lambda, n.x, g[0]), g[1] = fmaf(lambda, n.y, g[1]), g[2] = fmaf(lambda, n.z, g[2]);
    float3 const kn = MulR(k, n);
    H[0] = fmaf(kn.x, n.x, H[0]), H[1] = fmaf(kn.x, n.y, H[1]), H[2] = fmaf(kn.x, n.z, H[2]);
    H[3] = fmaf(kn.y, n.y, H[3]), H[4] = fmaf(kn.y, n.z, H[4]), H[5] = fmaf(kn.z, n.z, H[5]);
    // IPC smooth friction: tangent basis = (unnormalised edge, n x edge) as the reference has it
    T1                  = CrossR(n, T0);
    float3 const xtb    = Blend(bu, xtf[0], bv, xtf[1], bw, xtf[2]);
    float3 const dx     = SubR(SubR(xv, xtv), SubR(xb, xtb));
    float const u0 = DotR(T0, dx), u1 = DotR(T1, dx);
    float const unorm   = __fadd_rn(__fsqrt_rn(fmaf(u1, u1, __fmul_rn(u0, u0))), FLT_EPSILON);
    float const epsvh   = __fmul_rn(epsv, dt);
    float const muFl    = __fmul_rn(muF, fabsf(lambda));
    float const y       = __fdiv_rn(unorm, epsvh);
    float const f1      = (y < 1.f) ? fmaf(-y, y, __fmul_rn(2.f, y)) : 1.f;
    float const c       = __fdiv_rn(__fmul_rn(muFl, f1), unorm);
    float3 const Tu     = make_float3(fmaf(u1, T1.x, __fmul_rn(u0, T0.x)), fmaf(u1, T1.y, __fmul_rn(u0, T0.y)), fmaf(u1, T1.z, __fmul_rn(u0, T0.z)));
    g[0] = fmaf(c, Tu.x, g[0]), g[1] = fmaf(c, Tu.y, g[1]), g[2] = fmaf(c, Tu.z, g[2]);
    H[0] = fmaf(c, fmaf(T1.x, T1.x, __fmul_rn(T0.x, T0.x)), H[0]), H[1] = fmaf(c, fmaf(T1.x, T1.y, __fmul_rn(T0.x, T0.y)), H[1]);
    H[2] = fmaf(c, fmaf(T1.x, T1.z, __fmul_rn(T0.x, T0.z)), H[2]), H[3] = fmaf(c, fmaf(T1.y, T1.y, __fmul_rn(T0.y, T0.y)), H[3]);
    H[4] = fmaf(c, fmaf(T1.y, T1.z, __fmul_rn(T0.y, T0.z)), H[4]), H[5] = fmaf(c, fmaf(T1.z, T1.z, __fmul_rn(T0.z, T0.z)), H[5]);
}

// Area-scaled penalty of a vertex' contacts (ContactPenalty, gpu/impl/vbd/Kernels.cuh:80-114): the triangles listed in
// the vertex' row of fc (the non-negative entries, which come first), kC = XVA[v] muC / sum of their areas; contact c is
// penalised with kC * FA[f[c]].  Returns kC (0 without contacts).
__device__ __forceinline__ float ContactPenaltyScale(const int* __restrict__ fcv, const float* __restrict__ FA, float xva, float muC,
                                                    int f[kMaxContacts], int& nContacts)
{
    nContacts   = 0;
    float sumfa = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxContacts; ++c)
    {
        f[c] = __ldcg(fcv + c);
        if (f[c] >= 0)
            ++nContacts;
    }
    for (int c = 0; c < nContacts; ++c)
        sumfa = __fadd_rn(sumfa, __ldg(FA + f[c]));
    return nContacts > 0 ? __fdiv_rn(__fmul_rn(xva, muC), sumfa) : 0.f;
}

// ------------------------------------------------------------------------------------------
// active-set kernels.  Vertex and triangle ids are *internal* ids; q indexes the sorted query order.
// ------------------------------------------------------------------------------------------
struct ContactMesh {
    const int32_t* B;  // body of every internal vertex
    const int32_t* V;  // collision vertices (internal ids)
    const int4* F;     // collision triangles (internal vertex ids; .w = the triangle's body)
    uint32_t nCV, nF;
};

// scene bounds when the caller supplied none: min/max over positions (float atomics via ordered ints)
__device__ __forceinline__ int FloatToOrdered(float f)
{
    int const i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float OrderedToFloat(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void SceneBoundsReset(int* bounds)
{
    if (threadIdx.x < 3)
    {
        bounds[threadIdx.x]     = FloatToOrdered(FLT_MAX);
        bounds[3 + threadIdx.x] = FloatToOrdered(-FLT_MAX);
    }
}

__global__ void SceneBoundsReduce(const float4* pos, uint32_t n, int* bounds)
{
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        float4 const p = pos[i];
        lo[0] = fminf(lo[0], p.x), lo[1] = fminf(lo[1], p.y), lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x), hi[1] = fmaxf(hi[1], p.y), hi[2] = fmaxf(hi[2], p.z);
    }
    for (int d = 0; d < 3; ++d)
    {
        for (int o = 16; o > 0; o >>= 1)
        {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    }
    // one atomic per block and component: same-address atomics serialise
    __shared__ float sLo[32][3], sHi[32][3];
    uint32_t const warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nWarps = blockDim.x >> 5;
    if (lane == 0)
        for (int d = 0; d < 3; ++d)
            sLo[warp][d] = lo[d], sHi[warp][d] = hi[d];
    __syncthreads();
    if (threadIdx.x < 3)
    {
        float l = sLo[0][threadIdx.x], h = sHi[0][threadIdx.x];
        for (uint32_t w = 1; w < nWarps; ++w)
            l = fminf(l, sLo[w][threadIdx.x]), h = fmaxf(h, sHi[w][threadIdx.x]);
        atomicMin(&bounds[threadIdx.x], FloatToOrdered(l));
        atomicMax(&bounds[3 + threadIdx.x], FloatToOrdered(h));
    }
}

__global__ void SceneBoundsFinish(const int* bounds, WorldBox* w)
{
    if (threadIdx.x < 3)
    {
        float const lo = OrderedToFloat(bounds[threadIdx.x]), hi = OrderedToFloat(bounds[3 + threadIdx.x]);
        w->lo[threadIdx.x]  = lo;
        w->ext[threadIdx.x] = hi > lo ? hi - lo : 1.f;
    }
}

// predictor of the full step (gpu/impl/vbd/Integrator.cu:163-188): x1 = xt + dt v + dt^2 aext
__device__ __forceinline__ float3 Predict(float4 x, float4 v, float4 a, float dt)
{
    return make_float3(x.x + dt * v.x + dt * dt * a.x, x.y + dt * v.y + dt * dt * a.y, x.z + dt * v.z + dt * dt * a.z);
}

// swept boxes of the collision vertices, in the order given by ids (VertexTriangleMixedCcdDcd.cu:62-75)
__global__ void SweptPointBoxes(ContactMesh m, const uint32_t* ids, const float4* x, const float4* vel, const float4* aext, float dt,
                                float4* lo, float4* hi)
{
    uint32_t const q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m.nCV)
        return;
    int const i     = m.V[ids[q]];
    float4 const p0 = x[i];
    float3 const p1 = Predict(p0, vel[i], aext[i], dt);
    lo[q] = make_float4(fminf(p0.x, p1.x), fminf(p0.y, p1.y), fminf(p0.z, p1.z), 0.f);
    hi[q] = make_float4(fmaxf(p0.x, p1.x), fmaxf(p0.y, p1.y), fmaxf(p0.z, p1.z), 0.f);
}

// swept boxes of the triangles by triangle id (VertexTriangleMixedCcdDcd.cu:89-103); dt = 0: current boxes
__global__ void TriangleBoxes(ContactMesh m, const float4* x, const float4* vel, const float4* aext, float dt, float4* lo, float4* hi)
{
    uint32_t const f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= m.nF)
        return;
    int4 const t = m.F[f];
    int const id[3] = {t.x, t.y, t.z};
    float l[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, h[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int k = 0; k < 3; ++k)
    {
        float4 const p0 = x[id[k]];
        l[0] = fminf(l[0], p0.x), l[1] = fminf(l[1], p0.y), l[2] = fminf(l[2], p0.z);
        h[0] = fmaxf(h[0], p0.x), h[1] = fmaxf(h[1], p0.y), h[2] = fmaxf(h[2], p0.z);
        if (dt != 0.f)
        {
            float3 const p1 = Predict(p0, vel[id[k]], aext[id[k]], dt);
            l[0] = fminf(l[0], p1.x), l[1] = fminf(l[1], p1.y), l[2] = fminf(l[2], p1.z);
            h[0] = fmaxf(h[0], p1.x), h[1] = fmaxf(h[1], p1.y), h[2] = fmaxf(h[2], p1.z);
        }
    }
    // .w: the triangle's body, as the range [lo.w, hi.w] (integer bits) that BvhRefit merges up the tree -- a node whose
    // range is one body can be skipped as a whole by a query of that body (MarkActive)
    lo[f] = make_float4(l[0], l[1], l[2], __int_as_float(t.w));
    hi[f] = make_float4(h[0], h[1], h[2], __int_as_float(t.w));
}

// VertexTriangleMixedCcdDcd.cu:107-140: a vertex becomes active when its swept box overlaps the swept box
// of a triangle of another body; dupper = squared diagonal of the joint box (max over such triangles)
__global__ void MarkActive(ContactMesh m, BvhView t, const uint32_t* ids, const float4* qlo, const float4* qhi,
                           const float4* triLo, const float4* triHi, uint8_t* active, float* dupper)
{
    uint32_t const q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m.nCV)
        return;
    uint32_t const v = ids[q];
    int const i      = m.V[v];
    int const body   = m.B[i];
    float4 const lo = qlo[q], hi = qhi[q];
    float best = -1.f;
    // (nodes that hold triangles of the query's own body only are not descended into: their leaves would all be rejected)
    BvhForEachOverlap(t, lo, hi, [&](int, uint32_t f) {
        if (m.F[f].w == body)
            return;  // no self collision
        float4 const tl = triLo[f], th = triHi[f];
        float const dx = fmaxf(hi.x, th.x) - fminf(lo.x, tl.x), dy = fmaxf(hi.y, th.y) - fminf(lo.y, tl.y),
                    dz = fmaxf(hi.z, th.z) - fminf(lo.z, tl.z);
        best = fmaxf(best, dx * dx + dy * dy + dz * dz);
    }, body);
    if (best >= 0.f)
    {
        active[v] = 1;
        dupper[v] = best;
    }
}

__global__ void ActiveFlags(const uint32_t* ids, const uint8_t* active, uint32_t n, uint32_t* flags)
{
    uint32_t const q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q <= n)
        flags[q] = (q < n && active[ids[q]]) ? 1u : 0u;
}

// av = active vertices in sorted order, -1 elsewhere (VertexTriangleMixedCcdDcd.cu:141-146)
__global__ void CompactActive(const uint32_t* ids, const uint32_t* flags, const uint32_t* offsets, uint32_t n, int32_t* av, uint32_t* nActive)
{
    uint32_t const q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n)
    {
        if (flags[q])
            av[offsets[q]] = static_cast<int32_t>(ids[q]);
        if (q >= offsets[n])
            av[q] = -1;
    }
    if (q == 0)
        *nActive = offsets[n];
}

// k-nearest triangles of the active vertices (VertexTriangleMixedCcdDcd.cuh:108-162).
// mode 0: UpdateActiveSet -> nn rows, and the contact lists fc of the vertices (gpu/impl/vbd/Integrator.cu:250-273)
// mode 1: FinalizeActiveSet -> active[v] = vertex is on the negative side of (the last of) its nearest triangles
//
// One WARP per active vertex.  The depth-first branch and bound of BvhNearest (lbvh.cuh; Bvh.cuh:347-474) is walked by
// all lanes in lockstep (same loads: broadcasts), but leaves are not evaluated one by one: the triangles whose box
// passes the bound are collected, 32 at a time, their point-triangle distances -- triangle row, three corner positions,
// Ericson's region walk: what the query's time goes into -- are computed one per lane, and the reference's sequential
// update (nearest so far, ties within +-eps, at most 8, in discovery order) is then replayed over them in that order.
// While a batch is collected the bound is the one of the previous batch, so more leaves pass than in the one-by-one
// walk; a leaf that the tighter bound would have skipped has d >= box distance > bound and changes nothing when it is
// replayed -- the result is the sequential one, list order included.
__global__ void NearestTriangles(ContactMesh m, BvhView t, const int32_t* av, const uint32_t* nActive, const float4* x, const float* dupper,
                                 float eps, int mode, int32_t* nn, int32_t* fc, uint8_t* active)
{
    uint32_t const lane = threadIdx.x & 31u;
    uint32_t const nAct = *nActive;
    // (grid-stride over the active list: its length is known on the device only, the grid is sized for a full GPU)
    for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nAct; q += (gridDim.x * blockDim.x) >> 5)
    {
    int const v    = av[q];
    int const i    = m.V[v];
    int const body = m.B[i];
    float3 const p = F3(x[i]);
    int found[kMaxContacts];
    int count   = 0;
    float dmin  = dupper[v];
    int mine    = -1;  // the candidate this lane evaluates
    int pending = 0;   // candidates collected since the last evaluation
    auto evaluate = [&]() {
        float d = FLT_MAX;
        if (static_cast<int>(lane) < pending)
        {
            int4 const tri = m.F[mine];
            if (tri.w != body)
                d = PointTriangleDistance2(p, F3(x[tri.x]), F3(x[tri.y]), F3(x[tri.z]));
        }
        for (int c = 0; c < pending; ++c)
        {
            float const dc  = __shfl_sync(0xffffffffu, d, c);
            int const prim  = __shfl_sync(0xffffffffu, mine, c);
            float const lo = dmin - eps, hi = dmin + eps;
            if (dc < lo)
            {
                count          = 0;
                found[count++] = prim;
                dmin           = dc;
            }
            else if (dc <= hi && count < kMaxContacts)
                found[count++] = prim;
        }
        pending = 0;
    };
    int stack[kBvhStack];
    int top         = 0;
    stack[top++]    = 0;
    int const leaf0 = static_cast<int>(t.n) - 1;
    do
    {
        int const node = stack[--top];
        float const db = PointBoxDistance2(p, t.nodeLo[node], t.nodeHi[node]);
        if (db <= dmin + eps)
        {
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
                if (static_cast<int>(lane) == pending)
                    mine = prim;
                if (++pending == 32)
                    evaluate();
            }
        }
        if (top == 0 && pending > 0)
            evaluate();  // (may not push anything: the walk ends)
    } while (top > 0);
    if (lane != 0)
        continue;
    if (mode == 0)
    {
        for (int k = 0; k < kMaxContacts; ++k)
        {
            int const f               = k < count ? found[k] : -1;
            nn[v * kMaxContacts + k]  = f;
            fc[i * kMaxContacts + k]  = f;
        }
    }
    else if (count > 0)
    {
        int4 const tri  = m.F[found[count - 1]];
        float3 const A = F3(x[tri.x]), Bq = F3(x[tri.y]), C = F3(x[tri.z]);
        float3 const n  = Cross(Sub(Bq, A), Sub(C, A));
        active[v]       = Dot(Sub(p, A), n) < 0.f ? 1 : 0;  // sign of geometry/DistanceQueries.h PointPlane
    }
    }
}

__global__ void FillI32(int32_t* a, int32_t v, size_t n)
{
    size_t const i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        a[i] = v;
}

}  // namespace vbdx
