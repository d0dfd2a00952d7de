// XPBD integrator for sm_100a over the same contact pipeline as the VBD path.
// Behaviour: pbat::gpu::impl::xpbd::Integrator::Step (gpu/impl/xpbd/Integrator.cu:88-186) -- per Step the active set from the
// full-step predictor; per substep: multipliers reset, x = xt + h v + h^2 a, nearest triangles, `iterations` x (block
// Neo-Hookean constraints partition by partition, then the contact constraints of every active vertex, Jacobi through xb),
// v = (x - xt) / h -- with the per-constraint arithmetic of sim/xpbd/Kernels.h (ProjectBlockNeoHookean :78-141,
// ProjectVertexTriangle :169-246).
//
// Built from scratch for the GPU: the reference issues one thrust::for_each per partition (plus two per contact pass, plus
// three per substep), i.e. iterations x (partitions + 2) + 3 host-synchronous launches per substep.  Here ONE persistent
// cooperative launch runs a whole substep (a whole Step without contact): partitions are separated by software grid
// barriers; constraints are stored partition-major as packed 64-byte records so that a partition is one coalesced stream;
// positions are float4 with the inverse mass in .w (one gather per vertex instead of two).
#pragma once

#include "contact_host.cuh"
#include "step_kernel.cuh"

namespace vbdx {

struct XpbdParams {
    float4* x;            // xyz, inverse mass in w
    float4* xt;           // positions at the start of the substep
    float4* vel;
    const float4* aext;
    // constraints in (partition, cluster, caller order within the cluster) order; work item = cluster
    const int4* tetIds;       // vertex ids
    const float4* rec;        // 3 float4 per constraint: DmInv column c (xyz) | w = gammaSNH, alpha_D, alpha_H
    const float2* beta;       // damping of the two constraints
    float2* lambda;           // multipliers (reset per substep)
    const uint32_t* itemBegin;  // per work item: first constraint (nItems + 1)
    const uint32_t* partBegin;  // per partition: first work item (nPartitions + 1)
    int nPartitions, nV, nT;
    float sdt, sdt2;
    int iterations, substeps;
    int skipPreStep;          // contact path: the pre-step ran in its own launch (the detector sits between it and the solve)
    unsigned int* barrier;
    // contact (nCV == 0: none)
    int nCV;
    const uint32_t* nActive;  // device-side count of active collision vertices
    const int32_t* av;        // active collision vertices (indices into V)
    const int32_t* V;
    const int4* triF;
    const int32_t* nn;        // 8 nearest triangles per collision vertex, -1 terminated
    float4* xb;
    const float* muV;         // collision penalty per collision vertex
    const float* alphaC;      // compliance / damping / multiplier per collision vertex
    const float* betaC;
    float* lambdaC;
    float muS, muD;
};

__device__ __forceinline__ void XpbdPreStepVertex(XpbdParams const& p, uint32_t i)
{
    float4 const x = __ldcg(p.x + i), v = __ldcg(p.vel + i), a = __ldg(p.aext + i);
    p.xt[i]        = x;
    p.x[i]         = make_float4(fmaf(p.sdt2, a.x, fmaf(p.sdt, v.x, x.x)), fmaf(p.sdt2, a.y, fmaf(p.sdt, v.y, x.y)),
                                 fmaf(p.sdt2, a.z, fmaf(p.sdt, v.z, x.z)), x.w);
}

// ProjectBlockNeoHookean (sim/xpbd/Kernels.h:78-141) on constraint slot c
__device__ __forceinline__ void XpbdProjectTet(XpbdParams const& p, uint32_t c)
{
    int4 const t      = __ldg(p.tetIds + c);
    float4 const r0 = __ldg(p.rec + 3 * c), r1 = __ldg(p.rec + 3 * c + 1), r2 = __ldg(p.rec + 3 * c + 2);
    float2 const be = __ldg(p.beta + c);
    float2 lam      = __ldcg(p.lambda + c);
    float4 const x0 = __ldcg(p.x + t.x), x1 = __ldcg(p.x + t.y), x2 = __ldcg(p.x + t.z), x3 = __ldcg(p.x + t.w);
    float4 const y0 = __ldcg(p.xt + t.x), y1 = __ldcg(p.xt + t.y), y2 = __ldcg(p.xt + t.z), y3 = __ldcg(p.xt + t.w);
    float const D[3][3] = {{r0.x, r1.x, r2.x}, {r0.y, r1.y, r2.y}, {r0.z, r1.z, r2.z}};  // DmInv(row, col)
    float const gammaSNH = r0.w, at0 = r1.w / p.sdt2, at1 = r2.w / p.sdt2;
    float const g0 = at0 * be.x * p.sdt, g1 = at1 * be.y * p.sdt;
    float const Ds[3][3] = {{x1.x - x0.x, x2.x - x0.x, x3.x - x0.x}, {x1.y - x0.y, x2.y - x0.y, x3.y - x0.y}, {x1.z - x0.z, x2.z - x0.z, x3.z - x0.z}};
    float F[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
            F[r][cc] = Ds[r][0] * D[0][cc] + Ds[r][1] * D[1][cc] + Ds[r][2] * D[2][cc];
    float CD = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
            CD += F[r][cc] * F[r][cc];
    CD = sqrtf(CD);
    // PH = cofactor of F (columns: F1 x F2, F2 x F0, F0 x F1)
    float PH[3][3];
    PH[0][0] = F[1][1] * F[2][2] - F[2][1] * F[1][2], PH[1][0] = F[2][1] * F[0][2] - F[0][1] * F[2][2], PH[2][0] = F[0][1] * F[1][2] - F[1][1] * F[0][2];
    PH[0][1] = F[1][2] * F[2][0] - F[2][2] * F[1][0], PH[1][1] = F[2][2] * F[0][0] - F[0][2] * F[2][0], PH[2][1] = F[0][2] * F[1][0] - F[1][2] * F[0][0];
    PH[0][2] = F[1][0] * F[2][1] - F[2][0] * F[1][1], PH[1][2] = F[2][0] * F[0][1] - F[0][0] * F[2][1], PH[2][2] = F[0][0] * F[1][1] - F[1][0] * F[0][1];
    float const detF = F[0][0] * PH[0][0] + F[1][0] * PH[1][0] + F[2][0] * PH[2][0];
    float const CH   = detF - gammaSNH;
    // gradients w.r.t. x1..x3: columns of (F DmInv^T) / CD and of PH DmInv^T; x0: minus their sum
    float gD[4][3], gH[4][3];
#pragma unroll
    for (int a = 1; a < 4; ++a)
#pragma unroll
        for (int r = 0; r < 3; ++r)
        {
            gD[a][r] = (F[r][0] * D[a - 1][0] + F[r][1] * D[a - 1][1] + F[r][2] * D[a - 1][2]) / CD;
            gH[a][r] = PH[r][0] * D[a - 1][0] + PH[r][1] * D[a - 1][1] + PH[r][2] * D[a - 1][2];
        }
#pragma unroll
    for (int r = 0; r < 3; ++r)
    {
        gD[0][r] = -(gD[1][r] + gD[2][r] + gD[3][r]);
        gH[0][r] = -(gH[1][r] + gH[2][r] + gH[3][r]);
    }
    float const minv[4] = {x0.w, x1.w, x2.w, x3.w};
    float const dxt[4][3] = {{x0.x - y0.x, x0.y - y0.y, x0.z - y0.z}, {x1.x - y1.x, x1.y - y1.y, x1.z - y1.z},
                             {x2.x - y2.x, x2.y - y2.y, x2.z - y2.z}, {x3.x - y3.x, x3.y - y3.y, x3.z - y3.z}};
    float dotD = 0.f, dotH = 0.f, A00 = 0.f, A11 = 0.f, A01 = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
    {
        float nD = 0.f, nH = 0.f, dh = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r)
        {
            dotD += gD[a][r] * dxt[a][r];
            dotH += gH[a][r] * dxt[a][r];
            nD += gD[a][r] * gD[a][r], nH += gH[a][r] * gH[a][r], dh += gD[a][r] * gH[a][r];
        }
        A00 += minv[a] * nD, A11 += minv[a] * nH, A01 += minv[a] * dh;
    }
    float const b0 = -(CD + at0 * lam.x + g0 * dotD), b1 = -(CH + at1 * lam.y + g1 * dotH);
    float const D0 = 1.f + g0, D1 = 1.f + g1;
    float const a00 = D0 * A00 + at0, a11 = D1 * A11 + at1, a01 = A01 * D0, a10 = A01 * D1;
    float const det = a00 * a11 - a01 * a10;
    float const dl0 = (a11 * b0 - a01 * b1) / det, dl1 = (a00 * b1 - a10 * b0) / det;
    lam.x += dl0, lam.y += dl1;
    p.lambda[c] = lam;
    p.x[t.x] = make_float4(x0.x + minv[0] * (dl0 * gD[0][0] + dl1 * gH[0][0]), x0.y + minv[0] * (dl0 * gD[0][1] + dl1 * gH[0][1]),
                           x0.z + minv[0] * (dl0 * gD[0][2] + dl1 * gH[0][2]), x0.w);
    p.x[t.y] = make_float4(x1.x + minv[1] * (dl0 * gD[1][0] + dl1 * gH[1][0]), x1.y + minv[1] * (dl0 * gD[1][1] + dl1 * gH[1][1]),
                           x1.z + minv[1] * (dl0 * gD[1][2] + dl1 * gH[1][2]), x1.w);
    p.x[t.z] = make_float4(x2.x + minv[2] * (dl0 * gD[2][0] + dl1 * gH[2][0]), x2.y + minv[2] * (dl0 * gD[2][1] + dl1 * gH[2][1]),
                           x2.z + minv[2] * (dl0 * gD[2][2] + dl1 * gH[2][2]), x2.w);
    p.x[t.w] = make_float4(x3.x + minv[3] * (dl0 * gD[3][0] + dl1 * gH[3][0]), x3.y + minv[3] * (dl0 * gD[3][1] + dl1 * gH[3][1]),
                           x3.z + minv[3] * (dl0 * gD[3][2] + dl1 * gH[3][2]), x3.w);
}

// ProjectVertexTriangle (sim/xpbd/Kernels.h:169-246); returns true if the constraint was projected
__device__ __forceinline__ bool XpbdProjectVertexTriangle(float minvv, float3 xvt, float3 const xft[3], float3 const xf[3], float muC, float muS,
                                                          float muD, float atildec, float gammac, float& lambdac, float3& xv)
{
    if (minvv < 1e-10f)
        return false;
    float3 const T1 = Sub(xf[1], xf[0]), T2 = Sub(xf[2], xf[0]);
    float3 n        = Cross(T1, T2);
    float const dbl = sqrtf(Dot(n, n));
    if (dbl <= 1e-8f)
        return false;
    n               = Mul(1.f / dbl, n);
    float3 const xc = Sub(xv, Mul(Dot(n, Sub(xv, xf[0])), n));
    float3 const AP = Sub(xc, xf[0]);
    float const d00 = Dot(T1, T1), d01 = Dot(T1, T2), d11 = Dot(T2, T2), d20 = Dot(AP, T1), d21 = Dot(AP, T2);
    float const denom = d00 * d11 - d01 * d01;
    float const bv = (d11 * d20 - d01 * d21) / denom, bw = (d00 * d21 - d01 * d20) / denom, bu = 1.f - bv - bw;
    if (!(bu >= 0.f && bu <= 1.f && bv >= 0.f && bv <= 1.f && bw >= 0.f && bw <= 1.f))
        return false;
    float const C = muC * Dot(n, Sub(xv, xf[0]));
    if (C > 0.f)
        return false;
    float const D       = 1.f + gammac;
    float const dlambda = -(C + atildec * lambdac + gammac * Dot(n, Sub(xv, xvt))) / (D * minvv + atildec);
    float3 dx           = Mul(dlambda * minvv, n);
    xv                  = Add(xv, dx);
    lambdac += dlambda;
    float const d    = sqrtf(Dot(dx, dx));
    float3 const xb  = Add(Add(Mul(bu, xf[0]), Mul(bv, xf[1])), Mul(bw, xf[2]));
    float3 const xtb = Add(Add(Mul(bu, xft[0]), Mul(bv, xft[1])), Mul(bw, xft[2]));
    dx               = Sub(Sub(xv, xvt), Sub(xb, xtb));
    dx               = Sub(dx, Mul(Dot(n, dx), n));
    float const dxd  = sqrtf(Dot(dx, dx));
    if (dxd > muS * d)
        dx = Mul(fminf(muD * d / dxd, 1.f), dx);
    xv = Add(xv, dx);
    return true;
}

__global__ void XpbdPreStepKernel(const __grid_constant__ XpbdParams p)
{
    uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < static_cast<uint32_t>(p.nV))
        XpbdPreStepVertex(p, i);
    if (i < static_cast<uint32_t>(p.nT))
        p.lambda[i] = make_float2(0.f, 0.f);
    if (i < static_cast<uint32_t>(p.nCV))
        p.lambdaC[i] = 0.f;
}

// One persistent cooperative launch: `substeps` substeps (1 with contact) of pre-step, iterations x (partitions, contact), velocities.
__global__ void __launch_bounds__(256, 4) XpbdSolveKernel(const __grid_constant__ XpbdParams p)
{
    unsigned int target    = 0;
    uint32_t const gtid    = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t const gstride = gridDim.x * blockDim.x;
    for (int s = 0; s < p.substeps; ++s)
    {
        if (!p.skipPreStep)
        {
            for (uint32_t i = gtid; i < static_cast<uint32_t>(p.nV); i += gstride)
                XpbdPreStepVertex(p, i);
            for (uint32_t i = gtid; i < static_cast<uint32_t>(p.nT); i += gstride)
                p.lambda[i] = make_float2(0.f, 0.f);
            for (uint32_t i = gtid; i < static_cast<uint32_t>(p.nCV); i += gstride)
                p.lambdaC[i] = 0.f;
            GridBarrier(p.barrier, target);
        }
        for (int k = 0; k < p.iterations; ++k)
        {
            for (int q = 0; q < p.nPartitions; ++q)
            {
                uint32_t const ib = __ldg(p.partBegin + q), ie = __ldg(p.partBegin + q + 1);
                for (uint32_t it = ib + gtid; it < ie; it += gstride)
                {
                    uint32_t const cb = __ldg(p.itemBegin + it), ce = __ldg(p.itemBegin + it + 1);
                    XpbdProjectTet(p, cb);
                    for (uint32_t c = cb + 1; c < ce; ++c)  // a cluster's constraints one after the other: they share vertices
                    {
                        __threadfence_block();
                        XpbdProjectTet(p, c);
                    }
                }
                GridBarrier(p.barrier, target);
            }
            if (p.nCV > 0)
            {
                // ProjectCollisionConstraints (gpu/impl/xpbd/Integrator.cu:342-448): every active vertex against its nearest
                // triangles from the SAME x; results go to xb and are copied back after all of them are done
                uint32_t const na = *reinterpret_cast<volatile const uint32_t*>(p.nActive);
                for (uint32_t c = gtid; c < na; c += gstride)
                {
                    int const v = __ldcg(p.av + c), i = __ldg(p.V + v);
                    float4 const x4 = __ldcg(p.x + i);
                    float3 xv = F3(x4);
                    float3 const xvt = F3(__ldcg(p.xt + i));
                    float const atc = __ldg(p.alphaC + v) / p.sdt2, gc = atc * __ldg(p.betaC + v) * p.sdt, muc = __ldg(p.muV + v);
                    float lam = __ldcg(p.lambdaC + v);
                    for (int kk = 0; kk < kMaxContacts; ++kk)
                    {
                        int const f = __ldcg(p.nn + static_cast<size_t>(v) * kMaxContacts + kk);
                        if (f < 0)
                            break;
                        int4 const tri = __ldg(p.triF + f);
                        float3 const xf[3]  = {F3(__ldcg(p.x + tri.x)), F3(__ldcg(p.x + tri.y)), F3(__ldcg(p.x + tri.z))};
                        float3 const xft[3] = {F3(__ldcg(p.xt + tri.x)), F3(__ldcg(p.xt + tri.y)), F3(__ldcg(p.xt + tri.z))};
                        float l = lam;
                        if (XpbdProjectVertexTriangle(x4.w, xvt, xft, xf, muc, p.muS, p.muD, atc, gc, l, xv))
                            lam = l;
                    }
                    p.lambdaC[v] = lam;
                    p.xb[c]      = make_float4(xv.x, xv.y, xv.z, x4.w);
                }
                GridBarrier(p.barrier, target);
                for (uint32_t c = gtid; c < na; c += gstride)
                    p.x[__ldg(p.V + __ldcg(p.av + c))] = __ldcg(p.xb + c);
                GridBarrier(p.barrier, target);
            }
        }
        // IntegrateVelocity (sim/xpbd/Kernels.h:262-267)
        for (uint32_t i = gtid; i < static_cast<uint32_t>(p.nV); i += gstride)
        {
            float4 const x = __ldcg(p.x + i), xt = __ldcg(p.xt + i);
            p.vel[i] = make_float4((x.x - xt.x) / p.sdt, (x.y - xt.y) / p.sdt, (x.z - xt.z) / p.sdt, 0.f);
        }
        if (s + 1 < p.substeps)
            GridBarrier(p.barrier, target);
    }
}

}  // namespace vbdx
