// Device kernels of the two accelerators that wrap the sweep with whole-array updates:
//   Nesterov      NesterovIntegrator::Solve (sim/vbd/NesterovIntegrator.cpp:18-44)
//   trust region  TrustRegionIntegrator::SolveWithLinearAcceleratedPath (gpu/impl/vbd/TrustRegionIntegrator.cu:47-166):
//                 UpdateIterates (:366-388), SquaredStepSize (:481-497), TakeLinearStep (:499-515), RollbackLinearStep (:517-532)
// The sweeps in between are one-iteration launches of the persistent step kernel; the scalar recurrences (lambda/beta;
// parabola fit, clamp, accept/reject, radius) run on the host in double, like the reference's host code does.
// Arrays are float4 per vertex in internal order; `n` covers every vertex (constrained ones never move, so all of
// these updates leave them where they are).
#pragma once

#include "anderson.cuh"

namespace vbdx {

// y^k = x + beta (x - x^{k-1})    (x itself is NOT moved: the sweep starts from x, NesterovIntegrator.cpp:31-35)
__global__ void NesterovExtrapolate(int64_t n, const float4* pos, const float4* xkm1, float4* yk, float beta)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    float4 const x = pos[i], o = xkm1[i];
    yk[i] = make_float4(fmaf(beta, x.x - o.x, x.x), fmaf(beta, x.y - o.y, x.y), fmaf(beta, x.z - o.z, x.z), 0.f);
}

// x = y^k - alpha (x_swept - x^{k-1})    (NesterovIntegrator.cpp:36-42); also the snapshot a contact sweep reads next
__global__ void NesterovCorrect(int64_t n, float4* pos, const float4* xkm1, const float4* yk, float alpha, int64_t nActive, float4* snapNext)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nActive || i >= n)
        return;
    float4 x       = pos[i];
    float4 const o = xkm1[i], y = yk[i];
    x.x = fmaf(-alpha, x.x - o.x, y.x), x.y = fmaf(-alpha, x.y - o.y, y.y), x.z = fmaf(-alpha, x.z - o.z, y.z);
    pos[i] = x;
    if (snapNext != nullptr)
        snapNext[i] = x;
}

// out[0] += sum_i |x_i - x^{k-1}_i|^2   (double accumulation)
__global__ void SquaredStepSize(int64_t n, const float4* pos, const float4* xkm1, double* out)
{
    __shared__ double smem[32];
    double acc = 0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    {
        float4 const d = Sub(pos[i], xkm1[i]);
        acc += Dot3(d, d);
    }
    double const r = BlockSum(acc, smem);
    if (threadIdx.x == 0)
        atomicAdd(out, r);
}

// x = x^{k-1} + s (x - x^{k-1}):  s = t takes the accelerated step, s = 1/t rolls it back
__global__ void ScaleStep(int64_t n, float4* pos, const float4* xkm1, float s, int64_t nActive, float4* snapNext)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nActive || i >= n)
        return;
    float4 x       = pos[i];
    float4 const o = xkm1[i];
    x.x = fmaf(s, x.x - o.x, o.x), x.y = fmaf(s, x.y - o.y, o.y), x.z = fmaf(s, x.z - o.z, o.z);
    pos[i] = x;
    if (snapNext != nullptr)
        snapNext[i] = x;
}

}  // namespace vbdx
