// Backward-Euler objective and its gradient on the device, in double precision:
//   f(x) = 1/2 (x - xtilde)^T M (x - xtilde) + dt^2 sum_e wg_e psi(F_e(x))
// Integrator::ObjectiveFunction / ObjectiveFunctionGradient (sim/vbd/Integrator.cpp:138-200) with the Stable
// Neo-Hookean density of physics/StableNeoHookeanEnergy.h:738-755,
//   psi = mu/2 (|F|^2 - 3) + lambda/2 (det F - 1 - mu/lambda)^2,
// or the St. Venant-Kirchhoff density of physics/SaintVenantKirchhoffEnergy.h,
//   psi = mu tr(E^2) + lambda/2 tr(E)^2,  E = (F^T F - I)/2.
// These feed the iterate traces (TraceNextStep / ExportTrace, sim/vbd/Integrator.cpp:47-52,202-235) and the
// convergence checks of the reference's tests; they are not on the step's hot path.
// Inputs are 3 x nV column-major double arrays in the CALLER's vertex order.
#pragma once

#include <cstdint>

#include <cuda_runtime.h>

namespace vbdx {

__device__ __forceinline__ double BlockSum(double v, double* smem)
{
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
        smem[warp] = v;
    __syncthreads();
    double r = 0;
    if (warp == 0)
    {
        r = lane < static_cast<int>(blockDim.x >> 5) ? smem[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1)
            r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;  // valid in thread 0
}

// out[0] += 1/2 sum_i m_i |x_i - xtilde_i|^2;  grad (optional) = m_i (x_i - xtilde_i)
__global__ void ObjectiveKinetic(const double* x, const double* xtilde, const double* m, int64_t nV, double* out, double* grad)
{
    __shared__ double smem[32];
    double acc = 0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nV; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        for (int d = 0; d < 3; ++d)
        {
            double const dx = x[3 * i + d] - xtilde[3 * i + d];
            acc += 0.5 * m[i] * dx * dx;
            if (grad)
                grad[3 * i + d] = m[i] * dx;
        }
    double const r = BlockSum(acc, smem);
    if (threadIdx.x == 0)
        atomicAdd(out, r);
}

// F = (x_e - x_0) Jinv with Jinv row a = grad N_{a+1} (setup_kernels.cuh: ElementQuantities)
__device__ __forceinline__ void ElementF(const double* x, const int32_t* t, const double* Ji, double F[3][3])
{
    double D[3][3];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r)
            D[r][c] = x[3 * static_cast<int64_t>(t[c + 1]) + r] - x[3 * static_cast<int64_t>(t[0]) + r];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            F[r][c] = D[r][0] * Ji[c] + D[r][1] * Ji[3 + c] + D[r][2] * Ji[6 + c];
}

// out[0] += dt2 sum_e wg_e psi_e;  grad (optional, must hold the kinetic part already) += dt2 wg_e dpsi/dx
__global__ void ObjectiveElastic(
    const double* x,
    const int32_t* E,
    const double* Jinv,
    const double* vol,
    const double* lame,  // 2 x nT or null (mu0, lambda0)
    double mu0,
    double lambda0,
    int64_t nT,
    double dt2,
    int stvk,
    double* out,
    double* grad)
{
    __shared__ double smem[32];
    double acc = 0;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < nT; e += static_cast<int64_t>(gridDim.x) * blockDim.x)
    {
        int32_t t[4];
        for (int a = 0; a < 4; ++a)
            t[a] = E[4 * e + a];
        double const* Ji = Jinv + 9 * e;
        double F[3][3];
        ElementF(x, t, Ji, F);
        double const mu = lame ? lame[2 * e] : mu0, lam = lame ? lame[2 * e + 1] : lambda0;
        // cofactor matrix C (dJ/dF) and J
        double C[3][3];
        C[0][0] = F[1][1] * F[2][2] - F[1][2] * F[2][1];
        C[0][1] = F[1][2] * F[2][0] - F[1][0] * F[2][2];
        C[0][2] = F[1][0] * F[2][1] - F[1][1] * F[2][0];
        C[1][0] = F[0][2] * F[2][1] - F[0][1] * F[2][2];
        C[1][1] = F[0][0] * F[2][2] - F[0][2] * F[2][0];
        C[1][2] = F[0][1] * F[2][0] - F[0][0] * F[2][1];
        C[2][0] = F[0][1] * F[1][2] - F[0][2] * F[1][1];
        C[2][1] = F[0][2] * F[1][0] - F[0][0] * F[1][2];
        C[2][2] = F[0][0] * F[1][1] - F[0][1] * F[1][0];
        double const J = F[0][0] * C[0][0] + F[0][1] * C[0][1] + F[0][2] * C[0][2];
        double I2      = 0;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                I2 += F[r][c] * F[r][c];
        double const d = J - 1.0 - mu / lam;
        double const w = vol[e];
        // first Piola-Kirchhoff stress P = dpsi/dF
        double P[3][3];
        if (!stvk)
        {
            acc += w * (0.5 * lam * d * d + 0.5 * mu * (I2 - 3.0));
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    P[r][c] = mu * F[r][c] + lam * d * C[r][c];
        }
        else
        {
            double Eg[3][3], trE = 0, E2 = 0;
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                {
                    Eg[r][c] = 0.5 * (F[0][r] * F[0][c] + F[1][r] * F[1][c] + F[2][r] * F[2][c] - (r == c ? 1.0 : 0.0));
                    E2 += Eg[r][c] * Eg[r][c];
                }
            trE = Eg[0][0] + Eg[1][1] + Eg[2][2];
            acc += w * (mu * E2 + 0.5 * lam * trE * trE);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                {
                    double s = 0;  // (F S)_{rc},  S = 2 mu E + lambda tr(E) I
                    for (int k = 0; k < 3; ++k)
                        s += F[r][k] * (2.0 * mu * Eg[k][c] + (k == c ? lam * trE : 0.0));
                    P[r][c] = s;
                }
        }
        if (grad)
        {
            // g_a = P grad N_a, grad N_0 = -(grad N_1 + grad N_2 + grad N_3)
            double g0[3] = {0, 0, 0};
            for (int a = 0; a < 3; ++a)
                for (int r = 0; r < 3; ++r)
                {
                    double ga = 0;
                    for (int c = 0; c < 3; ++c)
                        ga += P[r][c] * Ji[3 * a + c];
                    ga *= dt2 * w;
                    atomicAdd(grad + 3 * static_cast<int64_t>(t[a + 1]) + r, ga);
                    g0[r] -= ga;
                }
            for (int r = 0; r < 3; ++r)
                atomicAdd(grad + 3 * static_cast<int64_t>(t[0]) + r, g0[r]);
        }
    }
    double const r = BlockSum(acc * dt2, smem);
    if (threadIdx.x == 0)
        atomicAdd(out, r);
}

}  // namespace vbdx
