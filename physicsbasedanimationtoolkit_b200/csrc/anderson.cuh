// Anderson acceleration of the VBD fixed-point iteration on the device.
// Behaviour: AndersonIntegrator::Solve (sim/vbd/AndersonIntegrator.cpp:24-58; GPU twin gpu/impl/vbd/AndersonIntegrator.cu:36-105):
//   x^{k-1} = x;  sweep;  G^k = x;  F^k = G^k - x^{k-1};
//   DG[:, (k-1) mod m] = G^k - G^{k-1};  DF[:, (k-1) mod m] = F^k - F^{k-1};
//   alpha = argmin |DF[:, :mk] alpha - F^k|  (mk = min(m, k));   x = G^k - DG[:, :mk] alpha
// The reference solves the least-squares problem with a rank-revealing orthogonal decomposition of the 3nV x mk
// window (Eigen CompleteOrthogonalDecomposition, threshold 1e-10; cuSOLVER QR on its GPU path).  Here the window never
// leaves the GPU and no dense factorisation of a tall matrix is needed: the mk x mk Gram matrix DF^T DF and DF^T F^k
// are accumulated in double precision by the same kernel that updates the window (only one row/column of the Gram
// matrix changes per iteration), and ONE thread factorises it by Cholesky with diagonal pivoting -- the pivots are the
// squares of the R diagonal of the column-pivoted QR the reference computes, so the same rank rule applies
// (pivot_j > threshold^2 * pivot_0, floored at what double precision resolves) -- followed by the minimum-norm
// solution of the rank-truncated system, as the complete orthogonal decomposition returns.
#pragma once

#include "diagnostics.cuh"

namespace vbdx {

constexpr int kMaxAndersonWindow = 16;

struct AndersonView {
    int64_t n;        // vertices (internal order)
    int m;            // window size
    float4* pos;      // current iterate
    float4* xkm1;     // iterate before the sweep
    float4* Gkm1;     // previous sweep result
    float4* Fkm1;     // previous residual
    float4* Fk;       // residual
    float4* DF;       // m columns of n
    float4* DG;       // m columns of n
    double* gram;     // m x m (persistent), row-major
    double* scratch;  // [0, m): DF[:, c] . DF[:, dkl];  [m, 2m): DF[:, c] . F^k
    double* alpha;    // m mixing weights
};

__device__ __forceinline__ float4 Sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, 0.f); }
__device__ __forceinline__ double Dot3(float4 a, float4 b)
{
    return static_cast<double>(a.x) * b.x + static_cast<double>(a.y) * b.y + static_cast<double>(a.z) * b.z;
}

// after the first sweep: G^0 = x, F^0 = G^0 - x^{-1}
__global__ void AndersonFirst(AndersonView a)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= a.n)
        return;
    float4 const x = a.pos[i];
    a.Gkm1[i]      = x;
    a.Fkm1[i]      = Sub(x, a.xkm1[i]);
}

// window update of iteration k (column dkl) fused with the Gram / right-hand-side accumulation
template <int kMaxCols>
__global__ void AndersonWindow(AndersonView a, int dkl, int mk)
{
    __shared__ double smem[32];
    double g[kMaxCols], r[kMaxCols];
    for (int c = 0; c < kMaxCols; ++c)
        g[c] = r[c] = 0.0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    {
        float4 const x  = a.pos[i];
        float4 const Fk = Sub(x, a.xkm1[i]);
        float4 const dg = Sub(x, a.Gkm1[i]);
        float4 const df = Sub(Fk, a.Fkm1[i]);
        a.DG[static_cast<int64_t>(dkl) * a.n + i] = dg;
        a.DF[static_cast<int64_t>(dkl) * a.n + i] = df;
        a.Gkm1[i]                                 = x;
        a.Fkm1[i]                                 = Fk;
        a.Fk[i]                                   = Fk;
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c)
            if (c < mk)
            {
                float4 const col = c == dkl ? df : a.DF[static_cast<int64_t>(c) * a.n + i];
                g[c] += Dot3(col, df);
                r[c] += Dot3(col, Fk);
            }
    }
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c)
        if (c < mk)
        {
            double const gs = BlockSum(g[c], smem);
            double const rs = BlockSum(r[c], smem);
            if (threadIdx.x == 0)
            {
                atomicAdd(a.scratch + c, gs);
                atomicAdd(a.scratch + a.m + c, rs);
            }
        }
}

// one thread: install the new Gram row/column, pivoted Cholesky, minimum-norm least-squares weights
__global__ void AndersonSolveSmall(AndersonView a, int dkl, int mk, double threshold)
{
    if (blockIdx.x != 0 || threadIdx.x != 0)
        return;
    int const m = a.m;
    for (int c = 0; c < mk; ++c)
        a.gram[c * m + dkl] = a.gram[dkl * m + c] = a.scratch[c];
    double A[kMaxAndersonWindow][kMaxAndersonWindow], R[kMaxAndersonWindow][kMaxAndersonWindow], b[kMaxAndersonWindow];
    int perm[kMaxAndersonWindow];
    for (int i = 0; i < mk; ++i)
    {
        perm[i] = i;
        b[i]    = a.scratch[m + i];
        for (int j = 0; j < mk; ++j)
            A[i][j] = a.gram[i * m + j], R[i][j] = 0.0;
    }
    // Cholesky with diagonal pivoting: P^T A P = R^T R, R upper trapezoidal (rank x mk)
    int rank      = 0;
    double pivot0 = 0.0;
    double const rel = fmax(threshold * threshold, 1e-13);  // pivots are squared column norms; double resolves ~1e-16
    for (int k = 0; k < mk; ++k)
    {
        int best = k;
        for (int j = k + 1; j < mk; ++j)
            if (A[j][j] > A[best][best])
                best = j;
        if (best != k)
        {
            for (int j = 0; j < mk; ++j)
            {
                double t = A[k][j];
                A[k][j] = A[best][j], A[best][j] = t;
            }
            for (int i = 0; i < mk; ++i)
            {
                double t = A[i][k];
                A[i][k] = A[i][best], A[i][best] = t;
            }
            for (int i = 0; i < rank; ++i)
            {
                double t = R[i][k];
                R[i][k] = R[i][best], R[i][best] = t;
            }
            int t = perm[k];
            perm[k] = perm[best], perm[best] = t;
            double tb = b[k];
            b[k] = b[best], b[best] = tb;
        }
        double const piv = A[k][k];
        if (k == 0)
            pivot0 = piv;
        if (!(piv > rel * pivot0) || !(piv > 0.0))
            break;
        double const d = sqrt(piv);
        R[k][k]        = d;
        for (int j = k + 1; j < mk; ++j)
            R[k][j] = A[k][j] / d;
        for (int i = k + 1; i < mk; ++i)
            for (int j = k + 1; j < mk; ++j)
                A[i][j] -= R[k][i] * R[k][j];
        ++rank;
    }
    double z[kMaxAndersonWindow], c[kMaxAndersonWindow];
    for (int i = 0; i < mk; ++i)
        z[i] = 0.0;
    if (rank > 0)
    {
        // R11^T c = (P^T b)[0:rank]
        for (int i = 0; i < rank; ++i)
        {
            double s = b[i];
            for (int j = 0; j < i; ++j)
                s -= R[j][i] * c[j];
            c[i] = s / R[i][i];
        }
        if (rank == mk)
        {
            for (int i = mk - 1; i >= 0; --i)
            {
                double s = c[i];
                for (int j = i + 1; j < mk; ++j)
                    s -= R[i][j] * z[j];
                z[i] = s / R[i][i];
            }
        }
        else
        {
            // minimum-norm z with W z = c, W = R[0:rank, 0:mk]:  z = W^T (W W^T)^{-1} c  (W W^T: small SPD, plain Cholesky)
            double S[kMaxAndersonWindow][kMaxAndersonWindow], u[kMaxAndersonWindow];
            for (int i = 0; i < rank; ++i)
                for (int j = 0; j < rank; ++j)
                {
                    double s = 0;
                    for (int l = 0; l < mk; ++l)
                        s += R[i][l] * R[j][l];
                    S[i][j] = s;
                }
            for (int k = 0; k < rank; ++k)
            {
                S[k][k] = sqrt(S[k][k]);
                for (int i = k + 1; i < rank; ++i)
                    S[i][k] /= S[k][k];
                for (int j = k + 1; j < rank; ++j)
                    for (int i = j; i < rank; ++i)
                        S[i][j] -= S[i][k] * S[j][k];
            }
            for (int i = 0; i < rank; ++i)
            {
                double s = c[i];
                for (int j = 0; j < i; ++j)
                    s -= S[i][j] * u[j];
                u[i] = s / S[i][i];
            }
            for (int i = rank - 1; i >= 0; --i)
            {
                double s = u[i];
                for (int j = i + 1; j < rank; ++j)
                    s -= S[j][i] * u[j];
                u[i] = s / S[i][i];
            }
            for (int l = 0; l < mk; ++l)
                for (int i = 0; i < rank; ++i)
                    z[l] += R[i][l] * u[i];
        }
    }
    for (int i = 0; i < m; ++i)
        a.alpha[i] = 0.0;
    for (int i = 0; i < mk; ++i)
        a.alpha[perm[i]] = z[i];
    for (int i = 0; i < 2 * m; ++i)
        a.scratch[i] = 0.0;  // ready for the next accumulation
}

// x = G^k - DG alpha; the result is also the next iteration's x^{k-1} (and, with contact, the snapshot the next
// sweep reads same-colour triangle corners from: snapNext, step_kernel.cuh)
__global__ void AndersonApply(AndersonView a, int mk, int64_t nActive, float4* snapNext)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= a.n)
        return;
    float4 x = a.pos[i];
    if (i < nActive)  // constrained vertices have identically zero window columns
    {
        double dx = 0, dy = 0, dz = 0;
        for (int c = 0; c < mk; ++c)
        {
            float4 const dg = a.DG[static_cast<int64_t>(c) * a.n + i];
            double const al = a.alpha[c];
            dx += al * dg.x, dy += al * dg.y, dz += al * dg.z;
        }
        x.x = static_cast<float>(x.x - dx), x.y = static_cast<float>(x.y - dy), x.z = static_cast<float>(x.z - dz);
        a.pos[i] = x;
        if (snapNext != nullptr)
            snapNext[i] = x;
    }
    a.xkm1[i] = x;
}

}  // namespace vbdx
