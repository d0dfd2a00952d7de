// Broyden ("limited-memory good Broyden") acceleration of the VBD iteration on the device.
// Behaviour: BroydenIntegrator::Solve (sim/vbd/BroydenIntegrator.cpp:41-77):
//   x^{k-1} = x;  sweep;  vbd(f_{k-1}) = x^{k-1} - x
//   for k = 1 .. iterations-1, column c = (k-1) mod m, mk = min(m, k):
//     X[:, c] = x - x^{k-1};  x^{k-1} = x;  sweep;  vbd(f_k) = x^{k-1} - x
//     GF[:, c] = vbd(f_k) - vbd(f_{k-1});  vbd(f_{k-1}) = vbd(f_k)
//     gamma = LSCG(GF[:, :mk], vbd(f_k); at most max(1, m-k) updates, tolerance 1e-10)
//     x -= (X[:, :mk] - GF[:, :mk]) gamma
// The reference hands the tall 3nV x mk window to Eigen::LeastSquaresConjugateGradient (identity preconditioner), i.e.
// conjugate gradients on the normal equations.  Every quantity that iteration looks at is a function of the mk x mk Gram
// matrix GF^T GF and of GF^T vbd(f_k):  A^T r = b - G y,  |A p|^2 = p^T G p.  So, as for Anderson (anderson.cuh), the
// window never leaves the GPU: the kernel that updates the window accumulates the one changed Gram row/column and the
// right-hand side in double, one thread runs the (at most m-1)-step CG recurrence on m numbers, and one kernel applies
// the correction.  The iteration cap, the zero start, the stopping rule |A^T r|^2 < tol^2 |A^T b|^2 and the update
// order are those of Eigen/src/IterativeLinearSolvers/LeastSquareConjugateGradient.h.
#pragma once

#include "anderson.cuh"

namespace vbdx {

struct BroydenView {
    int64_t n;        // vertices (internal order)
    int m;            // window size
    float4* pos;      // current iterate
    float4* xkm1;     // iterate before the sweep
    float4* fkm1;     // previous vbd step  x^{k-1} - x
    float4* fk;       // vbd step of this iteration
    float4* X;        // m columns of n: past steps
    float4* GF;       // m columns of n: differences of vbd steps
    double* gram;     // m x m (persistent), row-major
    double* scratch;  // [0, m): GF[:, c] . GF[:, col];  [m, 2m): GF[:, c] . vbd(f_k)
    double* gamma;    // m weights
};

// after the first sweep: vbd(f_0) = x^{-1} - x
__global__ void BroydenFirst(BroydenView a)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= a.n)
        return;
    a.fkm1[i] = Sub(a.xkm1[i], a.pos[i]);
}

// before the sweep of iteration k: X[:, col] = x - x^{k-1};  x^{k-1} = x
__global__ void BroydenBeforeSweep(BroydenView a, int col)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= a.n)
        return;
    float4 const x                          = a.pos[i];
    a.X[static_cast<int64_t>(col) * a.n + i] = Sub(x, a.xkm1[i]);
    a.xkm1[i]                               = x;
}

// after the sweep of iteration k: window column `col`, fused with the Gram / right-hand-side accumulation
template <int kMaxCols>
__global__ void BroydenWindow(BroydenView a, int col, int mk)
{
    __shared__ double smem[32];
    double g[kMaxCols], r[kMaxCols];
    for (int c = 0; c < kMaxCols; ++c)
        g[c] = r[c] = 0.0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    {
        float4 const fk = Sub(a.xkm1[i], a.pos[i]);
        float4 const df = Sub(fk, a.fkm1[i]);
        a.GF[static_cast<int64_t>(col) * a.n + i] = df;
        a.fkm1[i]                                 = fk;
        a.fk[i]                                   = fk;
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c)
            if (c < mk)
            {
                float4 const v = c == col ? df : a.GF[static_cast<int64_t>(c) * a.n + i];
                g[c] += Dot3(v, df);
                r[c] += Dot3(v, fk);
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

// one thread: install the new Gram row/column, then Eigen's least-squares CG in the mk-dimensional coefficient space
__global__ void BroydenSolveSmall(BroydenView a, int col, int mk, int maxIters, double tol)
{
    if (blockIdx.x != 0 || threadIdx.x != 0)
        return;
    int const m = a.m;
    for (int c = 0; c < mk; ++c)
        a.gram[c * m + col] = a.gram[col * m + c] = a.scratch[c];
    double y[kMaxAndersonWindow], nr[kMaxAndersonWindow], p[kMaxAndersonWindow], Gp[kMaxAndersonWindow];
    double rhsNorm2 = 0.0;
    for (int i = 0; i < mk; ++i)
    {
        y[i]  = 0.0;
        nr[i] = a.scratch[m + i];  // A^T (b - A 0)
        rhsNorm2 += nr[i] * nr[i];
    }
    double const threshold = tol * tol * rhsNorm2;
    if (rhsNorm2 != 0.0 && !(rhsNorm2 < threshold))
    {
        for (int i = 0; i < mk; ++i)
            p[i] = nr[i];
        double absNew = rhsNorm2;
        for (int it = 0; it < maxIters; ++it)
        {
            double pGp = 0.0;
            for (int i = 0; i < mk; ++i)
            {
                double s = 0.0;
                for (int j = 0; j < mk; ++j)
                    s += a.gram[i * m + j] * p[j];
                Gp[i] = s;
                pGp += p[i] * s;
            }
            double const alpha = absNew / pGp;  // |A p|^2 = p^T G p
            double res2        = 0.0;
            for (int i = 0; i < mk; ++i)
            {
                y[i] += alpha * p[i];
                nr[i] -= alpha * Gp[i];         // A^T (r - alpha A p)
                res2 += nr[i] * nr[i];
            }
            if (res2 < threshold)
                break;
            double const beta = res2 / absNew;
            absNew            = res2;
            for (int i = 0; i < mk; ++i)
                p[i] = nr[i] + beta * p[i];
        }
    }
    for (int i = 0; i < m; ++i)
        a.gamma[i] = i < mk ? y[i] : 0.0;
    for (int i = 0; i < 2 * m; ++i)
        a.scratch[i] = 0.0;  // ready for the next accumulation
}

// x -= (X - GF) gamma  (snapNext: see AndersonApply)
__global__ void BroydenApply(BroydenView a, int mk, int64_t nActive, float4* snapNext)
{
    int64_t const i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nActive)  // constrained vertices have identically zero window columns
        return;
    double dx = 0, dy = 0, dz = 0;
    for (int c = 0; c < mk; ++c)
    {
        float4 const xs = a.X[static_cast<int64_t>(c) * a.n + i];
        float4 const gf = a.GF[static_cast<int64_t>(c) * a.n + i];
        double const ga = a.gamma[c];
        dx += ga * (static_cast<double>(xs.x) - gf.x), dy += ga * (static_cast<double>(xs.y) - gf.y), dz += ga * (static_cast<double>(xs.z) - gf.z);
    }
    float4 x = a.pos[i];
    x.x = static_cast<float>(x.x - dx), x.y = static_cast<float>(x.y - dy), x.z = static_cast<float>(x.z - dz);
    a.pos[i] = x;
    if (snapNext != nullptr)
        snapNext[i] = x;
}

}  // namespace vbdx
