// TEST INFRASTRUCTURE (never linked into the product): runs the REFERENCE's own vertex-triangle contact code on the GPU,
// compiled from the headers where they lie under /root/reference --
//   pbat::sim::vbd::kernels::AccumulateVertexTriangleContact     sim/vbd/Kernels.h:223-302
//   pbat::gpu::impl::vbd::kernels::ContactPenalty<8>             gpu/impl/vbd/Kernels.cuh:80-114
// -- so that csrc/contact.cuh (scalar fp32 restatement inside the sweep's epilogue) can be compared with the function it
// restates on the same inputs (tests/test_gpu_contact.py).  The function is PBAT_HOST_DEVICE but cannot be called on the
// host (its expression templates dangle under g++: DESIGN.md section 7); on the device it is what the reference runs.
// Built by `make -C oracle contact_ref` into oracle/_ref/libcontact_ref.so (needs /root/reference; the built file travels).
#include "pbat/gpu/impl/vbd/Kernels.cuh"

#include <cstdint>
#include <cuda_runtime.h>

namespace mini = pbat::math::linalg::mini;

// in: 28 floats per pair = xtv(3) xv(3) xtf(3x3, column = triangle vertex) xf(3x3) dt k muF epsv
// out: 13 floats per pair = E, g(3), H(3x3 column-major)
__global__ void ContactPairs(int n, const float* in, float* out)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const float* p = in + 28 * i;
    mini::SVector<float, 3> xtv, xv;
    mini::SMatrix<float, 3, 3> xtf, xf;
    for (int r = 0; r < 3; ++r)
    {
        xtv(r) = p[r];
        xv(r)  = p[3 + r];
        for (int c = 0; c < 3; ++c)
        {
            xtf(r, c) = p[6 + 3 * c + r];
            xf(r, c)  = p[15 + 3 * c + r];
        }
    }
    mini::SVector<float, 3> g    = mini::Zeros<float, 3, 1>();
    mini::SMatrix<float, 3, 3> H = mini::Zeros<float, 3, 3>();
    float const E = pbat::sim::vbd::kernels::AccumulateVertexTriangleContact(xtv, xv, xtf, xf, p[24], p[25], p[26], p[27], &g, &H);
    float* o      = out + 13 * i;
    o[0]          = E;
    for (int r = 0; r < 3; ++r)
    {
        o[1 + r] = g(r);
        for (int c = 0; c < 3; ++c)
            o[4 + 3 * c + r] = H(r, c);
    }
}

// per vertex i: number of contacts and the 8 area-scaled penalties ContactPenalty<8> yields
__global__ void Penalties(int nVerts, int* fc, float* XVA, float* FA, float muC, int* nContacts, float* penalty)
{
    int const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nVerts)
        return;
    pbat::gpu::impl::vbd::kernels::ContactPenalty<8> cp{i, fc, XVA, FA, muC};
    nContacts[i] = cp.nContacts;
    for (int c = 0; c < 8; ++c)
        penalty[8 * i + c] = c < cp.nContacts ? cp.Penalty(c) : 0.f;
}

extern "C" int contact_ref_pairs(int n, const float* in, float* out)
{
    float *di = nullptr, *dout = nullptr;
    if (cudaMalloc(&di, sizeof(float) * 28 * n) != cudaSuccess || cudaMalloc(&dout, sizeof(float) * 13 * n) != cudaSuccess)
        return 1;
    cudaMemcpy(di, in, sizeof(float) * 28 * n, cudaMemcpyHostToDevice);
    ContactPairs<<<(n + 127) / 128, 128>>>(n, di, dout);
    cudaError_t const e = cudaDeviceSynchronize();
    cudaMemcpy(out, dout, sizeof(float) * 13 * n, cudaMemcpyDeviceToHost);
    cudaFree(di), cudaFree(dout);
    return e == cudaSuccess ? 0 : 2;
}

extern "C" int contact_ref_penalties(int nVerts, int nTris, const int* fc, const float* XVA, const float* FA, float muC, int* nContacts, float* penalty)
{
    int *dfc = nullptr, *dn = nullptr;
    float *dx = nullptr, *df = nullptr, *dp = nullptr;
    cudaMalloc(&dfc, sizeof(int) * 8 * nVerts), cudaMalloc(&dn, sizeof(int) * nVerts);
    cudaMalloc(&dx, sizeof(float) * nVerts), cudaMalloc(&df, sizeof(float) * nTris), cudaMalloc(&dp, sizeof(float) * 8 * nVerts);
    cudaMemcpy(dfc, fc, sizeof(int) * 8 * nVerts, cudaMemcpyHostToDevice);
    cudaMemcpy(dx, XVA, sizeof(float) * nVerts, cudaMemcpyHostToDevice);
    cudaMemcpy(df, FA, sizeof(float) * nTris, cudaMemcpyHostToDevice);
    Penalties<<<(nVerts + 127) / 128, 128>>>(nVerts, dfc, dx, df, muC, dn, dp);
    cudaError_t const e = cudaDeviceSynchronize();
    cudaMemcpy(nContacts, dn, sizeof(int) * nVerts, cudaMemcpyDeviceToHost);
    cudaMemcpy(penalty, dp, sizeof(float) * 8 * nVerts, cudaMemcpyDeviceToHost);
    cudaFree(dfc), cudaFree(dn), cudaFree(dx), cudaFree(df), cudaFree(dp);
    return e == cudaSuccess ? 0 : 2;
}
