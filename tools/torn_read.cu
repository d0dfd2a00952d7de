// Evidence for the 16-byte assumption of the barrier-free sweep and the halo exchange: a position and the
// number of its write travel as ONE 16-byte datum and readers compare only the tag.  A writer GPU hammers one slot per
// lane in the READER's memory over NVLink with (i, i, i, i), i = 1, 2, ...; the reader polls the slots the way the
// kernels do -- ld.relaxed.sys.b128, and cp.async 16 into shared memory -- and counts reads whose four words differ
// (a torn read).  Variants: the writer stores with st.relaxed.sys.b128 (what the product does) or with a plain st.v4.
//   nvcc -arch=sm_100a -o torn_read tools/torn_read.cu && ./torn_read [writer device] [reader device] [millions of writes]
// With one GPU both kernels run on it (st/ld .gpu paths through the same L2).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                      \
    do                                                                             \
    {                                                                              \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess)                                                     \
        {                                                                          \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));          \
            std::exit(2);                                                          \
        }                                                                          \
    } while (0)

__device__ __forceinline__ void StoreB128(float4* p, unsigned v)
{
    unsigned long long const w = (unsigned long long)v | ((unsigned long long)v << 32);
    asm volatile("{\n .reg .b128 t;\n mov.b128 t, {%1, %2};\n st.relaxed.sys.global.b128 [%0], t;\n}\n" ::"l"(p), "l"(w), "l"(w) : "memory");
}
__device__ __forceinline__ uint4 LoadB128(const float4* p)
{
    unsigned long long lo, hi;
    asm volatile("{\n .reg .b128 t;\n ld.relaxed.sys.global.b128 t, [%2];\n mov.b128 {%0, %1}, t;\n}\n" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
    return make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)hi, (unsigned)(hi >> 32));
}

__global__ void Writer(float4* slots, unsigned n, int plain, volatile unsigned* stop)
{
    float4* p = slots + blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned i = 1; i <= n; ++i)
    {
        if (plain)
            asm volatile("st.volatile.global.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(p), "r"(i) : "memory");  // four scalar accesses in the model
        else
            StoreB128(p, i);
    }
    __threadfence_system();
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *stop = 1u;
}

__global__ void Reader(const float4* slots, int viaCpAsync, volatile unsigned* stop, unsigned long long* torn, unsigned long long* reads,
                       unsigned long long* changes)
{
    __shared__ uint4 stage[256];
    const float4* p = slots + blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long t = 0, r = 0, c = 0;
    unsigned last = 0;
    while (*stop == 0u)
    {
        uint4 v;
        if (viaCpAsync)
        {
            unsigned const dst = (unsigned)__cvta_generic_to_shared(stage + threadIdx.x);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(p) : "memory");
            asm volatile("cp.async.wait_all;" ::: "memory");
            v = stage[threadIdx.x];
        }
        else
            v = LoadB128(p);
        ++r;
        t += !(v.x == v.y && v.y == v.z && v.z == v.w);
        c += v.w != last;
        last = v.w;
    }
    atomicAdd(torn, t), atomicAdd(reads, r), atomicAdd(changes, c);
}

int main(int argc, char** argv)
{
    int nDev = 0;
    CK(cudaGetDeviceCount(&nDev));
    int const wdev = argc > 1 ? std::atoi(argv[1]) : 0, rdev = argc > 2 ? std::atoi(argv[2]) : (nDev > 1 ? 1 : 0);
    unsigned const n = (argc > 3 ? std::atoi(argv[3]) : 2) * 1000000u;
    int const blocks = 8, threads = 256;
    CK(cudaSetDevice(rdev));
    float4* slots;
    unsigned* stop;
    unsigned long long* counters;
    CK(cudaMalloc(&slots, sizeof(float4) * blocks * threads));
    CK(cudaMallocHost(&stop, sizeof(unsigned)));
    CK(cudaMallocManaged(&counters, 3 * sizeof(unsigned long long)));
    if (wdev != rdev)
    {
        int ok = 0;
        CK(cudaDeviceCanAccessPeer(&ok, wdev, rdev));
        if (!ok)
        {
            std::printf("no peer access between devices %d and %d\n", wdev, rdev);
            return 0;
        }
        CK(cudaSetDevice(wdev));
        CK(cudaDeviceEnablePeerAccess(rdev, 0));
    }
    cudaStream_t sw, sr;
    CK(cudaSetDevice(wdev));
    CK(cudaStreamCreate(&sw));
    CK(cudaSetDevice(rdev));
    CK(cudaStreamCreate(&sr));
    int bad = 0;
    for (int plain = 0; plain < 2; ++plain)
        for (int via = 0; via < 2; ++via)
        {
            CK(cudaSetDevice(rdev));
            CK(cudaMemset(slots, 0, sizeof(float4) * blocks * threads));
            counters[0] = counters[1] = counters[2] = 0;
            *stop = 0;
            CK(cudaDeviceSynchronize());
            Reader<<<blocks, threads, 0, sr>>>(slots, via, stop, counters, counters + 1, counters + 2);
            CK(cudaSetDevice(wdev));
            Writer<<<blocks, threads, 0, sw>>>(slots, n, plain, stop);
            CK(cudaStreamSynchronize(sw));
            CK(cudaSetDevice(rdev));
            CK(cudaStreamSynchronize(sr));
            std::printf("writer GPU %d %-24s reader GPU %d %-22s: %llu reads, %llu saw a new value, %llu torn\n", wdev,
                        plain ? "st.volatile.v4.u32" : "st.relaxed.sys.b128", rdev, via ? "cp.async.cg 16" : "ld.relaxed.sys.b128", counters[1],
                        counters[2], counters[0]);
            if (!plain && counters[0] != 0)
                bad = 1;
        }
    return bad;
}
