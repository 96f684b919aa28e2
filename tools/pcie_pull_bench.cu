// Micro-benchmark: how fast can a running kernel pull a ~7 KB batch out of pinned host memory?
//  (1) 256 lanes, uncached 16-byte loads   (2) one TMA bulk copy host->shared   (3) 150 CTAs, 64 B each
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <stdint.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__global__ void pull_lanes(const uint4 *host, uint4 *dev, int n16, unsigned long long *ns, int vol)
{
    __syncthreads();
    unsigned long long t0 = gt();
    for (int i = threadIdx.x; i < n16; i += blockDim.x)
    {
        uint4 v;
        if (vol) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(host + i) : "memory");
        else v = __ldcg(host + i);
        dev[i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) *ns = gt() - t0;
}
__global__ void pull_tma(const void *host, uint4 *dev, int bytes, unsigned long long *ns)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned long long t0 = gt();
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sm)), "l"(host), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    unsigned long long t1 = gt();
    for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) dev[i] = reinterpret_cast<uint4*>(sm)[i];
    __syncthreads();
    if (threadIdx.x == 0) { ns[0] = t1 - t0; ns[1] = gt() - t0; }
}
__global__ void pull_per_cta(const uint4 *host, uint4 *dev, unsigned long long *ns)
{
    unsigned long long t0 = gt();
    if (threadIdx.x < 4)
    {
        uint4 v;
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(host + blockIdx.x * 4 + threadIdx.x) : "memory");
        dev[blockIdx.x * 4 + threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) ns[blockIdx.x] = gt() - t0;
}
int main()
{
    cudaSetDevice(0);
    const int bytes = 7168;
    void *h; cudaHostAlloc(&h, 65536, cudaHostAllocMapped); memset(h, 1, 65536);
    uint4 *d; cudaMalloc(&d, 65536);
    unsigned long long *ns; cudaMallocManaged(&ns, 8 * 1024);
    for (int rep = 0; rep < 3; ++rep)
    {
        pull_lanes<<<1, 256>>>((const uint4*)h, d, bytes / 16, ns, 1); cudaDeviceSynchronize();
        printf("lanes volatile 16B x %d: %.2f us\n", bytes / 16, ns[0] * 1e-3);
        pull_lanes<<<1, 256>>>((const uint4*)h, d, bytes / 16, ns, 0); cudaDeviceSynchronize();
        printf("lanes ld.cg    16B x %d: %.2f us\n", bytes / 16, ns[0] * 1e-3);
        pull_lanes<<<1, 1024>>>((const uint4*)h, d, bytes / 16, ns, 1); cudaDeviceSynchronize();
        printf("1024 lanes volatile: %.2f us\n", ns[0] * 1e-3);
        pull_tma<<<1, 256, 16384>>>(h, d, bytes, ns); cudaDeviceSynchronize();
        printf("TMA bulk %d B host->smem: %.2f us (+store to device %.2f us)  err=%s\n", bytes, ns[0] * 1e-3, ns[1] * 1e-3, cudaGetErrorString(cudaGetLastError()));
        pull_tma<<<1, 256, 16384>>>(h, d, 1024, ns); cudaDeviceSynchronize();
        printf("TMA bulk 1024 B host->smem: %.2f us\n", ns[0] * 1e-3);
        pull_per_cta<<<150, 64>>>((const uint4*)h, d, ns); cudaDeviceSynchronize();
        unsigned long long mx = 0, sum = 0; for (int i = 0; i < 150; ++i) { mx = ns[i] > mx ? ns[i] : mx; sum += ns[i]; }
        printf("150 CTAs x 64 B: mean %.2f us max %.2f us\n", sum / 150.0 * 1e-3, mx * 1e-3);
        pull_per_cta<<<600, 64>>>((const uint4*)h, d, ns); cudaDeviceSynchronize();
        mx = 0; sum = 0; for (int i = 0; i < 600; ++i) { mx = ns[i] > mx ? ns[i] : mx; sum += ns[i]; }
        printf("600 CTAs x 64 B: mean %.2f us max %.2f us\n", sum / 600.0 * 1e-3, mx * 1e-3);
    }
    return 0;
}
