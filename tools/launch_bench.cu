// Micro-benchmark: what a per-batch kernel launch costs on this box as a function of parameter size,
// cluster attribute, and how the host learns the kernel is done.  Build: nvcc -arch=sm_100a -O3 -o launch_bench launch_bench.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <chrono>
#include <cstdio>
#include <cstring>
namespace cg = cooperative_groups;
template <int N> struct Blob { unsigned char b[N]; };
template <int N> __global__ void k_params(const __grid_constant__ Blob<N> p, volatile int *flag, int seq)
{
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) { *flag = seq + p.b[(seq * 37) % N]; }
}
__global__ void k_zero_copy(const int *hostIn, volatile int *flag, int seq)
{
    if (threadIdx.x == 0) { int v = hostIn[blockIdx.y * 12]; if (blockIdx.x == 0 && blockIdx.y == 0) *flag = seq + v; }
}
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
template <int N> static void run(const char *name, int cluster, bool spinFlag, int gx, int gy)
{
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    int *hflag; cudaHostAlloc(&hflag, 64, cudaHostAllocMapped); *hflag = -1;
    Blob<N> blob; memset(&blob, 0, sizeof(blob));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchConfig_t cfg = cudaLaunchConfig_t();
    cfg.gridDim = dim3(gx, gy, 1); cfg.blockDim = dim3(256, 1, 1); cfg.stream = st; cfg.dynamicSmemBytes = 40000;
    cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = cluster > 0 ? 1 : 0;
    cudaFuncSetAttribute(k_params<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    const int iters = 2000;
    double evms = 0;
    for (int rep = 0; rep < 2; ++rep)
    {
        double t0 = now();
        for (int i = 0; i < iters; ++i)
        {
            if (rep == 1) cudaEventRecord(e0, st);
            cudaLaunchKernelEx(&cfg, k_params<N>, blob, (volatile int*)hflag, i);
            if (rep == 1) cudaEventRecord(e1, st);
            if (spinFlag) { while (*(volatile int*)hflag != i) { } }
            else cudaStreamSynchronize(st);
            if (rep == 1) { cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); evms += ms; }
        }
        double dt = now() - t0;
        if (rep == 0) printf("%-34s params %6d B cluster %d grid %dx%d %s: %.2f us per launch+wait", name, N, cluster, gx, gy, spinFlag ? "spin-on-host-flag" : "streamSync", dt / iters * 1e6);
        else printf("   | with events: %.2f us wall, %.2f us event time\n", dt / iters * 1e6, evms / iters * 1e3);
    }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(e));
}
int main()
{
    cudaSetDevice(0); cudaFree(0);
    run<64>("small params", 0, false, 2, 150);
    run<64>("small params", 0, true, 2, 150);
    run<64>("small params + cluster attr", 2, false, 2, 150);
    run<64>("small params + cluster attr", 2, true, 2, 150);
    run<4000>("4 KB params", 0, true, 2, 150);
    run<20000>("20 KB params", 0, false, 2, 150);
    run<20000>("20 KB params", 0, true, 2, 150);
    run<20000>("20 KB params + cluster", 2, true, 2, 150);
    run<20000>("20 KB params + cluster4", 4, true, 4, 75);
    run<64>("small params 1x1 grid", 0, true, 1, 1);
    // zero-copy read of proposals from pinned host memory
    {
        cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        int *hflag; cudaHostAlloc(&hflag, 64, cudaHostAllocMapped); *hflag = -1;
        int *hin; cudaHostAlloc(&hin, 20000, cudaHostAllocMapped); memset(hin, 0, 20000);
        const int iters = 2000; double t0 = now();
        for (int i = 0; i < iters; ++i) { k_zero_copy<<<dim3(2, 150), 256, 0, st>>>(hin, hflag, i); while (*(volatile int*)hflag != i) { } }
        printf("zero-copy proposal read, spin: %.2f us per launch+wait\n", (now() - t0) / iters * 1e6);
    }
    return 0;
}
