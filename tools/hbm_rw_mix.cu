// HBM throughput for different read:write mixes with plain coalesced 128-bit accesses (what a streaming kernel
// can expect): 1R:1W copy, 2R:1W (the vertical sweep's mix), 1R:2W (the left-to-right scan's mix), write only.
#include <cstdio>
#include <cuda_runtime.h>

template <int NR, int NW>
__global__ void k(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ c, uint4* __restrict__ d, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = make_uint4(1, 2, 3, 4);
        if (NR >= 1) v = a[i];
        if (NR >= 2) { const uint4 w = b[i]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
        if (NW >= 1) c[i] = v;
        if (NW >= 2) { v.x ^= 1; d[i] = v; }
    }
}

template <int NR, int NW>
void run(const char* name, uint4* a, uint4* b, uint4* c, uint4* d, size_t n)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) k<NR, NW><<<148 * 16, 256>>>(a, b, c, d, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-12s %7.1f GB/s\n", name, 5.0 * (NR + NW) * n * 16 / (ms * 1e-3) / 1e9);
}

int main()
{
    const size_t n = (size_t)3 << 30 >> 4;      // 3 GiB per array
    uint4 *a, *b, *c, *d;
    cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMalloc(&c, n * 16); cudaMalloc(&d, n * 16);
    cudaMemset(a, 1, n * 16); cudaMemset(b, 2, n * 16);
    run<1, 1>("1R:1W", a, b, c, d, n);
    run<2, 1>("2R:1W", a, b, c, d, n);
    run<1, 2>("1R:2W", a, b, c, d, n);
    run<0, 1>("0R:1W", a, b, c, d, n);
    run<1, 0>("1R:0W", a, b, c, d, n);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
