// Probe: min over an 8-lane group, butterfly of 3 shuffles vs one redux.sync with a per-group member mask.
// Checks that the two agree, then times a dependent chain (latency) and many warps (throughput).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned bfly8(unsigned m)
{
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = __vminu2(m, __shfl_xor_sync(0xffffffffu, m, o, 8));
    return __vminu2(m, __byte_perm(m, 0, 0x1032));
}
__device__ __forceinline__ unsigned redux8(unsigned m, unsigned mask)
{
    const unsigned h = min(m & 0xffffu, m >> 16);
    const unsigned r = __reduce_min_sync(mask, h);
    return r | (r << 16);
}

template <int MODE>
__global__ void k(unsigned* out, int iters, long long* cyc)
{
    const unsigned lane = threadIdx.x & 31, mask = 0xffu << (lane & 24);
    unsigned v = (threadIdx.x * 2654435761u + blockIdx.x * 40503u) | 0x00010001u;
    unsigned acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const unsigned m = MODE == 0 ? bfly8(v) : redux8(v, mask);
        acc += m;
        v = (v ^ (m + i)) * 0x9E3779B1u | 0x00010001u;      // dependent chain through the reduction result
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    unsigned *a, *b;
    long long* cyc;
    const int nb = 148 * 8, nt = 256, iters = 4096;
    cudaMalloc(&a, nb * nt * 4); cudaMalloc(&b, nb * nt * 4); cudaMallocManaged(&cyc, 8);
    k<0><<<nb, nt>>>(a, iters, cyc); k<1><<<nb, nt>>>(b, iters, cyc);
    cudaDeviceSynchronize();
    unsigned *ha = new unsigned[nb * nt], *hb = new unsigned[nb * nt];
    cudaMemcpy(ha, a, nb * nt * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb, b, nb * nt * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < nb * nt; ++i) bad += ha[i] != hb[i];
    printf("mismatches: %d of %d (%s)\n", bad, nb * nt, cudaGetErrorString(cudaGetLastError()));
    for (int mode = 0; mode < 2; ++mode) {
        if (mode == 0) k<0><<<1, 32>>>(a, iters, cyc); else k<1><<<1, 32>>>(a, iters, cyc);
        cudaDeviceSynchronize();
        printf("%s: %.1f cycles per dependent step (1 warp)\n", mode ? "redux " : "bfly  ", (double)*cyc / iters);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<nb, nt>>>(a, iters, cyc); else k<1><<<nb, nt>>>(a, iters, cyc);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s: %.3f ms for %d warps x %d steps -> %.2f warp-steps/clk/SM at 1.9 GHz\n", mode ? "redux " : "bfly  ", ms,
               nb * nt / 32, iters, (double)nb * nt / 32 * iters / (ms * 1e-3 * 1.9e9 * 148));
    }
    return 0;
}
