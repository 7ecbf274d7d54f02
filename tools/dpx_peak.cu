// Micro-benchmark: sustained issue rate of the packed 16x2 integer (DPX) instructions the SGM kernels are made
// of, on the whole chip.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dpx_peak tools/dpx_peak.cu
// Prints warp-instructions per clock per SM and lane-ops/s for each instruction class.
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
template <int OP>
__global__ void __launch_bounds__(256) k(unsigned* out, int iters, unsigned s0, unsigned s1)
{
    unsigned v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) v[i] = threadIdx.x * 2654435761u + i * 40503u + s0;
    unsigned b = s1 | 0x00010001u, c = s0 ^ 0x7fff7fffu;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) v[i] = __vminu2(v[i], b + i);                       // VIMNMX.U16x2
            if (OP == 1) v[i] = __viaddmin_u16x2(v[i], b, c);                // VIADDMNMX.U16x2
            if (OP == 2) v[i] = __vadd2(v[i], b);                            // VIADD.16x2
            if (OP == 3) v[i] = __byte_perm(v[i], b, 0x5432);                // PRMT
            if (OP == 4) v[i] = v[i] + b - c;                                // IADD3
            if (OP == 5) v[i] = __vimin3_u16x2(v[i], b, c + i);              // VIMNMX3.U16x2
            if (OP == 6) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1);       // SHFL.BFLY
            if (OP == 7) v[i] = __vimax_s16x2_relu(v[i], b);                 // VIMNMX.S16x2.RELU
            if (OP == 8) { v[i] = __vminu2(v[i], b + i); v[i] = v[i] * 3u + c; }   // ALU + IMAD (FMA pipe) mix
        }
        b += 0x00010001u;
    }
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r ^= v[i];
    if (r == 0x12345678u) out[threadIdx.x] = r;
}

template <int OP>
void run(const char* name, int perIter, unsigned* d, int sms, double clkGHz)
{
    const int iters = 4096, blocks = sms * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<blocks, 256>>>(d, 64, 1, 2);
    cudaEventRecord(a);
    k<OP><<<blocks, 256>>>(d, iters, 1, 2);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double warpInstr = (double)blocks * 8 * iters * CHAINS * perIter;
    const double perClkSM = warpInstr / (ms * 1e-3 * clkGHz * 1e9) / sms;
    printf("%-22s %8.3f ms  %6.3f warp-instr/clk/SM  %7.2f T lane-ops/s\n", name, ms, perClkSM, warpInstr * 32 / (ms * 1e-3) / 1e12);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
    unsigned* d;
    cudaMalloc(&d, 4096);
    run<0>("VIMNMX.U16x2", 1, d, p.multiProcessorCount, ghz);
    run<1>("VIADDMNMX.U16x2", 1, d, p.multiProcessorCount, ghz);
    run<2>("VIADD.16x2", 1, d, p.multiProcessorCount, ghz);
    run<3>("PRMT", 1, d, p.multiProcessorCount, ghz);
    run<4>("IADD3", 1, d, p.multiProcessorCount, ghz);
    run<5>("VIMNMX3.U16x2", 1, d, p.multiProcessorCount, ghz);
    run<6>("SHFL.BFLY", 1, d, p.multiProcessorCount, ghz);
    run<7>("VIMNMX.S16x2.RELU", 1, d, p.multiProcessorCount, ghz);
    run<8>("VIMNMX+IMAD pair", 2, d, p.multiProcessorCount, ghz);
    return 0;
}
