// Micro-benchmark: issue throughput of the packed-int16 DPX instructions the fill kernel is built
// from (VIADDMNMX.S16x2[.RELU], VIMNMX3.S16x2, VIMNMX.S16x2, VIADD.16x2) plus LDS/SHFL mixes, on one B200.
// Prints lane-ops per clock per SM and Tops/s.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 4096

template <int OP> __global__ void k(unsigned* out, unsigned seed)
{
    unsigned a[CHAINS];
    unsigned b = seed * 3 + threadIdx.x, c = seed * 7 + 1;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = seed + i * 77 + threadIdx.x;
    __shared__ unsigned sm[1024];
    if (OP >= 6) { for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * seed; __syncthreads(); }
    for (int it = 0; it < ITERS; ++it)
    {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
        {
            if (OP == 0) a[i] = __viaddmax_s16x2_relu(a[i], b, c);
            if (OP == 1) a[i] = __vimax3_s16x2(a[i], b, c);
            if (OP == 2) a[i] = __vmaxs2(a[i], b);
            if (OP == 3) a[i] = __vadd2(a[i], 0xfffafffa);
            if (OP == 4) a[i] = __viaddmax_s16x2(a[i], 0xffffffff, b);
            if (OP == 5) { a[i] = __viaddmax_s16x2_relu(a[i], b, c); a[i] = a[i] * 3 + b; } // DPX + IMAD (fma pipe)
            if (OP == 6) { a[i] = __viaddmax_s16x2_relu(a[i], sm[(a[i] + i) & 1023], c); }      // DPX + LDS
            if (OP == 7) { a[i] = __viaddmax_s16x2_relu(a[i], __shfl_up_sync(0xffffffffu, a[i], 1), c); } // DPX + SHFL
            if (OP == 8) { // the fill kernel's per-cell op mix (5 DPX)
                unsigned t = __viaddmax_s16x2_relu(a[i], b, c);
                unsigned tg = __vadd2(t, 0xfffafffa);
                c = __viaddmax_s16x2(c, 0xffffffff, tg);
                unsigned h = __vmaxs2(t, b);
                b = __viaddmax_s16x2(b, 0xffffffff, tg);
                a[i] = h;
            }
            if (OP == 9) a[i] = max((int)a[i] + (int)b, (int)c); // plain s32 add+max
        }
    }
    unsigned r = b ^ c;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP> void run(const char* name, int ops_per_iter, int sms, double clk_ghz)
{
    unsigned* d;
    int blocks = sms * 8, threads = 256;
    cudaMalloc(&d, blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(d, 1);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep)
    {
        cudaEventRecord(e0);
        k<OP><<<blocks, threads>>>(d, rep + 2);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double laneops = (double)blocks * threads * ITERS * CHAINS * ops_per_iter;
    double per_s = laneops / (best * 1e-3);
    printf("%-28s %8.3f ms  %8.2f Tlane-op/s  %7.1f lane-op/clk/SM (at %.3f GHz)\n", name, best, per_s / 1e12,
           per_s / sms / (clk_ghz * 1e9), clk_ghz);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz / 1e6;
    printf("device %s, %d SMs, max clock %.3f GHz\n", p.name, sms, ghz);
    run<0>("viaddmax_s16x2_relu", 1, sms, ghz);
    run<1>("vimax3_s16x2", 1, sms, ghz);
    run<2>("vmaxs2", 1, sms, ghz);
    run<3>("vadd2 imm", 1, sms, ghz);
    run<4>("viaddmax_s16x2 imm", 1, sms, ghz);
    run<5>("viaddmax + imad (2 ops)", 2, sms, ghz);
    run<6>("viaddmax + lds (2 ops)", 2, sms, ghz);
    run<7>("viaddmax + shfl (2 ops)", 2, sms, ghz);
    run<8>("cell mix (5 dpx)", 5, sms, ghz);
    run<9>("s32 add+max (1 fused?)", 1, sms, ghz);
    return 0;
}
