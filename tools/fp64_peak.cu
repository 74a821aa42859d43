// Practical FP64-pipe issue rate of a B200 SM for the instruction mix of the integrator
// (DADD / DMUL without contraction, 8 independent chains per thread), to put the "fp64_issue"
// roof of bench.py on a measured footing.   nvcc -arch=sm_100a -O3 -fmad=false fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, bool FMA>
__global__ void k(double* out, int iters, double c0, double c1)
{
    double x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        x[i] = threadIdx.x * 1e-3 + i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (FMA) {
                x[i] = fma(x[i], c0, c1);
                x[i] = fma(x[i], c0, c1);
            }
            else {
                x[i] = x[i] * c0;
                x[i] = x[i] + c1;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        s += x[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH, bool FMA>
void run(int threads, int ctas_per_sm, const char* name)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double* out;
    cudaMalloc(&out, (size_t)sms * ctas_per_sm * threads * 8);
    const int iters = 20000;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<CH, FMA><<<sms * ctas_per_sm, threads>>>(out, 100, 0.999999, 1e-9);
    cudaEventRecord(a);
    k<CH, FMA><<<sms * ctas_per_sm, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double winstr = (double)sms * ctas_per_sm * (threads / 32) * iters * CH * 2.0;
    const double per_s = winstr / (ms * 1e-3);
    printf("%-28s threads/SM %4d chains %d: %.3e warp-instr/s = %.3f per SM per cycle at %.0f MHz "
           "(%.3e lane-instr/s)\n", name, threads * ctas_per_sm, CH, per_s,
           per_s / sms / (clk * 1e3), clk / 1e3, per_s * 32);
    cudaFree(out);
}

int main()
{
    run<8, false>(512, 1, "DMUL+DADD");
    run<8, false>(1024, 1, "DMUL+DADD");
    run<4, false>(512, 1, "DMUL+DADD");
    run<2, false>(512, 1, "DMUL+DADD");
    run<8, false>(256, 1, "DMUL+DADD");
    run<8, false>(128, 1, "DMUL+DADD");
    run<8, true>(512, 1, "DFMA");
    run<8, true>(1024, 1, "DFMA");
    return 0;
}
