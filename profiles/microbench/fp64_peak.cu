// FP64 FMA-chain micro-benchmark: the measured peak of the FP64 pipe on this B200 (SURVEY.md §8d
// asks for the FP64 bound of the pair passes to be quoted against a MEASURED figure; it is not in
// MEASURED_PEAKS.json).  Every thread runs CHAINS independent DFMA chains; enough blocks to fill the
// chip several times.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_peak.cu -o fp64_peak
#include <cuda_runtime.h>
#include <stdio.h>

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double *out, double a, double b, int iters) {
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
static double run(int sms) {
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k_dfma<CHAINS><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * CHAINS * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaFree(out);
    return best;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double c4 = run<4>(p.multiProcessorCount), c8 = run<8>(p.multiProcessorCount), c16 = run<16>(p.multiProcessorCount);
    const double best = c4 > c8 ? (c4 > c16 ? c4 : c16) : (c8 > c16 ? c8 : c16);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %d, \"fp64_fma_tflops\": {\"chains4\": %.3f, \"chains8\": %.3f, \"chains16\": %.3f}, "
           "\"fp64_peak_tflops\": %.3f, \"dfma_per_clk_per_sm\": %.2f}\n",
           p.name, p.multiProcessorCount, clk / 1000, c4, c8, c16, best,
           best * 1e12 / 2.0 / p.multiProcessorCount / (clk * 1e3));
    return 0;
}
