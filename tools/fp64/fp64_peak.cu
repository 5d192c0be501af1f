// fp64_peak.cu — FP64 pipe micro-benchmark for the roofline of SURVEY.md 8(d): sustained DFMA rate
// and the rate of the reference's arithmetic (separate DMUL + DADD, no contraction) on this GPU.
// Measurement tool, not part of the library.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -fmad=false -o fp64_peak fp64_peak.cu ; prints one JSON line.
#include <cuda_runtime.h>
#include <cstdio>

template <bool FUSED>
__global__ void __launch_bounds__(256, 4) chains(double* out, double a, double b, int iters)
{
    double x[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = (double)(threadIdx.x + q) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (FUSED) x[q] = __fma_rn(x[q], a, b);
            else x[q] = __dadd_rn(__dmul_rn(x[q], a), b);
        }
    }
    double s = 0.;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += x[q];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <bool FUSED>
static double run(int sms, int iters)
{
    const int blocks = sms * 4, threads = 256;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    chains<FUSED><<<blocks, threads>>>(out, 0.999999, 1e-7, iters / 8);   // warm-up
    cudaDeviceSynchronize();
    double best = 0.;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        chains<FUSED><<<blocks, threads>>>(out, 0.999999, 1e-7, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;   // multiply + add per step
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaFree(out);
    return best;
}

int main()
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    const double fused = run<true>(p.multiProcessorCount, 1 << 16);
    const double split = run<false>(p.multiProcessorCount, 1 << 16);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_dfma_tflops\": %.2f, \"fp64_dmul_dadd_tflops\": %.2f, "
           "\"note\": \"2 flops per multiply-add step; 8 independent chains per thread, 1024 threads per SM\"}\n",
           p.name, p.multiProcessorCount, fused, split);
    return 0;
}
