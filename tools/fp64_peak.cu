// DFMA throughput microbenchmark: the fp64-pipe roofline denominator for k_heavy (SURVEY.md §8d asks
// for a measured peak; MEASURED_PEAKS.json has none for fp64).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = 1.0 + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}

int main()
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *out;
    cudaMalloc(&out, 8);
    const int iters = 20000, ILP = 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int bps = 1; bps <= 8; bps *= 2) {
        const int grid = sms * bps;
        k_dfma<ILP><<<grid, 256>>>(out, 1000, 0.999999, 1e-7);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            k_dfma<ILP><<<grid, 256>>>(out, iters, 0.999999, 1e-7);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double fma = (double)grid * 256 * ILP * iters;
        printf("{\"sms\": %d, \"blocks_per_sm\": %d, \"dfma_per_s\": %.4e, \"fp64_tflops\": %.3f, \"dfma_per_clk_per_sm_at_1965MHz\": %.2f}\n",
               sms, bps, fma / (best * 1e-3), 2.0 * fma / (best * 1e-3) / 1e12, fma / (best * 1e-3) / sms / 1.965e9);
    }
    return 0;
}
