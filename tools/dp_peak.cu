// Practical ceiling of the odds-form recurrence on one SM sub-partition: R DFMAs per step whose three operands are
// all vector registers (E[r-1], the per-read odds, E[r]), with and without the per-step shuffle / shared-memory load.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/dp_peak tools/dp_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int MODE>
__global__ void __launch_bounds__(128) k(double *out, int steps, const double2 *par)
{
    __shared__ double2 sp[4][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    sp[w][lane] = par[lane];
    __syncwarp();
    double E[R], T = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) E[r] = (r == 0 && lane == 0) ? 1.0 : 0.0;
    double2 c_next = sp[w][0];
    double in_next = 0.0;
    if (MODE >= 2) in_next = __shfl_up_sync(0xffffffffu, E[R - 1], 1);
    for (int j = 0; j < steps; ++j) {
        const double2 cc = c_next;
        const double in = (MODE >= 2) ? (lane == 0 ? 0.0 : in_next) : 1e-9;
        if (MODE >= 1) c_next = sp[w][(j + 1) & 31];
        const double top = E[R - 1];
        if (MODE >= 3) T = fma(top, cc.x, T * cc.y);
        E[R - 1] = fma(E[R - 2], cc.x, top);
        if (MODE >= 2) in_next = __shfl_up_sync(0xffffffffu, E[R - 1], 1);
#pragma unroll
        for (int r = R - 2; r >= 1; --r) E[r] = fma(E[r - 1], cc.x, E[r]);
        E[0] = fma(in, cc.x, E[0]);
    }
    double s = T;
#pragma unroll
    for (int r = 0; r < R; ++r) s += E[r];
    if (s == 123.456) out[0] = s;
}

template <int R, int MODE>
void run(int ctas_per_sm, const double2 *par, double *out)
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int steps = 200000;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<R, MODE><<<sms * ctas_per_sm, 128>>>(out, 1000, par);
    cudaEventRecord(a);
    k<R, MODE><<<sms * ctas_per_sm, 128>>>(out, steps, par);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double fp64_ops = (double)sms * ctas_per_sm * 4 * 32 * steps * (R + (MODE >= 3 ? 2 : 0));
    const double cyc_per_step = ms * 1e-3 * 1.965e9 / steps / ctas_per_sm;   // per warp-step per SMSP (4 warps per CTA = 1 per SMSP)
    printf("R=%d mode=%d warps/SMSP=%d: %.2f T fp64-instr-lanes/s (%.1f%% of 1.70e13), %.1f cycles per warp-step\n", R, MODE, ctas_per_sm,
           fp64_ops / (ms * 1e-3) / 1e12, 100.0 * fp64_ops / (ms * 1e-3) / 1.70e13, cyc_per_step);
}

int main()
{
    double2 h[32];
    for (int i = 0; i < 32; ++i) h[i] = make_double2(1e-3, 1.001);
    double2 *par; double *out;
    cudaMalloc(&par, sizeof(h)); cudaMalloc(&out, 8);
    cudaMemcpy(par, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int c : {1, 2, 4, 6, 8}) {
        run<8, 0>(c, par, out);
        run<8, 1>(c, par, out);
        run<8, 2>(c, par, out);
        run<8, 3>(c, par, out);
    }
    run<6, 3>(6, par, out);
    run<4, 3>(6, par, out);
    run<12, 3>(4, par, out);
    run<16, 3>(4, par, out);
    return 0;
}
