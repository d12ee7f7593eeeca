// FP64 peak micro-benchmark (denominator of the FP64 roofline; MEASURED_PEAKS.json has no FP64 entry).
// Each thread runs 8 independent DFMA chains; 8 warps per SMSP keep the FP64 pipe saturated.
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0,
           x7 = x0 + 7.0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b);
            x1 = fma(x1, a, b);
            x2 = fma(x2, a, b);
            x3 = fma(x3, a, b);
            x4 = fma(x4, a, b);
            x5 = fma(x5, a, b);
            x6 = fma(x6, a, b);
            x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// Returns achieved FP64 TFLOP/s (FMA = 2 flops) in *tflops, the kernel time in *ms.
extern "C" __attribute__((visibility("default"))) int pbu_measure_fp64_peak(int device, int iters, int reps, double *tflops, double *ms_out) {
    if (cudaSetDevice(device) != cudaSuccess) return -3;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    const int blocks = nsm * 8, threads = 256;
    double *out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return -3;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dfma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-6);   // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    if (reps > 0) {
        // burst: best single launch
        for (int r = 0; r < reps; ++r) {
            cudaEventRecord(e0);
            dfma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
    } else {
        // sustained (reps < 0): -reps launches back to back, mean time per launch -- what a kernel timed inside a
        // long step can expect once the clocks have settled under load
        const int n = -reps;
        cudaEventRecord(e0);
        for (int r = 0; r < n; ++r) dfma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-6);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms / n;
    }
    cudaError_t err = cudaGetLastError();
    cudaFree(out);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (err != cudaSuccess) return -3;
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return 0;
}
