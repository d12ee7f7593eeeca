// Micro-benchmarks behind the step-kernel design decisions (run on the B200 box: tools/ubench/run.sh).
//  1. DFMA issue rate vs independent chains per warp and warps per SM, with 3 distinct register operands
//  2. the cubic-Horner likelihood loop fed by broadcast shared-memory loads: U points in flight, Q chains per thread
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int ILP>
__global__ void __launch_bounds__(1024) dfma3_kernel(double *out, int iters, double a0, double b0) {
    double x[ILP], a[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 1e-9 + i; a[i] = a0 + i * 1e-9; b[i] = b0 + i * 1e-9; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a[i], b[i]);      // three distinct register operands per DFMA
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// cubic Horner + weighted residual + square, records [x, w, wy, pad] broadcast from shared memory
template <int U, int Q, int G = 1>
__global__ void __launch_bounds__(512) horner_kernel(double *out, const double *rec_g, int n, int reps) {
    extern __shared__ double rec[];
    for (int i = threadIdx.x; i < n * 4; i += blockDim.x) rec[i] = rec_g[i];
    __syncthreads();
    double th[Q][4], J[Q][U];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
#pragma unroll
        for (int j = 0; j < 4; ++j) th[q][j] = 1.0 + 1e-6 * (threadIdx.x + 7 * q + j);
#pragma unroll
        for (int u = 0; u < U; ++u) J[q][u] = 0.0;
    }
    for (int r = 0; r < reps; ++r) {
#pragma unroll 1
        for (int k = (G == 1) ? 0 : (int)((threadIdx.x & 31) / (32 / G)); k + (U - 1) * G < n; k += U * G) {
            double x[U], w[U], wy[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double2 v = *reinterpret_cast<const double2 *>(rec + (k + u * G) * 4);
                x[u] = v.x; w[u] = v.y; wy[u] = rec[(k + u * G) * 4 + 2];
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                double f[U];
#pragma unroll
                for (int u = 0; u < U; ++u) f[u] = th[q][3];
#pragma unroll
                for (int j = 2; j >= 0; --j)
#pragma unroll
                    for (int u = 0; u < U; ++u) f[u] = fma(f[u], x[u], th[q][j]);
#pragma unroll
                for (int u = 0; u < U; ++u) { const double t = fma(-f[u], w[u], wy[u]); J[q][u] = fma(t, t, J[q][u]); }
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) th[q][0] += 1e-9 * J[q][0];
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int u = 0; u < U; ++u) s += J[q][u];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < 3; ++i) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

template <int ILP>
static void run_dfma(double *out, int nsm) {
    for (int warps : {4, 8, 16, 32}) {
        const int iters = 2048;
        float ms = time_ms([&] { dfma3_kernel<ILP><<<nsm, warps * 32>>>(out, iters, 0.999999, 1e-6); });
        double fl = 2.0 * 8 * ILP * (double)iters * nsm * warps * 32;
        printf("dfma3 ilp=%d warps/SM=%2d  %.2f TFLOP/s\n", ILP, warps, fl / (ms * 1e-3) / 1e12);
    }
}

template <int U, int Q, int G = 1>
static void run_horner(double *out, const double *rec, int nsm) {
    const int n = 1000, reps = 200;
    CK(cudaFuncSetAttribute(horner_kernel<U, Q, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 32));
    for (int warps : {4, 8, 12, 14, 16}) {
        if (warps * 32 > 512) continue;
        float ms = time_ms([&] { horner_kernel<U, Q, G><<<nsm, warps * 32, n * 32>>>(out, rec, n, reps); });
        double instr = 5.0 * Q * (double)(n / (U * G) * U) * reps * nsm * warps * 32;   // DFMA thread-instructions
        printf("horner U=%d Q=%d G=%d warps/SM=%2d  %.2f TFLOP/s (2 flops per DFMA)  %.3f ms\n", U, Q, G, warps, 2 * instr / (ms * 1e-3) / 1e12, ms);
    }
}

int main() {
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *out, *rec;
    CK(cudaMalloc(&out, (size_t)nsm * 1024 * 8));
    CK(cudaMalloc(&rec, 1000 * 32));
    double h[4000];
    for (int i = 0; i < 1000; ++i) { h[4 * i] = -1.0 + 0.002 * i; h[4 * i + 1] = 10.0; h[4 * i + 2] = 3.0; h[4 * i + 3] = 0; }
    CK(cudaMemcpy(rec, h, sizeof h, cudaMemcpyHostToDevice));
    if (getenv("UBENCH_DFMA")) { run_dfma<1>(out, nsm); run_dfma<2>(out, nsm); run_dfma<4>(out, nsm); run_dfma<8>(out, nsm); }
    run_horner<4, 2>(out, rec, nsm); run_horner<4, 2, 2>(out, rec, nsm); run_horner<2, 4>(out, rec, nsm); run_horner<2, 4, 4>(out, rec, nsm);
    run_horner<3, 4, 4>(out, rec, nsm); run_horner<4, 4, 4>(out, rec, nsm); run_horner<2, 3>(out, rec, nsm); run_horner<3, 3>(out, rec, nsm);
    run_horner<2, 4, 8>(out, rec, nsm); run_horner<1, 4, 4>(out, rec, nsm); run_horner<2, 8, 8>(out, rec, nsm);
    return 0;
}
