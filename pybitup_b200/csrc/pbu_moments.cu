// K4: reduction of the per-chain running moments (kept-state Welford accumulators of the step kernels) over all
// chains of an engine.  This is the device half of what the reference does with np.mean / np.cov over its chain
// file (compute_covariance, metropolis_hastings_algorithms.py:295-307), extended to many chains: the sums below
// are additive over chains AND over GPUs, so the cross-GPU step is one small all-reduce of 2 + 4d + d(d+1)/2
// doubles (pooled covariance for the proposal / preconditioner, Gelman-Rubin R-hat).
//
// Output vector (doubles), delta_c = mean_c - mu for a caller-supplied reference point mu:
//   [0]                 number of chains with at least one kept state
//   [1]                 sum_c n_c
//   [2      .. 2+d)     sum_c n_c delta_c,i
//   [2+d    .. 2+2d)    sum_c delta_c,i
//   [2+2d   .. 2+3d)    sum_c delta_c,i^2
//   [2+3d   .. 2+4d)    sum_c s2_c,i            (within-chain variance M2_c,ii/(n_c-1), chains with n_c > 1)
//   [2+4d   .. +d(d+1)/2) sum_c (M2_c,ij + n_c delta_c,i delta_c,j)   packed upper triangle
// One block per output component walks all chains in a fixed order, so the result is deterministic.
#include <cstring>

#include "pbu_engine_internal.h"

using namespace pbu;

__global__ void __launch_bounds__(256) chain_sums_kernel(ChainState st, int d, const double *mu, double *out) {
    const int l = blockIdx.x;
    const int64_t C = st.n_chains;
    const int nt = d * (d + 1) / 2;
    int kind, i = 0, j = 0;
    if (l < 2) {
        kind = l;
    } else if (l < 2 + 4 * d) {
        kind = 2 + (l - 2) / d;
        i = (l - 2) % d;
    } else {
        kind = 6;
        int t = l - 2 - 4 * d;
        for (i = 0; i < d; ++i) {
            if (t < d - i) break;
            t -= d - i;
        }
        j = i + t;
    }
    (void)nt;
    const double mui = mu[i], muj = mu[j];
    double acc = 0.0;
    for (int64_t c = threadIdx.x; c < C; c += blockDim.x) {
        const double n = (double)st.wcount[c];
        double v = 0.0;
        if (n > 0.0) {
            const double di = st.wmean[(int64_t)i * C + c] - mui;
            switch (kind) {
                case 0: v = 1.0; break;
                case 1: v = n; break;
                case 2: v = n * di; break;
                case 3: v = di; break;
                case 4: v = di * di; break;
                case 5: v = (n > 1.0) ? st.wM2[(int64_t)tri_idx(d, i, i) * C + c] / (n - 1.0) : 0.0; break;
                default: {
                    const double dj = st.wmean[(int64_t)j * C + c] - muj;
                    v = st.wM2[(int64_t)tri_idx(d, i, j) * C + c] + n * di * dj;
                }
            }
        }
        acc += v;
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[l] = sh[0];
}

__global__ void reset_moments_kernel(ChainState st, int d) {
    const int64_t C = st.n_chains;
    const int nt = d * (d + 1) / 2;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < C; c += (int64_t)gridDim.x * blockDim.x) {
        st.wcount[c] = 0;
        for (int i = 0; i < d; ++i) st.wmean[(int64_t)i * C + c] = 0.0;
        for (int i = 0; i < nt; ++i) st.wM2[(int64_t)i * C + c] = 0.0;
    }
}

PBU_API int pbu_engine_chain_sums_len(const pbu_engine *e) { return e ? 2 + 4 * e->d + e->nt : -1; }

PBU_API int pbu_engine_chain_sums(pbu_engine *e, const double *mu_host, double *out_host) {
    if (!e || !out_host) return fail(PBU_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(e->device));
    const int L = 2 + 4 * e->d + e->nt;
    double *buf = nullptr;
    CUDA_TRY(cudaMalloc(&buf, (size_t)(L + e->d) * sizeof(double)));
    double mu[PBU_MAX_DIM] = {0};
    if (mu_host) memcpy(mu, mu_host, e->d * sizeof(double));
    cudaError_t ce = cudaMemcpyAsync(buf + L, mu, e->d * sizeof(double), cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess) {
        chain_sums_kernel<<<L, 256, 0, e->stream>>>(e->st, e->d, buf + L, buf);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(out_host, buf, (size_t)L * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    cudaFree(buf);
    if (ce != cudaSuccess) return fail(PBU_ERR_CUDA, "chain_sums: %s", cudaGetErrorString(ce));
    return PBU_OK;
}

// forget the kept-state moments accumulated so far (e.g. at the end of burn-in)
PBU_API int pbu_engine_reset_moments(pbu_engine *e) {
    if (!e) return fail(PBU_ERR_INVALID, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    reset_moments_kernel<<<296, 256, 0, e->stream>>>(e->st, e->d);
    CUDA_TRY(cudaGetLastError());
    return PBU_OK;
}
