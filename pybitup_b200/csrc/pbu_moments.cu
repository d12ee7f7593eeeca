// K4: reduction of the per-chain running moments (kept-state Welford accumulators of the step kernels) over all
// chains of an engine, and the pooled statistics built from them -- all device resident.  This is the device half of what
// the reference does with np.mean / np.cov over its chain file (compute_covariance,
// metropolis_hastings_algorithms.py:295-307) and of the covariance refresh of GradientBasedMCMC.adapt_covariance
// (:736-777), extended to many chains: the sums below are additive over chains AND over GPUs, so the cross-GPU step is one
// small all-reduce of 2 + 4d + d(d+1)/2 doubles (pooled covariance for the proposal / preconditioner, Gelman-Rubin R-hat).
//
// Sums vector (doubles), delta_c = mean_c - mu for a reference point mu:
//   [0]                 number of chains with at least one kept state
//   [1]                 sum_c n_c
//   [2      .. 2+d)     sum_c n_c delta_c,i
//   [2+d    .. 2+2d)    sum_c delta_c,i
//   [2+2d   .. 2+3d)    sum_c delta_c,i^2
//   [2+3d   .. 2+4d)    sum_c s2_c,i            (within-chain variance M2_c,ii/(n_c-1), chains with n_c > 1)
//   [2+4d   .. +d(d+1)/2) sum_c (M2_c,ij + n_c delta_c,i delta_c,j)   packed upper triangle
//
// Pipeline of one pooled-statistics call (host layer: pybitup_b200/engine.py::ChainEngine.pooled_statistics), every step
// asynchronous on the engine's stream, nothing allocated, nothing copied to the host:
//   chain_sums_dev(mu = 0) -> [NCCL all-reduce] -> pooled_mean_dev -> chain_sums_dev(mu) -> [NCCL all-reduce] ->
//   pooled_finalize_dev (mean, covariance, R-hat, proposal factor; optionally written into every chain's state).
// The first pass locates the pooled mean, the second takes the second moments about it, which removes the cancellation a
// one-pass sum of squares would have.  Summation order is fixed (slice partials added in slice order): deterministic.
#include <cstring>

#include "pbu_engine_internal.h"

using namespace pbu;

// grid (n_slices, L): block (b, l) sums component l over the chains of slice b
__global__ void __launch_bounds__(256) chain_sums_partial_kernel(ChainState st, int d, const double *mu, double *partial) {
    const int l = blockIdx.y, b = blockIdx.x, nb = gridDim.x;
    const int64_t C = st.n_chains;
    const int64_t per = (C + nb - 1) / nb;
    const int64_t c0 = (int64_t)b * per, c1 = (c0 + per < C) ? c0 + per : C;
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
    const double mui = mu ? mu[i] : 0.0, muj = mu ? mu[j] : 0.0;
    double acc = 0.0;
    for (int64_t c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
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
    if (threadIdx.x == 0) partial[(int64_t)l * nb + b] = sh[0];
}

__global__ void chain_sums_final_kernel(const double *partial, int nb, int L, double *out) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    double s = 0.0;
    for (int b = 0; b < nb; ++b) s += partial[(int64_t)l * nb + b];
    out[l] = s;
}

// mu_out = mu_in + sum_c n_c delta_c / sum_c n_c : the pooled mean located by the first pass
__global__ void pooled_mean_kernel(const double *sums, const double *mu_in, int d, double *mu_out) {
    const int i = threadIdx.x;
    if (i < d) mu_out[i] = (mu_in ? mu_in[i] : 0.0) + ((sums[1] > 0.0) ? sums[2 + i] / sums[1] : 0.0);
}

// Pooled statistics from the (all-reduced) sums about mu, the arithmetic of pybitup_b200/diagnostics.py::finalize, plus
// the proposal matrix V = scale (cov + eps I) (mha.py:463 with the pooled covariance in place of the recursion; :745-746
// for the Ito-SDE preconditioner), its upper Cholesky factor R in dpotf2 order and the inverse of R.  Layout of `stats`:
//   [0] chains  [1] n  [2] ok (1: V positive definite)  [3] reserved
//   [4 .. 4+d) mean | cov d*d row major | W d | B_over_n d | rhat d | V packed nt | R packed nt | inv(R) packed nt
__global__ void pooled_finalize_kernel(const double *sums, const double *mu, int d, double scale, double eps, double *stats) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nt = d * (d + 1) / 2;
    const double m = sums[0], n = sums[1];
    double *mean = stats + 4, *cov = mean + d, *W = cov + d * d, *Bn = W + d, *rhat = Bn + d, *V = rhat + d, *R = V + nt, *IL = R + nt;
    const double *s_ndelta = sums + 2, *s_delta = sums + 2 + d, *s_delta2 = sums + 2 + 2 * d, *s_s2 = sums + 2 + 3 * d, *s_S = sums + 2 + 4 * d;
    double shift[PBU_MAX_DIM];
    for (int i = 0; i < d; ++i) {
        shift[i] = s_ndelta[i] / n;
        mean[i] = (mu ? mu[i] : 0.0) + shift[i];
    }
    for (int i = 0; i < d; ++i)
        for (int j = i; j < d; ++j) {
            const double c = (s_S[tri_idx(d, i, j)] - n * shift[i] * shift[j]) / (n - 1.0);
            cov[i * d + j] = c;
            cov[j * d + i] = c;
            V[tri_idx(d, i, j)] = scale * (c + ((i == j) ? eps : 0.0));
        }
    const double n_c = n / m;
    for (int i = 0; i < d; ++i) {
        W[i] = s_s2[i] / m;
        const double cm = s_delta[i] / m;
        Bn[i] = (m > 1.0) ? (s_delta2[i] - m * cm * cm) / (m - 1.0) : nan("");
        rhat[i] = sqrt(((n_c - 1.0) / n_c * W[i] + Bn[i]) / W[i]);
    }
    bool ok = true;
    for (int j = 0; j < d; ++j) {                                    // dpotf2 order (pbu_common.cuh::chol_upper_packed)
        double dot = 0.0;
        for (int k = 0; k < j; ++k) dot = __fma_rn(R[tri_idx(d, k, j)], R[tri_idx(d, k, j)], dot);
        const double s = __dsub_rn(V[tri_idx(d, j, j)], dot);
        if (!(s > 0.0)) ok = false;
        const double ajj = __dsqrt_rn(s), rinv = __drcp_rn(ajj);
        R[tri_idx(d, j, j)] = ajj;
        for (int i = j + 1; i < d; ++i) {
            double t = 0.0;
            for (int k = 0; k < j; ++k) t = __fma_rn(R[tri_idx(d, k, j)], R[tri_idx(d, k, i)], t);
            R[tri_idx(d, j, i)] = __dmul_rn(__dsub_rn(V[tri_idx(d, j, i)], t), rinv);
        }
    }
    for (int j = 0; j < d; ++j) {                                    // back substitution (pbu_isde_kernel.cuh::inv_upper_packed)
        IL[tri_idx(d, j, j)] = 1.0 / R[tri_idx(d, j, j)];
        for (int i = j - 1; i >= 0; --i) {
            double s = 0.0;
            for (int k = i + 1; k <= j; ++k) s += R[tri_idx(d, i, k)] * IL[tri_idx(d, k, j)];
            IL[tri_idx(d, i, j)] = -s / R[tri_idx(d, i, i)];
        }
    }
    stats[0] = m;
    stats[1] = n;
    stats[2] = (ok && n > 1.0) ? 1.0 : 0.0;
    stats[3] = 0.0;
}

// dst[r][c] = vals[r] for every chain c, when *ok != 0 (a pooled covariance that is not positive definite leaves the
// proposal as it was; the flag stays in the statistics for the host to see)
__global__ void broadcast_rows_dev_kernel(double *dst, int64_t C, int n_rows, const double *vals, const double *ok) {
    if (*ok == 0.0) return;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    for (int r = 0; r < n_rows; ++r) dst[(int64_t)r * C + c] = vals[r];
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

static int sums_len(const pbu_engine *e) { return 2 + 4 * e->d + e->nt; }
static int n_slices(const pbu_engine *e) {
    const int64_t nb = (e->C + 2047) / 2048;
    return (int)(nb < 1 ? 1 : (nb > PBU_MOMENT_SLICES ? PBU_MOMENT_SLICES : nb));
}

PBU_API int pbu_engine_chain_sums_len(const pbu_engine *e) { return e ? sums_len(e) : -1; }
PBU_API int pbu_engine_pooled_stats_len(const pbu_engine *e) { return e ? 4 + 4 * e->d + e->d * e->d + 3 * e->nt : -1; }

PBU_API int pbu_engine_chain_sums_dev(pbu_engine *e, const double *mu_dev, double *out_dev, int n_components) {
    if (!e || !out_dev) return fail(PBU_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(e->device));
    const int nb = n_slices(e);
    const int L = (n_components > 0 && n_components < sums_len(e)) ? n_components : sums_len(e);
    chain_sums_partial_kernel<<<dim3(nb, L), 256, 0, e->stream>>>(e->st, e->d, mu_dev, e->d_mom_partial);
    chain_sums_final_kernel<<<(L + 127) / 128, 128, 0, e->stream>>>(e->d_mom_partial, nb, L, out_dev);
    CUDA_TRY(cudaGetLastError());
    return PBU_OK;
}

PBU_API int pbu_engine_pooled_mean_dev(pbu_engine *e, const double *sums_dev, const double *mu_in_dev, double *mu_out_dev) {
    if (!e || !sums_dev || !mu_out_dev) return fail(PBU_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(e->device));
    pooled_mean_kernel<<<1, 32, 0, e->stream>>>(sums_dev, mu_in_dev, e->d, mu_out_dev);
    CUDA_TRY(cudaGetLastError());
    return PBU_OK;
}

PBU_API int pbu_engine_pooled_finalize_dev(pbu_engine *e, const double *sums_dev, const double *mu_dev, double scale, double eps, int apply,
                                           double *stats_dev) {
    if (!e || !sums_dev || !stats_dev) return fail(PBU_ERR_INVALID, "null argument");
    if (apply && !e->initialised) return fail(PBU_ERR_INVALID, "pbu_engine_init_chains has not been called");
    CUDA_TRY(cudaSetDevice(e->device));
    const int d = e->d, nt = e->nt;
    pooled_finalize_kernel<<<1, 32, 0, e->stream>>>(sums_dev, mu_dev, d, scale, eps, stats_dev);
    if (apply) {
        const double *V = stats_dev + 4 + 4 * d + d * d, *R = V + nt, *IL = R + nt, *ok = stats_dev + 2;
        const unsigned grid = (unsigned)((e->C + 255) / 256);
        if (e->cfg.algo == PBU_ALGO_ISDE || e->cfg.algo == PBU_ALGO_HMC) {      // C_approx, inv_L_c (mha.py:773-777)
            broadcast_rows_dev_kernel<<<grid, 256, 0, e->stream>>>(e->d_Cmat, e->C, nt, V, ok);
            broadcast_rows_dev_kernel<<<grid, 256, 0, e->stream>>>(e->st.R, e->C, nt, IL, ok);
        } else {                                                                 // R = cholesky(V) (mha.py:486-487)
            broadcast_rows_dev_kernel<<<grid, 256, 0, e->stream>>>(e->st.R, e->C, nt, R, ok);
            // the non-adaptive step kernels take the shared factor as a kernel parameter: it comes back through pinned
            // memory and pbu_engine_run waits for this event before its next launch (nt + 1 doubles, no other host traffic)
            CUDA_TRY(cudaMemcpyAsync(e->h_R0_pinned, R, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
            CUDA_TRY(cudaMemcpyAsync(e->h_R0_pinned + nt, ok, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
            CUDA_TRY(cudaEventRecord(e->r0_event, e->stream));
            e->r0_pending = true;
        }
    }
    CUDA_TRY(cudaGetLastError());
    return PBU_OK;
}

// host-buffer form of the sums (one blocking copy; the device form above is what the samplers use)
PBU_API int pbu_engine_chain_sums(pbu_engine *e, const double *mu_host, double *out_host) {
    if (!e || !out_host) return fail(PBU_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(e->device));
    const int L = sums_len(e);
    double *mu_dev = nullptr;
    if (mu_host) {
        mu_dev = e->d_mom_scratch + L;
        CUDA_TRY(cudaMemcpyAsync(mu_dev, mu_host, e->d * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    }
    int rc = pbu_engine_chain_sums_dev(e, mu_dev, e->d_mom_scratch, 0);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_host, e->d_mom_scratch, (size_t)L * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
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
