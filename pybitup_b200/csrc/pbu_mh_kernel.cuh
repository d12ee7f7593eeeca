// K1: fused Metropolis-Hastings step kernel (RWMH / AM / DR / DRAM), many iterations per launch.
//
// Replaces, per chain and per iteration, the reference's Python loop
//   run_algorithm             metropolis_hastings_algorithms.py:158-182, :407-430
//   compute_new_val           :184-193   (d normals, new = cur + R @ z with R upper)
//   compute_acceptance_ratio  :195-207   (-> BayesianPosterior.compute_log_value,
//                                         bayesian_inference.py:44-64; Mixture prior,
//                                         distributions.py:456-479; Gaussian sum of squares,
//                                         bayesian_inference.py:495-542; Model.run_model :338-371)
//   accept_reject             :210-224, :520-567 (delayed rejection)
//   adapt_covariance          :432-487, :591-595
// with no host round trip per iteration.
//
// Mapping: G lanes of a warp own one chain (G = 1, 2, 8).  All G lanes carry the chain state
// redundantly and execute the scalar part of the step identically; the sum over the N data points
// is strided over the G lanes and combined with an xor-butterfly of warp shuffles, so every lane
// ends with the bit-identical sum.  The likelihood records are staged into shared memory with bulk
// TMA copies and read by all chains of the block: with G = 1 every shared-memory read is a warp
// broadcast.  Registers hold only what the data loop needs (state x, proposal, model constants,
// U records in flight); the adaptive-Metropolis state (X_av, V_i, R) is touched once per iteration
// and lives in global memory, chain-minor, where it stays L1/L2 resident.  That keeps the kernel
// under 128 registers so that one block of up to 512 threads (16 warps) fills an SM, and lets the
// host pick a block size that spreads the chains evenly over the 148 SMs (pbu_engine.cu::plan_launch).
#pragma once
#include "pbu_common.cuh"
#include "pbu_models.cuh"

namespace pbu {

constexpr int MH_MAX_BLOCK = 512;

// ------------------------------------------------------------------------------------------------
// Stage `bytes` (multiple of 16) from global into shared memory with bulk TMA copies; all threads
// of the block return once the data has landed.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_records(double *smem_rec, uint64_t *bar, const double *gsrc, uint32_t bytes) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bytes);
        const uint32_t CH = 32768u;
        for (uint32_t off = 0; off < bytes; off += CH) {
            const uint32_t n = (bytes - off < CH) ? (bytes - off) : CH;
            tma_load_1d(reinterpret_cast<char *>(smem_rec) + off, reinterpret_cast<const char *>(gsrc) + off, n, bar);
        }
    }
    mbar_wait(bar, 0);
}

template <int NC>
__device__ __forceinline__ void load_record(const double *smem_rec, int k, double (&rec)[NC]) {
    const double2 *r2 = reinterpret_cast<const double2 *>(smem_rec + (size_t)k * NC);
#pragma unroll
    for (int j = 0; j < NC / 2; ++j) {
        const double2 v = r2[j];
        rec[2 * j] = v.x;
        rec[2 * j + 1] = v.y;
    }
}

// ------------------------------------------------------------------------------------------------
// Gaussian sum of squares J(theta) = sum_runs sum_k ((y_run,k - f_k)/(std_y,k gamma))^2 over the records of
// one shared-memory tile (Likelihood.arg_gauss_likelihood, bayesian_inference.py:495-542), and, when
// GRAD, the gradient sum of Likelihood.compute_grad_log_value (:573-623).  Several experimental runs
// are folded on the host with the exact identity
//   sum_r (y_r - f)^2 = n_r (ybar - f)^2 + sum_r (y_r - ybar)^2
// so a record carries w = sqrt(n_r)/(std_y gamma) and w*ybar, and the second term is the constant
// prob.J_const: the device loop costs the same whatever n_runs is.
// Ju[U] are U independent partial sums (one per point in flight).
// ------------------------------------------------------------------------------------------------
template <class M, int G, bool GRAD>
__device__ __forceinline__ void tile_accumulate(const typename M::Ctx &ctx, typename M::Seq &seq, const double *tile, int n, int sub,
                                                double (&Ju)[M::UNROLL], double (&g)[GRAD ? M::NPARAM : 1]) {
    constexpr int NC = M::NC;
    constexpr int U = M::UNROLL;
    if (M::SEQUENTIAL) {
        for (int k = 0; k < n; ++k) {
            double rec[1][NC], t[1];
            load_record<NC>(tile, k, rec[0]);
            M::template resid_batch<1>(ctx, seq, rec, t);
            Ju[0] = fma(t[0], t[0], Ju[0]);
        }
    } else {
        int k = sub;
#pragma unroll 1
        for (; k + (U - 1) * G < n; k += U * G) {
            double rec[U][NC], t[U];
#pragma unroll
            for (int u = 0; u < U; ++u) load_record<NC>(tile, k + u * G, rec[u]);
            M::template resid_batch<U>(ctx, seq, rec, t);
#pragma unroll
            for (int u = 0; u < U; ++u) Ju[u] = fma(t[u], t[u], Ju[u]);
            if constexpr (GRAD) {
#pragma unroll
                for (int u = 0; u < U; ++u) M::grad_acc(ctx, rec[u], t[u], g);
            }
        }
#pragma unroll 1
        for (; k < n; k += G) {
            double rec[1][NC], t[1];
            load_record<NC>(tile, k, rec[0]);
            M::template resid_batch<1>(ctx, seq, rec, t);
            Ju[0] = fma(t[0], t[0], Ju[0]);
            if constexpr (GRAD) M::grad_acc(ctx, rec[0], t[0], g);
        }
    }
}

template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned group_mask) {
#pragma unroll
    for (int s = 1; s < G; s <<= 1) v += __shfl_xor_sync(group_mask, v, s);
    return v;
}

// Full parameter vector from the uncertain ones (Model.run_model, bayesian_inference.py:361-369).
template <class M, int D>
__device__ __forceinline__ void full_theta(const double (&x)[D], const ProblemDev &prob, double (&theta)[M::NPARAM]) {
#pragma unroll
    for (int q = 0; q < M::NPARAM; ++q) {
        double v = prob.theta_nom[q];
#pragma unroll
        for (int i = 0; i < D; ++i)
            if (prob.unc_idx[i] == q) v = x[i];
        theta[q] = v;
    }
}

template <int D>
__device__ __forceinline__ bool inside_box(const double (&x)[D], const ProblemDev &prob) {
    bool inside = true;
#pragma unroll
    for (int i = 0; i < D; ++i) inside = inside && !(x[i] < prob.lb[i] || x[i] > prob.ub[i]);   // bounds inclusive, distributions.py:284
    return inside;
}

// BayesianPosterior.compute_log_value (bayesian_inference.py:44-64) with the identity
// parametrisation: -inf outside the prior box WITHOUT evaluating the model (:53-55), otherwise
// log_prior - log(1) + (-(1/2) J).   Resident mode: all records are in shared memory.
template <class M, int D, int G>
__device__ __forceinline__ double log_posterior(const double (&x)[D], const ProblemDev &prob, const double *smem_rec, int sub,
                                                unsigned group_mask) {
    if (!inside_box<D>(x, prob)) return -INFINITY;
    double theta[M::NPARAM];
    full_theta<M, D>(x, prob, theta);
    typename M::Ctx ctx;
    M::begin(ctx, theta, prob);
    typename M::Seq seq;
    M::seq_init(ctx, seq, prob);
    double Ju[M::UNROLL], gdummy[1];
#pragma unroll
    for (int u = 0; u < M::UNROLL; ++u) Ju[u] = 0.0;
    tile_accumulate<M, G, false>(ctx, seq, smem_rec, prob.n_points, sub, Ju, gdummy);
    double J = 0.0;
#pragma unroll
    for (int u = 0; u < M::UNROLL; ++u) J += Ju[u];
    J = group_sum<G>(J, group_mask) + prob.J_const;
    return (prob.log_prior_const - 0.0) + (-0.5 * J);
}

// ------------------------------------------------------------------------------------------------
// Random draws for one proposal stage: D standard normals and one uniform.
// ------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void draw_philox(uint64_t seed, uint64_t chain, uint64_t it, uint32_t stage, double (&z)[D], double &u) {
    constexpr int NPAIR = (D + 1) / 2;
    constexpr int NW = 2 * NPAIR + 2;
    constexpr int NB = (NW + 3) / 4;
    uint32_t w[NB * 4];
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        uint4 ctr = make_uint4((uint32_t)chain, (uint32_t)it, (uint32_t)(it >> 32) ^ ((uint32_t)(chain >> 32) << 8) ^ (stage << 30),
                               (uint32_t)b);
        uint4 r = Philox::round10(ctr, key);
        w[4 * b] = r.x;
        w[4 * b + 1] = r.y;
        w[4 * b + 2] = r.z;
        w[4 * b + 3] = r.w;
    }
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) {
        double a, b;
        box_muller(w[2 * p], w[2 * p + 1], a, b);
        z[2 * p] = a;
        if (2 * p + 1 < D) z[2 * p + 1] = b;
    }
    u = uniform53(w[2 * NPAIR], w[2 * NPAIR + 1]);
}

template <int D>
__device__ __forceinline__ bool draw_injected(const StreamsDev &s, int64_t c, int64_t &cn, int64_t &cu, bool want_uniform, double (&z)[D],
                                              double &u) {
    if (cn + D > s.n_normals || (want_uniform && cu + 1 > s.n_uniforms)) {
#pragma unroll
        for (int i = 0; i < D; ++i) z[i] = 0.0;
        u = 1.0;
        return false;
    }
    const double *zn = s.normals + c * s.n_normals + cn;
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = zn[i];
    cn += D;
    if (want_uniform) {
        u = s.uniforms[c * s.n_uniforms + cu];
        cu += 1;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// The kernel.
// ------------------------------------------------------------------------------------------------
template <class M, int D, int G, bool ADAPT>
__global__ void __launch_bounds__(MH_MAX_BLOCK, 1) mh_step_kernel(const __grid_constant__ MhParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *smem_rec = reinterpret_cast<double *>(smem_raw + 128);
    constexpr int NT = tri_size(D);

    stage_records(smem_rec, bar, p.prob.records, (uint32_t)p.prob.n_points * (uint32_t)(M::NC * 8));

    const int64_t C = p.st.n_chains;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t c = gtid / G;
    const int sub = (int)(gtid % G);
    const bool live = c < C;
    if (!live) c = C - 1;                       // padding lanes shadow the last chain, never write
    const bool writer = live && (sub == 0);
    const int lane = threadIdx.x & 31;
    const unsigned group_mask = (G == 1) ? (1u << lane) : (((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << (lane & ~(G - 1)));

    // ---- chain state kept in registers across the launch
    double x[D], lp;
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = p.st.x[i * C + c];
    lp = p.st.lp[c];
    double mean_r = p.st.mean_r[c];
    int32_t n_rej = 0, n_eval = 0;              // per-launch increments of the int64 counters
    int32_t status = p.st.status[c];
    const bool injected = p.streams.normals != nullptr;
    int64_t cn = 0, cu = 0;
    if (injected) {
        cn = p.streams.cur_n[c];
        cu = p.streams.cur_u[c];
    }
    const uint64_t gchain = p.chain_offset + (uint64_t)c;
    // per-chain adaptive state in global memory (chain-minor): every lane of a group reads and, if live,
    // writes the same values, so each thread only ever reads back its own stores.
    double *const gR = p.st.R + c;
    double *const gV = p.st.V + c;
    double *const gxav = p.st.xav + c;

    for (int64_t s = 0; s < p.n_steps; ++s) {
        const int64_t it = p.it0 + s + 1;
        // Up to two proposal stages share ONE inlined copy of the posterior evaluation (code size):
        // stage 0 = the Metropolis proposal (mha.py:184-224), stage 1 = the delayed-rejection
        // proposal (mha.py:535-567), entered only by chains whose first proposal was rejected.
        double y1[D], lp1 = 0.0, lp2 = nan(""), r = 0.0, alpha = 0.0;
        int acc = 0;
#pragma unroll 1
        for (int stage = 0; stage < 2; ++stage) {
            double z[D], u;
            if (injected) {
                if (!draw_injected<D>(p.streams, c, cn, cu, true, z, u)) status |= ST_STREAM_EXHAUSTED;
            } else {
                draw_philox<D>(p.seed, gchain, (uint64_t)it, (uint32_t)stage, z, u);
            }
            // new = cur + scale * R @ z with R UPPER triangular (mha.py:193, :543)
            double y[D];
            const double scale = (stage == 0) ? 1.0 : p.gamma_dr;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                double sj = 0.0;
#pragma unroll
                for (int j = i; j < D; ++j) {
                    const double rij = ADAPT ? gR[(int64_t)tri_idx(D, i, j) * C] : p.R0[tri_idx(D, i, j)];
                    sj = fma(rij, z[j], sj);
                }
                y[i] = x[i] + scale * sj;
            }
            const double lpy = log_posterior<M, D, G>(y, p.prob, smem_rec, sub, group_mask);
            n_eval += 1;
            bool accept;
            if (stage == 0) {
#pragma unroll
                for (int i = 0; i < D; ++i) y1[i] = y[i];
                lp1 = lpy;
                r = (lp1 == -INFINITY) ? 0.0 : exp(lp1 - lp);              // :200-205
                if (r != r) {
                    status |= ST_NAN_RATIO;
                    if (p.nan_rejects) r = 0.0;
                }
                alpha = py_min1(r);                                         // :207
                accept = u < alpha;                                         // :214
            } else {
                lp2 = lpy;
                const double r12 = exp(lp1 - lp2);                          // :547
                const double a12 = py_min1(r12);
                // inv_V = inv_R * inv_R^T ELEMENTWISE = diag(1/R_ii^2)  (:517-518, :594-595)
                double M2 = 0.0, M4 = 0.0;
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    const double rii = ADAPT ? gR[(int64_t)tri_idx(D, i, i) * C] : p.R0[tri_idx(D, i, i)];
                    const double ir = 1.0 / rii;
                    const double iv = ir * ir;
                    const double d1 = y1[i] - y[i];
                    const double d2 = y1[i] - x[i];
                    M2 += (d1 * iv) * d1;
                    M4 += (d2 * iv) * d2;
                }
                double r2 = exp(lp2 - lp) * (exp(-0.5 * M2) / exp(-0.5 * M4)) * (1.0 - a12) / (1.0 - alpha);   // :555-556
                if (r2 != r2) {
                    status |= ST_NAN_RATIO;
                    if (p.nan_rejects) r2 = 0.0;
                }
                accept = u < py_min1(r2);                                   // :558-562
            }
            if (accept) {
#pragma unroll
                for (int i = 0; i < D; ++i) x[i] = y[i];
                lp = lpy;
                acc = stage + 1;
                break;
            }
            if (stage == 1 || !p.delayed) {
                n_rej += 1;                                                 // :224, :567
                break;
            }
        }

        // ---- covariance adaptation (mha.py:432-487)
        if (ADAPT) {
            const double fi = (double)it;
            const bool refresh = (it % p.updating_it == 0);                 // :470
            double Vc[NT];
            if (!p.am_welford) {
                // literal Haario/Smith recursion, evaluated in the reference's operation order
                // without FMA contraction; at it == 1 X_av_i aliases the state.
                double xa[D], xb[D];
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    xa[i] = (it == 1) ? x[i] : gxav[(int64_t)i * C];
                    xb[i] = __dadd_rn(x[i], __dmul_rn(fi / (fi + 1.0), __dadd_rn(xa[i], -x[i])));       // :455
                }
                const double c1 = (fi - 1.0) / fi, c2 = p.S_d / fi;
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = i; j < D; ++j) {
                        double t = __dadd_rn(__dmul_rn(fi, __dmul_rn(xa[i], xa[j])), -__dmul_rn(fi + 1.0, __dmul_rn(xb[i], xb[j])));
                        t = __dadd_rn(t, __dmul_rn(x[i], x[j]));
                        t = __dadd_rn(t, (i == j) ? p.eps_v : 0.0);
                        const int64_t o = (int64_t)tri_idx(D, i, j) * C;
                        const double v = __dadd_rn(__dmul_rn(c1, gV[o]), __dmul_rn(c2, t));             // :463
                        Vc[tri_idx(D, i, j)] = v;
                        if (live) gV[o] = v;
                    }
                if (live) {
#pragma unroll
                    for (int i = 0; i < D; ++i) gxav[(int64_t)i * C] = xb[i];
                }
            } else {
                // Welford over x_0..x_it: xav = running mean, V = sum of outer products of deviations
                const double inv_n = 1.0 / (fi + 1.0);
                const double sc = p.S_d / fi;                               // S_d * M2/(n-1), n = it+1
                double dl[D], mu[D];
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    const double m = gxav[(int64_t)i * C];
                    dl[i] = x[i] - m;
                    mu[i] = fma(dl[i], inv_n, m);
                }
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = i; j < D; ++j) {
                        const int64_t o = (int64_t)tri_idx(D, i, j) * C;
                        const double v = fma(dl[i], x[j] - mu[j], gV[o]);
                        Vc[tri_idx(D, i, j)] = sc * v + ((i == j) ? p.S_d * p.eps_v : 0.0);
                        if (live) gV[o] = v;
                    }
                if (live) {
#pragma unroll
                    for (int i = 0; i < D; ++i) gxav[(int64_t)i * C] = mu[i];
                }
            }
            if (refresh) {                                  // :486-487, :591-595
                double Rn[NT];
                if (chol_upper_packed<D>(Vc, Rn)) {
                    if (live) {
#pragma unroll
                        for (int i = 0; i < NT; ++i) gR[(int64_t)i * C] = Rn[i];
                    }
                } else {
                    status |= ST_NOT_PD;                    // scipy raises LinAlgError here (:487)
                }
            }
            mean_r += r;                                    // :419 (unclipped)
        } else {
            mean_r += alpha;                                // :170
        }

        // ---- outputs
        if (writer) {
            if (p.out.trace_lp1) p.out.trace_lp1[s * C + c] = lp1;
            if (p.out.trace_lp2) p.out.trace_lp2[s * C + c] = lp2;
            if (p.out.trace_acc) p.out.trace_acc[s * C + c] = (uint8_t)acc;
            if (it % p.thin == 0) {
                // running (Welford) moments of the KEPT states, held in global memory so that they cost
                // no registers in the data loop; input of the moment reduction (K4: pooled covariance, R-hat)
                {
                    const int64_t wc = p.st.wcount[c] + 1;
                    p.st.wcount[c] = wc;
                    const double inv_n = 1.0 / (double)wc;
                    double dl[D], mu[D];
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        const double m = p.st.wmean[i * C + c];
                        dl[i] = x[i] - m;
                        mu[i] = fma(dl[i], inv_n, m);
                        p.st.wmean[i * C + c] = mu[i];
                    }
#pragma unroll
                    for (int i = 0; i < D; ++i)
#pragma unroll
                        for (int j = i; j < D; ++j) {
                            double *q = &p.st.wM2[tri_idx(D, i, j) * C + c];
                            *q = fma(dl[i], x[j] - mu[j], *q);
                        }
                }
                const int64_t row = it / p.thin - p.it0 / p.thin - 1;
                if (p.out.samples) {
#pragma unroll
                    for (int i = 0; i < D; ++i) p.out.samples[(row * D + i) * C + c] = x[i];
                }
                if (p.out.lp_kept) p.out.lp_kept[row * C + c] = lp;
            }
        }
    }

    // ---- store state
    if (writer) {
#pragma unroll
        for (int i = 0; i < D; ++i) p.st.x[i * C + c] = x[i];
        p.st.lp[c] = lp;
        p.st.mean_r[c] = mean_r;
        p.st.n_rej[c] += n_rej;
        p.st.n_eval[c] += n_eval;
        p.st.status[c] = status;
        if (injected) {
            p.streams.cur_n[c] = cn;
            p.streams.cur_u[c] = cu;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Initial state: log-posterior of x0 (mha.py:102).  Also used for plain evaluation of S parameter
// vectors (one thread per vector, G = 1).
// ------------------------------------------------------------------------------------------------
template <class M, int D>
__global__ void __launch_bounds__(256) eval_logpost_kernel(const __grid_constant__ ProblemDev prob, const double *X /*[d][S]*/, int64_t S,
                                                           double *lp_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *smem_rec = reinterpret_cast<double *>(smem_raw + 128);
    stage_records(smem_rec, bar, prob.records, (uint32_t)prob.n_points * (uint32_t)(M::NC * 8));
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S) return;
    double x[D];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = X[i * S + t];
    lp_out[t] = log_posterior<M, D, 1>(x, prob, smem_rec, 0, 1u << (threadIdx.x & 31));
}

// Model.run_model for S parameter vectors: f_out[k][s] (point-major, coalesced over samples).
template <class M, int D>
__global__ void __launch_bounds__(256) eval_model_kernel(const __grid_constant__ ProblemDev prob, const double *X /*[d][S]*/, int64_t S,
                                                         double *f_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *smem_raw_rec = reinterpret_cast<double *>(smem_raw + 128);
    stage_records(smem_raw_rec, bar, prob.raw, (uint32_t)prob.n_points * (uint32_t)(M::NCR * 8));
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S) return;
    double x[D], theta[M::NPARAM];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = X[i * S + t];
    full_theta<M, D>(x, prob, theta);
    typename M::Ctx ctx;
    M::begin(ctx, theta, prob);
    typename M::Seq seq;
    M::seq_init(ctx, seq, prob);
    for (int k = 0; k < prob.n_points; ++k) {
        double raw[M::NCR];
        load_record<M::NCR>(smem_raw_rec, k, raw);
        f_out[(int64_t)k * S + t] = M::value(ctx, seq, raw);
    }
}

}  // namespace pbu
