// K2: fused Ito-SDE step kernel (ito_SDE.run_algorithm, metropolis_hastings_algorithms.py:932-1040,
// with GradientBasedMCMC.adapt_covariance :736-777), many iterations per launch.
//
// Per chain and iteration (reference line numbers):
//   adaptation window [starting_it, end_it+1]                              :741-777
//     it == starting_it : X_av = mean, V_i = S_d (cov + eps I) of ALL previous states   :742-746
//     it <= end_it      : Haario recursion on (X_av, V_i)                               :750-758
//     it % update_it==0 : C = V_i, L_c = chol(C) (upper), inv_L_c                       :767-777
//   WP    = sqrt(h) z                      (d normals, no uniform)         :990
//   xi_n  = xi + h/2 C P                                                   :991
//   g     = grad log-likelihood(xi_n)      (Analytical :573-623 | forward differences :1043-1088)
//   P'    = hfm/hfp P + h/hfp g + sqrt(f0)/hfp inv_L_c WP                  :998-1000
//   xi'   = xi_n + h/2 C P'                                                :1001
// The reference initialises the window from its chain CSV re-read from disk (6-decimal text);
// the kernel uses a Welford accumulation of the exact states instead (documented deviation).
//
// Mapping is the one of the MH kernel: G lanes per chain split the points, Q chains per thread share
// every record read, records staged through shared memory (resident or streamed in tiles).  The state
// (xi, P) stays in registers; the d x d preconditioner C, inv_L_c and the adaptation state live in global
// memory, chain-minor, touched once per iteration.
#pragma once
#include "pbu_mh_kernel.cuh"

namespace pbu {

struct IsdeParams {
    ProblemDev prob;
    ChainState st;            // x = xi, P, R = packed upper inv_L_c, V = packed V_i / Welford M2, xav = X_av / Welford mean
    StreamsDev streams;
    pbu_run_outputs out;
    double *Cmat;             // [d(d+1)/2][C] packed symmetric preconditioner C_approx per chain
    int64_t it0, n_steps, thin, n_groups;
    uint64_t seed, chain_offset;
    double h, f0, S_d, eps_id;
    int64_t starting_it, update_it, end_it;
    int32_t tile_points;
    int32_t fd_gradient;      // 1: forward differences with relative step 1e-10 (mha.py:726, :1069-1088)
    int32_t hmc_steps;        // HMC: leapfrog steps L (mha.py:819); the step size epsilon travels in h
    int32_t pad0;
    int32_t n_full_warps;     // mixed Ito-SDE kernels: warps of a block with G lanes per chain (the rest use 2G)
    int32_t groups_per_block; // mixed Ito-SDE kernels: chain groups per block
};

// symmetric packed matrix (upper triangle stored) times vector, reading the matrix from global memory
template <int D>
__device__ __forceinline__ void symv_global(const double *gC, int64_t C, const double (&v)[D], double (&out)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const int a = (i <= j) ? i : j, b = (i <= j) ? j : i;
            s = __dadd_rn(s, __dmul_rn(gC[(int64_t)tri_idx(D, a, b) * C], v[j]));     // np.matmul: plain dot product, no FMA
        }
        out[i] = s;
    }
}

// inverse of an upper-triangular packed matrix by back substitution (oracle: inv_upper_plain)
template <int D>
__device__ __forceinline__ void inv_upper_packed(const double (&R)[tri_size(D)], double (&X)[tri_size(D)]) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
        X[tri_idx(D, j, j)] = 1.0 / R[tri_idx(D, j, j)];
#pragma unroll
        for (int i = j - 1; i >= 0; --i) {
            double s = 0.0;
#pragma unroll
            for (int k = i + 1; k <= j; ++k) s += R[tri_idx(D, i, k)] * X[tri_idx(D, k, j)];
            X[tri_idx(D, i, j)] = -s / R[tri_idx(D, i, i)];
        }
    }
}

// GradientBasedMCMC.adapt_covariance (mha.py:736-777) for one chain: before the window a Welford accumulation of the
// states (rows 0..it-1 of the chain, which the reference re-reads from its CSV at :744); at starting_it the window
// opens with X_av = mean, V_i = S_d (cov + eps I); inside it the Haario recursion; every update_it iterations the
// preconditioner C, its upper Cholesky factor and the inverse of that are refreshed.  x = current_val.
template <int D>
__device__ __forceinline__ void adapt_preconditioner(const IsdeParams &p, int64_t it, bool upd, const double (&x)[D], bool live, double *gC,
                                                     double *gIL, double *gV, double *gxav, int64_t C, int32_t &status) {
    constexpr int NT = tri_size(D);
    if (it <= p.starting_it) {
        // Welford over rows 0..it-1 of the chain (the rows the reference re-reads from its CSV at :744)
        const double inv_n = 1.0 / (double)it;
        double dl[D], mu[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double m = (it == 1) ? 0.0 : gxav[(int64_t)i * C];
            dl[i] = x[i] - m;
            mu[i] = fma(dl[i], inv_n, m);
            if (live) gxav[(int64_t)i * C] = mu[i];
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = i; j < D; ++j) {
                const int64_t o = (int64_t)tri_idx(D, i, j) * C;
                double v = fma(dl[i], x[j] - mu[j], (it == 1) ? 0.0 : gV[o]);
                if (it == p.starting_it)        // V_i = S_d (cov + eps I), np.cov with ddof = 1   (:745-746)
                    v = p.S_d * (v / (double)(it - 1) + ((i == j) ? p.eps_id : 0.0));
                if (live) gV[o] = v;
            }
    } else if (it <= p.end_it) {
        const double fi = (double)it;
        double xa[D], xb[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            xa[i] = gxav[(int64_t)i * C];
            xb[i] = __dadd_rn(x[i], __dmul_rn(fi / (fi + 1.0), __dadd_rn(xa[i], -x[i])));      // :752
        }
        const double c1 = (fi - 1.0) / fi, c2 = p.S_d / fi;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = i; j < D; ++j) {
                double t = __dadd_rn(__dmul_rn(fi, __dmul_rn(xa[i], xa[j])), -__dmul_rn(fi + 1.0, __dmul_rn(xb[i], xb[j])));
                t = __dadd_rn(t, __dmul_rn(x[i], x[j]));
                t = __dadd_rn(t, (i == j) ? p.eps_id : 0.0);
                const int64_t o = (int64_t)tri_idx(D, i, j) * C;
                const double v = __dadd_rn(__dmul_rn(c1, gV[o]), __dmul_rn(c2, t));                       // :753
                if (live) gV[o] = v;
            }
        if (live) {
#pragma unroll
            for (int i = 0; i < D; ++i) gxav[(int64_t)i * C] = xb[i];
        }
    }
    if (it >= p.starting_it && it <= p.end_it + 1 && upd) {                         // :767-777
        double Vc[NT], Lc[NT], IL[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) Vc[i] = gV[(int64_t)i * C];
        if (chol_upper_packed<D>(Vc, Lc)) {
            inv_upper_packed<D>(Lc, IL);
            if (live) {
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    gC[(int64_t)i * C] = Vc[i];
                    gIL[(int64_t)i * C] = IL[i];
                }
            }
        } else {
            status |= ST_NOT_PD;
        }
    }

}

// Gradient of the log-likelihood (and the log-posterior value) at x for the Q chains of a thread.
template <class M, int D, int G, int Q, bool STREAM>
__device__ __forceinline__ void grad_log_like(const double (&yv)[Q][D], const ProblemDev &prob, TileStream<M::NC> &ts, int tile_points, int sub,
                                              unsigned group_mask, double (&grad)[Q][D], double (&lp)[Q]) {
    constexpr int U = (M::UNROLL > 2) ? 2 : M::UNROLL;
    constexpr int NP = M::NPARAM;
    typename M::Ctx ctx[Q];
    typename M::Seq seq[Q];
    double J[Q][U], g[Q][M::HAS_GRAD ? NP : 1], gd[Q][D];
    double x[Q][D], ldj[Q];        // X = parametrization_backward(Y) (bayesian_inference.py:84), -log(det_jac(X))
#pragma unroll
    for (int q = 0; q < Q; ++q) par_backward<D>(prob, yv[q], x[q], ldj[q]);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
#pragma unroll
        for (int u = 0; u < U; ++u) J[q][u] = 0.0;
#pragma unroll
        for (int j = 0; j < (M::HAS_GRAD ? NP : 1); ++j) g[q][j] = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) gd[q][i] = 0.0;
    }
    if constexpr (!STREAM) {
        // one pass per model of the likelihood (value :500-524, gradient :596-612); a single model is one block
        const double *const base = reinterpret_cast<const double *>(TileStream<M::NC>::smem() + 128);
        const int nb = prob.n_blocks;
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
#pragma unroll
            for (int q = 0; q < Q; ++q) begin_chain_block<M, D>(x[q], prob, b, ctx[q], seq[q]);
            tile_accumulate<M, G, Q, U, M::HAS_GRAD>(ctx, seq, base + (size_t)prob.blk[b].first * M::NC, prob.blk[b].count, sub, J, g);
            if constexpr (M::HAS_GRAD) {          // model-parameter gradient -> uncertain parameters, by this block's name map
#pragma unroll
                for (int q = 0; q < Q; ++q) {
#pragma unroll
                    for (int i = 0; i < D; ++i) {
#pragma unroll
                        for (int j = 0; j < NP; ++j)
                            if (D == NP ? (j == i) : (prob.blk[b].unc_idx[i] == j)) gd[q][i] += g[q][j];
                    }
#pragma unroll
                    for (int j = 0; j < NP; ++j) g[q][j] = 0.0;
                }
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < Q; ++q) begin_chain<M, D>(x[q], prob, ctx[q], seq[q]);        // streamed records: one model (host-checked)
        const double *const records = prob.records;
        const int n = prob.n_points;
        ts.begin_pass(records, n, tile_points);
        const int nt = TileStream<M::NC>::n_tiles(n, tile_points);
        for (int t = 0; t < nt; ++t) {
            int cnt;
            const double *tile = ts.acquire(n, tile_points, t, cnt);
            tile_accumulate<M, G, Q, U, M::HAS_GRAD>(ctx, seq, tile, cnt, sub, J, g);
            ts.release(records, n, tile_points, t);
        }
        if constexpr (M::HAS_GRAD) {
#pragma unroll
            for (int q = 0; q < Q; ++q)
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j < NP; ++j)
                        if (D == NP ? (j == i) : (prob.unc_idx[i] == j)) gd[q][i] = g[q][j];
        }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        double Jq = 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) Jq += J[q][u];
        Jq = group_sum<G>(Jq, group_mask) + prob.J_const;
        lp[q] = inside_box<D>(x[q], prob) ? (prob.log_prior_const + ldj[q]) + (-0.5 * Jq) : -INFINITY;
#pragma unroll
        for (int i = 0; i < D; ++i) grad[q][i] = group_sum<G>(gd[q][i], group_mask);
        par_grad<D>(prob, x[q], grad[q]);          // d/dY = inv_jac^T d/dX - grad log det_jac  (:86-89, :604-636)
    }
}

template <class M, int D, int G, int Q, bool STREAM>
__device__ __forceinline__ void isde_chain_loop(const IsdeParams &p, TileStream<M::NC> &ts, const int64_t group, const int sub) {
    constexpr int NT = tri_size(D);
    const int64_t C = p.st.n_chains;
    const int lane = threadIdx.x & 31;
    const unsigned group_mask = lane_group_mask<G>(lane);
    const bool injected = p.streams.normals != nullptr;

    int64_t c[Q];
    bool live[Q];
    double xi[Q][D], P[Q][D], lp[Q];
    int32_t status[Q], n_eval[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        c[q] = group + (int64_t)q * p.n_groups;
        live[q] = c[q] < C;
        if (!live[q]) c[q] = C - 1;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            xi[q][i] = p.st.x[i * C + c[q]];
            P[q][i] = p.st.P[i * C + c[q]];
        }
        lp[q] = p.st.lp[c[q]];
        status[q] = p.st.status[c[q]];
        n_eval[q] = 0;
    }
    const double hfm = 1.0 - p.h * p.f0 / 4.0, hfp = 1.0 + p.h * p.f0 / 4.0;      // :928-929
    const double sqrt_h = sqrt(p.h), sqrt_f0 = sqrt(p.f0);

    int32_t to_update = (int32_t)(p.update_it - p.it0 % p.update_it);     // countdowns instead of a 64-bit modulo per iteration
    int32_t to_keep = (int32_t)(p.thin - p.it0 % p.thin);
    int64_t row = 0;
    for (int64_t s = 0; s < p.n_steps; ++s) {
        const int64_t it = p.it0 + s + 1;
        const bool upd = (--to_update == 0);                            // it % update_it == 0
        if (upd) to_update = (int32_t)p.update_it;
        const bool keep = (--to_keep == 0);                             // it % thin == 0
        if (keep) to_keep = (int32_t)p.thin;
        double xn[Q][D], WP[Q][D];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int64_t cq = c[q];
            double *const gC = p.Cmat + cq;
            double *const gIL = p.st.R + cq;
            double *const gV = p.st.V + cq;
            double *const gxav = p.st.xav + cq;
            // ---- adaptation of the preconditioner (:736-777); current_val = xi_nm (:958)
            // one lane of the chain's group runs the recursion on the global-memory state, then the group synchronises (all
            // lanes doing it would race on V_i / X_av: a late lane would apply the update to already updated values)
            if (G == 1 || sub == 0) adapt_preconditioner<D>(p, it, upd, xi[q], live[q], gC, gIL, gV, gxav, C, status[q]);
            if constexpr (G > 1) __syncwarp(group_mask);

            // ---- first half step (:990-991)
            double z[D], udummy;
            if (injected) {
                if (!draw_injected<D, false>(p.streams, cq, live[q], live[q] && sub == 0, group_mask, z, udummy)) status[q] |= ST_STREAM_EXHAUSTED;
            } else {
                draw_philox<D>(p.seed, p.chain_offset + (uint64_t)cq, (uint64_t)it, 0u, z, udummy);
            }
            double CP[D];
            symv_global<D>(gC, C, P[q], CP);
#pragma unroll
            for (int i = 0; i < D; ++i) {
                WP[q][i] = __dmul_rn(sqrt_h, z[i]);
                xn[q][i] = __dadd_rn(xi[q][i], __dmul_rn(p.h / 2.0, CP[i]));
            }
        }

        // ---- gradient at xi_n (:994)
        double grad[Q][D], lpn[Q];
        if (!p.fd_gradient) {
            grad_log_like<M, D, G, Q, STREAM>(xn, p.prob, ts, p.tile_points, sub, group_mask, grad, lpn);
#pragma unroll
            for (int q = 0; q < Q; ++q) n_eval[q] += 1;
        } else {
            // computeGradientFD, forward scheme on the log-POSTERIOR (:1069-1088): d + 1 evaluations
            bool want[Q];
            double f0v[Q], fp[Q], xp[Q][D];
#pragma unroll
            for (int q = 0; q < Q; ++q) want[q] = true;
            log_posterior<M, D, G, Q, STREAM>(xn, want, p.prob, ts, p.tile_points, sub, group_mask, f0v);
#pragma unroll 1
            for (int i = 0; i < D; ++i) {
                double delta[Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
#pragma unroll
                    for (int j = 0; j < D; ++j) xp[q][j] = xn[q][j];
                    double xv = 0.0;
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        if (j == i) xv = xn[q][j];
                    delta[q] = __dmul_rn(xv, 1e-10);
                    if (delta[q] == 0.0) delta[q] = 1e-10;
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        if (j == i) xp[q][j] = __dadd_rn(xv, delta[q]);
                }
                log_posterior<M, D, G, Q, STREAM>(xp, want, p.prob, ts, p.tile_points, sub, group_mask, fp);
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const double gi = (fp[q] - f0v[q]) / delta[q];
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        if (j == i) grad[q][j] = gi;
                }
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                lpn[q] = f0v[q];
                n_eval[q] += D + 1;
            }
        }

        // ---- momentum and second half step (:998-1001)
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int64_t cq = c[q];
            const double *const gC = p.Cmat + cq;
            const double *const gIL = p.st.R + cq;
            double Pn[D];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                double lw = 0.0;                                   // (inv_L_c @ WP)_i, inv_L_c upper triangular
#pragma unroll
                for (int j = i; j < D; ++j) lw = __dadd_rn(lw, __dmul_rn(gIL[(int64_t)tri_idx(D, i, j) * C], WP[q][j]));
                const double a = __dmul_rn(hfm / hfp, P[q][i]);
                const double b = __dmul_rn(p.h / hfp, grad[q][i]);        // h/hfp * (-grad_phi), grad_phi = -grad
                const double cc = __dmul_rn(sqrt_f0 / hfp, lw);
                Pn[i] = __dadd_rn(__dadd_rn(a, b), cc);
            }
            double CP[D];
            symv_global<D>(gC, C, Pn, CP);
#pragma unroll
            for (int i = 0; i < D; ++i) {
                xi[q][i] = __dadd_rn(xn[q][i], __dmul_rn(p.h / 2.0, CP[i]));
                P[q][i] = Pn[i];
            }
            lp[q] = lpn[q];            // log-posterior at xi_n (the point the model was evaluated at)
        }
        // ---- MAP estimate: the reference evaluates the posterior once more, at the NEW state xi_np (:1016-1019)
        if (p.st.map_lp != nullptr) {
            bool wantm[Q];
            double lpm[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) wantm[q] = true;
            log_posterior<M, D, G, Q, STREAM>(xi, wantm, p.prob, ts, p.tile_points, sub, group_mask, lpm);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                n_eval[q] += 1;
                if (live[q] && sub == 0) track_map<D>(p.st, c[q], xi[q], lpm[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int64_t cq = c[q];
            if (live[q] && sub == 0) {
                if (p.out.trace_lp1) p.out.trace_lp1[s * C + cq] = lpn[q];
                if (keep) {
                    {
                        const int64_t wc = p.st.wcount[cq] + 1;
                        p.st.wcount[cq] = wc;
                        const double inv_n = 1.0 / (double)wc;
                        double dl[D], mu[D];
#pragma unroll
                        for (int i = 0; i < D; ++i) {
                            const double m = p.st.wmean[i * C + cq];
                            dl[i] = xi[q][i] - m;
                            mu[i] = fma(dl[i], inv_n, m);
                            p.st.wmean[i * C + cq] = mu[i];
                        }
#pragma unroll
                        for (int i = 0; i < D; ++i)
#pragma unroll
                            for (int j = i; j < D; ++j) {
                                double *w2 = &p.st.wM2[tri_idx(D, i, j) * C + cq];
                                *w2 = fma(dl[i], xi[q][j] - mu[j], *w2);
                            }
                    }
                    if (p.out.samples && cq < p.out.n_sample_chains) {
#pragma unroll
                        for (int i = 0; i < D; ++i) p.out.samples[(row * D + i) * p.out.n_sample_chains + cq] = xi[q][i];
                    }
                    if (p.out.aux && cq < p.out.n_sample_chains) {
#pragma unroll
                        for (int i = 0; i < D; ++i) p.out.aux[(row * D + i) * p.out.n_sample_chains + cq] = P[q][i];
                    }
                    if (p.out.lp_kept) p.out.lp_kept[row * C + cq] = lp[q];
                }
            }
        }
        if (keep) row += 1;
    }

#pragma unroll
    for (int q = 0; q < Q; ++q) {
        if (live[q] && sub == 0) {
            const int64_t cq = c[q];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                p.st.x[i * C + cq] = xi[q][i];
                p.st.P[i * C + cq] = P[q][i];
            }
            p.st.lp[cq] = lp[q];
            p.st.n_eval[cq] += n_eval[q];
            p.st.status[cq] = status[q];
        }
    }
}

// The kernel proper.  MIXED: the last warps of every block (from warp index p.n_full_warps on) give each chain 2G lanes
// instead of G, i.e. carry half as many chains -- half a unit of work, as in mh_step_kernel.  With equal warps the 2,048
// warps of 131,072 chains (G = 1, Q = 2) fill 128 blocks of 16 warps and leave 20 of the 148 SMs idle; three full warps and
// one half warp per scheduler (896 chains per block, 147 blocks) use them all.  Streamed records: every warp of either
// width walks the same tiles, the wide ones split each tile's points over their two lanes.
template <class M, int D, int G, int Q, bool STREAM, bool MIXED>
__global__ void __launch_bounds__(MH_MAX_BLOCK, M::MIN_BLOCKS) isde_step_kernel(const __grid_constant__ IsdeParams p) {
    TileStream<M::NC> ts;
    M::block_init();
    ts.init(p.prob.records, p.prob.n_points, p.tile_points);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if constexpr (!MIXED) {
        const int64_t warp_g = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
        isde_chain_loop<M, D, G, Q, STREAM>(p, ts, warp_g * (32 / G) + lane_chain<G>(lane), lane_sub<G>(lane));
    } else {
        const int64_t base = (int64_t)blockIdx.x * p.groups_per_block;
        if (warp < p.n_full_warps) {
            isde_chain_loop<M, D, G, Q, STREAM>(p, ts, base + warp * (32 / G) + lane_chain<G>(lane), lane_sub<G>(lane));
        } else {
            isde_chain_loop<M, D, 2 * G, Q, STREAM>(p, ts, base + p.n_full_warps * (32 / G) + (warp - p.n_full_warps) * (16 / G) + lane_chain<2 * G>(lane),
                                                   lane_sub<2 * G>(lane));
        }
    }
}


// ------------------------------------------------------------------------------------------------
// K2b: fused Hamiltonian Monte Carlo step kernel (HamiltonianMonteCarlo, mha.py:779-884, driven by
// MetropolisHastings.run_algorithm :158-182).  Per chain and iteration:
//   adaptation of the preconditioner as for ISDE                              :826, :736-777
//   p = z @ inv_L_c^T = inv_L_c z (d normals);  K_cur = sum((p C) p)/2        :829-831, :862-866
//   p += eps g_cur / 2                                                        :836
//   L times:  x += eps (p C);  g = grad(x);  p += eps g  (not after the last) :842-850
//   p += eps g / 2;  p = -p;  K_new = sum((p C) p)/2                          :852-855, :871-874
//   r = exp(U_cur - U_new + K_cur - K_new), U = -log posterior;  u < min(1, r) accepts          :876-881, :210-224
// The stored gradient of the current state (p.st.P) follows the state on acceptance (:221).  The log-posterior at the
// end of the trajectory comes from the same data pass as the last gradient.
// ------------------------------------------------------------------------------------------------
template <class M, int D, int G, int Q, bool STREAM>
__global__ void __launch_bounds__(MH_MAX_BLOCK, M::MIN_BLOCKS) hmc_step_kernel(const __grid_constant__ IsdeParams p) {
    TileStream<M::NC> ts;
    M::block_init();
    ts.init(p.prob.records, p.prob.n_points, p.tile_points);

    const int64_t C = p.st.n_chains;
    const int lane = threadIdx.x & 31;
    const int sub = lane_sub<G>(lane);
    const int64_t group = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 / G) + lane_chain<G>(lane);
    const unsigned group_mask = lane_group_mask<G>(lane);
    const bool injected = p.streams.normals != nullptr;

    int64_t c[Q];
    bool live[Q];
    double x[Q][D], gcur[Q][D], lp[Q], mean_r[Q];
    int32_t status[Q], n_eval[Q], n_rej[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        c[q] = group + (int64_t)q * p.n_groups;
        live[q] = c[q] < C;
        if (!live[q]) c[q] = C - 1;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            x[q][i] = p.st.x[i * C + c[q]];
            gcur[q][i] = p.st.P[i * C + c[q]];
        }
        lp[q] = p.st.lp[c[q]];
        mean_r[q] = p.st.mean_r[c[q]];
        status[q] = p.st.status[c[q]];
        n_eval[q] = 0;
        n_rej[q] = 0;
    }
    const double eps = p.h;

    // gradient (and log-posterior) at xn for all chains of the thread
    auto gradient = [&](const double (&xn)[Q][D], double (&g)[Q][D], double (&lpn)[Q]) {
        if (!p.fd_gradient) {
            grad_log_like<M, D, G, Q, STREAM>(xn, p.prob, ts, p.tile_points, sub, group_mask, g, lpn);
#pragma unroll
            for (int q = 0; q < Q; ++q) n_eval[q] += 0;       // compute_grad_log_value does not call compute_log_value
        } else {
            bool want[Q];
            double fp[Q], xp[Q][D];
#pragma unroll
            for (int q = 0; q < Q; ++q) want[q] = true;
            log_posterior<M, D, G, Q, STREAM>(xn, want, p.prob, ts, p.tile_points, sub, group_mask, lpn);
#pragma unroll 1
            for (int i = 0; i < D; ++i) {
                double delta[Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    double xv = 0.0;
#pragma unroll
                    for (int j = 0; j < D; ++j) {
                        xp[q][j] = xn[q][j];
                        if (j == i) xv = xn[q][j];
                    }
                    delta[q] = __dmul_rn(xv, 1e-10);
                    if (delta[q] == 0.0) delta[q] = 1e-10;
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        if (j == i) xp[q][j] = __dadd_rn(xv, delta[q]);
                }
                log_posterior<M, D, G, Q, STREAM>(xp, want, p.prob, ts, p.tile_points, sub, group_mask, fp);
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const double gi = (fp[q] - lpn[q]) / delta[q];
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        if (j == i) g[q][j] = gi;
                }
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) n_eval[q] += D + 1;
        }
    };

    int32_t to_update = (int32_t)(p.update_it - p.it0 % p.update_it);
    int32_t to_keep = (int32_t)(p.thin - p.it0 % p.thin);
    int64_t row = 0;
    for (int64_t s = 0; s < p.n_steps; ++s) {
        const int64_t it = p.it0 + s + 1;
        const bool upd = (--to_update == 0);
        if (upd) to_update = (int32_t)p.update_it;
        const bool keep = (--to_keep == 0);
        if (keep) to_keep = (int32_t)p.thin;
        double xn[Q][D], mom[Q][D], Kcur[Q], u[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int64_t cq = c[q];
            double *const gC = p.Cmat + cq;
            double *const gIL = p.st.R + cq;
            if (G == 1 || sub == 0) adapt_preconditioner<D>(p, it, upd, x[q], live[q], gC, gIL, p.st.V + cq, p.st.xav + cq, C, status[q]);
            if constexpr (G > 1) __syncwarp(group_mask);
            double z[D];
            if (injected) {
                if (!draw_injected<D, true>(p.streams, cq, live[q], live[q] && sub == 0, group_mask, z, u[q])) status[q] |= ST_STREAM_EXHAUSTED;
            } else {
                draw_philox<D>(p.seed, p.chain_offset + (uint64_t)cq, (uint64_t)it, 0u, z, u[q]);
            }
            double pc[D], CP[D];
#pragma unroll
            for (int i = 0; i < D; ++i) {                  // p = z @ inv_L_c^T, inv_L_c upper triangular (:830)
                double acc = 0.0;
#pragma unroll
                for (int j = i; j < D; ++j) acc = __dadd_rn(acc, __dmul_rn(z[j], gIL[(int64_t)tri_idx(D, i, j) * C]));
                pc[i] = acc;
            }
            symv_global<D>(gC, C, pc, CP);
            double k = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) k = __dadd_rn(k, __dmul_rn(CP[i], pc[i]));
            Kcur[q] = k / 2.0;                             // :862-866
#pragma unroll
            for (int i = 0; i < D; ++i) {
                mom[q][i] = __dadd_rn(pc[i], -(__dmul_rn(eps, -gcur[q][i]) / 2.0));      // :836
                xn[q][i] = x[q][i];
            }
        }
        double gnew[Q][D], lpn[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            lpn[q] = lp[q];
#pragma unroll
            for (int i = 0; i < D; ++i) gnew[q][i] = gcur[q][i];
        }
#pragma unroll 1
        for (int l = 0; l < p.hmc_steps; ++l) {
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                double CP[D];
                symv_global<D>(p.Cmat + c[q], C, mom[q], CP);
#pragma unroll
                for (int i = 0; i < D; ++i) xn[q][i] = __dadd_rn(xn[q][i], __dmul_rn(eps, CP[i]));            // :843
            }
            gradient(xn, gnew, lpn);
            if (l != p.hmc_steps - 1) {
#pragma unroll
                for (int q = 0; q < Q; ++q)
#pragma unroll
                    for (int i = 0; i < D; ++i) mom[q][i] = __dadd_rn(mom[q][i], -__dmul_rn(eps, -gnew[q][i]));  // :848
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int64_t cq = c[q];
            double pf[D], CP[D];
#pragma unroll
            for (int i = 0; i < D; ++i) pf[i] = -__dadd_rn(mom[q][i], -(__dmul_rn(eps, -gnew[q][i]) / 2.0));     // :852, :855
            symv_global<D>(p.Cmat + cq, C, pf, CP);
            double k = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) k = __dadd_rn(k, __dmul_rn(CP[i], pf[i]));
            const double Knew = k / 2.0;
            n_eval[q] += 1;                                // distr_fun(new_val) (:868)
            double r = pbu_exp_full(__dadd_rn(__dadd_rn(__dadd_rn(-lp[q], lpn[q]), Kcur[q]), -Knew));                    // :879
            if (r != r) {
                status[q] |= ST_NAN_RATIO;
                r = 0.0;
            }
            const double alpha = py_min1(r);
            if (u[q] < alpha) {
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    x[q][i] = xn[q][i];
                    gcur[q][i] = gnew[q][i];
                }
                lp[q] = lpn[q];
                if (p.st.map_lp != nullptr && live[q] && sub == 0) track_map<D>(p.st, cq, xn[q], lpn[q]);    // accept_reject -> estimate_max (:218)
            } else {
                n_rej[q] += 1;
            }
            mean_r[q] += alpha;                            // :170
            if (live[q] && sub == 0) {
                if (p.out.trace_lp1) p.out.trace_lp1[s * C + cq] = lpn[q];
                if (keep) {
                    const int64_t wc = p.st.wcount[cq] + 1;
                    p.st.wcount[cq] = wc;
                    const double inv_n = 1.0 / (double)wc;
                    double dl[D], mu[D];
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        const double m = p.st.wmean[i * C + cq];
                        dl[i] = x[q][i] - m;
                        mu[i] = fma(dl[i], inv_n, m);
                        p.st.wmean[i * C + cq] = mu[i];
                    }
#pragma unroll
                    for (int i = 0; i < D; ++i)
#pragma unroll
                        for (int j = i; j < D; ++j) {
                            double *w2 = &p.st.wM2[tri_idx(D, i, j) * C + cq];
                            *w2 = fma(dl[i], x[q][j] - mu[j], *w2);
                        }
                    if (p.out.samples && cq < p.out.n_sample_chains) {
#pragma unroll
                        for (int i = 0; i < D; ++i) p.out.samples[(row * D + i) * p.out.n_sample_chains + cq] = x[q][i];
                    }
                    if (p.out.lp_kept) p.out.lp_kept[row * C + cq] = lp[q];
                }
            }
        }
        if (keep) row += 1;
    }

#pragma unroll
    for (int q = 0; q < Q; ++q) {
        if (live[q] && sub == 0) {
            const int64_t cq = c[q];
#pragma unroll
            for (int i = 0; i < D; ++i) {
                p.st.x[i * C + cq] = x[q][i];
                p.st.P[i * C + cq] = gcur[q][i];
            }
            p.st.lp[cq] = lp[q];
            p.st.mean_r[cq] = mean_r[q];
            p.st.n_rej[cq] += n_rej[q];
            p.st.n_eval[cq] += n_eval[q];
            p.st.status[cq] = status[q];
        }
    }
}

// gradient of the log-posterior at S points (initial stored gradient of HMC, mha.py:815): G[i][s]
// gradient of the log-likelihood at one point per thread, analytical or by forward differences (:1069-1088)
template <class M, int D, bool ST>
__device__ __forceinline__ void eval_grad_point(const double (&x)[1][D], const ProblemDev &prob, TileStream<M::NC> &ts, int tile_points, unsigned mask,
                                                int fd_gradient, double (&g)[1][D], double (&lp)[1]) {
    if (!fd_gradient) {
        grad_log_like<M, D, 1, 1, ST>(x, prob, ts, tile_points, 0, mask, g, lp);
    } else {
        const bool want[1] = {true};
        double fp[1], xp[1][D];
        log_posterior<M, D, 1, 1, ST>(x, want, prob, ts, tile_points, 0, mask, lp);
        for (int i = 0; i < D; ++i) {
            double xv = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                xp[0][j] = x[0][j];
                if (j == i) xv = x[0][j];
            }
            double delta = __dmul_rn(xv, 1e-10);
            if (delta == 0.0) delta = 1e-10;
#pragma unroll
            for (int j = 0; j < D; ++j)
                if (j == i) xp[0][j] = __dadd_rn(xv, delta);
            log_posterior<M, D, 1, 1, ST>(xp, want, prob, ts, tile_points, 0, mask, fp);
#pragma unroll
            for (int j = 0; j < D; ++j)
                if (j == i) g[0][j] = (fp[0] - lp[0]) / delta;
        }
    }
}

template <class M, int D>
__global__ void __launch_bounds__(256) eval_grad_kernel(const __grid_constant__ ProblemDev prob, const double *X /*[d][S]*/, int64_t S,
                                                        double *G_out /*[d][S]*/, int tile_points, int fd_gradient) {
    TileStream<M::NC> ts;
    M::block_init();
    ts.init(prob.records, prob.n_points, tile_points);
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool livet = t < S;
    if (!livet) t = S - 1;
    double x[1][D], g[1][D], lp[1];
#pragma unroll
    for (int i = 0; i < D; ++i) x[0][i] = X[i * S + t];
    const unsigned mask = 1u << (threadIdx.x & 31);
    if (TileStream<M::NC>::resident(prob.n_points, tile_points))        // (uniform branch) the resident path walks the model blocks
        eval_grad_point<M, D, false>(x, prob, ts, tile_points, mask, fd_gradient, g, lp);
    else
        eval_grad_point<M, D, true>(x, prob, ts, tile_points, mask, fd_gradient, g, lp);
    if (livet) {
#pragma unroll
        for (int i = 0; i < D; ++i) G_out[i * S + t] = g[0][i];
    }
}

}  // namespace pbu
