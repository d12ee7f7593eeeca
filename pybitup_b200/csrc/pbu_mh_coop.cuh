// K1, cooperative layout: a TEAM of G lanes owns Q <= G chains.
//
// In the replicated layout of pbu_mh_kernel.cuh every lane of a G-lane group carries the whole state of its
// chain(s) and repeats the scalar part of the iteration (random draws, proposal, accept, adaptation).  Here the work is
// divided instead:
//   scalar part   lane s of the team (s < Q) OWNS chain s: it alone draws, proposes, accepts and adapts that chain;
//   data loop     the proposals of the Q chains are exchanged with warp shuffles, then every lane walks its share of
//                 the points (index = s mod G) evaluating ALL Q chains, so that each record read from shared memory
//                 and each x operand feeds Q chains (tools/ubench: U = 2 points x Q = 8 chains in flight reaches
//                 32.4 TFLOP/s, 95% of the DFMA peak, against 28.5 for two chains per thread);
//   reduction     the Q sums of squares are butterfly-reduced over the G lanes; the owner keeps its own.
// Nothing is computed twice, whatever G is.  Reference lines as in pbu_mh_kernel.cuh.
#pragma once
#include "pbu_mh_kernel.cuh"

namespace pbu {

template <int G>
__device__ __forceinline__ double team_bcast(double v, int owner_sub, int lane, unsigned team_mask) {
    return __shfl_sync(team_mask, v, lane_chain<G>(lane) + owner_sub * (32 / G));
}

template <class M, int D, int G, int Q, int ALGO, bool STREAM>
__device__ __forceinline__ void mh_coop_loop(const MhParams &p, TileStream<M::NC> &ts, const int64_t team, const int sub) {
    static_assert(Q <= G, "a team needs at least one lane per chain");
    constexpr bool ADAPT = (ALGO & 1) != 0;
    constexpr bool DELAYED = (ALGO & 2) != 0;
    constexpr int NT = tri_size(D);
    constexpr int U = (Q >= 4) ? ((M::UNROLL > 2) ? 2 : M::UNROLL) : M::UNROLL;
    constexpr bool PARK = !STREAM;               // the host reserves the slots behind the resident records (plan_launch)
    const int64_t C = p.st.n_chains;
    const int lane = threadIdx.x & 31;
    const unsigned team_mask = lane_group_mask<G>(lane);
    const bool injected = p.streams.normals != nullptr;

    // ---- the chain this lane owns
    const bool owner = sub < Q;
    int64_t c = team * Q + (owner ? sub : 0);
    const bool live = owner && c < C;
    if (c >= C) c = C - 1;                       // padding lanes shadow the last chain, never write
    double x[D], lp, mean_r;
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = p.st.x[i * C + c];
    lp = p.st.lp[c];
    mean_r = p.st.mean_r[c];
    int32_t n_rej = 0, n_eval = 0, status = p.st.status[c];
    // The adaptive state (V_i, X_av, R) of the own chain lives in global memory (chain-minor, stride C): its addresses are
    // rebuilt from the kernel parameters where they are used, so that no pointer stays live across the data loop (at the
    // 128-register cap every register the loop loses costs operand-reuse slots, i.e. DFMA issue rate).  Staging this
    // state in shared memory for the launch was tried and measured slower (3.18 against 2.96 ms per 100 iterations).
    const int64_t sS = C;
    int32_t to_refresh = ADAPT ? (int32_t)(p.updating_it - p.it0 % p.updating_it) : 0;
    int32_t to_keep = (int32_t)(p.thin - p.it0 % p.thin);
    int32_t row = 0;
    const int32_t n_steps = (int32_t)p.n_steps;      // a launch is at most 2^31 iterations (checked on the host)
    for (int32_t s = 0; s < n_steps; ++s) {
        const int64_t it = p.it0 + s + 1;
        double y[D] = {}, u = 1.0, lp1 = 0.0, lp2 = nan(""), r = 0.0, alpha = 0.0;
        int acc = 0;
        bool want = live;                    // lanes beyond the team's chains (and padding owners) only lend their arithmetic to the data loop
#pragma unroll 1
        for (int stage = 0; stage < (DELAYED ? 2 : 1); ++stage) {
            bool inb = false;
            if (want) {
                double z[D];
                if (injected) {
                    if (!draw_injected<D>(p.streams, c, live, live, 1u << lane, z, u)) status |= ST_STREAM_EXHAUSTED;
                } else {
                    draw_philox<D>(p.seed, p.chain_offset + (uint64_t)c, (uint64_t)it, (uint32_t)stage, z, u);
                }
                const double scale = (stage == 0) ? 1.0 : p.gamma_dr;
                double sj[D];                                                   // new = cur + scale * R @ z, R upper (:193, :543)
                upper_matvec<D>([&](int i, int j) { return ADAPT ? p.st.R[(int64_t)tri_idx(D, i, j) * sS + c] : p.R0[tri_idx(D, i, j)]; }, z, sj);
#pragma unroll
                for (int i = 0; i < D; ++i) y[i] = __dadd_rn(x[i], (stage == 0) ? sj[i] : __dmul_rn(scale, sj[i]));      // (1.0 * v == v exactly)
                n_eval += 1;
            }
            double Xown[D], ldj_own;                // X = parametrization_backward(Y) (bayesian_inference.py:50); X = Y when not re-parametrised
            par_backward<D>(p.prob, y, Xown, ldj_own);
            if (want) inb = inside_box<D>(Xown, p.prob);
            // ---- the own chain's state sits the data loop out in shared memory (one slot per thread, component stride
            // MH_MAX_BLOCK doubles): x, y, lp, mean_r, u are 2D + 3 register pairs the loop can spend on keeping all Q x U
            // dependency chains in flight instead (at the 128-register cap ptxas otherwise interleaves only two of them)
            volatile double *const park = PARK ? reinterpret_cast<volatile double *>(TileStream<M::NC>::smem() + p.state_smem_off) + threadIdx.x : nullptr;
            if constexpr (PARK) {
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    park[i * MH_MAX_BLOCK] = x[i];
                    park[(D + i) * MH_MAX_BLOCK] = y[i];
                }
                park[(2 * D) * MH_MAX_BLOCK] = lp;
                park[(2 * D + 1) * MH_MAX_BLOCK] = mean_r;
                park[(2 * D + 2) * MH_MAX_BLOCK] = u;
            }
            // ---- the team evaluates the proposals of its Q chains together
            typename M::Ctx ctx[Q];
            typename M::Seq seq[Q];
            double J[Q][U], gdummy[Q][1];
            bool any = false;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                double yq[D];
#pragma unroll
                for (int i = 0; i < D; ++i) yq[i] = team_bcast<G>(Xown[i], q, lane, team_mask);
                any = any || (__shfl_sync(team_mask, (int)(want && inb), lane_chain<G>(lane) + q * (32 / G)) != 0);
                begin_chain<M, D>(yq, p.prob, ctx[q], seq[q]);
#pragma unroll
                for (int uu = 0; uu < U; ++uu) J[q][uu] = 0.0;
            }
            if (STREAM || any) {
                if constexpr (!STREAM) {
                    tile_accumulate<M, G, Q, U, false>(ctx, seq, reinterpret_cast<const double *>(TileStream<M::NC>::smem() + 128), p.prob.n_points,
                                                       sub, J, gdummy);
                } else {
                    const double *const records = p.prob.records;
                    const int n = p.prob.n_points;
                    ts.begin_pass(records, n, p.tile_points);
                    const int nt = TileStream<M::NC>::n_tiles(n, p.tile_points);
                    for (int t = 0; t < nt; ++t) {
                        int cnt;
                        const double *tile = ts.acquire(n, p.tile_points, t, cnt);
                        if (any) tile_accumulate<M, G, Q, U, false>(ctx, seq, tile, cnt, sub, J, gdummy);
                        ts.release(records, n, p.tile_points, t);
                    }
                }
            }
            if constexpr (PARK) {
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    x[i] = park[i * MH_MAX_BLOCK];
                    y[i] = park[(D + i) * MH_MAX_BLOCK];
                }
                lp = park[(2 * D) * MH_MAX_BLOCK];
                mean_r = park[(2 * D + 1) * MH_MAX_BLOCK];
                u = park[(2 * D + 2) * MH_MAX_BLOCK];
            }
            double Jown = 0.0;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                double Jq = 0.0;
#pragma unroll
                for (int uu = 0; uu < U; ++uu) Jq += J[q][uu];
                Jq = group_sum<G>(Jq, team_mask);
                if (q == sub) Jown = Jq;
            }
            bool again_own = false;
            if (want) {
                double ldj = 0.0;                    // -log(det_jac), affine in Y: rebuilt here rather than kept across the data loop
                if (p.prob.par_any) {
                    ldj = p.prob.par_ldj_c0;
#pragma unroll
                    for (int i = 0; i < D; ++i) ldj = fma(p.prob.par_ldj_s[i], y[i], ldj);
                }
                const double lpy = inb ? (p.prob.log_prior_const + ldj) + (-0.5 * (Jown + p.prob.J_const)) : -INFINITY;     // :44-64
                bool accept;
                if (stage == 0) {
                    lp1 = lpy;
                    r = (lp1 == -INFINITY) ? 0.0 : pbu_exp_full(lp1 - lp);                      // :200-205
                    if (r != r) {
                        status |= ST_NAN_RATIO;
                        if (p.nan_rejects) r = 0.0;
                    }
                    alpha = py_min1(r);                                                 // :207
                    accept = u < alpha;                                                 // :214
                } else {
                    lp2 = lpy;
                    const double a12 = py_min1(pbu_exp_full(lp1 - lp2));                         // :547-548
                    double M2 = 0.0, M4 = 0.0;                                          // inv_V = diag(1/R_ii^2)
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        const double rii = ADAPT ? p.st.R[(int64_t)tri_idx(D, i, i) * sS + c] : p.R0[tri_idx(D, i, i)];
                        const double ir = 1.0 / rii;
                        const double iv = ir * ir;
                        const double y1i = p.st.P[(int64_t)i * C + c];     // first-stage proposal, parked in global memory
                        const double d1 = y1i - y[i];
                        const double d2 = y1i - x[i];
                        M2 += (d1 * iv) * d1;
                        M4 += (d2 * iv) * d2;
                    }
                    double r2 = pbu_exp_full(lp2 - lp) * (pbu_exp_full(-0.5 * M2) / pbu_exp_full(-0.5 * M4)) * (1.0 - a12) / (1.0 - alpha);     // :555-556
                    if (r2 != r2) {
                        status |= ST_NAN_RATIO;
                        if (p.nan_rejects) r2 = 0.0;
                    }
                    accept = u < py_min1(r2);                                           // :558-562
                }
                if (accept) {
#pragma unroll
                    for (int i = 0; i < D; ++i) x[i] = y[i];
                    lp = lpy;
                    acc = stage + 1;
                    want = false;
                    if (p.st.map_lp != nullptr && stage == 0 && live) track_map<D>(p.st, c, y, lpy);       // :218, Q10
                } else if (DELAYED && stage == 0) {
                    // keep the rejected first proposal for the second-stage ratio (:549-556) out of the registers: the
                    // (otherwise unused) momentum array is this chain's scratch
                    if (live) {
#pragma unroll
                        for (int i = 0; i < D; ++i) p.st.P[(int64_t)i * C + c] = y[i];
                    }
                    again_own = true;                                                   // :535
                } else {
                    n_rej += 1;                                                         // :224, :567
                    want = false;
                }
            }
            // the second stage runs when any chain of the team (of the block, when streaming) needs it
            bool again = __any_sync(team_mask, again_own);
            if (STREAM) again = __syncthreads_or(again) != 0;
            if (!again) break;
        }

        const bool refresh = ADAPT && (--to_refresh == 0);           // it % updating_it == 0  (:470)
        if (refresh) to_refresh = p.updating_it;
        const bool keep = (--to_keep == 0);
        if (keep) to_keep = (int32_t)p.thin;
        // ---- covariance adaptation of the own chain (mha.py:432-487)
        if (ADAPT) {
            const double fi = (double)it;
            double Vc[NT];
            if (!p.am_welford) {
                double xa[D], xb[D];
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    xa[i] = (it == 1) ? x[i] : p.st.xav[(int64_t)i * sS + c];                      // X_av aliases the state at it == 1 (Q3)
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
                        const int64_t o = (int64_t)tri_idx(D, i, j) * sS;
                        const double v = __dadd_rn(__dmul_rn(c1, p.st.V[o + c]), __dmul_rn(c2, t));             // :463
                        Vc[tri_idx(D, i, j)] = v;
                        if (live) p.st.V[o + c] = v;
                    }
                if (live) {
#pragma unroll
                    for (int i = 0; i < D; ++i) p.st.xav[(int64_t)i * sS + c] = xb[i];
                }
            } else {
                const double inv_n = 1.0 / (fi + 1.0);
                const double sc = p.S_d / fi;
                double dl[D], mu[D];
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    const double m = p.st.xav[(int64_t)i * sS + c];
                    dl[i] = x[i] - m;
                    mu[i] = fma(dl[i], inv_n, m);
                }
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = i; j < D; ++j) {
                        const int64_t o = (int64_t)tri_idx(D, i, j) * sS;
                        const double v = fma(dl[i], x[j] - mu[j], p.st.V[o + c]);
                        Vc[tri_idx(D, i, j)] = sc * v + ((i == j) ? p.S_d * p.eps_v : 0.0);
                        if (live) p.st.V[o + c] = v;
                    }
                if (live) {
#pragma unroll
                    for (int i = 0; i < D; ++i) p.st.xav[(int64_t)i * sS + c] = mu[i];
                }
            }
            if (refresh) {                                  // :486-487, :591-595
                double Rn[NT];
                if (chol_upper_packed<D>(Vc, Rn)) {
                    if (live) {
#pragma unroll
                        for (int i = 0; i < NT; ++i) p.st.R[(int64_t)i * sS + c] = Rn[i];
                    }
                } else {
                    status |= ST_NOT_PD;
                }
            }
            mean_r += r;                                    // :419 (unclipped, Q9)
        } else {
            mean_r += alpha;                                // :170
        }

        // ---- outputs of the own chain
        if (live) {
            if (p.out.trace_lp1) p.out.trace_lp1[(int64_t)s * C + c] = lp1;
            if (p.out.trace_lp2) p.out.trace_lp2[(int64_t)s * C + c] = lp2;
            if (p.out.trace_acc) p.out.trace_acc[(int64_t)s * C + c] = (uint8_t)acc;
            if (keep) {
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
                        double *w2 = &p.st.wM2[tri_idx(D, i, j) * C + c];
                        *w2 = fma(dl[i], x[j] - mu[j], *w2);
                    }
                if (p.out.samples && c < p.out.n_sample_chains) {
#pragma unroll
                    for (int i = 0; i < D; ++i) p.out.samples[((int64_t)row * D + i) * p.out.n_sample_chains + c] = x[i];
                }
                if (p.out.lp_kept) p.out.lp_kept[(int64_t)row * C + c] = lp;
            }
        }
        if (keep) row += 1;
    }

    if (live) {
#pragma unroll
        for (int i = 0; i < D; ++i) p.st.x[i * C + c] = x[i];
        p.st.lp[c] = lp;
        p.st.mean_r[c] = mean_r;
        p.st.n_rej[c] += n_rej;
        p.st.n_eval[c] += n_eval;
        p.st.status[c] = status;
    }
}

// Cooperative kernel: teams of G lanes own G chains.  MIXED: from warp p.n_full_warps on, teams are 2G lanes wide (and
// still own G chains): half a unit of work per warp, for scheduler-level balance (see mh_step_kernel).
template <class M, int D, int G, int ALGO, bool STREAM, bool MIXED>
__global__ void __launch_bounds__(MH_MAX_BLOCK, M::MIN_BLOCKS) mh_coop_kernel(const __grid_constant__ MhParams p) {
    TileStream<M::NC> ts;
    M::block_init();
    ts.init(p.prob.records, p.prob.n_points, p.tile_points);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // Phase stagger.  Every warp alternates a long data loop (the FP64 pipe at its throttle) with a scalar phase of ~1,200
    // mostly dependent instructions (draws, exp, adaptation, global loads) during which it issues almost no DFMA.  Warps of a
    // block start together and take equally long per iteration, so the four warps a scheduler holds would go through their
    // scalar phases TOGETHER and leave the pipe idle; started a quarter of an iteration apart they stay that way for the
    // whole launch, and one warp's scalar phase hides behind the other three's loops.
    if (p.stagger_cycles > 0) {
        const long long wait = (long long)((warp >> 2) & 3) * p.stagger_cycles;      // warps w, w+4, w+8, w+12 share a scheduler
        const long long t0 = clock64();
        while (clock64() - t0 < wait) {
        }
    }
    if constexpr (!MIXED) {
        const int64_t warp_g = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
        mh_coop_loop<M, D, G, G, ALGO, STREAM>(p, ts, warp_g * (32 / G) + lane_chain<G>(lane), lane_sub<G>(lane));
    } else {
        const int64_t base = (int64_t)blockIdx.x * p.groups_per_block;
        if (warp < p.n_full_warps) {
            mh_coop_loop<M, D, G, G, ALGO, STREAM>(p, ts, base + warp * (32 / G) + lane_chain<G>(lane), lane_sub<G>(lane));
        } else {
            mh_coop_loop<M, D, 2 * G, G, ALGO, STREAM>(p, ts, base + p.n_full_warps * (32 / G) + (warp - p.n_full_warps) * (16 / G) + lane_chain<2 * G>(lane),
                                                      lane_sub<2 * G>(lane));
        }
    }
}

}  // namespace pbu
