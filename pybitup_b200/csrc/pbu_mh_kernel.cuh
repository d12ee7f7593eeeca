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
#include <type_traits>
#include "pbu_common.cuh"
#include "pbu_models.cuh"

namespace pbu {

#ifndef PBU_MAX_BLOCK
#define PBU_MAX_BLOCK 512
#endif
constexpr int MH_MAX_BLOCK = PBU_MAX_BLOCK;
// models whose lookup tables are the conflict-free replicas (Model::Rep, 192 KB of shared memory): ONE block of up to 1024
// threads per SM
template <class M, class = void>
struct StepBounds {
    static constexpr int MAX_BLOCK = MH_MAX_BLOCK, MIN_BLOCKS = M::MIN_BLOCKS;
};
template <class M>
struct StepBounds<M, std::enable_if_t<M::IS_REP>> {
    static constexpr int MAX_BLOCK = 1024, MIN_BLOCKS = 1;
};

// ------------------------------------------------------------------------------------------------
// Shared-memory staging of the point records with bulk TMA copies (cp.async.bulk + mbarrier).
//   resident  : every record fits in shared memory; loaded once per launch, read by all chains of the
//               block for all iterations, no synchronisation after the first wait;
//   streaming : the records are cut into tiles that cycle through a ring of S slots once per posterior
//               evaluation ("pass").  All threads of the block walk the tiles together: wait for the
//               tile to land, use it, __syncthreads, then thread 0 refills the slot with the tile S
//               ahead, so the copy of tile t+1.. overlaps the arithmetic on tile t.  Global traffic is
//               one read of the records per block and pass, served by L2 (the record set is shared by
//               every block of every launch).
// ------------------------------------------------------------------------------------------------
// The stream keeps ONE register of state (tiles consumed so far); everything else is re-derived from kernel
// parameters (constant bank) and the compile-time record size, so the resident path costs the data loop nothing.
template <int NCOL>
struct TileStream {
    static constexpr int S = 2;                      // ring slots
    static constexpr uint32_t REC = NCOL * 8u;       // bytes per record
    uint32_t tseq;                                   // tiles consumed so far by this block (selects slot and mbarrier phase)

    static __device__ __forceinline__ unsigned char *smem() {
        extern __shared__ __align__(128) unsigned char pbu_smem[];
        return pbu_smem;                             // [S mbarriers, padded to 128 B][slot 0][slot 1]
    }
    static __device__ __forceinline__ uint64_t *full() { return reinterpret_cast<uint64_t *>(smem()); }
    static __device__ __forceinline__ char *slot_ptr(uint32_t slot, int tile_points) {
        return reinterpret_cast<char *>(smem() + 128) + (size_t)slot * (uint32_t)tile_points * REC;
    }
    static __device__ __forceinline__ bool resident(int n, int tile_points) { return tile_points >= n; }
    static __device__ __forceinline__ int n_tiles(int n, int tile_points) { return (n + tile_points - 1) / tile_points; }

    static __device__ __forceinline__ void issue(const double *records, int n, int tile_points, int tile, uint32_t slot) {
        const int first = tile * tile_points;
        const int cnt = (n - first < tile_points) ? (n - first) : tile_points;
        const uint32_t bytes = (uint32_t)cnt * REC;
        uint64_t *bar = full() + slot;
        mbar_expect_tx(bar, bytes);
        const uint32_t CH = 32768u;
        for (uint32_t off = 0; off < bytes; off += CH) {
            const uint32_t m = (bytes - off < CH) ? (bytes - off) : CH;
            tma_load_1d(slot_ptr(slot, tile_points) + off, reinterpret_cast<const char *>(records) + (size_t)first * REC + off, m, bar);
        }
    }
    __device__ __forceinline__ void init(const double *records, int n, int tile_points) {
        tseq = 0;
        if (threadIdx.x == 0) {
            for (int i = 0; i < S; ++i) mbar_init(full() + i, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (resident(n, tile_points)) {
            if (threadIdx.x == 0) issue(records, n, tile_points, 0, 0);
            mbar_wait(full(), 0);
        }
    }
    __device__ __forceinline__ void begin_pass(const double *records, int n, int tile_points) {
        if (resident(n, tile_points)) return;
        if (threadIdx.x == 0) {
            const int nt = n_tiles(n, tile_points);
            for (int i = 0; i < S && i < nt; ++i) issue(records, n, tile_points, i, (tseq + i) % S);
        }
    }
    // points of tile t of the current pass; blocks until the tile has landed
    __device__ __forceinline__ const double *acquire(int n, int tile_points, int t, int &count) {
        if (resident(n, tile_points)) {
            count = n;
            return reinterpret_cast<const double *>(smem() + 128);
        }
        const uint32_t slot = tseq % S;
        mbar_wait(full() + slot, (tseq / S) & 1u);
        const int first = t * tile_points;
        count = (n - first < tile_points) ? (n - first) : tile_points;
        return reinterpret_cast<const double *>(slot_ptr(slot, tile_points));
    }
    __device__ __forceinline__ void release(const double *records, int n, int tile_points, int t) {
        if (resident(n, tile_points)) return;
        __syncthreads();                                    // every thread is done with the slot
        if (threadIdx.x == 0 && t + S < n_tiles(n, tile_points)) issue(records, n, tile_points, t + S, tseq % S);
        tseq += 1;
    }
};

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
// GRAD, the gradient sum of Likelihood.compute_grad_log_value (:573-623), for the Q chains a thread
// owns: every record read from shared memory feeds Q chains, which divides the shared-memory
// wavefronts per FP64 operation by Q (a broadcast LDS still costs one wavefront per 8 bytes per
// lane, and that, not the bank bandwidth, is what binds a one-chain-per-thread loop).
// Several experimental runs are folded on the host with the exact identity
//   sum_r (y_r - f)^2 = n_r (ybar - f)^2 + sum_r (y_r - ybar)^2
// so a record carries w = sqrt(n_r)/(std_y gamma) and w*ybar, and the second term is the constant
// prob.J_const: the device loop costs the same whatever n_runs is.
// J[Q][U] are independent partial sums (one per chain and point in flight).
// ------------------------------------------------------------------------------------------------
template <class M, int G, int Q, int U, bool GRAD>
__device__ __forceinline__ void tile_accumulate(const typename M::Ctx (&ctx)[Q], typename M::Seq (&seq)[Q], const double *tile, int n,
                                                int sub, double (&J)[Q][U], double (&g)[Q][GRAD ? M::NPARAM : 1]) {
    constexpr int NC = M::NC;
    if (M::SEQUENTIAL) {
#pragma unroll 2
        for (int k = 0; k < n; ++k) {
            double rec[1][NC];
            load_record<NC>(tile, k, rec[0]);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                double t[1];
                M::template resid_batch<1>(ctx[q], seq[q], rec, t);
                J[q][0] = fma(t[0], t[0], J[q][0]);
            }
        }
    } else {
        int k = sub;
        if constexpr (M::HAS_MULTI && !GRAD && (Q > 1)) {
            // H batches of U points per trip: the loop bookkeeping shared by H * U points, the later batches' records in flight
            // while the first is evaluated (H = 2; 4 for 16-byte records, which take one load and two registers per point)
            constexpr int H = (NC <= 2) ? 4 : 2;
#pragma unroll 1
            for (; k + (H * U - 1) * G < n; k += H * U * G) {
                double rec[H][U][NC];
#pragma unroll
                for (int h = 0; h < H; ++h)
#pragma unroll
                    for (int u = 0; u < U; ++u) load_record<NC>(tile, k + (h * U + u) * G, rec[h][u]);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    double t[Q][U];
                    M::template resid_multi<Q, U>(ctx, rec[h], t);
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int q = 0; q < Q; ++q) J[q][u] = fma(t[q][u], t[q][u], J[q][u]);
                }
            }
        }
#pragma unroll 1
        for (; k + (U - 1) * G < n; k += U * G) {
            double rec[U][NC];
#pragma unroll
            for (int u = 0; u < U; ++u) load_record<NC>(tile, k + u * G, rec[u]);
            if constexpr (M::HAS_MULTI && !GRAD && (Q > 1)) {
                double t[Q][U];
                M::template resid_multi<Q, U>(ctx, rec, t);
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int q = 0; q < Q; ++q) J[q][u] = fma(t[q][u], t[q][u], J[q][u]);
            } else {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    double t[U];
                    M::template resid_batch<U>(ctx[q], seq[q], rec, t);
#pragma unroll
                    for (int u = 0; u < U; ++u) J[q][u] = fma(t[u], t[u], J[q][u]);
                    if constexpr (GRAD) {
#pragma unroll
                        for (int u = 0; u < U; ++u) M::grad_acc(ctx[q], rec[u], t[u], g[q]);
                    }
                }
            }
        }
#pragma unroll 1
        for (; k < n; k += G) {
            double rec[1][NC];
            load_record<NC>(tile, k, rec[0]);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                double t[1];
                M::template resid_batch<1>(ctx[q], seq[q], rec, t);
                J[q][0] = fma(t[0], t[0], J[q][0]);
                if constexpr (GRAD) M::grad_acc(ctx[q], rec[0], t[0], g[q]);
            }
        }
    }
}

// Lane layout of a warp when G lanes share a chain: the warp carries 32/G chains; lane l works for chain l % (32/G) and
// takes the points with index = (l / (32/G)) mod G.  Lanes that read the same record are then CONTIGUOUS (whole
// quarter-warps for G <= 4), which is what the shared-memory pipe wants: a 16-byte load is served one quarter-warp
// at a time, and a quarter whose lanes all read one address is a single broadcast wavefront.
template <int G>
__device__ __forceinline__ int lane_sub(int lane) { return lane / (32 / G); }
template <int G>
__device__ __forceinline__ int lane_chain(int lane) { return lane % (32 / G); }
template <int G>
__device__ __forceinline__ unsigned lane_group_mask(int lane) {
    unsigned m = 0u;
#pragma unroll
    for (int j = 0; j < G; ++j) m |= 1u << (lane_chain<G>(lane) + j * (32 / G));
    return m;
}
template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned group_mask) {
#pragma unroll
    for (int s = 1; s < G; s <<= 1) v += __shfl_xor_sync(group_mask, v, s * (32 / G));
    return v;
}

// Full parameter vector from the uncertain ones (Model.run_model, bayesian_inference.py:361-369).
template <class M, int D>
__device__ __forceinline__ void full_theta(const double (&x)[D], const double *theta_nom, const int32_t *unc_idx, double (&theta)[M::NPARAM]) {
#pragma unroll
    for (int q = 0; q < M::NPARAM; ++q) {
        double v = theta_nom[q];
#pragma unroll
        for (int i = 0; i < D; ++i)
            if (unc_idx[i] == q) v = x[i];
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

// Model constants of one chain from its uncertain parameters.  When every parameter is uncertain
// (D == NPARAM) the reference aliases param = var_param (bayesian_inference.py:368-369): theta IS x.
template <class M, int D>
__device__ __forceinline__ void begin_chain(const double (&x)[D], const ProblemDev &prob, typename M::Ctx &ctx, typename M::Seq &seq) {
    if constexpr (D == M::NPARAM) {
        M::begin(ctx, x, prob.consts);
    } else {
        double theta[M::NPARAM];
        full_theta<M, D>(x, prob.theta_nom, prob.unc_idx, theta);
        M::begin(ctx, theta, prob.consts);
    }
    M::seq_init(ctx, seq, prob.consts);
}
// ... of model block b of the likelihood (Likelihood.models, bayesian_inference.py:500-507): own nominal parameters,
// own name map, own constants.  The block index is a run-time value: the small vectors are read from the constant
// bank with an indexed load, once per evaluation.
template <class M, int D>
__device__ __forceinline__ void begin_chain_block(const double (&x)[D], const ProblemDev &prob, int b, typename M::Ctx &ctx, typename M::Seq &seq) {
    const ModelBlockDev &blk = prob.blk[b];
    if constexpr (D == M::NPARAM) {
        M::begin(ctx, x, blk.consts);
    } else {
        double theta[M::NPARAM];
        full_theta<M, D>(x, blk.theta_nom, blk.unc_idx, theta);
        M::begin(ctx, theta, blk.consts);
    }
    M::seq_init(ctx, seq, blk.consts);
}

// BayesianPosterior.compute_log_value (bayesian_inference.py:44-64) with the identity
// parametrisation, for the Q chains of a thread: -inf outside the prior box WITHOUT evaluating the
// model (:53-55), otherwise log_prior - log(1) + (-(1/2) J).  In resident mode the data pass is skipped
// when no chain of the thread needs it.
template <class M, int D, int G, int Q, bool STREAM>
__device__ __forceinline__ void log_posterior(const double (&yv)[Q][D], const bool (&want)[Q], const ProblemDev &prob, TileStream<M::NC> &ts,
                                              int tile_points, int sub, unsigned group_mask, double (&lp)[Q]) {
    constexpr int U = M::UNROLL;   // tools/ubench: U = 4 points x Q = 2 chains in flight reaches 83% of the DFMA peak at 12 warps/SM
    bool inb[Q], any = false;
    double x[Q][D], ldj[Q];        // X = parametrization_backward(Y), -log(det_jac(X))  (:50, :61)
#pragma unroll
    for (int q = 0; q < Q; ++q) par_backward<D>(prob, yv[q], x[q], ldj[q]);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        inb[q] = want[q] && inside_box<D>(x[q], prob);
        any = any || inb[q];
        lp[q] = -INFINITY;
    }
    if (!STREAM && !any) return;          // streaming: the whole block walks the tiles together
    typename M::Ctx ctx[Q];
    typename M::Seq seq[Q];
    double J[Q][U], gdummy[Q][1];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
#pragma unroll
        for (int u = 0; u < U; ++u) J[q][u] = 0.0;
    }
    if constexpr (!STREAM) {
        // one pass per model of the likelihood (bayesian_inference.py:500-524); a single model is one block
        const double *const base = reinterpret_cast<const double *>(TileStream<M::NC>::smem() + 128);
        const int nb = prob.n_blocks;
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            // block 0 is read from the top-level fields (compile-time constant-bank addresses): a single model pays nothing
            const double *recs = base;
            int cnt = prob.n_points;
            if (b == 0) {
#pragma unroll
                for (int q = 0; q < Q; ++q) begin_chain<M, D>(x[q], prob, ctx[q], seq[q]);
                if (nb > 1) cnt = prob.blk[0].count;
            } else {
#pragma unroll
                for (int q = 0; q < Q; ++q) begin_chain_block<M, D>(x[q], prob, b, ctx[q], seq[q]);
                recs = base + (size_t)prob.blk[b].first * M::NC;
                cnt = prob.blk[b].count;
            }
            tile_accumulate<M, G, Q, U, false>(ctx, seq, recs, cnt, sub, J, gdummy);
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
            if (any) tile_accumulate<M, G, Q, U, false>(ctx, seq, tile, cnt, sub, J, gdummy);
            ts.release(records, n, tile_points, t);
        }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        double Jq = 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u) Jq += J[q][u];
        Jq = group_sum<G>(Jq, group_mask) + prob.J_const;
        if (inb[q]) lp[q] = (prob.log_prior_const + ldj[q]) + (-0.5 * Jq);
    }
}

// ------------------------------------------------------------------------------------------------
// Random draws for one proposal stage: D standard normals and one uniform.
// ------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void draw_philox(uint64_t seed, uint64_t chain, uint64_t it, uint32_t stage, double (&z)[D], double &u) {
    if constexpr (D == 3 || D == 4) {
        // ONE Philox4x32-10 block per proposal stage: two Box-Muller pairs from the top 24 bits of the four words, the
        // acceptance uniform (32 bits) from their low bytes.  (Two blocks were a fifth of the instructions of a chain-step
        // of the small linear model, BASELINE configuration 1.)
        const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        const uint4 ctr = make_uint4((uint32_t)chain, (uint32_t)it, (uint32_t)(it >> 32) ^ ((uint32_t)(chain >> 32) << 8) ^ (stage << 30), 0u);
        const uint4 r = Philox::round10(ctr, key);
        double a, b;
        box_muller24(r.x, r.y, a, b);
        z[0] = a;
        z[1] = b;
        box_muller24(r.z, r.w, a, b);
        z[2] = a;
        if constexpr (D == 4) z[3] = b;
        u = uniform32_low_bytes(r.x, r.y, r.z, r.w);
        return;
    }
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

// Parity mode: the chain's next D normals and one uniform from the injected streams; the cursors live
// in global memory (every lane of a group reads them, the writer lane advances them after the draw).
// A padding group (live == false: its chain index is clamped onto the last real chain) must NOT read that chain's
// cursors: the real owner advances them concurrently, so two lanes of the padding group could see different values,
// disagree on the proposal and on whether it is inside the prior box, and part ways before the group's shuffles.
template <int D, bool UNIFORM = true>
__device__ __forceinline__ bool draw_injected(const StreamsDev &s, int64_t c, bool live, bool advance, unsigned group_mask, double (&z)[D], double &u) {
    if (!live) {
#pragma unroll
        for (int i = 0; i < D; ++i) z[i] = 0.0;
        u = 1.0;
        return true;
    }
    const int64_t cn = s.cur_n[c], cu = s.cur_u[c];
    __syncwarp(group_mask);
    const bool ok = !(cn + D > s.n_normals || (UNIFORM && cu + 1 > s.n_uniforms));
    if (ok) {
        const double *zn = s.normals + c * s.n_normals + cn;
#pragma unroll
        for (int i = 0; i < D; ++i) z[i] = zn[i];
        u = UNIFORM ? s.uniforms[c * s.n_uniforms + cu] : 1.0;
        if (advance) {
            s.cur_n[c] = cn + D;
            if (UNIFORM) s.cur_u[c] = cu + 1;
        }
    } else {
#pragma unroll
        for (int i = 0; i < D; ++i) z[i] = 0.0;
        u = 1.0;
    }
    __syncwarp(group_mask);
    return ok;
}

// The same without advancing the cursors, `ahead` proposal stages past them (speculative delayed rejection: both stages'
// draws are looked at before it is known whether the second is consumed; advance_injected then moves the cursors by the
// number of stages the reference would have drawn).
template <int D>
__device__ __forceinline__ bool peek_injected(const StreamsDev &s, int64_t c, bool live, int ahead, double (&z)[D], double &u) {
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = 0.0;
    u = 1.0;
    if (!live) return true;
    const int64_t cn = s.cur_n[c] + (int64_t)ahead * D, cu = s.cur_u[c] + ahead;
    if (cn + D > s.n_normals || cu + 1 > s.n_uniforms) return false;
    const double *zn = s.normals + c * s.n_normals + cn;
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = zn[i];
    u = s.uniforms[c * s.n_uniforms + cu];
    return true;
}
template <int D>
__device__ __forceinline__ void advance_injected(const StreamsDev &s, int64_t c, int stages) {
    s.cur_n[c] += (int64_t)stages * D;
    s.cur_u[c] += stages;
}

// ------------------------------------------------------------------------------------------------
// The kernel.  ALGO = pbu_algo_id of the MH family: bit 0 adaptive (AM, DRAM), bit 1 delayed
// rejection (DR, DRAM).  Q = chains per thread; thread group g owns chains g, g + n_groups, ...
// ------------------------------------------------------------------------------------------------
template <class M, int D, int G, int ALGO, int Q, bool STREAM>
__device__ __forceinline__ void mh_chain_loop(const MhParams &p, TileStream<M::NC> &ts, const int64_t group, const int sub) {
    constexpr bool ADAPT = (ALGO & 1) != 0;
    constexpr bool DELAYED = (ALGO & 2) != 0;
    // SPECULATIVE DELAYED REJECTION (sequential models, two lanes per chain).  The points of an ODE model cannot be split
    // over lanes, so few chains leave the GPU latency-bound (one serial RK4 recurrence per thread).  But the second-stage
    // proposal x + gamma R z2 (mha.py:543) does not depend on the OUTCOME of the first stage: lane 0 of a chain's pair
    // evaluates the first proposal while lane 1 evaluates the second, in the same pass over the points; the acceptance
    // logic then runs as written, on values that are already there.  Philox draws are keyed by (chain, iteration, stage)
    // and injected draws are consumed as the reference consumes them, so the chains are those of the one-lane kernel: the
    // second evaluation is wasted when the first proposal is accepted, and a chain-step costs ONE evaluation time, not 1.9.
    constexpr bool SPEC = M::SEQUENTIAL && DELAYED && (G == 2);
    constexpr int NT = tri_size(D);
    const int64_t C = p.st.n_chains;
    const int lane = threadIdx.x & 31;
    const unsigned group_mask = lane_group_mask<G>(lane);
    const bool injected = p.streams.normals != nullptr;

    // ---- chain state kept in registers across the launch
    int64_t c[Q];
    bool live[Q];
    double x[Q][D], lp[Q], mean_r[Q];
    int32_t n_rej[Q], n_eval[Q], status[Q];     // n_rej / n_eval: per-launch increments of the int64 counters
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        c[q] = group + (int64_t)q * p.n_groups;
        live[q] = c[q] < C;
        if (!live[q]) c[q] = C - 1;             // padding slots shadow the last chain, never write
#pragma unroll
        for (int i = 0; i < D; ++i) x[q][i] = p.st.x[i * C + c[q]];
        lp[q] = p.st.lp[c[q]];
        mean_r[q] = p.st.mean_r[c[q]];
        n_rej[q] = 0;
        n_eval[q] = 0;
        status[q] = p.st.status[c[q]];
    }
    // The per-chain adaptive state (X_av, V_i, R) lives in global memory, chain-minor: every lane of a group
    // reads and, if live, writes the same values, so each thread only ever reads back its own stores.

    // iteration counters kept as small countdowns: a 64-bit modulo per iteration costs ~100 instructions
    int32_t to_refresh = ADAPT ? (int32_t)(p.updating_it - p.it0 % p.updating_it) : 0;   // iterations until it % updating_it == 0
    int32_t to_keep = (int32_t)(p.thin - p.it0 % p.thin);                                 // iterations until it % thin == 0
    int64_t row = 0;                                                                      // kept rows written by this launch
    const bool tracing = p.out.trace_lp1 || p.out.trace_lp2 || p.out.trace_acc;         // (parity runs) tested once per iteration, not per pointer
    const int32_t n_steps = (int32_t)p.n_steps;      // a launch is at most 2^31 iterations (checked on the host)
    for (int32_t s = 0; s < n_steps; ++s) {
        const int64_t it = p.it0 + s + 1;
        // Up to two proposal stages share ONE inlined copy of the posterior evaluation (code size):
        // stage 0 = the Metropolis proposal (mha.py:184-224), stage 1 = the delayed-rejection
        // proposal (mha.py:535-567), entered only by chains whose first proposal was rejected.
        double y[Q][D], y1[DELAYED ? Q : 1][DELAYED ? D : 1], u[Q], lpy[Q];
        double lp1[Q], lp2[Q], r[Q], alpha[Q];
        int acc[Q];
        bool want[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            // padding slots (chain index clamped onto the last real chain) sit the iteration out: they would read that
            // chain's adaptive state and stream cursors while its owner updates them, and the lanes of one padding group
            // could then disagree on a proposal and leave the group's shuffles at different points
            want[q] = live[q];
            acc[q] = 0;
            lp1[q] = 0.0;
            lp2[q] = nan("");
            r[q] = 0.0;
            alpha[q] = 0.0;
            u[q] = 1.0;
        }
        double ysp[SPEC ? 2 : 1][SPEC ? Q : 1][SPEC ? D : 1], usp[SPEC ? 2 : 1][SPEC ? Q : 1], lpsp[SPEC ? 2 : 1][SPEC ? Q : 1];
        if constexpr (SPEC) {
            double ysel[Q][D], lps[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
#pragma unroll
                for (int st = 0; st < 2; ++st) {
                    double z[D];
                    usp[st][q] = 1.0;
                    if (injected) {
                        // the second stage's draws may lie past the end of the streams when the first stage accepts: only an
                        // exhausted stage that is actually consumed counts (flagged below)
                        if (!peek_injected<D>(p.streams, c[q], live[q], st, z, usp[st][q])) usp[st][q] = -1.0;
                    } else {
                        draw_philox<D>(p.seed, p.chain_offset + (uint64_t)c[q], (uint64_t)it, (uint32_t)st, z, usp[st][q]);
                    }
                    double sj[D];
                    upper_matvec<D>([&](int i, int j) { return ADAPT ? p.st.R[(int64_t)tri_idx(D, i, j) * C + c[q]] : p.R0[tri_idx(D, i, j)]; }, z, sj);
                    const double scale = (st == 0) ? 1.0 : p.gamma_dr;
#pragma unroll
                    for (int i = 0; i < D; ++i) ysp[st][q][i] = __dadd_rn(x[q][i], (st == 0) ? sj[i] : __dmul_rn(scale, sj[i]));      // (1.0 * v == v exactly)
                }
#pragma unroll
                for (int i = 0; i < D; ++i) ysel[q][i] = (sub == 0) ? ysp[0][q][i] : ysp[1][q][i];
            }
            __syncwarp(group_mask);       // both lanes have read the stream cursors before the writer lane moves them
            // each lane evaluates ITS proposal over all points (no lane splitting: lanes-per-chain 1 inside the evaluation)
            log_posterior<M, D, 1, Q, STREAM>(ysel, want, p.prob, ts, p.tile_points, 0, 1u << lane, lps);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                lpsp[0][q] = __shfl_sync(group_mask, lps[q], lane_chain<G>(lane));
                lpsp[1][q] = __shfl_sync(group_mask, lps[q], lane_chain<G>(lane) + 32 / G);
            }
        }
#pragma unroll 1
        for (int stage = 0; stage < (DELAYED ? 2 : 1); ++stage) {
            const double scale = (stage == 0) ? 1.0 : p.gamma_dr;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                if (!want[q]) continue;
                if constexpr (SPEC) {
#pragma unroll
                    for (int i = 0; i < D; ++i) y[q][i] = (stage == 0) ? ysp[0][SPEC ? q : 0][SPEC ? i : 0] : ysp[SPEC ? 1 : 0][SPEC ? q : 0][SPEC ? i : 0];
                    u[q] = (stage == 0) ? usp[0][SPEC ? q : 0] : usp[SPEC ? 1 : 0][SPEC ? q : 0];
                    if (u[q] < 0.0) {             // the injected streams ended before this stage
                        status[q] |= ST_STREAM_EXHAUSTED;
#pragma unroll
                        for (int i = 0; i < D; ++i) y[q][i] = x[q][i];
                        u[q] = 1.0;
                    }
                    n_eval[q] += 1;
                    continue;
                }
                double z[D];
                if (injected) {
                    if (!draw_injected<D>(p.streams, c[q], live[q], live[q] && sub == 0, group_mask, z, u[q])) status[q] |= ST_STREAM_EXHAUSTED;
                } else {
                    draw_philox<D>(p.seed, p.chain_offset + (uint64_t)c[q], (uint64_t)it, (uint32_t)stage, z, u[q]);
                }
                // new = cur + scale * R @ z with R UPPER triangular (mha.py:193, :543)
                double sj[D];
                upper_matvec<D>([&](int i, int j) { return ADAPT ? p.st.R[(int64_t)tri_idx(D, i, j) * C + c[q]] : p.R0[tri_idx(D, i, j)]; }, z, sj);
#pragma unroll
                for (int i = 0; i < D; ++i) y[q][i] = __dadd_rn(x[q][i], (stage == 0) ? sj[i] : __dmul_rn(scale, sj[i]));      // (1.0 * v == v exactly)
                n_eval[q] += 1;
            }
            if constexpr (SPEC) {
#pragma unroll
                for (int q = 0; q < Q; ++q) lpy[q] = (stage == 0) ? lpsp[0][SPEC ? q : 0] : lpsp[SPEC ? 1 : 0][SPEC ? q : 0];
            } else {
                log_posterior<M, D, G, Q, STREAM>(y, want, p.prob, ts, p.tile_points, sub, group_mask, lpy);
            }
            bool again = false;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                if (!want[q]) continue;
                bool accept;
                if (stage == 0) {
                    lp1[q] = lpy[q];
                    r[q] = (lp1[q] == -INFINITY) ? 0.0 : pbu_exp_full(lp1[q] - lp[q]);         // :200-205 (full-domain exp: the difference may be anything)
                    if (r[q] != r[q]) {
                        status[q] |= ST_NAN_RATIO;
                        if (p.nan_rejects) r[q] = 0.0;
                    }
                    alpha[q] = py_min1(r[q]);                                           // :207
                    accept = u[q] < alpha[q];                                           // :214
                } else {
                    lp2[q] = lpy[q];
                    const double r12 = pbu_exp_full(lp1[q] - lp2[q]);                            // :547
                    const double a12 = py_min1(r12);
                    // inv_V = inv_R * inv_R^T ELEMENTWISE = diag(1/R_ii^2)  (:517-518, :594-595)
                    double M2 = 0.0, M4 = 0.0;
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        const double rii = ADAPT ? p.st.R[(int64_t)tri_idx(D, i, i) * C + c[q]] : p.R0[tri_idx(D, i, i)];
                        const double ir = 1.0 / rii;
                        const double iv = ir * ir;
                        const double d1 = y1[DELAYED ? q : 0][DELAYED ? i : 0] - y[q][i];
                        const double d2 = y1[DELAYED ? q : 0][DELAYED ? i : 0] - x[q][i];
                        M2 += (d1 * iv) * d1;
                        M4 += (d2 * iv) * d2;
                    }
                    double r2 = pbu_exp_full(lp2[q] - lp[q]) * (pbu_exp_full(-0.5 * M2) / pbu_exp_full(-0.5 * M4)) * (1.0 - a12) / (1.0 - alpha[q]);   // :555-556
                    if (r2 != r2) {
                        status[q] |= ST_NAN_RATIO;
                        if (p.nan_rejects) r2 = 0.0;
                    }
                    accept = u[q] < py_min1(r2);                                        // :558-562
                }
                if (accept) {
#pragma unroll
                    for (int i = 0; i < D; ++i) x[q][i] = y[q][i];
                    lp[q] = lpy[q];
                    acc[q] = stage + 1;
                    want[q] = false;
                    if (p.st.map_lp != nullptr && stage == 0 && live[q] && sub == 0) track_map<D>(p.st, c[q], y[q], lpy[q]);   // :218, Q10
                } else if (DELAYED && stage == 0) {
#pragma unroll
                    for (int i = 0; i < D; ++i) y1[DELAYED ? q : 0][DELAYED ? i : 0] = y[q][i];
                    again = true;                                                       // :535
                } else {
                    n_rej[q] += 1;                                                      // :224, :567
                    want[q] = false;
                }
            }
            if (STREAM && !SPEC) again = __syncthreads_or(again) != 0;     // streaming: the block walks the tiles together
            if (!again) break;
        }
        if constexpr (SPEC) {
            // injected streams: the reference drew for one stage, or for two when the first proposal was rejected
            if (injected) {
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    if (live[q] && sub == 0) advance_injected<D>(p.streams, c[q], acc[q] == 1 ? 1 : 2);
                __syncwarp(group_mask);
            }
        }

        const bool refresh = ADAPT && (--to_refresh == 0);          // it % updating_it == 0  (:470)
        if (refresh) to_refresh = p.updating_it;
        const bool keep = (--to_keep == 0);                          // it % thin == 0
        if (keep) to_keep = (int32_t)p.thin;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int64_t cq = c[q];
            // ---- covariance adaptation (mha.py:432-487).  With G > 1 lanes per chain the group's lanes carry the chain
            // redundantly, but the adaptive state (V_i, X_av, R) lives in global memory and is read back only from there: ONE
            // lane (sub == 0) runs the recursion and stores it, then the group synchronises.  (All lanes doing it would
            // race: a lane that ran late would read values its neighbour has already updated and apply the update twice.)
            if (ADAPT && (G == 1 || sub == 0)) {
                double *const gR = p.st.R + cq;
                double *const gV = p.st.V + cq;
                double *const gxav = p.st.xav + cq;
                const double fi = (double)it;
                double Vc[NT];
                if (!p.am_welford) {
                    // literal Haario/Smith recursion, evaluated in the reference's operation order
                    // without FMA contraction; at it == 1 X_av_i aliases the state.
                    double xa[D], xb[D];
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                        xa[i] = (it == 1) ? x[q][i] : gxav[(int64_t)i * C];
                        xb[i] = __dadd_rn(x[q][i], __dmul_rn(fi / (fi + 1.0), __dadd_rn(xa[i], -x[q][i])));       // :455
                    }
                    const double c1 = (fi - 1.0) / fi, c2 = p.S_d / fi;
#pragma unroll
                    for (int i = 0; i < D; ++i)
#pragma unroll
                        for (int j = i; j < D; ++j) {
                            double t = __dadd_rn(__dmul_rn(fi, __dmul_rn(xa[i], xa[j])), -__dmul_rn(fi + 1.0, __dmul_rn(xb[i], xb[j])));
                            t = __dadd_rn(t, __dmul_rn(x[q][i], x[q][j]));
                            t = __dadd_rn(t, (i == j) ? p.eps_v : 0.0);
                            const int64_t o = (int64_t)tri_idx(D, i, j) * C;
                            const double v = __dadd_rn(__dmul_rn(c1, gV[o]), __dmul_rn(c2, t));             // :463
                            Vc[tri_idx(D, i, j)] = v;
                            if (live[q]) gV[o] = v;
                        }
                    if (live[q]) {
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
                        dl[i] = x[q][i] - m;
                        mu[i] = fma(dl[i], inv_n, m);
                    }
#pragma unroll
                    for (int i = 0; i < D; ++i)
#pragma unroll
                        for (int j = i; j < D; ++j) {
                            const int64_t o = (int64_t)tri_idx(D, i, j) * C;
                            const double v = fma(dl[i], x[q][j] - mu[j], gV[o]);
                            Vc[tri_idx(D, i, j)] = sc * v + ((i == j) ? p.S_d * p.eps_v : 0.0);
                            if (live[q]) gV[o] = v;
                        }
                    if (live[q]) {
#pragma unroll
                        for (int i = 0; i < D; ++i) gxav[(int64_t)i * C] = mu[i];
                    }
                }
                if (refresh) {                                  // :486-487, :591-595
                    double Rn[NT];
                    if (chol_upper_packed<D>(Vc, Rn)) {
                        if (live[q]) {
#pragma unroll
                            for (int i = 0; i < NT; ++i) gR[(int64_t)i * C] = Rn[i];
                        }
                    } else {
                        status[q] |= ST_NOT_PD;                 // scipy raises LinAlgError here (:487)
                    }
                }
            }
            if constexpr (ADAPT && G > 1) __syncwarp(group_mask);
            mean_r[q] += ADAPT ? r[q] : alpha[q];               // :419 (unclipped) / :170

            // ---- outputs
            if (live[q] && sub == 0) {
                if (tracing) {
                    if (p.out.trace_lp1) p.out.trace_lp1[s * C + cq] = lp1[q];
                    if (p.out.trace_lp2) p.out.trace_lp2[s * C + cq] = lp2[q];
                    if (p.out.trace_acc) p.out.trace_acc[s * C + cq] = (uint8_t)acc[q];
                }
                if (keep) {
                    // running (Welford) moments of the KEPT states, held in global memory so that they cost
                    // no registers in the data loop; input of the moment reduction (K4: pooled covariance, R-hat)
                    {
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

    // ---- store state
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        if (live[q] && sub == 0) {
            const int64_t cq = c[q];
#pragma unroll
            for (int i = 0; i < D; ++i) p.st.x[i * C + cq] = x[q][i];
            p.st.lp[cq] = lp[q];
            p.st.mean_r[cq] = mean_r[q];
            p.st.n_rej[cq] += n_rej[q];
            p.st.n_eval[cq] += n_eval[q];
            p.st.status[cq] = status[q];
        }
    }
}

// The kernel proper.  MIXED: the last warps of every block (from warp index p.n_full_warps on) give each chain 2G lanes
// instead of G, i.e. carry half as many chains.  With equal warps the 2048 warps of 65,536 chains (G = 2, Q = 2) cannot
// be spread evenly over the 592 warp schedulers of a B200 (3.46 each: the schedulers holding four set the pace);
// three full warps and one half warp per scheduler can (3.5 units each).
template <class M, int D, int G, int ALGO, int Q, bool STREAM, bool MIXED>
__global__ void __launch_bounds__(StepBounds<M>::MAX_BLOCK, StepBounds<M>::MIN_BLOCKS) mh_step_kernel(const __grid_constant__ MhParams p) {
    TileStream<M::NC> ts;
    M::block_init();
    ts.init(p.prob.records, p.prob.n_points, p.tile_points);
    const int lane = threadIdx.x & 31;
    if constexpr (!MIXED) {
        const int64_t warp_g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        mh_chain_loop<M, D, G, ALGO, Q, STREAM>(p, ts, warp_g * (32 / G) + lane_chain<G>(lane), lane_sub<G>(lane));
    } else {
        const int warp = threadIdx.x >> 5;
        const int64_t base = (int64_t)blockIdx.x * p.groups_per_block;
        if (warp < p.n_full_warps) {
            mh_chain_loop<M, D, G, ALGO, Q, STREAM>(p, ts, base + warp * (32 / G) + lane_chain<G>(lane), lane_sub<G>(lane));
        } else {
            mh_chain_loop<M, D, 2 * G, ALGO, Q, STREAM>(p, ts, base + p.n_full_warps * (32 / G) + (warp - p.n_full_warps) * (16 / G) + lane_chain<2 * G>(lane),
                                                       lane_sub<2 * G>(lane));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Initial state: log-posterior of x0 (mha.py:102).  Also used for plain evaluation of S parameter
// vectors (one thread per vector, G = 1).
// ------------------------------------------------------------------------------------------------
template <class M, int D>
__global__ void __launch_bounds__(256) eval_logpost_kernel(const __grid_constant__ ProblemDev prob, const double *X /*[d][S]*/, int64_t S,
                                                           double *lp_out, int tile_points) {
    TileStream<M::NC> ts;
    M::block_init();
    ts.init(prob.records, prob.n_points, tile_points);
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool livet = t < S;
    if (!livet) t = S - 1;                  // all threads of the block walk the tiles together
    double x[1][D], lp[1];
    const bool want[1] = {livet};
#pragma unroll
    for (int i = 0; i < D; ++i) x[0][i] = X[i * S + t];
    if (TileStream<M::NC>::resident(prob.n_points, tile_points))        // (uniform branch) the resident path walks the model blocks
        log_posterior<M, D, 1, 1, false>(x, want, prob, ts, tile_points, 0, 1u << (threadIdx.x & 31), lp);
    else
        log_posterior<M, D, 1, 1, true>(x, want, prob, ts, tile_points, 0, 1u << (threadIdx.x & 31), lp);
    if (livet) lp_out[t] = lp[0];
}

// Model.run_model for S parameter vectors: f_out[k][s] (point-major, coalesced over samples).
template <class M, int D>
__global__ void __launch_bounds__(256) eval_model_kernel(const __grid_constant__ ProblemDev prob, const double *X /*[d][S]*/, int64_t S,
                                                         double *f_out, int tile_points) {
    TileStream<M::NCR> ts;
    M::block_init();
    ts.init(prob.raw, prob.n_points, tile_points);
    const int n_tl = TileStream<M::NCR>::n_tiles(prob.n_points, tile_points);
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool livet = t < S;
    if (!livet) t = S - 1;
    double x[D];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = X[i * S + t];
    typename M::Ctx ctx;
    typename M::Seq seq;
    int b = 0, next = prob.blk[0].count;          // model blocks are contiguous point ranges: restart at each boundary
    begin_chain_block<M, D>(x, prob, 0, ctx, seq);
    ts.begin_pass(prob.raw, prob.n_points, tile_points);
    const bool one_model = prob.n_blocks <= 1;      // (uniform) the common case pays nothing for the model boundaries
    for (int tl = 0; tl < n_tl; ++tl) {
        int cnt;
        const double *tile = ts.acquire(prob.n_points, tile_points, tl, cnt);
        double *out = f_out + (int64_t)(tl * tile_points) * S + t;
        if (one_model) {
#pragma unroll 2
            for (int k = 0; k < cnt; ++k) {
                double raw[M::NCR];
                load_record<M::NCR>(tile, k, raw);
                const double f = M::value(ctx, seq, raw);
                if (livet) out[(int64_t)k * S] = f;
            }
        } else {
            for (int k = 0; k < cnt; ++k) {
                if (tl * tile_points + k == next) {
                    b += 1;
                    next += prob.blk[b].count;
                    begin_chain_block<M, D>(x, prob, b, ctx, seq);
                }
                double raw[M::NCR];
                load_record<M::NCR>(tile, k, raw);
                const double f = M::value(ctx, seq, raw);
                if (livet) out[(int64_t)k * S] = f;
            }
        }
        ts.release(prob.raw, prob.n_points, tile_points, tl);
    }
}

// The same for a model with conflict-free replicated lookup tables (M = Model::Rep, pbu_math.cuh): one block of 1024
// threads per SM (the replicas take 192 KB of its shared memory), persistent -- the tables are staged once and the block
// strides over the samples.  One model block, records resident (the host checks both before choosing this kernel).
template <class M, int D>
__global__ void __launch_bounds__(1024, 1) eval_model_rep_kernel(const __grid_constant__ ProblemDev prob, const double *X /*[d][S]*/, int64_t S,
                                                                 double *f_out, int tile_points) {
    TileStream<M::NCR> ts;
    M::block_init();
    ts.init(prob.raw, prob.n_points, tile_points);        // (synchronises the block: the tables are in place)
    int cnt;
    const double *tile = ts.acquire(prob.n_points, tile_points, 0, cnt);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < S; t += (int64_t)gridDim.x * blockDim.x) {
        double x[D];
#pragma unroll
        for (int i = 0; i < D; ++i) x[i] = X[i * S + t];
        typename M::Ctx ctx;
        typename M::Seq seq;
        begin_chain_block<M, D>(x, prob, 0, ctx, seq);
        double *out = f_out + t;
#pragma unroll 2
        for (int k = 0; k < cnt; ++k) {
            double raw[M::NCR];
            load_record<M::NCR>(tile, k, raw);
            out[(int64_t)k * S] = M::value(ctx, seq, raw);
        }
    }
}

// nullptr for models without replicated tables
template <class M, int D, class = void>
struct RepEval {
    static const void *fn() { return nullptr; }
};
template <class M, int D>
struct RepEval<M, D, std::void_t<typename M::Rep>> {
    static const void *fn() { return (const void *)&eval_model_rep_kernel<typename M::Rep, D>; }
};

}  // namespace pbu
