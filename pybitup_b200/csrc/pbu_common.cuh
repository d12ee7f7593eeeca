// Shared device-side definitions of the pybitup_b200 engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/pybitup_b200.h"

namespace pbu {

// ------------------------------------------------------------------------------------------------
// Problem as the kernels see it.  Passed by value as a __grid_constant__ kernel parameter, so the
// small per-problem vectors live in the constant bank and indexing them with unrolled loop
// counters costs nothing.
// ------------------------------------------------------------------------------------------------
// One model of the likelihood: its points are records first .. first + count - 1
struct ModelBlockDev {
    int32_t first, count;
    int32_t unc_idx[PBU_MAX_DIM];     // -1: uncertain parameter i is not a parameter of this model
    double theta_nom[PBU_MAX_PARAM];
    double consts[8];
};

struct ProblemDev {
    const double *records;   // dev [N][NC]  likelihood records, pre-scaled by w = sqrt(n_runs)/(std_y*gamma) (pbu_models.cuh)
    const double *raw;       // dev [N][NCR] raw design records (Model.run_model / propagation)
    int32_t n_points;
    int32_t n_runs;
    int32_t ncol;            // doubles per likelihood record (even) = Model::NC
    int32_t ncol_raw;        // doubles per raw record (even)        = Model::NCR
    double J_const;          // sum_k sum_runs ((y_run,k - ybar_k)/(std_y,k*gamma))^2, 0 for a single run
    int32_t n_param;
    int32_t n_unc;
    int32_t unc_idx[PBU_MAX_DIM];
    double theta_nom[PBU_MAX_PARAM];
    double lb[PBU_MAX_DIM];
    double ub[PBU_MAX_DIM];
    double log_prior_const;  // sum_i log(1/|ub_i-lb_i|)  (Mixture.compute_log_value, distributions.py:456-479)
    double consts[8];
    // Likelihood.models (bayesian_inference.py:500-524): blk[0] repeats the fields above for a single model.  The
    // replicated-layout kernels and the evaluation kernels walk the blocks; the cooperative kernel is single-model.
    int32_t n_blocks;
    int32_t par_any;         // 0: identity parametrisation (Model.parametrization_* of the base class, bayesian_inference.py:314-336)
    ModelBlockDev blk[PBU_MAX_MODELS];
    // Element-wise change of variables between the sampling space Y and the model space X (parametrization_backward,
    // :320-324; the posterior in Y, :44-64): kind 0  X_i = a_i + b_i Y_i;  kind 1  X_i = a_i + b_i exp(c_i Y_i).
    // -log(det_jac(X)) is affine in Y for these (and for a user hook that returns the constant 1): ldj_c0 + sum_i ldj_s_i Y_i.
    int32_t par_kind[PBU_MAX_DIM];
    double par_a[PBU_MAX_DIM], par_b[PBU_MAX_DIM], par_c[PBU_MAX_DIM];
    double par_ldj_c0, par_ldj_s[PBU_MAX_DIM];
};

// Chain state, structure-of-arrays with stride = n_chains (coalesced for one thread per chain).
struct ChainState {
    int64_t n_chains;
    double *x;        // [d][C]      current_val                      (mha.py:99)
    double *lp;       // [C]         distr_fun_current_val            (mha.py:102)
    double *R;        // [d(d+1)/2][C] packed upper Cholesky factor   (mha.py:143)
    double *xav;      // [d][C]      X_av_i                           (mha.py:405)
    double *V;        // [d(d+1)/2][C] packed upper V_i               (mha.py:404)
    double *mean_r;   // [C]                                          (mha.py:154,:170,:419)
    int64_t *n_rej;   // [C]         n_rejected                       (mha.py:146)
    int64_t *n_eval;  // [C]         posterior evaluations
    int32_t *status;  // [C]         bit0: Cholesky failed (not PD), bit1: stream exhausted, bit2: NaN ratio seen
    // Welford moments of the visited states (for pooled covariance / R-hat; K4 input)
    double *wmean;    // [d][C]
    double *wM2;      // [d(d+1)/2][C]
    int64_t *wcount;  // [C]
    // ISDE
    double *P;        // [d][C]      momentum P_nm                    (mha.py:920)
    // MAP tracking (estimate_max_distr, mha.py:138-139, :226-237, :1011-1019); nullptr = off
    double *map_x;    // [d][C]      arg_MAP
    double *map_lp;   // [C]         MAP_val (log)
};

// MetropolisHastings.estimate_max (mha.py:226-237): called on a FIRST-stage acceptance only (the delayed-rejection
// acceptance at :562-565 does not call it); the state stays in global memory, touched on improvement.
template <int D>
__device__ __forceinline__ void track_map(const ChainState &st, int64_t c, const double (&x_new)[D], double lp_new) {
    if (lp_new > st.map_lp[c]) {
        st.map_lp[c] = lp_new;
#pragma unroll
        for (int i = 0; i < D; ++i) st.map_x[(int64_t)i * st.n_chains + c] = x_new[i];
    }
}

struct StreamsDev {
    const double *normals;   // [C][n_normals]  (row per chain)
    const double *uniforms;  // [C][n_uniforms]
    int64_t n_normals;
    int64_t n_uniforms;
    int64_t *cur_n;          // [C] cursors
    int64_t *cur_u;          // [C]
};

struct MhParams {
    ProblemDev prob;
    ChainState st;
    StreamsDev streams;       // normals == nullptr -> Philox
    pbu_run_outputs out;
    int64_t it0;              // iterations already done (global counter, 0 at start)
    int64_t n_steps;
    int64_t thin;
    int64_t n_groups;         // thread groups in the grid; group g owns chains g, g + n_groups, ... (Q per group)
    uint64_t seed;
    uint64_t chain_offset;
    double eps_v;
    double gamma_dr;
    double S_d;               // 2.38^2/d (mha.py:402)
    int32_t updating_it;
    int32_t delayed;          // DR / DRAM
    int32_t am_welford;
    int32_t nan_rejects;
    int32_t resident;         // 1: all records fit in shared memory
    int32_t tile_points;      // streaming mode: points per shared-memory tile
    int32_t n_full_warps;     // mixed kernels: warps of a block with G lanes per chain (the rest use 2G)
    int32_t state_smem_off;   // cooperative adaptive kernels: byte offset of the per-thread (V_i, X_av, R) block in shared
                              // memory, 0 = the adaptive state stays in global memory
    int32_t groups_per_block; // mixed kernels: chain groups per block
    int32_t stagger_cycles;   // > 0: the warps that share a scheduler start this many cycles apart (pbu_mh_coop.cuh), so that the
                              // scalar phase of one overlaps the data loops of the others instead of all four idling the pipe together
    int32_t pad_stagger;
    double R0[PBU_MAX_DIM * (PBU_MAX_DIM + 1) / 2];   // packed upper Cholesky factor shared by all chains (non-adaptive samplers)
};

// X = parametrization_backward(Y) and -log(det_jac(X)) (bayesian_inference.py:50, :61).  `a + b * v` keeps numpy's two
// roundings.  Returns false when nothing is re-parametrised (X = Y, ldj = 0).
template <int D>
__device__ __forceinline__ void par_backward(const ProblemDev &prob, const double (&y)[D], double (&X)[D], double &ldj) {
    ldj = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) X[i] = y[i];
    if (prob.par_any) {
        ldj = prob.par_ldj_c0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double v = (prob.par_kind[i] == 1) ? exp(__dmul_rn(prob.par_c[i], y[i])) : y[i];
            X[i] = __dadd_rn(prob.par_a[i], __dmul_rn(prob.par_b[i], v));
            ldj = fma(prob.par_ldj_s[i], y[i], ldj);
        }
    }
}
// Gradient of the log-posterior with respect to Y from the gradient of the log-likelihood with respect to X
// (Likelihood.compute_grad_log_value with inv_jac = diag(dX/dY), :604-621, minus compute_grad_log_det_jac, :626-636).
template <int D>
__device__ __forceinline__ void par_grad(const ProblemDev &prob, const double (&X)[D], double (&g)[D]) {
    if (prob.par_any) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double dxdy = (prob.par_kind[i] == 1) ? __dmul_rn(prob.par_c[i], __dadd_rn(X[i], -prob.par_a[i])) : prob.par_b[i];
            g[i] = __dadd_rn(__dmul_rn(g[i], dxdy), prob.par_ldj_s[i]);
        }
    }
}

// status bits
enum { ST_NOT_PD = 1, ST_STREAM_EXHAUSTED = 2, ST_NAN_RATIO = 4 };

__host__ __device__ constexpr int tri_size(int d) { return d * (d + 1) / 2; }
// packed row-major upper triangle: (i, j>=i)
__host__ __device__ constexpr int tri_idx(int d, int i, int j) { return i * d - (i * (i - 1)) / 2 + (j - i); }

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter based: counter = (chain, iteration lo, iteration
// hi | stage<<28, block), key = seed.
// ------------------------------------------------------------------------------------------------
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __device__ __forceinline__ static uint4 round10(uint4 c, uint2 k) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
            uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
            c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
            k.x += W0;
            k.y += W1;
        }
        return c;
    }
};

// two standard normals from two 32-bit words: Box-Muller in fp32 (as curand_normal does), widened.  The special
// function unit does the transcendental work (lg2.approx, sin/cos.approx on an argument reduced to [-pi, pi), sqrt
// .approx): absolute error ~2^-21 on the uniform scale, i.e. the proposal draws are float-accurate normals, at a
// quarter of the instructions of logf / sincospif.  (Parity runs inject their draws and never come here.)
// sqrt(-2 ln u) for u in [2^-33, 1]: lg2.approx and sqrt.approx straight on the special function unit (one MUFU each).
// __log2f / __fsqrt_rn wrap them in a denormal rescue and a Newton step with a slow-path branch: 14 instructions more
// per pair, for digits a float-accurate normal does not have.  u is never denormal and the argument of the square root
// is 0 or at least 2^-23, so the .ftz forms lose nothing.
__device__ __forceinline__ float bm_radius(float u) {
    float l, s;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(-1.3862943611198906f * l));
    return s;
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, double &z0, double &z1) {
    const float u = (float)a * 2.3283064365386963e-10f + 1.1641532182693481e-10f;   // (0,1]
    const float ang = (float)(int32_t)b * 1.4629180792671596e-09f;                   // 2 pi * b / 2^32 in [-pi, pi)
    const float s = bm_radius(u);                                                    // sqrt(-2 ln u)
    z0 = (double)(s * __cosf(ang));
    z1 = (double)(s * __sinf(ang));
}

// The same from the top 24 bits of each word (exact int -> float conversions, no rounding that would look at the low byte):
// the low bytes of the four words of one Philox block then make an independent 32-bit uniform (draw_philox for d = 3, 4).
// The radius uniform has 24 bits: the normals are cut off at sqrt(-2 ln 2^-25) = 5.9 sigma (3.6e-9 of the mass).
__device__ __forceinline__ void box_muller24(uint32_t a, uint32_t b, double &z0, double &z1) {
    const float u = (float)(a >> 8) * 5.9604644775390625e-08f + 2.98023223876953125e-08f;    // (k + 1/2) 2^-24 in (0,1)
    const float ang = (float)((int32_t)b >> 8) * 3.7450702829238284e-07f;                     // 2 pi k / 2^24 in [-pi, pi)
    const float s = bm_radius(u);
    z0 = (double)(s * __cosf(ang));
    z1 = (double)(s * __sinf(ang));
}
// uniform in (0,1) with 32 random bits: the low bytes of four words
__device__ __forceinline__ double uniform32_low_bytes(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const uint32_t lo = __byte_perm(a, b, 0x0040), hi = __byte_perm(c, d, 0x0040);
    const uint32_t bits = __byte_perm(lo, hi, 0x5410);
    return fma((double)bits, 2.3283064365386963e-10, 1.1641532182693481e-10);                   // (k + 1/2) 2^-32
}

// uniform in (0,1) with 53 random bits
__device__ __forceinline__ double uniform53(uint32_t a, uint32_t b) {
    uint64_t bits = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11);   // 53 bits
    return ((double)bits + 0.5) * (1.0 / 9007199254740992.0);
}

// ------------------------------------------------------------------------------------------------
// 1-D bulk TMA copy global -> shared with mbarrier completion (cp.async.bulk, SASS UBLKCP).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// bytes must be a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// Small dense helpers on packed upper triangles, fully unrolled (D compile time).
// ------------------------------------------------------------------------------------------------
// Upper Cholesky V = R^T R in the operation order of LAPACK's unblocked dpotf2 ('U'), which is what
// scipy.linalg.cholesky runs at these sizes (mha.py:143, :487, :592): for row j, dot = sum_k R_kj^2 accumulated
// FROM ZERO with fused multiply-adds, ajj = sqrt(V_jj - dot) (one subtraction); then for every i > j the gemv term
// t = sum_k R_kj R_ki from zero with FMAs, and R_ji = (V_ji - t) * (1/ajj): subtract once, scale by the RECIPROCAL
// (dscal).  The intrinsics keep nvcc from re-contracting.  Up to d = 4 this is bit-identical to scipy/OpenBLAS
// (tests/test_oracle_golden.py::test_cholesky_potf2_order_is_scipys); from d = 5 on OpenBLAS's vectorised dot / gemv
// kernels sum in a lane order that depends on its build, so agreement is to rounding.  Returns false if not PD.
template <int D>
__device__ __forceinline__ bool chol_upper_packed(const double (&V)[tri_size(D)], double (&R)[tri_size(D)]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double dot = 0.0;
#pragma unroll
        for (int k = 0; k < j; ++k) dot = __fma_rn(R[tri_idx(D, k, j)], R[tri_idx(D, k, j)], dot);
        const double s = __dsub_rn(V[tri_idx(D, j, j)], dot);
        if (!(s > 0.0)) ok = false;
        const double ajj = __dsqrt_rn(s);
        R[tri_idx(D, j, j)] = ajj;
        const double rinv = __drcp_rn(ajj);
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < j; ++k) t = __fma_rn(R[tri_idx(D, k, j)], R[tri_idx(D, k, i)], t);
            R[tri_idx(D, j, i)] = __dmul_rn(__dsub_rn(V[tri_idx(D, j, i)], t), rinv);
        }
    }
    return ok;
}

// R @ z for the proposal new = cur + R @ z (mha.py:193, :543), R upper triangular: s_i = sum_{j >= i} R_ij z_j.
// np.matmul hands this to BLAS, whose small-size kernels do not sum left to right, and in adaptive Metropolis with a tiny
// eps_v an ulp in a proposal is amplified by cond(V_i) through the next Cholesky factor, so the ORDER decides whether a
// chain stays within 1e-10 of the reference.  Up to d = 4 (every fixture recorded from the reference) the order is the
// one numpy/OpenBLAS uses (tests/test_oracle_golden.py::test_proposal_matvec_order_is_numpys):
//   d = 2:  fma(a0, z0, a1 z1)            d = 3:  fma(a2, z2, fma(a0, z0, a1 z1))
//   d = 4:  (a0 z0 + a2 z2) + (a1 z1 + a3 z3), plain products (a 4-lane vector dot)
// where the zeros below the diagonal drop out exactly.  From d = 5 on the order depends on the BLAS build: an FMA chain.
// rget(i, j) returns R_ij for j >= i.
template <int D, class RGet>
__device__ __forceinline__ void upper_matvec(RGet rget, const double (&z)[D], double (&s)[D]) {
    if constexpr (D == 1) {
        s[0] = __dmul_rn(rget(0, 0), z[0]);
    } else if constexpr (D == 2) {
        s[0] = __fma_rn(rget(0, 0), z[0], __dmul_rn(rget(0, 1), z[1]));
        s[1] = __dmul_rn(rget(1, 1), z[1]);
    } else if constexpr (D == 3) {
        s[0] = __fma_rn(rget(0, 2), z[2], __fma_rn(rget(0, 0), z[0], __dmul_rn(rget(0, 1), z[1])));
        s[1] = __fma_rn(rget(1, 2), z[2], __dmul_rn(rget(1, 1), z[1]));
        s[2] = __dmul_rn(rget(2, 2), z[2]);
    } else if constexpr (D == 4) {
        s[0] = __dadd_rn(__dadd_rn(__dmul_rn(rget(0, 0), z[0]), __dmul_rn(rget(0, 2), z[2])),
                         __dadd_rn(__dmul_rn(rget(0, 1), z[1]), __dmul_rn(rget(0, 3), z[3])));
        s[1] = __dadd_rn(__dmul_rn(rget(1, 2), z[2]), __dadd_rn(__dmul_rn(rget(1, 1), z[1]), __dmul_rn(rget(1, 3), z[3])));
        s[2] = __dadd_rn(__dmul_rn(rget(2, 2), z[2]), __dmul_rn(rget(2, 3), z[3]));
        s[3] = __dmul_rn(rget(3, 3), z[3]);
    } else {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double sj = 0.0;
#pragma unroll
            for (int j = i; j < D; ++j) sj = fma(rget(i, j), z[j], sj);
            s[i] = sj;
        }
    }
}

// Python's min(1, r): r only when r < 1, so NaN -> 1
__device__ __forceinline__ double py_min1(double r) { return (r < 1.0) ? r : 1.0; }

}  // namespace pbu
