/*
 * pybitup_b200 -- C ABI of the B200-native batched-chain MCMC engine.
 *
 * This is the drop-in boundary for the MCMC hot path of jcoheur/pybitup.  The reference is pure
 * Python and has no FFI of its own; the entry points below are what a ctypes binding inside the
 * reference's `pybitup/inference_problem.py` (Sampler.sample, :75-109) and
 * `pybitup/solve_problem.py` (Propagation.propagate, :543-576) would bind -- see INTEGRATION.md.
 * Every function cites the reference interface it replaces (file:line relative to the reference
 * repository root).
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types; all floating point data is float64;
 *   - every function returns PBU_OK (0) or a negative pbu_status; pbu_last_error() gives the
 *     message of the last failure on the calling thread; no exceptions cross the ABI;
 *   - "host" pointers are ordinary host memory (pinned or pageable), "dev" pointers are CUDA
 *     device pointers on the engine's device (e.g. torch.Tensor.data_ptr());
 *   - one engine belongs to one CUDA device and is used from one host thread at a time
 *     (the reference's samplers are single threaded and not re-entrant either);
 *   - chains are independent units: engine g of G owns chains [chain_offset, chain_offset+n_chains)
 *     of the global chain index space, which also keys the Philox streams, so results do not
 *     depend on how chains are sharded over GPUs.
 */
#ifndef PYBITUP_B200_H
#define PYBITUP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBU_ABI_VERSION 1
#define PBU_MAX_PARAM 16 /* max length of the full model parameter vector (P)   */
#define PBU_MAX_DIM 16   /* max number of uncertain parameters (d)              */
#define PBU_MAX_MODELS 4 /* max models (Likelihood.models) of one posterior     */

typedef enum {
    PBU_OK = 0,
    PBU_ERR_INVALID = -1,     /* bad argument                                    */
    PBU_ERR_UNSUPPORTED = -2, /* (model, P, d, ...) combination has no kernel     */
    PBU_ERR_CUDA = -3,        /* CUDA runtime error (message has the detail)     */
    PBU_ERR_NOT_PD = -4,      /* a covariance was not positive definite (scipy LinAlgError in the
                                 reference: metropolis_hastings_algorithms.py:143, :487)          */
    PBU_ERR_STREAM_EXHAUSTED = -5 /* injected random streams too short for the run               */
} pbu_status;

/* Registered device forward models (the reference ships none: documentation/pybitup_doc.tex:280;
 * these replace user `bi.Model.fun_x`, bayesian_inference.py:306-308, :371). */
typedef enum {
    PBU_MODEL_LINEAR = 0,    /* f_k = sum_j Phi[k][j]*theta[j]; design = Phi (N x P)              */
    PBU_MODEL_POLY = 1,      /* Horner polynomial in x, theta = coefficients; design = x (N x 1) */
    PBU_MODEL_ARRHENIUS = 2, /* one-reaction pyrolysis ODE, RK4 over a temperature ramp;
                                theta = (log10 A, E, n, F); consts = {T0, dT, beta}; no design    */
    PBU_MODEL_HEAT = 3,      /* the reference's test model (test/test_pce/heat_conduction.py:27-41);
                                theta = (a,b,k,T_amb,phi,h); design = x; consts = {L}             */
    PBU_MODEL_SOIZE = 4      /* not a forward model: the 2-D test density SoizePDF (distributions.py:483-506)
                                written as two pseudo-residuals; design = selector column (N = 2)  */
} pbu_model_id;

/* Samplers (inference_problem.py:23-33, :79-107). */
typedef enum {
    PBU_ALGO_RWMH = 0, /* MetropolisHastings                         mha.py:26-368   */
    PBU_ALGO_AM = 1,   /* AdaptiveMetropolisHastings                 mha.py:370-487  */
    PBU_ALGO_DR = 2,   /* DelayedRejectionMetropolisHastings         mha.py:490-567  */
    PBU_ALGO_DRAM = 3, /* DelayedRejectionAdaptiveMetropolisHastings mha.py:570-595  */
    PBU_ALGO_ISDE = 4, /* ito_SDE                                    mha.py:886-1040 */
    PBU_ALGO_HMC = 5   /* HamiltonianMonteCarlo                      mha.py:779-884  */
} pbu_algo_id;

/* One Bayesian inverse problem: what Sampling.sample builds at solve_problem.py:184-337
 * (Data :271, Model parameters :200-202, Mixture-of-Uniform prior :322-324, Likelihood :332). */
typedef struct {
    int32_t model;            /* pbu_model_id                                                     */
    int32_t n_param;          /* P: length of theta_nom (Model.param_nom)                         */
    int32_t n_unc;            /* d: number of uncertain parameters                                */
    int32_t n_points;         /* N: data points                                                   */
    int32_t n_runs;           /* Data.n_runs (bayesian_inference.py:160)                          */
    int32_t n_design_cols;    /* columns of `design` (P for LINEAR, 1 for POLY/HEAT, 0 ARRHENIUS) */
    const double *theta_nom;  /* host [P]                                                         */
    const int32_t *unc_idx;   /* host [d]: position of uncertain parameter i inside theta         */
    const double *design;     /* host [N][n_design_cols] row major, may be NULL when 0 columns    */
    const double *y;          /* host [n_runs][N]    Data.y[run]                                   */
    const double *std_y;      /* host [N]            Data.std_y                                    */
    const double *lb;         /* host [d] Uniform.lb (distributions.py:276)                       */
    const double *ub;         /* host [d] Uniform.ub (distributions.py:277)                       */
    double gamma;             /* Likelihood.gamma (bayesian_inference.py:407)                     */
    double consts[8];         /* model constants, see pbu_model_id                                */
    double log_const;         /* used when has_log_const != 0: constant added to -J/2 INSTEAD of the log-prior
                                 constant (sampling an analytic density, solve_problem.py:134-164)  */
    int32_t has_log_const;
    /* Several models in one likelihood (Likelihood.models / .data are dicts keyed by model id, summed at
     * bayesian_inference.py:500-524).  n_models <= 1: one model, the fields above are all there is.  n_models > 1:
     * the points of the models are CONCATENATED in design / y / std_y (model m owns points model_first[m] ..
     * model_first[m] + model_count[m] - 1), all models are the same registered device model with the same P, and
     * each has its own nominal parameters, name map and constants; theta_nom / unc_idx / consts above are ignored. */
    int32_t n_models;
    const int32_t *model_first;     /* host [n_models]                                                          */
    const int32_t *model_count;     /* host [n_models]                                                          */
    const double *model_theta_nom;  /* host [n_models][P]                                                       */
    const int32_t *model_unc_idx;   /* host [n_models][d]: position inside that model's theta, -1 = not a
                                       parameter of that model (Model.run_model's name search finds nothing)    */
    const double *model_consts;     /* host [n_models][8]                                                       */
    /* Element-wise change of variables between the sampling space Y and the model space X, i.e. the user's
     * Model.parametrization_backward / _det_jac / _inv_jac hooks (bayesian_inference.py:314-336) lowered to the device:
     * par_kind NULL = identity.  kind 0: X_i = a_i + b_i Y_i;  kind 1: X_i = a_i + b_i exp(c_i Y_i).  The posterior in Y is
     * prior(X) - log(det_jac(X)) + loglike(X) (:44-64) with -log(det_jac(X)) = par_ldj_c0 + sum_i par_ldj_s[i] Y_i; its
     * gradient is diag(dX/dY) grad_X loglike + par_ldj_s (:84-89, :604-636).  The prior box [lb, ub] lives in X. */
    const int32_t *par_kind;        /* host [d] or NULL                                                         */
    const double *par_a;            /* host [d]                                                                 */
    const double *par_b;            /* host [d]                                                                 */
    const double *par_c;            /* host [d]                                                                 */
    const double *par_ldj_s;        /* host [d]                                                                 */
    double par_ldj_c0;
} pbu_problem;

/* Algorithm block: inference_problem.py:37-107 (JSON keys in SURVEY.md 9.2). */
typedef struct {
    int32_t algo;             /* pbu_algo_id                                                      */
    int32_t updating_it;      /* AMH/DRAM.updating_it          (mha.py:470)                       */
    double eps_v;             /* AMH/DRAM.eps_v                (mha.py:403)                       */
    double gamma_dr;          /* DR/DRAM.gamma                 (mha.py:514, :543)                 */
    double isde_h;            /* ISDE.h                        (mha.py:926)                       */
    double isde_f0;           /* ISDE.f0                       (mha.py:927)                       */
    int32_t isde_gradient;    /* 0 Analytical (mha.py:727-729), 1 Numerical forward FD (:724-726)  */
    int32_t adapt_start;      /* covariance.starting_it        (mha.py:632)  (ISDE)               */
    int32_t adapt_update;     /* covariance.update_it          (mha.py:633)  (ISDE)               */
    int32_t adapt_end;        /* covariance.end_it, <=0: n_iterations+1 (mha.py:635-638) (ISDE)   */
    int32_t am_welford;       /* 0: literal Haario recursion of the reference (mha.py:455,:463);
                                 1: Welford accumulation (cancellation free, same estimator)      */
    int32_t nan_rejects;      /* 0: reference semantics, min(1,nan)==1 accepts;
                                 1: a NaN acceptance ratio rejects                                */
    int32_t lanes_per_chain;  /* 0 auto; 1,2,4,8: lanes of a warp that share one chain's data sum */
    int32_t block_threads;    /* 0 auto                                                           */
    int32_t chains_per_thread;/* 0 auto; 1,2: chains that share every record a thread reads       */
    int32_t tile_points;      /* 0 auto: records resident in shared memory when they fit, else streamed in
                                 tiles; > 0 forces tiles of at most that many points (testing)            */
    uint64_t seed;            /* Philox key                                                       */
    uint64_t chain_offset;    /* global index of this engine's first chain (multi-GPU sharding)   */
    double hmc_step_size;     /* HMC.step_size (epsilon)       (mha.py:818)                       */
    int32_t hmc_num_steps;    /* HMC.num_steps (L)             (mha.py:819)                       */
    int32_t cooperative;      /* 0 auto; 1: cooperative teams (lanes_per_chain lanes own as many chains, each
                                 lane does the scalar part of one); -1: replicated layout only          */
    int32_t track_map;        /* 1: per-chain MAP estimate (opt_arg estimate_max_distr, mha.py:77, :138-139,
                                 :226-237; Ito-SDE :1011-1019 with its extra evaluation per iteration)   */
    int32_t reserved0;
} pbu_sampler_cfg;

/* Per-run outputs that live on the device; any pointer may be NULL (= not recorded).
 * Layouts are chain-minor so that one thread per chain writes coalesced. */
typedef struct {
    double *samples;     /* dev [n_keep][d][Cs]  state after every `thin`-th iteration (chain rows,
                            bayesian_inference.py:131-138) of the first Cs = n_sample_chains chains */
    double *lp_kept;     /* dev [n_keep][C]      log-posterior of the kept states                 */
    double *trace_lp1;   /* dev [n_steps][C]     log-posterior of the stage-1 proposal (parity)   */
    double *trace_lp2;   /* dev [n_steps][C]     ... of the stage-2 proposal, NaN if none         */
    uint8_t *trace_acc;  /* dev [n_steps][C]     0 rejected, 1 accepted stage 1, 2 stage 2        */
    double *aux;         /* dev [n_keep][d][Cs]  ISDE momentum P after the kept iterations
                            (aux_variables.csv, mha.py:1032)                                      */
    int64_t n_sample_chains; /* Cs: leading chains recorded in samples / aux; <= 0 means all C    */
} pbu_run_outputs;

typedef struct pbu_engine pbu_engine;

/* ---- library ------------------------------------------------------------------------------- */
int pbu_abi_version(void);
const char *pbu_last_error(void);
/* number of (model,P,d) kernel variants compiled in; fills `buf` with a human readable list */
int pbu_list_kernels(char *buf, int buf_len);

/* ---- engine life cycle: replaces Sampler.__init__ + sampler class __init__
 *      (inference_problem.py:15-73, metropolis_hastings_algorithms.py:74-155) ------------------- */
int pbu_engine_create(const pbu_problem *prob, const pbu_sampler_cfg *cfg, int64_t n_chains,
                      int device, pbu_engine **out);
int pbu_engine_destroy(pbu_engine *e);
/* CUDA stream all later work is issued on (a cudaStream_t, e.g. torch's current stream); NULL =
 * the legacy default stream. */
int pbu_engine_set_stream(pbu_engine *e, void *cuda_stream);

/* Initial state: x0 host [C][d] (MetropolisHastings.__init__: current_val = param_init, :95-99),
 * proposal covariance V0 host [d][d] shared by all chains (:80, :142-143; R = upper Cholesky).
 * Evaluates the initial log-posterior on the device (:102). For ISDE / HMC V0 is the preconditioner
 * C_approx (:653-711) and may be NULL for identity; HMC also evaluates the initial gradient (:815). */
int pbu_engine_init_chains(pbu_engine *e, const double *x0_host, const double *V0_host);

/* Parity mode: replace Philox by pre-generated streams (what `mha.random = obj` does to the
 * reference). normals host [C][n_normals], uniforms host [C][n_uniforms]; each chain
 * consumes its own rows sequentially (d normals then one uniform per stage). n == 0 restores Philox. */
int pbu_engine_set_streams(pbu_engine *e, const double *normals_host, int64_t n_normals,
                           const double *uniforms_host, int64_t n_uniforms);

/* ---- the hot loop: replaces run_algorithm (mha.py:158-182, :407-430, :932-1040) -------------- */
/* Advance every chain by n_steps iterations in ONE kernel launch (asynchronous on the engine's
 * stream). thin >= 1. Iteration numbering continues across calls. */
int pbu_engine_run(pbu_engine *e, int64_t n_steps, int64_t thin, const pbu_run_outputs *out);
/* number of rows `samples` needs for a run of n_steps starting at the engine's current iteration */
int64_t pbu_engine_n_keep(const pbu_engine *e, int64_t n_steps, int64_t thin);
int pbu_engine_synchronize(pbu_engine *e);

/* ---- state access (device -> host copies; synchronise the engine's stream) ------------------- */
int pbu_engine_get_state(pbu_engine *e, double *x_host /*[C][d]*/, double *lp_host /*[C]*/);
int pbu_engine_get_counters(pbu_engine *e, int64_t *n_rejected /*[C]*/, double *mean_r /*[C]*/,
                            int64_t *n_eval /*[C]*/, int32_t *status /*[C]*/);
/* MAP estimate of every chain (arg_MAP, MAP_val in log: mha.py:138-139, :232-234; written out at :352-355);
 * needs cfg.track_map = 1 */
int pbu_engine_get_map(pbu_engine *e, double *arg_map_host /*[C][d]*/, double *map_lp_host /*[C]*/);
/* the same in one pass, for callers with pinned staging buffers: every array is copied straight into the caller's buffer
 * and the stream is synchronised once; x comes back chain-minor [d][C] (its device layout).  Any pointer may be NULL.
 * (current_val, distr_fun_current_val, n_rejected, mean_r: mha.py:99, :102, :146, :154) */
int pbu_engine_read_state(pbu_engine *e, double *x_soa_host /*[d][C]*/, double *lp_host /*[C]*/, int64_t *n_rejected /*[C]*/,
                          double *mean_r /*[C]*/, int64_t *n_eval /*[C]*/, int32_t *status /*[C]*/);
/* per-chain proposal factor R (upper, V = R^T R; mha.py:143,:487), adaptive covariance V_i
 * (:463-466) and running mean X_av_i (:455,:467); any pointer may be NULL. Row-major [C][d][d]. */
int pbu_engine_get_proposal(pbu_engine *e, double *R_host, double *V_host, double *xav_host);
int64_t pbu_engine_iteration(const pbu_engine *e);
/* Ito-SDE state: momentum P_nm (mha.py:920) host [C][d] and preconditioner C_approx (:653-697, :773) host
 * [C][d][d]; pbu_engine_get_proposal returns inv_L_c in R_host, V_i and X_av_i as for the adaptive samplers. */
int pbu_engine_get_isde_state(pbu_engine *e, double *P_host, double *C_host);
/* launch geometry the engine planned for the step kernel: lanes of a warp per chain, threads per block,
 * chains per thread, blocks, dynamic shared memory bytes per block (diagnostics; no reference counterpart). */
int pbu_engine_get_launch(const pbu_engine *e, int32_t *lanes, int32_t *chains_per_thread, int32_t *block, int32_t *grid,
                          int64_t *smem_bytes);
/* which compiled variant of the step kernel that is: mixed warp widths, cooperative teams, replicated lookup tables,
 * uniform-weight records (0 / 1 each; diagnostics, no reference counterpart). */
int pbu_engine_get_variant(const pbu_engine *e, int32_t *mixed, int32_t *cooperative, int32_t *replicated_tables, int32_t *uniform_weight);

/* ---- evaluation only: BayesianPosterior.compute_log_value (bayesian_inference.py:44-64) and
 *      Model.run_model (:338-371) for S parameter vectors -------------------------------------- */
int pbu_eval_logpost(pbu_engine *e, const double *X_host /*[S][d]*/, int64_t S, double *lp_host /*[S]*/);
int pbu_eval_model(pbu_engine *e, const double *X_host /*[S][d]*/, int64_t S, double *f_host /*[S][N]*/);

/* ---- chain moments (K4): sums over this engine's chains of the running moments of the kept states; additive over
 *      chains and over GPUs.  Many-chain counterpart of compute_covariance (mha.py:295-307, np.mean / np.cov over
 *      the chain file); the host layer all-reduces them for the pooled covariance and the Gelman-Rubin R-hat.
 *      out_host [2 + 4d + d(d+1)/2] as laid out in pbu_moments.cu; mu_host [d] reference point (NULL = 0). */
int pbu_engine_chain_sums_len(const pbu_engine *e);
int pbu_engine_chain_sums(pbu_engine *e, const double *mu_host, double *out_host);
int pbu_engine_reset_moments(pbu_engine *e);
/* Device-resident form (what the samplers use; one process per GPU).  Everything is queued on the engine's stream, nothing is
 * allocated and nothing touches the host, so the caller can all-reduce the small vectors in place with NCCL between the
 * calls: chain_sums_dev(mu = NULL) -> all-reduce -> pooled_mean_dev -> chain_sums_dev(mu) -> all-reduce -> pooled_finalize_dev.
 *   chain_sums_dev     out_dev [pbu_engine_chain_sums_len] = this engine's sums about mu_dev [d] (NULL = 0); n_components > 0
 *                      computes the leading n_components entries only (2 + d locate the mean: all the first pass needs).
 *   pooled_mean_dev    mu_out_dev [d] = mu_in_dev (NULL = 0) + sum_c n_c delta_c / sum_c n_c from the all-reduced sums.
 *   pooled_finalize_dev  stats_dev [pbu_engine_pooled_stats_len] from the all-reduced sums about mu_dev:
 *        [0] chains [1] kept states [2] ok (proposal matrix positive definite) [3] reserved, then mean d | covariance d*d
 *        (np.cov of all kept states, mha.py:303-304) | W d | B/n d | Gelman-Rubin R-hat d | V packed | R packed | inv(R)
 *        packed, with V = scale (cov + eps I): the covariance refresh of mha.py:463 / :745-746 fed by ALL chains instead of
 *        one chain's recursion (scale = 2.38^2/d, eps = eps_v for the MH family; 1 and 1e-10 for the Ito-SDE
 *        preconditioner, :630-631).  apply != 0 writes the result into every chain: R = upper Cholesky of V for the MH
 *        family (:486-487), C_approx / inv_L_c for ISDE and HMC (:773-777) -- unless V is not positive definite, in which
 *        case the proposal stays as it was and ok = 0. */
int pbu_engine_pooled_stats_len(const pbu_engine *e);
int pbu_engine_chain_sums_dev(pbu_engine *e, const double *mu_dev, double *out_dev, int n_components);
int pbu_engine_pooled_mean_dev(pbu_engine *e, const double *sums_dev, const double *mu_in_dev, double *mu_out_dev);
int pbu_engine_pooled_finalize_dev(pbu_engine *e, const double *sums_dev, const double *mu_dev, double scale, double eps, int apply,
                                   double *stats_dev);
/* Replace the proposal covariance of EVERY chain (pooled-covariance mode): R = upper Cholesky of V host [d][d]
 * for the MH family (mha.py:486-487), C_approx / L_c / inv_L_c for ISDE (:773-777). */
int pbu_engine_set_proposal(pbu_engine *e, const double *V_host);

/* ---- Monte-Carlo propagation: replaces the sample loop of Propagation.propagate, emulator == "None"
 *      (solve_problem.py:543-576) and its statistics (:596-647).  The evaluations stay in device memory,
 *      point-major; mean / min / max and EXACT order statistics (for np.percentile, :643-645) are reduced on the
 *      device.  With samples sharded over GPUs the caller sums the moment and histogram outputs over ranks
 *      (one small all-reduce per call), which replaces the reference's MPI send of every evaluation (:578-592). */
typedef struct pbu_propagation pbu_propagation;
int pbu_propagation_create(pbu_engine *e, int64_t n_samples, pbu_propagation **out);
int pbu_propagation_destroy(pbu_propagation *p);
/* evaluate run_model for X host [S][d] (solve_problem.py:559-571) */
int pbu_propagation_eval_host(pbu_propagation *p, const double *X_host);
/* same with rows drawn on the device from N(mean, diag(std^2)); Philox keyed by (seed, sample_offset + row) */
int pbu_propagation_eval_gaussian(pbu_propagation *p, const double *mean_host /*[d]*/, const double *std_host /*[d]*/,
                                  uint64_t seed, uint64_t sample_offset);
/* device pointers: samples [d][S] and evaluations [N][S] */
int pbu_propagation_buffers(pbu_propagation *p, double **X_dev, double **f_dev);
/* per design point sum, min, max over this engine's samples, host [N] each (:607-627) */
int pbu_propagation_moments(pbu_propagation *p, double *sum_host, double *min_host, double *max_host);
/* exact order statistics by radix select.  ranks host [n_ranks] (<= 4): 0-based ranks in the sorted GLOBAL sample.
 * row_min / row_max host [N] (both or neither, may be NULL): the GLOBAL minimum / maximum of every design point (the
 * moments above, reduced over GPUs); the leading key bytes that all values share are then not scanned, and *first_pass
 * (may be NULL) receives the first pass to run (0 without bounds, 8 when every design point is constant).
 * For pass = first_pass..7: select_pass -> hist_host [N][n_ranks][256] (this engine's counts); sum over GPUs;
 * select_update.  On a single GPU both histogram pointers may be NULL: the counts then never leave the device and the
 * passes are queued back to back.  (np.percentile of fun_eval, solve_problem.py:639-647) */
int pbu_propagation_select_begin(pbu_propagation *p, const int64_t *ranks, int n_ranks, const double *row_min, const double *row_max,
                                 int *first_pass);
int pbu_propagation_select_pass(pbu_propagation *p, int pass, uint64_t *hist_host);
int pbu_propagation_select_update(pbu_propagation *p, int pass, const uint64_t *hist_total_host);
/* the same two steps with the histogram in a caller-owned DEVICE buffer [N][n_ranks][256] (one process per GPU: the
 * caller all-reduces it in place with NCCL on the engine's stream between the two calls; no host copy, no sync) */
int pbu_propagation_select_pass_dev(pbu_propagation *p, int pass, uint64_t *hist_dev);
int pbu_propagation_select_update_dev(pbu_propagation *p, int pass, const uint64_t *hist_total_dev);
int pbu_propagation_select_result(pbu_propagation *p, double *values_host /*[N][n_ranks]*/);
/* Bracketed select, the fast path for large S (same np.percentile, solve_problem.py:639-647, ONE pass over the evaluations
 * instead of up to eight).  n_q = 1 or 2 percentiles.
 *   1. bracket_sample: radix-selects a subsample (every stride-th chunk of 256 values of each design point) at ranks six
 *      binomial standard deviations either side of fractions[q] (in [0, 1]) -> brackets_host [N][n_q][2] = (a, b), -inf / +inf
 *      where the margin reaches the ends.  Over GPUs: min-reduce a, max-reduce b.
 *   2. bracket_pass: one pass with the global brackets: sum / min / max as pbu_propagation_moments (this replaces that
 *      call), below / inside host [N][n_q] = this engine's counts of values < a and in [a, b]; the values inside go to a
 *      candidate buffer of at least 4 cap_hint + 4096 slots per (point, percentile) (pbu_propagation_create takes it,
 *      sized for any percentile, from 2^18 samples on), *overflow != 0 if that was too small.  Where a == b (the
 *      subsample has a tie across the bracket, e.g. a constant design point) nothing is copied and nothing can overflow:
 *      if the rank test below holds the order statistic IS a, whatever select_result reports for that entry.  Over
 *      GPUs: sum the counts, max-reduce the flag.  The wanted global ranks r must satisfy below <= r < below + inside for
 *      every point, else (or on overflow) fall back to pbu_propagation_select_begin.
 *   3. bracket_select_begin with ranks_in_bracket host [N][2 n_q] (= r - below) and the global brackets, then select_pass /
 *      select_update / select_result exactly as above, over the candidates only. */
int pbu_propagation_bracket_sample(pbu_propagation *p, const double *fractions, int n_q, int stride, double *brackets_host);
int pbu_propagation_bracket_pass(pbu_propagation *p, const double *brackets_host, int n_q, int64_t cap_hint, double *sum_host, double *min_host,
                                 double *max_host, int64_t *below_host, int64_t *inside_host, int *overflow);
int pbu_propagation_bracket_select_begin(pbu_propagation *p, const int64_t *ranks_in_bracket, const double *brackets_host, int *first_pass);
/* evaluations of samples [s0, s0+n) on the host, sample-major [n][N] (fun_eval, :571) */
int pbu_propagation_get_evals(pbu_propagation *p, int64_t s0, int64_t n, double *f_host);
/* the samples [s0, s0+n) themselves, [n][d] on the host: the rows of eval_host, or those eval_gaussian drew (c_param, :559-565) */
int pbu_propagation_get_samples(pbu_propagation *p, int64_t s0, int64_t n, double *X_host);

/* ---- measurement helper: achieved FP64 FMA throughput of the device (TFLOP/s, FMA = 2 flops) from a
 *      register-resident DFMA loop; the denominator of the FP64 roofline in bench.py.  reps > 0: best of `reps`
 *      launches (burst); reps < 0: mean of -reps launches back to back (sustained) ---------------- */
int pbu_measure_fp64_peak(int device, int iters, int reps, double *tflops, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* PYBITUP_B200_H */
