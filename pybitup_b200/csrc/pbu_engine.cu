// Host side of the C ABI (include/pybitup_b200.h): engine life cycle, problem lowering to device
// records, kernel selection and launches.  No torch, no exceptions across the ABI.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "pbu_isde_kernel.cuh"
#include "pbu_registry.h"

using namespace pbu;

#include "pbu_engine_internal.h"

static thread_local std::string g_last_error;

int pbu_fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

template <class T>
static int dev_alloc(pbu_engine *e, T **p, size_t n) {
    void *q = nullptr;
    CUDA_TRY(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
    CUDA_TRY(cudaMemsetAsync(q, 0, std::max<size_t>(n, 1) * sizeof(T), e->stream));
    e->allocs.push_back(q);
    *p = static_cast<T *>(q);
    return PBU_OK;
}

// host [C][d] -> host [d][C]
static void transpose_to_soa(const double *src, int64_t C, int d, std::vector<double> &dst) {
    dst.resize((size_t)C * d);
    for (int64_t c = 0; c < C; ++c)
        for (int i = 0; i < d; ++i) dst[(size_t)i * C + c] = src[(size_t)c * d + i];
}
static void transpose_from_soa(const std::vector<double> &src, int64_t C, int d, double *dst) {
    for (int64_t c = 0; c < C; ++c)
        for (int i = 0; i < d; ++i) dst[(size_t)c * d + i] = src[(size_t)i * C + c];
}

static bool chol_upper_host(const double *V, int d, std::vector<double> &Rp) {
    // LAPACK dpotf2 order, as pbu::chol_upper_packed: dot and gemv sums from zero with FMAs, one subtraction, scaled by
    // the reciprocal of the pivot (volatile keeps the host compiler from contracting or reassociating)
    Rp.assign(tri_size(d), 0.0);
    for (int j = 0; j < d; ++j) {
        double dot = 0.0;
        for (int k = 0; k < j; ++k) dot = std::fma(Rp[tri_idx(d, k, j)], Rp[tri_idx(d, k, j)], dot);
        volatile double s = V[j * d + j] - dot;
        if (!(s > 0.0)) return false;
        const double ajj = std::sqrt(s);
        Rp[tri_idx(d, j, j)] = ajj;
        volatile double rinv = 1.0 / ajj;
        for (int i = j + 1; i < d; ++i) {
            double t = 0.0;
            for (int k = 0; k < j; ++k) t = std::fma(Rp[tri_idx(d, k, j)], Rp[tri_idx(d, k, i)], t);
            volatile double diff = V[j * d + i] - t;
            Rp[tri_idx(d, j, i)] = diff * rinv;
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
extern "C" __attribute__((visibility("default"))) int pbu_abi_version(void) { return PBU_ABI_VERSION; }
extern "C" __attribute__((visibility("default"))) const char *pbu_last_error(void) { return g_last_error.c_str(); }

extern "C" __attribute__((visibility("default"))) int pbu_list_kernels(char *buf, int buf_len) {
    std::string s;
    int n = 0;
    for (const MhKernelEntry *t = pbu_mh_kernels(&n); t && s.size() < 1u << 20 && n > 0; --n, ++t) {
        char line[128];
        snprintf(line, sizeof line, "mh model=%d P=%d d=%d lanes=%d chains_per_thread=%d algo=%d streamed=%d mixed=%d coop=%d\n", t->model, t->n_param,
                 t->n_unc, t->lanes, t->chains_per_thread, t->algo, t->streamed, t->mixed, t->coop);
        s += line;
    }
    int total = 0;
    pbu_mh_kernels(&total);
    if (buf && buf_len > 0) {
        strncpy(buf, s.c_str(), buf_len - 1);
        buf[buf_len - 1] = 0;
    }
    return total;
}

static const MhKernelEntry *find_mh(int model, int P, int d, int lanes, int q, int algo, int streamed, int mixed = 0, int coop = 0, int rep = 0, int uniform_w = 0) {
    int n = 0;
    const MhKernelEntry *t = pbu_mh_kernels(&n);
    for (int i = 0; i < n; ++i)
        if (t[i].model == model && t[i].n_param == P && t[i].n_unc == d && t[i].lanes == lanes && t[i].chains_per_thread == q &&
            t[i].algo == algo && t[i].streamed == streamed && t[i].mixed == mixed && t[i].coop == coop && t[i].rep == rep && t[i].uniform_w == uniform_w)
            return &t[i];
    return nullptr;
}
static const IsdeKernelEntry *find_isde(int model, int P, int d, int lanes, int q, int streamed, int mixed = 0) {
    int n = 0;
    const IsdeKernelEntry *t = pbu_isde_kernels(&n);
    for (int i = 0; i < n; ++i)
        if (t[i].model == model && t[i].n_param == P && t[i].n_unc == d && t[i].lanes == lanes && t[i].chains_per_thread == q &&
            t[i].streamed == streamed && t[i].mixed == mixed)
            return &t[i];
    return nullptr;
}
static const IsdeKernelEntry *find_hmc(int model, int P, int d, int lanes, int q, int streamed) {
    int n = 0;
    const IsdeKernelEntry *t = pbu_hmc_kernels(&n);
    for (int i = 0; i < n; ++i)
        if (t[i].model == model && t[i].n_param == P && t[i].n_unc == d && t[i].lanes == lanes && t[i].chains_per_thread == q &&
            t[i].streamed == streamed)
            return &t[i];
    return nullptr;
}
static const EvalKernelEntry *find_eval(int model, int P, int d) {
    int n = 0;
    const EvalKernelEntry *t = pbu_eval_kernels(&n);
    for (int i = 0; i < n; ++i)
        if (t[i].model == model && t[i].n_param == P && t[i].n_unc == d) return &t[i];
    return nullptr;
}

// Lower the problem to the two record streams of pbu_models.cuh.  Several runs are folded exactly:
// sum_r (y_r - f)^2 = n_r (ybar - f)^2 + sum_r (y_r - ybar)^2  (second term -> J_const).
struct HostBlock {          // one model of the likelihood (pbu_problem.n_models), or the whole problem
    int first, count;
    const double *theta_nom;
    const int32_t *unc_idx;
    const double *consts;
};

static void build_records(const pbu_problem *pr, const std::vector<HostBlock> &blocks, int nc, int ncr, std::vector<double> &rec,
                          std::vector<double> &raw, double &J_const) {
    const int N = pr->n_points, nr = pr->n_runs, P = pr->n_param;
    std::vector<int> block_of(N, 0);
    for (size_t b = 0; b < blocks.size(); ++b)
        for (int k = blocks[b].first; k < blocks[b].first + blocks[b].count; ++k) block_of[k] = (int)b;
    rec.assign((size_t)N * nc, 0.0);
    raw.assign((size_t)N * ncr, 0.0);
    J_const = 0.0;
    for (int k = 0; k < N; ++k) {
        double *r = &rec[(size_t)k * nc];
        double *q = &raw[(size_t)k * ncr];
        double w = 1.0 / (pr->std_y[k] * pr->gamma);
        double ybar = pr->y[k];
        if (nr > 1) {
            ybar = 0.0;
            for (int i = 0; i < nr; ++i) ybar += pr->y[(size_t)i * N + k];
            ybar /= (double)nr;
            for (int i = 0; i < nr; ++i) {
                const double t = (pr->y[(size_t)i * N + k] - ybar) * w;
                J_const += t * t;
            }
            w *= std::sqrt((double)nr);
        }
        const double wy = w * ybar;
        switch (pr->model) {
            case PBU_MODEL_LINEAR:
                for (int j = 0; j < P; ++j) {
                    q[j] = pr->design[(size_t)k * P + j];
                    r[j] = w * q[j];
                }
                r[P] = wy;
                break;
            case PBU_MODEL_ARRHENIUS: {          // every model has its own ramp; k counts from the start of the model's points
                const HostBlock &hb = blocks[block_of[k]];
                const double T0 = hb.consts[0], dT = hb.consts[1];
                const double kr = (double)(k - hb.first);
                q[0] = r[0] = 1.0 / (T0 + (kr + 0.5) * dT);
                q[1] = r[1] = 1.0 / (T0 + (kr + 1.0) * dT);
                r[2] = w;
                r[3] = wy;
                break;
            }
            default:   // POLY, HEAT: one design column x
                q[0] = r[0] = pr->design[k];
                r[1] = w;
                r[2] = wy;
                break;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Launch geometry.  One block is resident per SM slot; the chains' G-lane groups are spread so that
// the busiest SM has as little work as possible: block sizes in steps of one warp up to 512 threads
// are tried for every compiled lane count, with the occupancy the runtime reports for that kernel.
// Cost model (arbitrary units): the busiest SM runs ceil(blocks/n_sm) blocks, `occ` at a time; a
// thread costs N/G point evaluations plus a fixed per-iteration part; with fewer than WSAT resident
// warps the FP64 pipe is latency- rather than throughput-bound.
// ------------------------------------------------------------------------------------------------
// Relative FP64 throughput of the data loop against resident warps per SM, measured with tools/ubench
// (cubic Horner loop fed by broadcast shared-memory loads, U = 4 points in flight) for one and two chains
// per thread: two chains per thread reuse every loaded record and every operand-cache entry twice.
static double loop_efficiency(int q, double warps) {
    static const double w1[] = {0, 4, 8, 12, 16, 24, 32}, e1[] = {0.0, 0.19, 0.38, 0.54, 0.67, 0.83, 0.86};
    static const double w2[] = {0, 4, 8, 12, 16, 24, 32}, e2[] = {0.0, 0.74, 0.90, 0.99, 1.0, 1.0, 1.0};
    const double *w = (q >= 2) ? w2 : w1, *e = (q >= 2) ? e2 : e1;
    if (warps >= 32) return e[6];
    for (int i = 1; i < 7; ++i)
        if (warps <= w[i]) return e[i - 1] + (e[i] - e[i - 1]) * (warps - w[i - 1]) / (w[i] - w[i - 1]);
    return e[6];
}

static int plan_launch(pbu_engine *e, int model, bool sequential, int algo, int want_lanes, int want_q, int want_block, int want_coop) {
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, e->device);
    const double N = (double)e->prob.n_points;
    const double S0 = 60.0 + 15.0 * e->d;           // per-iteration scalar part (RNG, proposal, accept, adaptation), in DFMA-equivalents
    // (lanes, chains per thread or per team, mixed, coop): mixed = the last warps of a block use twice the lanes;
    // coop = teams of `lanes` lanes own `lanes` chains (pbu_mh_coop.cuh)
    // rep = the model's lookup tables as conflict-free replicas (192 KB of shared memory, one block of up to 1024 threads per SM)
    // uni = the model's uniform-weight form (records [x | w*y], one shared-memory load per point): cooperative kernels only
    const int cand[15][6] = {{1, 1, 0, 0, 0, 0}, {1, 2, 0, 0, 0, 0}, {2, 2, 0, 0, 0, 0}, {2, 2, 1, 0, 0, 0}, {4, 2, 0, 0, 0, 0}, {2, 1, 0, 0, 0, 0},
                             {4, 1, 0, 0, 0, 0}, {8, 1, 0, 0, 0, 0}, {4, 4, 0, 1, 0, 0}, {4, 4, 1, 1, 0, 0}, {1, 2, 1, 0, 0, 0}, {1, 1, 0, 0, 1, 0},
                             {2, 1, 0, 0, 1, 0}, {4, 4, 0, 1, 0, 1}, {4, 4, 1, 1, 0, 1}};
    const char *force = getenv("PBU_FORCE_VARIANT");      // "lanes,q,mixed,coop[,rep]": diagnostics
    const char *rep_env = getenv("PBU_REP_TABLES");       // 0: never the replicated tables (diagnostics)
    const char *uni_env = getenv("PBU_UNIFORM_W");        // 0: never the uniform-weight kernels (diagnostics)
    const bool hmc = algo == PBU_ALGO_HMC;
    const bool isde = algo == PBU_ALGO_ISDE || hmc;
    const int streamed = e->tile_points < e->prob.n_points ? 1 : 0;
    const double ops = isde ? 2.0 * e->ev->ops : (double)e->ev->ops;     // value + gradient
    double best = 1e300;
    bool found = false;
    for (const auto &gq : cand) {
        const int g = gq[0], q = gq[1], mixed = gq[2], coop = gq[3], rep = gq[4], uni = gq[5];
        if (rep && (isde || streamed || e->prob.n_blocks > 1 || (rep_env && atoi(rep_env) == 0))) continue;
        if (uni && (!e->d_records_u || (uni_env && atoi(uni_env) == 0))) continue;
        if (force) {
            int f[6] = {0, 0, 0, 0, 0, 0};
            sscanf(force, "%d,%d,%d,%d,%d,%d", &f[0], &f[1], &f[2], &f[3], &f[4], &f[5]);
            if (f[0] != g || f[1] != q || f[2] != mixed || f[3] != coop || f[4] != rep || f[5] != uni) continue;
        } else {
            if (want_coop > 0 && !coop) continue;
            if (want_coop < 0 && coop) continue;
            if (want_coop == 0 && coop && (want_lanes > 0 || want_q > 0 || want_block > 0)) continue;
            if (coop && mixed && want_block > 0) continue;
        }
        if (coop && (isde || streamed)) continue;
        if (coop && e->prob.n_blocks > 1) continue;          // the cooperative kernel evaluates one model
        if (want_lanes > 0 && g != want_lanes) continue;
        if (want_q > 0 && q != want_q) continue;
        // sequential (ODE) models: one lane per chain -- or, for delayed rejection, a pair of lanes that evaluates both
        // proposal stages at once (speculative delayed rejection, pbu_mh_kernel.cuh)
        const bool delayed = algo == PBU_ALGO_DR || algo == PBU_ALGO_DRAM;
        if (sequential && !(g == 1 || (g == 2 && q == 1 && delayed && !mixed && !coop))) continue;
        // mixed warp widths: the resident replicated / cooperative MH kernels, and the streamed Ito-SDE pass (not HMC)
        if (mixed && ((isde && (hmc || !streamed)) || (!isde && streamed) || (!force && (want_block > 0 || want_lanes > 0)))) continue;
        // delayed rejection: the second stage runs for a thread when ANY of its chains was rejected, so pairing
        // chains wastes evaluations; one chain per thread unless asked otherwise
        if ((algo == PBU_ALGO_DR || algo == PBU_ALGO_DRAM) && q != 1 && !coop && want_q <= 0 && want_lanes <= 0) continue;
        // wide states spill heavily with two chains per thread
        // (the streamed Ito-SDE / HMC pass with one lane per chain is the exception: its loop waits on the record loads, which
        // two chains per thread halve per FMA -- 9.8 against 10.5 ms at BASELINE configuration 4, spills and all)
        if (e->d > 6 && q != 1 && !(isde && streamed && g == 1) && !(force != nullptr) &&
            ((coop && want_coop <= 0) || (!coop && want_q <= 0 && want_lanes <= 0))) continue;
        const void *fn = nullptr;
        const MhKernelEntry *km = nullptr;
        const IsdeKernelEntry *ki = nullptr;
        if (isde) {
            ki = hmc ? find_hmc(model, e->P, e->d, g, q, streamed) : find_isde(model, e->P, e->d, g, q, streamed, mixed);
            if (ki) fn = ki->fn;
        } else {
            km = find_mh(model, e->P, e->d, g, q, algo, streamed, mixed, coop, rep, uni);
            if (km) fn = km->fn;
        }
        if (!fn) continue;
        int max_smem = 0;
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device);
        cudaFuncAttributes fattr;
        CUDA_TRY(cudaFuncGetAttributes(&fattr, fn));
        if ((int)fattr.sharedSizeBytes > PBU_STATIC_SMEM_RESERVE)
            return fail(PBU_ERR_UNSUPPORTED, "a step kernel uses %zu B of static shared memory, more than the %d B reserved", fattr.sharedSizeBytes, PBU_STATIC_SMEM_RESERVE);
        max_smem -= rep ? (int)fattr.sharedSizeBytes : PBU_STATIC_SMEM_RESERVE;      // (replicas: no tables in static shared memory)
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        const size_t rec_bytes_k = uni ? e->smem_bytes_u : e->smem_bytes;
        const size_t rec_aligned = (rec_bytes_k + 127) & ~(size_t)127;
        // replicated layout: a group of g lanes carries q chains per lane-set; cooperative: a team of g lanes owns q = g chains
        const int64_t groups_needed = (e->C + q - 1) / q;
        // 8 lanes reading 8 consecutive records collide in the shared-memory banks when the record stride is an even
        // number of 16-byte units (two wavefronts per load)
        const double conflict = (g >= 8 && (e->ev->nc / 2) % 2 == 0) ? 1.1 : 1.0;
        // per thread of a warp: every lane evaluates q chains on its share of the points; the scalar part is done for q chains
        // per lane in the replicated layout and for one chain per lane in the cooperative one
        // passes of a sequential model over its points per delayed-rejection step: with one lane per chain a warp walks the
        // points a second time as soon as ONE of its 32 chains was rejected at the first stage (practically always); the
        // speculative pair does one pass for its 16 chains.  Equal work per chain at full occupancy -- measured 5 % in favour
        // of the pair at 262,144 chains (tools/exp_block.py), 1.7x at 32,768 -- so the pair is preferred throughout.
        const double evals = (sequential && delayed) ? ((g == 2) ? 0.95 : 2.0) : 1.0;
        // replicated tables: no bank conflicts among the 32 independent table indices of a warp (PBU_REP_GAIN: diagnostics)
        double rep_gain = 0.9;
        if (const char *rg = getenv("PBU_REP_GAIN")) rep_gain = atof(rg);
        // uniform weight: 8 instead of 14 non-FP64 instructions per four points of the cooperative data loop
        const double work = q * (sequential ? N : std::ceil(N / g)) * ops * conflict * evals * (rep ? rep_gain : 1.0) * (uni ? 0.97 : 1.0) + (coop ? 1.0 : (double)q) * S0;
        for (int B = (rep ? 1024 : MH_MAX_BLOCK); B >= 32; B -= 32) {
            if (want_block > 0 && B != want_block) continue;
            // cooperative kernels park the own chain's state (x, y, lp, mean_r, u: 2d + 3 doubles per thread, component stride
            // MH_MAX_BLOCK) behind the records for the duration of the data loop (pbu_mh_coop.cuh)
            size_t smem_launch = rec_bytes_k;
            int state_off = 0;
            if (rep) {
                smem_launch = rec_aligned + 128 + PBU_REP_TABLE_BYTES;
                if (smem_launch > (size_t)max_smem) continue;
            }
            if (coop) {
                const size_t state = (size_t)(2 * e->d + 3) * sizeof(double) * MH_MAX_BLOCK;
                if (rec_aligned + state > (size_t)max_smem) continue;
                state_off = (int)rec_aligned;
                smem_launch = rec_aligned + state;
            }
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, B, smem_launch) != cudaSuccess || occ < 1) {
                cudaGetLastError();
                continue;
            }
            const int W = B / 32;
            for (int nwide = (mixed ? 4 : 0); nwide <= (mixed ? std::min(W - 4, 8) : 0); nwide += 4) {
                if (mixed && W % 4 != 0) break;
                const int nfull = W - nwide;
                const int gpb = nfull * (32 / g) + nwide * (16 / g);                  // chain groups per block
                const double blocks = std::ceil((double)groups_needed / gpb);
                const double per_sm = std::ceil(blocks / nsm);          // blocks the busiest SM runs, `occ` at a time
                const double conc = std::min(per_sm, (double)occ);
                const double warps = conc * W;
                // warps go round-robin to the 4 schedulers of an SM: the busiest one sets the pace.  A wide warp carries
                // half the points per lane, i.e. half a unit of work.
                const double sched_units = mixed ? (nfull / 4.0 + nwide / 8.0) : std::ceil(conc * W / 4.0) / conc;
                // a handful of warps on the whole GPU is latency-bound: only the serial work of a thread counts
                const bool tiny = blocks * W < nsm;
                const double eff = loop_efficiency(q, warps);
                const double t = tiny ? work : per_sm * sched_units * 4.0 * 32.0 * work / eff;
                if (t < best * (1.0 - 1e-3)) {
                    best = t;
                    e->mh = km;
                    e->isde = ki;
                    e->step_fn = fn;
                    e->lanes = g;
                    e->q = q;
                    e->block = B;
                    e->grid = (int)blocks;
                    e->n_full_warps = mixed ? nfull : 0;
                    e->groups_per_block = gpb;
                    // a warp's own FP64 pipe time per iteration (two cycles per DFMA-equivalent) = a quarter of the iteration of
                    // four warps sharing a scheduler: the start offset that spreads their scalar phases evenly
                    e->stagger_cycles = (coop && W >= 8 && !tiny) ? (int)std::min(2.0 * work, 1.0e6) : 0;
                    if (const char *sv = getenv("PBU_STAGGER")) e->stagger_cycles = coop ? atoi(sv) : 0;      // diagnostics
                    e->state_smem_off = state_off;
                    e->launch_smem = smem_launch;
                    e->uniform_w_kernel = uni != 0;
                    found = true;
                }
            }
        }
    }
    if (!found)
        return fail(PBU_ERR_UNSUPPORTED, "no step kernel variant fits model=%d P=%d d=%d lanes=%d chains_per_thread=%d block=%d algo=%d (smem %zu B)",
                    model, e->P, e->d, want_lanes, want_q, want_block, algo, e->smem_bytes);
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_create(const pbu_problem *pr, const pbu_sampler_cfg *cfg, int64_t n_chains, int device, pbu_engine **out) {
    if (!pr || !cfg || !out) return fail(PBU_ERR_INVALID, "null argument");
    if (n_chains < 1) return fail(PBU_ERR_INVALID, "n_chains must be >= 1");
    if (pr->n_unc < 1 || pr->n_unc > PBU_MAX_DIM || pr->n_param < pr->n_unc || pr->n_param > PBU_MAX_PARAM)
        return fail(PBU_ERR_INVALID, "bad parameter counts P=%d d=%d", pr->n_param, pr->n_unc);
    if (pr->n_points < 1 || pr->n_runs < 1) return fail(PBU_ERR_INVALID, "need n_points >= 1 and n_runs >= 1");
    if (!pr->y || !pr->std_y || !pr->lb || !pr->ub) return fail(PBU_ERR_INVALID, "null problem array");
    std::vector<HostBlock> blocks;
    if (pr->n_models <= 1) {
        if (!pr->theta_nom || !pr->unc_idx) return fail(PBU_ERR_INVALID, "null problem array");
        blocks.push_back({0, pr->n_points, pr->theta_nom, pr->unc_idx, pr->consts});
    } else {
        if (pr->n_models > PBU_MAX_MODELS) return fail(PBU_ERR_UNSUPPORTED, "at most %d models per posterior, got %d", PBU_MAX_MODELS, pr->n_models);
        if (!pr->model_first || !pr->model_count || !pr->model_theta_nom || !pr->model_unc_idx || !pr->model_consts)
            return fail(PBU_ERR_INVALID, "null model-block array");
        int at = 0;
        for (int m = 0; m < pr->n_models; ++m) {
            if (pr->model_first[m] != at || pr->model_count[m] < 1)
                return fail(PBU_ERR_INVALID, "model %d: the models' points must be contiguous, in order and non-empty", m);
            blocks.push_back({at, pr->model_count[m], pr->model_theta_nom + (size_t)m * pr->n_param, pr->model_unc_idx + (size_t)m * pr->n_unc,
                              pr->model_consts + (size_t)m * 8});
            at += pr->model_count[m];
        }
        if (at != pr->n_points) return fail(PBU_ERR_INVALID, "the models' points add up to %d, n_points is %d", at, pr->n_points);
    }
    int nd;
    switch (pr->model) {
        case PBU_MODEL_LINEAR: nd = pr->n_param; break;
        case PBU_MODEL_POLY: nd = 1; break;
        case PBU_MODEL_ARRHENIUS: nd = 2; break;
        case PBU_MODEL_HEAT: nd = 1; break;
        case PBU_MODEL_SOIZE: nd = 1; break;
        default: return fail(PBU_ERR_INVALID, "unknown model id %d", pr->model);
    }
    if (pr->model != PBU_MODEL_ARRHENIUS && (pr->n_design_cols != nd || !pr->design))
        return fail(PBU_ERR_INVALID, "model %d needs %d design column(s), got %d", pr->model, nd, pr->n_design_cols);
    for (size_t b = 0; b < blocks.size(); ++b) {
        const HostBlock &hb = blocks[b];
        if (pr->model == PBU_MODEL_ARRHENIUS && !(hb.consts[0] > 0.0 && hb.consts[1] > 0.0 && hb.consts[2] > 0.0))
            return fail(PBU_ERR_INVALID, "the Arrhenius model needs T0 > 0, dT > 0 (a heating ramp) and beta > 0");
        for (int i = 0; i < pr->n_unc; ++i) {
            // a posterior with several models may have uncertain parameters that a given model does not use (-1)
            const bool absent = blocks.size() > 1 && hb.unc_idx[i] == -1;
            if (!absent && (hb.unc_idx[i] < 0 || hb.unc_idx[i] >= pr->n_param)) return fail(PBU_ERR_INVALID, "unc_idx[%d] out of range", i);
        }
        if (pr->n_unc == pr->n_param)
            // all parameters uncertain: the reference aliases param = var_param (bayesian_inference.py:368-369), i.e. identity order
            for (int i = 0; i < pr->n_unc; ++i)
                if (hb.unc_idx[i] != i) return fail(PBU_ERR_INVALID, "with every parameter uncertain unc_idx must be 0..d-1 (Model.run_model aliases param = var_param)");
    }
    const bool adapt = cfg->algo == PBU_ALGO_AM || cfg->algo == PBU_ALGO_DRAM;
    if (adapt && cfg->updating_it < 1) return fail(PBU_ERR_INVALID, "updating_it must be >= 1");

    const EvalKernelEntry *ev = find_eval(pr->model, pr->n_param, pr->n_unc);
    if (!ev)
        return fail(PBU_ERR_UNSUPPORTED, "no device kernel for model=%d with P=%d parameters, d=%d uncertain (see pbu_list_kernels)",
                    pr->model, pr->n_param, pr->n_unc);

    CUDA_TRY(cudaSetDevice(device));
    pbu_engine *e = new pbu_engine();
    e->device = device;
    e->cfg = *cfg;
    e->C = n_chains;
    e->d = pr->n_unc;
    e->P = pr->n_param;
    e->nt = tri_size(e->d);
    e->ev = ev;

    if (cfg->block_threads != 0 && (cfg->block_threads % 32 != 0 || cfg->block_threads < 32 || cfg->block_threads > 1024)) {
        delete e;
        return fail(PBU_ERR_INVALID, "block_threads must be a multiple of 32 in [32, %d]", 1024);
    }
    if (cfg->algo == PBU_ALGO_HMC && (!(cfg->hmc_step_size > 0.0) || cfg->hmc_num_steps < 1)) {
        delete e;
        return fail(PBU_ERR_INVALID, "HMC needs step_size > 0 and num_steps >= 1");
    }
    if (cfg->algo == PBU_ALGO_ISDE || cfg->algo == PBU_ALGO_HMC) {
        if (cfg->algo == PBU_ALGO_ISDE && (!(cfg->isde_h > 0.0) || !(cfg->isde_f0 >= 0.0))) {
            delete e;
            return fail(PBU_ERR_INVALID, "ISDE needs h > 0 and f0 >= 0");
        }
        if (cfg->isde_gradient == 0) {
            const IsdeKernelEntry *k = find_isde(pr->model, pr->n_param, pr->n_unc, 1, 1, 0);
            if (!k || !k->has_grad) {
                delete e;
                return fail(PBU_ERR_UNSUPPORTED, "model %d has no analytical gradient on the device; use gradient \"Numerical\"", pr->model);
            }
            if (pr->n_runs != 1) {
                delete e;
                return fail(PBU_ERR_UNSUPPORTED, "the analytical gradient uses run 0 only (bayesian_inference.py:606-608); n_runs must be 1");
            }
        }
        if (cfg->adapt_update < 1) {
            delete e;
            return fail(PBU_ERR_INVALID, "covariance.update_it must be >= 1");
        }
    }

    // ---- problem records
    std::vector<double> rec, raw;
    double J_const;
    build_records(pr, blocks, ev->nc, ev->ncr, rec, raw, J_const);
    ProblemDev &pd = e->prob;
    memset(&pd, 0, sizeof pd);
    pd.n_points = pr->n_points;
    pd.n_runs = pr->n_runs;
    pd.ncol = ev->nc;
    pd.ncol_raw = ev->ncr;
    pd.J_const = J_const;
    pd.n_param = pr->n_param;
    pd.n_unc = pr->n_unc;
    for (int i = 0; i < PBU_MAX_DIM; ++i) pd.unc_idx[i] = -1;
    double lpc = 0.0;
    for (int i = 0; i < pr->n_unc; ++i) {
        pd.unc_idx[i] = blocks[0].unc_idx[i];
        pd.lb[i] = pr->lb[i];
        pd.ub[i] = pr->ub[i];
        lpc += std::log(1.0 / std::fabs(pr->ub[i] - pr->lb[i]));      // distributions.py:278, :475
    }
    pd.log_prior_const = pr->has_log_const ? pr->log_const : lpc;
    for (int i = 0; i < pr->n_param; ++i) pd.theta_nom[i] = blocks[0].theta_nom[i];
    for (int i = 0; i < 8; ++i) pd.consts[i] = blocks[0].consts[i];
    pd.par_any = 0;
    for (int i = 0; i < PBU_MAX_DIM; ++i) {
        pd.par_kind[i] = 0;
        pd.par_a[i] = 0.0;
        pd.par_b[i] = 1.0;
        pd.par_c[i] = 0.0;
        pd.par_ldj_s[i] = 0.0;
    }
    pd.par_ldj_c0 = 0.0;
    if (pr->par_kind) {
        if (!pr->par_a || !pr->par_b || !pr->par_c || !pr->par_ldj_s) {
            delete e;
            return fail(PBU_ERR_INVALID, "par_kind without par_a / par_b / par_c / par_ldj_s");
        }
        pd.par_any = 1;
        pd.par_ldj_c0 = pr->par_ldj_c0;
        for (int i = 0; i < pr->n_unc; ++i) {
            if (pr->par_kind[i] != 0 && pr->par_kind[i] != 1) {
                delete e;
                return fail(PBU_ERR_UNSUPPORTED, "parametrisation kind %d of parameter %d is not registered (0 affine, 1 exponential)", pr->par_kind[i], i);
            }
            pd.par_kind[i] = pr->par_kind[i];
            pd.par_a[i] = pr->par_a[i];
            pd.par_b[i] = pr->par_b[i];
            pd.par_c[i] = pr->par_c[i];
            pd.par_ldj_s[i] = pr->par_ldj_s[i];
        }
    }
    pd.n_blocks = (int)blocks.size();
    for (size_t b = 0; b < blocks.size(); ++b) {
        ModelBlockDev &mb = pd.blk[b];
        mb.first = blocks[b].first;
        mb.count = blocks[b].count;
        for (int i = 0; i < PBU_MAX_DIM; ++i) mb.unc_idx[i] = (i < pr->n_unc) ? blocks[b].unc_idx[i] : -1;
        for (int i = 0; i < pr->n_param; ++i) mb.theta_nom[i] = blocks[b].theta_nom[i];
        for (int i = 0; i < 8; ++i) mb.consts[i] = blocks[b].consts[i];
    }

    // resident when every record fits in shared memory, otherwise streamed in tiles through a 2-slot ring
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    max_smem -= PBU_STATIC_SMEM_RESERVE;          // a model's lookup tables live in static shared memory
    auto plan_tiles = [&](size_t rec_bytes, int want_tile, int &tile_points, size_t &smem) {
        const size_t all = (size_t)pr->n_points * rec_bytes;
        if (want_tile <= 0 && 128 + all <= (size_t)max_smem) {
            tile_points = pr->n_points;
            smem = 128 + all;
            return;
        }
        const size_t slot = std::min<size_t>(65536, ((size_t)max_smem - 128) / TileStream<2>::S);
        tile_points = (int)(slot / rec_bytes);
        if (want_tile > 0) tile_points = std::min(tile_points, want_tile);
        tile_points = std::max(1, std::min(tile_points, pr->n_points));
        if (tile_points >= pr->n_points) {
            tile_points = pr->n_points;
            smem = 128 + all;
        } else {
            smem = 128 + (size_t)TileStream<2>::S * tile_points * rec_bytes;
        }
    };
    plan_tiles((size_t)ev->nc * 8, cfg->tile_points, e->tile_points, e->smem_bytes);
    plan_tiles((size_t)ev->ncr * 8, cfg->tile_points, e->tile_points_raw, e->smem_raw_bytes);
    if (blocks.size() > 1 && e->tile_points < pr->n_points) {
        pbu_engine_destroy(e);
        return fail(PBU_ERR_UNSUPPORTED, "several models per posterior need all %d likelihood records resident in shared memory (%zu B); "
                    "streamed records are single-model", pr->n_points, (size_t)pr->n_points * ev->nc * 8);
    }
    int rc;
    if ((rc = dev_alloc(e, &e->d_records, rec.size())) != PBU_OK || (rc = dev_alloc(e, &e->d_raw, raw.size())) != PBU_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(e->d_records, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(e->d_raw, raw.data(), raw.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    pd.records = e->d_records;
    pd.raw = e->d_raw;
    // one weight for all points (and one model, resident records): the uniform-weight form of the records for the kernels of
    // Model::Uniform (pbu_models.cuh: PolyModelU)
    if (pr->model == PBU_MODEL_POLY && blocks.size() == 1 && e->tile_points >= pr->n_points) {
        bool same = true;
        for (int k = 1; k < pr->n_points && same; ++k) same = rec[(size_t)k * ev->nc + 1] == rec[1];
        if (same) {
            std::vector<double> ru((size_t)pr->n_points * 2);
            for (int k = 0; k < pr->n_points; ++k) {
                ru[2 * (size_t)k] = rec[(size_t)k * ev->nc];
                ru[2 * (size_t)k + 1] = rec[(size_t)k * ev->nc + 2];
            }
            if ((rc = dev_alloc(e, &e->d_records_u, ru.size())) != PBU_OK) return rc;
            CUDA_TRY(cudaMemcpyAsync(e->d_records_u, ru.data(), ru.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
            CUDA_TRY(cudaStreamSynchronize(e->stream));          // (ru is a local)
            e->w_uniform = rec[1];
            e->smem_bytes_u = 128 + ru.size() * sizeof(double);
        }
    }

    // ---- launch geometry
    if ((rc = plan_launch(e, pr->model, ev->sequential != 0, cfg->algo, cfg->lanes_per_chain, cfg->chains_per_thread, cfg->block_threads, cfg->cooperative)) != PBU_OK) {
        pbu_engine_destroy(e);
        return rc;
    }

    // ---- chain state
    const int64_t C = n_chains;
    ChainState &st = e->st;
    st.n_chains = C;
    if ((rc = dev_alloc(e, &st.x, (size_t)e->d * C)) || (rc = dev_alloc(e, &st.lp, (size_t)C)) ||
        (rc = dev_alloc(e, &st.R, (size_t)e->nt * C)) || (rc = dev_alloc(e, &st.xav, (size_t)e->d * C)) ||
        (rc = dev_alloc(e, &st.V, (size_t)e->nt * C)) || (rc = dev_alloc(e, &st.mean_r, (size_t)C)) ||
        (rc = dev_alloc(e, &st.n_rej, (size_t)C)) || (rc = dev_alloc(e, &st.n_eval, (size_t)C)) ||
        (rc = dev_alloc(e, &st.status, (size_t)C)) || (rc = dev_alloc(e, &st.wmean, (size_t)e->d * C)) ||
        (rc = dev_alloc(e, &st.wM2, (size_t)e->nt * C)) || (rc = dev_alloc(e, &st.wcount, (size_t)C)) ||
        (rc = dev_alloc(e, &st.P, (size_t)e->d * C)) ||
        (cfg->track_map && ((rc = dev_alloc(e, &st.map_x, (size_t)e->d * C)) || (rc = dev_alloc(e, &st.map_lp, (size_t)C)))) || (rc = dev_alloc(e, &e->d_Cmat, (size_t)((cfg->algo == PBU_ALGO_ISDE || cfg->algo == PBU_ALGO_HMC) ? e->nt : 0) * C)) || (rc = dev_alloc(e, &e->streams.cur_n, (size_t)C)) ||
        (rc = dev_alloc(e, &e->streams.cur_u, (size_t)C)) ||
        (rc = dev_alloc(e, &e->d_mom_partial, (size_t)(2 + 4 * e->d + e->nt) * PBU_MOMENT_SLICES)) ||
        (rc = dev_alloc(e, &e->d_mom_scratch, (size_t)(2 + 4 * e->d + e->nt) + e->d)))
        return rc;
    CUDA_TRY(cudaMallocHost((void **)&e->h_R0_pinned, (size_t)(e->nt + 1) * sizeof(double)));
    CUDA_TRY(cudaEventCreateWithFlags(&e->r0_event, cudaEventDisableTiming));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    *out = e;
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_destroy(pbu_engine *e) {
    if (!e) return PBU_OK;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    for (void *p : e->allocs) cudaFree(p);
    if (e->d_normals) cudaFree(e->d_normals);
    if (e->d_uniforms) cudaFree(e->d_uniforms);
    if (e->h_R0_pinned) cudaFreeHost(e->h_R0_pinned);
    if (e->r0_event) cudaEventDestroy(e->r0_event);
    delete e;
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_set_stream(pbu_engine *e, void *cuda_stream) {
    if (!e) return fail(PBU_ERR_INVALID, "null engine");
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->stream = static_cast<cudaStream_t>(cuda_stream);
    return PBU_OK;
}

static int launch_eval_lp(pbu_engine *e, const double *X_soa_dev, int64_t S, double *lp_dev) {
    CUDA_TRY(cudaFuncSetAttribute(e->ev->eval_lp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
    const int block = 128;
    const unsigned grid = (unsigned)((S + block - 1) / block);
    void *args[] = {(void *)&e->prob, (void *)&X_soa_dev, (void *)&S, (void *)&lp_dev, (void *)&e->tile_points};
    CUDA_TRY(cudaLaunchKernel(e->ev->eval_lp, dim3(grid), dim3(block), args, e->smem_bytes, e->stream));
    return PBU_OK;
}

// inverse of a packed upper-triangular matrix by back substitution (same order as pbu::inv_upper_packed)
static void inv_upper_host(const std::vector<double> &R, int d, std::vector<double> &X) {
    X.assign(tri_size(d), 0.0);
    for (int j = 0; j < d; ++j) {
        X[tri_idx(d, j, j)] = 1.0 / R[tri_idx(d, j, j)];
        for (int i = j - 1; i >= 0; --i) {
            double sacc = 0.0;
            for (int k = i + 1; k <= j; ++k) sacc += R[tri_idx(d, i, k)] * X[tri_idx(d, k, j)];
            X[tri_idx(d, i, j)] = -sacc / R[tri_idx(d, i, i)];
        }
    }
}

// fill rows of a chain-minor [n_rows][C] device array with one value per row (values travel as kernel arguments)
__global__ void fill_i64_kernel(int64_t *dst, int64_t n, int64_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}
// rows [C][d] -> chain-minor [d][C]
__global__ void aos_to_soa_kernel(const double *src, int64_t C, int d, double *dst) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    for (int i = 0; i < d; ++i) dst[(int64_t)i * C + c] = src[c * d + i];
}
struct RowValues {
    double v[tri_size(PBU_MAX_DIM)];
};
__global__ void broadcast_rows_kernel(double *dst, int64_t C, int n_rows, RowValues vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C) return;
    for (int r = 0; r < n_rows; ++r) dst[(int64_t)r * C + i] = vals.v[r];
}
static int broadcast_rows(pbu_engine *e, double *dev, const std::vector<double> &vals) {
    RowValues rv;
    if (vals.size() > sizeof(rv.v) / sizeof(double)) return fail(PBU_ERR_INVALID, "too many rows");
    for (size_t i = 0; i < vals.size(); ++i) rv.v[i] = vals[i];
    broadcast_rows_kernel<<<(unsigned)((e->C + 255) / 256), 256, 0, e->stream>>>(dev, e->C, (int)vals.size(), rv);
    CUDA_TRY(cudaGetLastError());
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_init_chains(pbu_engine *e, const double *x0_host, const double *V0_host) {
    if (!e || !x0_host) return fail(PBU_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(e->device));
    const int d = e->d, nt = e->nt;
    const int64_t C = e->C;
    int rc;
    // the caller's rows [C][d] go to the device as they are (asynchronous when they sit in pinned memory) and are turned
    // chain-minor there; the momentum array, cleared below, is the staging area
    const size_t n_state = (size_t)d * C;
    CUDA_TRY(cudaMemcpyAsync(e->st.P, x0_host, n_state * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    aos_to_soa_kernel<<<(unsigned)((C + 255) / 256), 256, 0, e->stream>>>(e->st.P, C, d, e->st.x);
    CUDA_TRY(cudaGetLastError());
    const bool adapt = e->cfg.algo == PBU_ALGO_AM || e->cfg.algo == PBU_ALGO_DRAM;
    const bool hmc = e->cfg.algo == PBU_ALGO_HMC;
    const bool isde = e->cfg.algo == PBU_ALGO_ISDE || hmc;
    std::vector<double> Vid;
    if (!V0_host) {
        Vid.assign((size_t)d * d, 0.0);
        for (int i = 0; i < d; ++i) Vid[i * d + i] = 1.0;
        V0_host = Vid.data();
    }
    std::vector<double> Rp, Vp((size_t)nt);
    bool pd = chol_upper_host(V0_host, d, Rp);
    if (!pd && !isde)
        return fail(PBU_ERR_NOT_PD, "proposal covariance is not positive definite (scipy.linalg.cholesky would raise LinAlgError)");
    for (int i = 0; i < d; ++i)
        for (int j = i; j < d; ++j) Vp[tri_idx(d, i, j)] = V0_host[i * d + j];
    if (isde) {
        // C_approx, L_c = cholesky(C) (upper), inv_L_c; a failed factorisation falls back to the identity
        // silently, as the reference's bare except does (mha.py:702-711)
        if (!pd) {
            for (int i = 0; i < d; ++i)
                for (int j = i; j < d; ++j) Vp[tri_idx(d, i, j)] = (i == j) ? 1.0 : 0.0;
            Rp = Vp;
        }
        std::vector<double> IL;
        inv_upper_host(Rp, d, IL);
        if ((rc = broadcast_rows(e, e->d_Cmat, Vp)) || (rc = broadcast_rows(e, e->st.R, IL))) return rc;
        CUDA_TRY(cudaMemsetAsync(e->st.V, 0, (size_t)nt * C * sizeof(double), e->stream));
        CUDA_TRY(cudaMemsetAsync(e->st.xav, 0, (size_t)d * C * sizeof(double), e->stream));
    } else {
        e->R0 = Rp;
        if ((rc = broadcast_rows(e, e->st.R, Rp))) return rc;
        if (adapt) {
            // V_i = V_0 (mha.py:404), X_av_i = param_init (:405); Welford mode starts from M2 = 0
            if (e->cfg.am_welford) std::fill(Vp.begin(), Vp.end(), 0.0);
            if ((rc = broadcast_rows(e, e->st.V, Vp))) return rc;
            CUDA_TRY(cudaMemcpyAsync(e->st.xav, e->st.x, n_state * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
        }
    }
    CUDA_TRY(cudaMemsetAsync(e->st.mean_r, 0, C * sizeof(double), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->st.n_rej, 0, C * sizeof(int64_t), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->st.status, 0, C * sizeof(int32_t), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->st.wmean, 0, (size_t)d * C * sizeof(double), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->st.wM2, 0, (size_t)nt * C * sizeof(double), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->st.wcount, 0, C * sizeof(int64_t), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->st.P, 0, (size_t)d * C * sizeof(double), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->streams.cur_n, 0, C * sizeof(int64_t), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->streams.cur_u, 0, C * sizeof(int64_t), e->stream));
    // initial posterior evaluation(s): once, twice for DRAM (double __init__, mha.py:588-589)
    // ... and twice for HMC (MetropolisHastings.__init__ :102 and HamiltonianMonteCarlo.__init__ :814)
    fill_i64_kernel<<<(unsigned)((C + 255) / 256), 256, 0, e->stream>>>(e->st.n_eval, C, (e->cfg.algo == PBU_ALGO_DRAM || hmc) ? 2 : 1);
    CUDA_TRY(cudaGetLastError());
    rc = launch_eval_lp(e, e->st.x, C, e->st.lp);
    if (rc) return rc;
    if (hmc) {      // distr_grad_log_like_current_val = computeGrad(current_val)  (:815), kept in st.P
        CUDA_TRY(cudaFuncSetAttribute(e->ev->eval_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
        const int block = 128;
        const unsigned grid = (unsigned)((C + block - 1) / block);
        int64_t S = C;
        int fd = e->cfg.isde_gradient != 0;
        void *gargs[] = {(void *)&e->prob, (void *)&e->st.x, (void *)&S, (void *)&e->st.P, (void *)&e->tile_points, (void *)&fd};
        CUDA_TRY(cudaLaunchKernel(e->ev->eval_grad, dim3(grid), dim3(block), gargs, e->smem_bytes, e->stream));
    }
    if (e->st.map_lp) {     // arg_MAP = current_val, MAP_val = distr_fun_current_val (mha.py:138-139)
        CUDA_TRY(cudaMemcpyAsync(e->st.map_x, e->st.x, n_state * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
        CUDA_TRY(cudaMemcpyAsync(e->st.map_lp, e->st.lp, C * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->iteration = 0;
    e->initialised = true;
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_set_proposal(pbu_engine *e, const double *V_host) {
    if (!e || !V_host) return fail(PBU_ERR_INVALID, "null argument");
    if (!e->initialised) return fail(PBU_ERR_INVALID, "pbu_engine_init_chains has not been called");
    CUDA_TRY(cudaSetDevice(e->device));
    const int d = e->d, nt = e->nt;
    std::vector<double> Rp, Vp((size_t)nt);
    if (!chol_upper_host(V_host, d, Rp)) return fail(PBU_ERR_NOT_PD, "covariance is not positive definite");
    for (int i = 0; i < d; ++i)
        for (int j = i; j < d; ++j) Vp[tri_idx(d, i, j)] = V_host[i * d + j];
    int rc;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->r0_pending = false;
    if (e->cfg.algo == PBU_ALGO_ISDE || e->cfg.algo == PBU_ALGO_HMC) {
        std::vector<double> IL;
        inv_upper_host(Rp, d, IL);
        if ((rc = broadcast_rows(e, e->d_Cmat, Vp)) || (rc = broadcast_rows(e, e->st.R, IL))) return rc;
    } else {
        e->R0 = Rp;
        if ((rc = broadcast_rows(e, e->st.R, Rp))) return rc;
    }
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_set_streams(pbu_engine *e, const double *normals_host, int64_t n_normals, const double *uniforms_host,
                                      int64_t n_uniforms) {
    if (!e) return fail(PBU_ERR_INVALID, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (e->d_normals) cudaFree(e->d_normals);
    if (e->d_uniforms) cudaFree(e->d_uniforms);
    e->d_normals = e->d_uniforms = nullptr;
    e->streams.normals = e->streams.uniforms = nullptr;
    e->streams.n_normals = e->streams.n_uniforms = 0;
    if (n_normals <= 0 || !normals_host) return PBU_OK;
    const bool need_u = e->cfg.algo != PBU_ALGO_ISDE;      // the Ito-SDE step draws normals only (mha.py:990)
    if (need_u && (!uniforms_host || n_uniforms <= 0)) return fail(PBU_ERR_INVALID, "uniform stream missing");
    if (!uniforms_host) n_uniforms = 0;
    CUDA_TRY(cudaMalloc(&e->d_normals, (size_t)e->C * n_normals * sizeof(double)));
    CUDA_TRY(cudaMalloc(&e->d_uniforms, std::max<size_t>(1, (size_t)e->C * n_uniforms) * sizeof(double)));
    CUDA_TRY(cudaMemcpy(e->d_normals, normals_host, (size_t)e->C * n_normals * sizeof(double), cudaMemcpyHostToDevice));
    if (n_uniforms > 0)
        CUDA_TRY(cudaMemcpy(e->d_uniforms, uniforms_host, (size_t)e->C * n_uniforms * sizeof(double), cudaMemcpyHostToDevice));
    e->streams.normals = e->d_normals;
    e->streams.uniforms = e->d_uniforms;
    e->streams.n_normals = n_normals;
    e->streams.n_uniforms = n_uniforms;
    CUDA_TRY(cudaMemsetAsync(e->streams.cur_n, 0, e->C * sizeof(int64_t), e->stream));
    CUDA_TRY(cudaMemsetAsync(e->streams.cur_u, 0, e->C * sizeof(int64_t), e->stream));
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int64_t pbu_engine_n_keep(const pbu_engine *e, int64_t n_steps, int64_t thin) {
    if (!e || thin < 1 || n_steps < 0) return -1;
    return (e->iteration + n_steps) / thin - e->iteration / thin;
}

extern "C" __attribute__((visibility("default"))) int64_t pbu_engine_iteration(const pbu_engine *e) { return e ? e->iteration : -1; }

extern "C" __attribute__((visibility("default"))) int pbu_engine_get_launch(const pbu_engine *e, int32_t *lanes, int32_t *chains_per_thread, int32_t *block, int32_t *grid,
                                                                                  int64_t *smem_bytes) {
    if (!e) return fail(PBU_ERR_INVALID, "null engine");
    if (lanes) *lanes = e->lanes;
    if (chains_per_thread) *chains_per_thread = e->q;
    if (block) *block = e->block;
    if (grid) *grid = e->grid;
    if (smem_bytes) *smem_bytes = (int64_t)e->launch_smem;
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_get_variant(const pbu_engine *e, int32_t *mixed, int32_t *cooperative, int32_t *replicated_tables,
                                                                                   int32_t *uniform_weight) {
    if (!e) return fail(PBU_ERR_INVALID, "null engine");
    if (mixed) *mixed = e->mh ? e->mh->mixed : (e->isde ? e->isde->mixed : 0);
    if (cooperative) *cooperative = e->mh ? e->mh->coop : 0;
    if (replicated_tables) *replicated_tables = e->mh ? e->mh->rep : 0;
    if (uniform_weight) *uniform_weight = e->uniform_w_kernel ? 1 : 0;
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_run(pbu_engine *e, int64_t n_steps, int64_t thin, const pbu_run_outputs *out) {
    if (!e) return fail(PBU_ERR_INVALID, "null engine");
    if (!e->initialised) return fail(PBU_ERR_INVALID, "pbu_engine_init_chains has not been called");
    if (n_steps < 0 || thin < 1 || n_steps > 2000000000LL) return fail(PBU_ERR_INVALID, "need 0 <= n_steps <= 2e9 per launch and thin >= 1");
    if (n_steps == 0) return PBU_OK;
    CUDA_TRY(cudaSetDevice(e->device));
    if (e->cfg.algo == PBU_ALGO_ISDE || e->cfg.algo == PBU_ALGO_HMC) {
        IsdeParams q;
        memset(&q, 0, sizeof q);
        q.prob = e->prob;
        q.st = e->st;
        q.streams = e->streams;
        if (out) q.out = *out;
        if (q.out.n_sample_chains <= 0 || q.out.n_sample_chains > e->C) q.out.n_sample_chains = e->C;
        q.Cmat = e->d_Cmat;
        q.it0 = e->iteration;
        q.n_steps = n_steps;
        q.thin = thin;
        q.n_groups = (int64_t)e->grid * e->groups_per_block;
        q.seed = e->cfg.seed;
        q.chain_offset = e->cfg.chain_offset;
        q.h = e->cfg.algo == PBU_ALGO_HMC ? e->cfg.hmc_step_size : e->cfg.isde_h;
        q.hmc_steps = e->cfg.hmc_num_steps;
        q.f0 = e->cfg.isde_f0;
        q.S_d = 1.0;                 // mha.py:630
        q.eps_id = 1e-10;            // mha.py:631
        q.starting_it = e->cfg.adapt_start > 0 ? e->cfg.adapt_start : INT64_MAX / 4;
        q.update_it = e->cfg.adapt_update > 0 ? e->cfg.adapt_update : 1;
        q.end_it = e->cfg.adapt_end > 0 ? e->cfg.adapt_end : INT64_MAX / 4;
        q.tile_points = e->tile_points;
        q.fd_gradient = e->cfg.isde_gradient != 0;
        q.n_full_warps = e->n_full_warps;
        q.groups_per_block = e->groups_per_block;
        void *iargs[] = {(void *)&q};
        CUDA_TRY(cudaLaunchKernel(e->step_fn, dim3((unsigned)e->grid), dim3(e->block), iargs, e->launch_smem, e->stream));
        e->iteration += n_steps;
        return PBU_OK;
    }
    if (e->r0_pending) {                 // pooled-covariance mode: the factor computed by pbu_engine_pooled_finalize_dev
        CUDA_TRY(cudaEventSynchronize(e->r0_event));
        e->r0_pending = false;
        if (e->h_R0_pinned[e->nt] != 0.0) e->R0.assign(e->h_R0_pinned, e->h_R0_pinned + e->nt);
    }
    MhParams p;
    memset(&p, 0, sizeof p);
    p.prob = e->prob;
    p.st = e->st;
    p.streams = e->streams;
    if (out) p.out = *out;
    if (p.out.n_sample_chains <= 0 || p.out.n_sample_chains > e->C) p.out.n_sample_chains = e->C;
    p.it0 = e->iteration;
    p.n_steps = n_steps;
    p.thin = thin;
    p.seed = e->cfg.seed;
    p.chain_offset = e->cfg.chain_offset;
    p.eps_v = e->cfg.eps_v;
    p.gamma_dr = e->cfg.gamma_dr;
    p.S_d = 2.38 * 2.38 / (double)e->d;     // mha.py:402
    p.updating_it = e->cfg.updating_it > 0 ? e->cfg.updating_it : 1;
    p.delayed = (e->cfg.algo == PBU_ALGO_DR || e->cfg.algo == PBU_ALGO_DRAM) ? 1 : 0;
    p.am_welford = e->cfg.am_welford;
    p.nan_rejects = e->cfg.nan_rejects;
    p.resident = e->tile_points >= e->prob.n_points;
    p.tile_points = e->tile_points;
    p.n_groups = (int64_t)e->grid * e->groups_per_block;
    p.n_full_warps = e->n_full_warps;
    p.groups_per_block = e->groups_per_block;
    p.stagger_cycles = e->stagger_cycles;
    p.state_smem_off = e->state_smem_off;
    for (size_t i = 0; i < e->R0.size(); ++i) p.R0[i] = e->R0[i];
    if (e->uniform_w_kernel) {          // the kernel of Model::Uniform: records [x | w*y], the weight in consts[7]
        p.prob.records = e->d_records_u;
        p.prob.ncol = 2;
        p.prob.consts[7] = p.prob.blk[0].consts[7] = e->w_uniform;
    }
    void *args[] = {(void *)&p};
    CUDA_TRY(cudaLaunchKernel(e->step_fn, dim3((unsigned)e->grid), dim3(e->block), args, e->launch_smem, e->stream));
    e->iteration += n_steps;
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_synchronize(pbu_engine *e) {
    if (!e) return fail(PBU_ERR_INVALID, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PBU_OK;
}

template <class T>
static int d2h(pbu_engine *e, const T *dev, size_t n, std::vector<T> &host) {
    host.resize(n);
    CUDA_TRY(cudaMemcpyAsync(host.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_get_state(pbu_engine *e, double *x_host, double *lp_host) {
    if (!e) return fail(PBU_ERR_INVALID, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    int rc;
    if (x_host) {
        std::vector<double> soa;
        if ((rc = d2h(e, e->st.x, (size_t)e->d * e->C, soa))) return rc;
        transpose_from_soa(soa, e->C, e->d, x_host);
    }
    if (lp_host) {
        CUDA_TRY(cudaMemcpyAsync(lp_host, e->st.lp, e->C * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    }
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_get_counters(pbu_engine *e, int64_t *n_rejected, double *mean_r, int64_t *n_eval, int32_t *status) {
    if (!e) return fail(PBU_ERR_INVALID, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    if (n_rejected) CUDA_TRY(cudaMemcpyAsync(n_rejected, e->st.n_rej, e->C * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
    if (mean_r) CUDA_TRY(cudaMemcpyAsync(mean_r, e->st.mean_r, e->C * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (n_eval) CUDA_TRY(cudaMemcpyAsync(n_eval, e->st.n_eval, e->C * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
    if (status) CUDA_TRY(cudaMemcpyAsync(status, e->st.status, e->C * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_get_map(pbu_engine *e, double *arg_map_host, double *map_lp_host) {
    if (!e) return fail(PBU_ERR_INVALID, "null engine");
    if (!e->st.map_lp) return fail(PBU_ERR_INVALID, "the engine was created without track_map");
    CUDA_TRY(cudaSetDevice(e->device));
    int rc;
    if (arg_map_host) {
        std::vector<double> soa;
        if ((rc = d2h(e, e->st.map_x, (size_t)e->d * e->C, soa))) return rc;
        transpose_from_soa(soa, e->C, e->d, arg_map_host);
    }
    if (map_lp_host) {
        CUDA_TRY(cudaMemcpyAsync(map_lp_host, e->st.map_lp, e->C * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    }
    return PBU_OK;
}

// State and counters in ONE pass: every array is copied straight into the caller's buffer (asynchronous when that is
// pinned memory), the stream is synchronised once.  x comes back chain-minor [d][C], the layout it has on the device.
extern "C" __attribute__((visibility("default"))) int pbu_engine_read_state(pbu_engine *e, double *x_soa_host, double *lp_host, int64_t *n_rejected,
                                                                                  double *mean_r, int64_t *n_eval, int32_t *status) {
    if (!e) return fail(PBU_ERR_INVALID, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    const size_t C = (size_t)e->C;
    if (x_soa_host) CUDA_TRY(cudaMemcpyAsync(x_soa_host, e->st.x, (size_t)e->d * C * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (lp_host) CUDA_TRY(cudaMemcpyAsync(lp_host, e->st.lp, C * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (n_rejected) CUDA_TRY(cudaMemcpyAsync(n_rejected, e->st.n_rej, C * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
    if (mean_r) CUDA_TRY(cudaMemcpyAsync(mean_r, e->st.mean_r, C * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (n_eval) CUDA_TRY(cudaMemcpyAsync(n_eval, e->st.n_eval, C * sizeof(int64_t), cudaMemcpyDeviceToHost, e->stream));
    if (status) CUDA_TRY(cudaMemcpyAsync(status, e->st.status, C * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_get_proposal(pbu_engine *e, double *R_host, double *V_host, double *xav_host) {
    if (!e) return fail(PBU_ERR_INVALID, "null engine");
    CUDA_TRY(cudaSetDevice(e->device));
    const int d = e->d;
    const int64_t C = e->C;
    int rc;
    std::vector<double> tri;
    auto unpack = [&](double *dst, bool symmetric, double scale, double diag_add) {
        for (int64_t c = 0; c < C; ++c)
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) {
                    double v = 0.0;
                    if (j >= i)
                        v = tri[(size_t)tri_idx(d, i, j) * C + c] * scale + (i == j ? diag_add : 0.0);
                    else if (symmetric)
                        v = tri[(size_t)tri_idx(d, j, i) * C + c] * scale;
                    dst[((size_t)c * d + i) * d + j] = v;
                }
    };
    if (R_host) {
        if ((rc = d2h(e, e->st.R, (size_t)e->nt * C, tri))) return rc;
        unpack(R_host, false, 1.0, 0.0);
    }
    if (V_host) {
        if ((rc = d2h(e, e->st.V, (size_t)e->nt * C, tri))) return rc;
        if (e->cfg.am_welford && e->iteration > 0) {
            const double S_d = 2.38 * 2.38 / (double)d;
            unpack(V_host, true, S_d / (double)e->iteration, S_d * e->cfg.eps_v);
        } else {
            unpack(V_host, true, 1.0, 0.0);
        }
    }
    if (xav_host) {
        std::vector<double> soa;
        if ((rc = d2h(e, e->st.xav, (size_t)d * C, soa))) return rc;
        transpose_from_soa(soa, C, d, xav_host);
    }
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_engine_get_isde_state(pbu_engine *e, double *P_host, double *C_host) {
    if (!e) return fail(PBU_ERR_INVALID, "null engine");
    if (e->cfg.algo != PBU_ALGO_ISDE && e->cfg.algo != PBU_ALGO_HMC) return fail(PBU_ERR_INVALID, "not an ISDE / HMC engine");
    CUDA_TRY(cudaSetDevice(e->device));
    const int d = e->d;
    const int64_t C = e->C;
    int rc;
    if (P_host) {
        std::vector<double> soa;
        if ((rc = d2h(e, e->st.P, (size_t)d * C, soa))) return rc;
        transpose_from_soa(soa, C, d, P_host);
    }
    if (C_host) {
        std::vector<double> tri;
        if ((rc = d2h(e, e->d_Cmat, (size_t)e->nt * C, tri))) return rc;
        for (int64_t c = 0; c < C; ++c)
            for (int i = 0; i < d; ++i)
                for (int j = 0; j < d; ++j) {
                    const int a = std::min(i, j), b = std::max(i, j);
                    C_host[((size_t)c * d + i) * d + j] = tri[(size_t)tri_idx(d, a, b) * C + c];
                }
    }
    return PBU_OK;
}

static int upload_soa(pbu_engine *e, const double *X_host, int64_t S, double **X_dev) {
    std::vector<double> soa;
    transpose_to_soa(X_host, S, e->d, soa);
    CUDA_TRY(cudaMalloc((void **)X_dev, soa.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(*X_dev, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice));
    return PBU_OK;
}

extern "C" __attribute__((visibility("default"))) int pbu_eval_logpost(pbu_engine *e, const double *X_host, int64_t S, double *lp_host) {
    if (!e || !X_host || !lp_host || S < 1) return fail(PBU_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(e->device));
    double *X_dev = nullptr, *lp_dev = nullptr;
    int rc = upload_soa(e, X_host, S, &X_dev);
    if (rc) return rc;
    CUDA_TRY(cudaMalloc((void **)&lp_dev, S * sizeof(double)));
    rc = launch_eval_lp(e, X_dev, S, lp_dev);
    if (!rc) {
        cudaError_t ce = cudaMemcpyAsync(lp_host, lp_dev, S * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        if (ce != cudaSuccess) rc = fail(PBU_ERR_CUDA, "eval_logpost: %s", cudaGetErrorString(ce));
    }
    cudaFree(X_dev);
    cudaFree(lp_dev);
    return rc;
}

extern "C" __attribute__((visibility("default"))) int pbu_eval_model(pbu_engine *e, const double *X_host, int64_t S, double *f_host) {
    if (!e || !X_host || !f_host || S < 1) return fail(PBU_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(e->device));
    const int N = e->prob.n_points;
    double *X_dev = nullptr, *f_dev = nullptr;
    int rc = upload_soa(e, X_host, S, &X_dev);
    if (rc) return rc;
    CUDA_TRY(cudaMalloc((void **)&f_dev, (size_t)S * N * sizeof(double)));
    CUDA_TRY(cudaFuncSetAttribute(e->ev->eval_model, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_raw_bytes));
    const int block = 128;
    const unsigned grid = (unsigned)((S + block - 1) / block);
    void *args[] = {(void *)&e->prob, (void *)&X_dev, (void *)&S, (void *)&f_dev, (void *)&e->tile_points_raw};
    cudaError_t ce = cudaLaunchKernel(e->ev->eval_model, dim3(grid), dim3(block), args, e->smem_raw_bytes, e->stream);
    std::vector<double> pm;
    if (ce == cudaSuccess) {
        pm.resize((size_t)S * N);
        ce = cudaMemcpyAsync(pm.data(), f_dev, pm.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    cudaFree(X_dev);
    cudaFree(f_dev);
    if (ce != cudaSuccess) return fail(PBU_ERR_CUDA, "eval_model: %s", cudaGetErrorString(ce));
    for (int64_t s = 0; s < S; ++s)
        for (int k = 0; k < N; ++k) f_host[(size_t)s * N + k] = pm[(size_t)k * S + s];
    return PBU_OK;
}
