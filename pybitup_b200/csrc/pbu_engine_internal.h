// Internals shared by the host-side translation units of the C ABI (not installed).
#pragma once
#include <string>
#include <vector>

#include "pbu_isde_kernel.cuh"
#include "pbu_registry.h"

// records the message returned by pbu_last_error() on this thread and returns `code`
int pbu_fail(int code, const char *fmt, ...);
#define fail pbu_fail

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return pbu_fail(PBU_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define PBU_API extern "C" __attribute__((visibility("default")))

// static shared memory a kernel may use besides the dynamic record tiles (the Arrhenius model's exp / log lookup
// tables, pbu_math.cuh: 12,288 B with their alignment slack); taken off the opt-in maximum when the tiles are planned
#define PBU_STATIC_SMEM_RESERVE 25600

// chain slices of the moment reduction (pbu_moments.cu): partial sums [L][slices], added in slice order
#define PBU_MOMENT_SLICES 64

struct pbu_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    pbu_sampler_cfg cfg{};
    pbu::ProblemDev prob{};
    int d = 0, P = 0, nt = 0;
    int64_t C = 0;
    int64_t iteration = 0;
    bool initialised = false;
    const pbu::MhKernelEntry *mh = nullptr;
    const pbu::IsdeKernelEntry *isde = nullptr;
    const void *step_fn = nullptr;   // the planned step kernel (mh->fn or isde->fn)
    int tile_points = 0;          // points per shared-memory tile of the likelihood records (= n_points when resident)
    int tile_points_raw = 0;      // same for the raw records
    double *d_Cmat = nullptr;     // ISDE: packed symmetric preconditioner per chain
    const pbu::EvalKernelEntry *ev = nullptr;
    int lanes = 1;
    int q = 1;                    // chains per thread
    int n_full_warps = 0;         // mixed kernels: warps per block with `lanes` lanes per chain, the rest use 2 * lanes
    int groups_per_block = 0;     // chain groups (G-lane teams) per block
    int stagger_cycles = 0;       // cooperative kernels: start offset between the warps of one scheduler (a quarter iteration)
    int state_smem_off = 0;       // cooperative adaptive kernels: offset of the per-thread adaptive state in shared memory
    size_t launch_smem = 0;       // dynamic shared memory of the planned step kernel (records + staged state)
    int block = 128;
    int grid = 1;
    size_t smem_bytes = 0;        // step / log-posterior kernels: barrier + likelihood records
    size_t smem_raw_bytes = 0;    // model-evaluation kernel: barrier + raw records
    std::vector<double> R0;       // packed upper Cholesky factor of the shared proposal covariance
    // K4 (pbu_moments.cu): persistent scratch, taken once when the engine is created
    double *d_mom_partial = nullptr;   // [L][PBU_MOMENT_SLICES] slice partials
    double *d_mom_scratch = nullptr;   // [L + d] sums and reference point of the host-buffer entry point
    // pooled-covariance mode of the non-adaptive MH kernels: the shared factor R0 is a kernel parameter, so the factor the
    // device computed comes back through pinned memory; pbu_engine_run waits for the event before its next launch
    double *h_R0_pinned = nullptr;     // [nt + 1]: R packed, ok flag
    cudaEvent_t r0_event = nullptr;
    bool r0_pending = false;
    // device allocations
    double *d_records = nullptr;
    // uniform-weight form of the likelihood records ([x | w*y], Model::Uniform), when every point has the same weight and a
    // model offers it; the step kernel the planner picked reads it when uniform_w_kernel is set
    double *d_records_u = nullptr;
    double w_uniform = 0.0;
    size_t smem_bytes_u = 0;
    bool uniform_w_kernel = false;
    double *d_raw = nullptr;
    pbu::ChainState st{};
    pbu::StreamsDev streams{};
    double *d_normals = nullptr, *d_uniforms = nullptr;
    std::vector<void *> allocs;
};

