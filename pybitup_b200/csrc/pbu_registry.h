// Table of compiled kernel variants; each pbu_inst_*.cu contributes a slice.
#pragma once
#include "pbu_common.cuh"

namespace pbu {

struct MhKernelEntry {
    int model, n_param, n_unc, lanes, chains_per_thread, algo;   // algo = pbu_algo_id (RWMH, AM, DR, DRAM)
    int streamed;             // 0: records resident in shared memory, 1: streamed in tiles
    int mixed;                // 1: the last warps of a block use 2 * lanes per chain
    int coop;                 // 1: cooperative teams (pbu_mh_coop.cuh): `lanes` lanes own `chains_per_thread` (= lanes) chains
    int rep;                  // 1: the model's lookup tables are conflict-free replicas in dynamic shared memory (Model::Rep): blocks of up to 1024 threads, one per SM
    int uniform_w;            // 1: Model::Uniform -- one weight for all points, folded into the coefficients; records [x | w*y] (pbu_models.cuh)
    const void *fn;   // __global__ void (*)(const MhParams)
};
struct IsdeKernelEntry {
    int model, n_param, n_unc, lanes, chains_per_thread, has_grad, streamed;
    int mixed;                // 1: the last warps of a block use 2 * lanes per chain (Ito-SDE only)
    const void *fn;   // __global__ void (*)(const IsdeParams)
};
struct EvalKernelEntry {
    int model, n_param, n_unc, sequential;
    int nc, ncr;              // doubles per likelihood / raw record (Model::NC, Model::NCR)
    int ops, wf;              // FP64 instructions / shared-memory wavefronts per point (launch planning)
    const void *eval_lp;      // __global__ void (*)(const ProblemDev, const double*, int64_t, double*)
    const void *eval_model;   // same signature
    const void *eval_grad;    // __global__ void (*)(const ProblemDev, const double*, int64_t, double*, int, int)
    const void *eval_model_rep;   // eval_model with conflict-free replicated lookup tables (1024-thread persistent blocks), or nullptr
};

}  // namespace pbu

const pbu::MhKernelEntry *pbu_mh_kernels(int *n);
const pbu::EvalKernelEntry *pbu_eval_kernels(int *n);
const pbu::IsdeKernelEntry *pbu_isde_kernels(int *n);
const pbu::IsdeKernelEntry *pbu_hmc_kernels(int *n);

// slices
#define PBU_DECLARE_SLICE(name)                                     \
    const pbu::MhKernelEntry *pbu_mh_slice_##name(int *n);          \
    const pbu::EvalKernelEntry *pbu_eval_slice_##name(int *n);      \
    const pbu::IsdeKernelEntry *pbu_isde_slice_##name(int *n);      \
    const pbu::IsdeKernelEntry *pbu_hmc_slice_##name(int *n);
PBU_DECLARE_SLICE(linear_a)
PBU_DECLARE_SLICE(linear_b)
PBU_DECLARE_SLICE(linear_c)
PBU_DECLARE_SLICE(linear_d)
PBU_DECLARE_SLICE(poly_a)
PBU_DECLARE_SLICE(poly_b)
PBU_DECLARE_SLICE(arrhenius)
PBU_DECLARE_SLICE(heat)
PBU_DECLARE_SLICE(soize)
PBU_DECLARE_SLICE(linear_e)
PBU_DECLARE_SLICE(linear_f)
PBU_DECLARE_SLICE(linear_g)
PBU_DECLARE_SLICE(linear_h)
PBU_DECLARE_SLICE(poly_c)
PBU_DECLARE_SLICE(poly_d)
PBU_DECLARE_SLICE(poly_e)
const pbu::MhKernelEntry *pbu_mh_slice_poly_u(int *n);
