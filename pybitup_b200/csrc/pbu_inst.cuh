// Helper macros to instantiate kernel variants and describe them in a table slice.
#pragma once
#include "pbu_isde_kernel.cuh"
#include "pbu_mh_coop.cuh"
#include "pbu_registry.h"

// S = 0: records resident in shared memory; S = 1: streamed in tiles (data sets larger than shared memory)
#define PBU_MH(MODEL, P, D, G, Q, A, S) \
    { MODEL::ID, P, D, G, Q, A, S, 0, 0, 0, 0, (const void *)&pbu::mh_step_kernel<MODEL, D, G, A, Q, (S != 0), false> }
// resident records, replicated lookup tables (MODEL::Rep)
#define PBU_MH_REP(MODEL, P, D, G, Q, A) \
    { MODEL::ID, P, D, G, Q, A, 0, 0, 0, 1, 0, (const void *)&pbu::mh_step_kernel<typename MODEL::Rep, D, G, A, Q, false, false> }
// mixed: the last warps of each block use 2G lanes per chain (scheduler-level load balance)
#define PBU_MH_MIXED(MODEL, P, D, G, Q, A) \
    { MODEL::ID, P, D, G, Q, A, 0, 1, 0, 0, 0, (const void *)&pbu::mh_step_kernel<MODEL, D, G, A, Q, false, true> }
// cooperative: teams of G lanes own G chains (pbu_mh_coop.cuh); MX = 1: the last warps of a block use 2G lanes
#define PBU_MH_COOP(MODEL, P, D, G, A, MX) \
    { MODEL::ID, P, D, G, G, A, 0, MX, 1, 0, 0, (const void *)&pbu::mh_coop_kernel<MODEL, D, G, A, false, (MX != 0)> }
// ... of the model's uniform-weight form (MODEL::Uniform)
#define PBU_MH_COOP_U(MODEL, P, D, G, A, MX) \
    { MODEL::ID, P, D, G, G, A, 0, MX, 1, 0, 1, (const void *)&pbu::mh_coop_kernel<typename MODEL::Uniform, D, G, A, false, (MX != 0)> }
#define PBU_MH_COOP_U_ALGOS(MODEL, P, D, G, MX) \
    PBU_MH_COOP_U(MODEL, P, D, G, 0, MX), PBU_MH_COOP_U(MODEL, P, D, G, 1, MX), PBU_MH_COOP_U(MODEL, P, D, G, 2, MX), PBU_MH_COOP_U(MODEL, P, D, G, 3, MX)
#define PBU_MH_COOP_ALGOS(MODEL, P, D, G, MX) \
    PBU_MH_COOP(MODEL, P, D, G, 0, MX), PBU_MH_COOP(MODEL, P, D, G, 1, MX), PBU_MH_COOP(MODEL, P, D, G, 2, MX), PBU_MH_COOP(MODEL, P, D, G, 3, MX)
#define PBU_MH_ALGOS(MODEL, P, D, G, Q, S) \
    PBU_MH(MODEL, P, D, G, Q, 0, S), PBU_MH(MODEL, P, D, G, Q, 1, S), PBU_MH(MODEL, P, D, G, Q, 2, S), PBU_MH(MODEL, P, D, G, Q, 3, S)
// data-parallel models.  Replicated layout (lanes per chain, chains per thread): (1,1) (1,2) (2,2), (2,2) mixed, (8,1) for
// very few chains; cooperative teams of 4 lanes x 4 chains, plain and mixed; streamed records: (1,1) (1,2) (8,1)
#define PBU_MH_ALL(MODEL, P, D)                                                                                        \
    PBU_MH_ALGOS(MODEL, P, D, 1, 1, 0), PBU_MH_ALGOS(MODEL, P, D, 1, 2, 0), PBU_MH_ALGOS(MODEL, P, D, 2, 2, 0),        \
        PBU_MH_ALGOS(MODEL, P, D, 8, 1, 0), PBU_MH_ALGOS(MODEL, P, D, 1, 1, 1), PBU_MH_ALGOS(MODEL, P, D, 1, 2, 1),    \
        PBU_MH_ALGOS(MODEL, P, D, 8, 1, 1), PBU_MH_MIXED(MODEL, P, D, 2, 2, 0), PBU_MH_MIXED(MODEL, P, D, 2, 2, 1),     \
        PBU_MH_COOP_ALGOS(MODEL, P, D, 4, 0), PBU_MH_COOP_ALGOS(MODEL, P, D, 4, 1)
#define PBU_MH_SEQ(MODEL, P, D) PBU_MH_ALGOS(MODEL, P, D, 1, 1, 0), PBU_MH_ALGOS(MODEL, P, D, 1, 1, 1)
// sequential (ODE) models, delayed rejection: pairs of lanes evaluate both proposal stages of a chain at once
// (speculative delayed rejection, pbu_mh_kernel.cuh)
#define PBU_MH_SEQ_SPEC(MODEL, P, D) \
    PBU_MH(MODEL, P, D, 2, 1, 2, 0), PBU_MH(MODEL, P, D, 2, 1, 3, 0), PBU_MH(MODEL, P, D, 2, 1, 2, 1), PBU_MH(MODEL, P, D, 2, 1, 3, 1)
// ... with the model's replicated lookup tables: one lane per chain (RWMH, AM), the speculative pair (DR, DRAM)
#define PBU_MH_SEQ_REP(MODEL, P, D) \
    PBU_MH_REP(MODEL, P, D, 1, 1, 0), PBU_MH_REP(MODEL, P, D, 1, 1, 1), PBU_MH_REP(MODEL, P, D, 2, 1, 2), PBU_MH_REP(MODEL, P, D, 2, 1, 3)
// Ito-SDE step kernel variants (lanes per chain, chains per thread, streamed).  (1, 2, streamed): every record read from the
// tile ring feeds two chains -- the layout of BASELINE configuration 4 when there are enough chains to fill the GPU
#define PBU_ISDE(MODEL, P, D, G, Q, S) \
    { MODEL::ID, P, D, G, Q, MODEL::HAS_GRAD ? 1 : 0, S, 0, (const void *)&pbu::isde_step_kernel<MODEL, D, G, Q, (S != 0), false> }
// streamed records, mixed warp widths (wide state vectors: the layout of BASELINE configuration 4)
#define PBU_ISDE_MIXED(MODEL, P, D, G, Q) \
    { MODEL::ID, P, D, G, Q, MODEL::HAS_GRAD ? 1 : 0, 1, 1, (const void *)&pbu::isde_step_kernel<MODEL, D, G, Q, true, true> }
#define PBU_ISDE_ALL(MODEL, P, D)                                                                                   \
    PBU_ISDE(MODEL, P, D, 1, 1, 0), PBU_ISDE(MODEL, P, D, 1, 2, 0), PBU_ISDE(MODEL, P, D, 2, 2, 0), PBU_ISDE(MODEL, P, D, 4, 1, 0), \
        PBU_ISDE(MODEL, P, D, 8, 1, 0), PBU_ISDE(MODEL, P, D, 1, 1, 1), PBU_ISDE(MODEL, P, D, 4, 1, 1), PBU_ISDE(MODEL, P, D, 8, 1, 1), \
        PBU_ISDE(MODEL, P, D, 1, 2, 1)
#define PBU_ISDE_SEQ(MODEL, P, D) PBU_ISDE(MODEL, P, D, 1, 1, 0), PBU_ISDE(MODEL, P, D, 1, 1, 1)
// Hamiltonian Monte Carlo step kernel variants (same table type as ISDE; `kind` 1)
#define PBU_HMC(MODEL, P, D, G, Q, S) \
    { MODEL::ID, P, D, G, Q, MODEL::HAS_GRAD ? 1 : 0, S, 0, (const void *)&pbu::hmc_step_kernel<MODEL, D, G, Q, (S != 0)> }
#define PBU_HMC_ALL(MODEL, P, D) PBU_HMC(MODEL, P, D, 1, 1, 0), PBU_HMC(MODEL, P, D, 4, 1, 0), PBU_HMC(MODEL, P, D, 8, 1, 0), PBU_HMC(MODEL, P, D, 1, 1, 1)
#define PBU_HMC_SEQ(MODEL, P, D) PBU_HMC(MODEL, P, D, 1, 1, 0), PBU_HMC(MODEL, P, D, 1, 1, 1)
#define PBU_EVAL(MODEL, P, D)                                                                 \
    { MODEL::ID, P, D, MODEL::SEQUENTIAL ? 1 : 0, MODEL::NC, MODEL::NCR, MODEL::OPS, MODEL::WF, (const void *)&pbu::eval_logpost_kernel<MODEL, D>, \
      (const void *)&pbu::eval_model_kernel<MODEL, D>, (const void *)&pbu::eval_grad_kernel<MODEL, D>, pbu::RepEval<MODEL, D>::fn() }

#define PBU_DEFINE_SLICE(name, MH_LIST, EVAL_LIST, ISDE_LIST, HMC_LIST)              \
    static const pbu::IsdeKernelEntry hmc_tab_##name[] = {HMC_LIST};                 \
    const pbu::IsdeKernelEntry *pbu_hmc_slice_##name(int *n) {                       \
        *n = (int)(sizeof(hmc_tab_##name) / sizeof(hmc_tab_##name[0]));              \
        return hmc_tab_##name;                                                       \
    }                                                                                \
    static const pbu::MhKernelEntry mh_tab_##name[] = {MH_LIST};                     \
    static const pbu::EvalKernelEntry ev_tab_##name[] = {EVAL_LIST};                 \
    static const pbu::IsdeKernelEntry isde_tab_##name[] = {ISDE_LIST};               \
    const pbu::IsdeKernelEntry *pbu_isde_slice_##name(int *n) {                      \
        *n = (int)(sizeof(isde_tab_##name) / sizeof(isde_tab_##name[0]));            \
        return isde_tab_##name;                                                      \
    }                                                                                \
    const pbu::MhKernelEntry *pbu_mh_slice_##name(int *n) {                          \
        *n = (int)(sizeof(mh_tab_##name) / sizeof(mh_tab_##name[0]));                \
        return mh_tab_##name;                                                        \
    }                                                                                \
    const pbu::EvalKernelEntry *pbu_eval_slice_##name(int *n) {                      \
        *n = (int)(sizeof(ev_tab_##name) / sizeof(ev_tab_##name[0]));                \
        return ev_tab_##name;                                                        \
    }
