// Helper macros to instantiate kernel variants and describe them in a table slice.
#pragma once
#include "pbu_mh_kernel.cuh"
#include "pbu_registry.h"

#define PBU_MH(MODEL, P, D, G, A) \
    { MODEL::ID, P, D, G, A, (const void *)&pbu::mh_step_kernel<MODEL, D, G, (A != 0)> }
// all lane counts x {fixed, adaptive}
#define PBU_MH_ALL(MODEL, P, D)                                                                                  \
    PBU_MH(MODEL, P, D, 1, 0), PBU_MH(MODEL, P, D, 1, 1), PBU_MH(MODEL, P, D, 2, 0), PBU_MH(MODEL, P, D, 2, 1),  \
        PBU_MH(MODEL, P, D, 8, 0), PBU_MH(MODEL, P, D, 8, 1)
#define PBU_MH_SEQ(MODEL, P, D) PBU_MH(MODEL, P, D, 1, 0), PBU_MH(MODEL, P, D, 1, 1)
#define PBU_EVAL(MODEL, P, D)                                                                 \
    { MODEL::ID, P, D, MODEL::SEQUENTIAL ? 1 : 0, MODEL::NC, MODEL::NCR, (const void *)&pbu::eval_logpost_kernel<MODEL, D>, \
      (const void *)&pbu::eval_model_kernel<MODEL, D> }

#define PBU_DEFINE_SLICE(name, MH_LIST, EVAL_LIST)                                   \
    static const pbu::MhKernelEntry mh_tab_##name[] = {MH_LIST};                     \
    static const pbu::EvalKernelEntry ev_tab_##name[] = {EVAL_LIST};                 \
    const pbu::MhKernelEntry *pbu_mh_slice_##name(int *n) {                          \
        *n = (int)(sizeof(mh_tab_##name) / sizeof(mh_tab_##name[0]));                \
        return mh_tab_##name;                                                        \
    }                                                                                \
    const pbu::EvalKernelEntry *pbu_eval_slice_##name(int *n) {                      \
        *n = (int)(sizeof(ev_tab_##name) / sizeof(ev_tab_##name[0]));                \
        return ev_tab_##name;                                                        \
    }
