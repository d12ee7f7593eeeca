// Concatenates the per-file kernel table slices.
#include <vector>
#include "pbu_registry.h"

template <class E, class F>
static void append(std::vector<E> &v, F slice) {
    int n = 0;
    const E *t = slice(&n);
    v.insert(v.end(), t, t + n);
}

const pbu::MhKernelEntry *pbu_mh_kernels(int *n) {
    static const std::vector<pbu::MhKernelEntry> all = [] {
        std::vector<pbu::MhKernelEntry> v;
        append(v, pbu_mh_slice_linear_a);
        append(v, pbu_mh_slice_linear_b);
        append(v, pbu_mh_slice_linear_c);
        append(v, pbu_mh_slice_linear_d);
        append(v, pbu_mh_slice_poly_a);
        append(v, pbu_mh_slice_poly_b);
        append(v, pbu_mh_slice_arrhenius);
        append(v, pbu_mh_slice_heat);
        append(v, pbu_mh_slice_soize);
        append(v, pbu_mh_slice_linear_e);
        append(v, pbu_mh_slice_linear_f);
        append(v, pbu_mh_slice_linear_g);
        append(v, pbu_mh_slice_linear_h);
        append(v, pbu_mh_slice_poly_c);
        append(v, pbu_mh_slice_poly_d);
        append(v, pbu_mh_slice_poly_e);
        append(v, pbu_mh_slice_poly_u);
        return v;
    }();
    *n = (int)all.size();
    return all.data();
}

const pbu::EvalKernelEntry *pbu_eval_kernels(int *n) {
    static const std::vector<pbu::EvalKernelEntry> all = [] {
        std::vector<pbu::EvalKernelEntry> v;
        append(v, pbu_eval_slice_linear_a);
        append(v, pbu_eval_slice_linear_b);
        append(v, pbu_eval_slice_linear_c);
        append(v, pbu_eval_slice_linear_d);
        append(v, pbu_eval_slice_poly_a);
        append(v, pbu_eval_slice_poly_b);
        append(v, pbu_eval_slice_arrhenius);
        append(v, pbu_eval_slice_heat);
        append(v, pbu_eval_slice_soize);
        append(v, pbu_eval_slice_linear_e);
        append(v, pbu_eval_slice_linear_f);
        append(v, pbu_eval_slice_linear_g);
        append(v, pbu_eval_slice_linear_h);
        append(v, pbu_eval_slice_poly_c);
        append(v, pbu_eval_slice_poly_d);
        append(v, pbu_eval_slice_poly_e);
        return v;
    }();
    *n = (int)all.size();
    return all.data();
}

const pbu::IsdeKernelEntry *pbu_isde_kernels(int *n) {
    static const std::vector<pbu::IsdeKernelEntry> all = [] {
        std::vector<pbu::IsdeKernelEntry> v;
        append(v, pbu_isde_slice_linear_a);
        append(v, pbu_isde_slice_linear_b);
        append(v, pbu_isde_slice_linear_c);
        append(v, pbu_isde_slice_linear_d);
        append(v, pbu_isde_slice_poly_a);
        append(v, pbu_isde_slice_poly_b);
        append(v, pbu_isde_slice_arrhenius);
        append(v, pbu_isde_slice_heat);
        append(v, pbu_isde_slice_soize);
        append(v, pbu_isde_slice_linear_e);
        append(v, pbu_isde_slice_linear_f);
        append(v, pbu_isde_slice_linear_g);
        append(v, pbu_isde_slice_linear_h);
        append(v, pbu_isde_slice_poly_c);
        append(v, pbu_isde_slice_poly_d);
        append(v, pbu_isde_slice_poly_e);
        return v;
    }();
    *n = (int)all.size();
    return all.data();
}

const pbu::IsdeKernelEntry *pbu_hmc_kernels(int *n) {
    static const std::vector<pbu::IsdeKernelEntry> all = [] {
        std::vector<pbu::IsdeKernelEntry> v;
        append(v, pbu_hmc_slice_linear_a);
        append(v, pbu_hmc_slice_linear_b);
        append(v, pbu_hmc_slice_linear_c);
        append(v, pbu_hmc_slice_linear_d);
        append(v, pbu_hmc_slice_poly_a);
        append(v, pbu_hmc_slice_poly_b);
        append(v, pbu_hmc_slice_arrhenius);
        append(v, pbu_hmc_slice_heat);
        append(v, pbu_hmc_slice_soize);
        append(v, pbu_hmc_slice_linear_e);
        append(v, pbu_hmc_slice_linear_f);
        append(v, pbu_hmc_slice_linear_g);
        append(v, pbu_hmc_slice_linear_h);
        append(v, pbu_hmc_slice_poly_c);
        append(v, pbu_hmc_slice_poly_d);
        append(v, pbu_hmc_slice_poly_e);
        return v;
    }();
    *n = (int)all.size();
    return all.data();
}
