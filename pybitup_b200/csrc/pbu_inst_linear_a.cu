// Kernel instantiations for one slice of the (model, P, d) table; see pbu_inst.cuh.
#include "pbu_inst.cuh"
using namespace pbu;
#define MH_LIST PBU_MH_ALL(LinearModel<1>, 1, 1), PBU_MH_ALL(LinearModel<2>, 2, 2), PBU_MH_ALL(LinearModel<3>, 3, 3)
#define EV_LIST PBU_EVAL(LinearModel<1>, 1, 1), PBU_EVAL(LinearModel<2>, 2, 2), PBU_EVAL(LinearModel<3>, 3, 3)
#define ISDE_LIST PBU_ISDE_ALL(LinearModel<1>, 1, 1), PBU_ISDE_ALL(LinearModel<2>, 2, 2), PBU_ISDE_ALL(LinearModel<3>, 3, 3)
#define HMC_LIST PBU_HMC_ALL(LinearModel<1>, 1, 1), PBU_HMC_ALL(LinearModel<2>, 2, 2), PBU_HMC_ALL(LinearModel<3>, 3, 3)
PBU_DEFINE_SLICE(linear_a, MH_LIST, EV_LIST, ISDE_LIST, HMC_LIST)
