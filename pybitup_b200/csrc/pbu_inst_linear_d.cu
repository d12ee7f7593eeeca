// Kernel instantiations for one slice of the (model, P, d) table; see pbu_inst.cuh.
#include "pbu_inst.cuh"
using namespace pbu;
#define MH_LIST PBU_MH_ALL(LinearModel<10>, 10, 10)
#define EV_LIST PBU_EVAL(LinearModel<10>, 10, 10)
#define ISDE_LIST PBU_ISDE_ALL(LinearModel<10>, 10, 10), PBU_ISDE_MIXED(LinearModel<10>, 10, 10, 1, 2)
#define HMC_LIST PBU_HMC_ALL(LinearModel<10>, 10, 10)
PBU_DEFINE_SLICE(linear_d, MH_LIST, EV_LIST, ISDE_LIST, HMC_LIST)
