// Kernel instantiations for one slice of the (model, P, d) table; see pbu_inst.cuh.
#include "pbu_inst.cuh"
using namespace pbu;
#define MH_LIST PBU_MH_ALL(PolyModel<2>, 2, 2), PBU_MH_ALL(PolyModel<3>, 3, 3), PBU_MH_ALL(PolyModel<4>, 4, 4)
#define EV_LIST PBU_EVAL(PolyModel<2>, 2, 2), PBU_EVAL(PolyModel<3>, 3, 3), PBU_EVAL(PolyModel<4>, 4, 4)
#define ISDE_LIST PBU_ISDE_ALL(PolyModel<2>, 2, 2), PBU_ISDE_ALL(PolyModel<3>, 3, 3), PBU_ISDE_ALL(PolyModel<4>, 4, 4)
#define HMC_LIST PBU_HMC_ALL(PolyModel<2>, 2, 2), PBU_HMC_ALL(PolyModel<3>, 3, 3), PBU_HMC_ALL(PolyModel<4>, 4, 4)
PBU_DEFINE_SLICE(poly_a, MH_LIST, EV_LIST, ISDE_LIST, HMC_LIST)
