// Kernel instantiations for one slice of the (model, P, d) table; see pbu_inst.cuh.
#include "pbu_inst.cuh"
using namespace pbu;
#define MH_LIST PBU_MH_ALL(PolyModel<5>, 5, 5), PBU_MH_ALL(PolyModel<6>, 6, 6)
#define EV_LIST PBU_EVAL(PolyModel<5>, 5, 5), PBU_EVAL(PolyModel<6>, 6, 6)
#define ISDE_LIST PBU_ISDE_ALL(PolyModel<5>, 5, 5), PBU_ISDE_ALL(PolyModel<6>, 6, 6)
#define HMC_LIST PBU_HMC_ALL(PolyModel<5>, 5, 5), PBU_HMC_ALL(PolyModel<6>, 6, 6)
PBU_DEFINE_SLICE(poly_b, MH_LIST, EV_LIST, ISDE_LIST, HMC_LIST)
