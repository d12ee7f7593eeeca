#include "pbu_inst.cuh"
using namespace pbu;
#define MH_LIST PBU_MH_ALL(PolyModel<2>, 2, 2), PBU_MH_ALL(PolyModel<3>, 3, 3), PBU_MH_ALL(PolyModel<4>, 4, 4)
#define EV_LIST PBU_EVAL(PolyModel<2>, 2, 2), PBU_EVAL(PolyModel<3>, 3, 3), PBU_EVAL(PolyModel<4>, 4, 4)
PBU_DEFINE_SLICE(poly_a, MH_LIST, EV_LIST)
