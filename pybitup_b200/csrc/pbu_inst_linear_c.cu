#include "pbu_inst.cuh"
using namespace pbu;
#define MH_LIST PBU_MH_ALL(LinearModel<8>, 8, 8), PBU_MH_ALL(LinearModel<10>, 10, 10)
#define EV_LIST PBU_EVAL(LinearModel<8>, 8, 8), PBU_EVAL(LinearModel<10>, 10, 10)
PBU_DEFINE_SLICE(linear_c, MH_LIST, EV_LIST)
