#include "pbu_inst.cuh"
using namespace pbu;
#define MH_LIST PBU_MH_ALL(HeatModel, 6, 1), PBU_MH_ALL(HeatModel, 6, 2), PBU_MH_ALL(HeatModel, 6, 3), PBU_MH_ALL(HeatModel, 6, 6)
#define EV_LIST PBU_EVAL(HeatModel, 6, 1), PBU_EVAL(HeatModel, 6, 2), PBU_EVAL(HeatModel, 6, 3), PBU_EVAL(HeatModel, 6, 6)
PBU_DEFINE_SLICE(heat, MH_LIST, EV_LIST)
