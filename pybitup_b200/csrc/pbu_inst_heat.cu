// Kernel instantiations for one slice of the (model, P, d) table; see pbu_inst.cuh.
#include "pbu_inst.cuh"
using namespace pbu;
#define MH_LIST PBU_MH_ALL(HeatModel, 6, 1), PBU_MH_ALL(HeatModel, 6, 2), PBU_MH_ALL(HeatModel, 6, 3), PBU_MH_ALL(HeatModel, 6, 6)
#define EV_LIST PBU_EVAL(HeatModel, 6, 1), PBU_EVAL(HeatModel, 6, 2), PBU_EVAL(HeatModel, 6, 3), PBU_EVAL(HeatModel, 6, 6)
#define ISDE_LIST PBU_ISDE_ALL(HeatModel, 6, 1), PBU_ISDE_ALL(HeatModel, 6, 2), PBU_ISDE_ALL(HeatModel, 6, 3), PBU_ISDE_ALL(HeatModel, 6, 6)
#define HMC_LIST PBU_HMC_ALL(HeatModel, 6, 1), PBU_HMC_ALL(HeatModel, 6, 2), PBU_HMC_ALL(HeatModel, 6, 3), PBU_HMC_ALL(HeatModel, 6, 6)
PBU_DEFINE_SLICE(heat, MH_LIST, EV_LIST, ISDE_LIST, HMC_LIST)
