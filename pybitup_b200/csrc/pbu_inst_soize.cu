// Kernel instantiations for one slice of the (model, P, d) table; see pbu_inst.cuh.
#include "pbu_inst.cuh"
using namespace pbu;
#define MH_LIST PBU_MH_ALGOS(SoizeModel, 2, 2, 1, 1, 0), PBU_MH_ALGOS(SoizeModel, 2, 2, 1, 2, 0)
#define EV_LIST PBU_EVAL(SoizeModel, 2, 2)
#define ISDE_LIST PBU_ISDE(SoizeModel, 2, 2, 1, 1, 0)
#define HMC_LIST PBU_HMC(SoizeModel, 2, 2, 1, 1, 0)
PBU_DEFINE_SLICE(soize, MH_LIST, EV_LIST, ISDE_LIST, HMC_LIST)
