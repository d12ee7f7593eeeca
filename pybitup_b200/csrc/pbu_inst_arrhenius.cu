// Kernel instantiations for one slice of the (model, P, d) table; see pbu_inst.cuh.
#include "pbu_inst.cuh"
using namespace pbu;
#define MH_LIST PBU_MH_SEQ(ArrheniusModel, 4, 1), PBU_MH_SEQ(ArrheniusModel, 4, 2), PBU_MH_SEQ(ArrheniusModel, 4, 3), PBU_MH_SEQ(ArrheniusModel, 4, 4), \
    PBU_MH_SEQ_SPEC(ArrheniusModel, 4, 1), PBU_MH_SEQ_SPEC(ArrheniusModel, 4, 2), PBU_MH_SEQ_SPEC(ArrheniusModel, 4, 3), PBU_MH_SEQ_SPEC(ArrheniusModel, 4, 4), \
    PBU_MH_SEQ_REP(ArrheniusModel, 4, 1), PBU_MH_SEQ_REP(ArrheniusModel, 4, 2), PBU_MH_SEQ_REP(ArrheniusModel, 4, 3), PBU_MH_SEQ_REP(ArrheniusModel, 4, 4)
#define EV_LIST PBU_EVAL(ArrheniusModel, 4, 1), PBU_EVAL(ArrheniusModel, 4, 2), PBU_EVAL(ArrheniusModel, 4, 3), PBU_EVAL(ArrheniusModel, 4, 4)
#define ISDE_LIST PBU_ISDE_SEQ(ArrheniusModel, 4, 1), PBU_ISDE_SEQ(ArrheniusModel, 4, 2), PBU_ISDE_SEQ(ArrheniusModel, 4, 3), PBU_ISDE_SEQ(ArrheniusModel, 4, 4)
#define HMC_LIST PBU_HMC_SEQ(ArrheniusModel, 4, 1), PBU_HMC_SEQ(ArrheniusModel, 4, 2), PBU_HMC_SEQ(ArrheniusModel, 4, 3), PBU_HMC_SEQ(ArrheniusModel, 4, 4)
PBU_DEFINE_SLICE(arrhenius, MH_LIST, EV_LIST, ISDE_LIST, HMC_LIST)
