// Kernel instantiations for one slice of the (model, P, d) table; see pbu_inst.cuh.  Models with FIXED parameters (d < P
// uncertain ones, Model.run_model, bayesian_inference.py:351-366) and the odd d = P sizes: one lane and one chain per thread.
#include "pbu_inst.cuh"
using namespace pbu;
#define MH_LIST PBU_MH_SEQ(PolyModel<6>, 6, 1), PBU_MH_SEQ(PolyModel<6>, 6, 2), PBU_MH_SEQ(PolyModel<6>, 6, 3), PBU_MH_SEQ(PolyModel<6>, 6, 4), PBU_MH_SEQ(PolyModel<6>, 6, 5)
#define EV_LIST PBU_EVAL(PolyModel<6>, 6, 1), PBU_EVAL(PolyModel<6>, 6, 2), PBU_EVAL(PolyModel<6>, 6, 3), PBU_EVAL(PolyModel<6>, 6, 4), PBU_EVAL(PolyModel<6>, 6, 5)
#define ISDE_LIST PBU_ISDE_SEQ(PolyModel<6>, 6, 1), PBU_ISDE_SEQ(PolyModel<6>, 6, 2), PBU_ISDE_SEQ(PolyModel<6>, 6, 3), PBU_ISDE_SEQ(PolyModel<6>, 6, 4), PBU_ISDE_SEQ(PolyModel<6>, 6, 5)
#define HMC_LIST PBU_HMC_SEQ(PolyModel<6>, 6, 1), PBU_HMC_SEQ(PolyModel<6>, 6, 2), PBU_HMC_SEQ(PolyModel<6>, 6, 3), PBU_HMC_SEQ(PolyModel<6>, 6, 4), PBU_HMC_SEQ(PolyModel<6>, 6, 5)
PBU_DEFINE_SLICE(poly_e, MH_LIST, EV_LIST, ISDE_LIST, HMC_LIST)
