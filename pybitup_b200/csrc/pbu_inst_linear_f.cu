// Kernel instantiations for one slice of the (model, P, d) table; see pbu_inst.cuh.  Models with FIXED parameters (d < P
// uncertain ones, Model.run_model, bayesian_inference.py:351-366) and the odd d = P sizes: one lane and one chain per thread.
#include "pbu_inst.cuh"
using namespace pbu;
#define MH_LIST PBU_MH_SEQ(LinearModel<5>, 5, 1), PBU_MH_SEQ(LinearModel<5>, 5, 2), PBU_MH_SEQ(LinearModel<5>, 5, 3), PBU_MH_SEQ(LinearModel<5>, 5, 4)
#define EV_LIST PBU_EVAL(LinearModel<5>, 5, 1), PBU_EVAL(LinearModel<5>, 5, 2), PBU_EVAL(LinearModel<5>, 5, 3), PBU_EVAL(LinearModel<5>, 5, 4)
#define ISDE_LIST PBU_ISDE_SEQ(LinearModel<5>, 5, 1), PBU_ISDE_SEQ(LinearModel<5>, 5, 2), PBU_ISDE_SEQ(LinearModel<5>, 5, 3), PBU_ISDE_SEQ(LinearModel<5>, 5, 4)
#define HMC_LIST PBU_HMC_SEQ(LinearModel<5>, 5, 1), PBU_HMC_SEQ(LinearModel<5>, 5, 2), PBU_HMC_SEQ(LinearModel<5>, 5, 3), PBU_HMC_SEQ(LinearModel<5>, 5, 4)
PBU_DEFINE_SLICE(linear_f, MH_LIST, EV_LIST, ISDE_LIST, HMC_LIST)
