// Cooperative step kernels of the polynomial model in its uniform-weight form (PolyModelU, pbu_models.cuh): one weight for all
// points, records [x | w*y].  Only the MH table has entries; evaluation, Ito-SDE and HMC use PolyModel.
#include "pbu_inst.cuh"
using namespace pbu;
#define PBU_U(P) PBU_MH_COOP_U_ALGOS(PolyModel<P>, P, P, 4, 0), PBU_MH_COOP_U_ALGOS(PolyModel<P>, P, P, 4, 1)
static const pbu::MhKernelEntry mh_tab_poly_u[] = {PBU_U(2), PBU_U(3), PBU_U(4), PBU_U(5), PBU_U(6)};
const pbu::MhKernelEntry *pbu_mh_slice_poly_u(int *n) {
    *n = (int)(sizeof(mh_tab_poly_u) / sizeof(mh_tab_poly_u[0]));
    return mh_tab_poly_u;
}
