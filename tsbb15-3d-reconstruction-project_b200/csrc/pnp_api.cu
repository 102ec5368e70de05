// (temporary skeleton: PnP + general-N entry points are filled in next)
#include "common.cuh"
using namespace rg;
extern "C" {
int rg_fmatrix_stls_host(void*, void*, int, const double*, const double*, double*) { set_error("not implemented"); return RG_ERR_ARG; }
int rg_pnp_ransac_host(void*, void*, int, int, const double*, const double*, int, int, const int*, double, int, int*, int*, double*, double*, unsigned char*, int*, double*, unsigned char*) { set_error("not implemented"); return RG_ERR_ARG; }
int rg_pnp_ransac_dev(void*, void*, int, int, const double*, const double*, int, int, const int*, double, int, int*, int*, double*, unsigned char*) { set_error("not implemented"); return RG_ERR_ARG; }
int rg_pnp_minimize_host(void*, void*, int, const double*, const double*, double*, double*) { set_error("not implemented"); return RG_ERR_ARG; }
int rg_pnp_score_count_host(void*, void*, int, const double*, const double*, int, const double*, double, int, int*) { set_error("not implemented"); return RG_ERR_ARG; }
}
