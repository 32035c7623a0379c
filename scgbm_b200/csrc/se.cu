// se.cu -- K8 placeholder (implemented next)
#include "common.cuh"
namespace scgbm {
void get_se_impl(const scgbm_se_args* a, scgbm_se_out* out) { (void)a; (void)out; throw Error(SCGBM_ERR_INVALID, "get_se: not built yet"); }
}
