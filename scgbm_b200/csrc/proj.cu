// proj.cu -- K9 placeholder (implemented next)
#include "common.cuh"
namespace scgbm {
void project_impl(const scgbm_proj_args* a, scgbm_proj_out* out) { (void)a; (void)out; throw Error(SCGBM_ERR_INVALID, "project: not built yet"); }
}
