/* TEST INFRASTRUCTURE ONLY -- a do-nothing stand-in for libscgbm, so that the R shim's marshalling (argument
 * unpacking, list assembly, element order, error paths) can be EXECUTED on a machine without a GPU
 * (tests/test_r_shim.py::test_shim_list_assembly_without_a_device).  It fills a few recognisable values and returns
 * success; it is never linked into anything but that test's harness. */
#include <string.h>
#include "scgbm.h"
int scgbm_fit(const scgbm_fit_args* a, scgbm_fit_out* o) {
  o->n_obj = 3; o->n_iter = 4; o->x_rank = a->M; o->D[0] = 42.0; o->obj[0] = -1.0; o->obj[1] = -0.5; o->obj[2] = -0.25;
  return 0;
}
int scgbm_fit_begin(const scgbm_fit_args* a, scgbm_fit_handle** h) { (void)a; *h = (scgbm_fit_handle*)1; return 0; }
int scgbm_fit_step(scgbm_fit_handle* h, int32_t n, int32_t* done) { (void)h; (void)n; *done = 1; return 0; }
int scgbm_fit_end(scgbm_fit_handle* h, scgbm_fit_out* o) {
  (void)h;
  if (o) { o->n_obj = 3; o->n_iter = 4; o->x_rank = 2; o->D[0] = 42.0; o->obj[0] = -1.0; o->obj[1] = -0.5; o->obj[2] = -0.25; }
  return 0;
}
int scgbm_get_se(const scgbm_se_args* a, scgbm_se_out* o) { o->se_scores[0] = a->W ? 1.0 : 2.0; return 0; }
int scgbm_project(const scgbm_proj_args* a, scgbm_proj_out* o) { (void)a; (void)o; return 0; }
int scgbm_sim_data(const scgbm_simdata_args* a, scgbm_simdata_out* o) { (void)a; (void)o; return 0; }
int scgbm_cci_perturb(const scgbm_cci_args* a, double* out) { out[0] = (double)a->reps; return 0; }
int scgbm_device_count(void) { return 0; }
const char* scgbm_last_error(void) { return "fake"; }
