/* TEST INFRASTRUCTURE ONLY -- drives r-package/src/shim.c the way R's .Call would, on top of the R-API stand-in
 * (rstub.c).  usage: harness <mode> <in.bin> <out.bin>
 *   in.bin : int32 I, J, M, max_iter, sparse(0/1), return_W(0/1), ngpu, then I*J int32 counts (column-major)
 *   out.bin: doubles -- n_names, then for each list element its length followed by its values (scalars as 1 value);
 *            the element NAMES go to stdout, one per line, so the test can check the reference's list order.
 * modes: "fit" = gbm.sc + get.se (from W, or from the kept factors when return_W = 0) + cci perturbation;
 *        "badcall" = get.se on a fit without W and without factors (must raise R's error, exit code 3). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"

SEXP scgbm_fit_R(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP scgbm_get_se_R(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP scgbm_cci_perturb_R(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP scgbm_device_count_R(void);

static SEXP by_name(SEXP list, const char* nm) {
  SEXP names = getAttrib(list, R_NamesSymbol);
  for (R_xlen_t k = 0; k < XLENGTH(list); ++k)
    if (!strcmp(CHAR(STRING_ELT(names, k)), nm)) return VECTOR_ELT(list, k);
  return R_NilValue;
}
static void put(FILE* f, SEXP v) {
  double n = (double)XLENGTH(v);
  fwrite(&n, sizeof n, 1, f);
  if (TYPEOF(v) == REALSXP) fwrite(REAL(v), sizeof(double), (size_t)XLENGTH(v), f);
  else for (R_xlen_t i = 0; i < XLENGTH(v); ++i) { double x = INTEGER(v)[i]; fwrite(&x, sizeof x, 1, f); }
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: harness <mode> <in.bin> <out.bin>\n"); return 2; }
  FILE* fi = fopen(argv[2], "rb");
  if (!fi) { perror("in.bin"); return 2; }
  int hdr[7];
  if (fread(hdr, sizeof(int), 7, fi) != 7) return 2;
  const int I = hdr[0], J = hdr[1], M = hdr[2], max_iter = hdr[3], sparse = hdr[4], return_W = hdr[5], ngpu = hdr[6];
  SEXP Yd = PROTECT(allocMatrix(INTSXP, I, J));
  if (fread(INTEGER(Yd), sizeof(int), (size_t)I * J, fi) != (size_t)I * J) return 2;
  fclose(fi);
  SEXP Y = Yd;
  if (sparse) {                                  /* as(Y, "dgCMatrix"): slots i, p, x, Dim */
    R_xlen_t nnz = 0;
    for (R_xlen_t k = 0; k < (R_xlen_t)I * J; ++k) nnz += INTEGER(Yd)[k] != 0;
    SEXP si = allocVector(INTSXP, nnz), sp = allocVector(INTSXP, J + 1), sx = allocVector(REALSXP, nnz), dim = allocVector(INTSXP, 2);
    R_xlen_t q = 0;
    for (int j = 0; j < J; ++j) {
      INTEGER(sp)[j] = (int)q;
      for (int i = 0; i < I; ++i) {
        const int v = INTEGER(Yd)[(R_xlen_t)j * I + i];
        if (v) { INTEGER(si)[q] = i; REAL(sx)[q] = v; ++q; }
      }
    }
    INTEGER(sp)[J] = (int)q; INTEGER(dim)[0] = I; INTEGER(dim)[1] = J;
    Y = rstub_new_s4("dgCMatrix");
    rstub_set_slot(Y, "i", si); rstub_set_slot(Y, "p", sp); rstub_set_slot(Y, "x", sx); rstub_set_slot(Y, "Dim", dim);
  }
  SEXP lgl_false = allocVector(LGLSXP, 1), lgl_W = allocVector(LGLSXP, 1), lgl_keep = allocVector(LGLSXP, 1);
  LOGICAL(lgl_false)[0] = 0; LOGICAL(lgl_W)[0] = return_W; LOGICAL(lgl_keep)[0] = !return_W;
  SEXP out = PROTECT(scgbm_fit_R(Y, ScalarInteger(M), ScalarInteger(max_iter), ScalarReal(1e-4), ScalarInteger(30), lgl_false,
                                 R_NilValue, ScalarInteger(1), lgl_W, lgl_false, ScalarInteger(-1), ScalarInteger(ngpu), lgl_keep));
  if (!strcmp(argv[1], "badcall")) {             /* return.W = FALSE and no kept factors: R's get.se cannot work either */
    scgbm_get_se_R(by_name(out, "U"), by_name(out, "scores"), R_NilValue, R_NilValue, R_NilValue, R_NilValue, R_NilValue,
                   R_NilValue, ScalarInteger(-1), ScalarReal(0.0), ScalarInteger(-1), ScalarInteger(1));
    return 0;                                    /* not reached: error() exits with code 3 */
  }
  SEXP se = PROTECT(scgbm_get_se_R(by_name(out, "U"), by_name(out, "scores"), by_name(out, "W"), by_name(out, "alpha"),
                                   by_name(out, "beta"), R_NilValue, by_name(out, "Xu"), by_name(out, "Xv"),
                                   return_W ? ScalarInteger(-1) : by_name(out, "x.rank"), ScalarReal(0.0), ScalarInteger(-1),
                                   ScalarInteger(ngpu)));
  SEXP pert = PROTECT(scgbm_cci_perturb_R(by_name(out, "scores"), VECTOR_ELT(se, 0), ScalarInteger(3), ScalarReal(21.0),
                                          ScalarInteger(0), ScalarInteger(-1)));
  FILE* fo = fopen(argv[3], "wb");
  if (!fo) { perror("out.bin"); return 2; }
  SEXP names = getAttrib(out, R_NamesSymbol);
  double nn = (double)XLENGTH(out);
  fwrite(&nn, sizeof nn, 1, fo);
  for (R_xlen_t k = 0; k < XLENGTH(out); ++k) { printf("%s\n", CHAR(STRING_ELT(names, k))); put(fo, VECTOR_ELT(out, k)); }
  put(fo, VECTOR_ELT(se, 0)); put(fo, VECTOR_ELT(se, 1)); put(fo, pert);
  fclose(fo);
  printf("warnings=%d devices=%d\n", rstub_warning_count, INTEGER(scgbm_device_count_R())[0]);
  UNPROTECT(4);
  return 0;
}
