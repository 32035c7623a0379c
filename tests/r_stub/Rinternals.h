/* TEST INFRASTRUCTURE ONLY -- a minimal stand-in for R's C API (R.h / Rinternals.h), just large enough to
 * COMPILE AND RUN r-package/src/shim.c without an R installation (R is absent from this image and from the GPU
 * box).  It implements, with R's own semantics, exactly the entry points the shim uses; tests/r_stub/rstub.c is
 * the tiny runtime behind it and tests/r_stub/harness.c drives the shim the way .Call would.  With a real R this
 * directory is simply not on the include path. */
#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef unsigned char Rbyte;

enum { NILSXP = 0, SYMSXP = 1, CHARSXP = 9, LGLSXP = 10, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, RAWSXP = 24, S4SXP = 25 };

typedef struct rstub_sexp* SEXP;
extern SEXP R_NilValue, R_DimSymbol, R_NamesSymbol, R_ClassSymbol;

SEXP Rf_allocVector(int type, R_xlen_t n);
SEXP Rf_allocMatrix(int type, int nrow, int ncol);
SEXP Rf_alloc3DArray(int type, int n1, int n2, int n3);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_install(const char*);
SEXP Rf_mkChar(const char*);
SEXP Rf_mkString(const char*);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarReal(double);
SEXP Rf_lengthgets(SEXP, R_xlen_t);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int Rf_asLogical(SEXP);
Rboolean Rf_isInteger(SEXP);
Rboolean Rf_isReal(SEXP);
Rboolean Rf_isNull(SEXP);
Rboolean Rf_isMatrix(SEXP);
Rboolean Rf_inherits(SEXP, const char*);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
R_xlen_t Rf_xlength(SEXP);
int TYPEOF(SEXP);
int* INTEGER(SEXP);
int* LOGICAL(SEXP);
double* REAL(SEXP);
Rbyte* RAW(SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
const char* R_CHAR(SEXP);
SEXP R_do_slot(SEXP obj, SEXP name);
char* R_alloc(size_t n, int size);
void Rprintf(const char*, ...);
void Rf_error(const char*, ...) __attribute__((noreturn));
void Rf_warning(const char*, ...);
void R_CheckUserInterrupt(void);
Rboolean R_ToplevelExec(void (*fun)(void*), void* data);

#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define alloc3DArray Rf_alloc3DArray
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
#define getAttrib Rf_getAttrib
#define setAttrib Rf_setAttrib
#define install Rf_install
#define mkChar Rf_mkChar
#define mkString Rf_mkString
#define ScalarInteger Rf_ScalarInteger
#define ScalarReal Rf_ScalarReal
#define lengthgets Rf_lengthgets
#define asInteger Rf_asInteger
#define asReal Rf_asReal
#define asLogical Rf_asLogical
#define isInteger Rf_isInteger
#define isReal Rf_isReal
#define isNull Rf_isNull
#define isMatrix Rf_isMatrix
#define inherits Rf_inherits
#define nrows Rf_nrows
#define ncols Rf_ncols
#define XLENGTH Rf_xlength
#define CHAR R_CHAR
#define error Rf_error
#define warning Rf_warning

/* ---- what only the harness (tests/r_stub/harness.c) uses: building inputs the way R would ---- */
SEXP rstub_new_s4(const char* klass);
void rstub_set_slot(SEXP obj, const char* name, SEXP v);
extern int rstub_warning_count;
extern char rstub_last_warning[1024];

#ifdef __cplusplus
}
#endif
#endif
