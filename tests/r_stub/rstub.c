/* TEST INFRASTRUCTURE ONLY -- the runtime behind tests/r_stub/Rinternals.h: vectors with attributes, a protect
 * counter, S4 slots as attributes, error() = message on stderr + exit(3) (R would longjmp to top level),
 * warning() recorded.  Memory is never freed: a harness run is one .Call. */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

struct rstub_attr { SEXP name, value; struct rstub_attr* next; };
struct rstub_sexp {
  int type;
  R_xlen_t n;
  void* data;                 /* int / double / Rbyte / SEXP[] / char[] */
  struct rstub_attr* attrs;
};

static struct rstub_sexp nil_obj = {NILSXP, 0, NULL, NULL};
SEXP R_NilValue = &nil_obj;
SEXP R_DimSymbol, R_NamesSymbol, R_ClassSymbol;
int rstub_warning_count = 0;
char rstub_last_warning[1024];
static int protect_depth = 0;

static SEXP new_obj(int type, R_xlen_t n, size_t elt) {
  SEXP s = (SEXP)calloc(1, sizeof(*s));
  s->type = type; s->n = n;
  s->data = calloc((size_t)(n > 0 ? n : 1), elt);
  if (!s || !s->data) { fprintf(stderr, "rstub: out of memory\n"); exit(4); }
  return s;
}
struct symtab { const char* name; SEXP sym; struct symtab* next; };
static struct symtab* symbols = NULL;
SEXP Rf_install(const char* name) {
  for (struct symtab* t = symbols; t; t = t->next) if (!strcmp(t->name, name)) return t->sym;
  SEXP s = new_obj(SYMSXP, (R_xlen_t)strlen(name) + 1, 1);
  strcpy((char*)s->data, name);
  struct symtab* t = (struct symtab*)malloc(sizeof(*t));
  t->name = (const char*)s->data; t->sym = s; t->next = symbols; symbols = t;
  return s;
}
__attribute__((constructor)) static void rstub_init(void) {
  R_DimSymbol = Rf_install("dim"); R_NamesSymbol = Rf_install("names"); R_ClassSymbol = Rf_install("class");
}
SEXP Rf_allocVector(int type, R_xlen_t n) {
  switch (type) {
    case LGLSXP: case INTSXP: return new_obj(type, n, sizeof(int));
    case REALSXP: return new_obj(type, n, sizeof(double));
    case RAWSXP: return new_obj(type, n, 1);
    case STRSXP: case VECSXP: {
      SEXP s = new_obj(type, n, sizeof(SEXP));
      for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
      return s;
    }
    default: Rf_error("rstub: allocVector of type %d", type);
  }
}
static SEXP with_dim(SEXP s, int nd, const int* d) {
  SEXP dim = Rf_allocVector(INTSXP, nd);
  for (int k = 0; k < nd; ++k) INTEGER(dim)[k] = d[k];
  Rf_setAttrib(s, R_DimSymbol, dim);
  return s;
}
SEXP Rf_allocMatrix(int type, int nrow, int ncol) {
  const int d[2] = {nrow, ncol};
  return with_dim(Rf_allocVector(type, (R_xlen_t)nrow * ncol), 2, d);
}
SEXP Rf_alloc3DArray(int type, int n1, int n2, int n3) {
  const int d[3] = {n1, n2, n3};
  return with_dim(Rf_allocVector(type, (R_xlen_t)n1 * n2 * n3), 3, d);
}
SEXP Rf_protect(SEXP s) { ++protect_depth; return s; }
void Rf_unprotect(int n) {
  protect_depth -= n;
  if (protect_depth < 0) { fprintf(stderr, "rstub: PROTECT stack imbalance\n"); exit(5); }
}
SEXP Rf_getAttrib(SEXP s, SEXP name) {
  if (!s || s == R_NilValue) return R_NilValue;
  for (struct rstub_attr* a = s->attrs; a; a = a->next) if (a->name == name) return a->value;
  return R_NilValue;
}
SEXP Rf_setAttrib(SEXP s, SEXP name, SEXP v) {
  for (struct rstub_attr* a = s->attrs; a; a = a->next) if (a->name == name) { a->value = v; return v; }
  struct rstub_attr* a = (struct rstub_attr*)malloc(sizeof(*a));
  a->name = name; a->value = v; a->next = s->attrs; s->attrs = a;
  return v;
}
SEXP Rf_mkChar(const char* c) {
  SEXP s = new_obj(CHARSXP, (R_xlen_t)strlen(c) + 1, 1);
  strcpy((char*)s->data, c);
  return s;
}
SEXP Rf_mkString(const char* c) { SEXP s = Rf_allocVector(STRSXP, 1); SET_STRING_ELT(s, 0, Rf_mkChar(c)); return s; }
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
SEXP Rf_lengthgets(SEXP s, R_xlen_t n) {
  SEXP r = Rf_allocVector(s->type, n);
  const size_t elt = s->type == REALSXP ? sizeof(double) : (s->type == RAWSXP ? 1 : (s->type == VECSXP || s->type == STRSXP ? sizeof(SEXP) : sizeof(int)));
  memcpy(r->data, s->data, (size_t)(n < s->n ? n : s->n) * elt);
  return r;
}
static void need(SEXP s, const char* what) { if (!s || s == R_NilValue || s->n < 1) Rf_error("rstub: %s of an empty object", what); }
int Rf_asInteger(SEXP s) { need(s, "asInteger"); return s->type == REALSXP ? (int)REAL(s)[0] : INTEGER(s)[0]; }
double Rf_asReal(SEXP s) { need(s, "asReal"); return s->type == REALSXP ? REAL(s)[0] : (double)INTEGER(s)[0]; }
int Rf_asLogical(SEXP s) { need(s, "asLogical"); return s->type == REALSXP ? REAL(s)[0] != 0.0 : INTEGER(s)[0] != 0; }
Rboolean Rf_isInteger(SEXP s) { return s && s->type == INTSXP ? TRUE : FALSE; }
Rboolean Rf_isReal(SEXP s) { return s && s->type == REALSXP ? TRUE : FALSE; }
Rboolean Rf_isNull(SEXP s) { return (!s || s == R_NilValue) ? TRUE : FALSE; }
Rboolean Rf_isMatrix(SEXP s) { SEXP d = Rf_getAttrib(s, R_DimSymbol); return (d != R_NilValue && d->n == 2) ? TRUE : FALSE; }
Rboolean Rf_inherits(SEXP s, const char* klass) {
  SEXP c = Rf_getAttrib(s, R_ClassSymbol);
  if (c == R_NilValue) return FALSE;
  for (R_xlen_t i = 0; i < c->n; ++i) if (!strcmp(R_CHAR(STRING_ELT(c, i)), klass)) return TRUE;
  return FALSE;
}
int Rf_nrows(SEXP s) { SEXP d = Rf_getAttrib(s, R_DimSymbol); if (d == R_NilValue) Rf_error("object is not a matrix"); return INTEGER(d)[0]; }
int Rf_ncols(SEXP s) { SEXP d = Rf_getAttrib(s, R_DimSymbol); if (d == R_NilValue || d->n < 2) Rf_error("object is not a matrix"); return INTEGER(d)[1]; }
R_xlen_t Rf_xlength(SEXP s) { return (!s || s == R_NilValue) ? 0 : s->n; }
int TYPEOF(SEXP s) { return s ? s->type : NILSXP; }
static void* typed(SEXP s, int t1, int t2, const char* acc) {
  if (!s || (s->type != t1 && s->type != t2)) Rf_error("%s() can only be applied to a matching vector, not a type-%d object", acc, s ? s->type : -1);
  return s->data;
}
int* INTEGER(SEXP s) { return (int*)typed(s, INTSXP, LGLSXP, "INTEGER"); }
int* LOGICAL(SEXP s) { return (int*)typed(s, LGLSXP, LGLSXP, "LOGICAL"); }
double* REAL(SEXP s) { return (double*)typed(s, REALSXP, REALSXP, "REAL"); }
Rbyte* RAW(SEXP s) { return (Rbyte*)typed(s, RAWSXP, RAWSXP, "RAW"); }
static void in_range(SEXP s, R_xlen_t i, const char* acc) {   /* R would corrupt memory here; the stand-in reports it */
  if (i < 0 || i >= s->n) Rf_error("%s: index %ld outside a vector of length %ld", acc, (long)i, (long)s->n);
}
SEXP VECTOR_ELT(SEXP s, R_xlen_t i) { SEXP* p = (SEXP*)typed(s, VECSXP, VECSXP, "VECTOR_ELT"); in_range(s, i, "VECTOR_ELT"); return p[i]; }
SEXP SET_VECTOR_ELT(SEXP s, R_xlen_t i, SEXP v) { SEXP* p = (SEXP*)typed(s, VECSXP, VECSXP, "SET_VECTOR_ELT"); in_range(s, i, "SET_VECTOR_ELT"); p[i] = v; return v; }
void SET_STRING_ELT(SEXP s, R_xlen_t i, SEXP v) { SEXP* p = (SEXP*)typed(s, STRSXP, STRSXP, "SET_STRING_ELT"); in_range(s, i, "SET_STRING_ELT"); p[i] = v; }
SEXP STRING_ELT(SEXP s, R_xlen_t i) { SEXP* p = (SEXP*)typed(s, STRSXP, STRSXP, "STRING_ELT"); in_range(s, i, "STRING_ELT"); return p[i]; }
const char* R_CHAR(SEXP s) { return (const char*)typed(s, CHARSXP, SYMSXP, "CHAR"); }
SEXP R_do_slot(SEXP obj, SEXP name) {
  SEXP v = Rf_getAttrib(obj, name);
  if (v == R_NilValue) Rf_error("no slot of name \"%s\" for this object", R_CHAR(name));
  return v;
}
char* R_alloc(size_t n, int size) { char* p = (char*)calloc(n ? n : 1, (size_t)size); if (!p) Rf_error("R_alloc failed"); return p; }
void Rprintf(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); fflush(stdout); }
void Rf_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  fprintf(stderr, "Error: "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n");
  va_end(ap);
  exit(3);
}
void Rf_warning(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(rstub_last_warning, sizeof(rstub_last_warning), fmt, ap);
  va_end(ap);
  ++rstub_warning_count;
  fprintf(stderr, "Warning message: %s\n", rstub_last_warning);
}
void R_CheckUserInterrupt(void) {}
Rboolean R_ToplevelExec(void (*fun)(void*), void* data) { fun(data); return TRUE; }
int R_registerRoutines(DllInfo* d, const void* c, const R_CallMethodDef* call, const void* f, const void* e) { (void)d; (void)c; (void)call; (void)f; (void)e; return 1; }
Rboolean R_useDynamicSymbols(DllInfo* d, Rboolean v) { (void)d; (void)v; return TRUE; }

SEXP rstub_new_s4(const char* klass) {
  SEXP s = new_obj(S4SXP, 0, 1);
  Rf_setAttrib(s, R_ClassSymbol, Rf_mkString(klass));
  return s;
}
void rstub_set_slot(SEXP obj, const char* name, SEXP v) { Rf_setAttrib(obj, Rf_install(name), v); }
