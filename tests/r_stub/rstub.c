/* Implementation of the stand-in R runtime declared in R.h / Rinternals.h (test infrastructure, see R.h). */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"
#include "R_ext/Rdynload.h"
#include "rstub.h"

struct SEXPREC { int type; R_xlen_t len; void* data; SEXP names; };
static struct SEXPREC nil_rec = {0, 0, NULL, NULL}, names_sym = {1, 0, NULL, NULL};
SEXP R_NilValue = &nil_rec, R_NamesSymbol = &names_sym;

jmp_buf rstub_error_jmp;        /* the driver's top level: Rf_error lands here */
int rstub_error_armed = 0;
char rstub_last_error[1024];
long rstub_interrupt_after = -1; /* >= 0: the (k+1)-th R_CheckUserInterrupt raises a user interrupt */
long rstub_interrupt_checks = 0;
static jmp_buf* toplevel_ctx = NULL;

SEXP Rf_allocVector(unsigned type, R_xlen_t n) {
  SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
  s->type = (int)type; s->len = n; s->names = R_NilValue;
  size_t el = type == REALSXP ? sizeof(double) : type == INTSXP ? sizeof(int) : sizeof(SEXP);
  s->data = calloc((size_t)(n > 0 ? n : 1), el);
  if (type == VECSXP || type == STRSXP) for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
  return s;
}
SEXP Rf_allocMatrix(unsigned type, int r, int c) { return Rf_allocVector(type, (R_xlen_t)r * c); }
SEXP Rf_mkChar(const char* c) {
  SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
  s->type = CHARSXP; s->len = (R_xlen_t)strlen(c); s->data = strdup(c); s->names = R_NilValue;
  return s;
}
SEXP Rf_mkNamed(unsigned type, const char** names) {
  R_xlen_t n = 0;
  while (names[n][0]) ++n;
  SEXP s = Rf_allocVector(type, n), nm = Rf_allocVector(STRSXP, n);
  for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)nm->data)[i] = Rf_mkChar(names[i]);
  s->names = nm;
  return s;
}
SEXP Rf_getAttrib(SEXP x, SEXP what) { return what == R_NamesSymbol ? x->names : R_NilValue; }
SEXP Rf_setAttrib(SEXP x, SEXP what, SEXP v) { if (what == R_NamesSymbol) x->names = v; return v; }
R_xlen_t XLENGTH(SEXP x) { return x->len; }
const char* CHAR(SEXP x) { return (const char*)x->data; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; return v; }
double* REAL(SEXP x) { if (x->type != REALSXP) Rf_error("REAL() on a non-double vector"); return (double*)x->data; }
int* INTEGER(SEXP x) { if (x->type != INTSXP) Rf_error("INTEGER() on a non-integer vector"); return (int*)x->data; }
double Rf_asReal(SEXP x) { return x->type == REALSXP ? ((double*)x->data)[0] : (double)((int*)x->data)[0]; }
int Rf_asInteger(SEXP x) { return x->type == INTSXP ? ((int*)x->data)[0] : (int)((double*)x->data)[0]; }
SEXP Rf_protect(SEXP s) { return s; }
void Rf_unprotect(int n) { (void)n; }
char* R_alloc(size_t n, int size) { return (char*)calloc(n ? n : 1, (size_t)size); }
void Rprintf(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); fflush(stdout); }
void Rf_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(rstub_last_error, sizeof(rstub_last_error), fmt, ap); va_end(ap);
  if (rstub_error_armed) longjmp(rstub_error_jmp, 1);
  fprintf(stderr, "Error: %s\n", rstub_last_error);
  exit(3);
}
void R_CheckUserInterrupt(void) {
  ++rstub_interrupt_checks;
  if (rstub_interrupt_after >= 0 && rstub_interrupt_checks > rstub_interrupt_after && toplevel_ctx) longjmp(*toplevel_ctx, 1);
}
Rboolean R_ToplevelExec(void (*fun)(void*), void* data) {
  jmp_buf ctx, *saved = toplevel_ctx;
  toplevel_ctx = &ctx;
  Rboolean ok = TRUE;
  if (setjmp(ctx) == 0) fun(data); else ok = FALSE;
  toplevel_ctx = saved;
  return ok;
}
int R_registerRoutines(DllInfo* d, const void* a, const R_CallMethodDef* c, const void* b, const void* e) {
  (void)d; (void)a; (void)b; (void)e; (void)c; return 1;
}
int R_useDynamicSymbols(DllInfo* d, int v) { (void)d; (void)v; return 1; }
