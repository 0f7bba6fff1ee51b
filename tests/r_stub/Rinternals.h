#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include <stddef.h>
#include "R.h"
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
enum { CHARSXP = 9, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19 };
extern SEXP R_NilValue, R_NamesSymbol;
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
R_xlen_t XLENGTH(SEXP);
const char* CHAR(SEXP);
SEXP STRING_ELT(SEXP, R_xlen_t);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
double Rf_asReal(SEXP);
int Rf_asInteger(SEXP);
double* REAL(SEXP);
int* INTEGER(SEXP);
SEXP Rf_allocVector(unsigned, R_xlen_t);
SEXP Rf_allocMatrix(unsigned, int, int);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
SEXP Rf_mkNamed(unsigned, const char**);
SEXP Rf_mkChar(const char*);
void Rf_error(const char*, ...) __attribute__((noreturn));
Rboolean R_ToplevelExec(void (*fun)(void*), void* data);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#endif
