/* tests/r_stub — a minimal stand-in for the R C API (R is absent from this image), just enough to COMPILE AND
 * RUN r-pkg/src/shim.c under a C driver: vectors, named lists, PROTECT as a no-op, Rprintf to stdout, Rf_error as a longjmp
 * to the driver, and R_ToplevelExec / R_CheckUserInterrupt with a programmable "user interrupt".  Test infrastructure only. */
#ifndef RSTUB_R_H
#define RSTUB_R_H
#include <stddef.h>
typedef enum { FALSE = 0, TRUE } Rboolean;
void Rprintf(const char*, ...);
char* R_alloc(size_t n, int size);
void R_CheckUserInterrupt(void);
#endif
