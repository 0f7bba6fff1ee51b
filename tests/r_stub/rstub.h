#ifndef RSTUB_H
#define RSTUB_H
#include <setjmp.h>
extern jmp_buf rstub_error_jmp;
extern int rstub_error_armed;
extern char rstub_last_error[1024];
extern long rstub_interrupt_after, rstub_interrupt_checks;
#endif
