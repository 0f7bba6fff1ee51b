/* Drives r-pkg/src/shim.c the way R would (through the stand-in runtime of this directory): one batch fit with progress
 * lines, one fit interrupted by the "user", one single-profile fit + generated quantities, and the whole FitOCT.R loop
 * body as one call.  Prints machine-readable lines that tests/test_gpu_rshim.py checks.  Needs a CUDA device. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"
#include "rstub.h"

SEXP foct_R_sample_batch(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP foct_R_sample(SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP foct_R_predict(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP foct_R_summary(SEXP, SEXP, SEXP);
SEXP foct_R_pipeline(SEXP, SEXP, SEXP, SEXP);

static SEXP num(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }

/* named list from (name, SEXP) pairs */
static SEXP mklist(int n, const char** names, SEXP* vals) {
  const char* nm[64];
  for (int i = 0; i < n; ++i) nm[i] = names[i];
  nm[n] = "";
  SEXP l = Rf_mkNamed(VECSXP, nm);
  for (int i = 0; i < n; ++i) SET_VECTOR_ELT(l, i, vals[i]);
  return l;
}
static SEXP get(SEXP l, const char* name) {
  SEXP names = Rf_getAttrib(l, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(l); ++i) if (!strcmp(CHAR(STRING_ELT(names, i)), name)) return VECTOR_ELT(l, i);
  return R_NilValue;
}

/* synthData.R-shaped profile j: x = 20..500, a + b exp(-x / (l0 (1 + m(x)))) + noise (an LCG normal: the numbers only
 * need to be plausible, the parity of the fit itself is checked elsewhere) */
static unsigned long long lcg = 88172645463325252ull;
static double unif(void) { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; return ((lcg >> 11) + 0.5) / 9007199254740992.0; }
static double gauss(void) { return sqrt(-2.0 * log(unif())) * cos(6.283185307179586 * unif()); }
static void profile(int j, int N, SEXP* x, SEXP* y, SEXP* uy) {
  *x = Rf_allocVector(REALSXP, N); *y = Rf_allocVector(REALSXP, N); *uy = Rf_allocVector(REALSXP, N);
  for (int i = 0; i < N; ++i) {
    const double xi = 20.0 + i, m = (j % 2) ? 10.0 * sin(xi / 50.0) / xi : 10.0 * sin(xi / 25.0) / xi;
    const double y0 = 2000.0 * exp(-xi / 150.0), sd = 0.5 * sqrt(y0 + 1.0);
    REAL(*x)[i] = xi;
    REAL(*y)[i] = 1000.0 + 2000.0 * exp(-xi / (150.0 * (1.0 + m))) + sd * gauss();
    REAL(*uy)[i] = sd;
  }
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 3, N = 481, Nn = 10, chains = 4, nb_warmup = 100, nb_iter = 200;
  SEXP xs = Rf_allocVector(VECSXP, n), ys = Rf_allocVector(VECSXP, n), us = Rf_allocVector(VECSXP, n);
  SEXP th0 = Rf_allocVector(REALSXP, 3 * n), S0 = Rf_allocVector(REALSXP, 9 * n);
  for (int j = 0; j < n; ++j) {
    SEXP x, y, uy;
    profile(j, N, &x, &y, &uy);
    SET_VECTOR_ELT(xs, j, x); SET_VECTOR_ELT(ys, j, y); SET_VECTOR_ELT(us, j, uy);
    const double t[3] = {1000.0, 2000.0, 300.0};
    for (int k = 0; k < 3; ++k) { REAL(th0)[3 * j + k] = t[k]; REAL(S0)[9 * j + 4 * k] = (0.05 * t[k]) * (0.05 * t[k]); }
  }
  const char* names[] = {"dataType", "Nn", "gridType", "rho", "lambda_rate", "theta0", "Sigma0", "prior_PD", "chains",
                         "nb_warmup", "nb_iter", "seed", "poll_ms", "save_warmup"};
  SEXP vals[] = {num(2), num(Nn), num(0), num(0.1), num(0.1), th0, S0, num(0), num(chains),
                 num(nb_warmup), num(nb_iter), num(1234), num(1), num(1)};
  SEXP ctl = mklist(14, names, vals);
  const int P_out = Nn + 7;

  /* ---- 1. batch fit with progress lines */
  rstub_error_armed = 1;
  if (setjmp(rstub_error_jmp)) { printf("UNEXPECTED_ERROR %s\n", rstub_last_error); return 2; }
  SEXP out = foct_R_sample_batch(num(0), xs, ys, us, ctl, R_NilValue);
  SEXP draws = get(out, "draws"), summ = get(out, "summary");
  printf("BATCH draws_len %ld expected %ld interrupt_checks %ld\n", (long)XLENGTH(draws), (long)n * nb_iter * chains * P_out,
         rstub_interrupt_checks);
  for (int j = 0; j < n; ++j)
    printf("BATCH_MEAN %d %.6f %.6f %.6f rhat %.4f\n", j, REAL(summ)[((size_t)j * P_out + 0) * 11], REAL(summ)[((size_t)j * P_out + 1) * 11],
           REAL(summ)[((size_t)j * P_out + 2) * 11], REAL(summ)[((size_t)j * P_out + 2) * 11 + 9]);

  /* ---- 2. the same fit, interrupted by the user at the 3rd poll: the shim must come back through Rf_error */
  rstub_interrupt_checks = 0; rstub_interrupt_after = 2;
  if (setjmp(rstub_error_jmp) == 0) {
    SEXP big_names[] = {0};
    (void)big_names;
    REAL(VECTOR_ELT(ctl, 10))[0] = 4000;  /* nb_iter: long enough to be interrupted for sure */
    foct_R_sample_batch(num(0), xs, ys, us, ctl, R_NilValue);
    printf("INTERRUPT not raised\n");
    return 2;
  } else {
    printf("INTERRUPT raised after %ld checks: %s\n", rstub_interrupt_checks, rstub_last_error);
  }
  rstub_interrupt_after = -1;
  REAL(VECTOR_ELT(ctl, 10))[0] = nb_iter;

  /* ---- 3. the library still works after a cancelled run: single profile + generated quantities for 2 draws */
  if (setjmp(rstub_error_jmp)) { printf("UNEXPECTED_ERROR %s\n", rstub_last_error); return 2; }
  SEXP one = foct_R_sample(num(0), VECTOR_ELT(xs, 0), VECTOR_ELT(ys, 0), VECTOR_ELT(us, 0), ctl);
  SEXP d1 = get(one, "draws");
  int same = XLENGTH(d1) == (R_xlen_t)nb_iter * chains * P_out;
  for (R_xlen_t i = 0; same && i < XLENGTH(d1); ++i) same = REAL(d1)[i] == REAL(draws)[i];  /* profile 0 of the batch, same id */
  printf("SINGLE equals_batch_profile0 %d\n", same);
  SEXP two = Rf_allocVector(REALSXP, 2 * P_out);
  memcpy(REAL(two), REAL(d1) + (size_t)(nb_iter - 1) * chains * P_out, 2 * P_out * sizeof(double));
  SEXP gq = foct_R_predict(num(0), VECTOR_ELT(xs, 0), VECTOR_ELT(ys, 0), VECTOR_ELT(us, 0), ctl, two);
  double rmax = 0.0;
  for (int i = 0; i < N; ++i) {
    const double m = REAL(get(gq, "m"))[i], r = REAL(get(gq, "resid"))[i];
    rmax = fmax(rmax, fabs(REAL(VECTOR_ELT(ys, 0))[i] - m - r));
  }
  printf("PREDICT n %ld max|y-m-resid| %.3g\n", (long)XLENGTH(get(gq, "m")), rmax);

  /* ---- 3b. summary(fit)$summary of the draws just returned, through foct_R_summary: equals the summary of the fit */
  const int n_post = nb_iter - nb_warmup;                      /* the fit's summary covers the post-warm-up draws */
  SEXP post = Rf_allocVector(REALSXP, (R_xlen_t)n_post * chains * P_out);
  memcpy(REAL(post), REAL(d1) + (size_t)nb_warmup * chains * P_out, (size_t)n_post * chains * P_out * sizeof(double));
  SEXP sm = foct_R_summary(post, num(n_post), num(chains));
  SEXP s1 = get(one, "summary");
  double smax = 0.0;
  for (int i = 0; i < P_out; ++i)
    for (int k = 0; k < 11; ++k) {
      const double a = REAL(sm)[(size_t)k * P_out + i], b = REAL(s1)[(size_t)i * 11 + k];
      if (a == a || b == b) smax = fmax(smax, fabs(a - b) / (fabs(b) + 1e-300));
    }
  printf("SUMMARY rows %ld max_rel_diff %.3g\n", (long)(XLENGTH(sm) / 11), smax);

  /* ---- 4. the loop body of FitOCT.R:84-124 for all profiles in one call */
  const char* pn[] = {"dataType", "Nn", "smooth_df", "priorType", "ru_theta", "rho_scale", "lambda_rate", "nb_warmup", "nb_iter",
                      "chains", "seed", "gate"};
  SEXP pv[] = {num(2), num(Nn), num(15), num(1), num(0.05), num(0), num(0.1), num(nb_warmup), num(nb_iter), num(chains), num(7), num(0)};
  SEXP pl = foct_R_pipeline(xs, ys, mklist(12, pn, pv), R_NilValue);
  printf("PIPELINE n_expgp %d uy_len %ld mono_theta %.3f %.3f %.3f\n", INTEGER(get(pl, "n_expgp"))[0], (long)XLENGTH(get(pl, "uy")),
         REAL(get(pl, "mono_theta"))[0], REAL(get(pl, "mono_theta"))[1], REAL(get(pl, "mono_theta"))[2]);
  printf("DONE\n");
  return 0;
}
