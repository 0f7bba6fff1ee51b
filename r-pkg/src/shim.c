/*
 * shim.c — R `.Call` adapter over include/fitoct_b200.h.  Logic-free on purpose:
 *   - inputs are R-owned, read-only; the library copies them to pinned host buffers itself;
 *   - outputs are fresh REALSXP vectors under PROTECT;
 *   - the core never throws or longjmps across the ABI: it returns an int status; on failure this shim
 *     unprotects and only then calls Rf_error with foct_last_error();
 *   - only the calling (R) thread touches the R API.
 * Cannot be compiled in this repository's image (no R.h); compile with `R CMD INSTALL r-pkg` on a box with R,
 * PKG_CPPFLAGS=-I<repo>/include  PKG_LIBS="-L<repo>/fitoct_b200 -lfitoct_b200".
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "fitoct_b200.h"

static double get_num(SEXP lst, const char* name, double dflt) {
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(lst); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return Rf_asReal(VECTOR_ELT(lst, i));
  return dflt;
}
static SEXP get_elt(SEXP lst, const char* name) {
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(lst); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(lst, i);
  return R_NilValue;
}

/* .Call("foct_R_sample", kind, x, y, uy, ctl)  — one profile, replaces rstan::sampling inside fitExpGP /
 * fitMonoExp (FitOCT.R:110-124).  ctl: named list with dataType, Nn, gridType, rho, lambda_rate, theta0,
 * Sigma0, prior_PD, chains, nb_warmup, nb_iter, seed, adapt_delta, max_treedepth. */
SEXP foct_R_sample(SEXP kind_, SEXP x, SEXP y, SEXP uy, SEXP ctl) {
  const int kind = Rf_asInteger(kind_);
  foct_problem P;
  memset(&P, 0, sizeof(P));
  P.N = (int)XLENGTH(x);
  P.x = REAL(x); P.y = REAL(y); P.uy = REAL(uy);
  P.dataType = (int)get_num(ctl, "dataType", 2);
  P.Nn = (int)get_num(ctl, "Nn", 10);
  P.gridType = (int)get_num(ctl, "gridType", 0);
  P.rho = get_num(ctl, "rho", 0.1);
  P.lambda_rate = get_num(ctl, "lambda_rate", 0.1);
  P.prior_PD = (int)get_num(ctl, "prior_PD", 0);
  P.id = (long long)get_num(ctl, "id", 0);
  SEXP th0 = get_elt(ctl, "theta0"), S0 = get_elt(ctl, "Sigma0");
  if (th0 != R_NilValue) memcpy(P.theta0, REAL(th0), 3 * sizeof(double));
  if (S0 != R_NilValue) memcpy(P.Sigma0, REAL(S0), 9 * sizeof(double)); /* symmetric: R column-major == row-major */

  foct_model_spec spec;
  foct_model_spec_default(&spec, kind);
  foct_sampler_cfg cfg;
  foct_sampler_cfg_default(&cfg);
  cfg.chains = (int)get_num(ctl, "chains", 4);
  cfg.n_warmup = (int)get_num(ctl, "nb_warmup", 500);
  cfg.n_iter = (int)get_num(ctl, "nb_iter", 1500);
  cfg.seed = (unsigned long long)get_num(ctl, "seed", 1234);
  cfg.adapt_delta = get_num(ctl, "adapt_delta", 0.8);
  cfg.max_treedepth = (int)get_num(ctl, "max_treedepth", 10);
  cfg.save_warmup = 1; /* traceplot(inc_warmup = TRUE), plotExpGP.R:46 */

  int D = 0, P_out = 0;
  if (foct_dims(kind, P.Nn, &D, &P_out)) Rf_error("fitoct_b200: %s", foct_last_error());
  const R_xlen_t rows = (R_xlen_t)cfg.n_iter * cfg.chains;
  SEXP draws = PROTECT(Rf_allocVector(REALSXP, rows * P_out));
  SEXP sp = PROTECT(Rf_allocVector(REALSXP, rows * FOCT_N_SAMPLER_PARAMS));
  SEXP summ = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)P_out * FOCT_N_SUMMARY_COLS));
  SEXP eps = PROTECT(Rf_allocVector(REALSXP, cfg.chains));
  SEXP invm = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)cfg.chains * D));
  foct_result R;
  memset(&R, 0, sizeof(R));
  R.draws = REAL(draws); R.sampler_params = REAL(sp); R.summary = REAL(summ);
  R.stepsize = REAL(eps); R.inv_metric = REAL(invm);
  const int rc = foct_sample(kind, &P, 1, &spec, &cfg, &R);
  if (rc) {
    UNPROTECT(5);
    Rf_error("fitoct_b200 error %d: %s", rc, foct_last_error());
  }
  const char* nm[] = {"draws", "sampler_params", "summary", "stepsize", "inv_metric", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(out, 0, draws); SET_VECTOR_ELT(out, 1, sp); SET_VECTOR_ELT(out, 2, summ);
  SET_VECTOR_ELT(out, 3, eps); SET_VECTOR_ELT(out, 4, invm);
  UNPROTECT(6);
  return out;
}

/* .Call("foct_R_monoexp_map", x, y, uy, dataType) -> list(theta, hessian, br, status, m, resid) */
SEXP foct_R_monoexp_map(SEXP x, SEXP y, SEXP uy, SEXP dataType) {
  foct_problem P;
  memset(&P, 0, sizeof(P));
  P.N = (int)XLENGTH(x);
  P.x = REAL(x); P.y = REAL(y); P.uy = REAL(uy);
  P.dataType = Rf_asInteger(dataType);
  P.rho = 1.0;
  foct_model_spec spec;
  foct_model_spec_default(&spec, FOCT_MONOEXP);
  SEXP th = PROTECT(Rf_allocVector(REALSXP, 3)), H = PROTECT(Rf_allocMatrix(REALSXP, 3, 3));
  SEXP br = PROTECT(Rf_allocVector(REALSXP, 1)), st = PROTECT(Rf_allocVector(INTSXP, 1));
  SEXP m = PROTECT(Rf_allocVector(REALSXP, P.N)), resid = PROTECT(Rf_allocVector(REALSXP, P.N));
  int rc = foct_monoexp_map(&P, 1, &spec, NULL, REAL(th), REAL(H), REAL(br), INTEGER(st));
  if (!rc) {
    double row[5] = {REAL(th)[0], REAL(th)[1], REAL(th)[2], REAL(br)[0], 0.0};
    rc = foct_predict(FOCT_MONOEXP, &P, &spec, row, 1, REAL(m), REAL(resid), NULL);
  }
  if (rc) {
    UNPROTECT(6);
    Rf_error("fitoct_b200 error %d: %s", rc, foct_last_error());
  }
  const char* nm[] = {"theta", "hessian", "br", "status", "m", "resid", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(out, 0, th); SET_VECTOR_ELT(out, 1, H); SET_VECTOR_ELT(out, 2, br);
  SET_VECTOR_ELT(out, 3, st); SET_VECTOR_ELT(out, 4, m); SET_VECTOR_ELT(out, 5, resid);
  UNPROTECT(7);
  return out;
}

/* .Call("foct_R_expgp_map", x, y, uy, ctl) -> list(par, hessian, status, m, resid, dL): fitExpGP(method = 'optim') */
SEXP foct_R_expgp_map(SEXP x, SEXP y, SEXP uy, SEXP ctl) {
  foct_problem P;
  memset(&P, 0, sizeof(P));
  P.N = (int)XLENGTH(x);
  P.x = REAL(x); P.y = REAL(y); P.uy = REAL(uy);
  P.dataType = (int)get_num(ctl, "dataType", 2);
  P.Nn = (int)get_num(ctl, "Nn", 10);
  P.gridType = (int)get_num(ctl, "gridType", 0);
  P.rho = get_num(ctl, "rho", 0.1);
  P.lambda_rate = get_num(ctl, "lambda_rate", 0.1);
  P.prior_PD = (int)get_num(ctl, "prior_PD", 0);
  SEXP th0 = get_elt(ctl, "theta0"), S0 = get_elt(ctl, "Sigma0");
  if (th0 != R_NilValue) memcpy(P.theta0, REAL(th0), 3 * sizeof(double));
  if (S0 != R_NilValue) memcpy(P.Sigma0, REAL(S0), 9 * sizeof(double));
  foct_model_spec spec;
  foct_model_spec_default(&spec, FOCT_EXPGP);
  const int D = P.Nn + 5, P_out = P.Nn + 7;
  SEXP par = PROTECT(Rf_allocVector(REALSXP, P_out)), H = PROTECT(Rf_allocMatrix(REALSXP, D, D));
  SEXP st = PROTECT(Rf_allocVector(INTSXP, 1));
  SEXP m = PROTECT(Rf_allocVector(REALSXP, P.N)), resid = PROTECT(Rf_allocVector(REALSXP, P.N)), dL = PROTECT(Rf_allocVector(REALSXP, P.N));
  int rc = foct_expgp_map(&P, 1, &spec, NULL, REAL(par), REAL(H), INTEGER(st));
  if (!rc) rc = foct_predict(FOCT_EXPGP, &P, &spec, REAL(par), 1, REAL(m), REAL(resid), REAL(dL));
  if (rc) {
    UNPROTECT(6);
    Rf_error("fitoct_b200 error %d: %s", rc, foct_last_error());
  }
  const char* nm[] = {"par", "hessian", "status", "m", "resid", "dL", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(out, 0, par); SET_VECTOR_ELT(out, 1, H); SET_VECTOR_ELT(out, 2, st);
  SET_VECTOR_ELT(out, 3, m); SET_VECTOR_ELT(out, 4, resid); SET_VECTOR_ELT(out, 5, dL);
  UNPROTECT(7);
  return out;
}

/* .Call("foct_R_estimate_noise", x, y, df, maxRate) -> list(uy, ySmooth, theta, info, status): FitOCTLib::estimateNoise
 * (FitOCT.R:89-91) */
SEXP foct_R_estimate_noise(SEXP x, SEXP y, SEXP df, SEXP maxRate) {
  foct_problem P;
  memset(&P, 0, sizeof(P));
  P.N = (int)XLENGTH(x);
  P.x = REAL(x); P.y = REAL(y); P.uy = NULL;
  P.dataType = 2;
  SEXP uy = PROTECT(Rf_allocVector(REALSXP, P.N)), ys = PROTECT(Rf_allocVector(REALSXP, P.N));
  SEXP th = PROTECT(Rf_allocVector(REALSXP, 2)), info = PROTECT(Rf_allocVector(REALSXP, 4));
  SEXP st = PROTECT(Rf_allocVector(INTSXP, 1));
  int rc = foct_estimate_noise(&P, 1, Rf_asReal(df), Rf_asReal(maxRate), REAL(uy), REAL(ys), REAL(th), REAL(info), INTEGER(st));
  if (rc) {
    UNPROTECT(5);
    Rf_error("fitoct_b200 error %d: %s", rc, foct_last_error());
  }
  const char* nm[] = {"uy", "ySmooth", "theta", "info", "status", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(out, 0, uy); SET_VECTOR_ELT(out, 1, ys); SET_VECTOR_ELT(out, 2, th);
  SET_VECTOR_ELT(out, 3, info); SET_VECTOR_ELT(out, 4, st);
  UNPROTECT(6);
  return out;
}

/* .Call("foct_R_birge_ci", ndf) -> c(lo, hi): the interval FitOCTLib::printBr compares br with (plotMonoExp.R:10) */
SEXP foct_R_birge_ci(SEXP ndf) {
  SEXP ci = PROTECT(Rf_allocVector(REALSXP, 2));
  int rc = foct_birge_ci(Rf_asReal(ndf), REAL(ci));
  if (rc) {
    UNPROTECT(1);
    Rf_error("fitoct_b200 error %d: %s", rc, foct_last_error());
  }
  UNPROTECT(1);
  return ci;
}

/* .Call("foct_R_exp_prior", x, y, uy, dataType, priorType, theta, hessian, ru_theta) -> list(theta0, Sigma0, ru):
 * FitOCTLib::estimateExpPrior (FitOCT.R:103-107); y = m + resid of the MonoExp fit */
SEXP foct_R_exp_prior(SEXP x, SEXP y, SEXP uy, SEXP dataType, SEXP priorType, SEXP theta, SEXP hessian, SEXP ru_theta) {
  foct_problem P;
  memset(&P, 0, sizeof(P));
  P.N = (int)XLENGTH(x);
  P.x = REAL(x); P.y = REAL(y); P.uy = REAL(uy);
  P.dataType = Rf_asInteger(dataType);
  SEXP t0 = PROTECT(Rf_allocVector(REALSXP, 3)), S0 = PROTECT(Rf_allocMatrix(REALSXP, 3, 3));
  SEXP ru = PROTECT(Rf_allocVector(REALSXP, 1));
  int rc = foct_estimate_exp_prior(&P, 1, Rf_asInteger(priorType), REAL(theta), REAL(hessian), Rf_asReal(ru_theta),
                                   REAL(t0), REAL(S0), REAL(ru));
  if (rc) {
    UNPROTECT(3);
    Rf_error("fitoct_b200 error %d: %s", rc, foct_last_error());
  }
  const char* nm[] = {"theta0", "Sigma0", "ru", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(out, 0, t0); SET_VECTOR_ELT(out, 1, S0); SET_VECTOR_ELT(out, 2, ru);
  UNPROTECT(4);
  return out;
}

/* .Call("foct_R_vb", x, y, uy, ctl) -> list(mean, draws, mu, omega, elbo, eta, iters, status): fitExpGP(method = 'vb')
 * (FitOCT.R:42).  ctl as for foct_R_sample plus rstan::vb's iter, grad_samples, elbo_samples, eval_elbo, output_samples,
 * adapt_engaged, adapt_iter, eta, tol_rel_obj and omega0. */
SEXP foct_R_vb(SEXP x, SEXP y, SEXP uy, SEXP ctl) {
  foct_problem P;
  memset(&P, 0, sizeof(P));
  P.N = (int)XLENGTH(x);
  P.x = REAL(x); P.y = REAL(y); P.uy = REAL(uy);
  P.dataType = (int)get_num(ctl, "dataType", 2);
  P.Nn = (int)get_num(ctl, "Nn", 10);
  P.gridType = (int)get_num(ctl, "gridType", 0);
  P.rho = get_num(ctl, "rho", 0.1);
  P.lambda_rate = get_num(ctl, "lambda_rate", 0.1);
  P.prior_PD = (int)get_num(ctl, "prior_PD", 0);
  SEXP th0 = get_elt(ctl, "theta0"), S0 = get_elt(ctl, "Sigma0");
  if (th0 != R_NilValue) memcpy(P.theta0, REAL(th0), 3 * sizeof(double));
  if (S0 != R_NilValue) memcpy(P.Sigma0, REAL(S0), 9 * sizeof(double));
  foct_model_spec spec;
  foct_model_spec_default(&spec, FOCT_EXPGP);
  foct_vb_cfg cfg;
  foct_vb_cfg_default(&cfg);
  cfg.iter = (int)get_num(ctl, "iter", cfg.iter);
  cfg.grad_samples = (int)get_num(ctl, "grad_samples", cfg.grad_samples);
  cfg.elbo_samples = (int)get_num(ctl, "elbo_samples", cfg.elbo_samples);
  cfg.eval_elbo = (int)get_num(ctl, "eval_elbo", cfg.eval_elbo);
  cfg.output_samples = (int)get_num(ctl, "output_samples", cfg.output_samples);
  cfg.adapt_engaged = (int)get_num(ctl, "adapt_engaged", cfg.adapt_engaged);
  cfg.adapt_iter = (int)get_num(ctl, "adapt_iter", cfg.adapt_iter);
  cfg.eta = get_num(ctl, "eta", cfg.eta);
  cfg.tol_rel_obj = get_num(ctl, "tol_rel_obj", cfg.tol_rel_obj);
  cfg.omega0 = get_num(ctl, "omega0", -3.0);
  cfg.seed = (unsigned long long)get_num(ctl, "seed", 1234);
  const int D = P.Nn + 5, P_out = P.Nn + 7;
  SEXP mean = PROTECT(Rf_allocVector(REALSXP, P_out));
  SEXP draws = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)cfg.output_samples * P_out));
  SEXP mu = PROTECT(Rf_allocVector(REALSXP, D)), om = PROTECT(Rf_allocVector(REALSXP, D));
  SEXP elbo = PROTECT(Rf_allocVector(REALSXP, 1)), eta = PROTECT(Rf_allocVector(REALSXP, 1));
  SEXP it = PROTECT(Rf_allocVector(INTSXP, 1)), st = PROTECT(Rf_allocVector(INTSXP, 1));
  foct_vb_result R;
  R.mean = REAL(mean); R.draws = REAL(draws); R.mu = REAL(mu); R.omega = REAL(om); R.elbo = REAL(elbo); R.eta = REAL(eta);
  R.iters = INTEGER(it); R.status = INTEGER(st);
  int rc = foct_vb(FOCT_EXPGP, &P, 1, &spec, &cfg, &R);
  if (rc) {
    UNPROTECT(8);
    Rf_error("fitoct_b200 error %d: %s", rc, foct_last_error());
  }
  const char* nm[] = {"mean", "draws", "mu", "omega", "elbo", "eta", "iters", "status", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(out, 0, mean); SET_VECTOR_ELT(out, 1, draws); SET_VECTOR_ELT(out, 2, mu); SET_VECTOR_ELT(out, 3, om);
  SET_VECTOR_ELT(out, 4, elbo); SET_VECTOR_ELT(out, 5, eta); SET_VECTOR_ELT(out, 6, it); SET_VECTOR_ELT(out, 7, st);
  UNPROTECT(9);
  return out;
}

static const R_CallMethodDef call_methods[] = {
    {"foct_R_sample", (DL_FUNC)&foct_R_sample, 5},
    {"foct_R_monoexp_map", (DL_FUNC)&foct_R_monoexp_map, 4},
    {"foct_R_expgp_map", (DL_FUNC)&foct_R_expgp_map, 4},
    {"foct_R_estimate_noise", (DL_FUNC)&foct_R_estimate_noise, 4},
    {"foct_R_birge_ci", (DL_FUNC)&foct_R_birge_ci, 1},
    {"foct_R_exp_prior", (DL_FUNC)&foct_R_exp_prior, 8},
    {"foct_R_vb", (DL_FUNC)&foct_R_vb, 4},
    {NULL, NULL, 0}};

void R_init_FitOCTb200(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
