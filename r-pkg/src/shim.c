/*
 * shim.c — R `.Call` adapter over include/fitoct_b200.h.  Logic-free on purpose:
 *   - inputs are R-owned, read-only; the library copies them to pinned host buffers itself;
 *   - outputs are fresh REALSXP vectors under PROTECT;
 *   - the core never throws or longjmps across the ABI: it returns an int status; on failure this shim
 *     unprotects and only then calls Rf_error with foct_last_error();
 *   - only the calling (R) thread touches the R API: sampling goes through foct_sample_cb, whose progress callback runs
 *     on the calling thread between polls — it prints rstan-style "Chain k: Iteration: i / n [ p%]" lines (what
 *     ShinyInterface/server.R:457-472 scrapes) and checks for a user interrupt (R_ToplevelExec around
 *     R_CheckUserInterrupt, so the longjmp of an interrupt never crosses the library); an interrupt cancels the
 *     kernels, the library frees its buffers and returns FOCT_ECANCELLED, and only then does the shim raise the error.
 * Cannot be compiled in this repository's image (no R.h); compile with `R CMD INSTALL r-pkg` on a box with R,
 * PKG_CPPFLAGS=-I<repo>/include  PKG_LIBS="-L<repo>/fitoct_b200 -lfitoct_b200".
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "fitoct_b200.h"

static SEXP get_elt(SEXP lst, const char* name);
static double get_num(SEXP lst, const char* name, double dflt) {
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(lst); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return Rf_asReal(VECTOR_ELT(lst, i));
  return dflt;
}
/* ---- progress + interrupts ---------------------------------------------------------------------- */
typedef struct { int chains, n_iter, n_warmup, last_pct, quiet, interrupted; } progress_t;
static void check_interrupt_fn(void* dummy) { (void)dummy; R_CheckUserInterrupt(); }
/* rstan prints one block per chain and the Shiny scraper computes ((chain - 1) * 100 + pct) / chains from the LAST line
 * (server.R:463-469).  All chains of a batch advance together here, so the overall fraction f is reported as the
 * chain k = floor(chains f) + 1 being pct = 100 (chains f - (k - 1)) % done: the scraper then shows exactly 100 f. */
static int progress_cb(double f, const char* phase, void* user) {
  progress_t* pg = (progress_t*)user;
  if (!R_ToplevelExec(check_interrupt_fn, NULL)) { pg->interrupted = 1; return 1; }
  const int overall = (int)(100.0 * f);
  if (!pg->quiet && overall != pg->last_pct) {
    pg->last_pct = overall;
    int k = (int)(f * pg->chains) + 1;
    if (k > pg->chains) k = pg->chains;
    int pct = (int)(100.0 * (f * pg->chains - (k - 1)) + 0.5);
    if (pct > 100) pct = 100;
    Rprintf("Chain %d: Iteration: %4d / %d [%3d%%]  (%s)\n", k, (int)(f * pg->n_iter), pg->n_iter, pct, phase);
  }
  return 0;
}

static void fill_problem(foct_problem* P, SEXP x, SEXP y, SEXP uy, SEXP ctl) {
  memset(P, 0, sizeof(*P));
  P->N = (int)XLENGTH(x);
  P->x = REAL(x); P->y = REAL(y); P->uy = uy == R_NilValue ? NULL : REAL(uy);
  P->dataType = (int)get_num(ctl, "dataType", 2);
  P->Nn = (int)get_num(ctl, "Nn", 10);
  P->gridType = (int)get_num(ctl, "gridType", 0);
  P->rho = get_num(ctl, "rho", 0.1);
  P->lambda_rate = get_num(ctl, "lambda_rate", 0.1);
  P->prior_PD = (int)get_num(ctl, "prior_PD", 0);
  P->id = (long long)get_num(ctl, "id", 0);
}
static void fill_cfg(foct_sampler_cfg* cfg, SEXP ctl) {
  foct_sampler_cfg_default(cfg);
  cfg->chains = (int)get_num(ctl, "chains", 4);
  cfg->n_warmup = (int)get_num(ctl, "nb_warmup", 500);
  cfg->n_iter = (int)get_num(ctl, "nb_iter", 1500);
  cfg->seed = (unsigned long long)get_num(ctl, "seed", 1234);
  cfg->adapt_delta = get_num(ctl, "adapt_delta", 0.8);
  cfg->max_treedepth = (int)get_num(ctl, "max_treedepth", 10);
  cfg->init_mode = (int)get_num(ctl, "init_mode", 0);
  cfg->rhat_target = get_num(ctl, "rhat_target", 0.0);
  cfg->max_extend = (int)get_num(ctl, "max_extend", 0);
  cfg->extend_iter = (int)get_num(ctl, "extend_iter", 0);
  cfg->save_warmup = (int)get_num(ctl, "save_warmup", 1); /* traceplot(inc_warmup = TRUE), plotExpGP.R:46 */
}
/* theta0 (3) / Sigma0 (9) of problem j from a vector (one profile) or a 3 x n / 9 x n matrix; a zero-length vector is
 * what as.numeric(NULL) gives and is NOT R_NilValue (it was memcpy'd from) */
static int copy_prior(foct_problem* P, SEXP th0, SEXP S0, R_xlen_t j) {
  if (th0 != R_NilValue && XLENGTH(th0) > 0) {
    if (XLENGTH(th0) < 3 * (j + 1)) return 1;
    memcpy(P->theta0, REAL(th0) + 3 * j, 3 * sizeof(double));
  }
  if (S0 != R_NilValue && XLENGTH(S0) > 0) {
    if (XLENGTH(S0) < 9 * (j + 1)) return 1;
    memcpy(P->Sigma0, REAL(S0) + 9 * j, 9 * sizeof(double)); /* symmetric: R column-major == row-major */
  }
  return 0;
}

static SEXP get_elt(SEXP lst, const char* name) {
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(lst); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(lst, i);
  return R_NilValue;
}

/* .Call("foct_R_sample_batch", kind, xs, ys, uys, ctl, devices) — n profiles in ONE call (a
 * batch-aware FitOCT.R hands all its files over at once instead of looping, FitOCT.R:74-124).  xs, ys, uys: lists of
 * numeric vectors (ragged lengths allowed); ctl as for foct_R_sample, with theta0 a 3 x n and Sigma0 a 9 x n matrix;
 * devices: integer vector of CUDA devices to shard over, or NULL.  Returns arrays whose R dims are the C layout
 * reversed: draws [P_out, chains, n_saved, n], sampler_params [6, chains, n_saved, n], summary [11, P_out, n], ...
 * Progress lines and interrupt checks as described at the top of this file. */
SEXP foct_R_sample_batch(SEXP kind_, SEXP xs, SEXP ys, SEXP uys, SEXP ctl, SEXP devices) {
  const int kind = Rf_asInteger(kind_);
  const R_xlen_t n = XLENGTH(xs);
  if (n < 1 || XLENGTH(ys) != n || XLENGTH(uys) != n) Rf_error("fitoct_b200: xs, ys, uys must be lists of the same length >= 1");
  foct_problem* P = (foct_problem*)R_alloc((size_t)n, sizeof(foct_problem));
  SEXP th0 = get_elt(ctl, "theta0"), S0 = get_elt(ctl, "Sigma0");
  const long long id0 = (long long)get_num(ctl, "id", 0);
  for (R_xlen_t j = 0; j < n; ++j) {
    SEXP x = VECTOR_ELT(xs, j), y = VECTOR_ELT(ys, j), uy = VECTOR_ELT(uys, j);
    if (XLENGTH(y) != XLENGTH(x) || XLENGTH(uy) != XLENGTH(x)) Rf_error("fitoct_b200: profile %d: x, y, uy differ in length", (int)j + 1);
    fill_problem(&P[j], x, y, uy, ctl);
    P[j].id = id0 + j;
    if (copy_prior(&P[j], th0, S0, j)) Rf_error("fitoct_b200: theta0 must hold 3 and Sigma0 9 numbers per profile");
  }
  if (kind == FOCT_EXPGP && (th0 == R_NilValue || XLENGTH(th0) == 0 || S0 == R_NilValue || XLENGTH(S0) == 0))
    Rf_error("fitoct_b200: fitExpGP needs theta0 and Sigma0 (FitOCT.R:116-117 passes estimateExpPrior's)");
  foct_model_spec spec;
  foct_model_spec_default(&spec, kind);
  foct_sampler_cfg cfg;
  fill_cfg(&cfg, ctl);
  if (devices != R_NilValue && XLENGTH(devices) > 0) { cfg.n_devices = (int)XLENGTH(devices); cfg.devices = INTEGER(devices); }
  int D = 0, P_out = 0;
  if (foct_dims(kind, P[0].Nn, &D, &P_out)) Rf_error("fitoct_b200: %s", foct_last_error());
  const int n_saved = cfg.save_warmup ? cfg.n_iter : cfg.n_iter - cfg.n_warmup;
  const int want_draws = (int)get_num(ctl, "draws", 1);
  const R_xlen_t rows = (R_xlen_t)n * n_saved * cfg.chains;
  int np = 0;
  SEXP draws = PROTECT(Rf_allocVector(REALSXP, want_draws ? rows * P_out : 0)); ++np;
  SEXP sp = PROTECT(Rf_allocVector(REALSXP, want_draws ? rows * FOCT_N_SAMPLER_PARAMS : 0)); ++np;
  SEXP summ = PROTECT(Rf_allocVector(REALSXP, n * P_out * FOCT_N_SUMMARY_COLS)); ++np;
  SEXP eps = PROTECT(Rf_allocVector(REALSXP, n * cfg.chains)); ++np;
  SEXP invm = PROTECT(Rf_allocVector(REALSXP, n * cfg.chains * D)); ++np;
  SEXP nlf = PROTECT(Rf_allocVector(REALSXP, n * cfg.chains * 2)); ++np;
  SEXP ndiv = PROTECT(Rf_allocVector(REALSXP, n * cfg.chains)); ++np;
  SEXP lastq = PROTECT(Rf_allocVector(REALSXP, n * cfg.chains * D)); ++np;
  SEXP next = PROTECT(Rf_allocVector(INTSXP, n)); ++np;
  foct_result R;
  memset(&R, 0, sizeof(R));
  if (want_draws) { R.draws = REAL(draws); R.sampler_params = REAL(sp); }
  R.summary = REAL(summ); R.stepsize = REAL(eps); R.inv_metric = REAL(invm); R.n_leapfrog = REAL(nlf);
  R.n_divergent = REAL(ndiv); R.last_q = REAL(lastq); R.n_extend = INTEGER(next);
  progress_t pg = {cfg.chains, cfg.n_iter, cfg.n_warmup, -1, (int)get_num(ctl, "quiet", 0), 0};
  const int rc = foct_sample_cb(kind, P, (int)n, &spec, &cfg, &R, progress_cb, &pg, (int)get_num(ctl, "poll_ms", 100));
  if (rc) {
    UNPROTECT(np);
    if (pg.interrupted) Rf_error("fitoct_b200: interrupted by the user (kernels cancelled, device buffers released)");
    Rf_error("fitoct_b200 error %d: %s", rc, foct_last_error());
  }
  const char* nm[] = {"draws", "sampler_params", "summary", "stepsize", "inv_metric", "n_leapfrog", "n_divergent",
                      "last_q", "n_extend", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm)); ++np;
  SET_VECTOR_ELT(out, 0, draws); SET_VECTOR_ELT(out, 1, sp); SET_VECTOR_ELT(out, 2, summ);
  SET_VECTOR_ELT(out, 3, eps); SET_VECTOR_ELT(out, 4, invm); SET_VECTOR_ELT(out, 5, nlf); SET_VECTOR_ELT(out, 6, ndiv);
  SET_VECTOR_ELT(out, 7, lastq); SET_VECTOR_ELT(out, 8, next);
  UNPROTECT(np);
  return out;
}

/* .Call("foct_R_sample", kind, x, y, uy, ctl)  — one profile, replaces rstan::sampling inside fitExpGP /
 * fitMonoExp (FitOCT.R:110-124).  ctl: named list with dataType, Nn, gridType, rho, lambda_rate, theta0,
 * Sigma0, prior_PD, chains, nb_warmup, nb_iter, seed, adapt_delta, max_treedepth, rhat_target, max_extend. */
SEXP foct_R_sample(SEXP kind_, SEXP x, SEXP y, SEXP uy, SEXP ctl) {
  SEXP xs = PROTECT(Rf_allocVector(VECSXP, 1)), ys = PROTECT(Rf_allocVector(VECSXP, 1)), us = PROTECT(Rf_allocVector(VECSXP, 1));
  SET_VECTOR_ELT(xs, 0, x); SET_VECTOR_ELT(ys, 0, y); SET_VECTOR_ELT(us, 0, uy);
  SEXP out = foct_R_sample_batch(kind_, xs, ys, us, ctl, R_NilValue);
  UNPROTECT(3);
  return out;
}

/* .Call("foct_R_predict", kind, x, y, uy, ctl, draws) -> list(m, resid, dL), each [N, n_draws]: the generated
 * quantities rstan would have saved per draw (plotMonoExp.R:15-16, plotExpGP.R:9-11), for draws = a P_out x n_draws
 * matrix of constrained rows. */
SEXP foct_R_predict(SEXP kind_, SEXP x, SEXP y, SEXP uy, SEXP ctl, SEXP draws) {
  const int kind = Rf_asInteger(kind_);
  foct_problem P;
  fill_problem(&P, x, y, uy, ctl);
  foct_model_spec spec;
  foct_model_spec_default(&spec, kind);
  spec.theta_prior = 1; /* the generated quantities do not involve the prior: no theta0 / Sigma0 needed */
  int D = 0, P_out = 0;
  if (foct_dims(kind, P.Nn, &D, &P_out)) Rf_error("fitoct_b200: %s", foct_last_error());
  if (XLENGTH(draws) < P_out || XLENGTH(draws) % P_out) Rf_error("fitoct_b200: draws must be a %d x n matrix", P_out);
  const int nd = (int)(XLENGTH(draws) / P_out);
  SEXP m = PROTECT(Rf_allocMatrix(REALSXP, P.N, nd)), r = PROTECT(Rf_allocMatrix(REALSXP, P.N, nd));
  SEXP dl = PROTECT(Rf_allocMatrix(REALSXP, P.N, nd));
  const int rc = foct_predict(kind, &P, &spec, REAL(draws), nd, REAL(m), REAL(r), REAL(dl));
  if (rc) {
    UNPROTECT(3);
    Rf_error("fitoct_b200 error %d: %s", rc, foct_last_error());
  }
  const char* nm[] = {"m", "resid", "dL", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm));
  SET_VECTOR_ELT(out, 0, m); SET_VECTOR_ELT(out, 1, r); SET_VECTOR_ELT(out, 2, dl);
  UNPROTECT(4);
  return out;
}

/* .Call("foct_R_summary", draws, n_draws, chains) -> [n_cols, 11] matrix: rstan's summary(fit)$summary / monitor() for
 * caller-supplied draws (plotExpGP.R:9-11 and server.R:88-104 read that table), computed by the kernel that summarises
 * a fit on the device.  draws: numeric of length n_draws * chains * n_cols in the order of foct_result.draws
 * (column fastest, then chain, then draw). */
SEXP foct_R_summary(SEXP draws, SEXP n_draws_, SEXP chains_) {
  const int n = Rf_asInteger(n_draws_), c = Rf_asInteger(chains_);
  if (n < 4 || c < 1 || XLENGTH(draws) < (R_xlen_t)n * c || XLENGTH(draws) % ((R_xlen_t)n * c))
    Rf_error("fitoct_b200: draws must hold n_draws * chains * n_cols values, n_draws >= 4");
  const int p = (int)(XLENGTH(draws) / ((R_xlen_t)n * c));
  SEXP tmp = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)p * FOCT_N_SUMMARY_COLS));
  const int rc = foct_summary(REAL(draws), 1, n, c, p, REAL(tmp));
  if (rc) {
    UNPROTECT(1);
    Rf_error("fitoct_b200 error %d: %s", rc, foct_last_error());
  }
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, p, FOCT_N_SUMMARY_COLS)); /* column-major for R */
  for (int i = 0; i < p; ++i)
    for (int k = 0; k < FOCT_N_SUMMARY_COLS; ++k) REAL(out)[(R_xlen_t)k * p + i] = REAL(tmp)[(R_xlen_t)i * FOCT_N_SUMMARY_COLS + k];
  UNPROTECT(2);
  return out;
}

/* .Call("foct_R_pipeline", xs, ys, ctl, devices) — the body of FitOCT.R's dataset loop (FitOCT.R:84-124) for ALL
 * datasets in one call: estimateNoise -> fitMonoExp -> printBr gate -> estimateExpPrior -> fitExpGP(method='sample') on
 * the profiles the gate lets through.  ctl carries ctrlParams.yaml's keys (smooth_df, priorType, ru_theta, Nn, gridType,
 * rho_scale, lambda_rate, nb_warmup, nb_iter, dataType).  Returns the per-profile pieces the report scripts read. */
SEXP foct_R_pipeline(SEXP xs, SEXP ys, SEXP ctl, SEXP devices) {
  const R_xlen_t n = XLENGTH(xs);
  if (n < 1 || XLENGTH(ys) != n) Rf_error("fitoct_b200: xs, ys must be lists of the same length >= 1");
  foct_problem* P = (foct_problem*)R_alloc((size_t)n, sizeof(foct_problem));
  R_xlen_t tot = 0;
  for (R_xlen_t j = 0; j < n; ++j) {
    fill_problem(&P[j], VECTOR_ELT(xs, j), VECTOR_ELT(ys, j), R_NilValue, ctl);
    if (XLENGTH(VECTOR_ELT(ys, j)) != P[j].N) Rf_error("fitoct_b200: profile %d: x and y differ in length", (int)j + 1);
    P[j].id = (long long)get_num(ctl, "id", 0) + j;
    tot += P[j].N;
  }
  foct_pipeline_cfg pc;
  foct_pipeline_cfg_default(&pc);
  pc.smooth_df = get_num(ctl, "smooth_df", pc.smooth_df);
  pc.prior_type = (int)get_num(ctl, "priorType", pc.prior_type);
  pc.ru_theta = get_num(ctl, "ru_theta", pc.ru_theta);
  pc.Nn = (int)get_num(ctl, "Nn", pc.Nn);
  pc.gridType = (int)get_num(ctl, "gridType", pc.gridType);
  pc.rho_scale = get_num(ctl, "rho_scale", pc.rho_scale);
  pc.lambda_rate = get_num(ctl, "lambda_rate", pc.lambda_rate);
  pc.gate = (int)get_num(ctl, "gate", pc.gate);
  foct_sampler_cfg cfg;
  fill_cfg(&cfg, ctl);
  cfg.save_warmup = (int)get_num(ctl, "save_warmup", 0);
  if (devices != R_NilValue && XLENGTH(devices) > 0) { cfg.n_devices = (int)XLENGTH(devices); cfg.devices = INTEGER(devices); }
  const int D = pc.Nn + 5, P_out = pc.Nn + 7, C = cfg.chains;
  const int n_saved = cfg.save_warmup ? cfg.n_iter : cfg.n_iter - cfg.n_warmup;
  const int want_draws = (int)get_num(ctl, "draws", 0);
  int np = 0;
#define NEWR(name, len) SEXP name = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)(len))); ++np
#define NEWI(name, len) SEXP name = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)(len))); ++np
  NEWR(uy, tot); NEWR(ysm, tot); NEWR(nth, n * 2); NEWR(mth, n * 3); NEWR(mH, n * 9); NEWR(mbr, n); NEWI(mst, n);
  NEWR(ci, n * 2); NEWI(alert, n); NEWR(t0, n * 3); NEWR(S0, n * 9); NEWR(ru, n); NEWI(idx, n);
  NEWR(draws, want_draws ? (R_xlen_t)n * n_saved * C * P_out : 0);
  NEWR(sp, want_draws ? (R_xlen_t)n * n_saved * C * 6 : 0);
  NEWR(summ, n * P_out * FOCT_N_SUMMARY_COLS); NEWR(eps, n * C); NEWR(invm, n * C * D); NEWI(next, n);
  foct_pipeline_out O;
  memset(&O, 0, sizeof(O));
  O.uy = REAL(uy); O.ySmooth = REAL(ysm); O.noise_theta = REAL(nth); O.mono_theta = REAL(mth); O.mono_hessian = REAL(mH);
  O.mono_br = REAL(mbr); O.mono_status = INTEGER(mst); O.br_ci = REAL(ci); O.alert = INTEGER(alert); O.theta0 = REAL(t0);
  O.Sigma0 = REAL(S0); O.ru = REAL(ru); O.expgp_index = INTEGER(idx);
  if (want_draws) { O.expgp.draws = REAL(draws); O.expgp.sampler_params = REAL(sp); }
  O.expgp.summary = REAL(summ); O.expgp.stepsize = REAL(eps); O.expgp.inv_metric = REAL(invm); O.expgp.n_extend = INTEGER(next);
  const int rc = foct_pipeline(P, (int)n, &pc, NULL, &cfg, &O);
  if (rc) {
    UNPROTECT(np);
    Rf_error("fitoct_b200 error %d: %s", rc, foct_last_error());
  }
  SEXP nex = PROTECT(Rf_allocVector(INTSXP, 1)); ++np;
  INTEGER(nex)[0] = O.n_expgp;
  const char* nm[] = {"uy", "ySmooth", "noise_theta", "mono_theta", "mono_hessian", "mono_br", "mono_status", "br_ci", "alert",
                      "theta0", "Sigma0", "ru", "n_expgp", "expgp_index", "draws", "sampler_params", "summary", "stepsize",
                      "inv_metric", "n_extend", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, nm)); ++np;
  SEXP v[] = {uy, ysm, nth, mth, mH, mbr, mst, ci, alert, t0, S0, ru, nex, idx, draws, sp, summ, eps, invm, next};
  for (int k = 0; k < 20; ++k) SET_VECTOR_ELT(out, k, v[k]);
  UNPROTECT(np);
  return out;
#undef NEWR
#undef NEWI
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
  if (copy_prior(&P, get_elt(ctl, "theta0"), get_elt(ctl, "Sigma0"), 0)) Rf_error("fitoct_b200: theta0 needs 3 and Sigma0 9 numbers");
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
  if (copy_prior(&P, get_elt(ctl, "theta0"), get_elt(ctl, "Sigma0"), 0)) Rf_error("fitoct_b200: theta0 needs 3 and Sigma0 9 numbers");
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
    {"foct_R_sample_batch", (DL_FUNC)&foct_R_sample_batch, 6},
    {"foct_R_predict", (DL_FUNC)&foct_R_predict, 6},
    {"foct_R_summary", (DL_FUNC)&foct_R_summary, 3},
    {"foct_R_pipeline", (DL_FUNC)&foct_R_pipeline, 4},
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
