/*
 * foct_oracle_vb.c — CPU restatement of Stan's mean-field ADVI (MODEL_SPEC §14) for method = 'vb' (FitOCT.R:42).
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see foct_oracle.h).  Included at the end of foct_oracle.c so that it
 * shares the model and RNG helpers.
 */

enum { SITE_VB_GRAD = 5, SITE_VB_ELBO = 6, SITE_VB_OUT = 7 };

typedef double (*vb_lpg_fn)(const void* ctx, const double* q, double* g, double* chi2);

static double vb_model_lpg(const void* ctx, const double* q, double* g, double* chi2) {
  return model_lpg((const model_t*)ctx, q, g, chi2, NULL);
}
static double vb_analytic_lpg(const void* ctx, const double* q, double* g, double* chi2) {
  if (chi2) *chi2 = 0.0;
  return analytic_lpg((const analytic_t*)ctx, q, g);
}

typedef struct {
  vb_lpg_fn f; const void* ctx; int D; rng_t rng; const foct_vb_cfg* cfg;
} vb_t;

static void vb_draw(const vb_t* V, const double* mu, const double* om, uint32_t it, uint32_t kind, uint32_t a,
                    uint32_t phase, double* eta, double* zeta) {
  for (int d = 0; d < V->D; ++d) {
    uint32_t r[4];
    rng_block(&V->rng, it, kind, a, (uint32_t)d, phase, r);
    eta[d] = foct_oracle_normal(r);
    zeta[d] = mu[d] + exp(om[d]) * eta[d];
  }
}

/* returns 0 ok, 1 non-finite */
static int vb_grad(const vb_t* V, const double* mu, const double* om, uint32_t it, uint32_t phase, double* gmu, double* gom) {
  const int D = V->D, S = V->cfg->grad_samples;
  for (int d = 0; d < D; ++d) gmu[d] = gom[d] = 0.0;
  int bad = 0;
  for (int s = 0; s < S; ++s) {
    double eta[MAXD], zeta[MAXD], g[MAXD];
    vb_draw(V, mu, om, it, SITE_VB_GRAD, (uint32_t)s, phase, eta, zeta);
    const double lp = V->f(V->ctx, zeta, g, NULL);
    if (!isfinite(lp)) bad = 1;
    for (int d = 0; d < D; ++d) { if (!isfinite(g[d])) bad = 1; gmu[d] += g[d]; gom[d] += g[d] * eta[d]; }
  }
  for (int d = 0; d < D; ++d) { gmu[d] /= S; gom[d] /= S; gom[d] = gom[d] * exp(om[d]) + 1.0; }
  return bad;
}

/* returns NAN on failure (too many dropped evaluations) */
static double vb_elbo(const vb_t* V, const double* mu, const double* om, uint32_t it, uint32_t phase) {
  const int D = V->D, n = V->cfg->elbo_samples;
  double sum = 0.0;
  int ok = 0, dropped = 0;
  for (uint32_t a = 0; ok < n; ++a) {
    double eta[MAXD], zeta[MAXD], g[MAXD];
    vb_draw(V, mu, om, it, SITE_VB_ELBO, a, phase, eta, zeta);
    const double lp = V->f(V->ctx, zeta, g, NULL);
    if (isfinite(lp)) { sum += lp; ++ok; }
    else if (++dropped >= n) return NAN;
  }
  double ent = 0.5 * D * (1.0 + log(6.283185307179586476925286766559));
  for (int d = 0; d < D; ++d) ent += om[d];
  return sum / n + ent;
}

static void vb_step(int D, double* mu, double* om, double* hmu, double* hom, const double* gmu, const double* gom, int k,
                    double eta_s) {
  const double sc = eta_s / sqrt((double)k);
  for (int d = 0; d < D; ++d) {
    hmu[d] = k == 1 ? gmu[d] * gmu[d] : 0.9 * hmu[d] + 0.1 * gmu[d] * gmu[d];
    hom[d] = k == 1 ? gom[d] * gom[d] : 0.9 * hom[d] + 0.1 * gom[d] * gom[d];
    mu[d] += sc * gmu[d] / (1.0 + sqrt(hmu[d]));
    om[d] += sc * gom[d] / (1.0 + sqrt(hom[d]));
  }
}

static int cmp_dbl_vb(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }

/* The ADVI driver on one target.  q0: start.  Outputs mu, om, elbo, eta, iters; returns status. */
static int vb_run(const vb_t* V, const double* q0, double* mu, double* om, double* elbo_out, double* eta_out, int* iters_out) {
  const foct_vb_cfg* c = V->cfg;
  const int D = V->D;
  double hmu[MAXD], hom[MAXD], gmu[MAXD], gom[MAXD];
  double eta_s = c->eta;
  *elbo_out = NAN; *iters_out = 0;
  if (c->adapt_engaged) {
    static const double seq[5] = {100.0, 10.0, 1.0, 0.1, 0.01};
    for (int d = 0; d < D; ++d) { mu[d] = q0[d]; om[d] = c->omega0; }
    const double elbo_init = vb_elbo(V, mu, om, 0, 1);
    if (isnan(elbo_init)) { *eta_out = NAN; return 2; }
    double elbo_best = -INFINITY, eta_best = 0.0;
    int found = 0;
    for (int e = 0; e < 5; ++e) {
      for (int d = 0; d < D; ++d) { mu[d] = q0[d]; om[d] = c->omega0; hmu[d] = hom[d] = 0.0; }
      for (int k = 1; k <= c->adapt_iter; ++k) {
        if (vb_grad(V, mu, om, (uint32_t)k, (uint32_t)(1 + e), gmu, gom)) for (int d = 0; d < D; ++d) gmu[d] = gom[d] = 0.0;
        vb_step(D, mu, om, hmu, hom, gmu, gom, k, seq[e]);
      }
      double elbo = vb_elbo(V, mu, om, (uint32_t)c->adapt_iter, (uint32_t)(1 + e));
      if (isnan(elbo)) elbo = -INFINITY;
      if (elbo < elbo_best && elbo_best > elbo_init) { found = 1; break; }
      if (e < 4) { elbo_best = elbo; eta_best = seq[e]; }
      else if (elbo > elbo_init) { eta_best = seq[e]; found = 1; }
    }
    if (!found) { *eta_out = NAN; return 2; }
    eta_s = eta_best;
  }
  *eta_out = eta_s;
  for (int d = 0; d < D; ++d) { mu[d] = q0[d]; om[d] = c->omega0; hmu[d] = hom[d] = 0.0; }
  int cap = (int)(0.1 * c->iter / c->eval_elbo);
  if (cap < 2) cap = 2;
  if (cap > 32) cap = 32;
  double cb[32]; int cbn = 0, cbpos = 0;
  double elbo = 0.0, elbo_prev;
  int status = 1, k;
  for (k = 1; k <= c->iter; ++k) {
    if (vb_grad(V, mu, om, (uint32_t)k, 0, gmu, gom)) { status = 2; break; }
    vb_step(D, mu, om, hmu, hom, gmu, gom, k, eta_s);
    if (k % c->eval_elbo == 0) {
      elbo_prev = elbo;
      elbo = vb_elbo(V, mu, om, (uint32_t)k, 0);
      if (isnan(elbo)) { status = 2; break; }
      const double delta = fabs((elbo_prev - elbo) / elbo);
      cb[cbpos] = delta; cbpos = (cbpos + 1) % cap; if (cbn < cap) ++cbn;
      double tmp[32], mean = 0.0;
      for (int i = 0; i < cbn; ++i) { tmp[i] = cb[i]; mean += cb[i]; }
      mean /= cbn;
      qsort(tmp, cbn, sizeof(double), cmp_dbl_vb);
      const double med = tmp[cbn / 2];
      if (mean < c->tol_rel_obj || med < c->tol_rel_obj) { status = 0; break; }
    }
  }
  *iters_out = k > c->iter ? c->iter : k;
  *elbo_out = elbo;
  return status;
}

static void vb_row(const model_t* M, const double* q, double* row) {
  const int Nn = M->Nn;
  double g[MAXD], chi2 = 0.0;
  model_lpg(M, q, g, &chi2, NULL);
  if (M->kind == FOCT_EXPGP) {
    for (int d = 0; d < 3 + Nn; ++d) row[d] = q[d];
    row[3 + Nn] = exp(q[3 + Nn]); row[4 + Nn] = exp(q[4 + Nn]);
    row[5 + Nn] = M->prior_PD ? NAN : chi2 / br_ndf(M);
    row[6 + Nn] = 0.0;
  } else {
    row[0] = q[0]; row[1] = q[1]; row[2] = q[2];
    row[3] = M->prior_PD ? NAN : chi2 / br_ndf(M);
    row[4] = 0.0;
  }
}

int foct_oracle_vb(int kind, const foct_problem* P, int n_problems, const foct_model_spec* spec, const foct_vb_cfg* cfg,
                   foct_vb_result* R) {
  for (int j = 0; j < n_problems; ++j) {
    model_t M;
    int rc = model_init(&M, kind, &P[j], spec, NULL);
    if (rc) return rc;
    const int D = M.D, Nn = M.Nn, P_out = kind == FOCT_EXPGP ? Nn + 7 : 5;
    vb_t V; V.f = vb_model_lpg; V.ctx = &M; V.D = D; V.cfg = cfg;
    rng_seed(&V.rng, cfg->seed, P[j].id, 0);
    double q0[MAXD], mu[MAXD], om[MAXD], elbo, eta; int iters;
    for (int d = 0; d < D; ++d) {
      uint32_t r[4]; double u[2];
      rng_block(&V.rng, 0, SITE_INIT, 0, (uint32_t)d, 0, r);
      if (cfg->init_mode == 2 && cfg->init) q0[d] = cfg->init[(size_t)j * D + d];
      else if (cfg->init_mode == 1) { foct_oracle_uniform2(r, u); q0[d] = -2.0 + 4.0 * u[0]; }
      else if (kind == FOCT_EXPGP) {
        if (d < 3) q0[d] = P[j].theta0[d];
        else if (d < 3 + Nn) q0[d] = 0.01 * foct_oracle_normal(r);
        else if (d == 3 + Nn) q0[d] = log(0.1);
        else q0[d] = 0.0;
      } else q0[d] = P[j].theta0[d];
    }
    const int st = vb_run(&V, q0, mu, om, &elbo, &eta, &iters);
    for (int d = 0; d < D; ++d) { R->mu[(size_t)j * D + d] = mu[d]; R->omega[(size_t)j * D + d] = om[d]; }
    if (R->elbo) R->elbo[j] = elbo;
    if (R->eta) R->eta[j] = eta;
    if (R->iters) R->iters[j] = iters;
    if (R->status) R->status[j] = st;
    vb_row(&M, mu, R->mean + (size_t)j * P_out);
    if (R->draws)
      for (int i = 0; i < cfg->output_samples; ++i) {
        double eta_v[MAXD], zeta[MAXD];
        vb_draw(&V, mu, om, 0, SITE_VB_OUT, (uint32_t)i, 0, eta_v, zeta);
        vb_row(&M, zeta, R->draws + ((size_t)j * cfg->output_samples + i) * P_out);
      }
    model_free(&M);
  }
  return 0;
}

/* Known-answer harness: ADVI on the analytic targets of foct_oracle_sample_analytic (0 iid normal, 2 dense-precision MVN).
 * q0 [D]; outputs mu, omega [D], info[3] = elbo, eta, iters; returns status. */
int foct_oracle_vb_analytic(int target, int D, const double* par, const foct_vb_cfg* cfg, const double* q0, double* mu,
                            double* omega, double* info) {
  if (D < 1 || D > MAXD) return FOCT_EINVAL;
  analytic_t A = {target, D, par};
  vb_t V; V.f = vb_analytic_lpg; V.ctx = &A; V.D = D; V.cfg = cfg;
  rng_seed(&V.rng, cfg->seed, 0, 0);
  double elbo, eta; int iters;
  const int st = vb_run(&V, q0, mu, omega, &elbo, &eta, &iters);
  if (info) { info[0] = elbo; info[1] = eta; info[2] = (double)iters; }
  return st;
}
