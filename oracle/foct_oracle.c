/*
 * foct_oracle.c — CPU restatement of the FitOCT hot path (model blocks + NUTS).  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED (see foct_oracle.h).  What each part follows:
 *   - model, gradient, generated quantities: MODEL_SPEC.md §1-6, whose evidence is
 *       mean function      /root/reference/synthData.R:22, ShinyInterface/ui.R:88, server.R:372
 *       control grid       /root/reference/ShinyInterface/server.R:626-635
 *       argument list      /root/reference/FitOCT.R:110-124
 *       parameter names    /root/reference/plotExpGP.R:9
 *   - sampler: Stan's published algorithm (Stan Reference Manual "MCMC Sampling"; Hoffman & Gelman 2014;
 *     Betancourt 2017), third-party and not under /root/reference: rstan/StanHeaders, version unpinned by
 *     the reference (FitOCT.R:4-10).  Restated here in recursive form, the way the manual describes it.
 *   - known-answer anchor available in the reference: Tests/testGamma.R:19-47 (Exponential(1/10) on a
 *     lower=0 parameter) -> foct_oracle_sample_analytic(target=1).
 *
 * Plain C99, sequential per (profile, chain); OpenMP only fans chains out over cores.
 */
#include "foct_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD FOCT_MAX_D

/* ------------------------------------------------------------------ RNG (MODEL_SPEC §7) */

void foct_oracle_philox(const uint32_t c_in[4], const uint32_t k_in[2], uint32_t out[4]) {
  uint32_t c0 = c_in[0], c1 = c_in[1], c2 = c_in[2], c3 = c_in[3];
  uint32_t k0 = k_in[0], k1 = k_in[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void foct_oracle_uniform2(const uint32_t r[4], double u[2]) {
  /* 53-bit uniforms in (0,1): never 0, never 1 */
  uint64_t a = (((uint64_t)r[0] << 32) | r[1]) >> 11;
  uint64_t b = (((uint64_t)r[2] << 32) | r[3]) >> 11;
  u[0] = ((double)a + 0.5) * 0x1.0p-53;
  u[1] = ((double)b + 0.5) * 0x1.0p-53;
}

double foct_oracle_normal(const uint32_t r[4]) {
  double u[2];
  foct_oracle_uniform2(r, u);
  return sqrt(-2.0 * log(u[0])) * cos(6.283185307179586476925286766559 * u[1]);
}

enum { SITE_MOM = 0, SITE_DIR = 1, SITE_MERGE = 2, SITE_INITEPS = 3, SITE_INIT = 4 };

typedef struct { uint32_t key[2]; } rng_t;

static void rng_seed(rng_t* g, unsigned long long seed, long long id, int chain) {
  unsigned long long stream = (unsigned long long)id * 64ull + (unsigned long long)chain;
  g->key[0] = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x85EBCA6Bu) ^ (uint32_t)(stream >> 32);
  g->key[1] = (uint32_t)stream;
}
static void rng_block(const rng_t* g, uint32_t it, uint32_t kind, uint32_t a, uint32_t b, uint32_t lvl,
                      uint32_t out[4]) {
  uint32_t c[4] = {it, kind | (a << 8), b, lvl};
  foct_oracle_philox(c, g->key, out);
}

/* ------------------------------------------------------------------ grid and basis (MODEL_SPEC §1) */

int foct_oracle_grid(int Nn, int gridType, double* xGP) {
  if (Nn < 1 || Nn > FOCT_MAX_NN) return FOCT_EINVAL;
  double dx = 1.0 / (Nn + 1);
  double lo = gridType == FOCT_GRID_INTERNAL ? 0.5 * dx : 0.0;
  double hi = gridType == FOCT_GRID_INTERNAL ? 1.0 - 0.5 * dx : 1.0;
  for (int k = 0; k < Nn; ++k) xGP[k] = Nn == 1 ? lo : lo + (hi - lo) * (double)k / (double)(Nn - 1);
  return 0;
}

static double kern(double a, double b, double rho, int kernel) {
  double d = a - b;
  double den = kernel == 0 ? 2.0 * rho * rho : rho * rho;
  return exp(-(d * d) / den);
}

int foct_oracle_basis(const foct_problem* P, const foct_model_spec* spec, double* B) {
  int N = P->N, Nn = P->Nn;
  if (Nn < 1 || Nn > FOCT_MAX_NN || N < 2) return FOCT_EINVAL;
  double xg[FOCT_MAX_NN], L[FOCT_MAX_NN][FOCT_MAX_NN];
  foct_oracle_grid(Nn, P->gridType, xg);
  double xmin = P->x[0], xmax = P->x[0];
  for (int i = 1; i < N; ++i) {
    if (P->x[i] < xmin) xmin = P->x[i];
    if (P->x[i] > xmax) xmax = P->x[i];
  }
  /* Cholesky of Kgg + jitter I (lower), row by row */
  for (int i = 0; i < Nn; ++i) {
    for (int j = 0; j <= i; ++j) {
      double s = kern(xg[i], xg[j], P->rho, spec->kernel) + (i == j ? spec->jitter : 0.0);
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      if (i == j) {
        if (!(s > 0.0)) return FOCT_EINVAL;
        L[i][i] = sqrt(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  }
  for (int i = 0; i < N; ++i) {
    double xp = (P->x[i] - xmin) / (xmax - xmin);
    double v[FOCT_MAX_NN];
    for (int k = 0; k < Nn; ++k) v[k] = kern(xp, xg[k], P->rho, spec->kernel);
    for (int k = 0; k < Nn; ++k) { /* L w = v */
      double s = v[k];
      for (int j = 0; j < k; ++j) s -= L[k][j] * v[j];
      v[k] = s / L[k][k];
    }
    for (int k = Nn - 1; k >= 0; --k) { /* L' b = w */
      double s = v[k];
      for (int j = k + 1; j < Nn; ++j) s -= L[j][k] * v[j];
      v[k] = s / L[k][k];
    }
    for (int k = 0; k < Nn; ++k) B[(size_t)k * N + i] = v[k];
  }
  return 0;
}

/* ------------------------------------------------------------------ model (MODEL_SPEC §2-6) */

typedef struct {
  int kind, N, Nn, D, prior_PD;
  double c;
  const double *x, *y, *uy, *B;
  double theta0[3], Pinv[9]; /* Sigma0^-1 */
  double lambda_rate, sum_log_uy;
  foct_model_spec spec;
  double* B_owned;
} model_t;

static int inv3(const double* S, double* Pi) {
  double a = S[0], b = S[1], c = S[2], d = S[3], e = S[4], f = S[5], g = S[6], h = S[7], i = S[8];
  double A = e * i - f * h, Bc = -(d * i - f * g), C = d * h - e * g;
  double det = a * A + b * Bc + c * C;
  if (!(fabs(det) > 0.0)) return FOCT_EINVAL;
  double id = 1.0 / det;
  Pi[0] = A * id;  Pi[1] = -(b * i - c * h) * id; Pi[2] = (b * f - c * e) * id;
  Pi[3] = Bc * id; Pi[4] = (a * i - c * g) * id;  Pi[5] = -(a * f - c * d) * id;
  Pi[6] = C * id;  Pi[7] = -(a * h - b * g) * id; Pi[8] = (a * e - b * d) * id;
  return 0;
}

static int model_init(model_t* M, int kind, const foct_problem* P, const foct_model_spec* spec,
                      const double* B) {
  memset(M, 0, sizeof(*M));
  M->kind = kind; M->N = P->N; M->Nn = kind == FOCT_EXPGP ? P->Nn : 0;
  M->D = kind == FOCT_EXPGP ? M->Nn + 5 : 3;
  M->prior_PD = P->prior_PD; M->c = (double)P->dataType;
  M->x = P->x; M->y = P->y; M->uy = P->uy;
  M->spec = *spec; M->lambda_rate = P->lambda_rate;
  memcpy(M->theta0, P->theta0, sizeof(M->theta0));
  if (P->N < 2 || (kind == FOCT_EXPGP && (P->Nn < 1 || P->Nn > FOCT_MAX_NN))) return FOCT_EINVAL;
  if (spec->theta_prior == 0 && inv3(P->Sigma0, M->Pinv)) return FOCT_EINVAL;
  double s = 0.0;
  for (int i = 0; i < P->N; ++i) s += log(P->uy[i]);
  M->sum_log_uy = s;
  if (kind == FOCT_EXPGP) {
    if (B) {
      M->B = B;
    } else {
      M->B_owned = (double*)malloc(sizeof(double) * (size_t)P->N * P->Nn);
      if (!M->B_owned) return FOCT_ENOMEM;
      int rc = foct_oracle_basis(P, spec, M->B_owned);
      if (rc) { free(M->B_owned); M->B_owned = NULL; return rc; }
      M->B = M->B_owned;
    }
  }
  return 0;
}
static void model_free(model_t* M) { free(M->B_owned); M->B_owned = NULL; }

/* lp, gradient, chi2 = sum((y-m)/uy)^2.  abs_t (optional) accumulates |summand| per component. */
static double model_lpg(const model_t* M, const double* q, double* g, double* chi2_out, double* abs_t) {
  const int N = M->N, Nn = M->Nn, D = M->D;
  const int gp = M->kind == FOCT_EXPGP;
  const double th1 = q[0], th2 = q[1], th3 = q[2];
  const double* yg = q + 3;
  const double lam = gp ? exp(q[3 + Nn]) : 1.0;
  const double sig = gp ? exp(q[4 + Nn]) : 1.0;
  double lp = 0.0, chi2 = 0.0;
  for (int d = 0; d < D; ++d) g[d] = 0.0;
  if (abs_t) for (int d = 0; d < D; ++d) abs_t[d] = 0.0;

  if (!M->prior_PD) {
    double szz = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0, gy[FOCT_MAX_NN];
    for (int k = 0; k < Nn; ++k) gy[k] = 0.0;
    for (int i = 0; i < N; ++i) {
      double dl = 0.0;
      for (int k = 0; k < Nn; ++k) dl += M->B[(size_t)k * N + i] * yg[k];
      double s = 1.0 + dl;
      double w = 1.0 / (sig * M->uy[i]);
      double m, e, t;
      if (M->spec.modulation == 0) {
        double r = 1.0 / (th3 * s);
        t = M->c * M->x[i] * r;
        e = exp(-t);
        m = th1 + th2 * e;
        double z = (M->y[i] - m) * w, gi = z * w;
        double qq = gi * th2 * e * t * r;
        szz += z * z; g1 += gi; g2 += gi * e; g3 += qq * s;
        for (int k = 0; k < Nn; ++k) gy[k] += qq * M->B[(size_t)k * N + i];
        if (abs_t) {
          abs_t[0] += fabs(gi); abs_t[1] += fabs(gi * e); abs_t[2] += fabs(qq * s);
          for (int k = 0; k < Nn; ++k) abs_t[3 + k] += fabs(th3 * qq * M->B[(size_t)k * N + i]);
          if (gp) abs_t[4 + Nn] += z * z;
        }
      } else {
        t = M->c * M->x[i] / th3;
        e = exp(-t);
        m = th1 + th2 * e * s;
        double z = (M->y[i] - m) * w, gi = z * w;
        szz += z * z; g1 += gi; g2 += gi * e * s; g3 += gi * e * s * t;
        for (int k = 0; k < Nn; ++k) gy[k] += gi * e * M->B[(size_t)k * N + i];
        if (abs_t) {
          abs_t[0] += fabs(gi); abs_t[1] += fabs(gi * e * s); abs_t[2] += fabs(gi * e * s * t * th2 / th3);
          for (int k = 0; k < Nn; ++k) abs_t[3 + k] += fabs(th2 * gi * e * M->B[(size_t)k * N + i]);
          if (gp) abs_t[4 + Nn] += z * z;
        }
      }
      double ru = (M->y[i] - m) / M->uy[i];
      chi2 += ru * ru;
    }
    lp += -0.5 * szz - (double)N * log(sig) - M->sum_log_uy;
    g[0] += g1; g[1] += g2;
    if (M->spec.modulation == 0) {
      g[2] += g3;
      for (int k = 0; k < Nn; ++k) g[3 + k] += th3 * gy[k];
    } else {
      g[2] += g3 * th2 / th3;
      for (int k = 0; k < Nn; ++k) g[3 + k] += th2 * gy[k];
    }
    if (gp) g[4 + Nn] += szz - (double)N;
  } else {
    chi2 = NAN;
  }

  if (M->spec.theta_prior == 0) {
    double d0 = th1 - M->theta0[0], d1 = th2 - M->theta0[1], d2 = th3 - M->theta0[2];
    const double* Pi = M->Pinv;
    double v0 = Pi[0] * d0 + Pi[1] * d1 + Pi[2] * d2;
    double v1 = Pi[3] * d0 + Pi[4] * d1 + Pi[5] * d2;
    double v2 = Pi[6] * d0 + Pi[7] * d1 + Pi[8] * d2;
    lp += -0.5 * (d0 * v0 + d1 * v1 + d2 * v2);
    g[0] -= v0; g[1] -= v1; g[2] -= v2;
    if (abs_t) { abs_t[0] += fabs(v0); abs_t[1] += fabs(v1); abs_t[2] += fabs(v2); }
  }
  if (gp) {
    double sy = 0.0;
    if (M->spec.ygp_prior == 0) {
      double il2 = 1.0 / (lam * lam);
      for (int k = 0; k < Nn; ++k) { sy += yg[k] * yg[k]; g[3 + k] -= yg[k] * il2; if (abs_t) abs_t[3 + k] += fabs(yg[k] * il2); }
      lp += -(double)Nn * q[3 + Nn] - 0.5 * sy * il2;
      g[3 + Nn] += sy * il2 - (double)Nn;
      if (abs_t) abs_t[3 + Nn] += sy * il2 + (double)Nn;
    } else {
      double il = 1.0 / lam;
      for (int k = 0; k < Nn; ++k) {
        sy += fabs(yg[k]);
        g[3 + k] -= (yg[k] > 0.0 ? 1.0 : (yg[k] < 0.0 ? -1.0 : 0.0)) * il;
        if (abs_t) abs_t[3 + k] += il;
      }
      lp += -(double)Nn * q[3 + Nn] - sy * il;
      g[3 + Nn] += sy * il - (double)Nn;
      if (abs_t) abs_t[3 + Nn] += sy * il + (double)Nn;
    }
    if (M->spec.lambda_prior == 0) { lp += q[3 + Nn] - M->lambda_rate * lam; g[3 + Nn] += 1.0 - M->lambda_rate * lam; }
    else { lp += -M->lambda_rate * lam; g[3 + Nn] += -M->lambda_rate * lam; }
    if (abs_t) abs_t[3 + Nn] += 2.0 + M->lambda_rate * lam;
    if (M->spec.sigma_sd > 0.0) {
      double u = (sig - M->spec.sigma_mean) / M->spec.sigma_sd;
      lp += -0.5 * u * u;
      g[4 + Nn] += -sig * u / M->spec.sigma_sd;
      if (abs_t) abs_t[4 + Nn] += fabs(sig * u / M->spec.sigma_sd);
    }
    lp += q[3 + Nn] + q[4 + Nn];
    g[3 + Nn] += 1.0; g[4 + Nn] += 1.0;
    if (abs_t) abs_t[4 + Nn] += (double)M->N + 2.0;
  }
  if (chi2_out) *chi2_out = chi2;
  return lp;
}

int foct_oracle_logp_grad(int kind, const foct_problem* P, const foct_model_spec* spec, const double* B,
                          const double* q, int n_q, double* lp, double* grad, double* chi2,
                          double* abs_terms) {
  model_t M;
  int rc = model_init(&M, kind, P, spec, B);
  if (rc) return rc;
  for (int j = 0; j < n_q; ++j) {
    double c2;
    lp[j] = model_lpg(&M, q + (size_t)j * M.D, grad + (size_t)j * M.D, &c2,
                      abs_terms ? abs_terms + (size_t)j * M.D : NULL);
    if (chi2) chi2[j] = c2;
  }
  model_free(&M);
  return 0;
}

static double br_ndf(const model_t* M) {
  if (M->spec.br_ndf == 1) return (double)M->N;
  return (double)(M->N - 3 - M->Nn);
}

int foct_oracle_predict(int kind, const foct_problem* P, const foct_model_spec* spec, const double* draws,
                        int n_draws, double* m_out, double* resid, double* dL) {
  model_t M;
  int rc = model_init(&M, kind, P, spec, NULL);
  if (rc) return rc;
  int P_out = kind == FOCT_EXPGP ? M.Nn + 7 : 5;
  for (int j = 0; j < n_draws; ++j) {
    const double* r = draws + (size_t)j * P_out;
    for (int i = 0; i < M.N; ++i) {
      double dl = 0.0;
      for (int k = 0; k < M.Nn; ++k) dl += M.B[(size_t)k * M.N + i] * r[3 + k];
      double m;
      if (spec->modulation == 0) m = r[0] + r[1] * exp(-M.c * M.x[i] / (r[2] * (1.0 + dl)));
      else m = r[0] + r[1] * exp(-M.c * M.x[i] / r[2]) * (1.0 + dl);
      if (m_out) m_out[(size_t)j * M.N + i] = m;
      if (resid) resid[(size_t)j * M.N + i] = M.y[i] - m;
      if (dL) dL[(size_t)j * M.N + i] = dl;
    }
  }
  model_free(&M);
  return 0;
}

/* ------------------------------------------------------------------ analytic targets */

typedef struct { int target, D; const double* par; } analytic_t;

static double analytic_lpg(const analytic_t* A, const double* q, double* g) {
  if (A->target == 0) { /* independent normal(0, sd_d) */
    double lp = 0.0;
    for (int d = 0; d < A->D; ++d) {
      double s = A->par[d];
      lp += -0.5 * q[d] * q[d] / (s * s);
      g[d] = -q[d] / (s * s);
    }
    return lp;
  }
  if (A->target == 2) { /* zero-mean multivariate normal with dense precision matrix par[D*D] */
    double lp = 0.0;
    for (int a = 0; a < A->D; ++a) {
      double s = 0.0;
      for (int b = 0; b < A->D; ++b) s += A->par[a * A->D + b] * q[b];
      g[a] = -s;
      lp += -0.5 * q[a] * s;
    }
    return lp;
  }
  /* Exponential(rate) on lambda = exp(q): lp = -rate*lambda + q  (Tests/testGamma.R:19-30) */
  double lam = exp(q[0]);
  g[0] = -A->par[0] * lam + 1.0;
  return -A->par[0] * lam + q[0];
}

/* ------------------------------------------------------------------ NUTS (MODEL_SPEC §7) */

typedef struct { double q[MAXD], p[MAXD], g[MAXD]; double V, chi2, H; } pspoint;

typedef struct {
  const model_t* M; const analytic_t* A; int D;
  rng_t rng; uint32_t it;
  double eps, invM[MAXD];
  int max_depth; double max_dH;
  pspoint z;       /* integrator state */
  int divergent, depth, n_leapfrog; double energy;
} nuts_t;

static void eval_potential(nuts_t* S, pspoint* z) {
  double c2 = NAN, lp;
  if (S->M) lp = model_lpg(S->M, z->q, z->g, &c2, NULL);
  else lp = analytic_lpg(S->A, z->q, z->g);
  z->V = -lp; z->chi2 = c2;
}
static double kinetic(const nuts_t* S, const double* p) {
  double k = 0.0;
  for (int d = 0; d < S->D; ++d) k += S->invM[d] * p[d] * p[d];
  return 0.5 * k;
}
static double hamiltonian(const nuts_t* S, const pspoint* z) { return z->V + kinetic(S, z->p); }

static void leapfrog(nuts_t* S, pspoint* z, double eps) {
  int D = S->D;
  for (int d = 0; d < D; ++d) z->p[d] += 0.5 * eps * z->g[d];       /* dp/dt = -dV/dq = +grad lp */
  for (int d = 0; d < D; ++d) z->q[d] += eps * S->invM[d] * z->p[d];
  eval_potential(S, z);
  for (int d = 0; d < D; ++d) z->p[d] += 0.5 * eps * z->g[d];
}

static double log_sum_exp2(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  double mx = a > b ? a : b, mn = a > b ? b : a;
  return mx + log1p(exp(mn - mx));
}

static int criterion(const nuts_t* S, const double* ps_minus, const double* ps_plus, const double* rho) {
  double a = 0.0, b = 0.0;
  for (int d = 0; d < S->D; ++d) { a += ps_plus[d] * rho[d]; b += ps_minus[d] * rho[d]; }
  return a > 0.0 && b > 0.0;
}
static void sharp(const nuts_t* S, const double* p, double* ps) {
  for (int d = 0; d < S->D; ++d) ps[d] = S->invM[d] * p[d];
}

static int build_tree(nuts_t* S, int depth, pspoint* z_propose, double* ps_beg, double* ps_end, double* rho,
                      double* p_beg, double* p_end, double H0, int sign, int* n_leapfrog,
                      double* log_sum_weight, double* sum_metro_prob, uint32_t leaf_base, int top_depth) {
  const int D = S->D;
  if (depth == 0) {
    leapfrog(S, &S->z, sign * S->eps);
    ++(*n_leapfrog);
    double h = hamiltonian(S, &S->z);
    if (isnan(h)) h = INFINITY;
    S->z.H = h;
    if (h - H0 > S->max_dH) S->divergent = 1;
    *log_sum_weight = log_sum_exp2(*log_sum_weight, H0 - h);
    *sum_metro_prob += (H0 - h > 0.0) ? 1.0 : exp(H0 - h);
    *z_propose = S->z;
    sharp(S, S->z.p, ps_beg);
    memcpy(ps_end, ps_beg, sizeof(double) * D);
    for (int d = 0; d < D; ++d) rho[d] += S->z.p[d];
    memcpy(p_beg, S->z.p, sizeof(double) * D);
    memcpy(p_end, S->z.p, sizeof(double) * D);
    return !S->divergent;
  }
  double lsw_init = -INFINITY, p_init_end[MAXD], ps_init_end[MAXD], rho_init[MAXD];
  memset(rho_init, 0, sizeof(rho_init));
  int valid_init = build_tree(S, depth - 1, z_propose, ps_beg, ps_init_end, rho_init, p_beg, p_init_end, H0,
                              sign, n_leapfrog, &lsw_init, sum_metro_prob, leaf_base, top_depth);
  if (!valid_init) return 0;

  pspoint z_propose_final = S->z;
  double lsw_final = -INFINITY, p_final_beg[MAXD], ps_final_beg[MAXD], rho_final[MAXD];
  memset(rho_final, 0, sizeof(rho_final));
  int valid_final = build_tree(S, depth - 1, &z_propose_final, ps_final_beg, ps_end, rho_final, p_final_beg,
                               p_end, H0, sign, n_leapfrog, &lsw_final, sum_metro_prob,
                               leaf_base + (1u << (depth - 1)), top_depth);
  if (!valid_final) return 0;

  double lsw_subtree = log_sum_exp2(lsw_init, lsw_final);
  *log_sum_weight = log_sum_exp2(*log_sum_weight, lsw_subtree);
  if (lsw_final > lsw_subtree) {
    *z_propose = z_propose_final;
  } else {
    /* merge draw site: one Philox block per (leaf, group of four levels); 32-bit uniform (MODEL_SPEC §7) */
    uint32_t r[4];
    rng_block(&S->rng, S->it, SITE_MERGE, (uint32_t)top_depth, leaf_base + (1u << depth) - 1u, (uint32_t)((depth - 1) >> 2), r);
    double u = ((double)r[(depth - 1) & 3] + 0.5) * 0x1.0p-32;
    if (u < exp(lsw_final - lsw_subtree)) *z_propose = z_propose_final;
  }
  double rho_sub[MAXD], rho_ext[MAXD];
  for (int d = 0; d < D; ++d) { rho_sub[d] = rho_init[d] + rho_final[d]; rho[d] += rho_sub[d]; }
  int persist = criterion(S, ps_beg, ps_end, rho_sub);
  for (int d = 0; d < D; ++d) rho_ext[d] = rho_init[d] + p_final_beg[d];
  persist &= criterion(S, ps_beg, ps_final_beg, rho_ext);
  for (int d = 0; d < D; ++d) rho_ext[d] = rho_final[d] + p_init_end[d];
  persist &= criterion(S, ps_init_end, ps_end, rho_ext);
  return persist;
}

/* One NUTS transition from (q, V, g) held in *cur (momentum resampled).  Returns accept_stat. */
static double nuts_transition(nuts_t* S, pspoint* cur) {
  const int D = S->D;
  for (int d = 0; d < D; ++d) {
    uint32_t r[4];
    rng_block(&S->rng, S->it, SITE_MOM, 0, (uint32_t)d, 0, r);
    cur->p[d] = foct_oracle_normal(r) / sqrt(S->invM[d]);
  }
  S->z = *cur;
  pspoint z_fwd = S->z, z_bck = S->z, z_sample = S->z, z_propose = S->z;
  double p_fwd_fwd[MAXD], ps_fwd_fwd[MAXD], p_fwd_bck[MAXD], ps_fwd_bck[MAXD];
  double p_bck_fwd[MAXD], ps_bck_fwd[MAXD], p_bck_bck[MAXD], ps_bck_bck[MAXD], rho[MAXD];
  size_t nb = sizeof(double) * D;
  memcpy(p_fwd_fwd, S->z.p, nb); sharp(S, S->z.p, ps_fwd_fwd);
  memcpy(p_fwd_bck, p_fwd_fwd, nb); memcpy(ps_fwd_bck, ps_fwd_fwd, nb);
  memcpy(p_bck_fwd, p_fwd_fwd, nb); memcpy(ps_bck_fwd, ps_fwd_fwd, nb);
  memcpy(p_bck_bck, p_fwd_fwd, nb); memcpy(ps_bck_bck, ps_fwd_fwd, nb);
  memcpy(rho, S->z.p, nb);
  double log_sum_weight = 0.0;
  double H0 = hamiltonian(S, &S->z);
  z_sample.H = H0;
  int n_leapfrog = 0;
  double sum_metro_prob = 0.0;
  S->depth = 0; S->divergent = 0;

  while (S->depth < S->max_depth) {
    double rho_fwd[MAXD], rho_bck[MAXD];
    memset(rho_fwd, 0, sizeof(rho_fwd)); memset(rho_bck, 0, sizeof(rho_bck));
    int valid_subtree;
    double lsw_subtree = -INFINITY;
    uint32_t r[4]; double u[2];
    rng_block(&S->rng, S->it, SITE_DIR, (uint32_t)S->depth, 0, 0, r);
    foct_oracle_uniform2(r, u);
    if (u[0] > 0.5) {
      S->z = z_fwd;
      memcpy(rho_bck, rho, nb);
      memcpy(p_bck_fwd, p_fwd_fwd, nb); memcpy(ps_bck_fwd, ps_fwd_fwd, nb);
      valid_subtree = build_tree(S, S->depth, &z_propose, ps_fwd_bck, ps_fwd_fwd, rho_fwd, p_fwd_bck, p_fwd_fwd,
                                 H0, 1, &n_leapfrog, &lsw_subtree, &sum_metro_prob, 0, S->depth);
      z_fwd = S->z;
    } else {
      S->z = z_bck;
      memcpy(rho_fwd, rho, nb);
      memcpy(p_fwd_bck, p_bck_bck, nb); memcpy(ps_fwd_bck, ps_bck_bck, nb);
      valid_subtree = build_tree(S, S->depth, &z_propose, ps_bck_fwd, ps_bck_bck, rho_bck, p_bck_fwd, p_bck_bck,
                                 H0, -1, &n_leapfrog, &lsw_subtree, &sum_metro_prob, 0, S->depth);
      z_bck = S->z;
    }
    if (!valid_subtree) break;
    ++S->depth;
    if (lsw_subtree > log_sum_weight) {
      z_sample = z_propose;
    } else if (u[1] < exp(lsw_subtree - log_sum_weight)) {
      z_sample = z_propose;
    }
    log_sum_weight = log_sum_exp2(log_sum_weight, lsw_subtree);
    for (int d = 0; d < D; ++d) rho[d] = rho_bck[d] + rho_fwd[d];
    int persist = criterion(S, ps_bck_bck, ps_fwd_fwd, rho);
    double rho_ext[MAXD];
    for (int d = 0; d < D; ++d) rho_ext[d] = rho_bck[d] + p_fwd_bck[d];
    persist &= criterion(S, ps_bck_bck, ps_fwd_bck, rho_ext);
    for (int d = 0; d < D; ++d) rho_ext[d] = rho_fwd[d] + p_bck_fwd[d];
    persist &= criterion(S, ps_bck_fwd, ps_fwd_fwd, rho_ext);
    if (!persist) break;
  }
  S->n_leapfrog = n_leapfrog;
  S->energy = z_sample.H;
  *cur = z_sample;
  return sum_metro_prob / (double)n_leapfrog;
}

/* Stan's init_stepsize heuristic.  `attempt` numbers the momentum refreshes for the RNG site. */
static void init_stepsize(nuts_t* S, const pspoint* cur) {
  if (S->eps == 0.0 || S->eps > 1e7 || isnan(S->eps)) return;
  const int D = S->D;
  uint32_t attempt = 0;
  int direction = 0;
  for (;;) {
    pspoint z = *cur;
    for (int d = 0; d < D; ++d) {
      uint32_t r[4];
      rng_block(&S->rng, S->it, SITE_INITEPS, attempt, (uint32_t)d, 0, r);
      z.p[d] = foct_oracle_normal(r) / sqrt(S->invM[d]);
    }
    ++attempt;
    double H0 = hamiltonian(S, &z);
    leapfrog(S, &z, S->eps);
    double h = hamiltonian(S, &z);
    if (isnan(h)) h = INFINITY;
    double dH = H0 - h;
    if (direction == 0) { direction = dH > log(0.8) ? 1 : -1; continue; }
    if (direction == 1 && !(dH > log(0.8))) break;
    if (direction == -1 && !(dH < log(0.8))) break;
    S->eps = direction == 1 ? 2.0 * S->eps : 0.5 * S->eps;
    if (S->eps > 1e7 || S->eps == 0.0 || attempt > 200) break; /* Stan throws here; we stop adapting */
  }
}

/* windowed adaptation state (Stan windowed_adaptation + welford_var_estimator + stepsize_adaptation) */
typedef struct {
  int num_warmup, init_buffer, term_buffer, base_window;
  int window_counter, window_size, next_window;
  double n, mean[MAXD], m2[MAXD];
  double mu, delta, gamma, kappa, t0, counter, s_bar, x_bar;
} adapt_t;

static void adapt_init(adapt_t* A, const foct_sampler_cfg* cfg, int D) {
  memset(A, 0, sizeof(*A));
  int ib = cfg->init_buffer > 0 ? cfg->init_buffer : 75;
  int tb = cfg->term_buffer > 0 ? cfg->term_buffer : 50;
  int bw = cfg->window > 0 ? cfg->window : 25;
  int nw = cfg->n_warmup;
  if (nw < 20) {
    A->num_warmup = 0; A->init_buffer = 0; A->term_buffer = 0; A->base_window = 0;
  } else {
    if (ib + bw + tb > nw) {
      ib = (int)(0.15 * nw); tb = (int)(0.1 * nw); bw = nw - (ib + tb);
    }
    A->num_warmup = nw; A->init_buffer = ib; A->term_buffer = tb; A->base_window = bw;
  }
  A->window_counter = 0; A->window_size = A->base_window;
  A->next_window = A->init_buffer + A->window_size - 1;
  A->delta = cfg->adapt_delta > 0 ? cfg->adapt_delta : 0.8;
  A->gamma = cfg->gamma > 0 ? cfg->gamma : 0.05;
  A->kappa = cfg->kappa > 0 ? cfg->kappa : 0.75;
  A->t0 = cfg->t0 > 0 ? cfg->t0 : 10.0;
  (void)D;
}
static void da_restart(adapt_t* A) { A->counter = 0; A->s_bar = 0; A->x_bar = 0; }
static double da_learn(adapt_t* A, double adapt_stat) {
  A->counter += 1.0;
  if (adapt_stat > 1.0) adapt_stat = 1.0;
  double eta = 1.0 / (A->counter + A->t0);
  A->s_bar = (1.0 - eta) * A->s_bar + eta * (A->delta - adapt_stat);
  double x = A->mu - A->s_bar * sqrt(A->counter) / A->gamma;
  double x_eta = pow(A->counter, -A->kappa);
  A->x_bar = (1.0 - x_eta) * A->x_bar + x_eta * x;
  return exp(x);
}
static int learn_variance(adapt_t* A, double* invM, const double* q, int D) {
  int in_window = A->window_counter >= A->init_buffer && A->window_counter < A->num_warmup - A->term_buffer &&
                  A->window_counter != A->num_warmup;
  if (in_window) {
    A->n += 1.0;
    for (int d = 0; d < D; ++d) {
      double delta = q[d] - A->mean[d];
      A->mean[d] += delta / A->n;
      A->m2[d] += (q[d] - A->mean[d]) * delta;
    }
  }
  int end_window = A->window_counter == A->next_window && A->window_counter != A->num_warmup;
  if (end_window) {
    /* compute_next_window */
    if (A->next_window != A->num_warmup - A->term_buffer - 1) {
      A->window_size *= 2;
      A->next_window = A->window_counter + A->window_size;
      if (A->next_window != A->num_warmup - A->term_buffer - 1) {
        int boundary = A->next_window + 2 * A->window_size;
        if (boundary >= A->num_warmup - A->term_buffer) A->next_window = A->num_warmup - A->term_buffer - 1;
      }
    }
    double n = A->n;
    for (int d = 0; d < D; ++d) {
      double var = A->m2[d] / (n - 1.0);
      invM[d] = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0));
      A->mean[d] = 0.0; A->m2[d] = 0.0;
    }
    A->n = 0.0;
    ++A->window_counter;
    return 1;
  }
  ++A->window_counter;
  return 0;
}

static void initial_point(nuts_t* S, const foct_problem* P, const foct_sampler_cfg* cfg, const double* init,
                          pspoint* cur) {
  const int D = S->D;
  for (int d = 0; d < D; ++d) {
    uint32_t r[4]; double u[2];
    rng_block(&S->rng, 0, SITE_INIT, 0, (uint32_t)d, 0, r);
    if (cfg->init_mode == 2 && init) {
      cur->q[d] = init[d];
    } else if (cfg->init_mode == 1 || !P) {
      foct_oracle_uniform2(r, u);
      cur->q[d] = -2.0 + 4.0 * u[0];
    } else {
      int Nn = S->M->Nn;
      if (d < 3) cur->q[d] = P->theta0[d];
      else if (d < 3 + Nn) cur->q[d] = 0.01 * foct_oracle_normal(r);
      else if (d == 3 + Nn) cur->q[d] = log(0.1);
      else cur->q[d] = 0.0;
    }
  }
  memset(cur->p, 0, sizeof(cur->p));
  eval_potential(S, cur);
}

/* Run one chain.  out_row(iter_saved) callbacks are replaced by direct writes through strides. */
typedef struct {
  double* draws; size_t draw_stride;   /* per saved iteration */
  double* sp; size_t sp_stride;
  int P_out;
  double* stepsize; double* inv_metric; double* n_leapfrog; double* n_divergent;
  double* last_q;                      /* [D] unconstrained state after the last transition */
  const double* invm_init;             /* [D] continuation: adapted metric, or NULL */
  const double* eps_init;              /* continuation: adapted step size, or NULL */
} chain_out;

static void write_draw(const nuts_t* S, const pspoint* cur, double accept, double eps_used, double* row,
                       double* sp) {
  if (row) {
    if (S->M) {
      const model_t* M = S->M;
      int Nn = M->Nn;
      row[0] = cur->q[0]; row[1] = cur->q[1]; row[2] = cur->q[2];
      if (M->kind == FOCT_EXPGP) {
        for (int k = 0; k < Nn; ++k) row[3 + k] = cur->q[3 + k];
        row[3 + Nn] = exp(cur->q[3 + Nn]);
        row[4 + Nn] = exp(cur->q[4 + Nn]);
        row[5 + Nn] = M->prior_PD ? NAN : cur->chi2 / br_ndf(M);
        row[6 + Nn] = -cur->V;
      } else {
        row[3] = M->prior_PD ? NAN : cur->chi2 / br_ndf(M);
        row[4] = -cur->V;
      }
    } else {
      for (int d = 0; d < S->D; ++d) row[d] = cur->q[d];
    }
  }
  if (sp) {
    sp[0] = accept; sp[1] = eps_used; sp[2] = (double)S->depth; sp[3] = (double)S->n_leapfrog;
    sp[4] = (double)S->divergent; sp[5] = S->energy;
  }
}

static void run_chain(nuts_t* S, const foct_problem* P, const foct_sampler_cfg* cfg, const double* init,
                      chain_out* O) {
  const int D = S->D;
  pspoint cur;
  memset(&cur, 0, sizeof(cur));
  for (int d = 0; d < D; ++d) S->invM[d] = O->invm_init ? O->invm_init[d] : 1.0;
  S->eps = cfg->stepsize0 > 0 ? cfg->stepsize0 : 1.0;
  if (O->eps_init) S->eps = *O->eps_init;
  S->max_depth = cfg->max_treedepth > 0 ? cfg->max_treedepth : 10;
  S->max_dH = 1000.0;
  S->it = 0;
  const uint32_t it0 = (uint32_t)cfg->iter_offset; /* continuation: Philox sites after those of the earlier run */
  initial_point(S, P, cfg, init, &cur);
  adapt_t A;
  adapt_init(&A, cfg, D);
  A.mu = log(10.0 * S->eps);
  da_restart(&A);
  double nlf[2] = {0, 0}, ndiv = 0;
  S->it = it0;
  if (cfg->n_warmup > 0) init_stepsize(S, &cur);
  for (int it = 0; it < cfg->n_iter; ++it) {
    S->it = it0 + (uint32_t)it;
    double eps_used = S->eps;
    double accept = nuts_transition(S, &cur);
    int warm = it < cfg->n_warmup;
    nlf[warm ? 0 : 1] += S->n_leapfrog;
    if (!warm) ndiv += S->divergent;
    int save_idx = cfg->save_warmup ? it : it - cfg->n_warmup;
    if (save_idx >= 0)
      write_draw(S, &cur, accept, eps_used, O->draws ? O->draws + (size_t)save_idx * O->draw_stride : NULL,
                 O->sp ? O->sp + (size_t)save_idx * O->sp_stride : NULL);
    if (warm) {
      S->eps = da_learn(&A, accept);
      if (learn_variance(&A, S->invM, cur.q, D)) {
        S->it = it0 + (uint32_t)(it + 1);
        init_stepsize(S, &cur);
        A.mu = log(10.0 * S->eps);
        da_restart(&A);
      }
      if (it == cfg->n_warmup - 1) S->eps = exp(A.x_bar);
    }
  }
  if (O->stepsize) *O->stepsize = S->eps;
  if (O->inv_metric) memcpy(O->inv_metric, S->invM, sizeof(double) * D);
  if (O->n_leapfrog) { O->n_leapfrog[0] = nlf[0]; O->n_leapfrog[1] = nlf[1]; }
  if (O->n_divergent) *O->n_divergent = ndiv;
  if (O->last_q) memcpy(O->last_q, cur.q, sizeof(double) * D);
}

int foct_oracle_sample(int kind, const foct_problem* P, int n_problems, const foct_model_spec* spec,
                       const foct_sampler_cfg* cfg, foct_result* R, int n_threads) {
  if (n_problems < 0 || cfg->chains < 1 || cfg->chains > FOCT_MAX_CHAINS || cfg->n_iter < cfg->n_warmup) return FOCT_EINVAL;
  const int C = cfg->chains;
  const int n_saved = cfg->save_warmup ? cfg->n_iter : cfg->n_iter - cfg->n_warmup;
  int rc_all = 0, used = 1;
#ifdef _OPENMP
  if (n_threads <= 0) n_threads = omp_get_max_threads();
  used = n_threads;
#else
  (void)n_threads;
#endif
  model_t* models = (model_t*)calloc((size_t)(n_problems > 0 ? n_problems : 1), sizeof(model_t));
  if (!models) return FOCT_ENOMEM;
#pragma omp parallel for schedule(dynamic, 1) num_threads(used)
  for (int j = 0; j < n_problems; ++j) {
    int rc = model_init(&models[j], kind, &P[j], spec, NULL);
    if (rc) {
#pragma omp critical
      rc_all = rc;
    }
  }
  if (!rc_all) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(used)
    for (int w = 0; w < n_problems * C; ++w) {
      int j = w / C, c = w % C;
      const model_t* M = &models[j];
      int D = M->D, P_out = kind == FOCT_EXPGP ? M->Nn + 7 : 5;
      nuts_t* S = (nuts_t*)calloc(1, sizeof(nuts_t));
      S->M = M; S->D = D;
      rng_seed(&S->rng, cfg->seed, P[j].id, c);
      chain_out O;
      memset(&O, 0, sizeof(O));
      O.P_out = P_out;
      if (R->draws) { O.draws = R->draws + ((size_t)j * n_saved * C + c) * P_out; O.draw_stride = (size_t)C * P_out; }
      if (R->sampler_params) { O.sp = R->sampler_params + ((size_t)j * n_saved * C + c) * 6; O.sp_stride = (size_t)C * 6; }
      if (R->stepsize) O.stepsize = R->stepsize + (size_t)j * C + c;
      if (R->inv_metric) O.inv_metric = R->inv_metric + ((size_t)j * C + c) * D;
      if (R->n_leapfrog) O.n_leapfrog = R->n_leapfrog + ((size_t)j * C + c) * 2;
      if (R->n_divergent) O.n_divergent = R->n_divergent + (size_t)j * C + c;
      if (R->last_q) O.last_q = R->last_q + ((size_t)j * C + c) * D;
      if (cfg->inv_metric_init) O.invm_init = cfg->inv_metric_init + ((size_t)j * C + c) * D;
      if (cfg->stepsize_init) O.eps_init = cfg->stepsize_init + (size_t)j * C + c;
      const double* init = cfg->init_mode == 2 && cfg->init ? cfg->init + ((size_t)j * C + c) * D : NULL;
      run_chain(S, &P[j], cfg, init, &O);
      free(S);
    }
    if (R->summary && R->draws) {
      int n_post = cfg->n_iter - cfg->n_warmup;
      int off = cfg->save_warmup ? cfg->n_warmup : 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(used)
      for (int j = 0; j < n_problems; ++j) {
        int P_out = kind == FOCT_EXPGP ? models[j].Nn + 7 : 5;
        foct_oracle_summary(R->draws + ((size_t)j * n_saved + off) * C * P_out, n_post, C, P_out,
                            R->summary + (size_t)j * P_out * FOCT_N_SUMMARY_COLS);
      }
    }
  }
  for (int j = 0; j < n_problems; ++j) model_free(&models[j]);
  free(models);
  return rc_all ? rc_all : used;
}

int foct_oracle_sample_analytic(int target, int D, const double* par, const foct_sampler_cfg* cfg,
                                double* draws, double* sampler_params) {
  if (D < 1 || D > MAXD) return FOCT_EINVAL;
  const int C = cfg->chains;
  const int n_saved = cfg->save_warmup ? cfg->n_iter : cfg->n_iter - cfg->n_warmup;
  (void)n_saved;
  analytic_t A = {target, D, par};
#pragma omp parallel for schedule(dynamic, 1)
  for (int c = 0; c < C; ++c) {
    nuts_t* S = (nuts_t*)calloc(1, sizeof(nuts_t));
    S->A = &A; S->D = D;
    rng_seed(&S->rng, cfg->seed, 0, c);
    chain_out O;
    memset(&O, 0, sizeof(O));
    O.P_out = D;
    if (draws) { O.draws = draws + (size_t)c * D; O.draw_stride = (size_t)C * D; }
    if (sampler_params) { O.sp = sampler_params + (size_t)c * 6; O.sp_stride = (size_t)C * 6; }
    foct_sampler_cfg cfg1 = *cfg;
    if (cfg1.init_mode == 0) cfg1.init_mode = 1;
    run_chain(S, NULL, &cfg1, cfg->init ? cfg->init + (size_t)c * D : NULL, &O);
    free(S);
  }
  return 0;
}

/* ------------------------------------------------------------------ summary (MODEL_SPEC §9) */

static int cmp_double(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}
typedef struct { double v; int idx; } vi_t;
static int cmp_vi(const void* a, const void* b) {
  const vi_t *x = (const vi_t*)a, *y = (const vi_t*)b;
  if (x->v < y->v) return -1;
  if (x->v > y->v) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

/* Stan compute_effective_sample_size on chains stored as x[c*n + t]. */
static double ess_stan(const double* x, int n, int C) {
  if (n < 4) return NAN;
  double* cm = (double*)malloc(sizeof(double) * C);
  double* acov = (double*)malloc(sizeof(double) * (size_t)C * n);
  double mean_var = 0.0;
  for (int c = 0; c < C; ++c) {
    const double* xc = x + (size_t)c * n;
    double m = 0.0;
    for (int t = 0; t < n; ++t) m += xc[t];
    m /= n; cm[c] = m;
    for (int s = 0; s < n; ++s) acov[(size_t)c * n + s] = NAN; /* lazily filled */
    double a0 = 0.0;
    for (int t = 0; t < n; ++t) a0 += (xc[t] - m) * (xc[t] - m);
    acov[(size_t)c * n] = a0 / n;
    mean_var += acov[(size_t)c * n] * n / (n - 1.0);
  }
  mean_var /= C;
  double var_plus = mean_var * (n - 1.0) / n;
  if (C > 1) {
    double mm = 0.0, v = 0.0;
    for (int c = 0; c < C; ++c) mm += cm[c];
    mm /= C;
    for (int c = 0; c < C; ++c) v += (cm[c] - mm) * (cm[c] - mm);
    var_plus += v / (C - 1.0);
  }
  double ess = NAN;
  if (var_plus > 0.0 && isfinite(var_plus)) {
    double* rho = (double*)calloc((size_t)n + 2, sizeof(double));
#define MEAN_ACOV(s, out)                                                     \
  do {                                                                        \
    double acc_ = 0.0;                                                        \
    for (int c_ = 0; c_ < C; ++c_) {                                          \
      const double* xc_ = x + (size_t)c_ * n;                                 \
      double a_ = 0.0, m_ = cm[c_];                                           \
      for (int t_ = 0; t_ + (s) < n; ++t_) a_ += (xc_[t_] - m_) * (xc_[t_ + (s)] - m_); \
      acc_ += a_ / n;                                                         \
    }                                                                         \
    (out) = acc_ / C;                                                         \
  } while (0)
    double a1;
    MEAN_ACOV(1, a1);
    double rho_even = 1.0, rho_odd = 1.0 - (mean_var - a1) / var_plus;
    rho[0] = rho_even; rho[1] = rho_odd;
    int s = 1;
    while (s < n - 4 && (rho_even + rho_odd) > 0.0) {
      double ae, ao;
      MEAN_ACOV(s + 1, ae);
      MEAN_ACOV(s + 2, ao);
      rho_even = 1.0 - (mean_var - ae) / var_plus;
      rho_odd = 1.0 - (mean_var - ao) / var_plus;
      if (rho_even + rho_odd >= 0.0) { rho[s + 1] = rho_even; rho[s + 2] = rho_odd; }
      s += 2;
    }
    int max_s = s;
    if (rho_even > 0.0) rho[max_s + 1] = rho_even;
    for (s = 1; s <= max_s - 3; s += 2) {
      if (rho[s + 1] + rho[s + 2] > rho[s - 1] + rho[s]) {
        rho[s + 1] = (rho[s - 1] + rho[s]) / 2.0;
        rho[s + 2] = rho[s + 1];
      }
    }
    double sum = 0.0;
    for (s = 0; s <= max_s; ++s) sum += rho[s];
    double tau = -1.0 + 2.0 * sum + rho[max_s + 1];
    double nt = (double)n * C;
    /* rstan's ess_rfun: tau_hat = max(tau_hat, 1 / log10(S)) - caps the ESS at S log10(S) and keeps a strongly
       antithetic column (tau <= 0) finite and positive */
    tau = fmax(tau, 1.0 / log10(nt));
    ess = nt / tau;
    free(rho);
#undef MEAN_ACOV
  }
  free(cm); free(acov);
  return ess;
}

/* Acklam/Wichura-grade inverse normal CDF is overkill here; use AS241 (PPND16), accurate to 1e-16. */
static double inv_norm_cdf(double p) {
  double q = p - 0.5, r, val;
  if (fabs(q) <= 0.425) {
    r = 0.180625 - q * q;
    val = q * (((((((2.5090809287301226727e3 * r + 3.3430575583588128105e4) * r + 6.7265770927008700853e4) * r + 4.5921953931549871457e4) * r + 1.3731693765509461125e4) * r + 1.9715909503065514427e3) * r + 1.3314166789178437745e2) * r + 3.3871328727963666080e0) /
          (((((((5.2264952788528545610e3 * r + 2.8729085735721942674e4) * r + 3.9307895800092710610e4) * r + 2.1213794301586595867e4) * r + 5.3941960214247511077e3) * r + 6.8718700749205790830e2) * r + 4.2313330701600911252e1) * r + 1.0);
    return val;
  }
  r = q < 0 ? p : 1.0 - p;
  r = sqrt(-log(r));
  if (r <= 5.0) {
    r -= 1.6;
    val = (((((((7.74545014278341407640e-4 * r + 2.27238449892691845833e-2) * r + 2.41780725177450611770e-1) * r + 1.27045825245236838258e0) * r + 3.64784832476320460504e0) * r + 5.76949722146069140550e0) * r + 4.63033784615654529590e0) * r + 1.42343711074968357734e0) /
          (((((((1.05075007164441684324e-9 * r + 5.47593808499534494600e-4) * r + 1.51986665636164571966e-2) * r + 1.48103976427480074590e-1) * r + 6.89767334985100004550e-1) * r + 1.67638483018380384940e0) * r + 2.05319162663775882187e0) * r + 1.0);
  } else {
    r -= 5.0;
    val = (((((((2.01033439929228813265e-7 * r + 2.71155556874348757815e-5) * r + 1.24266094738807843860e-3) * r + 2.65321895265761230930e-2) * r + 2.96560571828504891230e-1) * r + 1.78482653991729133580e0) * r + 5.46378491116411436990e0) * r + 6.65790464350110377720e0) /
          (((((((2.04426310338993978564e-15 * r + 1.42151175831644588870e-7) * r + 1.84631831751005468180e-5) * r + 7.86869131145613259100e-4) * r + 1.48753612908506148525e-2) * r + 1.36929880922735805310e-1) * r + 5.99832206555887937690e-1) * r + 1.0);
  }
  return q < 0 ? -val : val;
}

int foct_oracle_summary(const double* draws, int n, int C, int P, double* out) {
  int S = n * C;
  if (n < 1 || C < 1) return FOCT_EINVAL;
  double* col = (double*)malloc(sizeof(double) * S);    /* chain-major x[c*n+t] */
  double* sorted = (double*)malloc(sizeof(double) * S);
  int h = n / 2;                                         /* split-chain length */
  double* split = (double*)malloc(sizeof(double) * (size_t)(2 * C) * (h > 0 ? h : 1));
  vi_t* vi = (vi_t*)malloc(sizeof(vi_t) * (size_t)(2 * C) * (h > 0 ? h : 1));
  for (int p = 0; p < P; ++p) {
    double* o = out + (size_t)p * FOCT_N_SUMMARY_COLS;
    int bad = 0;
    for (int c = 0; c < C; ++c)
      for (int t = 0; t < n; ++t) {
        double v = draws[((size_t)t * C + c) * P + p];
        col[(size_t)c * n + t] = v;
        if (!isfinite(v)) bad = 1;
      }
    if (bad) { for (int k = 0; k < FOCT_N_SUMMARY_COLS; ++k) o[k] = NAN; continue; }
    double mean = 0.0;
    for (int i = 0; i < S; ++i) mean += col[i];
    mean /= S;
    double ss = 0.0;
    for (int i = 0; i < S; ++i) ss += (col[i] - mean) * (col[i] - mean);
    memcpy(sorted, col, sizeof(double) * S);
    qsort(sorted, S, sizeof(double), cmp_double);
    if (sorted[0] == sorted[S - 1]) ss = 0.0; /* constant column: sd 0, diagnostics NaN (rstan prints NaN) */
    double sd = S > 1 ? sqrt(ss / (S - 1.0)) : NAN;
    static const double probs[5] = {0.025, 0.25, 0.5, 0.75, 0.975};
    for (int k = 0; k < 5; ++k) { /* R type 7 */
      double hq = (S - 1) * probs[k];
      int lo = (int)floor(hq);
      int hi = lo + 1 < S ? lo + 1 : lo;
      o[3 + k] = sorted[lo] + (hq - lo) * (sorted[hi] - sorted[lo]);
    }
    double n_eff = (ss > 0.0) ? ess_stan(col, n, C) : NAN;
    /* split-Rhat */
    double rhat = NAN, bulk = NAN;
    if (h >= 2 && ss > 0.0) {
      int C2 = 2 * C;
      for (int c = 0; c < C; ++c) {
        memcpy(split + (size_t)(2 * c) * h, col + (size_t)c * n, sizeof(double) * h);
        memcpy(split + (size_t)(2 * c + 1) * h, col + (size_t)c * n + (n - h), sizeof(double) * h);
      }
      double W = 0.0, gm = 0.0, cmn[2 * FOCT_MAX_CHAINS];
      for (int c = 0; c < C2; ++c) {
        double m = 0.0, v = 0.0;
        for (int t = 0; t < h; ++t) m += split[(size_t)c * h + t];
        m /= h;
        for (int t = 0; t < h; ++t) v += (split[(size_t)c * h + t] - m) * (split[(size_t)c * h + t] - m);
        W += v / (h - 1.0); cmn[c] = m; gm += m;
      }
      W /= C2; gm /= C2;
      double Bv = 0.0;
      for (int c = 0; c < C2; ++c) Bv += (cmn[c] - gm) * (cmn[c] - gm);
      Bv /= (C2 - 1.0); /* variance of chain means (= B/n') */
      rhat = sqrt((W * (h - 1.0) / h + Bv) / W);
      /* bulk ESS: rank-normalise the split chains */
      int S2 = C2 * h;
      for (int i = 0; i < S2; ++i) { vi[i].v = split[i]; vi[i].idx = i; }
      qsort(vi, S2, sizeof(vi_t), cmp_vi);
      for (int r = 0; r < S2; ++r) split[vi[r].idx] = inv_norm_cdf(((double)(r + 1) - 0.375) / ((double)S2 + 0.25));
      bulk = ess_stan(split, h, C2);
    }
    o[0] = mean; o[1] = sd / sqrt(n_eff); o[2] = sd; o[8] = n_eff; o[9] = rhat; o[10] = bulk;
  }
  free(col); free(sorted); free(split); free(vi);
  return 0;
}


/* ------------------------------------------------------------------ ExpGP MAP (MODEL_SPEC §10)
 * BFGS on the unconstrained space without Jacobian terms (rstan::optimizing default, jacobian = FALSE), Armijo
 * backtracking, then a central finite-difference Hessian of the gradient.  Consumers: plotExpGP.R:13-17,
 * ShinyInterface/server.R:164-173. */
static double map_obj(const model_t* M, const double* q, double* g, double* chi2) {
  double lp = model_lpg(M, q, g, chi2, NULL);
  double jl = 0.0;
  if (M->kind == FOCT_EXPGP) {
    jl = q[3 + M->Nn] + q[4 + M->Nn];
    g[3 + M->Nn] -= 1.0; g[4 + M->Nn] -= 1.0;
  }
  for (int d = 0; d < M->D; ++d) g[d] = -g[d];
  return -(lp - jl);
}

int foct_oracle_expgp_map(const foct_problem* P, int n_problems, const foct_model_spec* spec, const double* init,
                          double* par, double* hessian, int* status) {
  for (int j = 0; j < n_problems; ++j) {
    model_t M;
    int rc = model_init(&M, FOCT_EXPGP, &P[j], spec, NULL);
    if (rc) return rc;
    const int D = M.D, Nn = M.Nn, P_out = Nn + 7;
    double q[MAXD], g[MAXD], h0[MAXD], H[MAXD][MAXD], chi2;
    for (int d = 0; d < D; ++d) {
      if (init) q[d] = init[(size_t)j * D + d];
      else if (d < 3) q[d] = P[j].theta0[d];
      else if (d < 3 + Nn) q[d] = 0.0;
      else if (d == 3 + Nn) q[d] = log(0.1);
      else q[d] = 0.0;
      if (d < 3) h0[d] = spec->theta_prior == 0 ? 1.0 / M.Pinv[d * 4] : (d == 2 ? 100.0 : 1.0e4);
      else if (d < 3 + Nn) h0[d] = 2.5e-3;
      else if (d == 3 + Nn) h0[d] = 0.25;
      else h0[d] = 0.01;
    }
    for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) H[a][b] = a == b ? h0[a] : 0.0;
    double f = map_obj(&M, q, g, &chi2);
    int st = 1;
    for (int it = 0; it < 1000; ++it) {
      double dv[MAXD], slope = 0.0;
      for (int a = 0; a < D; ++a) { double s = 0.0; for (int b = 0; b < D; ++b) s += H[a][b] * g[b]; dv[a] = -s; }
      for (int a = 0; a < D; ++a) slope += g[a] * dv[a];
      if (!(slope < 0.0)) {
        for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) H[a][b] = a == b ? h0[a] : 0.0;
        slope = 0.0;
        for (int a = 0; a < D; ++a) { dv[a] = -h0[a] * g[a]; slope += g[a] * dv[a]; }
        if (!(slope < 0.0)) { st = 0; break; }
      }
      double step = 1.0, fn = 0.0, gn[MAXD], qn[MAXD], c2n = 0.0;
      int ok = 0;
      for (int ls = 0; ls < 40; ++ls) {
        for (int a = 0; a < D; ++a) qn[a] = q[a] + step * dv[a];
        fn = map_obj(&M, qn, gn, &c2n);
        if (isfinite(fn) && fn <= f + 1e-4 * step * slope) { ok = 1; break; }
        step *= 0.5;
      }
      if (!ok) { st = 2; break; }
      double s[MAXD], y[MAXD], sy = 0.0, ss = 0.0, yy = 0.0;
      for (int a = 0; a < D; ++a) { s[a] = qn[a] - q[a]; y[a] = gn[a] - g[a]; sy += s[a] * y[a]; ss += s[a] * s[a]; yy += y[a] * y[a]; }
      double df = f - fn, fold = f;
      memcpy(q, qn, sizeof(double) * D); memcpy(g, gn, sizeof(double) * D); chi2 = c2n; f = fn;
      if (df <= 1e-13 * (fabs(fold) + 1.0)) { st = 0; break; }
      if (sy > 1e-12 * sqrt(ss * yy)) {
        double Hy[MAXD], yHy = 0.0;
        for (int a = 0; a < D; ++a) { double t = 0.0; for (int b = 0; b < D; ++b) t += H[a][b] * y[b]; Hy[a] = t; }
        for (int a = 0; a < D; ++a) yHy += y[a] * Hy[a];
        double c1 = (sy + yHy) / (sy * sy), isy = 1.0 / sy;
        for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) H[a][b] = H[a][b] + c1 * s[a] * s[b] - (Hy[a] * s[b] + s[a] * Hy[b]) * isy;
      }
    }
    double* row = par + (size_t)j * P_out;
    for (int d = 0; d < 3 + Nn; ++d) row[d] = q[d];
    row[3 + Nn] = exp(q[3 + Nn]); row[4 + Nn] = exp(q[4 + Nn]);
    row[5 + Nn] = M.prior_PD ? NAN : chi2 / br_ndf(&M);
    row[6 + Nn] = -f;
    if (status) status[j] = st;
    if (hessian) {
      for (int c = 0; c < D; ++c) {
        double h = 1e-5 * (fabs(q[c]) > 1.0 ? fabs(q[c]) : 1.0), qp[MAXD], qm[MAXD], gp[MAXD], gm[MAXD], cc;
        memcpy(qp, q, sizeof(double) * D); memcpy(qm, q, sizeof(double) * D);
        qp[c] += h; qm[c] -= h;
        map_obj(&M, qp, gp, &cc); map_obj(&M, qm, gm, &cc);
        for (int a = 0; a < D; ++a) hessian[((size_t)j * D + a) * D + c] = -(gp[a] - gm[a]) / (2.0 * h);
      }
    }
    model_free(&M);
  }
  return 0;
}

/* ------------------------------------------------------------------ MonoExp MAP (SURVEY a-12) */

/* -lp, gradient and exact Hessian of -lp for the mono-exponential (sigma == 1). */
static double mono_nlp(const model_t* M, const double* th, double* grad, double* H, double* chi2) {
  double f = 0.0, c2 = 0.0;
  for (int k = 0; k < 3; ++k) grad[k] = 0.0;
  for (int k = 0; k < 9; ++k) H[k] = 0.0;
  for (int i = 0; i < M->N; ++i) {
    double w = 1.0 / M->uy[i], w2 = w * w;
    double t = M->c * M->x[i] / th[2];
    double e = exp(-t);
    double m = th[0] + th[1] * e;
    double r = M->y[i] - m;
    double J[3] = {1.0, e, th[1] * e * t / th[2]};
    double m23 = e * t / th[2];
    double m33 = th[1] * e * (t * t - 2.0 * t) / (th[2] * th[2]);
    f += 0.5 * r * r * w2;
    c2 += r * r * w2;
    for (int a = 0; a < 3; ++a) {
      grad[a] -= r * w2 * J[a];
      for (int b = 0; b < 3; ++b) H[a * 3 + b] += w2 * J[a] * J[b];
    }
    H[1 * 3 + 2] -= w2 * r * m23; H[2 * 3 + 1] -= w2 * r * m23;
    H[2 * 3 + 2] -= w2 * r * m33;
  }
  if (M->spec.theta_prior == 0) {
    double d[3] = {th[0] - M->theta0[0], th[1] - M->theta0[1], th[2] - M->theta0[2]};
    for (int a = 0; a < 3; ++a) {
      double v = 0.0;
      for (int b = 0; b < 3; ++b) { v += M->Pinv[a * 3 + b] * d[b]; H[a * 3 + b] += M->Pinv[a * 3 + b]; }
      f += 0.5 * d[a] * v; grad[a] += v;
    }
  }
  if (chi2) *chi2 = c2;
  return f;
}

static int solve3(const double* A, const double* b, double* x) {
  double Ai[9];
  if (inv3(A, Ai)) return 1;
  for (int a = 0; a < 3; ++a) x[a] = Ai[a * 3] * b[0] + Ai[a * 3 + 1] * b[1] + Ai[a * 3 + 2] * b[2];
  return 0;
}

static void mono_start(const model_t* M, double* th) {
  double ymin = M->y[0], ymax = M->y[0];
  for (int i = 1; i < M->N; ++i) { if (M->y[i] < ymin) ymin = M->y[i]; if (M->y[i] > ymax) ymax = M->y[i]; }
  double th1 = ymin - 0.05 * (ymax - ymin);
  double sx = 0, sy = 0, sxx = 0, sxy = 0; int n = 0;
  for (int i = 0; i < M->N; ++i) {
    double v = M->y[i] - th1;
    if (v > 0.0) { double ly = log(v); sx += M->x[i]; sy += ly; sxx += M->x[i] * M->x[i]; sxy += M->x[i] * ly; ++n; }
  }
  double slope = (n * sxy - sx * sy) / (n * sxx - sx * sx);
  double icpt = (sy - slope * sx) / n;
  th[0] = th1; th[1] = exp(icpt);
  th[2] = slope < 0.0 ? -M->c / slope : (M->x[M->N - 1] - M->x[0]);
}

int foct_oracle_monoexp_map(const foct_problem* P, int n_problems, const foct_model_spec* spec,
                            const double* init, double* theta, double* hessian, double* br, int* status) {
  for (int j = 0; j < n_problems; ++j) {
    model_t M;
    int rc = model_init(&M, FOCT_MONOEXP, &P[j], spec, NULL);
    if (rc) return rc;
    double th[3], g[3], H[9], c2;
    if (init) memcpy(th, init + (size_t)j * 3, sizeof(th)); else mono_start(&M, th);
    double f = mono_nlp(&M, th, g, H, &c2);
    double mu = 1e-3;
    int st = 1;
    for (int it = 0; it < 200; ++it) {
      double A[9], rhs[3], step[3], tn[3], gn[3], Hn[9], c2n;
      memcpy(A, H, sizeof(A));
      for (int a = 0; a < 3; ++a) { A[a * 3 + a] += mu * fabs(H[a * 3 + a]) + 1e-300; rhs[a] = -g[a]; }
      if (solve3(A, rhs, step)) { mu *= 10.0; continue; }
      for (int a = 0; a < 3; ++a) tn[a] = th[a] + step[a];
      double fn = mono_nlp(&M, tn, gn, Hn, &c2n);
      if (isfinite(fn) && fn <= f) {
        double rel = 0.0;
        for (int a = 0; a < 3; ++a) { double r = fabs(step[a]) / (fabs(th[a]) + 1e-300); if (r > rel) rel = r; }
        memcpy(th, tn, sizeof(th)); memcpy(g, gn, sizeof(g)); memcpy(H, Hn, sizeof(H)); c2 = c2n; f = fn;
        mu = mu * 0.2 > 1e-12 ? mu * 0.2 : 1e-12;
        if (rel < 1e-12) { st = 0; break; }
      } else {
        mu *= 5.0;
        if (mu > 1e12) break;
      }
    }
    memcpy(theta + (size_t)j * 3, th, sizeof(th));
    if (hessian) for (int k = 0; k < 9; ++k) hessian[(size_t)j * 9 + k] = -H[k];
    if (br) br[j] = c2 / br_ndf(&M);
    if (status) status[j] = st;
    model_free(&M);
  }
  return 0;
}
#include "foct_oracle_vb.c"
