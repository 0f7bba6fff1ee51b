/*
 * foct_oracle_prep.c — CPU restatement of the steps either side of the hot path:
 * estimateNoise (smoothing spline + heteroscedastic noise fit), printBr (the gate) and estimateExpPrior.
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see foct_oracle.h): FitOCTLib's sources are absent, the algorithms
 * follow MODEL_SPEC.md §11-13 and the reference call sites cited there (FitOCT.R:89-107, Tests/statsSplineSmooth.R:21).
 * The spline machinery is pinned against scipy (tests/test_oracle_prep.py), qchisq against scipy.stats.
 */
#include "foct_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* R's .nknots.smspl (MODEL_SPEC §11). */
int foct_oracle_nknots(int n) {
  if (n < 50) return n;
  const double a1 = log2(50.0), a2 = log2(100.0), a3 = log2(140.0), a4 = log2(200.0);
  /* 1e-9: the documented break points (n = 200 -> 100, 800 -> 140, 3200 -> 200) must not fall to 99.999.. */
  if (n < 200) return (int)(exp2(a1 + (a2 - a1) * (n - 50) / 150.0) + 1e-9);
  if (n < 800) return (int)(exp2(a2 + (a3 - a2) * (n - 200) / 600.0) + 1e-9);
  if (n < 3200) return (int)(exp2(a3 + (a4 - a3) * (n - 800) / 2400.0) + 1e-9);
  return (int)(200.0 + pow((double)(n - 3200), 0.2) + 1e-9);
}

/* Values of the four cubic B-splines that are non-zero on [T[l], T[l+1]] (Cox-de Boor). */
static void bspl_val(const double* T, int l, double t, double N[4]) {
  double left[4], right[4];
  N[0] = 1.0;
  for (int k = 1; k <= 3; ++k) {
    left[k] = t - T[l + 1 - k];
    right[k] = T[l + k] - t;
    double saved = 0.0;
    for (int r = 0; r < k; ++r) {
      const double tmp = N[r] / (right[r + 1] + left[k - r]);
      N[r] = saved + right[r + 1] * tmp;
      saved = left[k - r] * tmp;
    }
    N[k] = saved;
  }
}

static double sdiv(double a, double b) { return b > 0.0 ? a / b : 0.0; }

/* Second derivatives of the same four B-splines, from the polynomial piece of interval l. */
static void bspl_d2(const double* T, int l, double t, double D2[4]) {
  /* order-2 (hat) values: B_{l-1,2}, B_{l,2} */
  const double h = T[l + 1] - T[l];
  const double b2[2] = {(T[l + 1] - t) / h, (t - T[l]) / h};
  /* first derivatives of the order-3 splines j = l-2 .. l */
  double d3[3];
  for (int a = 0; a < 3; ++a) {
    const int j = l - 2 + a;
    const double u = (j >= l - 1 && j <= l) ? b2[j - (l - 1)] : 0.0;          /* B_{j,2}   */
    const double v = (j + 1 >= l - 1 && j + 1 <= l) ? b2[j + 1 - (l - 1)] : 0.0; /* B_{j+1,2} */
    d3[a] = 2.0 * (sdiv(u, T[j + 2] - T[j]) - sdiv(v, T[j + 3] - T[j + 1]));
  }
  for (int a = 0; a < 4; ++a) {
    const int j = l - 3 + a;
    const double u = (a >= 1) ? d3[a - 1] : 0.0; /* (B_{j,3})'   */
    const double v = (a <= 2) ? d3[a] : 0.0;     /* (B_{j+1,3})' */
    D2[a] = 3.0 * (sdiv(u, T[j + 3] - T[j]) - sdiv(v, T[j + 4] - T[j + 1]));
  }
}

/* banded symmetric storage: M[d*nk + j] = M_{j, j+d}, d = 0..3 */
typedef struct {
  int N, nk, nknots;
  double* T;    /* nk + 4 knots */
  int* left;    /* interval of every point */
  double* XtX;  /* 4*nk */
  double* Om;   /* 4*nk */
  double* Xty;  /* nk */
  double* L;    /* 4*nk: L[d*nk+j] = L_{j+d,j}, d=1..3; L[j] = d_j */
  double* Z;    /* 4*nk band of the inverse */
  double* c;    /* nk */
  double r;
} spl_t;

static void spl_free(spl_t* S) {
  free(S->T); free(S->left); free(S->XtX); free(S->Om); free(S->Xty); free(S->L); free(S->Z); free(S->c);
}

static int spl_setup(spl_t* S, int N, const double* x, const double* y, int all_knots) {
  memset(S, 0, sizeof(*S));
  if (N < 4) return FOCT_EINVAL;
  for (int i = 1; i < N; ++i) if (!(x[i] > x[i - 1])) return FOCT_EINVAL;
  const int nknots = all_knots ? N : foct_oracle_nknots(N);
  const int nk = nknots + 2;
  S->N = N; S->nk = nk; S->nknots = nknots;
  S->T = calloc(nk + 4, sizeof(double));
  S->left = calloc(N, sizeof(int));
  S->XtX = calloc(4 * nk, sizeof(double)); S->Om = calloc(4 * nk, sizeof(double));
  S->Xty = calloc(nk, sizeof(double)); S->L = calloc(4 * nk, sizeof(double));
  S->Z = calloc(4 * nk, sizeof(double)); S->c = calloc(nk, sizeof(double));
  const double x0 = x[0], ir = 1.0 / (x[N - 1] - x[0]);
  /* knots: x[floor(seq(1, N, length.out = nknots))] */
  for (int k = 0; k < nknots; ++k) {
    int idx = nknots > 1 ? (int)floor(1.0 + (double)k * (double)(N - 1) / (double)(nknots - 1)) : 1;
    if (k == nknots - 1) idx = N;
    S->T[3 + k] = (x[idx - 1] - x0) * ir;
  }
  for (int k = 0; k < 3; ++k) { S->T[k] = S->T[3]; S->T[nk + 1 + k] = S->T[nk]; }
  /* design */
  int l = 3;
  for (int i = 0; i < N; ++i) {
    const double t = (x[i] - x0) * ir;
    while (l < nk - 1 && t >= S->T[l + 1]) ++l;
    S->left[i] = l;
    double B[4];
    bspl_val(S->T, l, t, B);
    for (int a = 0; a < 4; ++a) {
      S->Xty[l - 3 + a] += B[a] * y[i];
      for (int b = a; b < 4; ++b) S->XtX[(b - a) * nk + l - 3 + a] += B[a] * B[b];
    }
  }
  /* penalty: Simpson-exact integral of products of piecewise-linear second derivatives */
  for (int m = 3; m < nk; ++m) {
    const double h = S->T[m + 1] - S->T[m];
    if (!(h > 0.0)) continue;
    double A[4], Bv[4];
    bspl_d2(S->T, m, S->T[m], A);
    bspl_d2(S->T, m, S->T[m + 1], Bv);
    for (int a = 0; a < 4; ++a)
      for (int b = a; b < 4; ++b)
        S->Om[(b - a) * nk + m - 3 + a] += h / 6.0 * (2.0 * A[a] * A[b] + A[a] * Bv[b] + Bv[a] * A[b] + 2.0 * Bv[a] * Bv[b]);
  }
  double t1 = 0.0, t2 = 0.0;
  for (int j = 2; j <= nk - 4; ++j) { t1 += S->XtX[j]; t2 += S->Om[j]; }
  S->r = t1 / t2;
  return 0;
}

/* factor XtX + lam*Om, solve for the coefficients, return df = tr[(XtX + lam Om)^-1 XtX] */
static double spl_fit(spl_t* S, double lam) {
  const int nk = S->nk;
  double* L = S->L; double* Z = S->Z;
#define AB(d, j) (S->XtX[(d) * nk + (j)] + lam * S->Om[(d) * nk + (j)])
  for (int j = 0; j < nk; ++j) {
    double dj = AB(0, j);
    for (int k = (j - 3 > 0 ? j - 3 : 0); k < j; ++k) { const double ljk = L[(j - k) * nk + k]; dj -= ljk * ljk * L[k]; }
    L[j] = dj;
    for (int i = j + 1; i <= j + 3 && i < nk; ++i) {
      double s = AB(i - j, j);
      for (int k = (i - 3 > 0 ? i - 3 : 0); k < j; ++k) s -= L[(i - k) * nk + k] * L[(j - k) * nk + k] * L[k];
      L[(i - j) * nk + j] = s / dj;
    }
  }
#undef AB
  /* solve */
  double* c = S->c;
  for (int i = 0; i < nk; ++i) {
    double s = S->Xty[i];
    for (int k = (i - 3 > 0 ? i - 3 : 0); k < i; ++k) s -= L[(i - k) * nk + k] * c[k];
    c[i] = s;
  }
  for (int i = 0; i < nk; ++i) c[i] /= L[i];
  for (int i = nk - 1; i >= 0; --i) {
    double s = c[i];
    for (int k = i + 1; k <= i + 3 && k < nk; ++k) s -= L[(k - i) * nk + i] * c[k];
    c[i] = s;
  }
  /* band of the inverse */
  for (int i = nk - 1; i >= 0; --i) {
    const int jmax = i + 3 < nk - 1 ? i + 3 : nk - 1;
    for (int j = jmax; j >= i; --j) {
      double s = (i == j) ? 1.0 / L[i] : 0.0;
      for (int k = i + 1; k <= jmax; ++k) {
        const double zkj = k <= j ? Z[(j - k) * nk + k] : Z[(k - j) * nk + j];
        s -= L[(k - i) * nk + i] * zkj;
      }
      Z[(j - i) * nk + i] = s;
    }
  }
  double df = 0.0;
  for (int j = 0; j < nk; ++j) df += Z[j] * S->XtX[j];
  for (int d = 1; d <= 3; ++d)
    for (int j = 0; j + d < nk; ++j) df += 2.0 * Z[d * nk + j] * S->XtX[d * nk + j];
  return df;
}

static double spar_to_lambda(const spl_t* S, double spar) { return S->r * pow(256.0, 3.0 * spar - 1.0); }

/* info: spar, lambda, df reached, evaluations.  spar_fixed = NaN -> solve df(spar) = df. */
int foct_oracle_smooth_spline(int N, const double* x, const double* y, double df, int all_knots, double spar_fixed,
                              double* ySmooth, double* info) {
  spl_t S;
  int rc = spl_setup(&S, N, x, y, all_knots);
  if (rc) { spl_free(&S); return rc; }
  double spar, dfv;
  int evals = 0;
  if (!isnan(spar_fixed)) {
    spar = spar_fixed;
    dfv = spl_fit(&S, spar_to_lambda(&S, spar)); ++evals;
  } else {
    double a = -1.5, b = 1.5;
    double fa = spl_fit(&S, spar_to_lambda(&S, a)) - df; ++evals;
    if (fa <= 0.0) { spar = a; dfv = fa + df; }
    else {
      double fb = spl_fit(&S, spar_to_lambda(&S, b)) - df; ++evals;
      if (fb >= 0.0) { spar = b; dfv = fb + df; }
      else {
        int side = 0;
        spar = b; dfv = fb + df;
        for (int it = 0; it < 100; ++it) {
          const double c = (a * fb - b * fa) / (fb - fa);
          const double fc = spl_fit(&S, spar_to_lambda(&S, c)) - df; ++evals;
          spar = c; dfv = fc + df;
          if (fabs(fc) <= 1e-10) break;
          if (fc < 0.0) { b = c; fb = fc; if (side == -1) fa *= 0.5; side = -1; }
          else { a = c; fa = fc; if (side == 1) fb *= 0.5; side = 1; }
        }
      }
    }
  }
  for (int i = 0; i < N; ++i) {
    const double t = (x[i] - x[0]) * (1.0 / (x[N - 1] - x[0]));
    double B[4];
    const int l = S.left[i];
    bspl_val(S.T, l, t, B);
    ySmooth[i] = B[0] * S.c[l - 3] + B[1] * S.c[l - 2] + B[2] * S.c[l - 1] + B[3] * S.c[l];
  }
  if (info) { info[0] = spar; info[1] = spar_to_lambda(&S, spar); info[2] = dfv; info[3] = (double)evals; }
  spl_free(&S);
  return 0;
}

/* MODEL_SPEC §11 step 2: MLE of r_i ~ N(0, a1 exp(-x_i/a2)). */
int foct_oracle_noise_fit(int N, const double* x, const double* resid, double max_rate, double theta[2]) {
  const double vmin = 1.0 / max_rate, xN = x[N - 1];
  double sx = 0.0;
  for (int i = 0; i < N; ++i) sx += x[i];
  const double xbar = sx / N;
  double v = 0.0, S0 = 0.0;
  int at_bound = 0;
  for (int it = 0; it < 50; ++it) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < N; ++i) {
      const double w = resid[i] * resid[i] * exp(2.0 * v * (x[i] - xN));
      s0 += w; s1 += w * x[i]; s2 += w * x[i] * x[i];
    }
    S0 = s0;
    const double m1 = s1 / s0, var = s2 / s0 - m1 * m1;
    double dv = (xbar - m1) / (2.0 * var);
    double vn = v + dv;
    if (vn < vmin) { vn = vmin; if (at_bound) break; at_bound = 1; } else at_bound = 0;
    dv = vn - v;
    v = vn;
    if (fabs(dv) * xN <= 1e-12) break;
  }
  /* S0 at the final v */
  {
    double s0 = 0.0;
    for (int i = 0; i < N; ++i) s0 += resid[i] * resid[i] * exp(2.0 * v * (x[i] - xN));
    S0 = s0;
  }
  theta[0] = sqrt(S0 / N) * exp(v * xN);
  theta[1] = 1.0 / v;
  return 0;
}

/* estimateNoise for a batch; uy, ySmooth packed back to back in problem order; theta [n][2]; info [n][4]. */
int foct_oracle_estimate_noise(const foct_problem* P, int n, double df, double max_rate, double* uy, double* ySmooth,
                               double* theta, double* info) {
  size_t off = 0;
  for (int j = 0; j < n; ++j) {
    const int N = P[j].N;
    int rc = foct_oracle_smooth_spline(N, P[j].x, P[j].y, df, 0, NAN, ySmooth + off, info ? info + 4 * (size_t)j : NULL);
    if (rc) return rc;
    double* r = malloc(sizeof(double) * N);
    for (int i = 0; i < N; ++i) r[i] = P[j].y[i] - ySmooth[off + i];
    foct_oracle_noise_fit(N, P[j].x, r, max_rate, theta + 2 * (size_t)j);
    for (int i = 0; i < N; ++i) uy[off + i] = theta[2 * j] * exp(-P[j].x[i] / theta[2 * j + 1]);
    free(r);
    off += N;
  }
  return 0;
}

/* regularised lower incomplete gamma P(a, x) */
static double gammp(double a, double x) {
  if (x <= 0.0) return 0.0;
  const double gln = lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 10000; ++n) { ap += 1.0; del *= x / ap; sum += del; if (fabs(del) < fabs(sum) * 1e-17) break; }
    return sum * exp(-x + a * log(x) - gln);
  }
  double b = x + 1.0 - a, c = 1.0 / 1e-300, d = 1.0 / b, h = d;
  for (int i = 1; i < 10000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b; if (fabs(d) < 1e-300) d = 1e-300;
    c = b + an / c; if (fabs(c) < 1e-300) c = 1e-300;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < 1e-17) break;
  }
  return 1.0 - exp(-x + a * log(x) - gln) * h;
}

double foct_oracle_qchisq(double p, double ndf) {
  const double a = 0.5 * ndf;
  /* Newton on P(ndf/2, x/2) = p from the mean, safeguarded by bisection on a bracket */
  double lo = 0.0, hi = ndf + 40.0 * sqrt(2.0 * ndf) + 40.0;
  double x = ndf > 0 ? ndf : 1.0;
  for (int it = 0; it < 200; ++it) {
    const double f = gammp(a, 0.5 * x) - p;
    if (f > 0.0) hi = x; else lo = x;
    const double dens = 0.5 * exp(-0.5 * x + (a - 1.0) * log(0.5 * x) - lgamma(a));
    double xn = x - f / dens;
    if (!(xn > lo && xn < hi) || !isfinite(xn)) xn = 0.5 * (lo + hi);
    if (fabs(xn - x) <= 1e-14 * fabs(x)) { x = xn; break; }
    x = xn;
  }
  return x;
}

/* MODEL_SPEC §12.  ci[2] = qchisq({.025,.975}, ndf)/ndf; alert[j] = br[j] outside ci. */
int foct_oracle_print_br(const double* br, int n, double ndf, double ci[2], int* alert) {
  ci[0] = foct_oracle_qchisq(0.025, ndf) / ndf;
  ci[1] = foct_oracle_qchisq(0.975, ndf) / ndf;
  for (int j = 0; j < n; ++j) alert[j] = !(br[j] >= ci[0] && br[j] <= ci[1]);
  return 0;
}

static int cmp_d(const void* a, const void* b) { const double u = *(const double*)a, v = *(const double*)b; return (u > v) - (u < v); }

/* MODEL_SPEC §13.  priorType 0 = mono, 1 = abc.  theta_map [n][3], hessian [n][9] (of lp, negative definite).
 * Outputs theta0 [n][3], Sigma0 [n][9], ru [n] (the relative uncertainty used). */
int foct_oracle_exp_prior(const foct_problem* P, int n, int priorType, const double* theta_map, const double* hessian,
                          double ru_theta, double* theta0, double* Sigma0, double* ru_out) {
  for (int j = 0; j < n; ++j) {
    const double* th = theta_map + 3 * (size_t)j;
    const double* H = hessian + 9 * (size_t)j;
    double A[9], C[9];
    for (int k = 0; k < 9; ++k) A[k] = -H[k];
    const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
    C[0] = (A[4] * A[8] - A[5] * A[7]) / det; C[1] = (A[2] * A[7] - A[1] * A[8]) / det; C[2] = (A[1] * A[5] - A[2] * A[4]) / det;
    C[3] = (A[5] * A[6] - A[3] * A[8]) / det; C[4] = (A[0] * A[8] - A[2] * A[6]) / det; C[5] = (A[2] * A[3] - A[0] * A[5]) / det;
    C[6] = (A[3] * A[7] - A[4] * A[6]) / det; C[7] = (A[1] * A[6] - A[0] * A[7]) / det; C[8] = (A[0] * A[4] - A[1] * A[3]) / det;
    double sd[3], cor[9];
    for (int a = 0; a < 3; ++a) sd[a] = sqrt(C[a * 3 + a]);
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) cor[a * 3 + b] = a == b ? 1.0 : C[a * 3 + b] / (sd[a] * sd[b]);
    double ru = ru_theta;
    if (priorType == 1) {
      const int N = P[j].N;
      const double c = (double)P[j].dataType;
      double* ar = malloc(sizeof(double) * N);
      double sbar = 0.0;
      for (int i = 0; i < N; ++i) {
        const double t = c * P[j].x[i] / th[2], e = exp(-t);
        const double J[3] = {1.0, e, th[1] * e * t / th[2]};
        double v = 0.0;
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) v += J[a] * th[a] * cor[a * 3 + b] * th[b] * J[b];
        sbar += sqrt(v);
        ar[i] = fabs(P[j].y[i] - (th[0] + th[1] * e));
      }
      sbar /= N;
      qsort(ar, N, sizeof(double), cmp_d);
      const double h = 0.95 * (N - 1);
      const int lo = (int)floor(h);
      const double q95 = lo + 1 < N ? ar[lo] + (h - lo) * (ar[lo + 1] - ar[lo]) : ar[N - 1];
      ru = q95 / (1.96 * sbar);
      free(ar);
    }
    for (int a = 0; a < 3; ++a) {
      theta0[3 * (size_t)j + a] = th[a];
      for (int b = 0; b < 3; ++b) Sigma0[9 * (size_t)j + a * 3 + b] = (ru * th[a]) * cor[a * 3 + b] * (ru * th[b]);
    }
    if (ru_out) ru_out[j] = ru;
  }
  return 0;
}
