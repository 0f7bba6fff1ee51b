/*
 * foct_oracle.h — CPU restatement of the hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the reference's arithmetic for this path lives in FitOCTLib (Stan model blocks) and
 * rstan/Stan (NUTS), both external to /root/reference, unpinned and absent here; the
 * reference holds no golden vectors.  This oracle restates MODEL_SPEC.md (model) and Stan's published
 * diag_e NUTS + windowed adaptation (sampler).  It is checked against a 50-digit mpmath restatement
 * (oracle/mp_model.py, tests/golden/) and analytic targets, not against rstan.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call this.
 * The product (fitoct_b200/) never links or imports it.
 */
#ifndef FOCT_ORACLE_H
#define FOCT_ORACLE_H
#include "../include/fitoct_b200.h"
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Philox4x32-10 (Salmon et al. 2011): counter c[4], key k[2] -> out[4]. */
void foct_oracle_philox(const uint32_t c[4], const uint32_t k[2], uint32_t out[4]);
/* The two variates every draw site derives from one Philox block (MODEL_SPEC §7). */
void foct_oracle_uniform2(const uint32_t r[4], double u[2]);
double foct_oracle_normal(const uint32_t r[4]);

int foct_oracle_grid(int Nn, int gridType, double* xGP);
/* B_out [Nn][N] control-point major. */
int foct_oracle_basis(const foct_problem* P, const foct_model_spec* spec, double* B_out);

/* log density + analytic gradient (MODEL_SPEC §4-5).  B may be NULL (then built with foct_oracle_basis)
 * or a caller-supplied [Nn][N] basis, so that the kernel parity test can share one basis.
 * abs_terms (optional, [n_q][D]) receives sum_i |summand| of every gradient component: the conditioning
 * scale against which a 1e-12 relative tolerance is meaningful when a component cancels to ~0. */
int foct_oracle_logp_grad(int kind, const foct_problem* P, const foct_model_spec* spec, const double* B,
                          const double* q, int n_q, double* lp, double* grad, double* chi2,
                          double* abs_terms);

/* Batched CPU NUTS, one (profile, chain) per OpenMP thread.  Same result layout as foct_sample
 * (summary filled by foct_oracle_summary).  n_threads <= 0 => all cores. Returns threads used. */
int foct_oracle_sample(int kind, const foct_problem* P, int n_problems, const foct_model_spec* spec,
                       const foct_sampler_cfg* cfg, foct_result* R, int n_threads);

/* Sampler on a user-supplied diagonal-normal / exponential analytic target (tests of the NUTS logic):
 * target 0: independent normal, sd[d] given; target 1: Exponential(rate) on lambda>0 sampled on
 * q = log(lambda) with Jacobian (Tests/testGamma.R:19-30); target 2: zero-mean multivariate normal with dense
 * precision matrix par[D*D] (correlated: exercises the U-turn logic under a diagonal metric).
 * draws [n_saved][chains][D]. */
int foct_oracle_sample_analytic(int target, int D, const double* par, const foct_sampler_cfg* cfg,
                                double* draws, double* sampler_params);

/* rstan-style summary of draws[n_saved][chains][P] -> out[P][FOCT_N_SUMMARY_COLS] (MODEL_SPEC §9). */
int foct_oracle_summary(const double* draws, int n_saved, int chains, int P, double* out);

int foct_oracle_monoexp_map(const foct_problem* P, int n_problems, const foct_model_spec* spec,
                            const double* init, double* theta, double* hessian, double* br, int* status);

int foct_oracle_expgp_map(const foct_problem* P, int n_problems, const foct_model_spec* spec, const double* init,
                          double* par, double* hessian, int* status);

int foct_oracle_predict(int kind, const foct_problem* P, const foct_model_spec* spec, const double* draws,
                        int n_draws, double* m, double* resid, double* dL);

/* ---- steps either side of the path (foct_oracle_prep.c; MODEL_SPEC §11-13) ---- */
int foct_oracle_nknots(int n);
/* R-style penalised cubic regression spline.  info[4] = spar, lambda, df reached, evaluations.
 * spar_fixed = NaN: solve df(spar) = df; all_knots != 0: every x is a knot (natural smoothing spline). */
int foct_oracle_smooth_spline(int N, const double* x, const double* y, double df, int all_knots, double spar_fixed,
                              double* ySmooth, double* info);
int foct_oracle_noise_fit(int N, const double* x, const double* resid, double max_rate, double theta[2]);
int foct_oracle_estimate_noise(const foct_problem* P, int n, double df, double max_rate, double* uy, double* ySmooth,
                               double* theta, double* info);
double foct_oracle_qchisq(double p, double ndf);
int foct_oracle_print_br(const double* br, int n, double ndf, double ci[2], int* alert);
int foct_oracle_exp_prior(const foct_problem* P, int n, int priorType, const double* theta_map, const double* hessian,
                          double ru_theta, double* theta0, double* Sigma0, double* ru_out);

/* ---- method = 'vb': Stan's mean-field ADVI (foct_oracle_vb.c; MODEL_SPEC §14) ---- */
int foct_oracle_vb(int kind, const foct_problem* P, int n_problems, const foct_model_spec* spec, const foct_vb_cfg* cfg,
                   foct_vb_result* R);
int foct_oracle_vb_analytic(int target, int D, const double* par, const foct_vb_cfg* cfg, const double* q0, double* mu,
                            double* omega, double* info);

#ifdef __cplusplus
}
#endif
#endif
