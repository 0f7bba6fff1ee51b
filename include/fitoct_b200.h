/*
 * fitoct_b200.h — C ABI of the B200-native batched NUTS engine for the FitOCT decay models.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / R / C++ types.  Every entry point
 * cites the reference interface it replaces (paths are relative to the upstream rbocheux/FitOCT tree).
 * The mathematics is frozen in MODEL_SPEC.md; the reference-side binding (R `.Call` shim) is shown in
 * INTEGRATION.md and r-pkg/.
 *
 * Conventions
 *   - all floating point data is IEEE fp64, row-major, caller-owned; the library allocates nothing the
 *     caller must free except opaque `foct_plan` handles (freed with foct_plan_destroy);
 *   - every function returns 0 on success, a negative FOCT_E* code on failure; the message is available
 *     from foct_last_error() (thread-local);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with FOCT_ENODEV.
 */
#ifndef FITOCT_B200_H
#define FITOCT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FOCT_ABI_VERSION 2 /* 2: continuation / run-until-converged fields at the END of foct_sampler_cfg and foct_result */

#define FOCT_MAX_D 32        /* unconstrained dimensions live one-per-lane in a warp */
#define FOCT_MAX_NN 25       /* => D = Nn + 5 <= 30, P_out = Nn + 7 <= 32 (reference UI range is 5..20, ShinyInterface/ui.R:200-207) */
#define FOCT_MAX_CHAINS 8
#define FOCT_N_SAMPLER_PARAMS 6 /* accept_stat__, stepsize__, treedepth__, n_leapfrog__, divergent__, energy__ */
#define FOCT_N_SUMMARY_COLS 11  /* mean, se_mean, sd, 2.5%, 25%, 50%, 75%, 97.5%, n_eff, Rhat (rstan::summary), Bulk_ESS */

/* error codes */
#define FOCT_OK 0
#define FOCT_EINVAL (-1)  /* bad argument */
#define FOCT_ENODEV (-2)  /* no CUDA device / driver: the product has no CPU path */
#define FOCT_ECUDA (-3)   /* CUDA runtime error */
#define FOCT_ENOMEM (-4)
#define FOCT_ECANCELLED (-5) /* the caller asked the run to stop (foct_plan_cancel, a progress callback returning non-zero) */

/* model kinds */
#define FOCT_EXPGP 0   /* replaces FitOCTLib::fitExpGP   (FitOCT.R:110-124, priPost.R:2-16, server.R:408-426) */
#define FOCT_MONOEXP 1 /* replaces FitOCTLib::fitMonoExp (FitOCT.R:95, server.R:341-343) */

/* gridType (ShinyInterface/ui.R:211-218, server.R:626-635) */
#define FOCT_GRID_INTERNAL 0
#define FOCT_GRID_EXTREMAL 1

/* MODEL_SPEC.md §8 switch table.  Defaults from foct_model_spec_default(). */
typedef struct foct_model_spec {
  int modulation;    /* 0 = length (synthData.R:22), 1 = amplitude */
  int kernel;        /* 0 = stan_exp_quad exp(-d^2/(2 rho^2)), 1 = rmgauss exp(-d^2/rho^2) (server.R:648) */
  double jitter;     /* added to diag(Kgg); default 1e-9 */
  int ygp_prior;     /* 0 = normal(0, lambda), 1 = laplace(0, lambda) (Tests/lassoPrior.stan) */
  int lambda_prior;  /* 0 = gamma(2, lambda_rate), 1 = exponential(lambda_rate) (Tests/testGamma.R:27) */
  double sigma_mean; /* prior on the noise factor sigma: normal(sigma_mean, sigma_sd); sigma_sd <= 0 => flat */
  double sigma_sd;
  int theta_prior;   /* 0 = multi_normal(theta0, Sigma0), 1 = flat */
  int br_ndf;        /* 0 = N - n_params_of_mean, 1 = N */
} foct_model_spec;

/* One depth profile + the knobs the reference passes to fitExpGP (FitOCT.R:110-124). */
typedef struct foct_problem {
  int N;            /* points in this profile (ragged batches allowed) */
  const double* x;  /* depth                         (FitOCT.R:84-86) */
  const double* y;  /* signal */
  const double* uy; /* signal uncertainty from estimateNoise (FitOCT.R:89-90) */
  int dataType;     /* 1 amplitude / 2 intensity     (FitOCT.R:39,112) */
  int Nn;           /* control points; must be uniform over a batch; ignored for FOCT_MONOEXP */
  int gridType;     /* FOCT_GRID_*                   (FitOCT.R:114) */
  double rho;       /* GP length scale, already resolved by the caller (0 => 1/Nn done in R, FitOCT.R:119) */
  double lambda_rate;
  double theta0[3]; /* prior mean                    (FitOCT.R:116) */
  double Sigma0[9]; /* prior covariance, row-major   (FitOCT.R:117) */
  int prior_PD;     /* 1 => likelihood switched off  (FitOCT.R:122, priPost.R:14) */
  long long id;     /* RNG stream id of this profile: results do not depend on batch order or GPU count */
} foct_problem;

/* Sampler controls.  The reference exposes only nb_warmup / nb_iter (FitOCT.R:120-121); the rest are the
 * rstan `sampling(control=...)` knobs FitOCTLib hard-codes (unknown, SURVEY a-7) — rstan defaults here. */
typedef struct foct_sampler_cfg {
  int chains;              /* default 4 (server.R:469) */
  int n_warmup;            /* nb_warmup */
  int n_iter;              /* nb_iter = nb_warmup + nb_sample, as FitOCT.R:121 passes it */
  double adapt_delta;      /* 0.8 */
  int max_treedepth;       /* 10 */
  double stepsize0;        /* 1.0 */
  unsigned long long seed; /* FitOCT.R:15 seeds R; rstan then draws its seed from R's RNG */
  int init_mode;           /* MODEL_SPEC §7: 0 prior-centred, 1 U(-2,2), 2 caller-supplied */
  const double* init;      /* [n_problems][chains][D] unconstrained, only for init_mode == 2 */
  int save_warmup;         /* 1 => draws include warm-up (plotExpGP.R:46 traceplot(inc_warmup=TRUE)) */
  /* adaptation constants (0 => Stan defaults 0.05 / 0.75 / 10 / 75 / 50 / 25) */
  double gamma, kappa, t0;
  int init_buffer, term_buffer, window;
  /* devices to shard profiles over (independent shards, no collective); n_devices == 0 => current device */
  int n_devices;
  const int* devices;
  /* ---- ABI 2: continuation.  rstan cannot continue a stanfit; the reference reaches convergence by raising nb_warmup /
   * nb_sample and refitting (FitOCT.R:43-44).  Here a run can start from an adapted sampler instead of adapting one:
   * n_warmup = 0, init_mode = 2 with init = foct_result.last_q of the earlier run, inv_metric_init = its inv_metric,
   * stepsize_init = its stepsize, iter_offset = the iterations it did (keeps the Philox sites of the two runs apart). */
  const double* inv_metric_init; /* [n_problems][chains][D], NULL => unit metric */
  const double* stepsize_init;   /* [n_problems][chains], NULL => stepsize0 */
  int iter_offset;
  /* Run until converged (BASELINE north_star: "sampled to R-hat < 1.01"): after the n_iter iterations, every profile
   * whose largest split R-hat over the sampled parameters is >= rhat_target is continued on the device — same adapted
   * metric and step size, no new warm-up — for another extend_iter draws, at most max_extend times.  The summary then
   * covers ALL post-warm-up draws of the profile; the returned post-warm-up draws / sampler_params keep their shape and
   * are thinned evenly out of them (row t <- draw floor((t + 1) total / n_post) - 1).  rhat_target <= 0: off.  Needs
   * the summary output. */
  double rhat_target;
  int max_extend;
  int extend_iter; /* draws per continuation round; 0 => a quarter of n_iter - n_warmup (at least 50) */
} foct_sampler_cfg;

/* Caller-allocated outputs; any pointer may be NULL to skip that output.
 * n_saved = save_warmup ? n_iter : n_iter - n_warmup. */
typedef struct foct_result {
  double* draws;          /* [n_problems][n_saved][chains][P_out]  constrained, MODEL_SPEC §6 column order */
  double* sampler_params; /* [n_problems][n_saved][chains][6] */
  double* summary;        /* [n_problems][P_out][FOCT_N_SUMMARY_COLS] over post-warm-up draws, all chains */
  double* stepsize;       /* [n_problems][chains] adapted step size */
  double* inv_metric;     /* [n_problems][chains][D] adapted diagonal inverse metric */
  double* n_leapfrog;     /* [n_problems][chains][2] total leapfrog steps: {warm-up, sampling} */
  double* n_divergent;    /* [n_problems][chains] post-warm-up divergences */
  /* ---- ABI 2 */
  double* last_q;         /* [n_problems][chains][D] unconstrained state after the last transition (continuation) */
  int* n_extend;          /* [n_problems] continuation rounds the profile received (rhat_target) */
} foct_result;

int foct_version(void);
int foct_device_count(void);
const char* foct_last_error(void);

void foct_model_spec_default(foct_model_spec* spec, int kind);
void foct_sampler_cfg_default(foct_sampler_cfg* cfg);
/* D (unconstrained) and P_out (output columns) for a model kind. */
int foct_dims(int kind, int Nn, int* D, int* P_out);

/* Control-point grid `xGP` (returned as fitOut$xGP, plotExpGP.R:31; definition server.R:626-635).
 * Pure host arithmetic, no device needed. */
int foct_expgp_grid(int Nn, int gridType, double* xGP /*[Nn]*/);

/* GP conditional-mean basis B = K(xp,xGP) K(xGP,xGP)^-1, built on the device (MODEL_SPEC §1).
 * B_out is [Nn][N] (control-point major, the layout the kernels stage in shared memory). */
int foct_expgp_basis(const foct_problem* P, const foct_model_spec* spec, double* B_out);

/* Parity hook: log density and analytic gradient on the unconstrained space (MODEL_SPEC §4-5) for
 * n_q points per problem.  q, grad: [n_problems][n_q][D]; lp, chi2: [n_problems][n_q]
 * (chi2 = sum((y-m)/uy)^2, the numerator of br; may be NULL).  Replaces the generated Stan model class'
 * log_prob + reverse-mode gradient (SURVEY a-4, a-5). */
int foct_logp_grad(int kind, const foct_problem* P, int n_problems, const foct_model_spec* spec,
                   const double* q, int n_q, double* lp, double* grad, double* chi2);
/* The same for kind = FOCT_EXPGP under the name SURVEY §8(b) gives the parity hook. */
int foct_expgp_logp_grad(const foct_problem* P, int n_problems, const foct_model_spec* spec, const double* q, int n_q,
                         double* lp, double* grad, double* chi2);

/* One-shot batched NUTS: upload, sample on device, (optionally) summarise on device, download.
 * Replaces FitOCTLib::fitExpGP(method='sample') -> rstan::sampling (FitOCT.R:110-124) for a whole batch. */
int foct_sample(int kind, const foct_problem* P, int n_problems, const foct_model_spec* spec,
                const foct_sampler_cfg* cfg, foct_result* R);
/* The same call for a host that must stay responsive (the R shim polls, prints rstan-style progress lines
 * for ShinyInterface/server.R:457-472 and calls R_CheckUserInterrupt between polls).  The sampling runs on worker
 * threads, one per device; `progress` is called ON THE CALLING THREAD every poll_ms milliseconds (and once at the end)
 * with the fraction of chain-iterations completed and the phase ("Warmup" until every chain can have left warm-up,
 * then "Sampling", then "Extending" during the rhat_target rounds).  A non-zero return cancels the run: every chain
 * stops at its next iteration boundary, all buffers are released, and the call returns FOCT_ECANCELLED. */
typedef int (*foct_progress_fn)(double fraction, const char* phase, void* user);
int foct_sample_cb(int kind, const foct_problem* P, int n_problems, const foct_model_spec* spec,
                   const foct_sampler_cfg* cfg, foct_result* R, foct_progress_fn progress, void* user, int poll_ms);
/* Named as SURVEY §8(b) lists them. */
int foct_expgp_sample(const foct_problem* P, int n_problems, const foct_model_spec* spec,
                      const foct_sampler_cfg* cfg, foct_result* R);
int foct_monoexp_sample(const foct_problem* P, int n_problems, const foct_model_spec* spec,
                        const foct_sampler_cfg* cfg, foct_result* R);

/* MAP fit of the mono-exponential (the reference's only use of fitMonoExp: FitOCT.R:94-97,
 * plotMonoExp.R:14-16).  theta: [n][3]; hessian: [n][9] of lp at the optimum (negative definite, as rstan::optimizing(hessian=TRUE)); br: [n];
 * status: [n] (0 converged).  init may be NULL (then a log-linear start is used). */
int foct_monoexp_map(const foct_problem* P, int n_problems, const foct_model_spec* spec,
                     const double* init /*[n][3] or NULL*/, double* theta, double* hessian, double* br,
                     int* status);

/* MAP of fitExpGP: `method = 'optim'`, the Shiny default (ShinyInterface/ui.R:107-114; FitOCT.R:42; consumers
 * plotExpGP.R:13-17 read fit$par$theta / yGP / lambda / sigma / br, server.R:164-173 reads fit$hessian).  BFGS on
 * the unconstrained space without Jacobian terms (rstan::optimizing's default), then a finite-difference Hessian
 * of lp.  par: [n][P_out] constrained optimum in the MODEL_SPEC §6 column order (last column: lp at the optimum);
 * hessian: [n][D][D] (may be NULL); init: [n][D] unconstrained or NULL; status: 0 converged, 1 iteration limit,
 * 2 line search failed. */
int foct_expgp_map(const foct_problem* P, int n_problems, const foct_model_spec* spec, const double* init,
                   double* par, double* hessian, int* status);

/* Generated quantities for selected draws (SURVEY a-6): m, resid, dL at every depth.
 * draws: [n_draws][P_out] constrained rows of one problem; outputs [n_draws][N] (any may be NULL). */
int foct_predict(int kind, const foct_problem* P, const foct_model_spec* spec, const double* draws,
                 int n_draws, double* m, double* resid, double* dL);

/* rstan's summary(fit)$summary / monitor() for caller-supplied draws (plotExpGP.R:9-11, server.R:88-104 read it):
 * mean, se_mean, sd, 2.5/25/50/75/97.5 % quantiles, n_eff (Stan's Geyer estimator with rstan's tau clamp), split R-hat,
 * rank-normalised Bulk_ESS — the same kernel that summarises a fit inside foct_sample.
 * draws: [n_sets][n_draws][chains][n_cols] (the layout of foct_result.draws); summary: [n_sets][n_cols][FOCT_N_SUMMARY_COLS].
 * n_draws >= 4, 1 <= chains <= FOCT_MAX_CHAINS. */
int foct_summary(const double* draws, int n_sets, int n_draws, int chains, int n_cols, double* summary);

/* Plan API: the same path with device-resident inputs, for repeated runs and for timing the
 * sampling step without host<->device copies (bench.py `value`; `e2e` uses foct_sample). */
typedef struct foct_plan foct_plan;
int foct_plan_create(int kind, const foct_problem* P, int n_problems, const foct_model_spec* spec,
                     const foct_sampler_cfg* cfg, int want_draws, int want_summary, foct_plan** plan);
int foct_plan_run(foct_plan* plan, unsigned long long seed); /* async on the plan's stream */
int foct_plan_sync(foct_plan* plan, float* kernel_ms /* CUDA-event time of the sampling kernel, may be NULL */);
int foct_plan_fetch(foct_plan* plan, foct_result* R);
/* Non-blocking: *done = 1 once the kernels of the last foct_plan_run have finished; *fraction = chain-iterations
 * completed / total (either pointer may be NULL).  foct_plan_cancel asks the running kernels to stop at the next
 * iteration boundary of every chain; foct_plan_sync / foct_plan_fetch then return FOCT_ECANCELLED. */
int foct_plan_query(foct_plan* plan, int* done, double* fraction);
int foct_plan_cancel(foct_plan* plan);
/* CUDA-event durations of the last run (sampling kernel; summary kernel) and the launch geometry of the
 * sampling kernel — what bench.py's roofline is computed from.  Any pointer may be NULL. */
int foct_plan_timing(foct_plan* plan, float* sample_ms, float* summary_ms, int* grid, int* block,
                     int* blocks_per_sm, int* regs, int* smem_bytes);
/* Kernels the last foct_plan_run launched so far (sampling, summary, and those of the continuation rounds). */
int foct_plan_launches(foct_plan* plan);
void foct_plan_destroy(foct_plan* plan);

/* ---- the steps either side of the sampling path (MODEL_SPEC §11-13) ---------------------------
 * Per-point outputs of a batch are packed back to back in problem order (problem j starts at sum_{i<j} N_i). */

/* FitOCTLib::estimateNoise(x, y, df) (FitOCT.R:89-91, server.R:309-311): R-style smooth.spline at `df` equivalent
 * degrees of freedom, then the MLE of resid_i ~ N(0, a_1 exp(-x_i/a_2)).  P[j].uy is ignored.  theta: [n][2] = a_1, a_2
 * (plotNoise.R:4-6); info: [n][4] = spar, lambda, df reached, spline evaluations (may be NULL); status: [n] 0 ok,
 * 1 requested df not reachable on spar in [-1.5, 1.5], 3 x not strictly increasing (may be NULL). */
int foct_estimate_noise(const foct_problem* P, int n_problems, double df, double max_rate, double* uy, double* ySmooth,
                        double* theta, double* info, int* status);

/* 95 % interval of the Birge ratio, qchisq((.025,.975), ndf)/ndf.  Scalar host arithmetic, no device needed. */
int foct_birge_ci(double ndf, double* ci /*[2]*/);

/* FitOCTLib::printBr (plotMonoExp.R:10, plotExpGP.R:22; the gate at FitOCT.R:100): alert[j] = 1 iff br[j] lies outside
 * the interval for ndf = N_j - n_par (spec->br_ndf as in MODEL_SPEC §6).  ci: [n][2] (may be NULL). */
int foct_print_br(int kind, const foct_problem* P, int n_problems, const foct_model_spec* spec, const double* br,
                  double* ci, int* alert);

/* FitOCTLib::estimateExpPrior(x, uy, dataType, priorType, out = fitMonoExp result, ru_theta, eps) (FitOCT.R:103-107,
 * server.R:396-404).  theta_map: [n][3] and hessian: [n][9] as returned by foct_monoexp_map; outputs theta0 [n][3],
 * Sigma0 [n][9], ru [n] (the relative uncertainty applied; may be NULL). */
#define FOCT_PRIOR_MONO 0
#define FOCT_PRIOR_ABC 1
int foct_estimate_exp_prior(const foct_problem* P, int n_problems, int priorType, const double* theta_map,
                            const double* hessian, double ru_theta, double* theta0, double* Sigma0, double* ru);

/* The body of FitOCT.R's dataset loop (FitOCT.R:84-124) for a whole batch in one call: estimateNoise -> fitMonoExp
 * (MAP) -> printBr gate -> estimateExpPrior -> fitExpGP(method = 'sample') on the profiles the gate lets through.
 * The keys are ctrlParams.yaml's (FitOCT.R:37-53). */
typedef struct foct_pipeline_cfg {
  double smooth_df;   /* 15 */
  double max_rate;    /* 1e4 */
  int prior_type;     /* FOCT_PRIOR_ABC */
  double ru_theta;    /* 0.05 */
  int Nn;             /* 10 */
  int gridType;       /* FOCT_GRID_INTERNAL */
  double rho_scale;   /* 0 => 1/Nn (FitOCT.R:119) */
  double lambda_rate; /* 0.1 */
  int gate;           /* 1: only profiles with a Birge-ratio alert go on to fitExpGP (FitOCT.R:100); 0: all do */
} foct_pipeline_cfg;
void foct_pipeline_cfg_default(foct_pipeline_cfg* c);

typedef struct foct_pipeline_out {
  double* uy;            /* packed [sum N] */
  double* ySmooth;       /* packed [sum N] */
  double* noise_theta;   /* [n][2] */
  double* mono_theta;    /* [n][3] */
  double* mono_hessian;  /* [n][9] */
  double* mono_br;       /* [n] */
  int* mono_status;      /* [n] or NULL */
  double* br_ci;         /* [n][2] or NULL */
  int* alert;            /* [n] */
  double* theta0;        /* [n][3] */
  double* Sigma0;        /* [n][9] */
  double* ru;            /* [n] or NULL */
  int n_expgp;           /* out: number of profiles passed to fitExpGP */
  int* expgp_index;      /* [n]: profile index of ExpGP fit k, k < n_expgp */
  foct_result expgp;     /* buffers sized for n problems by the caller; the first n_expgp entries are filled */
  /* ABI 2: what happened to profile j ([n], may be NULL).  A failed or degenerate MonoExp fit (constant signal, singular
   * Hessian) is exactly what the Birge-ratio gate forwards to fitExpGP, and one such profile must not abort a directory
   * of 1e5: its prior is repaired or the profile is skipped, and the call goes on. */
  int* status;           /* FOCT_PIPE_* */
} foct_pipeline_out;
#define FOCT_PIPE_SAMPLED 0        /* fitExpGP ran with estimateExpPrior's prior */
#define FOCT_PIPE_GATED 1          /* MonoExp fit OK (printBr raised no alert): fitExpGP not needed (FitOCT.R:100) */
#define FOCT_PIPE_PRIOR_REPAIRED 2 /* Sigma0 was not finite / positive definite: fitExpGP ran with the diagonal prior
                                      diag((ru_theta theta_MAP)^2) (priorType 'mono' without the Hessian's correlations) */
#define FOCT_PIPE_SKIPPED 3        /* no usable MonoExp fit (non-finite theta or uy): nothing to centre a prior on */
int foct_pipeline(const foct_problem* P, int n_problems, const foct_pipeline_cfg* pc, const foct_model_spec* spec_gp,
                  const foct_sampler_cfg* cfg, foct_pipeline_out* out);

/* ---- method = 'vb' (FitOCT.R:42; rstan::vb): Stan's mean-field ADVI, MODEL_SPEC §14 ------------------------------ */
typedef struct foct_vb_cfg {
  int iter;            /* 10000 */
  int grad_samples;    /* 1 */
  int elbo_samples;    /* 100 */
  int eval_elbo;       /* 100 */
  int output_samples;  /* 1000 */
  int adapt_engaged;   /* 1 */
  int adapt_iter;      /* 50 */
  double eta;          /* 1.0; used as is when adapt_engaged = 0 */
  double tol_rel_obj;  /* 0.01 */
  unsigned long long seed;
  int init_mode;       /* as foct_sampler_cfg.init_mode */
  const double* init;  /* [n_problems][D] unconstrained, for init_mode 2 */
  double omega0;       /* initial log standard deviation of every component; Stan: 0 (MODEL_SPEC §14) */
} foct_vb_cfg;
void foct_vb_cfg_default(foct_vb_cfg* cfg);

typedef struct foct_vb_result { /* caller-allocated; mean, mu, omega required, the rest may be NULL */
  double* mean;   /* [n][P_out] constrained mean of the approximation, br at the mean, lp__ = 0 (Stan's first CSV row) */
  double* draws;  /* [n][output_samples][P_out] */
  double* mu;     /* [n][D] */
  double* omega;  /* [n][D] log standard deviations */
  double* elbo;   /* [n] last ELBO estimate */
  double* eta;    /* [n] step-size scale used */
  int* iters;     /* [n] main-loop iterations done */
  int* status;    /* [n] 0 converged, 1 iteration limit, 2 failed (non-finite gradient / ELBO, or no usable step size) */
} foct_vb_result;
int foct_vb(int kind, const foct_problem* P, int n_problems, const foct_model_spec* spec, const foct_vb_cfg* cfg,
            foct_vb_result* R);

/* Device buffers freed by the entry points above are cached (up to FOCT_POOL_MB, default 8192 MB) and reused by later
 * calls on the same device: an R session calls the fit once per profile.  This returns the cache to the driver. */
void foct_release_cache(void);

/* Measured fp64 FMA throughput of the device (DFMA-chain microbenchmark), the roofline denominator for
 * the sampling kernel (MEASURED_PEAKS.json has no fp64 figure). */
int foct_fp64_peak(int device, double* tflops, double* sm_mhz);

#ifdef __cplusplus
}
#endif
#endif /* FITOCT_B200_H */
