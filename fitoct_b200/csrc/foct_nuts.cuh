// foct_nuts.cuh — on-device multinomial NUTS with Stan-style windowed adaptation, one warp per chain.
//
// Restates Stan's published diag_e NUTS (MODEL_SPEC §7; SURVEY a-7..a-10) in ITERATIVE form: the recursive
// build_tree becomes a leaf loop with an O(depth) stack of pending "init" subtrees kept in per-lane local
// memory; merges, the three U-turn checks per merge, multinomial proposal selection and the random-number
// draw sites are exactly those of the recursion, so the CPU oracle (recursive) and this kernel consume the
// same Philox variates and build the same trees.
#pragma once
#include "foct_device.cuh"

namespace foct {

#ifndef FOCT_PARK
#define FOCT_PARK volatile
#endif

struct SamplerParams {
  const double* blobs;      // [n_problems][blob_stride] staged profile blobs
  size_t blob_stride;       // doubles per blob = (3 + NN) * npad
  int npad;
  const DevProblem* probs;
  int n_problems;
  DevSpec spec;
  int chains, n_warmup, n_iter, max_depth, save_warmup, init_mode;
  double adapt_delta, stepsize0, gamma, kappa, t0;
  int init_buffer, term_buffer, window;
  unsigned long long seed;
  const double* init;       // [n_problems][chains][D] or nullptr
  double* draws;            // [n_problems][n_saved][chains][P_out] or nullptr
  double* sparams;          // [n_problems][n_saved][chains][6] or nullptr
  double* stepsize;         // [n_problems][chains]
  double* inv_metric;       // [n_problems][chains][D]
  double* n_leapfrog;       // [n_problems][chains][2]
  double* n_divergent;      // [n_problems][chains]
  int* work_counter;        // dynamic profile scheduler
  const int* order;         // [n_problems] profile served by work item w (longest expected first), or nullptr
  // ---- continuation of an earlier run (foct_sampler_cfg.inv_metric_init / stepsize_init / iter_offset; the
  //      run-until-converged rounds of foct_plan_run): adapted state in, last state out
  const double* invm_init;  // [..][chains][D] or nullptr (unit metric)
  const double* eps_init;   // [..][chains] or nullptr (stepsize0)
  double* last_q;           // [..][chains][D] unconstrained state after the last transition, or nullptr
  int it_offset;            // added to the iteration index of every Philox site
  int accumulate;           // 1: n_leapfrog / n_divergent are added to what the buffers hold
  int save_stride;          // saved iterations a profile's draw block holds (0: n_saved of this launch)
  int save_offset;          // first row of this launch inside the block
  const int* slot_of;       // [..] draw block of profile j (nullptr: j)
  // ---- progress and cancellation (foct_plan_query / foct_plan_cancel / foct_sample_cb; SURVEY §8b: R polls and calls
  //      R_CheckUserInterrupt between polls, server.R:457-472 scrapes the progress)
  int shared_basis;              // every profile has the same depth grid (and dataType): the basis rows - in nuts2w_kernel c x too - are read through L1 from blob 0
  unsigned long long* progress;  // chain-iterations completed so far, or nullptr
  const int* cancel;             // != 0: every chain stops at its next iteration boundary, or nullptr
  // ---- time slicing of the work items (nuts2_kernel; all nullptr / 0: off)
  double* slice_state;              // [n_items][warps per CTA][FOCT_PAIR_STATE_DOUBLES][32]
  int* slice_done;                  // [n_items][warps per CTA] != 0: the chains of this warp have finished
  unsigned long long* slice_queue;  // [n_items] ring of suspended items, (ticket << 32 | item)
  unsigned* slice_ctl;              // tickets handed out | pushes | items finished
  int* slice_hist;                  // [FOCT_PROGRESS_BINS] waiting units by iterations done (nuts2w_kernel), or nullptr
  int slice_ticks;                  // gradient evaluations per slice
  int pair_kernel;                  // host: 1 = nuts2_kernel (two chains per warp), 0 = nuts_kernel
  int warp_units;                   // host: > 0 = nuts2w_kernel (warps claim (profile, chain pair) units on their own), the warps of a CTA
};

#define FOCT_PROGRESS_EVERY 8
__device__ __forceinline__ bool cancel_requested(const SamplerParams& K) {
  return K.cancel && *reinterpret_cast<const volatile int*>(K.cancel) != 0;
}
__device__ __forceinline__ void report_progress(const SamplerParams& K, int it_done, bool last, bool leader) {
  // one atomic per FOCT_PROGRESS_EVERY iterations of a chain, plus the remainder when the chain ends
  if (!K.progress || !leader) return;
  if (last) { const int r = it_done % FOCT_PROGRESS_EVERY; if (r) atomicAdd(K.progress, (unsigned long long)r); }
  else if (it_done % FOCT_PROGRESS_EVERY == 0) atomicAdd(K.progress, (unsigned long long)FOCT_PROGRESS_EVERY);
}

// Row of (profile, saved iteration, chain) in draws / sampler_params.
__device__ __forceinline__ size_t save_row(const SamplerParams& K, int prob, int n_saved, int save_idx, int chain) {
  const int stride = K.save_stride > 0 ? K.save_stride : n_saved;
  const int slot = K.slot_of ? K.slot_of[prob] : prob;
  return ((size_t)slot * stride + K.save_offset + save_idx) * K.chains + chain;
}

// log(exp(a) + exp(b)) and exp(b - lse) = w_b / (w_a + w_b) from ONE exponential (Stan computes
// log_sum_exp and then exp(lsw_final - lsw_subtree); same quantities).
__device__ __forceinline__ double lse_prob(double a, double b, double& prob_b) {
  const double d = b - a;
  const double ad = fabs(d);
  const double e = fexp(-ad);
  double lse = fmax(a, b) + flog1p_unit(e);
  prob_b = (d >= 0.0 ? 1.0 : e) * frcp(1.0 + e);  // 1 + e in [1, 2]: the branch-free reciprocal is exact enough
  if (!(ad >= 0.0)) { lse = -CUDART_INF; prob_b = 0.0; }  // both weights zero
  return lse;
}

// The three generalised U-turn checks of one merge (init ++ final), each a pair of dot products:
//   around the merged tree, init + first state of final, final + last state of init.
__device__ __forceinline__ bool merge_persists(double invM, double i_rho, double i_pbeg, double i_pend, double f_rho,
                                               double f_pbeg, double f_pend, int lane) {
  const double ps_b = invM * i_pbeg, ps_e = invM * f_pend, ps_fb = invM * f_pbeg, ps_ie = invM * i_pend;
  const double r_sub = i_rho + f_rho, r_b = i_rho + f_pbeg, r_c = f_rho + i_pend;
  double v[8];
  v[0] = ps_e * r_sub; v[1] = ps_b * r_sub;
  v[2] = ps_fb * r_b;  v[3] = ps_b * r_b;
  v[4] = ps_e * r_c;   v[5] = ps_ie * r_c;
  v[6] = lane == 0 ? 1.0 : 0.0; v[7] = v[6];
  const double s = warp_reduce_scatter<8>(v, lane);
  return __all_sync(FOCT_FULL, s > 0.0);
}

// One leapfrog step of the integrator state (q, p, g, V, chi2) of this lane.  Inlined: a separate (noinline)
// function with its own register allocation was measured 8 % slower (call ABI + state passed through local memory).
#ifndef FOCT_LEAPFROG_INLINE
#define FOCT_LEAPFROG_INLINE __forceinline__
#endif
template <int NN, int MOD, int TEAM = 1>
__device__ FOCT_LEAPFROG_INLINE void leapfrog(const double* __restrict__ blob, const DevProblem* P, const DevSpec* S,
                                      double eps, double invM, double* zq, double* zp, double* zg, double* zV,
                                      double* zc2, int lane, TeamCtx* tc = nullptr) {
  double p = fma(0.5 * eps, *zg, *zp);
  double q = fma(eps * invM, p, *zq);
  const Eval ev = warp_logp_grad<NN, MOD, 32, 0, TEAM>(blob, *P, *S, q, lane, nullptr, tc);
  p = fma(0.5 * eps, ev.g, p);
  *zq = q; *zp = p; *zg = ev.g; *zV = -ev.lp; *zc2 = ev.chi2;
}

// TEAM = 2: the chain is run by two warps (see TeamCtx in foct_device.cuh); `writer` is the member that stores results.
template <int NN, int MOD, int TEAM = 1>
__device__ void run_chain(const SamplerParams& K, const DevProblem& P, const double* __restrict__ blob, int prob,
                          int chain, int lane, TeamCtx* tc = nullptr) {
  const bool writer = TEAM == 1 || tc->member == 0;
  using DM = Dims<NN>;
  constexpr int D = DM::D;
  constexpr int P_OUT = DM::P_OUT;
  const bool act = lane < D;
  Rng rng;
  rng.seed(K.seed, P.id, chain);
  uint32_t rb[4];

  // ---- initial point (MODEL_SPEC §7 init_mode)
  double q = 0.0;
  if (act) {
    rng.block(0, SITE_INIT, 0, (uint32_t)lane, 0, rb);
    if (K.init_mode == 2 && K.init) {
      q = K.init[((size_t)prob * K.chains + chain) * D + lane];
    } else if (K.init_mode == 1) {
      q = -2.0 + 4.0 * u53(rb[0], rb[1]);
    } else {
      if (lane < 3) q = P.theta0[lane];
      else if (lane < 3 + NN) q = 0.01 * normal_from(rb);
      else if (lane == 3 + NN) q = log(0.1);
      else q = 0.0;
    }
  }
  Eval ev = warp_logp_grad<NN, MOD, 32, 0, TEAM>(blob, P, K.spec, q, lane, nullptr, tc);
  double g = ev.g, V = -ev.lp, c2 = ev.chi2;
  double invM = 1.0;
  double eps = K.stepsize0 > 0.0 ? K.stepsize0 : 1.0;
  if (K.invm_init && act) invM = K.invm_init[((size_t)prob * K.chains + chain) * D + lane];
  if (K.eps_init) eps = K.eps_init[(size_t)prob * K.chains + chain];
  const uint32_t it0 = (uint32_t)K.it_offset;

  // ---- adaptation state (Stan windowed_adaptation / welford_var_estimator / stepsize_adaptation)
  int a_num_warmup, a_init_buffer, a_term_buffer, a_base_window;
  {
    int ib = K.init_buffer > 0 ? K.init_buffer : 75, tb = K.term_buffer > 0 ? K.term_buffer : 50;
    int bw = K.window > 0 ? K.window : 25, nw = K.n_warmup;
    if (nw < 20) {
      a_num_warmup = a_init_buffer = a_term_buffer = a_base_window = 0;
    } else {
      if (ib + bw + tb > nw) { ib = (int)(0.15 * nw); tb = (int)(0.1 * nw); bw = nw - (ib + tb); }
      a_num_warmup = nw; a_init_buffer = ib; a_term_buffer = tb; a_base_window = bw;
    }
  }
  // State that is touched once per iteration or per doubling, never per leaf, is declared volatile: it then lives in
  // (L1-resident) local memory instead of occupying registers across the sweep, where register pressure decides whether
  // ptxas keeps the two points of an iteration interleaved (profiles/r1_microbench.txt, in-situ ablations).
  FOCT_PARK int a_counter = 0, a_wsize = a_base_window, a_next = a_init_buffer + a_base_window - 1;
  FOCT_PARK double w_n = 0.0, w_mean = 0.0, w_m2 = 0.0;
  const double da_delta = K.adapt_delta > 0.0 ? K.adapt_delta : 0.8;
  const double da_gamma = K.gamma > 0.0 ? K.gamma : 0.05, da_kappa = K.kappa > 0.0 ? K.kappa : 0.75;
  const double da_t0 = K.t0 > 0.0 ? K.t0 : 10.0;
  FOCT_PARK double da_mu = log(10.0 * eps), da_counter = 0.0, da_sbar = 0.0, da_xbar = 0.0;
  const int max_depth = K.max_depth > 0 ? (K.max_depth <= FOCT_STACK_LEVELS + 1 ? K.max_depth : FOCT_STACK_LEVELS + 1) : 10;
  const double log08 = log(0.8);

  FOCT_PARK double nlf_warm = 0.0, nlf_samp = 0.0, ndiv = 0.0;
  const int n_saved = K.save_warmup ? K.n_iter : K.n_iter - K.n_warmup;

  // Stan's init_stepsize heuristic; `it_site` tags the RNG sites of this call.
  auto init_stepsize = [&](uint32_t it_site) {
    if (!(eps > 0.0) || eps > 1e7) return;
    uint32_t attempt = 0;
    int direction = 0;
    for (;;) {
      double zq = q, zg = g, zV = V, zc2 = c2, zp = 0.0;
      if (act) {
        rng.block(it_site, SITE_INITEPS, attempt, (uint32_t)lane, 0, rb);
        zp = normal_from(rb) / sqrt(invM);
      }
      ++attempt;
      const double H0 = zV + 0.5 * warp_sum(invM * zp * zp);
      leapfrog<NN, MOD, TEAM>(blob, &P, &K.spec, eps, invM, &zq, &zp, &zg, &zV, &zc2, lane, tc);
      double h = zV + 0.5 * warp_sum(invM * zp * zp);
      if (isnan(h)) h = CUDART_INF;
      const double dH = H0 - h;
      if (direction == 0) { direction = dH > log08 ? 1 : -1; continue; }
      if (direction == 1 && !(dH > log08)) break;
      if (direction == -1 && !(dH < log08)) break;
      eps = direction == 1 ? 2.0 * eps : 0.5 * eps;
      if (eps > 1e7 || eps == 0.0 || attempt > 200) break;
    }
  };

  if (K.n_warmup > 0) init_stepsize(it0);

  // pending "init" subtrees, one slot per level (local memory; touched only at merges)
  double st_rho[FOCT_STACK_LEVELS], st_pbeg[FOCT_STACK_LEVELS], st_pend[FOCT_STACK_LEVELS];
  double st_qp[FOCT_STACK_LEVELS], st_gp[FOCT_STACK_LEVELS];
  double st_lsw[FOCT_STACK_LEVELS], st_V[FOCT_STACK_LEVELS], st_c2[FOCT_STACK_LEVELS], st_H[FOCT_STACK_LEVELS];

  int it_done = 0;
  for (int it = 0; it < K.n_iter; ++it) {
    if constexpr (TEAM == 2) {
      // the members must take the same decision: member 0 reads the flag, the team barrier publishes it (the next write
      // is ordered behind this read by the gradient barriers of the iteration in between)
      if (K.cancel) {
        if (tc->member == 0 && lane == 0) *tc->flag = cancel_requested(K) ? 1 : 0;
        team_sync(tc);
        if (*reinterpret_cast<volatile int*>(tc->flag)) break;
      }
    } else {
      if (cancel_requested(K)) break;
    }
    // ================================================================ one NUTS transition
    double p = 0.0;
    if (act) {
      rng.block(it0 + (uint32_t)it, SITE_MOM, 0, (uint32_t)lane, 0, rb);
      p = normal_from(rb) / sqrt(invM);
    }
    const double H0 = V + 0.5 * warp_sum(invM * p * p);
    // trajectory ends (q, p, g) and the running sample
    FOCT_PARK double fq = q, fp = p, fg = g, bq = q, bp = p, bg = g;
    FOCT_PARK double sq = q, sg = g, sV = V, sc2 = c2, sH = H0;
    FOCT_PARK double rho = p, lsw = 0.0;
    double sum_metro = 0.0;
    int n_leap = 0, depth = 0;
    bool divergent = false;

    while (depth < max_depth) {
      rng.block(it0 + (uint32_t)it, SITE_DIR, (uint32_t)depth, 0, 0, rb);
      const bool fwd = u53(rb[0], rb[1]) > 0.5;
      const double u_top = u53(rb[2], rb[3]);
      // integrator starts from the end being extended; the old trajectory is the "init" half of the top merge
      double zq = fwd ? fq : bq, zp = fwd ? fp : bp, zg = fwd ? fg : bg, zV = 0.0, zc2 = 0.0;
      FOCT_PARK double old_end_p = zp, other_end_p = fwd ? bp : fp;
      const double eps_s = fwd ? eps : -eps;
      // current (growing) subtree
      double c_lsw = -CUDART_INF, c_rho = 0.0, c_pbeg = 0.0, c_pend = 0.0, c_qp = 0.0, c_gp = 0.0;
      double c_V = 0.0, c_c2 = 0.0, c_H = 0.0;
      bool valid = true;
      const uint32_t n_leaves = 1u << depth;
      for (uint32_t n = 0; n < n_leaves; ++n) {
        FOCT_T(t_f0);
        leapfrog<NN, MOD, TEAM>(blob, &P, &K.spec, eps_s, invM, &zq, &zp, &zg, &zV, &zc2, lane, tc);
        ++n_leap;
        double h = zV + 0.5 * warp_sum(invM * zp * zp);
        if (isnan(h)) h = CUDART_INF;
        if (h - H0 > 1000.0) divergent = true;
        const double dw = H0 - h;
        sum_metro += dw > 0.0 ? 1.0 : fexp(dw);
        c_lsw = dw; c_rho = zp; c_pbeg = zp; c_pend = zp; c_qp = zq; c_gp = zg; c_V = zV; c_c2 = zc2; c_H = h;
        if (divergent) { valid = false; break; }
        // merge completed siblings upward: bit k of n set  <=>  slot k holds the init half
        FOCT_T(t_m0);
        int k = 0;
        int mb_group = -1;
        for (; (n >> k) & 1u; ++k) {
          double prob_final;
          const double lsw_sub = lse_prob(st_lsw[k], c_lsw, prob_final);
          // one Philox block serves the merges of four consecutive levels at this leaf (32-bit uniforms)
          if ((k >> 2) != mb_group) {
            mb_group = k >> 2;
            rng.block(it0 + (uint32_t)it, SITE_MERGE, (uint32_t)depth, n, (uint32_t)mb_group, rb);
          }
          const uint32_t w = (k & 3) == 0 ? rb[0] : ((k & 3) == 1 ? rb[1] : ((k & 3) == 2 ? rb[2] : rb[3]));
          const bool take_final = ((double)w + 0.5) * 0x1.0p-32 < prob_final;
          const double i_rho = st_rho[k], i_pbeg = st_pbeg[k], i_pend = st_pend[k];
          const bool persist = merge_persists(invM, i_rho, i_pbeg, i_pend, c_rho, c_pbeg, c_pend, lane);
          if (!take_final) { c_qp = st_qp[k]; c_gp = st_gp[k]; c_V = st_V[k]; c_c2 = st_c2[k]; c_H = st_H[k]; }
          c_lsw = lsw_sub; c_rho = i_rho + c_rho; c_pbeg = i_pbeg;
          if (!persist) { valid = false; break; }
        }
        FOCT_T(t_m1);
        FOCT_TADD(5, t_m0, t_m1);
#ifdef FOCT_TIMING
        if (lane == 0) { atomicAdd(&g_tim[6], 1ull); atomicAdd(&g_tim[7], (unsigned long long)k); }
#endif
        if (!valid) break;
        if (n + 1 < n_leaves) {
          st_rho[k] = c_rho; st_pbeg[k] = c_pbeg; st_pend[k] = c_pend; st_qp[k] = c_qp; st_gp[k] = c_gp;
          st_lsw[k] = c_lsw; st_V[k] = c_V; st_c2[k] = c_c2; st_H[k] = c_H;
        }
        FOCT_T(t_f1);
        FOCT_TADD(4, t_f0, t_f1);
      }
      if (fwd) { fq = zq; fp = zp; fg = zg; } else { bq = zq; bp = zp; bg = zg; }
      if (!valid) break;
      ++depth;
      {
        double prob_new;
        const double lsw_all = lse_prob(lsw, c_lsw, prob_new);
        // biased progressive sampling: accept with min(1, w_new / w_old) = min(1, prob_new / (1 - prob_new))
        const double ratio = c_lsw > lsw ? 1.0 : fexp(c_lsw - lsw);
        if (c_lsw > lsw || u_top < ratio) { sq = c_qp; sg = c_gp; sV = c_V; sc2 = c_c2; sH = c_H; }
        lsw = lsw_all;
      }
      const bool persist = merge_persists(invM, rho, other_end_p, old_end_p, c_rho, c_pbeg, c_pend, lane);
      rho += c_rho;
      if (!persist) break;
    }
    const double accept = sum_metro / (double)n_leap;
    q = sq; g = sg; V = sV; c2 = sc2;
    const double eps_used = eps;

    // ================================================================ bookkeeping, write-back, adaptation
    const bool warm = it < K.n_warmup;
    if (warm) nlf_warm += n_leap; else { nlf_samp += n_leap; ndiv += divergent ? 1.0 : 0.0; }
    const int save_idx = K.save_warmup ? it : it - K.n_warmup;
    if (save_idx >= 0 && writer) {
      const size_t row = save_row(K, prob, n_saved, save_idx, chain);
      if (K.draws) {
        double v;
        if (DM::GP) {
          v = (lane == 3 + NN || lane == 4 + NN) ? exp(q) : q;
        } else {
          v = q;
        }
        if (lane == D) v = P.prior_PD ? CUDART_NAN : c2 / P.br_ndf;
        if (lane == D + 1) v = -V;
        if (lane < P_OUT) K.draws[row * P_OUT + lane] = v;
      }
      if (K.sparams && lane < 6) {
        double v = accept;
        if (lane == 1) v = eps_used;
        if (lane == 2) v = (double)depth;
        if (lane == 3) v = (double)n_leap;
        if (lane == 4) v = divergent ? 1.0 : 0.0;
        if (lane == 5) v = sH;
        K.sparams[row * 6 + lane] = v;
      }
    }
    if (warm) {
      // dual averaging
      da_counter += 1.0;
      const double stat = accept > 1.0 ? 1.0 : accept;
      const double eta = 1.0 / (da_counter + da_t0);
      da_sbar = (1.0 - eta) * da_sbar + eta * (da_delta - stat);
      const double x = da_mu - da_sbar * sqrt(da_counter) / da_gamma;
      const double x_eta = pow(da_counter, -da_kappa);
      da_xbar = (1.0 - x_eta) * da_xbar + x_eta * x;
      eps = exp(x);
      // windowed variance
      const bool in_window = a_counter >= a_init_buffer && a_counter < a_num_warmup - a_term_buffer && a_counter != a_num_warmup;
      if (in_window) {
        w_n += 1.0;
        const double delta = q - w_mean;
        w_mean += delta / w_n;
        w_m2 += (q - w_mean) * delta;
      }
      const bool end_window = a_counter == a_next && a_counter != a_num_warmup;
      if (end_window) {
        if (a_next != a_num_warmup - a_term_buffer - 1) {
          a_wsize *= 2;
          a_next = a_counter + a_wsize;
          if (a_next != a_num_warmup - a_term_buffer - 1) {
            const int boundary = a_next + 2 * a_wsize;
            if (boundary >= a_num_warmup - a_term_buffer) a_next = a_num_warmup - a_term_buffer - 1;
          }
        }
        const double var = w_m2 / (w_n - 1.0);
        if (act) invM = (w_n / (w_n + 5.0)) * var + 1e-3 * (5.0 / (w_n + 5.0));
        w_n = 0.0; w_mean = 0.0; w_m2 = 0.0;
        ++a_counter;
        init_stepsize(it0 + (uint32_t)(it + 1));
        da_mu = log(10.0 * eps);
        da_counter = 0.0; da_sbar = 0.0; da_xbar = 0.0;
      } else {
        ++a_counter;
      }
      if (it == K.n_warmup - 1) eps = exp(da_xbar);
    }
    it_done = it + 1;
    report_progress(K, it_done, false, lane == 0 && writer);
  }
  report_progress(K, it_done, true, lane == 0 && writer);
  const size_t pc = (size_t)prob * K.chains + chain;
  if (!writer) return;
  if (lane == 0) {
    if (K.stepsize) K.stepsize[pc] = eps;
    if (K.n_leapfrog) {
      K.n_leapfrog[pc * 2] = (K.accumulate ? K.n_leapfrog[pc * 2] : 0.0) + nlf_warm;
      K.n_leapfrog[pc * 2 + 1] = (K.accumulate ? K.n_leapfrog[pc * 2 + 1] : 0.0) + nlf_samp;
    }
    if (K.n_divergent) K.n_divergent[pc] = (K.accumulate ? K.n_divergent[pc] : 0.0) + ndiv;
  }
  if (K.inv_metric && act) K.inv_metric[pc * D + lane] = invM;
  if (K.last_q && act) K.last_q[pc * D + lane] = q;
}

// Persistent CTAs of up to FOCT_CTA_CHAINS warps.  A work item is (profile, group of <= 4 chains); each CTA
// repeatedly claims an item from the atomic work counter, stages the profile blob into shared memory with
// one TMA bulk copy and runs its chains to completion.  Register budget follows the blob size: the more
// control points, the fewer CTAs fit an SM by shared memory, the more registers each thread may use.
#define FOCT_CTA_CHAINS 4
template <int NN>
struct NutsBounds {
#ifndef FOCT_MINB
#define FOCT_MINB 3
#endif
  static constexpr int MINB = NN <= 15 ? FOCT_MINB : 2;
};

template <int NN, int MOD>
__global__ void __launch_bounds__(32 * FOCT_CTA_CHAINS, NutsBounds<NN>::MINB) nuts_kernel(const SamplerParams K) {
  extern __shared__ __align__(128) double smem[];
  __shared__ uint64_t mbar;
  __shared__ int s_next;
  __shared__ DevProblem s_prob;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = (K.chains + FOCT_CTA_CHAINS - 1) / FOCT_CTA_CHAINS;
  const int n_items = K.n_problems * groups;
  fill_exptab();
  mbar_init(&mbar);
  uint32_t phase = 0;
  for (;;) {
    if (threadIdx.x == 0) s_next = atomicAdd(K.work_counter, 1);
    __syncthreads();
    const int w = s_next;
    if (w >= n_items) break;
    const int j = K.order ? K.order[w / groups] : w / groups, chain = (w % groups) * FOCT_CTA_CHAINS + warp;
    if (threadIdx.x == 0) s_prob = K.probs[j];
    stage_blob_tma(smem, K.blobs + (size_t)j * K.blob_stride, (uint32_t)(K.blob_stride * sizeof(double)), &mbar, phase, threadIdx.x == 0);
    __syncthreads();
    if (chain < K.chains) run_chain<NN, MOD>(K, s_prob, smem, j, chain, lane);
    __syncthreads();
  }
}

// The latency kernel: batches of at most one work item per SM (a single profile — what FitOCT.R's loop submits per call —
// or the continuation rounds of a few unconverged profiles).  Such a batch cannot fill the GPU; what counts is how fast
// ONE chain advances.  Each chain gets a team of two warps (TeamCtx), the CTA an SM of its own, every thread 255
// registers, and the sweep keeps four points in flight per lane (FOCT_UNROLL_LAT): 8 warps x 1 CTA per SM.
// CPC = chains per CTA: FOCT_CTA_CHAINS (an item is a profile's group of <= 4 chains, as in nuts_kernel), or 1 when even
// (profile, chain) items number at most one per SM - the single profile of FitOCT.R's loop then spreads its four chains
// over four SMs and every warp has a scheduler to itself.  The arithmetic of a chain is the same in both.
template <int NN, int MOD, int CPC = FOCT_CTA_CHAINS>
__global__ void __launch_bounds__(64 * CPC, 1) nuts_lat_kernel(const SamplerParams K) {
  extern __shared__ __align__(128) double smem[];
  __shared__ uint64_t mbar;
  __shared__ int s_next;
  __shared__ DevProblem s_prob;
  __shared__ double s_xch[CPC][2 * 64];
  __shared__ int s_flag[CPC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, slot = warp >> 1;
  const int groups = (K.chains + CPC - 1) / CPC;
  const int n_items = K.n_problems * groups;
  fill_exptab();
  mbar_init(&mbar);
  uint32_t phase = 0;
  TeamCtx tc;
  tc.member = warp & 1; tc.bar_id = 1 + slot; tc.parity = 0; tc.xch = &s_xch[slot][0]; tc.flag = &s_flag[slot];
  for (;;) {
    if (threadIdx.x == 0) s_next = atomicAdd(K.work_counter, 1);
    __syncthreads();
    const int w = s_next;
    if (w >= n_items) break;
    const int j = K.order ? K.order[w / groups] : w / groups, chain = (w % groups) * CPC + slot;
    if (threadIdx.x == 0) s_prob = K.probs[j];
    stage_blob_tma(smem, K.blobs + (size_t)j * K.blob_stride, (uint32_t)(K.blob_stride * sizeof(double)), &mbar, phase, threadIdx.x == 0);
    __syncthreads();
    if (chain < K.chains) run_chain<NN, MOD, 2>(K, s_prob, smem, j, chain, lane, &tc);
    __syncthreads();
  }
}

}  // namespace foct
