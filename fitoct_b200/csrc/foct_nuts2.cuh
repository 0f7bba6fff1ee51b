// foct_nuts2.cuh — the NUTS sampler with TWO chains per warp (D <= 16, i.e. Nn <= 11 and the mono-exponential).
//
// Why (profiles/r1_ncu_nuts_final2_summary.txt): with one chain per warp 17 of 32 lanes idle in all the
// O(D) vector work at Nn = 10 (D = 15), and 40 % of the warp instructions of a leaf are outside the data sweep
// (broadcast of q, priors, reduce-scatter, tree logic).  Here a chain is owned by a HALF-warp: lane l of the half owns
// component l of every D-vector, the sweep gives each half 16 points per pass, and every instruction outside the sweep
// serves two chains.  Both halves of a warp work on the same profile and are always at the same point of the sweep,
// so each LDS is a 16-word broadcast.
//
// Two chains can only share a warp if they share its instruction stream: the recursion of Stan's build_tree (already
// an iterative leaf loop in foct_nuts.cuh) is flattened once more into a STATE MACHINE whose tick is
//     PRE  (per chain: start an iteration / a doubling / a step-size trial)
//     one leapfrog step = one gradient evaluation            <- the whole warp, converged, always
//     POST (per chain: leaf bookkeeping, merges, U-turn checks, end of doubling / iteration, adaptation)
// so that the two chains take their gradient together whatever their tree depths, iteration numbers or adaptation
// phases are.  PRE and POST run under per-half control flow (the hardware diverges and reconverges; cross-lane calls
// there use the half's own mask).  The arithmetic, the draw sites of the Philox counters and the order of every
// decision are those of run_chain (foct_nuts.cuh), i.e. of Stan's base_nuts / adapt_diag_e_nuts (MODEL_SPEC §7).
#pragma once
#include "foct_nuts.cuh"

namespace foct {

enum PairMode { PM_ITER = 0, PM_DOUBLE = 1, PM_LEAF = 2, PM_EPS = 3, PM_EPS_EVAL = 4, PM_DONE = 5 };

template <int W>
__device__ __forceinline__ bool merge_persists_w(double invM, double i_rho, double i_pbeg, double i_pend, double f_rho,
                                                 double f_pbeg, double f_pend, int lane, unsigned mask) {
  const double ps_b = invM * i_pbeg, ps_e = invM * f_pend, ps_fb = invM * f_pbeg, ps_ie = invM * i_pend;
  const double r_sub = i_rho + f_rho, r_b = i_rho + f_pbeg, r_c = f_rho + i_pend;
  double v[8];
  v[0] = ps_e * r_sub; v[1] = ps_b * r_sub;
  v[2] = ps_fb * r_b;  v[3] = ps_b * r_b;
  v[4] = ps_e * r_c;   v[5] = ps_ie * r_c;
  v[6] = lane == 0 ? 1.0 : 0.0; v[7] = v[6];
  const double s = warp_reduce_scatter<8, W>(v, lane, mask);
  return __all_sync(mask, s > 0.0);
}

// lane32 = lane in the warp; this half's chain is `chain` (>= K.chains: the half has no chain and only keeps the other
// half company in the gradient evaluations).
template <int NN, int MOD, bool GB>
__device__ void run_pair(const SamplerParams& K, const DevProblem& P, const double* __restrict__ blob, int prob,
                         int chain, int lane32) {
  const double* __restrict__ gbasis = GB ? K.blobs : nullptr;  // blob 0: the basis every profile of the batch shares
  using DM = Dims<NN>;
  constexpr int D = DM::D;
  constexpr int P_OUT = DM::P_OUT;
  constexpr int W = 16;
  static_assert(D <= W, "run_pair needs D <= 16");
  const int lane = lane32 & (W - 1);
  const unsigned hm = 0xffffu << (lane32 & 16);
  const bool have = chain < K.chains;
  const bool act = lane < D;
  Rng rng;
  rng.seed(K.seed, P.id, chain);
  uint32_t rb[4];

  // ---- initial point (MODEL_SPEC §7 init_mode)
  double q = 0.0;
  if (act) {
    rng.block(0, SITE_INIT, 0, (uint32_t)lane, 0, rb);
    if (K.init_mode == 2 && K.init && have) {
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
  Eval ev = warp_logp_grad<NN, MOD, W, GB>(blob, P, K.spec, q, lane, gbasis);
  double g = ev.g, V = -ev.lp, c2 = ev.chi2;
  double invM = 1.0;
  FOCT_PARK double eps = K.stepsize0 > 0.0 ? K.stepsize0 : 1.0;
  if (have && K.invm_init && act) invM = K.invm_init[((size_t)prob * K.chains + chain) * D + lane];
  if (have && K.eps_init) eps = K.eps_init[(size_t)prob * K.chains + chain];
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

  // ---- per-iteration state (parked: touched once per iteration or per doubling)
  FOCT_PARK double fq = q, fp = 0.0, fg = g, bq = q, bp = 0.0, bg = g;
  FOCT_PARK double sq = q, sg = g, sV = V, sc2 = c2, sH = 0.0;
  FOCT_PARK double rho = 0.0, lsw = 0.0, u_top = 0.0, old_end_p = 0.0, other_end_p = 0.0;
  FOCT_PARK int depth = 0, fwd_i = 1, divergent_i = 0;
  FOCT_PARK int it = 0;
  // ---- step-size trial state (Stan's init_stepsize)
  FOCT_PARK int e_attempt = 0, e_direction = 0, e_after_window = 0;
  FOCT_PARK uint32_t e_site = it0;
  FOCT_PARK double e_H0 = 0.0;
  // ---- state of the running subtree / integrator (live across the gradient evaluation)
  double zq = q, zp = 0.0, zg = g, zV = V, zc2 = c2;
  double eps_s = eps, H0 = 0.0, sum_metro = 0.0;
  int n_leap = 0;
  uint32_t n = 0, n_leaves = 1;

  // pending "init" subtrees, one slot per level (local memory; touched only at merges)
  double st_rho[FOCT_STACK_LEVELS], st_pbeg[FOCT_STACK_LEVELS], st_pend[FOCT_STACK_LEVELS];
  double st_qp[FOCT_STACK_LEVELS], st_gp[FOCT_STACK_LEVELS];
  double st_lsw[FOCT_STACK_LEVELS], st_V[FOCT_STACK_LEVELS], st_c2[FOCT_STACK_LEVELS], st_H[FOCT_STACK_LEVELS];

  const size_t pc = (size_t)prob * K.chains + chain;
  auto finish_chain = [&]() {
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
  };

  int mode;
  if (!have) {
    mode = PM_DONE;
  } else if (K.n_warmup > 0 && eps > 0.0 && !(eps > 1e7)) {
    mode = PM_EPS;  // Stan's init_stepsize before the first transition
  } else {
    mode = PM_ITER;
  }

  for (;;) {
    __syncwarp();
    // ================================================================ PRE
    if (mode == PM_ITER && cancel_requested(K)) {
      report_progress(K, it, true, lane == 0);
      finish_chain();
      mode = PM_DONE;
    }
    if (mode == PM_ITER) {
      double p = 0.0;
      if (act) {
        rng.block(it0 + (uint32_t)it, SITE_MOM, 0, (uint32_t)lane, 0, rb);
        p = normal_from(rb) / sqrt(invM);
      }
      H0 = V + 0.5 * warp_sum<W>(invM * p * p, hm);
      fq = q; fp = p; fg = g; bq = q; bp = p; bg = g;
      sq = q; sg = g; sV = V; sc2 = c2; sH = H0;
      rho = p; lsw = 0.0;
      sum_metro = 0.0; n_leap = 0; depth = 0; divergent_i = 0;
      mode = PM_DOUBLE;
    }
    if (mode == PM_DOUBLE) {
      rng.block(it0 + (uint32_t)it, SITE_DIR, (uint32_t)depth, 0, 0, rb);
      const bool fwd = u53(rb[0], rb[1]) > 0.5;
      u_top = u53(rb[2], rb[3]);
      fwd_i = fwd ? 1 : 0;
      // integrator starts from the end being extended; the old trajectory is the "init" half of the top merge
      zq = fwd ? fq : bq; zp = fwd ? fp : bp; zg = fwd ? fg : bg;
      old_end_p = zp; other_end_p = fwd ? bp : fp;
      eps_s = fwd ? eps : -eps;
      n = 0; n_leaves = 1u << depth;
      mode = PM_LEAF;
    } else if (mode == PM_EPS) {
      zq = q; zg = g; zp = 0.0;
      if (act) {
        rng.block(e_site, SITE_INITEPS, (uint32_t)e_attempt, (uint32_t)lane, 0, rb);
        zp = normal_from(rb) / sqrt(invM);
      }
      e_attempt = e_attempt + 1;
      e_H0 = V + 0.5 * warp_sum<W>(invM * zp * zp, hm);
      eps_s = eps;
      mode = PM_EPS_EVAL;
    }
    __syncwarp();
    if (__all_sync(FOCT_FULL, mode == PM_DONE)) break;

    // ================================================================ one leapfrog step, both chains together
    {
      const double ph = fma(0.5 * eps_s, zg, zp);
      zq = fma(eps_s * invM, ph, zq);
      const Eval e2 = warp_logp_grad<NN, MOD, W, GB>(blob, P, K.spec, zq, lane, gbasis);
      zp = fma(0.5 * eps_s, e2.g, ph);
      zg = e2.g; zV = -e2.lp; zc2 = e2.chi2;
    }
    double h = zV + 0.5 * warp_sum<W>(invM * zp * zp);
    if (isnan(h)) h = CUDART_INF;

    // ================================================================ POST
    if (mode == PM_LEAF) {
      ++n_leap;
      const bool div = h - H0 > 1000.0;
      if (div) divergent_i = 1;
      const double dw = H0 - h;
      sum_metro += dw > 0.0 ? 1.0 : fexp(dw);
      double c_lsw = dw, c_rho = zp, c_pbeg = zp, c_pend = zp, c_qp = zq, c_gp = zg, c_V = zV, c_c2 = zc2, c_H = h;
      bool valid = !div;
      int k = 0;
      if (valid) {
        // merge completed siblings upward: bit k of n set  <=>  slot k holds the init half
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
          const bool persist = merge_persists_w<W>(invM, i_rho, i_pbeg, i_pend, c_rho, c_pbeg, c_pend, lane, hm);
          if (!take_final) { c_qp = st_qp[k]; c_gp = st_gp[k]; c_V = st_V[k]; c_c2 = st_c2[k]; c_H = st_H[k]; }
          c_lsw = lsw_sub; c_rho = i_rho + c_rho; c_pbeg = i_pbeg;
          if (!persist) { valid = false; break; }
        }
      }
      if (valid && n + 1 < n_leaves) {
        st_rho[k] = c_rho; st_pbeg[k] = c_pbeg; st_pend[k] = c_pend; st_qp[k] = c_qp; st_gp[k] = c_gp;
        st_lsw[k] = c_lsw; st_V[k] = c_V; st_c2[k] = c_c2; st_H[k] = c_H;
        ++n;
      } else {
        // ---- end of this doubling
        if (fwd_i) { fq = zq; fp = zp; fg = zg; } else { bq = zq; bp = zp; bg = zg; }
        bool iter_done = !valid;
        if (valid) {
          depth = depth + 1;
          double prob_new;
          const double lsw_old = lsw;
          const double lsw_all = lse_prob(lsw_old, c_lsw, prob_new);
          // biased progressive sampling: accept with min(1, w_new / w_old)
          const double ratio = c_lsw > lsw_old ? 1.0 : fexp(c_lsw - lsw_old);
          if (c_lsw > lsw_old || u_top < ratio) { sq = c_qp; sg = c_gp; sV = c_V; sc2 = c_c2; sH = c_H; }
          lsw = lsw_all;
          const double rho_old = rho;
          const bool persist = merge_persists_w<W>(invM, rho_old, other_end_p, old_end_p, c_rho, c_pbeg, c_pend, lane, hm);
          rho = rho_old + c_rho;
          iter_done = !persist || depth >= max_depth;
        }
        if (!iter_done) {
          mode = PM_DOUBLE;
        } else {
          // ============================================================ end of the transition
          const double accept = sum_metro / (double)n_leap;
          q = sq; g = sg; V = sV; c2 = sc2;
          const double eps_used = eps;
          const bool warm = it < K.n_warmup;
          if (warm) nlf_warm += n_leap; else { nlf_samp += n_leap; ndiv += divergent_i ? 1.0 : 0.0; }
          const int save_idx = K.save_warmup ? it : it - K.n_warmup;
          if (save_idx >= 0) {
            const size_t row = save_row(K, prob, n_saved, save_idx, chain);
            if (K.draws) {
              double v = q;
              if (DM::GP && (lane == 3 + NN || lane == 4 + NN)) v = exp(q);
              const double br = P.prior_PD ? CUDART_NAN : c2 / P.br_ndf;
              if (lane == D) v = br;
              if (lane == D + 1) v = -V;
              if (lane < P_OUT) K.draws[row * P_OUT + lane] = v;
              // columns beyond the half's 16 lanes can only be br (D = 16) and lp__, both uniform over the half
              if (lane + W < P_OUT) K.draws[row * P_OUT + lane + W] = lane + W == D ? br : -V;
            }
            if (K.sparams && lane < 6) {
              double v = accept;
              if (lane == 1) v = eps_used;
              if (lane == 2) v = (double)depth;
              if (lane == 3) v = (double)n_leap;
              if (lane == 4) v = divergent_i ? 1.0 : 0.0;
              if (lane == 5) v = sH;
              K.sparams[row * 6 + lane] = v;
            }
          }
          bool to_eps = false;
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
              // Stan: init_stepsize(), then set_mu(log(10 eps)) and restart() — continued in PM_EPS_EVAL
              e_site = it0 + (uint32_t)(it + 1); e_attempt = 0; e_direction = 0; e_after_window = 1;
              to_eps = eps > 0.0 && !(eps > 1e7);
              if (!to_eps) {
                da_mu = log(10.0 * eps);
                da_counter = 0.0; da_sbar = 0.0; da_xbar = 0.0;
              }
            } else {
              ++a_counter;
            }
            if (!to_eps && it == K.n_warmup - 1) eps = exp(da_xbar);
          }
          if (to_eps) {
            mode = PM_EPS;
          } else {
            it = it + 1;
            mode = it < K.n_iter ? PM_ITER : PM_DONE;
            report_progress(K, it, false, lane == 0);
            if (mode == PM_DONE) { report_progress(K, it, true, lane == 0); finish_chain(); }
          }
        }
      }
    } else if (mode == PM_EPS_EVAL) {
      const double dH = e_H0 - h;
      bool finished = false;
      if (e_direction == 0) {
        e_direction = dH > log08 ? 1 : -1;
      } else if ((e_direction == 1 && !(dH > log08)) || (e_direction == -1 && !(dH < log08))) {
        finished = true;
      } else {
        eps = e_direction == 1 ? 2.0 * eps : 0.5 * eps;
        if (eps > 1e7 || eps == 0.0 || e_attempt > 200) finished = true;
      }
      if (!finished) {
        mode = PM_EPS;
      } else if (e_after_window) {
        da_mu = log(10.0 * eps);
        da_counter = 0.0; da_sbar = 0.0; da_xbar = 0.0;
        if (it == K.n_warmup - 1) eps = exp(da_xbar);
        e_after_window = 0;
        it = it + 1;
        mode = it < K.n_iter ? PM_ITER : PM_DONE;
        report_progress(K, it, false, lane == 0);
        if (mode == PM_DONE) { report_progress(K, it, true, lane == 0); finish_chain(); }
      } else {
        mode = PM_ITER;  // the trial before the first transition
      }
    }
  }
}

// Persistent CTAs of two warps = four chains of one profile (chains 2w + h: warp w, half h).  A work item is
// (profile, group of <= 4 chains) as in nuts_kernel.  GB = false: the whole blob is staged, four CTAs share an SM (four
// staged profiles, up to 255 registers per thread).  GB = true (every profile of the batch has the same depth grid, hence
// the same basis): only cx | y | w are staged (12 KB), the basis rows come through L1 from one blob in global memory, and
// the number of resident CTAs is set by the registers instead of the shared memory.
#define FOCT_PAIR_CTA_CHAINS 4
#ifndef FOCT_PAIR_MINB
#define FOCT_PAIR_MINB 4
#endif
#ifndef FOCT_PAIR_MINB_GB
#define FOCT_PAIR_MINB_GB 6
#endif
template <int NN, int MOD, bool GB>
__global__ void __launch_bounds__(16 * FOCT_PAIR_CTA_CHAINS, GB ? FOCT_PAIR_MINB_GB : FOCT_PAIR_MINB)
nuts2_kernel(const SamplerParams K) {
  extern __shared__ __align__(128) double smem[];
  __shared__ uint64_t mbar;
  __shared__ int s_next;
  __shared__ DevProblem s_prob;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = (K.chains + FOCT_PAIR_CTA_CHAINS - 1) / FOCT_PAIR_CTA_CHAINS;
  const int n_items = K.n_problems * groups;
  mbar_init(&mbar);
  uint32_t phase = 0;
  for (;;) {
    if (threadIdx.x == 0) s_next = atomicAdd(K.work_counter, 1);
    __syncthreads();
    const int w = s_next;
    if (w >= n_items) break;
    const int j = K.order ? K.order[w / groups] : w / groups;
    const int chain = (w % groups) * FOCT_PAIR_CTA_CHAINS + 2 * warp + (lane >> 4);
    if (threadIdx.x == 0) s_prob = K.probs[j];
    if (GB) stage_rows_tma(smem, K.blobs + (size_t)j * K.blob_stride, K.npad / 32, 3 + NN, 3, &mbar, phase);
    else stage_blob_tma(smem, K.blobs + (size_t)j * K.blob_stride, (uint32_t)(K.blob_stride * sizeof(double)), &mbar, phase);
    __syncthreads();
    // a warp runs if at least its first half has a chain
    if ((w % groups) * FOCT_PAIR_CTA_CHAINS + 2 * warp < K.chains) run_pair<NN, MOD, GB>(K, s_prob, smem, j, chain, lane);
    __syncthreads();
  }
}

}  // namespace foct
