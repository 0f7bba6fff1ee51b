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

enum PairMode { PM_ITER = 0, PM_DOUBLE = 1, PM_LEAF = 2, PM_EPS = 3, PM_EPS_EVAL = 4, PM_DONE = 5, PM_INIT = 6 };

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

// Everything a half-warp carries from one tick to the next, in the order it is saved (time slicing, see nuts2_kernel).
#define FOCT_PAIR_DOUBLES(X)                                                                                              \
  X(q) X(g) X(V) X(c2) X(invM) X(eps) X(w_n) X(w_mean) X(w_m2) X(da_mu) X(da_counter) X(da_sbar) X(da_xbar) X(nlf_warm)  \
  X(nlf_samp) X(ndiv) X(fq) X(fp) X(fg) X(bq) X(bp) X(bg) X(sq) X(sg) X(sV) X(sc2) X(sH) X(rho) X(lsw) X(u_top)          \
  X(old_end_p) X(other_end_p) X(e_H0) X(zq) X(zp) X(zg) X(zV) X(zc2) X(eps_s) X(H0) X(sum_metro)
#define FOCT_PAIR_INTS(X)                                                                                                 \
  X(a_counter) X(a_wsize) X(a_next) X(depth) X(fwd_i) X(divergent_i) X(it) X(e_attempt) X(e_direction) X(e_after_window) \
  X(e_site) X(n_leap) X(n) X(n_leaves) X(mode)
#define FOCT_PAIR_N_DOUBLES 41
#define FOCT_PAIR_N_INTS 15
// doubles per lane of a saved warp: the scalars above, the ints packed two to a double, six stack arrays
// Levels of the subtree stack that nuts2w_kernel keeps in shared memory (run_pair, GB = 2).  Measured with y | w staged
// (8 KB per warp), 1776 profiles, gradients/s: 0 levels 3.27e8 (98 KB of shared memory per SM -> 132 KB carve-out..., L1 124 KB),
// 1 level 3.32e8 (same carve-out), 2 levels 3.18e8 and 3 levels 2.97e8 (the next carve-outs leave 92 / 60 KB of L1 for the
// 40 KB of basis rows and the remaining stack traffic): what the kernel needs most is L1 capacity.  profiles/r2_kernel_experiments.txt
#ifndef FOCT_STACK_SMEM
#define FOCT_STACK_SMEM 1
#endif
// Below seven control points ptxas answers the shared-memory stack with a schedule that runs the two points of a sweep
// iteration one after the other instead of interleaved (scripts/sass_sched.py: rcp at [9, 44] instead of [16, 23]; Nn = 5
// measured 13 % slower), so those instantiations keep the whole stack in local memory.
#define FOCT_STACK_SMEM_LEVELS(NN) ((NN) >= 7 ? FOCT_STACK_SMEM : 0)
#define FOCT_PAIR_STATE_DOUBLES (FOCT_PAIR_N_DOUBLES + (FOCT_PAIR_N_INTS + 1) / 2 + 6 * FOCT_STACK_LEVELS)

// lane32 = lane in the warp; this half's chain is `chain` (>= K.chains: the half has no chain and only keeps the other
// half company in the gradient evaluations).
// Time slicing: `sv` (nullptr: off) is this warp's block of FOCT_PAIR_STATE_DOUBLES x 32 doubles in global memory.  With
// `resume` the state is loaded from it instead of being initialised; after K.slice_ticks gradient evaluations the warp
// may give its place up: the state is stored there and the function returns the unit's progress bin (iterations done,
// in FOCT_PROGRESS_BINS-ths of the run); -1: both chains finished.  Who yields:
//   K.slice_hist == nullptr (CTA-level items): whoever finds another item waiting (round robin);
//   else: only a unit that is AHEAD, in iterations done, of some unit that waits.  Fits differ in cost by up to 40 %;
//   keeping them level in iterations instead of in gradient evaluations lets them all finish together, so the expensive
//   fits are not left to run alone at the end (an online approximation of longest-remaining-time-first).
#define FOCT_PROGRESS_BINS 64
template <int NN, int MOD, int GB>
__device__ int run_pair(const SamplerParams& K, const DevProblem& P, const double* __restrict__ blob, int prob,
                         int chain, int lane32, double* __restrict__ sv, bool resume, int n_items,
                         const double* __restrict__ gbasis,    // GB: blob 0, whose basis rows every profile of the batch shares
                         double* stack_smem = nullptr) {       // GB = 2: FOCT_STACK_SMEM levels x blockDim.x x 7 doubles
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
  const uint32_t it0 = (uint32_t)K.it_offset;

  // ---- adaptation constants (Stan windowed_adaptation / welford_var_estimator / stepsize_adaptation)
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
  const double da_delta = K.adapt_delta > 0.0 ? K.adapt_delta : 0.8;
  const double da_gamma = K.gamma > 0.0 ? K.gamma : 0.05, da_kappa = K.kappa > 0.0 ? K.kappa : 0.75;
  const double da_t0 = K.t0 > 0.0 ? K.t0 : 10.0;
  const int max_depth = K.max_depth > 0 ? (K.max_depth <= FOCT_STACK_LEVELS + 1 ? K.max_depth : FOCT_STACK_LEVELS + 1) : 10;
  const double log08 = log(0.8);
  const int n_saved = K.save_warmup ? K.n_iter : K.n_iter - K.n_warmup;

  // ---- chain state.  FOCT_PARK (volatile): touched once per iteration or per doubling, kept out of the registers
  double q = 0.0, g = 0.0, V = 0.0, c2 = 0.0, invM = 1.0;
  FOCT_PARK double eps = K.stepsize0 > 0.0 ? K.stepsize0 : 1.0;
  FOCT_PARK int a_counter = 0, a_wsize = a_base_window, a_next = a_init_buffer + a_base_window - 1;
  FOCT_PARK double w_n = 0.0, w_mean = 0.0, w_m2 = 0.0;
  FOCT_PARK double da_mu = 0.0, da_counter = 0.0, da_sbar = 0.0, da_xbar = 0.0;
  FOCT_PARK double nlf_warm = 0.0, nlf_samp = 0.0, ndiv = 0.0;
  // ---- per-iteration state
  FOCT_PARK double fq = 0.0, fp = 0.0, fg = 0.0, bq = 0.0, bp = 0.0, bg = 0.0;
  FOCT_PARK double sq = 0.0, sg = 0.0, sV = 0.0, sc2 = 0.0, sH = 0.0;
  FOCT_PARK double rho = 0.0, lsw = 0.0, u_top = 0.0, old_end_p = 0.0, other_end_p = 0.0;
  FOCT_PARK int depth = 0, fwd_i = 1, divergent_i = 0;
  FOCT_PARK int it = 0;
  // ---- step-size trial state (Stan's init_stepsize)
  FOCT_PARK int e_attempt = 0, e_direction = 0, e_after_window = 0;
  FOCT_PARK uint32_t e_site = it0;
  FOCT_PARK double e_H0 = 0.0;
  // ---- state of the running subtree / integrator (live across the gradient evaluation)
  double zq = 0.0, zp = 0.0, zg = 0.0, zV = 0.0, zc2 = 0.0;
  double eps_s = 0.0, H0 = 0.0, sum_metro = 0.0;
  int n_leap = 0;
  uint32_t n = 0, n_leaves = 1;
  int mode = PM_DONE;

  // pending "init" subtrees, one slot per level (local memory; touched only at merges)
  double st_rho[FOCT_STACK_LEVELS], st_pbeg[FOCT_STACK_LEVELS], st_pend[FOCT_STACK_LEVELS];
  double st_qp[FOCT_STACK_LEVELS], st_gp[FOCT_STACK_LEVELS];
  // the four scalars of a pending subtree (log weight, potential, chi2, H of its proposal: uniform over the half) share ONE
  // per-lane slot — lane 0 keeps lsw, lane 1 V, lane 2 c2, lane 3 H — instead of four replicated ones: a third less
  // local-memory traffic per leaf (the subtree stack is what misses L1: profiles/r2_ncu_nuts2w_summary.txt)
  double st_sc[FOCT_STACK_LEVELS];
  // The lowest levels are the busy ones (level k is written and read once per 2^(k+1) leaves): nuts2w_kernel keeps the
  // first FOCT_STACK_SMEM of them in shared memory (7 doubles per thread and level: 6 used, an odd word stride), the rest
  // stays in local memory.  75 % of the stack traffic of a leaf then neither misses L1 nor evicts the basis rows from it.
  constexpr int SL = GB == 2 ? FOCT_STACK_SMEM_LEVELS(NN) : 0;
  double* const sst = SL > 0 ? stack_smem + threadIdx.x * 7 : nullptr;
  const int SST = (int)blockDim.x * 7;
#define ST_LD(arr, row, k) ((SL > 0 && (k) < SL) ? sst[(k) * SST + (row)] : arr[k])
#define ST_ST(arr, row, k, v) do { if (SL > 0 && (k) < SL) sst[(k) * SST + (row)] = (v); else arr[k] = (v); } while (0)

#ifndef FOCT_TEST_NO_RESUME
  if (resume) {
    // (L2 loads: the block may have been written by another SM since this one last read it)
    const double* s = sv + lane32;
#define X(v) v = __ldcg(s); s += 32;
    FOCT_PAIR_DOUBLES(X)
#undef X
    const int* si = reinterpret_cast<const int*>(sv + FOCT_PAIR_N_DOUBLES * 32) + lane32;
#define X(v) v = __ldcg(si); si += 32;
    FOCT_PAIR_INTS(X)
#undef X
    s = sv + (FOCT_PAIR_N_DOUBLES + (FOCT_PAIR_N_INTS + 1) / 2) * 32 + lane32;
#pragma unroll 1
    for (int k = 0; k < FOCT_STACK_LEVELS; ++k, s += 6 * 32) {
      ST_ST(st_rho, 0, k, __ldcg(s)); ST_ST(st_pbeg, 1, k, __ldcg(s + 32)); ST_ST(st_pend, 2, k, __ldcg(s + 64));
      ST_ST(st_qp, 3, k, __ldcg(s + 96)); ST_ST(st_gp, 4, k, __ldcg(s + 128)); ST_ST(st_sc, 5, k, __ldcg(s + 160));
    }
  } else
#endif
  {
    // ---- initial point (MODEL_SPEC §7 init_mode)
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
    // The gradient at the initial point is taken by the first tick of the loop below (mode PM_INIT: a leapfrog step of
    // length zero leaves q where it is), so that the kernel carries ONE inlined copy of the sweep: with a second one for
    // this evaluation the code outgrew the instruction cache (7.8 k instructions = 125 KB ran 4-5 % slower than 7.6 k,
    // profiles/r2_kernel_experiments.txt).
    if (have && K.invm_init && act) invM = K.invm_init[((size_t)prob * K.chains + chain) * D + lane];
    if (have && K.eps_init) eps = K.eps_init[(size_t)prob * K.chains + chain];
    zq = q; zp = 0.0; zg = 0.0; eps_s = 0.0;
    mode = have ? PM_INIT : PM_DONE;
  }

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

  int ticks = 0;
  for (;;) {
    __syncwarp();
#ifndef FOCT_TEST_NO_SUSPEND
    if (sv && ticks >= K.slice_ticks) {
      // end of the slice
      const unsigned head = *reinterpret_cast<const volatile unsigned*>(K.slice_ctl);
      const unsigned pushed = *reinterpret_cast<const volatile unsigned*>(K.slice_ctl + 1);
      bool yield = head < (unsigned)n_items + pushed;  // somebody waits
      int bin = 0;
      if (K.slice_hist) {
        int f = mode == PM_DONE ? 0x7fffffff : it;
        f = min(f, __shfl_xor_sync(FOCT_FULL, f, 16));
        bin = min((int)(((long long)f * FOCT_PROGRESS_BINS) / max(K.n_iter, 1)), FOCT_PROGRESS_BINS - 1);
        if (yield) {
          // the least advanced waiting unit: one that has not started yet (bin 0), else the lowest occupied bin
          int qmin = 0;
          if (head >= (unsigned)n_items) {
            const int h0 = *reinterpret_cast<const volatile int*>(K.slice_hist + lane32);
            const int h1 = *reinterpret_cast<const volatile int*>(K.slice_hist + 32 + lane32);
            const unsigned b0 = __ballot_sync(FOCT_FULL, h0 > 0), b1 = __ballot_sync(FOCT_FULL, h1 > 0);
            qmin = b0 ? __ffs((int)b0) - 1 : (b1 ? 32 + __ffs((int)b1) - 1 : FOCT_PROGRESS_BINS);
          }
          yield = bin > qmin;
        }
      }
      if (yield) {
        double* s = sv + lane32;
#define X(v) __stcg(s, (double)v); s += 32;
        FOCT_PAIR_DOUBLES(X)
#undef X
        int* si = reinterpret_cast<int*>(sv + FOCT_PAIR_N_DOUBLES * 32) + lane32;
#define X(v) __stcg(si, (int)v); si += 32;
        FOCT_PAIR_INTS(X)
#undef X
        s = sv + (FOCT_PAIR_N_DOUBLES + (FOCT_PAIR_N_INTS + 1) / 2) * 32 + lane32;
#pragma unroll 1
        for (int k = 0; k < FOCT_STACK_LEVELS; ++k, s += 6 * 32) {
          __stcg(s, ST_LD(st_rho, 0, k)); __stcg(s + 32, ST_LD(st_pbeg, 1, k)); __stcg(s + 64, ST_LD(st_pend, 2, k));
          __stcg(s + 96, ST_LD(st_qp, 3, k)); __stcg(s + 128, ST_LD(st_gp, 4, k)); __stcg(s + 160, ST_LD(st_sc, 5, k));
        }
        __threadfence();
        return bin;
      }
      ticks = 0;
    }
    ++ticks;
#endif
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
    if (__all_sync(FOCT_FULL, mode == PM_DONE)) return -1;

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
          const double i_sc = ST_LD(st_sc, 5, k);
          const double lsw_sub = lse_prob(bcast<W>(i_sc, 0, hm), c_lsw, prob_final);
          // one Philox block serves the merges of four consecutive levels at this leaf (32-bit uniforms)
          if ((k >> 2) != mb_group) {
            mb_group = k >> 2;
            rng.block(it0 + (uint32_t)it, SITE_MERGE, (uint32_t)depth, n, (uint32_t)mb_group, rb);
          }
          const uint32_t w = (k & 3) == 0 ? rb[0] : ((k & 3) == 1 ? rb[1] : ((k & 3) == 2 ? rb[2] : rb[3]));
          const bool take_final = ((double)w + 0.5) * 0x1.0p-32 < prob_final;
          const double i_rho = ST_LD(st_rho, 0, k), i_pbeg = ST_LD(st_pbeg, 1, k), i_pend = ST_LD(st_pend, 2, k);
          const bool persist = merge_persists_w<W>(invM, i_rho, i_pbeg, i_pend, c_rho, c_pbeg, c_pend, lane, hm);
          if (!take_final) {
            c_qp = ST_LD(st_qp, 3, k); c_gp = ST_LD(st_gp, 4, k);
            c_V = bcast<W>(i_sc, 1, hm); c_c2 = bcast<W>(i_sc, 2, hm); c_H = bcast<W>(i_sc, 3, hm);
          }
          c_lsw = lsw_sub; c_rho = i_rho + c_rho; c_pbeg = i_pbeg;
          if (!persist) { valid = false; break; }
        }
      }
      if (valid && n + 1 < n_leaves) {
        ST_ST(st_rho, 0, k, c_rho); ST_ST(st_pbeg, 1, k, c_pbeg); ST_ST(st_pend, 2, k, c_pend); ST_ST(st_qp, 3, k, c_qp);
        ST_ST(st_gp, 4, k, c_gp);
        ST_ST(st_sc, 5, k, lane == 0 ? c_lsw : (lane == 1 ? c_V : (lane == 2 ? c_c2 : c_H)));
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
    } else if (mode == PM_INIT) {
      g = zg; V = zV; c2 = zc2;
      da_mu = log(10.0 * eps);
      mode = (K.n_warmup > 0 && eps > 0.0 && !(eps > 1e7)) ? PM_EPS   // Stan's init_stepsize before the first transition
                                                            : PM_ITER;
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
//
// TIME SLICING (K.slice_state != nullptr; the host turns it on when there are more work items than resident CTAs).  A fit
// is ~3e5 gradient evaluations long and cannot be split, so with a plain work counter the last items of a batch run
// alone on a mostly idle GPU: 1000 profiles on 888 resident CTAs took as long as 1250 (profiles/r2_kernel_experiments.txt).
// Here a CTA gives its item up after K.slice_ticks gradient evaluations if another item is waiting: the warps store their
// chain state (FOCT_PAIR_STATE_DOUBLES x 32 doubles per warp) to global memory, the item goes to the back of a ring queue
// and the CTA takes the item at the front.  All items advance at the same rate (round robin), the GPU stays full until
// fewer items than CTAs are left, and what is left then is the same small remainder of every long fit.  The result of
// a fit does not depend on where or when its slices run (per-profile Philox keys; the state is saved and restored bit for
// bit), which tests/test_gpu_parity.py checks against the unsliced kernel.
//   ctl[0] tickets handed out; tickets < n_items are the items themselves (in K.order), ticket n_items + p is the p-th push
//   ctl[1] pushes so far       ctl[2] items finished
//   queue[p % n_items] = (ticket << 32 | item): a slot cannot be reused before it is consumed, because the items between
//   two pushes into the same slot would have to be n_items + 1 different ones
#define FOCT_PAIR_CTA_CHAINS 4
#ifndef FOCT_PAIR_MINB
#define FOCT_PAIR_MINB 4
#endif
#ifndef FOCT_PAIR_MINB_GB
#define FOCT_PAIR_MINB_GB 6
#endif
template <int NN, int MOD, int GB>
__global__ void __launch_bounds__(16 * FOCT_PAIR_CTA_CHAINS, GB ? FOCT_PAIR_MINB_GB : FOCT_PAIR_MINB)
nuts2_kernel(const SamplerParams K) {
  extern __shared__ __align__(128) double smem[];
  __shared__ uint64_t mbar;
  __shared__ int s_next, s_resume;
  __shared__ DevProblem s_prob;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool leader = threadIdx.x == 0;
  const int groups = (K.chains + FOCT_PAIR_CTA_CHAINS - 1) / FOCT_PAIR_CTA_CHAINS;
  const int n_items = K.n_problems * groups;
  const bool sliced = K.slice_state != nullptr;
  constexpr int CTA_WARPS = FOCT_PAIR_CTA_CHAINS / 2;
  const double* gbasis = GB ? K.blobs : nullptr;  // blob 0: the basis every profile of the batch shares
  fill_exptab();
  mbar_init(&mbar);
  uint32_t phase = 0;
  for (;;) {
    if (leader) {
      int w = -1, res = 0;
      if (!sliced) {
        w = atomicAdd(K.work_counter, 1);
        if (w >= n_items) w = -1;
      } else {
        const unsigned t = atomicAdd(K.slice_ctl, 1u);
        if (t < (unsigned)n_items) {
          w = (int)t;
        } else {
          // wait for the push this ticket stands for, or for the last item to finish
          const volatile unsigned long long* slot = K.slice_queue + (t - (unsigned)n_items) % (unsigned)n_items;
          for (;;) {
            const unsigned long long v = *slot;
            if ((unsigned)(v >> 32) == t) { w = (int)(v & 0xffffffffu); res = 1; break; }
            if (*reinterpret_cast<const volatile unsigned*>(K.slice_ctl + 2) >= (unsigned)n_items) break;
            __nanosleep(2000);
          }
          __threadfence();
        }
      }
      s_next = w; s_resume = res;
    }
    __syncthreads();
    const int w = s_next;
    const bool resume = s_resume != 0;
    if (w < 0) break;
    const int j = K.order ? K.order[w / groups] : w / groups;
    const int chain = (w % groups) * FOCT_PAIR_CTA_CHAINS + 2 * warp + (lane >> 4);
    if (leader) s_prob = K.probs[j];
    if (GB) stage_rows_tma(smem, K.blobs + (size_t)j * K.blob_stride, K.npad / 32, 3 + NN, 3, &mbar, phase, leader);
    else stage_blob_tma(smem, K.blobs + (size_t)j * K.blob_stride, (uint32_t)(K.blob_stride * sizeof(double)), &mbar, phase, leader);
    __syncthreads();
    // a warp runs if at least its first half has a chain
    bool fin = true;
    if ((w % groups) * FOCT_PAIR_CTA_CHAINS + 2 * warp < K.chains) {
      double* sv = sliced ? K.slice_state + ((size_t)w * CTA_WARPS + warp) * (FOCT_PAIR_STATE_DOUBLES * 32) : nullptr;
      int* wdone = sliced ? K.slice_done + (size_t)w * CTA_WARPS + warp : nullptr;
      if (!(resume && __ldcg(wdone) != 0)) {
        fin = run_pair<NN, MOD, GB>(K, s_prob, smem, j, chain, lane, sv, resume, n_items, gbasis) < 0;
        if (sliced && lane == 0) __stcg(wdone, fin ? 1 : 0);
      }
    }
    const int all_fin = __syncthreads_and(fin ? 1 : 0);
    if (sliced && leader) {
      if (all_fin) {
        atomicAdd(K.slice_ctl + 2, 1u);
      } else {
        __threadfence();  // the state stored by this CTA's warps before the barrier is visible before the queue entry is
        const unsigned p = atomicAdd(K.slice_ctl + 1, 1u);
        *reinterpret_cast<volatile unsigned long long*>(K.slice_queue + p % (unsigned)n_items) =
            ((unsigned long long)((unsigned)n_items + p) << 32) | (unsigned)w;
      }
    }
  }
}

// The shared-basis kernel with WARPS as the scheduling unit.  A work unit is (profile, pair of chains); each warp claims
// units on its own, stages the cx | y | w rows of its profile into its own 12 KB of shared memory (own mbarrier, TMA
// issued by lane 0), and runs, suspends and resumes them exactly like a CTA of nuts2_kernel does with an item — but
// without any CTA-wide barrier: the two warps of a CTA share nothing but the SM.  Against CTA-level items this removes
// the wait of a profile's faster pair for its slower one (the four chains of a profile differ by ~5 % in gradient
// evaluations) and lets the time slicing act per pair.
//   ctl[0] tickets | ctl[1] pushes | ctl[2] units finished      (as in nuts2_kernel, with units for items)
// FOCT_CX_SHARED: with one depth grid (and one dataType) for the batch the row c x is shared as well and comes through L1
// from blob 0 next to the basis rows; a warp stages y | w only (8 KB).  Less shared memory = more L1 for the basis rows and
// the subtree stacks in local memory (L1 hit rates at 12 KB per warp: global 94 %, local 35 %: profiles/r2_ncu_nuts2w_pf_summary.txt).
#ifndef FOCT_CX_SHARED
#define FOCT_CX_SHARED 1
#endif
template <int NN, int MOD>
__global__ void __launch_bounds__(64, FOCT_PAIR_MINB_GB) nuts2w_kernel(const SamplerParams K) {
  extern __shared__ __align__(128) double smem[];
  __shared__ uint64_t mbar[2];
  __shared__ DevProblem s_prob[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool leader = lane == 0;
  const int upp = (K.chains + 1) / 2;  // units (chain pairs) per profile
  const int n_units = K.n_problems * upp;
  const bool sliced = K.slice_state != nullptr;
  constexpr int GBW = FOCT_CX_SHARED ? 2 : 1, SR = FOCT_CX_SHARED ? 2 : 3;  // staged rows per block: y | w, or c x | y | w
  double* rows = smem + (size_t)warp * SR * K.npad;
  double* stack_smem = smem + (size_t)2 * SR * K.npad;  // [FOCT_STACK_SMEM levels][64 threads][7]
  fill_exptab();
  if (leader) {
    const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(&mbar[warp]);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t phase = 0;
  for (;;) {
    int u = -1, res = 0;
    if (leader) {
      if (!sliced) {
        u = atomicAdd(K.work_counter, 1);
        if (u >= n_units) u = -1;
      } else {
        const unsigned t = atomicAdd(K.slice_ctl, 1u);
        if (t < (unsigned)n_units) {
          u = (int)t;
        } else {
          // wait for the push this ticket stands for, or for the last unit to finish
          const volatile unsigned long long* slot = K.slice_queue + (t - (unsigned)n_units) % (unsigned)n_units;
          for (;;) {
            const unsigned long long v = *slot;
            if ((unsigned)(v >> 32) == t) {
              u = (int)((v & 0xffffffffu) >> 6); res = 1;
              if (K.slice_hist) atomicSub(K.slice_hist + (int)(v & 63u), 1);  // no longer waiting
              break;
            }
            if (*reinterpret_cast<const volatile unsigned*>(K.slice_ctl + 2) >= (unsigned)n_units) break;
            __nanosleep(2000);
          }
          __threadfence();
        }
      }
    }
    u = __shfl_sync(FOCT_FULL, u, 0);
    res = __shfl_sync(FOCT_FULL, res, 0);
    if (u < 0) break;
    const int r = u / upp, pair = u - r * upp;
    const int j = K.order ? K.order[r] : r;
    const int chain = 2 * pair + (lane >> 4);
    if (leader) s_prob[warp] = K.probs[j];
    stage_rows_tma(rows, K.blobs + (size_t)j * K.blob_stride + (3 - SR) * 32, K.npad / 32, 3 + NN, SR, &mbar[warp], phase, leader);
    __syncwarp();
    double* sv = sliced ? K.slice_state + (size_t)u * (FOCT_PAIR_STATE_DOUBLES * 32) : nullptr;
    const int bin = run_pair<NN, MOD, GBW>(K, s_prob[warp], rows, j, chain, lane, sv, res != 0, n_units, K.blobs, stack_smem);
    __syncwarp();
    if (sliced && leader) {
      if (bin < 0) {
        atomicAdd(K.slice_ctl + 2, 1u);
      } else {
        __threadfence();  // the state stored by this warp's lanes (before the __syncwarp) is visible before the queue entry is
        if (K.slice_hist) atomicAdd(K.slice_hist + bin, 1);
        const unsigned p = atomicAdd(K.slice_ctl + 1, 1u);
        // (unit << 6 | progress bin: whoever takes the unit removes it from the histogram of the waiting ones)
        *reinterpret_cast<volatile unsigned long long*>(K.slice_queue + p % (unsigned)n_units) =
            ((unsigned long long)((unsigned)n_units + p) << 32) | ((unsigned)u << 6) | (unsigned)bin;
      }
    }
  }
}

}  // namespace foct
