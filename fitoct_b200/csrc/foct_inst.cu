// foct_inst.cu — instantiates the sampling and parity kernels for ONE control-point count
// (-DFOCT_INST_NN=<NN>, NN = 0 is the mono-exponential) and both modulation sites.
#include "foct_launch.h"
#include "foct_nuts2.cuh"

#ifndef FOCT_INST_NN
#error "compile with -DFOCT_INST_NN=<NN>"
#endif

namespace foct {

// Parity hook kernel: one CTA per profile, each warp evaluates log density + gradient at q points.
template <int NN, int MOD>
__global__ void __launch_bounds__(128) logp_kernel(const LogpParams K) {
  extern __shared__ __align__(128) double smem[];
  fill_exptab();
  __shared__ uint64_t mbar;
  __shared__ DevProblem s_prob;
  constexpr int D = Dims<NN>::D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  mbar_init(&mbar);
  uint32_t phase = 0;
  for (int j = blockIdx.x; j < K.n_problems; j += gridDim.x) {
    if (threadIdx.x == 0) s_prob = K.probs[j];
    stage_blob_tma(smem, K.blobs + (size_t)j * K.blob_stride, (uint32_t)(K.blob_stride * sizeof(double)), &mbar, phase, threadIdx.x == 0);
    __syncthreads();
    if constexpr (D <= 16) {
      if (K.width == 16) {
        // the half-warp evaluation used by the two-chains-per-warp sampler: each half takes its own q, both halves
        // of a warp call together (full-mask shuffles of width 16)
        const int l = lane & 15, half = lane >> 4;
        for (int base = 0; base < K.n_q; base += 2 * nwarp) {
          const int iq = base + 2 * warp + half;
          const bool ok = iq < K.n_q;
          const size_t r = (size_t)j * K.n_q + (ok ? iq : 0);
          const double qd = ok && l < D ? K.q[r * D + l] : 0.0;
          const Eval ev = warp_logp_grad<NN, MOD, 16>(smem, s_prob, K.spec, qd, l);
          if (ok && l < D) K.grad[r * D + l] = ev.g;
          if (ok && l == 0) {
            K.lp[r] = ev.lp;
            if (K.chi2) K.chi2[r] = ev.chi2;
          }
        }
        __syncthreads();
        continue;
      }
      if constexpr (NN > 0) {
        if (K.width == 17) {
          // the evaluation of nuts2w_kernel, the kernel of BASELINE-size batches: half-warps, only y | w in the staged
          // rows (copied behind the blob), c x and the basis rows through L1 from the profile's blob in global memory,
          // software-pipelined sweep (sweep_points_pf) where the model has one (length modulation)
          double* rows = smem + K.blob_stride;
          for (int e = threadIdx.x; e < 2 * K.npad; e += blockDim.x) {
            const int blk = e / 64, r = (e % 64) / 32, i = e % 32;
            rows[e] = smem[(size_t)blk * (3 + NN) * 32 + (1 + r) * 32 + i];
          }
          __syncthreads();
          const double* gb = K.blobs + (size_t)j * K.blob_stride;
          const int l = lane & 15, half = lane >> 4;
          for (int base = 0; base < K.n_q; base += 2 * nwarp) {
            const int iq = base + 2 * warp + half;
            const bool ok = iq < K.n_q;
            const size_t r = (size_t)j * K.n_q + (ok ? iq : 0);
            const double qd = ok && l < D ? K.q[r * D + l] : 0.0;
            const Eval ev = warp_logp_grad<NN, MOD, 16, 2>(rows, s_prob, K.spec, qd, l, gb);
            if (ok && l < D) K.grad[r * D + l] = ev.g;
            if (ok && l == 0) {
              K.lp[r] = ev.lp;
              if (K.chi2) K.chi2[r] = ev.chi2;
            }
          }
          __syncthreads();
          continue;
        }
      }
    }
    for (int iq = warp; iq < K.n_q; iq += nwarp) {
      const size_t r = (size_t)j * K.n_q + iq;
      const double qd = lane < D ? K.q[r * D + lane] : 0.0;
      const Eval ev = warp_logp_grad<NN, MOD>(smem, s_prob, K.spec, qd, lane);
      if (lane < D) K.grad[r * D + lane] = ev.g;
      if (lane == 0) {
        K.lp[r] = ev.lp;
        if (K.chi2) K.chi2[r] = ev.chi2;
      }
    }
    __syncthreads();
  }
}


// MAP of the ExpGP / MonoExp posterior by BFGS on the unconstrained space WITHOUT the Jacobian terms (what
// rstan::optimizing does by default: jacobian = FALSE), one warp per profile, lane d owns row d of the inverse
// Hessian approximation.  Followed by a central finite-difference Hessian of the gradient at the optimum, the
// way rstan builds `fit$hessian` (FitOCT.R:42 method 'optim'; ShinyInterface/server.R:164-173 reads it).
// MODEL_SPEC §10.  (The CPU checker used by the tests runs the identical iteration.)
template <int NN, int MOD>
__global__ void __launch_bounds__(32) map_bfgs_kernel(const MapParams K) {
  extern __shared__ __align__(128) double smem[];
  fill_exptab();
  __shared__ uint64_t mbar;
  __shared__ DevProblem s_prob;
  using DM = Dims<NN>;
  constexpr int D = DM::D;
  constexpr int P_OUT = DM::P_OUT;
  const int lane = threadIdx.x;
  const bool act = lane < D;
  mbar_init(&mbar);
  uint32_t phase = 0;
  for (int j = blockIdx.x; j < K.n_problems; j += gridDim.x) {
    if (lane == 0) s_prob = K.probs[j];
    stage_blob_tma(smem, K.blobs + (size_t)j * K.blob_stride, (uint32_t)(K.blob_stride * sizeof(double)), &mbar, phase, threadIdx.x == 0);
    __syncthreads();
    const DevProblem& P = s_prob;
    // objective f = -(lp - Jacobian); jac_d = 1 on the log-lambda / log-sigma lanes
    const double jac = (DM::GP && (lane == 3 + NN || lane == 4 + NN)) ? 1.0 : 0.0;
    auto eval = [&](double qd, double& f, double& gd, double& chi2) {
      const Eval ev = warp_logp_grad<NN, MOD>(smem, P, K.spec, qd, lane);
      const double jl = DM::GP ? bcast(qd, 3 + NN) + bcast(qd, 4 + NN) : 0.0;
      f = -(ev.lp - jl);
      gd = act ? -(ev.g - jac) : 0.0;
      chi2 = ev.chi2;
    };
    double q = 0.0;
    if (act) {
      if (K.init) q = K.init[(size_t)j * D + lane];
      else if (lane < 3) q = P.theta0[lane];
      else if (lane < 3 + NN) q = 0.0;
      else if (lane == 3 + NN) q = log(0.1);
      else q = 0.0;
    }
    // initial inverse Hessian: prior variances of theta, (0.05)^2 for the control values, 0.25 / 0.01 for the logs
    double h0 = 1.0;
    if (lane < 3) h0 = K.spec.theta_prior == 0 ? 1.0 / P.Pinv[lane * 4] : (lane == 2 ? 100.0 : 1.0e4);
    else if (lane < 3 + NN) h0 = 2.5e-3;
    else if (lane == 3 + NN) h0 = 0.25;
    else h0 = 0.01;
    double H[D];
#pragma unroll
    for (int c = 0; c < D; ++c) H[c] = (c == lane) ? h0 : 0.0;
    double f, g, chi2;
    eval(q, f, g, chi2);
    int status = 1, it = 0;
    for (; it < K.max_iter; ++it) {
      // d = -H g
      double dvec = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) dvec = fma(H[c], bcast(g, c), dvec);
      dvec = act ? -dvec : 0.0;
      double slope = warp_sum(g * dvec);
      if (!(slope < 0.0)) {  // not a descent direction: restart from the diagonal
#pragma unroll
        for (int c = 0; c < D; ++c) H[c] = (c == lane) ? h0 : 0.0;
        dvec = act ? -h0 * g : 0.0;
        slope = warp_sum(g * dvec);
        if (!(slope < 0.0)) { status = 0; break; }  // zero gradient
      }
      double step = 1.0, fn = 0.0, gn = 0.0, c2n = 0.0, qn = q;
      bool ok = false;
      for (int ls = 0; ls < 40; ++ls) {
        qn = fma(step, dvec, q);
        eval(qn, fn, gn, c2n);
        if (isfinite(fn) && fn <= f + 1e-4 * step * slope) { ok = true; break; }
        step *= 0.5;
      }
      if (!ok) { status = 2; break; }
      const double s = qn - q, y = gn - g;
      const double sy = warp_sum(s * y);
      const double df = f - fn;
      q = qn; g = gn; chi2 = c2n;
      const double fold = f;
      f = fn;
      if (df <= 1e-13 * (fabs(fold) + 1.0)) { status = 0; ++it; break; }
      const double ss = warp_sum(s * s), yy = warp_sum(y * y);
      if (sy > 1e-12 * sqrt(ss * yy)) {
        double Hy = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) Hy = fma(H[c], bcast(y, c), Hy);
        if (!act) Hy = 0.0;
        const double yHy = warp_sum(y * Hy);
        const double c1 = (sy + yHy) / (sy * sy), isy = 1.0 / sy;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const double sc = bcast(s, c), Hyc = bcast(Hy, c);
          H[c] = H[c] + c1 * s * sc - (Hy * sc + s * Hyc) * isy;
        }
      }
    }
    // outputs: constrained parameters, br, lp without Jacobian
    {
      double v = q;
      if (DM::GP && (lane == 3 + NN || lane == 4 + NN)) v = exp(q);
      if (lane == D) v = P.prior_PD ? CUDART_NAN : chi2 / P.br_ndf;
      if (lane == D + 1) v = -f;
      if (lane < P_OUT) K.par[(size_t)j * P_OUT + lane] = v;
      if (lane == 0) K.status[j] = status;
    }
    if (K.hessian) {
      for (int c = 0; c < D; ++c) {
        const double qc = bcast(q, c);
        const double h = 1e-5 * fmax(1.0, fabs(qc));
        double fp, gp, fm, gm, cc;
        eval(lane == c ? q + h : q, fp, gp, cc);
        eval(lane == c ? q - h : q, fm, gm, cc);
        if (act) K.hessian[((size_t)j * D + lane) * D + c] = -(gp - gm) / (2.0 * h);  // d2 lp / dq_lane dq_c
      }
    }
    __syncthreads();
  }
}

// method = 'vb': Stan's mean-field ADVI (MODEL_SPEC §14; FitOCT.R:42), one warp per profile, lane d owns component d of
// mu / omega and of the step-size history.  Every Monte-Carlo draw costs one fused log-density + gradient sweep; the
// draws come from per-site Philox counters, so the CPU checker consumes the same variates.
template <int NN, int MOD>
__global__ void __launch_bounds__(32) vb_kernel(const VbParams K) {
  extern __shared__ __align__(128) double smem[];
  fill_exptab();
  __shared__ uint64_t mbar;
  __shared__ DevProblem s_prob;
  using DM = Dims<NN>;
  constexpr int D = DM::D;
  constexpr int P_OUT = DM::P_OUT;
  constexpr uint32_t SITE_VB_GRAD = 5, SITE_VB_ELBO = 6, SITE_VB_OUT = 7;
  const int lane = threadIdx.x;
  const bool act = lane < D;
  mbar_init(&mbar);
  uint32_t phase_bit = 0;
  for (int j = blockIdx.x; j < K.n_problems; j += gridDim.x) {
    if (lane == 0) s_prob = K.probs[j];
    stage_blob_tma(smem, K.blobs + (size_t)j * K.blob_stride, (uint32_t)(K.blob_stride * sizeof(double)), &mbar, phase_bit, threadIdx.x == 0);
    __syncthreads();
    const DevProblem& P = s_prob;
    Rng rng;
    rng.seed(K.seed, P.id, 0);
    uint32_t rb[4];
    auto draw = [&](double mu, double om, uint32_t it, uint32_t kind, uint32_t a, uint32_t ph, double& eta) -> double {
      eta = 0.0;
      if (!act) return 0.0;
      rng.block(it, kind, a, (uint32_t)lane, ph, rb);
      eta = normal_from(rb);
      return mu + exp(om) * eta;
    };
    // gradient estimate; returns true when a draw was non-finite
    auto grad = [&](double mu, double om, uint32_t it, uint32_t ph, double& gmu, double& gom) -> bool {
      gmu = 0.0; gom = 0.0;
      bool bad = false;
      for (int s = 0; s < K.grad_samples; ++s) {
        double eta;
        const double zeta = draw(mu, om, it, SITE_VB_GRAD, (uint32_t)s, ph, eta);
        const Eval ev = warp_logp_grad<NN, MOD>(smem, P, K.spec, zeta, lane);
        bad |= !isfinite(ev.lp) || __any_sync(FOCT_FULL, act && !isfinite(ev.g));
        if (act) { gmu += ev.g; gom += ev.g * eta; }
      }
      gmu /= K.grad_samples; gom /= K.grad_samples;
      gom = act ? gom * exp(om) + 1.0 : 0.0;
      return bad;
    };
    auto elbo_est = [&](double mu, double om, uint32_t it, uint32_t ph) -> double {
      double sum = 0.0;
      int ok = 0, dropped = 0;
      for (uint32_t a = 0; ok < K.elbo_samples; ++a) {
        double eta;
        const double zeta = draw(mu, om, it, SITE_VB_ELBO, a, ph, eta);
        const Eval ev = warp_logp_grad<NN, MOD>(smem, P, K.spec, zeta, lane);
        if (isfinite(ev.lp)) { sum += ev.lp; ++ok; }
        else if (++dropped >= K.elbo_samples) return CUDART_NAN;
      }
      const double ent = 0.5 * D * (1.0 + 1.8378770664093454836) + warp_sum(act ? om : 0.0);
      return sum / K.elbo_samples + ent;
    };
    auto step = [&](double& mu, double& om, double& hmu, double& hom, double gmu, double gom, int k, double eta_s) {
      const double sc = eta_s / sqrt((double)k);
      hmu = k == 1 ? gmu * gmu : 0.9 * hmu + 0.1 * gmu * gmu;
      hom = k == 1 ? gom * gom : 0.9 * hom + 0.1 * gom * gom;
      if (act) {
        mu += sc * gmu / (1.0 + sqrt(hmu));
        om += sc * gom / (1.0 + sqrt(hom));
      }
    };
    // start (init_mode as the sampler's, MODEL_SPEC §7)
    double q0 = 0.0;
    if (act) {
      rng.block(0, SITE_INIT, 0, (uint32_t)lane, 0, rb);
      if (K.init_mode == 2 && K.init) q0 = K.init[(size_t)j * D + lane];
      else if (K.init_mode == 1) q0 = -2.0 + 4.0 * u53(rb[0], rb[1]);
      else if (lane < 3) q0 = P.theta0[lane];
      else if (lane < 3 + NN) q0 = 0.01 * normal_from(rb);
      else if (lane == 3 + NN) q0 = log(0.1);
      else q0 = 0.0;
    }
    double mu = q0, om = act ? K.omega0 : 0.0, hmu = 0.0, hom = 0.0, gmu, gom;
    double eta_s = K.eta, elbo = CUDART_NAN;
    int status = 1, iters = 0;
    bool failed = false;
    if (K.adapt_engaged) {
      const double elbo_init = elbo_est(mu, om, 0, 1);
      if (isnan(elbo_init)) failed = true;
      double elbo_best = -CUDART_INF, eta_best = 0.0;
      bool found = false;
      for (int e = 0; e < 5 && !failed; ++e) {
        const double eta_e = e == 0 ? 100.0 : e == 1 ? 10.0 : e == 2 ? 1.0 : e == 3 ? 0.1 : 0.01;
        mu = q0; om = act ? K.omega0 : 0.0; hmu = 0.0; hom = 0.0;
        for (int k = 1; k <= K.adapt_iter; ++k) {
          if (grad(mu, om, (uint32_t)k, (uint32_t)(1 + e), gmu, gom)) { gmu = 0.0; gom = 0.0; }
          step(mu, om, hmu, hom, gmu, gom, k, eta_e);
        }
        double el = elbo_est(mu, om, (uint32_t)K.adapt_iter, (uint32_t)(1 + e));
        if (isnan(el)) el = -CUDART_INF;
        if (el < elbo_best && elbo_best > elbo_init) { found = true; break; }
        if (e < 4) { elbo_best = el; eta_best = eta_e; }
        else if (el > elbo_init) { eta_best = eta_e; found = true; }
      }
      if (!found) failed = true;
      eta_s = eta_best;
    }
    if (failed) {
      status = 2; eta_s = CUDART_NAN;
    } else {
      mu = q0; om = act ? K.omega0 : 0.0; hmu = 0.0; hom = 0.0;
      int cap = (int)(0.1 * K.iter / K.eval_elbo);
      cap = cap < 2 ? 2 : (cap > 32 ? 32 : cap);
      double cbv = 0.0;  // lane i holds entry i of the circular buffer of relative ELBO changes
      int cbn = 0, cbpos = 0;
      double elbo_prev;
      elbo = 0.0;
      int k = 1;
      for (; k <= K.iter; ++k) {
        if (grad(mu, om, (uint32_t)k, 0, gmu, gom)) { status = 2; break; }
        step(mu, om, hmu, hom, gmu, gom, k, eta_s);
        if (k % K.eval_elbo == 0) {
          elbo_prev = elbo;
          elbo = elbo_est(mu, om, (uint32_t)k, 0);
          if (isnan(elbo)) { status = 2; break; }
          const double delta = fabs((elbo_prev - elbo) / elbo);
          if (lane == cbpos) cbv = delta;
          cbpos = (cbpos + 1) % cap;
          if (cbn < cap) ++cbn;
          const bool in = lane < cbn;
          const double mean = warp_sum(in ? cbv : 0.0) / cbn;
          int rank = 0;  // upper median = the entry with cbn/2 entries ordered before it
          for (int i = 0; i < cbn; ++i) {
            const double vi = bcast(cbv, i);
            rank += (vi < cbv) || (vi == cbv && i < lane);
          }
          const unsigned who = __ballot_sync(FOCT_FULL, in && rank == cbn / 2);
          const double med = bcast(cbv, __ffs(who) - 1);
          if (mean < K.tol_rel_obj || med < K.tol_rel_obj) { status = 0; break; }
        }
      }
      iters = k > K.iter ? K.iter : k;
    }
    if (act) { K.mu[(size_t)j * D + lane] = mu; K.omega[(size_t)j * D + lane] = om; }
    if (lane == 0) {
      if (K.elbo) K.elbo[j] = elbo;
      if (K.eta_out) K.eta_out[j] = eta_s;
      if (K.iters) K.iters[j] = iters;
      if (K.status) K.status[j] = status;
    }
    auto write_row = [&](double q, double* row) {
      const Eval ev = warp_logp_grad<NN, MOD>(smem, P, K.spec, q, lane);
      double v = q;
      if (DM::GP && (lane == 3 + NN || lane == 4 + NN)) v = exp(q);
      if (lane == D) v = P.prior_PD ? CUDART_NAN : ev.chi2 / P.br_ndf;
      if (lane == D + 1) v = 0.0;  // lp__ = 0 in ADVI output, as Stan writes it
      if (lane < P_OUT) row[lane] = v;
    };
    write_row(mu, K.mean + (size_t)j * P_OUT);
    if (K.draws)
      for (int i = 0; i < K.output_samples; ++i) {
        double eta;
        const double zeta = draw(mu, om, 0, SITE_VB_OUT, (uint32_t)i, 0, eta);
        write_row(zeta, K.draws + ((size_t)j * K.output_samples + i) * P_OUT);
      }
    __syncthreads();
  }
}

template <int NN>
static cudaError_t launch_vb(int mod, int grid, size_t smem, cudaStream_t st, const VbParams& K) {
  cudaError_t e;
  if (mod == 0) {
    e = cudaFuncSetAttribute(vb_kernel<NN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    vb_kernel<NN, 0><<<grid, 32, smem, st>>>(K);
  } else {
    e = cudaFuncSetAttribute(vb_kernel<NN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    vb_kernel<NN, 1><<<grid, 32, smem, st>>>(K);
  }
  return cudaGetLastError();
}

template <int NN>
static cudaError_t launch_map(int mod, int grid, size_t smem, cudaStream_t st, const MapParams& K) {
  cudaError_t e;
  if (mod == 0) {
    e = cudaFuncSetAttribute(map_bfgs_kernel<NN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    map_bfgs_kernel<NN, 0><<<grid, 32, smem, st>>>(K);
  } else {
    e = cudaFuncSetAttribute(map_bfgs_kernel<NN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    map_bfgs_kernel<NN, 1><<<grid, 32, smem, st>>>(K);
  }
  return cudaGetLastError();
}

// Two sampling kernels.  Nn <= 11 (D <= 16) and the mono-exponential can run two chains per warp (foct_nuts2.cuh): more
// gradients per second on a full GPU (3.0e8 against 2.3e8 at Nn = 10), but a chain advances at 60-70 % of the speed it has
// with a warp of its own.  So the one-chain-per-warp kernel serves the batches that fit the GPU in one go (measured, B200,
// 500 + 1000 iterations: 1 profile 1.44 s against 1.98 s, 148 profiles 1.70 against 2.54 s, 444 profiles 2.21 against
// 3.02 s; profiles/r2_kernel_experiments.txt) and the continuation rounds of a few unconverged profiles.
// FOCT_NO_PAIR=1 / FOCT_FORCE_PAIR=1 pin the choice (A/B runs, tests of the time slicing on small batches).
template <int NN>
static bool pair_capable() { return Dims<NN>::D <= 16; }
static int pair_env() {
  if (std::getenv("FOCT_NO_PAIR")) return 0;
  if (std::getenv("FOCT_FORCE_PAIR")) return 1;
  return -1;
}
template <int NN>
static int nuts_block(int chains, bool pair) {
  if (pair) return 32 * ((std::min(chains, FOCT_PAIR_CTA_CHAINS) + 1) / 2);
  return 32 * std::min(chains, FOCT_CTA_CHAINS);
}

template <int NN, int MOD>
static cudaError_t launch_nuts_mod(int grid, int block, size_t smem, cudaStream_t st, const SamplerParams& K) {
  cudaError_t e;
  if constexpr (Dims<NN>::D <= 16) {
    if (K.pair_kernel) {
      if (K.shared_basis && K.warp_units) {
        if constexpr (NN > 0) {
          e = cudaFuncSetAttribute(nuts2w_kernel<NN, MOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return e;
          nuts2w_kernel<NN, MOD><<<grid, block, smem, st>>>(K);
        }
      } else if (K.shared_basis) {
        e = cudaFuncSetAttribute(nuts2_kernel<NN, MOD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        nuts2_kernel<NN, MOD, 1><<<grid, block, smem, st>>>(K);
      } else {
        e = cudaFuncSetAttribute(nuts2_kernel<NN, MOD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        nuts2_kernel<NN, MOD, 0><<<grid, block, smem, st>>>(K);
      }
      return cudaGetLastError();
    }
  }
  // At most one work item per SM: the latency kernel (two warps per chain, one CTA per SM).  FOCT_NO_LAT=1: A/B runs and
  // the tests that pin the one-chain-per-warp kernel.
  if (grid <= sm_count() && !std::getenv("FOCT_NO_LAT")) {
    const size_t smem_blob = K.blob_stride * sizeof(double);  // (the latency kernel stages whole blobs whatever `smem` says)
    e = cudaFuncSetAttribute(nuts_lat_kernel<NN, MOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_blob);
    if (e == cudaSuccess) {
      // Few enough (profile, chain) pairs for a CTA - an SM - per chain: the four chains of a single profile run on four
      // SMs, every warp with a scheduler to itself (measured: 0.78 s against 0.95 s with the four chains on one SM);
      // else two chains per CTA (four warps = one per scheduler) while that still gives every CTA its own SM.
      const long long chain_items = (long long)K.n_problems * K.chains, pair_items = (long long)K.n_problems * ((K.chains + 1) / 2);
      if (chain_items <= sm_count() && K.chains > 1 &&
          cudaFuncSetAttribute(nuts_lat_kernel<NN, MOD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_blob) == cudaSuccess) {
        nuts_lat_kernel<NN, MOD, 1><<<(int)chain_items, 64, smem_blob, st>>>(K);
        return cudaGetLastError();
      }
      if (pair_items <= sm_count() && K.chains > 2 &&
          cudaFuncSetAttribute(nuts_lat_kernel<NN, MOD, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_blob) == cudaSuccess) {
        nuts_lat_kernel<NN, MOD, 2><<<(int)pair_items, 128, smem_blob, st>>>(K);
        return cudaGetLastError();
      }
      nuts_lat_kernel<NN, MOD><<<grid, 2 * block, smem_blob, st>>>(K);
      return cudaGetLastError();
    }
    cudaGetLastError();  // (a blob that leaves no room for the team buffers: the one-chain-per-warp kernel below)
  }
  e = cudaFuncSetAttribute(nuts_kernel<NN, MOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  nuts_kernel<NN, MOD><<<grid, block, smem, st>>>(K);
  return cudaGetLastError();
}

template <int NN>
static cudaError_t launch_nuts(int mod, int grid, int block, size_t smem, cudaStream_t st, const SamplerParams& K) {
  return mod == 0 ? launch_nuts_mod<NN, 0>(grid, block, smem, st, K) : launch_nuts_mod<NN, 1>(grid, block, smem, st, K);
}

template <int NN>
static cudaError_t launch_logp(int mod, int grid, int block, size_t smem, cudaStream_t st, const LogpParams& K) {
  cudaError_t e;
  if (mod == 0) {
    e = cudaFuncSetAttribute(logp_kernel<NN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    logp_kernel<NN, 0><<<grid, block, smem, st>>>(K);
  } else {
    e = cudaFuncSetAttribute(logp_kernel<NN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    logp_kernel<NN, 1><<<grid, block, smem, st>>>(K);
  }
  return cudaGetLastError();
}

template <class KernelT>
static cudaError_t occupancy_of(KernelT kernel, int block, size_t smem, int* blocks_per_sm, int* regs) {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, block, smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncGetAttributes(&fa, kernel);
  if (e == cudaSuccess && regs) *regs = fa.numRegs;
  return e;
}

// Launch geometry of the sampling kernel for `chains` chains per profile: threads per CTA, resident CTAs per SM,
// chains a CTA serves per work item, registers per thread.
// shared_basis in: the batch has one depth grid; out: whether this kernel reads the basis from global memory (then `smem`
// must be the size of the cx | y | w part only: the caller passes both sizes).
template <int NN>
static cudaError_t nuts_occupancy(int mod, int chains, long long n_items, int n_sm, size_t smem_full, size_t smem_rows,
                                  int* shared_basis, size_t* smem, int* block, int* blocks_per_sm, int* cta_chains, int* regs,
                                  size_t* slice_bytes, int* pair_kernel, int* warp_units) {
  *smem = smem_full;
  *warp_units = 0;
  *slice_bytes = 0;  // per work item: saved state of a CTA's warps (time slicing); 0 = this kernel runs items to completion
  *pair_kernel = 0;
  *cta_chains = FOCT_CTA_CHAINS;
  // geometry of the one-chain-per-warp kernel: what a batch must exceed to be worth two chains per warp
  *block = nuts_block<NN>(chains, false);
  int one_bpsm = 0, one_regs = 0;
  // (Staging c x | y | w only and reading the basis rows through L1, as the two-chains-per-warp kernels do, was measured
  // on this kernel too: Nn = 15 1.94e8 against 1.98e8 gradients/s, Nn = 20 1.41e8 against 1.53e8, Nn = 10 2.26e8 against
  // 2.32e8 - a warp that owns its chain reads 32 distinct points per row, and LDS serves that better than L1.  Removed.)
  cudaError_t e = mod == 0 ? occupancy_of(nuts_kernel<NN, 0>, *block, *smem, &one_bpsm, &one_regs)
                           : occupancy_of(nuts_kernel<NN, 1>, *block, *smem, &one_bpsm, &one_regs);
  if (e != cudaSuccess) {
    // (a profile too long for its blob to be staged whole can still run on the shared-basis kernel)
    if (!pair_capable<NN>()) return e;
    cudaGetLastError();
    one_bpsm = 0;
  }
  if constexpr (Dims<NN>::D <= 16) {
    const int env = pair_env();
    const long long groups = (chains + FOCT_CTA_CHAINS - 1) / FOCT_CTA_CHAINS;
    // (measured crossover: 600 profiles still finish sooner in two ragged waves of the one-chain kernel, 3.45 s against 3.65 s)
    const bool pair = one_bpsm < 1 || (env >= 0 ? env == 1 : 10 * n_items * groups > 16 * (long long)n_sm * one_bpsm);
    if (pair) {
      *pair_kernel = 1;
      *block = nuts_block<NN>(chains, true);
      *cta_chains = FOCT_PAIR_CTA_CHAINS;
      *slice_bytes = (size_t)(FOCT_PAIR_CTA_CHAINS / 2) * FOCT_PAIR_STATE_DOUBLES * 32 * sizeof(double);
      // One depth grid for the batch = one GP basis for every profile: only cx | y | w are staged per item (12 KB), the
      // basis rows come through L1 from blob 0, and twelve warps fit an SM instead of eight.  With 64-bit loads that bought
      // nothing (2.58e8 vs 2.61e8 gradients/s at full waves: twice the L1 requests of the staged variant's LDS); with one
      // 128-bit load per row and point pair it is 3.0e8 (profiles/r2_kernel_experiments.txt).  FOCT_NO_SHARED_BASIS=1: A/B.
      const bool use_gb = std::getenv("FOCT_NO_SHARED_BASIS") == nullptr;  // (read per plan: tests flip it)
      if (*shared_basis && NN > 0 && use_gb) {
        *shared_basis = 1;
        if constexpr (NN > 0) {
          if (!std::getenv("FOCT_CTA_ITEMS")) {  // (A/B: the CTA-level items of nuts2_kernel<.., 1>)
            // warps as the scheduling unit: two warps per CTA whatever the chain count, each with its own staged rows
            *slice_bytes = (size_t)FOCT_PAIR_STATE_DOUBLES * 32 * sizeof(double);  // per unit = per warp
            *warp_units = 2;
            *block = 64;
            *smem = 2 * (FOCT_CX_SHARED ? smem_rows / 3 * 2 : smem_rows);  // per warp: y | w (c x is shared too), or c x | y | w
            if (FOCT_CX_SHARED) *smem += (size_t)FOCT_STACK_SMEM_LEVELS(NN) * 64 * 7 * sizeof(double);  // the busiest levels of the subtree stacks
            return mod == 0 ? occupancy_of(nuts2w_kernel<NN, 0>, *block, *smem, blocks_per_sm, regs)
                            : occupancy_of(nuts2w_kernel<NN, 1>, *block, *smem, blocks_per_sm, regs);
          }
        }
        *smem = smem_rows;
        return mod == 0 ? occupancy_of(nuts2_kernel<NN, 0, 1>, *block, *smem, blocks_per_sm, regs)
                        : occupancy_of(nuts2_kernel<NN, 1, 1>, *block, *smem, blocks_per_sm, regs);
      }
      *shared_basis = 0;
      return mod == 0 ? occupancy_of(nuts2_kernel<NN, 0, 0>, *block, *smem, blocks_per_sm, regs)
                      : occupancy_of(nuts2_kernel<NN, 1, 0>, *block, *smem, blocks_per_sm, regs);
    }
  }
  *shared_basis = 0;
  *blocks_per_sm = one_bpsm;
  *regs = one_regs;
  return cudaSuccess;
}

#define FOCT_CAT_(a, b) a##b
#define FOCT_CAT(a, b) FOCT_CAT_(a, b)

const InstEntry* FOCT_CAT(foct_inst_, FOCT_INST_NN)() {
  static const InstEntry e = {FOCT_INST_NN, &launch_nuts<FOCT_INST_NN>, &launch_logp<FOCT_INST_NN>,
                              &nuts_occupancy<FOCT_INST_NN>, &launch_map<FOCT_INST_NN>, &launch_vb<FOCT_INST_NN>};
  return &e;
}

}  // namespace foct

#ifdef FOCT_TIMING
extern "C" int foct_debug_timing(unsigned long long* out, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out, foct::g_tim, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(foct::g_tim, z, sizeof(z)); }
  return (int)e;
}
#endif
