// foct_inst.cu — instantiates the sampling and parity kernels for ONE control-point count
// (-DFOCT_INST_NN=<NN>, NN = 0 is the mono-exponential) and both modulation sites.
#include "foct_launch.h"

#ifndef FOCT_INST_NN
#error "compile with -DFOCT_INST_NN=<NN>"
#endif

namespace foct {

// Parity hook kernel: one CTA per profile, each warp evaluates log density + gradient at q points.
template <int NN, int MOD>
__global__ void __launch_bounds__(128) logp_kernel(const LogpParams K) {
  extern __shared__ __align__(128) double smem[];
  __shared__ uint64_t mbar;
  __shared__ DevProblem s_prob;
  constexpr int D = Dims<NN>::D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  mbar_init(&mbar);
  uint32_t phase = 0;
  for (int j = blockIdx.x; j < K.n_problems; j += gridDim.x) {
    if (threadIdx.x == 0) s_prob = K.probs[j];
    stage_blob_tma(smem, K.blobs + (size_t)j * K.blob_stride, (uint32_t)(K.blob_stride * sizeof(double)), &mbar, phase);
    __syncthreads();
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
    stage_blob_tma(smem, K.blobs + (size_t)j * K.blob_stride, (uint32_t)(K.blob_stride * sizeof(double)), &mbar, phase);
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

template <int NN>
static cudaError_t launch_nuts(int mod, int grid, int block, size_t smem, cudaStream_t st, const SamplerParams& K) {
  cudaError_t e;
  if (mod == 0) {
    e = cudaFuncSetAttribute(nuts_kernel<NN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    nuts_kernel<NN, 0><<<grid, block, smem, st>>>(K);
  } else {
    e = cudaFuncSetAttribute(nuts_kernel<NN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    nuts_kernel<NN, 1><<<grid, block, smem, st>>>(K);
  }
  return cudaGetLastError();
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

template <int NN>
static cudaError_t nuts_occupancy(int mod, int block, size_t smem, int* blocks_per_sm, int* regs) {
  cudaFuncAttributes fa;
  cudaError_t e;
  if (mod == 0) {
    e = cudaFuncSetAttribute(nuts_kernel<NN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, nuts_kernel<NN, 0>, block, smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&fa, nuts_kernel<NN, 0>);
  } else {
    e = cudaFuncSetAttribute(nuts_kernel<NN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, nuts_kernel<NN, 1>, block, smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&fa, nuts_kernel<NN, 1>);
  }
  if (e == cudaSuccess && regs) *regs = fa.numRegs;
  return e;
}

#define FOCT_CAT_(a, b) a##b
#define FOCT_CAT(a, b) FOCT_CAT_(a, b)

const InstEntry* FOCT_CAT(foct_inst_, FOCT_INST_NN)() {
  static const InstEntry e = {FOCT_INST_NN, &launch_nuts<FOCT_INST_NN>, &launch_logp<FOCT_INST_NN>,
                              &nuts_occupancy<FOCT_INST_NN>, &launch_map<FOCT_INST_NN>};
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
