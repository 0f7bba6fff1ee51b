// foct_summary.cuh — on-device posterior summaries (MODEL_SPEC §9; SURVEY a-11, build-plan step 7):
// mean, se_mean, sd, five quantiles, n_eff (Stan's Geyer estimator), split-Rhat and rank-normalised
// Bulk_ESS, one CTA per (profile, output column).  Replaces what rstan::summary computes on the host from
// the stanfit (ShinyInterface/server.R:88-104) so that a 1e5-profile batch never ships its draws.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/fitoct_b200.h"

namespace foct {

constexpr int SUM_THREADS = 256;

// Sum of two values over the CTA, result broadcast to every thread.
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch /*[2*8+2]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // scratch reuse
  if (l == 0) { scratch[w] = a; scratch[8 + w] = b; }
  __syncthreads();
  if (threadIdx.x < 32) {
    double x = l < SUM_THREADS / 32 ? scratch[l] : 0.0, y = l < SUM_THREADS / 32 ? scratch[8 + l] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      x += __shfl_xor_sync(0xffffffffu, x, o);
      y += __shfl_xor_sync(0xffffffffu, y, o);
    }
    if (l == 0) { scratch[16] = x; scratch[17] = y; }
  }
  __syncthreads();
  a = scratch[16]; b = scratch[17];
}

// AS241 PPND16 inverse normal CDF (same algorithm as the oracle).
__device__ inline double inv_norm_cdf(double p) {
  const double q = p - 0.5;
  double r, val;
  if (fabs(q) <= 0.425) {
    r = 0.180625 - q * q;
    return q * (((((((2.5090809287301226727e3 * r + 3.3430575583588128105e4) * r + 6.7265770927008700853e4) * r + 4.5921953931549871457e4) * r + 1.3731693765509461125e4) * r + 1.9715909503065514427e3) * r + 1.3314166789178437745e2) * r + 3.3871328727963666080e0) /
           (((((((5.2264952788528545610e3 * r + 2.8729085735721942674e4) * r + 3.9307895800092710610e4) * r + 2.1213794301586595867e4) * r + 5.3941960214247511077e3) * r + 6.8718700749205790830e2) * r + 4.2313330701600911252e1) * r + 1.0);
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

// Stan's compute_effective_sample_size over CC chains of length len stored chain-major in x, which is
// CENTRED IN PLACE.  Streaming form of the Geyer initial-positive / initial-monotone sequence estimator.
__device__ double ess_block(double* x, int len, int CC, double* scratch) {
  if (len < 4) return CUDART_NAN;
  const int S = len * CC;
  double mean_var = 0.0, cm_sum = 0.0;
  double cmv[2 * FOCT_MAX_CHAINS];
  for (int c = 0; c < CC; ++c) {
    double s = 0.0, dummy = 0.0;
    for (int t = threadIdx.x; t < len; t += SUM_THREADS) s += x[c * len + t];
    block_sum2(s, dummy, scratch);
    const double m = s / len;
    double a0 = 0.0;
    dummy = 0.0;
    for (int t = threadIdx.x; t < len; t += SUM_THREADS) {
      const double d = x[c * len + t] - m;
      x[c * len + t] = d;
      a0 = fma(d, d, a0);
    }
    block_sum2(a0, dummy, scratch);
    mean_var += (a0 / len) * len / (len - 1.0);
    cm_sum += m;
    cmv[c] = m;
  }
  __syncthreads();
  mean_var /= CC;
  double var_plus = mean_var * (len - 1.0) / len;
  if (CC > 1) {
    const double mbar = cm_sum / CC;
    double v = 0.0;
    for (int c = 0; c < CC; ++c) v += (cmv[c] - mbar) * (cmv[c] - mbar);
    var_plus += v / (CC - 1.0);
  }
  if (!(var_plus > 0.0) || !isfinite(var_plus)) return CUDART_NAN;
  auto pair_acov = [&](int lag_a, int lag_b, double& ra, double& rb) {
    double a = 0.0, b = 0.0;
    for (int e = threadIdx.x; e < S; e += SUM_THREADS) {
      const int t = e % len;
      const double v = x[e];
      if (t + lag_a < len) a = fma(v, x[e + lag_a], a);
      if (t + lag_b < len) b = fma(v, x[e + lag_b], b);
    }
    block_sum2(a, b, scratch);
    ra = 1.0 - (mean_var - a / len / CC) / var_plus;
    rb = 1.0 - (mean_var - b / len / CC) / var_plus;
  };
  double rho_even = 1.0, rho_odd, dummy;
  pair_acov(1, 1, rho_odd, dummy);
  double sum_pairs = rho_even + rho_odd, prevP = sum_pairs;
  double last_raw = 0.0, last_adj = 0.0;
  bool last_stored = false;
  int s = 1;
  while (s < len - 4 && (rho_even + rho_odd) > 0.0) {
    pair_acov(s + 1, s + 2, rho_even, rho_odd);
    last_stored = false;
    if (rho_even + rho_odd >= 0.0) {
      const double raw = rho_even + rho_odd;
      const double adj = raw > prevP ? prevP : raw;
      sum_pairs += adj;
      prevP = adj;
      last_raw = raw; last_adj = adj; last_stored = true;
    }
    s += 2;
  }
  // Stan's monotone pass stops one pair short of max_s: a stored final pair keeps its raw value
  if (last_stored) sum_pairs += last_raw - last_adj;
  const double extra = rho_even > 0.0 ? rho_even : 0.0;
  const double nt = (double)S;
  // rstan's ess_rfun: tau_hat = max(tau_hat, 1 / log10(S)) — caps the ESS at S log10(S) and keeps a strongly antithetic
  // column (tau <= 0) finite and positive
  const double tau = fmax(-1.0 + 2.0 * sum_pairs + extra, 1.0 / log10(nt));
  return nt / tau;
}

// Dynamic shared memory layout: A[S] doubles | W[Spad] doubles | I[Spad] ints.
// in_slot / out_row (either may be nullptr = identity): item group g reads the draw block in_slot[g] and writes summary
// row out_row[g] — the run-until-converged rounds summarise a subset of profiles out of their extension blocks.
__global__ void __launch_bounds__(SUM_THREADS) summary_kernel(const double* __restrict__ draws, int n_saved, int off,
                                                              int n, int C, int P_out, int Spad, int n_items,
                                                              double* __restrict__ out, double* __restrict__ gwork,
                                                              const int* __restrict__ in_slot,
                                                              const int* __restrict__ out_row) {
  extern __shared__ __align__(16) unsigned char sraw[];
  __shared__ double scratch[18];
  __shared__ int s_bad;
  __shared__ int drop_pos[FOCT_MAX_CHAINS];
  const int S = n * C;
  double *A, *W;
  int* I;
  if (gwork) {  // global-memory workspace for very long chains
    unsigned char* base = (unsigned char*)gwork + (size_t)blockIdx.x * ((size_t)S * 8 + (size_t)Spad * 12);
    A = (double*)base; W = A + S; I = (int*)(W + Spad);
  } else {
    A = (double*)sraw; W = A + S; I = (int*)(W + Spad);
  }
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
  const int grp = item / P_out, p = item % P_out;
  const int prob = in_slot ? in_slot[grp] : grp;
  double* o = out + ((size_t)(out_row ? out_row[grp] : grp) * P_out + p) * FOCT_N_SUMMARY_COLS;
  __syncthreads();
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  // 1. gather the column, chain-major
  const double* src = draws + ((size_t)prob * n_saved + off) * C * P_out + p;
  bool bad = false;
  for (int e = threadIdx.x; e < S; e += SUM_THREADS) {
    const int c = e / n, t = e % n;
    const double v = src[((size_t)t * C + c) * P_out];
    A[e] = v;
    if (!isfinite(v)) bad = true;
  }
  if (bad) s_bad = 1;
  __syncthreads();
  if (s_bad) {
    if (threadIdx.x < FOCT_N_SUMMARY_COLS) o[threadIdx.x] = CUDART_NAN;
    continue;
  }
  // 2. mean, sd
  double sm = 0.0, dummy = 0.0;
  for (int e = threadIdx.x; e < S; e += SUM_THREADS) sm += A[e];
  block_sum2(sm, dummy, scratch);
  const double mean = sm / S;
  double ss = 0.0;
  dummy = 0.0;
  for (int e = threadIdx.x; e < S; e += SUM_THREADS) { const double d = A[e] - mean; ss = fma(d, d, ss); }
  block_sum2(ss, dummy, scratch);

  // 4. bitonic sort of (value, index)
  for (int e = threadIdx.x; e < Spad; e += SUM_THREADS) { W[e] = e < S ? A[e] : CUDART_INF; I[e] = e; }
  __syncthreads();
  for (int k = 2; k <= Spad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int e = threadIdx.x; e < Spad; e += SUM_THREADS) {
        const int x = e ^ j;
        if (x > e) {
          const double a = W[e], b = W[x];
          const int ia = I[e], ib = I[x];
          const bool up = (e & k) == 0;
          const bool gt = a > b || (a == b && ia > ib);
          if (gt == up) { W[e] = b; W[x] = a; I[e] = ib; I[x] = ia; }
        }
      }
      __syncthreads();
    }
  }
  // 5. quantiles (R type 7)
  if (threadIdx.x < 5) {
    const double probs[5] = {0.025, 0.25, 0.5, 0.75, 0.975};
    const double hq = (S - 1) * probs[threadIdx.x];
    const int lo = (int)floor(hq);
    const int hi = lo + 1 < S ? lo + 1 : lo;
    o[3 + threadIdx.x] = W[lo] + (hq - lo) * (W[hi] - W[lo]);
  }
  // 3. n_eff on the unsplit chains (a column whose min == max is constant: sd 0, diagnostics NaN, like rstan)
  __syncthreads();
  const bool constant = W[0] == W[S - 1];
  __syncthreads();
  double n_eff = CUDART_NAN;
  if (constant) ss = 0.0;
  if (!constant) {
    for (int e = threadIdx.x; e < S; e += SUM_THREADS) W[e] = A[e];
    __syncthreads();
    n_eff = ess_block(W, n, C, scratch);
  }
  __syncthreads();
  // 6. split-Rhat on the raw values
  const int h = n / 2;
  double rhat = CUDART_NAN, bulk = CUDART_NAN;
  const double sd = S > 1 ? sqrt(ss / (S - 1.0)) : CUDART_NAN;
  if (h >= 2 && !constant) {
    const int C2 = 2 * C;
    double Wv = 0.0, gm = 0.0, gm2 = 0.0;
    double cmn[2 * FOCT_MAX_CHAINS];
    for (int c2 = 0; c2 < C2; ++c2) {
      const int c = c2 >> 1, t0 = (c2 & 1) ? n - h : 0;
      double s1 = 0.0;
      dummy = 0.0;
      for (int t = threadIdx.x; t < h; t += SUM_THREADS) s1 += A[c * n + t0 + t];
      block_sum2(s1, dummy, scratch);
      const double m = s1 / h;
      double v = 0.0;
      dummy = 0.0;
      for (int t = threadIdx.x; t < h; t += SUM_THREADS) { const double d = A[c * n + t0 + t] - m; v = fma(d, d, v); }
      block_sum2(v, dummy, scratch);
      Wv += v / (h - 1.0);
      cmn[c2] = m;
      gm += m;
    }
    Wv /= C2; gm /= C2;
    for (int c2 = 0; c2 < C2; ++c2) gm2 += (cmn[c2] - gm) * (cmn[c2] - gm);
    const double Bv = gm2 / (C2 - 1.0);
    rhat = sqrt((Wv * (h - 1.0) / h + Bv) / Wv);
    // 7. Bulk_ESS: rank-normalise the split chains (dropped middle draws of odd-length chains get no rank)
    __syncthreads();
    const bool odd = (n & 1) != 0;
    if (odd) {
      for (int r = threadIdx.x; r < S; r += SUM_THREADS) {
        const int e = I[r];
        if (e % n == h) drop_pos[e / n] = r;
      }
    }
    __syncthreads();
    const int S2 = C2 * h;
    for (int r = threadIdx.x; r < S; r += SUM_THREADS) {
      const int e = I[r], c = e / n, t = e % n;
      if (odd && t == h) continue;
      int r2 = r;
      if (odd)
        for (int cc = 0; cc < C; ++cc) r2 -= drop_pos[cc] < r ? 1 : 0;
      const int si = t < h ? (2 * c) * h + t : (2 * c + 1) * h + (t - (n - h));
      A[si] = inv_norm_cdf(((double)(r2 + 1) - 0.375) / ((double)S2 + 0.25));
    }
    __syncthreads();
    bulk = ess_block(A, h, C2, scratch);
  }
  if (threadIdx.x == 0) {
    o[0] = mean; o[1] = sd / sqrt(n_eff); o[2] = sd; o[8] = n_eff; o[9] = rhat; o[10] = bulk;
  }
  }  // item loop
}

// Launch one CTA per (profile, column).  Workspace lives in shared memory when it fits, else in a
// caller-independent global scratch buffer allocated here and released after the kernel.
static cudaError_t launch_summary(const double* d_draws, int n_problems, int n_saved, int off, int n_post, int C,
                                  int P_out, double* d_summary, cudaStream_t st, const int* in_slot = nullptr,
                                  const int* out_row = nullptr) {
  const int S = n_post * C;
  int Spad = 1;
  while (Spad < S) Spad <<= 1;
  const size_t bytes = (size_t)S * 8 + (size_t)Spad * 12;
  const int n_items = n_problems * P_out;
  cudaError_t e;
  if (bytes <= 200 * 1024) {
    e = cudaFuncSetAttribute(summary_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    const int grid = n_items < sm_count() * 64 ? n_items : sm_count() * 64;
    summary_kernel<<<grid, SUM_THREADS, bytes, st>>>(d_draws, n_saved, off, n_post, C, P_out, Spad, n_items, d_summary, nullptr, in_slot, out_row);
    return cudaGetLastError();
  }
  double* gwork = nullptr;
  const int grid = n_items < sm_count() * 4 ? n_items : sm_count() * 4;
  e = cudaMallocAsync((void**)&gwork, bytes * (size_t)grid, st);
  if (e != cudaSuccess) return e;
  summary_kernel<<<grid, SUM_THREADS, 0, st>>>(d_draws, n_saved, off, n_post, C, P_out, Spad, n_items, d_summary, gwork, in_slot, out_row);
  e = cudaGetLastError();
  cudaFreeAsync(gwork, st);
  return e;
}

}  // namespace foct
