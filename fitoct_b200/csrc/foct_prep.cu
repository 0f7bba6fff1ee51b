// foct_prep.cu — the steps either side of the sampling path, batched on the device (MODEL_SPEC §11-13):
//   estimateNoise     FitOCT.R:89-91   smoothing spline at `df` -> residuals -> uy(x) = a_1 exp(-x/a_2)
//   printBr (gate)    FitOCT.R:100, plotMonoExp.R:10   Birge ratio against the reduced chi-square 95 % interval
//   estimateExpPrior  FitOCT.R:103-107 theta0 / Sigma0 for fitExpGP from the MonoExp MAP fit ('mono' | 'abc')
// and foct_pipeline(), the body of FitOCT.R's dataset loop for a whole batch (noise -> MonoExp MAP -> gate -> prior ->
// fitExpGP on the gated profiles).  One warp per profile; the banded spline algebra lives in shared memory.
// No CPU compute path: every entry point but foct_birge_ci (a scalar function of ndf) needs a CUDA device.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include <math_constants.h>

#include "../../include/fitoct_b200.h"
#include "foct_launch.h"

namespace foct {

#define PREP_FULL 0xffffffffu

struct PrepMeta {
  size_t off;      // x at up[off], y at up[off + N]
  size_t out_off;  // packed per-point outputs start here
  int N, nknots, dataType, pad;
};

// R's .nknots.smspl (MODEL_SPEC §11); host only.
static int nknots_smspl(int n) {
  if (n < 50) return n;
  const double a1 = std::log2(50.0), a2 = std::log2(100.0), a3 = std::log2(140.0), a4 = std::log2(200.0);
  if (n < 200) return (int)(std::exp2(a1 + (a2 - a1) * (n - 50) / 150.0) + 1e-9);
  if (n < 800) return (int)(std::exp2(a2 + (a3 - a2) * (n - 200) / 600.0) + 1e-9);
  if (n < 3200) return (int)(std::exp2(a3 + (a4 - a3) * (n - 800) / 2400.0) + 1e-9);
  return (int)(200.0 + std::pow((double)(n - 3200), 0.2) + 1e-9);
}

// 0-based index of the data point that is inner knot k: x[floor(seq(1, N, length.out = nknots))]
__device__ __forceinline__ int knot_index(int k, int N, int nkn) {
  if (k >= nkn - 1) return N - 1;
  return (int)floor(1.0 + (double)k * (double)(N - 1) / (double)(nkn - 1)) - 1;
}

// the four cubic B-splines that are non-zero on knot interval l (Cox - de Boor)
__device__ __forceinline__ void bspl_val(const double* T, int l, double t, double B[4]) {
  double left[4], right[4];
  B[0] = 1.0;
#pragma unroll
  for (int k = 1; k <= 3; ++k) {
    left[k] = t - T[l + 1 - k];
    right[k] = T[l + k] - t;
    double saved = 0.0;
#pragma unroll
    for (int r = 0; r < k; ++r) {
      const double tmp = B[r] / (right[r + 1] + left[k - r]);
      B[r] = saved + right[r + 1] * tmp;
      saved = left[k - r] * tmp;
    }
    B[k] = saved;
  }
}

__device__ __forceinline__ double sdiv(double a, double b) { return b > 0.0 ? a / b : 0.0; }

// their second derivatives, from the polynomial piece of interval l
__device__ void bspl_d2(const double* T, int l, double t, double D2[4]) {
  const double h = T[l + 1] - T[l];
  const double b2[2] = {(T[l + 1] - t) / h, (t - T[l]) / h};
  double d3[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int j = l - 2 + a;
    const double u = a >= 1 ? b2[a - 1] : 0.0;  // B_{j,2}, non-zero for j = l-1, l
    const double v = a <= 1 ? b2[a] : 0.0;      // B_{j+1,2}
    d3[a] = 2.0 * (sdiv(u, T[j + 2] - T[j]) - sdiv(v, T[j + 3] - T[j + 1]));
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int j = l - 3 + a;
    const double u = a >= 1 ? d3[a - 1] : 0.0;
    const double v = a <= 2 ? d3[a] : 0.0;
    D2[a] = 3.0 * (sdiv(u, T[j + 3] - T[j]) - sdiv(v, T[j + 4] - T[j + 1]));
  }
}

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PREP_FULL, v, o);
  return v;
}

// Banded symmetric storage M[d*nk + j] = M_{j,j+d}, d = 0..3.  Lane 0 factors XtX + lam*Om = L D L^T, solves for the
// coefficients and accumulates df = tr[(XtX + lam Om)^-1 XtX] from the band of the inverse (Takahashi recurrence).
// Two sweeps with every recurrence carried in registers (a window of three columns), so the dependent chain per column
// is a few FMAs and one division instead of ~40 shared-memory round trips:
//   up:   column j of L and d, fused with the forward substitution;   stores l1,l2,l3,d,w=z/d per column
//   down: back substitution fused with the inverse band (6-value window) and the trace
// Ls: [5*nk] = l1 | l2 | l3 | d | w.  Returns df to all lanes; the coefficients are left in c.
__device__ double spl_fit(int nk, const double* XtX, const double* Om, const double* Xty, double* Ls, double* c,
                          double lam, int lane) {
  double df = 0.0;
  if (lane == 0) {
    double* l1 = Ls; double* l2 = Ls + nk; double* l3 = Ls + 2 * nk; double* dd = Ls + 3 * nk; double* w = Ls + 4 * nk;
    // window: columns j-1 (a), j-2 (b), j-3 (e)
    double a1 = 0, a2 = 0, a3 = 0, ad = 0, b2 = 0, b3 = 0, bd = 0, e3 = 0, ed = 0;
    double za = 0, zb = 0, ze = 0;  // z_{j-1}, z_{j-2}, z_{j-3}
    for (int j = 0; j < nk; ++j) {
      const double A0 = XtX[j] + lam * Om[j];
      const double A1 = XtX[nk + j] + lam * Om[nk + j];
      const double A2 = XtX[2 * nk + j] + lam * Om[2 * nk + j];
      const double A3 = XtX[3 * nk + j] + lam * Om[3 * nk + j];
      double dj = A0;
      dj -= e3 * e3 * ed;
      dj -= b2 * b2 * bd;
      dj -= a1 * a1 * ad;
      double s1 = A1;
      s1 -= b3 * b2 * bd;
      s1 -= a2 * a1 * ad;
      double s2 = A2;
      s2 -= a3 * a1 * ad;
      const double n1 = j + 1 < nk ? s1 / dj : 0.0;
      const double n2 = j + 2 < nk ? s2 / dj : 0.0;
      const double n3 = j + 3 < nk ? A3 / dj : 0.0;
      double z = Xty[j];
      z -= e3 * ze;
      z -= b2 * zb;
      z -= a1 * za;
      l1[j] = n1; l2[j] = n2; l3[j] = n3; dd[j] = dj; w[j] = z / dj;
      e3 = b3; ed = bd; b2 = a2; b3 = a3; bd = ad; a1 = n1; a2 = n2; a3 = n3; ad = dj;
      ze = zb; zb = za; za = z;
    }
    // window of the inverse band on rows/cols i+1..i+3, and of the coefficients
    double z11 = 0, z12 = 0, z13 = 0, z22 = 0, z23 = 0, z33 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (int i = nk - 1; i >= 0; --i) {
      const double m1 = l1[i], m2 = l2[i], m3 = l3[i];
      double ci = w[i];
      ci -= m1 * c1;
      ci -= m2 * c2;
      ci -= m3 * c3;
      c[i] = ci;
      double y3 = 0.0;  // Z_{i,i+3}
      y3 -= m1 * z13; y3 -= m2 * z23; y3 -= m3 * z33;
      double y2 = 0.0;  // Z_{i,i+2}
      y2 -= m1 * z12; y2 -= m2 * z22; y2 -= m3 * z23;
      double y1 = 0.0;  // Z_{i,i+1}
      y1 -= m1 * z11; y1 -= m2 * z12; y1 -= m3 * z13;
      double y0 = 1.0 / dd[i];  // Z_{i,i}
      y0 -= m1 * y1; y0 -= m2 * y2; y0 -= m3 * y3;
      df += y0 * XtX[i];
      if (i + 1 < nk) df += 2.0 * y1 * XtX[nk + i];
      if (i + 2 < nk) df += 2.0 * y2 * XtX[2 * nk + i];
      if (i + 3 < nk) df += 2.0 * y3 * XtX[3 * nk + i];
      z33 = z22; z23 = z12; z22 = z11; z13 = y2; z12 = y1; z11 = y0;
      c3 = c2; c2 = c1; c1 = ci;
    }
  }
  __syncwarp();
  return __shfl_sync(PREP_FULL, df, 0);
}

// One warp per profile.  Shared memory per CTA: (16*nkmax + 4) doubles: T | XtX 4 | Om 4 | Xty | L 5 | c.
__global__ void __launch_bounds__(32) noise_kernel(const double* __restrict__ up, const PrepMeta* __restrict__ meta, int n,
                                                   int nkmax, double df_target, double max_rate, double* __restrict__ uy,
                                                   double* __restrict__ ys, double* __restrict__ theta,
                                                   double* __restrict__ info, int* __restrict__ status) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x;
  double* T = sm;
  double* XtX = T + nkmax + 4;
  double* Om = XtX + 4 * nkmax;
  double* Xty = Om + 4 * nkmax;
  double* L = Xty + nkmax;   // l1 | l2 | l3 | d | w per column
  double* c = L + 5 * nkmax;
  for (int j = blockIdx.x; j < n; j += gridDim.x) {
    const PrepMeta M = meta[j];
    const int N = M.N, nkn = M.nknots, nk = nkn + 2;
    const double* x = up + M.off;
    const double* y = x + N;
    double* ysj = ys + M.out_off;
    double* uyj = uy + M.out_off;
    int bad = 0;
    for (int i = lane + 1; i < N; i += 32) bad |= !(x[i] > x[i - 1]);
    if (__any_sync(PREP_FULL, bad)) {  // x must increase strictly (R collapses ties; the FitOCT depth grids have none)
      for (int i = lane; i < N; i += 32) { ysj[i] = CUDART_NAN; uyj[i] = CUDART_NAN; }
      if (lane == 0) { status[j] = 3; theta[2 * j] = theta[2 * j + 1] = CUDART_NAN; }
      continue;
    }
    const double x0 = x[0], ir = 1.0 / (x[N - 1] - x[0]);
    for (int k = lane; k < nkn; k += 32) T[3 + k] = (x[knot_index(k, N, nkn)] - x0) * ir;
    for (int k = lane; k < 4 * nk; k += 32) { XtX[k] = 0.0; Om[k] = 0.0; }
    for (int k = lane; k < nk; k += 32) Xty[k] = 0.0;
    __syncwarp();
    if (lane < 3) { T[lane] = T[3]; T[nk + 1 + lane] = T[nk]; }
    __syncwarp();
    // design and penalty, one knot interval per lane; intervals m and m' share basis functions iff |m - m'| <= 3,
    // so the adds into the band go in four phases (deterministic summation order)
    for (int base = 3; base < nk; base += 32) {
      const int m = base + lane;
      const bool act = m < nk;
      double a[10], b[4], o[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) { a[k] = 0.0; o[k] = 0.0; }
#pragma unroll
      for (int k = 0; k < 4; ++k) b[k] = 0.0;
      if (act) {
        const int i0 = knot_index(m - 3, N, nkn), i1 = (m == nk - 1) ? N : knot_index(m - 2, N, nkn);
        for (int i = i0; i < i1; ++i) {
          const double t = (x[i] - x0) * ir;
          double B[4];
          bspl_val(T, m, t, B);
          const double yi = y[i];
          int q = 0;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            b[u] += B[u] * yi;
#pragma unroll
            for (int v = u; v < 4; ++v) a[q++] += B[u] * B[v];
          }
        }
        const double h = T[m + 1] - T[m];
        double A2[4], B2[4];
        bspl_d2(T, m, T[m], A2);
        bspl_d2(T, m, T[m + 1], B2);
        int q = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = u; v < 4; ++v)
            o[q++] = h / 6.0 * (2.0 * A2[u] * A2[v] + A2[u] * B2[v] + B2[u] * A2[v] + 2.0 * B2[u] * B2[v]);
      }
      for (int ph = 0; ph < 4; ++ph) {
        if (act && (m & 3) == ph) {
          int q = 0;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            Xty[m - 3 + u] += b[u];
#pragma unroll
            for (int v = u; v < 4; ++v) {
              XtX[(v - u) * nk + m - 3 + u] += a[q];
              Om[(v - u) * nk + m - 3 + u] += o[q];
              ++q;
            }
          }
        }
        __syncwarp();
      }
    }
    double r = 0.0;
    if (lane == 0) {
      double t1 = 0.0, t2 = 0.0;
      for (int k = 2; k <= nk - 4; ++k) { t1 += XtX[k]; t2 += Om[k]; }
      r = t1 / t2;
    }
    r = __shfl_sync(PREP_FULL, r, 0);
    // df(spar) = df_target on spar in [-1.5, 1.5], Illinois false position (MODEL_SPEC §11)
    double spar, dfv;
    int evals = 0;
    {
      double sa = -1.5, sb = 1.5;
      double fa = spl_fit(nk, XtX, Om, Xty, L, c, r * pow(256.0, 3.0 * sa - 1.0), lane) - df_target;
      ++evals;
      if (fa <= 0.0) {
        spar = sa; dfv = fa + df_target;
      } else {
        double fb = spl_fit(nk, XtX, Om, Xty, L, c, r * pow(256.0, 3.0 * sb - 1.0), lane) - df_target;
        ++evals;
        spar = sb; dfv = fb + df_target;
        if (fb < 0.0) {
          int side = 0;
          for (int it = 0; it < 100; ++it) {
            const double sc = (sa * fb - sb * fa) / (fb - fa);
            const double fc = spl_fit(nk, XtX, Om, Xty, L, c, r * pow(256.0, 3.0 * sc - 1.0), lane) - df_target;
            ++evals;
            spar = sc; dfv = fc + df_target;
            if (fabs(fc) <= 1e-10) break;
            if (fc < 0.0) { sb = sc; fb = fc; if (side == -1) fa *= 0.5; side = -1; }
            else { sa = sc; fa = fc; if (side == 1) fb *= 0.5; side = 1; }
          }
        }
      }
    }
    // fitted values
    for (int base = 3; base < nk; base += 32) {
      const int m = base + lane;
      if (m < nk) {
        const int i0 = knot_index(m - 3, N, nkn), i1 = (m == nk - 1) ? N : knot_index(m - 2, N, nkn);
        for (int i = i0; i < i1; ++i) {
          double B[4];
          bspl_val(T, m, (x[i] - x0) * ir, B);
          ysj[i] = B[0] * c[m - 3] + B[1] * c[m - 2] + B[2] * c[m - 1] + B[3] * c[m];
        }
      }
    }
    __syncwarp();
    // heteroscedastic noise fit on the residuals: Newton on the profile score in v = 1/a_2
    const double vmin = 1.0 / max_rate, xN = x[N - 1];
    double sx = 0.0;
    for (int i = lane; i < N; i += 32) sx += x[i];
    const double xbar = wsum(sx) / N;
    double v = 0.0;
    int at_bound = 0;
    for (int it = 0; it < 50; ++it) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int i = lane; i < N; i += 32) {
        const double res = y[i] - ysj[i];
        const double w = res * res * exp(2.0 * v * (x[i] - xN));
        s0 += w; s1 += w * x[i]; s2 += w * x[i] * x[i];
      }
      s0 = wsum(s0); s1 = wsum(s1); s2 = wsum(s2);
      const double m1 = s1 / s0, var = s2 / s0 - m1 * m1;
      double vn = v + (xbar - m1) / (2.0 * var);
      bool stop = false;
      if (vn < vmin) { vn = vmin; if (at_bound) stop = true; at_bound = 1; } else at_bound = 0;
      if (stop) break;
      const double dv = vn - v;
      v = vn;
      if (fabs(dv) * xN <= 1e-12) break;
    }
    double S0 = 0.0;
    for (int i = lane; i < N; i += 32) {
      const double res = y[i] - ysj[i];
      S0 += res * res * exp(2.0 * v * (x[i] - xN));
    }
    S0 = wsum(S0);
    const double a1 = sqrt(S0 / N) * exp(v * xN), a2 = 1.0 / v;
    for (int i = lane; i < N; i += 32) uyj[i] = a1 * exp(-x[i] / a2);
    if (lane == 0) {
      theta[2 * j] = a1; theta[2 * j + 1] = a2;
      if (info) { info[4 * j] = spar; info[4 * j + 1] = r * pow(256.0, 3.0 * spar - 1.0); info[4 * j + 2] = dfv; info[4 * j + 3] = (double)evals; }
      status[j] = (fabs(dfv - df_target) <= 1e-8) ? 0 : 1;  // 1: requested df not reachable on spar in [-1.5, 1.5]
    }
    __syncwarp();
  }
}

// k-th smallest (0-based) of n non-negative doubles: MSB-first radix select on the bit patterns, counts by the warp.
__device__ double warp_select(const double* a, int n, int k, int lane) {
  unsigned long long prefix = 0ull, mask = 0ull;
  for (int bit = 63; bit >= 0; --bit) {
    const unsigned long long m2 = mask | (1ull << bit);
    unsigned cnt = 0;
    for (int i = lane; i < n; i += 32) cnt += (((unsigned long long)__double_as_longlong(a[i]) & m2) == prefix);
    cnt = __reduce_add_sync(PREP_FULL, cnt);
    if ((unsigned)k >= cnt) { k -= (int)cnt; prefix |= (1ull << bit); }
    mask = m2;
  }
  return __longlong_as_double((long long)prefix);
}

// One warp per profile: the printBr gate (br against ci) and estimateExpPrior.  prior_type: -1 none, 0 mono, 1 abc.
__global__ void __launch_bounds__(128) gate_prior_kernel(const double* __restrict__ up, const PrepMeta* __restrict__ meta,
                                                         int n, int prior_type, const double* __restrict__ theta_map,
                                                         const double* __restrict__ hessian, double ru_theta,
                                                         const double* __restrict__ br, const double* __restrict__ ci,
                                                         double* __restrict__ absr, int* __restrict__ alert,
                                                         double* __restrict__ theta0, double* __restrict__ Sigma0,
                                                         double* __restrict__ ru_out) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n) return;
  if (br && lane == 0) alert[j] = !(br[j] >= ci[2 * j] && br[j] <= ci[2 * j + 1]);
  if (prior_type < 0) return;
  const PrepMeta M = meta[j];
  const int N = M.N;
  const double* x = up + M.off;
  const double* y = x + N;
  double th[3], A[9], C[9], sd[3], cor[9];
  for (int a = 0; a < 3; ++a) th[a] = theta_map[3 * (size_t)j + a];
  for (int k = 0; k < 9; ++k) A[k] = -hessian[9 * (size_t)j + k];
  const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
  C[0] = (A[4] * A[8] - A[5] * A[7]) / det; C[1] = (A[2] * A[7] - A[1] * A[8]) / det; C[2] = (A[1] * A[5] - A[2] * A[4]) / det;
  C[3] = (A[5] * A[6] - A[3] * A[8]) / det; C[4] = (A[0] * A[8] - A[2] * A[6]) / det; C[5] = (A[2] * A[3] - A[0] * A[5]) / det;
  C[6] = (A[3] * A[7] - A[4] * A[6]) / det; C[7] = (A[1] * A[6] - A[0] * A[7]) / det; C[8] = (A[0] * A[4] - A[1] * A[3]) / det;
  for (int a = 0; a < 3; ++a) sd[a] = sqrt(C[a * 3 + a]);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) cor[a * 3 + b] = a == b ? 1.0 : C[a * 3 + b] / (sd[a] * sd[b]);
  double ru = ru_theta;
  if (prior_type == 1) {
    double* ar = absr + M.out_off;
    const double cc = (double)M.dataType;
    double sbar = 0.0;
    for (int i = lane; i < N; i += 32) {
      const double t = cc * x[i] / th[2], e = exp(-t);
      const double J[3] = {1.0, e, th[1] * e * t / th[2]};
      double v = 0.0;
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) v += J[a] * th[a] * cor[a * 3 + b] * th[b] * J[b];
      sbar += sqrt(v);
      ar[i] = fabs(y[i] - (th[0] + th[1] * e));
    }
    sbar = wsum(sbar) / N;
    __syncwarp();
    const double h = 0.95 * (N - 1);
    const int lo = (int)floor(h);
    const double vlo = warp_select(ar, N, lo < N - 1 ? lo : N - 1, lane);
    double q95 = vlo;
    if (lo + 1 < N) {
      const double vhi = warp_select(ar, N, lo + 1, lane);
      q95 = vlo + (h - lo) * (vhi - vlo);
    }
    ru = q95 / (1.96 * sbar);
  }
  if (lane == 0) {
    for (int a = 0; a < 3; ++a) {
      theta0[3 * (size_t)j + a] = th[a];
      for (int b = 0; b < 3; ++b) Sigma0[9 * (size_t)j + a * 3 + b] = (ru * th[a]) * cor[a * 3 + b] * (ru * th[b]);
    }
    if (ru_out) ru_out[j] = ru;
  }
}

// ---------------------------------------------------------------- host side
// regularised lower incomplete gamma P(a, x): series below a + 1, Lentz continued fraction above
static double reg_gamma_p(double a, double x) {
  if (x <= 0.0) return 0.0;
  const double lg = std::lgamma(a), pre = std::exp(-x + a * std::log(x) - lg);
  if (x < a + 1.0) {
    double term = 1.0 / a, sum = term, ap = a;
    for (int it = 0; it < 20000; ++it) {
      ap += 1.0;
      term *= x / ap;
      sum += term;
      if (std::fabs(term) < std::fabs(sum) * 1e-17) break;
    }
    return sum * pre;
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, cc = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 20000; ++i) {
    const double an = -(double)i * ((double)i - a);
    b += 2.0;
    d = an * d + b; if (std::fabs(d) < tiny) d = tiny;
    cc = b + an / cc; if (std::fabs(cc) < tiny) cc = tiny;
    d = 1.0 / d;
    const double del = d * cc;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-17) break;
  }
  return 1.0 - pre * h;
}

static double chisq_quantile(double p, double ndf) {
  const double a = 0.5 * ndf;
  double lo = 0.0, hi = ndf + 40.0 * std::sqrt(2.0 * ndf) + 40.0, x = ndf;
  for (int it = 0; it < 200; ++it) {
    const double f = reg_gamma_p(a, 0.5 * x) - p;
    if (f > 0.0) hi = x; else lo = x;
    const double dens = 0.5 * std::exp(-0.5 * x + (a - 1.0) * std::log(0.5 * x) - std::lgamma(a));
    double xn = x - f / dens;
    if (!(xn > lo && xn < hi) || !std::isfinite(xn)) xn = 0.5 * (lo + hi);
    const bool done = std::fabs(xn - x) <= 1e-14 * std::fabs(x);
    x = xn;
    if (done) break;
  }
  return x;
}

struct PrepUpload {
  double* up = nullptr;  // pinned staging: x | y per problem
  std::vector<PrepMeta> meta;
  size_t total = 0;
  int nkmax = 0;
  double* d_up = nullptr;
  PrepMeta* d_meta = nullptr;
  ~PrepUpload() { device_cache_free(d_up); device_cache_free(d_meta); pinned_cache_free(up); }
};

static int prep_upload(const foct_problem* P, int n, PrepUpload& U) {
  if (!P || n < 1) return fail(FOCT_EINVAL, "empty batch or NULL argument");
  size_t tot = 0;
  for (int j = 0; j < n; ++j) {
    if (P[j].N < 4 || !P[j].x || !P[j].y) return fail(FOCT_EINVAL, "problem %d: N=%d or NULL data", j, P[j].N);
    if (P[j].dataType != 1 && P[j].dataType != 2) return fail(FOCT_EINVAL, "problem %d: dataType=%d not in {1,2}", j, P[j].dataType);
    tot += (size_t)P[j].N;
  }
  U.total = tot;
  if (pinned_cache_alloc((void**)&U.up, 2 * tot * sizeof(double)) != cudaSuccess) return fail(FOCT_ENOMEM, "pinned staging buffer of %zu bytes", 2 * tot * sizeof(double));
  U.meta.resize(n);
  size_t off = 0, out = 0;
  for (int j = 0; j < n; ++j) {
    const int N = P[j].N;
    std::memcpy(&U.up[off], P[j].x, N * sizeof(double));
    std::memcpy(&U.up[off + N], P[j].y, N * sizeof(double));
    PrepMeta& M = U.meta[j];
    M.off = off; M.out_off = out; M.N = N; M.nknots = nknots_smspl(N); M.dataType = P[j].dataType; M.pad = 0;
    U.nkmax = std::max(U.nkmax, M.nknots + 2);
    off += 2 * (size_t)N;
    out += (size_t)N;
  }
  if (device_cache_alloc((void**)&U.d_up, 2 * tot * sizeof(double)) != cudaSuccess ||
      device_cache_alloc((void**)&U.d_meta, n * sizeof(PrepMeta)) != cudaSuccess)
    return fail(FOCT_ENOMEM, "device allocation of %zu bytes failed", 2 * tot * sizeof(double));
  if (cudaMemcpy(U.d_up, U.up, 2 * tot * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(U.d_meta, U.meta.data(), n * sizeof(PrepMeta), cudaMemcpyHostToDevice) != cudaSuccess)
    return fail(FOCT_ECUDA, "upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

struct DevBuf {  // returns the block to the cache on scope exit
  void* p = nullptr;
  ~DevBuf() { device_cache_free(p); }
  template <class T> T* as() { return static_cast<T*>(p); }
  cudaError_t alloc(size_t bytes) { return device_cache_alloc(&p, bytes ? bytes : 8); }
};

static double ndf_of(int kind, const foct_problem& P, const foct_model_spec* spec) {
  if (spec && spec->br_ndf == 1) return (double)P.N;
  return (double)(P.N - 3 - (kind == FOCT_EXPGP ? P.Nn : 0));
}

static int gate_prior(int kind, const foct_problem* P, int n, const foct_model_spec* spec, int prior_type,
                      const double* theta_map, const double* hessian, double ru_theta, const double* br, double* ci_out,
                      int* alert, double* theta0, double* Sigma0, double* ru) {
  if (int rc = check_device()) return rc;
  PrepUpload U;
  if (int rc = prep_upload(P, n, U)) return rc;
  std::vector<double> ci;
  if (br) {
    ci.resize(2 * (size_t)n);
    std::map<double, std::pair<double, double>> cache;
    for (int j = 0; j < n; ++j) {
      const double ndf = ndf_of(kind, P[j], spec);
      if (!(ndf >= 1.0)) return fail(FOCT_EINVAL, "problem %d: no degrees of freedom left (N=%d)", j, P[j].N);
      auto it = cache.find(ndf);
      if (it == cache.end())
        it = cache.emplace(ndf, std::make_pair(chisq_quantile(0.025, ndf) / ndf, chisq_quantile(0.975, ndf) / ndf)).first;
      ci[2 * j] = it->second.first; ci[2 * j + 1] = it->second.second;
    }
    if (ci_out) std::memcpy(ci_out, ci.data(), ci.size() * sizeof(double));
  }
  DevBuf d_th, d_H, d_br, d_ci, d_absr, d_alert, d_t0, d_S0, d_ru;
#define CUP(call) if ((call) != cudaSuccess) return fail(FOCT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(cudaGetLastError()))
  if (prior_type >= 0) {
    CUP(d_th.alloc(3 * (size_t)n * 8)); CUP(d_H.alloc(9 * (size_t)n * 8));
    CUP(d_t0.alloc(3 * (size_t)n * 8)); CUP(d_S0.alloc(9 * (size_t)n * 8)); CUP(d_ru.alloc((size_t)n * 8));
    CUP(d_absr.alloc(U.total * 8));
    CUP(cudaMemcpy(d_th.p, theta_map, 3 * (size_t)n * 8, cudaMemcpyHostToDevice));
    CUP(cudaMemcpy(d_H.p, hessian, 9 * (size_t)n * 8, cudaMemcpyHostToDevice));
  }
  if (br) {
    CUP(d_br.alloc((size_t)n * 8)); CUP(d_ci.alloc(2 * (size_t)n * 8)); CUP(d_alert.alloc((size_t)n * sizeof(int)));
    CUP(cudaMemcpy(d_br.p, br, (size_t)n * 8, cudaMemcpyHostToDevice));
    CUP(cudaMemcpy(d_ci.p, ci.data(), 2 * (size_t)n * 8, cudaMemcpyHostToDevice));
  }
  gate_prior_kernel<<<(n + 3) / 4, 128>>>(U.d_up, U.d_meta, n, prior_type, d_th.as<double>(), d_H.as<double>(), ru_theta,
                                          br ? d_br.as<double>() : nullptr, d_ci.as<double>(), d_absr.as<double>(),
                                          d_alert.as<int>(), d_t0.as<double>(), d_S0.as<double>(), d_ru.as<double>());
  CUP(cudaGetLastError());
  CUP(cudaDeviceSynchronize());
  if (br) CUP(cudaMemcpy(alert, d_alert.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  if (prior_type >= 0) {
    CUP(cudaMemcpy(theta0, d_t0.p, 3 * (size_t)n * 8, cudaMemcpyDeviceToHost));
    CUP(cudaMemcpy(Sigma0, d_S0.p, 9 * (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (ru) CUP(cudaMemcpy(ru, d_ru.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  }
#undef CUP
  return 0;
}

}  // namespace foct

using namespace foct;

extern "C" int foct_birge_ci(double ndf, double* ci) {
  if (!ci || !(ndf >= 1.0)) return fail(FOCT_EINVAL, "ndf=%g must be >= 1 and ci non-NULL", ndf);
  ci[0] = chisq_quantile(0.025, ndf) / ndf;
  ci[1] = chisq_quantile(0.975, ndf) / ndf;
  return 0;
}

extern "C" int foct_estimate_noise(const foct_problem* P, int n, double df, double max_rate, double* uy, double* ySmooth,
                                   double* theta, double* info, int* status) {
  if (int rc = check_device()) return rc;
  if (!uy || !ySmooth || !theta) return fail(FOCT_EINVAL, "NULL output");
  if (!(df > 1.0)) return fail(FOCT_EINVAL, "df=%g must exceed 1", df);
  if (!(max_rate > 0.0)) max_rate = 1e4;
  Trace tr("foct_estimate_noise");
  PrepUpload U;
  if (int rc = prep_upload(P, n, U)) return rc;
  tr.mark("pack + upload");
  for (int j = 0; j < n; ++j)
    if (df > (double)(U.meta[j].nknots + 2)) return fail(FOCT_EINVAL, "problem %d: df=%g exceeds the %d spline coefficients", j, df, U.meta[j].nknots + 2);
  DevBuf d_uy, d_ys, d_th, d_info, d_st;
#define CUP(call) if ((call) != cudaSuccess) return fail(FOCT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(cudaGetLastError()))
  CUP(d_uy.alloc(U.total * 8)); CUP(d_ys.alloc(U.total * 8)); CUP(d_th.alloc(2 * (size_t)n * 8));
  CUP(d_info.alloc(4 * (size_t)n * 8)); CUP(d_st.alloc((size_t)n * sizeof(int)));
  const size_t smem = (16 * (size_t)U.nkmax + 4) * sizeof(double);
  if (smem > 200 * 1024) return fail(FOCT_EINVAL, "profile too long: %d spline coefficients do not fit in shared memory", U.nkmax);
  CUP(cudaFuncSetAttribute(noise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 0;
  CUP(cudaGetDevice(&dev));
  CUP(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(n, sms * 16);
  noise_kernel<<<grid, 32, smem>>>(U.d_up, U.d_meta, n, U.nkmax, df, max_rate, d_uy.as<double>(), d_ys.as<double>(),
                                   d_th.as<double>(), d_info.as<double>(), d_st.as<int>());
  tr.mark("alloc + launch");
  CUP(cudaGetLastError());
  CUP(cudaDeviceSynchronize());
  tr.mark("kernel");
  CUP(cudaMemcpy(uy, d_uy.p, U.total * 8, cudaMemcpyDeviceToHost));
  CUP(cudaMemcpy(ySmooth, d_ys.p, U.total * 8, cudaMemcpyDeviceToHost));
  CUP(cudaMemcpy(theta, d_th.p, 2 * (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (info) CUP(cudaMemcpy(info, d_info.p, 4 * (size_t)n * 8, cudaMemcpyDeviceToHost));
  if (status) CUP(cudaMemcpy(status, d_st.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  tr.mark("download");
#undef CUP
  return 0;
}

extern "C" int foct_print_br(int kind, const foct_problem* P, int n, const foct_model_spec* spec, const double* br,
                             double* ci, int* alert) {
  if (!br || !alert) return fail(FOCT_EINVAL, "NULL br or alert");
  if (kind != FOCT_EXPGP && kind != FOCT_MONOEXP) return fail(FOCT_EINVAL, "unknown model kind %d", kind);
  return gate_prior(kind, P, n, spec, -1, nullptr, nullptr, 0.0, br, ci, alert, nullptr, nullptr, nullptr);
}

extern "C" int foct_estimate_exp_prior(const foct_problem* P, int n, int prior_type, const double* theta_map,
                                       const double* hessian, double ru_theta, double* theta0, double* Sigma0, double* ru) {
  if (!theta_map || !hessian || !theta0 || !Sigma0) return fail(FOCT_EINVAL, "NULL argument");
  if (prior_type != FOCT_PRIOR_MONO && prior_type != FOCT_PRIOR_ABC) return fail(FOCT_EINVAL, "priorType=%d not in {mono, abc}", prior_type);
  return gate_prior(FOCT_MONOEXP, P, n, nullptr, prior_type, theta_map, hessian, ru_theta, nullptr, nullptr, nullptr, theta0,
                    Sigma0, ru);
}

extern "C" void foct_pipeline_cfg_default(foct_pipeline_cfg* c) {
  if (!c) return;
  c->smooth_df = 15.0; c->max_rate = 1e4; c->prior_type = FOCT_PRIOR_ABC; c->ru_theta = 0.05;
  c->Nn = 10; c->gridType = FOCT_GRID_INTERNAL; c->rho_scale = 0.0; c->lambda_rate = 0.1; c->gate = 1;
}

// The body of FitOCT.R's dataset loop (FitOCT.R:84-124) for a batch, every numerical step on the device.
extern "C" int foct_pipeline(const foct_problem* P, int n, const foct_pipeline_cfg* pc, const foct_model_spec* spec_gp,
                             const foct_sampler_cfg* cfg, foct_pipeline_out* out) {
  if (int rc = check_device()) return rc;
  if (!P || n < 1 || !pc || !cfg || !out) return fail(FOCT_EINVAL, "NULL argument or empty batch");
  if (!out->uy || !out->ySmooth || !out->noise_theta || !out->mono_theta || !out->mono_hessian || !out->mono_br ||
      !out->alert || !out->theta0 || !out->Sigma0 || !out->expgp_index)
    return fail(FOCT_EINVAL, "foct_pipeline_out has NULL buffers");
  if (pc->Nn < 1 || pc->Nn > FOCT_MAX_NN) return fail(FOCT_EINVAL, "Nn=%d outside 1..%d", pc->Nn, FOCT_MAX_NN);
  out->n_expgp = 0;
  // the small steps run on the first device the caller listed (fitExpGP itself shards over all of them)
  if (cfg->n_devices > 0 && cfg->devices && cudaSetDevice(cfg->devices[0]) != cudaSuccess)
    return fail(FOCT_ECUDA, "cudaSetDevice(%d) failed: %s", cfg->devices[0], cudaGetErrorString(cudaGetLastError()));
  // 1. estimateNoise (FitOCT.R:89)
  std::vector<int> st(n);
  if (int rc = foct_estimate_noise(P, n, pc->smooth_df, pc->max_rate, out->uy, out->ySmooth, out->noise_theta, nullptr, st.data())) return rc;
  for (int j = 0; j < n; ++j)
    if (st[j] == 3) return fail(FOCT_EINVAL, "problem %d: x must increase strictly", j);
  // 2. fitMonoExp MAP (FitOCT.R:95) with the estimated uy
  std::vector<foct_problem> Q(P, P + n);
  size_t off = 0;
  std::vector<double> ones;  // a profile whose noise fit failed (uy not finite / not positive) still rides along in the MAP batch
  for (int j = 0; j < n; ++j) {
    Q[j].uy = out->uy + off; Q[j].Nn = 0;
    bool ok = true;
    for (int i = 0; i < P[j].N && ok; ++i) ok = out->uy[off + i] > 0.0 && std::isfinite(out->uy[off + i]);
    if (!ok) {
      int maxN = 0;
      for (int k = 0; k < n; ++k) maxN = std::max(maxN, P[k].N);
      if (ones.empty()) ones.assign((size_t)maxN, 1.0);
      Q[j].uy = ones.data();  // flagged FOCT_PIPE_SKIPPED in step 4 (which looks at out->uy)
    }
    off += (size_t)P[j].N;
  }
  foct_model_spec sm;
  foct_model_spec_default(&sm, FOCT_MONOEXP);
  if (int rc = foct_monoexp_map(Q.data(), n, &sm, nullptr, out->mono_theta, out->mono_hessian, out->mono_br, out->mono_status)) return rc;
  // 3. printBr gate (FitOCT.R:98-100) and estimateExpPrior (FitOCT.R:103-107)
  if (int rc = gate_prior(FOCT_MONOEXP, Q.data(), n, &sm, pc->prior_type, out->mono_theta, out->mono_hessian, pc->ru_theta,
                          out->mono_br, out->br_ci, out->alert, out->theta0, out->Sigma0, out->ru))
    return rc;
  // 4. fitExpGP on the profiles the gate lets through (FitOCT.R:110-124).  Failures are handled PER PROFILE: a MonoExp
  //    fit that did not converge has a large Birge ratio, so it is exactly what the gate forwards, and its Hessian may be
  //    singular or indefinite (Sigma0 = NaN) — one such profile must not abort the whole directory.
  auto finite3 = [](const double* v, int k) { for (int i = 0; i < k; ++i) if (!std::isfinite(v[i])) return false; return true; };
  auto spd3 = [](const double* S) {  // Cholesky of a symmetric 3 x 3
    if (!(S[0] > 0.0)) return false;
    const double l10 = S[3] / std::sqrt(S[0]), l20 = S[6] / std::sqrt(S[0]);
    const double d1 = S[4] - l10 * l10;
    if (!(d1 > 0.0)) return false;
    const double l21 = (S[7] - l20 * l10) / std::sqrt(d1);
    return S[8] - l20 * l20 - l21 * l21 > 0.0;
  };
  std::vector<foct_problem> G;
  size_t uoff = 0;
  for (int j = 0; j < n; ++j) {
    const size_t u0 = uoff;
    uoff += (size_t)P[j].N;
    if (out->status) out->status[j] = FOCT_PIPE_GATED;
    if (pc->gate && !out->alert[j]) continue;
    double* th0 = out->theta0 + 3 * (size_t)j;
    double* S0 = out->Sigma0 + 9 * (size_t)j;
    const double* thm = out->mono_theta + 3 * (size_t)j;
    int stj = FOCT_PIPE_SAMPLED;
    bool uy_ok = true;
    for (int i = 0; i < P[j].N && uy_ok; ++i) uy_ok = out->uy[u0 + i] > 0.0 && std::isfinite(out->uy[u0 + i]);
    if (!uy_ok || !finite3(thm, 3) || !(thm[2] != 0.0)) {
      if (out->status) out->status[j] = FOCT_PIPE_SKIPPED;
      continue;
    }
    if (!finite3(th0, 3) || !finite3(S0, 9) || !spd3(S0)) {
      for (int k = 0; k < 9; ++k) S0[k] = 0.0;
      for (int k = 0; k < 3; ++k) {
        th0[k] = thm[k];
        const double sd = pc->ru_theta * std::fabs(thm[k]);
        S0[4 * k] = sd > 0.0 ? sd * sd : 1.0;
      }
      if (out->ru) out->ru[j] = pc->ru_theta;
      stj = FOCT_PIPE_PRIOR_REPAIRED;
    }
    if (out->status) out->status[j] = stj;
    foct_problem g = Q[j];
    g.Nn = pc->Nn; g.gridType = pc->gridType;
    g.rho = pc->rho_scale == 0.0 ? 1.0 / pc->Nn : pc->rho_scale;
    g.lambda_rate = pc->lambda_rate; g.prior_PD = 0;
    std::memcpy(g.theta0, th0, sizeof(g.theta0));
    std::memcpy(g.Sigma0, S0, sizeof(g.Sigma0));
    out->expgp_index[G.size()] = j;
    G.push_back(g);
  }
  out->n_expgp = (int)G.size();
  if (G.empty()) return 0;
  foct_model_spec sg;
  if (spec_gp) sg = *spec_gp; else foct_model_spec_default(&sg, FOCT_EXPGP);
  return foct_sample(FOCT_EXPGP, G.data(), (int)G.size(), &sg, cfg, &out->expgp);
}
