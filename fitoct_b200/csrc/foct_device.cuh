// foct_device.cuh — device-side building blocks: Philox, warp reductions, the fused fp64
// log-density + analytic-gradient sweep (MODEL_SPEC §3-5, SURVEY a-4/a-5).
//
// Layout contract: a lane group owns one chain - a warp (nuts_kernel), half a warp (nuts2*_kernel: two
// chains per warp) or a team of two warps that both hold the whole state (nuts_lat_kernel); lane d (< D) of the group owns
// component d of every D-vector (q, p, grad, rho, ...); group-uniform scalars are held redundantly by all lanes.  The profile
// lives in shared memory as a "blob" of npass = Npad/32 pass blocks; block j holds, for points 32j..32j+31,
// the rows cx | y | w | B_0 .. B_{NN-1}, 32 doubles each, so that one warp pass reads every operand at a
// compile-time offset from a single per-lane pointer (conflict-free LDS.64, no address arithmetic).
// Padded points carry w = 0 and contribute exactly nothing.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/fitoct_b200.h"

#define FOCT_FULL 0xffffffffu
#define FOCT_STACK_LEVELS 12  // subtree depth <= max_treedepth - 1 <= 12

namespace foct {
#ifdef FOCT_TIMING
__device__ unsigned long long g_tim[8];  // 0 grad total, 1 sweep loop, 2 reduce, 3 prologue, 4 leaf total, 5 merge, 6 leaves, 7 merges
#define FOCT_T(var) const long long var = clock64()
#define FOCT_TADD(i, a, b) do { if (lane == 0) atomicAdd(&g_tim[i], (unsigned long long)((b) - (a))); } while (0)
#else
#define FOCT_T(var)
#define FOCT_TADD(i, a, b)
#endif

// Per-profile constants the kernels read from global memory (written by the setup kernel).
struct DevProblem {
  int N, npass, prior_PD, Nn;
  double c;  // dataType
  double theta0[3];
  double Pinv[9];  // Sigma0^-1
  double lambda_rate;
  double sum_log_uy;
  double br_ndf;
  long long id;
  double xmin, xscale;  // xp = (x - xmin) * xscale
  double rho;
  int gridType, pad_;
};

// Batch-global model switches (MODEL_SPEC §8), passed by value as kernel parameters.
struct DevSpec {
  int ygp_prior, lambda_prior, theta_prior;
  double sigma_mean, sigma_sd, sigma_inv_sd;
};

// ---------------------------------------------------------------- RNG (MODEL_SPEC §7)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { SITE_MOM = 0, SITE_DIR = 1, SITE_MERGE = 2, SITE_INITEPS = 3, SITE_INIT = 4 };

struct Rng {
  uint32_t k0, k1;
  __device__ void seed(unsigned long long s, long long id, int chain) {
    unsigned long long stream = (unsigned long long)id * 64ull + (unsigned long long)chain;
    k0 = (uint32_t)s ^ ((uint32_t)(s >> 32) * 0x85EBCA6Bu) ^ (uint32_t)(stream >> 32);
    k1 = (uint32_t)stream;
  }
  __device__ __forceinline__ void block(uint32_t it, uint32_t kind, uint32_t a, uint32_t b, uint32_t lvl,
                                        uint32_t (&r)[4]) const {
    philox4x32_10(it, kind | (a << 8), b, lvl, k0, k1, r);
  }
};

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  unsigned long long a = (((unsigned long long)hi << 32) | lo) >> 11;
  return ((double)a + 0.5) * 0x1.0p-53;
}
// (one copy of log / cos / sqrt per kernel instead of one per call site when FOCT_NORMAL_NOINLINE: the sampling kernels are
// sensitive to their code size, see foct_nuts2.cuh)
#ifdef FOCT_NORMAL_NOINLINE
static __device__ __noinline__ double normal_from(const uint32_t (&r)[4]) {
#else
__device__ __forceinline__ double normal_from(const uint32_t (&r)[4]) {
#endif
  double u0 = u53(r[0], r[1]), u1 = u53(r[2], r[3]);
  return sqrt(-2.0 * log(u0)) * cos(6.283185307179586476925286766559 * u1);
}

// ---------------------------------------------------------------- blob layout
// Index (in doubles) of row r (0 cx, 1 y, 2 w, 3+k basis k) of point i inside a blob.
__host__ __device__ __forceinline__ size_t blob_index(int i, int r, int NN) {
  return ((size_t)(i >> 5) * (size_t)(3 + NN) + (size_t)r) * 32u + (size_t)(i & 31);
}

// ---------------------------------------------------------------- fp64 exp
// exp(x) = 2^m T[j] E(r): n = rint(32 x / ln 2) = 32 m + j by the magic-number trick, |r| <= ln2/64, E the degree-6
// Taylor polynomial (0.03 ulp truncation error), T = 2^(j/32) from a 256-byte table read through L1 with a
// lane-dependent index (at most two cache lines); 12 fp64 instructions and 9 constants.  (Round 1 went through the
// libdevice exp - ~50 issue slots with its per-call constant materialisation - and a table-free degree-11 Estrin
// variant with 18 instructions and 14 constants; the table version won by 3-4 % because fewer constants stay in
// registers across the sweep.)  Constants are literals so that ptxas may feed them as immediate / uniform operands.
#define FEXP_MAGIC 6755399441055744.0     /* 1.5 * 2^52 */
// Constants with a zero low word are fp64 IMMEDIATES (no register, two register-file reads per DFMA):
//  - 32/ln2 only selects n; rounded to 21 bits it moves n by at most one near a tie, i.e. |r| <= 1.02 ln2/64;
//  - Cody-Waite split of ln2/32 with a 21-bit head, so that n * head is exact and head is an immediate;
//  - 1/120 and 1/720 multiply r^5 <= 1.5e-10 and r^6 <= 1.7e-12: their 6e-8 relative rounding is far below 1 ulp.
#define FEXPT_L2E32 0x1.71547p+5
#define FEXPT_NHI (-0x1.62e43p-6)
#define FEXPT_C5 0x1.11111p-7
#define FEXPT_C6 0x1.6c16cp-10
static __device__ const double FEXP_TAB[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};
// The three constants that need all 53 bits cannot be immediates.  Round 1 wrote every constant as a literal and ptxas
// parked each in a vector register pair for the whole sweep (14 registers; a DFMA with three register-file operands
// issues every 3.0 cycles instead of 2.06, profiles/r1_microbench.txt).  From __constant__ memory ptxas can feed them
// through uniform registers (`DFMA R2, R4, UR8, R2`) where it has them to spare.
__constant__ double FEXP_K[4] = {
    0x1.05c610ca86c39p-34 /* head - ln2/32 */, 0x1.5555555555555p-3 /* 1/6 */, 0x1.5555555555555p-5 /* 1/24 */, 0.0};

// E(r) = exp(r) on |r| <= ln2/64: degree-6 Taylor polynomial in Horner form — six DFMAs with an immediate or uniform
// operand each (12.4 pipe cycles; round 1's Estrin form was seven operations, three of them with three register
// operands: 19 cycles.  The kernel is bound by fp64 issue, not by the length of this chain).
__device__ __forceinline__ double fexp_poly(double r) {
  double p = fma(r, FEXPT_C6, FEXPT_C5);
  p = fma(p, r, FEXP_K[2]);
  p = fma(p, r, FEXP_K[1]);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  return fma(p, r, 1.0);
}

// Assemble 2^m * p and saturate on the INTEGER pipe: |x| >= 700 gives 0 / +inf (exp(-700) ~ 1e-304 is below anything
// the model can resolve; a positive argument that large only arises from a negative decay length, i.e. a state that is
// non-finite anyway), NaN propagates through the polynomial.  hx = high word of x.  (Round 1 compared x in fp64:
// two DSETP on the fp64 pipe and four FSEL per point.)
__device__ __forceinline__ double fexp_scale(double p, int n, int hx) {
  int hi = __double2hiint(p) + ((n >> 5) << 20), lo = __double2loint(p);
  // |x| in [700, inf]  <=>  0x4085E000 <= (hx & 0x7fffffff) <= 0x7ff00000; NaN lies above
  const bool big = (unsigned)((hx & 0x7fffffff) - 0x4085E000) <= (unsigned)(0x7ff00000 - 0x4085E000);
  hi = big ? (hx < 0 ? 0 : 0x7ff00000) : hi;
  lo = big ? 0 : lo;
  return __hiloint2double(hi, lo);
}

// exp(x), branch-free (ptxas can interleave the points of the sweep).
__device__ __forceinline__ double fexp(double x) {
  const double t = fma(x, FEXPT_L2E32, FEXP_MAGIC);
  const double nd = t - FEXP_MAGIC;
  const int n = __double2loint(t);
  const double tj = __ldg(&FEXP_TAB[n & 31]);
  double r = fma(nd, FEXPT_NHI, x);
  r = fma(nd, FEXP_K[0], r);
  return fexp_scale(fexp_poly(r) * tj, n, __double2hiint(x));
}

// log(1 + e) for 0 <= e <= 1 (the log-sum-exp of the multinomial tree weights), absolute error ~2e-16: u = 1 + e in [1, 2]
// is split as u = (1 + f) / ic_j with ic_j = 1 / (1 + (j + 1/2) / 32) from a 32-entry table (|f| <= 1/64), so that
// log u = -log(ic_j) + log1p(f) with an 8-term series.  ~20 instructions; the libdevice log1p it replaces was 75
// executed instructions per leaf and 4 percent of the sampling kernel's stall samples (profiles/r2_ncu_pair_v1_regions.txt).
static __device__ const double2 FLOG_TAB[32] = {
    {0x1.f81f81f81f820p-1, 0x1.fc0a8b0fc03c4p-7}, {0x1.e9131abf0b767p-1, 0x1.77458f632dcffp-5},
    {0x1.dae6076b981dbp-1, 0x1.341d7961bd1d0p-4}, {0x1.cd85689039b0bp-1, 0x1.a926d3a4ad562p-4},
    {0x1.c0e070381c0e0p-1, 0x1.0d77e7cd08e5bp-3}, {0x1.b4e81b4e81b4fp-1, 0x1.44d2b6ccb7d1cp-3},
    {0x1.a98ef606a63bep-1, 0x1.7ab890210d907p-3}, {0x1.9ec8e951033d9p-1, 0x1.af3c94e80bff3p-3},
    {0x1.948b0fcd6e9e0p-1, 0x1.e27076e2af2e8p-3}, {0x1.8acb90f6bf3aap-1, 0x1.0a324e27390e2p-2},
    {0x1.8181818181818p-1, 0x1.22941fbcf7966p-2}, {0x1.78a4c8178a4c8p-1, 0x1.3a64c556945eap-2},
    {0x1.702e05c0b8170p-1, 0x1.51aad872df82ep-2}, {0x1.6816816816817p-1, 0x1.686c81e9b14adp-2},
    {0x1.6058160581606p-1, 0x1.7eaf83b82afc2p-2}, {0x1.58ed2308158edp-1, 0x1.947941c2116fbp-2},
    {0x1.51d07eae2f815p-1, 0x1.a9cec9a9a084ap-2}, {0x1.4afd6a052bf5bp-1, 0x1.beb4d9da71b7ap-2},
    {0x1.446f86562d9fbp-1, 0x1.d32fe7e00ebd5p-2}, {0x1.3e22cbce4a902p-1, 0x1.e744261d68789p-2},
    {0x1.3813813813814p-1, 0x1.faf588f78f31dp-2}, {0x1.323e34a2b10bfp-1, 0x1.0723e5c1cdf41p-1},
    {0x1.2c9fb4d812ca0p-1, 0x1.109f39e2d4c96p-1}, {0x1.27350b8812735p-1, 0x1.19ee6b467c96fp-1},
    {0x1.21fb78121fb78p-1, 0x1.23130d7bebf43p-1}, {0x1.1cf06ada2811dp-1, 0x1.2c0e9ed448e8cp-1},
    {0x1.1811811811812p-1, 0x1.34e289d9ce1d2p-1}, {0x1.135c81135c811p-1, 0x1.3d9026a7156fbp-1},
    {0x1.0ecf56be69c90p-1, 0x1.4618bc21c5ec2p-1}, {0x1.0a6810a6810a7p-1, 0x1.4e7d811b75bb0p-1},
    {0x1.0624dd2f1a9fcp-1, 0x1.56bf9d5b3f399p-1}, {0x1.0204081020408p-1, 0x1.5ee02a9241676p-1}};
__constant__ double FLOG_K[2] = {0x1.5555555555555p-2 /* 1/3 */, 0x1.999999999999ap-3 /* 1/5 */};
__device__ __forceinline__ double flog1p_unit(double e) {
  const double u = 1.0 + e;
  int j = (__double2hiint(u) - 0x3ff00000) >> 15;
  j = max(0, min(j, 31));  // u == 2 lands on the last entry; a NaN on any entry
  const double2 t = __ldg(&FLOG_TAB[j]);
  const double f = fma(u, t.x, -1.0);
  double p = fma(f, -0.125, 0x1.24924p-3 /* 1/7 */);   // the truncated 1/7 and 1/6 multiply f^7 and f^6 <= 1.5e-11
  p = fma(p, f, -0x1.55555p-3 /* 1/6 */);
  p = fma(p, f, FLOG_K[1]);
  p = fma(p, f, -0.25);
  p = fma(p, f, FLOG_K[0]);
  p = fma(p, f, -0.5);
  p = fma(p, f, 1.0);
  return fma(p, f, t.y);
}

// Branch-free reciprocal: MUFU.RCP64H seed (~20 bits), one cubic and one quadratic Newton step (the fast
// path of the CUDA division, without its exponent-range slow path: a == 0 or denormal yields NaN, i.e. a
// non-finite state, instead of +-inf).
__device__ __forceinline__ double frcp(double a) {
  double x0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(a));
  double e = fma(-a, x0, 1.0);
  e = fma(e, e, e);
  const double x1 = fma(x0, e, x0);
  const double e2 = fma(-a, x1, 1.0);
  return fma(x1, e2, x1);
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ double selp(bool c, double a, double b) {
  double r;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b), "r"((int)c));
  return r;
}
// Lane groups: a chain is owned by a group of W lanes — the whole warp (W = 32) or one half of it (W = 16, two chains
// per warp: foct_nuts2.cuh).  Shuffles carry the group width, so a butterfly or a broadcast never leaves its group;
// `mask` names the lanes that execute the call together (FOCT_FULL when the warp is converged, the half's own 16 bits
// inside per-chain control flow).
template <int W = 32>
__device__ __forceinline__ double warp_sum(double v, unsigned mask = FOCT_FULL) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, W);
  return v;
}
template <int W = 32>
__device__ __forceinline__ double bcast(double v, int src, unsigned mask = FOCT_FULL) { return __shfl_sync(mask, v, src, W); }
// Transposed ("reduce-scatter") reduction of KP per-lane accumulators over a group of W lanes: KP-1 + log2(W/KP)
// 64-bit shuffles instead of KP log2 W.  On return, the full sum of accumulator index a is returned to lane a of the
// group (a < KP); lanes >= KP get accumulator (lane mod KP).  `lane` is the index inside the group.
template <int KP, int W = 32>
__device__ __forceinline__ double warp_reduce_scatter(double (&v)[KP], int lane, unsigned mask = FOCT_FULL) {
  static_assert((KP == 8 || KP == 16 || KP == 32) && KP <= W, "KP");
  int width = KP;
#pragma unroll
  for (int off = W / 2; off >= 1; off >>= 1) {
    if (width > 1) {
      const int half = width / 2;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const double keep = selp(upper, v[j + half], v[j]);
        const double send = selp(upper, v[j], v[j + half]);
        v[j] = keep + __shfl_xor_sync(mask, send, off, W);
      }
      width = half;
    } else {
      v[0] += __shfl_xor_sync(mask, v[0], off, W);
    }
  }
  // lane L now holds accumulator index L >> log2(W / KP).  Route index a to lane a.
  constexpr int SH = KP == W ? 0 : (2 * KP == W ? 1 : 2);
  if (SH == 0) return v[0];
  return __shfl_sync(mask, v[0], (lane << SH) & (W - 1), W);
}

// ---------------------------------------------------------------- model sweep
// Result of one log-density/gradient evaluation, in the warp layout.
struct Eval {
  double g;     // lane d: d lp / d q_d   (0 for lanes >= D)
  double lp;    // uniform
  double chi2;  // uniform: sum ((y-m)/uy)^2  (NaN when prior_PD)
};

template <int NN>
struct Dims {
  static constexpr bool GP = NN > 0;
  static constexpr int D = GP ? NN + 5 : 3;
  static constexpr int P_OUT = GP ? NN + 7 : 5;
  static constexpr int KP = D <= 8 ? 8 : (D <= 16 ? 16 : 32);
};

// FOCT_EXPTAB_SMEM: the sweeps read the 2^(j/32) table of exp() from a per-CTA copy in shared memory (every kernel that
// evaluates the model calls fill_exptab() first) instead of through L1 from global memory: two integer instructions less
// per point, a shorter and steadier latency on the critical path of the point, and the L1 request path — the co-bottleneck
// of the sampling kernels — is left to the basis rows.  Measured on nuts2w_kernel: +3.1 % gradients/s at full waves, 1000
// profiles 4.21 s -> 4.08 s, bit-identical draws (profiles/r2_kernel_experiments.txt).
#ifndef FOCT_EXPTAB_SMEM
#define FOCT_EXPTAB_SMEM 1
#endif
__shared__ double s_exptab[32];
__device__ __forceinline__ void fill_exptab() {
#if FOCT_EXPTAB_SMEM
  if (threadIdx.x < 32) s_exptab[threadIdx.x] = FEXP_TAB[threadIdx.x];
  __syncthreads();  // (blocks are a whole number of warps; one-warp blocks included)
#endif
}
__device__ __forceinline__ double exptab(int n) {
#if FOCT_EXPTAB_SMEM
  return s_exptab[n & 31];
#else
  return __ldg(&FEXP_TAB[n & 31]);
#endif
}

// U data points of the sweep processed in LOCKSTEP, statement by statement (ptxas keeps source order inside a
// basic block, so writing the U independent dependency chains interleaved is what actually puts U chains in
// flight; two inlined copies of a one-point body were scheduled back to back — profiles/r1_ncu_nuts_v2).
// Lane-private loads at compile-time offsets from pp, ~52 fp64 instructions per point, no branches.
// Raw sums go to acc[] (theta2 / theta3 factors are applied once, after the reduction).
// W = 32: point u of the iteration sits one pass block further (pp + u * block); W = 16 (half-warp lane groups, pp
// carries the lane index inside the group): the points are the two 16-point halves of consecutive 32-point blocks.
// GB: the basis rows are not in the staged block (which then holds cx | y | w only) but read through L1 from `pg`, a
// blob in global memory that every profile of the batch shares (same depth grid => same basis).  (A copy
// of the basis rows in shared memory, shared by three two-warp sub-CTAs of a 192-thread CTA, was measured 11 % slower at
// the same twelve warps per SM: profiles/r2_kernel_experiments.txt.)
#ifndef FOCT_PAIR_LDS128
#define FOCT_PAIR_LDS128 1
#endif
template <int NN, int MOD, int KP, int ZI, int U, int W = 32, int GB = 0>
__device__ __forceinline__ void sweep_points(const double* __restrict__ pp0, double th1, double th2, double th3, double r3,
                                             const double (&yg)[NN > 0 ? NN : 1], double (&acc)[KP],
                                             const double* __restrict__ pg = nullptr) {
  static_assert(W == 32 || (W == 16 && U % 2 == 0), "half-warp groups take the two halves of a block together");
  constexpr int SROWS = GB == 2 ? 2 : (GB ? 3 : 3 + NN);   // rows of a staged block (GB = 2: y | w only, c x comes from the shared blob too)
  constexpr int YR = GB == 2 ? 0 : 1;                      // row of y inside a staged block
  const double* __restrict__ pp = pp0;
#define FOCT_PT(u) (W == 32 ? (u) * SROWS * 32 : ((u) >> 1) * SROWS * 32 + ((u) & 1) * 16)
#define FOCT_PTG(u) (W == 32 ? (u) * (3 + NN) * 32 : ((u) >> 1) * (3 + NN) * 32 + ((u) & 1) * 16)
  double b[U][NN > 0 ? NN : 1];
  double dl0[U], dl1[U], s[U], cx[U], y[U], ws[U], r[U], t[U];
#pragma unroll
  for (int u = 0; u < U; ++u) { dl0[u] = 1.0; dl1[u] = 0.0; }  // s = 1 + dL: the 1 rides in the first partial sum
#if FOCT_PAIR_LDS128
  if constexpr (W == 16) {
    // Half-warp groups: lane l takes the NEIGHBOURING points 2l, 2l+1 of a 32-point block (pp carries 2l), so one
    // 128-bit load per row serves both points in flight.  An LDS.64 costs ~2.6 issue cycles next to the fp64 stream,
    // an LDS.128 ~1.6 for twice the data (scripts/micro/issue_cost.cu): the loads were a quarter of the sweep's issue time.
#pragma unroll
    for (int k = 0; k < NN; ++k) {
#pragma unroll
      for (int v = 0; v < U / 2; ++v) {
        const double2 bb = GB ? __ldg(reinterpret_cast<const double2*>(pg + v * (3 + NN) * 32 + (3 + k) * 32))
                              : *reinterpret_cast<const double2*>(pp + v * SROWS * 32 + (3 + k) * 32);
        b[2 * v][k] = bb.x; b[2 * v + 1][k] = bb.y;
        if (k & 1) { dl1[2 * v] = fma(bb.x, yg[k], dl1[2 * v]); dl1[2 * v + 1] = fma(bb.y, yg[k], dl1[2 * v + 1]); }
        else { dl0[2 * v] = fma(bb.x, yg[k], dl0[2 * v]); dl0[2 * v + 1] = fma(bb.y, yg[k], dl0[2 * v + 1]); }
      }
    }
#pragma unroll
    for (int v = 0; v < U / 2; ++v) {
      const double2 c2 = GB == 2 ? __ldg(reinterpret_cast<const double2*>(pg + v * (3 + NN) * 32))
                                 : *reinterpret_cast<const double2*>(pp + v * SROWS * 32);
      const double2 y2 = *reinterpret_cast<const double2*>(pp + v * SROWS * 32 + YR * 32);
      const double2 w2 = *reinterpret_cast<const double2*>(pp + v * SROWS * 32 + (YR + 1) * 32);
      cx[2 * v] = c2.x; cx[2 * v + 1] = c2.y; y[2 * v] = y2.x; y[2 * v + 1] = y2.y; ws[2 * v] = w2.x; ws[2 * v + 1] = w2.y;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) s[u] = dl0[u] + dl1[u];
  } else
#endif
  {
#pragma unroll
    for (int k = 0; k < NN; ++k) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        b[u][k] = GB ? __ldg(pg + FOCT_PTG(u) + (3 + k) * 32) : pp[FOCT_PT(u) + (3 + k) * 32];
        if (k & 1) dl1[u] = fma(b[u][k], yg[k], dl1[u]); else dl0[u] = fma(b[u][k], yg[k], dl0[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s[u] = dl0[u] + dl1[u];
      cx[u] = GB == 2 ? __ldg(pg + FOCT_PTG(u)) : pp[FOCT_PT(u)];
      y[u] = pp[FOCT_PT(u) + YR * 32];
      ws[u] = pp[FOCT_PT(u) + (YR + 1) * 32];  // 1/uy; the 1/sigma^2 common to every sum is applied once after the reduction
    }
  }
  // reciprocal of the local decay length (length modulation with a GP); otherwise r3 is hoisted
  if (MOD == 0 && NN > 0) {
    double a[U], x0[U], e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { a[u] = th3 * s[u]; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0[u]) : "d"(a[u])); }
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = fma(-a[u], x0[u], 1.0);
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = fma(e[u], e[u], e[u]);
#ifdef FOCT_RCP_FULL
#pragma unroll
    for (int u = 0; u < U; ++u) x0[u] = fma(x0[u], e[u], x0[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = fma(-a[u], x0[u], 1.0);
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = fma(x0[u], e[u], x0[u]);
#else
    // the 20-bit MUFU seed plus ONE cubic step already leaves a relative error of ~2^-60 + rounding (<= 2 ulp):
    // the second Newton step of the CUDA division only buys the last bit, at two more dependent DFMAs per point
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = fma(x0[u], e[u], x0[u]);
#endif
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = r3;
  }
#pragma unroll
  for (int u = 0; u < U; ++u) t[u] = cx[u] * r[u];
  // e = exp(-t): table-driven (see FEXP_TAB), staged across the U points
  double tt[U], nd[U], rr[U], pe[U], tj[U], e[U];
  int n[U];
#pragma unroll
  for (int u = 0; u < U; ++u) tt[u] = fma(-t[u], FEXPT_L2E32, FEXP_MAGIC);
#pragma unroll
  for (int u = 0; u < U; ++u) { nd[u] = tt[u] - FEXP_MAGIC; n[u] = __double2loint(tt[u]); tj[u] = exptab(n[u]); }
#pragma unroll
  for (int u = 0; u < U; ++u) rr[u] = fma(nd[u], FEXPT_NHI, -t[u]);
#pragma unroll
  for (int u = 0; u < U; ++u) rr[u] = fma(nd[u], FEXP_K[0], rr[u]);
#pragma unroll
  for (int u = 0; u < U; ++u) pe[u] = fma(rr[u], FEXPT_C6, FEXPT_C5);
#pragma unroll
  for (int u = 0; u < U; ++u) pe[u] = fma(pe[u], rr[u], FEXP_K[2]);
#pragma unroll
  for (int u = 0; u < U; ++u) pe[u] = fma(pe[u], rr[u], FEXP_K[1]);
#pragma unroll
  for (int u = 0; u < U; ++u) pe[u] = fma(pe[u], rr[u], 0.5);
#pragma unroll
  for (int u = 0; u < U; ++u) pe[u] = fma(pe[u], rr[u], 1.0);
#pragma unroll
  for (int u = 0; u < U; ++u) pe[u] = fma(pe[u], rr[u], 1.0);
#pragma unroll
  for (int u = 0; u < U; ++u) e[u] = fexp_scale(pe[u] * tj[u], n[u], __double2hiint(t[u]) ^ 0x80000000);
  if (MOD == 0) {
    double m[U], z[U], gi[U], ge[U], qq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) m[u] = fma(th2, e[u], th1);
#pragma unroll
    for (int u = 0; u < U; ++u) z[u] = (y[u] - m[u]) * ws[u];
#pragma unroll
    for (int u = 0; u < U; ++u) gi[u] = z[u] * ws[u];
#pragma unroll
    for (int u = 0; u < U; ++u) ge[u] = gi[u] * e[u];
    // qq s = ge t / theta3 (r s = 1/theta3): the theta3 sum needs neither qq nor s, and its 1/theta3 moves behind the reduction
    double gt[U];
#pragma unroll
    for (int u = 0; u < U; ++u) gt[u] = ge[u] * t[u];
#pragma unroll
    for (int u = 0; u < U; ++u) qq[u] = gt[u] * r[u];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      acc[ZI] = fma(z[u], z[u], acc[ZI]);
      acc[0] += gi[u];
      acc[1] += ge[u];
      acc[2] += gt[u];
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < NN; ++k) acc[3 + k] = fma(qq[u], b[u][k], acc[3 + k]);
  } else {
    double es[U], m[U], z[U], gi[U], ge[U], ges[U];
#pragma unroll
    for (int u = 0; u < U; ++u) es[u] = e[u] * s[u];
#pragma unroll
    for (int u = 0; u < U; ++u) m[u] = fma(th2, es[u], th1);
#pragma unroll
    for (int u = 0; u < U; ++u) z[u] = (y[u] - m[u]) * ws[u];
#pragma unroll
    for (int u = 0; u < U; ++u) gi[u] = z[u] * ws[u];
#pragma unroll
    for (int u = 0; u < U; ++u) { ge[u] = gi[u] * e[u]; ges[u] = gi[u] * es[u]; }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      acc[ZI] = fma(z[u], z[u], acc[ZI]);
      acc[0] += gi[u];
      acc[1] += ges[u];
      acc[2] = fma(ges[u], t[u], acc[2]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < NN; ++k) acc[3 + k] = fma(ge[u], b[u][k], acc[3 + k]);
  }
}

#undef FOCT_PT
#undef FOCT_PTG

// Software-pipelined form of the half-warp / shared-basis sweep (W = 16, GB = 1, two points per lane, length modulation):
// the basis rows of a block arrive in `bb` (loaded one iteration AHEAD), and each row of the NEXT block is requested right
// after its last use in the transposed accumulation, so the L1 latency of the ten 128-bit basis loads is covered by the
// tail of this iteration instead of being exposed at the head of the next one (ncu: "long scoreboard" was the second
// largest stall of nuts2w_kernel, 1.0-1.7 cycles per issued instruction).  Same arithmetic in the same order as
// sweep_points<..., U = 2, W = 16, GB = 1>: bit-identical results.
#ifndef FOCT_PREFETCH
#define FOCT_PREFETCH 1
#endif
#ifndef FOCT_CX_AHEAD_MAXNN
#define FOCT_CX_AHEAD_MAXNN 9
#endif
template <int NN, int KP, int ZI, bool CXG>
__device__ __forceinline__ void sweep_points_pf(const double* __restrict__ pp, const double* __restrict__ pgc,
                                                const double* __restrict__ pgn, double th1, double th2, double th3,
                                                const double (&yg)[NN], double (&acc)[KP], double2 (&bb)[NN], double2& cxn) {
  constexpr int U = 2;
  // CXG: the depth row c x is the same for every profile of the batch, so it is read from the shared blob like the basis
  // rows, and only y | w are staged per warp: 8 KB instead of 12 KB of shared memory per warp.  With few control points the
  // iteration is too short to cover the load between its head and the use of c x (Nn = 5 lost 13 %): there the pair is
  // requested one iteration ahead like the basis rows (`cxn`, 4 more registers live across the iteration); from
  // FOCT_CX_AHEAD_MAXNN + 1 control points on the dot products of the A phase cover it and the registers are not there.
  constexpr bool AHEAD = CXG && NN <= FOCT_CX_AHEAD_MAXNN;
  double2 c2g;
  if constexpr (AHEAD) c2g = cxn;
  else if constexpr (CXG) c2g = __ldg(reinterpret_cast<const double2*>(pgc));
  double dl0[U], dl1[U], cx[U], y[U], ws[U], r[U], t[U], a[U], x0[U], e0[U];
#pragma unroll
  for (int u = 0; u < U; ++u) { dl0[u] = 1.0; dl1[u] = 0.0; }  // s = 1 + dL: the 1 rides in the first partial sum
#pragma unroll
  for (int k = 0; k < NN; ++k) {
    if (k & 1) { dl1[0] = fma(bb[k].x, yg[k], dl1[0]); dl1[1] = fma(bb[k].y, yg[k], dl1[1]); }
    else { dl0[0] = fma(bb[k].x, yg[k], dl0[0]); dl0[1] = fma(bb[k].y, yg[k], dl0[1]); }
  }
  {
    constexpr int YR = CXG ? 0 : 1;
    double2 c2;
    if constexpr (CXG) c2 = c2g; else c2 = *reinterpret_cast<const double2*>(pp);
    const double2 y2 = *reinterpret_cast<const double2*>(pp + YR * 32);
    const double2 w2 = *reinterpret_cast<const double2*>(pp + (YR + 1) * 32);
    cx[0] = c2.x; cx[1] = c2.y; y[0] = y2.x; y[1] = y2.y; ws[0] = w2.x; ws[1] = w2.y;
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    a[u] = th3 * (dl0[u] + dl1[u]);
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0[u]) : "d"(a[u]));
  }
#pragma unroll
  for (int u = 0; u < U; ++u) e0[u] = fma(-a[u], x0[u], 1.0);
#pragma unroll
  for (int u = 0; u < U; ++u) e0[u] = fma(e0[u], e0[u], e0[u]);
#pragma unroll
  for (int u = 0; u < U; ++u) r[u] = fma(x0[u], e0[u], x0[u]);
#pragma unroll
  for (int u = 0; u < U; ++u) t[u] = cx[u] * r[u];
  double tt[U], nd[U], rr[U], pe[U], tj[U], e[U];
  int n[U];
#pragma unroll
  for (int u = 0; u < U; ++u) tt[u] = fma(-t[u], FEXPT_L2E32, FEXP_MAGIC);
#pragma unroll
  for (int u = 0; u < U; ++u) { nd[u] = tt[u] - FEXP_MAGIC; n[u] = __double2loint(tt[u]); tj[u] = exptab(n[u]); }
#pragma unroll
  for (int u = 0; u < U; ++u) rr[u] = fma(nd[u], FEXPT_NHI, -t[u]);
#pragma unroll
  for (int u = 0; u < U; ++u) rr[u] = fma(nd[u], FEXP_K[0], rr[u]);
#pragma unroll
  for (int u = 0; u < U; ++u) pe[u] = fma(rr[u], FEXPT_C6, FEXPT_C5);
#pragma unroll
  for (int u = 0; u < U; ++u) pe[u] = fma(pe[u], rr[u], FEXP_K[2]);
#pragma unroll
  for (int u = 0; u < U; ++u) pe[u] = fma(pe[u], rr[u], FEXP_K[1]);
#pragma unroll
  for (int u = 0; u < U; ++u) pe[u] = fma(pe[u], rr[u], 0.5);
#pragma unroll
  for (int u = 0; u < U; ++u) pe[u] = fma(pe[u], rr[u], 1.0);
#pragma unroll
  for (int u = 0; u < U; ++u) pe[u] = fma(pe[u], rr[u], 1.0);
#pragma unroll
  for (int u = 0; u < U; ++u) e[u] = fexp_scale(pe[u] * tj[u], n[u], __double2hiint(t[u]) ^ 0x80000000);
  double m[U], z[U], gi[U], ge[U], gt[U], qq[U];
#pragma unroll
  for (int u = 0; u < U; ++u) m[u] = fma(th2, e[u], th1);
#pragma unroll
  for (int u = 0; u < U; ++u) z[u] = (y[u] - m[u]) * ws[u];
#pragma unroll
  for (int u = 0; u < U; ++u) gi[u] = z[u] * ws[u];
#pragma unroll
  for (int u = 0; u < U; ++u) ge[u] = gi[u] * e[u];
#pragma unroll
  for (int u = 0; u < U; ++u) gt[u] = ge[u] * t[u];
#pragma unroll
  for (int u = 0; u < U; ++u) qq[u] = gt[u] * r[u];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    acc[ZI] = fma(z[u], z[u], acc[ZI]);
    acc[0] += gi[u];
    acc[1] += ge[u];
    acc[2] += gt[u];
  }
  // transposed accumulation, two rows at a time so that the two dependent updates of an accumulator are not adjacent;
  // a row's registers are refilled for the next block as soon as both of its points are consumed
#pragma unroll
  for (int k = 0; k < NN; k += 2) {
    acc[3 + k] = fma(qq[0], bb[k].x, acc[3 + k]);
    if (k + 1 < NN) acc[4 + k] = fma(qq[0], bb[k + 1].x, acc[4 + k]);
    acc[3 + k] = fma(qq[1], bb[k].y, acc[3 + k]);
    if (k + 1 < NN) acc[4 + k] = fma(qq[1], bb[k + 1].y, acc[4 + k]);
    bb[k] = __ldg(reinterpret_cast<const double2*>(pgn + (3 + k) * 32));
    if (k + 1 < NN) bb[k + 1] = __ldg(reinterpret_cast<const double2*>(pgn + (4 + k) * 32));
  }
  if constexpr (AHEAD) cxn = __ldg(reinterpret_cast<const double2*>(pgn));
}

// Two warps per chain (TEAM = 2, the latency kernel of small batches: nuts_lat_kernel).  Both warps of a team execute the
// whole chain redundantly — same state, same Philox sites, same decisions — but each sweeps only its half of the profile's
// blocks; the reduce-scattered partial sums are exchanged through shared memory and added in a fixed order (member 0's +
// member 1's), so both members continue with bit-identical values and never diverge.  One named barrier per gradient
// (the exchange buffer is double-buffered by parity, which makes the write-after-read of the next gradient safe).
struct TeamCtx {
  int member;    // 0 / 1; member 0 is the one that writes results
  int bar_id;    // named barrier of the team (1 + chain slot of the CTA), 64 threads
  int parity;
  double* xch;   // shared memory: [2 parities][2 members][32]
  int* flag;     // shared memory: the team's cancel decision
};
__device__ __forceinline__ void team_sync(const TeamCtx* tc) { asm volatile("bar.sync %0, 64;" ::"r"(tc->bar_id) : "memory"); }

// Fused sweep over the staged profile.  qd = this lane's component of q.  Everything a lane needs from
// the other lanes is fetched with shuffles up front; the per-point loop touches only shared memory and
// processes two passes (64 points per warp) per iteration so that two independent dependency chains are
// in flight per warp (the fp64 pipe was latency-, not throughput-bound with one: profiles/r1_*v1*).
// W = 32: one chain per warp.  W = 16: one chain per half-warp — `lane` is then the index inside the half, both halves
// must call together (full-mask shuffles of width 16) and walk the same profile, so every LDS is a 16-word broadcast.
template <int NN, int MOD, int W = 32, int GB = 0, int TEAM = 1>
__device__ __forceinline__ Eval warp_logp_grad(const double* __restrict__ blob, const DevProblem& P, const DevSpec& S,
                                               double qd, int lane, const double* __restrict__ gbasis = nullptr,
                                               TeamCtx* tc = nullptr) {
  static_assert(W == 32 || Dims<NN>::D <= 16, "a half-warp holds at most 16 components");
  static_assert(TEAM == 1 || (TEAM == 2 && W == 32 && GB == 0), "teams of two whole warps on a staged blob");
  FOCT_T(t_g0);
  using DM = Dims<NN>;
  constexpr int D = DM::D;
  constexpr int KP = DM::KP;
#ifndef FOCT_UNROLL
#define FOCT_UNROLL 2
#endif
#ifndef FOCT_UNROLL_MAXNN
#define FOCT_UNROLL_MAXNN 12
#endif
#ifndef FOCT_UNROLL_LAT
#define FOCT_UNROLL_LAT 4
#endif
  // above 12 control points the second basis row spills (measured: Nn=15 is 9 % faster with 1); the latency kernel has
  // 255 registers per thread and one warp per scheduler: four points in flight, two above 12 control points
  constexpr int UNROLL = TEAM == 2 ? (NN <= FOCT_UNROLL_MAXNN ? FOCT_UNROLL_LAT : (NN <= 20 ? 2 : 1))
                                   : (NN <= FOCT_UNROLL_MAXNN ? FOCT_UNROLL : 1);
  const double th1 = bcast<W>(qd, 0), th2 = bcast<W>(qd, 1), th3 = bcast<W>(qd, 2);
  double yg[NN > 0 ? NN : 1];
#pragma unroll
  for (int k = 0; k < NN; ++k) yg[k] = bcast<W>(qd, 3 + k);
  const double qlam = DM::GP ? bcast<W>(qd, 3 + NN) : 0.0;
  const double qsig = DM::GP ? bcast<W>(qd, 4 + NN) : 0.0;
  const double lam = DM::GP ? fexp(qlam) : 1.0;
  const double sig = DM::GP ? fexp(qsig) : 1.0;
  const double isig = DM::GP ? frcp(sig) : 1.0;
  const double isig2 = isig * isig;

  // priors + Jacobians (O(D), evaluated redundantly by every lane; each lane keeps its own component).  Done BEFORE
  // the sweep, while q is fresh in registers: only two doubles stay live across the loop (computed after it, this
  // block cost 1460 cycles per leaf in register reloads - scripts/timing_probe.py).
  double pr_lp = 0.0, pr_g = 0.0;
  if (S.theta_prior == 0) {
    const double d0 = th1 - P.theta0[0], d1 = th2 - P.theta0[1], d2 = th3 - P.theta0[2];
    const double v0 = P.Pinv[0] * d0 + P.Pinv[1] * d1 + P.Pinv[2] * d2;
    const double v1 = P.Pinv[3] * d0 + P.Pinv[4] * d1 + P.Pinv[5] * d2;
    const double v2 = P.Pinv[6] * d0 + P.Pinv[7] * d1 + P.Pinv[8] * d2;
    pr_lp += -0.5 * (d0 * v0 + d1 * v1 + d2 * v2);
    if (lane == 0) pr_g -= v0;
    if (lane == 1) pr_g -= v1;
    if (lane == 2) pr_g -= v2;
  }
  if (DM::GP) {
    double sy = 0.0;
    if (S.ygp_prior == 0) {
      const double il1 = frcp(lam);
      const double il2 = il1 * il1;
#pragma unroll
      for (int k = 0; k < NN; ++k) sy = fma(yg[k], yg[k], sy);
      pr_lp += -(double)NN * qlam - 0.5 * sy * il2;
      if (lane >= 3 && lane < 3 + NN) pr_g -= qd * il2;
      if (lane == 3 + NN) pr_g += sy * il2 - (double)NN;
    } else {
      const double il = frcp(lam);
#pragma unroll
      for (int k = 0; k < NN; ++k) sy += fabs(yg[k]);
      pr_lp += -(double)NN * qlam - sy * il;
      if (lane >= 3 && lane < 3 + NN) pr_g -= (qd > 0.0 ? 1.0 : (qd < 0.0 ? -1.0 : 0.0)) * il;
      if (lane == 3 + NN) pr_g += sy * il - (double)NN;
    }
    const double rl = P.lambda_rate * lam;
    if (S.lambda_prior == 0) {
      pr_lp += qlam - rl;
      if (lane == 3 + NN) pr_g += 1.0 - rl;
    } else {
      pr_lp += -rl;
      if (lane == 3 + NN) pr_g += -rl;
    }
    if (S.sigma_sd > 0.0) {
      const double u = (sig - S.sigma_mean) * S.sigma_inv_sd;
      pr_lp += -0.5 * u * u;
      if (lane == 4 + NN) pr_g += -sig * u * S.sigma_inv_sd;
    }
    pr_lp += qlam + qsig;
    if (lane == 3 + NN || lane == 4 + NN) pr_g += 1.0;
  }

  double acc[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) acc[k] = 0.0;
  // acc[0..2] = d/d theta (unscaled), acc[3..3+NN) = d/d yGP (unscaled), acc[ZI] = sum z^2
  constexpr int ZI = DM::GP ? 4 + NN : KP - 1;
  Eval out;
  double zz = 0.0;
  if (!P.prior_PD) {
    const double r3 = frcp(th3);  // (a full division drags its exponent-range slow path into every leaf)
    constexpr int ROWS = GB == 2 ? 2 : (GB ? 3 : 3 + NN);     // rows of a staged block
    constexpr int GROWS = 3 + NN;             // rows of a block of the shared (global) blob
    // (half-warp groups with 128-bit loads: lane l owns the neighbouring points 2l, 2l+1 of each block)
    const int lane_off = (W == 16 && FOCT_PAIR_LDS128) ? 2 * lane : lane;
    const double* pp = blob + lane_off;
    const double* pg = GB ? gbasis + lane_off : nullptr;
    FOCT_T(t_l0);
    FOCT_TADD(3, t_g0, t_l0);
    // (A profile of 481 points is 15 blocks and ONE point, so the 16th iteration runs 1/32 full.  Evaluating such tail
    // points by the lane group as a whole — control point k in lane 3 + k, the scalar chain repeated by every lane — was
    // measured 2.5 % SLOWER than the padded iteration: it is one more serial dependency chain of the same length, and
    // the kernel is bound by latency, not by issue slots.  profiles/r2_kernel_experiments.txt)
    int npass = P.npass;
    int pass = 0;
    if constexpr (TEAM == 2) {  // this member's half of the blocks
      const int half = (npass + 1) >> 1;
      if (tc->member) { pass = half; pp += (size_t)half * ROWS * 32; } else { npass = half; }
    }
    if constexpr (W == 16) {
      // 32-point blocks: the two 16-point halves of a block are two points in flight per lane; FOCT_UNROLL16 = 4 takes
      // two blocks per iteration (two warps per scheduler is all the staged profiles leave room for: the instruction-
      // level parallelism has to come from the loop body)
#ifndef FOCT_UNROLL16
#define FOCT_UNROLL16 4
#endif
#ifndef FOCT_UNROLL16_GB
#define FOCT_UNROLL16_GB 2
#endif
      constexpr int U16 = NN <= 11 ? (GB ? FOCT_UNROLL16_GB : FOCT_UNROLL16) : 2;
      if constexpr (FOCT_PREFETCH && GB && MOD == 0 && NN > 0 && NN <= 11 && U16 == 2 && FOCT_PAIR_LDS128) {
        double2 bb[NN];
#pragma unroll
        for (int k = 0; k < NN; ++k) bb[k] = __ldg(reinterpret_cast<const double2*>(pg + (3 + k) * 32));
        double2 cxn = make_double2(0.0, 0.0);
        if constexpr (GB == 2 && NN <= FOCT_CX_AHEAD_MAXNN) cxn = __ldg(reinterpret_cast<const double2*>(pg));
#pragma unroll 1
        for (; pass < npass; ++pass, pp += ROWS * 32) {
          const double* pgc = pg;
          if (pass + 1 < npass) pg += GROWS * 32;  // the last iteration re-requests its own block: harmless, stays in bounds
          sweep_points_pf<NN, KP, ZI, GB == 2>(pp, pgc, pg, th1, th2, th3, yg, acc, bb, cxn);
        }
      } else {
        if (U16 > 2) {
#pragma unroll 1
          for (; pass + U16 / 2 <= npass; pass += U16 / 2, pp += (U16 / 2) * ROWS * 32, pg += (U16 / 2) * GROWS * 32)
            sweep_points<NN, MOD, KP, ZI, U16, W, GB>(pp, th1, th2, th3, r3, yg, acc, pg);
        }
#pragma unroll 1
        for (; pass < npass; ++pass, pp += ROWS * 32, pg += GROWS * 32)
          sweep_points<NN, MOD, KP, ZI, 2, W, GB>(pp, th1, th2, th3, r3, yg, acc, pg);
      }
    } else {
      if (UNROLL >= 2) {
#pragma unroll 1
        for (; pass + UNROLL <= npass; pass += UNROLL, pp += UNROLL * ROWS * 32, pg += UNROLL * GROWS * 32)
          sweep_points<NN, MOD, KP, ZI, UNROLL, W, GB>(pp, th1, th2, th3, r3, yg, acc, pg);
      }
#pragma unroll 1
      for (; pass < npass; ++pass, pp += ROWS * 32, pg += GROWS * 32)
        sweep_points<NN, MOD, KP, ZI, 1, W, GB>(pp, th1, th2, th3, r3, yg, acc, pg);
    }
    FOCT_T(t_l1);
    FOCT_TADD(1, t_l0, t_l1);
    double red = warp_reduce_scatter<KP, W>(acc, lane);
    if constexpr (TEAM == 2) {
      double* buf = tc->xch + tc->parity * 64;
      buf[tc->member * 32 + lane] = red;
      team_sync(tc);
      red = buf[lane] + buf[32 + lane];
      tc->parity ^= 1;
    }
    red *= isig2;  // every sum carries the factor 1/sigma^2
    zz = bcast<W>(red, ZI);
    FOCT_T(t_l2);
    FOCT_TADD(2, t_l1, t_l2);
    // scale the raw sums into gradient components
    double scale = 1.0;
    if (MOD == 0) {
      if (lane == 2) scale = th2 * r3;
      if (lane >= 3 && lane < 3 + NN) scale = th2 * th3;
    } else {
      if (lane == 2) scale = th2 * r3;
      if (lane >= 3 && lane < 3 + NN) scale = th2;
    }
    out.g = red * scale;
    if (DM::GP && lane == 4 + NN) out.g = zz - (double)P.N;
    if (lane >= D) out.g = 0.0;
    out.lp = -0.5 * zz - (double)P.N * qsig - P.sum_log_uy;
    out.chi2 = zz * sig * sig;  // = sum ((y-m)/uy)^2
  } else {
    out.g = 0.0;
    out.lp = 0.0;
    out.chi2 = CUDART_NAN;
  }

  out.lp += pr_lp;
  out.g += pr_g;
  FOCT_T(t_g1);
  FOCT_TADD(0, t_g0, t_g1);
  return out;
}

// Stage one profile blob (contiguous in global memory) into shared memory with a TMA bulk copy
// (cp.async.bulk, SASS UBLKCP) completed on an mbarrier.  Called by all threads of the CTA.
__device__ __forceinline__ void stage_blob_tma(double* smem_dst, const double* gsrc, uint32_t bytes, uint64_t* mbar,
                                               uint32_t& phase, bool leader) {
  const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(mbar);
  if (leader) {
    const uint32_t dst_s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    // (the block may have been read through the generic proxy by the previous work item)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(bytes) : "memory");
    // chunks of <= 64 KB keep every copy well inside any per-instruction limit
    uint32_t off = 0;
    while (off < bytes) {
      uint32_t n = bytes - off;
      if (n > 65536u) n = 65536u;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s + off),
          "l"((const char*)gsrc + off), "r"(n), "r"(mbar_s)
          : "memory");
      off += n;
    }
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mbar_s), "r"(phase)
        : "memory");
  }
  phase ^= 1u;
}

// Stage only the first `rows` rows (cx | y | w) of every 32-point block of a blob: one bulk copy per block, all
// completing on the same mbarrier.  Used when the basis rows are shared by the whole batch and read through L1.
__device__ __forceinline__ void stage_rows_tma(double* smem_dst, const double* gsrc, int nblocks, int src_rows, int rows,
                                               uint64_t* mbar, uint32_t& phase, bool leader) {
  const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(mbar);
  if (leader) {
    const uint32_t dst_s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const uint32_t blk = (uint32_t)rows * 32u * (uint32_t)sizeof(double);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(blk * (uint32_t)nblocks) : "memory");
    for (int j = 0; j < nblocks; ++j)
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s + (uint32_t)j * blk),
          "l"((const char*)(gsrc + (size_t)j * src_rows * 32)), "r"(blk), "r"(mbar_s)
          : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mbar_s), "r"(phase)
        : "memory");
  }
  phase ^= 1u;
}

__device__ __forceinline__ void mbar_init(uint64_t* mbar) {
  if (threadIdx.x == 0) {
    const uint32_t mbar_s = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

}  // namespace foct
