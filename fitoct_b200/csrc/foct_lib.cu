// foct_lib.cu — host side of the C ABI (include/fitoct_b200.h): argument checking, device-resident plans,
// the setup kernel that builds the per-profile shared-memory blobs (incl. the GP basis, MODEL_SPEC §1),
// multi-GPU sharding with one host thread per device (no collective: SURVEY §8e), and small helper kernels.
//
// There is no CPU compute path in this file: every compute entry point needs a CUDA device.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "foct_launch.h"
#include "foct_summary.cuh"

namespace foct {
#ifdef FOCT_VARIANT_NNS
// scripts/build_variant.sh: a small build for kernel A/B runs that carries only Nn = 0 and Nn = 10
FOCT_DECL_INST(0) FOCT_DECL_INST(10)
static const InstEntry* inst_for(int NN) { return NN == 0 ? foct_inst_0() : (NN == 10 ? foct_inst_10() : nullptr); }
#else
FOCT_DECL_INST(0)
FOCT_DECL_INST(1)  FOCT_DECL_INST(2)  FOCT_DECL_INST(3)  FOCT_DECL_INST(4)  FOCT_DECL_INST(5)
FOCT_DECL_INST(6)  FOCT_DECL_INST(7)  FOCT_DECL_INST(8)  FOCT_DECL_INST(9)  FOCT_DECL_INST(10)
FOCT_DECL_INST(11) FOCT_DECL_INST(12) FOCT_DECL_INST(13) FOCT_DECL_INST(14) FOCT_DECL_INST(15)
FOCT_DECL_INST(16) FOCT_DECL_INST(17) FOCT_DECL_INST(18) FOCT_DECL_INST(19) FOCT_DECL_INST(20)
FOCT_DECL_INST(21) FOCT_DECL_INST(22) FOCT_DECL_INST(23) FOCT_DECL_INST(24) FOCT_DECL_INST(25)

static const InstEntry* inst_for(int NN) {
  typedef const InstEntry* (*getter)();
  static const getter table[FOCT_MAX_NN + 1] = {
      foct_inst_0,  foct_inst_1,  foct_inst_2,  foct_inst_3,  foct_inst_4,  foct_inst_5,  foct_inst_6,
      foct_inst_7,  foct_inst_8,  foct_inst_9,  foct_inst_10, foct_inst_11, foct_inst_12, foct_inst_13,
      foct_inst_14, foct_inst_15, foct_inst_16, foct_inst_17, foct_inst_18, foct_inst_19, foct_inst_20,
      foct_inst_21, foct_inst_22, foct_inst_23, foct_inst_24, foct_inst_25};
  if (NN < 0 || NN > FOCT_MAX_NN) return nullptr;
  return table[NN]();
}
#endif

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(FOCT_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)


// ------------------------------------------------------------------ device-memory cache
// cudaMalloc / cudaFree cost milliseconds (cudaFree also synchronises the device) and an R session calls the fit once per
// profile, so freed blocks are kept and handed back to later requests of a similar size on the same device.  At most
// FOCT_POOL_MB (default 8192) MB stay cached; foct_release_cache() returns everything to the driver.
namespace {
struct DevicePool {
  std::mutex mu;
  std::unordered_map<void*, std::pair<int, size_t>> live;
  std::multimap<std::pair<int, size_t>, void*> idle;
  size_t idle_bytes = 0;
  size_t cap() const {
    static const size_t c = [] {
      const char* e = std::getenv("FOCT_POOL_MB");
      return (size_t)((e ? std::atof(e) : 8192.0) * 1024.0 * 1024.0);
    }();
    return c;
  }
  void release(int dev) {  // dev < 0: every device.  Called with mu held.
    for (auto it = idle.begin(); it != idle.end();) {
      if (dev < 0 || it->first.first == dev) { idle_bytes -= it->first.second; cudaFree(it->second); it = idle.erase(it); }
      else ++it;
    }
  }
};
DevicePool g_pool;
}  // namespace

template <class T>
static cudaError_t pool_malloc(T** out, size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  bytes = (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
  std::lock_guard<std::mutex> lk(g_pool.mu);
  auto it = g_pool.idle.lower_bound({dev, bytes});
  if (it != g_pool.idle.end() && it->first.first == dev && it->first.second <= bytes + bytes / 4 + (1u << 20)) {
    void* p = it->second;
    g_pool.live[p] = it->first;
    g_pool.idle_bytes -= it->first.second;
    g_pool.idle.erase(it);
    *out = static_cast<T*>(p);
    return cudaSuccess;
  }
  void* p = nullptr;
  e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {  // give the cached blocks back and retry once
    cudaGetLastError();
    g_pool.release(dev);
    e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return e;
  }
  g_pool.live[p] = {dev, bytes};
  *out = static_cast<T*>(p);
  return cudaSuccess;
}

static void pool_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pool.mu);
  auto it = g_pool.live.find(p);
  if (it == g_pool.live.end()) { cudaFree(p); return; }
  const std::pair<int, size_t> key = it->second;
  g_pool.live.erase(it);
  if (g_pool.idle_bytes + key.second <= g_pool.cap()) {
    g_pool.idle.emplace(key, p);
    g_pool.idle_bytes += key.second;
  } else {
    cudaFree(p);
  }
}


// Pinned staging buffers are cached the same way: cudaMallocHost / cudaFreeHost map and unmap the pages in every CUDA
// context of the process and were measured at 0.4 - 0.7 s per call of foct_sample inside a process that also runs torch.
namespace {
struct PinnedPool {
  std::mutex mu;
  std::unordered_map<void*, size_t> live;
  std::multimap<size_t, void*> idle;
  size_t idle_bytes = 0;
  static constexpr size_t kCap = (size_t)512 << 20;
};
PinnedPool g_pinned;
}  // namespace

template <class T>
static cudaError_t pinned_malloc(T** out, size_t bytes) {
  bytes = (std::max<size_t>(bytes, 1) + 4095) & ~(size_t)4095;
  {
    std::lock_guard<std::mutex> lk(g_pinned.mu);
    auto it = g_pinned.idle.lower_bound(bytes);
    if (it != g_pinned.idle.end() && it->first <= 2 * bytes + (1u << 20)) {
      void* p = it->second;
      g_pinned.live[p] = it->first;
      g_pinned.idle_bytes -= it->first;
      g_pinned.idle.erase(it);
      *out = static_cast<T*>(p);
      return cudaSuccess;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMallocHost(&p, bytes);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(g_pinned.mu);
  g_pinned.live[p] = bytes;
  *out = static_cast<T*>(p);
  return cudaSuccess;
}

static void pinned_free(void* p) {
  if (!p) return;
  std::unique_lock<std::mutex> lk(g_pinned.mu);
  auto it = g_pinned.live.find(p);
  if (it == g_pinned.live.end()) { lk.unlock(); cudaFreeHost(p); return; }
  const size_t bytes = it->second;
  g_pinned.live.erase(it);
  if (g_pinned.idle_bytes + bytes <= PinnedPool::kCap) {
    g_pinned.idle.emplace(bytes, p);
    g_pinned.idle_bytes += bytes;
    return;
  }
  lk.unlock();
  cudaFreeHost(p);
}

// shared with foct_prep.cu (declared in foct_launch.h)
cudaError_t device_cache_alloc(void** p, size_t bytes) { return pool_malloc(p, bytes); }
void device_cache_free(void* p) { pool_free(p); }
cudaError_t pinned_cache_alloc(void** p, size_t bytes) { return pinned_malloc(p, bytes); }
void pinned_cache_free(void* p) { pinned_free(p); }

extern "C" void foct_release_cache(void) {
  {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    g_pool.release(-1);
  }
  std::lock_guard<std::mutex> lk(g_pinned.mu);
  for (auto& kv : g_pinned.idle) cudaFreeHost(kv.second);
  g_pinned.idle.clear();
  g_pinned.idle_bytes = 0;
}

// ------------------------------------------------------------------ setup kernel
// Host-side description of one profile inside the concatenated upload buffer.
struct HostMeta {
  size_t off;  // offset (doubles) of x in the upload buffer; y at off+N, uy at off+2N
  int N, dataType, gridType, prior_PD;
  double rho, lambda_rate, theta0[3], Pinv[9];
  long long id;
};

__device__ __forceinline__ double gp_kern(double a, double b, double rho, int kernel) {
  const double d = a - b;
  const double den = kernel == 0 ? 2.0 * rho * rho : rho * rho;
  return exp(-(d * d) / den);
}

// One CTA per profile: depth normalisation, Cholesky of Kgg (thread 0; Nn <= 25), per-point triangular
// solves B_i = Kgg^-1 k(xp_i, xGP), and the blob cx | y | w | B[k][i] written straight into global memory
// in the layout the sampling kernel bulk-copies into shared memory.
__global__ void __launch_bounds__(128) setup_kernel(const double* __restrict__ up, const HostMeta* __restrict__ meta,
                                                    int n_problems, int NN, int npad, size_t blob_stride,
                                                    double* __restrict__ blobs, DevProblem* __restrict__ probs,
                                                    int kernel, double jitter, int br_mode, int* __restrict__ status) {
  __shared__ double sL[FOCT_MAX_NN * FOCT_MAX_NN];
  __shared__ double sxg[FOCT_MAX_NN];
  __shared__ double red[3][128];
  __shared__ int s_bad;
  for (int j = blockIdx.x; j < n_problems; j += gridDim.x) {
    const HostMeta M = meta[j];
    const int N = M.N;
    const double* x = up + M.off;
    const double* y = x + N;
    const double* uy = y + N;
    double mn = CUDART_INF, mx = -CUDART_INF, slu = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      mn = fmin(mn, x[i]);
      mx = fmax(mx, x[i]);
      slu += log(uy[i]);
    }
    red[0][threadIdx.x] = mn; red[1][threadIdx.x] = mx; red[2][threadIdx.x] = slu;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) {
        red[0][threadIdx.x] = fmin(red[0][threadIdx.x], red[0][threadIdx.x + s]);
        red[1][threadIdx.x] = fmax(red[1][threadIdx.x], red[1][threadIdx.x + s]);
        red[2][threadIdx.x] += red[2][threadIdx.x + s];
      }
      __syncthreads();
    }
    const double xmin = red[0][0], xmax = red[1][0];
    if (threadIdx.x == 0) {
      // control grid (server.R:626-635) and Cholesky of Kgg + jitter I
      const double dx = 1.0 / (NN + 1);
      const double lo = M.gridType == FOCT_GRID_INTERNAL ? 0.5 * dx : 0.0;
      const double hi = M.gridType == FOCT_GRID_INTERNAL ? 1.0 - 0.5 * dx : 1.0;
      for (int k = 0; k < NN; ++k) sxg[k] = NN == 1 ? lo : lo + (hi - lo) * (double)k / (double)(NN - 1);
      for (int i = 0; i < NN; ++i)
        for (int jj = 0; jj <= i; ++jj) {
          double s = gp_kern(sxg[i], sxg[jj], M.rho, kernel) + (i == jj ? jitter : 0.0);
          for (int k = 0; k < jj; ++k) s -= sL[i * NN + k] * sL[jj * NN + k];
          if (i == jj) {
            if (!(s > 0.0)) s_bad = 1;
            sL[i * NN + i] = sqrt(s);
          } else {
            sL[i * NN + jj] = s / sL[jj * NN + jj];
          }
        }
      DevProblem P;
      P.N = N; P.npass = (N + 31) / 32; P.prior_PD = M.prior_PD; P.Nn = NN;
      P.c = (double)M.dataType;
      for (int k = 0; k < 3; ++k) P.theta0[k] = M.theta0[k];
      for (int k = 0; k < 9; ++k) P.Pinv[k] = M.Pinv[k];
      P.lambda_rate = M.lambda_rate;
      P.sum_log_uy = red[2][0];
      P.br_ndf = br_mode == 1 ? (double)N : (double)(N - 3 - NN);
      P.id = M.id;
      P.xmin = xmin; P.xscale = 1.0 / (xmax - xmin);
      P.rho = M.rho; P.gridType = M.gridType; P.pad_ = 0;
      probs[j] = P;
      if (s_bad || !(xmax > xmin)) atomicExch(status, j + 1);
    }
    __syncthreads();
    double* blob = blobs + (size_t)j * blob_stride;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
      const bool in = i < N;
      blob[blob_index(i, 0, NN)] = in ? (double)M.dataType * x[i] : 0.0;
      blob[blob_index(i, 1, NN)] = in ? y[i] : 0.0;
      blob[blob_index(i, 2, NN)] = in ? 1.0 / uy[i] : 0.0;
      if (NN > 0) {
        double v[FOCT_MAX_NN];
        if (in) {
          const double xp = (x[i] - xmin) / (xmax - xmin);
          for (int k = 0; k < NN; ++k) v[k] = gp_kern(xp, sxg[k], M.rho, kernel);
          for (int k = 0; k < NN; ++k) {
            double s = v[k];
            for (int jj = 0; jj < k; ++jj) s -= sL[k * NN + jj] * v[jj];
            v[k] = s / sL[k * NN + k];
          }
          for (int k = NN - 1; k >= 0; --k) {
            double s = v[k];
            for (int jj = k + 1; jj < NN; ++jj) s -= sL[jj * NN + k] * v[jj];
            v[k] = s / sL[k * NN + k];
          }
        }
        for (int k = 0; k < NN; ++k) blob[blob_index(i, 3 + k, NN)] = in ? v[k] : 0.0;
      }
    }
    __syncthreads();
  }
}

// Generated quantities (SURVEY a-6) for selected draws of one profile.
__global__ void predict_kernel(const double* __restrict__ blob, int npad, int N, int NN, int mod, int P_out,
                               const double* __restrict__ draws, int n_draws, double* m_out, double* resid,
                               double* dL) {
  const size_t tot = (size_t)n_draws * N;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(t / N), i = (int)(t % N);
    const double* r = draws + (size_t)j * P_out;
    double dl = 0.0;
    for (int k = 0; k < NN; ++k) dl = fma(blob[blob_index(i, 3 + k, NN)], r[3 + k], dl);
    const double cx = blob[blob_index(i, 0, NN)];
    const double m = mod == 0 ? r[0] + r[1] * exp(-cx / (r[2] * (1.0 + dl))) : r[0] + r[1] * exp(-cx / r[2]) * (1.0 + dl);
    if (m_out) m_out[t] = m;
    if (resid) resid[t] = blob[blob_index(i, 1, NN)] - m;
    if (dL) dL[t] = dl;
  }
}

// DFMA-chain microbenchmark: 8 independent chains per thread, the roofline denominator.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
    v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
}

// ---------------------------------------------------------------- MonoExp MAP (SURVEY a-12)
// One warp per profile: damped Newton (Levenberg-Marquardt on the exact Hessian) on -lp of the
// mono-exponential with sigma == 1; the same iteration the oracle runs, lanes striding over the points.
__device__ __forceinline__ bool dev_inv3(const double* S, double* Pi) {
  const double a = S[0], b = S[1], c = S[2], d = S[3], e = S[4], f = S[5], g = S[6], h = S[7], i = S[8];
  const double A = e * i - f * h, Bc = -(d * i - f * g), C = d * h - e * g;
  const double det = a * A + b * Bc + c * C;
  if (!(fabs(det) > 0.0)) return false;
  const double id = 1.0 / det;
  Pi[0] = A * id;  Pi[1] = -(b * i - c * h) * id; Pi[2] = (b * f - c * e) * id;
  Pi[3] = Bc * id; Pi[4] = (a * i - c * g) * id;  Pi[5] = -(a * f - c * d) * id;
  Pi[6] = C * id;  Pi[7] = -(a * h - b * g) * id; Pi[8] = (a * e - b * d) * id;
  return true;
}

// f = -lp, g[3], H[9] (of -lp), chi2; all lanes receive all values.
__device__ void mono_nlp(const double* __restrict__ blob, int NN, const DevProblem& P, int theta_prior,
                         const double* th, double& f, double* g, double* H, double& chi2, int lane) {
  double v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 0.0;
  for (int i = lane; i < P.N; i += 32) {
    const double w = blob[blob_index(i, 2, NN)], w2 = w * w;
    const double t = blob[blob_index(i, 0, NN)] / th[2];
    const double e = exp(-t);
    const double m = th[0] + th[1] * e;
    const double r = blob[blob_index(i, 1, NN)] - m;
    const double J0 = 1.0, J1 = e, J2 = th[1] * e * t / th[2];
    const double m23 = e * t / th[2];
    const double m33 = th[1] * e * (t * t - 2.0 * t) / (th[2] * th[2]);
    v[0] += 0.5 * r * r * w2;
    v[1] -= r * w2 * J0; v[2] -= r * w2 * J1; v[3] -= r * w2 * J2;
    v[4] += w2 * J0 * J0; v[5] += w2 * J0 * J1; v[6] += w2 * J0 * J2;
    v[7] += w2 * J1 * J1; v[8] += w2 * J1 * J2 - w2 * r * m23;
    v[9] += w2 * J2 * J2 - w2 * r * m33;
    v[10] += r * r * w2;
  }
  const double red = warp_reduce_scatter<16>(v, lane);
  double a[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) a[k] = __shfl_sync(FOCT_FULL, red, k);
  f = a[0]; g[0] = a[1]; g[1] = a[2]; g[2] = a[3];
  H[0] = a[4]; H[1] = a[5]; H[2] = a[6]; H[3] = a[5]; H[4] = a[7]; H[5] = a[8]; H[6] = a[6]; H[7] = a[8]; H[8] = a[9];
  chi2 = a[10];
  if (theta_prior == 0) {
    const double d[3] = {th[0] - P.theta0[0], th[1] - P.theta0[1], th[2] - P.theta0[2]};
    for (int r = 0; r < 3; ++r) {
      double vv = 0.0;
      for (int c = 0; c < 3; ++c) { vv += P.Pinv[r * 3 + c] * d[c]; H[r * 3 + c] += P.Pinv[r * 3 + c]; }
      f += 0.5 * d[r] * vv; g[r] += vv;
    }
  }
}

__global__ void __launch_bounds__(128) map_kernel(const double* __restrict__ blobs, size_t blob_stride, int NN,
                                                  const DevProblem* __restrict__ probs, int n, int theta_prior,
                                                  const double* __restrict__ init, double* theta_out,
                                                  double* hess_out, double* br_out, int* status_out) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n) return;
  const DevProblem P = probs[j];
  const double* blob = blobs + (size_t)j * blob_stride;
  double th[3], g[3], H[9], c2, f;
  if (init) {
    th[0] = init[j * 3]; th[1] = init[j * 3 + 1]; th[2] = init[j * 3 + 2];
  } else {
    // log-linear start: theta1 just below min(y), regress log(y - theta1) on x
    double ymin = CUDART_INF, ymax = -CUDART_INF;
    for (int i = lane; i < P.N; i += 32) { ymin = fmin(ymin, blob[blob_index(i, 1, NN)]); ymax = fmax(ymax, blob[blob_index(i, 1, NN)]); }
    for (int o = 16; o > 0; o >>= 1) {
      ymin = fmin(ymin, __shfl_xor_sync(FOCT_FULL, ymin, o));
      ymax = fmax(ymax, __shfl_xor_sync(FOCT_FULL, ymax, o));
    }
    const double th1 = ymin - 0.05 * (ymax - ymin);
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = lane; i < P.N; i += 32) {
      const double yy = blob[blob_index(i, 1, NN)] - th1;
      if (yy > 0.0) {
        const double x = blob[blob_index(i, 0, NN)] / P.c, ly = log(yy);
        v[0] += x; v[1] += ly; v[2] += x * x; v[3] += x * ly; v[4] += 1.0;
      }
    }
    const double red = warp_reduce_scatter<8>(v, lane);
    const double sx = __shfl_sync(FOCT_FULL, red, 0), sy = __shfl_sync(FOCT_FULL, red, 1);
    const double sxx = __shfl_sync(FOCT_FULL, red, 2), sxy = __shfl_sync(FOCT_FULL, red, 3);
    const double nn = __shfl_sync(FOCT_FULL, red, 4);
    const double slope = (nn * sxy - sx * sy) / (nn * sxx - sx * sx);
    const double icpt = (sy - slope * sx) / nn;
    th[0] = th1; th[1] = exp(icpt);
    th[2] = slope < 0.0 ? -P.c / slope : (blob[blob_index(P.N - 1, 0, NN)] - blob[0]) / P.c;
  }
  mono_nlp(blob, NN, P, theta_prior, th, f, g, H, c2, lane);
  double mu = 1e-3;
  int st = 1;
  for (int it = 0; it < 200; ++it) {
    double A[9], Ai[9], step[3], tn[3], gn[3], Hn[9], c2n, fn;
    for (int k = 0; k < 9; ++k) A[k] = H[k];
    for (int a = 0; a < 3; ++a) A[a * 3 + a] += mu * fabs(H[a * 3 + a]) + 1e-300;
    if (!dev_inv3(A, Ai)) { mu *= 10.0; continue; }
    for (int a = 0; a < 3; ++a) step[a] = -(Ai[a * 3] * g[0] + Ai[a * 3 + 1] * g[1] + Ai[a * 3 + 2] * g[2]);
    for (int a = 0; a < 3; ++a) tn[a] = th[a] + step[a];
    mono_nlp(blob, NN, P, theta_prior, tn, fn, gn, Hn, c2n, lane);
    if (isfinite(fn) && fn <= f) {
      double rel = 0.0;
      for (int a = 0; a < 3; ++a) rel = fmax(rel, fabs(step[a]) / (fabs(th[a]) + 1e-300));
      for (int a = 0; a < 3; ++a) { th[a] = tn[a]; g[a] = gn[a]; }
      for (int k = 0; k < 9; ++k) H[k] = Hn[k];
      c2 = c2n; f = fn;
      mu = mu * 0.2 > 1e-12 ? mu * 0.2 : 1e-12;
      if (rel < 1e-12) { st = 0; break; }
    } else {
      mu *= 5.0;
      if (mu > 1e12) break;
    }
  }
  if (lane == 0) {
    for (int a = 0; a < 3; ++a) theta_out[j * 3 + a] = th[a];
    if (hess_out) for (int k = 0; k < 9; ++k) hess_out[(size_t)j * 9 + k] = -H[k];
    if (br_out) br_out[j] = c2 / P.br_ndf;
    if (status_out) status_out[j] = st;
  }
}

static int inv3(const double* S, double* Pi) {
  const double a = S[0], b = S[1], c = S[2], d = S[3], e = S[4], f = S[5], g = S[6], h = S[7], i = S[8];
  const double A = e * i - f * h, Bc = -(d * i - f * g), C = d * h - e * g;
  const double det = a * A + b * Bc + c * C;
  if (!(std::fabs(det) > 0.0) || !std::isfinite(det)) return 1;
  const double id = 1.0 / det;
  Pi[0] = A * id;  Pi[1] = -(b * i - c * h) * id; Pi[2] = (b * f - c * e) * id;
  Pi[3] = Bc * id; Pi[4] = (a * i - c * g) * id;  Pi[5] = -(a * f - c * d) * id;
  Pi[6] = C * id;  Pi[7] = -(a * h - b * g) * id; Pi[8] = (a * e - b * d) * id;
  return 0;
}

}  // namespace foct

using namespace foct;

// ------------------------------------------------------------------ plan
struct foct_plan {
  int kind = 0, n = 0, NN = 0, D = 0, P_out = 0, npad = 0, device = 0;
  size_t blob_stride = 0;
  foct_model_spec spec{};
  foct_sampler_cfg cfg{};
  int n_saved = 0, n_post = 0;
  bool want_draws = false, want_summary = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  double *d_blobs = nullptr, *d_draws = nullptr, *d_sparams = nullptr, *d_summary = nullptr, *d_stepsize = nullptr,
         *d_invm = nullptr, *d_nleap = nullptr, *d_ndiv = nullptr, *d_init = nullptr;
  DevProblem* d_probs = nullptr;
  int* d_counter = nullptr;
  int* d_order = nullptr;  // work-item order: longest expected fits first (see plan_order)
  const InstEntry* inst = nullptr;
  int grid = 0, block = 0, blocks_per_sm = 0, regs = 0, n_sm = 0;
  int groups = 1;  // work units per profile: groups of <= 4 chains (CTA-level items) or chain pairs (nuts2w_kernel)
  int wpc = 1;     // work units a CTA has in flight (2: the warps of nuts2w_kernel claim units on their own)
  int warp_units = 0, alt_groups = 1;
  size_t smem = 0;
  bool ran = false;
  // continuation inputs (cfg.inv_metric_init / stepsize_init) and the state a continuation starts from
  double *d_invm_init = nullptr, *d_eps_init = nullptr, *d_lastq = nullptr;
  // run until converged (cfg.rhat_target): extension blocks of the profiles that were continued
  bool extend_pending = false, thinned = false;
  int x_cap = 0, x_slots = 0, x_ext = 0;
  double *d_xdraws = nullptr, *d_xsparams = nullptr;
  int *d_slot_of = nullptr, *d_sel = nullptr, *d_sel_slots = nullptr;
  std::vector<int> n_extend, slot_of;
  cudaEvent_t evx0 = nullptr, evx1 = nullptr, evx2 = nullptr;
  float ext_sample_ms = 0.f, ext_summary_ms = 0.f;
  unsigned long long seed = 0;
  // progress / cancellation: device counter and flag, read and written from a side stream while the kernel runs
  unsigned long long* d_progress = nullptr;
  int* d_cancel = nullptr;
  unsigned long long* h_progress = nullptr;  // pinned
  cudaStream_t side = nullptr;
  bool cancelled = false;
  int shared_basis = 0;
  int launches = 0;  // kernels of the last run
  // time slicing of the work items (nuts2_kernel): saved chain state, ring queue, per-warp done flags
  double* d_slice_state = nullptr;
  unsigned long long* d_slice_queue = nullptr;
  int* d_slice_done = nullptr;
  int slice_ticks = 0;
  // which sampling kernel (foct_inst.cu: two chains per warp for batches larger than the GPU holds at once), and the
  // geometry of the one-chain-per-warp kernel for continuation rounds of few profiles
  int pair_kernel = 0;
  int alt_block = 0, alt_blocks_per_sm = 0;
  size_t alt_smem = 0;
};

// Copy the post-warm-up draws of the selected profiles into their extension blocks.
__global__ void gather_blocks_kernel(const double* __restrict__ src, size_t src_stride, size_t src_off, double* __restrict__ dst,
                                     size_t dst_stride, size_t len, const int* __restrict__ sel, const int* __restrict__ slots, int m) {
  for (int k = blockIdx.y; k < m; k += gridDim.y) {
    const double* a = src + (size_t)sel[k] * src_stride + src_off;
    double* b = dst + (size_t)slots[k] * dst_stride;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
  }
}
// Thin an extension block back into the profile's post-warm-up rows: row t <- row floor((t + 1) total / n_post) - 1 of
// the block, total = n_post + n_extend * ext draws.
__global__ void thin_blocks_kernel(const double* __restrict__ xsrc, size_t x_stride, double* __restrict__ dst, size_t dst_stride,
                                   size_t dst_off, int n_post, int ext, size_t row_len, const int* __restrict__ slot_of,
                                   const int* __restrict__ n_extend, int n) {
  for (int j = blockIdx.y; j < n; j += gridDim.y) {
    const int e = n_extend[j];
    if (e <= 0) continue;
    const long long total = (long long)n_post + (long long)e * ext;
    const double* a = xsrc + (size_t)slot_of[j] * x_stride;
    double* b = dst + (size_t)j * dst_stride + dst_off;
    const size_t tot = (size_t)n_post * row_len;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
      const size_t t = i / row_len, c = i % row_len;
      const size_t src = (size_t)(((long long)(t + 1) * total) / n_post - 1);
      b[i] = a[src * row_len + c];
    }
  }
}

static void plan_free(foct_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  // foct_plan_run is asynchronous: the kernels may still be reading and writing these buffers, and pool_free hands
  // them to the next pool_malloc of any thread without the implicit synchronisation cudaFree would have done
  if (p->stream && p->ran) cudaStreamSynchronize(p->stream);
  if (p->side) { cudaStreamSynchronize(p->side); cudaStreamDestroy(p->side); }
  pool_free(p->d_progress); pool_free(p->d_cancel); pinned_free(p->h_progress);
  pool_free(p->d_invm_init); pool_free(p->d_eps_init); pool_free(p->d_lastq);
  pool_free(p->d_xdraws); pool_free(p->d_xsparams); pool_free(p->d_slot_of); pool_free(p->d_sel); pool_free(p->d_sel_slots);
  if (p->evx0) cudaEventDestroy(p->evx0);
  if (p->evx1) cudaEventDestroy(p->evx1);
  if (p->evx2) cudaEventDestroy(p->evx2);
  pool_free(p->d_blobs); pool_free(p->d_draws); pool_free(p->d_sparams); pool_free(p->d_summary);
  pool_free(p->d_stepsize); pool_free(p->d_invm); pool_free(p->d_nleap); pool_free(p->d_ndiv); pool_free(p->d_init);
  pool_free(p->d_probs); pool_free(p->d_counter); pool_free(p->d_order);
  pool_free(p->d_slice_state); pool_free(p->d_slice_queue); pool_free(p->d_slice_done);
  if (p->ev0) cudaEventDestroy(p->ev0);
  if (p->ev1) cudaEventDestroy(p->ev1);
  if (p->ev2) cudaEventDestroy(p->ev2);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
}

// Streaming multiprocessors of the current device (grids are sized in multiples of it; 148 on B200).
int foct::sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    cached_dev = dev; cached = n;
  }
  return cached;
}

int foct::check_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n < 1)
    return fail(FOCT_ENODEV, "no CUDA device available (%s); fitoct_b200 has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  return 0;
}

static DevSpec dev_spec(const foct_model_spec& s) {
  DevSpec d;
  d.ygp_prior = s.ygp_prior; d.lambda_prior = s.lambda_prior; d.theta_prior = s.theta_prior;
  d.sigma_mean = s.sigma_mean; d.sigma_sd = s.sigma_sd; d.sigma_inv_sd = s.sigma_sd > 0.0 ? 1.0 / s.sigma_sd : 0.0;
  return d;
}

// Validate a batch, upload it and build blobs + DevProblem[] on `device`.
static int build_device_batch(int kind, const foct_problem* P, int n, const foct_model_spec* spec, int device,
                              cudaStream_t st, int* NN_out, int* npad_out, size_t* stride_out, double** d_blobs,
                              DevProblem** d_probs) {
  if (!P || n < 1 || !spec) return fail(FOCT_EINVAL, "empty batch or NULL argument");
  if (kind != FOCT_EXPGP && kind != FOCT_MONOEXP) return fail(FOCT_EINVAL, "unknown model kind %d", kind);
  const int NN = kind == FOCT_EXPGP ? P[0].Nn : 0;
  if (kind == FOCT_EXPGP && (NN < 1 || NN > FOCT_MAX_NN)) return fail(FOCT_EINVAL, "Nn=%d outside 1..%d", NN, FOCT_MAX_NN);
  int maxN = 0;
  size_t total = 0;
  for (int j = 0; j < n; ++j) {
    if (P[j].N < 4 || !P[j].x || !P[j].y || !P[j].uy) return fail(FOCT_EINVAL, "problem %d: N=%d or NULL data", j, P[j].N);
    if (kind == FOCT_EXPGP && P[j].Nn != NN) return fail(FOCT_EINVAL, "problem %d: Nn=%d differs from batch Nn=%d", j, P[j].Nn, NN);
    if (P[j].dataType != 1 && P[j].dataType != 2) return fail(FOCT_EINVAL, "problem %d: dataType=%d not in {1,2}", j, P[j].dataType);
    if (kind == FOCT_EXPGP && !(P[j].rho > 0.0)) return fail(FOCT_EINVAL, "problem %d: rho must be > 0 (resolve rho_scale==0 to 1/Nn as FitOCT.R:119 does)", j);
    if (P[j].N - 3 - NN < 1 && spec->br_ndf == 0) return fail(FOCT_EINVAL, "problem %d: N=%d too small for %d parameters", j, P[j].N, 3 + NN);
    maxN = std::max(maxN, P[j].N);
    total += 3 * (size_t)P[j].N;
  }
  const int npad = (maxN + 31) / 32 * 32;
  const size_t stride = (size_t)(3 + NN) * npad;
  if (stride * sizeof(double) > 220 * 1024) return fail(FOCT_EINVAL, "profile of %d points x %d control points does not fit shared memory", maxN, NN);

  double* h_up = nullptr;
  HostMeta* h_meta = nullptr;
  CU(cudaSetDevice(device));
  CU(pinned_malloc(&h_up, total * sizeof(double)));
  if (pinned_malloc(&h_meta, (size_t)n * sizeof(HostMeta)) != cudaSuccess) { pinned_free(h_up); return fail(FOCT_ENOMEM, "pinned alloc"); }
  size_t off = 0;
  int bad = -1;
  for (int j = 0; j < n; ++j) {
    const int N = P[j].N;
    std::memcpy(h_up + off, P[j].x, N * sizeof(double));
    std::memcpy(h_up + off + N, P[j].y, N * sizeof(double));
    std::memcpy(h_up + off + 2 * (size_t)N, P[j].uy, N * sizeof(double));
    HostMeta& M = h_meta[j];
    M.off = off; M.N = N; M.dataType = P[j].dataType; M.gridType = P[j].gridType; M.prior_PD = P[j].prior_PD;
    M.rho = P[j].rho; M.lambda_rate = P[j].lambda_rate; M.id = P[j].id;
    for (int k = 0; k < 3; ++k) M.theta0[k] = P[j].theta0[k];
    for (int k = 0; k < 9; ++k) M.Pinv[k] = 0.0;
    if (spec->theta_prior == 0 && inv3(P[j].Sigma0, M.Pinv)) bad = j;
    for (int i = 0; i < N; ++i)
      if (!(P[j].uy[i] > 0.0)) bad = j;
    off += 3 * (size_t)N;
  }
  int rc = 0;
  double* d_up = nullptr;
  HostMeta* d_meta = nullptr;
  int* d_status = nullptr;
  *d_blobs = nullptr; *d_probs = nullptr;
  do {
    if (bad >= 0) { rc = fail(FOCT_EINVAL, "problem %d: singular Sigma0 or non-positive uy", bad); break; }
#define CUB(call) if ((call) != cudaSuccess) { rc = fail(FOCT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(cudaGetLastError())); break; }
    CUB(pool_malloc(&d_up, total * sizeof(double)));
    CUB(pool_malloc(&d_meta, (size_t)n * sizeof(HostMeta)));
    CUB(pool_malloc(&d_status, sizeof(int)));
    CUB(pool_malloc(d_blobs, (size_t)n * stride * sizeof(double)));
    CUB(pool_malloc(d_probs, (size_t)n * sizeof(DevProblem)));
    CUB(cudaMemcpyAsync(d_up, h_up, total * sizeof(double), cudaMemcpyHostToDevice, st));
    CUB(cudaMemcpyAsync(d_meta, h_meta, (size_t)n * sizeof(HostMeta), cudaMemcpyHostToDevice, st));
    CUB(cudaMemsetAsync(d_status, 0, sizeof(int), st));
    const int grid = std::min(n, sm_count() * 8);
    setup_kernel<<<grid, 128, 0, st>>>(d_up, d_meta, n, NN, npad, stride, *d_blobs, *d_probs, spec->kernel, spec->jitter,
                                       spec->br_ndf, d_status);
    CUB(cudaGetLastError());
    int h_status = 0;
    CUB(cudaMemcpyAsync(&h_status, d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUB(cudaStreamSynchronize(st));
    if (h_status) { rc = fail(FOCT_EINVAL, "problem %d: degenerate depth grid or Kgg not positive definite (rho, jitter?)", h_status - 1); break; }
#undef CUB
  } while (0);
  pool_free(d_up); pool_free(d_meta); pool_free(d_status);
  pinned_free(h_up); pinned_free(h_meta);
  if (rc) { pool_free(*d_blobs); pool_free(*d_probs); *d_blobs = nullptr; *d_probs = nullptr; return rc; }
  *NN_out = NN; *npad_out = npad; *stride_out = stride;
  return 0;
}

// ------------------------------------------------------------------ ABI: trivia
extern "C" int foct_version(void) { return FOCT_ABI_VERSION; }
extern "C" const char* foct_last_error(void) { return g_err.c_str(); }
extern "C" int foct_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
extern "C" void foct_model_spec_default(foct_model_spec* s, int kind) {
  s->modulation = 0; s->kernel = 0; s->jitter = 1e-9; s->ygp_prior = 0; s->lambda_prior = 0;
  s->sigma_mean = 1.0; s->sigma_sd = 0.1; s->theta_prior = kind == FOCT_EXPGP ? 0 : 1; s->br_ndf = 0;
}
extern "C" void foct_sampler_cfg_default(foct_sampler_cfg* c) {
  std::memset(c, 0, sizeof(*c));
  c->chains = 4; c->n_warmup = 500; c->n_iter = 1500;  // FitOCT.R:43-44, server.R:469
  c->adapt_delta = 0.8; c->max_treedepth = 10; c->stepsize0 = 1.0; c->seed = 1234;
}
extern "C" int foct_dims(int kind, int Nn, int* D, int* P_out) {
  if (kind == FOCT_EXPGP) {
    if (Nn < 1 || Nn > FOCT_MAX_NN) return fail(FOCT_EINVAL, "Nn=%d outside 1..%d", Nn, FOCT_MAX_NN);
    if (D) *D = Nn + 5;
    if (P_out) *P_out = Nn + 7;
  } else if (kind == FOCT_MONOEXP) {
    if (D) *D = 3;
    if (P_out) *P_out = 5;
  } else {
    return fail(FOCT_EINVAL, "unknown model kind %d", kind);
  }
  return 0;
}
extern "C" int foct_expgp_grid(int Nn, int gridType, double* xGP) {
  if (Nn < 1 || Nn > FOCT_MAX_NN || !xGP) return fail(FOCT_EINVAL, "Nn=%d outside 1..%d", Nn, FOCT_MAX_NN);
  const double dx = 1.0 / (Nn + 1);
  const double lo = gridType == FOCT_GRID_INTERNAL ? 0.5 * dx : 0.0, hi = gridType == FOCT_GRID_INTERNAL ? 1.0 - 0.5 * dx : 1.0;
  for (int k = 0; k < Nn; ++k) xGP[k] = Nn == 1 ? lo : lo + (hi - lo) * (double)k / (double)(Nn - 1);
  return 0;
}

// ------------------------------------------------------------------ ABI: basis / logp / predict
extern "C" int foct_expgp_basis(const foct_problem* P, const foct_model_spec* spec, double* B_out) {
  if (int rc = check_device()) return rc;
  if (!B_out) return fail(FOCT_EINVAL, "NULL output");
  int dev = 0, NN, npad;
  size_t stride;
  double* d_blobs;
  DevProblem* d_probs;
  CU(cudaGetDevice(&dev));
  if (int rc = build_device_batch(FOCT_EXPGP, P, 1, spec, dev, 0, &NN, &npad, &stride, &d_blobs, &d_probs)) return rc;
  std::vector<double> h(stride);
  cudaError_t e = cudaMemcpy(h.data(), d_blobs, stride * sizeof(double), cudaMemcpyDeviceToHost);
  pool_free(d_blobs); pool_free(d_probs);
  if (e == cudaSuccess)
    for (int k = 0; k < NN; ++k)
      for (int i = 0; i < P->N; ++i) B_out[(size_t)k * P->N + i] = h[blob_index(i, 3 + k, NN)];
  if (e != cudaSuccess) return fail(FOCT_ECUDA, "copy of basis failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int foct_logp_grad(int kind, const foct_problem* P, int n, const foct_model_spec* spec, const double* q,
                              int n_q, double* lp, double* grad, double* chi2) {
  if (int rc = check_device()) return rc;
  if (!q || !lp || !grad || n_q < 1) return fail(FOCT_EINVAL, "NULL q/lp/grad or n_q < 1");
  int dev = 0, NN, npad;
  size_t stride;
  double* d_blobs;
  DevProblem* d_probs;
  CU(cudaGetDevice(&dev));
  if (int rc = build_device_batch(kind, P, n, spec, dev, 0, &NN, &npad, &stride, &d_blobs, &d_probs)) return rc;
  const int D = kind == FOCT_EXPGP ? NN + 5 : 3;
  const size_t nq = (size_t)n * n_q;
  double *d_q = nullptr, *d_lp = nullptr, *d_g = nullptr, *d_c2 = nullptr;
  int rc = 0;
  do {
#define CUB(call) if ((call) != cudaSuccess) { rc = fail(FOCT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(cudaGetLastError())); break; }
    CUB(pool_malloc(&d_q, nq * D * sizeof(double)));
    CUB(pool_malloc(&d_g, nq * D * sizeof(double)));
    CUB(pool_malloc(&d_lp, nq * sizeof(double)));
    CUB(pool_malloc(&d_c2, nq * sizeof(double)));
    CUB(cudaMemcpy(d_q, q, nq * D * sizeof(double), cudaMemcpyHostToDevice));
    LogpParams K;
    K.blobs = d_blobs; K.blob_stride = stride; K.npad = npad; K.probs = d_probs; K.n_problems = n;
    K.spec = dev_spec(*spec); K.q = d_q; K.n_q = n_q; K.lp = d_lp; K.grad = d_g; K.chi2 = d_c2;
    {  // the parity hook exercises the evaluations the samplers use.  D <= 16: half-warps on a staged blob (default, the
       // pair kernels on ragged grids), FOCT_LOGP_WIDTH=17: half-warps in the shared-basis layout of nuts2w_kernel (the kernel
       // of BASELINE-size batches, software-pipelined sweep), FOCT_LOGP_WIDTH=32: full warps (one chain per warp / latency kernel)
      const char* wenv = std::getenv("FOCT_LOGP_WIDTH");
      const int w = wenv ? std::atoi(wenv) : 16;
      K.width = (w == 32 || w == 17) ? w : 16;
    }
    const InstEntry* inst = inst_for(NN);
    const size_t logp_smem = stride * sizeof(double) + (K.width == 17 ? (size_t)2 * npad * sizeof(double) : 0);
    CUB(inst->launch_logp(spec->modulation, std::min(n, sm_count() * 4), 128, logp_smem, 0, K));
    CUB(cudaDeviceSynchronize());
    CUB(cudaMemcpy(lp, d_lp, nq * sizeof(double), cudaMemcpyDeviceToHost));
    CUB(cudaMemcpy(grad, d_g, nq * D * sizeof(double), cudaMemcpyDeviceToHost));
    if (chi2) CUB(cudaMemcpy(chi2, d_c2, nq * sizeof(double), cudaMemcpyDeviceToHost));
#undef CUB
  } while (0);
  pool_free(d_q); pool_free(d_g); pool_free(d_lp); pool_free(d_c2); pool_free(d_blobs); pool_free(d_probs);
  return rc;
}

extern "C" int foct_summary(const double* draws, int n_sets, int n_draws, int chains, int n_cols, double* summary) {
  if (int rc = check_device()) return rc;
  if (!draws || !summary) return fail(FOCT_EINVAL, "NULL draws / summary");
  if (n_sets < 1 || n_cols < 1 || n_draws < 4 || chains < 1 || chains > FOCT_MAX_CHAINS)
    return fail(FOCT_EINVAL, "need n_sets, n_cols >= 1, n_draws >= 4, 1 <= chains <= %d", FOCT_MAX_CHAINS);
  const size_t nd = (size_t)n_sets * n_draws * chains * n_cols, ns = (size_t)n_sets * n_cols * FOCT_N_SUMMARY_COLS;
  double *d_dr = nullptr, *d_su = nullptr;
  int rc = 0;
  do {
#define CUB(call) if ((call) != cudaSuccess) { rc = fail(FOCT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(cudaGetLastError())); break; }
    CUB(pool_malloc(&d_dr, nd * sizeof(double)));
    CUB(pool_malloc(&d_su, ns * sizeof(double)));
    CUB(cudaMemcpy(d_dr, draws, nd * sizeof(double), cudaMemcpyHostToDevice));
    CUB(launch_summary(d_dr, n_sets, n_draws, 0, n_draws, chains, n_cols, d_su, nullptr));
    CUB(cudaDeviceSynchronize());
    CUB(cudaMemcpy(summary, d_su, ns * sizeof(double), cudaMemcpyDeviceToHost));
#undef CUB
  } while (0);
  pool_free(d_dr); pool_free(d_su);
  return rc;
}

extern "C" int foct_predict(int kind, const foct_problem* P, const foct_model_spec* spec, const double* draws,
                            int n_draws, double* m, double* resid, double* dL) {
  if (int rc = check_device()) return rc;
  if (!draws || n_draws < 1) return fail(FOCT_EINVAL, "no draws");
  int dev = 0, NN, npad;
  size_t stride;
  double* d_blobs;
  DevProblem* d_probs;
  CU(cudaGetDevice(&dev));
  if (int rc = build_device_batch(kind, P, 1, spec, dev, 0, &NN, &npad, &stride, &d_blobs, &d_probs)) return rc;
  const int P_out = kind == FOCT_EXPGP ? NN + 7 : 5;
  const size_t tot = (size_t)n_draws * P->N;
  double *d_dr = nullptr, *d_m = nullptr, *d_r = nullptr, *d_dl = nullptr;
  int rc = 0;
  do {
#define CUB(call) if ((call) != cudaSuccess) { rc = fail(FOCT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(cudaGetLastError())); break; }
    CUB(pool_malloc(&d_dr, (size_t)n_draws * P_out * sizeof(double)));
    CUB(pool_malloc(&d_m, tot * sizeof(double)));
    CUB(pool_malloc(&d_r, tot * sizeof(double)));
    CUB(pool_malloc(&d_dl, tot * sizeof(double)));
    CUB(cudaMemcpy(d_dr, draws, (size_t)n_draws * P_out * sizeof(double), cudaMemcpyHostToDevice));
    predict_kernel<<<(int)std::min<size_t>((tot + 255) / 256, (size_t)sm_count() * 16), 256>>>(d_blobs, npad, P->N, NN, spec->modulation, P_out,
                                                                                d_dr, n_draws, d_m, d_r, d_dl);
    CUB(cudaGetLastError());
    CUB(cudaDeviceSynchronize());
    if (m) CUB(cudaMemcpy(m, d_m, tot * sizeof(double), cudaMemcpyDeviceToHost));
    if (resid) CUB(cudaMemcpy(resid, d_r, tot * sizeof(double), cudaMemcpyDeviceToHost));
    if (dL) CUB(cudaMemcpy(dL, d_dl, tot * sizeof(double), cudaMemcpyDeviceToHost));
#undef CUB
  } while (0);
  pool_free(d_dr); pool_free(d_m); pool_free(d_r); pool_free(d_dl); pool_free(d_blobs); pool_free(d_probs);
  return rc;
}

// ------------------------------------------------------------------ ABI: plan
static int validate_cfg(const foct_sampler_cfg* c) {
  if (!c) return fail(FOCT_EINVAL, "NULL sampler cfg");
  if (c->chains < 1 || c->chains > FOCT_MAX_CHAINS) return fail(FOCT_EINVAL, "chains=%d outside 1..%d", c->chains, FOCT_MAX_CHAINS);
  if (c->n_warmup < 0 || c->n_iter < c->n_warmup || c->n_iter < 1) return fail(FOCT_EINVAL, "need 0 <= n_warmup <= n_iter, n_iter >= 1 (nb_iter = nb_warmup + nb_sample, FitOCT.R:121)");
  if (c->max_treedepth > FOCT_STACK_LEVELS + 1) return fail(FOCT_EINVAL, "max_treedepth=%d > %d", c->max_treedepth, FOCT_STACK_LEVELS + 1);
  if (c->init_mode < 0 || c->init_mode > 2 || (c->init_mode == 2 && !c->init)) return fail(FOCT_EINVAL, "bad init_mode / init");
  if (c->iter_offset < 0) return fail(FOCT_EINVAL, "iter_offset=%d < 0", c->iter_offset);
  if (c->rhat_target > 0.0 && (c->max_extend < 0 || c->max_extend > 64)) return fail(FOCT_EINVAL, "max_extend=%d outside 0..64", c->max_extend);
  if (c->rhat_target > 0.0 && !(c->rhat_target > 1.0)) return fail(FOCT_EINVAL, "rhat_target=%g must exceed 1", c->rhat_target);
  if (c->extend_iter < 0) return fail(FOCT_EINVAL, "extend_iter=%d < 0", c->extend_iter);
  return 0;
}

static int plan_create_on(int device, int kind, const foct_problem* P, int n, const foct_model_spec* spec,
                          const foct_sampler_cfg* cfg, const double* init_slice, const double* invm_slice,
                          const double* eps_slice, int want_draws, int want_summary, foct_plan** out) {
  if (!spec || !P || !out) return fail(FOCT_EINVAL, "NULL problem / spec / plan pointer");
  if (int rc = validate_cfg(cfg)) return rc;
  if (cfg->rhat_target > 0.0 && cfg->max_extend > 0 && !want_summary)
    return fail(FOCT_EINVAL, "rhat_target needs the summary output (the continuation is decided from the split R-hat)");
  foct_plan* p = new foct_plan();
  p->device = device; p->kind = kind; p->n = n; p->spec = *spec; p->cfg = *cfg;
  p->cfg.init = nullptr; p->cfg.devices = nullptr; p->cfg.n_devices = 0;
  p->cfg.inv_metric_init = nullptr; p->cfg.stepsize_init = nullptr;
  p->want_draws = want_draws != 0; p->want_summary = want_summary != 0;
  p->n_post = cfg->n_iter - cfg->n_warmup;
  p->n_saved = cfg->save_warmup ? cfg->n_iter : p->n_post;
#define CUP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int rc_ = fail(FOCT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); plan_free(p); return rc_; } } while (0)
  Trace tr("plan_create");
  CUP(cudaSetDevice(device));
  CUP(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  CUP(cudaEventCreate(&p->ev0));
  CUP(cudaEventCreate(&p->ev1));
  CUP(cudaEventCreate(&p->ev2));
  CUP(cudaEventCreate(&p->evx0));
  CUP(cudaEventCreate(&p->evx1));
  CUP(cudaEventCreate(&p->evx2));
  CUP(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
  CUP(pool_malloc(&p->d_progress, sizeof(unsigned long long)));
  CUP(pool_malloc(&p->d_cancel, sizeof(int)));
  CUP(pinned_malloc(&p->h_progress, 64));
  *p->h_progress = 0ull;
  tr.mark("stream + events");
  if (int rc = build_device_batch(kind, P, n, spec, device, p->stream, &p->NN, &p->npad, &p->blob_stride, &p->d_blobs, &p->d_probs)) {
    plan_free(p);
    return rc;
  }
  tr.mark("pack + upload + setup kernel");
  p->D = kind == FOCT_EXPGP ? p->NN + 5 : 3;
  p->P_out = p->D + 2;
  const size_t pc = (size_t)n * cfg->chains;
  const bool need_draws = p->want_draws || p->want_summary;
  if (need_draws) {
    CUP(pool_malloc(&p->d_draws, pc * p->n_saved * p->P_out * sizeof(double)));
    if (p->want_draws) CUP(pool_malloc(&p->d_sparams, pc * p->n_saved * 6 * sizeof(double)));
  }
  if (p->want_summary) CUP(pool_malloc(&p->d_summary, (size_t)n * p->P_out * FOCT_N_SUMMARY_COLS * sizeof(double)));
  CUP(pool_malloc(&p->d_stepsize, pc * sizeof(double)));
  CUP(pool_malloc(&p->d_invm, pc * p->D * sizeof(double)));
  CUP(pool_malloc(&p->d_nleap, pc * 2 * sizeof(double)));
  CUP(pool_malloc(&p->d_ndiv, pc * sizeof(double)));
  CUP(pool_malloc(&p->d_counter, (4 + 64) * sizeof(int)));  // work counter; with time slicing: tickets | pushes | finished | - | waiting units by progress bin
  if (cfg->init_mode == 2) {
    CUP(pool_malloc(&p->d_init, pc * p->D * sizeof(double)));
    CUP(cudaMemcpy(p->d_init, init_slice, pc * p->D * sizeof(double), cudaMemcpyHostToDevice));
  }
  CUP(pool_malloc(&p->d_lastq, pc * p->D * sizeof(double)));
  if (invm_slice) {
    CUP(pool_malloc(&p->d_invm_init, pc * p->D * sizeof(double)));
    CUP(cudaMemcpy(p->d_invm_init, invm_slice, pc * p->D * sizeof(double), cudaMemcpyHostToDevice));
  }
  if (eps_slice) {
    CUP(pool_malloc(&p->d_eps_init, pc * sizeof(double)));
    CUP(cudaMemcpy(p->d_eps_init, eps_slice, pc * sizeof(double), cudaMemcpyHostToDevice));
  }
  tr.mark("output buffers");
  p->inst = inst_for(p->NN);
  if (!p->inst) { plan_free(p); return fail(FOCT_EINVAL, "this build carries no kernels for Nn=%d", p->NN); }
  // One depth grid for the whole batch (the profiles of one acquisition: same x after selX, FitOCT.R:85-86) means one
  // GP basis: the sampling kernel then stages only cx | y | w per profile and reads the basis of blob 0 through L1.
  int shared_basis = kind == FOCT_EXPGP && n > 1;
  for (int j = 1; j < n && shared_basis; ++j)
    shared_basis = P[j].N == P[0].N && P[j].gridType == P[0].gridType && P[j].rho == P[0].rho && P[j].dataType == P[0].dataType &&
                   std::memcmp(P[j].x, P[0].x, (size_t)P[0].N * sizeof(double)) == 0;
  int cta_chains = FOCT_CTA_CHAINS;
  size_t slice_bytes = 0;
  int n_sm = 0;  // (cudaGetDeviceProperties costs milliseconds; one attribute does not)
  CUP(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
  CUP(p->inst->nuts_occupancy(spec->modulation, cfg->chains, n, n_sm, p->blob_stride * sizeof(double), (size_t)3 * p->npad * sizeof(double),
                              &shared_basis, &p->smem, &p->block, &p->blocks_per_sm, &cta_chains, &p->regs, &slice_bytes, &p->pair_kernel, &p->warp_units));
  p->shared_basis = shared_basis;
  if (const char* env = std::getenv("FOCT_SMEM_PAD")) p->smem += (size_t)std::atoi(env);  // experiments: shrink L1 by unused shared memory
  if (p->blocks_per_sm < 1) { plan_free(p); return fail(FOCT_EINVAL, "sampling kernel does not fit an SM (smem %zu B)", p->smem); }
  if (p->pair_kernel) {
    // (n_items = 0: the geometry of the one-chain-per-warp kernel, if the blob fits an SM at all)
    int sb = 0, cc = 0, rg = 0, pk = 0, wu = 0;
    size_t sl = 0;
    if (p->inst->nuts_occupancy(spec->modulation, cfg->chains, 0, n_sm, p->blob_stride * sizeof(double), (size_t)3 * p->npad * sizeof(double),
                                &sb, &p->alt_smem, &p->alt_block, &p->alt_blocks_per_sm, &cc, &rg, &sl, &pk, &wu) != cudaSuccess || pk) {
      cudaGetLastError();
      p->alt_blocks_per_sm = 0;
    }
  }
  p->alt_groups = (cfg->chains + FOCT_CTA_CHAINS - 1) / FOCT_CTA_CHAINS;
  const int groups = p->warp_units ? (cfg->chains + 1) / 2 : (cfg->chains + cta_chains - 1) / cta_chains;
  p->groups = groups; p->n_sm = n_sm; p->wpc = p->warp_units ? p->warp_units : 1;  // (warp_units = warps per CTA of nuts2w_kernel)
  p->grid = (int)std::min<long long>(((long long)n * groups + p->wpc - 1) / p->wpc, (long long)n_sm * p->blocks_per_sm);
  if (const char* env = std::getenv("FOCT_MAX_GRID")) p->grid = std::max(1, std::min(p->grid, std::atoi(env)));  // tests: few CTAs
  const long long resident = (long long)p->grid * p->wpc;  // work units in flight at once
  // More work items than resident CTAs: the CTAs share them in time slices (nuts2_kernel) instead of running each to
  // completion, so that the last fits of the batch do not run alone.  FOCT_SLICE_TICKS=0 turns it off (A/B runs).
  if (slice_bytes > 0 && (long long)n * groups > resident) {
    int ticks = 4096;  // gradient evaluations per slice: ~1 % of a fit, ~50 ms; a state round trip is 160 KB per item
    if (const char* env = std::getenv("FOCT_SLICE_TICKS")) ticks = std::atoi(env);
    if (ticks > 0) {
      const size_t items = (size_t)n * groups;
      CUP(pool_malloc(&p->d_slice_state, items * slice_bytes));
      CUP(pool_malloc(&p->d_slice_queue, items * sizeof(unsigned long long)));
      CUP(pool_malloc(&p->d_slice_done, items * 2 * sizeof(int)));
      p->slice_ticks = ticks;
    }
  }
  // Longest-processing-time-first scheduling.  Fits differ in cost by up to 4x (an unmodulated profile needs a third of
  // the leapfrogs of a strongly modulated one), and with more work items than resident CTAs the order in which the
  // persistent CTAs claim them decides how long the last ones run alone.  The Birge ratio of a mono-exponential MAP fit
  // (map_kernel on the blobs already on the device, ~0.2 ms per 1000 profiles) ranks the fits by expected cost (Spearman
  // 0.9 against the leapfrog count); the result of a fit does not depend on when it runs (per-profile Philox keys).
  tr.mark("occupancy + device properties");
  if (kind == FOCT_EXPGP && (long long)n * groups > resident && !std::getenv("FOCT_NO_LPT")) {
    double *d_th = nullptr, *d_br = nullptr;
    int* d_st = nullptr;
    CUP(pool_malloc(&d_th, (size_t)n * 3 * sizeof(double)));
    CUP(pool_malloc(&d_br, (size_t)n * sizeof(double)));
    CUP(pool_malloc(&d_st, (size_t)n * sizeof(int)));
    map_kernel<<<(n + 3) / 4, 128, 0, p->stream>>>(p->d_blobs, p->blob_stride, p->NN, p->d_probs, n, /*theta_prior flat*/ 1,
                                                  nullptr, d_th, nullptr, d_br, d_st);
    std::vector<double> br(n);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(br.data(), d_br, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    pool_free(d_th); pool_free(d_br); pool_free(d_st);
    if (e != cudaSuccess) { plan_free(p); return fail(FOCT_ECUDA, "cost ranking failed: %s", cudaGetErrorString(e)); }
    std::vector<int> order(n);
    for (int j = 0; j < n; ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      const double ca = std::isfinite(br[a]) ? br[a] : 1e300, cb = std::isfinite(br[b]) ? br[b] : 1e300;  // failed fit: assume costly
      return ca > cb;
    });
    CUP(pool_malloc(&p->d_order, (size_t)n * sizeof(int)));
    CUP(cudaMemcpy(p->d_order, order.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    tr.mark("cost ranking");
  }
#undef CUP
  *out = p;
  return 0;
}

extern "C" int foct_plan_create(int kind, const foct_problem* P, int n, const foct_model_spec* spec,
                                const foct_sampler_cfg* cfg, int want_draws, int want_summary, foct_plan** plan) {
  if (int rc = check_device()) return rc;
  if (!plan) return fail(FOCT_EINVAL, "NULL plan pointer");
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (cfg && cfg->n_devices == 1 && cfg->devices) dev = cfg->devices[0];
  return plan_create_on(dev, kind, P, n, spec, cfg, cfg ? cfg->init : nullptr, cfg ? cfg->inv_metric_init : nullptr,
                        cfg ? cfg->stepsize_init : nullptr, want_draws, want_summary, plan);
}

static void plan_params(const foct_plan* p, unsigned long long seed, SamplerParams& K) {
  std::memset(&K, 0, sizeof(K));
  K.blobs = p->d_blobs; K.blob_stride = p->blob_stride; K.npad = p->npad; K.probs = p->d_probs; K.n_problems = p->n;
  K.spec = dev_spec(p->spec);
  const foct_sampler_cfg& c = p->cfg;
  K.chains = c.chains; K.n_warmup = c.n_warmup; K.n_iter = c.n_iter; K.max_depth = c.max_treedepth > 0 ? c.max_treedepth : 10;
  K.save_warmup = c.save_warmup; K.init_mode = c.init_mode;
  K.adapt_delta = c.adapt_delta; K.stepsize0 = c.stepsize0; K.gamma = c.gamma; K.kappa = c.kappa; K.t0 = c.t0;
  K.init_buffer = c.init_buffer; K.term_buffer = c.term_buffer; K.window = c.window;
  K.seed = seed; K.init = p->d_init;
  K.draws = p->d_draws; K.sparams = p->d_sparams; K.stepsize = p->d_stepsize; K.inv_metric = p->d_invm;
  K.n_leapfrog = p->d_nleap; K.n_divergent = p->d_ndiv; K.work_counter = p->d_counter; K.order = p->d_order;
  K.invm_init = p->d_invm_init; K.eps_init = p->d_eps_init; K.last_q = p->d_lastq; K.it_offset = c.iter_offset;
  K.progress = p->d_progress; K.cancel = p->d_cancel; K.shared_basis = p->shared_basis;
  K.slice_state = p->d_slice_state; K.slice_queue = p->d_slice_queue; K.slice_done = p->d_slice_done;
  K.slice_ctl = reinterpret_cast<unsigned*>(p->d_counter); K.slice_ticks = p->slice_ticks;
  K.pair_kernel = p->pair_kernel; K.warp_units = p->warp_units;
  // progress-aware yielding (foct_nuts2.cuh); FOCT_SLICE_PRIO=0: plain round robin (A/B runs)
  const char* prio_env = std::getenv("FOCT_SLICE_PRIO");
  const bool prio = !(prio_env && std::atoi(prio_env) == 0);
  K.slice_hist = (p->warp_units && p->d_slice_state && prio) ? p->d_counter + 4 : nullptr;
}

extern "C" int foct_plan_run(foct_plan* p, unsigned long long seed) {
  if (!p) return fail(FOCT_EINVAL, "NULL plan");
  CU(cudaSetDevice(p->device));
  SamplerParams K;
  plan_params(p, seed, K);
  const foct_sampler_cfg& c = p->cfg;
  CU(cudaMemsetAsync(p->d_counter, 0, (4 + 64) * sizeof(int), p->stream));
  if (p->d_slice_queue) CU(cudaMemsetAsync(p->d_slice_queue, 0xff, (size_t)p->n * p->groups * sizeof(unsigned long long), p->stream));
  CU(cudaMemsetAsync(p->d_progress, 0, sizeof(unsigned long long), p->stream));
  CU(cudaMemsetAsync(p->d_cancel, 0, sizeof(int), p->stream));
  p->cancelled = false;
  CU(cudaEventRecord(p->ev0, p->stream));
  CU(p->inst->launch_nuts(p->spec.modulation, p->grid, p->block, p->smem, p->stream, K));
  CU(cudaEventRecord(p->ev1, p->stream));
  if (p->want_summary && p->n_post >= 2) {
    const int off = c.save_warmup ? c.n_warmup : 0;
    CU(launch_summary(p->d_draws, p->n, p->n_saved, off, p->n_post, c.chains, p->P_out, p->d_summary, p->stream));
  }
  CU(cudaEventRecord(p->ev2, p->stream));
  p->ran = true;
  p->launches = 1 + (p->want_summary && p->n_post >= 2 ? 1 : 0);
  p->seed = seed;
  p->thinned = false;
  p->ext_sample_ms = p->ext_summary_ms = 0.f;
  p->n_extend.assign(p->n, 0);
  p->extend_pending = c.rhat_target > 0.0 && c.max_extend > 0 && p->want_summary && p->n_post >= 2;
  return 0;
}

// Run until converged (cfg.rhat_target): continue the profiles whose largest split R-hat over the sampled parameters
// is still >= the target — same adapted metric and step size, start = the last state, Philox sites after the ones
// already used — for another n_post draws per round, and summarise all their post-warm-up draws again.  Host logic
// between device rounds: called by the first sync / fetch / timing after foct_plan_run.
static int plan_extend(foct_plan* p) {
  if (!p->extend_pending) return 0;
  p->extend_pending = false;
  if (p->cancelled) return 0;
  const foct_sampler_cfg& c = p->cfg;
  const int n = p->n, C = c.chains, P_out = p->P_out, D = p->D, n_post = p->n_post;
  CU(cudaStreamSynchronize(p->stream));
  std::vector<double> rh((size_t)n * P_out * FOCT_N_SUMMARY_COLS);
  std::vector<int> sel, slots;
  auto select = [&](const std::vector<int>& among) {
    std::vector<int> out;
    for (int j : among) {
      double mx = 0.0;
      for (int d = 0; d < D; ++d) {
        const double r = rh[((size_t)j * P_out + d) * FOCT_N_SUMMARY_COLS + 9];
        if (r > mx) mx = r;  // NaN (a constant or non-finite column) never selects
      }
      if (mx >= c.rhat_target) out.push_back(j);
    }
    return out;
  };
  CU(cudaMemcpy(rh.data(), p->d_summary, rh.size() * sizeof(double), cudaMemcpyDeviceToHost));
  std::vector<int> all(n);
  for (int j = 0; j < n; ++j) all[j] = j;
  sel = select(all);
  if (sel.empty()) return 0;
  // extension blocks: one per profile selected in the first round
  const int m0 = (int)sel.size();
  const int ext = c.extend_iter > 0 ? c.extend_iter : std::max(50, (n_post + 3) / 4);
  p->x_ext = ext;
  p->x_cap = n_post + c.max_extend * ext;
  p->x_slots = m0;
  p->slot_of.assign(n, -1);
  for (int k = 0; k < m0; ++k) p->slot_of[sel[k]] = k;
  const size_t row = (size_t)C * P_out, blk = (size_t)p->x_cap * row;
  pool_free(p->d_xdraws); pool_free(p->d_xsparams); pool_free(p->d_slot_of); pool_free(p->d_sel); pool_free(p->d_sel_slots);
  p->d_xdraws = p->d_xsparams = nullptr; p->d_slot_of = p->d_sel = p->d_sel_slots = nullptr;
  CU(pool_malloc(&p->d_xdraws, (size_t)m0 * blk * sizeof(double)));
  if (p->d_sparams) CU(pool_malloc(&p->d_xsparams, (size_t)m0 * p->x_cap * C * 6 * sizeof(double)));
  CU(pool_malloc(&p->d_slot_of, (size_t)n * sizeof(int)));
  CU(pool_malloc(&p->d_sel, (size_t)m0 * sizeof(int)));
  CU(pool_malloc(&p->d_sel_slots, (size_t)m0 * sizeof(int)));
  CU(cudaMemcpyAsync(p->d_slot_of, p->slot_of.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, p->stream));
  const int off = c.save_warmup ? c.n_warmup : 0;
  CU(cudaEventRecord(p->evx0, p->stream));
  for (int e = 1; e <= c.max_extend && !sel.empty(); ++e) {
    const int m = (int)sel.size();
    slots.resize(m);
    for (int k = 0; k < m; ++k) slots[k] = p->slot_of[sel[k]];
    CU(cudaMemcpyAsync(p->d_sel, sel.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    CU(cudaMemcpyAsync(p->d_sel_slots, slots.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    if (e == 1) {
      const dim3 grid(32, std::min(m, 65535));
      gather_blocks_kernel<<<grid, 256, 0, p->stream>>>(p->d_draws, (size_t)p->n_saved * row, (size_t)off * row, p->d_xdraws, blk,
                                                        (size_t)n_post * row, p->d_sel, p->d_sel_slots, m);
      if (p->d_xsparams)
        gather_blocks_kernel<<<grid, 256, 0, p->stream>>>(p->d_sparams, (size_t)p->n_saved * C * 6, (size_t)off * C * 6, p->d_xsparams,
                                                          (size_t)p->x_cap * C * 6, (size_t)n_post * C * 6, p->d_sel, p->d_sel_slots, m);
      CU(cudaGetLastError());
      p->launches += p->d_xsparams ? 2 : 1;
    }
    p->launches += 2;
    SamplerParams K;
    plan_params(p, p->seed, K);
    K.n_problems = m; K.order = p->d_sel;
    K.n_warmup = 0; K.n_iter = ext; K.save_warmup = 0;
    K.init_mode = 2; K.init = p->d_lastq; K.invm_init = p->d_invm; K.eps_init = p->d_stepsize;
    K.it_offset = c.iter_offset + c.n_iter + (e - 1) * ext;
    K.draws = p->d_xdraws; K.sparams = p->d_xsparams; K.slot_of = p->d_slot_of;
    K.save_stride = p->x_cap; K.save_offset = n_post + (e - 1) * ext; K.accumulate = 1;
    CU(cudaMemsetAsync(p->d_counter, 0, (4 + 64) * sizeof(int), p->stream));
    if ((long long)m * p->groups <= (long long)p->n_sm * p->blocks_per_sm * p->wpc) K.slice_state = nullptr;  // every unit has its CTA / warp
    else if (p->d_slice_queue) CU(cudaMemsetAsync(p->d_slice_queue, 0xff, (size_t)m * p->groups * sizeof(unsigned long long), p->stream));
    cudaEvent_t a0, a1, a2;
    CU(cudaEventCreate(&a0)); CU(cudaEventCreate(&a1)); CU(cudaEventCreate(&a2));
    CU(cudaEventRecord(a0, p->stream));
    // few profiles left: a warp per chain (a chain advances ~1.5x faster than with half a warp, and the GPU is not full)
    const bool one_chain = p->pair_kernel && p->alt_blocks_per_sm > 0 && (long long)m * p->alt_groups <= (long long)p->n_sm * p->alt_blocks_per_sm;
    if (one_chain) { K.pair_kernel = 0; K.warp_units = 0; K.shared_basis = 0; K.slice_state = nullptr; }
    const int bpsm = one_chain ? p->alt_blocks_per_sm : p->blocks_per_sm, wpc = one_chain ? 1 : p->wpc;
    const int grid = (int)std::min<long long>(((long long)m * (one_chain ? p->alt_groups : p->groups) + wpc - 1) / wpc, (long long)p->n_sm * bpsm);
    CU(p->inst->launch_nuts(p->spec.modulation, grid, one_chain ? p->alt_block : p->block, one_chain ? p->alt_smem : p->smem, p->stream, K));
    CU(cudaEventRecord(a1, p->stream));
    CU(launch_summary(p->d_xdraws, m, p->x_cap, 0, n_post + e * ext, C, P_out, p->d_summary, p->stream, p->d_sel_slots, p->d_sel));
    CU(cudaEventRecord(a2, p->stream));
    CU(cudaMemcpyAsync(rh.data(), p->d_summary, rh.size() * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    float t1 = 0.f, t2 = 0.f;
    cudaEventElapsedTime(&t1, a0, a1); cudaEventElapsedTime(&t2, a1, a2);
    cudaEventDestroy(a0); cudaEventDestroy(a1); cudaEventDestroy(a2);
    p->ext_sample_ms += t1; p->ext_summary_ms += t2;
    for (int j : sel) p->n_extend[j] = e;
    sel = select(sel);
  }
  return 0;
}

extern "C" int foct_plan_sync(foct_plan* p, float* kernel_ms) {
  if (!p) return fail(FOCT_EINVAL, "NULL plan");
  CU(cudaSetDevice(p->device));
  CU(cudaStreamSynchronize(p->stream));
  if (p->cancelled) return fail(FOCT_ECANCELLED, "the run was cancelled");
  if (int rc = plan_extend(p)) return rc;
  if (kernel_ms) {
    *kernel_ms = 0.f;
    if (p->ran) { CU(cudaEventElapsedTime(kernel_ms, p->ev0, p->ev1)); *kernel_ms += p->ext_sample_ms; }
  }
  return 0;
}

extern "C" int foct_plan_timing(foct_plan* p, float* sample_ms, float* summary_ms, int* grid, int* block,
                                int* blocks_per_sm, int* regs, int* smem_bytes) {
  if (!p) return fail(FOCT_EINVAL, "NULL plan");
  CU(cudaSetDevice(p->device));
  if (p->ran) {
    CU(cudaEventSynchronize(p->ev2));
    if (int rc = plan_extend(p)) return rc;
    if (sample_ms) { CU(cudaEventElapsedTime(sample_ms, p->ev0, p->ev1)); *sample_ms += p->ext_sample_ms; }
    if (summary_ms) { CU(cudaEventElapsedTime(summary_ms, p->ev1, p->ev2)); *summary_ms += p->ext_summary_ms; }
  }
  if (grid) *grid = p->grid;
  if (block) *block = p->block;
  if (blocks_per_sm) *blocks_per_sm = p->blocks_per_sm;
  if (regs) *regs = p->regs;
  if (smem_bytes) *smem_bytes = (int)p->smem;
  return 0;
}

extern "C" int foct_plan_fetch(foct_plan* p, foct_result* R) {
  if (!p || !R) return fail(FOCT_EINVAL, "NULL plan/result");
  if (!p->ran) return fail(FOCT_EINVAL, "plan has not been run");
  CU(cudaSetDevice(p->device));
  CU(cudaStreamSynchronize(p->stream));
  if (p->cancelled) return fail(FOCT_ECANCELLED, "the run was cancelled");
  if (int rc = plan_extend(p)) return rc;
  const size_t pc = (size_t)p->n * p->cfg.chains;
  if (p->x_slots > 0 && !p->thinned && p->want_draws && (R->draws || R->sampler_params)) {
    // continued profiles: their post-warm-up rows become every (1 + n_extend)-th draw of the extension block
    int* d_next = nullptr;
    CU(pool_malloc(&d_next, (size_t)p->n * sizeof(int)));
    CU(cudaMemcpyAsync(d_next, p->n_extend.data(), (size_t)p->n * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    const int C = p->cfg.chains, off = p->cfg.save_warmup ? p->cfg.n_warmup : 0;
    const size_t row = (size_t)C * p->P_out;
    const dim3 grid(32, std::min(p->n, 65535));
    thin_blocks_kernel<<<grid, 256, 0, p->stream>>>(p->d_xdraws, (size_t)p->x_cap * row, p->d_draws, (size_t)p->n_saved * row,
                                                    (size_t)off * row, p->n_post, p->x_ext, row, p->d_slot_of, d_next, p->n);
    if (p->d_xsparams)
      thin_blocks_kernel<<<grid, 256, 0, p->stream>>>(p->d_xsparams, (size_t)p->x_cap * C * 6, p->d_sparams, (size_t)p->n_saved * C * 6,
                                                      (size_t)off * C * 6, p->n_post, p->x_ext, (size_t)C * 6, p->d_slot_of, d_next, p->n);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    pool_free(d_next);
    if (e != cudaSuccess) return fail(FOCT_ECUDA, "thinning of the continued draws failed: %s", cudaGetErrorString(e));
    p->thinned = true;
  }
  if (R->draws && p->n_saved > 0) {
    if (!p->want_draws) return fail(FOCT_EINVAL, "plan was created without draws");
    CU(cudaMemcpy(R->draws, p->d_draws, pc * p->n_saved * p->P_out * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (R->sampler_params && p->n_saved > 0) {
    if (!p->want_draws) return fail(FOCT_EINVAL, "plan was created without draws");
    CU(cudaMemcpy(R->sampler_params, p->d_sparams, pc * p->n_saved * 6 * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (R->summary) {
    if (!p->want_summary) return fail(FOCT_EINVAL, "plan was created without summary");
    const size_t ns = (size_t)p->n * p->P_out * FOCT_N_SUMMARY_COLS;
    if (p->n_post >= 2) CU(cudaMemcpy(R->summary, p->d_summary, ns * sizeof(double), cudaMemcpyDeviceToHost));
    else for (size_t i = 0; i < ns; ++i) R->summary[i] = std::nan("");  // fewer than two post-warm-up draws: nothing to summarise
  }
  if (R->stepsize) CU(cudaMemcpy(R->stepsize, p->d_stepsize, pc * sizeof(double), cudaMemcpyDeviceToHost));
  if (R->inv_metric) CU(cudaMemcpy(R->inv_metric, p->d_invm, pc * p->D * sizeof(double), cudaMemcpyDeviceToHost));
  if (R->n_leapfrog) CU(cudaMemcpy(R->n_leapfrog, p->d_nleap, pc * 2 * sizeof(double), cudaMemcpyDeviceToHost));
  if (R->n_divergent) CU(cudaMemcpy(R->n_divergent, p->d_ndiv, pc * sizeof(double), cudaMemcpyDeviceToHost));
  if (R->last_q) CU(cudaMemcpy(R->last_q, p->d_lastq, pc * p->D * sizeof(double), cudaMemcpyDeviceToHost));
  if (R->n_extend) for (int j = 0; j < p->n; ++j) R->n_extend[j] = p->n_extend.empty() ? 0 : p->n_extend[j];
  return 0;
}

extern "C" int foct_plan_query(foct_plan* p, int* done, double* fraction) {
  if (!p) return fail(FOCT_EINVAL, "NULL plan");
  CU(cudaSetDevice(p->device));
  const cudaError_t q = p->ran ? cudaStreamQuery(p->stream) : cudaSuccess;
  if (q != cudaSuccess && q != cudaErrorNotReady) return fail(FOCT_ECUDA, "cudaStreamQuery failed: %s", cudaGetErrorString(q));
  if (done) *done = q == cudaSuccess;
  if (fraction) {
    *fraction = 0.0;
    if (p->ran) {
      CU(cudaMemcpyAsync(p->h_progress, p->d_progress, sizeof(unsigned long long), cudaMemcpyDeviceToHost, p->side));
      CU(cudaStreamSynchronize(p->side));
      *fraction = (double)*p->h_progress / ((double)p->n * p->cfg.chains * p->cfg.n_iter);
    }
  }
  return 0;
}

extern "C" int foct_plan_cancel(foct_plan* p) {
  if (!p) return fail(FOCT_EINVAL, "NULL plan");
  if (!p->ran) return 0;
  CU(cudaSetDevice(p->device));
  static const int one = 1;
  CU(cudaMemcpyAsync(p->d_cancel, &one, sizeof(int), cudaMemcpyHostToDevice, p->side));
  CU(cudaStreamSynchronize(p->side));
  p->cancelled = true;
  return 0;
}

extern "C" int foct_plan_launches(foct_plan* p) { return p ? p->launches : 0; }
extern "C" void foct_plan_destroy(foct_plan* p) { plan_free(p); }

// ------------------------------------------------------------------ ABI: one-shot sampling, sharded over devices
// What foct_sample_cb's polling thread shares with the per-device workers: the workers publish the chain-iterations their
// current plan has completed, the caller's thread turns that into a fraction and may raise `cancel`.
struct SampleJob {
  std::atomic<long long> finished_iters{0};   // chain-iterations of chunks that are complete
  std::atomic<long long> running_iters[64];   // ... of the chunk each worker is running now
  std::atomic<int> extending{0};              // workers inside the rhat_target rounds
  std::atomic<bool> cancel{false};
  SampleJob() { for (auto& a : running_iters) a.store(0); }
};

static int sample_chunk(int device, int kind, const foct_problem* P, int first, int n, const foct_model_spec* spec,
                        const foct_sampler_cfg* cfg, foct_result* R, int D, int P_out, SampleJob* job, int worker) {
  const int C = cfg->chains;
  const int n_saved = cfg->save_warmup ? cfg->n_iter : cfg->n_iter - cfg->n_warmup;
  foct_plan* p = nullptr;
  const double* init_slice = cfg->init_mode == 2 && cfg->init ? cfg->init + (size_t)first * C * D : nullptr;
  const double* invm_slice = cfg->inv_metric_init ? cfg->inv_metric_init + (size_t)first * C * D : nullptr;
  const double* eps_slice = cfg->stepsize_init ? cfg->stepsize_init + (size_t)first * C : nullptr;
  Trace tr("foct_sample");
  if (job && job->cancel.load()) return fail(FOCT_ECANCELLED, "the run was cancelled");
  int rc = plan_create_on(device, kind, P + first, n, spec, cfg, init_slice, invm_slice, eps_slice,
                          R->draws || R->sampler_params, R->summary != nullptr, &p);
  if (rc) return rc;
  tr.mark("plan (alloc, pack, upload, setup kernel)");
  rc = foct_plan_run(p, cfg->seed);
  if (!rc && job) {
    // keep the caller's thread informed while the kernel runs; pass a cancellation on to the device
    for (;;) {
      int done = 0;
      double frac = 0.0;
      rc = foct_plan_query(p, &done, &frac);
      if (rc) break;
      job->running_iters[worker].store((long long)(frac * (double)n * C * cfg->n_iter));
      if (job->cancel.load() && !p->cancelled) { rc = foct_plan_cancel(p); if (rc) break; }
      if (done) break;
      std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    if (!rc && p->extend_pending && !p->cancelled) {
      job->extending.fetch_add(1);
      rc = foct_plan_sync(p, nullptr);
      job->extending.fetch_sub(1);
    }
  }
  if (!rc && tr.on) { foct_plan_sync(p, nullptr); tr.mark("kernels"); }
  if (!rc) {
    foct_result S = *R;
    const size_t pc0 = (size_t)first * C;
    if (S.draws) S.draws += pc0 * n_saved * P_out;
    if (S.sampler_params) S.sampler_params += pc0 * n_saved * 6;
    if (S.summary) S.summary += (size_t)first * P_out * FOCT_N_SUMMARY_COLS;
    if (S.stepsize) S.stepsize += pc0;
    if (S.inv_metric) S.inv_metric += pc0 * D;
    if (S.n_leapfrog) S.n_leapfrog += pc0 * 2;
    if (S.n_divergent) S.n_divergent += pc0;
    if (S.last_q) S.last_q += pc0 * D;
    if (S.n_extend) S.n_extend += first;
    rc = foct_plan_fetch(p, &S);
    tr.mark("fetch");
  }
  plan_free(p);
  if (job) {
    job->running_iters[worker].store(0);
    if (!rc) job->finished_iters.fetch_add((long long)n * C * cfg->n_iter);
  }
  tr.mark("free");
  return rc;
}

// One device's shard, processed in chunks so that the device-resident draws (needed by the summary kernel even
// when the caller wants summaries only: n x chains x n_saved x P_out doubles) stay below a memory budget —
// 1e5 profiles are 54 GB of draws; 1e6 would not fit 180 GB.  Chunks are whole multiples of the
// resident CTA count where possible, so chunking does not add partial waves.
static int sample_shard(int device, int kind, const foct_problem* P, int first, int n, const foct_model_spec* spec,
                        const foct_sampler_cfg* cfg, foct_result* R, int D, int P_out, SampleJob* job, int worker) {
  const int n_saved = cfg->save_warmup ? cfg->n_iter : cfg->n_iter - cfg->n_warmup;
  const double per_profile = (double)cfg->chains * std::max(1, n_saved) * (P_out + 6) * sizeof(double);
  double budget = 48.0 * 1024 * 1024 * 1024;
  if (const char* env = std::getenv("FOCT_DRAW_BUDGET_MB")) budget = std::atof(env) * 1024 * 1024;
  long long chunk = (long long)(budget / per_profile);
  {  // whole waves of resident CTAs where the chunk is many waves anyway (~4-6 CTAs per SM)
    cudaSetDevice(device);
    const long long wave = (long long)sm_count() * 4;
    if (chunk >= 4 * wave) chunk -= chunk % wave;
  }
  chunk = std::max<long long>(1, std::min<long long>(chunk, n));
  for (int off = 0; off < n; off += (int)chunk) {
    const int m = (int)std::min<long long>(chunk, n - off);
    if (int rc = sample_chunk(device, kind, P, first + off, m, spec, cfg, R, D, P_out, job, worker)) return rc;
  }
  return 0;
}

static int sample_impl(int kind, const foct_problem* P, int n, const foct_model_spec* spec, const foct_sampler_cfg* cfg,
                       foct_result* R, foct_progress_fn progress, void* user, int poll_ms) {
  if (int rc = check_device()) return rc;
  if (!P || !spec || !cfg || !R || n < 1) return fail(FOCT_EINVAL, "NULL argument or empty batch");
  if (int rc = validate_cfg(cfg)) return rc;
  int D, P_out;
  if (int rc = foct_dims(kind, P[0].Nn, &D, &P_out)) return rc;
  std::vector<int> devs;
  if (cfg->n_devices > 0 && cfg->devices) devs.assign(cfg->devices, cfg->devices + cfg->n_devices);
  else { int d = 0; CU(cudaGetDevice(&d)); devs.push_back(d); }
  const int ndev_avail = foct_device_count();
  for (int d : devs) if (d < 0 || d >= ndev_avail) return fail(FOCT_EINVAL, "device %d not present (%d devices)", d, ndev_avail);
  const int G = (int)std::min<size_t>(std::min<size_t>(devs.size(), (size_t)n), 64);
  if (G == 1 && !progress) return sample_shard(devs[0], kind, P, 0, n, spec, cfg, R, D, P_out, nullptr, 0);
  // independent contiguous shards, one host thread per GPU, no collective
  SampleJob job;
  std::vector<std::thread> th;
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  std::atomic<int> running{G};
  for (int gidx = 0; gidx < G; ++gidx) {
    const int first = (int)((long long)n * gidx / G), last = (int)((long long)n * (gidx + 1) / G);
    th.emplace_back([&, gidx, first, last]() {
      rcs[gidx] = sample_shard(devs[gidx], kind, P, first, last - first, spec, cfg, R, D, P_out, progress ? &job : nullptr, gidx);
      if (rcs[gidx]) errs[gidx] = g_err;
      running.fetch_sub(1);
    });
  }
  if (progress) {
    // the callback runs on THIS thread (R's API is single-threaded): poll, report, pass a cancellation on
    const double total = (double)n * cfg->chains * cfg->n_iter;
    const auto tick = std::chrono::milliseconds(poll_ms > 0 ? poll_ms : 100);
    auto next = std::chrono::steady_clock::now() + tick;
    while (running.load() > 0) {
      std::this_thread::sleep_for(std::chrono::milliseconds(2));
      if (std::chrono::steady_clock::now() < next) continue;
      next += tick;
      long long it = job.finished_iters.load();
      for (int g = 0; g < G; ++g) it += job.running_iters[g].load();
      const double f = std::min(1.0, (double)it / total);
      // chains advance at similar rates: warm-up is over for all of them once the mean iteration count is past it
      const char* phase = job.extending.load() > 0 ? "Extending" : (f * cfg->n_iter < cfg->n_warmup ? "Warmup" : "Sampling");
      if (!job.cancel.load() && progress(f, phase, user)) job.cancel.store(true);
    }
  }
  for (auto& t : th) t.join();
  if (job.cancel.load()) return fail(FOCT_ECANCELLED, "the run was cancelled by the progress callback");
  for (int gidx = 0; gidx < G; ++gidx)
    if (rcs[gidx]) { g_err = errs[gidx]; return rcs[gidx]; }
  if (progress) progress(1.0, "Sampling", user);
  return 0;
}

extern "C" int foct_sample(int kind, const foct_problem* P, int n, const foct_model_spec* spec,
                           const foct_sampler_cfg* cfg, foct_result* R) {
  return sample_impl(kind, P, n, spec, cfg, R, nullptr, nullptr, 0);
}
extern "C" int foct_sample_cb(int kind, const foct_problem* P, int n, const foct_model_spec* spec,
                              const foct_sampler_cfg* cfg, foct_result* R, foct_progress_fn progress, void* user, int poll_ms) {
  return sample_impl(kind, P, n, spec, cfg, R, progress, user, poll_ms);
}
extern "C" int foct_expgp_logp_grad(const foct_problem* P, int n, const foct_model_spec* spec, const double* q, int n_q,
                                    double* lp, double* grad, double* chi2) {
  return foct_logp_grad(FOCT_EXPGP, P, n, spec, q, n_q, lp, grad, chi2);
}

extern "C" int foct_expgp_sample(const foct_problem* P, int n, const foct_model_spec* spec, const foct_sampler_cfg* cfg,
                                 foct_result* R) {
  return foct_sample(FOCT_EXPGP, P, n, spec, cfg, R);
}
extern "C" int foct_monoexp_sample(const foct_problem* P, int n, const foct_model_spec* spec, const foct_sampler_cfg* cfg,
                                   foct_result* R) {
  return foct_sample(FOCT_MONOEXP, P, n, spec, cfg, R);
}


// ------------------------------------------------------------------ ABI: MonoExp MAP
extern "C" int foct_monoexp_map(const foct_problem* P, int n, const foct_model_spec* spec, const double* init,
                                double* theta, double* hessian, double* br, int* status) {
  if (int rc = check_device()) return rc;
  if (!theta) return fail(FOCT_EINVAL, "NULL theta output");
  int dev = 0, NN, npad;
  size_t stride;
  double* d_blobs;
  DevProblem* d_probs;
  CU(cudaGetDevice(&dev));
  if (int rc = build_device_batch(FOCT_MONOEXP, P, n, spec, dev, 0, &NN, &npad, &stride, &d_blobs, &d_probs)) return rc;
  double *d_init = nullptr, *d_th = nullptr, *d_H = nullptr, *d_br = nullptr;
  int* d_st = nullptr;
  int rc = 0;
  do {
#define CUB(call) if ((call) != cudaSuccess) { rc = fail(FOCT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(cudaGetLastError())); break; }
    CUB(pool_malloc(&d_th, (size_t)n * 3 * sizeof(double)));
    CUB(pool_malloc(&d_H, (size_t)n * 9 * sizeof(double)));
    CUB(pool_malloc(&d_br, (size_t)n * sizeof(double)));
    CUB(pool_malloc(&d_st, (size_t)n * sizeof(int)));
    if (init) {
      CUB(pool_malloc(&d_init, (size_t)n * 3 * sizeof(double)));
      CUB(cudaMemcpy(d_init, init, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice));
    }
    map_kernel<<<(n + 3) / 4, 128>>>(d_blobs, stride, 0, d_probs, n, spec->theta_prior, d_init, d_th, d_H, d_br, d_st);
    CUB(cudaGetLastError());
    CUB(cudaDeviceSynchronize());
    CUB(cudaMemcpy(theta, d_th, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (hessian) CUB(cudaMemcpy(hessian, d_H, (size_t)n * 9 * sizeof(double), cudaMemcpyDeviceToHost));
    if (br) CUB(cudaMemcpy(br, d_br, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    if (status) CUB(cudaMemcpy(status, d_st, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
#undef CUB
  } while (0);
  pool_free(d_init); pool_free(d_th); pool_free(d_H); pool_free(d_br); pool_free(d_st); pool_free(d_blobs); pool_free(d_probs);
  return rc;
}


// ------------------------------------------------------------------ ABI: ExpGP MAP (method = 'optim')
extern "C" int foct_expgp_map(const foct_problem* P, int n, const foct_model_spec* spec, const double* init,
                              double* par, double* hessian, int* status) {
  if (int rc = check_device()) return rc;
  if (!par) return fail(FOCT_EINVAL, "NULL par output");
  int dev = 0, NN, npad;
  size_t stride;
  double* d_blobs;
  DevProblem* d_probs;
  CU(cudaGetDevice(&dev));
  if (int rc = build_device_batch(FOCT_EXPGP, P, n, spec, dev, 0, &NN, &npad, &stride, &d_blobs, &d_probs)) return rc;
  const int D = NN + 5, P_out = NN + 7;
  double *d_init = nullptr, *d_par = nullptr, *d_H = nullptr;
  int* d_st = nullptr;
  int rc = 0;
  do {
#define CUB(call) if ((call) != cudaSuccess) { rc = fail(FOCT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(cudaGetLastError())); break; }
    CUB(pool_malloc(&d_par, (size_t)n * P_out * sizeof(double)));
    CUB(pool_malloc(&d_st, (size_t)n * sizeof(int)));
    if (hessian) CUB(pool_malloc(&d_H, (size_t)n * D * D * sizeof(double)));
    if (init) {
      CUB(pool_malloc(&d_init, (size_t)n * D * sizeof(double)));
      CUB(cudaMemcpy(d_init, init, (size_t)n * D * sizeof(double), cudaMemcpyHostToDevice));
    }
    MapParams K;
    K.blobs = d_blobs; K.blob_stride = stride; K.npad = npad; K.probs = d_probs; K.n_problems = n;
    K.spec = dev_spec(*spec); K.init = d_init; K.par = d_par; K.hessian = d_H; K.status = d_st; K.max_iter = 1000;
    const InstEntry* inst = inst_for(NN);
    CUB(inst->launch_map(spec->modulation, std::min(n, sm_count() * 4), stride * sizeof(double), 0, K));
    CUB(cudaDeviceSynchronize());
    CUB(cudaMemcpy(par, d_par, (size_t)n * P_out * sizeof(double), cudaMemcpyDeviceToHost));
    if (hessian) CUB(cudaMemcpy(hessian, d_H, (size_t)n * D * D * sizeof(double), cudaMemcpyDeviceToHost));
    if (status) CUB(cudaMemcpy(status, d_st, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
#undef CUB
  } while (0);
  pool_free(d_init); pool_free(d_par); pool_free(d_H); pool_free(d_st); pool_free(d_blobs); pool_free(d_probs);
  return rc;
}

// ------------------------------------------------------------------ ABI: method = 'vb' (ADVI, MODEL_SPEC §14)
extern "C" void foct_vb_cfg_default(foct_vb_cfg* c) {
  if (!c) return;
  std::memset(c, 0, sizeof(*c));
  c->iter = 10000; c->grad_samples = 1; c->elbo_samples = 100; c->eval_elbo = 100; c->output_samples = 1000;
  c->adapt_engaged = 1; c->adapt_iter = 50; c->eta = 1.0; c->tol_rel_obj = 0.01; c->seed = 1234ull; c->init_mode = 0;
  c->init = nullptr; c->omega0 = 0.0;
}

extern "C" int foct_vb(int kind, const foct_problem* P, int n, const foct_model_spec* spec, const foct_vb_cfg* cfg,
                       foct_vb_result* R) {
  if (int rc = check_device()) return rc;
  if (!cfg || !R || !R->mean || !R->mu || !R->omega) return fail(FOCT_EINVAL, "NULL cfg / result (mean, mu, omega are required)");
  if (cfg->iter < 1 || cfg->grad_samples < 1 || cfg->elbo_samples < 1 || cfg->eval_elbo < 1 || cfg->output_samples < 0 ||
      (cfg->adapt_engaged && cfg->adapt_iter < 1) || !(cfg->tol_rel_obj > 0.0) || (!cfg->adapt_engaged && !(cfg->eta > 0.0)))
    return fail(FOCT_EINVAL, "vb: iter, grad_samples, elbo_samples, eval_elbo, adapt_iter >= 1; eta, tol_rel_obj > 0");
  if (cfg->init_mode < 0 || cfg->init_mode > 2 || (cfg->init_mode == 2 && !cfg->init)) return fail(FOCT_EINVAL, "vb: bad init_mode / init");
  int dev = 0, NN, npad;
  size_t stride;
  double* d_blobs;
  DevProblem* d_probs;
  CU(cudaGetDevice(&dev));
  if (int rc = build_device_batch(kind, P, n, spec, dev, 0, &NN, &npad, &stride, &d_blobs, &d_probs)) return rc;
  const int D = kind == FOCT_EXPGP ? NN + 5 : 3, P_out = kind == FOCT_EXPGP ? NN + 7 : 5;
  const size_t nd = (size_t)n * cfg->output_samples * P_out;
  double *d_init = nullptr, *d_mean = nullptr, *d_draws = nullptr, *d_mu = nullptr, *d_om = nullptr, *d_elbo = nullptr, *d_eta = nullptr;
  int *d_it = nullptr, *d_st = nullptr;
  int rc = 0;
  do {
#define CUB(call) if ((call) != cudaSuccess) { rc = fail(FOCT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(cudaGetLastError())); break; }
    CUB(pool_malloc(&d_mean, (size_t)n * P_out * sizeof(double)));
    CUB(pool_malloc(&d_mu, (size_t)n * D * sizeof(double)));
    CUB(pool_malloc(&d_om, (size_t)n * D * sizeof(double)));
    CUB(pool_malloc(&d_elbo, (size_t)n * sizeof(double)));
    CUB(pool_malloc(&d_eta, (size_t)n * sizeof(double)));
    CUB(pool_malloc(&d_it, (size_t)n * sizeof(int)));
    CUB(pool_malloc(&d_st, (size_t)n * sizeof(int)));
    if (R->draws && nd) CUB(pool_malloc(&d_draws, nd * sizeof(double)));
    if (cfg->init_mode == 2) {
      CUB(pool_malloc(&d_init, (size_t)n * D * sizeof(double)));
      CUB(cudaMemcpy(d_init, cfg->init, (size_t)n * D * sizeof(double), cudaMemcpyHostToDevice));
    }
    VbParams K;
    K.blobs = d_blobs; K.blob_stride = stride; K.npad = npad; K.probs = d_probs; K.n_problems = n; K.spec = dev_spec(*spec);
    K.iter = cfg->iter; K.grad_samples = cfg->grad_samples; K.elbo_samples = cfg->elbo_samples; K.eval_elbo = cfg->eval_elbo;
    K.output_samples = cfg->output_samples; K.adapt_engaged = cfg->adapt_engaged; K.adapt_iter = cfg->adapt_iter;
    K.init_mode = cfg->init_mode; K.eta = cfg->eta; K.tol_rel_obj = cfg->tol_rel_obj; K.omega0 = cfg->omega0; K.seed = cfg->seed;
    K.init = d_init; K.mean = d_mean; K.draws = d_draws; K.mu = d_mu; K.omega = d_om; K.elbo = d_elbo; K.eta_out = d_eta;
    K.iters = d_it; K.status = d_st;
    const InstEntry* inst = inst_for(NN);
    CUB(inst->launch_vb(spec->modulation, std::min(n, sm_count() * 8), stride * sizeof(double), 0, K));
    CUB(cudaDeviceSynchronize());
    CUB(cudaMemcpy(R->mean, d_mean, (size_t)n * P_out * sizeof(double), cudaMemcpyDeviceToHost));
    CUB(cudaMemcpy(R->mu, d_mu, (size_t)n * D * sizeof(double), cudaMemcpyDeviceToHost));
    CUB(cudaMemcpy(R->omega, d_om, (size_t)n * D * sizeof(double), cudaMemcpyDeviceToHost));
    if (d_draws) CUB(cudaMemcpy(R->draws, d_draws, nd * sizeof(double), cudaMemcpyDeviceToHost));
    if (R->elbo) CUB(cudaMemcpy(R->elbo, d_elbo, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    if (R->eta) CUB(cudaMemcpy(R->eta, d_eta, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    if (R->iters) CUB(cudaMemcpy(R->iters, d_it, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    if (R->status) CUB(cudaMemcpy(R->status, d_st, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
#undef CUB
  } while (0);
  pool_free(d_init); pool_free(d_mean); pool_free(d_draws); pool_free(d_mu); pool_free(d_om); pool_free(d_elbo); pool_free(d_eta);
  pool_free(d_it); pool_free(d_st); pool_free(d_blobs); pool_free(d_probs);
  return rc;
}

// ------------------------------------------------------------------ ABI: fp64 peak
extern "C" int foct_fp64_peak(int device, double* tflops, double* sm_mhz) {
  if (int rc = check_device()) return rc;
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  const int grid = prop.multiProcessorCount * 8, block = 256, iters = 1 << 16;
  double* d_out = nullptr;
  CU(pool_malloc(&d_out, (size_t)grid * block * sizeof(double)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    CU(cudaEventRecord(e0));
    dfma_peak_kernel<<<grid, block>>>(d_out, iters, 0.999999, 1e-9);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); pool_free(d_out);
  const double flops = 2.0 * 8.0 * (double)iters * (double)grid * block;
  if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
  if (sm_mhz) *sm_mhz = prop.clockRate / 1000.0;
  return 0;
}
