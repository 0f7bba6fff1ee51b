// foct_launch.h — host-callable launchers, one translation unit per control-point count NN
// (foct_inst.cu compiled with -DFOCT_INST_NN=<NN>); foct_lib.cu dispatches through foct_inst_table().
#pragma once
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "foct_nuts.cuh"

namespace foct {

struct LogpParams {
  const double* blobs;
  size_t blob_stride;
  int npad;
  const DevProblem* probs;
  int n_problems;
  DevSpec spec;
  const double* q;  // [n_problems][n_q][D]
  int n_q;
  double* lp;    // [n_problems][n_q]
  double* grad;  // [n_problems][n_q][D]
  double* chi2;  // [n_problems][n_q] or nullptr
  int width;     // evaluation layout: 16 = half-warps on a staged blob (D <= 16), 17 = half-warps in nuts2w_kernel's shared-basis layout, else full warps
};

struct MapParams {
  const double* blobs;
  size_t blob_stride;
  int npad;
  const DevProblem* probs;
  int n_problems;
  DevSpec spec;
  const double* init;  // [n_problems][D] unconstrained or nullptr
  double* par;         // [n_problems][P_out] constrained optimum + br + lp (no Jacobian)
  double* hessian;     // [n_problems][D][D] of lp without Jacobian, or nullptr
  int* status;         // [n_problems] 0 converged, 1 iteration limit, 2 line search failed
  int max_iter;
};

struct VbParams {  // method = 'vb' (MODEL_SPEC §14)
  const double* blobs;
  size_t blob_stride;
  int npad;
  const DevProblem* probs;
  int n_problems;
  DevSpec spec;
  int iter, grad_samples, elbo_samples, eval_elbo, output_samples, adapt_engaged, adapt_iter, init_mode;
  double eta, tol_rel_obj, omega0;
  unsigned long long seed;
  const double* init;  // [n_problems][D] or nullptr
  double *mean, *draws, *mu, *omega, *elbo, *eta_out;
  int *iters, *status;
};

struct InstEntry {
  int NN;  // 0 = mono-exponential
  cudaError_t (*launch_nuts)(int mod, int grid, int block, size_t smem, cudaStream_t st, const SamplerParams& K);
  cudaError_t (*launch_logp)(int mod, int grid, int block, size_t smem, cudaStream_t st, const LogpParams& K);
  // launch geometry for a batch of n_items profiles on a device with n_sm SMs; n_items = 0: the one-chain-per-warp kernel
  cudaError_t (*nuts_occupancy)(int mod, int chains, long long n_items, int n_sm, size_t smem_full, size_t smem_rows,
                                int* shared_basis, size_t* smem, int* block, int* blocks_per_sm, int* cta_chains, int* regs,
                                size_t* slice_bytes, int* pair_kernel, int* warp_units);
  cudaError_t (*launch_map)(int mod, int grid, size_t smem, cudaStream_t st, const MapParams& K);
  cudaError_t (*launch_vb)(int mod, int grid, size_t smem, cudaStream_t st, const VbParams& K);
};

// shared by the host translation units (defined in foct_lib.cu)
int fail(int code, const char* fmt, ...);  // records the thread-local message behind foct_last_error(), returns code
int check_device();
int sm_count();  // SMs of the current device
// cached device / pinned-host buffers (foct_lib.cu): freed blocks are kept for later calls, see foct_release_cache()
cudaError_t device_cache_alloc(void** p, size_t bytes);
void device_cache_free(void* p);
cudaError_t pinned_cache_alloc(void** p, size_t bytes);
void pinned_cache_free(void* p);

// FOCT_TRACE=1: host-phase wall times of the host entry points on stderr
struct Trace {
  bool on;
  const char* what;
  std::chrono::steady_clock::time_point t0;
  explicit Trace(const char* w) : on(std::getenv("FOCT_TRACE") != nullptr), what(w), t0(std::chrono::steady_clock::now()) {}
  void mark(const char* phase) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[foct trace] %s: %s %.3f ms\n", what, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

#define FOCT_DECL_INST(NN) const InstEntry* foct_inst_##NN();

}  // namespace foct
