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
      const Eval ev = warp_logp_grad<NN, MOD>(smem, smem + K.blob_stride + (size_t)warp * K.npad, s_prob, K.spec, qd, lane);
      if (lane < D) K.grad[r * D + lane] = ev.g;
      if (lane == 0) {
        K.lp[r] = ev.lp;
        if (K.chi2) K.chi2[r] = ev.chi2;
      }
    }
    __syncthreads();
  }
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
                              &nuts_occupancy<FOCT_INST_NN>};
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
