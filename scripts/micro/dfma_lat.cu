// Microbenchmark: dependent-issue latency and throughput of DFMA / LDS.64 / SHFL on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b, long long* cyc) {
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x + i;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl(double* out, int iters, long long* cyc) {
  double v = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) v += __shfl_xor_sync(0xffffffffu, v, 1 + (r & 15));
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lds(double* out, int iters, long long* cyc) {
  __shared__ double s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 7) % 1024;
  __syncthreads();
  int idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) idx = (int)s[idx];
  }
  long long t1 = clock64();
  out[threadIdx.x] = idx;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 1 << 24); cudaMalloc(&c, 8);
  long long h;
  const int iters = 4096;
#define RUN(ILP, BLK, GRD) { k_dfma<ILP><<<GRD, BLK>>>(d, iters, 0.999, 1e-9, c); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
  printf("DFMA ilp=%d block=%d grid=%d: %.2f cycles per dependent DFMA step (per warp), %.2f cycles/DFMA issued per SMSP\n", ILP, BLK, GRD, (double)h / (iters * 8.0), (double)h / (iters * 8.0 * ILP * (BLK >= 128 ? BLK / 128.0 : 1.0))); }
  RUN(1, 32, 1) RUN(2, 32, 1) RUN(4, 32, 1) RUN(8, 32, 1)
  RUN(1, 128, 1) RUN(1, 256, 1) RUN(1, 384, 1) RUN(1, 512, 1) RUN(2, 384, 1) RUN(2, 512, 1) RUN(4, 512, 1)
  k_shfl<<<1, 32>>>(d, iters, c); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("SHFL64+DADD dependent: %.2f cycles per step\n", (double)h / (iters * 8.0));
  k_lds<<<1, 32>>>(d, iters, c); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("LDS.64 dependent (incl. cvt): %.2f cycles per step\n", (double)h / (iters * 8.0));
  return 0;
}
