// DFMA throughput with 1 vs 3 distinct register operands per instruction, and mixed with LDS / SHFL (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters, const double* in, long long* cyc) {
  __shared__ double sm[32 * 16];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
  __syncthreads();
  double a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = in[i] + threadIdx.x; b[i] = in[8 + i]; c[i] = in[16 + i]; }
  const double* p = sm + (threadIdx.x & 31);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fma(a[i], b[0], c[0]);            // 1 varying operand
      if (MODE == 1) a[i] = fma(b[i], c[i], a[i]);            // 3 distinct register operands
      if (MODE == 2) { a[i] = fma(b[i], c[i], a[i]); b[i] = p[(i & 7) * 32]; }  // + one LDS.64 per DFMA
      if (MODE == 3) { a[i] = fma(b[i], c[i], a[i]); if ((i & 3) == 0) b[i] = p[(i & 7) * 32]; }  // LDS per 4 DFMA
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *d, *in; long long* c; cudaMalloc(&d, 1 << 24); cudaMalloc(&in, 256); cudaMalloc(&c, 8);
  double h_in[24]; for (int i = 0; i < 24; ++i) h_in[i] = 0.5 + 0.01 * i; cudaMemcpy(in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
  long long h; const int iters = 4096;
#define RUN(MODE, BLK) { k<MODE><<<148, BLK>>>(d, iters, in, c); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyHostToDevice == 0 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToHost); \
  printf("mode %d block %d: %.2f cycles per DFMA per SMSP\n", MODE, BLK, (double)h / (iters * 8.0 * (BLK / 128.0))); }
  RUN(0, 512) RUN(1, 512) RUN(2, 512) RUN(3, 512) RUN(0, 384) RUN(1, 384) RUN(2, 384) RUN(3, 384)
  return 0;
}
