// Issue cost (cycles per warp-instruction per SMSP at 4 warps/SMSP) of the instruction kinds the sweep uses.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters, const double* in, long long* cyc) {
  __shared__ __align__(16) double sm[32 * 32];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
  __syncthreads();
  double a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = in[i] + threadIdx.x; b[i] = in[8 + i]; c[i] = in[16 + i]; }
  const double* p = sm + (threadIdx.x & 31);
  const double2* p2 = reinterpret_cast<const double2*>(sm) + (threadIdx.x & 31);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fma(b[i], c[i], a[i]);                       // DFMA 3 regs
      if (MODE == 1) a[i] = fma(a[i], b[i], 1.25);                       // DFMA 2 regs + immediate
      if (MODE == 2) a[i] = a[i] * b[i];                                 // DMUL
      if (MODE == 3) a[i] = a[i] + b[i];                                 // DADD
      if (MODE == 4) { a[i] = fma(b[i], c[i], a[i]); b[i] = p[i * 32]; }             // DFMA + LDS.64
      if (MODE == 5) { a[i] = fma(b[i], c[i], a[i]); if (!(i & 1)) { double2 v = p2[i * 16]; b[i] = v.x; b[i + 1] = v.y; } }  // 2 DFMA + LDS.128
      if (MODE == 6) { a[i] = fma(b[i], c[i], a[i]); b[i] = __shfl_xor_sync(0xffffffffu, b[i], 1); }  // DFMA + 2 SHFL
      if (MODE == 7) a[i] = fma(a[i], a[i], a[i]);                       // DFMA 1 reg
      if (MODE == 8) a[i] = fma(a[i], b[0], c[i]);                       // DFMA 3 regs, one shared across instrs
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
  const char* names[] = {"DFMA 3 regs", "DFMA 2 regs + imm", "DMUL", "DADD", "DFMA + LDS.64", "2 DFMA + LDS.128", "DFMA + SHFL64", "DFMA 1 reg", "DFMA 3 regs (1 shared)"};
#define RUN(MODE) { k<MODE><<<148, 512>>>(d, iters, in, c); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
  printf("%-24s %.2f cycles per loop slot per SMSP\n", names[MODE], (double)h / (iters * 8.0 * 4.0)); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
  return 0;
}
