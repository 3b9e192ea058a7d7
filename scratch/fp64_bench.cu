// Microbenchmark: per-SM throughput of the fp64 operations the conformer kernels lean on (B200, sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-3, b = 1.000001, c = 0.999999, d = a + 1.0, e = a + 2.0, f = a + 3.0;
  __shared__ double sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = 0.0;
  __syncthreads();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) { a = fma(a, b, c); d = fma(d, b, c); e = fma(e, b, c); f = fma(f, b, c); }
    if (OP == 1) { a = a / b + c; d = d / b + c; e = e / b + c; f = f / b + c; }
    if (OP == 2) { a = sqrt(a) + c; d = sqrt(d) + c; e = sqrt(e) + c; f = sqrt(f) + c; }
    if (OP == 3) { a = __drcp_rn(a) + c; d = __drcp_rn(d) + c; e = __drcp_rn(e) + c; f = __drcp_rn(f) + c; }
    if (OP == 4) { a = rsqrt(a) + c; d = rsqrt(d) + c; e = rsqrt(e) + c; f = rsqrt(f) + c; }
    if (OP == 5) { atomicAdd(&sm[(threadIdx.x * 7 + i) & 63], a); atomicAdd(&sm[(threadIdx.x * 13 + i) & 63], d); atomicAdd(&sm[(threadIdx.x * 3 + i) & 63], e); atomicAdd(&sm[(threadIdx.x * 5 + i) & 63], f); }
    if (OP == 6) { a = (double)(__double2int_rn(a) + i) * b; d = (double)(__double2int_rn(d) + i) * b; e = (double)(__double2int_rn(e) + i) * b; f = (double)(__double2int_rn(f) + i) * b; }
    if (OP == 7) { a = acos(fmin(1.0, fmax(-1.0, a * 0.1))) + c; d = acos(fmin(1.0, fmax(-1.0, d * 0.1))) + c; e = cos(e) + c; f = cos(f) + c; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f + sm[threadIdx.x & 63];
}
template <int OP>
void run(const char* name) {
  double* out; cudaMalloc(&out, 148 * 4 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<OP><<<148 * 4, 256>>>(out, 100, 1.5);
  cudaEventRecord(e0);
  k<OP><<<148 * 4, 256>>>(out, iters, 1.5);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = 148.0 * 4 * 256 * iters * 4;
  printf("%-28s %8.3f ms  %8.2f Gop/s  %6.2f op/clk/SM (at 1.965 GHz)\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<0>("DFMA"); run<1>("ddiv (+DADD)"); run<2>("dsqrt (+DADD)"); run<3>("drcp (+DADD)"); run<4>("drsqrt (+DADD)");
  run<5>("smem atomicAdd(double)"); run<6>("d2i + i2d + DMUL"); run<7>("acos/cos (+DADD)");
  return 0;
}
