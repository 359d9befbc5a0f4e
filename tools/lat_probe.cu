// Single-warp dependent-chain latencies on B200 (calibration for the single-CTA linear-algebra kernels).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void chain(double* out, long long* cyc, double a, double b) {
  double x = a + threadIdx.x * 1e-9;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, b, a);
    if (OP == 1) x = x * b;
    if (OP == 2) x = __shfl_sync(0xffffffffu, x, (i & 31));
    if (OP == 3) x = rsqrt(x) + a;
    if (OP == 4) x = 1.0 / x + a;
    if (OP == 5) x = sqrt(x) + a;
    if (OP == 6) x = log(x) + 2.0;
    if (OP == 7) { float y = (float)x; y = fmaf(y, 1.0001f, 0.5f); x = y; }
    if (OP == 8) x = __shfl_sync(0xffffffffu, x * b, (i & 31));
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, double a, double b) {
  double* out; long long* cyc; cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 8);
  chain<OP><<<1, 32>>>(out, cyc, a, b); cudaDeviceSynchronize();
  chain<OP><<<1, 32>>>(out, cyc, a, b); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %7.1f cycles per dependent op\n", name, (double)h / N);
}
int main() {
  run<0>("DFMA", 1.0, 0.999);
  run<1>("DMUL", 1.0, 1.0000001);
  run<2>("SHFL.64", 1.0, 1.0);
  run<3>("rsqrt(double)+DADD", 1.5, 1.0);
  run<4>("1/x (double)+DADD", 1.5, 1.0);
  run<5>("sqrt(double)+DADD", 1.5, 1.0);
  run<6>("log(double)+DADD", 1.5, 1.0);
  run<7>("cvt f64->f32, FFMA, cvt back", 1.5, 1.0);
  run<8>("DMUL+SHFL.64", 1.0, 1.0000001);
  return 0;
}
