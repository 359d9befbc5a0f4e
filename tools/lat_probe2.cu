// Single-warp latency of the forward-substitution step used by the AM update kernel.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void fwd(double* out, long long* cyc, const double* Lg, int reps) {
  const int lane = threadIdx.x;
  double lrow[32];
#pragma unroll
  for (int jj = 0; jj < 32; ++jj) lrow[jj] = (jj <= lane) ? Lg[jj * 32 + lane] : 0.0;
  double rdv = 1.0 / lrow[0 + 0 * (lane)]; 
#pragma unroll
  for (int jj = 0; jj < 32; ++jj) if (jj == lane) rdv = 1.0 / lrow[jj];
  double res = 1.0 + lane, wsum = 0;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
      const double wj = __shfl_sync(0xffffffffu, res * rdv, jj);
      if (lane == jj) wsum += wj;
      res = fma(-lrow[jj], wj, res);
    }
    res += 1.0;
  }
  long long t1 = clock64();
  out[lane] = res + wsum;
  if (lane == 0) *cyc = t1 - t0;
}
__global__ void shflchain(double* out, long long* cyc, int reps) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) x = __shfl_sync(0xffffffffu, x, (jj * 7 + 3) & 31) + 1.0;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *out, *L; long long* cyc; cudaMalloc(&out, 256); cudaMalloc(&cyc, 8); cudaMalloc(&L, 1024 * 8);
  double h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 0.01 * (i % 7); for (int i = 0; i < 32; ++i) h[i * 32 + i] = 2.0;
  cudaMemcpy(L, h, sizeof(h), cudaMemcpyHostToDevice);
  long long c;
  fwd<<<1, 32>>>(out, cyc, L, 100); cudaDeviceSynchronize();
  fwd<<<1, 32>>>(out, cyc, L, 100); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("forward-substitution step (DMUL -> SHFL.64 -> DFMA): %.1f cycles per column\n", (double)c / 3200);
  shflchain<<<1, 32>>>(out, cyc, 100); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("SHFL.64 + DADD chain: %.1f cycles per step\n", (double)c / 3200);
  return 0;
}
