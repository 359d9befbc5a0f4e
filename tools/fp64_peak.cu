// FP64 peak probes for B200 (sm_100a): register-only DMMA (mma.sync m8n8k4 f64) and DFMA loops.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
// These are yard-sticks for the metric-pass roofline, not product code.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int NT>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters, double seed) {
  double c[NT][2];
  double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
#pragma unroll
  for (int t = 0; t < NT; ++t) { c[t][0] = 0.0; c[t][1] = 0.0; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < NT; ++t) s += c[t][0] + c[t][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters, double seed) {
  double c[NT];
  double a = seed + threadIdx.x * 1e-9, b = 1.0 - 1e-12;
#pragma unroll
  for (int t = 0; t < NT; ++t) c[t] = t;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < NT; ++t) c[t] = fma(c[t], b, a);
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < NT; ++t) s += c[t];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s sms %d clock %d kHz\n", prop.name, sms, prop.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256 * 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 20000;
  for (int bps = 1; bps <= 8; bps *= 2) {
    int grid = sms * bps;
    // DMMA, 16 independent accumulators per warp
    dmma_loop<16><<<grid, 256>>>(out, 100, 1.0); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
      CK(cudaEventRecord(e0)); dmma_loop<16><<<grid, 256>>>(out, iters, 1.0); CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double flops = (double)grid * 8 /*warps*/ * iters * 16.0 * 512.0;
    printf("DMMA m8n8k4  blocks/SM %d (%2d warps/SM): %8.3f ms  %7.2f TFLOP/s\n", bps, bps * 8, best, flops / best * 1e-9);
    dfma_loop<16><<<grid, 256>>>(out, 100, 1.0); CK(cudaDeviceSynchronize());
    best = 1e30f;
    for (int r = 0; r < 3; ++r) {
      CK(cudaEventRecord(e0)); dfma_loop<16><<<grid, 256>>>(out, iters, 1.0); CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    flops = (double)grid * 256 * iters * 16.0 * 2.0;
    printf("DFMA         blocks/SM %d (%2d warps/SM): %8.3f ms  %7.2f TFLOP/s\n", bps, bps * 8, best, flops / best * 1e-9);
  }
  // 12 warps per CTA, 1 CTA per SM (the metric kernel's shape)
  {
    int grid = sms;
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
      CK(cudaEventRecord(e0)); dmma_loop<16><<<grid, 384>>>(out, iters, 1.0); CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double flops = (double)grid * 12 * iters * 16.0 * 512.0;
    printf("DMMA m8n8k4  12 warps/SM: %8.3f ms  %7.2f TFLOP/s\n", best, flops / best * 1e-9);
  }
  return 0;
}
