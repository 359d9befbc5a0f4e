// synth.cuh -- device-side synthetic data and random tape: counter-based Philox4x32-10
// keyed by the GLOBAL (row, col), so 1/2/4/8-GPU row shardings see the same matrix.  Bit-exact twin of
// oracle/logistic.py (synth_X_block, synth_uniform_rows) -- integer arithmetic plus exact conversions only.
#pragma once
#include "common.cuh"

namespace gamc {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

__host__ __device__ inline double synth_x(uint64_t seed, uint64_t row, uint32_t col) {
  const Philox4 r = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), col, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
  const long long s = (long long)r.x + (long long)r.y + (long long)r.z + (long long)r.w - (1LL << 33);
  return (double)s * (1.7320508075688772 / 4294967296.0);
}

__host__ __device__ inline double philox_u53(uint32_t a, uint32_t b) {
  const uint64_t bits = ((uint64_t)a << 21) | ((uint64_t)b >> 11);
  return (double)bits * (1.0 / 9007199254740992.0);
}

// X[i, j] for local rows i in [0, nrows), global row = row0 + i; column-major with leading dimension ld.
static __global__ void k_synth_X(double* __restrict__ X, long long ld, long long row0, long long nrows, int p, uint64_t seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  for (int j = blockIdx.y; j < p; j += gridDim.y) X[i + (long long)j * ld] = synth_x(seed, (uint64_t)(row0 + i), (uint32_t)j);
}

// y[i] = 1{u_i < sigma(x_i . theta_true)}; eta accumulated sequentially in j WITHOUT fma contraction
// (matches the NumPy loop in oracle/logistic.py:synth_logistic).
static __global__ void k_synth_y(const double* __restrict__ X, long long ld, long long row0, long long nrows, int p,
                          const double* __restrict__ theta_true, uint64_t seed, double* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  double eta = 0.0;
  for (int j = 0; j < p; ++j) eta = __dadd_rn(eta, __dmul_rn(X[i + (long long)j * ld], theta_true[j]));
  const uint64_t row = (uint64_t)(row0 + i);
  const Philox4 r = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), 0xFFFFFFFFu, 1u, (uint32_t)seed, (uint32_t)(seed >> 32));
  const double u = philox_u53(r.x, r.y);
  const double e = exp(-fabs(eta));
  const double sig = (eta >= 0.0) ? 1.0 / (1.0 + e) : e / (1.0 + e);
  y[i] = (u < sig) ? 1.0 : 0.0;
}

// Device-generated random tape (throughput mode): slot t -> u_sched, u_mix, u_acc in (0,1), z ~ N(0,1)
// by Box-Muller on Philox outputs.  Counter (t_lo, t_hi, j, stream).
static __global__ void k_tape_generate(long long n, int p, uint64_t seed, long long t0, double* __restrict__ usched,
                                double* __restrict__ umix, double* __restrict__ uacc, double* __restrict__ z) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int pairs = (p + 1) / 2;
  const long long total = n * (long long)(pairs + 1);
  if (idx >= total) return;
  const long long t = idx / (pairs + 1);
  const int j = (int)(idx % (pairs + 1));
  const uint64_t gt = (uint64_t)(t0 + t);
  if (j == pairs) {
    const Philox4 r = philox4x32_10((uint32_t)gt, (uint32_t)(gt >> 32), 0xFFFFFFFEu, 2u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const Philox4 r2 = philox4x32_10((uint32_t)gt, (uint32_t)(gt >> 32), 0xFFFFFFFDu, 2u, (uint32_t)seed, (uint32_t)(seed >> 32));
    usched[t] = philox_u53(r.x, r.y);
    umix[t] = philox_u53(r.z, r.w);
    double ua = philox_u53(r2.x, r2.y);
    uacc[t] = ua > 0.0 ? ua : 1.0 / 9007199254740992.0;
    return;
  }
  const Philox4 r = philox4x32_10((uint32_t)gt, (uint32_t)(gt >> 32), (uint32_t)j, 3u, (uint32_t)seed, (uint32_t)(seed >> 32));
  double u1 = philox_u53(r.x, r.y), u2 = philox_u53(r.z, r.w);
  if (u1 <= 0.0) u1 = 1.0 / 9007199254740992.0;
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  z[t * p + 2 * j] = rad * cs;
  if (2 * j + 1 < p) z[t * p + 2 * j + 1] = rad * sn;
}

}  // namespace gamc
