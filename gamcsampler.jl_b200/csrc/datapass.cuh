// datapass.cuh -- Kernel A: one HBM pass over the row-sharded design matrix.
//
// Computes, for the logistic target of examples/logistic_regression/src/GAMC/simulations.jl:25-37,
//   loglik  = sum_i y_i eta_i - log(1+exp(eta_i)),  eta = X theta            (ploglikelihood :25-28)
//   grad    = X'(y - sigma(eta))                                             (pgradlogtarget :32, likelihood part)
//   w_i     = sigma(eta_i)(1 - sigma(eta_i))                                 (ptensorlogtarget :35-36, the weights)
// in ONE read of X (the reference traverses X once per callback, three times per evaluation).
//
// Layout: X is column-major N x p.  A "unit" is a TMA box of RB rows x CU columns (RB = 32*RPL rows,
// CU = 64/RPL columns, 16 KB) staged in a ring of shared-memory slots by a producer warp; a "row block"
// is NU = ceil(p/CU... see below) consecutive units covering all columns of RB rows.  Eight consumer
// warps split the columns of every unit (8 columns... CU/8 each); lane l owns rows l*RPL..l*RPL+RPL-1, so every shared load
// is a conflict-free 256/512-byte row segment of one column.
//   phase 1: eta partials per warp -> cross-warp sum through shared memory -> sigma, softplus per row
//   phase 2 (GRAD): the same units, still resident, are re-read: gacc[col] += x * (y - sigma)
// Per-thread gradient accumulators live in registers across ALL row blocks of the CTA and are reduced
// over lanes once at the end; per-CTA partials are written in a fixed order (deterministic).
// HBM-bound: algorithmic bytes = 8 N p + 8 N (+ 8 N for the weights when GRAD).
#pragma once
#include "common.cuh"

namespace gamc {

constexpr int DP_CONSUMER_WARPS = 8;
constexpr int DP_THREADS = (DP_CONSUMER_WARPS + 1) * 32;
constexpr int DP_UNIT_DOUBLES = 2048;          // 16 KB
constexpr int DP_NSLOT = 13;                   // ring depth (208 KB)

struct DataPassSmem {
  alignas(128) double ring[DP_NSLOT][DP_UNIT_DOUBLES];
  double theta[512];
  double red[2][DP_CONSUMER_WARPS][64];
  alignas(8) uint64_t full_bar[DP_NSLOT];
  alignas(8) uint64_t empty_bar[DP_NSLOT];
};

// NU  = units per row block (columns padded to NU*CU, zero-filled by TMA out-of-bounds handling)
// RPL = rows per lane (1: RB = 32 rows, CU = 64 columns;  2: RB = 64 rows, CU = 32 columns)
template <int NU, int RPL, bool GRAD>
__global__ void __launch_bounds__(DP_THREADS, 1)
k_datapass(const __grid_constant__ CUtensorMap mapX, const double* __restrict__ y,
           const double* __restrict__ theta, int p, long long N, double* __restrict__ w_out,
           double* __restrict__ part_ll, double* __restrict__ part_g, int part_g_stride, const int* skip_flag) {
  if (skip_flag && *reinterpret_cast<const volatile int*>(skip_flag)) return;   // gamc_config.reuse_tensor: nothing to recompute
  constexpr int RB = 32 * RPL;
  constexpr int CU = 64 / RPL;
  constexpr int CPW = CU / DP_CONSUMER_WARPS;  // columns per warp per unit
  extern __shared__ __align__(128) unsigned char smem_raw[];
  DataPassSmem& S = *reinterpret_cast<DataPassSmem*>(smem_raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long nrb = (N + RB - 1) / RB;

  for (int c = tid; c < NU * CU; c += DP_THREADS) S.theta[c] = (c < p) ? theta[c] : 0.0;
  if (tid == 0) {
    for (int s = 0; s < DP_NSLOT; ++s) { mbar_init(&S.full_bar[s], 1); mbar_init(&S.empty_bar[s], DP_CONSUMER_WARPS); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == DP_CONSUMER_WARPS) {
    // ===== producer warp: one elected lane issues the TMA boxes =====
    if (lane == 0) {
      tma_prefetch_desc(&mapX);
      unsigned q = 0;
      for (long long rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
#pragma unroll 1
        for (int u = 0; u < NU; ++u, ++q) {
          const unsigned slot = q % DP_NSLOT, round = q / DP_NSLOT;
          mbar_wait(&S.empty_bar[slot], (round & 1) ^ 1);
          mbar_arrive_expect_tx(&S.full_bar[slot], DP_UNIT_DOUBLES * 8);
          tma_load_2d(S.ring[slot], &mapX, &S.full_bar[slot], (int)(rb * RB), u * CU);
        }
      }
    }
    return;
  }

  // ===== consumer warps =====
  double gacc[GRAD ? NU * CPW : 1];
#pragma unroll
  for (int i = 0; i < (GRAD ? NU * CPW : 1); ++i) gacc[i] = 0.0;
  double ll = 0.0;
  unsigned q = 0;
  int it = 0;
  for (long long rb = blockIdx.x; rb < nrb; rb += gridDim.x, ++it) {
    const long long row0 = rb * RB + (long long)lane * RPL;
    double yv[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) yv[r] = (row0 + r < N) ? __ldg(&y[row0 + r]) : 0.0;

    // ---- phase 1: eta partial over this warp's columns
    double eta[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) eta[r] = 0.0;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const unsigned qq = q + u, slot = qq % DP_NSLOT, round = qq / DP_NSLOT;
      mbar_wait(&S.full_bar[slot], round & 1);
      const double* base = &S.ring[slot][(warp * CPW) * RB + lane * RPL];
      const double* th = &S.theta[u * CU + warp * CPW];
#pragma unroll
      for (int jj = 0; jj < CPW; ++jj) {
        const double t = th[jj];
        if (RPL == 1) {
          eta[0] = fma(base[jj * RB], t, eta[0]);
        } else {
          const double2 x = *reinterpret_cast<const double2*>(&base[jj * RB]);
          eta[0] = fma(x.x, t, eta[0]);
          eta[RPL - 1] = fma(x.y, t, eta[RPL - 1]);
        }
      }
      if (!GRAD) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.empty_bar[slot]);
      }
    }
    // ---- cross-warp sum of eta (double-buffered scratch, one named barrier per row block)
    double* red = &S.red[it & 1][0][0];
#pragma unroll
    for (int r = 0; r < RPL; ++r) red[warp * 64 + lane * RPL + r] = eta[r];
    asm volatile("bar.sync 1, %0;" ::"n"(DP_CONSUMER_WARPS * 32) : "memory");
    if (GRAD || warp == 0) {
      double rres[RPL];
#pragma unroll
      for (int r = 0; r < RPL; ++r) {
        double e = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < DP_CONSUMER_WARPS; ++w2) e += red[w2 * 64 + lane * RPL + r];
        double sig, sp;
        sigmoid_softplus(e, sig, sp);
        const bool valid = (row0 + r) < N;
        rres[r] = valid ? (yv[r] - sig) : 0.0;
        if (warp == 0) {
          ll += valid ? (yv[r] * e - sp) : 0.0;
          if (GRAD && valid) w_out[row0 + r] = sig * (1.0 - sig);
        }
      }
      if (GRAD) {
        // ---- phase 2: gradient partials from the still-resident units
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          const unsigned qq = q + u, slot = qq % DP_NSLOT;
          const double* base = &S.ring[slot][(warp * CPW) * RB + lane * RPL];
#pragma unroll
          for (int jj = 0; jj < CPW; ++jj) {
            if (RPL == 1) {
              gacc[GRAD ? u * CPW + jj : 0] = fma(base[jj * RB], rres[0], gacc[GRAD ? u * CPW + jj : 0]);
            } else {
              const double2 x = *reinterpret_cast<const double2*>(&base[jj * RB]);
              double a = gacc[GRAD ? u * CPW + jj : 0];
              a = fma(x.x, rres[0], a);
              a = fma(x.y, rres[RPL - 1], a);
              gacc[GRAD ? u * CPW + jj : 0] = a;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&S.empty_bar[slot]);
        }
      }
    }
    q += NU;
  }

  // ---- per-CTA partials, fixed order
  if (warp == 0) {
    ll = warp_sum(ll);
    if (lane == 0) part_ll[blockIdx.x] = ll;
  }
  if (GRAD) {
#pragma unroll
    for (int u = 0; u < NU; ++u) {
#pragma unroll
      for (int jj = 0; jj < CPW; ++jj) {
        const double v = warp_sum(gacc[GRAD ? u * CPW + jj : 0]);
        if (lane == 0) part_g[(size_t)blockIdx.x * part_g_stride + u * CU + warp * CPW + jj] = v;
      }
    }
  }
}

}  // namespace gamc
