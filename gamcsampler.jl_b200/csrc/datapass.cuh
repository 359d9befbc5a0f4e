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

// ---------------------------------------------------------------------------------------------------
// k_datapass_ll3<NU>: the log-likelihood at THREE parameter vectors in one read of X (am_mode 3: the proposal of an AM step and
// the two candidate proposals of the NEXT AM step -- one per accept/reject outcome -- so that two consecutive AM steps cost one
// HBM pass).  Same TMA ring and the same order of summation per vector as k_datapass<NU, 1, false>.
// part_ll[v * part_stride + cta]: per-CTA partial of vector v.
// ---------------------------------------------------------------------------------------------------
constexpr int DP3_NSLOT = 12;                  // ring depth (192 KB: three parameter vectors and three scratch sets need room)
struct DataPassLL3Smem {
  alignas(128) double ring[DP3_NSLOT][DP_UNIT_DOUBLES];
  double theta[3][512];
  double red[2][3][DP_CONSUMER_WARPS][32];
  alignas(8) uint64_t full_bar[DP3_NSLOT];
  alignas(8) uint64_t empty_bar[DP3_NSLOT];
};

template <int NU>
__global__ void __launch_bounds__(DP_THREADS, 1)
k_datapass_ll3(const __grid_constant__ CUtensorMap mapX, const double* __restrict__ y, const double* __restrict__ th0,
               const double* __restrict__ th1, const double* __restrict__ th2, int p, long long N, double* __restrict__ part_ll, int part_stride) {
  constexpr int RB = 32, CU = 64, CPW = CU / DP_CONSUMER_WARPS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  DataPassLL3Smem& S = *reinterpret_cast<DataPassLL3Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long nrb = (N + RB - 1) / RB;
  for (int c = tid; c < NU * CU; c += DP_THREADS) {
    S.theta[0][c] = (c < p) ? th0[c] : 0.0; S.theta[1][c] = (c < p) ? th1[c] : 0.0; S.theta[2][c] = (c < p) ? th2[c] : 0.0;
  }
  if (tid == 0) {
    for (int s = 0; s < DP3_NSLOT; ++s) { mbar_init(&S.full_bar[s], 1); mbar_init(&S.empty_bar[s], DP_CONSUMER_WARPS); }
    mbar_fence_init();
  }
  __syncthreads();
  if (warp == DP_CONSUMER_WARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&mapX);
      unsigned q = 0;
      for (long long rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
#pragma unroll 1
        for (int u = 0; u < NU; ++u, ++q) {
          const unsigned slot = q % DP3_NSLOT, round = q / DP3_NSLOT;
          mbar_wait(&S.empty_bar[slot], (round & 1) ^ 1);
          mbar_arrive_expect_tx(&S.full_bar[slot], DP_UNIT_DOUBLES * 8);
          tma_load_2d(S.ring[slot], &mapX, &S.full_bar[slot], (int)(rb * RB), u * CU);
        }
      }
    }
    return;
  }
  double ll0 = 0.0;                              // warp v < 3: the scalar of vector v
  unsigned q = 0;
  int it = 0;
  for (long long rb = blockIdx.x; rb < nrb; rb += gridDim.x, ++it) {
    const long long row0 = rb * RB + lane;
    const double yv = (row0 < N) ? __ldg(&y[row0]) : 0.0;
    double e0 = 0.0, e1 = 0.0, e2 = 0.0;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const unsigned qq = q + u, slot = qq % DP3_NSLOT, round = qq / DP3_NSLOT;
      mbar_wait(&S.full_bar[slot], round & 1);
      const double* base = &S.ring[slot][(warp * CPW) * RB + lane];
      const int c0 = u * CU + warp * CPW;
#pragma unroll
      for (int jj = 0; jj < CPW; ++jj) {
        const double x = base[jj * RB];
        e0 = fma(x, S.theta[0][c0 + jj], e0);
        e1 = fma(x, S.theta[1][c0 + jj], e1);
        e2 = fma(x, S.theta[2][c0 + jj], e2);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty_bar[slot]);
    }
    double* red = &S.red[it & 1][0][0][0];
    red[(0 * DP_CONSUMER_WARPS + warp) * 32 + lane] = e0;
    red[(1 * DP_CONSUMER_WARPS + warp) * 32 + lane] = e1;
    red[(2 * DP_CONSUMER_WARPS + warp) * 32 + lane] = e2;
    asm volatile("bar.sync 1, %0;" ::"n"(DP_CONSUMER_WARPS * 32) : "memory");
    if (warp < 3) {                              // warp v finishes vector v
      double e = 0.0;
#pragma unroll
      for (int w2 = 0; w2 < DP_CONSUMER_WARPS; ++w2) e += red[(warp * DP_CONSUMER_WARPS + w2) * 32 + lane];
      double sig, sp;
      sigmoid_softplus(e, sig, sp);
      ll0 += (row0 < N) ? (yv * e - sp) : 0.0;
    }
    q += NU;
  }
  if (warp < 3) {
    const double v = warp_sum(ll0);
    if (lane == 0) part_ll[(size_t)warp * part_stride + blockIdx.x] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// k_datapass_co<NU>: the log-likelihood + gradient + weights pass of a level-2 evaluation, built to run CONCURRENTLY with the
// metric kernel on the same SMs (one pass over X out of HBM for the whole evaluation: the metric kernel follows a bounded
// number of rows behind and finds its slabs in L2).  The metric pass is bound by the FP64 tensor pipe and leaves HBM idle; this
// pass is bound by HBM and needs ~2 % of the FP64 pipe -- but the two only share an SM if this kernel is small:
//   * ring of exactly NU units (one 32-row block, 64 KB at p = 256) instead of 13 slots;
//   * gradient accumulators: ONE column per thread (phase 2 reads the resident block with lane = column and a rotated row
//     index, conflict-free, instead of lane = row with NU*8 accumulators);
//   * sigma on warp 0 only; its softplus (needed for the scalar only) is evaluated after the residuals are published, off
//     the other warps' critical path;
//   * eight warps of 48 registers (see the register budget below).
// Same formulas as k_datapass; the summation order of eta differs (short dependent chains), so results agree to rounding.
// Hand-over to the metric kernel: after the weights of a row block are in global memory the CTA publishes
// (epoch << 32 | blocks done) in dp_progress[cta]; the metric kernel's producer warp waits for the blocks of a slab before it
// loads the slab's weights.  Throttle (keeps the metric kernel's re-read of X inside L2): the producer does not run more than
// `lag_rows` ahead of the metric frontier (*mt_front, the most advanced metric CTA); the wait is bounded (100 us without an
// answer -> free run for the rest of the kernel), so this kernel never depends on the metric kernel being resident.
// Launched first; signals griddepcontrol.launch_dependents at once so that the metric kernel (launched with programmatic
// stream serialisation) starts while this one runs.
// ---------------------------------------------------------------------------------------------------
template <int NU>
struct DataPassCoSmem {
  alignas(128) double ring[NU][DP_UNIT_DOUBLES];
  alignas(16) double theta[NU * 64];
  double red[2][DP_CONSUMER_WARPS][32];
  alignas(16) double r[2][32];
  alignas(8) uint64_t full_bar[NU];
  alignas(8) uint64_t empty_bar[NU];
};

// Register budget: the four scheduler partitions of an SM own 16 K registers each and a warp lives on one of them; the metric CTA
// (13 warps x 104 registers) puts four warps on one partition and leaves it 3072 registers.  Eight warps of 48 registers (two per
// partition) are what fits next to it -- so no dedicated producer warp: lane 0 of the first warp of each unit re-issues the unit's TMA
// load once both of its warps are done with the slot.
constexpr int DPC_THREADS = DP_CONSUMER_WARPS * 32;

template <int NU>
__global__ void __launch_bounds__(DPC_THREADS, 5)   // 65536 / (5 x 256) -> 48 registers
k_datapass_co(const __grid_constant__ CUtensorMap mapX, const double* __restrict__ y,
              const double* __restrict__ theta, int p, long long N, double* __restrict__ w_out,
              double* __restrict__ part_ll, double* __restrict__ part_g, int part_g_stride, const int* skip_flag,
              unsigned long long* __restrict__ dp_progress, const unsigned long long* mt_front, unsigned epoch, long long lag_rows,
              unsigned long long* dbg) {
  if (skip_flag && *reinterpret_cast<const volatile int*>(skip_flag)) return;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (dbg && threadIdx.x == 0) {       // GAMC_CO_DEBUG: when and where the CTAs of the pair ran
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    atomicMin(&dbg[0], global_timer_ns());
    atomicAdd(&dbg[8 + smid], 1ull);
  }
  constexpr int RB = 32, CU = 64, CPW = CU / DP_CONSUMER_WARPS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  DataPassCoSmem<NU>& S = *reinterpret_cast<DataPassCoSmem<NU>*>(smem_raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long nrb = (N + RB - 1) / RB;

  for (int c = tid; c < NU * CU; c += DPC_THREADS) S.theta[c] = (c < p) ? theta[c] : 0.0;
  if (tid == 0) {
    for (int s = 0; s < NU; ++s) { mbar_init(&S.full_bar[s], 1); mbar_init(&S.empty_bar[s], 2); }
    mbar_fence_init();
  }
  __syncthreads();

  const int my_u = warp >> 1;                       // phase 2: this warp's unit and this thread's column in it
  const int my_c = (warp & 1) * 32 + lane;
  const bool issuer = lane == 0 && (warp & 1) == 0 && my_u < NU;   // re-loads unit my_u
  bool free_run = (mt_front == nullptr);
  long long front_ok = lag_rows;                     // rows below this may be loaded without looking at the frontier again
  if (issuer && blockIdx.x < nrb) {
    if (warp == 0) tma_prefetch_desc(&mapX);
    mbar_arrive_expect_tx(&S.full_bar[my_u], DP_UNIT_DOUBLES * 8);
    tma_load_2d(S.ring[my_u], &mapX, &S.full_bar[my_u], (int)((long long)blockIdx.x * RB), my_u * CU);
  }
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;     // gradient of column my_u * 64 + my_c over all row blocks of the CTA
  double ll = 0.0;
  unsigned k = 0;
  for (long long rb = blockIdx.x; rb < nrb; rb += gridDim.x, ++k) {
    const long long row0 = rb * RB + lane;
    if (dbg && tid == 0 && blockIdx.x == 0) {      // GAMC_CO_DEBUG: progress curve of the data pass (quarters of its row blocks)
      const unsigned nk = (unsigned)((nrb + gridDim.x - 1) / gridDim.x);
      if (k == nk / 4) dbg[4] = global_timer_ns();
      if (k == nk / 2) dbg[5] = global_timer_ns();
      if (k == (3 * nk) / 4) dbg[6] = global_timer_ns();
    }
    // ---- phase 1: eta partial over this warp's columns (lane = row).  Next to the metric kernel every FP64 instruction queues
    //      behind the DMMAs of three or four warps (~100 cycles), so dependent chains are kept short: four partial sums here,
    //      a tree over the warps' partials below (the summation order therefore differs from k_datapass: equal to rounding).
    double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      mbar_wait(&S.full_bar[u], k & 1);
      const double* base = &S.ring[u][(warp * CPW) * RB + lane];
      const double* th = &S.theta[u * CU + warp * CPW];
#pragma unroll
      for (int jj = 0; jj < CPW; jj += 4) {
        const double2 ta = *reinterpret_cast<const double2*>(&th[jj]);
        const double2 tb = *reinterpret_cast<const double2*>(&th[jj + 2]);
        e0 = fma(base[jj * RB], ta.x, e0);
        e1 = fma(base[(jj + 1) * RB], ta.y, e1);
        e2 = fma(base[(jj + 2) * RB], tb.x, e2);
        e3 = fma(base[(jj + 3) * RB], tb.y, e3);
      }
    }
    double* red = &S.red[k & 1][0][0];
    red[warp * 32 + lane] = (e0 + e1) + (e2 + e3);
    asm volatile("bar.sync 1, %0;" ::"n"(DPC_THREADS) : "memory");
    double e = 0.0, ex = 0.0, yv = 0.0;
    bool valid = false;
    if (warp == 0) {
      e = ((red[lane] + red[32 + lane]) + (red[64 + lane] + red[96 + lane])) + ((red[128 + lane] + red[160 + lane]) + (red[192 + lane] + red[224 + lane]));
      valid = row0 < N;
      yv = valid ? __ldg(&y[row0]) : 0.0;
      ex = exp(-fabs(e));
      const double inv = 1.0 / (1.0 + ex);
      const double sig = (e >= 0.0) ? inv : ex * inv;
      S.r[k & 1][lane] = valid ? (yv - sig) : 0.0;
      if (valid) w_out[row0] = sig * (1.0 - sig);
    }
    asm volatile("bar.sync 2, %0;" ::"n"(DPC_THREADS) : "memory");
    if (warp == 0) {
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        st_release_u64(&dp_progress[blockIdx.x], ((unsigned long long)epoch << 32) | (unsigned long long)(k + 1));
      }
      const double sp = fmax(e, 0.0) + log1p(ex);
      ll += valid ? (yv * e - sp) : 0.0;
    }
    // ---- phase 2: gradient, lane = column of the resident block; row index rotated by the lane (conflict-free)
    if (my_u < NU) {
      const double* col = &S.ring[my_u][my_c * RB];
      const double* rr = S.r[k & 1];
      // 16-byte loads of row pairs; the pair index is rotated by the lane so that each quarter-warp covers all banks
#pragma unroll
      for (int i = 0; i < RB; i += 4) {
        const int ra = (2 * lane + i) & 31, rb2 = (2 * lane + i + 2) & 31;
        const double2 xa = *reinterpret_cast<const double2*>(&col[ra]);
        const double2 qa = *reinterpret_cast<const double2*>(&rr[ra]);
        const double2 xb = *reinterpret_cast<const double2*>(&col[rb2]);
        const double2 qb = *reinterpret_cast<const double2*>(&rr[rb2]);
        s0 = fma(xa.x, qa.x, s0);
        s1 = fma(xa.y, qa.y, s1);
        s2 = fma(xb.x, qb.x, s2);
        s3 = fma(xb.y, qb.y, s3);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty_bar[my_u]);
      const long long rbn = rb + gridDim.x;
      if (issuer && rbn < nrb) {
        mbar_wait(&S.empty_bar[my_u], k & 1);             // both warps of the unit are done with the slot
        const long long row_first = rbn * RB;
        if (!free_run && row_first > front_ok) {
          // the cached bound does not cover this block: look at the frontier again (measured: re-reading it for every block costs
          // 45 us per evaluation, a hysteresis band on top of the cache nothing), wait for 100 us at most
          const unsigned long long t0 = global_timer_ns();
          for (;;) {
            const unsigned long long v = *reinterpret_cast<const volatile unsigned long long*>(mt_front);
            const long long front = ((unsigned)(v >> 32) == epoch) ? (long long)(v & 0xffffffffull) : 0;
            front_ok = front + lag_rows;
            if (row_first <= front_ok) break;
            if (global_timer_ns() - t0 > 100000ull) { free_run = true; break; }
            __nanosleep(256);
          }
        }
        mbar_arrive_expect_tx(&S.full_bar[my_u], DP_UNIT_DOUBLES * 8);
        tma_load_2d(S.ring[my_u], &mapX, &S.full_bar[my_u], (int)(rbn * RB), my_u * CU);
      }
      __syncwarp();
    }
  }

  if (warp == 0) {
    ll = warp_sum(ll);
    if (lane == 0) part_ll[blockIdx.x] = ll;
  }
  if (my_u < NU) part_g[(size_t)blockIdx.x * part_g_stride + my_u * CU + my_c] = (s0 + s1) + (s2 + s3);
  if (dbg && tid == 0) atomicMax(&dbg[1], global_timer_ns());
}

}  // namespace gamc
