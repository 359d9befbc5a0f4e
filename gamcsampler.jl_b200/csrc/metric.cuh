// metric.cuh -- Kernel B: Fisher metric partials  G_lik = X' diag(w) X  on the FP64 tensor cores.
//
// Replaces `broadcast(*, r.*(1-r), X)' * X` of examples/logistic_regression/src/GAMC/simulations.jl:36
// (an N x p temporary followed by a full, symmetry-unaware dgemm) by a weighted SYRK that only forms the
// lower triangle, in 32 x 32 "tasks" (column block bi x column block bj, bi >= bj), one task per warp,
// 4 x 4 DMMA.8x8x4 accumulator tiles per task held in registers for the whole K (row) range of the CTA.
//
// Operand layout: the K dimension of the contraction is the observation index i, which is the
// CONTIGUOUS dimension of column-major X, so both MMA operands (A = X' "row-major", B = X "col-major")
// are read straight out of the same shared-memory slab: a TMA box of 36 rows x 32 columns per column
// block.  36 (= 4 mod 16 in units of 8-byte bank pairs) makes the fragment loads -- 8 columns x 4 rows
// per half-warp -- conflict-free without swizzling; rows beyond N are zero-filled by TMA.
// Work split: grid = nkinds x nsplit CTAs; a "kind" owns up to 12 tasks and the <= MAXS column blocks they
// touch (plan built on the host, metric_plan.h); split s processes row slabs s, s+nsplit, ...; partial
// tiles go to a workspace and are summed in fixed split order by k_reduce (deterministic).
// Compute-bound: algorithmic flops = N p (p+1); the 8x8 tiles above the diagonal of diagonal tasks are skipped.
#pragma once
#include "common.cuh"

namespace gamc {

#ifndef GAMC_MT_SLAB
#define GAMC_MT_SLAB 36
#endif
#ifndef GAMC_MT_NSTAGE
#define GAMC_MT_NSTAGE 3
#endif
constexpr int MT_SLAB = GAMC_MT_SLAB;          // rows per slab (K-extent per stage); 4 mod 16 keeps the fragment loads conflict-free
constexpr int MT_BLK = 32;                   // columns per block
constexpr int MT_MAXS = 8;                   // max column blocks resident per CTA
constexpr int MT_MAXT = 12;                  // tasks (consumer warps) per CTA
constexpr int MT_THREADS = (MT_MAXT + 1) * 32;
constexpr int MT_NSTAGE = GAMC_MT_NSTAGE;
constexpr int MT_SLOT_DOUBLES = MT_SLAB * MT_BLK;  // 1152 doubles = 9216 B

struct MetricPlan {
  int nslots;
  int ntasks;
  int slot_block[MT_MAXS];
  int task_sa[MT_MAXT];   // slot holding column block bi (rows of the output task)
  int task_sb[MT_MAXT];   // slot holding column block bj (columns of the output task)
  int task_id[MT_MAXT];   // global task index (position in the workspace)
  int task_diag[MT_MAXT];
};

constexpr int MT_WPAD = (MT_SLAB + 15) / 16 * 16;
template <int NS>
struct MetricSmemT {
  alignas(128) double slab[NS][MT_MAXS][MT_SLOT_DOUBLES];
  alignas(128) double w[NS][MT_WPAD];
  alignas(8) uint64_t full_bar[NS];
  alignas(8) uint64_t empty_bar[NS];
  MetricPlan plan;
};
using MetricSmem = MetricSmemT<MT_NSTAGE>;
constexpr int MT_NSTAGE_CO = 2;              // stages of the co-scheduled variant (shares the SM with k_datapass_co)

template <bool DIAG>
__device__ __forceinline__ void metric_slab(const double* __restrict__ A, const double* __restrict__ B,
                                            const double* __restrict__ wv, int lane, double (&acc)[4][4][2]) {
  const int g = lane >> 2, c = lane & 3;
  const double* ap = A + g * MT_SLAB + c;
  const double* bp = B + g * MT_SLAB + c;
#pragma unroll
  for (int k4 = 0; k4 < MT_SLAB / 4; ++k4) {
    const double wk = wv[k4 * 4 + c];
    double a[4], b[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      a[t] = ap[t * 8 * MT_SLAB + k4 * 4] * wk;
      b[t] = bp[t * 8 * MT_SLAB + k4 * 4];
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        if (!DIAG || mt >= nt) dmma_8x8x4(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
}

// Hand-over block of the co-scheduled variant (CO): the weights of a slab come from k_datapass_co, which runs at the same time.
struct MetricCo {
  const double* w;                         // weights sigma (1 - sigma), written by k_datapass_co during this kernel
  const unsigned long long* dp_progress;   // [dp_grid] (epoch << 32 | 32-row blocks finished by data-pass CTA c; it owns blocks c, c + dp_grid, ...)
  unsigned long long* mt_front;            // (epoch << 32 | first row not yet swept by the most advanced metric CTA): the data pass throttles on it; [1] = hand-over error word
  long long N;
  int dp_grid;
  unsigned epoch;
  unsigned long long* dbg;                 // GAMC_CO_DEBUG: time stamps / CTA placement (else null)
};

// NS = pipeline stages.  CO = false: the weights are a finished vector (TMA, mapW).  CO = true: launched with programmatic
// stream serialisation right after k_datapass_co and running next to it; the producer warp waits until the data pass has
// published the row blocks of a slab, copies the slab's weights to shared memory itself (L2 loads: they were written by other
// SMs during this kernel) and then issues the TMA loads of X; the kernel ends with griddepcontrol.wait, so it cannot complete
// before the data pass has.
template <int NS, bool CO>
__global__ void __launch_bounds__(MT_THREADS, 1)
k_metric_t(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW,
           const MetricPlan* __restrict__ plans, const int* __restrict__ cta_kind, const int* __restrict__ cta_split,
           const int* __restrict__ kind_nsplit, long long nslabs, double* __restrict__ ws, int max_nsplit, const int* skip_flag,
           const MetricCo co) {
  if (skip_flag && *reinterpret_cast<const volatile int*>(skip_flag)) return;   // gamc_config.reuse_tensor
  extern __shared__ __align__(128) unsigned char smem_raw[];
  MetricSmemT<NS>& S = *reinterpret_cast<MetricSmemT<NS>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kind = cta_kind[blockIdx.x], split = cta_split[blockIdx.x], nsplit = kind_nsplit[kind];
  if (CO && co.dbg && tid == 0) {
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    atomicMin(&co.dbg[2], global_timer_ns());
    atomicAdd(&co.dbg[8 + smid], 1ull << 32);
  }

  {
    const int* src = reinterpret_cast<const int*>(&plans[kind]);
    int* dst = reinterpret_cast<int*>(&S.plan);
    for (int i = tid; i < (int)(sizeof(MetricPlan) / sizeof(int)); i += MT_THREADS) dst[i] = src[i];
  }
  if (tid == 0) {
    const int nt = plans[kind].ntasks;   // only warps that own a task take part in the empty-barrier protocol
    for (int s = 0; s < NS; ++s) { mbar_init(&S.full_bar[s], 1); mbar_init(&S.empty_bar[s], nt); }
    mbar_fence_init();
  }
  __syncthreads();
  const int nslots = S.plan.nslots, ntasks = S.plan.ntasks;

  if (warp == MT_MAXT) {
    if (!CO) {
      if (lane == 0) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapW);
        unsigned it = 0;
        for (long long slab = split; slab < nslabs; slab += nsplit, ++it) {
          const unsigned st = it % NS, round = it / NS;
          mbar_wait(&S.empty_bar[st], (round & 1) ^ 1);
          mbar_arrive_expect_tx(&S.full_bar[st], (uint32_t)(nslots * MT_SLOT_DOUBLES * 8 + MT_SLAB * 8));
          for (int s = 0; s < nslots; ++s)
            tma_load_2d(S.slab[st][s], &mapX, &S.full_bar[st], (int)(slab * MT_SLAB), S.plan.slot_block[s] * MT_BLK);
          tma_load_1d(S.w[st], &mapW, &S.full_bar[st], (int)(slab * MT_SLAB));
        }
      }
    } else {
      if (lane == 0) tma_prefetch_desc(&mapX);
      unsigned it = 0;
      for (long long slab = split; slab < nslabs; slab += nsplit, ++it) {
        const unsigned st = it % NS, round = it / NS;
        if (lane == 0) mbar_wait(&S.empty_bar[st], (round & 1) ^ 1);
        __syncwarp();
        const long long r0 = slab * MT_SLAB;
        const long long r1 = (r0 + MT_SLAB < co.N ? r0 + MT_SLAB : co.N) - 1;
        // the 32-row blocks of the data pass that cover rows r0..r1 (at most three): one lane polls each
        for (long long rb = (r0 >> 5) + lane; rb <= (r1 >> 5); rb += 32) {
          const unsigned long long need = ((unsigned long long)co.epoch << 32) | (unsigned long long)(rb / co.dp_grid + 1);
          const unsigned long long* flag = co.dp_progress + (rb % co.dp_grid);
          if (ld_acquire_u64(flag) < need) {
            // bounded (the data pass is resident before this kernel starts and never waits for it without a time limit, so this
            // only trips on a programming error): raise the hand-over error word instead of hanging the device
            const unsigned long long t0 = global_timer_ns();
            while (ld_acquire_u64(flag) < need) {
              if (global_timer_ns() - t0 > 4000000000ull) { co.mt_front[1] = 1ull; break; }
              __nanosleep(64);
            }
          }
        }
        __syncwarp();
        for (int i = lane; i < MT_WPAD; i += 32) S.w[st][i] = (i < MT_SLAB && r0 + i < co.N) ? __ldcg(&co.w[r0 + i]) : 0.0;
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_expect_tx(&S.full_bar[st], (uint32_t)(nslots * MT_SLOT_DOUBLES * 8));
          for (int s = 0; s < nslots; ++s)
            tma_load_2d(S.slab[st][s], &mapX, &S.full_bar[st], (int)(slab * MT_SLAB), S.plan.slot_block[s] * MT_BLK);
        }
      }
      asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    return;
  }

  // consumer warps (one task each); surplus warps of a partially filled kind retire
  if (warp >= ntasks) { if (CO) asm volatile("griddepcontrol.wait;" ::: "memory"); return; }
  const int sa = S.plan.task_sa[warp], sb = S.plan.task_sb[warp];
  const bool diag = S.plan.task_diag[warp] != 0;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  unsigned it = 0;
  for (long long slab = split; slab < nslabs; slab += nsplit, ++it) {
    const unsigned st = it % NS, round = it / NS;
    mbar_wait(&S.full_bar[st], round & 1);
    if (diag) metric_slab<true>(S.slab[st][sa], S.slab[st][sb], S.w[st], lane, acc);
    else      metric_slab<false>(S.slab[st][sa], S.slab[st][sb], S.w[st], lane, acc);
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&S.empty_bar[st]);
      if (CO && warp == 0) {   // sweep frontier (the most advanced metric CTA) for the data pass's throttle
        const long long fr = (slab + 1) * MT_SLAB;
        atomicMax(co.mt_front, ((unsigned long long)co.epoch << 32) | (unsigned long long)(fr < 0xffffffffll ? fr : 0xffffffffll));
      }
    }
  }

  {
    // partial tile -> workspace [task][split][32 x 32 row-major]
    double* out = ws + ((size_t)S.plan.task_id[warp] * max_nsplit + split) * (MT_BLK * MT_BLK);
    const int g = lane >> 2, c = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        double2 v = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        *reinterpret_cast<double2*>(&out[(mt * 8 + g) * MT_BLK + nt * 8 + 2 * c]) = v;
      }
  }
  if (CO) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------
// k_metric_fused: the whole level-2 evaluation (log-likelihood, gradient, metric) in ONE pass, for p <= 8 column blocks.
// Every CTA keeps all column blocks of its slab resident (blocks a kind's tasks do not use are loaded as well), so four
// "weight warps" -- three, on different scheduler partitions, because the FP64 pipe is saturated by the DMMAs and side work
// must be spread -- can form eta = X theta (two or three column blocks each), sigma, the metric weights sigma (1 - sigma) and the residual
// y - sigma for the 36 rows of the slab as soon as it lands, one stage ahead of the DMMA warps, which wait for the weights
// instead of for the TMA.  The gradient X'(y - sigma) rides along in the warps that own a DIAGONAL task: their A fragments are
// exactly the 32 columns of one block, they issue 10 instead of 16 DMMAs per k-step, and each column block has exactly one
// diagonal task in the whole plan.  The separate gradient pass over X (k_datapass<GRAD>) disappears.
// Per-CTA partials (fixed order, deterministic): part_ll[cta] (slab s is accounted for by the CTA of kind s mod nkinds that
// sweeps it), part_g[cta][column].
// ---------------------------------------------------------------------------------------------------
constexpr int MTF_NW = 3;                                     // weight warps: warps 13, 14, 15 = scheduler partitions 1, 2, 3 (16 warps in all:
                                                              // a 17th would put five warps on one partition and cap the kernel at 96 registers)
constexpr int MTF_THREADS = (MT_MAXT + 1 + MTF_NW) * 32;      // task warps, producer, weight warps
constexpr int MTF_PAD = (MT_SLAB + 15) / 16 * 16;
constexpr int MTF_RPW = (MT_SLAB + MTF_NW - 1) / MTF_NW;      // rows whose sigmoid each weight warp evaluates

struct MetricSmemF {
  alignas(128) double slab[MT_NSTAGE][MT_MAXS][MT_SLOT_DOUBLES];
  alignas(128) double yv[MT_NSTAGE][MTF_PAD];
  alignas(128) double w[MT_NSTAGE][MTF_PAD];
  alignas(128) double r[MT_NSTAGE][MTF_PAD];
  alignas(128) double eta[MTF_NW][MTF_PAD];
  double theta[MT_MAXS * MT_BLK];
  double llpart[MTF_NW];
  alignas(8) uint64_t full_bar[MT_NSTAGE];
  alignas(8) uint64_t ready_bar[MT_NSTAGE];
  alignas(8) uint64_t empty_bar[MT_NSTAGE];
  int slot_of_block[MT_MAXS];
  MetricPlan plan;
};

template <bool DIAG>
__device__ __forceinline__ void metric_slab_fused(const double* __restrict__ A, const double* __restrict__ B, const double* __restrict__ wv,
                                                  const double* __restrict__ rv, int lane, double (&acc)[4][4][2], double (&ga)[4]) {
  const int g = lane >> 2, c = lane & 3;
  const double* ap = A + g * MT_SLAB + c;
  const double* bp = B + g * MT_SLAB + c;
#pragma unroll
  for (int k4 = 0; k4 < MT_SLAB / 4; ++k4) {
    const double wk = wv[k4 * 4 + c];
    double a[4], b[4];
    if (DIAG) {
      const double rk = rv[k4 * 4 + c];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const double x = ap[t * 8 * MT_SLAB + k4 * 4];
        ga[t] = fma(x, rk, ga[t]);                       // gradient of column t*8+g, rows k = c (mod 4)
        a[t] = x * wk;
        b[t] = bp[t * 8 * MT_SLAB + k4 * 4];
      }
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        a[t] = ap[t * 8 * MT_SLAB + k4 * 4] * wk;
        b[t] = bp[t * 8 * MT_SLAB + k4 * 4];
      }
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        if (!DIAG || mt >= nt) dmma_8x8x4(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
}

__global__ void __launch_bounds__(MTF_THREADS, 1)
k_metric_fused(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY,
               const MetricPlan* __restrict__ plans, const int* __restrict__ cta_kind, const int* __restrict__ cta_split,
               const int* __restrict__ kind_nsplit, long long nslabs, double* __restrict__ ws, int max_nsplit,
               const double* __restrict__ theta, int p, long long N, double* __restrict__ part_ll, double* __restrict__ part_g,
               int part_g_stride, int nkinds, const int* skip_flag) {
  if (skip_flag && *reinterpret_cast<const volatile int*>(skip_flag)) return;   // gamc_config.reuse_tensor
  extern __shared__ __align__(128) unsigned char smem_raw[];
  MetricSmemF& S = *reinterpret_cast<MetricSmemF*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kind = cta_kind[blockIdx.x], split = cta_split[blockIdx.x], nsplit = kind_nsplit[kind];
  const int nblk = (p + MT_BLK - 1) / MT_BLK;

  {
    const int* src = reinterpret_cast<const int*>(&plans[kind]);
    int* dst = reinterpret_cast<int*>(&S.plan);
    for (int i = tid; i < (int)(sizeof(MetricPlan) / sizeof(int)); i += MTF_THREADS) dst[i] = src[i];
  }
  for (int j = tid; j < MT_MAXS * MT_BLK; j += MTF_THREADS) S.theta[j] = (j < p) ? theta[j] : 0.0;
  for (int j = tid; j < part_g_stride; j += MTF_THREADS) part_g[(size_t)blockIdx.x * part_g_stride + j] = 0.0;
  if (tid == 0) {
    const int nt = plans[kind].ntasks;
    for (int s = 0; s < MT_NSTAGE; ++s) { mbar_init(&S.full_bar[s], 1); mbar_init(&S.ready_bar[s], MTF_NW); mbar_init(&S.empty_bar[s], nt + MTF_NW); }
    mbar_fence_init();
  }
  __syncthreads();
  const int nslots = S.plan.nslots, ntasks = S.plan.ntasks;
  if (tid < MT_MAXS) {
    int sl = 0;
    for (int s = 0; s < nslots; ++s) if (S.plan.slot_block[s] == tid) sl = s;
    S.slot_of_block[tid] = sl;
  }
  __syncthreads();

  if (warp == MT_MAXT) {
    // ===== producer: X blocks + the y slab
    if (lane == 0) {
      tma_prefetch_desc(&mapX);
      tma_prefetch_desc(&mapY);
      unsigned it = 0;
      for (long long slab = split; slab < nslabs; slab += nsplit, ++it) {
        const unsigned st = it % MT_NSTAGE, round = it / MT_NSTAGE;
        mbar_wait(&S.empty_bar[st], (round & 1) ^ 1);
        mbar_arrive_expect_tx(&S.full_bar[st], (uint32_t)(nslots * MT_SLOT_DOUBLES * 8 + MT_SLAB * 8));
        for (int s = 0; s < nslots; ++s)
          tma_load_2d(S.slab[st][s], &mapX, &S.full_bar[st], (int)(slab * MT_SLAB), S.plan.slot_block[s] * MT_BLK);
        tma_load_1d(S.yv[st], &mapY, &S.full_bar[st], (int)(slab * MT_SLAB));
      }
    }
    return;
  }

  if (warp > MT_MAXT) {
    // ===== weight warps: q handles column blocks q, q+3, q+6 of eta = X theta (lane = row; rows 32..35 spread over the
    //       lanes), then the sigmoid of rows q*12 .. q*12+11
    const int q = warp - MT_MAXT - 1;
    const int lo = lane & 3, hi = lane >> 2;
    double ll = 0.0;
    unsigned it = 0;
    for (long long slab = split; slab < nslabs; slab += nsplit, ++it) {
      const unsigned st = it % MT_NSTAGE, round = it / MT_NSTAGE;
      mbar_wait(&S.full_bar[st], round & 1);
      double e0[4] = {0.0, 0.0, 0.0, 0.0}, e1 = 0.0;
      for (int b = q; b < nblk; b += MTF_NW) {
        const double* sl = S.slab[st][S.slot_of_block[b]];
        const double* th = S.theta + b * MT_BLK;
#pragma unroll 8
        for (int jj = 0; jj < MT_BLK; ++jj) e0[jj & 3] = fma(sl[jj * MT_SLAB + lane], th[jj], e0[jj & 3]);
#pragma unroll
        for (int m = 0; m < MT_BLK / 8; ++m) e1 = fma(sl[(hi + 8 * m) * MT_SLAB + 32 + lo], th[hi + 8 * m], e1);   // row 32+lo, columns hi+8m
      }
      e1 += __shfl_xor_sync(0xffffffffu, e1, 4);
      e1 += __shfl_xor_sync(0xffffffffu, e1, 8);
      e1 += __shfl_xor_sync(0xffffffffu, e1, 16);
      S.eta[q][lane] = (e0[0] + e0[1]) + (e0[2] + e0[3]);
      if (lane < MT_SLAB - 32) S.eta[q][32 + lane] = e1;
      asm volatile("bar.sync 2, %0;" ::"n"(MTF_NW * 32) : "memory");
      // every slab is swept by one CTA of each kind: the kind with index slab % nkinds adds its log-likelihood terms, the
      // others skip the log1p (the FP64 instructions of these warps queue behind the DMMAs: every one saved counts)
      const bool need_ll = (int)(slab % nkinds) == kind;
      if (lane < MTF_RPW) {
        const int k = q * MTF_RPW + lane;
        if (k < MT_SLAB) {
          double eta = 0.0;
#pragma unroll
          for (int qq = 0; qq < MTF_NW; ++qq) eta += S.eta[qq][k];
          const double e = exp(-fabs(eta));
          const double inv = 1.0 / (1.0 + e);
          const double sig = (eta >= 0.0) ? inv : e * inv;
          const bool valid = slab * MT_SLAB + k < N;
          const double y = S.yv[st][k];
          S.w[st][k] = valid ? sig * (1.0 - sig) : 0.0;
          S.r[st][k] = valid ? (y - sig) : 0.0;
          if (need_ll && valid) ll += y * eta - (fmax(eta, 0.0) + log1p(e));
        }
      }
      asm volatile("bar.sync 3, %0;" ::"n"(MTF_NW * 32) : "memory");   // eta scratch free again; w, r of this warp written
      if (lane == 0) { mbar_arrive(&S.ready_bar[st]); mbar_arrive(&S.empty_bar[st]); }
    }
    ll = warp_sum(ll);
    if (lane == 0) S.llpart[q] = ll;
    asm volatile("bar.sync 2, %0;" ::"n"(MTF_NW * 32) : "memory");
    if (q == 0 && lane == 0) {
      double t = 0.0;
#pragma unroll
      for (int qq = 0; qq < MTF_NW; ++qq) t += S.llpart[qq];
      part_ll[blockIdx.x] = t;
    }
    return;
  }

  // ===== DMMA warps (one task each); surplus warps of a partially filled kind retire
  if (warp >= ntasks) return;
  const int sa = S.plan.task_sa[warp], sb = S.plan.task_sb[warp];
  const bool diag = S.plan.task_diag[warp] != 0;
  double acc[4][4][2];
  double ga[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  unsigned it = 0;
  for (long long slab = split; slab < nslabs; slab += nsplit, ++it) {
    const unsigned st = it % MT_NSTAGE, round = it / MT_NSTAGE;
    mbar_wait(&S.ready_bar[st], round & 1);
    if (diag) metric_slab_fused<true>(S.slab[st][sa], S.slab[st][sb], S.w[st], S.r[st], lane, acc, ga);
    else      metric_slab_fused<false>(S.slab[st][sa], S.slab[st][sb], S.w[st], S.r[st], lane, acc, ga);
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty_bar[st]);
  }

  {
    double* out = ws + ((size_t)S.plan.task_id[warp] * max_nsplit + split) * (MT_BLK * MT_BLK);
    const int g = lane >> 2, c = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        double2 v = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        *reinterpret_cast<double2*>(&out[(mt * 8 + g) * MT_BLK + nt * 8 + 2 * c]) = v;
      }
    if (diag) {
      const int blk = S.plan.slot_block[sa];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        double v = ga[t];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        const int col = blk * MT_BLK + t * 8 + g;
        if (c == 0 && col < part_g_stride) part_g[(size_t)blockIdx.x * part_g_stride + col] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// k_reduce: deterministic second stage.  Sums per-CTA log-likelihood / gradient partials of Kernel A and
// the per-split metric tiles of Kernel B in a fixed order into the packed evaluation buffer
//   E = [ ll, g(0..p-1), pad, G(p x p column-major, both triangles written, exactly symmetric) ]
// (the "landing buffer"), which is also the all-reduce payload in multi-GPU jobs.
// ---------------------------------------------------------------------------------------------------
struct TaskCoord { int bi, bj; };

__global__ void __launch_bounds__(256)
k_reduce(double* __restrict__ E, int p, int goff,
         const double* __restrict__ part_ll, const double* __restrict__ part_g, int part_g_stride, int nparts,
         int want_grad, const double* __restrict__ ws, const TaskCoord* __restrict__ coords, int ntasks,
         const int* __restrict__ task_nsplit, int max_nsplit, const int* skip_flag) {
  if (skip_flag && *reinterpret_cast<const volatile int*>(skip_flag)) return;
  const int nblk_tensor = ws ? ntasks * 4 : 0;  // 256 threads per block, 1024 elements per task
  if ((int)blockIdx.x < nblk_tensor) {
    const int task = blockIdx.x >> 2;
    const int e = ((blockIdx.x & 3) << 8) + threadIdx.x;   // element within the 32x32 tile, row-major
    const int r = e >> 5, c = e & 31;
    const int bi = coords[task].bi, bj = coords[task].bj;
    if (bi == bj && r < c) return;
    const int gi = bi * MT_BLK + r, gj = bj * MT_BLK + c;
    if (gi >= p || gj >= p) return;
    double s = 0.0;
    const double* src = ws + (size_t)task * max_nsplit * (MT_BLK * MT_BLK) + e;
    const int nsplit = task_nsplit[task];
    int k = 0;
    for (; k + 16 <= nsplit; k += 16) {         // sixteen independent loads in flight (the kernel is bound by L2 latency), added in split order
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = src[(size_t)(k + u) * (MT_BLK * MT_BLK)];
#pragma unroll
      for (int u = 0; u < 16; ++u) s += v[u];
    }
    {
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = (k + u < nsplit) ? src[(size_t)(k + u) * (MT_BLK * MT_BLK)] : 0.0;
#pragma unroll
      for (int u = 0; u < 16; ++u) if (k + u < nsplit) s += v[u];
    }
    double* G = E + goff;
    G[gi + (size_t)gj * p] = s;
    G[gj + (size_t)gi * p] = s;
    return;
  }
  // gradient blocks (32 columns each): thread (column, group) adds every RG_GROUPS-th per-CTA partial with all its loads in flight,
  // the groups are then combined in a fixed order -- deterministic, and one L2 round trip deep instead of nparts / 8
  constexpr int RG_GROUPS = 8;
  __shared__ double gsum[RG_GROUPS][33];
  const int gb = (int)blockIdx.x - nblk_tensor;       // 0 .. ceil(p / 32): the last one only sums the scalar
  const int ngb = (p + 31) / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (gb == ngb || !want_grad) {
    if (warp == 0 && (gb == ngb || gb == 0)) {
      double s = 0.0;
      for (int k = lane; k < nparts; k += 32) s += part_ll[k];
      s = warp_sum(s);
      if (lane == 0) E[0] = s;
    }
    return;
  }
  const int j = gb * 32 + lane;
  double s = 0.0;
  if (j < p) {
    for (int k0 = warp; k0 < nparts; k0 += RG_GROUPS * 24) {
      double v[24];
#pragma unroll
      for (int u = 0; u < 24; ++u) { const int k = k0 + u * RG_GROUPS; v[u] = (k < nparts) ? part_g[(size_t)k * part_g_stride + j] : 0.0; }
#pragma unroll
      for (int u = 0; u < 24; ++u) s += v[u];
    }
  }
  gsum[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && j < p) {
    double t = 0.0;
#pragma unroll
    for (int g = 0; g < RG_GROUPS; ++g) t += gsum[g][lane];
    E[1 + j] = t;
  }
}

}  // namespace gamc
