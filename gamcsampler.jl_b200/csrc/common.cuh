// common.cuh -- shared device helpers for the GAMC sm_100a kernels: PTX wrappers (mbarrier, TMA bulk
// tensor copies, FP64 DMMA), warp/block reductions, the device-resident sampler state.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>

namespace gamc {

// Phase clocks (debug builds only, -DGAMC_PHASE_CLOCKS): thread 0 accumulates clock64 deltas between PHASE() marks of
// the single-CTA step kernels; read back with gamc_debug_phases().  Compiled out of the product library.
#ifdef GAMC_PHASE_CLOCKS
static __device__ long long g_phase_acc[64];
static __device__ long long g_phase_cnt[64];
static __device__ long long g_phase_last;
#define GAMC_PHASE_BEGIN() do { __syncthreads(); if (threadIdx.x == 0) g_phase_last = clock64(); } while (0)
#define GAMC_PHASE(k) do { __syncthreads(); if (threadIdx.x == 0) { const long long now_ = clock64(); \
    g_phase_acc[k] += now_ - g_phase_last; g_phase_cnt[k] += 1; g_phase_last = now_; } } while (0)
// per-warp variant (no barrier): GAMC_WCLK_BEGIN(t) then GAMC_WCLK(t, k) from one lane of the warp being timed
#define GAMC_WCLK_BEGIN(t) long long t = clock64()
#define GAMC_WCLK(t, k) do { if ((threadIdx.x & 31) == 0) { const long long now_ = clock64(); g_phase_acc[k] += now_ - t; g_phase_cnt[k] += 1; t = now_; } } while (0)
#else
#define GAMC_PHASE_BEGIN() do { } while (0)
#define GAMC_PHASE(k) do { } while (0)
#define GAMC_WCLK_BEGIN(t) do { } while (0)
#define GAMC_WCLK(t, k) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------------
// Device-resident sampler state: the scalar fields of MuvGAMCState / GAMCTune
// (reference src/samplers/GAMC.jl:7-27, src/tuners/GAMCTuner.jl:1-7).
// ---------------------------------------------------------------------------------------------------
struct DevState {
  long long count;
  long long updatetensorcount;
  double step, sqrtstep;
  long long tot_acc, tot_prop, tot_totprop;
  double tot_rate;
  long long sm_acc, sm_prop, sm_totprop;
  double sm_rate;
  long long am_acc, am_prop, am_totprop;
  double am_rate;
  double smfreq, amfreq;
  double logtarget;   // pstate.logtarget (current point)
  double ratio;       // sstate.ratio of the last transition
  int cur;            // which evaluation slot (0/1) holds (ll, grad, G) / U / firstterm of the current point
  int last_accept;
  int error_flag, error_pivot;
  long long error_iter;
  long long chain_saved;
  double sumlog[2];   // sum_i log U_ii of the factor in each slot
  double mix_scratch;
  int lcur;           // which of the three AM factor buffers holds chol(C) (am_mode 2 rotates them; otherwise 0)
  int spec_err[2];    // speculative factor update of outcome 0 (accept) / 1 (reject) failed
  long long tune_events;   // burn-in tuning events so far (records of the tune log)
  int tensor_valid;   // slot `cur` (gradient, metric, factor, G^-1 grad) belongs to the current point: no AM accept since it was filled
};

// Tuner / range constants handed to the step kernels by value.
struct TuneCfg {
  int tuning;        // job.tuner.verbose || totaltuner isa AcceptanceRateMCTuner (iterate/GAMC.jl:4,199)
  int count_total;   // totaltuner.verbose || isa AcceptanceRateMCTuner          (:114,287)
  int is_accrate;    // totaltuner isa AcceptanceRateMCTuner                     (:270)
  int verbose_sm, verbose_am;
  int period;
  long long burnin;
  double targetrate, score_k;
  double minorscale, sqrtminorscale, c;
  double lambda;     // prior variance of the logistic target
  int p;
  int am_mode;
  int reuse_tensor;  // gamc_config.reuse_tensor
};

// Warm the L1 line of a global address this CTA will read later in the kernel (the single-CTA step kernels are chains of
// dependent L2 round trips; issuing them together at the top collapses them into one).
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// gpu-scope flag hand-over between kernels that run at the same time (k_datapass_co -> k_metric_t<.., true>)
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the whole block (blockDim.x multiple of 32, <= 1024).  `red` = 33 doubles of shared memory.
// Fixed order => deterministic.  Result returned to every thread.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = (lane < nw) ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// ---------------------------------------------------------------------------------------------------
// mbarrier + TMA (cp.async.bulk.tensor) wrappers.  SASS: SYNCS.* / UTMALDG.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// 2-D tiled TMA load: box at (c0 = innermost coordinate, c1) of the tensor map -> shared memory.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// Same with an L2 cache-policy hint (e.g. evict-first for data that is streamed once).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0) {
  asm volatile(
      "cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// FP64 tensor-core MMA, D(8x8) += A(8x4,row) * B(4x8,col).  SASS: DMMA.8x8x4 (the native FP64 MMA shape on
// sm_100a; larger PTX shapes decompose into it).  Fragment layout (PTX ISA): a: row = lane>>2, k = lane&3;
// b: k = lane&3, col = lane>>2; c/d: row = lane>>2, cols = 2*(lane&3) + {0,1}.
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Numerically stable logistic pieces.  eta -> (sigma(eta), softplus(eta)); restates
// `1./(1+exp.(-Xp))` and `log.(1+exp.(Xp))` of examples/logistic_regression/src/GAMC/simulations.jl:27,32,35.
__device__ __forceinline__ void sigmoid_softplus(double eta, double& sig, double& sp) {
  const double e = exp(-fabs(eta));
  const double inv = 1.0 / (1.0 + e);
  sig = (eta >= 0.0) ? inv : e * inv;
  sp = fmax(eta, 0.0) + log1p(e);
}

}  // namespace gamc
