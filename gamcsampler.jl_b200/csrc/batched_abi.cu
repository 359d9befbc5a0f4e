// batched_abi.cu -- C ABI of the batched-chains path (include/gamc_b200.h, "batched chains"): many independent GAMC chains of a
// small-dimensional target, one persistent kernel per run (batched.cuh).  Own translation unit: its state hangs off the handle's
// ext[GAMC_EXT_BATCHED] slot.
#include <cmath>
#include <cstring>
#include <algorithm>
#include "ctx.h"
#include "common.cuh"
#include "synth.cuh"
#include "batched.cuh"

using namespace gamc;

namespace {
constexpr int BT_LOGI_MAXN = 2048;   // observations of the batched logistic target (weights + residuals live in shared memory)

struct BatchedExt {
  BatchedCfg bcfg{}; long long chain_offset = 0; bool target = false, ready = false;
  double* d_const = nullptr;    // target constants: Sinv | [t, obs, sigma] | [X, y]
  double* d_blob = nullptr; double* d_chain = nullptr; unsigned char* d_accept = nullptr; double* d_tape = nullptr; long long tape_doubles = 0;
};

void free_run_state(BatchedExt* x) {
  cudaFree(x->d_blob); cudaFree(x->d_chain); cudaFree(x->d_accept); cudaFree(x->d_tape);
  x->d_blob = nullptr; x->d_chain = nullptr; x->d_accept = nullptr; x->d_tape = nullptr; x->tape_doubles = 0;
  x->ready = false;
}

void destroy_batched(gamc_ctx_base* b) {
  BatchedExt* x = static_cast<BatchedExt*>(b->ext[GAMC_EXT_BATCHED]);
  if (!x) return;
  free_run_state(x);
  cudaFree(x->d_const);
  delete x;
  b->ext[GAMC_EXT_BATCHED] = nullptr;
}

BatchedExt* ext_of(gamc_ctx_base* b) {
  if (!b->ext[GAMC_EXT_BATCHED]) { b->ext[GAMC_EXT_BATCHED] = new BatchedExt(); b->ext_free[GAMC_EXT_BATCHED] = destroy_batched; }
  return static_cast<BatchedExt*>(b->ext[GAMC_EXT_BATCHED]);
}

template <int TGT>
gamc_status launch_batched_t(gamc_ctx_base* h, BatchedExt* x, int mode, long long n, const double* d_theta0) {
  const int nt = (TGT == 0 || TGT == 3) ? 32 : 128;
  const size_t smem = bt_smem_bytes(x->bcfg.blob_doubles, nt, TGT == 3 ? 2 * x->bcfg.lg_n : 0);
  { gamc_status s = ensure_attr(h, (const void*)k_batched_gamc<TGT>, 200 * 1024); if (s != GAMC_OK) return s; }
  ProfScope ps(h, GAMC_K_LINALG);
  k_batched_gamc<TGT><<<x->bcfg.nchains, nt, smem, h->stream>>>(x->bcfg, mode, n, d_theta0);
  CUDA_TRY(h, cudaGetLastError());
  return GAMC_OK;
}

gamc_status launch_batched(gamc_ctx_base* h, BatchedExt* x, int mode, long long n, const double* d_theta0) {
  if (x->bcfg.target == 0) return launch_batched_t<0>(h, x, mode, n, d_theta0);
  if (x->bcfg.target == 3) return launch_batched_t<3>(h, x, mode, n, d_theta0);
  return x->bcfg.rv_npl == 2 ? launch_batched_t<2>(h, x, mode, n, d_theta0) : launch_batched_t<1>(h, x, mode, n, d_theta0);
}

// fresh target: drop the previous constants and run state
gamc_status reset_target(gamc_ctx_base* h, BatchedExt* x, size_t const_doubles) {
  cudaSetDevice(h->device);
  free_run_state(x);
  cudaFree(x->d_const); x->d_const = nullptr; x->target = false;
  CUDA_TRY(h, cudaMalloc(&x->d_const, sizeof(double) * const_doubles));
  x->bcfg = BatchedCfg{};
  return GAMC_OK;
}
}  // namespace

extern "C" {

gamc_status gamc_batched_target_mvt(gamc_handle hh, int32_t n, double nu, const double* Sinv, double logconst) {
  if (!hh) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  REQUIRE(h, Sinv && n >= 1 && nu > 0, GAMC_INVALID_ARG, "gamc_batched_target_mvt: bad argument");
  REQUIRE(h, n <= BT_MAXP, GAMC_UNSUPPORTED_SHAPE, "batched chains support p <= 32 (larger t targets: gamc_bigd_*)");
  BatchedExt* x = ext_of(h);
  { gamc_status s = reset_target(h, x, (size_t)n * n); if (s != GAMC_OK) return s; }
  CUDA_TRY(h, cudaMemcpyAsync(x->d_const, Sinv, sizeof(double) * n * n, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  x->bcfg.p = n; x->bcfg.ld = (n & 1) ? n : n + 1; x->bcfg.target = 0;
  x->bcfg.nu = nu; x->bcfg.logconst = logconst; x->bcfg.Sinv = x->d_const;
  x->target = true;
  return GAMC_OK;
}

gamc_status gamc_batched_target_rv(gamc_handle hh, const double* t, const double* obs, const double* sigma, int64_t N, int32_t nplanets) {
  if (!hh) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  REQUIRE(h, t && obs && sigma && N >= 1, GAMC_INVALID_ARG, "gamc_batched_target_rv: bad argument");
  REQUIRE(h, nplanets == 1 || nplanets == 2, GAMC_UNSUPPORTED_SHAPE, "radial-velocity target supports 1 or 2 planets");
  BatchedExt* x = ext_of(h);
  { gamc_status s = reset_target(h, x, 3 * (size_t)N); if (s != GAMC_OK) return s; }
  CUDA_TRY(h, cudaMemcpyAsync(x->d_const, t, sizeof(double) * N, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(x->d_const + N, obs, sizeof(double) * N, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(x->d_const + 2 * N, sigma, sizeof(double) * N, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  double lognorm = -0.5 * (double)N * log(2.0 * M_PI);          // stat_model.jl:27,33 (data constants, summed once on the host)
  for (int64_t i = 0; i < N; ++i) lognorm -= 0.5 * log(sigma[i] * sigma[i]);
  BatchedCfg& C = x->bcfg;
  C.p = 5 * nplanets + 1; C.ld = (C.p & 1) ? C.p : C.p + 1; C.target = 1;
  C.rv_t = x->d_const; C.rv_obs = x->d_const + N; C.rv_sig = x->d_const + 2 * N; C.rv_n = N; C.rv_npl = nplanets;
  C.rv_lognorm = lognorm;
  x->target = true;
  return GAMC_OK;
}

gamc_status gamc_batched_target_logistic(gamc_handle hh, const double* X, int64_t ldX, const double* y, int64_t N, int32_t p, double lambda) {
  if (!hh) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  REQUIRE(h, X && y && N >= 1 && p >= 1 && ldX >= N && lambda > 0, GAMC_INVALID_ARG, "gamc_batched_target_logistic: bad argument");
  REQUIRE(h, p <= BT_MAXP && N <= BT_LOGI_MAXN, GAMC_UNSUPPORTED_SHAPE,
          "batched logistic chains support p <= 32 and N <= 2048 (larger problems: gamc_target_logistic, one chain on the whole GPU)");
  BatchedExt* x = ext_of(h);
  { gamc_status s = reset_target(h, x, (size_t)N * (p + 1)); if (s != GAMC_OK) return s; }
  CUDA_TRY(h, cudaMemcpy2DAsync(x->d_const, sizeof(double) * N, X, sizeof(double) * ldX, sizeof(double) * N, p, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(x->d_const + (size_t)N * p, y, sizeof(double) * N, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  BatchedCfg& C = x->bcfg;
  C.p = p; C.ld = (p & 1) ? p : p + 1; C.target = 3;
  C.lg_X = x->d_const; C.lg_y = x->d_const + (size_t)N * p; C.lg_n = (int)N; C.lg_lambda = lambda;
  x->target = true;
  return GAMC_OK;
}

gamc_status gamc_batched_init(gamc_handle hh, const gamc_config* cfg, int32_t nchains, const double* theta0, int32_t transform, double softabs_alpha) {
  if (!hh) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  BatchedExt* x = ext_of(h);
  REQUIRE(h, x->target, GAMC_NOT_READY, "no batched target: call gamc_batched_target_* first");
  REQUIRE(h, cfg && theta0 && nchains >= 1, GAMC_INVALID_ARG, "null argument");
  REQUIRE(h, cfg->driftstep > 0, GAMC_INVALID_ARG, "Drift step is not positive");
  REQUIRE(h, cfg->minorscale > 0, GAMC_INVALID_ARG, "Constant minorscale must be positive");
  REQUIRE(h, cfg->c >= 0, GAMC_INVALID_ARG, "Constant c must be non-negative");
  REQUIRE(h, cfg->t0 > 0, GAMC_INVALID_ARG, "t0 is not positive");
  REQUIRE(h, cfg->schedule >= 0 && cfg->schedule <= GAMC_SCHED_RAND_QUAD_DECAY, GAMC_INVALID_ARG, "unknown schedule");
  REQUIRE(h, cfg->period >= 1 && cfg->nsteps >= 1 && cfg->burnin >= 0, GAMC_INVALID_ARG, "bad period / nsteps / burnin");
  REQUIRE(h, transform == 0 || transform == 1, GAMC_INVALID_ARG, "transform must be 0 (none) or 1 (softabs)");
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  free_run_state(x);
  BatchedCfg& C = x->bcfg;
  const int p = C.p;
  C.nchains = nchains; C.chain_offset = x->chain_offset; C.transform = transform; C.softabs_alpha = softabs_alpha;
  C.driftstep = cfg->driftstep; C.minorscale = cfg->minorscale; C.sqrtminorscale = sqrt(cfg->minorscale); C.c = cfg->c;
  C.t0 = cfg->t0; C.schedule = cfg->schedule; C.sp0 = cfg->sched_params[0]; C.sp1 = cfg->sched_params[1]; C.sp2 = cfg->sched_params[2];
  C.is_accrate = cfg->total_tuner_kind == 1;
  C.tuning = (cfg->verbose_total || cfg->verbose_smmala || cfg->verbose_am || C.is_accrate) ? 1 : 0;
  C.count_total = (cfg->verbose_total || C.is_accrate) ? 1 : 0;
  C.verbose_sm = cfg->verbose_smmala; C.verbose_am = cfg->verbose_am; C.period = cfg->period;
  C.nsteps = cfg->nsteps; C.burnin = cfg->burnin; C.targetrate = cfg->targetrate; C.score_k = cfg->score_k;
  C.blob_doubles = bt_blob_doubles(p, C.ld);
  C.chain_cap = std::max<long long>(0, cfg->nsteps - cfg->burnin);
  CUDA_TRY(h, cudaMalloc(&x->d_blob, sizeof(double) * (size_t)nchains * C.blob_doubles));
  if (C.chain_cap > 0) {
    CUDA_TRY(h, cudaMalloc(&x->d_chain, sizeof(double) * (size_t)nchains * C.chain_cap * p));
    CUDA_TRY(h, cudaMalloc(&x->d_accept, (size_t)nchains * C.chain_cap));
  }
  C.blob = x->d_blob; C.chain_value = x->d_chain; C.chain_accept = x->d_accept;
  C.tape_usched = C.tape_umix = C.tape_uacc = C.tape_z = nullptr; C.tape_len = 0; C.seed = 0;
  double* d_t0 = nullptr;
  CUDA_TRY(h, cudaMalloc(&d_t0, sizeof(double) * (size_t)nchains * p));
  CUDA_TRY(h, cudaMemcpyAsync(d_t0, theta0, sizeof(double) * (size_t)nchains * p, cudaMemcpyHostToDevice, h->stream));
  gamc_status s = launch_batched(h, x, 1, 0, d_t0);
  if (s != GAMC_OK) { cudaFree(d_t0); return s; }
  // initial-point check (src/samplers/GAMC.jl:168-170) for every chain
  std::vector<double> err(nchains);
  cudaError_t e = cudaMemcpy2DAsync(err.data(), sizeof(double), x->d_blob + BS_ERR, sizeof(double) * C.blob_doubles, sizeof(double), nchains,
                                    cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_t0);
  if (e != cudaSuccess) { h->err = std::string("batched init: ") + cudaGetErrorString(e); return GAMC_CUDA_ERROR; }
  for (int c = 0; c < nchains; ++c)
    if (err[c] != 0.0) { h->err = "Log-target, gradient or tensor not finite / not positive definite at the initial value of chain " + std::to_string(c); return GAMC_NONFINITE_INIT; }
  x->ready = true;
  return GAMC_OK;
}

gamc_status gamc_batched_run(gamc_handle hh, int64_t n, uint64_t seed, const double* u_sched, const double* u_mix, const double* z, const double* u_acc) {
  if (!hh) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  BatchedExt* x = ext_of(h);
  REQUIRE(h, x->ready, GAMC_NOT_READY, "batched sampler not initialised");
  REQUIRE(h, n >= 0, GAMC_INVALID_ARG, "n < 0");
  cudaSetDevice(h->device);
  BatchedCfg& C = x->bcfg;
  const bool tape = u_sched || u_mix || z || u_acc;
  if (tape) {
    REQUIRE(h, u_sched && u_mix && z && u_acc, GAMC_INVALID_ARG, "either all tape pointers or none");
    const long long per = (long long)C.nchains * n;
    const long long need = per * (3 + C.p);
    if (need > x->tape_doubles) {
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      cudaFree(x->d_tape); x->d_tape = nullptr; x->tape_doubles = 0;
      CUDA_TRY(h, cudaMalloc(&x->d_tape, sizeof(double) * need));
      x->tape_doubles = need;
    }
    CUDA_TRY(h, cudaMemcpyAsync(x->d_tape, u_sched, sizeof(double) * per, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(x->d_tape + per, u_mix, sizeof(double) * per, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(x->d_tape + 2 * per, u_acc, sizeof(double) * per, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(x->d_tape + 3 * per, z, sizeof(double) * per * C.p, cudaMemcpyHostToDevice, h->stream));
    C.tape_usched = x->d_tape; C.tape_umix = x->d_tape + per; C.tape_uacc = x->d_tape + 2 * per; C.tape_z = x->d_tape + 3 * per;
    C.tape_len = n;
  } else {
    C.tape_usched = C.tape_umix = C.tape_uacc = C.tape_z = nullptr; C.tape_len = 0;
  }
  C.seed = seed;
  return launch_batched(h, x, 0, n, nullptr);
}

gamc_status gamc_batched_set_chain_offset(gamc_handle hh, int64_t chain_offset) {
  if (!hh) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  REQUIRE(h, chain_offset >= 0, GAMC_INVALID_ARG, "chain_offset < 0");
  BatchedExt* x = ext_of(h);
  x->chain_offset = chain_offset;
  x->bcfg.chain_offset = chain_offset;
  return GAMC_OK;
}

gamc_status gamc_batched_chain_get(gamc_handle hh, int32_t chain_first, int32_t nchains, double* value, uint8_t* accept) {
  if (!hh) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  BatchedExt* x = ext_of(h);
  REQUIRE(h, x->ready, GAMC_NOT_READY, "batched sampler not initialised");
  const BatchedCfg& C = x->bcfg;
  REQUIRE(h, chain_first >= 0 && nchains >= 0 && chain_first + nchains <= C.nchains, GAMC_INVALID_ARG, "chain range out of bounds");
  cudaSetDevice(h->device);
  if (nchains == 0 || C.chain_cap == 0) { CUDA_TRY(h, cudaStreamSynchronize(h->stream)); return GAMC_OK; }
  if (value) CUDA_TRY(h, cudaMemcpyAsync(value, x->d_chain + (size_t)chain_first * C.chain_cap * C.p, sizeof(double) * (size_t)nchains * C.chain_cap * C.p, cudaMemcpyDeviceToHost, h->stream));
  if (accept) CUDA_TRY(h, cudaMemcpyAsync(accept, x->d_accept + (size_t)chain_first * C.chain_cap, (size_t)nchains * C.chain_cap, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return GAMC_OK;
}

gamc_status gamc_batched_get_state(gamc_handle hh, int32_t chain, gamc_state_scalars* out, double* theta) {
  if (!hh || !out) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  BatchedExt* x = ext_of(h);
  REQUIRE(h, x->ready, GAMC_NOT_READY, "batched sampler not initialised");
  const BatchedCfg& C = x->bcfg;
  REQUIRE(h, chain >= 0 && chain < C.nchains, GAMC_INVALID_ARG, "chain out of range");
  cudaSetDevice(h->device);
  std::vector<double> b(BS_NSCAL + BT_MAXP);
  CUDA_TRY(h, cudaMemcpyAsync(b.data(), x->d_blob + (size_t)chain * C.blob_doubles, sizeof(double) * (BS_NSCAL + BT_MAXP), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  memset(out, 0, sizeof(*out));
  out->count = (int64_t)b[BS_COUNT]; out->updatetensorcount = (int64_t)b[BS_UPD];
  out->step = b[BS_STEP]; out->sqrttunestep = b[BS_SQSTEP];
  out->total_accepted = (int64_t)b[BS_TACC]; out->total_proposed = (int64_t)b[BS_TPROP]; out->total_totproposed = (int64_t)b[BS_TTOT]; out->total_rate = b[BS_TRATE];
  out->smmala_accepted = (int64_t)b[BS_SMACC]; out->smmala_proposed = (int64_t)b[BS_SMPROP]; out->smmala_totproposed = (int64_t)b[BS_SMTOT];
  out->am_accepted = (int64_t)b[BS_AMACC]; out->am_proposed = (int64_t)b[BS_AMPROP]; out->am_totproposed = (int64_t)b[BS_AMTOT];
  out->smmalafrequency = NAN; out->amfrequency = NAN;
  out->logtarget = b[BS_LT]; out->ratio = b[BS_RATIO];
  out->presentupdatetensor = b[BS_PRESENT] != 0.0; out->pastupdatetensor = b[BS_PAST] != 0.0;
  out->error_flag = (int32_t)b[BS_ERR]; out->chain_saved = (int64_t)b[BS_SAVED];
  if (theta) memcpy(theta, b.data() + BS_NSCAL + BV_THETA * BT_MAXP, sizeof(double) * C.p);
  return GAMC_OK;
}

}  // extern "C"
