// bigd_abi.cu -- C ABI of the large-dimension batched t-target path (include/gamc_b200.h, "batched chains, large dimension";
// kernels in bigd.cuh).  Own translation unit; state behind ext[GAMC_EXT_BIGD].
#include <cmath>
#include <cstring>
#include <algorithm>
#include <vector>
#include "ctx.h"
#include "common.cuh"
#include "synth.cuh"
#include "bigd.cuh"

using namespace gamc;

namespace {
struct BigdExt {
  BdCfg cfg{}; long long chain_offset = 0; bool target = false, ready = false;
  double* d_const = nullptr;     // [td | te | ld | le | rld | poles]
  double* d_vecs = nullptr; double* d_scal = nullptr; double* d_W = nullptr; double* d_B = nullptr; double* d_delta = nullptr; double* d_S = nullptr;
  double* d_L = nullptr; double* d_rdiag = nullptr; double* d_chain = nullptr; unsigned char* d_accept = nullptr;
  double* d_tape = nullptr; long long tape_doubles = 0;
};

void free_run_state(BigdExt* x) {
  cudaFree(x->d_vecs); cudaFree(x->d_scal); cudaFree(x->d_W); cudaFree(x->d_B); cudaFree(x->d_delta); cudaFree(x->d_S);
  cudaFree(x->d_L); cudaFree(x->d_rdiag); cudaFree(x->d_chain); cudaFree(x->d_accept); cudaFree(x->d_tape);
  x->d_vecs = x->d_scal = x->d_W = x->d_B = x->d_delta = x->d_S = x->d_L = x->d_rdiag = x->d_chain = x->d_tape = nullptr;
  x->d_accept = nullptr; x->tape_doubles = 0; x->ready = false;
}
void destroy_bigd(gamc_ctx_base* b) {
  BigdExt* x = static_cast<BigdExt*>(b->ext[GAMC_EXT_BIGD]);
  if (!x) return;
  free_run_state(x);
  cudaFree(x->d_const);
  delete x;
  b->ext[GAMC_EXT_BIGD] = nullptr;
}
BigdExt* ext_of(gamc_ctx_base* b) {
  if (!b->ext[GAMC_EXT_BIGD]) { b->ext[GAMC_EXT_BIGD] = new BigdExt(); b->ext_free[GAMC_EXT_BIGD] = destroy_bigd; }
  return static_cast<BigdExt*>(b->ext[GAMC_EXT_BIGD]);
}
int warp_grid(int nchains) { return (nchains + BD_WPB - 1) / BD_WPB; }
int block_threads(int d) { return (d + 31) & ~31; }

gamc_status check_errors(gamc_ctx_base* h, BigdExt* x, bool at_init) {
  const BdCfg& C = x->cfg;
  std::vector<double> err(C.nchains);
  CUDA_TRY(h, cudaMemcpy2DAsync(err.data(), sizeof(double), x->d_scal + BDS_ERR, sizeof(double) * BDS_N, sizeof(double), C.nchains, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  for (int c = 0; c < C.nchains; ++c) {
    if (err[c] == 0.0) continue;
    if (err[c] == 6.0) {
      h->err = "chain " + std::to_string(c) + ": the softabs correction needs more than max_factors = " + std::to_string(C.ncmax) +
               " product factors of 32 columns (the point is far from the mode: many eigenvalues of the tensor lie below 21/alpha); raise max_factors";
      return GAMC_UNSUPPORTED_SHAPE;
    }
    if (at_init || err[c] == 2.0) {
      h->err = "Log-target, gradient or tensor not finite / not positive definite at the initial value of chain " + std::to_string(c);
      return GAMC_NONFINITE_INIT;
    }
    h->err = "matrix is not positive definite; Cholesky factorization failed (chain " + std::to_string(c) + ")";
    return GAMC_NOT_POSDEF;
  }
  return GAMC_OK;
}
}  // namespace

extern "C" {

gamc_status gamc_bigd_target_mvt(gamc_handle hh, int32_t n, double nu, const double* tdiag, const double* toff) {
  if (!hh) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  REQUIRE(h, tdiag && (toff || n == 1) && n >= 1 && nu > 0, GAMC_INVALID_ARG, "gamc_bigd_target_mvt: bad argument");
  REQUIRE(h, n <= BD_MAXD, GAMC_UNSUPPORTED_SHAPE, "the structured t-target path supports 1 <= n <= 1024");
  cudaSetDevice(h->device);
  BigdExt* x = ext_of(h);
  free_run_state(x);
  cudaFree(x->d_const); x->d_const = nullptr; x->target = false;
  // bidiagonal Cholesky factor of the index-reversed tridiagonal (host, O(n), once)
  std::vector<double> c(6 * (size_t)n, 0.0);
  double* td = c.data(); double* te = td + n; double* ld = te + n; double* le = ld + n; double* rld = le + n; double* poles = rld + n;
  for (int i = 0; i < n; ++i) td[i] = tdiag[i];
  for (int i = 0; i + 1 < n; ++i) te[i] = toff[i];
  auto tdm = [&](int i) { return td[n - 1 - i]; };
  auto tem = [&](int i) { return te[n - 2 - i]; };
  double sumlog = 0.0, glo = INFINITY, ghi = -INFINITY;
  for (int i = 0; i < n; ++i) {
    const double piv = tdm(i) - (i > 0 ? le[i - 1] * le[i - 1] : 0.0);
    REQUIRE(h, piv > 0.0 && std::isfinite(piv), GAMC_NOT_POSDEF, "gamc_bigd_target_mvt: the inverse scale matrix is not positive definite");
    ld[i] = sqrt(piv); rld[i] = 1.0 / ld[i];
    if (i + 1 < n) le[i] = tem(i) / ld[i];
    sumlog += log(ld[i]);
    const double off = (i > 0 ? fabs(te[i - 1]) : 0.0) + (i + 1 < n ? fabs(te[i]) : 0.0);
    glo = std::min(glo, td[i] - off); ghi = std::max(ghi, td[i] + off);
  }
  // eigenvalues of T, ascending, by bisection on the Sturm count (host, O(n^2), once): the poles of the secular function
  {
    auto count_below = [&](double mu) {
      int neg = 0;
      double D = td[0] - mu;
      if (fabs(D) < 1e-290) D = -1e-290;
      neg += D < 0.0;
      for (int i = 1; i < n; ++i) {
        D = (td[i] - mu) - te[i - 1] * te[i - 1] / D;
        if (fabs(D) < 1e-290) D = -1e-290;
        neg += D < 0.0;
      }
      return neg;
    };
    for (int k = 0; k < n; ++k) {
      double lo = glo - 1e-3 * std::max(1.0, fabs(glo)), hi = ghi + 1e-3 * std::max(1.0, fabs(ghi));
      if (k > 0) lo = poles[k - 1];
      for (int it = 0; it < 200 && hi - lo > 2e-16 * std::max(fabs(lo), fabs(hi)); ++it) {
        const double mid = 0.5 * (lo + hi);
        if (count_below(mid) >= k + 1) hi = mid; else lo = mid;
      }
      poles[k] = 0.5 * (lo + hi);
    }
  }
  CUDA_TRY(h, cudaMalloc(&x->d_const, sizeof(double) * c.size()));
  CUDA_TRY(h, cudaMemcpyAsync(x->d_const, c.data(), sizeof(double) * c.size(), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  BdCfg& C = x->cfg;
  C = BdCfg{};
  C.d = n; C.nu = nu;
  C.td = x->d_const; C.te = C.td + n; C.ld = C.te + n; C.le = C.ld + n; C.rld = C.le + n; C.poles = C.rld + n;
  C.sumlog_ld = sumlog; C.gersh_lo = glo; C.gersh_hi = ghi;
  // logpdf(MvTDist(nu, 0, Sigma_t), x) = logconst - (nu+n)/2 log1p(x'Tx/nu);  -1/2 logdet Sigma_t = +1/2 logdet T = sum log ld
  C.logconst = lgamma((nu + n) / 2.0) - lgamma(nu / 2.0) - 0.5 * n * log(nu * M_PI) + sumlog;
  x->target = true;
  return GAMC_OK;
}

gamc_status gamc_bigd_set_chain_offset(gamc_handle hh, int64_t chain_offset) {
  if (!hh) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  REQUIRE(h, chain_offset >= 0, GAMC_INVALID_ARG, "chain_offset < 0");
  BigdExt* x = ext_of(h);
  x->chain_offset = chain_offset; x->cfg.chain_offset = chain_offset;
  return GAMC_OK;
}

gamc_status gamc_bigd_init(gamc_handle hh, const gamc_config* cfg, int32_t nchains, const double* theta0, int32_t transform, double softabs_alpha,
                           int32_t max_factors) {
  if (!hh) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  BigdExt* x = ext_of(h);
  REQUIRE(h, x->target, GAMC_NOT_READY, "no target: call gamc_bigd_target_mvt first");
  REQUIRE(h, cfg && theta0 && nchains >= 1, GAMC_INVALID_ARG, "null argument");
  REQUIRE(h, cfg->driftstep > 0, GAMC_INVALID_ARG, "Drift step is not positive");
  REQUIRE(h, cfg->minorscale > 0, GAMC_INVALID_ARG, "Constant minorscale must be positive");
  REQUIRE(h, cfg->c >= 0, GAMC_INVALID_ARG, "Constant c must be non-negative");
  REQUIRE(h, cfg->t0 > 0, GAMC_INVALID_ARG, "t0 is not positive");
  REQUIRE(h, cfg->schedule >= 0 && cfg->schedule <= GAMC_SCHED_RAND_QUAD_DECAY, GAMC_INVALID_ARG, "unknown schedule");
  REQUIRE(h, cfg->period >= 1 && cfg->nsteps >= 1 && cfg->burnin >= 0, GAMC_INVALID_ARG, "bad period / nsteps / burnin");
  REQUIRE(h, transform == 0 || (transform == 1 && softabs_alpha > 0), GAMC_INVALID_ARG, "transform must be 0 (none) or 1 (softabs, alpha > 0)");
  REQUIRE(h, max_factors >= 0 && max_factors <= 64, GAMC_INVALID_ARG, "max_factors out of range");
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  free_run_state(x);
  BdCfg& C = x->cfg;
  const int d = C.d;
  C.nchains = nchains; C.chain_offset = x->chain_offset; C.transform = transform; C.alpha = transform ? softabs_alpha : 1000.0;
  C.ncmax = max_factors > 0 ? max_factors : std::min(5, (d + 1 + BD_CH - 1) / BD_CH);
  C.driftstep = cfg->driftstep; C.minorscale = cfg->minorscale; C.sqrtminorscale = sqrt(cfg->minorscale); C.c = cfg->c;
  C.t0 = cfg->t0; C.schedule = cfg->schedule; C.sp0 = cfg->sched_params[0]; C.sp1 = cfg->sched_params[1]; C.sp2 = cfg->sched_params[2];
  C.is_accrate = cfg->total_tuner_kind == 1;
  C.tuning = (cfg->verbose_total || cfg->verbose_smmala || cfg->verbose_am || C.is_accrate) ? 1 : 0;
  C.count_total = (cfg->verbose_total || C.is_accrate) ? 1 : 0;
  C.verbose_sm = cfg->verbose_smmala; C.verbose_am = cfg->verbose_am; C.period = cfg->period;
  C.nsteps = cfg->nsteps; C.burnin = cfg->burnin; C.targetrate = cfg->targetrate; C.score_k = cfg->score_k;
  C.chain_cap = std::max<long long>(0, cfg->nsteps - cfg->burnin);
  const size_t nc = (size_t)nchains, gen = nc * 2 * C.ncmax * (size_t)d * BD_CH;
  CUDA_TRY(h, cudaMalloc(&x->d_vecs, sizeof(double) * nc * BDV_N * d));
  CUDA_TRY(h, cudaMalloc(&x->d_scal, sizeof(double) * nc * BDS_N));
  CUDA_TRY(h, cudaMalloc(&x->d_W, sizeof(double) * gen));
  CUDA_TRY(h, cudaMalloc(&x->d_B, sizeof(double) * gen));
  CUDA_TRY(h, cudaMalloc(&x->d_delta, sizeof(double) * nc * 2 * C.ncmax * d));
  CUDA_TRY(h, cudaMalloc(&x->d_S, sizeof(double) * nc * 2 * C.ncmax * BD_CH));
  CUDA_TRY(h, cudaMalloc(&x->d_L, sizeof(double) * nc * (size_t)d * d));
  CUDA_TRY(h, cudaMalloc(&x->d_rdiag, sizeof(double) * nc * d));
  if (C.chain_cap > 0) {
    CUDA_TRY(h, cudaMalloc(&x->d_chain, sizeof(double) * nc * C.chain_cap * d));
    CUDA_TRY(h, cudaMalloc(&x->d_accept, nc * C.chain_cap));
  }
  CUDA_TRY(h, cudaMemsetAsync(x->d_vecs, 0, sizeof(double) * nc * BDV_N * d, h->stream));
  C.vecs = x->d_vecs; C.scal = x->d_scal; C.W = x->d_W; C.B = x->d_B; C.delta = x->d_delta; C.S = x->d_S; C.L = x->d_L; C.rdiagL = x->d_rdiag;
  C.chain_value = x->d_chain; C.chain_accept = x->d_accept;
  C.tape_usched = C.tape_umix = C.tape_uacc = C.tape_z = nullptr; C.tape_len = 0; C.seed = 0;
  double* d_t0 = nullptr;
  CUDA_TRY(h, cudaMalloc(&d_t0, sizeof(double) * nc * d));
  CUDA_TRY(h, cudaMemcpyAsync(d_t0, theta0, sizeof(double) * nc * d, cudaMemcpyHostToDevice, h->stream));
  const size_t smem = bd_warp_smem_bytes(d);
  gamc_status st = ensure_attr(h, (const void*)k_bd_init, smem);
  if (st == GAMC_OK) st = ensure_attr(h, (const void*)k_bd_smmala_pre, smem);
  if (st == GAMC_OK) st = ensure_attr(h, (const void*)k_bd_smmala_post, smem);
  if (st == GAMC_OK) st = ensure_attr(h, (const void*)k_bd_am_post, smem);
  if (st != GAMC_OK) { cudaFree(d_t0); return st; }
  {
    ProfScope ps(h, GAMC_K_LINALG);
    k_bd_init<<<warp_grid(nchains), 32 * BD_WPB, smem, h->stream>>>(C, d_t0);
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_t0);
  if (e != cudaSuccess) { h->err = std::string("bigd init: ") + cudaGetErrorString(e); return GAMC_CUDA_ERROR; }
  st = check_errors(h, x, true);
  if (st != GAMC_OK) return st;
  x->ready = true;
  return GAMC_OK;
}

gamc_status gamc_bigd_run(gamc_handle hh, int64_t n, uint64_t seed, const double* u_sched, const double* u_mix, const double* z, const double* u_acc) {
  if (!hh) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  BigdExt* x = ext_of(h);
  REQUIRE(h, x->ready, GAMC_NOT_READY, "bigd sampler not initialised");
  REQUIRE(h, n >= 0, GAMC_INVALID_ARG, "n < 0");
  cudaSetDevice(h->device);
  BdCfg& C = x->cfg;
  const int d = C.d;
  const bool tape = u_sched || u_mix || z || u_acc;
  if (tape) {
    REQUIRE(h, u_sched && u_mix && z && u_acc, GAMC_INVALID_ARG, "either all tape pointers or none");
    const long long per = (long long)C.nchains * n;
    const long long need = per * (3 + d);
    if (need > x->tape_doubles) {
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      cudaFree(x->d_tape); x->d_tape = nullptr; x->tape_doubles = 0;
      CUDA_TRY(h, cudaMalloc(&x->d_tape, sizeof(double) * need));
      x->tape_doubles = need;
    }
    CUDA_TRY(h, cudaMemcpyAsync(x->d_tape, u_sched, sizeof(double) * per, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(x->d_tape + per, u_mix, sizeof(double) * per, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(x->d_tape + 2 * per, u_acc, sizeof(double) * per, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(x->d_tape + 3 * per, z, sizeof(double) * per * d, cudaMemcpyHostToDevice, h->stream));
    C.tape_usched = x->d_tape; C.tape_umix = x->d_tape + per; C.tape_uacc = x->d_tape + 2 * per; C.tape_z = x->d_tape + 3 * per;
    C.tape_len = n;
  } else {
    C.tape_usched = C.tape_umix = C.tape_uacc = C.tape_z = nullptr; C.tape_len = 0;
  }
  C.seed = seed;
  const size_t smem = bd_warp_smem_bytes(d);
  const int wg = warp_grid(C.nchains), bt = block_threads(d);
  for (long long it = 0; it < n; ++it) {
    // one transition of every chain: the branch is decided per chain on the device; each kernel returns at once for the chains
    // of the other branch (src/samplers/iterate/GAMC.jl:8-12,128)
    { ProfScope ps(h, GAMC_K_LINALG); k_bd_begin<<<(C.nchains + 127) / 128, 128, 0, h->stream>>>(C, it); }
    { ProfScope ps(h, GAMC_K_SMMALA_PRE); k_bd_smmala_pre<<<wg, 32 * BD_WPB, smem, h->stream>>>(C, it); }
    { ProfScope ps(h, GAMC_K_SMMALA_POST); k_bd_smmala_post<<<wg, 32 * BD_WPB, smem, h->stream>>>(C); }
    { ProfScope ps(h, GAMC_K_HANDOVER); k_bd_handover<<<C.nchains, bt, 0, h->stream>>>(C); }
    { ProfScope ps(h, GAMC_K_AM_PRE); k_bd_am_pre<<<C.nchains, bt, 0, h->stream>>>(C, it); }
    { ProfScope ps(h, GAMC_K_AM_POST); k_bd_am_post<<<wg, 32 * BD_WPB, smem, h->stream>>>(C); }
    CUDA_TRY(h, cudaGetLastError());
  }
  return GAMC_OK;
}

gamc_status gamc_bigd_chain_get(gamc_handle hh, int32_t chain_first, int32_t nchains, double* value, uint8_t* accept) {
  if (!hh) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  BigdExt* x = ext_of(h);
  REQUIRE(h, x->ready, GAMC_NOT_READY, "bigd sampler not initialised");
  const BdCfg& C = x->cfg;
  REQUIRE(h, chain_first >= 0 && nchains >= 0 && chain_first + nchains <= C.nchains, GAMC_INVALID_ARG, "chain range out of bounds");
  cudaSetDevice(h->device);
  gamc_status st = check_errors(h, x, false);        // synchronises
  if (st != GAMC_OK) return st;
  if (nchains == 0 || C.chain_cap == 0) return GAMC_OK;
  if (value) CUDA_TRY(h, cudaMemcpyAsync(value, x->d_chain + (size_t)chain_first * C.chain_cap * C.d, sizeof(double) * (size_t)nchains * C.chain_cap * C.d, cudaMemcpyDeviceToHost, h->stream));
  if (accept) CUDA_TRY(h, cudaMemcpyAsync(accept, x->d_accept + (size_t)chain_first * C.chain_cap, (size_t)nchains * C.chain_cap, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return GAMC_OK;
}

gamc_status gamc_bigd_get_state(gamc_handle hh, int32_t chain, gamc_state_scalars* out, double* theta) {
  if (!hh || !out) return GAMC_INVALID_ARG;
  gamc_ctx_base* h = gamc_base(hh);
  BigdExt* x = ext_of(h);
  REQUIRE(h, x->ready, GAMC_NOT_READY, "bigd sampler not initialised");
  const BdCfg& C = x->cfg;
  REQUIRE(h, chain >= 0 && chain < C.nchains, GAMC_INVALID_ARG, "chain out of range");
  cudaSetDevice(h->device);
  double b[BDS_N];
  CUDA_TRY(h, cudaMemcpyAsync(b, x->d_scal + (size_t)chain * BDS_N, sizeof(b), cudaMemcpyDeviceToHost, h->stream));
  if (theta) CUDA_TRY(h, cudaMemcpyAsync(theta, x->d_vecs + ((size_t)chain * BDV_N + BDV_THETA) * C.d, sizeof(double) * C.d, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  memset(out, 0, sizeof(*out));
  out->count = (int64_t)b[BDS_COUNT]; out->updatetensorcount = (int64_t)b[BDS_UPD];
  out->step = b[BDS_STEP]; out->sqrttunestep = b[BDS_SQSTEP];
  out->total_accepted = (int64_t)b[BDS_TACC]; out->total_proposed = (int64_t)b[BDS_TPROP]; out->total_totproposed = (int64_t)b[BDS_TTOT]; out->total_rate = b[BDS_TRATE];
  out->smmala_accepted = (int64_t)b[BDS_SMACC]; out->smmala_proposed = (int64_t)b[BDS_SMPROP]; out->smmala_totproposed = (int64_t)b[BDS_SMTOT];
  out->am_accepted = (int64_t)b[BDS_AMACC]; out->am_proposed = (int64_t)b[BDS_AMPROP]; out->am_totproposed = (int64_t)b[BDS_AMTOT];
  out->smmalafrequency = NAN; out->amfrequency = NAN;
  out->logtarget = b[BDS_LT]; out->ratio = b[BDS_RATIO];
  out->presentupdatetensor = b[BDS_PRESENT] != 0.0; out->pastupdatetensor = b[BDS_PAST] != 0.0;
  out->error_flag = (int32_t)b[BDS_ERR]; out->chain_saved = (int64_t)b[BDS_SAVED];
  out->error_pivot = (int32_t)b[BDS_R0 + (int)b[BDS_CUR]];      // number of eigenvalues softabs changed at the current point (diagnostic)
  return GAMC_OK;
}

}  // extern "C"
