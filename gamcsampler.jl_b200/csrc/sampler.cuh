// sampler.cuh -- the arithmetic body of GAMC's `iterate!` (reference src/samplers/iterate/GAMC.jl:1-293)
// as single-CTA step kernels around the data-pass kernels.  The host knows the branch sequence in advance
// (the schedule coin depends only on (i, nsteps, u_sched), src/samplers/GAMC.jl:264-337), so it enqueues
//     SMMALA:  [eval(theta) if stale] k_smmala_pre  -> eval(theta*) -> k_smmala_post
//     AM:      [G^-1 from U if previous step was SMMALA] k_am_pre -> loglik(theta*) -> k_am_post
// without ever synchronising; accept/reject, state swap, running means, tuner and chain output are decided
// on the device.
#pragma once
#include "common.cuh"
#include "linalg.cuh"

namespace gamc {

// One mailbox slot of the peer all-reduce (see k_am_post) and the per-launch arguments of that path.
struct PeerMail { double val; unsigned long long seq; };
struct PeerArgs { PeerMail* const* peers; int nranks; int rank; unsigned long long seq; unsigned long long timeout_ns; };
// Peer landing (see k_land): peerR[r] = rank r's pair of reduced evaluation buffers, roff = offset of this evaluation's buffer.
struct LandPeer { double* const* peerR; PeerMail* const* peers; int nranks; int rank; unsigned long long seq; size_t roff; unsigned long long timeout_ns; };
// Wall-clock nanoseconds (independent of the SM clock): the peer waits are bounded in real time (gamc_comm_set_timeout).
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ double ld_peer(const double* q) { return *reinterpret_cast<const volatile double*>(q); }
// Element `off` of every rank's reduced buffer, added in rank order (identical bits on every rank).  The loads of up to eight
// ranks are issued together -- one NVLink round trip per batch instead of one per rank -- and only the additions are ordered.
__device__ __forceinline__ double peer_sum(double* const* peerR, int nr, size_t off) {
  double s = 0.0;
  for (int r0 = 0; r0 < nr; r0 += 8) {
    double t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = (r0 + u < nr) ? ld_peer(peerR[r0 + u] + off) : 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) if (r0 + u < nr) s += t[u];
  }
  return s;
}

// After k_reduce: tell every peer that this rank's reduced evaluation `seq` is in place (mailbox slots 2n .. 4n).
__global__ void k_peer_flag(PeerMail* const* peers, int nranks, int rank, unsigned long long seq) {
  if ((int)threadIdx.x < nranks) {
    __threadfence_system();
    volatile PeerMail* dst = peers[threadIdx.x] + 2 * nranks + (int)(seq & 1ull) * nranks + rank;
    dst->seq = seq;
  }
}

struct Bufs {
  DevState* st;
  double* R; int goff;                   // landing buffer of an evaluation = all-reduce payload: [ll, g(p), pad, G(p x p)]
  double* Gs; size_t Gstride;            // Gs[slot]: p x p  pstate.tensorlogtarget of the point held in the slot
  double* gs;                            // [2][p]           pstate.gradlogtarget
  double* U; size_t Ustride;             // U[slot]: p x p, upper triangle = reverse Cholesky factor of G[slot]
  double* rdiagU;                        // [2][p] reciprocal diagonal (M-space order)
  double* invU; size_t invUstride;       // [2][la_invd_doubles(p)] inverted 32 x 32 diagonal blocks of the factor (block solves)
  double* fterm;                         // [2][p] G^-1 grad ("firstterm")
  double* land;                          // [3p + 8] k_land output: per-column non-finite flags, q1 and q2 partials; [3p] log-target, [3p+1] gradient flag
  double* theta; double* theta_prop; double* mu; double* lastmean; double* secondlastmean;
  double* C;                             // p x p  sstate.oldinvtensor while in AM mode (AM covariance)
  double* Lc;                            // p x p  lower Cholesky factor of eps*C (am_mode 0) / of C (am_mode 1)
  double* rdiagL;                        // p  (am_mode 2: three buffers of p, like Lc)
  double* accbuf;                        // [2][p] unscaled proposal products L' z of the two speculative outcomes (am_mode 2)
  double* cand;                          // [2][p] candidate proposals of the NEXT AM step, one per outcome of this one (am_mode 3)
  size_t Lstride;                        // p*p: distance between the AM factor buffers
  double* W;                             // p x p  scratch: triangular inverse
  double* chain_value; unsigned char* chain_accept; long long chain_cap;
  double* chain_logtarget;               // [chain_cap] pstate.logtarget of every saved sample
  double* tune_log; int tune_log_cap;    // [cap][16] one record per burn-in tuning event (gamc_tune_log_get)
  const double* tape_usched; const double* tape_umix; const double* tape_z; const double* tape_uacc;
};

// shared memory of the step kernels: 4 vectors of p doubles + reduction scratch + LA scratch
struct StepSmem {
  double* v0; double* v1; double* v2; double* v3; double* red;
  LaScratch la;
};
__host__ __device__ inline size_t step_smem_bytes(int p) {
  const int pp = (p + 3) & ~3;
  return sizeof(double) * (4 * (size_t)pp + 40) + la_scratch_bytes(p);
}
__device__ inline StepSmem step_smem_carve(unsigned char* base, int p) {
  StepSmem s;
  const int pp = (p + 3) & ~3;
  double* d = reinterpret_cast<double*>(base);
  s.v0 = d; s.v1 = d + pp; s.v2 = d + 2 * pp; s.v3 = d + 3 * pp; s.red = d + 4 * pp;
  s.la = la_scratch_carve(reinterpret_cast<unsigned char*>(d + 4 * pp + 40), p);
  return s;
}

__device__ __forceinline__ void raise_error(DevState* st, int code, int pivot) {
  if (threadIdx.x == 0 && st->error_flag == 0) {
    st->error_flag = code; st->error_pivot = pivot; st->error_iter = st->count;
  }
}

// plogprior of examples/logistic_regression/src/GAMC/simulations.jl:30; all threads get the value.
__device__ inline double log_prior(const double* theta, int p, double lambda, double* red) {
  double s = 0.0;
  for (int i = threadIdx.x; i < p; i += blockDim.x) s = fma(theta[i], theta[i], s);
  s = block_sum(s, red);
  if (!(lambda > 0.0)) return 0.0;          // generic (callback) target: the host's values already contain the prior
  return -0.5 * (s / lambda + p * log(2.0 * 3.14159265358979323846 * lambda));
}

// Add the prior terms to a freshly reduced evaluation slot (after the all-reduce, once):
//   ll += logprior, g -= theta/lambda, G_ii += 1/lambda   (simulations.jl:30,32,36).  Returns the log-target.
__device__ inline double add_prior(double* E, int goff, const double* theta, int p, double lambda, double* red) {
  const double lp = log_prior(theta, p, lambda, red);
  for (int i = threadIdx.x; i < p; i += blockDim.x) {
    E[1 + i] -= theta[i] / lambda;
    E[goff + (size_t)i * p + i] += 1.0 / lambda;
  }
  __syncthreads();
  return E[0] + lp;
}

// ---------------------------------------------------------------------------------------------------
// k_land: move a landed (and, multi-GPU, all-reduced) level-2 evaluation R = [ll, g, G] into an evaluation slot, on the
// whole GPU (one CTA per column) instead of inside the single-CTA step kernels:
//   Gs[slot] <- G + I/lambda, U[slot] <- its upper triangle (the input of the reverse Cholesky), gs[slot] <- g - theta/lambda,
//   land[3p] <- ll + logprior(theta)                         (prior terms of simulations.jl:30,32,36, added once, after the reduction)
// and the O(p^2) pieces of the SMMALA acceptance ratio that do not depend on the factorisation (mode 2 only):
//   land[p + j]  = d_j sum_i d_i G(theta)_ij,   d = theta* - mu                (iterate/GAMC.jl:47-54)
//   land[2p + j] = a_j sum_i a_i G(theta*)_ij,  a = theta - theta*             (:60-67, see k_smmala_post)
// land[j] = 1 if column j of G has a non-finite entry.  Column partials are summed in a fixed order by the consumer.
// mode: 0 = initial point (slot 0), 1 = re-evaluation of the current point (slot cur), 2 = proposal (slot 1 - cur).
// ---------------------------------------------------------------------------------------------------
// Multi-GPU with peer landing (lp.nranks > 1): the kernel IS the all-reduce -- it waits for every rank's "reduced" flag
// and adds the ranks' buffers in rank order while landing them (loads over NVLink peer memory; identical bits on every
// rank; no NCCL call and no second pass over the payload).
constexpr int LAND_THREADS = 128;
__global__ void __launch_bounds__(LAND_THREADS) k_land(Bufs B, TuneCfg T, int mode, LandPeer lp) {
  __shared__ double red[40];
  DevState* st = B.st;
  if (st->error_flag) return;
  if (mode == 1 && T.reuse_tensor && st->tensor_valid) return;   // the evaluation was skipped: slot `cur` is still current
  const int p = T.p, j = blockIdx.x;
  const int nr = lp.nranks > 1 ? lp.nranks : 1;
  if (nr > 1) {
    if ((int)threadIdx.x < nr) {
      volatile PeerMail* src = lp.peers[lp.rank] + 2 * nr + (int)(lp.seq & 1ull) * nr + threadIdx.x;
      const unsigned long long t0 = globaltimer_ns();
      while (src->seq != lp.seq) {
        if (globaltimer_ns() - t0 > lp.timeout_ns) {   // the peer never published this evaluation (gamc_comm_set_timeout)
          if (st->error_flag == 0) { st->error_flag = 5 /*GAMC_NCCL_ERROR*/; st->error_pivot = threadIdx.x; st->error_iter = st->count; }
          break;
        }
        __nanosleep(100);
      }
      __threadfence_system();
    }
    __syncthreads();
  }
  const int cur = st->cur;
  const int slot = mode == 0 ? 0 : (mode == 1 ? cur : 1 - cur);
  const double* G = B.R + B.goff + (size_t)j * p;
  const size_t poff = lp.roff + B.goff + (size_t)j * p;
  const double* Gc = B.Gs + (size_t)cur * B.Gstride + (size_t)j * p;
  double* Gd = B.Gs + (size_t)slot * B.Gstride + (size_t)j * p;
  double* U = B.U + (size_t)slot * B.Ustride + (size_t)j * p;
  const double ilam = T.lambda > 0.0 ? 1.0 / T.lambda : 0.0;   // lambda <= 0: generic target, prior already included
  double nf = 0.0, q1 = 0.0, q2 = 0.0;
  for (int i = threadIdx.x; i < p; i += LAND_THREADS) {
    double v;
    if (nr > 1) v = peer_sum(lp.peerR, nr, poff + i);
    else v = G[i];
    if (i == j) v += ilam;
    Gd[i] = v;
    if (i <= j) U[i] = v;
    if (!isfinite(v)) nf = 1.0;
    if (mode == 2) {
      const double tp = B.theta_prop[i];
      q1 += (tp - B.mu[i]) * Gc[i];
      q2 += (B.theta[i] - tp) * v;
    }
  }
  nf = block_sum(nf, red);
  if (mode == 2) {
    q1 = block_sum(q1, red);
    q2 = block_sum(q2, red);
    if (threadIdx.x == 0) {
      const double tp = B.theta_prop[j];
      B.land[p + j] = q1 * (tp - B.mu[j]);
      B.land[2 * p + j] = q2 * (B.theta[j] - tp);
    }
  }
  if (threadIdx.x == 0) B.land[j] = nf != 0.0 ? 1.0 : 0.0;
  if (j == 0) {
    const double* th = mode == 2 ? B.theta_prop : B.theta;
    double gnf = 0.0;
    for (int i = threadIdx.x; i < p; i += LAND_THREADS) {
      double gl;
      if (nr > 1) gl = peer_sum(lp.peerR, nr, lp.roff + 1 + i);
      else gl = B.R[1 + i];
      const double g = gl - th[i] * ilam;
      B.gs[(size_t)slot * p + i] = g;
      if (!isfinite(g)) gnf = 1.0;
    }
    gnf = block_sum(gnf, red);
    const double lpr = log_prior(th, p, T.lambda, red);
    if (threadIdx.x == 0) {
      double ll;
      if (nr > 1) { ll = 0.0; for (int r = 0; r < nr; ++r) ll += ld_peer(lp.peerR[r] + lp.roff); }
      else ll = B.R[0];
      B.land[3 * p] = ll + lpr; B.land[3 * p + 1] = gnf;
    }
  }
}

// all threads: 1 if the landed evaluation (k_land) is finite everywhere
__device__ inline int landed_finite(const Bufs& B, int p, double* red) {
  double bad = 0.0;
  for (int i = threadIdx.x; i < p; i += blockDim.x) bad += B.land[i];
  bad = block_sum(bad, red);
  return bad == 0.0 && B.land[3 * p + 1] == 0.0 && isfinite(B.land[3 * p]);
}

// Factor the metric k_land left in a slot: U[slot] <- reverse Cholesky of G (in place), fterm[slot] <- G^-1 g.
// Returns 0 on success, the failed (M-space, 1-based) pivot otherwise.
__device__ inline int factor_slot(const Bufs& B, int slot, int p, StepSmem& S) {
  double* U = B.U + (size_t)slot * B.Ustride;
  const double* g = B.gs + (size_t)slot * p;
  double* invd = B.invU + (size_t)slot * B.invUstride;
  GAMC_PHASE(16);
  MatView<true> V{U, p, p};
  int bad = 0;
  const double sl = chol_blocked<true>(V, B.rdiagU + (size_t)slot * p, invd, S.la, &bad);
  GAMC_PHASE(17);
  if (bad) return bad;
  if (threadIdx.x == 0) B.st->sumlog[slot] = sl;
  // firstterm = G^-1 g = Lm^-T (Lm^-1 g')   in M space
  for (int i = threadIdx.x; i < p; i += blockDim.x) S.v3[i] = g[p - 1 - i];
  __syncthreads();
  trsv_lower<true>(V, invd, S.v3, S.la);
  GAMC_PHASE(18);
  trsv_lower_t<true>(V, invd, S.v3, S.la);
  for (int i = threadIdx.x; i < p; i += blockDim.x) B.fterm[(size_t)slot * p + i] = S.v3[p - 1 - i];
  __syncthreads();
  GAMC_PHASE(19);
  return 0;
}

// Bookkeeping shared by both branches: running means (iterate/GAMC.jl:193-197), burn-in tuning (:199-292),
// chain output (Klara run loop: save when count > burnin).
__device__ inline void step_tail(const Bufs& B, const TuneCfg& T, int accepted) {
  DevState* st = B.st;
  const int p = T.p;
  const long long t = st->count;
  for (int i = threadIdx.x; i < p; i += blockDim.x) {
    const double lm = B.lastmean[i];
    B.secondlastmean[i] = lm;                                           // :193
    B.lastmean[i] = ((double)(t - 1) * lm + B.theta[i]) / (double)t;    // recursive_mean! :195
  }
  if (t > T.burnin) {
    const long long pos = st->chain_saved;
    if (pos < B.chain_cap) {
      for (int i = threadIdx.x; i < p; i += blockDim.x) B.chain_value[(size_t)pos * p + i] = B.theta[i];
      if (threadIdx.x == 0) { B.chain_accept[pos] = (unsigned char)accepted; B.chain_logtarget[pos] = st->logtarget; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st->last_accept = accepted;
    if (t > T.burnin && st->chain_saved < B.chain_cap) st->chain_saved += 1;
    if (T.tuning) {
      if (st->tot_totprop <= T.burnin && (st->tot_prop % T.period) == 0) {
        if (T.count_total) st->tot_rate = (double)st->tot_acc / (double)st->tot_prop;       // rate! :205
        if (T.verbose_sm) {
          st->smfreq = (double)st->sm_prop / (double)st->tot_prop;
          st->sm_rate = (double)st->sm_acc / (double)st->sm_prop;
        }
        if (T.verbose_am) {
          st->amfreq = (double)st->am_prop / (double)st->tot_prop;
          st->am_rate = (double)st->am_acc / (double)st->am_prop;
        }
        if (T.is_accrate) {                                                                  // tune! :270-273
          st->step *= 2.0 / (1.0 + exp(-T.score_k * (st->tot_rate - T.targetrate)));
          st->sqrtstep = sqrt(st->step);
        }
        if (B.tune_log && st->tune_events < B.tune_log_cap) {       // what the verbose report prints (:218-268), before the resets
          double* r = B.tune_log + 16 * st->tune_events;
          r[0] = (double)t; r[1] = (double)st->tot_acc; r[2] = (double)st->tot_prop; r[3] = st->tot_rate;
          r[4] = (double)st->sm_acc; r[5] = (double)st->sm_prop; r[6] = st->sm_rate; r[7] = st->smfreq;
          r[8] = (double)st->am_acc; r[9] = (double)st->am_prop; r[10] = st->am_rate; r[11] = st->amfreq;
          r[12] = st->step; r[13] = r[14] = r[15] = 0.0;
        }
        st->tune_events += 1;
        if (T.verbose_sm) { st->sm_totprop += st->sm_prop; st->sm_prop = 0; st->sm_acc = 0; st->sm_rate = nan(""); }
        if (T.verbose_am) { st->am_totprop += st->am_prop; st->am_prop = 0; st->am_acc = 0; st->am_rate = nan(""); }
        st->tot_totprop += st->tot_prop;                                                     // :283
        st->tot_prop = 0;
        if (T.count_total) { st->tot_acc = 0; st->tot_rate = nan(""); }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Initialisation: `initialize!` + `sampler_state` (src/samplers/GAMC.jl:157-228) after eval(theta0) into slot 0.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LA_THREADS, 1) k_sampler_init(Bufs B, TuneCfg T) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StepSmem S = step_smem_carve(smem_raw, T.p);
  DevState* st = B.st;
  const int p = T.p;
  const double lt = B.land[3 * p];                  // k_land(mode 0) ran on the evaluation at theta0
  if (!landed_finite(B, p, S.red)) { raise_error(st, 2 /*GAMC_NONFINITE_INIT*/, 0); return; }
  const int bad = factor_slot(B, 0, p, S);
  if (bad) { raise_error(st, 3 /*GAMC_NOT_POSDEF*/, p + 1 - bad); return; }
  for (int i = threadIdx.x; i < p; i += blockDim.x) B.lastmean[i] = B.theta[i];   // GAMC.jl:218
  if (threadIdx.x == 0) { st->logtarget = lt; st->cur = 0; st->tensor_valid = 1; }
}

// ---------------------------------------------------------------------------------------------------
// SMMALA branch, part 1 (iterate/GAMC.jl:2-35): counters, optional re-factorisation at theta, proposal.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LA_THREADS, 1) k_smmala_pre(Bufs B, TuneCfg T, long long tape_idx, int refactor) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StepSmem S = step_smem_carve(smem_raw, T.p);
  DevState* st = B.st;
  const int p = T.p;
  if (st->error_flag) return;
  const int cur = st->cur;
  if (threadIdx.x == 0) {
    st->count += 1;                                   // :2
    if (T.tuning) st->tot_prop += 1;                  // :4-6
    st->updatetensorcount += 1;                       // :13
    if (T.verbose_sm) st->sm_prop += 1;               // :15-17
  }
  GAMC_PHASE_BEGIN();
  const int still_valid = T.reuse_tensor && st->tensor_valid;   // gamc_config.reuse_tensor: the re-evaluation was skipped on the device
  __syncthreads();
  if (refactor && !still_valid) {                     // :19-31 (previous step was AM: grad/metric stale)
    if (threadIdx.x == 0) st->logtarget = B.land[3 * p];   // k_land(mode 1) ran on the re-evaluation at theta
    const int bad = factor_slot(B, cur, p, S);
    if (bad) { raise_error(st, 3, p + 1 - bad); return; }
    if (threadIdx.x == 0) st->tensor_valid = 1;
  }
  __syncthreads();
  GAMC_PHASE(refactor ? 32 : 33);
  const double eps = st->step, sq = st->sqrtstep;
  // y' = Lm^-T z'  (cholinvtensor * randn(p), :35)
  const double* z = B.tape_z + (size_t)tape_idx * p;
  for (int i = threadIdx.x; i < p; i += blockDim.x) S.v0[i] = z[p - 1 - i];
  __syncthreads();
  MatView<true> V{B.U + (size_t)cur * B.Ustride, p, p};
  trsv_lower_t<true>(V, B.invU + (size_t)cur * B.invUstride, S.v0, S.la);
  for (int i = threadIdx.x; i < p; i += blockDim.x) {
    const double m = B.theta[i] + 0.5 * eps * B.fterm[(size_t)cur * p + i];   // :33
    B.mu[i] = m;
    B.theta_prop[i] = m + sq * S.v0[p - 1 - i];                                // :35
  }
  GAMC_PHASE(34);
}

// ---------------------------------------------------------------------------------------------------
// SMMALA branch, part 2 (iterate/GAMC.jl:37-127 + tail): ratio, accept/reject, state swap.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LA_THREADS, 1) k_smmala_post(Bufs B, TuneCfg T, long long tape_idx) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StepSmem S = step_smem_carve(smem_raw, T.p);
  DevState* st = B.st;
  const int p = T.p;
  if (st->error_flag) return;
  const int cur = st->cur, prop = 1 - cur;
  const double eps = st->step;
  GAMC_PHASE_BEGIN();
  const double lt_p = B.land[3 * p];                 // k_land(mode 2) ran on the evaluation at theta*
  int ok = landed_finite(B, p, S.red);               // non-finite target / gradient / metric at the proposal: reject
  GAMC_PHASE(1);
  double ratio = -INFINITY;
  int bad = 0;
  if (ok) bad = factor_slot(B, prop, p, S);          // newinvtensor / newfirstterm (:43,56) via one factorisation
  if (bad > 0) { raise_error(st, 3, p + 1 - bad); return; }
  if (ok) {
    GAMC_PHASE(2);
    // q1 = d' G d, d = theta* - mu (:47-54): column partials from k_land, summed in a fixed order
    double t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0;
    for (int i = threadIdx.x; i < p; i += blockDim.x) {
      t1 += B.land[p + i];
      t2 += B.land[2 * p + i];
      // mu* = theta* + eps/2 f*, d2 = theta - mu* = a - eps/2 f* with a = theta - theta*; since G* f* = g*:
      //   q2 = d2' G* d2 = a' G* a - eps a' g* + eps^2/4 f*' g*                         (:58-67)
      const double gp = B.gs[(size_t)prop * p + i];
      t3 = fma(B.theta[i] - B.theta_prop[i], gp, t3);
      t4 = fma(B.fterm[(size_t)prop * p + i], gp, t4);
    }
    const double q1 = block_sum(t1, S.red);
    const double aGa = block_sum(t2, S.red);
    const double ag = block_sum(t3, S.red);
    const double fg = block_sum(t4, S.red);
    const double q2 = aGa - eps * ag + 0.25 * eps * eps * fg;
    GAMC_PHASE(4);
    const double plog = p * log(eps);
    const double ld_old = plog - 2.0 * st->sumlog[cur];   // logdet(eps * G^-1)
    const double ld_new = plog - 2.0 * st->sumlog[prop];
    ratio = lt_p - st->logtarget;                                                       // :45
    ratio += 0.5 * (ld_old + q1 / eps);
    ratio -= 0.5 * (ld_new + q2 / eps);
  }
  const double u = B.tape_uacc[tape_idx];
  const int accept = (ratio > 0.0) || (ratio > log(u));                                 // :69
  __syncthreads();
  if (accept) {
    for (int i = threadIdx.x; i < p; i += blockDim.x) B.theta[i] = B.theta_prop[i];     // :70
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st->ratio = ratio;
    if (accept) {
      st->logtarget = lt_p;                                                             // :98
      st->cur = prop;                          // grad, tensor, invtensor, chol, firstterm swap (:72-96)
      if (T.count_total) st->tot_acc += 1;                                              // :114-116
      if (T.verbose_sm) st->sm_acc += 1;
    }
  }
  __syncthreads();
  step_tail(B, T, accept);
  GAMC_PHASE(5);
}

// ---------------------------------------------------------------------------------------------------
// AM branch, part 1 (iterate/GAMC.jl:128-146): covariance recursion, proposal from the two-component GMM.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LA_THREADS, 1) k_am_pre(Bufs B, TuneCfg T, long long tape_idx) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StepSmem S = step_smem_carve(smem_raw, T.p);
  DevState* st = B.st;
  const int p = T.p;
  if (st->error_flag) return;
  if (threadIdx.x == 0) {
    st->count += 1;
    if (T.tuning) st->tot_prop += 1;
    if (T.verbose_am) st->am_prop += 1;               // :129-131
  }
  __syncthreads();
  const long long t = st->count;
  const double k = (double)(t - 2);
  const double eps = st->step;
  // Klara covariance!(C, C, k, x, lastmean, secondlastmean)  (:133-140), then Hermitian (:142):
  //   C = ((k-1) C + k m2 m2' - (k+1) m1 m1' + x x')/k
  for (int i = threadIdx.x; i < p; i += blockDim.x) { S.v0[i] = B.theta[i]; S.v1[i] = B.lastmean[i]; S.v2[i] = B.secondlastmean[i]; }
  __syncthreads();
  const size_t n = (size_t)p * p;
  for (size_t e = threadIdx.x; e < n; e += blockDim.x) {
    const int i = (int)(e % p), j = (int)(e / p);
    if (i >= j) {
      const double c = ((k - 1.0) * B.C[e] + k * (S.v2[i] * S.v2[j]) - (k + 1.0) * (S.v1[i] * S.v1[j]) + S.v0[i] * S.v0[j]) / k;
      B.C[e] = c;
      B.C[(size_t)j + (size_t)i * p] = c;
      B.Lc[e] = eps * c;                              // MvNormal(theta, eps*C) -> Cholesky (set_gmm, GAMC.jl:188); am_mode 0 uses buffer 0 only
    }
  }
  __syncthreads();
  MatView<false> V{B.Lc, p, p};
  int bad = 0;
  chol_blocked<false>(V, B.rdiagL, nullptr, S.la, &bad);
  if (bad) { raise_error(st, 3, bad); return; }
  const double* z = B.tape_z + (size_t)tape_idx * p;
  for (int i = threadIdx.x; i < p; i += blockDim.x) S.v3[i] = z[i];
  __syncthreads();
  const double umix = B.tape_umix[tape_idx];
  if (umix <= 1.0 - T.c) {                            // rand(MixtureModel): core component (:146)
    trmv_lower(B.Lc, p, p, S.v3, S.v1);
    __syncthreads();
    for (int i = threadIdx.x; i < p; i += blockDim.x) B.theta_prop[i] = S.v0[i] + S.v1[i];
  } else {
    for (int i = threadIdx.x; i < p; i += blockDim.x) B.theta_prop[i] = S.v0[i] + T.sqrtminorscale * S.v3[i];
  }
}

// ---------------------------------------------------------------------------------------------------
// AM branch, part 1, fast path (am_mode = 1).  With m1 = (k m2 + x)/(k+1) (how lastmean is maintained,
// iterate/GAMC.jl:193-195) Klara's three-term recursion is algebraically the rank-one update
//     C_k = ((k-1)/k) C_{k-1} + u u',   u = (x - m2)/sqrt(k+1),
// so the factor the proposal needs is  chol(C_k) = (a L) chol(I + w w'),  a = sqrt((k-1)/k), (a L) w = u,
// and chol(I + w w') has the closed form (s_j = 1 + sum_{i<j} w_i^2)
//     K_jj = sqrt(s_{j+1}/s_j),   K_ij = w_i w_j / sqrt(s_j s_{j+1})  (i > j),
// which gives, with the running partial products P_ij = sum_{k<=j} a L_ik w_k (the same quantity the forward
// substitution for w needs),
//     L'_ij = a L_ij K_jj + g_j (u_i - P_ij),   g_j = w_j / sqrt(s_j s_{j+1}).
// One ascending sweep over 32-column panels does the triangular solve, the prefix sums, the O(p^2) update and
// the proposal product L' z: no sqrt/divide on a dependent chain, L read once and written once.  Panels are
// staged in shared memory with cp.async (double-buffered) so the single CTA is bound by throughput, not by
// memory latency.  B.Lc holds chol(C) (unscaled by eps; the proposal is theta + sqrt(eps) L z).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__host__ __device__ inline size_t am_r1_smem_bytes(int p, int nbuf) {
  const int pp = (p + 3) & ~3;
  return sizeof(double) * ((size_t)nbuf * 32 * pp + 2 * (size_t)pp + 4 * 32 + 2 * 32 * 33 + 32 + 64);
}
__host__ __device__ inline int am_r1_nbuf(int p) { return am_r1_smem_bytes(p, 2) <= 200 * 1024 ? 2 : 1; }

// spec = 0: the step itself (in place on the current factor buffer; bumps the counters; writes the proposal).
// spec = 1 (am_mode 2): SPECULATIVE update for the NEXT AM step, launched on a side stream while the data pass of the
// current step runs: CTA o in {0, 1} assumes the current proposal is accepted (o = 0) or rejected (o = 1), reads the
// current factor, writes the updated factor to buffer (lcur + 1 + o) % 3 and the unscaled product L' z to accbuf[o];
// k_am_post picks the outcome that happened.  Same arithmetic as spec = 0, so the chain is bit-identical.
// spec = 2 (am_mode 3): as spec = 1, on the main stream BEFORE the data pass, and also writes the two candidate proposals of the next
// step (B.cand) so that the pass can evaluate the log-likelihood at them together with this step's proposal.
__global__ void __launch_bounds__(LA_THREADS, 1) k_am_pre_r1(Bufs B, TuneCfg T, long long tape_idx, int nbuf, int spec) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  DevState* st = B.st;
  const int p = T.p, pp = (p + 3) & ~3;
  if (st->error_flag) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int outcome = spec ? (int)blockIdx.x : 0;
  const int lcur = st->lcur;
  const double* Lin = B.Lc + (size_t)lcur * B.Lstride;
  double* Lout = spec ? B.Lc + (size_t)((lcur + 1 + outcome) % 3) * B.Lstride : B.Lc + (size_t)lcur * B.Lstride;
  const double* rdin = B.rdiagL + (size_t)lcur * p;
  double* rdout = spec ? B.rdiagL + (size_t)((lcur + 1 + outcome) % 3) * p : B.rdiagL + (size_t)lcur * p;
  const double* xsrc = (spec && outcome == 0) ? B.theta_prop : B.theta;          // the state the next step would start from
  const double* m2src = spec ? B.lastmean : B.secondlastmean;                     // secondlastmean of the NEXT step = lastmean now
  double* panel = reinterpret_cast<double*>(smem_raw);          // [nbuf][32][pp]  (column-major: element (row, jj) at jj*pp + row)
  double* zs = panel + (size_t)nbuf * 32 * pp;                   // z
  double* bw = zs + 2 * (size_t)pp;                              // per-block: w[32], Kd[32], g[32], sigma
  double* Ls = bw + 4 * 32;                                      // [32][33] a*L of the diagonal block (transposed scratch)
  double* Ts = Ls + 32 * 33;                                     // [32][33] residuals u_i - P_ij of the diagonal block
  double* accd = Ts + 32 * 33;                                   // [32] proposal-product contributions of the diagonal block
  int* flag = reinterpret_cast<int*>(accd + 32);
  if (tid == 0) {
    if (!spec) {
      st->count += 1;
      if (T.tuning) st->tot_prop += 1;
      if (T.verbose_am) st->am_prop += 1;
    }
    flag[0] = 0;
  }
  const int nblk = (p + 31) >> 5;
  // cp.async staging of panel b by the threads [t0, t0 + nt) (each thread waits for its own groups)
  auto issue_panel = [&](int b, int t, int nt) {
    const int j0 = b << 5, nb = min(32, p - j0), rows = p - j0;
    double* dst = panel + (size_t)(b % nbuf) * 32 * pp;
    const double* src = Lin + (size_t)j0 + (size_t)j0 * p;
    if ((p & 1) == 0) {   // 16-byte copies: column starts are 16-byte aligned when p is even (j0 is a multiple of 32)
      for (int jj = 0; jj < nb; ++jj)
        for (int r = 2 * t; r < rows; r += 2 * nt) cp_async16(dst + (size_t)jj * pp + r, src + (size_t)jj * p + r);
    } else {
      for (int jj = 0; jj < nb; ++jj)
        for (int r = t; r < rows; r += nt) cp_async8(dst + (size_t)jj * pp + r, src + (size_t)jj * p + r);
    }
  };
  issue_panel(0, tid, LA_THREADS);
  cp_async_commit();
  __syncthreads();
  const long long t = st->count + (spec ? 1 : 0);                             // iteration the update belongs to
  const double k = (double)(t - 2);
  if (!(k > 1.0)) {                                                            // C_1 = u u' has rank one: the reference's chol throws
    cp_async_wait<0>();
    if (spec) { if (tid == 0) st->spec_err[outcome] = 1; } else raise_error(st, 3, 1);
    return;
  }
  const double alpha = sqrt((k - 1.0) / k), beta = 1.0 / sqrt(k + 1.0), ralpha = 1.0 / alpha;
  const double* z = B.tape_z + (size_t)tape_idx * p;
  const int i = tid;
  double th = 0.0, ui = 0.0, P = 0.0, acc = 0.0, rdv = 0.0;
  if (i < p) {
    th = xsrc[i];
    ui = (th - m2src[i]) * beta;
    rdv = rdin[i] * ralpha;              // 1 / (a L_ii)
    zs[i] = z[i];
  }
  double sigma = 1.0;                     // s_{j0}: carried by every thread identically
  for (int b = 0; b < nblk; ++b) {
    const int j0 = b << 5, nb = min(32, p - j0);
    cp_async_wait<0>();
    __syncthreads();                      // panel b has landed; everybody is done with block b-1
    const double* pan = panel + (size_t)(b % nbuf) * 32 * pp;
    if (warp == b) {
      // ---- diagonal block (one warp, lane = row j0+lane): forward substitution for w.
      // Dependent chain per column = mul -> shfl -> fma.  The residual after column jj is exactly
      // u_i - P_i,jj, the factor the update needs: park it (and a*L) in transposed scratch for the helpers.
      double lrow[32];
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        lrow[jj] = (jj <= lane && jj < nb && i < p) ? alpha * pan[(size_t)jj * pp + lane] : 0.0;
        Ls[jj * 33 + lane] = lrow[jj];
      }
      double res = ui - P, wmine = 0.0;
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        if (jj < nb) {
          const double wj = __shfl_sync(0xffffffffu, res * rdv, jj);          // w_j = (u_j - P_j)/(a L_jj)
          if (lane == jj) wmine = wj;
          res = fma(-lrow[jj], wj, res);                                      // lrow[jj] = 0 for jj > lane
          Ts[jj * 33 + lane] = res;
        }
      }
      P = ui - res;
      // prefix sums s_j and the closed-form factor entries, one column per lane (no serial sqrt/divide)
      double sc = wmine * wmine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const double n = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += n; }
      const double incl = sigma + sc, excl = incl - wmine * wmine;
      const double rs = rsqrt(excl * incl);
      bw[lane] = wmine; bw[32 + lane] = incl * rs; bw[64 + lane] = wmine * rs;
      const double snew = __shfl_sync(0xffffffffu, incl, nb - 1);
      if (lane == 0) bw[96] = snew;
    } else if (nbuf > 1 && b + 1 < nblk) {
      // ---- meanwhile the other warps stage the next panel
      const int t2 = (warp < b ? warp : warp - 1) * 32 + lane;
      issue_panel(b + 1, t2, LA_THREADS - 32);
    }
    cp_async_commit();
    __syncthreads();
    sigma = bw[96];
    if (i >= j0 + 32 && i < p) {
      // ---- rows below the diagonal block: one thread per row, residual res = u_i - P_i carried in a register
      const double* prow = pan + (i - j0);
      double* out = Lout + (size_t)i + (size_t)j0 * p;
      double res = ui - P, a0 = 0.0, a1 = 0.0;
      if (nb == 32) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const double l = alpha * prow[(size_t)jj * pp];
          res = fma(-l, bw[jj], res);
          const double ln = fma(l, bw[32 + jj], bw[64 + jj] * res);
          out[(size_t)jj * p] = ln;
          if (jj & 1) a1 = fma(ln, zs[j0 + jj], a1); else a0 = fma(ln, zs[j0 + jj], a0);
        }
      } else {
        for (int jj = 0; jj < nb; ++jj) {
          const double l = alpha * prow[(size_t)jj * pp];
          res = fma(-l, bw[jj], res);
          const double ln = fma(l, bw[32 + jj], bw[64 + jj] * res);
          out[(size_t)jj * p] = ln;
          a0 = fma(ln, zs[j0 + jj], a0);
        }
      }
      P = ui - res;
      acc += a0 + a1;
    }
    {
      // ---- the diagonal block itself, spread over all warps (row r per warp, lane = column):
      // L'_rj = l K_jj + g_j (u_r - P_rj) for j <= r, independent per element.
      for (int r = warp; r < nb; r += LA_THREADS / 32) {
        const int gi = j0 + r;
        double contrib = 0.0;
        if (lane <= r && lane < nb) {
          const double ln = fma(Ls[lane * 33 + r], bw[32 + lane], bw[64 + lane] * Ts[lane * 33 + r]);
          Lout[(size_t)gi + (size_t)(j0 + lane) * p] = ln;
          contrib = ln * zs[j0 + lane];
          if (lane == r) rdout[gi] = 1.0 / ln;
        }
        contrib = warp_sum(contrib);
        if (lane == 0) accd[r] = contrib;
      }
    }
    __syncthreads();
    if (warp == b && i < p) acc += accd[lane];
    if (nbuf == 1 && b + 1 < nblk) { issue_panel(b + 1, tid, LA_THREADS); cp_async_commit(); }   // committed here: wait_group at the loop top only covers committed groups
  }
  const int fin = isfinite(sigma) ? 0 : 1;
  if (__syncthreads_or(fin)) {
    if (spec) { if (tid == 0) st->spec_err[outcome] = 1; } else raise_error(st, 3, 1);
    return;
  }
  if (spec) {
    if (i < p) {
      B.accbuf[(size_t)outcome * p + i] = acc;                              // scaled by sqrt(eps) once the outcome is known
      if (spec == 2) {
        // am_mode 3: the candidate proposal itself, exactly as k_am_post will form it after the decision (the host pairs two AM
        // steps only when the first is not a tuning event, so sqrt(eps) is the one the next iterate! sees)
        const double umix = B.tape_umix[tape_idx];
        B.cand[(size_t)outcome * p + i] = (umix <= 1.0 - T.c) ? fma(st->sqrtstep, acc, th) : fma(T.sqrtminorscale, zs[i], th);   // (explicit fma: the three places that form an AM proposal must round alike)
      }
    }
    if (tid == 0) st->spec_err[outcome] = 0;
    return;
  }
  if (i < p) {
    const double umix = B.tape_umix[tape_idx];
    if (umix <= 1.0 - T.c) B.theta_prop[i] = fma(st->sqrtstep, acc, th);    // core component  N(theta, eps C)
    else B.theta_prop[i] = fma(T.sqrtminorscale, zs[i], th);                // stabilising component
  }
}

// chol(inv(G)) = U^-T as a natural-order lower factor:  L(a,b) = W(p-1-b, p-1-a), W = Lm^-1 (k_triinv).
__global__ void __launch_bounds__(256) k_w_to_l(DevState* st, const double* __restrict__ W, int p, double* __restrict__ L, double* __restrict__ rdiagL) {
  if (blockIdx.x == 0 && threadIdx.x == 0) st->lcur = 0;       // the seeded factor lives in buffer 0
  const size_t n = (size_t)p * p;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int a = (int)(e % p), b = (int)(e / p);
    const double v = (a >= b) ? W[(size_t)(p - 1 - b) + (size_t)(p - 1 - a) * p] : 0.0;
    L[e] = v;
    if (a == b) rdiagL[a] = 1.0 / v;
  }
}

// C = s * L L' (state read-out of sstate.oldinvtensor in am_mode 1); one thread per entry.
__global__ void __launch_bounds__(256) k_llt(const double* __restrict__ L, int p, double* __restrict__ C) {
  const size_t n = (size_t)p * p;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int a = (int)(e % p), b = (int)(e / p);
    const int m = min(a, b);
    double s = 0.0;
    for (int k = 0; k <= m; ++k) s = fma(L[(size_t)a + (size_t)k * p], L[(size_t)b + (size_t)k * p], s);
    C[e] = s;
  }
}

// ---------------------------------------------------------------------------------------------------
// AM branch, part 2 (iterate/GAMC.jl:148-191 + tail).  The two GMM log-densities (:152,156) use the same
// covariance with the roles of theta and theta* exchanged; d -> -d gives bit-identical quadratic forms and
// log-sum-exp, so the two terms are the same number a and the device omits them.  The reference forms
// (r - a) + a, which equals r up to ONE rounding of the ratio (not bit for bit); accept decisions agree in
// every tested run (the oracle evaluates both terms literally; tests/test_oracle_sampler.py asserts a is
// bit-identical on both sides).
// ---------------------------------------------------------------------------------------------------
// next_spec = 1 (am_mode 2): the next step is an AM step whose factor update was computed speculatively for both outcomes
// while the data pass ran; after the decision this kernel selects the factor buffer of the outcome that happened and
// performs the head of the next `iterate!` (counters, proposal) so that the next data pass can start immediately.
// part_ll / nparts: the per-CTA log-likelihood partials of the data pass are summed here, in k_reduce's order (bitwise the
// same value), so the AM step launches neither k_reduce nor NCCL; nparts = 0 reads the (all-)reduced R[0] instead.
// pa (multi-GPU): one-shot all-reduce of that scalar over NVLink peer memory -- every rank stores (value, sequence number)
// into its slot of every peer's mailbox, waits until its own mailbox holds this step's sequence number from every rank,
// and adds the values in rank order: the same bits on every rank.  Slots alternate with the parity of the sequence
// number (a rank can be at most one step ahead of the slowest one).
// pair_second = 1 (am_mode 3): second AM step of a pair.  Its proposal was written by the first step's k_am_post; the data pass
// of the pair has already evaluated the log-likelihood at both candidates: part_ll belongs to the candidate of an ACCEPTED first
// step, part_alt to the candidate of a rejected one (st->last_accept tells which).  The candidate is compared with the proposal
// bit for bit: a difference (the host paired across a tuning event) raises an error instead of using a stale value.
__global__ void __launch_bounds__(LA_THREADS, 1) k_am_post(Bufs B, TuneCfg T, long long tape_idx, int next_spec,
                                                           const double* __restrict__ part_ll, int nparts, PeerArgs pa,
                                                           const double* __restrict__ part_alt, int pair_second) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StepSmem S = step_smem_carve(smem_raw, T.p);
  DevState* st = B.st;
  const int p = T.p;
  if (st->error_flag) return;
  if (pair_second) {
    const int o = st->last_accept ? 0 : 1;
    if (o) part_ll = part_alt;
    int bad = 0;
    for (int i = threadIdx.x; i < p; i += blockDim.x) bad |= (B.theta_prop[i] != B.cand[(size_t)o * p + i]);
    if (__syncthreads_or(bad)) { raise_error(st, 4 /*GAMC_CUDA_ERROR: internal*/, 0); return; }
  }
  // everything this kernel reads later, requested now: one L2 round trip instead of five dependent ones
  for (int i = threadIdx.x * 16; i < p; i += blockDim.x * 16) {      // one request per 128-byte line
    prefetch_l1(B.theta_prop + i); prefetch_l1(B.theta + i); prefetch_l1(B.lastmean + i);
    if (next_spec) { prefetch_l1(B.accbuf + i); prefetch_l1(B.accbuf + p + i); prefetch_l1(B.tape_z + (size_t)(tape_idx + 1) * p + i); }
  }
  if (threadIdx.x == 0) { prefetch_l1(B.tape_uacc + tape_idx); if (next_spec) prefetch_l1(B.tape_umix + tape_idx + 1); }
  if (nparts > 0) {
    if (threadIdx.x < 32) {
      double s = 0.0;
      for (int k = threadIdx.x; k < nparts; k += 32) s += part_ll[k];
      s = warp_sum(s);
      if (threadIdx.x == 0) S.red[39] = s;
    }
    __syncthreads();
    if (pa.nranks > 1) {
      const double mine = S.red[39];
      const int par = (int)(pa.seq & 1ull);
      if ((int)threadIdx.x < pa.nranks) {
        volatile PeerMail* dst = pa.peers[threadIdx.x] + par * pa.nranks + pa.rank;
        dst->val = mine;
        __threadfence_system();
        dst->seq = pa.seq;
        volatile PeerMail* src = pa.peers[pa.rank] + par * pa.nranks + threadIdx.x;
        const unsigned long long t0 = globaltimer_ns();
        bool got = true;
        while (src->seq != pa.seq) {
          if (globaltimer_ns() - t0 > pa.timeout_ns) { got = false; break; }   // the peer never reached this step (gamc_comm_set_timeout)
          __nanosleep(100);
        }
        __threadfence_system();
        S.red[threadIdx.x] = got ? src->val : nan("");
        if (!got && st->error_flag == 0) { st->error_flag = 5 /*GAMC_NCCL_ERROR*/; st->error_pivot = threadIdx.x; st->error_iter = st->count; }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0.0;
        for (int r = 0; r < pa.nranks; ++r) s += S.red[r];
        S.red[39] = s;
      }
      __syncthreads();
    }
  }
  const double ll_p = nparts > 0 ? S.red[39] : B.R[0];
  const double lt_p = ll_p + log_prior(B.theta_prop, p, T.lambda, S.red);
  double ratio = isfinite(lt_p) ? (lt_p - st->logtarget) : -INFINITY;                   // :150
  const double u = B.tape_uacc[tape_idx];
  const int accept = (ratio > 0.0) || (ratio > log(u));                                 // :158
  __syncthreads();
  if (accept) {
    for (int i = threadIdx.x; i < p; i += blockDim.x) B.theta[i] = B.theta_prop[i];     // :159
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st->ratio = ratio;
    if (accept) {
      st->logtarget = lt_p;                                                             // :161
      st->tensor_valid = 0;                    // gradient / metric of the slot now belong to an older point
      if (T.count_total) st->tot_acc += 1;
      if (T.verbose_am) st->am_acc += 1;
    }
  }
  __syncthreads();
  step_tail(B, T, accept);
  if (next_spec) {
    __syncthreads();
    const int o = accept ? 0 : 1;
    if (st->spec_err[o]) { raise_error(st, 3, 1); return; }
    const long long tn = tape_idx + 1;
    const double umix = B.tape_umix[tn];
    const double* z = B.tape_z + (size_t)tn * p;
    const double sq = st->sqrtstep;                                           // after this step's tuning, as the next iterate! sees it
    for (int i = threadIdx.x; i < p; i += blockDim.x) {
      const double th = B.theta[i];
      B.theta_prop[i] = (umix <= 1.0 - T.c) ? fma(sq, B.accbuf[(size_t)o * p + i], th) : fma(T.sqrtminorscale, z[i], th);
    }
    if (threadIdx.x == 0) {
      st->lcur = (st->lcur + 1 + o) % 3;
      st->count += 1;                                                         // head of the next iterate! (iterate/GAMC.jl:2-6,129-131)
      if (T.tuning) st->tot_prop += 1;
      if (T.verbose_am) st->am_prop += 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// MALA (identity metric): Klara's comparison sampler, the baseline of the reference's efficiency tables
// (examples/logistic_regression/src/MALA/simulations.jl:37-41, examples/logistic_regression/src/table.jl:9-17).  Needs only
// the fused log-likelihood + gradient pass (no metric, no factorisation):
//   mu = theta + eps/2 grad;  theta* = mu + sqrt(eps) z;  mu* = theta* + eps/2 grad*
//   log alpha = lt* - lt + |theta* - mu|^2/(2 eps) - |theta - mu*|^2/(2 eps)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mala_init(Bufs B, TuneCfg T) {
  __shared__ double red[40];
  DevState* st = B.st;
  const int p = T.p;
  const double lp = log_prior(B.theta, p, T.lambda, red);
  double bad = 0.0;
  for (int i = threadIdx.x; i < p; i += blockDim.x) {
    const double g = B.R[1 + i] - (T.lambda > 0.0 ? B.theta[i] / T.lambda : 0.0);
    B.gs[i] = g;
    B.lastmean[i] = B.theta[i];
    if (!isfinite(g)) bad = 1.0;
  }
  bad = block_sum(bad, red);
  const double lt = B.R[0] + lp;
  if (threadIdx.x == 0) {
    st->logtarget = lt; st->cur = 0;
    if (bad != 0.0 || !isfinite(lt)) { st->error_flag = 2; st->error_iter = 0; }
  }
}

__global__ void __launch_bounds__(256) k_mala_pre(Bufs B, TuneCfg T, long long tape_idx) {
  DevState* st = B.st;
  const int p = T.p;
  if (st->error_flag) return;
  if (threadIdx.x == 0) { st->count += 1; if (T.tuning) st->tot_prop += 1; }
  const double eps = st->step, sq = st->sqrtstep;
  const double* z = B.tape_z + (size_t)tape_idx * p;
  for (int i = threadIdx.x; i < p; i += blockDim.x) {
    const double m = B.theta[i] + 0.5 * eps * B.gs[i];
    B.mu[i] = m;
    B.theta_prop[i] = m + sq * z[i];
  }
}

__global__ void __launch_bounds__(256) k_mala_post(Bufs B, TuneCfg T, long long tape_idx) {
  __shared__ double red[40];
  DevState* st = B.st;
  const int p = T.p;
  if (st->error_flag) return;
  const double eps = st->step;
  const double lt_p = B.R[0] + log_prior(B.theta_prop, p, T.lambda, red);
  double q1 = 0.0, q2 = 0.0, bad = 0.0;
  for (int i = threadIdx.x; i < p; i += blockDim.x) {
    const double gp = B.R[1 + i] - (T.lambda > 0.0 ? B.theta_prop[i] / T.lambda : 0.0);
    B.gs[p + i] = gp;                                     // proposal-slot gradient
    if (!isfinite(gp)) bad = 1.0;
    const double d1 = B.theta_prop[i] - B.mu[i];
    const double d2 = B.theta[i] - (B.theta_prop[i] + 0.5 * eps * gp);
    q1 = fma(d1, d1, q1); q2 = fma(d2, d2, q2);
  }
  q1 = block_sum(q1, red); q2 = block_sum(q2, red); bad = block_sum(bad, red);
  double ratio = -INFINITY;
  if (bad == 0.0 && isfinite(lt_p)) ratio = lt_p - st->logtarget + q1 / (2.0 * eps) - q2 / (2.0 * eps);
  const int accept = (ratio > 0.0) || (ratio > log(B.tape_uacc[tape_idx]));
  __syncthreads();
  if (accept) for (int i = threadIdx.x; i < p; i += blockDim.x) { B.theta[i] = B.theta_prop[i]; B.gs[i] = B.gs[p + i]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    st->ratio = ratio;
    if (accept) { st->logtarget = lt_p; if (T.count_total) st->tot_acc += 1; }
  }
  __syncthreads();
  step_tail(B, T, accept);
}

// ---------------------------------------------------------------------------------------------------
// SMMALA -> AM hand-over: sstate.oldinvtensor = inv(G) seeds the AM covariance recursion
// (iterate/GAMC.jl:26,92,133).  W = Lm^-1 (M space), then inv(G) = W' W.
// Blocked by 32: one CTA per block column J, W_JJ = Y_J (the inverted diagonal blocks chol_blocked left behind) and
//   W_IJ = -Y_I * sum_{K=J}^{I-1} L_IK W_KJ     for I = J+1 ..,
// every product on the FP64 tensor cores (8 warps x 2 DMMA 8x8 tiles); block column J of W stays in shared memory and
// the next L block is fetched into registers while the current one is multiplied.
// ---------------------------------------------------------------------------------------------------
constexpr int TI_LD = 36;                 // leading dimension of the 32 x 32 operand tiles in shared memory (4 mod 16: conflict-free fragments)
constexpr int TI_TILE = 32 * TI_LD;
__host__ __device__ inline size_t triinv_smem_bytes(int p) { return sizeof(double) * (size_t)((p + 31) / 32 + 3) * TI_TILE; }

__global__ void __launch_bounds__(256) k_triinv(const DevState* st, const double* Ubase, size_t Ustride,
                                                const double* invBase, size_t invStride, int p, double* W) {
  extern __shared__ __align__(16) double ti_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (p + 31) / 32, J = blockIdx.x;
  const int cur = st->cur;
  MatView<true> L{const_cast<double*>(Ubase) + (size_t)cur * Ustride, p, p};
  const double* invd = invBase + (size_t)cur * invStride;
  double* Wsm = ti_smem;                          // [I - J][k][n]: block (I, J) of W, row k, column n
  double* Asm = ti_smem + (size_t)nblk * TI_TILE; // [k][r]: current left operand (L_IK or Y_I), column k, row r
  double* Ssm = Asm + TI_TILE;                    // [k][n]: S = sum_K L_IK W_KJ
  const int fr = lane >> 2, fk = lane & 3;
  const int mi = warp >> 1, n0 = (warp & 1) * 2;  // this warp's tiles: (mi, n0) and (mi, n0 + 1)
  // W_JJ = Y_J: layout A of invd is [j * 32 + i] = inv(i, j)
  for (int e = tid; e < 1024; e += 256) {
    const int i = e & 31, j = e >> 5;
    const double v = invd[(size_t)J * LA_INVD_STRIDE + j * 32 + i];
    Wsm[i * TI_LD + j] = v;
    const int gi = 32 * J + i, gj = 32 * J + j;
    if (gi < p && gj < p) W[(size_t)gi + (size_t)gj * p] = v;
  }
  double pre[4];
  auto fetch_L = [&](int I, int K) {              // L_IK -> registers, element (r, k) with r = e % 32 (coalesced along rows)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = tid + 256 * u, r = e & 31, k = e >> 5;
      const int gi = 32 * I + r, gj = 32 * K + k;
      pre[u] = (gi < p && gj < p) ? L(gi, gj) : 0.0;
    }
  };
  auto fetch_Y = [&](int I) {                     // Y_I as a left operand: A[r = i][k = j] = inv(i, j)
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int e = tid + 256 * u; pre[u] = invd[(size_t)I * LA_INVD_STRIDE + e]; }   // e = j * 32 + i
  };
  auto stage = [&]() {
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int e = tid + 256 * u, r = e & 31, k = e >> 5; Asm[k * TI_LD + r] = pre[u]; }
  };
  if (J + 1 < nblk) fetch_L(J + 1, J);
  __syncthreads();
  for (int I = J + 1; I < nblk; ++I) {
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int K = J; K < I; ++K) {
      stage();                                    // L_IK (fetched earlier) -> shared
      __syncthreads();
      if (K + 1 < I) fetch_L(I, K + 1); else fetch_Y(I);
      const double* Bm = Wsm + (size_t)(K - J) * TI_TILE;
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        const double af = Asm[(4 * k4 + fk) * TI_LD + 8 * mi + fr];
        const double b0 = Bm[(4 * k4 + fk) * TI_LD + 8 * n0 + fr];
        const double b1 = Bm[(4 * k4 + fk) * TI_LD + 8 * (n0 + 1) + fr];
        dmma_8x8x4(acc[0][0], acc[0][1], af, b0);
        dmma_8x8x4(acc[1][0], acc[1][1], af, b1);
      }
      __syncthreads();
    }
    // S -> shared as a right operand, Y_I -> shared as the left one
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int e = 0; e < 2; ++e) Ssm[(8 * mi + fr) * TI_LD + 8 * (n0 + t) + 2 * fk + e] = acc[t][e];
    stage();
    __syncthreads();
    if (I + 1 < nblk) fetch_L(I + 1, J);
    double out[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
      const double af = Asm[(4 * k4 + fk) * TI_LD + 8 * mi + fr];
      const double b0 = Ssm[(4 * k4 + fk) * TI_LD + 8 * n0 + fr];
      const double b1 = Ssm[(4 * k4 + fk) * TI_LD + 8 * (n0 + 1) + fr];
      dmma_8x8x4(out[0][0], out[0][1], af, b0);
      dmma_8x8x4(out[1][0], out[1][1], af, b1);
    }
    double* Wb = Wsm + (size_t)(I - J) * TI_TILE;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = 8 * mi + fr, c = 8 * (n0 + t) + 2 * fk + e;
        const double v = -out[t][e];
        Wb[r * TI_LD + c] = v;
        const int gi = 32 * I + r, gj = 32 * J + c;
        if (gi < p && gj < p) W[(size_t)gi + (size_t)gj * p] = v;
      }
    __syncthreads();
  }
}

// C(natural) = inv(G):  Cm(a,b) = sum_{k >= max(a,b)} W[k][a] W[k][b];  C[p-1-a][p-1-b] = Cm(a,b).
__global__ void __launch_bounds__(256) k_wtw(const double* __restrict__ W, int p, double* __restrict__ C) {
  __shared__ double sa[16][17], sb[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int a0 = blockIdx.x * 16, b0 = blockIdx.y * 16;
  if (b0 > a0) return;  // lower tiles only; mirrored on output
  double acc = 0.0;
  for (int k0 = a0; k0 < p; k0 += 16) {   // k >= max(a,b) >= a0 (W is lower triangular: zero above)
    const int k = k0 + tx;   // tx runs along k: coalesced column reads of W
    sa[tx][ty] = (k < p && a0 + ty < p) ? W[(size_t)k + (size_t)(a0 + ty) * p] : 0.0;
    sb[tx][ty] = (k < p && b0 + ty < p) ? W[(size_t)k + (size_t)(b0 + ty) * p] : 0.0;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) acc = fma(sa[kk][ty], sb[kk][tx], acc);
    __syncthreads();
  }
  const int a = a0 + ty, b = b0 + tx;
  if (a < p && b < p) {
    C[(size_t)(p - 1 - a) + (size_t)(p - 1 - b) * p] = acc;
    C[(size_t)(p - 1 - b) + (size_t)(p - 1 - a) * p] = acc;
  }
}

}  // namespace gamc
