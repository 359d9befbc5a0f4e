/*
 * gamc_b200.h -- C ABI of the B200-native GAMC hot path (libgamc_b200.so).
 *
 * Drop-in boundary for papamarkou/GAMCSampler.jl's per-iteration hot path: evaluation of log-target,
 * gradient and metric tensor, and the SMMALA / AM transition that consumes them.  Plain pointers and
 * sizes only; all matrices are COLUMN-MAJOR FP64 (Julia's native layout, so a Julia `Matrix{Float64}` /
 * `Vector{Float64}` is passed by pointer with no copy on the host side).  The caller owns every host
 * buffer; the library owns all device memory behind the handle.  One handle = one CUDA device + one
 * stream; calls on one handle are serialised, different handles are independent.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference repo).
 * The ccall binding a GAMCSampler.jl maintainer would add is in INTEGRATION.md and
 * gamcsampler.jl_b200/julia/GAMCSamplerB200.jl.
 */
#ifndef GAMC_B200_H
#define GAMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gamc_ctx* gamc_handle;

/* Error behaviour of the reference: `@assert` -> AssertionError (src/samplers/GAMC.jl:51-67,135-138,168-170),
 * PosDefException from `chol` (src/samplers/iterate/GAMC.jl:30,94 and the MvNormal constructor behind
 * set_gmm, src/samplers/GAMC.jl:188), SingularException from `inv` (:26,43). */
typedef enum {
  GAMC_OK = 0,
  GAMC_INVALID_ARG = 1,       /* AssertionError: bad sampler parameter / null pointer / bad size          */
  GAMC_NONFINITE_INIT = 2,    /* "Log-target not finite: initial value out of support" (GAMC.jl:168-170) */
  GAMC_NOT_POSDEF = 3,        /* PosDefException / SingularException; pivot via gamc_last_error_string   */
  GAMC_CUDA_ERROR = 4,
  GAMC_NCCL_ERROR = 5,
  GAMC_UNSUPPORTED_SHAPE = 6, /* p outside the compiled kernel range (1..512)                            */
  GAMC_NOT_READY = 7          /* call order violated (no target / sampler not initialised / no tape)     */
} gamc_status;

/* Schedules: `update!` functions of src/samplers/GAMC.jl:264-337 (decay formulas src/stats/schedule.jl:1-7). */
typedef enum {
  GAMC_SCHED_AM_ONLY = 0,        /* am_only_update!          :264                         */
  GAMC_SCHED_SMMALA_ONLY = 1,    /* smmala_only_update!      :336                         */
  GAMC_SCHED_MOD = 2,            /* mod_update!(n=10)        :277   params[0] = n          */
  GAMC_SCHED_RAND = 3,           /* rand_update!(p=0.5)      :327   params[0] = p          */
  GAMC_SCHED_COS = 4,            /* cos_update!(a,b,c)       :267   params = a,b,c         */
  GAMC_SCHED_RAND_EXP_DECAY = 5, /* rand_exp_decay_update!   :287   params = a(10), b(0)   */
  GAMC_SCHED_RAND_LIN_DECAY = 6, /* rand_lin_decay_update!   :297   params = a(1e-2), b    */
  GAMC_SCHED_RAND_POW_DECAY = 7, /* rand_pow_decay_update!   :307   params = a(1e-3), b    */
  GAMC_SCHED_RAND_QUAD_DECAY = 8 /* rand_quad_decay_update!  :317   params = a(1e-5), b    */
} gamc_schedule;

/* Sampler + tuner + range configuration.  Mirrors `GAMC(; update, transform, driftstep, minorscale, c, t0)`
 * (src/samplers/GAMC.jl:126-151), `GAMCTuner(smmalatuner, amtuner, totaltuner)` (src/tuners/GAMCTuner.jl:15-23)
 * with Klara's VanillaMCTuner / AcceptanceRateMCTuner, and `BasicMCRange(nsteps, burnin)`
 * (examples/logistic_regression/src/GAMC/simulations.jl:57). */
typedef struct {
  double driftstep;        /* > 0                                                                  */
  double minorscale;       /* > 0: covariance scale of the stabilising mixture component           */
  double c;                /* >= 0: weight of the stabilising component                            */
  int32_t t0;              /* > 0: first t0 iterations are always SMMALA                           */
  int32_t schedule;        /* gamc_schedule                                                        */
  double sched_params[3];  /* NaN = reference default for that schedule                            */
  int32_t total_tuner_kind;/* 0 = VanillaMCTuner, 1 = AcceptanceRateMCTuner                        */
  double targetrate;       /* AcceptanceRateMCTuner target (0.35 in the examples)                  */
  double score_k;          /* logistic_rate_score steepness (Klara default 7)                      */
  int32_t period;          /* tuner period (Klara default 100)                                     */
  int32_t verbose_total;   /* VanillaMCTuner(verbose=...) flags: they switch the counters on       */
  int32_t verbose_smmala;  /*   (src/samplers/iterate/GAMC.jl:4-6,15-17,129-131)                   */
  int32_t verbose_am;
  int64_t nsteps;          /* BasicMCRange.nsteps (the `tot` argument of the schedules)            */
  int64_t burnin;          /* BasicMCRange.burnin                                                  */
  int32_t am_mode;         /* 0 = literal covariance recursion + Cholesky each AM step,
                              1 = rank-one Cholesky update of the same recursion (fast path),
                              2 = as 1, computed for both accept/reject outcomes on a side stream while
                                  the data pass runs (bit-identical chain, shorter critical path),
                              3 = two consecutive AM steps share ONE pass over X: the pass evaluates the log-likelihood at this
                                  step's proposal and at both candidate proposals of the next step (same arithmetic per
                                  vector; steps that are burn-in tuning events are not paired)                       */
  int32_t sampler_kind;    /* 0 = GAMC, 1 = MALA (Klara's comparison sampler, examples/logistic_regression/src/MALA/simulations.jl:37;
                              identity metric: only driftstep, the total tuner and the range are used)          */
  int32_t reuse_tensor;    /* 0 = literal: an SMMALA step that follows an AM step always re-evaluates gradient and metric at
                              the current point (src/samplers/iterate/GAMC.jl:19-31);
                              1 = skip that re-evaluation (on the device, no host round trip) when no AM proposal has been
                                  accepted since they were last computed: same inputs, deterministic kernels, so the chain is
                                  bit-identical (tested) -- an optimisation the reference does not have; off by default and
                                  never part of bench.py's headline value; single-GPU jobs only                      */
  int32_t reserved_;       /* keep zero */
} gamc_config;

/* Snapshot of the scalar sampler state: fields of MuvGAMCState / GAMCTune that the reference exposes
 * (src/samplers/GAMC.jl:7-27, src/tuners/GAMCTuner.jl:1-7; read after a run at
 * examples/logistic_regression/src/GAMC/simulations.jl:85-86). */
typedef struct {
  int64_t count;
  int64_t updatetensorcount;
  double step;             /* tune.totaltune.step */
  double sqrttunestep;
  int64_t total_accepted, total_proposed, total_totproposed;
  double total_rate;
  int64_t smmala_accepted, smmala_proposed, smmala_totproposed;
  int64_t am_accepted, am_proposed, am_totproposed;
  double smmalafrequency, amfrequency;
  double logtarget;        /* pstate.logtarget */
  double ratio;            /* sstate.ratio of the last transition */
  int32_t presentupdatetensor, pastupdatetensor;
  int32_t error_flag;      /* 0, or GAMC_NOT_POSDEF raised on device during the run */
  int32_t error_pivot;     /* 1-based pivot index of the failed factorisation */
  int64_t error_iteration; /* iteration count at which it was raised */
  int64_t chain_saved;     /* number of samples currently in the device chain buffer */
} gamc_state_scalars;

/* Vectors / matrices of the state that can be copied out (gamc_get_array). */
typedef enum {
  GAMC_ARR_VALUE = 0,          /* pstate.value            p        */
  GAMC_ARR_GRADLOGTARGET = 1,  /* pstate.gradlogtarget    p        */
  GAMC_ARR_TENSORLOGTARGET = 2,/* pstate.tensorlogtarget  p x p    */
  GAMC_ARR_LASTMEAN = 3,       /* sstate.lastmean         p        */
  GAMC_ARR_SECONDLASTMEAN = 4, /* sstate.secondlastmean   p        */
  GAMC_ARR_OLDINVTENSOR = 5,   /* sstate.oldinvtensor     p x p  (G^-1 after SMMALA, AM covariance after AM) */
  GAMC_ARR_CHOLINVTENSOR = 6,  /* sstate.cholinvtensor    p x p lower (L L' = G^-1)                          */
  GAMC_ARR_OLDFIRSTTERM = 7,   /* sstate.oldfirstterm     p        */
  GAMC_ARR_PROPOSED_VALUE = 8, /* sstate.pstate.value     p        */
  /* monitored extras of the current point that iterate! keeps in step with the value when they are in outopts[:monitor]
   * (src/samplers/iterate/GAMC.jl:74-90,100-106); for the logistic target they follow from (value, gradient, tensor) and lambda */
  GAMC_ARR_GRADLOGLIKELIHOOD = 9,   /* pstate.gradloglikelihood    p      */
  GAMC_ARR_GRADLOGPRIOR = 10,       /* pstate.gradlogprior         p      */
  GAMC_ARR_TENSORLOGLIKELIHOOD = 11,/* pstate.tensorloglikelihood  p x p  */
  GAMC_ARR_TENSORLOGPRIOR = 12      /* pstate.tensorlogprior       p x p  */
} gamc_array_id;

/* ---- lifecycle ------------------------------------------------------------------------------------- */
gamc_status gamc_create(gamc_handle* out, int32_t device);
gamc_status gamc_destroy(gamc_handle h);
const char* gamc_last_error_string(gamc_handle h);
/* Use the caller's CUDA stream (cudaStream_t passed as void*) instead of the handle's own. */
gamc_status gamc_set_stream(gamc_handle h, void* cuda_stream);
gamc_status gamc_sync(gamc_handle h);
/* "gamc_b200 <version> sm_100a" */
const char* gamc_version(void);

/* ---- target: Bayesian logistic regression -----------------------------------------------------------
 * Replaces the user callbacks ploglikelihood / plogprior / pgradlogtarget / ptensorlogtarget of
 * examples/logistic_regression/src/GAMC/simulations.jl:25-37 and the model data `[lambda, X, y, p]`
 * (:39-48,69).  X is N x p column-major with leading dimension ldX >= N, y is N (0/1), lambda the prior
 * VARIANCE.  One H2D copy; the handle keeps the device copy.  In a multi-GPU job each rank passes ITS row
 * shard and the prior terms are added once after the all-reduce. */
gamc_status gamc_target_logistic(gamc_handle h, const double* X, int64_t ldX, const double* y,
                                 int64_t N, int32_t p, double lambda);
/* Same, for data already on the device (no copy; the caller keeps the buffers alive; ldX must be even
 * and dX 16-byte aligned for the TMA descriptor). */
gamc_status gamc_target_logistic_device(gamc_handle h, const double* dX, int64_t ldX, const double* dy,
                                        int64_t N, int32_t p, double lambda);
/* Synthetic shard generated on the device: rows [row0, row0+nrows) of the global problem,
 * X[i,j] = standardised Irwin-Hall(4) from Philox4x32-10 keyed by (seed; i, j), y ~ Bernoulli(sigma(x.theta_true)).
 * Bit-identical to oracle/logistic.py:synth_logistic for any sharding. */
gamc_status gamc_target_logistic_synth(gamc_handle h, uint64_t seed, int64_t row0, int64_t nrows, int32_t p,
                                       const double* theta_true, double lambda);
/* Route of the metric  X' diag(sigma (1 - sigma)) X  (`ptensorlogtarget`, simulations.jl:35-36) for the logistic targets set up
 * AFTER this call:
 *   GAMC_METRIC_AUTO: the integer route when the shard of the design matrix has at least 2^22 entries, else FP64;
 *   GAMC_METRIC_F64:  weighted SYRK on the FP64 tensor cores (DMMA), elementwise accurate to rounding;
 *   GAMC_METRIC_I8:   error-free int8 digit products on the 5th-generation tensor cores (tcgen05, TMEM accumulators): diag(sqrt w) X
 *                     is cut into `slices` signed base-256 digit planes (2..7; 0 = 6) and the digit products D_a' D_b with
 *                     a + b <= smax (0 = slices + 1) are summed exactly in integers; with the defaults the result agrees with the
 *                     FP64 evaluation to 1e-13 of the largest entry of the metric, deterministically (no dependence on
 *                     scheduling).  Needs slices/8 times the memory of X for the digit planes.
 * The environment variable GAMC_METRIC_I8 ("0", "1" or "slices,smax") overrides the route at target set-up. */
typedef enum { GAMC_METRIC_AUTO = 0, GAMC_METRIC_F64 = 1, GAMC_METRIC_I8 = 2 } gamc_metric_route;
gamc_status gamc_set_metric_route(gamc_handle h, int32_t route, int32_t slices, int32_t smax);
/* The route the CURRENT target uses (GAMC_METRIC_F64 or GAMC_METRIC_I8) and, for the integer route, its digit planes and smax. */
gamc_status gamc_get_metric_route(gamc_handle h, int32_t* route, int32_t* slices, int32_t* smax);
/* Copy rows [r0, r0+n) of the device-resident design matrix / labels back (n x p column-major, ld = n). */
gamc_status gamc_target_read_rows(gamc_handle h, int64_t r0, int64_t n, double* X_out, double* y_out);

/* ---- target evaluation ------------------------------------------------------------------------------
 * Replace the Klara closures `parameter.logtarget!(pstate)` (call site src/samplers/iterate/GAMC.jl:148)
 * and `parameter.uptotensorlogtarget!(pstate)` (:20,37; src/samplers/GAMC.jl:163,241).
 * theta: p host doubles.  logtarget = loglikelihood + logprior; grad: p; G: p x p column-major, full
 * symmetric.  Any output pointer may be NULL.  In a multi-GPU job every rank calls with the same theta
 * and receives the all-reduced result. */
gamc_status gamc_eval_logtarget(gamc_handle h, const double* theta, double* logtarget);
gamc_status gamc_eval_upto_tensor(gamc_handle h, const double* theta, double* logtarget, double* grad, double* G);
/* Likelihood-only partial sums of THIS rank's shard, before any all-reduce and without prior terms
 * (what the rank contributes to the reduction): ll, g_lik[p], G_lik[p x p]. */
gamc_status gamc_eval_partials(gamc_handle h, const double* theta, double* ll, double* g_lik, double* G_lik);

/* ---- generic target ----------------------------------------------------------------
 * Replaces, for an arbitrary model, the Klara closures `parameter.logtarget!(pstate)` and
 * `parameter.uptotensorlogtarget!(pstate)` called at src/samplers/iterate/GAMC.jl:20,37,148 and
 * src/samplers/GAMC.jl:163: the HOST evaluates the model, the sampler's dense algebra, proposal, accept/reject,
 * running means, tuner and chain stay on the device.  One host round trip per evaluation -- meant for the
 * small-p examples and for checking models, not for the data-parallel path.
 * fn(theta[p], p, level, &logtarget, grad[p], G[p*p], user): level 0 = log-target only, 1 = + gradient,
 * 2 = + metric tensor (column-major, full symmetric, AFTER any `sampler.transform`, iterate/GAMC.jl:22-24,39-41).
 * All values are those of the full log-target (prior included).  Return 0; anything else aborts the run with
 * GAMC_INVALID_ARG.  Non-finite values at a proposal reject it, at the initial point raise GAMC_NONFINITE_INIT. */
typedef int32_t (*gamc_target_fn)(const double* theta, int32_t p, int32_t level, double* logtarget, double* grad, double* G, void* user);
gamc_status gamc_target_callback(gamc_handle h, int32_t p, gamc_target_fn fn, void* user);
/* The same closures for a model that lives ON THE DEVICE ("user-supplied device callback"): `enqueue` is called
 * on the host but must only ENQUEUE work on `cuda_stream` (kernels, cuBLAS calls, CUDA.jl launches ...) that reads the p doubles at
 * d_theta and writes the evaluation into d_out, laid out as the library's landing buffer:
 *   d_out[0] = log-target, d_out[1 .. p] = gradient (level >= 1), d_out[goff + i + j*p] = tensor (level 2, column-major, after
 *   any `sampler.transform`), goff = (p + 2) & ~1.
 * No host synchronisation happens per evaluation: the sampler's kernels are ordered after the user's work by the stream, so a
 * whole gamc_run stays asynchronous.  Values are those of the full log-target (prior included). */
typedef int32_t (*gamc_target_enqueue_fn)(const double* d_theta, int32_t p, int32_t level, double* d_out, void* cuda_stream, void* user);
gamc_status gamc_target_device_callback(gamc_handle h, int32_t p, gamc_target_enqueue_fn enqueue, void* user);

/* ---- multi-GPU: observations row-sharded, one process per GPU -------------------------
 * gamc_comm_unique_id fills a 128-byte NCCL unique id on one rank; the host side (torch.distributed /
 * MPI / Julia Distributed) broadcasts it; every rank then calls gamc_comm_init.  After that, every
 * target evaluation all-reduces [ll, g_lik, G_lik] (NCCL, sum, in place, on the handle's stream).
 * gamc_comm_init also maps a small mailbox of every peer through CUDA IPC (NVLink peer memory): the scalar
 * log-likelihood of an adaptive-Metropolis step is then all-reduced inside the accept/reject kernel itself
 * (one-shot: every rank writes its partial to every peer, sums in rank order) instead of through NCCL; and
 * gamc_sampler_init maps every peer's pair of reduced-evaluation buffers, so the [ll, g, G] payload of
 * SMMALA steps is summed over the ranks by the landing kernel itself (loads over NVLink, rank order).
 * gamc_comm_peer_path reports which of the two is active: bit 0 = AM scalar, bit 1 = SMMALA payload
 * (0 = NCCL carries both; e.g. GAMC_NO_P2P=1, GAMC_NO_P2P_LAND=1, no peer access, more than 32 ranks).
 * The standalone gamc_eval_* entry points always use NCCL. */
gamc_status gamc_comm_unique_id(uint8_t id_out[128]);
gamc_status gamc_comm_init(gamc_handle h, const uint8_t id[128], int32_t nranks, int32_t rank);
gamc_status gamc_comm_peer_path(gamc_handle h, int32_t* active);
/* How long a rank waits inside a step kernel for a peer's contribution before it gives up with GAMC_NCCL_ERROR (wall-clock
 * seconds, measured with %globaltimer; default 30 s, or GAMC_PEER_TIMEOUT_S).  Ranks are aligned by a stream-ordered NCCL
 * barrier at every gamc_sampler_init and at the start of every gamc_run, so the wait only has to cover kernel-time skew. */
gamc_status gamc_comm_set_timeout(gamc_handle h, double seconds);

/* ---- sampler ----------------------------------------------------------------------------------------
 * gamc_sampler_init = Klara `initialize!` + `sampler_state` for GAMC (src/samplers/GAMC.jl:157-176,202-228):
 * evaluates target/gradient/metric at theta0, checks finiteness, factorises the metric, seeds lastmean,
 * counters and tuner state (tuner_state, :195-200).  Calling it again is `reset!` (:234-260). */
gamc_status gamc_sampler_init(gamc_handle h, const gamc_config* cfg, const double* theta0);

/* `GAMC.update!::Function` (src/samplers/GAMC.jl:127; the examples pass closures, examples/logistic_regression/src/GAMC/
 * simulations.jl:51): a user-supplied schedule.  fn(i, tot, u, user) returns 1 for an SMMALA step and 0 for an AM step at
 * iteration i of tot; u is the iteration's schedule uniform from the tape (use it in place of `rand(Bernoulli(.))` so that the
 * run stays reproducible).  It replaces gamc_config.schedule until fn = NULL is set.  Evaluated on the host, ahead of the
 * device, when gamc_run builds the launch sequence -- it can depend on (i, tot, u) and on host state, not on the chain. */
typedef int32_t (*gamc_schedule_fn)(int64_t i, int64_t tot, double u, void* user);
gamc_status gamc_set_schedule_callback(gamc_handle h, gamc_schedule_fn fn, void* user);

/* Restart a chain from saved host state (Klara has no checkpointing, `reset!` src/samplers/GAMC.jl:234-260 only
 * re-initialises): after gamc_sampler_init, overwrite the counters / tuner state (count, updatetensorcount, step, the three
 * BasicMCTune triples, present/pastupdatetensor) from `st`, the value, lastmean and secondlastmean from the p-vectors, and --
 * when the last transition was an AM step (st->pastupdatetensor == 0) -- sstate.oldinvtensor (the running AM covariance, p x p)
 * from `oldinvtensor`.  Log-target, gradient, metric and their factorisation are re-evaluated at `theta`.  The tape position
 * restarts at 0: pass the tape of the REMAINING iterations to gamc_tape_set. */
gamc_status gamc_set_state(gamc_handle h, const gamc_state_scalars* st, const double* theta, const double* lastmean,
                           const double* secondlastmean, const double* oldinvtensor);

/* Random tape for the next n iterations (host arrays): u_sched[n], u_mix[n], z[p x n column-major],
 * u_acc[n].  Replaces the reference's in-line `rand(Bernoulli(.))`, `rand(proposal)`, `randn(p)`, `rand()`
 * (src/samplers/GAMC.jl:276-334, src/samplers/iterate/GAMC.jl:35,69,146,158) with host-drawn numbers in
 * fixed per-iteration slots.  H2D copy. */
gamc_status gamc_tape_set(gamc_handle h, int64_t n, const double* u_sched, const double* u_mix,
                          const double* z, const double* u_acc);
/* Device-generated tape (Philox, throughput mode): same slots, no host traffic. */
gamc_status gamc_tape_generate(gamc_handle h, int64_t n, uint64_t seed);

/* Read slots [first, first+n) of the resident tape back (host arrays as in gamc_tape_set; any pointer may be NULL): what a
 * device-generated tape drew, e.g. to replay a run on the host. */
gamc_status gamc_tape_get(gamc_handle h, int64_t first, int64_t n, double* u_sched, double* u_mix, double* z, double* u_acc);

/* Enqueue n `iterate!` transitions (src/samplers/iterate/GAMC.jl:1-293) reading the resident tape;
 * asynchronous.  Samples with iteration index > burnin are appended to the device chain buffer. */
gamc_status gamc_run(gamc_handle h, int64_t n);
/* Copy saved samples [first, first+n) out: value is p x n column-major (`chain.value`), accept n bytes
 * (`chain.diagnosticvalues`, examples/logistic_regression/src/GAMC/simulations.jl:81-82).  Synchronises. */
gamc_status gamc_chain_get(gamc_handle h, int64_t first, int64_t n, double* value, uint8_t* accept);
/* Convenience: gamc_tape_set + gamc_run + gamc_chain_get of everything saved during these n iterations
 * (n_saved_out receives the count).  This is the host-buffer end-to-end call. */
gamc_status gamc_iterate(gamc_handle h, int64_t n, const double* u_sched, const double* u_mix, const double* z,
                         const double* u_acc, double* value, uint8_t* accept, int64_t* n_saved_out);
/* Which branch each of the n tape iterations would take (1 = SMMALA, 0 = AM), from the resident tape and
 * the current count -- the schedule coin depends only on (i, nsteps, u_sched). */
gamc_status gamc_peek_kinds(gamc_handle h, int64_t n, uint8_t* kinds_out);

gamc_status gamc_get_state(gamc_handle h, gamc_state_scalars* out);
gamc_status gamc_get_array(gamc_handle h, int32_t array_id, double* out);
/* pstate.loglikelihood and pstate.logprior of the current point (monitored extras, src/samplers/iterate/GAMC.jl:100-106). */
gamc_status gamc_get_monitor(gamc_handle h, double* loglikelihood, double* logprior);
/* pstate.logtarget of the saved samples [first, first+n) (Klara `outopts[:monitor]` containing :logtarget). */
gamc_status gamc_chain_get_logtarget(gamc_handle h, int64_t first, int64_t n, double* logtarget);
/* The burn-in report of src/samplers/iterate/GAMC.jl:199-292, one record per tuning event, recorded on the device (the
 * reference prints it when a tuner is verbose; the front-end prints the same lines from these records).  Record layout, 16
 * doubles: [count, total accepted, total proposed, total rate, smmala accepted, smmala proposed, smmala rate, smmala frequency,
 * am accepted, am proposed, am rate, am frequency, step after tune!, 0, 0, 0] -- counters as they stood before reset_burnin!.
 * Copies up to max_records records into out; *n_records receives the number recorded so far. */
gamc_status gamc_tune_log_get(gamc_handle h, double* out, int64_t max_records, int64_t* n_records);


/* Batched logistic regression (the as-shipped example: N = 200, p = 4, 10 chains; examples/logistic_regression/src/GAMC/
 * simulations.jl:13-37,68-91): the same callbacks as gamc_target_logistic, evaluated inside the persistent batched kernel.
 * p <= 32, N <= 2048. */
gamc_status gamc_batched_target_logistic(gamc_handle h, const double* X, int64_t ldX, const double* y, int64_t N, int32_t p, double lambda);

/* ---- batched chains (small p): many independent chains, one persistent kernel per run -----------------------------
 * Replaces the reference's sequential `while i <= nchains ... run(job)` loops (examples/t_distribution/src/GAMC/simulations.jl:56-79)
 * and, per chain, the same `iterate!` as above -- with the target evaluated inside the kernel.
 * gamc_batched_target_mvt: the target of examples/t_distribution/src/GAMC/simulations.jl:17-33,
 *   logpdf(MvTDist(nu, 0, Sigma_t), x) = logconst - (nu+n)/2 log1p(x' Sinv x / nu); Sinv = inv(Sigma_t) (n x n, symmetric),
 *   gradient and tensor (= -Hessian) in closed form (the reference uses ForwardDiff, :33).
 * transform: 0 = none, 1 = Klara softabs(H, alpha) (`transform=H -> softabs(H, 1000.)`, :39).
 * theta0: p x nchains column-major.  Chains are independent: in a multi-GPU job every rank simply owns its own chains. */
gamc_status gamc_batched_target_mvt(gamc_handle h, int32_t n, double nu, const double* Sinv, double logconst);
/* gamc_batched_target_rv: the Keplerian radial-velocity target of examples/radial_velocity/src/stat_model.jl:18-63
 * (calc_model_rv rv_model_keplerian.jl:9-51, Kepler solver kepler_eqn.jl:18-56, parameter packing rv_model_param.jl:45-96):
 * theta = [log1p(P), log1p(K), e cos w, e sin w, w + M0] per planet + offset; t, obs, sigma: N epochs.  Gradient and
 * tensor (= -Hessian) by second-order forward-mode dual numbers in the kernel (the reference uses ForwardDiff, order 2). */
gamc_status gamc_batched_target_rv(gamc_handle h, const double* t, const double* obs, const double* sigma, int64_t N, int32_t nplanets);
gamc_status gamc_batched_init(gamc_handle h, const gamc_config* cfg, int32_t nchains, const double* theta0,
                              int32_t transform, double softabs_alpha);
/* n transitions of every chain.  Tape pointers may all be NULL (device Philox keyed by (seed, chain, iteration));
 * otherwise u_sched/u_mix/u_acc are [nchains][n] and z is [nchains][n][p] (C order). */
gamc_status gamc_batched_run(gamc_handle h, int64_t n, uint64_t seed, const double* u_sched, const double* u_mix,
                             const double* z, const double* u_acc);
/* value: [nchains][cap][p] doubles with cap = nsteps - burnin (sample s of chain c at value[(c*cap + s)*p]); accept [nchains][cap]. */
/* Multi-GPU: chains are independent, every rank owns a block of them; the global index of this rank's first chain keys the
 * device random streams so that ranks draw distinct numbers.  Call before gamc_batched_init (default 0). */
gamc_status gamc_batched_set_chain_offset(gamc_handle h, int64_t chain_offset);
gamc_status gamc_batched_chain_get(gamc_handle h, int32_t chain_first, int32_t nchains, double* value, uint8_t* accept);
gamc_status gamc_batched_get_state(gamc_handle h, int32_t chain, gamc_state_scalars* out, double* theta);

/* ---- batched chains, LARGE dimension: the t target at d up to 1024 (BASELINE config: d = 1024, 4096 chains per GPU) ----------
 * The same experiment as gamc_batched_target_mvt (examples/t_distribution/src/GAMC/simulations.jl:17-39,56-79) for dimensions
 * where the reference's dense route (ForwardDiff Hessian, `eig` inside softabs, `inv`, `chol`) is O(d^3) per chain and step.  The
 * inverse scale matrix T = inv(Sigma_t) must be symmetric TRIDIAGONAL -- it is for the reference's Sigma_ij = c^|i-j| -- and is given
 * by its diagonal tdiag[n] and off-diagonal toff[n-1]:
 *   logpdf(MvTDist(nu, 0, Sigma_t), x) = lgamma((nu+n)/2) - lgamma(nu/2) - n/2 log(nu pi) + 1/2 logdet T - (nu+n)/2 log1p(x'Tx/nu).
 * SMMALA steps use the structure (tensor = a (T - beta v v'); softabs changes only its few eigenvalues below 21/alpha, found by
 * bisection on the LDL' inertia count; reverse Cholesky factor in product form): O(d) memory and O(d r^2) work per evaluation
 * instead of O(d^3).  AM steps keep the dense d x d Cholesky factor of the proposal covariance per chain in HBM (8 MB at d = 1024)
 * and update it in one streamed O(d^2) sweep.  One warp per chain (SMMALA) / one CTA per chain (AM).
 * max_factors: product factors of 32 columns allocated per evaluation (0 = default 5: up to 159 eigenvalues below 21/alpha, enough
 * for the reference's N(0, 2^2) initial values at d = 1024); a chain that needs more raises GAMC_UNSUPPORTED_SHAPE.
 * theta0: [nchains][n] (row c = initial value of chain c).  Tape / chain layouts as for gamc_batched_run / gamc_batched_chain_get. */
gamc_status gamc_bigd_target_mvt(gamc_handle h, int32_t n, double nu, const double* tdiag, const double* toff);
gamc_status gamc_bigd_set_chain_offset(gamc_handle h, int64_t chain_offset);
gamc_status gamc_bigd_init(gamc_handle h, const gamc_config* cfg, int32_t nchains, const double* theta0, int32_t transform, double softabs_alpha,
                           int32_t max_factors);
gamc_status gamc_bigd_run(gamc_handle h, int64_t n, uint64_t seed, const double* u_sched, const double* u_mix, const double* z, const double* u_acc);
gamc_status gamc_bigd_chain_get(gamc_handle h, int32_t chain_first, int32_t nchains, double* value, uint8_t* accept);
/* out->error_pivot reports how many eigenvalues softabs changed at the chain's current point (diagnostic). */
gamc_status gamc_bigd_get_state(gamc_handle h, int32_t chain, gamc_state_scalars* out, double* theta);

/* ---- measurement hooks ------------------------------------------------------------------------------
 * With profiling on, every kernel launched by gamc_run / gamc_eval_* is bracketed by CUDA events on the
 * handle's stream; gamc_profile_get returns launches and summed milliseconds per kernel family. */
typedef enum {
  GAMC_K_LOGLIK = 0,     /* log-likelihood only data pass (AM step)              */
  GAMC_K_LOGLIK_GRAD = 1,/* fused log-likelihood + gradient (+ weights) data pass */
  GAMC_K_METRIC = 2,     /* X' diag(w) X on FP64 DMMA                             */
  GAMC_K_REDUCE = 3,     /* deterministic partial reductions                      */
  GAMC_K_LINALG = 4,     /* Cholesky / solves / proposal / MH accept kernels      */
  GAMC_K_ALLREDUCE = 5,  /* NCCL all-reduce / peer flag                           */
  GAMC_K_LAND = 6,       /* k_land: prior terms, slot fill, ratio partials (= the peer all-reduce in multi-GPU jobs) */
  GAMC_K_SMMALA_PRE = 7, /* k_smmala_pre: [factorisation of a re-evaluated point,] proposal                          */
  GAMC_K_SMMALA_POST = 8,/* k_smmala_post: factorisation of the proposal's metric, ratio, accept/reject, tail        */
  GAMC_K_AM_PRE = 9,     /* k_am_pre / k_am_pre_r1 on the critical path (speculative ones run on a side stream)      */
  GAMC_K_AM_POST = 10,   /* k_am_post: partial sums, peer all-reduce of the scalar, accept/reject, tail               */
  GAMC_K_HANDOVER = 11,  /* SMMALA -> AM hand-over: k_triinv, k_wtw, k_w_to_l                                        */
  GAMC_K_COUNT = 12
} gamc_kernel_family;
gamc_status gamc_profile_enable(gamc_handle h, int32_t on);
gamc_status gamc_profile_get(gamc_handle h, int32_t family, int64_t* launches, double* total_ms);
gamc_status gamc_profile_reset(gamc_handle h);
/* Total kernels launched by this handle since creation (for bench.py's gpu_launches). */
gamc_status gamc_launch_count(gamc_handle h, int64_t* out);
/* FP64 yard-sticks of THIS device, measured now: a register-only `mma.sync.m8n8k4.f64` (DMMA) loop and a DFMA loop on every
 * SM (a few milliseconds).  bench.py quotes the metric pass against the DMMA figure of the same lease. */
gamc_status gamc_probe_fp64_peak(gamc_handle h, double* dmma_tflops, double* dfma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* GAMC_B200_H */
