/* chain50.c -- the C ABI of include/gamc_b200.h driven from plain C99, no Python and no CUDA headers: a small logistic
 * problem, `gamc_sampler_init`, a host-drawn tape, 50 `iterate!` transitions in one `gamc_iterate` call, state read-out.
 * Build: gcc -std=c99 -I include tests/cpp/chain50.c -L gamcsampler.jl_b200 -lgamc_b200 -Wl,-rpath,$PWD/gamcsampler.jl_b200 -lm
 * Prints "CHAIN50 OK ..." and returns 0 when every call succeeded and the outputs are sane. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gamc_b200.h"

static uint64_t s_state = 0x9E3779B97F4A7C15ull;
static double u01(void) {                      /* splitmix64 -> (0, 1) */
  uint64_t z = (s_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return ((double)(z >> 11) + 0.5) / 9007199254740992.0;
}
static double gauss(void) { return sqrt(-2.0 * log(u01())) * cos(6.283185307179586 * u01()); }

#define CK(call)                                                                                   \
  do {                                                                                             \
    gamc_status st_ = (call);                                                                      \
    if (st_ != GAMC_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, (int)st_, gamc_last_error_string(h)); return 1; } \
  } while (0)

static int32_t every_fifth(int64_t i, int64_t tot, double u, void* user) { (void)tot; (void)u; (void)user; return (i % 5) == 0; }

int main(void) {
  enum { N = 500, P = 6, NSTEPS = 50, BURNIN = 10 };
  gamc_handle h = NULL;
  if (gamc_create(&h, 0) != GAMC_OK) { fprintf(stderr, "gamc_create failed: no sm_100 device (there is no CPU fallback)\n"); return 2; }
  double* X = (double*)malloc(sizeof(double) * N * P);
  double* y = (double*)malloc(sizeof(double) * N);
  double theta_true[P], theta0[P];
  for (int j = 0; j < P; ++j) theta_true[j] = gauss() / sqrt((double)P);
  for (int j = 0; j < P; ++j) for (int i = 0; i < N; ++i) X[i + (size_t)j * N] = gauss();      /* column-major */
  for (int i = 0; i < N; ++i) {
    double eta = 0.0;
    for (int j = 0; j < P; ++j) eta += X[i + (size_t)j * N] * theta_true[j];
    y[i] = u01() < 1.0 / (1.0 + exp(-eta)) ? 1.0 : 0.0;
  }
  CK(gamc_target_logistic(h, X, N, y, N, P, 100.0));
  double lt = 0.0, grad[P], G[P * P];
  CK(gamc_eval_upto_tensor(h, theta_true, &lt, grad, G));
  for (int a = 0; a < P; ++a) for (int b = 0; b < P; ++b) if (G[a + b * P] != G[b + a * P]) { fprintf(stderr, "metric not symmetric\n"); return 1; }

  gamc_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.driftstep = 0.02; cfg.minorscale = 0.001; cfg.c = 0.01; cfg.t0 = 3;
  cfg.schedule = GAMC_SCHED_RAND_EXP_DECAY; cfg.sched_params[0] = 10.0; cfg.sched_params[1] = 0.0; cfg.sched_params[2] = NAN;
  cfg.total_tuner_kind = 1; cfg.targetrate = 0.35; cfg.score_k = 7.0; cfg.period = 100;
  cfg.nsteps = NSTEPS; cfg.burnin = BURNIN; cfg.am_mode = 2;
  for (int j = 0; j < P; ++j) theta0[j] = theta_true[j] + 0.05 * gauss();
  CK(gamc_sampler_init(h, &cfg, theta0));

  double u_sched[NSTEPS], u_mix[NSTEPS], u_acc[NSTEPS], z[NSTEPS * P], value[NSTEPS * P];
  uint8_t accept[NSTEPS];
  for (int t = 0; t < NSTEPS; ++t) { u_sched[t] = u01(); u_mix[t] = u01(); u_acc[t] = u01(); for (int j = 0; j < P; ++j) z[t * P + j] = gauss(); }
  int64_t nsaved = -1;
  CK(gamc_iterate(h, NSTEPS, u_sched, u_mix, z, u_acc, value, accept, &nsaved));
  gamc_state_scalars st;
  CK(gamc_get_state(h, &st));
  if (nsaved != NSTEPS - BURNIN || st.count != NSTEPS || st.chain_saved != nsaved || st.error_flag != 0) { fprintf(stderr, "bad counters\n"); return 1; }
  int nacc = 0;
  for (int t = 0; t < nsaved; ++t) { nacc += accept[t]; for (int j = 0; j < P; ++j) if (!isfinite(value[t * P + j])) { fprintf(stderr, "non-finite sample\n"); return 1; } }
  double ll = 0.0, lp = 0.0;
  CK(gamc_get_monitor(h, &ll, &lp));
  if (fabs(ll + lp - st.logtarget) > 1e-9 * fabs(st.logtarget)) { fprintf(stderr, "monitor: loglikelihood + logprior != logtarget\n"); return 1; }

  /* a user schedule function (GAMC.update!::Function): SMMALA on every fifth iteration */
  CK(gamc_set_schedule_callback(h, every_fifth, NULL));
  CK(gamc_sampler_init(h, &cfg, theta0));
  CK(gamc_tape_set(h, NSTEPS, u_sched, u_mix, z, u_acc));
  CK(gamc_run(h, NSTEPS));
  CK(gamc_sync(h));
  CK(gamc_get_state(h, &st));
  if (st.updatetensorcount != 3 + (NSTEPS / 5 - 0)) { fprintf(stderr, "schedule callback: updatetensorcount %lld\n", (long long)st.updatetensorcount); return 1; }
  int64_t launches = 0;
  CK(gamc_launch_count(h, &launches));
  printf("CHAIN50 OK %s logtarget %.6f accepted %d/%d kernels %lld\n", gamc_version(), st.logtarget, nacc, (int)nsaved, (long long)launches);
  gamc_destroy(h);
  free(X); free(y);
  return 0;
}
