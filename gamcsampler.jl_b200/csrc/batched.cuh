// batched.cuh -- batched-chains GAMC for small parameter dimension (p <= 32): many independent chains, ONE
// persistent kernel launch for a whole run.  Each CTA owns one chain; its complete sampler state (theta, gradient,
// metric, inverse metric, Cholesky factors, AM covariance, running means, tuner) lives in shared memory; the whole
// `iterate!` loop of the reference (src/samplers/iterate/GAMC.jl:1-293) -- schedule coin, SMMALA or AM branch, target
// evaluation, optional `transform` (Klara softabs), dense algebra, accept/reject, means, burn-in tuning, chain output --
// runs inside the kernel, so there is no launch or synchronisation per iteration.
//
// Targets evaluated in the kernel:
//   TARGET_MVT  multivariate-t target of examples/t_distribution/src/GAMC/simulations.jl:17-33 (closed-form gradient and
//               Hessian in place of the reference's ForwardDiff sweep; tensor = -Hessian, then softabs(., alpha), :39)
//   TARGET_RV   Keplerian radial-velocity target of examples/radial_velocity/src/*.jl (stat_model.jl:18-63), epochs spread
//               over the CTA's threads, derivatives by second-order forward-mode dual numbers through the Kepler solver
// Warp 0 (lane = coordinate) does the small dense algebra; matrices are row-major with an odd leading dimension.
// Random numbers: a host tape (parity tests) or Philox4x32-10 keyed by (chain, iteration) (throughput mode).
#pragma once
#include "common.cuh"
#include "synth.cuh"

namespace gamc {

constexpr int BT_MAXP = 32;

struct BatchedCfg {
  int p, ld;                    // dimension, leading dimension (odd)
  int nchains;
  long long chain_offset;       // global index of this rank's first chain (keys the Philox streams; chains are sharded across ranks)
  int target;                   // 0 = mvt, 1 = rv, 3 = logistic regression
  int transform;                // 0 = none, 1 = softabs
  double softabs_alpha;
  // sampler / tuner / range (same meaning as gamc_config)
  double driftstep, minorscale, sqrtminorscale, c;
  int t0, schedule; double sp0, sp1, sp2;
  int tuning, count_total, is_accrate, verbose_sm, verbose_am, period;
  long long nsteps, burnin;
  double targetrate, score_k;
  // mvt target
  double nu, logconst; const double* Sinv;   // n x n row-major == column-major (symmetric)
  // rv target
  const double* rv_t; const double* rv_obs; const double* rv_sig; long long rv_n; int rv_npl; double rv_lognorm;
  // logistic target (small N, small p): X is N x p column-major, y in {0,1}, lambda = prior variance
  const double* lg_X; const double* lg_y; int lg_n; double lg_lambda;
  // tape (may be null -> Philox with `seed`)
  const double* tape_usched; const double* tape_umix; const double* tape_uacc; const double* tape_z;   // [chain][t] / [chain][t][p]
  long long tape_len;           // iterations per chain on the tape
  unsigned long long seed;
  // output
  double* chain_value; unsigned char* chain_accept; long long chain_cap;   // [chain][cap][p], [chain][cap]
  double* blob; int blob_doubles;                                        // persistent per-chain state
};

// ---- per-chain state layout in shared memory (doubles) ------------------------------------------------------------
enum { BV_THETA = 0, BV_PROP, BV_GC, BV_GP, BV_FC, BV_FP, BV_MU, BV_M1, BV_M2, BV_Z, BV_T1, BV_T2, BV_ROT0, BV_ROT1, BV_ROT2, BV_NVEC };   // T2, ROT0..2 are contiguous: softabs scratch
// The evaluation scratch W aliases the proposal-slot metric (the Jacobi input is dead once its eigenvalues are saved) and the
// AM factor aliases the proposal-slot factor (never live together): 6 matrices per chain instead of 8.
enum { BM_GC = 0, BM_IC, BM_LC, BM_GP, BM_IP, BM_LP, BM_NMAT, BM_LAM = BM_LP, BM_W = BM_GP };
enum { BS_LT = 0, BS_LTP, BS_LDC, BS_LDP, BS_STEP, BS_SQSTEP, BS_RATIO, BS_COUNT, BS_UPD, BS_TACC, BS_TPROP, BS_TTOT, BS_TRATE,
       BS_SMACC, BS_SMPROP, BS_SMTOT, BS_AMACC, BS_AMPROP, BS_AMTOT, BS_PAST, BS_PRESENT, BS_ERR, BS_SAVED, BS_NSCAL = 24 };

__host__ __device__ inline int bt_blob_doubles(int p, int ld) { return BS_NSCAL + BV_NVEC * BT_MAXP + BM_NMAT * p * ld; }

struct BState {
  double* s; double* v; double* m; int p, ld;
  __device__ double* vec(int k) const { return v + k * BT_MAXP; }
  __device__ double* mat(int k) const { return m + (size_t)k * p * ld; }
};

// =====================================================================================================
// warp-level small dense algebra (all 32 lanes call; lane = row; p <= 32)
// =====================================================================================================
// in-place lower Cholesky; returns sum log L_ii, NaN on failure (uniform across the warp)
__device__ inline double w_chol(double* A, int p, int ld, int lane) {
  double slog = 0.0;
  bool bad = false;
  for (int j = 0; j < p; ++j) {
    const double d = A[j * ld + j];
    if (!(d > 0.0) || !isfinite(d)) { bad = true; break; }
    const double rd = rsqrt(d);
    slog += 0.5 * log(d);
    if (lane == j) A[j * ld + j] = d * rd;
    else if (lane > j && lane < p) A[lane * ld + j] *= rd;
    __syncwarp();
    if (lane > j && lane < p) {
      const double lij = A[lane * ld + j];
      for (int c = j + 1; c <= lane; ++c) A[lane * ld + c] = fma(-lij, A[c * ld + j], A[lane * ld + c]);
    }
    __syncwarp();
  }
  return bad ? nan("") : slog;
}

// Inv = (L L')^-1, lane j computes column j with two substitutions, in place in Inv.  Result symmetric (averaged).
__device__ inline void w_inv_from_chol(const double* L, double* Inv, int p, int ld, int lane) {
  if (lane < p) {
    const int j = lane;
    for (int i = 0; i < p; ++i) {            // L y = e_j
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s = fma(-L[i * ld + k], Inv[k * ld + j], s);
      Inv[i * ld + j] = s / L[i * ld + i];
    }
    for (int i = p - 1; i >= 0; --i) {       // L' x = y
      double s = Inv[i * ld + j];
      for (int k = i + 1; k < p; ++k) s = fma(-L[k * ld + i], Inv[k * ld + j], s);
      Inv[i * ld + j] = s / L[i * ld + i];
    }
  }
  __syncwarp();
  if (lane < p) for (int c = 0; c < lane; ++c) { const double a = 0.5 * (Inv[lane * ld + c] + Inv[c * ld + lane]); Inv[lane * ld + c] = a; Inv[c * ld + lane] = a; }
  __syncwarp();
}

// y = A x (A p x p row-major, x/y length p); lane = row
__device__ inline void w_matvec(const double* A, const double* x, double* y, int p, int ld, int lane) {
  if (lane < p) {
    double s = 0.0;
    for (int j = 0; j < p; ++j) s = fma(A[lane * ld + j], x[j], s);
    y[lane] = s;
  }
  __syncwarp();
}
// y = L x with L lower
__device__ inline void w_trmv_lower(const double* L, const double* x, double* y, int p, int ld, int lane) {
  if (lane < p) {
    double s = 0.0;
    for (int j = 0; j <= lane; ++j) s = fma(L[lane * ld + j], x[j], s);
    y[lane] = s;
  }
  __syncwarp();
}
__device__ inline double w_dot(const double* a, const double* b, int p, int lane) {
  double s = (lane < p) ? a[lane] * b[lane] : 0.0;
  return warp_sum(s);
}

// Jacobi eigen-decomposition of the symmetric matrix A (destroyed: eigenvalues end on its diagonal), Q = vectors.
// Parallel (round-robin tournament) ordering: each step rotates p/2 DISJOINT index pairs at once -- lanes k < p/2 compute
// the rotation of pair k, then A <- A J (lane = row), A <- J' A (lane = column), Q <- Q J (lane = row), all conflict-free.
// rot = 3 * 16 doubles of shared scratch (cos, sin, pair indices).
__device__ inline void w_jacobi(double* A, double* Q, int p, int ld, int lane, double* rot) {
  if (lane < p) for (int j = 0; j < p; ++j) Q[lane * ld + j] = (lane == j) ? 1.0 : 0.0;
  __syncwarp();
  if (p < 2) return;
  const int m = (p + 1) & ~1;            // players (one dummy if p is odd)
  const int half = m >> 1;
  double* rc = rot; double* rs = rot + 16; int* rp = reinterpret_cast<int*>(rot + 32);   // rp[2k], rp[2k+1] = pair k
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0.0, dg = 0.0;
    if (lane < p) { for (int j = 0; j < p; ++j) { const double a = A[lane * ld + j]; if (j != lane) off = fma(a, a, off); else dg = a * a; } }
    off = warp_sum(off); dg = warp_sum(dg);
    if (off <= 1e-27 * dg || off == 0.0) break;   // |offdiag|_F <= 3e-14 |diag|_F: rounding level for p <= 32
    // round-robin positions of this lane's pair, advanced incrementally (player 0 fixed, players 1..m-1 rotate)
    int ra = (lane == 0) ? 0 : lane, rb = m - 1 - lane;          // step 0: pair k = (k, m-1-k); values in 1..m-1 rotate by +1 mod (m-1)
    for (int step = 0; step < m - 1; ++step) {
      if (lane < half) {
        int a = ra, b = rb;
        if (a > b) { const int t = a; a = b; b = t; }
        double c = 1.0, sn = 0.0;
        if (b < p) {
          const double apq = A[a * ld + b];
          const double app = A[a * ld + a], aqq = A[b * ld + b];
          // skip rotations that cannot change the diagonal at working precision
          if (fabs(apq) > 1e-19 * (fabs(app) + fabs(aqq))) {
            // t = tan(phi) = sgn(d) h / (|d| + sqrt(d^2 + h^2)), d = aqq - app, h = 2 apq   (one rsqrt, one divide)
            const double d = aqq - app, h2 = 2.0 * apq;
            const double q = fma(d, d, h2 * h2);
            const double r = q * rsqrt(q);
            const double t = (d >= 0.0 ? h2 : -h2) / (fabs(d) + r);
            c = rsqrt(fma(t, t, 1.0)); sn = t * c;
          }
        }
        rc[lane] = c; rs[lane] = sn; rp[2 * lane] = a; rp[2 * lane + 1] = (b < p && sn != 0.0) ? b : -1;
        // next step's pair
        if (lane != 0) ra = (ra == m - 1) ? 1 : ra + 1;
        rb = (rb == m - 1) ? 1 : rb + 1;
      }
      __syncwarp();
      if (lane < p) {                     // A <- A J and Q <- Q J: lane = row
        double* Ar = A + lane * ld; double* Qr = Q + lane * ld;
        for (int k = 0; k < half; ++k) {
          const int b = rp[2 * k + 1];
          if (b < 0) continue;
          const int a = rp[2 * k];
          const double c = rc[k], sn = rs[k];
          const double x = Ar[a], y = Ar[b];
          Ar[a] = c * x - sn * y; Ar[b] = fma(sn, x, c * y);
          const double qx = Qr[a], qy = Qr[b];
          Qr[a] = c * qx - sn * qy; Qr[b] = fma(sn, qx, c * qy);
        }
      }
      __syncwarp();
      if (lane < p) {                     // A <- J' A: lane = column
        for (int k = 0; k < half; ++k) {
          const int b = rp[2 * k + 1];
          if (b < 0) continue;
          const int a = rp[2 * k];
          const double c = rc[k], sn = rs[k];
          const double x = A[a * ld + lane], y = A[b * ld + lane];
          A[a * ld + lane] = c * x - sn * y; A[b * ld + lane] = fma(sn, x, c * y);
        }
      }
      __syncwarp();
    }
  }
}

// Klara softabs: G <- Q diag(lam coth(alpha lam)) Q'; also Inv <- Q diag(1/f) Q' and returns logdet(G) = sum log f.
// A holds the input (destroyed), Q scratch, outputs to G and Inv.
__device__ inline double w_softabs(double* A, double* Q, double* G, double* Inv, double alpha, int p, int ld, int lane, double* fbuf) {
  double* finv = fbuf + 32 + 48;   // fbuf: [f(32) | rotation scratch (48) | 1/f (32)] = T2, ROT0, ROT1, ROT2
  w_jacobi(A, Q, p, ld, lane, fbuf + 32);
  double lf = 0.0;
  if (lane < p) {
    const double lam = A[lane * ld + lane];
    const double x = alpha * lam;
    const double f = (fabs(x) < 1e-8) ? 1.0 / alpha : lam / tanh(x);
    fbuf[lane] = f;
    finv[lane] = 1.0 / f;
    lf = log(f);
  }
  lf = warp_sum(lf);
  __syncwarp();
  if (lane < p) {
    for (int j = 0; j <= lane; ++j) {
      double g = 0.0, iv = 0.0;
      for (int k = 0; k < p; ++k) { const double qq = Q[lane * ld + k] * Q[j * ld + k]; g = fma(qq, fbuf[k], g); iv = fma(qq, finv[k], iv); }
      G[lane * ld + j] = g; G[j * ld + lane] = g;
      Inv[lane * ld + j] = iv; Inv[j * ld + lane] = iv;
    }
  }
  __syncwarp();
  return lf;
}

// =====================================================================================================
// targets
// =====================================================================================================
// multivariate t: lt = logconst - (nu+n)/2 log1p(q/nu), q = x'Sinv x; grad = -a v; tensor(-H) = a (Sinv - 2 v v'/(nu+q)), a = (nu+n)/(nu+q)
__device__ inline double mvt_eval(const BatchedCfg& C, const double* x, double* g, double* H, double* vtmp, int level, int lane) {
  const int p = C.p, ld = C.ld;
  double vi = 0.0;
  if (lane < p) for (int j = 0; j < p; ++j) vi = fma(C.Sinv[lane * p + j], x[j], vi);
  const double q = warp_sum(lane < p ? vi * x[lane] : 0.0);
  const double lt = C.logconst - 0.5 * (C.nu + p) * log1p(q / C.nu);
  if (level >= 1) {
    const double a = (C.nu + p) / (C.nu + q);
    if (lane < p) { g[lane] = -a * vi; vtmp[lane] = vi; }
    __syncwarp();
    if (level >= 2 && lane < p) {
      const double cc = 2.0 / (C.nu + q);
      for (int j = 0; j < p; ++j) H[lane * ld + j] = a * (C.Sinv[lane * p + j] - cc * vi * vtmp[j]);
    }
    __syncwarp();
  }
  return lt;
}

// Bayesian logistic regression of examples/logistic_regression/src/GAMC/simulations.jl:25-37 for the as-shipped sizes (N = 200,
// p = 4): ploglikelihood + plogprior, pgradlogtarget, ptensorlogtarget.  One warp; lanes stride the observations, every sum ends in
// a warp reduction.  wr = [w(N) | r(N)] shared scratch: weights sigma(1-sigma) and residuals y - sigma of the first pass.
__device__ inline double logi_eval(const BatchedCfg& C, const double* x, double* g, double* H, double* wr, int level, int lane) {
  const int p = C.p, ld = C.ld, N = C.lg_n;
  const double* X = C.lg_X; const double* y = C.lg_y;
  double ll = 0.0;
  for (int n = lane; n < N; n += 32) {
    double eta = 0.0;
    for (int j = 0; j < p; ++j) eta = fma(X[n + (size_t)j * N], x[j], eta);
    double sig, sp;
    sigmoid_softplus(eta, sig, sp);
    ll += y[n] * eta - sp;
    if (level >= 1) { wr[n] = sig * (1.0 - sig); wr[N + n] = y[n] - sig; }
  }
  ll = warp_sum(ll);
  const double tt = warp_sum(lane < p ? x[lane] * x[lane] : 0.0);
  const double lt = ll - 0.5 * (tt / C.lg_lambda + p * log(2.0 * 3.14159265358979323846 * C.lg_lambda));
  if (level >= 1) {
    __syncwarp();
    for (int j = 0; j < p; ++j) {
      double s = 0.0;
      for (int n = lane; n < N; n += 32) s = fma(X[n + (size_t)j * N], wr[N + n], s);
      s = warp_sum(s);
      if (lane == 0) g[j] = s - x[j] / C.lg_lambda;
    }
    if (level >= 2) {
      for (int i = 0; i < p; ++i)
        for (int j = 0; j <= i; ++j) {
          double s = 0.0;
          for (int n = lane; n < N; n += 32) s = fma(X[n + (size_t)i * N] * wr[n], X[n + (size_t)j * N], s);
          s = warp_sum(s);
          if (lane == 0) { if (i == j) s += 1.0 / C.lg_lambda; H[i * ld + j] = s; H[j * ld + i] = s; }
        }
    }
    __syncwarp();
  }
  return lt;
}

// ---- second-order forward-mode dual numbers for the RV target (value, NP first and NP(NP+1)/2 second derivatives) -------
template <int NP>
struct Dual2 {
  static constexpr int NH = NP * (NP + 1) / 2;
  double v; double g[NP]; double h[NH];
  __device__ static int hidx(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
};
template <int NP> __device__ inline Dual2<NP> d_const(double c) { Dual2<NP> r; r.v = c; for (int i = 0; i < NP; ++i) r.g[i] = 0; for (int i = 0; i < Dual2<NP>::NH; ++i) r.h[i] = 0; return r; }
template <int NP> __device__ inline Dual2<NP> d_var(double c, int k) { Dual2<NP> r = d_const<NP>(c); r.g[k] = 1.0; return r; }
// y = f(x) with f' = f1, f'' = f2
template <int NP> __device__ inline Dual2<NP> d_unary(const Dual2<NP>& x, double f0, double f1, double f2) {
  Dual2<NP> r; r.v = f0;
  for (int i = 0; i < NP; ++i) r.g[i] = f1 * x.g[i];
  int k = 0;
  for (int i = 0; i < NP; ++i) for (int j = 0; j <= i; ++j, ++k) r.h[k] = fma(f1, x.h[k], f2 * x.g[i] * x.g[j]);
  return r;
}
template <int NP> __device__ inline Dual2<NP> operator+(const Dual2<NP>& a, const Dual2<NP>& b) { Dual2<NP> r; r.v = a.v + b.v; for (int i = 0; i < NP; ++i) r.g[i] = a.g[i] + b.g[i]; for (int i = 0; i < Dual2<NP>::NH; ++i) r.h[i] = a.h[i] + b.h[i]; return r; }
template <int NP> __device__ inline Dual2<NP> operator-(const Dual2<NP>& a, const Dual2<NP>& b) { Dual2<NP> r; r.v = a.v - b.v; for (int i = 0; i < NP; ++i) r.g[i] = a.g[i] - b.g[i]; for (int i = 0; i < Dual2<NP>::NH; ++i) r.h[i] = a.h[i] - b.h[i]; return r; }
template <int NP> __device__ inline Dual2<NP> operator*(const Dual2<NP>& a, const Dual2<NP>& b) {
  Dual2<NP> r; r.v = a.v * b.v;
  for (int i = 0; i < NP; ++i) r.g[i] = fma(a.v, b.g[i], b.v * a.g[i]);
  int k = 0;
  for (int i = 0; i < NP; ++i) for (int j = 0; j <= i; ++j, ++k) r.h[k] = a.v * b.h[k] + b.v * a.h[k] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
  return r;
}
template <int NP> __device__ inline Dual2<NP> operator*(double c, const Dual2<NP>& a) { Dual2<NP> r; r.v = c * a.v; for (int i = 0; i < NP; ++i) r.g[i] = c * a.g[i]; for (int i = 0; i < Dual2<NP>::NH; ++i) r.h[i] = c * a.h[i]; return r; }
template <int NP> __device__ inline Dual2<NP> operator+(const Dual2<NP>& a, double c) { Dual2<NP> r = a; r.v += c; return r; }
template <int NP> __device__ inline Dual2<NP> operator+(double c, const Dual2<NP>& a) { Dual2<NP> r = a; r.v += c; return r; }
template <int NP> __device__ inline Dual2<NP> operator-(double c, const Dual2<NP>& a) { Dual2<NP> r; r.v = c - a.v; for (int i = 0; i < NP; ++i) r.g[i] = -a.g[i]; for (int i = 0; i < Dual2<NP>::NH; ++i) r.h[i] = -a.h[i]; return r; }
template <int NP> __device__ inline Dual2<NP> d_recip(const Dual2<NP>& x) { const double r = 1.0 / x.v; return d_unary(x, r, -r * r, 2.0 * r * r * r); }
template <int NP> __device__ inline Dual2<NP> operator/(const Dual2<NP>& a, const Dual2<NP>& b) { return a * d_recip(b); }
template <int NP> __device__ inline Dual2<NP> d_sin(const Dual2<NP>& x) { double s, c; sincos(x.v, &s, &c); return d_unary(x, s, c, -s); }
template <int NP> __device__ inline Dual2<NP> d_cos(const Dual2<NP>& x) { double s, c; sincos(x.v, &s, &c); return d_unary(x, c, -s, -c); }
template <int NP> __device__ inline Dual2<NP> d_sqrt(const Dual2<NP>& x) { const double s = sqrt(x.v); return d_unary(x, s, 0.5 / s, -0.25 / (s * x.v)); }
template <int NP> __device__ inline Dual2<NP> d_exp(const Dual2<NP>& x) { const double e = exp(x.v); return d_unary(x, e, e, e); }
// atan2(y, x): via u = y/x is singular at x = 0; use the gradient of atan2 directly: d = (x dy - y dx)/(x^2+y^2)
template <int NP> __device__ inline Dual2<NP> d_atan2(const Dual2<NP>& y, const Dual2<NP>& x) {
  const Dual2<NP> r2 = x * x + y * y;
  const Dual2<NP> ir2 = d_recip(r2);
  const Dual2<NP> ay = x * ir2;              // d atan2 / dy
  const Dual2<NP> ax = (-1.0) * (y * ir2);   // d atan2 / dx
  Dual2<NP> r; r.v = atan2(y.v, x.v);
  for (int i = 0; i < NP; ++i) r.g[i] = fma(ay.v, y.g[i], ax.v * x.g[i]);
  int k = 0;
  for (int i = 0; i < NP; ++i) for (int j = 0; j <= i; ++j, ++k)
    r.h[k] = ay.v * y.h[k] + ax.v * x.h[k] + ay.g[j] * y.g[i] + ax.g[j] * x.g[i];
  return r;
}

}  // namespace gamc

namespace gamc {

// =====================================================================================================
// Keplerian radial-velocity target (examples/radial_velocity/src/*.jl), block-cooperative over epochs
// =====================================================================================================
// calc_ecc_anom_itterative_laguerre (kepler_eqn.jl:43-56): Danby guess (:18-23) + Laguerre n = 5 (:30-40), M reduced mod 2 pi
// Returns E and its sine / cosine.  One full sincos at the starting guess; the later (small) Laguerre corrections update
// (sin E, cos E) by the angle-addition formulas with a short Taylor series for the correction angle.
__device__ inline double rv_kepler_sc(double M, double ecc, double& sE, double& cE) {
  const double twopi = 6.283185307179586476925286766559;
  M = M - twopi * floor(M * (1.0 / twopi));
  double E = (M < 3.14159265358979323846) ? M + 0.85 * ecc : M - 0.85 * ecc;
  sincos(E, &sE, &cE);
  for (int it = 0; it < 60; ++it) {
    const double es = ecc * sE, ec = ecc * cE;
    const double F = (E - es) - M, Fp = 1.0 - ec, Fpp = es;
    const double root = sqrt(fabs(4.0 * (4.0 * Fp * Fp - 5.0 * F * Fpp)));
    const double denom = (Fp > 0.0) ? Fp + root : Fp - root;
    const double d = -5.0 * F / denom;             // E_new = E + d
    E += d;
    if (fabs(d) <= 0.1) {
      const double d2 = d * d;
      const double sd = d * (1.0 - d2 * (1.0 / 6.0) * (1.0 - d2 * (1.0 / 20.0) * (1.0 - d2 * (1.0 / 42.0) * (1.0 - d2 * (1.0 / 72.0)))));
      const double cd = 1.0 - d2 * 0.5 * (1.0 - d2 * (1.0 / 12.0) * (1.0 - d2 * (1.0 / 30.0) * (1.0 - d2 * (1.0 / 56.0) * (1.0 - d2 * (1.0 / 90.0)))));
      const double sn = sE * cd + cE * sd;
      cE = cE * cd - sE * sd;
      sE = sn;
    } else {
      sincos(E, &sE, &cE);
    }
    if (fabs(d) < 1e-10) break;                    // reference: 1e-8 (kepler_eqn.jl:43); Laguerre converges cubically
  }
  return E;
}
__device__ inline double rv_kepler(double M, double ecc) { double s, c; return rv_kepler_sc(M, ecc, s, c); }

// The same solver for U independent epochs at once: the per-epoch work is one long dependent chain (sincos -> sqrt ->
// divide -> update), so a thread that interleaves several epochs hides most of that latency (instruction-level parallelism).
template <int U>
__device__ inline void rv_kepler_sc_batch(const double (&Min)[U], double ecc, double (&sE)[U], double (&cE)[U]) {
  const double twopi = 6.283185307179586476925286766559;
  double M[U], E[U];
  bool done[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    M[u] = Min[u] - twopi * floor(Min[u] * (1.0 / twopi));
    E[u] = (M[u] < 3.14159265358979323846) ? M[u] + 0.85 * ecc : M[u] - 0.85 * ecc;
    sincos(E[u], &sE[u], &cE[u]);
    done[u] = false;
  }
  for (int it = 0; it < 60; ++it) {
    // branch-free across the U epochs so that their dependent chains interleave; converged epochs take a zero step
    bool all = true, big = false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double es = ecc * sE[u], ec = ecc * cE[u];
      const double F = (E[u] - es) - M[u], Fp = 1.0 - ec, Fpp = es;
      const double root = sqrt(fabs(4.0 * (4.0 * Fp * Fp - 5.0 * F * Fpp)));
      const double denom = (Fp > 0.0) ? Fp + root : Fp - root;
      double d = -5.0 * F / denom;
      d = done[u] ? 0.0 : d;
      E[u] += d;
      const double d2 = d * d;
      const double sd = d * (1.0 - d2 * (1.0 / 6.0) * (1.0 - d2 * (1.0 / 20.0) * (1.0 - d2 * (1.0 / 42.0) * (1.0 - d2 * (1.0 / 72.0)))));
      const double cd = 1.0 - d2 * 0.5 * (1.0 - d2 * (1.0 / 12.0) * (1.0 - d2 * (1.0 / 30.0) * (1.0 - d2 * (1.0 / 56.0) * (1.0 - d2 * (1.0 / 90.0)))));
      const double sn = sE[u] * cd + cE[u] * sd;
      cE[u] = cE[u] * cd - sE[u] * sd;
      sE[u] = sn;
      big = big || (fabs(d) > 0.1);
      done[u] = done[u] || (fabs(d) < 1e-10);
      all = all && done[u];
    }
    if (big) {                                     // rare (first step from a poor guess): the series is not accurate enough
#pragma unroll
      for (int u = 0; u < U; ++u) sincos(E[u], &sE[u], &cE[u]);
    }
    if (all) break;
  }
}

// Quantities of one planet that do not depend on the epoch (rv_model_keplerian.jl:15-19,31-33), computed once per
// evaluation and shared through shared memory: plain values, and (for level 2) their second-order duals over the
// planet's nonlinear parameters (theta1 = log1p P, h, k, w + M0).
struct RvPlain { double K, Kp, k, ecc, w, n, j, rj1, wm, pad; };
struct RvDual { Dual2<4> ecc, w, n, j, rj1; };
constexpr int RV_CONST_DOUBLES = 2 * ((int)(sizeof(RvPlain) + sizeof(RvDual)) / 8);   // two planets

__device__ inline void rv_planet_constants(const double* th, RvPlain* pc, RvDual* dc, int level) {
  typedef Dual2<4> D;
  const double twopi = 6.283185307179586476925286766559;
  const double P = exp(th[0]) - 1.0, Kp = exp(th[1]), h = th[2], k = th[3];
  const double ecc = sqrt(h * h + k * k), w = atan2(k, h);
  const double j = sqrt((1.0 - ecc) * (1.0 + ecc));
  pc->K = Kp - 1.0; pc->Kp = Kp; pc->k = k; pc->ecc = ecc; pc->w = w; pc->n = twopi / P; pc->j = j; pc->rj1 = 1.0 / (1.0 + j); pc->wm = th[4];
  if (level >= 2) {
    const D Pd = d_exp(d_var<4>(th[0], 0)) + (-1.0);
    const D hd = d_var<4>(h, 1), kd = d_var<4>(k, 2);
    const D ed = d_sqrt(hd * hd + kd * kd);
    dc->ecc = ed;
    dc->w = d_atan2(kd, hd);
    dc->n = twopi * d_recip(Pd);
    const D jd = d_sqrt((1.0 - ed) * (1.0 + ed));
    dc->j = jd;
    dc->rj1 = d_recip(jd + 1.0);
  }
}

// velocity of one planet at epoch t, plain doubles (rv_model_keplerian.jl:20-35)
__device__ inline double rv_planet(const RvPlain& c, double t) {
  const double M = t * c.n - (c.wm - c.w), lam = c.w + M;
  double sE, cE;
  rv_kepler_sc(M, c.ecc, sE, cE);
  const double pe = c.ecc * sE, q = c.ecc * cE;
  return c.K * c.j / (1.0 - q) * (cos(lam + pe) - c.k * q * c.rj1);
}

// Phi(theta1, h, k, wpM0; t) = velocity / K with exact first and second derivatives.  The eccentric anomaly is solved in
// plain doubles; its derivatives follow from implicit differentiation of Kepler's equation E - e sin E = M:
//   D E_i = M_i + s e_i,                                   D = 1 - e cos E, s = sin E, c = cos E
//   D E_ij = M_ij + s e_ij + c (E_j e_i + E_i e_j) - e s E_i E_j
// (the reference differentiates through its Laguerre iterations, which converge to the same derivatives).
__device__ inline Dual2<4> rv_phi_dual(const RvPlain& pc, const RvDual& dc, double kval, double t) {
  typedef Dual2<4> D;
  const double twopi = 6.283185307179586476925286766559;
  D M = t * dc.n + dc.w;                 // M = t n - (wpM0 - w)
  M.v -= pc.wm; M.g[3] -= 1.0;
  const D lam = dc.w + M;
  (void)twopi;
  double s, c;
  const double E0 = rv_kepler_sc(M.v, pc.ecc, s, c);
  const double rD = 1.0 / (1.0 - pc.ecc * c);
  D E; E.v = E0;
#pragma unroll
  for (int i = 0; i < 4; ++i) E.g[i] = (M.g[i] + s * dc.ecc.g[i]) * rD;
  {
    int kk = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j, ++kk)
        E.h[kk] = (M.h[kk] + s * dc.ecc.h[kk] + c * (E.g[j] * dc.ecc.g[i] + E.g[i] * dc.ecc.g[j]) - pc.ecc * s * E.g[i] * E.g[j]) * rD;
  }
  const D sE = d_unary(E, s, c, -s), cE = d_unary(E, c, -s, -c);
  const D q = dc.ecc * cE;
  const D arg = lam + dc.ecc * sE;
  double sa, ca; sincos(arg.v, &sa, &ca);
  const D cosarg = d_unary(arg, ca, -sa, -ca);
  const D kd = d_var<4>(kval, 2);
  return dc.j * d_recip(1.0 - q) * (cosarg - kd * q * dc.rj1);
}

constexpr int RV_MAXP = 11;
constexpr int RV_MAXRED = 1 + RV_MAXP + RV_MAXP * (RV_MAXP + 1) / 2;   // 78

// Per-thread partial sums over this thread's epochs.  level 0: acc[0] = sum r^2/s^2 (chi-square).
// level 2: also acc[1+a] = sum r z_a / s^2 and acc[1+p+idx(a,b)] = sum (z_a z_b + r z_ab)/s^2 (lower triangle, row-major).
template <int NPL>
__device__ inline void rv_partial(const BatchedCfg& C, const double* th, const RvPlain* pc, const RvDual* dc, int level, int tid, int nt, double* acc) {
  constexpr int p = 5 * NPL + 1;
  constexpr int NR = 1 + p + p * (p + 1) / 2;
#pragma unroll
  for (int i = 0; i < NR; ++i) acc[i] = 0.0;
  const double offset = th[5 * NPL];
  long long e0 = tid;
  if (level < 2) {
    // log-target only (AM steps): GAMC_RV_BATCH independent epochs per trip
#ifndef GAMC_RV_BATCH
#define GAMC_RV_BATCH 2     // measured at 4 CTAs/SM on C5: 1 -> 82.7 k, 2 -> 83.2 k, 4 -> 71.0 k chain-iterations/s
#endif
    constexpr int U = GAMC_RV_BATCH;
    for (; e0 + (long long)(U - 1) * nt < C.rv_n; e0 += (long long)U * nt) {
      double t[U], z[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { t[u] = C.rv_t[e0 + (long long)u * nt]; z[u] = offset; }
#pragma unroll
      for (int pl = 0; pl < NPL; ++pl) {
        const RvPlain& c = pc[pl];
        double M[U], sE[U], cE[U];
#pragma unroll
        for (int u = 0; u < U; ++u) M[u] = t[u] * c.n - (c.wm - c.w);
        rv_kepler_sc_batch<U>(M, c.ecc, sE, cE);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const double pe = c.ecc * sE[u], q = c.ecc * cE[u];
          z[u] += c.K * c.j / (1.0 - q) * (cos(c.w + M[u] + pe) - c.k * q * c.rj1);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const double sg = C.rv_sig[e0 + (long long)u * nt];
        const double r = z[u] - C.rv_obs[e0 + (long long)u * nt];
        acc[0] = fma(r * r, 1.0 / (sg * sg), acc[0]);
      }
    }
  }
  for (long long e = e0; e < C.rv_n; e += nt) {
    const double t = C.rv_t[e], o = C.rv_obs[e], sg = C.rv_sig[e];
    const double is2 = 1.0 / (sg * sg);
    if (level < 2) {
      double z = offset;
#pragma unroll
      for (int pl = 0; pl < NPL; ++pl) z += rv_planet(pc[pl], t);
      const double r = z - o;
      acc[0] = fma(r * r, is2, acc[0]);
    } else {
      double za[p];
      double zab[NPL][15];       // per planet: second derivatives over (theta1, theta2, h, k, wpM0), lower triangle
      double z = offset;
      za[5 * NPL] = 1.0;
#pragma unroll
      for (int pl = 0; pl < NPL; ++pl) {
        const Dual2<4> phi = rv_phi_dual(pc[pl], dc[pl], pc[pl].k, t);
        const double Kp = pc[pl].Kp, K = pc[pl].K;                // K' = K'' = exp(theta2)
        z = fma(K, phi.v, z);
        // local order: 0 = theta1, 1 = theta2 (K), 2 = h, 3 = k, 4 = wpM0;  phi's variables: 0 = theta1, 1 = h, 2 = k, 3 = wpM0
        const int map[5] = {0, -1, 1, 2, 3};
#pragma unroll
        for (int a = 0; a < 5; ++a) za[5 * pl + a] = (a == 1) ? Kp * phi.v : K * phi.g[map[a]];
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b) {
            double v;
            if (a == 1 && b == 1) v = Kp * phi.v;
            else if (a == 1) v = Kp * phi.g[map[b]];
            else if (b == 1) v = Kp * phi.g[map[a]];
            else v = K * phi.h[Dual2<4>::hidx(map[a], map[b])];
            zab[pl][a * (a + 1) / 2 + b] = v;
          }
      }
      const double r = z - o;
      acc[0] = fma(r * r, is2, acc[0]);
      const double ris2 = r * is2;
#pragma unroll
      for (int a = 0; a < p; ++a) acc[1 + a] = fma(ris2, za[a], acc[1 + a]);
#pragma unroll
      for (int a = 0; a < p; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          double v = za[a] * za[b] * is2;
          if (a < 5 * NPL && (a / 5) == (b / 5)) { const int la = a % 5, lb = b % 5; v = fma(ris2, zab[a / 5][la * (la + 1) / 2 + lb], v); }
          acc[1 + p + a * (a + 1) / 2 + b] += v;
        }
    }
  }
}

// Block-wide RV evaluation; called by ALL threads of the CTA.  Results (level 2): lt returned to every thread,
// g (length p) and H = tensor = -Hessian of the log-target (p x ld) written by warp 0.  red = RV_MAXRED * nwarps doubles.
template <int NPL>
__device__ inline double rv_eval_block(const BatchedCfg& C, const double* x, double* g, double* H, int level, double* red, double* bcast, double* rvconst) {
  constexpr int p = 5 * NPL + 1;
  constexpr int NR = 1 + p + p * (p + 1) / 2;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  bool valid = true;
#pragma unroll
  for (int pl = 0; pl < NPL; ++pl) { const double h = x[5 * pl + 2], k = x[5 * pl + 3]; if (!(h * h + k * k < 1.0)) valid = false; }   // is_valid, rv_model_param.jl:83-96
  if (!valid) return -INFINITY;                                                     // stat_model.jl:21
  RvPlain* pc = reinterpret_cast<RvPlain*>(rvconst);
  RvDual* dc = reinterpret_cast<RvDual*>(rvconst + 2 * (sizeof(RvPlain) / 8));
  if (tid < NPL) rv_planet_constants(x + 5 * tid, pc + tid, dc + tid, level);
  __syncthreads();
  double acc[NR];
  rv_partial<NPL>(C, x, pc, dc, level, tid, nt, acc);
  const int nred = (level >= 2) ? NR : 1;
  for (int i = 0; i < nred; ++i) { const double v = warp_sum(acc[i]); if (lane == 0) red[warp * RV_MAXRED + i] = v; }
  __syncthreads();
  if (warp == 0) {
    for (int i = lane; i < nred; i += 32) { double s = 0.0; for (int w2 = 0; w2 < nw; ++w2) s += red[w2 * RV_MAXRED + i]; red[i] = s; }
    __syncwarp();
    if (lane == 0) {
      // log-likelihood normalisation (stat_model.jl:27-35) and modified-Jeffreys prior (:39-61; log prior = const - sum(theta_P + theta_K))
      double lognorm = C.rv_lognorm;
      double lp = -2.0 * log(6.283185307179586476925286766559);
      for (int pl = 0; pl < NPL; ++pl) lp -= x[5 * pl] + x[5 * pl + 1] + 2.0 * log(log1p(10000.0));
      bcast[0] = -0.5 * red[0] + lognorm + lp;
    }
    if (level >= 2) {
      if (lane < p) {
        double gv = -red[1 + lane];
        if (lane < 5 * NPL && (lane % 5) < 2) gv -= 1.0;          // d logprior / d theta_P, theta_K
        g[lane] = gv;
        for (int b = 0; b <= lane; ++b) { const double v = red[1 + p + lane * (lane + 1) / 2 + b]; H[lane * C.ld + b] = v; H[b * C.ld + lane] = v; }
      }
    }
  }
  __syncthreads();
  return bcast[0];
}

}  // namespace gamc

namespace gamc {

// =====================================================================================================
// the persistent batched-chains kernel
// =====================================================================================================
__device__ inline bool bt_schedule(const BatchedCfg& C, long long i, double u) {
  const double di = (double)i, dt = (double)C.nsteps;
  auto par = [](double v, double d) { return isnan(v) ? d : v; };
  switch (C.schedule) {
    case 0: return false;
    case 1: return true;
    case 2: { const long long n = (long long)par(C.sp0, 10.0); return n > 0 && (i % n) == 0; }
    case 3: return u <= par(C.sp0, 0.5);
    case 4: return u <= par(C.sp1, 0.2) * cos(par(C.sp0, 10.0) * di * 3.14159265358979323846 / dt) + par(C.sp2, 0.2);
    case 5: { const double b = par(C.sp1, 0.0); return u <= (1 - b) * exp(-par(C.sp0, 10.0) * di / dt) + b; }
    case 6: { const double b = par(C.sp1, 0.0); return u <= (1 - b) * (1 / (1 + par(C.sp0, 1e-2) * di)) + b; }
    case 7: { const double b = par(C.sp1, 0.0); return u <= (1 - b) * pow(par(C.sp0, 1e-3), di / dt) + b; }
    case 8: { const double b = par(C.sp1, 0.0); return u <= (1 - b) * (1 / (1 + par(C.sp0, 1e-5) * (di * di))) + b; }
    default: return true;
  }
}

// random numbers of (chain, iteration t): u_sched, u_mix, u_acc broadcast; z[p] into smem (warp 0)
__device__ inline void bt_draw(const BatchedCfg& C, int chain, long long t, long long tape_pos, double* z, double& us, double& um, double& ua, int lane) {
  const int p = C.p;
  const uint32_t gchain = (uint32_t)(C.chain_offset + chain);
  if (C.tape_z) {
    const size_t o = (size_t)chain * C.tape_len + tape_pos;
    us = C.tape_usched[o]; um = C.tape_umix[o]; ua = C.tape_uacc[o];
    if (lane < p) z[lane] = C.tape_z[o * p + lane];
  } else {
    const Philox4 r = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), gchain, 0x51u, (uint32_t)C.seed, (uint32_t)(C.seed >> 32));
    const Philox4 r2 = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), gchain, 0x52u, (uint32_t)C.seed, (uint32_t)(C.seed >> 32));
    us = philox_u53(r.x, r.y); um = philox_u53(r.z, r.w); ua = philox_u53(r2.x, r2.y);
    if (ua <= 0.0) ua = 1.0 / 9007199254740992.0;
    if (lane < p) {
      const Philox4 q = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), gchain, 0x100u + (uint32_t)lane, (uint32_t)C.seed, (uint32_t)(C.seed >> 32));
      double u1 = philox_u53(q.x, q.y); const double u2 = philox_u53(q.z, q.w);
      if (u1 <= 0.0) u1 = 1.0 / 9007199254740992.0;
      double sn, cs; sincospi(2.0 * u2, &sn, &cs);
      z[lane] = sqrt(-2.0 * log(u1)) * cs;
    }
  }
  __syncwarp();
}

// Target evaluation by warp 0 (mvt).  level: 0 = log-target, 2 = up to tensor.  On level 2 the point's slot
// (G, Inv, L, grad, firstterm, logdet of Inv) is filled: `transform`, inverse, Cholesky of the inverse, G^-1 grad --
// the reference's :20-30 / :37-56 sequence for one point.  Returns the log-target; *ok = 0 if anything is non-finite / not PD.
struct BtWork { double* red; double* bcast; int* cmd; double* rvconst; };

// master side of the block-cooperative RV evaluation: wake the worker warps, evaluate together
template <int NPL>
__device__ inline double rv_eval_master(const BatchedCfg& C, const BState& S, const BtWork& Wk, int xvec, int gradv, double* H, int level, int lane) {
  if (lane == 0) { Wk.cmd[0] = level + 1; Wk.cmd[1] = xvec; }
  __syncthreads();                                   // barrier A: releases the workers
  return rv_eval_block<NPL>(C, S.vec(xvec), S.vec(gradv), H, level, Wk.red, Wk.bcast, Wk.rvconst);
}

template <int TGT>
__device__ inline double bt_eval_point(const BatchedCfg& C, const BState& S, const BtWork& Wk, int xvec, int level, int gslot, int islot, int lslot,
                                       int gradv, int ftermv, int ldscal, int lane, int* ok) {
  const double* x = S.vec(xvec);
  const int p = C.p, ld = C.ld;
  double* W = S.mat(BM_W);
  double* G = S.mat(gslot); double* Inv = S.mat(islot); double* L = S.mat(lslot);
  double* g = S.vec(gradv);
  *ok = 1;
  double lt;
  if (TGT == 0) lt = mvt_eval(C, x, g, (level >= 2) ? W : nullptr, S.vec(BV_T1), level >= 2 ? 2 : 0, lane);
  else if (TGT == 3) lt = logi_eval(C, x, g, W, Wk.rvconst, level >= 2 ? 2 : 0, lane);
  else lt = rv_eval_master<(TGT == 2 ? 2 : 1)>(C, S, Wk, xvec, gradv, W, level >= 2 ? 2 : 0, lane);
  if (!isfinite(lt)) { *ok = 0; return lt; }
  if (level >= 2) {   // finite gradient / tensor (initialize! asserts, GAMC.jl:168-170; otherwise a rejection)
    double bad = 0.0;
    if (lane < p) { if (!isfinite(g[lane])) bad = 1.0; for (int j = 0; j < p; ++j) if (!isfinite(W[lane * ld + j])) bad = 1.0; }
    if (warp_sum(bad) != 0.0) { *ok = 0; return lt; }
  }
  if (level < 2) return lt;
  double logdetG;
  if (C.transform == 1) {
    // W holds tensor = -Hessian; softabs -> G, Inv (L used as the eigenvector scratch)
    logdetG = w_softabs(W, L, G, Inv, C.softabs_alpha, p, ld, lane, S.vec(BV_T2));
  } else {
    if (lane < p) for (int j = 0; j < p; ++j) { G[lane * ld + j] = W[lane * ld + j]; L[lane * ld + j] = W[lane * ld + j]; }
    __syncwarp();
    const double sl = w_chol(L, p, ld, lane);           // G = Lg Lg'
    if (isnan(sl)) { *ok = 0; return lt; }
    logdetG = 2.0 * sl;
    w_inv_from_chol(L, Inv, p, ld, lane);
  }
  if (!isfinite(logdetG)) { *ok = 0; return lt; }
  // cholinvtensor = chol(Hermitian(inv))'  (lower)
  if (lane < p) for (int j = 0; j < p; ++j) L[lane * ld + j] = (j <= lane) ? Inv[lane * ld + j] : 0.0;
  __syncwarp();
  const double sl2 = w_chol(L, p, ld, lane);
  if (isnan(sl2)) { *ok = 0; return lt; }
  w_matvec(Inv, g, S.vec(ftermv), p, ld, lane);         // firstterm = inv * grad
  if (lane == 0) S.s[ldscal] = -logdetG;                 // logdet(inv(G))
  __syncwarp();
  return lt;
}

__device__ inline void bt_tail(const BatchedCfg& C, const BState& S, int chain, int accepted, int lane) {
  const int p = C.p;
  double* s = S.s;
  const long long t = (long long)s[BS_COUNT];
  if (lane < p) {
    const double lm = S.vec(BV_M1)[lane];
    S.vec(BV_M2)[lane] = lm;
    S.vec(BV_M1)[lane] = ((double)(t - 1) * lm + S.vec(BV_THETA)[lane]) / (double)t;
  }
  if (t > C.burnin) {
    const long long pos = (long long)s[BS_SAVED];
    if (pos < C.chain_cap) {
      if (C.chain_value && lane < p) C.chain_value[((size_t)chain * C.chain_cap + pos) * p + lane] = S.vec(BV_THETA)[lane];
      if (C.chain_accept && lane == 0) C.chain_accept[(size_t)chain * C.chain_cap + pos] = (unsigned char)accepted;
    }
  }
  __syncwarp();
  if (lane == 0) {
    if (t > C.burnin && s[BS_SAVED] < (double)C.chain_cap) s[BS_SAVED] += 1.0;
    if (C.tuning) {
      if ((long long)s[BS_TTOT] <= C.burnin && ((long long)s[BS_TPROP] % C.period) == 0) {
        if (C.count_total) s[BS_TRATE] = s[BS_TACC] / s[BS_TPROP];
        if (C.is_accrate) { s[BS_STEP] *= 2.0 / (1.0 + exp(-C.score_k * (s[BS_TRATE] - C.targetrate))); s[BS_SQSTEP] = sqrt(s[BS_STEP]); }
        if (C.verbose_sm) { s[BS_SMTOT] += s[BS_SMPROP]; s[BS_SMPROP] = 0; s[BS_SMACC] = 0; }
        if (C.verbose_am) { s[BS_AMTOT] += s[BS_AMPROP]; s[BS_AMPROP] = 0; s[BS_AMACC] = 0; }
        s[BS_TTOT] += s[BS_TPROP]; s[BS_TPROP] = 0;
        if (C.count_total) { s[BS_TACC] = 0; s[BS_TRATE] = nan(""); }
      }
    }
  }
  __syncwarp();
}

// mode: 0 = run n_iters transitions, 1 = initialise (initialize! + sampler_state) from theta0 [p x nchains]
__host__ __device__ inline size_t bt_smem_bytes(int blob_doubles, int nthreads, int extra_doubles = 0) {
  return sizeof(double) * ((size_t)blob_doubles + (size_t)RV_MAXRED * (nthreads / 32) + 8 + (nthreads > 32 ? RV_CONST_DOUBLES : 0) + (size_t)extra_doubles);
}

template <int TGT>
#ifndef GAMC_RV_MINBLOCKS
#define GAMC_RV_MINBLOCKS 4     // resident CTAs per SM the RV kernels are compiled for: 128 registers (spills) but twice the warps of the uncapped build; measured 51.9 k (2 CTAs/SM) -> 60.9 k (3) -> 70.9 k (4) chain-iterations/s on C5
#endif
__global__ void __launch_bounds__((TGT == 0 || TGT == 3) ? 32 : 128, (TGT == 0 || TGT == 3) ? 1 : GAMC_RV_MINBLOCKS) k_batched_gamc(BatchedCfg C, int mode, long long n_iters, const double* __restrict__ theta0) {
  extern __shared__ __align__(16) double bsm[];
  const int chain = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (chain >= C.nchains) return;
  const int p = C.p, ld = C.ld;
  BState S{bsm, bsm + BS_NSCAL, bsm + BS_NSCAL + BV_NVEC * BT_MAXP, p, ld};
  double* s = S.s;
  double* blob = C.blob + (size_t)chain * C.blob_doubles;
  BtWork Wk;
  Wk.red = bsm + C.blob_doubles;
  Wk.bcast = Wk.red + RV_MAXRED * (blockDim.x >> 5);
  Wk.cmd = reinterpret_cast<int*>(Wk.bcast + 4);
  Wk.rvconst = Wk.bcast + 8;
  constexpr bool RV = (TGT == 1 || TGT == 2);
  if (RV && warp != 0) {
    // worker warps: wait for an evaluation request from warp 0, take part in it, repeat
    while (true) {
      __syncthreads();                               // barrier A
      const int c = Wk.cmd[0];
      if (c == 0) break;
      rv_eval_block<(TGT == 2 ? 2 : 1)>(C, S.vec(Wk.cmd[1]), nullptr, nullptr, c - 1, Wk.red, Wk.bcast, Wk.rvconst);
    }
    return;
  }
  auto release_workers = [&]() { if (RV) { if (lane == 0) Wk.cmd[0] = 0; __syncthreads(); } };
  if (mode == 1) {
    for (int i = lane; i < C.blob_doubles; i += 32) bsm[i] = 0.0;
    __syncwarp();
    if (lane < p) { S.vec(BV_THETA)[lane] = theta0[(size_t)chain * p + lane]; S.vec(BV_M1)[lane] = S.vec(BV_THETA)[lane]; }
    __syncwarp();
    int ok;
    const double lt = bt_eval_point<TGT>(C, S, Wk, BV_THETA, 2, BM_GC, BM_IC, BM_LC, BV_GC, BV_FC, BS_LDC, lane, &ok);
    if (lane == 0) {
      s[BS_LT] = lt; s[BS_STEP] = C.driftstep; s[BS_SQSTEP] = sqrt(C.driftstep);
      s[BS_TTOT] = C.period; s[BS_SMTOT] = C.period; s[BS_AMTOT] = C.period; s[BS_TRATE] = nan("");
      s[BS_PRESENT] = 1.0; s[BS_PAST] = 0.0; s[BS_ERR] = ok ? 0.0 : 2.0;   // GAMC_NONFINITE_INIT / not PD at the initial point
    }
    __syncwarp();
    for (int i = lane; i < C.blob_doubles; i += 32) blob[i] = bsm[i];
    release_workers();
    return;
  }
  for (int i = lane; i < C.blob_doubles; i += 32) bsm[i] = blob[i];
  __syncwarp();
  double* theta = S.vec(BV_THETA); double* prop = S.vec(BV_PROP); double* z = S.vec(BV_Z);
  for (long long it = 0; it < n_iters; ++it) {
    if (s[BS_ERR] != 0.0) break;
    const long long t = (long long)s[BS_COUNT] + 1;
    double us, um, ua;
    bt_draw(C, chain, t, it, z, us, um, ua, lane);
    bool present = s[BS_PRESENT] != 0.0;
    if (t > C.t0) present = bt_schedule(C, t, us);                                 // iterate/GAMC.jl:8-10
    const bool past = s[BS_PAST] != 0.0;
    __syncwarp();
    if (lane == 0) {
      s[BS_COUNT] = (double)t;
      if (C.tuning) s[BS_TPROP] += 1.0;
      s[BS_PRESENT] = present ? 1.0 : 0.0;
    }
    const double eps = s[BS_STEP], sq = s[BS_SQSTEP];
    int accept = 0;
    if (present) {                                                                 // ---- SMMALA branch (:12-127)
      if (lane == 0) { s[BS_UPD] += 1.0; if (C.verbose_sm) s[BS_SMPROP] += 1.0; }
      int ok = 1;
      if (!past) {                                                                 // :19-31
        const double lt = bt_eval_point<TGT>(C, S, Wk, BV_THETA, 2, BM_GC, BM_IC, BM_LC, BV_GC, BV_FC, BS_LDC, lane, &ok);
        if (lane == 0) s[BS_LT] = lt;
        if (!ok) { if (lane == 0) s[BS_ERR] = 3.0; __syncwarp(); break; }
      }
      __syncwarp();
      if (lane < p) S.vec(BV_MU)[lane] = theta[lane] + 0.5 * eps * S.vec(BV_FC)[lane];       // :33
      __syncwarp();
      w_trmv_lower(S.mat(BM_LC), z, S.vec(BV_T1), p, ld, lane);
      if (lane < p) prop[lane] = S.vec(BV_MU)[lane] + sq * S.vec(BV_T1)[lane];               // :35
      __syncwarp();
      const double lt_p = bt_eval_point<TGT>(C, S, Wk, BV_PROP, 2, BM_GP, BM_IP, BM_LP, BV_GP, BV_FP, BS_LDP, lane, &ok);   // :37-43,56
      double ratio = -INFINITY;
      if (ok) {
        // ratio (:45-67): logdet(eps*inv) = p log eps + logdet(inv)
        if (lane < p) S.vec(BV_T1)[lane] = prop[lane] - S.vec(BV_MU)[lane];
        __syncwarp();
        w_matvec(S.mat(BM_GC), S.vec(BV_T1), S.vec(BV_T2), p, ld, lane);
        const double q1 = w_dot(S.vec(BV_T1), S.vec(BV_T2), p, lane);
        if (lane < p) S.vec(BV_T1)[lane] = theta[lane] - (prop[lane] + 0.5 * eps * S.vec(BV_FP)[lane]);
        __syncwarp();
        w_matvec(S.mat(BM_GP), S.vec(BV_T1), S.vec(BV_T2), p, ld, lane);
        const double q2 = w_dot(S.vec(BV_T1), S.vec(BV_T2), p, lane);
        const double plog = p * log(eps);
        ratio = lt_p - s[BS_LT];
        ratio += 0.5 * ((plog + s[BS_LDC]) + q1 / eps);
        ratio -= 0.5 * ((plog + s[BS_LDP]) + q2 / eps);
      }
      accept = (ratio > 0.0) || (ratio > log(ua));                                 // :69
      __syncwarp();
      if (accept) {
        // copy the proposal slot into the current slot (:70-98)
        for (int e = lane; e < p * ld; e += 32) { S.mat(BM_GC)[e] = S.mat(BM_GP)[e]; S.mat(BM_IC)[e] = S.mat(BM_IP)[e]; S.mat(BM_LC)[e] = S.mat(BM_LP)[e]; }
        if (lane < p) { theta[lane] = prop[lane]; S.vec(BV_GC)[lane] = S.vec(BV_GP)[lane]; S.vec(BV_FC)[lane] = S.vec(BV_FP)[lane]; }
        if (lane == 0) { s[BS_LT] = lt_p; s[BS_LDC] = s[BS_LDP]; if (C.count_total) s[BS_TACC] += 1.0; if (C.verbose_sm) s[BS_SMACC] += 1.0; }
      }
      if (lane == 0) s[BS_RATIO] = ratio;
      __syncwarp();
    } else {                                                                       // ---- AM branch (:128-191)
      if (lane == 0 && C.verbose_am) s[BS_AMPROP] += 1.0;
      const double k = (double)(t - 2);
      double* Cm = S.mat(BM_IC);                                                   // sstate.oldinvtensor
      double* Lam = S.mat(BM_LAM);
      if (lane < p) {                                                              // covariance! (:133-140) + Hermitian (:142)
        const double xi = theta[lane], m1i = S.vec(BV_M1)[lane], m2i = S.vec(BV_M2)[lane];
        for (int j = 0; j <= lane; ++j) {
          const double cij = ((k - 1.0) * Cm[lane * ld + j] + k * (m2i * S.vec(BV_M2)[j]) - (k + 1.0) * (m1i * S.vec(BV_M1)[j]) + xi * theta[j]) / k;
          Lam[lane * ld + j] = eps * cij;
          S.mat(BM_W)[lane * ld + j] = cij;
        }
      }
      __syncwarp();
      if (lane < p) for (int j = 0; j <= lane; ++j) { const double cij = S.mat(BM_W)[lane * ld + j]; Cm[lane * ld + j] = cij; Cm[j * ld + lane] = cij; }
      __syncwarp();
      const double sl = w_chol(Lam, p, ld, lane);                                  // MvNormal(theta, eps*C) (set_gmm, GAMC.jl:188)
      if (isnan(sl)) { if (lane == 0) s[BS_ERR] = 3.0; __syncwarp(); break; }
      if (um <= 1.0 - C.c) {                                                       // :146 core component
        w_trmv_lower(Lam, z, S.vec(BV_T1), p, ld, lane);
        if (lane < p) prop[lane] = theta[lane] + S.vec(BV_T1)[lane];
      } else if (lane < p) prop[lane] = theta[lane] + C.sqrtminorscale * z[lane];
      __syncwarp();
      int ok;
      const double lt_p = bt_eval_point<TGT>(C, S, Wk, BV_PROP, 0, BM_GP, BM_IP, BM_LP, BV_GP, BV_FP, BS_LDP, lane, &ok);   // :148
      const double ratio = ok ? (lt_p - s[BS_LT]) : -INFINITY;                     // GMM terms cancel exactly (:150-156)
      accept = (ratio > 0.0) || (ratio > log(ua));                                 // :158
      __syncwarp();
      if (accept) {
        if (lane < p) theta[lane] = prop[lane];
        if (lane == 0) { s[BS_LT] = lt_p; if (C.count_total) s[BS_TACC] += 1.0; if (C.verbose_am) s[BS_AMACC] += 1.0; }
      }
      if (lane == 0) s[BS_RATIO] = ratio;
      __syncwarp();
    }
    if (lane == 0) s[BS_PAST] = present ? 1.0 : 0.0;                               // :197
    __syncwarp();
    bt_tail(C, S, chain, accept, lane);
  }
  __syncwarp();
  for (int i = lane; i < C.blob_doubles; i += 32) blob[i] = bsm[i];
  release_workers();
}

}  // namespace gamc
