// bigd.cuh -- batched GAMC chains of the multivariate-t target in LARGE dimension (33 <= d <= 1024; BASELINE config 4: d = 1024,
// 4096 chains per GPU).  Reference: examples/t_distribution/src/GAMC/simulations.jl:17-39 (target, `transform = softabs(H, 1000)`),
// :56-79 (the sequential chain loop), and the transition src/samplers/iterate/GAMC.jl:1-293.
//
// The reference handles this target densely: ForwardDiff Hessian (O(d^2) duals), `eig` inside softabs, `inv`, `chol`, two `logdet`s
// per SMMALA step -- O(d^3) per chain and step, hopeless at d = 1024 x 4096 chains.  Here the STRUCTURE of the target carries the
// SMMALA branch (tests/bigd_ref.py states the algebra in NumPy; tests/test_bigd_math.py checks it against the dense oracle):
//   * T = inv(Sigma_t) is tridiagonal; tensor = -Hessian = a (T - beta v v'), v = T x           (rank-one modification)
//   * softabs changes only the eigenvalues below 21/alpha: G = a T + V S V', V = [lowest eigenvectors ..., v]. The eigenvalues
//     come from bisection on the inertia count of the LDL' recurrence of T - mu plus the sign of the secular function, the
//     eigenvectors from (T - mu)^-1 v: O(d) per sweep, no dense matrix, no eigen-decomposition
//   * reverse Cholesky factor of G in product form  Lm = sqrt(a) L_T K_1 ... K_m,  K = chol(I + W S W') from a row sweep with a
//     32 x 32 state held in the registers of one warp (lane = row of the state)
//   * G^-1 grad, chol(G^-1) z, logdet and the quadratic forms of the acceptance ratio are recurrences over the d rows
// ONE WARP PER CHAIN runs all of that: 4096 chains = 4096 warps, all resident at once.
// The AM branch needs the dense Cholesky factor of the proposal covariance (d x d per chain, 8 MB at d = 1024: the HBM-resident
// part of the state, 32 GB for 4096 chains): seeded from the structure at the SMMALA -> AM hand-over (bd_handover), then updated
// by the same closed-form rank-one sweep as the single-chain path (sampler.cuh k_am_pre_r1), one CTA per chain.
#pragma once
#include "common.cuh"
#include "synth.cuh"

namespace gamc {

constexpr int BD_CH = 32;        // columns per product factor (= warp width: lane <-> column)
constexpr int BD_WPB = 4;        // chains (warps) per CTA in the warp-per-chain kernels
constexpr int BD_MAXD = 1024;
constexpr int BD_NBISECT = 60;

enum { BDV_THETA = 0, BDV_PROP, BDV_MU, BDV_M1, BDV_M2, BDV_G0, BDV_G1, BDV_F0, BDV_F1, BDV_V0, BDV_V1, BDV_N };   // G/F/V: per evaluation slot
enum { BDS_LT = 0, BDS_LTP, BDS_A0, BDS_A1, BDS_BETA0, BDS_BETA1, BDS_SL0, BDS_SL1, BDS_NCH0, BDS_NCH1, BDS_OK1, BDS_CUR, BDS_STEP, BDS_SQSTEP,
       BDS_RATIO, BDS_COUNT, BDS_UPD, BDS_TACC, BDS_TPROP, BDS_TTOT, BDS_TRATE, BDS_SMACC, BDS_SMPROP, BDS_SMTOT, BDS_AMACC, BDS_AMPROP,
       BDS_AMTOT, BDS_PAST, BDS_PRESENT, BDS_ERR, BDS_SAVED, BDS_US, BDS_UM, BDS_UA, BDS_R0, BDS_R1, BDS_N = 40 };

struct BdCfg {
  int d, nchains, ncmax;
  long long chain_offset;
  int transform; double alpha;
  // target: natural-order tridiagonal T = inv(Sigma_t) (td, te), bidiagonal Cholesky factor of the index-reversed T (ld, le, 1/ld)
  const double* td; const double* te; const double* ld; const double* le; const double* rld;
  const double* poles;           // eigenvalues of T, ascending (host, once): the secular roots interlace them
  double nu, logconst, sumlog_ld, gersh_lo, gersh_hi;
  // sampler / tuner / range (same meaning as gamc_config)
  double driftstep, minorscale, sqrtminorscale, c;
  int t0, schedule; double sp0, sp1, sp2;
  int tuning, count_total, is_accrate, verbose_sm, verbose_am, period;
  long long nsteps, burnin;
  double targetrate, score_k;
  // tape (null -> Philox with `seed`): [chain][t], z [chain][t][d]
  const double* tape_usched; const double* tape_umix; const double* tape_uacc; const double* tape_z;
  long long tape_len; unsigned long long seed;
  // state
  double* vecs;                  // [nchains][BDV_N][d]
  double* scal;                  // [nchains][BDS_N]
  double* W; double* B;          // [nchains][2][ncmax][d][32]   product-factor generators of the two evaluation slots
  double* delta;                 // [nchains][2][ncmax][d]
  double* S;                     // [nchains][2][ncmax][32]
  double* L; double* rdiagL;     // [nchains][d*d] dense AM factor (natural order, column-major, lower), [nchains][d]
  double* chain_value; unsigned char* chain_accept; long long chain_cap;   // [chain][cap][d], [chain][cap]
};

struct BdChain {                 // per-warp view of one chain
  const BdCfg* C; int chain, lane, d;
  double* s;
  __device__ double* vec(int k) const { return C->vecs + ((size_t)chain * BDV_N + k) * d; }
  __device__ size_t slot_chunk(int slot, int k) const { return ((size_t)chain * 2 + slot) * C->ncmax + k; }
  __device__ double* Wc(int slot, int k) const { return C->W + slot_chunk(slot, k) * (size_t)d * BD_CH; }
  __device__ double* Bc(int slot, int k) const { return C->B + slot_chunk(slot, k) * (size_t)d * BD_CH; }
  __device__ double* Dc(int slot, int k) const { return C->delta + slot_chunk(slot, k) * (size_t)d; }
  __device__ double* Sc(int slot, int k) const { return C->S + slot_chunk(slot, k) * BD_CH; }
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// index-reversed ("M space") view of the tridiagonal T: row i' = d-1-i
__device__ __forceinline__ double bd_tdm(const BdCfg& C, int i) { return C.td[C.d - 1 - i]; }
__device__ __forceinline__ double bd_tem(const BdCfg& C, int i) { return C.te[C.d - 2 - i]; }   // between rows i and i+1 of M space

// s(lam) - lam for Klara's softabs s(lam) = lam coth(alpha lam) (1/alpha at 0), accurate for every sign and size of lam
__device__ inline double bd_softabs_gap(double lam, double alpha) {
  const double x = alpha * lam;
  if (fabs(x) < 1e-8) return 1.0 / alpha - lam;
  if (x > 0.0) return x < 350.0 ? 2.0 * lam / expm1(2.0 * x) : 0.0;
  return -2.0 * lam + (x > -350.0 ? (-2.0 * lam) / expm1(-2.0 * x) : 0.0);
}

// ---- random numbers of (chain, iteration): same conventions as the small-p batched kernel (batched.cuh bt_draw) ----------------
__device__ inline void bd_uniforms(const BdCfg& C, int chain, long long t, long long tape_pos, double& us, double& um, double& ua) {
  if (C.tape_z) {
    const size_t o = (size_t)chain * C.tape_len + tape_pos;
    us = C.tape_usched[o]; um = C.tape_umix[o]; ua = C.tape_uacc[o];
    return;
  }
  const uint32_t gchain = (uint32_t)(C.chain_offset + chain);
  const Philox4 r = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), gchain, 0x51u, (uint32_t)C.seed, (uint32_t)(C.seed >> 32));
  const Philox4 r2 = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), gchain, 0x52u, (uint32_t)C.seed, (uint32_t)(C.seed >> 32));
  us = philox_u53(r.x, r.y); um = philox_u53(r.z, r.w); ua = philox_u53(r2.x, r2.y);
  if (ua <= 0.0) ua = 1.0 / 9007199254740992.0;
}
__device__ inline double bd_normal(const BdCfg& C, int chain, long long t, long long tape_pos, int i) {
  if (C.tape_z) return C.tape_z[((size_t)chain * C.tape_len + tape_pos) * C.d + i];
  const uint32_t gchain = (uint32_t)(C.chain_offset + chain);
  const Philox4 q = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), gchain, 0x100u + (uint32_t)i, (uint32_t)C.seed, (uint32_t)(C.seed >> 32));
  double u1 = philox_u53(q.x, q.y); const double u2 = philox_u53(q.z, q.w);
  if (u1 <= 0.0) u1 = 1.0 / 9007199254740992.0;
  double sn, cs; sincospi(2.0 * u2, &sn, &cs);
  return sqrt(-2.0 * log(u1)) * cs;
}

__device__ inline bool bd_schedule(const BdCfg& C, long long i, double u) {       // src/samplers/GAMC.jl:264-337
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

// =====================================================================================================
// warp-level pieces (all 32 lanes call; xs / vm are per-warp shared-memory vectors of length d in M-space order)
// =====================================================================================================
// y = T x in natural order (global vectors), returns q = x'y to every lane
__device__ inline double bd_tmatvec(const BdCfg& C, const double* x, double* y, int lane) {
  const int d = C.d;
  double q = 0.0;
  for (int i = lane; i < d; i += 32) {
    double v = C.td[i] * x[i];
    if (i > 0) v = fma(C.te[i - 1], x[i - 1], v);
    if (i + 1 < d) v = fma(C.te[i], x[i + 1], v);
    y[i] = v;
    q = fma(x[i], v, q);
  }
  return warp_sum(q);
}

// One LDL' sweep of T' - mu (each lane its own mu): number of eigenvalues of A = T' - beta v v' below mu = negative pivots +
// [f(mu) < 0], f the secular function 1 - beta v'(T'-mu)^-1 v.
__device__ inline int bd_count(const BdCfg& C, const double* vm, double beta, double mu) {
  const int d = C.d;
  const double pivmin = 1e-290;
  double D = bd_tdm(C, 0) - mu;
  if (fabs(D) < pivmin) D = -pivmin;
  double rD = 1.0 / D, z = vm[0];
  int neg = D < 0.0;
  double acc = z * z * rD;
  for (int i = 1; i < d; ++i) {
    const double e = bd_tem(C, i - 1);
    const double l = e * rD;
    D = (bd_tdm(C, i) - mu) - l * e;
    if (fabs(D) < pivmin) D = -pivmin;
    rD = 1.0 / D;
    z = fma(-l, z, vm[i]);
    neg += D < 0.0;
    acc = fma(z * z, rD, acc);
  }
  return neg + ((1.0 - beta * acc) < 0.0 ? 1 : 0);
}

// The secular function f(mu) = 1 - beta v'(T'-mu)^-1 v itself (each lane its own mu), from the same sweep.
__device__ inline double bd_secular(const BdCfg& C, const double* vm, double beta, double mu) {
  const int d = C.d;
  const double pivmin = 1e-290;
  double D = bd_tdm(C, 0) - mu;
  if (fabs(D) < pivmin) D = -pivmin;
  double rD = 1.0 / D, z = vm[0];
  double acc = z * z * rD;
  for (int i = 1; i < d; ++i) {
    const double e = bd_tem(C, i - 1);
    const double l = e * rD;
    D = (bd_tdm(C, i) - mu) - l * e;
    if (fabs(D) < pivmin) D = -pivmin;
    rD = 1.0 / D;
    z = fma(-l, z, vm[i]);
    acc = fma(z * z, rD, acc);
  }
  return 1.0 - beta * acc;
}

// (sqrt(a) L_T) u = b in place on the warp's vector (first-order recurrence; every lane runs it, lane 0 stores)
__device__ inline void bd_bidiag_solve(const BdCfg& C, double sa, double* xs, int lane) {
  const int d = C.d;
  const double rsa = 1.0 / sa;
  __syncwarp();
  if (lane == 0) {
    double u = xs[0] * rsa * C.rld[0];
    xs[0] = u;
    for (int i = 1; i < d; ++i) { u = (xs[i] * rsa - C.le[i - 1] * u) * C.rld[i]; xs[i] = u; }
  }
  __syncwarp();
}
// (sqrt(a) L_T)' y = b
__device__ inline void bd_bidiag_solve_t(const BdCfg& C, double sa, double* xs, int lane) {
  const int d = C.d;
  const double rsa = 1.0 / sa;
  __syncwarp();
  if (lane == 0) {
    double y = xs[d - 1] * rsa * C.rld[d - 1];
    xs[d - 1] = y;
    for (int i = d - 2; i >= 0; --i) { y = (xs[i] * rsa - C.le[i] * y) * C.rld[i]; xs[i] = y; }
  }
  __syncwarp();
}
// y = (sqrt(a) L_T)' x (no recurrence)
__device__ inline void bd_bidiag_mult_t(const BdCfg& C, double sa, double* xs, int lane) {
  const int d = C.d;
  __syncwarp();
  double mine[BD_MAXD / 32];
#pragma unroll
  for (int k = 0; k < BD_MAXD / 32; ++k) {
    const int i = lane + 32 * k;
    mine[k] = (i < d) ? sa * (C.ld[i] * xs[i] + (i + 1 < d ? C.le[i] * xs[i + 1] : 0.0)) : 0.0;
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < BD_MAXD / 32; ++k) { const int i = lane + 32 * k; if (i < d) xs[i] = mine[k]; }
  __syncwarp();
}

// The generator rows of a factor stream from HBM (256 B per row and matrix, one row per lane-wide load), and every recurrence
// below walks the d rows in order: to keep the walks off the memory latency, rows are fetched in blocks of BD_PF and the NEXT
// block's loads are issued before the current block is consumed (register double buffering).
constexpr int BD_PF = 8;
struct BdRows { double w[BD_PF], b[BD_PF], rs[BD_PF]; };
// rows i0 .. i0+BD_PF-1 (ascending, dir = +1) or i0 .. i0-BD_PF+1 (descending, dir = -1), clamped to [0, d)
__device__ __forceinline__ void bd_fetch(BdRows& R, const double* __restrict__ W, const double* __restrict__ B, const double* __restrict__ dl,
                                         int d, int i0, int dir, int lane) {
#pragma unroll
  for (int u = 0; u < BD_PF; ++u) {
    const int i = min(max(i0 + dir * u, 0), d - 1);
    R.w[u] = W[(size_t)i * BD_CH + lane]; R.b[u] = B[(size_t)i * BD_CH + lane]; R.rs[u] = dl[i];
  }
}

// K y = u in place: y_i = (u_i - w_i'sigma)/sqrt(delta_i), sigma += b_i y_i   (lane <-> component of sigma)
__device__ inline void bd_k_solve(const double* __restrict__ W, const double* __restrict__ B, const double* __restrict__ dl, int d, double* xs, int lane) {
  double sig = 0.0;
  BdRows nxt;
  bd_fetch(nxt, W, B, dl, d, 0, 1, lane);
  for (int i0 = 0; i0 < d; i0 += BD_PF) {
    const BdRows cur = nxt;
    if (i0 + BD_PF < d) bd_fetch(nxt, W, B, dl, d, i0 + BD_PF, 1, lane);
#pragma unroll
    for (int u = 0; u < BD_PF; ++u) {
      if (i0 + u < d) {
        const double dot = warp_sum(cur.w[u] * sig);
        const double y = (xs[i0 + u] - dot) * rsqrt(cur.rs[u]);
        sig = fma(cur.b[u], y, sig);
        if (lane == 0) xs[i0 + u] = y;
      }
    }
  }
  __syncwarp();
}
// K' w = b in place (backward): w_j = (b_j - b_j'tau)/sqrt(delta_j), tau += w_j (generator row) w_j
__device__ inline void bd_k_solve_t(const double* __restrict__ W, const double* __restrict__ B, const double* __restrict__ dl, int d, double* xs, int lane) {
  double tau = 0.0;
  BdRows nxt;
  bd_fetch(nxt, W, B, dl, d, d - 1, -1, lane);
  for (int j0 = d - 1; j0 >= 0; j0 -= BD_PF) {
    const BdRows cur = nxt;
    if (j0 - BD_PF >= 0) bd_fetch(nxt, W, B, dl, d, j0 - BD_PF, -1, lane);
#pragma unroll
    for (int u = 0; u < BD_PF; ++u) {
      if (j0 - u >= 0) {
        const double dot = warp_sum(cur.b[u] * tau);
        const double y = (xs[j0 - u] - dot) * rsqrt(cur.rs[u]);
        tau = fma(cur.w[u], y, tau);
        if (lane == 0) xs[j0 - u] = y;
      }
    }
  }
  __syncwarp();
}
// out = K' w in place (backward): out_j = sqrt(delta_j) w_j + b_j'tau, tau += w_j (generator row) w_j
__device__ inline void bd_k_mult_t(const double* __restrict__ W, const double* __restrict__ B, const double* __restrict__ dl, int d, double* xs, int lane) {
  double tau = 0.0;
  BdRows nxt;
  bd_fetch(nxt, W, B, dl, d, d - 1, -1, lane);
  for (int j0 = d - 1; j0 >= 0; j0 -= BD_PF) {
    const BdRows cur = nxt;
    if (j0 - BD_PF >= 0) bd_fetch(nxt, W, B, dl, d, j0 - BD_PF, -1, lane);
#pragma unroll
    for (int u = 0; u < BD_PF; ++u) {
      if (j0 - u >= 0) {
        const double x = xs[j0 - u];
        const double dot = warp_sum(cur.b[u] * tau);
        tau = fma(cur.w[u], x, tau);
        if (lane == 0) xs[j0 - u] = fma(sqrt(cur.rs[u]), x, dot);
      }
    }
  }
  __syncwarp();
}

// The same three recurrences for a factor with at most NC live generator columns (NC = 4: the typical-set regime, where a factor
// holds the lowest eigenvector and v).  Every lane carries the whole NC-vector and reads the generator rows as broadcasts, so a
// row step is a short chain of FMAs with no shuffle reduction on it (about four times shorter than the 32-column version).
template <int NC>
__device__ inline void bd_k_solve_nc(const double* __restrict__ W, const double* __restrict__ B, const double* __restrict__ dl, int d, double* xs, int lane) {
  double sig[NC];
#pragma unroll
  for (int m = 0; m < NC; ++m) sig[m] = 0.0;
  __syncwarp();
  for (int i0 = 0; i0 < d; i0 += 4) {
    double w[4][NC], b[4][NC], rs[4], xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = min(i0 + u, d - 1);
#pragma unroll
      for (int m = 0; m < NC; ++m) { w[u][m] = W[(size_t)i * BD_CH + m]; b[u][m] = B[(size_t)i * BD_CH + m]; }
      rs[u] = rsqrt(dl[i]); xv[u] = xs[i];
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + u < d) {
        double dot = 0.0;
#pragma unroll
        for (int m = 0; m < NC; ++m) dot = fma(w[u][m], sig[m], dot);
        const double y = (xv[u] - dot) * rs[u];
#pragma unroll
        for (int m = 0; m < NC; ++m) sig[m] = fma(b[u][m], y, sig[m]);
        if (lane == 0) xs[i0 + u] = y;
      }
    }
  }
  __syncwarp();
}
template <int NC>
__device__ inline void bd_k_solve_t_nc(const double* __restrict__ W, const double* __restrict__ B, const double* __restrict__ dl, int d, double* xs, int lane) {
  double tau[NC];
#pragma unroll
  for (int m = 0; m < NC; ++m) tau[m] = 0.0;
  __syncwarp();
  for (int j0 = d - 1; j0 >= 0; j0 -= 4) {
    double w[4][NC], b[4][NC], rs[4], xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = max(j0 - u, 0);
#pragma unroll
      for (int m = 0; m < NC; ++m) { w[u][m] = W[(size_t)j * BD_CH + m]; b[u][m] = B[(size_t)j * BD_CH + m]; }
      rs[u] = rsqrt(dl[j]); xv[u] = xs[j];
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (j0 - u >= 0) {
        double dot = 0.0;
#pragma unroll
        for (int m = 0; m < NC; ++m) dot = fma(b[u][m], tau[m], dot);
        const double y = (xv[u] - dot) * rs[u];
#pragma unroll
        for (int m = 0; m < NC; ++m) tau[m] = fma(w[u][m], y, tau[m]);
        if (lane == 0) xs[j0 - u] = y;
      }
    }
  }
  __syncwarp();
}
template <int NC>
__device__ inline void bd_k_mult_t_nc(const double* __restrict__ W, const double* __restrict__ B, const double* __restrict__ dl, int d, double* xs, int lane) {
  double tau[NC];
#pragma unroll
  for (int m = 0; m < NC; ++m) tau[m] = 0.0;
  __syncwarp();
  for (int j0 = d - 1; j0 >= 0; j0 -= 4) {
    double w[4][NC], b[4][NC], sd[4], xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = max(j0 - u, 0);
#pragma unroll
      for (int m = 0; m < NC; ++m) { w[u][m] = W[(size_t)j * BD_CH + m]; b[u][m] = B[(size_t)j * BD_CH + m]; }
      sd[u] = sqrt(dl[j]); xv[u] = xs[j];
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (j0 - u >= 0) {
        double dot = 0.0;
#pragma unroll
        for (int m = 0; m < NC; ++m) dot = fma(b[u][m], tau[m], dot);
#pragma unroll
        for (int m = 0; m < NC; ++m) tau[m] = fma(w[u][m], xv[u], tau[m]);
        if (lane == 0) xs[j0 - u] = fma(sd[u], xv[u], dot);
      }
    }
  }
  __syncwarp();
}

// Lm y = b, Lm' y = b, y = Lm' x for the product-form factor of `slot` (nch factors, ncol live columns in all); xs in M-space order
__device__ __forceinline__ bool bd_small(int ncol, int k) { return ncol - k * BD_CH <= 4; }
__device__ inline void bd_solve(const BdChain& Q, int slot, int nch, int ncol, double sa, double* xs) {
  bd_bidiag_solve(*Q.C, sa, xs, Q.lane);
  for (int k = 0; k < nch; ++k) {
    if (bd_small(ncol, k)) bd_k_solve_nc<4>(Q.Wc(slot, k), Q.Bc(slot, k), Q.Dc(slot, k), Q.d, xs, Q.lane);
    else bd_k_solve(Q.Wc(slot, k), Q.Bc(slot, k), Q.Dc(slot, k), Q.d, xs, Q.lane);
  }
}
__device__ inline void bd_solve_t(const BdChain& Q, int slot, int nch, int ncol, double sa, double* xs) {
  for (int k = nch - 1; k >= 0; --k) {
    if (bd_small(ncol, k)) bd_k_solve_t_nc<4>(Q.Wc(slot, k), Q.Bc(slot, k), Q.Dc(slot, k), Q.d, xs, Q.lane);
    else bd_k_solve_t(Q.Wc(slot, k), Q.Bc(slot, k), Q.Dc(slot, k), Q.d, xs, Q.lane);
  }
  bd_bidiag_solve_t(*Q.C, sa, xs, Q.lane);
}
__device__ inline void bd_mult_t(const BdChain& Q, int slot, int nch, int ncol, double sa, double* xs) {
  bd_bidiag_mult_t(*Q.C, sa, xs, Q.lane);
  for (int k = 0; k < nch; ++k) {
    if (bd_small(ncol, k)) bd_k_mult_t_nc<4>(Q.Wc(slot, k), Q.Bc(slot, k), Q.Dc(slot, k), Q.d, xs, Q.lane);
    else bd_k_mult_t(Q.Wc(slot, k), Q.Bc(slot, k), Q.Dc(slot, k), Q.d, xs, Q.lane);
  }
}
// NOTE on bd_mult_t: Lm' = K_m' ... K_1' (sqrt(a) L_T)'  -- the bidiagonal factor is applied first, then K_1', ..., K_m'.

// =====================================================================================================
// Level-2 evaluation at x = vec(xv) into evaluation slot `slot`: log-target, gradient, v, product-form factor of the softabs metric,
// G^-1 grad ("firstterm"), sum log diag.  Returns the log-target; *ok = 0 if anything is non-finite / not positive definite /
// (*ok = -1: needs more than ncmax product factors).  vm, xs: per-warp shared vectors (length d); bc: per-warp 2 x 32 broadcast rows.
// =====================================================================================================
__device__ inline double bd_eval2(const BdChain& Q, int xv, int slot, double* vm, double* xs, double* bc, int* ok) {
  const BdCfg& C = *Q.C;
  const int d = Q.d, lane = Q.lane;
  const double* x = Q.vec(xv);
  double* v = Q.vec(BDV_V0 + slot);
  double* g = Q.vec(BDV_G0 + slot);
  *ok = 1;
  const double q = bd_tmatvec(C, x, v, lane);
  const double lt = C.logconst - 0.5 * (C.nu + d) * log1p(q / C.nu);
  const double a = (C.nu + d) / (C.nu + q), beta = 2.0 / (C.nu + q);
  if (!isfinite(lt) || !(a > 0.0)) { *ok = 0; return lt; }
  __syncwarp();
  double vv = 0.0;
  for (int i = lane; i < d; i += 32) { const double vi = v[i]; g[i] = -a * vi; vm[d - 1 - i] = vi; vv = fma(vi, vi, vv); }
  vv = warp_sum(vv);
  __syncwarp();
  // ---- how many eigenvalues of the tensor a (T - beta v v') does softabs change?  (none without the transform)
  const double thr = 21.0 / C.alpha;
  int r = 0;
  if (C.transform == 1) r = bd_count(C, vm, beta, thr / a);
  const int ncol = r + 1;                                  // + the column v itself (the rank-one part of the tensor)
  const int nch = (ncol + BD_CH - 1) / BD_CH;
  if (nch > C.ncmax) { *ok = -1; return lt; }              // more product factors than allocated (gamc_bigd_init max_factors)
  const double sa = sqrt(a);
  double sumlog = 0.5 * d * log(a) + C.sumlog_ld;
  const double lo0 = C.gersh_lo - beta * vv - 1e-3 * fmax(1.0, fabs(C.gersh_lo));
  for (int k = 0; k < nch; ++k) {
    double* W = Q.Wc(slot, k); double* B = Q.Bc(slot, k); double* dl = Q.Dc(slot, k);
    const int col = k * BD_CH + lane;                      // this lane's column of V
    const bool is_root = col < r, is_v = col == r;
    double Sj = 0.0, scale = 1.0;
    // ---- roots of the secular equation: eigenvalue number col+1 of A = T' - beta v v' for the column of lane `lane`
    double mu = 0.0;
    const bool any_root = k * BD_CH < r;                   // the last factor may hold the column v only
    if (any_root) {
      // (g+1)-section on the count: the nr roots of this factor share the warp, g = 32 / nr lanes each.  Every step evaluates the
      // count at g interior points of the root's bracket at once (one lane each) and keeps the sub-interval in which the count
      // reaches the root's index: log2(g+1) bits per sweep instead of one -- 12 sweeps instead of 60 when softabs changes a single
      // eigenvalue (the regime a chain lives in once it has reached the typical set).
      const int nr = min(BD_CH, r - k * BD_CH);
      if (nr > 4) {
        // Many roots in this factor (far from the mode): lane <-> root j.  Root j lies strictly between the poles d_{j-1} and d_j of
        // the secular function (the eigenvalues of T: constants, computed once on the host), where f falls monotonically from +inf to
        // -inf; phi = f (mu - d_{j-1})(d_j - mu) / gap is finite and smooth there, so a bracketed secant iteration (Illinois variant of
        // regula falsi) on phi converges superlinearly: about ten sweeps instead of sixty bisection steps.
        const bool act = lane < nr;
        const int j = k * BD_CH + lane;
        const double U = C.poles[min(j, d - 1)];
        const double Lp = (j == 0) ? lo0 : C.poles[min(j, d - 1) - (j > 0 ? 1 : 0)];
        const double gap = U - Lp;
        auto phi = [&](double m2) {
          const double f = bd_secular(C, vm, beta, m2);
          return (j == 0) ? f * (U - m2) : f * ((m2 - Lp) * (U - m2) / gap);
        };
        double xa = (j == 0) ? Lp : fma(1e-10, gap, Lp), xb = fma(-1e-10, gap, U);
        double fa = phi(xa), fb = phi(xb);
        double x = 0.5 * (xa + xb);
        bool conv = !act;
        if (act && !(fa > 0.0)) { x = xa; conv = true; }            // the root sits within 1e-10 gap of the lower pole
        if (act && !conv && !(fb < 0.0)) { x = xb; conv = true; }   // ... of the upper pole
        const double tol = 4e-16 * fmax(fabs(Lp), fabs(U));
        int side = 0;
        for (int it = 0; it < 48; ++it) {
          if (__all_sync(0xffffffffu, conv)) break;
          double xn = (xa * fb - xb * fa) / (fb - fa);
          if (!(xn > xa && xn < xb)) xn = 0.5 * (xa + xb);
          if (conv) xn = x;
          const double fx = phi(xn);                                   // converged lanes run along (uniform control flow)
          if (!conv) {
            const double step = fabs(xn - x);
            x = xn;
            if (fx > 0.0) { xa = xn; fa = fx; if (side == 1) fb *= 0.5; side = 1; }
            else if (fx < 0.0) { xb = xn; fb = fx; if (side == -1) fa *= 0.5; side = -1; }
            else conv = true;
            if (step <= tol || (xb - xa) <= tol) conv = true;
          }
        }
        mu = x;
      } else {
      const int g = 32 / nr;
      const int grp = lane / g, sub = lane - grp * g;
      const bool act = grp < nr;
      const int want = k * BD_CH + grp + 1;
      const unsigned gmask = (g == 32 ? 0xffffffffu : ((1u << g) - 1u)) << (act ? grp * g : 0);
      double lo = lo0, hi = fmin(thr / a, C.gersh_hi);
      const int nit = (int)ceil((double)BD_NBISECT / log2((double)(g + 1)));
      for (int it = 0; it < nit; ++it) {
        const double hstep = (hi - lo) / (double)(g + 1);
        const double mid = fma((double)(sub + 1), hstep, lo);
        const int cnt = bd_count(C, vm, beta, mid);        // idle lanes run along (uniform control flow)
        const unsigned hit = __ballot_sync(0xffffffffu, act && cnt >= want);
        const unsigned bits = act ? (hit & gmask) >> (grp * g) : 0u;
        const int f = bits ? (__ffs(bits) - 1) : g;        // first interior point at which the count has reached `want`
        const double nlo = fma((double)f, hstep, lo);
        hi = (f < g) ? fma((double)(f + 1), hstep, lo) : hi;
        lo = nlo;
      }
      const double mu_grp = 0.5 * (lo + hi);
      mu = __shfl_sync(0xffffffffu, mu_grp, min(lane, nr - 1) * g);    // lane j <-> column j of this factor
      }
    }
    // ---- eigenvector x = (T' - mu)^-1 v' (unnormalised) by the LDL' recurrence: forward sweep parks y and the multipliers in the
    //      chunk's own W / B storage, backward sweep overwrites W with x
    if (!any_root) {
      for (int i = 0; i < d; ++i) W[(size_t)i * BD_CH + lane] = is_v ? vm[i] : 0.0;
      if (is_v) Sj = -a * beta;
      __syncwarp();
    } else {
      const double pivmin = 1e-290;
      double D = bd_tdm(C, 0) - mu;
      if (fabs(D) < pivmin) D = -pivmin;
      double rD = 1.0 / D, z = vm[0];
      if (is_root) { W[lane] = z * rD; B[lane] = 0.0; }
      for (int i = 1; i < d; ++i) {
        const double e = bd_tem(C, i - 1);
        const double l = e * rD;
        D = (bd_tdm(C, i) - mu) - l * e;
        if (fabs(D) < pivmin) D = -pivmin;
        rD = 1.0 / D;
        z = fma(-l, z, vm[i]);
        if (is_root) { W[(size_t)i * BD_CH + lane] = z * rD; B[(size_t)i * BD_CH + lane] = l; }
      }
      __syncwarp();
      double xn = 0.0, amax = 0.0, ssq = 0.0;
      for (int i0 = d - 1; i0 >= 0; i0 -= BD_PF) {
        double yv[BD_PF], lv[BD_PF];
#pragma unroll
        for (int u = 0; u < BD_PF; ++u) {
          const int i = max(i0 - u, 0);
          yv[u] = W[(size_t)i * BD_CH + lane];
          lv[u] = (i + 1 < d) ? B[(size_t)(i + 1) * BD_CH + lane] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < BD_PF; ++u) {
          const int i = i0 - u;
          if (i >= 0) {
            if (is_root) {
              xn = fma(-lv[u], xn, yv[u]);
              W[(size_t)i * BD_CH + lane] = xn;
              amax = fmax(amax, fabs(xn)); ssq = fma(xn, xn, ssq);
            } else {
              W[(size_t)i * BD_CH + lane] = is_v ? vm[i] : 0.0;
            }
          }
        }
      }
      if (is_root) {
        int ex; frexp(amax, &ex);
        scale = ldexp(1.0, -ex);                           // power of two: the column is used as x * scale (exact)
        Sj = bd_softabs_gap(a * mu, C.alpha) / (ssq * scale * scale);
        if (!isfinite(Sj) || !(ssq > 0.0)) Sj = nan("");
      } else if (is_v) {
        Sj = -a * beta;
      }
      __syncwarp();
    }
    if (lane < BD_CH) Q.Sc(slot, k)[lane] = Sj;
    // ---- W <- (sqrt(a) L_T)^-1 (x scale): bidiagonal recurrence, lane <-> column
    {
      const double rsa = scale / sa;
      double u = 0.0;
      for (int i0 = 0; i0 < d; i0 += BD_PF) {
        double wv[BD_PF];
#pragma unroll
        for (int q2 = 0; q2 < BD_PF; ++q2) wv[q2] = W[(size_t)min(i0 + q2, d - 1) * BD_CH + lane];
#pragma unroll
        for (int q2 = 0; q2 < BD_PF; ++q2) {
          const int i = i0 + q2;
          if (i < d) {
            u = (wv[q2] * rsa - (i > 0 ? C.le[i - 1] : 0.0) * u) * C.rld[i];
            W[(size_t)i * BD_CH + lane] = u;
          }
        }
      }
      __syncwarp();
    }
    // ---- W <- K_{k-1}^-1 ... K_0^-1 W: lane <-> column, every lane carries its own sigma (32 doubles); generator rows of the
    //      earlier factor are broadcast through shared memory
    for (int c = 0; c < k; ++c) {
      const double* Wp = Q.Wc(slot, c); const double* Bp = Q.Bc(slot, c); const double* dp = Q.Dc(slot, c);
      double sig[BD_CH];
#pragma unroll
      for (int m = 0; m < BD_CH; ++m) sig[m] = 0.0;
      for (int i0 = 0; i0 < d; i0 += 4) {
        double wp[4], bp[4], dv[4], wo[4];
#pragma unroll
        for (int q2 = 0; q2 < 4; ++q2) {
          const int i = min(i0 + q2, d - 1);
          wp[q2] = Wp[(size_t)i * BD_CH + lane]; bp[q2] = Bp[(size_t)i * BD_CH + lane]; dv[q2] = dp[i]; wo[q2] = W[(size_t)i * BD_CH + lane];
        }
#pragma unroll
        for (int q2 = 0; q2 < 4; ++q2) {
          const int i = i0 + q2;
          if (i < d) {
            bc[lane] = wp[q2]; bc[BD_CH + lane] = bp[q2];
            __syncwarp();
            double dot = 0.0;
#pragma unroll
            for (int m = 0; m < BD_CH; ++m) dot = fma(bc[m], sig[m], dot);
            const double y = (wo[q2] - dot) * rsqrt(dv[q2]);
#pragma unroll
            for (int m = 0; m < BD_CH; ++m) sig[m] = fma(bc[BD_CH + m], y, sig[m]);
            W[(size_t)i * BD_CH + lane] = y;
            __syncwarp();
          }
        }
      }
    }
    // ---- row sweep: chol(I + W S W').  With at most four live columns (the typical-set regime) every lane carries the whole 4 x 4
    //      state and reads the generator rows as broadcasts: no shared-memory broadcast, no shuffle reduction on the row's chain.
    const int live = min(BD_CH, ncol - k * BD_CH);
    if (live <= 4) {
      constexpr int NC = 4;
      double M[NC][NC];
#pragma unroll
      for (int p2 = 0; p2 < NC; ++p2)
#pragma unroll
        for (int q2 = 0; q2 < NC; ++q2) M[p2][q2] = 0.0;
#pragma unroll
      for (int p2 = 0; p2 < NC; ++p2) M[p2][p2] = __shfl_sync(0xffffffffu, Sj, p2);
      int bad = 0;
      double slog = 0.0, prod = 1.0;
      for (int i0 = 0; i0 < d && !bad; i0 += 4) {
        double wv[4][NC];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int m = 0; m < NC; ++m) wv[u][m] = W[(size_t)min(i0 + u, d - 1) * BD_CH + m];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u;
          if (i < d && !bad) {
            double t[NC];
            double dlt = 1.0;
#pragma unroll
            for (int p2 = 0; p2 < NC; ++p2) {
              double a2 = 0.0;
#pragma unroll
              for (int q2 = 0; q2 < NC; ++q2) a2 = fma(M[p2][q2], wv[u][q2], a2);
              t[p2] = a2;
              dlt = fma(wv[u][p2], a2, dlt);
            }
            if (!(dlt > 0.0) || !isfinite(dlt)) {
              bad = 1;
            } else {
              const double rs = rsqrt(dlt), rr = rs * rs;
              if (lane < NC) B[(size_t)i * BD_CH + lane] = (lane == 0 ? t[0] : lane == 1 ? t[1] : lane == 2 ? t[2] : t[3]) * rs;
              else B[(size_t)i * BD_CH + lane] = 0.0;
#pragma unroll
              for (int p2 = 0; p2 < NC; ++p2)
#pragma unroll
                for (int q2 = 0; q2 < NC; ++q2) M[p2][q2] = fma(-t[p2] * rr, t[q2], M[p2][q2]);
              if (lane == 0) dl[i] = dlt;
              prod *= dlt;
              if ((i & 7) == 7) { slog += log(prod); prod = 1.0; }      // one logarithm per eight rows
            }
          }
        }
      }
      if (bad) { *ok = 0; return lt; }
      sumlog += 0.5 * (slog + log(prod));
    } else {
      double M[BD_CH];
#pragma unroll
      for (int m = 0; m < BD_CH; ++m) M[m] = (m == lane) ? Sj : 0.0;
      int bad = 0;
      double slog = 0.0;
      double wnext[BD_PF];
#pragma unroll
      for (int q2 = 0; q2 < BD_PF; ++q2) wnext[q2] = W[(size_t)min(q2, d - 1) * BD_CH + lane];
      for (int i0 = 0; i0 < d && !bad; i0 += BD_PF) {
        double wcur[BD_PF];
#pragma unroll
        for (int q2 = 0; q2 < BD_PF; ++q2) wcur[q2] = wnext[q2];
        if (i0 + BD_PF < d) {             // next block's rows in flight while this block is swept
#pragma unroll
          for (int q2 = 0; q2 < BD_PF; ++q2) wnext[q2] = W[(size_t)min(i0 + BD_PF + q2, d - 1) * BD_CH + lane];
        }
#pragma unroll
        for (int q2 = 0; q2 < BD_PF; ++q2) {
          const int i = i0 + q2;
          if (i < d && !bad) {
            const double w = wcur[q2];
            bc[lane] = w;
            __syncwarp();
            double t = 0.0;
#pragma unroll
            for (int m = 0; m < BD_CH; ++m) t = fma(M[m], bc[m], t);
            const double dlt = 1.0 + warp_sum(w * t);
            if (!(dlt > 0.0) || !isfinite(dlt)) {          // uniform: dlt is the same on every lane
              bad = 1;
            } else {
              const double rs = rsqrt(dlt);
              B[(size_t)i * BD_CH + lane] = t * rs;
              bc[BD_CH + lane] = t;
              __syncwarp();
              const double tr = t * (rs * rs);
#pragma unroll
              for (int m = 0; m < BD_CH; ++m) M[m] = fma(-tr, bc[BD_CH + m], M[m]);
              if (lane == 0) dl[i] = dlt;
              slog += log(dlt);
            }
            __syncwarp();
          }
        }
      }
      if (bad) { *ok = 0; return lt; }
      sumlog += 0.5 * slog;
    }
    __syncwarp();
  }
  // ---- firstterm = G^-1 grad = P Lm^-T Lm^-1 P grad
  for (int i = lane; i < d; i += 32) xs[d - 1 - i] = g[i];
  __syncwarp();
  bd_solve(Q, slot, nch, ncol, sa, xs);
  bd_solve_t(Q, slot, nch, ncol, sa, xs);
  double* f = Q.vec(BDV_F0 + slot);
  double bad = 0.0;
  for (int i = lane; i < d; i += 32) { const double fi = xs[d - 1 - i]; f[i] = fi; if (!isfinite(fi)) bad = 1.0; }
  if (warp_sum(bad) != 0.0 || !isfinite(sumlog)) { *ok = 0; return lt; }
  if (lane == 0) { Q.s[BDS_A0 + slot] = a; Q.s[BDS_BETA0 + slot] = beta; Q.s[BDS_SL0 + slot] = sumlog; Q.s[BDS_NCH0 + slot] = (double)nch; Q.s[BDS_R0 + slot] = (double)r; }
  __syncwarp();
  return lt;
}

// running means, chain output, burn-in tuning (iterate/GAMC.jl:193-292): warp version of step_tail
__device__ inline void bd_tail(const BdChain& Q, int accepted) {
  const BdCfg& C = *Q.C;
  const int d = Q.d, lane = Q.lane;
  double* s = Q.s;
  const long long t = (long long)s[BDS_COUNT];
  const double* th = Q.vec(BDV_THETA);
  double* m1 = Q.vec(BDV_M1); double* m2 = Q.vec(BDV_M2);
  for (int i = lane; i < d; i += 32) { const double lm = m1[i]; m2[i] = lm; m1[i] = ((double)(t - 1) * lm + th[i]) / (double)t; }
  if (t > C.burnin) {
    const long long pos = (long long)s[BDS_SAVED];
    if (pos < C.chain_cap) {
      double* out = C.chain_value + ((size_t)Q.chain * C.chain_cap + pos) * d;
      for (int i = lane; i < d; i += 32) out[i] = th[i];
      if (lane == 0) C.chain_accept[(size_t)Q.chain * C.chain_cap + pos] = (unsigned char)accepted;
    }
  }
  __syncwarp();
  if (lane == 0) {
    if (t > C.burnin && s[BDS_SAVED] < (double)C.chain_cap) s[BDS_SAVED] += 1.0;
    if (C.tuning) {
      if ((long long)s[BDS_TTOT] <= C.burnin && ((long long)s[BDS_TPROP] % C.period) == 0) {
        if (C.count_total) s[BDS_TRATE] = s[BDS_TACC] / s[BDS_TPROP];
        if (C.is_accrate) { s[BDS_STEP] *= 2.0 / (1.0 + exp(-C.score_k * (s[BDS_TRATE] - C.targetrate))); s[BDS_SQSTEP] = sqrt(s[BDS_STEP]); }
        if (C.verbose_sm) { s[BDS_SMTOT] += s[BDS_SMPROP]; s[BDS_SMPROP] = 0; s[BDS_SMACC] = 0; }
        if (C.verbose_am) { s[BDS_AMTOT] += s[BDS_AMPROP]; s[BDS_AMPROP] = 0; s[BDS_AMACC] = 0; }
        s[BDS_TTOT] += s[BDS_TPROP]; s[BDS_TPROP] = 0;
        if (C.count_total) { s[BDS_TACC] = 0; s[BDS_TRATE] = nan(""); }
      }
    }
    s[BDS_PAST] = s[BDS_PRESENT];                                                   // :197
  }
  __syncwarp();
}

// =====================================================================================================
// kernels.  Warp-per-chain kernels: blockDim = 32 * BD_WPB, dynamic shared memory = BD_WPB * (2 d + 64) doubles.
// =====================================================================================================
__host__ __device__ inline size_t bd_warp_smem_bytes(int d) { return sizeof(double) * (size_t)BD_WPB * (2 * (size_t)d + 2 * BD_CH); }

struct BdWarp { BdChain Q; double* vm; double* xs; double* bc; bool active; };
__device__ inline BdWarp bd_warp_setup(const BdCfg& C, double* smem) {
  BdWarp w;
  const int warp = threadIdx.x >> 5;
  w.Q.C = &C; w.Q.lane = threadIdx.x & 31; w.Q.d = C.d;
  w.Q.chain = blockIdx.x * BD_WPB + warp;
  w.active = w.Q.chain < C.nchains;
  w.Q.s = C.scal + (size_t)(w.active ? w.Q.chain : 0) * BDS_N;
  double* base = smem + (size_t)warp * (2 * (size_t)C.d + 2 * BD_CH);
  w.vm = base; w.xs = base + C.d; w.bc = base + 2 * (size_t)C.d;
  return w;
}

// initialize! + sampler_state (src/samplers/GAMC.jl:157-228) for every chain; theta0 [chain][d]
__global__ void __launch_bounds__(32 * BD_WPB) k_bd_init(BdCfg C, const double* __restrict__ theta0) {
  extern __shared__ __align__(16) double bd_smem[];
  BdWarp w = bd_warp_setup(C, bd_smem);
  if (!w.active) return;
  const BdChain& Q = w.Q;
  const int d = C.d, lane = Q.lane;
  double* s = Q.s;
  for (int i = lane; i < BDS_N; i += 32) s[i] = 0.0;
  double* th = Q.vec(BDV_THETA); double* m1 = Q.vec(BDV_M1);
  for (int i = lane; i < d; i += 32) { const double v = theta0[(size_t)Q.chain * d + i]; th[i] = v; m1[i] = v; Q.vec(BDV_M2)[i] = 0.0; }
  __syncwarp();
  int ok;
  const double lt = bd_eval2(Q, BDV_THETA, 0, w.vm, w.xs, w.bc, &ok);
  if (lane == 0) {
    s[BDS_LT] = lt; s[BDS_CUR] = 0.0; s[BDS_STEP] = C.driftstep; s[BDS_SQSTEP] = sqrt(C.driftstep);
    s[BDS_TTOT] = C.period; s[BDS_SMTOT] = C.period; s[BDS_AMTOT] = C.period; s[BDS_TRATE] = nan("");
    s[BDS_PRESENT] = 1.0; s[BDS_PAST] = 0.0;
    if (ok != 1) s[BDS_ERR] = ok < 0 ? 6.0 : 2.0;           // more factors needed than allocated / GAMC_NONFINITE_INIT (or not PD) at the initial point
  }
}

// head of iterate! (:2-10): counters, the iteration's uniforms, the schedule coin.  One thread per chain.
__global__ void __launch_bounds__(128) k_bd_begin(BdCfg C, long long tape_pos) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= C.nchains) return;
  double* s = C.scal + (size_t)chain * BDS_N;
  if (s[BDS_ERR] != 0.0) return;
  const long long t = (long long)s[BDS_COUNT] + 1;
  double us, um, ua;
  bd_uniforms(C, chain, t, tape_pos, us, um, ua);
  bool present = s[BDS_PRESENT] != 0.0;
  if (t > C.t0) present = bd_schedule(C, t, us);
  s[BDS_COUNT] = (double)t;
  if (C.tuning) s[BDS_TPROP] += 1.0;
  s[BDS_PRESENT] = present ? 1.0 : 0.0;
  s[BDS_US] = us; s[BDS_UM] = um; s[BDS_UA] = ua;
  if (present) { s[BDS_UPD] += 1.0; if (C.verbose_sm) s[BDS_SMPROP] += 1.0; }
  else if (C.verbose_am) s[BDS_AMPROP] += 1.0;
}

// SMMALA branch, first half (:19-35): re-evaluation at theta if the previous step was AM, then the proposal
__global__ void __launch_bounds__(32 * BD_WPB) k_bd_smmala_pre(BdCfg C, long long tape_pos) {
  extern __shared__ __align__(16) double bd_smem[];
  BdWarp w = bd_warp_setup(C, bd_smem);
  if (!w.active) return;
  const BdChain& Q = w.Q;
  double* s = Q.s;
  if (s[BDS_ERR] != 0.0 || s[BDS_PRESENT] == 0.0) return;
  const int d = C.d, lane = Q.lane;
  const int cur = (int)s[BDS_CUR];
  if (s[BDS_PAST] == 0.0) {                                                         // :19-31
    int ok;
    const double lt = bd_eval2(Q, BDV_THETA, cur, w.vm, w.xs, w.bc, &ok);
    if (ok != 1) { if (lane == 0) s[BDS_ERR] = ok < 0 ? 6.0 : 3.0; return; }        // capacity / PosDefException analogue
    if (lane == 0) s[BDS_LT] = lt;
    __syncwarp();
  }
  const double eps = s[BDS_STEP], sq = s[BDS_SQSTEP];
  const long long t = (long long)s[BDS_COUNT];
  // theta* = mu + sqrt(eps) chol(inv(G)) z,  chol(inv(G)) z = P Lm^-T P z                                   (:33-35)
  for (int i = lane; i < d; i += 32) w.xs[d - 1 - i] = bd_normal(C, Q.chain, t, tape_pos, i);
  __syncwarp();
  bd_solve_t(Q, cur, (int)s[BDS_NCH0 + cur], (int)s[BDS_R0 + cur] + 1, sqrt(s[BDS_A0 + cur]), w.xs);
  const double* th = Q.vec(BDV_THETA); const double* f = Q.vec(BDV_F0 + cur);
  double* mu = Q.vec(BDV_MU); double* prop = Q.vec(BDV_PROP);
  for (int i = lane; i < d; i += 32) { const double m = th[i] + 0.5 * eps * f[i]; mu[i] = m; prop[i] = m + sq * w.xs[d - 1 - i]; }
}

// SMMALA branch, second half (:37-127): evaluation at the proposal, ratio, accept/reject, tail
__global__ void __launch_bounds__(32 * BD_WPB) k_bd_smmala_post(BdCfg C) {
  extern __shared__ __align__(16) double bd_smem[];
  BdWarp w = bd_warp_setup(C, bd_smem);
  if (!w.active) return;
  const BdChain& Q = w.Q;
  double* s = Q.s;
  if (s[BDS_ERR] != 0.0 || s[BDS_PRESENT] == 0.0) return;
  const int d = C.d, lane = Q.lane;
  const int cur = (int)s[BDS_CUR], prp = 1 - cur;
  const double eps = s[BDS_STEP];
  int ok;
  const double lt_p = bd_eval2(Q, BDV_PROP, prp, w.vm, w.xs, w.bc, &ok);             // :37-43,56
  if (ok < 0) { if (lane == 0) s[BDS_ERR] = 6.0; return; }                          // more product factors needed than allocated
  double ratio = -INFINITY;
  double* th = Q.vec(BDV_THETA); const double* prop = Q.vec(BDV_PROP); const double* mu = Q.vec(BDV_MU);
  if (ok == 1) {
    // q1 = d'G d = |Lm' P d|^2, d = theta* - mu                                                              (:47-54)
    for (int i = lane; i < d; i += 32) w.xs[d - 1 - i] = prop[i] - mu[i];
    __syncwarp();
    bd_mult_t(Q, cur, (int)s[BDS_NCH0 + cur], (int)s[BDS_R0 + cur] + 1, sqrt(s[BDS_A0 + cur]), w.xs);
    double q1 = 0.0;
    for (int i = lane; i < d; i += 32) q1 = fma(w.xs[i], w.xs[i], q1);
    q1 = warp_sum(q1);
    __syncwarp();
    // q2 = (theta - mu*)'G*(theta - mu*), mu* = theta* + eps/2 G*^-1 grad*                                   (:58-67)
    const double* fp = Q.vec(BDV_F0 + prp);
    for (int i = lane; i < d; i += 32) w.xs[d - 1 - i] = th[i] - (prop[i] + 0.5 * eps * fp[i]);
    __syncwarp();
    bd_mult_t(Q, prp, (int)s[BDS_NCH0 + prp], (int)s[BDS_R0 + prp] + 1, sqrt(s[BDS_A0 + prp]), w.xs);
    double q2 = 0.0;
    for (int i = lane; i < d; i += 32) q2 = fma(w.xs[i], w.xs[i], q2);
    q2 = warp_sum(q2);
    const double plog = d * log(eps);
    const double ld_old = plog - 2.0 * s[BDS_SL0 + cur], ld_new = plog - 2.0 * s[BDS_SL0 + prp];   // logdet(eps inv(G))
    ratio = lt_p - s[BDS_LT];
    ratio += 0.5 * (ld_old + q1 / eps);
    ratio -= 0.5 * (ld_new + q2 / eps);
  }
  const int accept = (ratio > 0.0) || (ratio > log(s[BDS_UA]));                     // :69
  __syncwarp();
  if (accept) for (int i = lane; i < d; i += 32) th[i] = prop[i];
  __syncwarp();
  if (lane == 0) {
    s[BDS_RATIO] = ratio;
    if (accept) { s[BDS_LT] = lt_p; s[BDS_CUR] = (double)prp; if (C.count_total) s[BDS_TACC] += 1.0; if (C.verbose_sm) s[BDS_SMACC] += 1.0; }
  }
  __syncwarp();
  bd_tail(Q, accept);
}

// AM branch, second half (:148-191 + tail): log-target at the proposal, accept/reject.  The two GMM log-densities cancel exactly.
__global__ void __launch_bounds__(32 * BD_WPB) k_bd_am_post(BdCfg C) {
  extern __shared__ __align__(16) double bd_smem[];
  BdWarp w = bd_warp_setup(C, bd_smem);
  if (!w.active) return;
  const BdChain& Q = w.Q;
  double* s = Q.s;
  if (s[BDS_ERR] != 0.0 || s[BDS_PRESENT] != 0.0) return;
  const int d = C.d, lane = Q.lane;
  double* th = Q.vec(BDV_THETA); const double* prop = Q.vec(BDV_PROP);
  const double q = bd_tmatvec(C, prop, w.xs, lane);
  const double lt_p = C.logconst - 0.5 * (C.nu + d) * log1p(q / C.nu);
  const double ratio = isfinite(lt_p) ? (lt_p - s[BDS_LT]) : -INFINITY;             // :150-156
  const int accept = (ratio > 0.0) || (ratio > log(s[BDS_UA]));                     // :158
  __syncwarp();
  if (accept) for (int i = lane; i < d; i += 32) th[i] = prop[i];
  __syncwarp();
  if (lane == 0) {
    s[BDS_RATIO] = ratio;
    if (accept) { s[BDS_LT] = lt_p; if (C.count_total) s[BDS_TACC] += 1.0; if (C.verbose_am) s[BDS_AMACC] += 1.0; }
  }
  __syncwarp();
  bd_tail(Q, accept);
}

// ---------------------------------------------------------------------------------------------------
// SMMALA -> AM hand-over (:26,92,133): sstate.oldinvtensor = inv(G) seeds the AM covariance; its lower Cholesky factor
// chol(inv(G)) = U^-T is the index reversal of Lm^-1.  One CTA per chain, thread c = column c of X = Lm^-1 (forward solves on the
// unit vectors, all columns in lockstep over the rows so that every store is coalesced), one pass per product factor.
// X(j, c) lands at L[(d-1-c) + (d-1-j) d].
// ---------------------------------------------------------------------------------------------------
// One pass of the hand-over for one product factor with NC live columns (NC = 4 covers the typical-set regime, where a factor
// holds the lowest eigenvector and v: the other 30 generator columns are zero and are skipped).
template <int NC>
__device__ __forceinline__ void bd_handover_pass(const BdCfg& C, const double* __restrict__ W, const double* __restrict__ B, const double* __restrict__ dl,
                                                 double* __restrict__ L, double* __restrict__ rdiag, int k, bool last, double rsa, double (*rows)[2 * BD_CH + 2]) {
  const int c = threadIdx.x, d = C.d;
  double sig[NC];
#pragma unroll
  for (int m = 0; m < NC; ++m) sig[m] = 0.0;
  double u = 0.0;                                      // bidiagonal recurrence of pass 0
  for (int i0 = 0; i0 < d; i0 += 16) {
    __syncthreads();
    for (int e = threadIdx.x; e < 16 * (2 * NC + 1); e += blockDim.x) {
      const int rr = e / (2 * NC + 1), m = e % (2 * NC + 1), i = i0 + rr;
      if (i < d) rows[rr][m] = m < NC ? W[(size_t)i * BD_CH + m] : (m < 2 * NC ? B[(size_t)i * BD_CH + (m - NC)] : rsqrt(dl[i]));
    }
    __syncthreads();
    if (c >= d) continue;
    for (int rr = 0; rr < 16 && i0 + rr < d; ++rr) {
      const int i = i0 + rr;
      if (i < c) continue;                             // X is lower triangular: rows above the diagonal are zero and stay zero
      const size_t addr = (size_t)(d - 1 - c) + (size_t)(d - 1 - i) * d;
      double x;
      if (k == 0) { u = ((i == c ? rsa : 0.0) - C.le[i > 0 ? i - 1 : 0] * (i > c ? u : 0.0)) * C.rld[i]; x = u; }
      else x = L[addr];
      double dot = 0.0;
#pragma unroll
      for (int m = 0; m < NC; ++m) dot = fma(rows[rr][m], sig[m], dot);
      const double y = (x - dot) * rows[rr][2 * NC];
#pragma unroll
      for (int m = 0; m < NC; ++m) sig[m] = fma(rows[rr][NC + m], y, sig[m]);
      L[addr] = y;
      if (last && i == c) rdiag[d - 1 - c] = 1.0 / y;
    }
  }
}

__global__ void __launch_bounds__(BD_MAXD) k_bd_handover(BdCfg C) {
  __shared__ double rows[16][2 * BD_CH + 2];
  const int chain = blockIdx.x, d = C.d;
  double* s = C.scal + (size_t)chain * BDS_N;
  if (s[BDS_ERR] != 0.0 || s[BDS_PRESENT] != 0.0 || s[BDS_PAST] == 0.0) return;
  const int cur = (int)s[BDS_CUR];
  const int nch = (int)s[BDS_NCH0 + cur];
  const int ncol = (int)s[BDS_R0 + cur] + 1;           // live columns over all factors (the eigenvectors softabs changed, then v)
  const double rsa = 1.0 / sqrt(s[BDS_A0 + cur]);
  double* L = C.L + (size_t)chain * d * d;
  double* rdiag = C.rdiagL + (size_t)chain * d;
  const size_t sc0 = ((size_t)chain * 2 + cur) * C.ncmax;
  for (int k = 0; k < nch; ++k) {
    const double* W = C.W + (sc0 + k) * (size_t)d * BD_CH;
    const double* B = C.B + (sc0 + k) * (size_t)d * BD_CH;
    const double* dl = C.delta + (sc0 + k) * (size_t)d;
    const int live = min(BD_CH, ncol - k * BD_CH);
    if (live <= 4) bd_handover_pass<4>(C, W, B, dl, L, rdiag, k, k == nch - 1, rsa, rows);
    else bd_handover_pass<BD_CH>(C, W, B, dl, L, rdiag, k, k == nch - 1, rsa, rows);
  }
}

// ---------------------------------------------------------------------------------------------------
// AM branch, first half (:128-146): rank-one update of the dense factor + proposal.  The same closed form as the single-chain
// path (sampler.cuh k_am_pre_r1: C_k = ((k-1)/k) C_{k-1} + u u', chol(I + w w') in closed form, one ascending sweep over
// 32-column blocks), one CTA per chain, thread = row, the factor streamed from / to HBM in place.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BD_MAXD) k_bd_am_pre(BdCfg C, long long tape_pos) {
  __shared__ double zs[BD_MAXD];
  __shared__ double bw[4 * 32];
  __shared__ double Ls[32 * 33];
  __shared__ double Ts[32 * 33];
  __shared__ double accd[32];
  const int chain = blockIdx.x, d = C.d;
  double* s = C.scal + (size_t)chain * BDS_N;
  if (s[BDS_ERR] != 0.0 || s[BDS_PRESENT] != 0.0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  double* L = C.L + (size_t)chain * d * d;
  double* rd = C.rdiagL + (size_t)chain * d;
  const double* theta = C.vecs + ((size_t)chain * BDV_N + BDV_THETA) * d;
  const double* m2 = C.vecs + ((size_t)chain * BDV_N + BDV_M2) * d;
  double* prop = C.vecs + ((size_t)chain * BDV_N + BDV_PROP) * d;
  const long long t = (long long)s[BDS_COUNT];
  const double k = (double)(t - 2);
  if (!(k > 1.0)) { __syncthreads(); if (tid == 0) s[BDS_ERR] = 3.0; return; }        // C_1 has rank one: the reference's chol throws
  const double alpha = sqrt((k - 1.0) / k), beta = 1.0 / sqrt(k + 1.0), ralpha = 1.0 / alpha;
  const int i = tid;
  double th = 0.0, ui = 0.0, P = 0.0, acc = 0.0, rdv = 0.0;
  if (i < d) {
    th = theta[i];
    ui = (th - m2[i]) * beta;
    rdv = rd[i] * ralpha;
    zs[i] = bd_normal(C, chain, t, tape_pos, i);
  }
  double sigma = 1.0;
  const int nblk = (d + 31) >> 5;
  // the first diagonal block, towards L2
  if (warp == 0 && i < d) for (int jj = 0; jj <= lane && jj < 32; ++jj) prefetch_l2(&L[(size_t)lane + (size_t)jj * d]);
  for (int b = 0; b < nblk; ++b) {
    const int j0 = b << 5, nb = min(32, d - j0);
    // While the diagonal warp runs its serial chain, the rows below fetch their 32 columns of this block (and the next diagonal
    // block) into L2: the loads after the barrier then see L2 latency instead of HBM latency.  One request per 32-byte sector.
    if (i >= j0 + 32 && i < d) {
      if ((lane & 3) == 0) {
#pragma unroll 8
        for (int jj = 0; jj < 32; ++jj) if (jj < nb) prefetch_l2(&L[(size_t)i + (size_t)(j0 + jj) * d]);
      }
      if (warp == b + 1) for (int jj = 0; jj <= lane && jj < 32 && j0 + 32 + jj < d; ++jj) prefetch_l2(&L[(size_t)i + (size_t)(j0 + 32 + jj) * d]);
    }
    __syncthreads();
    if (warp == b) {
      // diagonal block: the lane's row of a L goes to shared memory (the register budget at 1024 threads per CTA is 64)
#pragma unroll 8
      for (int jj = 0; jj < 32; ++jj)
        Ls[jj * 33 + lane] = (jj <= lane && jj < nb && i < d) ? alpha * L[(size_t)(j0 + lane) + (size_t)(j0 + jj) * d] : 0.0;
      __syncwarp();
      double res = ui - P, wmine = 0.0;
#pragma unroll 8
      for (int jj = 0; jj < 32; ++jj) {
        if (jj < nb) {
          const double wj = __shfl_sync(0xffffffffu, res * rdv, jj);
          if (lane == jj) wmine = wj;
          res = fma(-Ls[jj * 33 + lane], wj, res);
          Ts[jj * 33 + lane] = res;
        }
      }
      P = ui - res;
      double sc = wmine * wmine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const double n = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += n; }
      const double incl = sigma + sc, excl = incl - wmine * wmine;
      const double rs = rsqrt(excl * incl);
      bw[lane] = wmine; bw[32 + lane] = incl * rs; bw[64 + lane] = wmine * rs;
      const double snew = __shfl_sync(0xffffffffu, incl, nb - 1);
      if (lane == 0) bw[96] = snew;
    }
    __syncthreads();
    sigma = bw[96];
    if (i >= j0 + 32 && i < d) {
      double res = ui - P, a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int h0 = 0; h0 < 32; h0 += 16) {              // 16 columns at a time: their loads are in flight together
        double lv[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) lv[jj] = (h0 + jj < nb) ? L[(size_t)i + (size_t)(j0 + h0 + jj) * d] : 0.0;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          if (h0 + jj < nb) {
            const int c = h0 + jj;
            const double l = alpha * lv[jj];
            res = fma(-l, bw[c], res);
            const double ln = fma(l, bw[32 + c], bw[64 + c] * res);
            L[(size_t)i + (size_t)(j0 + c) * d] = ln;
            if (jj & 1) a1 = fma(ln, zs[j0 + c], a1); else a0 = fma(ln, zs[j0 + c], a0);
          }
        }
      }
      P = ui - res;
      acc += a0 + a1;
    }
    for (int r = warp; r < nb; r += nwarps) {
      const int gi = j0 + r;
      double contrib = 0.0;
      if (lane <= r && lane < nb) {
        const double ln = fma(Ls[lane * 33 + r], bw[32 + lane], bw[64 + lane] * Ts[lane * 33 + r]);
        L[(size_t)gi + (size_t)(j0 + lane) * d] = ln;
        contrib = ln * zs[j0 + lane];
        if (lane == r) rd[gi] = 1.0 / ln;
      }
      contrib = warp_sum(contrib);
      if (lane == 0) accd[r] = contrib;
    }
    __syncthreads();
    if (warp == b && i < d) acc += accd[lane];
  }
  const int fin = isfinite(sigma) ? 0 : 1;
  if (__syncthreads_or(fin)) { if (tid == 0) s[BDS_ERR] = 3.0; return; }
  if (i < d) {
    if (s[BDS_UM] <= 1.0 - C.c) prop[i] = th + s[BDS_SQSTEP] * acc;                 // core component  N(theta, eps C)   (:146)
    else prop[i] = th + C.sqrtminorscale * zs[i];                                   // stabilising component
  }
}

}  // namespace gamc
