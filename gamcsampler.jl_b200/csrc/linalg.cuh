// linalg.cuh -- single-CTA dense FP64 linear algebra on p x p matrices resident in global memory (L2).
//
// Replaces the Julia Base / LAPACK calls of the reference transition
// (src/samplers/iterate/GAMC.jl:26-30,43,49-51,56,62-64,94; src/samplers/GAMC.jl:222-224):
//   inv(G), chol(Hermitian(inv(G)))', inv(G)*grad, logdet(eps*inv(G))
// by ONE factorisation per evaluated point.  The lower factor L of G^-1 (L L' = G^-1) that the reference
// obtains from inv + chol equals U^-T where G = U U' with U UPPER triangular (a "reverse" Cholesky =
// ordinary Cholesky of the index-reversed matrix); then
//   L z = U^-T z,  G^-1 grad = U^-T (U^-1 grad),  logdet(eps G^-1) = p log eps - 2 sum_i log U_ii.
// All routines work on a view with an optional index reversal (REV): in "M space" (i' = p-1-i) the
// reverse factor is an ordinary lower Cholesky factor Lm, so one set of lower-triangular routines serves
// both the SMMALA metric (REV = true) and the AM covariance (REV = false).
//
// Launch shape: one CTA of LA_THREADS threads.  Vectors live in shared memory in M-space order.
#pragma once
#include "common.cuh"

namespace gamc {

constexpr int LA_THREADS = 512;   // 128 registers per thread: the 32-wide row fragments stay in registers
constexpr int LA_NB = 32;
constexpr int LA_DLD = 33;  // padded leading dimension of the 32 x 32 diagonal block in shared memory
constexpr int LA_YLD = 36;  // leading dimension of the inverted diagonal block in shared memory (16-byte aligned rows)
constexpr int LA_INVD_STRIDE = 2 * LA_NB * LA_NB;   // doubles per diagonal block in the global inverse buffer (two layouts)

__host__ __device__ inline size_t la_invd_doubles(int p) { return (size_t)((p + LA_NB - 1) / LA_NB) * LA_INVD_STRIDE; }

template <bool REV>
struct MatView {
  double* a;
  int ld, p;
  __device__ __forceinline__ double& operator()(int i, int j) const {
    return REV ? a[(size_t)(p - 1 - i) + (size_t)(p - 1 - j) * ld] : a[(size_t)i + (size_t)j * ld];
  }
};

// Shared-memory scratch of the factorisation and the triangular solves.
struct LaScratch {
  double* D;      // LA_NB * LA_DLD      diagonal block before factorisation, row r at D[r * LA_DLD]
  double* Dn;     // 2 * LA_NB * LA_DLD  staged copies of the next diagonal block (written by the trailing update)
  double* LT;     // LA_NB * LA_DLD      its factor by columns, LT[j * LA_DLD + r] = Lkk(r, j)
  double* Y;      // LA_NB * LA_YLD      its inverse by rows, Y[c * LA_YLD + k] = (Lkk^-1)(c, k)
  double* rd;     // LA_NB               reciprocal diagonal of the block
  double* xs;     // LA_NB               block solution of the triangular solves
  double* P;      // LA_NB * pld         solved panel, k-major: P[c * pld + t]
  int* flag;      // [0] failed pivot, [1] factored columns published by the chain warp, [2] inverse rows published (monotone)
  int pld;
};

__host__ __device__ inline size_t la_scratch_bytes(int p) {
  const int pld = ((p + 3) & ~3) + 4;
  return sizeof(double) * (4 * LA_NB * LA_DLD + LA_NB * LA_YLD + 2 * LA_NB + (size_t)LA_NB * pld) + 16;
}

__device__ inline LaScratch la_scratch_carve(unsigned char* base, int p) {
  LaScratch s;
  s.pld = ((p + 3) & ~3) + 4;
  s.Y = reinterpret_cast<double*>(base);          // first: keeps the 16-byte alignment the double2 reads need
  s.P = s.Y + LA_NB * LA_YLD;
  s.D = s.P + (size_t)LA_NB * s.pld;
  s.Dn = s.D + LA_NB * LA_DLD;
  s.LT = s.Dn + 2 * LA_NB * LA_DLD;
  s.rd = s.LT + LA_NB * LA_DLD;
  s.xs = s.rd + LA_NB;
  s.flag = reinterpret_cast<int*>(s.xs + LA_NB);
  return s;
}

__device__ __forceinline__ void la_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void la_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void la_spin_until(volatile int* flag, int target) {
  while (*flag < target) __nanosleep(40);
  asm volatile("" ::: "memory");
}

// Warp roles of chol_blocked (16 warps).  Scheduler partition 0 (warps 0, 4, 8, 12) is kept free of trailing-update
// work so that the serial column chain of warp 0 is not slowed down by FP64 pipe contention.
constexpr int LA_TEAM = 128;                                   // warps 0, 4, 8, 12: diagonal-block update; warp 0: chain; warp 4: inverter
constexpr int LA_NTRAIL = LA_THREADS - LA_TEAM;                // 12 warps: trailing update + panel solve
__device__ __forceinline__ bool la_is_team(int warp) { return (warp & 3) == 0; }
__device__ __forceinline__ int la_trail_index(int warp, int lane) { return (warp - 1 - (warp >> 2)) * 32 + lane; }   // dense 0..383

// In-place blocked right-looking Cholesky of the lower triangle (M space) of A: A = Lm Lm', one CTA, software-pipelined
// at column granularity.  In pass `it`:
//   team (warps 0,4,8,12)  diagonal block `it` -= Pn Pn' (Pn = first 32 rows of the solved panel it-1)
//   warp 0                 factors it column by column (row per lane, in registers; dependent chain per column:
//                          shfl -> rsqrt -> mul -> sts/lds -> fma, no divide, no sqrt) and publishes each column
//   warp 4                 inverts the factor one column behind (lane j solves Lkk y = e_j) and publishes each row
//   the other 12 warps     rank-32 trailing update with panel it-1 on the FP64 tensor cores (DMMA m8n8k4), then the
//                          panel solve X = A_panel Lkk^-T, one row per thread, one column behind the inverter
// so the serial column chain is the only thing on the critical path of a pass.
// Also writes rdiag[i] = 1 / Lm_ii (M-space order, global), the inverted diagonal blocks (invd, may be null; both
// layouts [j*32+i] and [i*32+j], for the coalesced block solves of trsv_lower / trsv_lower_t) and returns
// sum_i log Lm_ii.  On a non-positive pivot returns NaN and sets *fail_pivot = 1-based M-space pivot (else 0).
// Must be called by all LA_THREADS threads.
template <bool REV>
__device__ double chol_blocked(MatView<REV> A, double* rdiag, double* invd, LaScratch S, int* fail_pivot) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = A.p;
  const int nblk = (p + LA_NB - 1) / LA_NB;
  volatile int* vflag = S.flag;
  if (tid == 0) { S.flag[0] = 0; S.flag[1] = 0; S.flag[2] = 0; }
  __syncthreads();
  for (int it = 0; it < nblk; ++it) {
    const int kn = it * LA_NB;                       // first row / column of diagonal block `it`
    const int nbn = min(LA_NB, p - kn);
    const int mprev = p - kn;                        // rows of panel it-1 (= trailing size), when it > 0
    const int mnext = p - kn - nbn;                  // rows of panel it
    if (la_is_team(warp)) {
      // ---- team: diagonal block `it` = A(block) - Pn Pn', 8 columns per warp.  From pass 2 on the block comes from
      //      the shared staging copy the trailing update of the previous pass left (no L2 round trip).
      const int c0 = (warp >> 2) * 8;
      GAMC_WCLK_BEGIN(w0);
      double av[8];
      if (it >= 2) {
        const double* Ds = S.Dn + (it & 1) * LA_NB * LA_DLD;
#pragma unroll
        for (int b = 0; b < 8; ++b) av[b] = Ds[lane * LA_DLD + c0 + b];
      } else {
#pragma unroll
        for (int b = 0; b < 8; ++b) { const int c = c0 + b; av[b] = (c <= lane && lane < nbn) ? A(kn + lane, kn + c) : 0.0; }
      }
      if (it > 0) {
        double sacc[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) sacc[b] = 0.0;
#pragma unroll 8
        for (int k = 0; k < LA_NB; ++k) {
          const double pr = S.P[k * S.pld + lane];
          const double2 q0 = *reinterpret_cast<const double2*>(&S.P[k * S.pld + c0]);
          const double2 q1 = *reinterpret_cast<const double2*>(&S.P[k * S.pld + c0 + 2]);
          const double2 q2 = *reinterpret_cast<const double2*>(&S.P[k * S.pld + c0 + 4]);
          const double2 q3 = *reinterpret_cast<const double2*>(&S.P[k * S.pld + c0 + 6]);
          sacc[0] = fma(pr, q0.x, sacc[0]); sacc[1] = fma(pr, q0.y, sacc[1]);
          sacc[2] = fma(pr, q1.x, sacc[2]); sacc[3] = fma(pr, q1.y, sacc[3]);
          sacc[4] = fma(pr, q2.x, sacc[4]); sacc[5] = fma(pr, q2.y, sacc[5]);
          sacc[6] = fma(pr, q3.x, sacc[6]); sacc[7] = fma(pr, q3.y, sacc[7]);
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) av[b] -= sacc[b];
      }
#pragma unroll
      for (int b = 0; b < 8; ++b) { const int c = c0 + b; S.D[lane * LA_DLD + c] = (c <= lane && lane < nbn) ? av[b] : 0.0; }
      la_bar_sync(1, LA_TEAM);
      if (it > 0) la_bar_arrive(2, LA_THREADS);                // the team no longer reads the solved panel it-1
      if (warp == 0) {
        GAMC_WCLK(w0, 52);
        // ---- column chain
        const int r = lane;
        double a[LA_NB];
#pragma unroll
        for (int c = 0; c < LA_NB; ++c) a[c] = S.D[r * LA_DLD + c];
        int bad = 0;
        double myrd = 0.0;
        // software-pipelined: the pivot broadcast and rsqrt of column j+1 are issued before the bulk update of column j,
        // so the long-latency part of the next column overlaps the (issue-bound) FMAs of this one
        double djj = __shfl_sync(0xffffffffu, a[0], 0);
        double rdv = rsqrt(djj);
#pragma unroll
        for (int j = 0; j < LA_NB; ++j) {
          if (j < nbn) {
            if (!(djj > 0.0) || !isfinite(djj)) { if (!bad) bad = kn + j + 1; }
            const double l = a[j] * rdv;             // lane j: sqrt(d_jj); lanes > j: l_ij; lanes < j: 0
            a[j] = l;
            if (r == j) { myrd = rdv; S.rd[j] = rdv; }
            S.LT[j * LA_DLD + r] = l;
            if (j + 1 < LA_NB) {                     // next pivot column first, through a shuffle (shortest dependent path)
              const double ln = __shfl_sync(0xffffffffu, l, j + 1);
              a[j + 1] = fma(-l, ln, a[j + 1]);
              djj = __shfl_sync(0xffffffffu, a[j + 1], j + 1);
              rdv = rsqrt(djj);
            }
            __syncwarp();
            if (r == 0) vflag[1] = it * LA_NB + j + 1;   // column j visible to the inverter (shared stores of a warp retire in order)
#pragma unroll
            for (int c = j + 2; c < LA_NB; ++c) a[c] = fma(-l, S.LT[j * LA_DLD + c], a[c]);   // a[c] stays unused for c > r
          }
        }
        GAMC_WCLK(w0, 53);
        if (bad && lane == 0) S.flag[0] = bad;
      } else if (warp == 4) {
        // ---- inverter: column `lane` of Lkk^-1 by forward substitution on e_lane, one column behind the chain;
        //      after step i every lane holds inv(i, lane): row i of the inverse is published for the panel solve
        const int r = lane;
        double y[LA_NB];
#pragma unroll
        for (int i = 0; i < LA_NB; ++i) y[i] = (i == r && r < nbn) ? 1.0 : 0.0;
#pragma unroll
        for (int i = 0; i < LA_NB; ++i) {
          if (i < nbn) {
            la_spin_until(vflag + 1, it * LA_NB + i + 1);
            const double yi = y[i] * S.rd[i];
            y[i] = yi;
            S.Y[i * LA_YLD + r] = yi;
            __syncwarp();
            if (r == 0) vflag[2] = it * LA_NB + i + 1;
#pragma unroll
            for (int k = i + 1; k < LA_NB; ++k) if (k < nbn) y[k] = fma(-S.LT[i * LA_DLD + k], yi, y[k]);
          }
        }
        GAMC_WCLK(w0, 55);
      } else if (warp == 8) {
        // ---- publisher of the factor: column j of the diagonal block to global memory as soon as the chain releases it
        for (int j = 0; j < nbn; ++j) {
          la_spin_until(vflag + 1, it * LA_NB + j + 1);
          if (lane >= j && lane < nbn) A(kn + lane, kn + j) = S.LT[j * LA_DLD + lane];
        }
        if (lane < nbn) rdiag[kn + lane] = S.rd[lane];
      } else if (invd) {
        // ---- publisher of the inverse (warp 12): row i in layout B as the inverter releases it, layout A at the end
        double* gB = invd + (size_t)it * LA_INVD_STRIDE + LA_NB * LA_NB;   // layout B: [i * 32 + j] = inv(i, j)
        double* gA = invd + (size_t)it * LA_INVD_STRIDE;                   // layout A: [j * 32 + i] = inv(i, j)
        for (int i = 0; i < nbn; ++i) {
          la_spin_until(vflag + 2, it * LA_NB + i + 1);
          gB[i * LA_NB + lane] = S.Y[i * LA_YLD + lane];
        }
        for (int j = 0; j < nbn; ++j) gA[j * LA_NB + lane] = (lane < nbn) ? S.Y[lane * LA_YLD + j] : 0.0;
      }
    } else {
      const int tt = la_trail_index(warp, lane);
      GAMC_WCLK_BEGIN(w2);
      if (it > 0) {
        // ---- trailing update A22 -= P P' (lower triangle) with panel it-1 on the FP64 tensor cores: one 32 x 32
        //      output block (4 x 4 DMMA m8n8k4 tiles, 8 k-steps) per warp task, minus block (0,0) = diagonal block
        //      `it` (the team's).  Fragments come straight from the k-major panel (conflict-free for pld = 4 mod 16).
        //      Block (1,1), the next diagonal block, is also staged in shared memory for the team's next pass.
        const int nb32 = (mprev + LA_NB - 1) / LA_NB;
        const int ntask = nb32 * (nb32 + 1) / 2;
        const int tw = tt >> 5;                                  // dense trailing-warp index 0..11
        const int fr = lane >> 2, fk = lane & 3;
        double* Dst = S.Dn + ((it + 1) & 1) * LA_NB * LA_DLD;
        for (int q = 1 + tw; q < ntask; q += LA_NTRAIL / 32) {
          // column-major triangular index -> (bi >= bj)
          int bj = 0, rem = q;
          while (rem >= nb32 - bj) { rem -= nb32 - bj; ++bj; }
          const int bi = bj + rem;
          const bool stage = (bi == 1 && bj == 1);
          double acc[4][4][2];
#pragma unroll
          for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
          const double* pa = S.P + 32 * bi + fr + (size_t)fk * S.pld;
          const double* pb = S.P + 32 * bj + fr + (size_t)fk * S.pld;
#pragma unroll 2
          for (int k4 = 0; k4 < LA_NB / 4; ++k4) {
            double af[4], bf[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) { af[mi] = pa[(size_t)(4 * k4) * S.pld + 8 * mi]; bf[mi] = pb[(size_t)(4 * k4) * S.pld + 8 * mi]; }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
              for (int ni = 0; ni < 4; ++ni) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
          }
          const int i0 = kn + 32 * bi + fr, j0 = kn + 32 * bj + 2 * fk;
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) {
            double old[4][2];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int i = i0 + 8 * mi, j = j0 + 8 * ni + e;
                old[ni][e] = (i < p && j < p && i >= j) ? A(i, j) : 0.0;
              }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int i = i0 + 8 * mi, j = j0 + 8 * ni + e;
                const double v = old[ni][e] - acc[mi][ni][e];
                if (i < p && j < p && i >= j) A(i, j) = v;
                if (stage) Dst[(fr + 8 * mi) * LA_DLD + 2 * fk + 8 * ni + e] = (i < p && j < p && i >= j) ? v : 0.0;
              }
          }
        }
        la_bar_sync(2, LA_THREADS);                  // every reader of the solved panel it-1 is done
        if (warp == 1) GAMC_WCLK(w2, 56);
      }
      // ---- panel solve X = A_panel Lkk^-T on the tensor cores: X(t, c) = sum_{k <= c} A_panel(t, k) inv(c, k).
      //      One 8-row block per warp task; A fragments straight from global memory (issued before waiting for the
      //      inverter), B fragments from the shared inverse (conflict-free for LA_YLD = 36); k-blocks above the
      //      diagonal of the inverse are skipped.
      if (mnext > 0) {
        const int fr = lane >> 2, fk = lane & 3;
        const int nrb = (mnext + 7) >> 3;
        constexpr int NW = LA_NTRAIL / 32;
        for (int rb0 = tt >> 5; rb0 < nrb; rb0 += 3 * NW) {       // three row blocks per warp in flight: one L2 round trip
          double af[3][LA_NB / 4];
#pragma unroll
          for (int s3 = 0; s3 < 3; ++s3) {
            const int t = 8 * (rb0 + s3 * NW) + fr;
#pragma unroll
            for (int k4 = 0; k4 < LA_NB / 4; ++k4) af[s3][k4] = (t < mnext) ? A(kn + LA_NB + t, kn + 4 * k4 + fk) : 0.0;
          }
          la_spin_until(vflag + 2, it * LA_NB + nbn);
#pragma unroll
          for (int s3 = 0; s3 < 3; ++s3) {
            const int rb = rb0 + s3 * NW;
            if (rb < nrb) {
              const int t = 8 * rb + fr;
              double acc[4][2];
#pragma unroll
              for (int ni = 0; ni < 4; ++ni) { acc[ni][0] = 0.0; acc[ni][1] = 0.0; }
#pragma unroll
              for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int k4 = 0; k4 <= 2 * ni + 1; ++k4)
                  dmma_8x8x4(acc[ni][0], acc[ni][1], af[s3][k4], S.Y[(8 * ni + fr) * LA_YLD + 4 * k4 + fk]);
#pragma unroll
              for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int c = 8 * ni + 2 * fk + e;
                  S.P[c * S.pld + t] = acc[ni][e];
                  if (t < mnext) A(kn + LA_NB + t, kn + c) = acc[ni][e];
                }
            }
          }
        }
      }
      if (warp == 1) GAMC_WCLK(w2, 57);
    }
    __syncthreads();
    if (REV) GAMC_PHASE(48);
    if (S.flag[0]) break;
  }
  // sum_i log Lm_ii = -sum_i log rdiag_i, in parallel at the end (keeps the logarithm off the column chain)
  const int bad = S.flag[0];
  double sl = 0.0;
  if (!bad) for (int i = tid; i < p; i += LA_THREADS) sl -= log(rdiag[i]);
  const double out = block_sum(sl, S.D);
  __syncthreads();
  if (fail_pivot) *fail_pivot = bad;
  return bad ? nan("") : out;
}

// Forward substitution Lm x = b in place on b (shared memory, M-space order).  invd = the inverted diagonal blocks
// written by chol_blocked.  Right-looking by block column: the last warp solves the diagonal block as a 32 x 32
// matrix-vector product with the stored inverse while threads 0..479 already hold "their" row of the block column in
// registers (prefetched one step ahead), so one step costs two barriers and no exposed L2 round trip.
template <bool REV>
__device__ void trsv_lower(MatView<REV> L, const double* invd, double* b, LaScratch S) {
  constexpr int RT = LA_THREADS - 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool solver = warp == (LA_THREADS / 32 - 1);
  const int p = L.p;
  const int nblk = (p + LA_NB - 1) / LA_NB;
  double reg[LA_NB];     // solver warp: row `lane` of the inverted block; others: row kb+32+tid of the block column
  auto prefetch = [&](int blk) {
    const int kb = blk * LA_NB;
    if (solver) {
      const double* gA = invd + (size_t)blk * LA_INVD_STRIDE;
#pragma unroll
      for (int j = 0; j < LA_NB; ++j) reg[j] = gA[j * LA_NB + lane];
    } else {
      const int i = kb + LA_NB + tid;
      if (i < p) {
#pragma unroll
        for (int c = 0; c < LA_NB; ++c) reg[c] = L(i, kb + c);
      }
    }
  };
  prefetch(0);
  __syncthreads();
  for (int blk = 0; blk < nblk; ++blk) {
    const int kb = blk * LA_NB;
    const int nb = min(LA_NB, p - kb);
    if (solver) {
      double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
#pragma unroll
      for (int j = 0; j < LA_NB; j += 4) {
        if (j < nb) x0 = fma(reg[j], b[kb + j], x0);
        if (j + 1 < nb) x1 = fma(reg[j + 1], b[kb + j + 1], x1);
        if (j + 2 < nb) x2 = fma(reg[j + 2], b[kb + j + 2], x2);
        if (j + 3 < nb) x3 = fma(reg[j + 3], b[kb + j + 3], x3);
      }
      const double x = (x0 + x1) + (x2 + x3);
      __syncwarp();
      if (lane < nb) { b[kb + lane] = x; S.xs[lane] = x; }
    }
    __syncthreads();
    if (!solver) {
      const int i = kb + LA_NB + tid;
      if (i < p) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int c = 0; c < LA_NB; c += 4) {
          s0 = fma(reg[c], S.xs[c], s0); s1 = fma(reg[c + 1], S.xs[c + 1], s1);
          s2 = fma(reg[c + 2], S.xs[c + 2], s2); s3 = fma(reg[c + 3], S.xs[c + 3], s3);
        }
        b[i] -= (s0 + s1) + (s2 + s3);
      }
      for (int i2 = i + RT; i2 < p; i2 += RT) {     // p > 512 only
        double s = 0.0;
        for (int c = 0; c < LA_NB; ++c) s = fma(L(i2, kb + c), S.xs[c], s);
        b[i2] -= s;
      }
    }
    if (blk + 1 < nblk) prefetch(blk + 1);
    __syncthreads();
  }
}

// Backward substitution Lm' x = b in place on b (shared memory, M-space order); same scheme, blocks in descending order,
// thread j < kb holds column j of the block row.
template <bool REV>
__device__ void trsv_lower_t(MatView<REV> L, const double* invd, double* b, LaScratch S) {
  constexpr int RT = LA_THREADS - 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool solver = warp == (LA_THREADS / 32 - 1);
  const int p = L.p;
  const int nblk = (p + LA_NB - 1) / LA_NB;
  double reg[LA_NB];     // solver warp: column `lane` of the inverted block; others: column tid of the block row
  auto prefetch = [&](int blk) {
    const int kb = blk * LA_NB;
    const int nb = min(LA_NB, p - kb);
    if (solver) {
      const double* gB = invd + (size_t)blk * LA_INVD_STRIDE + LA_NB * LA_NB;
#pragma unroll
      for (int i = 0; i < LA_NB; ++i) reg[i] = gB[i * LA_NB + lane];
    } else if (tid < kb) {
      // the 32 entries of column `tid` in this block row are contiguous in memory: 16-byte loads halve the number of
      // sector requests of this uncoalesced (one line per lane) access -- the LSU, not latency, bounds this loop
      if (REV && nb == LA_NB && (p & 1) == 0 && L.ld == p) {
#pragma unroll
        for (int c = 0; c < LA_NB; c += 2) {
          const double2 v = *reinterpret_cast<const double2*>(&L(kb + c + 1, tid));   // reversed view: (kb+c+1, tid) sits just below (kb+c, tid)
          reg[c + 1] = v.x; reg[c] = v.y;
        }
      } else {
#pragma unroll
        for (int c = 0; c < LA_NB; ++c) reg[c] = (c < nb) ? L(kb + c, tid) : 0.0;
      }
    }
  };
  prefetch(nblk - 1);
  __syncthreads();
  for (int blk = nblk - 1; blk >= 0; --blk) {
    const int kb = blk * LA_NB;
    const int nb = min(LA_NB, p - kb);
    if (solver) {
      double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
#pragma unroll
      for (int i = 0; i < LA_NB; i += 4) {
        if (i < nb) x0 = fma(reg[i], b[kb + i], x0);
        if (i + 1 < nb) x1 = fma(reg[i + 1], b[kb + i + 1], x1);
        if (i + 2 < nb) x2 = fma(reg[i + 2], b[kb + i + 2], x2);
        if (i + 3 < nb) x3 = fma(reg[i + 3], b[kb + i + 3], x3);
      }
      const double x = (x0 + x1) + (x2 + x3);
      __syncwarp();
      if (lane < nb) { b[kb + lane] = x; S.xs[lane] = x; }
      if (lane >= nb) S.xs[lane] = 0.0;
    }
    __syncthreads();
    if (!solver) {
      if (tid < kb) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int c = 0; c < LA_NB; c += 4) {
          s0 = fma(reg[c], S.xs[c], s0); s1 = fma(reg[c + 1], S.xs[c + 1], s1);
          s2 = fma(reg[c + 2], S.xs[c + 2], s2); s3 = fma(reg[c + 3], S.xs[c + 3], s3);
        }
        b[tid] -= (s0 + s1) + (s2 + s3);
      }
      for (int j2 = tid + RT; j2 < kb; j2 += RT) {  // p > 512 only
        double s = 0.0;
        for (int c = 0; c < nb; ++c) s = fma(L(kb + c, j2), S.xs[c], s);
        b[j2] -= s;
      }
    }
    if (blk > 0) prefetch(blk - 1);
    __syncthreads();
  }
}

// Streaming helper for the p x p passes of the step kernels: U independent loads in flight per thread (the single CTA
// is L2-latency bound, not bandwidth bound), then f(e, value) on each.
template <int U, typename F>
__device__ __forceinline__ void flat_stream(const double* src, int n, F f) {
  for (int base = 0; base < n; base += U * (int)blockDim.x) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const int e = base + u * (int)blockDim.x + (int)threadIdx.x; v[u] = (e < n) ? src[e] : 0.0; }
#pragma unroll
    for (int u = 0; u < U; ++u) { const int e = base + u * (int)blockDim.x + (int)threadIdx.x; if (e < n) f(e, v[u]); }
  }
}

// x' A x for a full p x p column-major matrix, x in shared memory; `red` = block_sum scratch.  All threads get the value.
__device__ inline double quad_form(const double* A, int p, const double* x, double* red) {
  double q = 0.0;
  flat_stream<16>(A, p * p, [&](int e, double v) {
    const int j = e / p, i = e - j * p;
    q = fma(v * x[i], x[j], q);
  });
  return block_sum(q, red);
}

// y = L x for a lower-triangular column-major L (natural order); x, y in shared memory.
__device__ inline void trmv_lower(const double* __restrict__ L, int ld, int p, const double* x, double* y) {
  for (int i = threadIdx.x; i < p; i += blockDim.x) {
    double s0 = 0.0, s1 = 0.0;
    int j = 0;
    for (; j + 1 <= i; j += 2) {
      s0 = fma(L[(size_t)i + (size_t)j * ld], x[j], s0);
      s1 = fma(L[(size_t)i + (size_t)(j + 1) * ld], x[j + 1], s1);
    }
    if (j <= i) s0 = fma(L[(size_t)i + (size_t)j * ld], x[j], s0);
    y[i] = s0 + s1;
  }
}

}  // namespace gamc
