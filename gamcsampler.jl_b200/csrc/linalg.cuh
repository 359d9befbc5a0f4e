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

template <bool REV>
struct MatView {
  double* a;
  int ld, p;
  __device__ __forceinline__ double& operator()(int i, int j) const {
    return REV ? a[(size_t)(p - 1 - i) + (size_t)(p - 1 - j) * ld] : a[(size_t)i + (size_t)j * ld];
  }
};

// Shared-memory scratch for the factorisation: D (diagonal block), rd (reciprocal diagonal of the block),
// P (panel, k-major: P[c * pld + t]), flag.
struct LaScratch {
  double* D;      // LA_NB * LA_DLD
  double* rd;     // LA_NB
  double* P;      // LA_NB * pld
  int* flag;      // 2 ints
  int pld;
};

__host__ __device__ inline size_t la_scratch_bytes(int p) {
  const int pld = ((p + 3) & ~3) + 4;
  return sizeof(double) * (LA_NB * LA_DLD + LA_NB + (size_t)LA_NB * pld) + 16;
}

__device__ inline LaScratch la_scratch_carve(unsigned char* base, int p) {
  LaScratch s;
  s.pld = ((p + 3) & ~3) + 4;
  s.D = reinterpret_cast<double*>(base);
  s.rd = s.D + LA_NB * LA_DLD;
  s.P = s.rd + LA_NB;
  s.flag = reinterpret_cast<int*>(s.P + (size_t)LA_NB * s.pld);
  return s;
}

// In-place blocked right-looking Cholesky of the lower triangle (M space) of A: A = Lm Lm'.
// Also writes rdiag[i] = 1 / Lm_ii (M-space order, shared or global) and returns sum_i log Lm_ii.
// On a non-positive pivot returns NaN and sets *fail_pivot = 1-based M-space pivot index (else 0).
// Must be called by all LA_THREADS threads.
template <bool REV>
__device__ double chol_blocked(MatView<REV> A, double* rdiag, LaScratch S, int* fail_pivot) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = A.p;
  double sumlog = 0.0;  // meaningful in thread 0 of warp 0; broadcast at the end
  if (tid == 0) { S.flag[0] = 0; }
  __syncthreads();
  for (int kb = 0; kb < p; kb += LA_NB) {
    const int nb = min(LA_NB, p - kb);
    const int m = p - kb - nb;
    // (a) diagonal block -> shared
    for (int e = tid; e < LA_NB * LA_NB; e += LA_THREADS) {
      const int r = e & 31, c = e >> 5;
      if (r < nb && c < nb && r >= c) S.D[r * LA_DLD + c] = A(kb + r, kb + c);
    }
    __syncthreads();
    // (b) factor the diagonal block with warp 0: right-looking, lane = row, the row lives in registers; the
    //     freshly scaled column is broadcast through shared memory.  Dependent chain per column:
    //     shfl -> rsqrt -> mul -> (sts/lds) -> fma; no divide, no sqrt, the logarithms are taken in parallel afterwards.
    if (warp == 0) {
      const int r = lane;
      double a[LA_NB];
#pragma unroll
      for (int c = 0; c < LA_NB; ++c) a[c] = (c <= r && r < nb && c < nb) ? S.D[r * LA_DLD + c] : 0.0;
      int bad = 0;
      double* colbuf = S.rd;   // LA_NB doubles, reused as the reciprocal diagonal at the end
      double myrd = 0.0;
#pragma unroll
      for (int j = 0; j < LA_NB; ++j) {
        if (j < nb) {
          const double djj = __shfl_sync(0xffffffffu, a[j], j);
          if (!(djj > 0.0) || !isfinite(djj)) { if (!bad) bad = kb + j + 1; }
          const double rdv = rsqrt(djj);
          const double l = a[j] * rdv;               // lane j: sqrt(d_jj); lanes > j: l_ij; lanes < j: 0
          a[j] = l;
          if (r == j) myrd = rdv;
          colbuf[r] = l;
          __syncwarp();
#pragma unroll
          for (int c = j + 1; c < LA_NB; ++c) a[c] = fma(-l, colbuf[c], a[c]);   // a[c] is zero (and stays unused) for c > r
          __syncwarp();
        }
      }
#pragma unroll
      for (int c = 0; c < LA_NB; ++c) if (c <= r && r < nb && c < nb) S.D[r * LA_DLD + c] = a[c];
      __syncwarp();
      if (r < nb) { S.rd[r] = myrd; sumlog -= log(myrd); }
      if (bad && lane == 0) S.flag[0] = bad;
    }
    __syncthreads();
    if (S.flag[0]) break;
    // write back the factored diagonal block and the reciprocal diagonal
    for (int e = tid; e < LA_NB * LA_NB; e += LA_THREADS) {
      const int r = e & 31, c = e >> 5;
      if (r < nb && c < nb && r >= c) A(kb + r, kb + c) = S.D[r * LA_DLD + c];
    }
    if (tid < nb) rdiag[kb + tid] = S.rd[tid];
    if (m > 0) {
      // (c) panel solve: row i = kb+nb+t, x = a Lkk^-T, one row per thread
      for (int t = tid; t < m; t += LA_THREADS) {
        const int i = kb + nb + t;
        double x[LA_NB];
#pragma unroll
        for (int c = 0; c < LA_NB; ++c) x[c] = (c < nb) ? A(i, kb + c) : 0.0;
#pragma unroll
        for (int c = 0; c < LA_NB; ++c) {
          if (c < nb) {
            const double xc = x[c] * S.rd[c];
            x[c] = xc;
#pragma unroll
            for (int k = c + 1; k < LA_NB; ++k) if (k < nb) x[k] = fma(-xc, S.D[k * LA_DLD + c], x[k]);
          }
        }
#pragma unroll
        for (int c = 0; c < LA_NB; ++c) {
          if (c < nb) { A(i, kb + c) = x[c]; }
          S.P[c * S.pld + t] = x[c];
        }
      }
      // zero the padding rows of the panel so that 4x4 micro-tiles may over-read
      for (int t = m + tid; t < ((m + 3) & ~3); t += LA_THREADS)
        for (int c = 0; c < LA_NB; ++c) S.P[c * S.pld + t] = 0.0;
      __syncthreads();
      // (d) trailing update A22 -= P P' (lower triangle), 4 x 4 register micro-tiles
      const int mt = (m + 3) >> 2;
      const int ntile = mt * (mt + 1) / 2;
      for (int q = tid; q < ntile; q += LA_THREADS) {
        // decode column-major triangular index: column tj, row ti >= tj
        int tj = (int)(((2.0 * mt + 1.0) - sqrt((2.0 * mt + 1.0) * (2.0 * mt + 1.0) - 8.0 * (double)q)) * 0.5);
        if (tj < 0) tj = 0;
        while (tj > 0 && tj * mt - tj * (tj - 1) / 2 > q) --tj;
        while ((tj + 1) * mt - (tj + 1) * tj / 2 <= q) ++tj;
        const int ti = tj + (q - (tj * mt - tj * (tj - 1) / 2));
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
        const double* pa = S.P + 4 * ti;
        const double* pb = S.P + 4 * tj;
#pragma unroll 4
        for (int c = 0; c < LA_NB; ++c) {
          const double2 a01 = *reinterpret_cast<const double2*>(pa + c * S.pld);
          const double2 a23 = *reinterpret_cast<const double2*>(pa + c * S.pld + 2);
          const double2 b01 = *reinterpret_cast<const double2*>(pb + c * S.pld);
          const double2 b23 = *reinterpret_cast<const double2*>(pb + c * S.pld + 2);
          const double av[4] = {a01.x, a01.y, a23.x, a23.y};
          const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
        const int i0 = kb + nb + 4 * ti, j0 = kb + nb + 4 * tj;
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const int i = i0 + a, j = j0 + b;
            if (i < p && j < p && i >= j) A(i, j) -= acc[a][b];
          }
      }
    }
    __syncthreads();
  }
  // broadcast result (each lane of warp 0 holds the logs of "its" diagonal entries)
  __syncthreads();
  if (warp == 0) {
    sumlog = warp_sum(sumlog);
    if (lane == 0) S.D[0] = sumlog;
  }
  __syncthreads();
  const int bad = S.flag[0];
  const double out = S.D[0];
  __syncthreads();
  if (fail_pivot) *fail_pivot = bad;
  return bad ? nan("") : out;
}

// Forward substitution Lm x = b in place on b (shared memory, M-space order), rdiag = 1/Lm_ii.
template <bool REV>
__device__ void trsv_lower(MatView<REV> L, const double* rdiag, double* b, LaScratch S) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = L.p;
  for (int kb = 0; kb < p; kb += LA_NB) {
    const int nb = min(LA_NB, p - kb);
    for (int e = tid; e < LA_NB * LA_NB; e += LA_THREADS) {
      const int r = e & 31, c = e >> 5;
      if (r < nb && c < nb && r > c) S.D[r * LA_DLD + c] = L(kb + r, kb + c);
    }
    __syncthreads();
    if (warp == 0) {
      double br = (lane < nb) ? b[kb + lane] : 0.0;
      const double rdv = (lane < nb) ? rdiag[kb + lane] : 0.0;
      for (int j = 0; j < nb; ++j) {
        const double xj = __shfl_sync(0xffffffffu, br * rdv, j);
        if (lane == j) br = xj;
        else if (lane > j && lane < nb) br = fma(-S.D[lane * LA_DLD + j], xj, br);
      }
      if (lane < nb) { b[kb + lane] = br; S.rd[lane] = br; }
    }
    __syncthreads();
    for (int i = kb + nb + tid; i < p; i += LA_THREADS) {
      double s0 = b[i], s1 = 0.0;
#pragma unroll 8
      for (int c = 0; c < nb; c += 2) {
        s0 = fma(-L(i, kb + c), S.rd[c], s0);
        if (c + 1 < nb) s1 = fma(-L(i, kb + c + 1), S.rd[c + 1], s1);
      }
      b[i] = s0 + s1;
    }
    __syncthreads();
  }
}

// Backward substitution Lm' x = b in place on b (shared memory, M-space order).
template <bool REV>
__device__ void trsv_lower_t(MatView<REV> L, const double* rdiag, double* b, LaScratch S) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = L.p;
  const int nblk = (p + LA_NB - 1) / LA_NB;
  for (int blk = nblk - 1; blk >= 0; --blk) {
    const int kb = blk * LA_NB;
    const int nb = min(LA_NB, p - kb);
    for (int e = tid; e < LA_NB * LA_NB; e += LA_THREADS) {
      const int r = e & 31, c = e >> 5;
      if (r < nb && c < nb && r > c) S.D[r * LA_DLD + c] = L(kb + r, kb + c);
    }
    __syncthreads();
    if (warp == 0) {
      // solve Lkk' x = b: x_j = (b_j - sum_{i>j} L_ij x_i) / L_jj, j descending; lane = j
      double bj = (lane < nb) ? b[kb + lane] : 0.0;
      const double rdv = (lane < nb) ? rdiag[kb + lane] : 0.0;
      for (int i = nb - 1; i >= 0; --i) {
        const double xi = __shfl_sync(0xffffffffu, bj * rdv, i);
        if (lane == i) bj = xi;
        else if (lane < i) bj = fma(-S.D[i * LA_DLD + lane], xi, bj);
      }
      if (lane < nb) { b[kb + lane] = bj; S.rd[lane] = bj; }
    }
    __syncthreads();
    for (int j = tid; j < kb; j += LA_THREADS) {
      double s0 = b[j], s1 = 0.0;
#pragma unroll 8
      for (int c = 0; c < nb; c += 2) {
        s0 = fma(-L(kb + c, j), S.rd[c], s0);
        if (c + 1 < nb) s1 = fma(-L(kb + c + 1, j), S.rd[c + 1], s1);
      }
      b[j] = s0 + s1;
    }
    __syncthreads();
  }
}

// y = A x for a full symmetric p x p column-major matrix (natural order), x and y in shared memory.
__device__ inline void symv_full(const double* __restrict__ A, int ld, int p, const double* x, double* y) {
  for (int i = threadIdx.x; i < p; i += blockDim.x) {
    double s0 = 0.0, s1 = 0.0;
    int j = 0;
    for (; j + 1 < p; j += 2) {
      s0 = fma(A[(size_t)i + (size_t)j * ld], x[j], s0);
      s1 = fma(A[(size_t)i + (size_t)(j + 1) * ld], x[j + 1], s1);
    }
    if (j < p) s0 = fma(A[(size_t)i + (size_t)j * ld], x[j], s0);
    y[i] = s0 + s1;
  }
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
