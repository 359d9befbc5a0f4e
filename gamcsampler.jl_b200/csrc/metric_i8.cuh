// metric_i8.cuh -- the Fisher metric  G_lik = X' diag(w) X  on the INTEGER tensor cores (tcgen05.mma kind::i8, TMEM
// accumulators, TMA-fed), an error-free-transformation ("Ozaki") alternative to the FP64 DMMA kernel of metric.cuh.
//
// Replaces `broadcast(*, r.*(1-r), X)' * X` of examples/logistic_regression/src/GAMC/simulations.jl:36.
//
// FP64 has no tcgen05 kind, and the DMMA pipe bounds metric.cuh at 37 TFLOP/s.  Here S = diag(sqrt w) X is brought to
// fixed point per column, S_nj = q_nj 2^(e_j - E) (E = 8 k - 2 bits, 2^e_j >= max_n |x_nj| / 2 >= max_n |S_nj| because
// w = sigma (1 - sigma) <= 1/4), and q is cut into k signed base-256 digits d_1 (most significant) .. d_k in [-128, 127],
// one int8 plane per digit.  Then, EXACTLY in integers,
//     (S'S)_ij = 2^(e_i + e_j - 2E) sum_{a, b} 256^(2k - a - b) (D_a' D_b)_ij ,
// and the products with a + b > smax are dropped (their weight is 2^(4 - 8 (a + b))).  With k = 6, smax = 7 (21 digit
// products) the result agrees with the FP64 evaluation to 1e-13 of the largest entry (tests/test_gpu_i8.py, numpy
// statement of the scheme in tests/i8_ref.py); the products themselves are exact, so the result does not depend on the
// order of summation inside the tensor core.
//
// Kernels: k_i8_colmax / k_i8_scales (once per target: column exponents), k_i8_slice (per evaluation: digit planes of
// diag(sqrt w) X, HBM-bound), k_metric_i8 (the int8 GEMMs: warp-specialised TMA producer / single-thread tcgen05.mma
// issuer / TMEM epilogue with exact int64 atomics), k_reduce_i8 (weights, column scales, landing buffer).
#pragma once
#include "common.cuh"
#include "datapass.cuh"

namespace gamc {

#ifdef GAMC_PHASE_CLOCKS
#define I8_CLK_DECL(v) long long v = 0
#define I8_CLK_T(t) const long long t = clock64()
#define I8_CLK_ADD(v, t0) v += clock64() - (t0)
#define I8_CLK_OUT(k, v) do { if ((threadIdx.x & 31) == 0) { atomicAdd((unsigned long long*)&g_phase_acc[k], (unsigned long long)(v)); atomicAdd((unsigned long long*)&g_phase_cnt[k], 1ull); } } while (0)
#else
#define I8_CLK_DECL(v) do { } while (0)
#define I8_CLK_T(t) do { } while (0)
#define I8_CLK_ADD(v, t0) do { } while (0)
#define I8_CLK_OUT(k, v) do { } while (0)
#endif

constexpr int I8_TILE = 128;                       // columns of X per operand tile = edge of an output tile
#ifndef GAMC_I8_KC
#define GAMC_I8_KC 128
#endif
constexpr int I8_KC = GAMC_I8_KC;                  // rows (the contraction index) per chunk = one swizzle row of 64 or 128 bytes
static_assert(I8_KC == 64 || I8_KC == 128, "chunk = one SWIZZLE_64B or SWIZZLE_128B row");
constexpr int I8_TILE_BYTES = I8_TILE * I8_KC;     // 8 KB (16 KB): finer tiles leave more of the ring in flight while a chunk is consumed
constexpr int I8_MAXT = 6;                         // operand tiles per stage
constexpr int I8_MAXP = 12;                        // digit products per kind
constexpr int I8_MAXACC = 4;                       // accumulators of 128 TMEM columns: all 512 columns of the SM
constexpr int I8_MAXK = 7;                         // digit planes
constexpr int I8_THREADS = 192;                    // warps 0-3: epilogue (TMEM lane quarters 0-3); warp 4: TMA producer; warp 5: MMA issuer, TMEM allocator
constexpr int I8_SLICE_ROWS = 128, I8_SLICE_COLS = 64;

struct I8Kind {
  int bi, bj;                  // output tile (column blocks of X), bi >= bj
  int tile;                    // index of the output tile (position of its accumulators in the int64 buffer)
  int ntiles;                  // operand tiles per chunk
  int tile_slice[I8_MAXT];     // digit plane (0 = most significant)
  int tile_blk[I8_MAXT];       // column block
  int npairs;
  int pair_a[I8_MAXP], pair_b[I8_MAXP], pair_acc[I8_MAXP];   // operand tiles and accumulator of each digit product
  int nacc;
  int acc_slot[I8_MAXACC];     // group slot of accumulator g inside the tile's block of the int64 buffer (see below)
  int max_chunks;              // chunks one int32 accumulation may span (head room of 2^31)
};
// Group slots of an output tile in the int64 buffer: slot s - 2 holds sum_{a + b = s} D_a' D_b over the ORDERED digit pairs the
// planner issued with that weight; for a diagonal tile only a <= b is issued and the a < b products go to slot (smax - 1) + s - 2,
// which the reduction adds together with its transpose (D_a' D_b + D_b' D_a).
struct I8Item { int kind, chunk0, nchunks; };

constexpr int I8_RING = 13 * 16384 / I8_TILE_BYTES;   // operand tile slots (208 KB ring; a chunk takes the kind's ntiles consecutive slots)
constexpr int I8_IQ = 4;                           // items in flight between the producer and the other roles
constexpr int I8_CQ = 16;                          // chunk barriers (more than the I8_RING chunks that can be in flight)
struct I8Ctl {
  alignas(8) uint64_t full[I8_CQ];                 // chunk cs: all its operand tiles have landed
  alignas(8) uint64_t done[I8_CQ];                 // chunk cs: its MMAs have read the tiles (tcgen05.commit)
  int nt_of[I8_CQ];                                // producer: tiles of the chunks in flight
  alignas(8) uint64_t item_full[I8_IQ], item_empty[I8_IQ];
  alignas(8) uint64_t tmem_full, tmem_empty;
  uint32_t tmem_base;
  int item[I8_IQ];
  I8Kind kind;                                     // the MMA issuer's copy of the current item's kind
  uint32_t pk[I8_MAXP];                            // its digit products, packed: A tile | B tile << 8 | accumulator << 16
};
constexpr size_t I8_SMEM_BYTES = (size_t)I8_RING * I8_TILE_BYTES + 1024 + sizeof(I8Ctl);

// ---- tcgen05 / TMEM wrappers (SASS: UTCIMMA / UTCBAR / LDTM) ------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// One lane of a CONVERGED warp (the form ptxas recognises as a single-thread region: operands of the tcgen05 / TMA instructions
// issued under it move to uniform registers once instead of through a per-lane waterfall loop).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_alloc_512(uint32_t* dst_smem) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(512u) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t taddr) {    // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(512u) : "memory");
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]', int8 x int8 -> int32, M = 128, N = 128, K = 32 per instruction.
// ptxas serves every thread that executes tcgen05.mma through a loop (ELECT, operands to uniform registers with R2UR, UTCIMMA),
// and the R2UR moves are what limits the issue rate of a 64-cycle instruction: only the two 32-bit words that really vary -- the
// low words of the shared-memory descriptors -- are passed in registers; the high words (stride byte offset 1024 B, version 1,
// SWIZZLE_128B), the instruction descriptor and the accumulator column are immediates (the CTA owns all 512 TMEM columns, so
// its allocation starts at column 0, lane 0: checked at kernel start).
//   shared-memory matrix descriptor of a K-major operand tile written by TMA with the 128-byte (64-byte) swizzle: rows of I8_KC
//   bytes, 8-row groups 8 I8_KC bytes apart (stride byte offset, bits 32-45), leading byte offset 1 (unused), version 1 (bits 46-47),
//   layout 2 = SWIZZLE_128B or 4 = SWIZZLE_64B (bits 61-63);
//   instruction descriptor: D = s32 (bits 4-5 = 2), A and B signed 8 bit (bits 7-9, 10-12 = 1), both K-major (bits 15, 16 = 0),
//   N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__device__ __forceinline__ constexpr uint32_t tc_idesc_i8(int m, int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
constexpr uint32_t TC_DESC_HI_SW128 = (uint32_t)(8 * I8_KC / 16) | (1u << 14) | ((I8_KC == 128 ? 2u : 4u) << 29);   // bits 32-63 of the descriptor (layout 4 = SWIZZLE_64B)
__device__ __forceinline__ uint32_t tc_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
template <int ACC>
__device__ __forceinline__ void tc_mma_i8(uint32_t desc_a_lo, uint32_t desc_b_lo, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 td, id;\n\t"
      "setp.ne.b32 p, %2, 0;\n\t"
      "mov.b64 da, {%0, %3};\n\t"
      "mov.b64 db, {%1, %3};\n\t"
      "mov.b32 td, %4;\n\t"
      "mov.b32 id, %5;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [td], da, db, id, p;\n\t"
      "}" ::"r"(desc_a_lo), "r"(desc_b_lo), "r"(accumulate), "n"(TC_DESC_HI_SW128), "n"(ACC * I8_TILE), "n"(tc_idesc_i8(I8_TILE, I8_TILE)));
}
// the four K = 32 steps of one digit product on one 128-row chunk into accumulator ACC
template <int ACC>
__device__ __forceinline__ void tc_product(uint32_t da_lo, uint32_t db_lo, uint32_t accumulate) {
  tc_mma_i8<ACC>(da_lo, db_lo, accumulate);
#pragma unroll
  for (int kk = 1; kk < I8_KC / 32; ++kk) tc_mma_i8<ACC>(da_lo + 2 * kk, db_lo + 2 * kk, 1u);   // 32 bytes of K further: start address + 32 B (>> 4 = 2)
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// ---- column exponents (once per target) ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_i8_colmax(const double* __restrict__ X, long long ld, long long N, int p,
                                                   unsigned long long* __restrict__ cmax) {
  const int j = blockIdx.x;
  double m = 0.0;
  for (long long r = (long long)blockIdx.y * 256 + threadIdx.x; r < N; r += (long long)gridDim.y * 256) m = fmax(m, fabs(X[r + (size_t)j * ld]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&cmax[j], (unsigned long long)__double_as_longlong(m));   // non-negative doubles order like their bit patterns
}

// cs[j] = 2^(E - e_j) (fixed-point scale of column j), cg[j] = 2^e_j, 2^e_j >= max|x_j| / 2
__global__ void k_i8_scales(const unsigned long long* __restrict__ cmax, int p, int ebits, double* __restrict__ cs, double* __restrict__ cg) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p) return;
  const double m = __longlong_as_double((long long)cmax[j]);
  int e = 0;
  if (m > 0.0 && isfinite(m)) { frexp(m, &e); e -= 1; }       // m = f 2^(e+1), f in [0.5, 1)  =>  2^e >= m / 2
  cs[j] = ldexp(1.0, ebits - e);
  cg[j] = ldexp(1.0, e);
}

// ---- digit planes of diag(sqrt w) X (per evaluation) -----------------------------------------------------------------------
// Digit-plane layout in HBM: [chunk of I8_KC rows][plane][column][I8_KC bytes], so that an operand tile of the GEMM (one plane, 128
// columns, one chunk) is 8 KB (16 KB) of CONTIGUOUS memory (one DRAM page run instead of 128 segments a column apart).
// Block = 128 rows (one chunk) x 64 columns; warp = 8 columns; lane = 4 consecutive rows: 1 KB coalesced loads of X per warp and
// column, 128-byte coalesced stores per warp, column and plane; rows past N are written as zeros.  q' = rint(x sqrt(w) 2^(E - e_j)) + B with B = sum_{i < k-1} 128 256^i:
// the bytes of q' are the digits, offset by 128 except the top one (balanced digits without a carry chain).
__global__ void __launch_bounds__(256)
k_i8_slice(const double* __restrict__ X, long long ld, const double* __restrict__ w, long long N, int p, const double* __restrict__ cs,
           int k, int8_t* __restrict__ S, double bias, const int* skip_flag) {
  if (skip_flag && *reinterpret_cast<const volatile int*>(skip_flag)) return;
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const long long r = (long long)blockIdx.x * I8_SLICE_ROWS + 4 * lane;
  const size_t chunk = (size_t)(r / I8_KC);
  const int coff = (int)(r % I8_KC);
  double sw[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) sw[u] = (r + u < N) ? sqrt(w[r + u]) : 0.0;
  const int j0 = blockIdx.y * I8_SLICE_COLS + wq * 8;
#pragma unroll 2
  for (int jj = 0; jj < 8; ++jj) {
    const int j = j0 + jj;
    if (j >= p) break;
    const double* col = X + (size_t)j * ld + r;
    double x[4];
    if (r + 3 < N) {
      const double2 v0 = *reinterpret_cast<const double2*>(col), v1 = *reinterpret_cast<const double2*>(col + 2);
      x[0] = v0.x; x[1] = v0.y; x[2] = v1.x; x[3] = v1.y;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = (r + u < N) ? col[u] : 0.0;
    }
    const double c = cs[j];
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long q = __double2ll_rn(fma(x[u] * sw[u], c, bias));
      lo[u] = (uint32_t)q; hi[u] = (uint32_t)((unsigned long long)q >> 32);
    }
    for (int s = 0; s < k; ++s) {
      const int b = k - 1 - s;                       // byte of q' that holds digit s + 1
      const uint32_t sel = (uint32_t)(b & 3) | ((uint32_t)(4 + (b & 3)) << 4);
      uint32_t t01, t23;
      if (b < 4) { t01 = __byte_perm(lo[0], lo[1], sel); t23 = __byte_perm(lo[2], lo[3], sel); }
      else       { t01 = __byte_perm(hi[0], hi[1], sel); t23 = __byte_perm(hi[2], hi[3], sel); }
      uint32_t word = __byte_perm(t01, t23, 0x5410);
      if (s > 0) word ^= 0x80808080u;
      *reinterpret_cast<uint32_t*>(S + ((chunk * k + s) * p + j) * I8_KC + coff) = word;
    }
  }
}

// ---- gradient pass + digit planes in ONE read of X ------------------------------------------------------------------------
// k_datapass_i8<NU>: k_datapass<NU, 1, true> (log-likelihood, gradient, weights; same TMA ring, same phase 1, so the scalar is
// bit-identical) whose phase 2 also cuts diag(sqrt w) X into digit planes while the row block is still resident in shared memory:
// the level-2 evaluation of the integer route reads X once (k_i8_slice would read it a second time).  Phase 2 mapping: lane =
// 4 consecutive rows (lane & 7) x 2 of the warp's 8 columns per unit (lane >> 3, + 4): 32-byte shared loads without bank
// conflicts, and the four digits of one plane and column go out as one 32-bit store (a quarter-warp writes one 32-byte sector).
constexpr int DPI_WARPS = 16;                      // consumer warps: the digit extraction is latency-bound, so twice the warps of k_datapass
constexpr int DPI_THREADS = (DPI_WARPS + 1) * 32;
struct DataPassI8Smem {
  alignas(128) double ring[DP_NSLOT][DP_UNIT_DOUBLES];
  double theta[512];
  double cs[512];
  double red[2][DPI_WARPS][32];
  alignas(8) uint64_t full_bar[DP_NSLOT];
  alignas(8) uint64_t empty_bar[DP_NSLOT];
};

template <int NU>
__global__ void __launch_bounds__(DPI_THREADS, 1)
k_datapass_i8(const __grid_constant__ CUtensorMap mapX, const double* __restrict__ y, const double* __restrict__ theta, int p, long long N,
              double* __restrict__ w_out, double* __restrict__ part_ll, double* __restrict__ part_g, int part_g_stride,
              const double* __restrict__ cs, int k, int8_t* __restrict__ S8, double bias, const int* skip_flag) {
  if (skip_flag && *reinterpret_cast<const volatile int*>(skip_flag)) return;
  constexpr int RB = 32, CU = 64, CPW = CU / DPI_WARPS, CPL = CPW / 4;   // columns per warp and unit; per lane
  extern __shared__ __align__(128) unsigned char smem_raw[];
  DataPassI8Smem& S = *reinterpret_cast<DataPassI8Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long nrb = (N + RB - 1) / RB;

  for (int c = tid; c < NU * CU; c += DPI_THREADS) { S.theta[c] = (c < p) ? theta[c] : 0.0; S.cs[c] = (c < p) ? cs[c] : 0.0; }
  if (tid == 0) {
    for (int s = 0; s < DP_NSLOT; ++s) { mbar_init(&S.full_bar[s], 1); mbar_init(&S.empty_bar[s], DPI_WARPS); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == DPI_WARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&mapX);
      unsigned q = 0;
      for (long long rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
#pragma unroll 1
        for (int u = 0; u < NU; ++u, ++q) {
          const unsigned slot = q % DP_NSLOT, round = q / DP_NSLOT;
          mbar_wait(&S.empty_bar[slot], (round & 1) ^ 1);
          mbar_arrive_expect_tx(&S.full_bar[slot], DP_UNIT_DOUBLES * 8);
          tma_load_2d(S.ring[slot], &mapX, &S.full_bar[slot], (int)(rb * RB), u * CU);
        }
      }
    }
    return;
  }

  const int rg = lane & 7, cg = lane >> 3;
  // bytes 0 .. k-2 of q' carry digit + 128: one XOR turns them into the signed digits
  unsigned long long xmask = 0;
  for (int b = 0; b + 1 < k; ++b) xmask |= 0x80ull << (8 * b);
  const uint32_t xlo = (uint32_t)xmask, xhi = (uint32_t)(xmask >> 32);
  const size_t plane_stride = (size_t)p * I8_KC;
  double gacc[NU * CPL];
#pragma unroll
  for (int i = 0; i < NU * CPL; ++i) gacc[i] = 0.0;
  double ll = 0.0;
  unsigned q = 0;
  int it = 0;
  for (long long rb = blockIdx.x; rb < nrb; rb += gridDim.x, ++it) {
    const long long row0 = rb * RB + lane;
    const double yv = (row0 < N) ? __ldg(&y[row0]) : 0.0;
    // ---- phase 1 (as k_datapass): eta partial over this warp's columns, lane = row
    double eta = 0.0;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const unsigned qq = q + u, slot = qq % DP_NSLOT, round = qq / DP_NSLOT;
      mbar_wait(&S.full_bar[slot], round & 1);
      const double* base = &S.ring[slot][(warp * CPW) * RB + lane];
      const double* th = &S.theta[u * CU + warp * CPW];
#pragma unroll
      for (int jj = 0; jj < CPW; ++jj) eta = fma(base[jj * RB], th[jj], eta);
    }
    double* red = &S.red[it & 1][0][0];
    red[warp * 32 + lane] = eta;
    asm volatile("bar.sync 1, %0;" ::"n"(DPI_WARPS * 32) : "memory");
    double e = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < DPI_WARPS; ++w2) e += red[w2 * 32 + lane];
    // sigma as in sigmoid_softplus (common.cuh); the softplus is needed for the scalar only: warp 0
    const double ex = exp(-fabs(e));
    const double inv = 1.0 / (1.0 + ex);
    const double sig = (e >= 0.0) ? inv : ex * inv;
    const bool valid = row0 < N;
    const double rres = valid ? (yv - sig) : 0.0;
    const double wv = valid ? sig * (1.0 - sig) : 0.0;
    if (warp == 0) {
      const double sp = fmax(e, 0.0) + log1p(ex);
      ll += valid ? (yv * e - sp) : 0.0;
      if (valid) w_out[row0] = wv;
    }
    const double sw = sqrt(wv);
    // ---- phase 2: gradient partials and digit planes from the still-resident units; lane = rows 4 rg .. 4 rg + 3
    double rr[4], ss[4];
#pragma unroll
    for (int u4 = 0; u4 < 4; ++u4) { rr[u4] = __shfl_sync(0xffffffffu, rres, 4 * rg + u4); ss[u4] = __shfl_sync(0xffffffffu, sw, 4 * rg + u4); }
    const long long row = rb * RB + 4 * rg;
    int8_t* out0 = S8 + ((size_t)(row / I8_KC) * k + (size_t)(k - 1)) * plane_stride + (size_t)(row % I8_KC);   // plane k-1 = byte 0 of q'
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const unsigned qq = q + u, slot = qq % DP_NSLOT;
#pragma unroll
      for (int jj = 0; jj < CPL; ++jj) {
        const int cl = warp * CPW + cg + 4 * jj;
        const double2 xa = *reinterpret_cast<const double2*>(&S.ring[slot][cl * RB + 4 * rg]);
        const double2 xb = *reinterpret_cast<const double2*>(&S.ring[slot][cl * RB + 4 * rg + 2]);
        double a = gacc[u * CPL + jj];
        a = fma(xa.x, rr[0], a); a = fma(xa.y, rr[1], a); a = fma(xb.x, rr[2], a); a = fma(xb.y, rr[3], a);
        gacc[u * CPL + jj] = a;
        const int j = u * CU + cl;
        if (j < p) {
          const double c = S.cs[j];
          const long long q0 = __double2ll_rn(fma(xa.x * ss[0], c, bias)), q1 = __double2ll_rn(fma(xa.y * ss[1], c, bias));
          const long long q2 = __double2ll_rn(fma(xb.x * ss[2], c, bias)), q3 = __double2ll_rn(fma(xb.y * ss[3], c, bias));
          const uint32_t l0 = (uint32_t)q0 ^ xlo, l1 = (uint32_t)q1 ^ xlo, l2 = (uint32_t)q2 ^ xlo, l3 = (uint32_t)q3 ^ xlo;
          const uint32_t h0 = (uint32_t)((unsigned long long)q0 >> 32) ^ xhi, h1 = (uint32_t)((unsigned long long)q1 >> 32) ^ xhi;
          const uint32_t h2 = (uint32_t)((unsigned long long)q2 >> 32) ^ xhi, h3 = (uint32_t)((unsigned long long)q3 >> 32) ^ xhi;
          int8_t* o = out0 + (size_t)j * I8_KC;
#pragma unroll
          for (int b = 0; b < I8_MAXK; ++b) {
            if (b < k) {
              const uint32_t sel = (uint32_t)(b & 3) | ((uint32_t)(4 + (b & 3)) << 4);
              const uint32_t t01 = b < 4 ? __byte_perm(l0, l1, sel) : __byte_perm(h0, h1, sel);
              const uint32_t t23 = b < 4 ? __byte_perm(l2, l3, sel) : __byte_perm(h2, h3, sel);
              *reinterpret_cast<uint32_t*>(o - (size_t)b * plane_stride) = __byte_perm(t01, t23, 0x5410);
            }
          }
        }
      }
      if (NU > 4) {                                  // wide rows: the ring holds less than two row blocks, give every unit back at once
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.empty_bar[slot]);
      }
    }
    if (NU <= 4) {                                   // released together (no barrier between the units: their work interleaves)
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < NU; ++u) mbar_arrive(&S.empty_bar[(q + u) % DP_NSLOT]);
      }
    }
    q += NU;
  }

  if (warp == 0) {
    ll = warp_sum(ll);
    if (lane == 0) part_ll[blockIdx.x] = ll;
  }
#pragma unroll
  for (int u = 0; u < NU; ++u) {
#pragma unroll
    for (int jj = 0; jj < CPL; ++jj) {
      double v = gacc[u * CPL + jj];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      if (rg == 0) part_g[(size_t)blockIdx.x * part_g_stride + u * CU + warp * CPW + cg + 4 * jj] = v;
    }
  }
}

// ---- the int8 GEMMs ---------------------------------------------------------------------------------------------------------
// Persistent grid (one CTA per SM).  Work items (kind, block of consecutive 128-row chunks), ordered by rows so that all CTAs sweep
// the digit planes together (the L2 serves the re-reads), are claimed from a global counter by the producer warp and handed to the
// MMA issuer and the epilogue warps through a small ring.  Operand tiles travel through a ring of I8_RING 16 KB slots with a
// full / done mbarrier pair per chunk in flight, so the pipeline depth adapts to the kind's number of tiles and runs on across item boundaries.
// An item accumulates in int32 in tensor memory; its epilogue adds the accumulators to the tile's int64 group slots in global
// memory with integer atomics: exact, hence independent of which CTA took which item and in which order -- the evaluation is
// bit-reproducible although the schedule is dynamic.
__global__ void __launch_bounds__(I8_THREADS, 1)
k_metric_i8(const __grid_constant__ CUtensorMap mapS, const I8Kind* __restrict__ kinds, const I8Item* __restrict__ items, int nitems,
            int* __restrict__ counter, long long* __restrict__ acc64, int nslots, const int* skip_flag) {
  if (skip_flag && *reinterpret_cast<const volatile int*>(skip_flag)) return;
  extern __shared__ unsigned char i8_smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t pad = (1024u - (smem_u32(i8_smem_raw) & 1023u)) & 1023u;      // SWIZZLE_128B tiles: 1024-byte aligned
  unsigned char* tiles = i8_smem_raw + pad;
  I8Ctl& C = *reinterpret_cast<I8Ctl*>(tiles + (size_t)I8_RING * I8_TILE_BYTES);
  if (tid == 0) {
    for (int s = 0; s < I8_CQ; ++s) { mbar_init(&C.full[s], 1); mbar_init(&C.done[s], 1); }
    for (int s = 0; s < I8_IQ; ++s) { mbar_init(&C.item_full[s], 1); mbar_init(&C.item_empty[s], 5); }   // consumers: MMA issuer + 4 epilogue warps
    mbar_init(&C.tmem_full, 1);
    mbar_init(&C.tmem_empty, 4);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc_512(&C.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = C.tmem_base;
  if (tmem != 0) __trap();                           // all 512 columns are ours: the MMA issue path hard-codes the base

  if (warp == 4) {
    // ===== producer: claims items, feeds the tile ring (the warp stays converged; one elected lane issues)
    if (elect_one()) tma_prefetch_desc(&mapS);
    unsigned n = 0;                                  // tile slots handed out so far
    unsigned cs = 0, head = 0;                       // chunks issued / chunks whose slots have been taken back
    int nfree = I8_RING;
    I8_CLK_DECL(clk_pw); I8_CLK_DECL(clk_pt); I8_CLK_T(tp0);
    for (unsigned iq = 0;; ++iq) {
      const int e = iq % I8_IQ;
      mbar_wait(&C.item_empty[e], ((iq / I8_IQ) & 1) ^ 1);
      int it = 0;
      if (lane == 0) it = atomicAdd(counter, 1);
      it = __shfl_sync(0xffffffffu, it, 0);
      if (lane == 0) { C.item[e] = it < nitems ? it : -1; mbar_arrive(&C.item_full[e]); }   // (arrive = release: the item index is visible to the waiters)
      if (it >= nitems) break;
      const I8Item im = items[it];
      const I8Kind* kd = kinds + im.kind;
      const int ntiles = kd->ntiles;
      int tsl[I8_MAXT], tbk[I8_MAXT];
#pragma unroll
      for (int t = 0; t < I8_MAXT; ++t) { tsl[t] = t < ntiles ? kd->tile_slice[t] : 0; tbk[t] = t < ntiles ? kd->tile_blk[t] * I8_TILE : 0; }
      for (int c = 0; c < im.nchunks; ++c, ++cs) {
        { I8_CLK_T(tw);
          while (nfree < ntiles) {                   // take back the slots of the oldest chunks in flight
            mbar_wait(&C.done[head % I8_CQ], (head / I8_CQ) & 1);
            nfree += C.nt_of[head % I8_CQ];
            ++head;
          }
          I8_CLK_ADD(clk_pw, tw); }
        nfree -= ntiles;
        if (elect_one()) {
          C.nt_of[cs % I8_CQ] = ntiles;
          uint64_t* fb = &C.full[cs % I8_CQ];
          mbar_arrive_expect_tx(fb, (uint32_t)ntiles * I8_TILE_BYTES);
#pragma unroll
          for (int t = 0; t < I8_MAXT; ++t) {
            if (t < ntiles) {
              unsigned sl = n % I8_RING + t; sl -= sl >= I8_RING ? I8_RING : 0;
              tma_load_4d(tiles + (size_t)sl * I8_TILE_BYTES, &mapS, fb, 0, tbk[t], tsl[t], im.chunk0 + c);
            }
          }
        }
        __syncwarp();
        n += ntiles;
      }
    }
    I8_CLK_ADD(clk_pt, tp0); I8_CLK_OUT(62, clk_pw); I8_CLK_OUT(63, clk_pt);
    __syncwarp();
  } else if (warp == 5) {
    // ===== MMA issuer (one thread issues; the warp copies the kind of each item to shared memory together)
    unsigned n = 0, nitem = 0, cs = 0;
    I8_CLK_DECL(clk_wf); I8_CLK_DECL(clk_wt); I8_CLK_DECL(clk_is); I8_CLK_DECL(clk_tot); I8_CLK_T(tm0);
    for (unsigned iq = 0;; ++iq) {
      const int e = iq % I8_IQ;
      mbar_wait(&C.item_full[e], (iq / I8_IQ) & 1);
      const int it = C.item[e];
      __syncwarp();
      if (lane == 0) mbar_arrive(&C.item_empty[e]);
      if (it < 0) break;
      const I8Item im = items[it];
      {
        const int* src = reinterpret_cast<const int*>(kinds + im.kind);
        int* dst = reinterpret_cast<int*>(&C.kind);
        for (int i = lane; i < (int)(sizeof(I8Kind) / sizeof(int)); i += 32) dst[i] = src[i];
        if (lane < I8_MAXP) C.pk[lane] = (uint32_t)kinds[im.kind].pair_a[lane] | ((uint32_t)kinds[im.kind].pair_b[lane] << 8) | ((uint32_t)kinds[im.kind].pair_acc[lane] << 16);
      }
      __syncwarp();
      {
        const int ntiles = C.kind.ntiles, npairs = C.kind.npairs;
        if (nitem > 0) { I8_CLK_T(tw); mbar_wait(&C.tmem_empty, (nitem - 1) & 1); tc_fence_after(); I8_CLK_ADD(clk_wt, tw); }   // the previous item's accumulators have been read
        uint32_t started = 0;
        const uint32_t sbase = smem_u32(tiles);
        const uint32_t d0 = tc_desc_lo(sbase);
        for (int c = 0; c < im.nchunks; ++c) {
          const unsigned base = n % I8_RING;
          { I8_CLK_T(tw); mbar_wait(&C.full[cs % I8_CQ], (cs / I8_CQ) & 1); I8_CLK_ADD(clk_wf, tw); }
          tc_fence_after();
          I8_CLK_T(ti);
          if (elect_one()) {
            for (int q = 0; q < npairs; ++q) {
              const uint32_t pk = C.pk[q];
              unsigned sa = base + (pk & 0xFFu), sb = base + ((pk >> 8) & 0xFFu);
              sa -= sa >= I8_RING ? I8_RING : 0; sb -= sb >= I8_RING ? I8_RING : 0;
              const uint32_t da = d0 + sa * (I8_TILE_BYTES >> 4), db = d0 + sb * (I8_TILE_BYTES >> 4);
              const uint32_t g = pk >> 16;
              const uint32_t accum = (started >> g) & 1u;
              if (g == 0) tc_product<0>(da, db, accum);
              else if (g == 1) tc_product<1>(da, db, accum);
              else if (g == 2) tc_product<2>(da, db, accum);
              else tc_product<3>(da, db, accum);
              started |= 1u << g;
            }
            tc_commit(&C.done[cs % I8_CQ]);          // the chunk's slots are free once these MMAs have read them
            if (c == im.nchunks - 1) tc_commit(&C.tmem_full);
          }
          __syncwarp();
          I8_CLK_ADD(clk_is, ti);
          // `started` is per lane: after the first chunk every accumulator of the kind has been written
          started = 0xFu;
          n += ntiles; ++cs;
        }
      }
      ++nitem;
      __syncwarp();
    }
    I8_CLK_ADD(clk_tot, tm0); I8_CLK_OUT(58, clk_wf); I8_CLK_OUT(59, clk_wt); I8_CLK_OUT(60, clk_is); I8_CLK_OUT(61, clk_tot);
  } else {
    // ===== epilogue: warp q reads TMEM lanes 32 q .. 32 q + 31 (= rows i of the output tile), thread = row
    const int i = warp * 32 + lane;
    unsigned nitem = 0;
    for (unsigned iq = 0;; ++iq) {
      const int e = iq % I8_IQ;
      mbar_wait(&C.item_full[e], (iq / I8_IQ) & 1);
      const int it = C.item[e];
      __syncwarp();
      if (lane == 0) mbar_arrive(&C.item_empty[e]);
      if (it < 0) break;
      const I8Kind* kd = kinds + items[it].kind;
      const int nacc = kd->nacc;
      long long* tbase = acc64 + (size_t)kd->tile * nslots * (I8_TILE * I8_TILE);
      int slot[I8_MAXACC];
#pragma unroll
      for (int g = 0; g < I8_MAXACC; ++g) slot[g] = g < nacc ? kd->acc_slot[g] : 0;
      mbar_wait(&C.tmem_full, nitem & 1);
      tc_fence_after();
      const bool diag = kd->bi == kd->bj;
      for (int g = 0; g < nacc; ++g) {
        unsigned long long* out = reinterpret_cast<unsigned long long*>(tbase + (size_t)slot[g] * (I8_TILE * I8_TILE)) + i;
        // a + b products with a = b on a diagonal tile are symmetric: only i >= j is read by the reduction
        const bool lower_only = diag && slot[g] < nslots / 2;
        for (int cb = 0; cb < I8_TILE / 32; ++cb) {
          if (lower_only && cb > warp) break;
          uint32_t r[32];
          tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * I8_TILE + cb * 32), r);
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (r[c] && !(lower_only && cb * 32 + c > i)) atomicAdd(out + (size_t)(cb * 32 + c) * I8_TILE, (unsigned long long)(long long)(int)r[c]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&C.tmem_empty);
      ++nitem;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc_512(tmem); }
}

// ---- exact int64 group sums -> landing buffer (both triangles, exactly symmetric).
//      G_ij = 2^(e_i + e_j) sum_s 2^(4 - 8 s) [slot_s(i, j) (+ slotA_s(i, j) + slotA_s(j, i) on diagonal tiles)].
//      (The sums and the item counter are cleared by a memset node before the next evaluation.)
__global__ void __launch_bounds__(256)
k_reduce_i8(double* __restrict__ E, int p, int goff, const long long* __restrict__ acc64, int smax, const int* __restrict__ tile_bi,
            const int* __restrict__ tile_bj, const double* __restrict__ cg, const int* skip_flag) {
  if (skip_flag && *reinterpret_cast<const volatile int*>(skip_flag)) return;
  constexpr int BPT = I8_TILE * I8_TILE / 256;       // blocks per tile
  const int t = blockIdx.x / BPT;
  const int e = (blockIdx.x % BPT) * 256 + threadIdx.x;
  const int i = e & (I8_TILE - 1), j = e >> 7;
  const int bi = tile_bi[t], bj = tile_bj[t];
  const int gi = bi * I8_TILE + i, gj = bj * I8_TILE + j;
  const bool diag = bi == bj;
  if (gi >= p || gj >= p || (diag && i < j)) return;
  const int ng = smax - 1, nslots = 2 * ng;
  const long long* tb = acc64 + (size_t)t * nslots * (I8_TILE * I8_TILE);
  const int et = i * I8_TILE + j;                    // the transposed element of the tile
  double s = 0.0;
  for (int q = ng - 1; q >= 0; --q) {                // smallest weights first
    long long v = tb[(size_t)q * (I8_TILE * I8_TILE) + e];
    if (diag) v += tb[(size_t)(ng + q) * (I8_TILE * I8_TILE) + e] + tb[(size_t)(ng + q) * (I8_TILE * I8_TILE) + et];
    s = fma((double)v, ldexp(1.0, 4 - 8 * (q + 2)), s);
  }
  s *= cg[gi] * cg[gj];
  double* G = E + goff;
  G[gi + (size_t)gj * p] = s;
  G[gj + (size_t)gi * p] = s;
}

}  // namespace gamc
