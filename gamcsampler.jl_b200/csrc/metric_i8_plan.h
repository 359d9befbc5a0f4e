// metric_i8_plan.h -- host-side work decomposition of the integer-tensor-core metric (metric_i8.cuh).
//
// G_lik = S'S with S = diag(sqrt w) X cut into `nslices` signed base-256 digit planes D_1 (most significant) .. D_k:
//   G_ij = 2^(e_i + e_j) * sum_{a + b <= smax} 2^(4 - 8 (a + b)) (D_a' D_b)_ij .
// The lower triangle of G is cut into 128 x 128 output tiles (bi >= bj).  An off-diagonal tile needs every ORDERED digit
// pair (a, b); a diagonal tile only a <= b, because D_b' D_a = (D_a' D_b)' lies in the same tile: the a < b products are
// kept in separate group slots and the reduction adds them together with their transpose.  The digit products of a tile are
// grouped into "kinds".  A kind keeps at most I8_MAXACC accumulators in tensor memory (one per group slot: products with
// the same a + b carry the same weight and are summed exactly in int32), reads at most I8_MAXT operand tiles (digit plane
// x column block) per 128-row chunk, and issues at most I8_MAXP products per chunk.  Work items = (kind, block of chunks),
// ordered by rows (all kinds of a block next to each other, the expensive ones first) and claimed dynamically by the
// persistent CTAs, so that all CTAs sweep the digit planes together (the 126 MB L2 serves the re-reads); the blocks
// shrink towards the end of the rows so that the CTAs finish together.
#pragma once
#include <vector>
#include <algorithm>
#include <utility>
#include "metric_i8.cuh"

namespace gamc {

struct I8Layout {
  int nslices = 0, smax = 0, nblk = 0, nslots = 0;
  std::vector<I8Kind> kinds;
  std::vector<int> kind_cost;        // modelled cycles per chunk
  std::vector<I8Item> items;
  std::vector<int> tile_bi, tile_bj; // output tiles
};

// modelled cost of one 128-row chunk: tensor pipe (128 x 128 x 128 int8 MACs per product at 8192 MAC/clk = 256 clk) against
// the L2 -> SM path (16 KB per operand tile at ~40 B/clk)
inline int i8_cost(int npairs, int ntiles) { return std::max(256 * npairs, 410 * ntiles); }

struct I8Prod { int a, b, slot; };

// Greedy cover of the products of one output tile by kinds.
inline void i8_plan_tile(int bi, int bj, int tile, int k, int smax, std::vector<I8Kind>& out) {
  const bool diag = bi == bj;
  const int ng = smax - 1;
  std::vector<I8Prod> rem;
  for (int a = 1; a <= k; ++a)
    for (int b = 1; b <= k; ++b) {
      if (a + b > smax || (diag && a > b)) continue;
      rem.push_back({a, b, (diag && a < b) ? ng + (a + b - 2) : (a + b - 2)});
    }
  while (!rem.empty()) {
    double best_score = -1.0;
    std::vector<I8Prod> best;
    // candidate = the remaining products inside a rectangle of digit planes [a0, a1] x [b0, b1]; its products are taken in the
    // order of their weight until the accumulators, the operand tiles or the product list are exhausted
    for (int a0 = 1; a0 <= k; ++a0)
      for (int a1 = a0; a1 <= k; ++a1)
        for (int b0 = 1; b0 <= k; ++b0)
          for (int b1 = b0; b1 <= k; ++b1)
            for (int s0 = 2; s0 <= smax; ++s0) {
              std::vector<I8Prod> cand;
              std::vector<int> slots, planes_a, planes_b;
              auto add_unique = [](std::vector<int>& v, int x) { if (std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x); };
              for (int s = s0; s <= smax; ++s)
                for (auto& pr : rem) {
                  if (pr.a + pr.b != s || pr.a < a0 || pr.a > a1 || pr.b < b0 || pr.b > b1) continue;
                  std::vector<int> sl = slots, pa = planes_a, pb = planes_b;
                  add_unique(sl, pr.slot); add_unique(pa, pr.a);
                  if (diag) add_unique(pa, pr.b); else add_unique(pb, pr.b);
                  if ((int)sl.size() > I8_MAXACC || (int)(pa.size() + pb.size()) > I8_MAXT || (int)cand.size() + 1 > I8_MAXP) continue;
                  slots = sl; planes_a = pa; planes_b = pb;
                  cand.push_back(pr);
                }
              if (cand.empty()) continue;
              const int ntiles = (int)(planes_a.size() + planes_b.size());
              const double score = (double)cand.size() / i8_cost((int)cand.size(), ntiles) + 1e-9 * cand.size();
              if (score > best_score) { best_score = score; best = cand; }
            }
    I8Kind kd{};
    kd.bi = bi; kd.bj = bj; kd.tile = tile;
    auto tile_index = [&](int digit, int blk) {
      for (int t = 0; t < kd.ntiles; ++t) if (kd.tile_slice[t] == digit - 1 && kd.tile_blk[t] == blk) return t;
      kd.tile_slice[kd.ntiles] = digit - 1; kd.tile_blk[kd.ntiles] = blk;
      return kd.ntiles++;
    };
    auto acc_index = [&](int slot) {
      for (int g = 0; g < kd.nacc; ++g) if (kd.acc_slot[g] == slot) return g;
      kd.acc_slot[kd.nacc] = slot;
      return kd.nacc++;
    };
    int per_acc[I8_MAXACC] = {0, 0, 0, 0};
    for (auto& pr : best) {
      const int q = kd.npairs++;
      kd.pair_a[q] = tile_index(pr.a, bi);
      kd.pair_b[q] = tile_index(pr.b, bj);
      kd.pair_acc[q] = acc_index(pr.slot);
      per_acc[kd.pair_acc[q]] += 1;
    }
    // int32 head room: |digit| <= 128, so one product adds at most 2^14 per row; an item must end before 2^31 can be reached
    const int mp = *std::max_element(per_acc, per_acc + I8_MAXACC);
    kd.max_chunks = std::max(1, (int)((1ll << 31) / ((1ll << 14) * I8_KC * mp)) - 1);
    out.push_back(kd);
    std::vector<I8Prod> next;
    for (auto& pr : rem) {
      bool taken = false;
      for (auto& b : best) taken = taken || (b.a == pr.a && b.b == pr.b);
      if (!taken) next.push_back(pr);
    }
    rem.swap(next);
  }
}

// block_chunks: chunks per work item while plenty of rows remain (<= every kind's max_chunks); the last rows are cut finer.
inline I8Layout build_i8_layout(int p, int nslices, int smax, int nctas, long long nchunks, int block_chunks = 0) {
  I8Layout L;
  L.nslices = nslices; L.smax = smax; L.nblk = (p + I8_TILE - 1) / I8_TILE; L.nslots = 2 * (smax - 1);
  for (int bi = 0; bi < L.nblk; ++bi)
    for (int bj = 0; bj <= bi; ++bj) {
      i8_plan_tile(bi, bj, (int)L.tile_bi.size(), nslices, smax, L.kinds);
      L.tile_bi.push_back(bi); L.tile_bj.push_back(bj);
    }
  const int nk = (int)L.kinds.size();
  long long total = 0;
  int cap = 1 << 30;
  for (auto& kd : L.kinds) { L.kind_cost.push_back(i8_cost(kd.npairs, kd.ntiles)); total += L.kind_cost.back(); cap = std::min(cap, kd.max_chunks); }
  // default block: about eight items per CTA (every item ends with an epilogue of up to 64 K integer atomics)
  if (block_chunks <= 0) block_chunks = (int)std::max<long long>(32, nchunks * nk / (8ll * std::max(1, nctas)));
  cap = std::min(cap, block_chunks);
  std::vector<int> order(nk);
  for (int q = 0; q < nk; ++q) order[q] = q;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return L.kind_cost[x] > L.kind_cost[y]; });
  // Row blocks: `cap` chunks while more than ~2 rounds of work per CTA remain, then geometrically smaller (guided scheduling):
  // a block of c chunks costs total * c cycles over all kinds, shared by nctas CTAs.
  long long c0 = 0;
  while (c0 < nchunks) {
    const long long remaining = nchunks - c0;
    long long bc = std::min<long long>(cap, remaining);
    // guided scheduling: an item of average cost is at most 1 / nctas of the remaining work
    const long long fine = std::max<long long>(4, remaining * nk / std::max(1, nctas));
    bc = std::min(bc, fine);
    for (int q : order) L.items.push_back(I8Item{q, (int)c0, (int)bc});
    c0 += bc;
  }
  return L;
}

}  // namespace gamc
