// metric_i8_plan.h -- host-side work decomposition of the integer-tensor-core metric (metric_i8.cuh).
//
// G_lik = S'S with S = diag(sqrt w) X cut into `nslices` signed base-256 digit planes D_1 (most significant) .. D_k:
//   G_ij = 2^(e_i + e_j) * sum_{a + b <= smax} 2^(4 - 8 (a + b)) (D_a' D_b)_ij .
// The lower triangle of G is cut into 128 x 128 output tiles (bi >= bj); the digit products (a, b) of a tile are
// grouped into CTA "kinds".  A kind keeps at most I8_MAXACC accumulators in tensor memory (one per value of a + b:
// products with the same a + b carry the same weight and are summed exactly in int32), reads at most I8_MAXT operand
// tiles (digit plane x column block) per 128-row chunk, and issues at most I8_MAXP products per chunk.  Every kind
// gets a number of CTAs proportional to its cost per chunk, and the CTAs of a kind take every nsplit-th chunk, so all
// kinds sweep the rows together and the 126 MB L2 serves the re-reads of the digit planes.
#pragma once
#include <vector>
#include <algorithm>
#include <utility>
#include "metric_i8.cuh"

namespace gamc {

struct I8Layout {
  int nslices = 0, smax = 0, nblk = 0;
  std::vector<I8Kind> kinds;
  std::vector<int> kind_cost;        // modelled cycles per chunk
  std::vector<int> kind_nsplit;      // CTAs of each kind
  std::vector<int> cta_kind, cta_split;
  std::vector<int> tile_bi, tile_bj; // output tiles
  std::vector<int> tile_off, tile_n, tile_ctas;   // CTAs that hold a partial sum of tile t: tile_ctas[tile_off[t] .. + tile_n[t])
  int max_tiles = 0;
};

// modelled cost of one 128-row chunk: tensor pipe (128 x 128 x 128 int8 MACs per product at 8192 MAC/clk = 256 clk) against
// the L2 -> SM path (16 KB per operand tile at ~40 B/clk)
inline int i8_cost(int npairs, int ntiles) { return std::max(256 * npairs, 410 * ntiles); }

// Greedy cover of the products of one output tile.  diag: the A and B operands are digit planes of the SAME column block.
inline void i8_plan_tile(int bi, int bj, int k, int smax, std::vector<I8Kind>& out) {
  const bool diag = bi == bj;
  std::vector<std::pair<int, int>> rem;
  for (int a = 1; a <= k; ++a)
    for (int b = 1; b <= k; ++b)
      if (a + b <= smax) rem.push_back({a, b});
  while (!rem.empty()) {
    double best_score = -1.0;
    std::vector<std::pair<int, int>> best;
    for (int s0 = 2; s0 <= smax; ++s0)
      for (int a0 = 1; a0 <= k; ++a0)
        for (int a1 = a0; a1 <= k; ++a1)
          for (int b0 = 1; b0 <= k; ++b0)
            for (int b1 = b0; b1 <= k; ++b1) {
              int ntiles;
              if (diag) { const int lo = std::min(a0, b0), hi = std::max(a1, b1); ntiles = 0; for (int d = lo; d <= hi; ++d) ntiles += ((d >= a0 && d <= a1) || (d >= b0 && d <= b1)); }
              else ntiles = (a1 - a0 + 1) + (b1 - b0 + 1);
              if (ntiles > I8_MAXT) continue;
              std::vector<std::pair<int, int>> cand;
              for (auto& pr : rem)
                if (pr.first >= a0 && pr.first <= a1 && pr.second >= b0 && pr.second <= b1 && pr.first + pr.second >= s0 &&
                    pr.first + pr.second < s0 + I8_MAXACC)
                  cand.push_back(pr);
              if (cand.empty() || (int)cand.size() > I8_MAXP) continue;
              const double score = (double)cand.size() / i8_cost((int)cand.size(), ntiles) + 1e-9 * cand.size();
              if (score > best_score) { best_score = score; best = cand; }
            }
    I8Kind kd{};
    kd.bi = bi; kd.bj = bj;
    auto tile_index = [&](int digit, int blk) {
      for (int t = 0; t < kd.ntiles; ++t) if (kd.tile_slice[t] == digit - 1 && kd.tile_blk[t] == blk) return t;
      kd.tile_slice[kd.ntiles] = digit - 1; kd.tile_blk[kd.ntiles] = blk;
      return kd.ntiles++;
    };
    auto acc_index = [&](int s) {
      for (int g = 0; g < kd.nacc; ++g) if (kd.acc_sum[g] == s) return g;
      kd.acc_sum[kd.nacc] = s;
      return kd.nacc++;
    };
    std::sort(best.begin(), best.end(), [](const std::pair<int, int>& x, const std::pair<int, int>& y) {
      return x.first + x.second != y.first + y.second ? x.first + x.second < y.first + y.second : x.first < y.first; });
    int per_acc[I8_MAXACC] = {0, 0, 0, 0};
    for (auto& pr : best) {
      const int q = kd.npairs++;
      kd.pair_a[q] = tile_index(pr.first, bi);
      kd.pair_b[q] = tile_index(pr.second, bj);
      kd.pair_acc[q] = acc_index(pr.first + pr.second);
      per_acc[kd.pair_acc[q]] += 1;
    }
    // int32 head room: |digit| <= 128, so one product adds at most 2^14 per row; flush before 2^31 can be reached
    const int mp = *std::max_element(per_acc, per_acc + I8_MAXACC);
    kd.flush_chunks = std::max(1, (int)((1ll << 31) / ((1ll << 14) * I8_KC * mp)) - 1);
    out.push_back(kd);
    std::vector<std::pair<int, int>> next;
    for (auto& pr : rem) if (std::find(best.begin(), best.end(), pr) == best.end()) next.push_back(pr);
    rem.swap(next);
  }
}

inline I8Layout build_i8_layout(int p, int nslices, int smax, int ncta_target, long long nchunks) {
  I8Layout L;
  L.nslices = nslices; L.smax = smax; L.nblk = (p + I8_TILE - 1) / I8_TILE;
  std::vector<int> kind_tile;
  for (int bi = 0; bi < L.nblk; ++bi)
    for (int bj = 0; bj <= bi; ++bj) {
      const size_t before = L.kinds.size();
      i8_plan_tile(bi, bj, nslices, smax, L.kinds);
      for (size_t q = before; q < L.kinds.size(); ++q) kind_tile.push_back((int)L.tile_bi.size());
      L.tile_bi.push_back(bi); L.tile_bj.push_back(bj);
    }
  const int nk = (int)L.kinds.size();
  long long total = 0;
  for (auto& kd : L.kinds) { L.kind_cost.push_back(i8_cost(kd.npairs, kd.ntiles)); total += L.kind_cost.back(); L.max_tiles = std::max(L.max_tiles, kd.ntiles); }
  // CTAs per kind proportional to cost (largest-remainder rule), at least one, no more than there are chunks
  const int budget = std::max(ncta_target, nk);
  L.kind_nsplit.assign(nk, 1);
  int used = nk;
  std::vector<double> want(nk);
  for (int q = 0; q < nk; ++q) want[q] = (double)budget * L.kind_cost[q] / (double)total;
  for (int q = 0; q < nk; ++q) { const int extra = std::max(0, (int)want[q] - 1); L.kind_nsplit[q] += extra; used += extra; }
  while (used < budget) {
    int best = 0; double gap = -1e300;
    for (int q = 0; q < nk; ++q) { const double g = want[q] - L.kind_nsplit[q]; if (g > gap) { gap = g; best = q; } }
    L.kind_nsplit[best] += 1; used += 1;
  }
  for (int q = 0; q < nk; ++q) L.kind_nsplit[q] = (int)std::max<long long>(1, std::min<long long>(L.kind_nsplit[q], nchunks));
  // CTA order: splits of all kinds interleaved, so that co-resident CTAs work on neighbouring chunks
  int maxs = *std::max_element(L.kind_nsplit.begin(), L.kind_nsplit.end());
  for (int s = 0; s < maxs; ++s)
    for (int q = 0; q < nk; ++q)
      if (s < L.kind_nsplit[q]) { L.cta_kind.push_back(q); L.cta_split.push_back(s); }
  const int nt = (int)L.tile_bi.size();
  L.tile_off.assign(nt, 0); L.tile_n.assign(nt, 0);
  for (int t = 0; t < nt; ++t) {
    L.tile_off[t] = (int)L.tile_ctas.size();
    for (int c = 0; c < (int)L.cta_kind.size(); ++c)
      if (kind_tile[L.cta_kind[c]] == t) L.tile_ctas.push_back(c);
    L.tile_n[t] = (int)L.tile_ctas.size() - L.tile_off[t];
  }
  return L;
}

}  // namespace gamc
