// Host-only invariants of the planner of the integer-tensor-core metric (gamcsampler.jl_b200/csrc/metric_i8_plan.h).
// Built and run by tests/test_host_logic.py::test_i8_planner_invariants (no GPU needed: only host code runs).
// With an argument "print" it lists the plan of p = 256, 6 digit planes, digit products a + b <= 7.
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <tuple>
#include "metric_i8_plan.h"
using namespace gamc;

static int check(int p, int k, int smax, int ncta, long long nchunks, bool print) {
  int bad = 0;
  const I8Layout L = build_i8_layout(p, k, smax, ncta, nchunks);
  const int nblk = (p + I8_TILE - 1) / I8_TILE;
  std::map<std::tuple<int, int, int, int>, int> seen;   // (bi, bj, a, b) -> count
  for (size_t q = 0; q < L.kinds.size(); ++q) {
    const I8Kind& kd = L.kinds[q];
    if (kd.bi < kd.bj || kd.bi >= nblk || kd.ntiles < 1 || kd.ntiles > I8_MAXT || kd.npairs < 1 || kd.npairs > I8_MAXP || kd.nacc < 1 ||
        kd.nacc > I8_MAXACC || kd.flush_chunks < 1) { printf("p=%d k=%d: kind %zu shape\n", p, k, q); ++bad; continue; }
    if (I8_SMEM_TILES / kd.ntiles < 2) { printf("p=%d k=%d: kind %zu cannot double-buffer\n", p, k, q); ++bad; }
    int per_acc[I8_MAXACC] = {0, 0, 0, 0};
    for (int t = 0; t < kd.npairs; ++t) {
      const int ta = kd.pair_a[t], tb = kd.pair_b[t], g = kd.pair_acc[t];
      if (ta < 0 || ta >= kd.ntiles || tb < 0 || tb >= kd.ntiles || g < 0 || g >= kd.nacc) { printf("p=%d: pair index\n", p); ++bad; continue; }
      if (kd.tile_blk[ta] != kd.bi || kd.tile_blk[tb] != kd.bj) { printf("p=%d: operand block\n", p); ++bad; }
      const int a = kd.tile_slice[ta] + 1, b = kd.tile_slice[tb] + 1;
      if (a < 1 || a > k || b < 1 || b > k || a + b > smax || kd.acc_sum[g] != a + b) { printf("p=%d: product (%d,%d) acc %d\n", p, a, b, kd.acc_sum[g]); ++bad; }
      seen[{kd.bi, kd.bj, a, b}] += 1;
      per_acc[g] += 1;
    }
    for (int g = 0; g < kd.nacc; ++g)
      if ((long long)per_acc[g] * kd.flush_chunks * I8_KC * (1ll << 14) >= (1ll << 31)) { printf("p=%d: int32 head room of kind %zu\n", p, q); ++bad; }
    if (print) {
      printf("kind %2zu tile (%d,%d) tiles %d pairs %2d acc %d flush %4d cost %5d ctas %3d :", q, kd.bi, kd.bj, kd.ntiles, kd.npairs, kd.nacc, kd.flush_chunks,
             L.kind_cost[q], L.kind_nsplit[q]);
      for (int t = 0; t < kd.npairs; ++t) printf(" (%d,%d)", kd.tile_slice[kd.pair_a[t]] + 1, kd.tile_slice[kd.pair_b[t]] + 1);
      printf("\n");
    }
  }
  int want = 0;
  for (int a = 1; a <= k; ++a) for (int b = 1; b <= k; ++b) want += (a + b <= smax);
  for (int bi = 0; bi < nblk; ++bi)
    for (int bj = 0; bj <= bi; ++bj) {
      int have = 0;
      for (int a = 1; a <= k; ++a) for (int b = 1; b <= k; ++b) {
        const auto it = seen.find({bi, bj, a, b});
        const int c = it == seen.end() ? 0 : it->second;
        if (c != (a + b <= smax ? 1 : 0)) { printf("p=%d k=%d smax=%d: product (%d,%d) of tile (%d,%d) covered %d times\n", p, k, smax, a, b, bi, bj, c); ++bad; }
        have += c;
      }
      if (have != want) ++bad;
    }
  // CTA tables
  std::set<std::pair<int, int>> ctas;
  for (size_t c = 0; c < L.cta_kind.size(); ++c) {
    const int q = L.cta_kind[c], s = L.cta_split[c];
    if (q < 0 || q >= (int)L.kinds.size() || s < 0 || s >= L.kind_nsplit[q] || !ctas.insert({q, s}).second) { printf("p=%d: CTA table\n", p); ++bad; }
  }
  size_t total = 0;
  for (int n : L.kind_nsplit) total += n;
  if (total != L.cta_kind.size()) { printf("p=%d: CTA count\n", p); ++bad; }
  std::set<int> listed;
  for (size_t t = 0; t < L.tile_bi.size(); ++t)
    for (int u = 0; u < L.tile_n[t]; ++u) {
      const int c = L.tile_ctas[L.tile_off[t] + u];
      const I8Kind& kd = L.kinds[L.cta_kind[c]];
      if (kd.bi != L.tile_bi[t] || kd.bj != L.tile_bj[t] || !listed.insert(c).second) { printf("p=%d: tile list\n", p); ++bad; }
    }
  if (listed.size() != L.cta_kind.size()) { printf("p=%d: tile lists do not cover the CTAs\n", p); ++bad; }
  if (print) printf("p=%d k=%d smax=%d: %zu kinds, %zu CTAs\n", p, k, smax, L.kinds.size(), L.cta_kind.size());
  return bad;
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "print")) { check(256, 6, 7, 148, 8192, true); check(512, 6, 7, 148, 131072, true); return 0; }
  int bad = 0;
  for (int p : {1, 4, 64, 128, 129, 256, 300, 384, 479, 512})
    for (int k = 2; k <= I8_MAXK; ++k)
      for (int smax : {k, k + 1, k + 2, 2 * k})
        for (long long nchunks : {1ll, 32ll, 8192ll})
          bad += check(p, k, smax, 148, nchunks, false);
  printf(bad ? "FAILED %d\n" : "OK\n", bad);
  return bad ? 1 : 0;
}
