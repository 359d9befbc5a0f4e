// Host-only invariants of the planner of the integer-tensor-core metric (gamcsampler.jl_b200/csrc/metric_i8_plan.h).
// Built and run by tests/test_host_logic.py::test_i8_planner_invariants (no GPU needed: only host code runs).
// With an argument "print" it lists the plans of p = 256 and p = 512 (6 digit planes, digit products a + b <= 7).
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
  const int nblk = (p + I8_TILE - 1) / I8_TILE, ng = smax - 1;
  std::map<std::tuple<int, int, int, int>, int> seen;   // (bi, bj, a, b) -> count
  if (L.nslots != 2 * ng) { printf("nslots\n"); ++bad; }
  for (size_t q = 0; q < L.kinds.size(); ++q) {
    const I8Kind& kd = L.kinds[q];
    if (kd.bi < kd.bj || kd.bi >= nblk || kd.ntiles < 1 || kd.ntiles > I8_MAXT || kd.npairs < 1 || kd.npairs > I8_MAXP || kd.nacc < 1 ||
        kd.nacc > I8_MAXACC || kd.max_chunks < 1 || kd.tile < 0 || kd.tile >= (int)L.tile_bi.size() || L.tile_bi[kd.tile] != kd.bi ||
        L.tile_bj[kd.tile] != kd.bj) { printf("p=%d k=%d: kind %zu shape\n", p, k, q); ++bad; continue; }
    if (2 * kd.ntiles > I8_RING) { printf("p=%d k=%d: kind %zu cannot double-buffer\n", p, k, q); ++bad; }
    const bool diag = kd.bi == kd.bj;
    int per_acc[I8_MAXACC] = {0, 0, 0, 0};
    std::set<int> slots;
    for (int g = 0; g < kd.nacc; ++g) if (kd.acc_slot[g] < 0 || kd.acc_slot[g] >= L.nslots || !slots.insert(kd.acc_slot[g]).second) { printf("p=%d: slot table\n", p); ++bad; }
    for (int t = 0; t < kd.npairs; ++t) {
      const int ta = kd.pair_a[t], tb = kd.pair_b[t], g = kd.pair_acc[t];
      if (ta < 0 || ta >= kd.ntiles || tb < 0 || tb >= kd.ntiles || g < 0 || g >= kd.nacc) { printf("p=%d: pair index\n", p); ++bad; continue; }
      if (kd.tile_blk[ta] != kd.bi || kd.tile_blk[tb] != kd.bj) { printf("p=%d: operand block\n", p); ++bad; }
      const int a = kd.tile_slice[ta] + 1, b = kd.tile_slice[tb] + 1;
      const int want_slot = (diag && a < b) ? ng + a + b - 2 : a + b - 2;
      if (a < 1 || a > k || b < 1 || b > k || a + b > smax || (diag && a > b) || kd.acc_slot[g] != want_slot) { printf("p=%d: product (%d,%d) slot %d\n", p, a, b, kd.acc_slot[g]); ++bad; }
      seen[{kd.bi, kd.bj, a, b}] += 1;
      per_acc[g] += 1;
    }
    for (int g = 0; g < kd.nacc; ++g)
      if ((long long)per_acc[g] * kd.max_chunks * I8_KC * (1ll << 14) >= (1ll << 31)) { printf("p=%d: int32 head room of kind %zu\n", p, q); ++bad; }
    if (print) {
      printf("kind %2zu tile (%d,%d) tiles %d pairs %2d acc %d max %4d cost %5d :", q, kd.bi, kd.bj, kd.ntiles, kd.npairs, kd.nacc, kd.max_chunks, L.kind_cost[q]);
      for (int t = 0; t < kd.npairs; ++t) printf(" (%d,%d)", kd.tile_slice[kd.pair_a[t]] + 1, kd.tile_slice[kd.pair_b[t]] + 1);
      printf("\n");
    }
  }
  for (int bi = 0; bi < nblk; ++bi)
    for (int bj = 0; bj <= bi; ++bj)
      for (int a = 1; a <= k; ++a) for (int b = 1; b <= k; ++b) {
        const auto it = seen.find({bi, bj, a, b});
        const int c = it == seen.end() ? 0 : it->second;
        const int want = (a + b <= smax && !(bi == bj && a > b)) ? 1 : 0;
        if (c != want) { printf("p=%d k=%d smax=%d: product (%d,%d) of tile (%d,%d) covered %d times\n", p, k, smax, a, b, bi, bj, c); ++bad; }
      }
  // items: every kind covers every chunk exactly once, in increasing row order, never longer than the int32 head room allows
  std::vector<long long> next(L.kinds.size(), 0);
  long long last_c0 = 0;
  for (const I8Item& im : L.items) {
    if (im.kind < 0 || im.kind >= (int)L.kinds.size() || im.nchunks < 1 || im.nchunks > L.kinds[im.kind].max_chunks || im.chunk0 != next[im.kind] ||
        im.chunk0 < last_c0) { printf("p=%d: item (%d, %d, %d)\n", p, im.kind, im.chunk0, im.nchunks); ++bad; break; }
    next[im.kind] += im.nchunks; last_c0 = im.chunk0;
  }
  for (size_t q = 0; q < L.kinds.size(); ++q) if (next[q] != nchunks) { printf("p=%d: kind %zu covers %lld of %lld chunks\n", p, q, next[q], nchunks); ++bad; }
  if (print) {
    long long total = 0;
    for (int c : L.kind_cost) total += c;
    printf("p=%d k=%d smax=%d: %zu kinds, %zu items, modelled %lld cycles per chunk = %.3f ms at 1.9 GHz on %d SMs\n", p, k, smax, L.kinds.size(), L.items.size(),
           total, (double)total * nchunks / ncta / 1.9e6, ncta);
  }
  return bad;
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "print")) { check(256, 6, 7, 148, 8192, true); check(512, 6, 7, 148, 131072, true); check(256, 5, 6, 148, 8192, true); return 0; }
  int bad = 0;
  for (int p : {1, 4, 64, 128, 129, 256, 300, 384, 479, 512})
    for (int k = 2; k <= I8_MAXK; ++k)
      for (int smax : {k, k + 1, k + 2, 2 * k})
        for (long long nchunks : {1ll, 32ll, 8192ll})
          bad += check(p, k, smax, 148, nchunks, false);
  printf(bad ? "FAILED %d\n" : "OK\n", bad);
  return bad ? 1 : 0;
}
