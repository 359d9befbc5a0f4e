// Host-only invariants of the metric planner (gamcsampler.jl_b200/csrc/metric_plan.h) for every supported p.
// Built and run by tests/test_host_logic.py::test_metric_planner_invariants (no GPU needed: only host code runs).
#include <cstdio>
#include <set>
#include <utility>
#include "metric_plan.h"
using namespace gamc;

int main() {
  int bad = 0;
  for (int p = 1; p <= 512; ++p) {
    MetricLayout L = build_metric_layout(p);
    const int nblk = (p + MT_BLK - 1) / MT_BLK;
    const long long nslabs = 29128;
    assign_splits(L, 148, nslabs);
    std::set<std::pair<int, int>> seen;
    int ndiag = 0;
    std::set<int> ids;
    for (const MetricPlan& k : L.plans) {
      if (k.nslots < 1 || k.nslots > MT_MAXS || k.ntasks < 1 || k.ntasks > MT_MAXT) { printf("p=%d: kind shape %d/%d\n", p, k.nslots, k.ntasks); ++bad; }
      for (int t = 0; t < k.ntasks; ++t) {
        if (k.task_sa[t] < 0 || k.task_sa[t] >= k.nslots || k.task_sb[t] < 0 || k.task_sb[t] >= k.nslots) { printf("p=%d: slot index\n", p); ++bad; continue; }
        const int bi = k.slot_block[k.task_sa[t]], bj = k.slot_block[k.task_sb[t]];
        if (bi < bj || bi >= nblk) { printf("p=%d: task (%d,%d) not in the lower triangle\n", p, bi, bj); ++bad; }
        if (!seen.insert({bi, bj}).second) { printf("p=%d: task (%d,%d) twice\n", p, bi, bj); ++bad; }
        if ((bi == bj) != (k.task_diag[t] != 0)) { printf("p=%d: diag flag of (%d,%d)\n", p, bi, bj); ++bad; }
        ndiag += k.task_diag[t] != 0;
        const int id = k.task_id[t];
        if (id < 0 || id >= (int)L.coords.size() || L.coords[id].bi != bi || L.coords[id].bj != bj || !ids.insert(id).second) { printf("p=%d: task id %d\n", p, id); ++bad; }
      }
    }
    if ((int)seen.size() != nblk * (nblk + 1) / 2) { printf("p=%d: %zu of %d block pairs covered\n", p, seen.size(), nblk * (nblk + 1) / 2); ++bad; }
    if (ndiag != nblk) { printf("p=%d: %d diagonal tasks for %d blocks\n", p, ndiag, nblk); ++bad; }
    // CTA table: every (kind, split) exactly once, split < nsplit of the kind, grid within the SM budget
    std::set<std::pair<int, int>> ctas;
    for (size_t c = 0; c < L.cta_kind.size(); ++c) {
      const int k = L.cta_kind[c], s = L.cta_split[c];
      if (k < 0 || k >= (int)L.plans.size() || s < 0 || s >= L.kind_nsplit[k] || !ctas.insert({k, s}).second) { printf("p=%d: cta %zu\n", p, c); ++bad; }
    }
    size_t want = 0;
    for (int n : L.kind_nsplit) { want += n; if (n < 1 || n > L.max_nsplit) { printf("p=%d: nsplit %d\n", p, n); ++bad; } }
    if (ctas.size() != want || L.cta_kind.size() > 148 + L.plans.size()) { printf("p=%d: grid %zu (want %zu)\n", p, L.cta_kind.size(), want); ++bad; }
  }
  printf("%s\n", bad ? "FAIL" : "OK");
  return bad ? 1 : 0;
}
