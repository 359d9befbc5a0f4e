// metric_plan.h -- host-side work decomposition for Kernel B (metric.cuh).
// The lower triangle of the p x p metric is cut into 32 x 32 tasks (bi >= bj); tasks are grouped into CTA
// "kinds" of at most MT_MAXT tasks touching at most MT_MAXS distinct column blocks, walking the triangle
// in 4-row super-rows so that a kind's tasks share column blocks (each resident block is loaded once per
// slab and read by several warps).
#pragma once
#include <vector>
#include <algorithm>
#include "metric.cuh"

namespace gamc {

struct MetricLayout {
  std::vector<MetricPlan> plans;
  std::vector<TaskCoord> coords;   // global task index -> (bi, bj)
  std::vector<int> kind_cost;      // per kind: DMMA tiles per k4-step on its busiest SM sub-partition
  std::vector<int> kind_nsplit;    // per kind: number of CTAs (row splits), proportional to its cost
  std::vector<int> cta_kind, cta_split;   // CTA index -> (kind, split); kinds interleaved so that they sweep the rows together
  std::vector<int> task_nsplit;    // per global task: nsplit of its kind
  int max_nsplit = 0;
};

// Warp w runs on SM sub-partition w % 4.  A diagonal task issues 10 of the 16 DMMA tiles per k4-step, so the tasks
// of a kind are placed on warps such that the four sub-partitions carry equal DMMA work (longest-processing-time rule).
inline int balance_kind(MetricPlan& pl) {
  const int nt = pl.ntasks;
  std::vector<int> order(nt);
  for (int i = 0; i < nt; ++i) order[i] = i;
  auto cost = [&](int t) { return pl.task_diag[t] ? 10 : 16; };
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
  int load[4] = {0, 0, 0, 0}, cnt[4] = {0, 0, 0, 0};
  MetricPlan out = pl;
  for (int w = 0; w < MT_MAXT; ++w) { out.task_sa[w] = out.task_sb[w] = out.task_id[w] = out.task_diag[w] = 0; }
  std::vector<std::vector<int>> bucket(4);
  for (int t : order) {
    int best = -1;
    for (int s = 0; s < 4; ++s) if (cnt[s] < MT_MAXT / 4 && (best < 0 || load[s] < load[best])) best = s;
    load[best] += cost(t); cnt[best] += 1; bucket[best].push_back(t);
  }
  // warps must be dense (warp index < ntasks): fill warp slots round by round
  int w = 0;
  std::vector<int> pos(4, 0);
  int placed = 0;
  std::vector<int> warp_task(MT_MAXT, -1);
  for (int round = 0; round < MT_MAXT / 4; ++round)
    for (int s = 0; s < 4; ++s)
      if (pos[s] < (int)bucket[s].size()) warp_task[round * 4 + s] = bucket[s][pos[s]++];
  // compact: if some slot inside [0, nt) is empty (ragged kinds), move the highest placed task down
  for (w = 0; w < nt; ++w) {
    if (warp_task[w] >= 0) continue;
    for (int v = MT_MAXT - 1; v > w; --v) if (warp_task[v] >= 0) { warp_task[w] = warp_task[v]; warp_task[v] = -1; break; }
  }
  for (w = 0; w < nt; ++w) {
    const int t = warp_task[w];
    out.task_sa[w] = pl.task_sa[t]; out.task_sb[w] = pl.task_sb[t]; out.task_id[w] = pl.task_id[t]; out.task_diag[w] = pl.task_diag[t];
    ++placed;
  }
  (void)placed;
  pl = out;
  int smsp[4] = {0, 0, 0, 0};
  for (w = 0; w < nt; ++w) smsp[w % 4] += pl.task_diag[w] ? 10 : 16;
  return std::max(std::max(smsp[0], smsp[1]), std::max(smsp[2], smsp[3]));
}

// Distribute `num_ctas` CTAs over the kinds proportionally to their cost (each kind >= 1, none more than nslabs).
inline void assign_splits(MetricLayout& L, int num_ctas, long long nslabs) {
  const int nk = (int)L.plans.size();
  L.kind_nsplit.assign(nk, 1);
  int used = nk;
  while (used < num_ctas) {
    int best = -1;
    for (int k = 0; k < nk; ++k) {
      if (L.kind_nsplit[k] >= nslabs) continue;
      if (best < 0 || (long long)L.kind_cost[k] * L.kind_nsplit[best] > (long long)L.kind_cost[best] * L.kind_nsplit[k]) best = k;
    }
    if (best < 0) break;
    L.kind_nsplit[best] += 1; ++used;
  }
  L.max_nsplit = *std::max_element(L.kind_nsplit.begin(), L.kind_nsplit.end());
  L.cta_kind.clear(); L.cta_split.clear();
  for (int s = 0; s < L.max_nsplit; ++s)
    for (int k = 0; k < nk; ++k)
      if (s < L.kind_nsplit[k]) { L.cta_kind.push_back(k); L.cta_split.push_back(s); }
  L.task_nsplit.assign(L.coords.size(), 1);
  for (int k = 0; k < nk; ++k)
    for (int t = 0; t < L.plans[k].ntasks; ++t) L.task_nsplit[L.plans[k].task_id[t]] = L.kind_nsplit[k];
}

inline MetricLayout build_metric_layout(int p) {
  MetricLayout L;
  const int nb = (p + MT_BLK - 1) / MT_BLK;
  const int SUPER = 4;
  MetricPlan cur{};
  auto flush = [&]() {
    if (cur.ntasks > 0) L.plans.push_back(cur);
    cur = MetricPlan{};
  };
  auto slot_of = [&](int blk) -> int {
    for (int s = 0; s < cur.nslots; ++s) if (cur.slot_block[s] == blk) return s;
    return -1;
  };
  for (int r0 = 0; r0 < nb; r0 += SUPER) {
    const int r1 = std::min(nb, r0 + SUPER);
    for (int bj = 0; bj < r1; ++bj) {
      for (int bi = std::max(bj, r0); bi < r1; ++bi) {
        int need = (slot_of(bi) < 0) + ((bj != bi && slot_of(bj) < 0) ? 1 : 0);
        if (cur.ntasks == MT_MAXT || cur.nslots + need > MT_MAXS) flush();
        if (slot_of(bi) < 0) cur.slot_block[cur.nslots++] = bi;
        if (slot_of(bj) < 0) cur.slot_block[cur.nslots++] = bj;
        const int t = cur.ntasks++;
        cur.task_sa[t] = slot_of(bi);
        cur.task_sb[t] = slot_of(bj);
        cur.task_diag[t] = (bi == bj);
        cur.task_id[t] = (int)L.coords.size();
        L.coords.push_back(TaskCoord{bi, bj});
      }
    }
  }
  flush();
  for (auto& pl : L.plans) L.kind_cost.push_back(balance_kind(pl));
  return L;
}

}  // namespace gamc
