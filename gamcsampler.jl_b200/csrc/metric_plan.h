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
};

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
  return L;
}

}  // namespace gamc
