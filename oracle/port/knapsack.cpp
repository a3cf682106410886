// TEST INFRASTRUCTURE -- restatement of AMReX's cost-weighted box -> rank assignment used by
// PeleLM::getFuncCountDM (Source/PeleLM.cpp:5529-5623):
//   knapsack()              amrex/Src/Base/AMReX_DistributionMapping.cpp:576-716
//   KnapSackDoIt / ProcessorMap  :718-870 (buckets sorted heaviest first, then mapped onto the
//                           rank order given by LeastUsedCPUs :191-235; ties keep rank order)
// Pinned against the reference's own knapsack() (oracle/ref_knapsack.cpp) in tests/test_knapsack.py.
// Uses the same standard-library primitives (std::sort, std::priority_queue, std::lower_bound,
// std::rotate) so that tie-breaking is the libstdc++ one the reference gets.
#include "oracle.h"

#include <algorithm>
#include <queue>
#include <utility>
#include <vector>

namespace {
struct Ball {
  int id;
  long w;
  bool operator<(const Ball& o) const { return w > o.w; }
};
struct Bin {
  std::vector<Ball>* v;
  long w;
  bool operator<(const Bin& o) const { return w > o.w; }
  int size() const { return static_cast<int>(v->size()); }
  void push_back(const Ball& b)
  {
    w += b.w;
    v->push_back(b);
  }
};
} // namespace

extern "C" void orc_knapsack(const long* wgts, int nboxes, int nprocs, double max_eff, int do_full, int nmax,
                             int* proc_of_box, double* efficiency)
{
  if (nboxes <= nprocs || nprocs < 2) { // RoundRobinProcessorMap, sort = false ordering
    for (int i = 0; i < nboxes; ++i) proc_of_box[i] = i % (nprocs > 0 ? nprocs : 1);
    if (efficiency) *efficiency = 1.0;
    return;
  }
  std::vector<Ball> lb;
  lb.reserve(nboxes);
  for (int i = 0; i < nboxes; ++i) lb.push_back(Ball{i, wgts[i]});
  std::sort(lb.begin(), lb.end());

  std::vector<std::vector<Ball>> store(nprocs);
  std::priority_queue<Bin> q;
  for (int i = 0; i < nprocs; ++i) q.push(Bin{&store[i], 0});
  std::vector<Bin> bins;
  bins.reserve(nprocs);
  for (int i = 0; i < nboxes; ++i) {
    if (!q.empty()) {
      Bin b = q.top();
      q.pop();
      b.push_back(lb[i]);
      if (b.size() < nmax) q.push(b);
      else bins.push_back(b);
    } else {
      bins[i % nprocs].push_back(lb[i]);
    }
  }
  double max_weight = 0, sum_weight = 0;
  for (auto const& b : bins) {
    sum_weight += static_cast<double>(b.w);
    max_weight = std::max(static_cast<double>(b.w), max_weight);
  }
  while (!q.empty()) {
    Bin b = q.top();
    q.pop();
    if (b.size() > 0) {
      sum_weight += static_cast<double>(b.w);
      max_weight = std::max(static_cast<double>(b.w), max_weight);
      bins.push_back(b);
    }
  }
  double eff = sum_weight / (nprocs * max_weight);
  std::sort(bins.begin(), bins.end());

  if (eff < max_eff && do_full && bins.size() > 1 && bins.begin()->size() > 1) {
    bool swapped = true;
    while (swapped && eff < max_eff && bins.begin()->size() > 1) {
      swapped = false;
      auto top = bins.begin();
      auto bottom = bins.end() - 1;
      const long w_top = top->w, w_bottom = bottom->w;
      for (auto b1 = top->v->begin(); b1 != top->v->end() && !swapped; ++b1) {
        for (auto b2 = bottom->v->begin(); b2 != bottom->v->end(); ++b2) {
          const long dw = b1->w - b2->w;
          const long w_top_new = w_top - dw, w_bottom_new = w_bottom + dw;
          if (w_top_new < w_top && w_bottom_new < w_top) {
            std::swap(*b1, *b2);
            top->w -= dw;
            bottom->w += dw;
            if (top + 1 == bottom) {
              if (*bottom < *top) std::swap(*top, *bottom);
            } else {
              auto it = std::lower_bound(top + 1, bottom, *bottom);
              std::rotate(it, bottom, bottom + 1);
              it = std::lower_bound(top + 1, bottom + 1, *top);
              std::rotate(top, top + 1, it);
            }
            max_weight = static_cast<double>(bins.begin()->w);
            eff = sum_weight / (nprocs * max_weight);
            swapped = true;
            break;
          }
        }
      }
    }
  }
  // KnapSackDoIt: buckets by total weight, heaviest first (stable), bucket i -> rank ord[i] = i
  std::vector<std::pair<long, int>> order;
  for (int i = 0; i < static_cast<int>(bins.size()); ++i) order.push_back({bins[i].w, i});
  std::stable_sort(order.begin(), order.end(),
                   [](const std::pair<long, int>& a, const std::pair<long, int>& b) { return a.first > b.first; });
  for (int i = 0; i < static_cast<int>(order.size()); ++i)
    for (auto const& ball : *bins[order[i].second].v) proc_of_box[ball.id] = i;
  if (efficiency) *efficiency = eff;
}
