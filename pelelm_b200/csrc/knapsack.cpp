// Host-side cost-weighted box -> rank assignment for the chemistry load balance.
// Behaviour to reproduce: PeleLM::getFuncCountDM (Source/PeleLM.cpp:5529-5623) ->
// DistributionMapping::KnapSackProcessorMap (amrex/Src/Base/AMReX_DistributionMapping.cpp:846-870) ->
// knapsack() (:576-716): longest-processing-time greedy into the lightest bin, then pairwise swaps
// between the heaviest and the lightest bin while efficiency < max_efficiency; finally buckets are
// handed to ranks heaviest-first (KnapSackDoIt :718-844, rank order = identity when all ranks hold
// equal FAB bytes). Every rank runs this on the all-gathered cost vector and gets the same answer.
#include <algorithm>
#include <cstdint>
#include <functional>
#include <numeric>
#include <queue>
#include <vector>

#include "../../include/plm_chem.h"

namespace {

struct Item {
  int box;
  int64_t w;
};
struct Bucket {
  std::vector<Item> items;
  int64_t w = 0;
};
// ordering used for both the initial sort of boxes and for buckets: heavier first
struct Heavier {
  bool operator()(const Item& a, const Item& b) const { return a.w > b.w; }
};
struct BucketRef {
  int id;
  int64_t w;
  bool operator<(const BucketRef& o) const { return w > o.w; }  // priority_queue top = lightest
};

}  // namespace

// ord[i] = the rank that takes the i-th bucket (heaviest first) or, with at most one box per rank, box i (round robin):
// identity, or the ranks by ascending FAB bytes (DistributionMapping::LeastUsedCPUs, the reference's sort = true default)
static int knapsack_core(const int64_t* wgts, int nboxes, int nranks, double max_efficiency, int do_full_knapsack, int nmax,
                         const int* ord, int* rank_of_box, double* efficiency)
{
  if (!wgts || !rank_of_box || nboxes < 0 || nranks < 1) return PLM_ERR_ARG;
  if (ord) {  // must be a permutation of the ranks
    std::vector<char> seen(nranks, 0);
    for (int i = 0; i < nranks; ++i) {
      if (ord[i] < 0 || ord[i] >= nranks || seen[ord[i]]) return PLM_ERR_ARG;
      seen[ord[i]] = 1;
    }
  }
  if (nboxes <= nranks || nranks < 2) {  // RoundRobinProcessorMap -> RoundRobinDoIt (:365-447): box i -> ord[i % nranks]
    for (int i = 0; i < nboxes; ++i) rank_of_box[i] = ord ? ord[i % nranks] : i % nranks;
    if (efficiency) *efficiency = 1.0;
    return PLM_OK;
  }
  std::vector<Item> balls(nboxes);
  for (int i = 0; i < nboxes; ++i) balls[i] = Item{i, wgts[i]};
  std::sort(balls.begin(), balls.end(), [](const Item& a, const Item& b) { return a.w > b.w; });

  std::vector<Bucket> bk(nranks);
  std::priority_queue<BucketRef> pq;
  for (int i = 0; i < nranks; ++i) pq.push(BucketRef{i, 0});
  std::vector<int> order;  // bucket ids in the order they were closed / drained
  for (int i = 0; i < nboxes; ++i) {
    if (!pq.empty()) {
      BucketRef t = pq.top();
      pq.pop();
      bk[t.id].items.push_back(balls[i]);
      bk[t.id].w += balls[i].w;
      t.w = bk[t.id].w;
      if (int(bk[t.id].items.size()) < nmax) pq.push(t);
      else order.push_back(t.id);
    } else {
      Bucket& b = bk[order[i % nranks]];
      b.items.push_back(balls[i]);
      b.w += balls[i].w;
    }
  }
  double sum_w = 0, max_w = 0;
  for (int id : order) {
    sum_w += double(bk[id].w);
    max_w = std::max(max_w, double(bk[id].w));
  }
  while (!pq.empty()) {
    const BucketRef t = pq.top();
    pq.pop();
    if (!bk[t.id].items.empty()) {
      sum_w += double(bk[t.id].w);
      max_w = std::max(max_w, double(bk[t.id].w));
      order.push_back(t.id);
    }
  }
  double eff = sum_w / (nranks * max_w);
  // sorted view, heaviest first (same comparator and std::sort as the reference: weight() > rhs.weight())
  std::vector<BucketRef> v;
  for (int id : order) v.push_back(BucketRef{id, bk[id].w});
  std::sort(v.begin(), v.end());

  if (eff < max_efficiency && do_full_knapsack && v.size() > 1 && bk[v.front().id].items.size() > 1) {
    for (;;) {
      if (!(eff < max_efficiency && bk[v.front().id].items.size() > 1)) break;
      auto top = v.begin();
      auto bot = v.end() - 1;
      Bucket& T = bk[top->id];
      Bucket& B = bk[bot->id];
      bool swapped = false;
      for (size_t a = 0; a < T.items.size() && !swapped; ++a)
        for (size_t b = 0; b < B.items.size(); ++b) {
          const int64_t dw = T.items[a].w - B.items[b].w;
          const int64_t wt = top->w - dw, wb = bot->w + dw;
          if (wt < top->w && wb < top->w) {
            std::swap(T.items[a], B.items[b]);
            T.w -= dw;
            B.w += dw;
            top->w = T.w;
            bot->w = B.w;
            if (top + 1 == bot) {
              if (*bot < *top) std::swap(*top, *bot);
            } else {
              auto it = std::lower_bound(top + 1, bot, *bot);
              std::rotate(it, bot, bot + 1);
              it = std::lower_bound(top + 1, bot + 1, *top);
              std::rotate(top, top + 1, it);
            }
            max_w = double(v.front().w);
            eff = sum_w / (nranks * max_w);
            swapped = true;
            break;
          }
        }
      if (!swapped) break;
    }
  }
  // heaviest bucket -> ord[0], next -> ord[1], ... (stable on ties; KnapSackDoIt :760-815)
  std::vector<int> idx(v.size());
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return v[a].w > v[b].w; });
  for (size_t i = 0; i < idx.size(); ++i)
    for (const Item& it : bk[v[idx[i]].id].items) rank_of_box[it.box] = ord ? ord[i] : int(i);
  if (efficiency) *efficiency = eff;
  return PLM_OK;
}

extern "C" int plm_knapsack(const int64_t* wgts, int nboxes, int nranks, double max_efficiency, int do_full_knapsack,
                            int nmax, int* rank_of_box, double* efficiency)
{
  return knapsack_core(wgts, nboxes, nranks, max_efficiency, do_full_knapsack, nmax, nullptr, rank_of_box, efficiency);
}

extern "C" int plm_knapsack_ordered(const int64_t* wgts, int nboxes, int nranks, double max_efficiency, int do_full_knapsack,
                                    int nmax, const int* rank_order, int* rank_of_box, double* efficiency)
{
  if (!rank_order) return PLM_ERR_ARG;
  return knapsack_core(wgts, nboxes, nranks, max_efficiency, do_full_knapsack, nmax, rank_order, rank_of_box, efficiency);
}

// DistributionMapping::LeastUsedCPUs (AMReX_DistributionMapping.cpp:191-235): ranks by ascending kilobytes held in FABs,
// rank order kept among equals (std::stable_sort with LIpairLT)
extern "C" int plm_least_used_ranks(const int64_t* kbytes, int nranks, int* rank_order)
{
  if (!kbytes || !rank_order || nranks < 1) return PLM_ERR_ARG;
  std::iota(rank_order, rank_order + nranks, 0);
  std::stable_sort(rank_order, rank_order + nranks, [&](int a, int b) { return kbytes[a] < kbytes[b]; });
  return PLM_OK;
}
