// RTree/split.hpp -- host side (header-only C++17) of rtree-b200.
//
// Node-overflow split policies, selected through Config::split_algorithm exactly as
// in the reference (include/RTree/rtree.hpp:249-256).  A policy is any type with
//
//   template <class NodeType>
//   static NodeType* split(NodeType* node, typename NodeType::value_type extra,
//                          NodeType* pair);
//
// which distributes node's MAX_ENTRIES entries plus `extra` over `node` and the
// empty `pair`, both ending with >= MIN_ENTRIES entries, and returns `pair`.
//
//   RStarSplit     -- Beckmann et al. 1990, in the variant the reference ships
//                     (include/RTree/rstar_split.hpp:17-170): the split axis is the
//                     one holding the single distribution of least margin sum.
//   QuadraticSplit -- Guttman 1984 (include/RTree/quadratic_split.hpp:16-240).
//
// Both are written from scratch on a small scratch array of (bound, original slot)
// records rather than by shuffling the node's entries in place.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <limits>
#include <utility>

#include "geometry.hpp"
#include "node.hpp"

namespace eh
{
namespace rtree
{

namespace detail
{
/// Move the M+1 entries of an overflowing node into a scratch array (the extra
/// entry last), leaving the node empty.
template <typename NodeType>
struct overflow_set
{
  using value_type = typename NodeType::value_type;
  constexpr static int COUNT = static_cast<int>(NodeType::MAX_ENTRIES) + 1;
  static_vector<value_type, NodeType::MAX_ENTRIES + 1> entries;

  overflow_set(NodeType* node, value_type&& extra)
  {
    for (size_type i = 0; i < node->size(); ++i)
      entries.emplace_back(std::move(node->at(i)));
    entries.emplace_back(std::move(extra));
    node->clear();
  }
};
} // namespace detail

struct RStarSplit
{
  template <typename NodeType>
  static NodeType* split(NodeType* node, typename NodeType::value_type extra, NodeType* pair)
  {
    using geometry_type = typename NodeType::geometry_type;
    using traits = geometry_traits<geometry_type>;
    using scalar = typename traits::scalar_type;
    constexpr int M = static_cast<int>(NodeType::MAX_ENTRIES);
    constexpr int m = static_cast<int>(NodeType::MIN_ENTRIES);
    constexpr int COUNT = M + 1;

    detail::overflow_set<NodeType> set(node, std::move(extra));
    auto& entries = set.entries;

    // order[] is a permutation of the entries; sorting permutes indices, not entries
    std::array<int, COUNT> order;
    for (int i = 0; i < COUNT; ++i)
      order[i] = i;
    auto sort_along = [&](int axis)
    {
      std::sort(order.begin(), order.end(),
                [&](int a, int b)
                {
                  const scalar amin = helper::min_point(entries[a].first, axis);
                  const scalar bmin = helper::min_point(entries[b].first, axis);
                  if (amin < bmin)
                    return true;
                  if (amin == bmin)
                    return helper::max_point(entries[a].first, axis)
                           < helper::max_point(entries[b].first, axis);
                  return false;
                });
    };

    // prefix[i] covers order[0..i], suffix[i] covers order[i..M]
    static_vector<geometry_type, NodeType::MAX_ENTRIES + 1> prefix, suffix;
    auto sweep = [&]()
    {
      prefix.clear();
      suffix.clear();
      for (int i = 0; i < COUNT; ++i)
      {
        prefix.emplace_back(entries[order[i]].first);
        suffix.emplace_back(entries[order[i]].first);
      }
      for (int i = 1; i < COUNT; ++i)
        helper::enlarge_to(prefix[i], prefix[i - 1]);
      for (int i = COUNT - 2; i >= 0; --i)
        helper::enlarge_to(suffix[i], suffix[i + 1]);
    };

    // 1. split axis: the axis that owns the one distribution with least margin sum,
    //    ties broken by the smaller overlap
    int best_axis = -1;
    int last_sorted_axis = -1;
    scalar best_margin = std::numeric_limits<scalar>::max();
    scalar best_overlap = std::numeric_limits<scalar>::max();
    for (int axis = 0; axis < traits::DIM; ++axis)
    {
      sort_along(axis);
      last_sorted_axis = axis;
      sweep();
      for (int k = m; k <= M - m + 1; ++k)
      {
        // groups: order[0..k) and order[k..M]
        const scalar margin_sum = helper::margin(prefix[k - 1]) + helper::margin(suffix[k]);
        const scalar overlap = helper::intersection_area(prefix[k - 1], suffix[k]);
        if (best_axis == -1 || margin_sum < best_margin
            || (margin_sum == best_margin && overlap < best_overlap))
        {
          best_axis = axis;
          best_margin = margin_sum;
          best_overlap = overlap;
        }
      }
    }

    // 2. split index on that axis: least overlap, ties by least area sum
    if (best_axis != last_sorted_axis)
    {
      sort_along(best_axis);
      sweep();
    }
    int best_k = -1;
    scalar least_overlap = std::numeric_limits<scalar>::max();
    scalar least_area = std::numeric_limits<scalar>::max();
    for (int k = m; k <= M - m + 1; ++k)
    {
      const scalar area_sum = helper::area(prefix[k - 1]) + helper::area(suffix[k]);
      const scalar overlap = helper::intersection_area(prefix[k - 1], suffix[k]);
      if (best_k == -1 || overlap < least_overlap
          || (overlap == least_overlap && area_sum < least_area))
      {
        best_k = k;
        least_overlap = overlap;
        least_area = area_sum;
      }
    }

    for (int i = 0; i < best_k; ++i)
      node->insert(std::move(entries[order[i]]));
    for (int i = best_k; i < COUNT; ++i)
      pair->insert(std::move(entries[order[i]]));
    return pair;
  }
};

struct QuadraticSplit
{
  template <typename NodeType>
  static NodeType* split(NodeType* node, typename NodeType::value_type extra, NodeType* pair)
  {
    using geometry_type = typename NodeType::geometry_type;
    using traits = geometry_traits<geometry_type>;
    using scalar = typename traits::scalar_type;
    constexpr int M = static_cast<int>(NodeType::MAX_ENTRIES);
    constexpr int m = static_cast<int>(NodeType::MIN_ENTRIES);
    constexpr int COUNT = M + 1;

    detail::overflow_set<NodeType> set(node, std::move(extra));
    auto& entries = set.entries;

    // PickSeeds: the pair that wastes most area when put together; ties go to the
    // pair with the smaller intersection
    int seed1 = 0, seed2 = 1;
    {
      scalar worst = std::numeric_limits<scalar>::lowest();
      scalar worst_inter = 0;
      bool have = false;
      for (int i = 0; i < COUNT - 1; ++i)
      {
        const geometry_type gi(entries[i].first);
        const scalar ai = helper::area(gi);
        for (int j = i + 1; j < COUNT; ++j)
        {
          const geometry_type gj(entries[j].first);
          const scalar waste = helper::enlarged_area(gi, gj) - ai - helper::area(gj);
          if (!have || waste > worst)
          {
            have = true;
            worst = waste;
            worst_inter = helper::intersection_area(gi, gj);
            seed1 = i;
            seed2 = j;
          }
          else if (waste == worst)
          {
            const scalar inter = helper::intersection_area(gi, gj);
            if (inter < worst_inter)
            {
              worst_inter = inter;
              seed1 = i;
              seed2 = j;
            }
          }
        }
      }
    }

    std::array<int, COUNT> group; // -1 unassigned, 0 -> node, 1 -> pair
    group.fill(-1);
    group[seed1] = 0;
    group[seed2] = 1;
    geometry_type bound0(entries[seed1].first);
    geometry_type bound1(entries[seed2].first);
    int count0 = 1, count1 = 1, left = COUNT - 2;

    while (left > 0)
    {
      // a group that needs every remaining entry to reach m takes them all
      if (count0 + left == m)
      {
        for (int i = 0; i < COUNT; ++i)
          if (group[i] == -1)
            group[i] = 0;
        break;
      }
      if (count1 + left == m)
      {
        for (int i = 0; i < COUNT; ++i)
          if (group[i] == -1)
            group[i] = 1;
        break;
      }
      // PickNext: the entry with the strongest preference for one group
      int pick = -1, pick_to = 0;
      scalar strongest = std::numeric_limits<scalar>::lowest();
      for (int i = 0; i < COUNT; ++i)
      {
        if (group[i] != -1)
          continue;
        const scalar d0 = helper::enlarged_area(bound0, entries[i].first) - helper::area(bound0);
        const scalar d1 = helper::enlarged_area(bound1, entries[i].first) - helper::area(bound1);
        const scalar diff = d0 < d1 ? d1 - d0 : d0 - d1;
        if (pick == -1 || diff > strongest)
        {
          pick = i;
          strongest = diff;
          pick_to = d0 < d1 ? 0 : 1;
        }
      }
      group[pick] = pick_to;
      if (pick_to == 0)
      {
        helper::enlarge_to(bound0, entries[pick].first);
        ++count0;
      }
      else
      {
        helper::enlarge_to(bound1, entries[pick].first);
        ++count1;
      }
      --left;
    }

    for (int i = 0; i < COUNT; ++i)
    {
      if (group[i] == 1)
        pair->insert(std::move(entries[i]));
      else
        node->insert(std::move(entries[i]));
    }
    return pair;
  }
};

} // namespace rtree
} // namespace eh
