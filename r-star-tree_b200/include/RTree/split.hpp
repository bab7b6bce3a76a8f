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
    constexpr int EXTRA = M; // index of the overflowing entry in the scratch set

    detail::overflow_set<NodeType> set(node, std::move(extra));
    auto& entries = set.entries;
    auto waste_of = [&](int i, int j)
    {
      const geometry_type gi(entries[i].first), gj(entries[j].first);
      return helper::enlarged_area(gi, gj) - helper::area(gi) - helper::area(gj);
    };
    auto overlap_of = [&](int i, int j)
    {
      const geometry_type gi(entries[i].first), gj(entries[j].first);
      return helper::intersection_area(gi, gj);
    };

    // PickSeeds: the pair that wastes most area when grouped; equal waste goes to the pair
    // with the smaller intersection.  Pairs among the old entries are examined first, then
    // the overflowing entry against each of them (the order decides exact ties).
    int seed_a = 0, seed_b = 0;
    scalar worst = std::numeric_limits<scalar>::lowest();
    auto consider = [&](int i, int j)
    {
      const scalar waste = waste_of(i, j);
      if (waste > worst || (waste == worst && overlap_of(i, j) < overlap_of(seed_a, seed_b)))
      {
        worst = waste;
        seed_a = i;
        seed_b = j;
      }
    };
    for (int i = 0; i < M - 1; ++i)
      for (int j = i + 1; j < M; ++j)
        consider(i, j);
    for (int j = 0; j < M; ++j)
      consider(EXTRA, j);

    // `left` lists the entries still in (or heading for) the old node, in slot order; its
    // first `settled` ones are assigned.  `right` is the new node.  Removing from `left`
    // moves its last element into the hole, so the final slot order is well defined.
    std::array<int, COUNT> left, right;
    int n_left = M, n_right = 0, settled = 1;
    for (int i = 0; i < M; ++i)
      left[i] = i;
    auto take_from_left = [&](int pos)
    {
      right[n_right++] = left[pos];
      left[pos] = left[n_left - 1];
      --n_left;
    };
    if (seed_a == EXTRA)
    {
      right[n_right++] = EXTRA;
      std::swap(left[0], left[seed_b]);
    }
    else
    {
      take_from_left(seed_b); // seed_a < seed_b, so slot seed_a is untouched
      std::swap(left[0], left[seed_a]);
      left[n_left++] = EXTRA;
    }
    geometry_type bound_left(entries[left[0]].first);
    geometry_type bound_right(entries[right[0]].first);

    while (settled + n_right < COUNT)
    {
      const int open = COUNT - (settled + n_right);
      if (settled + open == m)
        break; // everything still open stays left
      if (n_right + open == m)
      {
        for (int k = 0; k < open; ++k)
          take_from_left(n_left - 1);
        break;
      }
      // PickNext: the open entry with the strongest preference for one side
      int pick = settled;
      bool to_left = true;
      scalar strongest = std::numeric_limits<scalar>::lowest();
      for (int pos = settled; pos < n_left; ++pos)
      {
        auto const& g = entries[left[pos]].first;
        const scalar d_left = helper::enlarged_area(bound_left, g) - helper::area(bound_left);
        const scalar d_right = helper::enlarged_area(bound_right, g) - helper::area(bound_right);
        const scalar preference = d_left < d_right ? d_right - d_left : d_left - d_right;
        if (preference > strongest)
        {
          strongest = preference;
          pick = pos;
          to_left = d_left < d_right;
        }
      }
      if (to_left)
      {
        helper::enlarge_to(bound_left, entries[left[pick]].first);
        std::swap(left[settled], left[pick]);
        ++settled;
      }
      else
      {
        helper::enlarge_to(bound_right, entries[left[pick]].first);
        take_from_left(pick);
      }
    }

    for (int i = 0; i < n_left; ++i)
      node->insert(std::move(entries[left[i]]));
    for (int i = 0; i < n_right; ++i)
      pair->insert(std::move(entries[right[i]]));
    return pair;
  }
};

} // namespace rtree
} // namespace eh
