// examples/range_and_knn.cpp -- the reference's sample (example/sample/sample.cpp: build a tree, search
// a range) moved onto the B200 engine, in the reference's own language.  Host side: this repository's
// header-only eh::rtree (same API as the reference); device side: the C ABI of include/rtree_b200.h.
//
//   g++ -std=c++17 -O2 -I r-star-tree_b200/include -I include examples/range_and_knn.cpp
//       -L r-star-tree_b200/csrc -lrtree_b200 -Wl,-rpath,$PWD/r-star-tree_b200/csrc -o range_and_knn
//
// It checks every GPU answer against RTree::search() on the host tree and exits non-zero on any
// difference, so it doubles as a test (tests/test_gpu_cpp_example.py).
#include <RTree.hpp>
#include <rtree_b200.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

using point = eh::rtree::point_t<float, 3>;
using box = eh::rtree::aabb_t<point>;
using tree_t = eh::rtree::RTree<box, point, std::uint32_t>; // DefaultConfig: 4 / 8 / 3, RStarSplit

static void die(const char* what)
{
  std::fprintf(stderr, "%s: %s\n", what, rt_last_error());
  std::exit(2);
}

int main(int argc, char** argv)
{
  const unsigned n_points = argc > 1 ? std::atoi(argv[1]) : 200000;
  const unsigned n_queries = argc > 2 ? std::atoi(argv[2]) : 50000;
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> u(0.f, 1.f);

  // ---- build on the CPU, exactly as with the reference -------------------------------------
  tree_t tree;
  for (std::uint32_t i = 0; i < n_points; ++i)
    tree.insert({ point(u(rng), u(rng), u(rng)), i });

  // ---- hand the tree to the GPU --------------------------------------------------------------
  auto image = tree.flatten_device(eh::rtree::id_source::mapped_value);
  rt_tree_desc desc = image.desc();
  rt_tree* gpu = nullptr;
  if (rt_upload(&desc, 0, &gpu) != RT_OK)
    die("rt_upload");

  // ---- a batch of range queries ----------------------------------------------------------------
  std::vector<float> boxes(std::size_t(n_queries) * 6);
  for (unsigned q = 0; q < n_queries; ++q)
    for (int d = 0; d < 3; ++d)
    {
      const float c = u(rng), h = 0.5f * 0.04f * u(rng);
      boxes[q * 6 + d] = c - h;
      boxes[q * 6 + 3 + d] = c + h;
    }
  rt_hits hits;
  if (rt_query_boxes(gpu, boxes.data(), n_queries, RT_POINT_IN_BOX, 0, &hits) != RT_OK)
    die("rt_query_boxes");

  // the same queries through RTree::search(filter, functor) on the host tree
  unsigned long long mismatches = 0, total = 0;
  for (unsigned q = 0; q < n_queries; ++q)
  {
    const box query(point(boxes[q * 6], boxes[q * 6 + 1], boxes[q * 6 + 2]),
                    point(boxes[q * 6 + 3], boxes[q * 6 + 4], boxes[q * 6 + 5]));
    std::vector<std::uint32_t> want;
    tree.search(
        [&](box const& b)
        {
          for (int d = 0; d < 3; ++d)
            if (b.max_[d] < query.min_[d] || b.min_[d] > query.max_[d])
              return 0;
          return 1;
        },
        [&](tree_t::value_type const& v)
        {
          bool inside = true;
          for (int d = 0; d < 3; ++d)
            inside = inside && !(v.first[d] < query.min_[d]) && !(v.first[d] > query.max_[d]);
          if (inside)
            want.push_back(v.second);
          return false;
        });
    std::sort(want.begin(), want.end());
    const std::uint64_t begin = hits.offsets[q], end = hits.offsets[q + 1];
    total += want.size();
    if (end - begin != want.size() || !std::equal(want.begin(), want.end(), hits.ids + begin))
      ++mismatches;
  }
  std::printf("range queries: %u queries, %llu hits, %llu mismatching hit lists\n", n_queries, total, mismatches);

  // ---- the 4 nearest neighbours of the first 1000 query centres -------------------------------
  const unsigned n_knn = std::min(n_queries, 1000u), k = 4;
  std::vector<float> centres(std::size_t(n_knn) * 3);
  for (unsigned q = 0; q < n_knn; ++q)
    for (int d = 0; d < 3; ++d)
      centres[q * 3 + d] = 0.5f * (boxes[q * 6 + d] + boxes[q * 6 + 3 + d]);
  rt_knn_result knn;
  if (rt_query_knn(gpu, centres.data(), n_knn, k, 0, &knn) != RT_OK)
    die("rt_query_knn");
  unsigned long long knn_bad = 0;
  for (unsigned q = 0; q < n_knn; ++q)
  {
    // brute force over the host tree's values (begin() .. end() iterate them all)
    std::vector<std::pair<float, std::uint32_t>> all;
    for (auto it = tree.begin(); it != tree.end(); ++it)
    {
      float d2 = 0.f;
      for (int d = 0; d < 3; ++d)
      {
        const float e = it->first[d] - centres[q * 3 + d];
        const float sq = e * e;
        d2 = d2 + sq;
      }
      all.emplace_back(d2, it->second);
    }
    std::partial_sort(all.begin(), all.begin() + k, all.end());
    for (unsigned j = 0; j < k; ++j)
      if (knn.ids[q * k + j] != all[j].second)
        ++knn_bad;
    if (q == 20) // brute force is O(N) per query: a few are enough here
      break;
  }
  std::printf("kNN: %u queries answered, %llu ids differing from brute force on the first 21\n", n_knn, knn_bad);

  rt_free(gpu);
  return (mismatches == 0 && knn_bad == 0) ? 0 : 1;
}
