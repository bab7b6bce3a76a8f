// tests/cpp/api_surface.cpp -- the public API of eh::rtree::RTree as the reference's README documents it
// (README.md:89-347), driven with a USER-DEFINED geometry type bound through the geometry_traits customisation
// point (README.md:155-192), a user Config with QuadraticSplit (README.md:128-153), the tri-state search filter
// and the aborting functor (README.md:194-235), iterators and node pointers (README.md:237-262), flatten
// (README.md:264-335), rebound / rebalance (README.md:342-347), erase, copy and move.
// It prints a digest of everything it sees.  tests/test_reference_unit_tests.py compiles this file twice --
// against the reference's headers and against r-star-tree_b200/include -- and the two outputs must be equal
// byte for byte.  Lines under RTB_ONLY_OURS exercise members the reference declares but cannot instantiate
// (SURVEY.md appendix B6 / B9); they are printed with a "ours:" prefix and compared with values derived from
// the portable calls instead.
#include <RTree.hpp>

#include <algorithm>
#include <cstdint>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <utility>
#include <vector>

// ---- a user geometry: 2-D box / point over float arrays ------------------------------------------
struct Pt
{
  float c[2];
};
struct Rect
{
  float lo[2], hi[2];
  Rect() : lo { 0, 0 }, hi { 0, 0 } {}
  Rect(float x0, float y0, float x1, float y1) : lo { x0, y0 }, hi { x1, y1 } {}
  Rect(Pt const& p) : lo { p.c[0], p.c[1] }, hi { p.c[0], p.c[1] } {} // key -> degenerate box (implicit, as README says)
};

namespace eh
{
namespace rtree
{
template <>
struct geometry_traits<Rect>
{
  constexpr static int DIM = 2;
  using scalar_type = float;
  static scalar_type min_point(Rect const& g, int axis) { return g.lo[axis]; }
  static scalar_type max_point(Rect const& g, int axis) { return g.hi[axis]; }
  static void set_min_point(Rect& g, int axis, scalar_type v) { g.lo[axis] = v; }
  static void set_max_point(Rect& g, int axis, scalar_type v) { g.hi[axis] = v; }
};
template <>
struct geometry_traits<Pt>
{
  constexpr static int DIM = 2;
  using scalar_type = float;
  static scalar_type min_point(Pt const& g, int axis) { return g.c[axis]; }
  static scalar_type max_point(Pt const& g, int axis) { return g.c[axis]; }
};
} // namespace rtree
} // namespace eh

namespace er = eh::rtree;

struct WideQuadratic
{
  constexpr static er::size_type MIN_ENTRIES = 3;
  constexpr static er::size_type MAX_ENTRIES = 10;
  constexpr static er::size_type REINSERT_COUNT = 4;
  using split_algorithm = er::QuadraticSplit;
};

// order-sensitive digest of a sequence of words
struct Digest
{
  uint64_t h = 1469598103934665603ull;
  void word(uint64_t v)
  {
    h ^= v;
    h *= 1099511628211ull;
  }
  void real(float v)
  {
    uint32_t bits;
    std::memcpy(&bits, &v, 4);
    word(bits);
  }
};

template <typename Tree>
uint64_t digest_flat(Tree const& tree)
{
  auto flat = tree.flatten();
  Digest d;
  d.word(flat.leaf_level);
  d.word(flat.root);
  for (auto const& n : flat.nodes)
    d.word(n.offset), d.word(n.size), d.word(n.parent);
  for (auto const& b : flat.children_bound)
    for (int a = 0; a < 2; ++a)
      d.real(er::geometry_traits<typename Tree::geometry_type>::min_point(b, a)),
          d.real(er::geometry_traits<typename Tree::geometry_type>::max_point(b, a));
  for (auto c : flat.children)
    d.word(c);
  for (auto v : flat.data)
    d.word((uint64_t)v);
  std::printf("  flatten: leaf_level %u nodes %zu entries %zu data %zu digest %016llx\n", (unsigned)flat.leaf_level,
              flat.nodes.size(), flat.children.size(), flat.data.size(), (unsigned long long)d.h);
  return d.h;
}

template <typename Tree>
void walk_nodes(Tree& tree)
{
  // README "Directly accessing node pointer": root(), leaf_level(), children of nodes and leaves
  Digest d;
  size_t internal = 0, leaves = 0, values = 0;
  struct Item
  {
    typename Tree::node_type* node;
    int level;
  };
  std::vector<Item> todo { { tree.root(), 0 } };
  while (!todo.empty())
  {
    Item it = todo.back();
    todo.pop_back();
    if (it.level == tree.leaf_level())
    {
      auto* leaf = it.node->as_leaf();
      ++leaves;
      for (typename Tree::leaf_type::value_type const& child : *leaf)
        d.word((uint64_t)child.second), ++values;
    }
    else
    {
      ++internal;
      for (typename Tree::node_type::value_type const& child : *it.node)
      {
        d.real(child.first.lo[0]), d.real(child.first.hi[1]);
        todo.push_back({ child.second->as_node(), it.level + 1 });
      }
    }
  }
  std::printf("  node walk: internal %zu leaves %zu values %zu digest %016llx\n", internal, leaves, values,
              (unsigned long long)d.h);
  // level iterators
  for (int level = 0; level < tree.leaf_level(); ++level)
  {
    size_t count = 0, entries = 0;
    for (auto ni = tree.node_begin(level); ni != tree.node_end(level); ++ni)
      ++count, entries += ni->size();
    std::printf("  level %d: %zu nodes, %zu entries\n", level, count, entries);
  }
  size_t count = 0, entries = 0;
  for (auto li = tree.leaf_begin(); li != tree.leaf_end(); ++li)
    ++count, entries += li->size();
  std::printf("  leaf level %d: %zu leaves, %zu values\n", tree.leaf_level(), count, entries);
}

template <typename Tree>
void queries(Tree& tree, std::mt19937& rng)
{
  std::uniform_real_distribution<float> centre(-20.f, 20.f), half(0.2f, 4.f);
  Digest d;
  size_t hits = 0, filter_calls = 0;
  for (int q = 0; q < 200; ++q)
  {
    const float cx = centre(rng), cy = centre(rng), h = half(rng);
    const Rect box(cx - h, cy - h, cx + h, cy + h);
    auto filter = [&](Rect const& b) -> int
    {
      ++filter_calls;
      for (int a = 0; a < 2; ++a)
        if (b.hi[a] < box.lo[a] || b.lo[a] > box.hi[a])
          return 0;
      return 1;
    };
    std::vector<int> found;
    tree.search(filter,
                [&](typename Tree::value_type const& v) -> bool
                {
                  bool inside = true;
                  for (int a = 0; a < 2; ++a)
                    inside = inside && !(v.first.c[a] < box.lo[a]) && !(v.first.c[a] > box.hi[a]);
                  if (inside)
                    found.push_back(v.second);
                  return false;
                });
    hits += found.size();
    for (int v : found)
      d.word((uint64_t)v); // in delivery order: the traversal order is part of the behaviour
    d.word(0xFFFFull);
  }
  std::printf("  200 range queries: %zu hits, %zu filter calls, digest %016llx\n", hits, filter_calls,
              (unsigned long long)d.h);

  // the functor aborts at its third value; the filter aborts the whole search with -1
  int seen = 0;
  tree.search([](Rect const&) -> int { return 1; }, [&](typename Tree::value_type const&) -> bool { return ++seen == 3; });
  int visited = 0, delivered = 0;
  tree.search([&](Rect const&) -> int { return ++visited >= 4 ? -1 : 1; },
              [&](typename Tree::value_type const&) -> bool
              {
                ++delivered;
                return false;
              });
  std::printf("  functor abort after %d values; filter abort after %d bounds with %d values delivered\n", seen, visited,
              delivered);
  // const overload
  Tree const& frozen = tree;
  size_t all = 0;
  frozen.search([](Rect const&) -> int { return 1; },
                [&](typename Tree::value_type const&) -> bool
                {
                  ++all;
                  return false;
                });
  std::printf("  const search over everything: %zu values, size() %zu\n", all, (size_t)frozen.size());
}

template <typename Tree>
void scenario(const char* title, uint32_t seed)
{
  std::printf("%s\n", title);
  std::mt19937 rng(seed);
  std::normal_distribution<float> radius(0.f, 8.f);
  std::uniform_real_distribution<float> angle(0.f, 6.2831853f);
  Tree tree;
  std::printf("  fresh: size %zu leaf_level %d begin==end %d\n", (size_t)tree.size(), tree.leaf_level(),
              (int)(tree.begin() == tree.end()));
  for (int i = 0; i < 1500; ++i)
  {
    const float r = radius(rng), t = angle(rng);
    const Pt p { { r * std::cos(t), r * std::sin(t) } };
    if (i % 2)
      tree.insert({ p, i });
    else
      tree.emplace(p, i);
  }
  std::printf("  built: size %zu leaf_level %d\n", (size_t)tree.size(), tree.leaf_level());
  digest_flat(tree);
  walk_nodes(tree);
  queries(tree, rng);

  // iteration order = flatten data order
  {
    Digest d;
    for (typename Tree::value_type const& v : tree)
      d.word((uint64_t)v.second);
    std::printf("  iteration digest %016llx\n", (unsigned long long)d.h);
  }
  // moving objects: change keys through the iterators, rebound every value, then rebalance
  for (auto it = tree.begin(); it != tree.end(); ++it)
  {
    it->first.c[0] += 0.25f * (float)(it->second % 7);
    it->first.c[1] -= 0.125f * (float)(it->second % 5);
  }
  for (auto it = tree.begin(); it != tree.end(); ++it)
    tree.rebound(it);
  std::printf("  after rebound:\n");
  digest_flat(tree);
  tree.rebalance();
  std::printf("  after rebalance: size %zu leaf_level %d\n", (size_t)tree.size(), tree.leaf_level());
  digest_flat(tree);

  // erase every value whose id is divisible by 3 (one at a time: an erase invalidates iterators)
  size_t erased = 0;
  for (int id = 0; id < 1500; id += 3)
  {
    for (auto it = tree.begin(); it != tree.end(); ++it)
    {
      if (it->second == id)
      {
        tree.erase(it);
        ++erased;
        break;
      }
    }
  }
  std::printf("  erased %zu: size %zu leaf_level %d\n", erased, (size_t)tree.size(), tree.leaf_level());
  digest_flat(tree);

  // copy, move, assignment
  Tree copy(tree);
  Tree moved(std::move(copy));
  Tree assigned;
  assigned = moved;
  std::printf("  copy / move / assign: sizes %zu %zu\n", (size_t)moved.size(), (size_t)assigned.size());
  const uint64_t a = digest_flat(moved), b = digest_flat(assigned), c = digest_flat(tree);
  std::printf("  copies equal the original: %d\n", (int)(a == c && b == c));
  auto flat_moved = assigned.flatten_move();
  std::printf("  flatten_move: %zu nodes, %zu data\n", flat_moved.nodes.size(), flat_moved.data.size());

#ifdef RTB_ONLY_OURS
  // members the reference declares but cannot instantiate
  size_t by_iterator = 0;
  tree.search_iterator([](Rect const&) -> int { return 1; },
                       [&](typename Tree::iterator it) -> bool
                       {
                         by_iterator += it->second >= 0 ? 1 : 0;
                         return false;
                       });
  Tree const& frozen = tree;
  size_t const_nodes = 0;
  for (int level = 0; level < frozen.leaf_level(); ++level)
    for (auto ni = frozen.node_cbegin(level); ni != frozen.node_cend(level); ++ni)
      ++const_nodes;
  std::printf("ours: search_iterator saw %zu values (size %zu); node_cbegin walked %zu internal nodes\n", by_iterator,
              (size_t)tree.size(), const_nodes);
#endif
  tree.clear();
  std::printf("  cleared: size %zu leaf_level %d\n", (size_t)tree.size(), tree.leaf_level());
}

int main()
{
  scenario<er::RTree<Rect, Pt, int>>("user geometry, DefaultConfig (4 / 8 / 3, RStarSplit)", 7u);
  scenario<er::RTree<Rect, Pt, int, WideQuadratic>>("user geometry, user Config (3 / 10 / 4, QuadraticSplit)", 8u);
  return 0;
}
