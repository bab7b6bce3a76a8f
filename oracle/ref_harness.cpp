// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A C-ABI shim around the UNMODIFIED reference library (header-only C++17,
// included from where it lies: -I/root/reference/include).  It builds trees with
// the reference's own RTree::insert(), runs the reference's own
// RTree::search(geometry_filter, data_functor) (include/RTree/rtree.hpp:1141-1146)
// and RTree::flatten() (include/RTree/rtree.hpp:966-974), and hands the results
// out as plain arrays so that tests/ and bench.py's cpu_baseline / --impl reference
// leg can compare against / time the real thing.  The compiled library goes to
// oracle/_ref/ (git-ignored, travels to the GPU box).  Nothing in the product path
// (r-star-tree_b200/, include/) may load it.
//
// Functors used:
//   a2  box filter   : closed-interval overlap, the idiom of
//                      example/sample/sample.cpp:48-54 and test/rtree.cpp:403
//   a3  leaf test    : geometry_traits<aabb>::is_inside(aabb, point)
//                      (include/RTree/aabb.hpp:206-210, 1-D: :131-134)
//   box keys         : INTERSECTS = the same closed overlap on the key box,
//                      CONTAINS = test/rtree.cpp:390 written per axis
//   a12 ray filter   : slab test defined in oracle/slab.h (shared with oracle.c),
//                      expressed through search() with a stateful filter, which is
//                      legal because functors are held by reference
//                      (include/RTree/rtree.hpp:1145).
#include <RTree.hpp>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "slab.h"
#include "knn.h"

namespace er = eh::rtree;

namespace
{

enum
{
  MODE_POINT_IN_BOX = 0,
  MODE_INTERSECTS = 1,
  MODE_CONTAINS = 2,
  // OR-ed in: the node filter, and the key test of box keys under MODE_INTERSECTS, is the reference's
  // own geometry_traits<aabb>::is_overlap (include/RTree/aabb.hpp:225-236: strict for N-D boxes;
  // :147-159: closed in 1-D; is_overlap(aabb, point) = is_inside, :219-223)
  MODE_STRICT = 0x100
};
enum
{
  RAY_ALL = 0,
  RAY_CLOSEST = 1,
  RAY_ANY = 2
};

struct QuadraticConfig16
{
  constexpr static er::size_type MIN_ENTRIES = 8;
  constexpr static er::size_type MAX_ENTRIES = 16;
  constexpr static er::size_type REINSERT_COUNT = 6;
  using split_algorithm = er::QuadraticSplit;
};
struct QuadraticConfig8
{
  constexpr static er::size_type MIN_ENTRIES = 4;
  constexpr static er::size_type MAX_ENTRIES = 8;
  constexpr static er::size_type REINSERT_COUNT = 3;
  using split_algorithm = er::QuadraticSplit;
};

struct ITree
{
  virtual ~ITree() {}
  virtual void insert(const void* keys, uint64_t n, uint32_t first_id) = 0;
  virtual int64_t set_keys(const void* keys, uint64_t n) = 0;
  virtual uint64_t erase_ids(const uint32_t* ids, uint64_t n) = 0;
  virtual uint32_t leaf_level() const = 0;
  virtual uint64_t size() const = 0;
  virtual void rebalance() = 0;
  virtual void flatten() = 0;
  virtual void flat_sizes(uint64_t* out) const = 0;
  virtual void flat_copy(uint32_t* nodes, void* bounds, uint32_t* children,
                         uint32_t* data) const = 0;
  virtual int64_t search_boxes(const void* q, uint64_t n, int mode, int threads,
                               uint64_t* offsets, uint32_t** ids,
                               double* seconds) const = 0;
  virtual int64_t search_rays(const float* rays, uint64_t n, int mode,
                              int threads, uint64_t* offsets, uint32_t** ids,
                              float* t_out, uint32_t* id_out, uint8_t* any_out,
                              double* seconds) const = 0;
  virtual int64_t search_knn(const float* points, uint64_t n, uint32_t k, int threads,
                             uint32_t* ids_out, float* d2_out, double* seconds) const = 0;
};

// ---- geometry adapters: how a C array row becomes a reference key / box ------

template <typename S, unsigned D>
struct NdAdapter
{
  using point = er::point_t<S, D>;
  using box = er::aabb_t<point>;
  static point make_point(const S* p)
  {
    point r;
    for (unsigned i = 0; i < D; ++i)
      r[i] = p[i];
    return r;
  }
  static box make_box(const S* p)
  {
    return box(make_point(p), make_point(p + D));
  }
  static S lo(box const& b, unsigned d) { return b.min_[d]; }
  static S hi(box const& b, unsigned d) { return b.max_[d]; }
};
template <typename S>
struct OneDAdapter
{
  using point = S;
  using box = er::aabb_t<S>;
  static point make_point(const S* p) { return p[0]; }
  static box make_box(const S* p) { return box(p[0], p[1]); }
  static S lo(box const& b, unsigned) { return b.min_; }
  static S hi(box const& b, unsigned) { return b.max_; }
};

template <typename A, typename S, unsigned D, bool KEY_IS_BOX, typename Config>
struct TreeImpl : ITree
{
  using box = typename A::box;
  using point = typename A::point;
  using key = typename std::conditional<KEY_IS_BOX, box, point>::type;
  using tree_t = er::RTree<box, key, uint32_t, Config>;
  using value_t = typename tree_t::value_type;
  using flat_t = typename tree_t::flatten_result_t;

  tree_t tree;
  flat_t flat;
  bool have_flat = false;

  static key make_key(const S* p, std::true_type) { return A::make_box(p); }
  static key make_key(const S* p, std::false_type) { return A::make_point(p); }
  constexpr static unsigned KEY_SCALARS = KEY_IS_BOX ? 2 * D : D;

  void insert(const void* keys, uint64_t n, uint32_t first_id) override
  {
    const S* p = static_cast<const S*>(keys);
    for (uint64_t i = 0; i < n; ++i)
    {
      tree.insert(value_t(make_key(p + i * KEY_SCALARS,
                                   std::integral_constant<bool, KEY_IS_BOX> {}),
                          first_id + (uint32_t)i));
    }
    have_flat = false;
  }
  // "Dealing with moving objects" (reference README.md:342-347): keys are changed in place through
  // the iterators (iteration order = flatten_result_t::data order) and RTree::rebound(iterator)
  // (include/RTree/rtree.hpp:497-500) refits the ancestors; no topology change
  int64_t set_keys(const void* keys, uint64_t n) override
  {
    const S* p = static_cast<const S*>(keys);
    uint64_t i = 0;
    for (auto it = tree.begin(); it != tree.end(); ++it, ++i)
    {
      if (i >= n)
        return -1;
      it->first = make_key(p + i * KEY_SCALARS, std::integral_constant<bool, KEY_IS_BOX> {});
    }
    if (i != n)
      return -1;
    for (auto it = tree.begin(); it != tree.end(); ++it)
      tree.rebound(it);
    have_flat = false;
    return (int64_t)i;
  }
  // the reference's own RTree::erase(iterator) (include/RTree/rtree.hpp:372-457) on the first value
  // whose mapped id matches, one id after the other in the order given
  uint64_t erase_ids(const uint32_t* ids, uint64_t n) override
  {
    uint64_t erased = 0;
    for (uint64_t i = 0; i < n; ++i)
    {
      for (auto it = tree.begin(); it != tree.end(); ++it)
      {
        if (it->second == ids[i])
        {
          tree.erase(it);
          ++erased;
          break;
        }
      }
    }
    have_flat = false;
    return erased;
  }
  uint32_t leaf_level() const override { return (uint32_t)tree.leaf_level(); }
  uint64_t size() const override { return tree.size(); }
  void rebalance() override
  {
    tree.rebalance();
    have_flat = false;
  }
  void flatten() override
  {
    flat = tree.flatten();
    have_flat = true;
  }
  void flat_sizes(uint64_t* out) const override
  {
    out[0] = flat.nodes.size();
    out[1] = flat.children.size();
    out[2] = flat.data.size();
    out[3] = flat.leaf_level;
    out[4] = flat.root;
  }
  void flat_copy(uint32_t* nodes, void* bounds, uint32_t* children,
                 uint32_t* data) const override
  {
    for (size_t i = 0; i < flat.nodes.size(); ++i)
    {
      nodes[3 * i + 0] = flat.nodes[i].offset;
      nodes[3 * i + 1] = flat.nodes[i].size;
      nodes[3 * i + 2] = flat.nodes[i].parent;
    }
    S* b = static_cast<S*>(bounds);
    for (size_t i = 0; i < flat.children_bound.size(); ++i)
    {
      for (unsigned d = 0; d < D; ++d)
      {
        b[i * 2 * D + d] = A::lo(flat.children_bound[i], d);
        b[i * 2 * D + D + d] = A::hi(flat.children_bound[i], d);
      }
    }
    std::copy(flat.children.begin(), flat.children.end(), children);
    std::copy(flat.data.begin(), flat.data.end(), data);
  }

  // leaf predicates -----------------------------------------------------------
  static bool leaf_hit(box const& q, point const& k, int mode)
  {
    if (mode & MODE_STRICT)
      return er::geometry_traits<box>::is_overlap(q, k); // = is_inside for a point (aabb.hpp:219-223)
    // a3: the reference's own predicate
    return er::geometry_traits<box>::is_inside(q, k);
  }
  static bool leaf_hit(box const& q, box const& k, int mode)
  {
    if (mode == (MODE_INTERSECTS | MODE_STRICT))
      return er::geometry_traits<box>::is_overlap(q, k);
    if ((mode & ~MODE_STRICT) == MODE_CONTAINS)
    {
      for (unsigned d = 0; d < D; ++d)
      {
        // per-axis "not greater", the form of the reference's N-D less_equal
        // (include/RTree/aabb.hpp:249-259)
        if (A::lo(q, d) > A::lo(k, d) || A::hi(k, d) > A::hi(q, d))
          return false;
      }
      return true;
    }
    for (unsigned d = 0; d < D; ++d)
    {
      if (A::hi(k, d) < A::lo(q, d) || A::lo(k, d) > A::hi(q, d))
        return false;
    }
    return true;
  }

  template <typename Emit>
  void one_box_query(box const& q, int mode, Emit&& emit) const
  {
    const bool strict = (mode & MODE_STRICT) != 0;
    auto filter = [&q, strict](box const& b) -> int
    {
      if (strict)
        return er::geometry_traits<box>::is_overlap(b, q) ? 1 : 0;
      for (unsigned d = 0; d < D; ++d)
      {
        if (A::hi(b, d) < A::lo(q, d) || A::lo(b, d) > A::hi(q, d))
          return 0;
      }
      return 1;
    };
    auto functor = [&](value_t const& v) -> bool
    {
      if (leaf_hit(q, v.first, mode))
        emit(v.second);
      return false;
    };
    tree.search(filter, functor);
  }

  int64_t search_boxes(const void* qv, uint64_t n, int mode, int threads,
                       uint64_t* offsets, uint32_t** ids_out,
                       double* seconds) const override
  {
    const S* qs = static_cast<const S*>(qv);
    const bool want_ids = offsets != nullptr;
    std::vector<std::vector<uint32_t>> per_query;
    if (want_ids)
      per_query.resize(n);
    int64_t total = 0;
#ifdef _OPENMP
    if (threads <= 0)
      threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1024) num_threads(threads) reduction(+ : total)
    for (int64_t i = 0; i < (int64_t)n; ++i)
    {
      box q = A::make_box(qs + i * 2 * D);
      if (want_ids)
      {
        std::vector<uint32_t>& out = per_query[i];
        one_box_query(q, mode, [&](uint32_t id) { out.push_back(id); });
        total += (int64_t)out.size();
      }
      else
      {
        int64_t c = 0;
        one_box_query(q, mode, [&](uint32_t) { ++c; });
        total += c;
      }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds)
      *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (want_ids)
    {
      uint32_t* ids = (uint32_t*)std::malloc(sizeof(uint32_t)
                                             * (size_t)std::max<int64_t>(total, 1));
      uint64_t pos = 0;
      for (uint64_t i = 0; i < n; ++i)
      {
        offsets[i] = pos;
        std::sort(per_query[i].begin(), per_query[i].end());
        std::copy(per_query[i].begin(), per_query[i].end(), ids + pos);
        pos += per_query[i].size();
      }
      offsets[n] = pos;
      *ids_out = ids;
    }
    return total;
  }

  // rays ------------------------------------------------------------------------
  template <unsigned DD = D>
  typename std::enable_if<(DD == 3) && std::is_same<S, float>::value, int64_t>::type
  rays_impl(const float* rays, uint64_t n, int mode, int threads,
            uint64_t* offsets, uint32_t** ids_out, float* t_out,
            uint32_t* id_out, uint8_t* any_out, double* seconds) const
  {
    std::vector<std::vector<uint32_t>> per_query;
    if (mode == RAY_ALL && offsets)
      per_query.resize(n);
    int64_t total = 0;
#ifdef _OPENMP
    if (threads <= 0)
      threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    auto t0c = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1024) num_threads(threads) reduction(+ : total)
    for (int64_t i = 0; i < (int64_t)n; ++i)
    {
      const float* r = rays + i * 7;
      auto key_lohi = [](key const& k, float* lo, float* hi)
      {
        box kb(k);
        for (unsigned d = 0; d < 3; ++d)
        {
          lo[d] = A::lo(kb, d);
          hi[d] = A::hi(kb, d);
        }
      };
      if (mode == RAY_ALL)
      {
        int64_t c = 0;
        auto filter = [&](box const& b) -> int
        {
          float lo[3] = { A::lo(b, 0), A::lo(b, 1), A::lo(b, 2) };
          float hi[3] = { A::hi(b, 0), A::hi(b, 1), A::hi(b, 2) };
          float t;
          return rt_slab_hit(lo, hi, r, r + 3, r[6], &t) ? 1 : 0;
        };
        auto functor = [&](value_t const& v) -> bool
        {
          float lo[3], hi[3], t;
          key_lohi(v.first, lo, hi);
          if (rt_slab_hit(lo, hi, r, r + 3, r[6], &t))
          {
            ++c;
            if (!per_query.empty())
              per_query[i].push_back(v.second);
          }
          return false;
        };
        tree.search(filter, functor);
        total += c;
      }
      else if (mode == RAY_CLOSEST)
      {
        float t_best = r[6];
        uint32_t id_best = 0xFFFFFFFFu;
        bool found = false;
        auto filter = [&](box const& b) -> int
        {
          float lo[3] = { A::lo(b, 0), A::lo(b, 1), A::lo(b, 2) };
          float hi[3] = { A::hi(b, 0), A::hi(b, 1), A::hi(b, 2) };
          float t;
          if (!rt_slab_hit(lo, hi, r, r + 3, r[6], &t))
            return 0;
          return (t <= t_best) ? 1 : 0; // non-strict: ties survive
        };
        auto functor = [&](value_t const& v) -> bool
        {
          float lo[3], hi[3], t;
          key_lohi(v.first, lo, hi);
          if (rt_slab_hit(lo, hi, r, r + 3, r[6], &t))
          {
            if (!found || t < t_best || (t == t_best && v.second < id_best))
            {
              found = true;
              t_best = t;
              id_best = v.second;
            }
          }
          return false;
        };
        tree.search(filter, functor);
        if (t_out)
          t_out[i] = found ? t_best : RT_SLAB_MISS_T;
        if (id_out)
          id_out[i] = id_best;
        total += found ? 1 : 0;
      }
      else
      {
        bool found = false;
        auto filter = [&](box const& b) -> int
        {
          float lo[3] = { A::lo(b, 0), A::lo(b, 1), A::lo(b, 2) };
          float hi[3] = { A::hi(b, 0), A::hi(b, 1), A::hi(b, 2) };
          float t;
          return rt_slab_hit(lo, hi, r, r + 3, r[6], &t) ? 1 : 0;
        };
        auto functor = [&](value_t const& v) -> bool
        {
          float lo[3], hi[3], t;
          key_lohi(v.first, lo, hi);
          if (rt_slab_hit(lo, hi, r, r + 3, r[6], &t))
          {
            found = true;
            return true; // abort: include/RTree/rtree.hpp:994-997
          }
          return false;
        };
        tree.search(filter, functor);
        if (any_out)
          any_out[i] = found ? 1 : 0;
        total += found ? 1 : 0;
      }
    }
    auto t1c = std::chrono::steady_clock::now();
    if (seconds)
      *seconds = std::chrono::duration<double>(t1c - t0c).count();
    if (mode == RAY_ALL && offsets)
    {
      uint32_t* ids = (uint32_t*)std::malloc(sizeof(uint32_t)
                                             * (size_t)std::max<int64_t>(total, 1));
      uint64_t pos = 0;
      for (uint64_t i = 0; i < n; ++i)
      {
        offsets[i] = pos;
        std::sort(per_query[i].begin(), per_query[i].end());
        std::copy(per_query[i].begin(), per_query[i].end(), ids + pos);
        pos += per_query[i].size();
      }
      offsets[n] = pos;
      *ids_out = ids;
    }
    return total;
  }
  template <unsigned DD = D>
  typename std::enable_if<!((DD == 3) && std::is_same<S, float>::value), int64_t>::type
  rays_impl(const float*, uint64_t, int, int, uint64_t*, uint32_t**, float*,
            uint32_t*, uint8_t*, double*) const
  {
    return -1;
  }
  int64_t search_rays(const float* rays, uint64_t n, int mode, int threads,
                      uint64_t* offsets, uint32_t** ids, float* t_out,
                      uint32_t* id_out, uint8_t* any_out,
                      double* seconds) const override
  {
    return rays_impl(rays, n, mode, threads, offsets, ids, t_out, id_out,
                     any_out, seconds);
  }

  // kNN (oracle/knn.h): the reference's search() with a pruning filter and a k-best functor ----
  template <typename SS = S>
  typename std::enable_if<std::is_same<SS, float>::value, int64_t>::type
  knn_impl(const float* points, uint64_t n, uint32_t k, int threads, uint32_t* ids_out, float* d2_out,
           double* seconds) const
  {
#ifdef _OPENMP
    if (threads <= 0)
      threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    auto t0c = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads)
    for (int64_t i = 0; i < (int64_t)n; ++i)
    {
      const float* q = points + i * D;
      std::vector<std::pair<float, uint32_t>> best; // unordered; `worst` valid once full
      best.reserve(k);
      size_t worst = 0;
      auto find_worst = [&]()
      {
        worst = 0;
        for (size_t j = 1; j < best.size(); ++j)
          if (rt_knn_less(best[worst].first, best[worst].second, best[j].first, best[j].second))
            worst = j;
      };
      auto bound = [&]() -> float { return best.size() < k ? INFINITY : best[worst].first; };
      auto filter = [&](box const& b) -> int
      {
        float lo[D], hi[D];
        for (unsigned d = 0; d < D; ++d)
          lo[d] = A::lo(b, d), hi[d] = A::hi(b, d);
        return rt_knn_dist2(lo, hi, q, (int)D) <= bound() ? 1 : 0; // non-strict: ties survive
      };
      auto functor = [&](value_t const& v) -> bool
      {
        box kb(v.first);
        float lo[D], hi[D];
        for (unsigned d = 0; d < D; ++d)
          lo[d] = A::lo(kb, d), hi[d] = A::hi(kb, d);
        const float d2 = rt_knn_dist2(lo, hi, q, (int)D);
        if (best.size() < k)
        {
          best.emplace_back(d2, v.second);
          if (best.size() == k)
            find_worst();
        }
        else if (rt_knn_less(d2, v.second, best[worst].first, best[worst].second))
        {
          best[worst] = { d2, v.second };
          find_worst();
        }
        return false;
      };
      tree.search(filter, functor);
      std::sort(best.begin(), best.end());
      for (uint32_t j = 0; j < k; ++j)
      {
        ids_out[i * k + j] = j < best.size() ? best[j].second : 0xFFFFFFFFu;
        d2_out[i * k + j] = j < best.size() ? best[j].first : INFINITY;
      }
    }
    auto t1c = std::chrono::steady_clock::now();
    if (seconds)
      *seconds = std::chrono::duration<double>(t1c - t0c).count();
    return 0;
  }
  template <typename SS = S>
  typename std::enable_if<!std::is_same<SS, float>::value, int64_t>::type
  knn_impl(const float*, uint64_t, uint32_t, int, uint32_t*, float*, double*) const
  {
    return -1;
  }
  int64_t search_knn(const float* points, uint64_t n, uint32_t k, int threads, uint32_t* ids_out,
                     float* d2_out, double* seconds) const override
  {
    return knn_impl(points, n, k, threads, ids_out, d2_out, seconds);
  }
};

ITree* make_tree(int kind)
{
  using er::DefaultConfig;
  switch (kind)
  {
  case 0: return new TreeImpl<OneDAdapter<int>, int, 1, false, DefaultConfig>();
  case 1: return new TreeImpl<OneDAdapter<double>, double, 1, false, DefaultConfig>();
  case 2: return new TreeImpl<NdAdapter<double, 2>, double, 2, false, DefaultConfig>();
  case 3: return new TreeImpl<NdAdapter<float, 2>, float, 2, false, DefaultConfig>();
  case 4: return new TreeImpl<NdAdapter<float, 3>, float, 3, false, DefaultConfig>();
  case 5: return new TreeImpl<NdAdapter<float, 3>, float, 3, true, DefaultConfig>();
  case 6: return new TreeImpl<NdAdapter<float, 8>, float, 8, false, QuadraticConfig16>();
  case 7: return new TreeImpl<NdAdapter<float, 2>, float, 2, true, DefaultConfig>();
  case 8: return new TreeImpl<NdAdapter<int, 2>, int, 2, false, DefaultConfig>();
  case 9: return new TreeImpl<NdAdapter<float, 3>, float, 3, false, QuadraticConfig8>();
  case 10: return new TreeImpl<OneDAdapter<float>, float, 1, false, DefaultConfig>();
  case 11: return new TreeImpl<NdAdapter<float, 4>, float, 4, false, DefaultConfig>();
  case 12: return new TreeImpl<NdAdapter<double, 3>, double, 3, false, DefaultConfig>();
  case 13: return new TreeImpl<NdAdapter<int, 3>, int, 3, false, DefaultConfig>();
  default: return nullptr;
  }
}

} // namespace

extern "C"
{
  void* ref_tree_create(int kind) { return make_tree(kind); }
  void ref_tree_destroy(void* h) { delete static_cast<ITree*>(h); }
  void ref_tree_insert(void* h, const void* keys, uint64_t n, uint32_t first_id)
  {
    static_cast<ITree*>(h)->insert(keys, n, first_id);
  }
  uint32_t ref_tree_leaf_level(void* h)
  {
    return static_cast<ITree*>(h)->leaf_level();
  }
  uint64_t ref_tree_size(void* h) { return static_cast<ITree*>(h)->size(); }
  void ref_tree_rebalance(void* h) { static_cast<ITree*>(h)->rebalance(); }
  void ref_tree_flatten(void* h, uint64_t* sizes5)
  {
    static_cast<ITree*>(h)->flatten();
    static_cast<ITree*>(h)->flat_sizes(sizes5);
  }
  void ref_tree_flat_copy(void* h, uint32_t* nodes, void* bounds,
                          uint32_t* children, uint32_t* data)
  {
    static_cast<ITree*>(h)->flat_copy(nodes, bounds, children, data);
  }
  // offsets == NULL: only count + time (the CPU-baseline leg).
  int64_t ref_search_boxes(void* h, const void* q, uint64_t n, int mode,
                           int threads, uint64_t* offsets, uint32_t** ids,
                           double* seconds)
  {
    return static_cast<ITree*>(h)->search_boxes(q, n, mode, threads, offsets,
                                                ids, seconds);
  }
  int64_t ref_search_rays(void* h, const float* rays, uint64_t n, int mode,
                          int threads, uint64_t* offsets, uint32_t** ids,
                          float* t_out, uint32_t* id_out, uint8_t* any_out,
                          double* seconds)
  {
    return static_cast<ITree*>(h)->search_rays(rays, n, mode, threads, offsets,
                                               ids, t_out, id_out, any_out,
                                               seconds);
  }
  int64_t ref_tree_set_keys(void* h, const void* keys, uint64_t n)
  {
    return static_cast<ITree*>(h)->set_keys(keys, n);
  }
  uint64_t ref_tree_erase_ids(void* h, const uint32_t* ids, uint64_t n)
  {
    return static_cast<ITree*>(h)->erase_ids(ids, n);
  }
  int64_t ref_search_knn(void* h, const float* points, uint64_t n, uint32_t k, int threads,
                         uint32_t* ids_out, float* d2_out, double* seconds)
  {
    return static_cast<ITree*>(h)->search_knn(points, n, k, threads, ids_out, d2_out, seconds);
  }
  void ref_free(void* p) { std::free(p); }
  int ref_max_threads()
  {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
  }
}
