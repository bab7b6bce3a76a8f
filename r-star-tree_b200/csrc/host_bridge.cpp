// host_bridge.cpp -- C entry points over the header-only host library (RTree.hpp).
//
// The host API is a C++ template library, exactly like the reference's.  Python (the
// test and bench driver) cannot instantiate templates, so this file instantiates the
// configurations of BASELINE.json ("kinds") and exposes them as handles:
// build a tree from an array of keys with RTree::insert(), flatten(), flatten_device(),
// and the CPU RTree::search().  It contains no GPU code and no query-path logic of its
// own; the GPU side is librtree_b200.so (rt_api.cu).
//
// kind  tree type                                                         (BASELINE config)
//   0   RTree<aabb_t<int>, int, u32>                 1-D i32, test/rtree.cpp "Flatten"
//   1   RTree<aabb_t<double>, double, u32>           1-D f64, example/sample
//   2   RTree<aabb<point<double,2>>, point, u32>     config 1 (visualize_2d)
//   3   RTree<aabb<point<float,2>>, point, u32>
//   4   RTree<aabb<point<float,3>>, point, u32>      config 2 / 5
//   5   RTree<aabb<point<float,3>>, aabb, u32>       config 3 (triangle AABBs)
//   6   RTree<aabb<point<float,8>>, point, u32, {8,16,6,Quadratic}>   config 4
//   7   RTree<aabb<point<float,2>>, aabb, u32>
//   8   RTree<aabb<point<int,2>>, point, u32>
//   9   RTree<aabb<point<float,3>>, point, u32, {4,8,3,Quadratic}>
//  10   RTree<aabb_t<float>, float, u32>
//  11   RTree<aabb<point<float,4>>, point, u32>
//  12   RTree<aabb<point<double,3>>, point, u32>
//  13   RTree<aabb<point<int,3>>, point, u32>
#include <RTree.hpp>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

namespace rt = eh::rtree;

namespace
{
thread_local std::string g_error;

struct Quadratic16
{
  constexpr static rt::size_type MIN_ENTRIES = 8;
  constexpr static rt::size_type MAX_ENTRIES = 16;
  constexpr static rt::size_type REINSERT_COUNT = 6;
  using split_algorithm = rt::QuadraticSplit;
};
struct Quadratic8
{
  constexpr static rt::size_type MIN_ENTRIES = 4;
  constexpr static rt::size_type MAX_ENTRIES = 8;
  constexpr static rt::size_type REINSERT_COUNT = 3;
  using split_algorithm = rt::QuadraticSplit;
};

struct TreeHandle
{
  virtual ~TreeHandle() {}
  virtual void insert(const void* keys, uint64_t n, uint32_t first_id) = 0;
  virtual int64_t set_keys(const void* keys, uint64_t n) = 0;
  virtual uint64_t erase_ids(const uint32_t* ids, uint64_t n) = 0;
  virtual void clear() = 0;
  virtual void rebalance() = 0;
  virtual uint32_t leaf_level() const = 0;
  virtual uint64_t size() const = 0;
  virtual int check_invariants() const = 0;
  virtual void flatten(uint64_t* sizes5) = 0;
  virtual void flat_copy(uint32_t* nodes, void* bounds, uint32_t* children, uint32_t* data) const = 0;
  virtual int64_t search_boxes(const void* q, uint64_t n, int mode, uint64_t* offsets,
                               uint32_t** ids) const = 0;
  virtual rt::device_tree_t* flatten_device(int ids_from_mapped) const = 0;
  virtual rt::device_tree_t* flatten_device_from_arrays(uint32_t leaf_level, const uint32_t* nodes,
                                                        uint64_t n_nodes, const void* bounds,
                                                        const uint32_t* children, uint64_t n_children,
                                                        const uint32_t* data, uint64_t n_data,
                                                        int ids_from_mapped) const = 0;
};

// how one row of a C array becomes a key / a box of the tree's types
template <typename S, unsigned D>
struct Nd
{
  using point = rt::point_t<S, D>;
  using box = rt::aabb_t<point>;
  static point to_point(const S* p)
  {
    point r;
    for (unsigned d = 0; d < D; ++d)
      r[d] = p[d];
    return r;
  }
  static box to_box(const S* p) { return box(to_point(p), to_point(p + D)); }
};
template <typename S>
struct Scalar1d
{
  using point = S;
  using box = rt::aabb_t<S>;
  static point to_point(const S* p) { return p[0]; }
  static box to_box(const S* p) { return box(p[0], p[1]); }
};

template <typename G, typename S, unsigned D, bool BOXKEY, typename Config>
struct Tree : TreeHandle
{
  using box = typename G::box;
  using point = typename G::point;
  using key = typename std::conditional<BOXKEY, box, point>::type;
  using tree_t = rt::RTree<box, key, uint32_t, Config>;
  using value_t = typename tree_t::value_type;
  using flat_t = typename tree_t::flatten_result_t;
  using gt = rt::geometry_traits<box>;
  constexpr static unsigned KEY_SCALARS = BOXKEY ? 2 * D : D;

  tree_t tree;
  flat_t flat;

  static key to_key(const S* p, std::true_type) { return G::to_box(p); }
  static key to_key(const S* p, std::false_type) { return G::to_point(p); }

  void insert(const void* keys, uint64_t n, uint32_t first_id) override
  {
    const S* p = static_cast<const S*>(keys);
    for (uint64_t i = 0; i < n; ++i)
      tree.emplace(to_key(p + i * KEY_SCALARS, std::integral_constant<bool, BOXKEY> {}),
                   first_id + static_cast<uint32_t>(i));
  }
  // keys changed in place through the iterators (iteration order = flatten data order), then
  // rebound() for every value: the host-side form of what rt_refit() does on the device
  int64_t set_keys(const void* keys, uint64_t n) override
  {
    const S* p = static_cast<const S*>(keys);
    uint64_t i = 0;
    for (auto it = tree.begin(); it != tree.end(); ++it, ++i)
    {
      if (i >= n)
        return -1;
      it->first = to_key(p + i * KEY_SCALARS, std::integral_constant<bool, BOXKEY> {});
    }
    if (i != n)
      return -1;
    for (auto it = tree.begin(); it != tree.end(); ++it)
      tree.rebound(it);
    return (int64_t)i;
  }
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
    return erased;
  }
  void clear() override { tree.clear(); }
  void rebalance() override { tree.rebalance(); }
  uint32_t leaf_level() const override { return static_cast<uint32_t>(tree.leaf_level()); }
  uint64_t size() const override { return tree.size(); }

  // structural invariants of reference test/rtree.cpp:32-95 ("Insert"): fill bounds,
  // parent back-pointers, parent bound == recomputed child bound, iteration count
  int check_invariants() const override
  {
    const int L = tree.leaf_level();
    for (int level = 0; level < L; ++level)
    {
      for (auto ni = tree.node_begin(level); ni != tree.node_end(level); ++ni)
      {
        auto const* node = *ni;
        if (node != tree.root() && (node->size() < tree_t::MIN_ENTRIES || node->size() > tree_t::MAX_ENTRIES))
          return 1;
        if (node == tree.root() && node->parent() != nullptr)
          return 2;
        for (auto const& c : *node)
        {
          if (c.second->parent() != node)
            return 3;
          box b = (level + 1 == L) ? c.second->as_leaf()->calculate_bound()
                                   : c.second->as_node()->calculate_bound();
          for (unsigned d = 0; d < D; ++d)
          {
            if (gt::min_point(b, d) != gt::min_point(c.first, d)
                || gt::max_point(b, d) != gt::max_point(c.first, d))
              return 4;
          }
        }
      }
    }
    uint64_t leaves_total = 0;
    for (auto li = tree.leaf_begin(); li != tree.leaf_end(); ++li)
    {
      auto const* leaf = *li;
      if (L > 0 && (leaf->size() < tree_t::MIN_ENTRIES || leaf->size() > tree_t::MAX_ENTRIES))
        return 5;
      leaves_total += leaf->size();
    }
    uint64_t iterated = 0;
    for (auto it = tree.begin(); it != tree.end(); ++it)
      ++iterated;
    if (iterated != leaves_total || iterated != tree.size())
      return 6;
    return 0;
  }

  void flatten(uint64_t* sizes5) override
  {
    flat = tree.flatten();
    sizes5[0] = flat.nodes.size();
    sizes5[1] = flat.children.size();
    sizes5[2] = flat.data.size();
    sizes5[3] = flat.leaf_level;
    sizes5[4] = flat.root;
  }
  void flat_copy(uint32_t* nodes, void* bounds, uint32_t* children, uint32_t* data) const override
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
        b[i * 2 * D + d] = gt::min_point(flat.children_bound[i], d);
        b[i * 2 * D + D + d] = gt::max_point(flat.children_bound[i], d);
      }
    }
    std::copy(flat.children.begin(), flat.children.end(), children);
    std::copy(flat.data.begin(), flat.data.end(), data);
  }

  // The host library's own CPU search() with the canonical functors -- host API
  // parity only; the GPU path never comes through here.
  int64_t search_boxes(const void* qv, uint64_t n, int mode, uint64_t* offsets,
                       uint32_t** ids_out) const override
  {
    const S* qs = static_cast<const S*>(qv);
    std::vector<uint32_t> all;
    for (uint64_t i = 0; i < n; ++i)
    {
      const box q = G::to_box(qs + i * 2 * D);
      const size_t begin = all.size();
      offsets[i] = begin;
      auto filter = [&q](box const& b) -> int
      {
        for (unsigned d = 0; d < D; ++d)
        {
          if (gt::max_point(b, d) < gt::min_point(q, d) || gt::min_point(b, d) > gt::max_point(q, d))
            return 0;
        }
        return 1;
      };
      auto functor = [&](value_t const& v) -> bool
      {
        const box k(v.first);
        bool hit = true;
        for (unsigned d = 0; d < D && hit; ++d)
        {
          if (mode == (int)RT_CONTAINS || !BOXKEY)
            hit = !(gt::min_point(k, d) < gt::min_point(q, d) || gt::max_point(k, d) > gt::max_point(q, d));
          else
            hit = !(gt::max_point(k, d) < gt::min_point(q, d) || gt::min_point(k, d) > gt::max_point(q, d));
        }
        if (hit)
          all.push_back(v.second);
        return false;
      };
      tree.search(filter, functor);
      std::sort(all.begin() + begin, all.end());
    }
    offsets[n] = all.size();
    uint32_t* ids = static_cast<uint32_t*>(std::malloc(sizeof(uint32_t) * std::max<size_t>(all.size(), 1)));
    std::copy(all.begin(), all.end(), ids);
    *ids_out = ids;
    return static_cast<int64_t>(all.size());
  }

  rt::device_tree_t* flatten_device(int ids_from_mapped) const override
  {
    return new rt::device_tree_t(tree.flatten_device(
        ids_from_mapped ? rt::id_source::mapped_value : rt::id_source::data_index));
  }
  // feed a flatten_result_t that came from elsewhere (e.g. the reference's flatten())
  rt::device_tree_t* flatten_device_from_arrays(uint32_t leaf_level, const uint32_t* nodes,
                                                uint64_t n_nodes, const void* bounds,
                                                const uint32_t* children, uint64_t n_children,
                                                const uint32_t* data, uint64_t n_data,
                                                int ids_from_mapped) const override
  {
    flat_t f;
    f.leaf_level = leaf_level;
    f.root = 0;
    f.nodes.resize(n_nodes);
    for (uint64_t i = 0; i < n_nodes; ++i)
      f.nodes[i] = { nodes[3 * i], nodes[3 * i + 1], nodes[3 * i + 2] };
    const S* b = static_cast<const S*>(bounds);
    f.children_bound.reserve(n_children);
    for (uint64_t i = 0; i < n_children; ++i)
      f.children_bound.push_back(G::to_box(b + i * 2 * D));
    f.children.assign(children, children + n_children);
    f.data.assign(data, data + n_data);
    return new rt::device_tree_t(rt::flatten_device<Config::MAX_ENTRIES, BOXKEY>(
        f, ids_from_mapped ? rt::id_source::mapped_value : rt::id_source::data_index));
  }
};

TreeHandle* make_tree(int kind)
{
  using rt::DefaultConfig;
  switch (kind)
  {
  case 0: return new Tree<Scalar1d<int32_t>, int32_t, 1, false, DefaultConfig>();
  case 1: return new Tree<Scalar1d<double>, double, 1, false, DefaultConfig>();
  case 2: return new Tree<Nd<double, 2>, double, 2, false, DefaultConfig>();
  case 3: return new Tree<Nd<float, 2>, float, 2, false, DefaultConfig>();
  case 4: return new Tree<Nd<float, 3>, float, 3, false, DefaultConfig>();
  case 5: return new Tree<Nd<float, 3>, float, 3, true, DefaultConfig>();
  case 6: return new Tree<Nd<float, 8>, float, 8, false, Quadratic16>();
  case 7: return new Tree<Nd<float, 2>, float, 2, true, DefaultConfig>();
  case 8: return new Tree<Nd<int32_t, 2>, int32_t, 2, false, DefaultConfig>();
  case 9: return new Tree<Nd<float, 3>, float, 3, false, Quadratic8>();
  case 10: return new Tree<Scalar1d<float>, float, 1, false, DefaultConfig>();
  case 11: return new Tree<Nd<float, 4>, float, 4, false, DefaultConfig>();
  case 12: return new Tree<Nd<double, 3>, double, 3, false, DefaultConfig>();
  case 13: return new Tree<Nd<int32_t, 3>, int32_t, 3, false, DefaultConfig>();
  default: return nullptr;
  }
}

template <typename F>
auto guarded(F&& f, decltype(f()) on_error) -> decltype(f())
{
  try
  {
    return f();
  }
  catch (std::exception const& e)
  {
    g_error = e.what();
  }
  catch (...)
  {
    g_error = "unknown C++ exception";
  }
  return on_error;
}

} // namespace

extern "C"
{
  const char* rtb_last_error() { return g_error.c_str(); }
  void* rtb_tree_create(int kind) { return make_tree(kind); }
  void rtb_tree_destroy(void* h) { delete static_cast<TreeHandle*>(h); }
  int rtb_tree_insert(void* h, const void* keys, uint64_t n, uint32_t first_id)
  {
    return guarded([&] { static_cast<TreeHandle*>(h)->insert(keys, n, first_id); return 0; }, -1);
  }
  int64_t rtb_tree_set_keys(void* h, const void* keys, uint64_t n)
  {
    return guarded([&] { return static_cast<TreeHandle*>(h)->set_keys(keys, n); }, (int64_t)-1);
  }
  int64_t rtb_tree_erase_ids(void* h, const uint32_t* ids, uint64_t n)
  {
    return guarded([&] { return (int64_t) static_cast<TreeHandle*>(h)->erase_ids(ids, n); }, (int64_t)-1);
  }
  void rtb_tree_clear(void* h) { static_cast<TreeHandle*>(h)->clear(); }
  void rtb_tree_rebalance(void* h) { static_cast<TreeHandle*>(h)->rebalance(); }
  uint32_t rtb_tree_leaf_level(void* h) { return static_cast<TreeHandle*>(h)->leaf_level(); }
  uint64_t rtb_tree_size(void* h) { return static_cast<TreeHandle*>(h)->size(); }
  int rtb_tree_check_invariants(void* h) { return static_cast<TreeHandle*>(h)->check_invariants(); }
  void rtb_tree_flatten(void* h, uint64_t* sizes5) { static_cast<TreeHandle*>(h)->flatten(sizes5); }
  void rtb_tree_flat_copy(void* h, uint32_t* nodes, void* bounds, uint32_t* children, uint32_t* data)
  {
    static_cast<TreeHandle*>(h)->flat_copy(nodes, bounds, children, data);
  }
  int64_t rtb_tree_search_boxes(void* h, const void* q, uint64_t n, int mode, uint64_t* offsets,
                                uint32_t** ids)
  {
    return guarded([&] { return static_cast<TreeHandle*>(h)->search_boxes(q, n, mode, offsets, ids); },
                   (int64_t)-1);
  }
  void rtb_free(void* p) { std::free(p); }

  // RTree::flatten_device() of the tree behind `h`
  void* rtb_flatten_device(void* h, int ids_from_mapped)
  {
    return guarded([&]() -> void* { return static_cast<TreeHandle*>(h)->flatten_device(ids_from_mapped); },
                   (void*)nullptr);
  }
  // flatten_device() applied to reference-shaped flatten arrays of tree kind `kind`
  void* rtb_flatten_device_from_arrays(int kind, uint32_t leaf_level, const uint32_t* nodes,
                                       uint64_t n_nodes, const void* bounds, const uint32_t* children,
                                       uint64_t n_children, const uint32_t* data, uint64_t n_data,
                                       int ids_from_mapped)
  {
    return guarded(
        [&]() -> void*
        {
          TreeHandle* t = make_tree(kind);
          if (t == nullptr)
            throw std::invalid_argument("unknown tree kind");
          rt::device_tree_t* image = nullptr;
          try
          {
            image = t->flatten_device_from_arrays(leaf_level, nodes, n_nodes, bounds, children,
                                                  n_children, data, n_data, ids_from_mapped);
          }
          catch (...)
          {
            delete t;
            throw;
          }
          delete t;
          return image;
        },
        (void*)nullptr);
  }
  void rtb_device_tree_desc(void* image, rt_tree_desc* out)
  {
    *out = static_cast<rt::device_tree_t*>(image)->desc();
  }
  // new (breadth-first) node index -> index in flatten_result_t::nodes
  const uint32_t* rtb_device_tree_bfs_to_dfs(void* image)
  {
    return static_cast<rt::device_tree_t*>(image)->bfs_to_dfs.data();
  }
  void rtb_device_tree_free(void* image) { delete static_cast<rt::device_tree_t*>(image); }
  // on-disk form of the image (RTree/device_layout.hpp: save_device_tree / load_device_tree)
  int rtb_device_tree_save(void* image, const char* path)
  {
    return guarded(
        [&]() -> int
        {
          rt::save_device_tree(*static_cast<rt::device_tree_t*>(image), path);
          return 0;
        },
        -1);
  }
  void* rtb_device_tree_load(const char* path)
  {
    return guarded([&]() -> void* { return new rt::device_tree_t(rt::load_device_tree(path)); }, (void*)nullptr);
  }
}
