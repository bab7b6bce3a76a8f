// RTree/rtree.hpp -- host side (header-only C++17) of rtree-b200.
//
// eh::rtree::RTree<Geometry, Key, Mapped, Config, Allocator>: the reference's public
// API (reference: include/RTree/rtree.hpp:38-1176, README.md:89-126) written from
// scratch.  The CPU builds the tree (insert / forced reinsert / split), the GPU
// answers the queries:
//
//   build side   insert, emplace, erase, clear, rebound, rebalance, iterators
//                -- sequential pointer code, a PRODUCER of trees only
//   read side    search()        reference rtree.hpp:1141-1152  (CPU, for host users)
//                flatten()       reference rtree.hpp:966-981    (the hand-off format)
//                flatten_device()  NEW: flatten() re-laid as level-ordered SoA blocks
//                                  for rt_upload() (see device_layout.hpp)
//
// Behavioural notes are in SURVEY.md appendix A.  Known reference defects are fixed
// rather than reproduced: search_iterator works with iterator functors (B6),
// flatten_move really moves (B7), the converting constructor initialises the root
// (B8), node_cbegin compiles (B9).
#pragma once

#include <algorithm>
#include <iterator>
#include <limits>
#include <memory>
#include <type_traits>
#include <utility>
#include <vector>

#include "device_layout.hpp"
#include "geometry.hpp"
#include "node.hpp"
#include "split.hpp"

namespace eh
{
namespace rtree
{

struct DefaultConfig
{
  /// minimum number of entries in a node (root exempt)
  constexpr static size_type MIN_ENTRIES = 4;
  /// maximum number of entries in a node == width of a device block
  constexpr static size_type MAX_ENTRIES = 8;
  /// entries taken out and re-inserted on the first overflow of an insert()
  constexpr static size_type REINSERT_COUNT = 3;
  /// RStarSplit or QuadraticSplit (or any type with the same static split())
  using split_algorithm = RStarSplit;
};

template <typename GeometryType, typename KeyType, typename MappedType,
          typename Config = DefaultConfig, template <typename _T> class Allocator = std::allocator>
class RTree
{
public:
  using size_type = ::eh::rtree::size_type;
  using geometry_type = GeometryType;
  using key_type = KeyType;
  using mapped_type = MappedType;
  using value_type = std::pair<key_type, mapped_type>;
  using traits = geometry_traits<geometry_type>;
  using scalar_type = typename traits::scalar_type;

  constexpr static size_type MAX_ENTRIES = Config::MAX_ENTRIES;
  constexpr static size_type MIN_ENTRIES = Config::MIN_ENTRIES;
  constexpr static size_type REINSERT_COUNT = Config::REINSERT_COUNT;
  static_assert(MIN_ENTRIES >= 1, "Invalid MIN_ENTRIES count");
  static_assert(MIN_ENTRIES <= MAX_ENTRIES / 2, "Invalid MIN_ENTRIES count");
  static_assert(MAX_ENTRIES + 1 - REINSERT_COUNT >= MIN_ENTRIES, "Invalid REINSERT_COUNT count");
  static_assert(MAX_ENTRIES + 1 - REINSERT_COUNT <= MAX_ENTRIES, "Invalid REINSERT_COUNT count");

  using node_base_type = static_node_base_t<GeometryType, KeyType, MappedType, MIN_ENTRIES, MAX_ENTRIES>;
  using node_type = static_node_t<GeometryType, KeyType, MappedType, MIN_ENTRIES, MAX_ENTRIES>;
  using leaf_type = static_leaf_node_t<GeometryType, KeyType, MappedType, MIN_ENTRIES, MAX_ENTRIES>;

  using node_allocator_type = Allocator<node_type>;
  using leaf_allocator_type = Allocator<leaf_type>;

  using iterator = iterator_t<leaf_type>;
  using const_iterator = iterator_t<leaf_type const>;
  using node_iterator = node_iterator_t<node_type>;
  using const_node_iterator = node_iterator_t<node_type const>;
  using leaf_iterator = node_iterator_t<leaf_type>;
  using const_leaf_iterator = node_iterator_t<leaf_type const>;

  // ---- flatten(): the reference's hand-off format, field for field ---------------
  // (reference: include/RTree/rtree.hpp:841-870)
  struct flatten_node_t
  {
    /// first slot of this node in children / children_bound
    size_type offset;
    /// number of children
    size_type size;
    /// index of the parent node (0 for the root)
    size_type parent;
  };
  struct flatten_result_t
  {
    /// level of the leaf nodes (root is level 0)
    size_type leaf_level;
    /// always 0
    size_type root;
    /// every node, leaves included, in depth-first pre-order
    std::vector<flatten_node_t> nodes;
    /// bound of every child slot (for a leaf: the key, widened to geometry_type)
    std::vector<geometry_type> children_bound;
    /// internal node: index of the child node; leaf: index into data
    std::vector<size_type> children;
    /// the mapped values in left-to-right leaf order
    std::vector<mapped_type> data;
  };

protected:
  node_base_type* _root = nullptr;
  int _leaf_level = 0;

public:
  // ---- node allocation (also used by the node types) --------------------------------
  template <typename NodeT>
  NodeT* construct_node()
  {
    using alloc_t = Allocator<NodeT>;
    alloc_t alloc;
    NodeT* p = std::allocator_traits<alloc_t>::allocate(alloc, 1);
    ::new (static_cast<void*>(p)) NodeT();
    return p;
  }
  template <typename NodeT>
  void destroy_node(NodeT* p)
  {
    using alloc_t = Allocator<NodeT>;
    alloc_t alloc;
    p->~NodeT();
    std::allocator_traits<alloc_t>::deallocate(alloc, p, 1);
  }

protected:
  void release_all()
  {
    if (_root == nullptr)
      return;
    if (_leaf_level == 0)
    {
      destroy_node(_root->as_leaf());
    }
    else
    {
      _root->as_node()->delete_recursive(_leaf_level, *this);
      destroy_node(_root->as_node());
    }
    _root = nullptr;
    _leaf_level = 0;
  }
  void init_root()
  {
    if (_root == nullptr)
    {
      _root = construct_node<leaf_type>();
      _leaf_level = 0;
    }
  }

  /// ChooseSubtree (Guttman CL1-CL4): walk down to `target_level` always taking the
  /// child needing the least area enlargement, ties to the smaller area.
  node_type* choose_insert_target(geometry_type const& bound, int target_level)
  {
    node_type* n = _root->as_node();
    for (int level = 0; level < target_level; ++level)
    {
      size_type chosen = 0;
      scalar_type least_growth = std::numeric_limits<scalar_type>::max();
      scalar_type chosen_area = 0;
      bool have = false;
      for (size_type i = 0; i < n->size(); ++i)
      {
        geometry_type const& g = n->at(i).first;
        const scalar_type a = helper::area(g);
        const scalar_type growth = helper::enlarged_area(g, bound) - a;
        if (!have || growth < least_growth || (growth == least_growth && a < chosen_area))
        {
          have = true;
          chosen = i;
          least_growth = growth;
          chosen_area = a;
        }
      }
      n = n->at(chosen).second->as_node();
    }
    return n;
  }

  static bool same_bound(geometry_type const& a, geometry_type const& b)
  {
    for (int d = 0; d < traits::DIM; ++d)
    {
      if (!(helper::min_point(a, d) == helper::min_point(b, d))
          || !(helper::max_point(a, d) == helper::max_point(b, d)))
        return false;
    }
    return true;
  }

  /// recompute the bound stored for `n` in its parent and carry the change towards
  /// the root; stops as soon as a stored bound is already right, because every
  /// bound above depends on it alone
  template <typename NodeT>
  void rebound_upward(NodeT* n)
  {
    if (n->_parent == nullptr)
      return;
    {
      geometry_type& stored = n->_parent->at(n->_index_on_parent).first;
      geometry_type fresh = n->calculate_bound();
      if (same_bound(stored, fresh))
        return;
      stored = fresh;
    }
    for (node_type* a = n->_parent; a->_parent != nullptr; a = a->_parent)
    {
      geometry_type& stored = a->_parent->at(a->_index_on_parent).first;
      geometry_type fresh = a->calculate_bound();
      if (same_bound(stored, fresh))
        return;
      stored = fresh;
    }
  }

  /// put `entry` into `target`, treating overflow: the first overflow of an insert()
  /// re-inserts (unless target is the root or an only child), any other one splits
  template <typename NodeT>
  void insert_node(NodeT* target, typename NodeT::value_type entry, bool may_reinsert)
  {
    NodeT* sibling = nullptr;
    if (target->size() == MAX_ENTRIES)
    {
      if (may_reinsert && static_cast<node_base_type*>(target) != _root
          && target->parent()->size() > 1)
      {
        reinsert(target, std::move(entry));
      }
      else
      {
        sibling = construct_node<NodeT>();
        Config::split_algorithm::split(target, std::move(entry), sibling);
      }
    }
    else
    {
      target->insert(std::move(entry));
    }
    rebound_upward(target);

    if (sibling != nullptr)
    {
      if (static_cast<node_base_type*>(target) == _root)
      {
        node_type* new_root = construct_node<node_type>();
        new_root->insert({ target->calculate_bound(), target });
        new_root->insert({ sibling->calculate_bound(), sibling });
        _root = new_root;
        ++_leaf_level;
      }
      else
      {
        insert_node(target->parent(), { sibling->calculate_bound(), sibling }, true);
      }
    }
  }

  /// R*-tree forced reinsert: keep the M+1-p entries nearest to the node centre,
  /// re-insert the p farthest ones (nearest of those first) from the top
  template <typename NodeT>
  void reinsert(NodeT* node, typename NodeT::value_type extra)
  {
    using entry_t = typename NodeT::value_type;
    // height above the leaves stays valid while re-insertions grow the tree
    const int height = std::is_same<NodeT, leaf_type>::value ? 0 : _leaf_level - node->level_recursive();

    geometry_type centre_of = node->calculate_bound();
    helper::enlarge_to(centre_of, extra.first);

    detail::overflow_set<NodeT> set(node, std::move(extra));
    constexpr int COUNT = static_cast<int>(MAX_ENTRIES) + 1;
    std::pair<scalar_type, int> order[COUNT];
    for (int i = 0; i < COUNT; ++i)
      order[i] = { helper::distance_center(centre_of, set.entries[i].first), i };
    std::sort(order, order + COUNT,
              [](std::pair<scalar_type, int> const& a, std::pair<scalar_type, int> const& b)
              { return a.first < b.first; });

    constexpr int KEEP = COUNT - static_cast<int>(REINSERT_COUNT);
    for (int i = 0; i < KEEP; ++i)
      node->insert(std::move(set.entries[order[i].second]));
    rebound_upward(node);

    for (int i = KEEP; i < COUNT; ++i)
    {
      entry_t& e = set.entries[order[i].second];
      NodeT* target = static_cast<NodeT*>(static_cast<node_base_type*>(
          choose_insert_target(geometry_type(e.first), _leaf_level - height)));
      insert_node(target, std::move(e), false);
    }
  }

public:
  // ---- construction ---------------------------------------------------------------
  RTree() { init_root(); }
  template <typename Iterator>
  RTree(Iterator first, Iterator last)
  {
    init_root();
    for (; first != last; ++first)
      insert(*first);
  }
  RTree(RTree const& rhs) { copy_from(rhs); }
  RTree(RTree&& rhs) noexcept
      : _root(rhs._root)
      , _leaf_level(rhs._leaf_level)
  {
    rhs._root = nullptr;
    rhs._leaf_level = 0;
  }
  /// rebuild from a tree with another Config / Allocator
  template <typename Config2, template <typename> class Allocator2>
  RTree(RTree<GeometryType, KeyType, MappedType, Config2, Allocator2> const& rhs)
  {
    init_root();
    for (auto const& v : rhs)
      insert(v);
  }
  RTree& operator=(RTree const& rhs)
  {
    if (this != &rhs)
    {
      release_all();
      copy_from(rhs);
    }
    return *this;
  }
  RTree& operator=(RTree&& rhs) noexcept
  {
    if (this != &rhs)
    {
      release_all();
      _root = rhs._root;
      _leaf_level = rhs._leaf_level;
      rhs._root = nullptr;
      rhs._leaf_level = 0;
    }
    return *this;
  }
  ~RTree() { release_all(); }

protected:
  void copy_from(RTree const& rhs)
  {
    if (rhs._root == nullptr)
    {
      init_root();
      return;
    }
    _leaf_level = rhs._leaf_level;
    if (rhs._leaf_level == 0)
      _root = rhs._root->as_leaf()->clone_recursive(*this);
    else
      _root = rhs._root->as_node()->clone_recursive(rhs._leaf_level, *this);
  }

public:
  // ---- modifiers ------------------------------------------------------------------
  void insert(value_type value)
  {
    init_root();
    leaf_type* leaf = static_cast<leaf_type*>(
        static_cast<node_base_type*>(choose_insert_target(geometry_type(value.first), _leaf_level)));
    insert_node(leaf, std::move(value), true);
  }
  template <typename... Args>
  void emplace(Args&&... args)
  {
    insert(value_type(std::forward<Args>(args)...));
  }

  /// remove one value; under-full nodes are dissolved and their content re-inserted
  void erase(iterator pos)
  {
    leaf_type* leaf = pos._leaf;
    leaf->erase(leaf->begin() + pos._index);

    // (node, height above the leaves) of every node unlinked on the way up
    std::vector<std::pair<node_base_type*, int>> orphans;
    node_base_type* n = leaf;
    int height = 0;
    while (n != _root)
    {
      node_type* p = n->_parent;
      const size_type fill = (height == 0) ? n->as_leaf()->size() : n->as_node()->size();
      if (fill < MIN_ENTRIES)
      {
        p->erase(n);
        n->_parent = nullptr;
        orphans.emplace_back(n, height);
      }
      else
      {
        p->at(n->_index_on_parent).first
            = (height == 0) ? n->as_leaf()->calculate_bound() : n->as_node()->calculate_bound();
      }
      n = p;
      ++height;
    }
    shrink_root();
    for (auto& o : orphans)
      reattach(o.first, o.second);
  }
  void erase(const_iterator pos)
  {
    erase(iterator(const_cast<leaf_type*>(pos._leaf), pos._index));
  }

protected:
  /// an internal root with one child hands over to it; an empty one becomes a leaf
  void shrink_root()
  {
    while (_leaf_level > 0)
    {
      node_type* r = _root->as_node();
      if (r->size() == 1)
      {
        node_base_type* only = r->at(0).second;
        only->_parent = nullptr;
        only->_index_on_parent = 0;
        r->clear();
        destroy_node(r);
        _root = only;
        --_leaf_level;
      }
      else if (r->size() == 0)
      {
        destroy_node(r);
        _root = nullptr;
        _leaf_level = 0;
        init_root();
      }
      else
      {
        break;
      }
    }
  }
  /// give the content of an unlinked node (of the given height) back to the tree
  void reattach(node_base_type* n, int height)
  {
    if (height == 0)
    {
      leaf_type* leaf = n->as_leaf();
      for (auto& v : *leaf)
        insert(std::move(v));
      leaf->clear();
      destroy_node(leaf);
      return;
    }
    node_type* node = n->as_node();
    if (_leaf_level >= height)
    {
      // children (height-1) go below a node of height `height`
      for (auto& c : *node)
      {
        node_type* target = choose_insert_target(c.first, _leaf_level - height);
        insert_node(target, std::move(c), true);
      }
    }
    else
    {
      // the tree became shorter than this subtree: dissolve one more level
      for (auto& c : *node)
      {
        c.second->_parent = nullptr;
        reattach(c.second, height - 1);
      }
    }
    node->clear();
    destroy_node(node);
  }

public:
  void clear()
  {
    release_all();
    init_root();
  }

  /// a key was modified in place through `pos`: refit the bounds above it
  void rebound(iterator pos) { rebound_upward(pos._leaf); }
  /// refit the bounds above a node (node_type* or leaf_type*)
  template <typename NodeT>
  void rebound(NodeT* n)
  {
    rebound_upward(n);
  }

  /// rebuild by re-inserting every value (better balance after many updates)
  void rebalance()
  {
    RTree fresh(std::make_move_iterator(begin()), std::make_move_iterator(end()));
    *this = std::move(fresh);
  }

  // ---- observers ------------------------------------------------------------------
  /// number of values; O(N)
  size_type size() const
  {
    if (_root == nullptr)
      return 0;
    return _leaf_level == 0 ? _root->as_leaf()->size() : _root->as_node()->size_recursive(_leaf_level);
  }
  bool empty() const { return size() == 0; }
  node_type* root() { return _root->as_node(); }
  node_type const* root() const { return _root->as_node(); }
  int leaf_level() const { return _leaf_level; }

protected:
  node_base_type* leftmost(int level) const
  {
    node_base_type* n = _root;
    for (int l = 0; l < level; ++l)
      n = n->as_node()->at(0).second;
    return n;
  }

public:
  // ---- iteration ------------------------------------------------------------------
  iterator begin()
  {
    leaf_type* leaf = leftmost(_leaf_level)->as_leaf();
    return leaf->empty() ? iterator() : iterator(leaf, size_type(0));
  }
  const_iterator begin() const { return cbegin(); }
  const_iterator cbegin() const
  {
    leaf_type const* leaf = leftmost(_leaf_level)->as_leaf();
    return leaf->empty() ? const_iterator() : const_iterator(leaf, size_type(0));
  }
  iterator end() { return iterator(); }
  const_iterator end() const { return const_iterator(); }
  const_iterator cend() const { return const_iterator(); }

  /// nodes of one internal level (0 <= level < leaf_level())
  node_iterator node_begin(int level) { return node_iterator(leftmost(level)->as_node()); }
  const_node_iterator node_begin(int level) const { return node_cbegin(level); }
  const_node_iterator node_cbegin(int level) const
  {
    return const_node_iterator(leftmost(level)->as_node());
  }
  node_iterator node_end(int) { return node_iterator(); }
  const_node_iterator node_end(int) const { return const_node_iterator(); }
  const_node_iterator node_cend(int) const { return const_node_iterator(); }

  leaf_iterator leaf_begin() { return leaf_iterator(leftmost(_leaf_level)->as_leaf()); }
  const_leaf_iterator leaf_begin() const { return leaf_cbegin(); }
  const_leaf_iterator leaf_cbegin() const
  {
    return const_leaf_iterator(leftmost(_leaf_level)->as_leaf());
  }
  leaf_iterator leaf_end() { return leaf_iterator(); }
  const_leaf_iterator leaf_end() const { return const_leaf_iterator(); }
  const_leaf_iterator leaf_cend() const { return const_leaf_iterator(); }

  // ---- search (CPU) ---------------------------------------------------------------
protected:
  /// depth-first; filter(bound) -> 1 descend / 0 skip / -1 abort; at the leaves every
  /// value goes to the functor (no filtering), true aborts.  The root's own bound is
  /// never tested.  `make_arg(leaf, i)` shapes what the functor receives.
  template <typename NodePtr, typename Filter, typename Functor, typename MakeArg>
  bool walk(Filter& filter, Functor& functor, MakeArg& make_arg, NodePtr node, int level) const
  {
    if (level == _leaf_level)
    {
      auto* leaf = node->as_leaf();
      for (size_type i = 0; i < leaf->size(); ++i)
      {
        if (functor(make_arg(leaf, i)))
          return true;
      }
      return false;
    }
    auto* inner = node->as_node();
    for (size_type i = 0; i < inner->size(); ++i)
    {
      const int verdict = filter(inner->at(i).first);
      if (verdict == -1)
        return true;
      if (verdict == 1)
      {
        NodePtr child = inner->at(i).second;
        if (walk(filter, functor, make_arg, child, level + 1))
          return true;
      }
    }
    return false;
  }

public:
  template <typename GeometryFilter, typename DataFunctor>
  void search(GeometryFilter&& filter, DataFunctor&& functor)
  {
    auto arg = [](leaf_type* leaf, size_type i) -> value_type& { return leaf->at(i); };
    walk<node_base_type*>(filter, functor, arg, _root, 0);
  }
  template <typename GeometryFilter, typename ConstDataFunctor>
  void search(GeometryFilter&& filter, ConstDataFunctor&& functor) const
  {
    auto arg = [](leaf_type const* leaf, size_type i) -> value_type const& { return leaf->at(i); };
    walk<node_base_type const*>(filter, functor, arg, _root, 0);
  }
  /// like search(), the functor receives an iterator (so it may erase / rebound later)
  template <typename GeometryFilter, typename ItFunctor>
  void search_iterator(GeometryFilter&& filter, ItFunctor&& functor)
  {
    auto arg = [](leaf_type* leaf, size_type i) { return iterator(leaf, i); };
    walk<node_base_type*>(filter, functor, arg, _root, 0);
  }
  template <typename GeometryFilter, typename ConstItFunctor>
  void search_iterator(GeometryFilter&& filter, ConstItFunctor&& functor) const
  {
    auto arg = [](leaf_type const* leaf, size_type i) { return const_iterator(leaf, i); };
    walk<node_base_type const*>(filter, functor, arg, _root, 0);
  }

protected:
  // ---- flatten ----------------------------------------------------------------------
  template <bool Move, typename NodePtr>
  size_type flatten_walk(flatten_result_t& out, NodePtr node, size_type parent, int level) const
  {
    const size_type me = static_cast<size_type>(out.nodes.size());
    const size_type first_slot = static_cast<size_type>(out.children.size());
    if (level == _leaf_level)
    {
      auto* leaf = node->as_leaf();
      out.nodes.push_back({ first_slot, leaf->size(), parent });
      for (size_type i = 0; i < leaf->size(); ++i)
      {
        out.children_bound.push_back(geometry_type(leaf->at(i).first));
        out.children.push_back(static_cast<size_type>(out.data.size()));
        if (Move)
          out.data.emplace_back(std::move(const_cast<mapped_type&>(leaf->at(i).second)));
        else
          out.data.emplace_back(leaf->at(i).second);
      }
      return me;
    }
    auto* inner = node->as_node();
    out.nodes.push_back({ first_slot, inner->size(), parent });
    // the whole child block is reserved BEFORE descending, so blocks are contiguous
    // in node-index order
    for (size_type i = 0; i < inner->size(); ++i)
    {
      out.children_bound.push_back(inner->at(i).first);
      out.children.push_back(0);
    }
    for (size_type i = 0; i < inner->size(); ++i)
    {
      NodePtr below = inner->at(i).second;
      const size_type child = flatten_walk<Move>(out, below, me, level + 1);
      out.children[first_slot + i] = child;
    }
    return me;
  }

public:
  /// linearise the tree into four dense host vectors (depth-first pre-order)
  template <bool Move = false>
  flatten_result_t flatten() const
  {
    static_assert(!Move, "use the non-const flatten_move() to move the mapped values out");
    flatten_result_t out;
    out.leaf_level = static_cast<size_type>(_leaf_level);
    out.root = 0;
    if (_root != nullptr)
      flatten_walk<false, node_base_type const*>(out, _root, 0, 0);
    return out;
  }
  /// same, moving the mapped values out of the tree (which keeps its keys)
  flatten_result_t flatten_move()
  {
    flatten_result_t out;
    out.leaf_level = static_cast<size_type>(_leaf_level);
    out.root = 0;
    if (_root != nullptr)
      flatten_walk<true, node_base_type*>(out, _root, 0, 0);
    return out;
  }
  /// const trees can only copy (the reference's const flatten_move, appendix B7)
  flatten_result_t flatten_move() const { return flatten<false>(); }

  // ---- flatten_device: the hand-off to the GPU ----------------------------------------
  /// flatten() re-laid as level-ordered, MAX_ENTRIES-wide SoA blocks, ready for
  /// rt_upload().  Hits are reported as indices into flatten().data ...
  device_tree_t flatten_device() const
  {
    return ::eh::rtree::flatten_device<MAX_ENTRIES, key_is_box()>(flatten(), id_source::data_index);
  }
  /// ... or, for integral mapped types of at most 32 bits, as the mapped values.
  device_tree_t flatten_device(id_source ids) const
  {
    return ::eh::rtree::flatten_device<MAX_ENTRIES, key_is_box()>(flatten(), ids);
  }
  /// whether keys are boxes (two corners) rather than points
  constexpr static bool key_is_box()
  {
    return detail::key_is_box<geometry_type, key_type>::value;
  }
};

} // namespace rtree
} // namespace eh
