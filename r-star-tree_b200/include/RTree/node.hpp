// RTree/node.hpp -- host side (header-only C++17) of rtree-b200.
//
// In-place node storage and the iterators over it, written from scratch with the
// reference's public names so user code ports unchanged:
//   static_vector<T,N>                       (reference: include/RTree/static_vector.hpp)
//   static_node_base_t / static_node_t / static_leaf_node_t
//                                            (reference: include/RTree/static_node.hpp)
//   iterator_t / node_iterator_t             (reference: include/RTree/iterator.hpp)
//
// This is the pointer-linked BUILD-side structure.  Queries on the GPU never touch
// it: flatten() + flatten_device() (rtree.hpp, device_layout.hpp) re-lay it out.
#pragma once

#include <cstddef>
#include <iterator>
#include <new>
#include <type_traits>
#include <utility>

#include "geometry.hpp"

namespace eh
{
namespace rtree
{

// ---------------------------------------------------------------------------------
// static_vector: at most N elements constructed inside the object, no heap.
// ---------------------------------------------------------------------------------
template <typename T, size_type N>
class static_vector
{
public:
  using value_type = T;
  using size_type = ::eh::rtree::size_type;
  using iterator = T*;
  using const_iterator = T const*;
  using reference = T&;
  using const_reference = T const&;

private:
  alignas(T) unsigned char _raw[sizeof(T) * N];
  size_type _count = 0;

  T* slot(size_type i) { return reinterpret_cast<T*>(_raw) + i; }
  T const* slot(size_type i) const { return reinterpret_cast<T const*>(_raw) + i; }

public:
  static_vector() = default;
  static_vector(static_vector const& rhs)
  {
    for (size_type i = 0; i < rhs._count; ++i)
      ::new (slot(i)) T(rhs[i]);
    _count = rhs._count;
  }
  static_vector(static_vector&& rhs) noexcept(std::is_nothrow_move_constructible<T>::value)
  {
    for (size_type i = 0; i < rhs._count; ++i)
      ::new (slot(i)) T(std::move(rhs[i]));
    _count = rhs._count;
    rhs.clear();
  }
  static_vector& operator=(static_vector const& rhs)
  {
    if (this != &rhs)
    {
      clear();
      for (size_type i = 0; i < rhs._count; ++i)
        ::new (slot(i)) T(rhs[i]);
      _count = rhs._count;
    }
    return *this;
  }
  static_vector& operator=(static_vector&& rhs) noexcept(std::is_nothrow_move_constructible<T>::value)
  {
    if (this != &rhs)
    {
      clear();
      for (size_type i = 0; i < rhs._count; ++i)
        ::new (slot(i)) T(std::move(rhs[i]));
      _count = rhs._count;
      rhs.clear();
    }
    return *this;
  }
  ~static_vector() { clear(); }

  constexpr static size_type capacity() { return N; }
  size_type size() const { return _count; }
  bool empty() const { return _count == 0; }

  template <typename... Args>
  T& emplace_back(Args&&... args)
  {
    EH_RTREE_ASSERT_SILENT(_count < N);
    T* p = ::new (slot(_count)) T(std::forward<Args>(args)...);
    ++_count;
    return *p;
  }
  void push_back(T const& v) { emplace_back(v); }
  void push_back(T&& v) { emplace_back(std::move(v)); }
  void pop_back()
  {
    EH_RTREE_ASSERT_SILENT(_count > 0);
    --_count;
    slot(_count)->~T();
  }
  void clear()
  {
    while (_count > 0)
      pop_back();
  }
  /// order-preserving removal
  iterator erase(const_iterator pos)
  {
    const size_type at = static_cast<size_type>(pos - begin());
    for (size_type i = at; i + 1 < _count; ++i)
      *slot(i) = std::move(*slot(i + 1));
    pop_back();
    return begin() + at;
  }

  T& operator[](size_type i) { return *slot(i); }
  T const& operator[](size_type i) const { return *slot(i); }
  T& at(size_type i) { return *slot(i); }
  T const& at(size_type i) const { return *slot(i); }
  T& front() { return *slot(0); }
  T const& front() const { return *slot(0); }
  T& back() { return *slot(_count - 1); }
  T const& back() const { return *slot(_count - 1); }
  T* data() { return slot(0); }
  T const* data() const { return slot(0); }

  iterator begin() { return slot(0); }
  const_iterator begin() const { return slot(0); }
  const_iterator cbegin() const { return slot(0); }
  iterator end() { return slot(_count); }
  const_iterator end() const { return slot(_count); }
  const_iterator cend() const { return slot(_count); }
};

// ---------------------------------------------------------------------------------
// Nodes.  Internal nodes hold (bound, child*), leaves hold (key, mapped).  Only the
// tree knows a node's level, so both kinds share one base and are reached through
// as_node() / as_leaf().
// ---------------------------------------------------------------------------------
template <typename G, typename K, typename V, size_type MinE, size_type MaxE>
struct static_node_t;
template <typename G, typename K, typename V, size_type MinE, size_type MaxE>
struct static_leaf_node_t;

template <typename G, typename K, typename V, size_type MinE, size_type MaxE>
struct static_node_base_t
{
  using node_base_type = static_node_base_t;
  using node_type = static_node_t<G, K, V, MinE, MaxE>;
  using leaf_type = static_leaf_node_t<G, K, V, MinE, MaxE>;
  using size_type = ::eh::rtree::size_type;
  using geometry_type = G;
  using key_type = K;
  using mapped_type = V;

  constexpr static size_type MIN_ENTRIES = MinE;
  constexpr static size_type MAX_ENTRIES = MaxE;

  node_type* _parent = nullptr;
  size_type _index_on_parent = 0;

  node_type* parent() const { return _parent; }
  bool is_root() const { return _parent == nullptr; }

  node_type* as_node() { return static_cast<node_type*>(this); }
  node_type const* as_node() const { return static_cast<node_type const*>(this); }
  leaf_type* as_leaf() { return static_cast<leaf_type*>(this); }
  leaf_type const* as_leaf() const { return static_cast<leaf_type const*>(this); }

  /// number of ancestors (root = 0); O(height)
  int level_recursive() const
  {
    int level = 0;
    for (node_type const* p = _parent; p != nullptr; p = p->_parent)
      ++level;
    return level;
  }

  /// right neighbour on the same level (crossing parents), nullptr at the end
  node_base_type* next();
  node_base_type const* next() const
  {
    return const_cast<static_node_base_t*>(this)->next();
  }
  /// left neighbour on the same level, nullptr at the beginning
  node_base_type* prev();
  node_base_type const* prev() const
  {
    return const_cast<static_node_base_t*>(this)->prev();
  }
};

template <typename G, typename K, typename V, size_type MinE, size_type MaxE>
struct static_node_t : static_node_base_t<G, K, V, MinE, MaxE>
{
  using base = static_node_base_t<G, K, V, MinE, MaxE>;
  using typename base::geometry_type;
  using typename base::node_base_type;
  using typename base::size_type;
  using value_type = std::pair<geometry_type, node_base_type*>;
  using storage_type = static_vector<value_type, MaxE>;
  using iterator = typename storage_type::iterator;
  using const_iterator = typename storage_type::const_iterator;

  storage_type _child;

  size_type size() const { return _child.size(); }
  bool empty() const { return _child.empty(); }
  value_type& at(size_type i) { return _child[i]; }
  value_type const& at(size_type i) const { return _child[i]; }
  value_type& operator[](size_type i) { return _child[i]; }
  value_type const& operator[](size_type i) const { return _child[i]; }
  value_type& front() { return _child.front(); }
  value_type& back() { return _child.back(); }
  iterator begin() { return _child.begin(); }
  const_iterator begin() const { return _child.begin(); }
  iterator end() { return _child.end(); }
  const_iterator end() const { return _child.end(); }

  /// append a child and adopt it
  void insert(value_type entry)
  {
    entry.second->_parent = this;
    entry.second->_index_on_parent = _child.size();
    _child.push_back(std::move(entry));
  }
  /// drop the links only; the children stay alive
  void clear() { _child.clear(); }
  void pop_back() { _child.pop_back(); }
  void swap(size_type a, size_type b)
  {
    if (a == b)
      return;
    std::swap(_child[a], _child[b]);
    _child[a].second->_index_on_parent = a;
    _child[b].second->_index_on_parent = b;
  }
  /// O(1) unlink: the last child moves into the hole
  void erase(iterator pos)
  {
    const size_type at = static_cast<size_type>(pos - begin());
    if (at + 1 != _child.size())
    {
      _child[at] = std::move(_child.back());
      _child[at].second->_index_on_parent = at;
    }
    _child.pop_back();
  }
  void erase(node_base_type* child) { erase(begin() + child->_index_on_parent); }

  geometry_type calculate_bound() const
  {
    geometry_type bound = _child[0].first;
    for (size_type i = 1; i < _child.size(); ++i)
      helper::enlarge_to(bound, _child[i].first);
    return bound;
  }

  /// number of values below this node, which sits `leaf_level` levels above the leaves
  size_type size_recursive(int leaf_level) const
  {
    size_type total = 0;
    for (value_type const& c : _child)
    {
      total += (leaf_level == 1) ? c.second->as_leaf()->size()
                                 : c.second->as_node()->size_recursive(leaf_level - 1);
    }
    return total;
  }
  /// destroy the whole subtree below (not this node itself)
  template <typename Tree>
  void delete_recursive(int leaf_level, Tree& tree)
  {
    for (value_type& c : _child)
    {
      if (leaf_level == 1)
      {
        tree.destroy_node(c.second->as_leaf());
      }
      else
      {
        c.second->as_node()->delete_recursive(leaf_level - 1, tree);
        tree.destroy_node(c.second->as_node());
      }
    }
    _child.clear();
  }
  /// deep copy of the subtree rooted here
  template <typename Tree>
  static_node_t* clone_recursive(int leaf_level, Tree& tree) const
  {
    static_node_t* copy = tree.template construct_node<static_node_t>();
    for (value_type const& c : _child)
    {
      if (leaf_level == 1)
        copy->insert({ c.first, c.second->as_leaf()->clone_recursive(tree) });
      else
        copy->insert({ c.first, c.second->as_node()->clone_recursive(leaf_level - 1, tree) });
    }
    return copy;
  }
};

template <typename G, typename K, typename V, size_type MinE, size_type MaxE>
struct static_leaf_node_t : static_node_base_t<G, K, V, MinE, MaxE>
{
  using base = static_node_base_t<G, K, V, MinE, MaxE>;
  using typename base::geometry_type;
  using typename base::key_type;
  using typename base::mapped_type;
  using typename base::size_type;
  using value_type = std::pair<key_type, mapped_type>;
  using storage_type = static_vector<value_type, MaxE>;
  using iterator = typename storage_type::iterator;
  using const_iterator = typename storage_type::const_iterator;

  storage_type _child;

  size_type size() const { return _child.size(); }
  bool empty() const { return _child.empty(); }
  value_type& at(size_type i) { return _child[i]; }
  value_type const& at(size_type i) const { return _child[i]; }
  value_type& operator[](size_type i) { return _child[i]; }
  value_type const& operator[](size_type i) const { return _child[i]; }
  value_type& front() { return _child.front(); }
  value_type& back() { return _child.back(); }
  iterator begin() { return _child.begin(); }
  const_iterator begin() const { return _child.begin(); }
  iterator end() { return _child.end(); }
  const_iterator end() const { return _child.end(); }

  void insert(value_type entry) { _child.push_back(std::move(entry)); }
  void clear() { _child.clear(); }
  void pop_back() { _child.pop_back(); }
  void swap(size_type a, size_type b)
  {
    if (a != b)
      std::swap(_child[a], _child[b]);
  }
  void erase(iterator pos)
  {
    const size_type at = static_cast<size_type>(pos - begin());
    if (at + 1 != _child.size())
      _child[at] = std::move(_child.back());
    _child.pop_back();
  }

  geometry_type calculate_bound() const
  {
    geometry_type bound(_child[0].first);
    for (size_type i = 1; i < _child.size(); ++i)
      helper::enlarge_to(bound, _child[i].first);
    return bound;
  }
  template <typename Tree>
  void delete_recursive(Tree&)
  {
    _child.clear();
  }
  template <typename Tree>
  static_leaf_node_t* clone_recursive(Tree& tree) const
  {
    static_leaf_node_t* copy = tree.template construct_node<static_leaf_node_t>();
    for (value_type const& v : _child)
      copy->insert(v);
    return copy;
  }
};

template <typename G, typename K, typename V, size_type MinE, size_type MaxE>
typename static_node_base_t<G, K, V, MinE, MaxE>::node_base_type*
static_node_base_t<G, K, V, MinE, MaxE>::next()
{
  // climb until some ancestor has a right sibling, step right, come back down-left
  int climbed = 0;
  node_base_type* n = this;
  while (n->_parent != nullptr && n->_index_on_parent + 1 == n->_parent->size())
  {
    n = n->_parent;
    ++climbed;
  }
  if (n->_parent == nullptr)
    return nullptr;
  n = n->_parent->at(n->_index_on_parent + 1).second;
  for (; climbed > 0; --climbed)
    n = n->as_node()->at(0).second;
  return n;
}
template <typename G, typename K, typename V, size_type MinE, size_type MaxE>
typename static_node_base_t<G, K, V, MinE, MaxE>::node_base_type*
static_node_base_t<G, K, V, MinE, MaxE>::prev()
{
  int climbed = 0;
  node_base_type* n = this;
  while (n->_parent != nullptr && n->_index_on_parent == 0)
  {
    n = n->_parent;
    ++climbed;
  }
  if (n->_parent == nullptr)
    return nullptr;
  n = n->_parent->at(n->_index_on_parent - 1).second;
  for (; climbed > 0; --climbed)
    n = n->as_node()->at(n->as_node()->size() - 1).second;
  return n;
}

// ---------------------------------------------------------------------------------
// Iterators.
// ---------------------------------------------------------------------------------

/// bidirectional iterator over the values (key, mapped) of all leaves, left to right
template <typename LeafType>
class iterator_t
{
public:
  using leaf_type = LeafType;
  using value_type = typename std::conditional<std::is_const<LeafType>::value,
                                               typename LeafType::value_type const,
                                               typename LeafType::value_type>::type;
  using reference = value_type&;
  using pointer = value_type*;
  using difference_type = std::ptrdiff_t;
  using iterator_category = std::bidirectional_iterator_tag;

  leaf_type* _leaf = nullptr;
  size_type _index = 0;

  iterator_t() = default;
  iterator_t(leaf_type* leaf, size_type index)
      : _leaf(leaf)
      , _index(index)
  {
  }
  /// (value pointer, owning leaf): the reference's constructor shape
  iterator_t(pointer value, leaf_type* leaf)
      : _leaf(leaf)
      , _index(static_cast<size_type>(value - &leaf->at(0)))
  {
  }
  /// iterator -> const_iterator
  template <typename OtherLeaf,
            typename = typename std::enable_if<std::is_same<OtherLeaf const, LeafType>::value>::type>
  iterator_t(iterator_t<OtherLeaf> const& rhs)
      : _leaf(rhs._leaf)
      , _index(rhs._index)
  {
  }

  reference operator*() const { return _leaf->at(_index); }
  pointer operator->() const { return &_leaf->at(_index); }
  leaf_type* leaf() const { return _leaf; }

  iterator_t& operator++()
  {
    if (_index + 1 < _leaf->size())
    {
      ++_index;
      return *this;
    }
    auto* n = _leaf->next();
    while (n != nullptr && n->as_leaf()->empty())
      n = n->next();
    _leaf = n ? const_cast<leaf_type*>(static_cast<LeafType const*>(n->as_leaf())) : nullptr;
    _index = 0;
    return *this;
  }
  iterator_t operator++(int)
  {
    iterator_t old = *this;
    ++*this;
    return old;
  }
  iterator_t& operator--()
  {
    if (_index > 0)
    {
      --_index;
      return *this;
    }
    auto* n = _leaf->prev();
    while (n != nullptr && n->as_leaf()->empty())
      n = n->prev();
    _leaf = n ? const_cast<leaf_type*>(static_cast<LeafType const*>(n->as_leaf())) : nullptr;
    _index = _leaf ? _leaf->size() - 1 : 0;
    return *this;
  }
  iterator_t operator--(int)
  {
    iterator_t old = *this;
    --*this;
    return old;
  }
  template <typename L>
  bool operator==(iterator_t<L> const& rhs) const
  {
    return _leaf == rhs._leaf && _index == rhs._index;
  }
  template <typename L>
  bool operator!=(iterator_t<L> const& rhs) const
  {
    return !(*this == rhs);
  }
};

/// bidirectional iterator over the nodes of ONE level; dereferences to NodeType*
template <typename NodeType>
class node_iterator_t
{
public:
  using node_type = NodeType;
  using value_type = node_type*;
  using reference = value_type;
  using pointer = value_type;
  using difference_type = std::ptrdiff_t;
  using iterator_category = std::bidirectional_iterator_tag;

  node_type* _node = nullptr;

  node_iterator_t() = default;
  node_iterator_t(node_type* node)
      : _node(node)
  {
  }
  template <typename Other,
            typename = typename std::enable_if<std::is_same<Other const, NodeType>::value>::type>
  node_iterator_t(node_iterator_t<Other> const& rhs)
      : _node(rhs._node)
  {
  }

  value_type operator*() const { return _node; }
  value_type operator->() const { return _node; }

  node_iterator_t& operator++()
  {
    auto* n = _node->next();
    _node = n ? const_cast<node_type*>(static_cast<NodeType const*>(n)) : nullptr;
    return *this;
  }
  node_iterator_t operator++(int)
  {
    node_iterator_t old = *this;
    ++*this;
    return old;
  }
  node_iterator_t& operator--()
  {
    auto* n = _node->prev();
    _node = n ? const_cast<node_type*>(static_cast<NodeType const*>(n)) : nullptr;
    return *this;
  }
  node_iterator_t operator--(int)
  {
    node_iterator_t old = *this;
    --*this;
    return old;
  }
  template <typename N>
  bool operator==(node_iterator_t<N> const& rhs) const
  {
    return _node == rhs._node;
  }
  template <typename N>
  bool operator!=(node_iterator_t<N> const& rhs) const
  {
    return _node != rhs._node;
  }
};

} // namespace rtree
} // namespace eh
