// RTree/geometry.hpp -- host side (header-only C++17) of rtree-b200.
//
// The geometry customisation point of the reference API, written from scratch:
//   geometry_traits<G>            (reference: include/RTree/geometry_traits.hpp:13-41)
//   helper::{min_point,...,distance_center}  (reference: geometry_traits.hpp:43-178)
//   point_t<T,Dim>, aabb_t<P>     (reference: include/RTree/aabb.hpp:12-122)
//   traits of the stock types     (reference: include/RTree/aabb.hpp:124-316)
//
// Same names, argument meaning and results as the reference so user code and custom
// traits port unchanged.  Two deliberate differences, both where the reference does
// not compile: the N-D is_inside(aabb, aabb) below really
// tests containment.
#pragma once

#include <algorithm>
#include <cstdint>
#include <iostream>
#include <type_traits>

#ifndef NDEBUG
#ifndef EH_RTREE_ASSERT
#define EH_RTREE_ASSERT(cond, msg)                                                        \
  do                                                                                      \
  {                                                                                       \
    if (!(cond))                                                                          \
      std::cerr << "Assertion failed: " << msg << "\nfile " << __FILE__ << "\nline "      \
                << __LINE__ << std::endl;                                                 \
  } while (0)
#endif
#ifndef EH_RTREE_ASSERT_SILENT
#define EH_RTREE_ASSERT_SILENT(cond) EH_RTREE_ASSERT(cond, "")
#endif
#else
#ifndef EH_RTREE_ASSERT
#define EH_RTREE_ASSERT(cond, msg) ((void)0)
#endif
#ifndef EH_RTREE_ASSERT_SILENT
#define EH_RTREE_ASSERT_SILENT(cond) ((void)0)
#endif
#endif

namespace eh
{
namespace rtree
{

// every index in a flattened tree is 32 bit (reference: include/RTree/global.hpp:35)
using size_type = std::uint32_t;

// ---------------------------------------------------------------------------------
// Customisation point.  The primary template treats G as a 1-D arithmetic value.
// ---------------------------------------------------------------------------------
template <typename G>
struct geometry_traits
{
  constexpr static int DIM = 1;
  using scalar_type = G;
  static scalar_type min_point(G const& g, int) { return g; }
  static scalar_type max_point(G const& g, int) { return g; }
  static void set_min_point(G& g, int, scalar_type v) { g = v; }
  static void set_max_point(G& g, int, scalar_type v) { g = v; }
};

namespace helper
{
template <typename G>
inline typename geometry_traits<G>::scalar_type min_point(G const& g, int axis)
{
  return geometry_traits<G>::min_point(g, axis);
}
template <typename G>
inline typename geometry_traits<G>::scalar_type max_point(G const& g, int axis)
{
  return geometry_traits<G>::max_point(g, axis);
}
template <typename G>
inline void set_min_point(G& g, int axis, typename geometry_traits<G>::scalar_type const& v)
{
  geometry_traits<G>::set_min_point(g, axis, v);
}
template <typename G>
inline void set_max_point(G& g, int axis, typename geometry_traits<G>::scalar_type const& v)
{
  geometry_traits<G>::set_max_point(g, axis, v);
}

/// grow `a` until it covers `b`
template <typename A, typename B>
inline void enlarge_to(A& a, B const& b)
{
  static_assert(geometry_traits<A>::DIM == geometry_traits<B>::DIM, "Dimension not match");
  for (int d = 0; d < geometry_traits<A>::DIM; ++d)
  {
    const auto lo = std::min(min_point(a, d), min_point(b, d));
    const auto hi = std::max(max_point(a, d), max_point(b, d));
    set_min_point(a, d, lo);
    set_max_point(a, d, hi);
  }
}
/// area (hyper-volume) of the union box of a and b
template <typename A, typename B>
inline typename geometry_traits<A>::scalar_type enlarged_area(A const& a, B const& b)
{
  static_assert(geometry_traits<A>::DIM == geometry_traits<B>::DIM, "Dimension not match");
  typename geometry_traits<A>::scalar_type v = 1;
  for (int d = 0; d < geometry_traits<A>::DIM; ++d)
  {
    v *= std::max(max_point(a, d), max_point(b, d)) - std::min(min_point(a, d), min_point(b, d));
  }
  return v;
}
/// area of the intersection, 0 when the boxes are disjoint on any axis
template <typename A, typename B>
inline typename geometry_traits<A>::scalar_type intersection_area(A const& a, B const& b)
{
  static_assert(geometry_traits<A>::DIM == geometry_traits<B>::DIM, "Dimension not match");
  typename geometry_traits<A>::scalar_type v = 1;
  for (int d = 0; d < geometry_traits<A>::DIM; ++d)
  {
    if (min_point(a, d) > max_point(b, d) || max_point(a, d) < min_point(b, d))
    {
      return 0;
    }
    v *= std::min(max_point(a, d), max_point(b, d)) - std::max(min_point(a, d), min_point(b, d));
  }
  return v;
}
template <typename G>
inline typename geometry_traits<G>::scalar_type area(G const& g)
{
  typename geometry_traits<G>::scalar_type v = 1;
  for (int d = 0; d < geometry_traits<G>::DIM; ++d)
  {
    v *= max_point(g, d) - min_point(g, d);
  }
  return v;
}
/// sum of the edge lengths
template <typename G>
inline typename geometry_traits<G>::scalar_type margin(G const& g)
{
  typename geometry_traits<G>::scalar_type v = 0;
  for (int d = 0; d < geometry_traits<G>::DIM; ++d)
  {
    v += max_point(g, d) - min_point(g, d);
  }
  return v;
}
/// squared distance between (twice the) centres; only used to ORDER entries for the
/// forced reinsert, so neither the factor 2 nor a sqrt matters
template <typename A, typename B>
inline typename geometry_traits<A>::scalar_type distance_center(A const& a, B const& b)
{
  static_assert(geometry_traits<A>::DIM == geometry_traits<B>::DIM, "Dimension not match");
  typename geometry_traits<A>::scalar_type v = 0;
  for (int d = 0; d < geometry_traits<A>::DIM; ++d)
  {
    const typename geometry_traits<A>::scalar_type diff
        = (max_point(a, d) + min_point(a, d)) - (max_point(b, d) + min_point(b, d));
    v += diff * diff;
  }
  return v;
}
} // namespace helper

// ---------------------------------------------------------------------------------
// Stock N-D point and axis-aligned box.
// ---------------------------------------------------------------------------------
template <typename ScalarType, unsigned int Dim>
class point_t
{
public:
  using size_type = unsigned int;
  using scalar_type = ScalarType;
  using value_type = ScalarType;

protected:
  scalar_type _data[Dim];

public:
  point_t() = default;
  point_t(point_t const&) = default;
  point_t& operator=(point_t const&) = default;

  /// point_t<float,3> p{x, y, z};
  template <typename T0, typename... Ts,
            typename = typename std::enable_if<std::is_convertible<T0, ScalarType>::value>::type>
  point_t(T0 first, Ts... rest)
  {
    static_assert(sizeof...(rest) + 1 == Dim, "Constructor Dimension not match");
    const scalar_type tmp[Dim] = { static_cast<scalar_type>(first), static_cast<scalar_type>(rest)... };
    for (size_type i = 0; i < Dim; ++i)
    {
      _data[i] = tmp[i];
    }
  }
  template <typename Iterator>
  void assign(Iterator first, Iterator last)
  {
    for (size_type i = 0; first != last && i < Dim; ++i, ++first)
    {
      _data[i] = *first;
    }
  }
  constexpr static size_type size() { return Dim; }
  scalar_type& operator[](size_type i) { return _data[i]; }
  scalar_type operator[](size_type i) const { return _data[i]; }
  scalar_type* data() { return _data; }
  scalar_type const* data() const { return _data; }
  scalar_type* begin() { return _data; }
  scalar_type const* begin() const { return _data; }
  scalar_type* end() { return _data + Dim; }
  scalar_type const* end() const { return _data + Dim; }
};

/// {min_, max_}; implicitly constructible from a point (a degenerate box), which is
/// what lets flatten() store point keys in children_bound
template <typename PointType>
struct aabb_t
{
  PointType min_, max_;

  aabb_t(PointType const& p)
      : min_(p)
      , max_(p)
  {
  }
  aabb_t(PointType const& lo, PointType const& hi)
      : min_(lo)
      , max_(hi)
  {
  }
};

// 1-D box over a plain arithmetic type: closed intervals everywhere
// (reference: include/RTree/aabb.hpp:124-196)
template <typename ArithmeticType>
struct geometry_traits<aabb_t<ArithmeticType>>
{
  using scalar_type = ArithmeticType;
  using AABB = aabb_t<ArithmeticType>;
  constexpr static int DIM = 1;

  template <typename P>
  static bool is_inside(AABB const& box, P const& p)
  {
    return box.min_ <= p && p <= box.max_;
  }
  template <typename P>
  static bool is_inside(AABB const& box, aabb_t<P> const& inner)
  {
    return box.min_ <= inner.min_ && inner.max_ <= box.max_;
  }
  template <typename P>
  static bool is_overlap(AABB const& box, P const& p)
  {
    return is_inside(box, p);
  }
  template <typename P>
  static bool is_overlap(AABB const& a, aabb_t<P> const& b)
  {
    return !(b.min_ > a.max_) && !(a.min_ > b.max_);
  }
  static AABB merge(AABB const& a, ArithmeticType p)
  {
    return { std::min(a.min_, p), std::max(a.max_, p) };
  }
  static AABB merge(AABB const& a, AABB const& b)
  {
    return { std::min(a.min_, b.min_), std::max(a.max_, b.max_) };
  }
  static scalar_type area(AABB const& a) { return a.max_ - a.min_; }

  static scalar_type min_point(AABB const& b, int) { return b.min_; }
  static scalar_type max_point(AABB const& b, int) { return b.max_; }
  static void set_min_point(AABB& b, int, scalar_type v) { b.min_ = v; }
  static void set_max_point(AABB& b, int, scalar_type v) { b.max_ = v; }
};

// N-D box over point_t (reference: include/RTree/aabb.hpp:198-297).
// is_overlap(aabb, aabb) keeps the reference's STRICT semantics (touching boxes do
// not overlap); the query engine never uses it -- the device
// filter is the closed form of the reference's sample / test.
template <typename T, unsigned int Dim>
struct geometry_traits<aabb_t<point_t<T, Dim>>>
{
  using scalar_type = T;
  using Point = point_t<T, Dim>;
  using AABB = aabb_t<Point>;
  constexpr static int DIM = Dim;

  /// every coordinate strictly smaller
  static bool less_than(Point const& a, Point const& b)
  {
    for (unsigned int d = 0; d < Dim; ++d)
    {
      if (a[d] >= b[d])
        return false;
    }
    return true;
  }
  /// no coordinate greater
  static bool less_equal(Point const& a, Point const& b)
  {
    for (unsigned int d = 0; d < Dim; ++d)
    {
      if (a[d] > b[d])
        return false;
    }
    return true;
  }
  static bool is_inside(AABB const& box, Point const& p)
  {
    return less_equal(box.min_, p) && less_equal(p, box.max_);
  }
  static bool is_inside(AABB const& box, AABB const& inner)
  {
    return less_equal(box.min_, inner.min_) && less_equal(inner.max_, box.max_);
  }
  static bool is_overlap(AABB const& box, Point const& p) { return is_inside(box, p); }
  static bool is_overlap(AABB const& a, AABB const& b)
  {
    return less_than(b.min_, a.max_) && less_than(a.min_, b.max_);
  }
  static Point min(Point const& a, Point const& b)
  {
    Point r;
    for (unsigned int d = 0; d < Dim; ++d)
      r[d] = std::min(a[d], b[d]);
    return r;
  }
  static Point max(Point const& a, Point const& b)
  {
    Point r;
    for (unsigned int d = 0; d < Dim; ++d)
      r[d] = std::max(a[d], b[d]);
    return r;
  }

  static scalar_type min_point(AABB const& b, int axis) { return b.min_[axis]; }
  static scalar_type max_point(AABB const& b, int axis) { return b.max_[axis]; }
  static void set_min_point(AABB& b, int axis, T v) { b.min_[axis] = v; }
  static void set_max_point(AABB& b, int axis, T v) { b.max_[axis] = v; }
};

// a point used as a key is its own min and max corner
template <typename T, unsigned int Dim>
struct geometry_traits<point_t<T, Dim>>
{
  using scalar_type = T;
  constexpr static int DIM = Dim;
  static scalar_type min_point(point_t<T, Dim> const& p, int axis) { return p[axis]; }
  static scalar_type max_point(point_t<T, Dim> const& p, int axis) { return p[axis]; }
};

} // namespace rtree
} // namespace eh
