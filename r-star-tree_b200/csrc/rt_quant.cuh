// rt_quant.cuh -- the quantised node image of float32 point-keyed trees.
//
// The exact traversal reads two-corner float records (2 D + 1 words per slot of an internal node:
// 256 bytes per node in 3-D with W = 8, 1088 bytes in 8-D with W = 16) and spends two FSETP per
// axis and child.  For float32 point-keyed trees rt_upload also builds a second image in which EVERY
// slot of EVERY node -- child bounds of internal nodes and the points of leaves alike -- is a record of
// D + 1 words: 2 D conservative 15-bit grid coordinates, two per word, and the link,
//     halves  h = q(hi_0) .. q(hi_{D-1}), n(lo_0) .. n(lo_{D-1})          a_j = G | h_2j | h_2j+1 << 16
// with q(v) = clamp(floor((v - lo_root) * scale) + 1, 1, 32767), n(v) = 32768 - q(v) and guard bits
// G = 0x80008000 (D odd: the last half-word pairs q(hi_{D-1}) with n(lo_0), and so on -- the list is
// simply packed in order).  A query box is brought onto the same grid once,
//     halves  b = q(qlo_0) .. q(qlo_{D-1}), n(qhi_0) .. n(qhi_{D-1})      b_j = b_2j | b_2j+1 << 16
// and "all 2 D interval tests pass" is  (AND_j (a_j - b_j)) & G == G : every half-word difference
// keeps its guard bit exactly when a >= b, and never borrows from its neighbour.
//
// q is monotone non-decreasing (round-to-nearest subtract and multiply by a non-negative scale,
// floor, clamp) and the SAME device function quantises node bounds and query bounds, so
//     hi >= qlo  implies  q(hi) >= q(qlo)      and      lo <= qhi  implies  q(lo) <= q(qhi):
// the quantised test passes whenever the reference's closed-overlap filter (reference
// example/sample/sample.cpp:48-54) passes.  It is a CONSERVATIVE filter: internal nodes may be
// visited that the exact filter would skip (about one grid cell per face), and leaf slots that pass
// are candidates which are then tested exactly on the original float record (reference
// include/RTree/aabb.hpp:206-210).  Hit sets are therefore identical to the exact path.
// Empty slots carry a = 0 in every half (below every b >= 1) and can never pass.  A NaN query
// bound does not constrain (as in the reference's comparisons): lower bounds map to 1, upper
// bounds to 32767.  In 3-D a node is one 128-byte line instead of two; in 8-D (W = 16) 576 bytes
// instead of 1088.  A point leaf's quantised block has the stride of its float block (D + 1 words
// per slot in both), so the float record behind a candidate sits at a constant distance.
//
// 8-D leaves are stored PACKED: a point's record would say the same thing in its q-halves and its
// n-halves, so only s_j = q(p_2j) | q(p_2j+1) << 16 (j < 4) is kept, in the first 16-byte row; the
// traversal derives G | s_j and 0x00010000 - s_j (= G | n(p) in each half) in registers and never
// fetches the second row of a leaf (rt_traverse_box.cuh).  An empty leaf slot is s-half 0x8000.
#pragma once

#include "rt_common.cuh"

namespace rtb
{
constexpr uint32_t QUANT_GUARD = 0x80008000u;

__device__ __forceinline__ uint32_t quant(float v, float lo, float scale)
{
  float t = floorf(__fmul_rn(__fsub_rn(v, lo), scale)) + 1.0f;
  t = fminf(fmaxf(t, 1.0f), 32767.0f); // NaN -> 1
  return (uint32_t)t;
}
// upper query bound: NaN must not constrain
__device__ __forceinline__ uint32_t quant_upper(float v, float lo, float scale)
{
  return (v != v) ? 32767u : quant(v, lo, scale);
}

// 2 D half-words, in order, two per word
template <int D>
__device__ __forceinline__ void quant_pack(const uint32_t (&h)[2 * D], uint32_t guard, uint32_t (&w)[D])
{
#pragma unroll
  for (int j = 0; j < D; ++j)
    w[j] = guard | h[2 * j] | (h[2 * j + 1] << 16);
}

// the D words of a slot with corners lo / hi
template <int D>
__device__ __forceinline__ void quant_slot(const float (&lo)[D], const float (&hi)[D], const float* glo,
                                           const float* gscale, uint32_t (&a)[D])
{
  uint32_t h[2 * D];
#pragma unroll
  for (int d = 0; d < D; ++d)
  {
    // a NaN bound must not constrain (the reference's comparisons are all false on NaN): an upper
    // bound goes to the top of the grid, a lower bound to the bottom
    h[d] = quant_upper(hi[d], glo[d], gscale[d]);
    h[D + d] = 32768u - quant(lo[d], glo[d], gscale[d]);
  }
  quant_pack<D>(h, QUANT_GUARD, a);
}

// the D words of a query box
template <int D>
__device__ __forceinline__ void quant_query(const float (&qlo)[D], const float (&qhi)[D], const float* glo,
                                            const float* gscale, uint32_t (&b)[D])
{
  uint32_t h[2 * D];
#pragma unroll
  for (int d = 0; d < D; ++d)
  {
    h[d] = quant(qlo[d], glo[d], gscale[d]);
    h[D + d] = 32768u - quant_upper(qhi[d], glo[d], gscale[d]);
  }
  quant_pack<D>(h, 0u, b);
}

// the guard bits that survive: all of them (== QUANT_GUARD) iff the slot passes
template <int D>
__device__ __forceinline__ uint32_t quant_pass_bits(const uint32_t* a, const uint32_t (&b)[D])
{
  uint32_t t = a[0] - b[0];
#pragma unroll
  for (int j = 1; j < D; ++j)
    t &= a[j] - b[j];
  return t & QUANT_GUARD;
}
template <int D>
__device__ __forceinline__ bool quant_pass(const uint32_t* a, const uint32_t (&b)[D])
{
  return quant_pass_bits<D>(a, b) == QUANT_GUARD;
}

// Strictly inside on the grid: every half-word of a exceeds its half-word of b (a >= b + 1).  For a
// leaf slot (a from the point p, b from the query box) this means q(p) > q(qlo) and q(p) < q(qhi) on
// every axis, and because q is monotone non-decreasing that implies qlo < p < qhi: the exact test of
// the reference (include/RTree/aabb.hpp:206-210) is known to pass without looking at the floats.
// Only candidates in a boundary cell of the query (about 2 % of them) need the float record.
template <int D>
__device__ __forceinline__ bool quant_strictly_inside(const uint32_t* a, const uint32_t (&b)[D])
{
  uint32_t t = a[0] - b[0] - 0x00010001u;
#pragma unroll
  for (int j = 1; j < D; ++j)
    t &= a[j] - b[j] - 0x00010001u;
  return (t & QUANT_GUARD) == QUANT_GUARD;
}

// byte offset of word j of slot c in a block of `words`-word records (layout: include/rtree_b200.h)
__host__ __device__ __forceinline__ uint32_t block_word_offset(uint32_t width, uint32_t words, uint32_t c, uint32_t j)
{
  const uint32_t full = words / 4, r = words % 4;
  if (j < 4 * full)
    return (j / 4) * width * 16 + c * 16 + (j % 4) * 4;
  const uint32_t tail = r == 1 ? 4u : (r == 2 ? 8u : 16u);
  return full * width * 16 + c * tail + (j - 4 * full) * 4;
}
} // namespace rtb
