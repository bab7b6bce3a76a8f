// rt_quant.cuh -- the quantised node image of the headline tree kind (3-D float32 points, W = 8).
//
// The exact traversal reads 256 bytes per internal node (two 128-byte rows of float corners) and
// spends six FSETP per child.  For this tree kind rt_upload also builds a second image in which
// EVERY slot of EVERY node -- child bounds of internal nodes and the points of leaves alike -- is one
// 16-byte record
//     A0 = G | q(hi_x) | q(hi_y) << 16     A1 = G | n(lo_x) | n(lo_y) << 16
//     A2 = G | q(hi_z) | n(lo_z) << 16     link (child block / 16, or the leaf entry's id)
// with q(v) = clamp(floor((v - lo_root) * scale) + 1, 1, 32767), n(v) = 32768 - q(v) and guard bits
// G = 0x80008000.  A query box is brought onto the same grid once,
//     B0 = q(qlo_x) | q(qlo_y) << 16   B1 = n(qhi_x) | n(qhi_y) << 16   B2 = q(qlo_z) | n(qhi_z) << 16
// and "all six interval tests pass" is  ((A0-B0) & (A1-B1) & (A2-B2) & G) == G : every half-word
// difference keeps its guard bit exactly when a >= b, and never borrows from its neighbour.
//
// q is monotone non-decreasing (round-to-nearest subtract and multiply by a non-negative scale,
// floor, clamp) and the SAME device function quantises node bounds and query bounds, so
//     hi >= qlo  implies  q(hi) >= q(qlo)      and      lo <= qhi  implies  q(lo) <= q(qhi):
// the quantised test passes whenever the reference's closed-overlap filter (reference
// example/sample/sample.cpp:48-54) passes.  It is a CONSERVATIVE filter: internal nodes may be
// visited that the exact filter would skip (about one grid cell per face: < 1 % more), and leaf
// slots that pass are candidates which are then tested exactly on the original float record
// (reference include/RTree/aabb.hpp:206-210).  Hit sets are therefore identical to the exact path.
// Empty slots carry a = 0 in every half (below every b >= 1) and can never pass.  A NaN query
// bound does not constrain (as in the reference's comparisons): lower bounds map to 1, upper
// bounds to 32767.  One node is one 128-byte line: half the L1 wavefronts of the exact layout.
#pragma once

#include "rt_common.cuh"

namespace rtb
{
constexpr uint32_t QUANT_GUARD = 0x80008000u;
constexpr uint32_t QUANT_BLOCK = 128; // bytes per node in the quantised image (W = 8 slots x 16 bytes)

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

// the three A words of a slot with corners lo / hi
__device__ __forceinline__ void quant_slot(const float (&lo)[3], const float (&hi)[3], const float (&glo)[3],
                                           const float (&gscale)[3], uint32_t (&a)[3])
{
  uint32_t qh[3], nl[3];
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    qh[d] = quant(hi[d], glo[d], gscale[d]);
    nl[d] = 32768u - quant(lo[d], glo[d], gscale[d]);
  }
  a[0] = QUANT_GUARD | qh[0] | (qh[1] << 16);
  a[1] = QUANT_GUARD | nl[0] | (nl[1] << 16);
  a[2] = QUANT_GUARD | qh[2] | (nl[2] << 16);
}

// the three B words of a query box
__device__ __forceinline__ void quant_query(const float (&qlo)[3], const float (&qhi)[3], const float (&glo)[3],
                                            const float (&gscale)[3], uint32_t (&b)[3])
{
  uint32_t ql[3], nh[3];
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    ql[d] = quant(qlo[d], glo[d], gscale[d]);
    nh[d] = 32768u - quant_upper(qhi[d], glo[d], gscale[d]);
  }
  b[0] = ql[0] | (ql[1] << 16);
  b[1] = nh[0] | (nh[1] << 16);
  b[2] = ql[2] | (nh[2] << 16);
}

__device__ __forceinline__ bool quant_pass(const uint4& r, const uint32_t (&b)[3])
{
  const uint32_t t = (r.x - b[0]) & (r.y - b[1]) & (r.z - b[2]);
  return (t & QUANT_GUARD) == QUANT_GUARD;
}
} // namespace rtb
