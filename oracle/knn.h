/* oracle/knn.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * k-nearest-neighbour queries.  The reference has no kNN code; like the
 * ray queries it is RTree::search() (include/RTree/rtree.hpp:984-1019) driven by stateful
 * functors, which the API allows because functors are held by reference (rtree.hpp:1145):
 *
 *   state   : the k best (d2, id) pairs seen so far, "worst" = the largest of them in
 *             lexicographic (d2, id) order (d2 = +inf, id = 0xFFFFFFFF while fewer than k)
 *   filter  : descend into a child iff  dist2(q, child bound) <= worst.d2   (non-strict: ties
 *             survive, so the result does not depend on the traversal order)
 *   functor : for a key with (d2, id) < worst, replace the worst pair
 *   result  : the k pairs ascending by (d2, id); missing ones are (+inf, 0xFFFFFFFF)
 *
 * dist2(q, [lo, hi]) is defined operation for operation so every consumer rounds alike
 * (float32, round to nearest, axis 0 first, no fused multiply-add):
 *     e  = (q < lo) ? lo - q : ((q > hi) ? q - hi : 0)
 *     d2 = d2 + e * e
 * A point key has lo == hi.  NaN coordinates are outside the contract.
 */
#ifndef RT_ORACLE_KNN_H
#define RT_ORACLE_KNN_H

static inline float rt_knn_dist2(const float* lo, const float* hi, const float* q, int dim)
{
  float d2 = 0.0f;
  int d;
  for (d = 0; d < dim; ++d)
  {
    const float e = (q[d] < lo[d]) ? lo[d] - q[d] : ((q[d] > hi[d]) ? q[d] - hi[d] : 0.0f);
    const float sq = e * e;
    d2 = d2 + sq;
  }
  return d2;
}

/* (d2, id) < (wd2, wid) */
static inline int rt_knn_less(float d2, unsigned id, float wd2, unsigned wid)
{
  return d2 < wd2 || (d2 == wd2 && id < wid);
}

#endif
