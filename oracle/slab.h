/* oracle/slab.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The ray/AABB slab test of SURVEY.md section 8 (a12).  The reference has no ray
 * code at all (only a screenshot, example/teapot_diffusive_reflection.png), so this
 * header IS the definition the reference's search() is driven with in
 * oracle/ref_harness.cpp and that oracle/oracle.c restates on the flat form.
 *
 *   ray  = { o[3], inv_d[3], tmax },  inv_d computed ON THE HOST as 1.0f/d so every
 *          consumer sees the same bits;  tmin = 0
 *   axis : a = (lo - o) * inv ;  c = (hi - o) * inv
 *          near = (c < a) ? c : a ;  far = (c < a) ? a : c
 *          t0 = (near > t0) ? near : t0 ;  t1 = (far < t1) ? far : t1
 *   hit  iff t0 <= t1 ; entry distance = t0
 *
 * Only sub, mul and compare-select: no operation pair an FMA could contract, and
 * every NaN (0 * inf) falls through the selects the same way everywhere.
 */
#ifndef RT_ORACLE_SLAB_H
#define RT_ORACLE_SLAB_H

#include <math.h>

#define RT_SLAB_MISS_T ((float)INFINITY)

static inline int rt_slab_hit(const float* lo, const float* hi, const float* o,
                              const float* inv, float tmax, float* t_entry)
{
  float t0 = 0.0f;
  float t1 = tmax;
  int d;
  for (d = 0; d < 3; ++d)
  {
    const float a = (lo[d] - o[d]) * inv[d];
    const float c = (hi[d] - o[d]) * inv[d];
    const float tn = (c < a) ? c : a;
    const float tf = (c < a) ? a : c;
    t0 = (tn > t0) ? tn : t0;
    t1 = (tf < t1) ? tf : t1;
  }
  *t_entry = t0;
  return t0 <= t1;
}

#endif
