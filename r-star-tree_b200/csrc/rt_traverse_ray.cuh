// rt_traverse_ray.cuh -- warp-cooperative ray queries on the same block layout (sm_100a).
//
// The reference has no ray code; a ray query is RTree::search() (reference
// include/RTree/rtree.hpp:984-1019) driven by a slab-test filter, which the API allows
// because functors are held by reference (rtree.hpp:1145) and may be stateful:
//   all-hits : filter = slab hit; functor collects the key ids whose AABB is hit
//   closest  : filter = slab hit && t <= t_best (non-strict: ties survive);
//              functor keeps (t, id) minimal, ties to the smaller id
//   any-hit  : functor returns true on the first hit = abort (rtree.hpp:994-997)
//
// Slab test -- identical operation for operation to oracle/slab.h so that host and
// device round the same way (only sub, mul, compare-select: nothing an FMA could fuse):
//   a = (lo - o) * inv ; c = (hi - o) * inv ; near/far by (c < a)
//   t0 = (near > t0) ? near : t0 ; t1 = (far < t1) ? far : t1 ; hit iff t0 <= t1
//
// Lane mapping, stack discipline and hit compaction are those of rt_traverse_box.cuh.
// The closest-hit result is independent of traversal order: a key with the final
// (t, id) can never be pruned because every ancestor bound has t_entry <= t.
#pragma once

#include "rt_block.cuh"
#include "rt_common.cuh"

namespace rtb
{

enum : uint32_t
{
  RAY_COUNT = 0,
  RAY_COUNT_STATS = 1,
  RAY_WRITE = 2,
  RAY_CLOSEST = 3,
  RAY_ANY = 4,
  RAY_CLOSEST_STATS = 5, // closest-hit / any-hit that also count the nodes they visit (RT_COLLECT_VISITS)
  RAY_ANY_STATS = 6
};

template <int W, bool BOXKEY, uint32_t KMODE>
__global__ void __launch_bounds__(TRAVERSE_THREADS)
ray_traverse_kernel(const RayLaunch p)
{
  // the *_STATS modes are their base mode plus the visit counters
  constexpr uint32_t RMODE = KMODE == RAY_COUNT_STATS     ? RAY_COUNT
                             : KMODE == RAY_CLOSEST_STATS ? RAY_CLOSEST
                             : KMODE == RAY_ANY_STATS     ? RAY_ANY
                                                          : KMODE;
  constexpr bool STATS = KMODE != RMODE;
  constexpr int G = 32 / W;
  constexpr int D = 3;
  extern __shared__ uint32_t smem[];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int slot = lane % W;
  const uint32_t grp = lane / W;
  const uint32_t lt_mask = (1u << lane) - 1u;

  // per warp: G words holding the null node, the stack, the hit buffer (see rt_traverse_box.cuh)
  const uint32_t stack_cap = 32u * p.tree.num_levels;
  const uint32_t warp_mem = smem_addr(smem + warp * (HIT_BUF + G + stack_cap));
  const uint32_t stack = warp_mem + (HIT_BUF + G) * 4u;
  const uint32_t hitbuf = warp_mem;
  const uint32_t stack_pop = stack + (grp - G) * 4u; // group g pops stack index sp - G + g
  const uint32_t null_link = p.tree.null_link;
  if (lane < G)
    sts_u32(stack - G * 4u + lane * 4u, null_link);
  __syncwarp();

  const unsigned char* lane_blocks = opaque(p.tree.blocks + slot * 16); // this lane's element of a row
  const uint32_t leaf_link_begin = p.tree.leaf_link_begin;

  unsigned long long visits_int = 0, visits_leaf = 0, rays_hit = 0, steps = 0;

  uint32_t range = (p.work.by_sm ? sm_id() : blockIdx.x) % p.work.ranges, tries = 0;
  for (;;)
  {
    // per_range is a multiple of chunk, so a chunk never crosses its range; n < 2^32 (checked
    // by the API) but ranges * per_range may exceed it by a little, hence the 64-bit position
    const uint32_t c = claim_chunk(p.work, range, lane);
    const unsigned long long begin = (unsigned long long)range * p.work.per_range + c;
    if (c >= p.work.per_range || begin >= p.n)
    {
      // this range is used up: go on to the next one, until all have been seen empty
      if (++tries >= p.work.ranges)
        break;
      range = range + 1 == p.work.ranges ? 0 : range + 1;
      continue;
    }
    const unsigned long long last = min(begin + p.work.chunk, (unsigned long long)p.n);

    // (a 64-bit induction variable on purpose: with a 32-bit one ptxas moves the traversal loop's
    // stack pointer and ballots out of the uniform datapath, which costs 20 % -- measured)
    for (unsigned long long i = begin; i < last; ++i)
    {
      const unsigned long long qi = p.order ? (unsigned long long)p.order[i] : i;
      float org[D], inv[D];
#pragma unroll
      for (int d = 0; d < D; ++d)
      {
        org[d] = __ldg(p.rays + qi * 7 + d);
        inv[d] = __ldg(p.rays + qi * 7 + 3 + d);
      }
      const float tmax = __ldg(p.rays + qi * 7 + 6);

      unsigned long long out_base = 0;
      uint32_t expect = 0;
      bool buffered = false;
      if (RMODE == RAY_WRITE)
      {
        out_base = p.offsets[qi];
        expect = (uint32_t)min(p.offsets[qi + 1] - out_base, (unsigned long long)0xFFFFFFFFu);
        buffered = p.sorted && expect <= HIT_BUF;
      }

      float t_best = tmax;
      uint32_t id_best = RT_INVALID_ID;
      bool found = false;
      uint32_t count = 0;
      uint32_t sp4 = 4u; // bytes on the stack (warp-uniform)
      if (lane == 0)
        sts_u32(stack, 0u);
      __syncwarp();

      while (sp4 != 0)
      {
        // with fewer than G nodes left the lower groups read the null node (all slots empty)
        const uint32_t node = lds_u32(stack_pop + sp4);
        sp4 -= min(sp4, (uint32_t)(G * 4));
        RTB_LOOP_SYNC();
        RTB_CHECK(node <= null_link, "ray traversal: popped a link beyond the image");

        const bool is_leaf = node >= leaf_link_begin;
        uint32_t link;
        float lo[D], hi[D];
        load_entry<float, D, W, BOXKEY>(lane_blocks + (unsigned long long)node * 16u, slot, is_leaf, lo, hi, link);
        float t0 = 0.0f;
        float t1 = tmax;
#pragma unroll
        for (int d = 0; d < D; ++d)
        {
          const float a = __fmul_rn(__fsub_rn(lo[d], org[d]), inv[d]);
          const float c = __fmul_rn(__fsub_rn(hi[d], org[d]), inv[d]);
          const float tn = (c < a) ? c : a;
          const float tf = (c < a) ? a : c;
          t0 = (tn > t0) ? tn : t0;
          t1 = (tf < t1) ? tf : t1;
        }
        bool hit = (link != RT_INVALID_ID) && (t0 <= t1);
        if (RMODE == RAY_CLOSEST)
          hit = hit && (is_leaf || t0 <= t_best);

        const uint32_t m_all = __ballot_sync(0xFFFFFFFFu, hit);
        const uint32_t m_leaf = __ballot_sync(0xFFFFFFFFu, hit && is_leaf);
        const uint32_t m_int = m_all & ~m_leaf;

        if (STATS)
        {
          visits_leaf += __popc(__ballot_sync(0xFFFFFFFFu, is_leaf && slot == 0 && node != null_link));
          visits_int += __popc(__ballot_sync(0xFFFFFFFFu, !is_leaf && slot == 0));
          ++steps;
        }

        if (hit && !is_leaf)
          sts_u32(stack + sp4 + __popc(m_int & lt_mask) * 4u, link);
        sp4 += __popc(m_int) * 4u;
        RTB_CHECK(sp4 <= stack_cap * 4u, "ray traversal: the stack outgrew 32 entries per level");

        if (m_leaf != 0)
        {
          if (RMODE == RAY_CLOSEST)
          {
            // t0 >= +0 and never NaN, so its bit pattern orders like the value
            const bool cand = hit && is_leaf;
            const uint32_t t_bits = __reduce_min_sync(0xFFFFFFFFu, cand ? __float_as_uint(t0) : 0xFFFFFFFFu);
            const uint32_t id_min = __reduce_min_sync(
                0xFFFFFFFFu, (cand && __float_as_uint(t0) == t_bits) ? link : RT_INVALID_ID);
            const float t = __uint_as_float(t_bits);
            if (!found || t < t_best || (t == t_best && id_min < id_best))
            {
              found = true;
              t_best = t;
              id_best = id_min;
            }
          }
          else if (RMODE == RAY_ANY)
          {
            found = true;
            sp4 = 0; // abort the whole search
          }
          else
          {
            if (RMODE == RAY_WRITE && hit && is_leaf)
            {
              const uint32_t at = count + __popc(m_leaf & lt_mask);
              if (buffered)
              {
                if (at < HIT_BUF)
                  sts_u32(hitbuf + at * 4u, link);
              }
              else if (at < expect)
              {
                p.ids[out_base + at] = link;
              }
            }
            count += __popc(m_leaf);
          }
        }
        RTB_LOOP_SYNC();
      }
      __syncwarp(); // the hit buffer is read across lanes below

      if (RMODE == RAY_COUNT)
      {
        if (lane == 0)
        {
          p.counts[qi] = count;
          if (count > HIT_BUF)
            atomicMax(p.max_list, count);
        }
      }
      else if (RMODE == RAY_WRITE)
      {
        if (buffered)
        {
          uint32_t v = ((uint32_t)lane < count) ? lds_u32(hitbuf + lane * 4u) : RT_INVALID_ID;
          v = warp_sort32(v, lane);
          if ((uint32_t)lane < count && (uint32_t)lane < expect)
            p.ids[out_base + lane] = v;
        }
        else if (p.sorted && expect > 1 && lane == 0)
        {
          p.big_list[atomicAdd(p.big_count, 1u)] = (uint32_t)qi;
        }
      }
      else if (RMODE == RAY_CLOSEST)
      {
        if (lane == 0)
        {
          p.t_out[qi] = found ? t_best : __uint_as_float(0x7F800000u);
          p.id_out[qi] = id_best;
          rays_hit += found ? 1 : 0;
        }
      }
      else // RAY_ANY
      {
        if (lane == 0)
        {
          p.any_out[qi] = found ? 1 : 0;
          rays_hit += found ? 1 : 0;
        }
      }
      __syncwarp();
    }
  }

  if (STATS && lane == 0)
  {
    atomicAdd(p.visit_stats + 0, visits_int);
    atomicAdd(p.visit_stats + 1, visits_leaf);
    atomicAdd(p.visit_stats + 3, steps);
  }
  if ((RMODE == RAY_CLOSEST || RMODE == RAY_ANY) && lane == 0 && rays_hit != 0)
    atomicAdd(p.visit_stats + 2, rays_hit);
}

} // namespace rtb
