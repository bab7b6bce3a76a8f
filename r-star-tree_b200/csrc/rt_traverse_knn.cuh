// rt_traverse_knn.cuh -- warp-cooperative k-nearest-neighbour queries on the float blocks (sm_100a).
//
// The reference has no kNN code; the query is RTree::search() (reference
// include/RTree/rtree.hpp:984-1019) driven by stateful functors (legal: they are held by
// reference, rtree.hpp:1145), defined operation for operation in oracle/knn.h:
//   filter  : descend iff dist2(q, child bound) <= worst.d2 (non-strict: ties survive, so the
//             answer does not depend on the traversal order)
//   functor : a key with (d2, id) < worst replaces the worst of the k best pairs
//   result  : the k pairs ascending by (d2, id), missing ones (+inf, RT_INVALID_ID)
//   dist2   : per axis e = (q < lo) ? lo - q : ((q > hi) ? q - hi : 0); d2 = d2 + e * e
//             (float32, round to nearest, axis 0 first, never fused)
//
// Mapping and stack discipline are those of rt_traverse_box.cuh (one warp per query, G = 32/W
// nodes per step, one lane per child slot, null node below the stack).  Differences:
//   * a stack entry is (link, dist2 of the node's bound): entries that the shrinking bound has
//     overtaken since they were pushed are dropped when popped;
//   * the children of one node are pushed far-ones-first, so the nearest are popped next and the
//     bound tightens early (children of different nodes keep their order: the stack stays sorted
//     by level, which is what bounds its depth); "far" = more than 4 x the distance of the node's
//     nearest child, found with a butterfly over the node's lanes;
//   * the k best pairs are kept SORTED across the lanes (lane i holds the i-th best, k <= 32): a
//     candidate that beats the worst is inserted with one shuffle-up of the pairs behind it, the
//     worst is simply lane k-1's pair, and the result needs no sort at the end.  The candidates of a
//     step are offered nearest first (one warp min-reduction each).  (Round 1 kept them
//     unsorted, found the worst with two warp reductions per accepted candidate and sorted 64-bit
//     pairs over the warp per query: 13 k instructions per query, now about a third.)
#pragma once

#include "rt_block.cuh"
#include "rt_common.cuh"

namespace rtb
{

__device__ __forceinline__ void lds_u32x2(uint32_t addr, uint32_t& a, uint32_t& b)
{
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts_u32x2(uint32_t addr, uint32_t a, uint32_t b)
{
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

template <int D, int W, bool BOXKEY, bool COLLECT = false> // COLLECT: also count the node visits (RT_COLLECT_VISITS)
__global__ void __launch_bounds__(TRAVERSE_THREADS)
knn_traverse_kernel(const KnnLaunch p)
{
  constexpr int G = 32 / W;
  constexpr uint32_t INF_BITS = 0x7F800000u;
  extern __shared__ uint32_t smem[];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int slot = lane % W;
  const uint32_t grp = lane / W;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t grp_mask = (W == 32 ? 0xFFFFFFFFu : ((1u << W) - 1u)) << (grp * W);

  // per warp: G null entries, then the stack of (link, dist2 bits) pairs
  const uint32_t stack_cap = 32u * p.tree.num_levels;
  const uint32_t warp_mem = smem_addr(smem + warp * 2 * (G + stack_cap));
  const uint32_t stack = warp_mem + G * 8u;
  const uint32_t stack_pop = stack + (grp - G) * 8u;
  const uint32_t null_link = p.tree.null_link;
  if (lane < G)
    sts_u32x2(warp_mem + lane * 8u, null_link, 0u);
  __syncwarp();

  const unsigned char* lane_blocks = opaque(p.tree.blocks + slot * 16);
  const uint32_t leaf_link_begin = p.tree.leaf_link_begin;
  const uint32_t k = p.k;
  const int last = (int)k - 1; // the lane that holds the worst of the k best

  unsigned long long visits_int = 0, visits_leaf = 0;

  uint32_t range = (p.work.by_sm ? sm_id() : blockIdx.x) % p.work.ranges, tries = 0;
  for (;;)
  {
    const uint32_t c = claim_chunk(p.work, range, lane);
    const unsigned long long begin = (unsigned long long)range * p.work.per_range + c;
    if (c >= p.work.per_range || begin >= p.n)
    {
      if (++tries >= p.work.ranges)
        break;
      range = range + 1 == p.work.ranges ? 0 : range + 1;
      continue;
    }
    const unsigned long long end = min(begin + p.work.chunk, (unsigned long long)p.n);

    for (unsigned long long i = begin; i < end; ++i)
    {
      const unsigned long long qi = p.order ? (unsigned long long)p.order[i] : i;
      float q[D];
#pragma unroll
      for (int d = 0; d < D; ++d)
        q[d] = __ldg(p.points + qi * D + d);

      // the k best, sorted ascending by (d2 bits, id) across lanes 0 .. k-1 (d2 >= +0: bits order)
      uint32_t bd = INF_BITS, bi = RT_INVALID_ID;
      uint32_t wd = INF_BITS, wi = RT_INVALID_ID; // the worst of them = lane k-1's pair

      uint32_t sp8 = 8u; // bytes on the stack
      if (lane == 0)
        sts_u32x2(stack, 0u, 0u); // the root; its own bound is never tested (reference rtree.hpp:984)
      __syncwarp();

      while (sp8 != 0)
      {
        uint32_t node, node_d2;
        lds_u32x2(stack_pop + sp8, node, node_d2);
        sp8 -= min(sp8, (uint32_t)(G * 8));
        RTB_LOOP_SYNC();
        RTB_CHECK(node <= null_link, "kNN traversal: popped a link beyond the image");
        if (node_d2 > wd) // the bound has moved past this node since it was pushed
          node = null_link;

        const bool is_leaf = node >= leaf_link_begin;
        uint32_t link;
        float lo[D], hi[D];
        load_entry<float, D, W, BOXKEY>(lane_blocks + (unsigned long long)node * 16u, slot, is_leaf, lo, hi, link);
        float d2 = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d)
        {
          // oracle/knn.h: e = (q < lo) ? lo - q : ((q > hi) ? q - hi : 0).  For lo <= hi at most one of
          // the two differences is positive, so the select chain equals max(lo - q, q - hi, 0) bit for
          // bit (the zero is +0 either way: fmaxf(-0, +0) = +0), at two instructions fewer per axis
          const float e = fmaxf(fmaxf(__fsub_rn(lo[d], q[d]), __fsub_rn(q[d], hi[d])), 0.0f);
          d2 = __fadd_rn(d2, __fmul_rn(e, e));
        }
        const uint32_t d2b = __float_as_uint(d2);
        const bool valid = link != RT_INVALID_ID;

        if constexpr (COLLECT)
        {
          visits_leaf += __popc(__ballot_sync(0xFFFFFFFFu, is_leaf && slot == 0 && node != null_link));
          visits_int += __popc(__ballot_sync(0xFFFFFFFFu, !is_leaf && slot == 0));
        }

        // ---- children to descend: those the bound has not excluded, far ones of a node first ------
        const bool push = valid && !is_leaf && d2b <= wd;
        const uint32_t m_push = __ballot_sync(0xFFFFFFFFu, push);
        if (m_push != 0)
        {
          uint32_t dmin = push ? d2b : 0xFFFFFFFFu; // nearest pushed child of this lane's node
#pragma unroll
          for (int s = 1; s < W; s <<= 1)
            dmin = min(dmin, __shfl_xor_sync(0xFFFFFFFFu, dmin, s));
          const bool is_far = push && d2 > __fmul_rn(__uint_as_float(dmin), 4.0f);
          const uint32_t m_far = __ballot_sync(0xFFFFFFFFu, is_far);
          const uint32_t below = __popc(m_push & ~grp_mask & lt_mask); // pushes of the lower groups
          const uint32_t far_g = m_far & grp_mask, near_g = m_push & ~m_far & grp_mask;
          const uint32_t at = below + (is_far ? __popc(far_g & lt_mask) : __popc(far_g) + __popc(near_g & lt_mask));
          if (push)
            sts_u32x2(stack + sp8 + at * 8u, link, d2b);
          sp8 += __popc(m_push) * 8u;
          RTB_CHECK(sp8 <= stack_cap * 8u, "kNN traversal: the stack outgrew 32 entries per level");
        }

        // ---- leaf entries that beat the worst of the k best: inserted one at a time, nearest first ----
        // (nearest first: every insertion tightens the bound as much as it can, so fewer of the step's
        // later candidates still qualify, and the first one whose distance exceeds the bound ends the
        // step -- 77 insertions per query in lane order at k = 8 on the config-2 tree)
        bool cand = valid && is_leaf && (d2b < wd || (d2b == wd && link < wi));
        uint32_t m_cand = __ballot_sync(0xFFFFFFFFu, cand);
        while (m_cand != 0)
        {
          const uint32_t cd = __reduce_min_sync(0xFFFFFFFFu, cand ? d2b : 0xFFFFFFFFu);
          if (cd > wd)
            break; // no remaining candidate can beat the worst any more
          const int src = __ffs(__ballot_sync(0xFFFFFFFFu, cand && d2b == cd)) - 1;
          const uint32_t ci = __shfl_sync(0xFFFFFFFFu, link, src);
          if (lane == src)
            cand = false;
          m_cand &= ~(1u << src);
          if (cd < wd || ci < wi) // (cd == wd here otherwise) still better than the worst: warp-uniform
          {
            // every pair behind the candidate's place moves up one lane, the lane at its place takes it
            const bool behind = bd > cd || (bd == cd && bi > ci);
            const uint32_t up_d = __shfl_up_sync(0xFFFFFFFFu, bd, 1);
            const uint32_t up_i = __shfl_up_sync(0xFFFFFFFFu, bi, 1);
            const bool prev_behind = lane > 0 && (up_d > cd || (up_d == cd && up_i > ci));
            if (behind)
            {
              bd = prev_behind ? up_d : cd;
              bi = prev_behind ? up_i : ci;
            }
            wd = __shfl_sync(0xFFFFFFFFu, bd, last);
            wi = __shfl_sync(0xFFFFFFFFu, bi, last);
          }
        }
        RTB_LOOP_SYNC();
      }

      // ---- the k pairs, already ascending by (d2, id) ---------------------------------------------
      if ((uint32_t)lane < k)
      {
        p.ids_out[qi * k + lane] = bi;
        p.d2_out[qi * k + lane] = __uint_as_float(bd);
      }
      __syncwarp();
    }
  }

  if (COLLECT && lane == 0)
  {
    atomicAdd(p.visit_stats + 0, visits_int);
    atomicAdd(p.visit_stats + 1, visits_leaf);
  }
}

} // namespace rtb
