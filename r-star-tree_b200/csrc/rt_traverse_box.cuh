// rt_traverse_box.cuh -- the hot kernel: warp-cooperative AABB range queries (sm_100a).
//
// Replaces a host loop over RTree::search(filter, functor) (reference
// include/RTree/rtree.hpp:1141-1146; recursion :984-1019) with
//   filter  = closed overlap of the child bound with the query box
//             (reference example/sample/sample.cpp:48-54, test/rtree.cpp:403)
//   functor = is_inside(query box, key)   (reference include/RTree/aabb.hpp:206-210,
//             less_equal :249-259; flat form: test/rtree.cpp:390)
// on the level-ordered SoA blocks of RTree::flatten_device().
//
// Mapping.  One warp answers one query at a time.  A block is W (8 or 16) slots wide, so
// the 32 lanes cover G = 32/W nodes per step: lane = (group g, slot c) loads child c of
// the g-th node popped from the warp's stack -- one 128-bit load per 4 record words, the
// W lanes of a group reading W*16 contiguous bytes (rt_block.cuh), i.e. fully coalesced:
// a 3-D float internal node is two such loads, a point leaf one.  The
// per-axis interval tests run in registers, __ballot_sync gathers the W-bit hit masks of
// all G nodes at once, and __popc over the lower lanes gives every hit lane its slot on
// the stack (children to descend) or in the hit list (leaf entries), so both compactions
// are branch-free and deterministic.
//
// Stack.  Per warp in shared memory.  Popping the top G entries and pushing children in
// lane order keeps the stack sorted by level, hence at most 32 entries per level are
// ever live: capacity 32 * num_levels, no overflow path needed.  Group g pops index
// sp - G + g; below index 0 sit G words naming the "null" block rt_upload appends (all
// slots empty), so a step with fewer than G nodes left needs no activity mask.
//
// Work.  Persistent grid; the queries (Morton order after the coherence pass) are cut into
// one contiguous range per SM and the warps of an SM claim a few queries at a time from
// their range, stealing from the other ranges at the end (WorkRanges, rt_common.cuh): the
// 64 warps of an SM walk neighbouring queries and share their nodes in L1.
//
// Memory.  Internal levels are a small, hot prefix of the block buffer; they and the
// leaves stream through L1/L2 via the read-only path (ld.global.nc).  Nothing is written
// but counts / hit ids.  QUANT (3-D float32 point trees): the same traversal on the
// quantised image of rt_quant.cuh -- 16 bytes per slot, one integer test per slot.
//
// Output.  PASS_COUNT(+STATS): hits per query.  PASS_STASH: hits per query, and lists of
// at most HIT_BUF hits are sorted in registers (bitonic over the warp) and parked in a
// stash so that the second traversal is only needed for the rare longer lists.
// PASS_WRITE: hit ids go to their final CSR position; short lists are sorted in the warp.
#pragma once

#include <type_traits>

#include "rt_block.cuh"
#include "rt_common.cuh"
#include "rt_quant.cuh"

namespace rtb
{

// a fresh slab of the stash pool for this warp (warp-uniform); none left: slab_left = 0, and the
// queries that needed one are answered by the second traversal
__device__ __forceinline__ void claim_stash_slab(const BoxLaunch& p, int lane, unsigned long long& slab_pos,
                                                 uint32_t& slab_left)
{
  unsigned long long got = 0;
  if (lane == 0)
    got = atomicAdd(p.stash_cursor, (unsigned long long)p.stash_slab);
  got = __shfl_sync(0xFFFFFFFFu, got, 0);
  const bool ok = got + p.stash_slab <= p.stash_capacity;
  slab_pos = ok ? got : slab_pos;
  slab_left = ok ? p.stash_slab : 0u;
}

// QUANT: traverse the quantised image (rt_quant.cuh) -- one 16-byte record per slot, one integer
// test for internal children and leaf points alike.  A leaf candidate strictly inside the query on the
// grid is a hit as it stands; only candidates in a boundary cell of the query (about 2 %) are looked up
// on the float record and tested with the reference's comparisons.
// PACKED (QUANT, dim <= 3): the queries arrive in processing order, already quantised, 16 bytes each
// (BoxLaunch::packed) -- one 128-bit load per query instead of an index, six floats and six
// quantisations per warp.
//
// (The traversal loop has no explicit warp barrier: RTB_LOOP_SYNC() in rt_block.cuh says why.)
// RTB_STAGE_TOP = 1 (experiment, north_star's "top levels staged in shared memory"): every CTA copies
// the quantised blocks of the top levels (TreeDev::stage_links) into shared memory and reads them
// from there.  Measured and left off.
#ifndef RTB_STAGE_TOP
#define RTB_STAGE_TOP 0
#endif


// STRICT (RT_STRICT_OVERLAP, float blocks only): the node filter -- and the key test of box keys under
// RT_INTERSECTS -- is the reference's N-D geometry_traits<aabb>::is_overlap(aabb, aabb)
// (include/RTree/aabb.hpp:225-236, less_than :238-248): boxes that only touch do not overlap.  Point
// keys keep is_overlap(aabb, point) = is_inside (aabb.hpp:219-223), RT_CONTAINS its containment test.
template <typename S, int D, int W, bool BOXKEY, bool CONTAINS, uint32_t PASS, bool QUANT = false, bool STRICT = false>
__global__ void __launch_bounds__(TRAVERSE_THREADS, QUANT ? 8 : 1) // quantised kernels: 32 registers, 64 warps per SM
box_traverse_kernel(const BoxLaunch p)
{
  static_assert(!(QUANT && STRICT), "the strict filter runs on the float blocks");
  static_assert(!QUANT || (std::is_same<S, float>::value && !BOXKEY),
                "the quantised image exists for float32 point-keyed trees");
  constexpr int G = 32 / W;
  constexpr bool PACKED = QUANT && D <= 3;
  extern __shared__ uint32_t smem[];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int slot = lane % W;
  const uint32_t grp = lane / W;
  uint32_t lt_mask = (1u << lane) - 1u;
  asm volatile("" : "+r"(lt_mask)); // kept in its register: ptxas otherwise re-derives it from %tid in the leaf path

  // per warp: the hit buffer, G words holding the null node, the stack (32-bit shared addresses;
  // the hit buffer sits at a constant distance below the stack so it needs no register of its own)
  const uint32_t stack_cap = 32u * p.tree.num_levels;
  uint32_t warp_mem = smem_addr(smem + warp * (HIT_BUF + G + stack_cap));
  // one register for the hit buffer, the stack and the null words
  constexpr bool ABS_SP = D <= 3 || RTB_ABS_SP_WIDE; // see the stack pointer in the query loop
  if constexpr (ABS_SP)
    warp_mem = __reduce_min_sync(0xFFFFFFFFu, warp_mem); // a warp reduction: ptxas then knows it to be warp-uniform
  else
    asm volatile("" : "+r"(warp_mem));
  const uint32_t stack = warp_mem + (HIT_BUF + G) * 4u;
  constexpr uint32_t HITBUF_BELOW = (HIT_BUF + G) * 4u; // hit buffer = stack - HITBUF_BELOW
  // group g pops the word at stack index sp - G + g: with fewer than G
  // nodes left the lower groups read the null node below the stack (a block of empty slots),
  // so popping needs no "is this group active" logic at all
  // (an offset from the stack pointer; see the query loop for what that is counted from)
  const uint32_t sp_base = ABS_SP ? warp_mem : 0u;
  const uint32_t pop_offset = warp_mem - sp_base + HITBUF_BELOW + (grp - G) * 4u;
  const uint32_t null_link = QUANT ? p.tree.qnull_link : p.tree.null_link;
  if (lane < G)
    sts_u32(stack - G * 4u + lane * 4u, null_link);
  __syncwarp();
#if RTB_STAGE_TOP
  // the top of the quantised image, once per CTA, behind the warps' stacks
  const uint32_t stage_links = QUANT ? p.tree.stage_links : 0u;
  const uint32_t stage_base = smem_addr(smem + (TRAVERSE_THREADS / 32) * (HIT_BUF + G + stack_cap)) + slot * 16u;
  if constexpr (QUANT)
  {
    for (uint32_t k = threadIdx.x; k < stage_links; k += TRAVERSE_THREADS)
    {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.tree.qblocks) + k);
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage_base - slot * 16u + k * 16u), "r"(v.x),
                   "r"(v.y), "r"(v.z), "r"(v.w)
                   : "memory");
    }
    __syncthreads();
  }
#endif

  // this lane's element of a row
  const unsigned char* lane_blocks = opaque((QUANT ? p.tree.qblocks : p.tree.blocks) + slot * 16);
  // 8-D quantised records: where the lane's link sits behind the two rows (4-byte elements), kept in a
  // register (ptxas otherwise re-derives it from %tid in every step)
  uint32_t link_offset = 2 * W * 16 - slot * 12;
  asm volatile("" : "+r"(link_offset));
  // QUANT: the float record behind a leaf slot, relative to the lane's quantised-image pointer
  const long long leaf_delta = QUANT ? (long long)(p.tree.leaf_blocks - p.tree.qblocks) : 0ll;
  const uint32_t leaf_link_begin = QUANT ? p.tree.qleaf_link_begin : p.tree.leaf_link_begin;
  const S* __restrict__ queries = static_cast<const S*>(p.queries);

  unsigned long long visits_int = 0, visits_leaf = 0, steps = 0, lookups = 0;
  // stash slab owned by this warp (PASS_STASH)
  unsigned long long slab_pos = 0;
  uint32_t slab_left = 0;

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
      unsigned long long qi;  // the query's index in the caller's array
      uint32_t defer_key;     // what the deferred list holds for it: see BoxLaunch::order
      S qlo[D], qhi[D];
      uint32_t qb[D]; // the query on the quantisation grid (QUANT)
      if constexpr (PACKED)
      {
        const unsigned long long pos = p.order ? (unsigned long long)p.order[i] : i;
        const uint4 v = __ldg(p.packed + pos);
        qb[0] = v.x, qb[1] = v.y;
        if constexpr (D == 3)
          qb[2] = v.z;
        qi = D == 3 ? v.w : v.z;
        defer_key = (uint32_t)pos;
      }
      else
      {
        qi = p.order ? (unsigned long long)p.order[i] : i;
        defer_key = (uint32_t)qi;
#pragma unroll
        for (int d = 0; d < D; ++d)
        {
          qlo[d] = __ldg(queries + qi * (2 * D) + d);
          qhi[d] = __ldg(queries + qi * (2 * D) + D + d);
        }
        if constexpr (QUANT)
          quant_query<D>(qlo, qhi, p.tree.qlo, p.tree.qscale, qb);
      }

      unsigned long long out_base = 0;
      uint32_t expect = 0;
      bool buffered = (PASS == PASS_STASH);
      if (PASS == PASS_WRITE)
      {
        out_base = p.offsets[qi];
        expect = (uint32_t)min(p.offsets[qi + 1] - out_base, (unsigned long long)0xFFFFFFFFu);
        buffered = p.sorted && expect <= HIT_BUF;
      }

      // PASS_STASH: every query starts with at least a quarter slab of room in the warp's slab (and
      // never less than the hit buffer), so lists up to that length are stashed without further ado
      if (PASS == PASS_STASH && slab_left < max(HIT_BUF, p.stash_slab / 4))
        claim_stash_slab(p, lane, slab_pos, slab_left);
      uint32_t count = 0;
      // the stack pointer = sp_base + the bytes on the stack (warp-uniform: it only ever sees ballots and
      // warp_mem).  Narrow records count it from the warp's shared memory, i.e. keep it as an address, so
      // that neither the pop nor the push has to add the warp's base: two instructions per step
      // (profiles/r2_box_traverse_f32_3d_w8_stash_quant.sass).  For the wide (8-D) kernel see RTB_ABS_SP_WIDE
      uint32_t spw = sp_base + 4u;
      if (lane == 0)
        sts_u32(stack, 0u); // the root; its own bound is never tested (reference rtree.hpp:984)
      __syncwarp();

      while (spw != sp_base)
      {
        const uint32_t node = lds_u32(spw + pop_offset);
        if constexpr (ABS_SP)
          spw = max(spw - (uint32_t)(G * 4), sp_base); // (shared addresses of dynamic memory are >= 1024: no wrap)
        else
          spw -= min(spw, (uint32_t)(G * 4));
        RTB_LOOP_SYNC();
        RTB_CHECK(node <= null_link, "box traversal: popped a link beyond the image");

        const bool is_leaf = node >= leaf_link_begin;
        uint32_t link;
        bool hit;
        uint32_t hit_bits = 0; // QUANT: the guard bits that survived; all of them = the slot passes
        using QRec = Record<D + 1, W>;
        uint32_t w[QRec::PADDED]; // the slot's quantised record (QUANT)
        if constexpr (QUANT)
        {
          if constexpr (D == 8)
          {
            // 8-D records are two 16-byte rows + a link row.  A LEAF slot holds a point, whose two
            // rows would say the same thing twice (q(p) and 32768 - q(p)); the image stores only the
            // first, G | q(p_2j) | q(p_2j+1) << 16 (rt_quant.cu), and the second is derived with one
            // subtraction per word.  More than half of this kernel's node visits are leaves and it is
            // bound by L1 wavefronts, so not fetching their second row is 2 of 4.5 wavefronts per leaf.
            // An empty leaf slot is G | 0: its first row already fails every test.
            static_assert(QRec::NV == 2 && QRec::R == 1, "8-D quantised record: two rows and the link");
            const unsigned char* rec = lane_blocks + (unsigned long long)node * 16u;
            const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rec));
            uint4 r1;
            if (!is_leaf)
            {
              r1 = __ldg(reinterpret_cast<const uint4*>(rec + W * 16));
            }
            else
            {
              // each half: 0x18000 - (0x8000 | q) = 0x8000 | (32768 - q); no borrow between the halves
              r1.x = 0x80018000u - r0.x, r1.y = 0x80018000u - r0.y;
              r1.z = 0x80018000u - r0.z, r1.w = 0x80018000u - r0.w;
            }
            // (a leaf slot's link is the point's id: only a hit needs it, and the leaf section fetches it then --
            // 13 hits in the 29,000 leaf slots a config-4 query looks at)
            // (a predicated load that leaves the register alone otherwise: nothing reads a leaf lane's `link`
            // before the leaf section has set it)
            asm volatile("{\n\t"
                         ".reg .pred p;\n\t"
                         "setp.lt.u32 p, %1, %2;\n\t"
                         "@p ld.global.nc.u32 %0, [%3];\n\t"
                         "}"
                         : "=r"(link)
                         : "r"(node), "r"(leaf_link_begin), "l"(rec + link_offset));
            w[0] = r0.x, w[1] = r0.y, w[2] = r0.z, w[3] = r0.w;
            w[4] = r1.x, w[5] = r1.y, w[6] = r1.z, w[7] = r1.w;
            w[D] = link;
            hit_bits = quant_pass_bits<D>(w, qb);
          }
          else
          {
#if RTB_STAGE_TOP
            if (D == 3 && node < stage_links)
            {
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                           : "r"(stage_base + node * 16u)
                           : "memory");
            }
            else
#endif
              QRec::load(lane_blocks + (unsigned long long)node * 16u, slot, w);
            link = w[D];
            hit_bits = quant_pass_bits<D>(w, qb); // empty slots never pass
          }
          hit = hit_bits == QUANT_GUARD;
        }
        else
        {
          S lo[D], hi[D];
          load_entry<S, D, W, BOXKEY>(lane_blocks + (unsigned long long)node * 16u, slot, is_leaf, lo, hi, link);
          hit = link != RT_INVALID_ID;
#pragma unroll
          for (int d = 0; d < D; ++d)
          {
            if (STRICT && (!is_leaf || (BOXKEY && !CONTAINS)))
            {
              // the reference's strict overlap: false as soon as one axis has q.min >= hi or lo >= q.max
              hit = hit && !(qlo[d] >= hi[d]) && !(lo[d] >= qhi[d]);
            }
            else if (BOXKEY && CONTAINS)
            {
              // leaf: key box inside the query box; internal: closed overlap
              const S a = is_leaf ? lo[d] : hi[d];
              const S b = is_leaf ? hi[d] : lo[d];
              hit = hit && !(a < qlo[d]) && !(b > qhi[d]);
            }
            else
            {
              // closed overlap; for a point key (lo == hi) this IS the reference's
              // is_inside: !(p < qmin) && !(p > qmax) per axis
              hit = hit && !(hi[d] < qlo[d]) && !(lo[d] > qhi[d]);
            }
          }
        }

        uint32_t m_leaf, m_int, spw_next;
        if constexpr (QUANT && PASS != PASS_COUNT_STATS)
        {
          // the two ballots and the push, spelled out: three compares for (leaf?, hit && leaf, hit &&
          // !leaf), two votes, and the children that passed go onto the stack in lane order, which keeps
          // it sorted by level (left to nvcc this part costs four instructions more per step)
          asm volatile("{\n\t"
                       ".reg .pred pl, pleaf, pint;\n\t"
                       ".reg .b32 t;\n\t"
                       "setp.ge.u32 pl, %3, %4;\n\t"
                       "setp.eq.and.b32 pleaf, %5, 0x80008000, pl;\n\t"
                       "setp.eq.and.b32 pint, %5, 0x80008000, !pl;\n\t"
                       "vote.sync.ballot.b32 %0, pleaf, 0xffffffff;\n\t"
                       "vote.sync.ballot.b32 %1, pint, 0xffffffff;\n\t"
                       "and.b32 t, %1, %6;\n\t"
                       "popc.b32 t, t;\n\t"
                       "mad.lo.u32 t, t, 4, %7;\n\t"
                       "@pint st.shared.u32 [t+%9], %8;\n\t"
                       "popc.b32 t, %1;\n\t"
                       "mad.lo.u32 %2, t, 4, %10;\n\t"
                       "}"
                       : "=r"(m_leaf), "=r"(m_int), "=r"(spw_next)
                       : "r"(node), "r"(leaf_link_begin), "r"(hit_bits), "r"(lt_mask), "r"(spw + (warp_mem - sp_base)),
                         "r"(link), "n"(HITBUF_BELOW), "r"(spw)
                       : "memory");
          spw = spw_next;
        }
        else
        {
          m_leaf = __ballot_sync(0xFFFFFFFFu, hit && is_leaf);
          m_int = __ballot_sync(0xFFFFFFFFu, hit && !is_leaf);

          if (PASS == PASS_COUNT_STATS)
          {
            visits_leaf += __popc(__ballot_sync(0xFFFFFFFFu, is_leaf && slot == 0 && node != null_link));
            visits_int += __popc(__ballot_sync(0xFFFFFFFFu, !is_leaf && slot == 0));
            ++steps;
          }

          // children to descend: compact onto the stack in lane order (keeps it level-sorted)
          if (hit && !is_leaf)
            sts_u32(spw + (warp_mem - sp_base) + HITBUF_BELOW + __popc(m_int & lt_mask) * 4u, link);
          spw += __popc(m_int) * 4u;
        }

        RTB_CHECK(spw - sp_base <= stack_cap * 4u, "box traversal: the stack outgrew 32 entries per level");

        // leaf entries that satisfy the predicate
        if (m_leaf != 0)
        {
          // the lanes in `m_hits` (this one iff `leaf_hit`) hold a hit: into the hit buffer / slab / list
          auto record_hits = [&](const bool leaf_hit, const uint32_t m_hits)
          {
            if (PASS == PASS_STASH || PASS == PASS_WRITE)
            {
              if (leaf_hit)
              {
                const uint32_t at = count + __popc(m_hits & lt_mask);
                if (buffered)
                {
                  if (at < HIT_BUF)
                    sts_u32(warp_mem + at * 4u, link);
                  else if (PASS == PASS_STASH && at < slab_left)
                  {
                    RTB_CHECK(slab_pos + at < p.stash_capacity, "box traversal: spill beyond the stash pool");
                    p.stash[slab_pos + at] = link;
                  } // beyond the hit buffer: straight into the warp's slab,
                                                   // at its final place there (sorted after the gather)
                }
                else if (PASS == PASS_WRITE && at < expect)
                {
                  p.ids[out_base + at] = link;
                }
              }
            }
            count += __popc(m_hits);
            if (PASS == PASS_COUNT && p.stop_at_first && count != 0)
            {
              // the functor-returns-true abort of reference rtree.hpp:994-997: which hit came first
              // depends on the traversal order, "there is one" does not
              count = 1;
              spw = sp_base;
            }
          };
          if constexpr (QUANT)
          {
            // a candidate strictly inside the query on the grid is a hit; one in a boundary cell of
            // the query gets the reference's exact test on its float record
            const bool candidate = hit && is_leaf;
            bool unsure = false;
            if constexpr (D <= 3)
            {
              unsure = candidate && !quant_strictly_inside<D>(w, qb);
            }
            else if (candidate)
            {
              // wide records are not kept in registers across the ballots (8 registers = a CTA per SM):
              // the few lanes with a candidate read theirs again, an L1 hit
              uint32_t again[QRec::PADDED];
              const unsigned char* rec = lane_blocks + (unsigned long long)node * 16u;
              const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rec));
              link = __ldg(reinterpret_cast<const uint32_t*>(rec + link_offset));
              again[0] = r0.x, again[1] = r0.y, again[2] = r0.z, again[3] = r0.w;
              again[4] = 0x80018000u - r0.x, again[5] = 0x80018000u - r0.y;
              again[6] = 0x80018000u - r0.z, again[7] = 0x80018000u - r0.w;
              unsure = !quant_strictly_inside<D>(again, qb);
            }
            const uint32_t m_unsure = __ballot_sync(0xFFFFFFFFu, unsure);
            if (m_unsure == 0)
            {
              // the common case gets its own copy of the bookkeeping, so that "this lane holds a hit"
              // stays a predicate instead of travelling through a register to the join
              record_hits(candidate, m_leaf);
            }
            else
            {
              bool leaf_hit = candidate;
              if (unsure)
              {
                // the float block of a point leaf has the stride of its quantised block
                uint32_t kw[QRec::PADDED];
                QRec::load(lane_blocks + leaf_delta + (unsigned long long)(node - leaf_link_begin) * 16u, slot, kw);
                bool inside = true;
                unsigned long long at_q = qi * (2 * D);
                asm volatile("" : "+l"(at_q)); // keep the address of the float box out of the loop's registers
#pragma unroll
                for (int d = 0; d < D; ++d)
                {
                  // the float box is not kept in registers: re-read (an L1 hit) in this rare path
                  const S k_d = __uint_as_float(kw[d]);
                  const S lo_d = __ldg(queries + at_q + d);
                  const S hi_d = __ldg(queries + at_q + D + d);
                  inside = inside && !(k_d < lo_d) && !(k_d > hi_d);
                }
                leaf_hit = inside;
              }
              if (PASS == PASS_COUNT_STATS)
                lookups += __popc(m_unsure);
              record_hits(leaf_hit, __ballot_sync(0xFFFFFFFFu, leaf_hit));
            }
          }
          else
          {
            record_hits(hit && is_leaf, m_leaf);
          }
        }
        RTB_LOOP_SYNC();
      }
      __syncwarp(); // the hit buffer is read across lanes below

      // ---- end of query ------------------------------------------------------------
      if (PASS == PASS_COUNT || PASS == PASS_COUNT_STATS)
      {
        if (lane == 0)
        {
          p.counts[qi] = count;
          if (count > HIT_BUF)
            atomicMax(p.max_list, count); // the write pass will hand this list to the segment sort
        }
      }
      else if (PASS == PASS_STASH)
      {
        uint32_t where = NO_STASH;
        if (count == 0)
        {
          where = 0;
        }
        else if (count <= HIT_BUF)
        {
          if (slab_left >= count)
          {
            RTB_CHECK(slab_pos + count <= p.stash_capacity, "box traversal: short list beyond the stash pool");
            uint32_t v = ((uint32_t)lane < count) ? lds_u32(warp_mem + lane * 4u) : RT_INVALID_ID;
            if (p.sorted)
              v = warp_sort32(v, lane, count);
            if ((uint32_t)lane < count)
              p.stash[slab_pos + lane] = v;
            where = (uint32_t)slab_pos;
            slab_pos += count;
            slab_left -= count;
          }
        }
        else if (count <= slab_left)
        {
          // a list that outgrew the hit buffer but not the slab: ids HIT_BUF.. are in the slab
          // already (every one of them had at < slab_left), the first HIT_BUF join them
          RTB_CHECK(slab_pos + count <= p.stash_capacity, "box traversal: long list beyond the stash pool");
          p.stash[slab_pos + lane] = lds_u32(warp_mem + lane * 4u);
          where = (uint32_t)slab_pos;
          slab_pos += count;
          slab_left -= count;
          if (p.sorted && lane == 0)
          {
            p.big_list[atomicAdd(p.big_count, 1u)] = (uint32_t)qi; // sorted after the gather
            atomicMax(p.max_list, count);
          }
        }
        if (lane == 0)
        {
          p.counts[qi] = count;
          p.stash_pos[qi] = where;
          if (where == NO_STASH)
          {
            RTB_CHECK(*p.deferred_count < p.n, "box traversal: more deferred queries than queries");
            p.deferred_list[atomicAdd(p.deferred_count, 1u)] = defer_key;
            atomicMax(p.max_list, count); // the second traversal writes it, the segment sort orders it
          }
        }
      }
      else // PASS_WRITE
      {
        if (buffered)
        {
          uint32_t v = ((uint32_t)lane < count) ? lds_u32(warp_mem + lane * 4u) : RT_INVALID_ID;
          v = warp_sort32(v, lane, count);
          if ((uint32_t)lane < count && (uint32_t)lane < expect)
            p.ids[out_base + lane] = v;
        }
        else if (p.sorted && expect > 1 && lane == 0)
        {
          p.big_list[atomicAdd(p.big_count, 1u)] = (uint32_t)qi;
        }
      }
      __syncwarp();
    }
  }

  if (PASS == PASS_COUNT_STATS && lane == 0)
  {
    atomicAdd(p.visit_stats + 0, visits_int);
    atomicAdd(p.visit_stats + 1, visits_leaf);
    atomicAdd(p.visit_stats + 3, steps);
    if (QUANT)
      atomicAdd(p.confirm_stats, lookups);
  }
}

template <typename S, int D, int W, bool BOXKEY, bool CONTAINS, uint32_t PASS, bool QUANT = false, bool STRICT = false>
cudaError_t launch_box_pass(const BoxLaunch& p)
{
  // the (dim, width) pairs of quant_supported() in rt_quant.cu
  constexpr bool CAN_QUANT = std::is_same<S, float>::value && !BOXKEY
                             && ((D == 2 && W == 8) || (D == 3 && W == 8) || (D == 8 && W == 16));
  if constexpr (CAN_QUANT && !QUANT && !STRICT)
  {
    if (p.tree.qblocks != nullptr)
      return launch_box_pass<S, D, W, BOXKEY, CONTAINS, PASS, true>(p);
  }
  // (a 1-D tree's is_overlap is closed in the reference, include/RTree/aabb.hpp:147-159: nothing to do)
  if constexpr (!STRICT && !QUANT && D > 1)
  {
    if (p.strict)
      return launch_box_pass<S, D, W, BOXKEY, CONTAINS, PASS, false, true>(p);
  }
  auto kernel = box_traverse_kernel<S, D, W, BOXKEY, CONTAINS, PASS, QUANT, STRICT>;
  size_t smem = traverse_smem_bytes(p.tree.num_levels, W);
#if RTB_STAGE_TOP
  if (QUANT)
    smem += (size_t)p.tree.stage_links * 16u;
#endif
  cudaError_t err = cudaSuccess;
  if (smem > 48 * 1024)
  {
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess)
      return err;
  }
  BoxLaunch q = p;
  int grid = 0;
  if ((err = plan_work((const void*)kernel, smem, p.n, p.grid, p.stream, q.work, grid)) != cudaSuccess)
    return err;
  kernel<<<grid, TRAVERSE_THREADS, smem, p.stream>>>(q);
  return cudaGetLastError();
}

template <typename S, int D, int W, bool BOXKEY, bool CONTAINS>
cudaError_t launch_box(const BoxLaunch& p)
{
  switch (p.pass)
  {
  case PASS_COUNT: return launch_box_pass<S, D, W, BOXKEY, CONTAINS, PASS_COUNT>(p);
  case PASS_COUNT_STATS: return launch_box_pass<S, D, W, BOXKEY, CONTAINS, PASS_COUNT_STATS>(p);
  case PASS_STASH: return launch_box_pass<S, D, W, BOXKEY, CONTAINS, PASS_STASH>(p);
  case PASS_WRITE: return launch_box_pass<S, D, W, BOXKEY, CONTAINS, PASS_WRITE>(p);
  default: return cudaErrorInvalidValue;
  }
}

// the instantiation list of one scalar type: (dim, width) pairs, point and box keys
#define RTB_BOX_CASE(S, D, W)                                                                 \
  if (dim == D && width == W)                                                                 \
  {                                                                                           \
    if (key_kind == RT_KEY_POINT)                                                             \
      return &launch_box<S, D, W, false, false>;                                              \
    if (mode == RT_CONTAINS)                                                                  \
      return &launch_box<S, D, W, true, true>;                                                \
    return &launch_box<S, D, W, true, false>;                                                 \
  }

} // namespace rtb
