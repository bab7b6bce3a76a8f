// rt_block.cuh -- device-side view of the block layout of include/rtree_b200.h, and the
// small warp / shared-memory primitives the traversal kernels share (internal header).
#pragma once

#include "rt_common.cuh"

// No __syncwarp() inside the traversal loop (RTB_LOOP_SYNCWARP = 0).  The stack is written and read by
// different lanes, so the loop needs the lanes of a warp to execute it together -- and they do: the
// whole warp enters a step converged (the loop condition is made of ballot results and a warp reduction), the pop is one LDS for
// all lanes, every step has two full-mask vote.sync between the pop and the push, the push is a
// predicated STS (no branch), and the only divergent code (the leaf section's `if (lane has a hit)`)
// sits inside reconvergence barriers (BSSY / BSYNC in the SASS, profiles/r2_*.sass) that close
// before the next pop.  A converged warp's shared-memory instructions are executed in program order
// by the LSU, so pop -> push -> pop needs no further barrier; the two explicit ones cost two issue
// slots per step (3.6 % of the kernel, measured).  RTB_LOOP_SYNCWARP = 1 puts them back (`make
// variants` builds that as librtree_b200_sync.so for comparison).  The per-query barriers stay.
#ifndef RTB_LOOP_SYNCWARP
#define RTB_LOOP_SYNCWARP 0
#endif
#if RTB_LOOP_SYNCWARP
#define RTB_LOOP_SYNC() __syncwarp()
#else
#define RTB_LOOP_SYNC() ((void)0)
#endif

// Whether the wide-record (8-D) box kernel also keeps its stack pointer as a shared address
// (rt_traverse_box.cuh).  The narrow kernels always do; for the wide one ptxas then moves the pointer
// into the uniform datapath, which is measured separately (`make variants`: librtree_b200_relwide.so
// is the build with 0).
#ifndef RTB_ABS_SP_WIDE
#define RTB_ABS_SP_WIDE 1
#endif

// RTB_DEBUG_CHECKS = 1 (`make checked` -> librtree_b200_checked.so, never loaded unless RTB_LIB names it): the
// traversal kernels assert what their design argues -- a stack never outgrows 32 entries per level, a popped link
// names a block of the image, stash / list cursors stay inside their buffers -- and trap with a message
// otherwise.  compute-sanitizer is closed on this pool; the GPU suite run against this build stands in for it
// (profiles/r2_v6_pytest_gpu_checked_build.log).  0 in the product: the macro expands to nothing.
#ifndef RTB_DEBUG_CHECKS
#define RTB_DEBUG_CHECKS 0
#endif
#if RTB_DEBUG_CHECKS
#include <cstdio>
#define RTB_CHECK(cond, what)                                                                       \
  do                                                                                                \
  {                                                                                                 \
    if (!(cond))                                                                                    \
    {                                                                                               \
      printf("[rtb check] %s:%d block %d thread %d: %s\n", __FILE__, __LINE__, (int)blockIdx.x,    \
             (int)threadIdx.x, what);                                                               \
      __trap();                                                                                     \
    }                                                                                               \
  } while (0)
#else
#define RTB_CHECK(cond, what) ((void)0)
#endif

namespace rtb
{

// ---- slot records: structure of arrays of 16-byte vectors -------------------------------
// A record of WORDS 32-bit words is stored as WORDS/4 full rows (one uint4 per slot) plus a
// tail row of 4-, 8- or 16-byte elements; lane `slot` fetches its record with one vector
// load per row, the W lanes of a group covering W*16 contiguous bytes.
template <int WORDS, int W>
struct Record
{
  static constexpr int NV = WORDS / 4;
  static constexpr int R = WORDS % 4;
  static constexpr int PADDED = (WORDS + 3) / 4 * 4;
  static constexpr int TAIL_BYTES = R == 0 ? 0 : (R == 1 ? 4 : (R == 2 ? 8 : 16));
  static constexpr uint32_t BYTES = W * (16 * NV + TAIL_BYTES);

  // `lane_blk` = block address + slot * 16 (the lane's element in a 16-byte row)
  __device__ __forceinline__ static void load(const unsigned char* __restrict__ lane_blk, int slot,
                                              uint32_t (&w)[PADDED])
  {
#pragma unroll
    for (int k = 0; k < NV; ++k)
    {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(lane_blk + k * W * 16));
      w[4 * k + 0] = v.x, w[4 * k + 1] = v.y, w[4 * k + 2] = v.z, w[4 * k + 3] = v.w;
    }
    // tail elements are TAIL_BYTES wide: step back from the 16-byte lane offset
    const unsigned char* tail = lane_blk + NV * W * 16 - slot * (16 - TAIL_BYTES);
    if (R == 1)
    {
      w[4 * NV] = __ldg(reinterpret_cast<const uint32_t*>(tail));
    }
    else if (R == 2)
    {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(tail));
      w[4 * NV] = v.x, w[4 * NV + 1] = v.y;
    }
    else if (R == 3)
    {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(tail));
      w[4 * NV] = v.x, w[4 * NV + 1] = v.y, w[4 * NV + 2] = v.z, w[4 * NV + 3] = v.w;
    }
  }
};

template <typename S>
struct ScalarWords;
template <>
struct ScalarWords<float>
{
  static constexpr int N = 1;
  __device__ __forceinline__ static float get(const uint32_t* w) { return __uint_as_float(w[0]); }
};
template <>
struct ScalarWords<int32_t>
{
  static constexpr int N = 1;
  __device__ __forceinline__ static int32_t get(const uint32_t* w) { return (int32_t)w[0]; }
};
template <>
struct ScalarWords<double>
{
  static constexpr int N = 2;
  __device__ __forceinline__ static double get(const uint32_t* w) { return __hiloint2double((int)w[1], (int)w[0]); }
};

// Load the record of `slot` from the block whose lane address (block + slot * 16) is
// `blk` and split it into corners + link.  Record = lo corner, link, hi corner; point-keyed
// leaves store the prefix "corner, link" only (lo == hi), so both kinds read the first rows
// alike and only two-corner records fetch the rest.
template <typename S, int D, int W, bool BOXKEY>
__device__ __forceinline__ void load_entry(const unsigned char* __restrict__ blk, int slot, bool is_leaf,
                                           S (&lo)[D], S (&hi)[D], uint32_t& link)
{
  constexpr int SW = ScalarWords<S>::N;
  using Two = Record<2 * D * SW + SW, W>;
  using One = Record<D * SW + 1, W>;
  if (!BOXKEY && is_leaf)
  {
    uint32_t w[One::PADDED];
    One::load(blk, slot, w);
#pragma unroll
    for (int d = 0; d < D; ++d)
      lo[d] = hi[d] = ScalarWords<S>::get(w + d * SW);
    link = w[D * SW];
  }
  else
  {
    uint32_t w[Two::PADDED];
    Two::load(blk, slot, w);
#pragma unroll
    for (int d = 0; d < D; ++d)
    {
      lo[d] = ScalarWords<S>::get(w + d * SW);
      hi[d] = ScalarWords<S>::get(w + (D + 1 + d) * SW);
    }
    link = w[D * SW];
  }
}

// Hide a pointer's provenance from the optimiser so it stays in one register pair (ptxas
// otherwise re-adds the uniform base inside the traversal loop).
__device__ __forceinline__ const unsigned char* opaque(const unsigned char* p)
{
  asm volatile("" : "+l"(p));
  return p;
}

// ---- shared memory through 32-bit addresses (keeps the address math out of the loop) ----
__device__ __forceinline__ uint32_t smem_addr(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v)
{
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t sm_id()
{
  uint32_t id;
  asm("mov.u32 %0, %%smid;" : "=r"(id));
  return id;
}

// Claim the next chunk of queries of `range` for this warp (see WorkRanges): the position inside
// the range, at or beyond per_range when the range is used up.
__device__ __forceinline__ uint32_t claim_chunk(const WorkRanges& w, uint32_t range, int lane)
{
  uint32_t c = 0;
  if (lane == 0)
    c = atomicAdd(w.cursors + range, w.chunk);
  return __shfl_sync(0xFFFFFFFFu, c, 0);
}

// ascending bitonic sort of one value per lane; only the first `count` lanes hold values, the
// others RT_INVALID_ID (the largest key): with count <= 8 / 16 the network over the first 8 / 16 lanes
// already leaves everything in place, so the later merge stages are skipped (count is warp-uniform)
__device__ __forceinline__ uint32_t warp_sort32(uint32_t v, int lane, uint32_t count = 32u)
{
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1)
  {
    if (k == 16 && count <= 8u)
      break;
    if (k == 32 && count <= 16u)
      break;
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1)
    {
      const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, v, j);
      const bool ascending = (lane & k) == 0;
      const bool lower = (lane & j) == 0;
      v = (lower == ascending) ? min(v, other) : max(v, other);
    }
  }
  return v;
}

} // namespace rtb
