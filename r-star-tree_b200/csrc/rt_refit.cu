// rt_refit.cu -- rt_refit(): new keys into the leaf slots, then every bound recomputed bottom-up,
// on the device; the topology (links, slot occupancy) is untouched.
//
// The reference's way of "dealing with moving objects" (reference README.md:342-347) is to change
// keys in place through the iterators and call RTree::rebound(iterator)
// (include/RTree/rtree.hpp:479-500), which recomputes the bound of the leaf and of every ancestor as
// the union of its children's bounds.  Done for every value that is: every slot bound = component-wise
// min / max over the child's entries, level by level from the leaves up -- exact arithmetic, so the
// refitted image equals flatten_device() of the refitted host tree byte for byte (tested).
//
// Keys arrive in flatten_result_t::data order (= iteration order of RTree::begin()..end(), which is
// the left-to-right order of the leaves and therefore the order of the leaf blocks): entry c of leaf j
// has data index first[j] + c, first[] = exclusive scan of the leaves' entry counts.
#include "rt_common.cuh"

namespace rtb
{
namespace
{
struct Layout
{
  unsigned char* blocks;
  uint64_t leaf_offset;
  uint32_t internal_stride, leaf_stride, internal_words, leaf_words;
  uint32_t width, dim, key_is_box, leaf_link_begin;
};

// byte offset of word j of slot c in a block of `words`-word records (include/rtree_b200.h)
__device__ __forceinline__ uint32_t word_offset(uint32_t width, uint32_t words, uint32_t c, uint32_t j)
{
  const uint32_t full = words / 4, r = words % 4;
  if (j < 4 * full)
    return (j / 4) * width * 16 + c * 16 + (j % 4) * 4;
  const uint32_t tail = r == 1 ? 4u : (r == 2 ? 8u : 16u);
  return full * width * 16 + c * tail + (j - 4 * full) * 4;
}

template <typename S>
__device__ __forceinline__ S load_scalar(const unsigned char* block, uint32_t width, uint32_t words, uint32_t c,
                                         uint32_t j)
{
  constexpr uint32_t SW = sizeof(S) / 4;
  uint32_t raw[SW];
#pragma unroll
  for (uint32_t k = 0; k < SW; ++k)
    raw[k] = *reinterpret_cast<const uint32_t*>(block + word_offset(width, words, c, j + k));
  S v;
  memcpy(&v, raw, sizeof v);
  return v;
}
template <typename S>
__device__ __forceinline__ void store_scalar(unsigned char* block, uint32_t width, uint32_t words, uint32_t c,
                                             uint32_t j, S v)
{
  constexpr uint32_t SW = sizeof(S) / 4;
  uint32_t raw[SW];
  memcpy(raw, &v, sizeof v);
#pragma unroll
  for (uint32_t k = 0; k < SW; ++k)
    *reinterpret_cast<uint32_t*>(block + word_offset(width, words, c, j + k)) = raw[k];
}

// entries per leaf (slots are filled from 0, empty ones carry RT_INVALID_ID)
template <typename S>
__global__ void __launch_bounds__(256) refit_leaf_counts(Layout L, uint32_t n_leaves, uint32_t* counts)
{
  constexpr uint32_t SW = sizeof(S) / 4;
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_leaves)
    return;
  const unsigned char* block = L.blocks + L.leaf_offset + (uint64_t)j * L.leaf_stride;
  uint32_t n = 0;
  for (uint32_t c = 0; c < L.width; ++c)
    n += *reinterpret_cast<const uint32_t*>(block + word_offset(L.width, L.leaf_words, c, L.dim * SW)) != RT_INVALID_ID;
  counts[j] = n;
}

// one thread per (leaf, slot): the new key of that entry
template <typename S>
__global__ void __launch_bounds__(256) refit_scatter_keys(Layout L, uint32_t n_leaves, const unsigned long long* first,
                                                          const S* keys)
{
  constexpr uint32_t SW = sizeof(S) / 4;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t j = (uint32_t)(tid / L.width), c = (uint32_t)(tid % L.width);
  if (j >= n_leaves)
    return;
  unsigned char* block = L.blocks + L.leaf_offset + (uint64_t)j * L.leaf_stride;
  if (*reinterpret_cast<const uint32_t*>(block + word_offset(L.width, L.leaf_words, c, L.dim * SW)) == RT_INVALID_ID)
    return;
  const uint64_t index = first[j] + c;
  const S* key = keys + index * L.dim * (L.key_is_box ? 2 : 1);
  for (uint32_t d = 0; d < L.dim; ++d)
  {
    store_scalar<S>(block, L.width, L.leaf_words, c, d * SW, key[d]);
    if (L.key_is_box)
      store_scalar<S>(block, L.width, L.leaf_words, c, (L.dim + 1 + d) * SW, key[L.dim + d]);
  }
}

// one thread per (internal node of one level, slot): bound = union of the child's entries
template <typename S>
__global__ void __launch_bounds__(256) refit_level(Layout L, uint32_t node_begin, uint32_t node_end)
{
  constexpr uint32_t SW = sizeof(S) / 4;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t node = node_begin + (uint32_t)(tid / L.width), c = (uint32_t)(tid % L.width);
  if (node >= node_end)
    return;
  unsigned char* block = L.blocks + (uint64_t)node * L.internal_stride;
  const uint32_t link = *reinterpret_cast<const uint32_t*>(block + word_offset(L.width, L.internal_words, c, L.dim * SW));
  if (link == RT_INVALID_ID)
    return;
  const unsigned char* child = L.blocks + (uint64_t)link * 16u;
  const bool child_is_leaf = link >= L.leaf_link_begin;
  const uint32_t cwords = child_is_leaf ? L.leaf_words : L.internal_words;
  const bool one_corner = child_is_leaf && !L.key_is_box;
  for (uint32_t d = 0; d < L.dim; ++d)
  {
    bool any = false;
    S lo = S(), hi = S();
    for (uint32_t e = 0; e < L.width; ++e)
    {
      if (*reinterpret_cast<const uint32_t*>(child + word_offset(L.width, cwords, e, L.dim * SW)) == RT_INVALID_ID)
        continue;
      const S elo = load_scalar<S>(child, L.width, cwords, e, d * SW);
      const S ehi = one_corner ? elo : load_scalar<S>(child, L.width, cwords, e, (L.dim + 1 + d) * SW);
      // the reference's aabb merge: min of the mins, max of the maxes (first entry initialises)
      lo = (!any || elo < lo) ? elo : lo;
      hi = (!any || ehi > hi) ? ehi : hi;
      any = true;
    }
    if (any)
    {
      store_scalar<S>(block, L.width, L.internal_words, c, d * SW, lo);
      store_scalar<S>(block, L.width, L.internal_words, c, (L.dim + 1 + d) * SW, hi);
    }
  }
}

template <typename S>
cudaError_t refit_typed(const Layout& L, const uint32_t* levels, uint32_t num_levels, const void* keys,
                        uint32_t* counts, unsigned long long* first, void* scan_scratch, cudaStream_t stream,
                        uint32_t* launches)
{
  const uint32_t leaf_begin = levels[num_levels - 1], n_leaves = levels[num_levels] - leaf_begin;
  refit_leaf_counts<S><<<(n_leaves + 255) / 256, 256, 0, stream>>>(L, n_leaves, counts);
  *launches += 1;
  cudaError_t err = launch_exclusive_scan(counts, n_leaves, first, scan_scratch, stream, launches);
  if (err != cudaSuccess)
    return err;
  const uint64_t leaf_threads = (uint64_t)n_leaves * L.width;
  refit_scatter_keys<S><<<(unsigned)((leaf_threads + 255) / 256), 256, 0, stream>>>(L, n_leaves, first,
                                                                                 static_cast<const S*>(keys));
  *launches += 1;
  for (uint32_t level = num_levels - 1; level-- > 0;)
  {
    const uint64_t threads = (uint64_t)(levels[level + 1] - levels[level]) * L.width;
    refit_level<S><<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(L, levels[level], levels[level + 1]);
    *launches += 1;
  }
  return cudaGetLastError();
}
} // namespace

size_t refit_scratch_bytes(uint32_t n_leaves)
{
  return ((size_t)n_leaves + 4) * sizeof(uint32_t) + ((size_t)n_leaves + 2) * sizeof(unsigned long long) + 512
         + scan_scratch_bytes(n_leaves);
}

cudaError_t launch_refit(unsigned char* blocks, const rt_tree_desc& d, const uint32_t* levels, const void* keys,
                         void* scratch, cudaStream_t stream, uint32_t* launches)
{
  Layout L;
  L.blocks = blocks;
  L.leaf_offset = d.leaf_offset;
  L.internal_stride = d.internal_stride;
  L.leaf_stride = d.leaf_stride;
  L.internal_words = rt_record_words(d.dim, d.scalar, 1);
  L.leaf_words = rt_record_words(d.dim, d.scalar, d.key_kind == RT_KEY_BOX);
  L.width = d.width;
  L.dim = d.dim;
  L.key_is_box = d.key_kind == RT_KEY_BOX;
  L.leaf_link_begin = (uint32_t)(d.leaf_offset / 16);
  const uint32_t n_leaves = levels[d.num_levels] - levels[d.num_levels - 1];
  unsigned char* base = static_cast<unsigned char*>(scratch);
  uint32_t* counts = reinterpret_cast<uint32_t*>(base);
  base += (((size_t)n_leaves + 4) * sizeof(uint32_t) + 255) / 256 * 256;
  unsigned long long* first = reinterpret_cast<unsigned long long*>(base);
  base += (((size_t)n_leaves + 2) * sizeof(unsigned long long) + 255) / 256 * 256;
  switch (d.scalar)
  {
  case RT_F32: return refit_typed<float>(L, levels, d.num_levels, keys, counts, first, base, stream, launches);
  case RT_F64: return refit_typed<double>(L, levels, d.num_levels, keys, counts, first, base, stream, launches);
  case RT_I32: return refit_typed<int32_t>(L, levels, d.num_levels, keys, counts, first, base, stream, launches);
  default: return cudaErrorInvalidValue;
  }
}
} // namespace rtb
