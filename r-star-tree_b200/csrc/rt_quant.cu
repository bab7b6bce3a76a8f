// rt_quant.cu -- builds the quantised image (rt_quant.cuh) from the uploaded float blocks, on the
// device, with the same quantiser the traversal applies to the queries.
#include "rt_quant.cuh"

namespace rtb
{
namespace
{
struct Grid
{
  float lo[8], scale[8];
};

struct Shape
{
  const unsigned char* blocks; // the float image
  uint64_t leaf_offset;
  uint32_t internal_stride, leaf_stride, internal_words, leaf_words;
  uint32_t width, dim, num_nodes, num_internal;
  uint32_t qstride; // bytes per node in the quantised image (internal and leaf alike)
};

__device__ __forceinline__ float float_at(const unsigned char* block, uint32_t width, uint32_t words, uint32_t c,
                                          uint32_t j)
{
  return __uint_as_float(*reinterpret_cast<const uint32_t*>(block + block_word_offset(width, words, c, j)));
}

// one warp: the grid spans the bounds of the root's children (= of all keys)
__global__ void quant_grid_kernel(Shape s, Grid* grid)
{
  const uint32_t lane = threadIdx.x;
  const bool root_is_leaf = s.num_internal == 0;
  const unsigned char* block = s.blocks + (root_is_leaf ? s.leaf_offset : 0);
  const uint32_t words = root_is_leaf ? s.leaf_words : s.internal_words;
  const bool valid = lane < s.width
                     && *reinterpret_cast<const uint32_t*>(block + block_word_offset(s.width, words, lane, s.dim))
                            != RT_INVALID_ID;
  for (uint32_t d = 0; d < s.dim; ++d)
  {
    float l = __uint_as_float(0x7F800000u), h = __uint_as_float(0xFF800000u);
    if (valid)
    {
      l = float_at(block, s.width, words, lane, d);
      h = root_is_leaf ? l : float_at(block, s.width, words, lane, s.dim + 1 + d);
    }
    for (int k = 16; k > 0; k >>= 1)
    {
      l = fminf(l, __shfl_xor_sync(0xFFFFFFFFu, l, k));
      h = fmaxf(h, __shfl_xor_sync(0xFFFFFFFFu, h, k));
    }
    if (lane == 0)
    {
      const float extent = h - l;
      // an empty, degenerate or unbounded axis gets scale 0: every value lands in cell 1 and the
      // axis never rejects anything (still conservative)
      const bool usable = extent > 0.0f && extent < __uint_as_float(0x7F800000u);
      grid->lo[d] = usable ? l : 0.0f;
      grid->scale[d] = usable ? 32766.0f / extent : 0.0f;
    }
  }
}

// one thread per (node, slot); node == num_nodes is the null block
template <int D>
__global__ void __launch_bounds__(256) quant_build_kernel(Shape s, const Grid* grid, unsigned char* qblocks)
{
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t node = (uint32_t)(tid / s.width), slot = (uint32_t)(tid % s.width);
  if (node > s.num_nodes)
    return;
  // wide records (8-D): leaf slots are stored packed, see below; their empty form is the same word
  constexpr bool PACKED_LEAVES = D == 8;
  uint32_t a[D];
#pragma unroll
  for (int j = 0; j < D; ++j)
    a[j] = QUANT_GUARD; // an empty slot: a = 0 in every half (a packed leaf slot: s-half 0x8000, never passes)
  uint32_t link = RT_INVALID_ID;
  if (node < s.num_nodes)
  {
    const bool leaf = node >= s.num_internal;
    const unsigned char* block = leaf ? s.blocks + s.leaf_offset + (uint64_t)(node - s.num_internal) * s.leaf_stride
                                      : s.blocks + (uint64_t)node * s.internal_stride;
    const uint32_t words = leaf ? s.leaf_words : s.internal_words;
    const uint32_t raw = *reinterpret_cast<const uint32_t*>(block + block_word_offset(s.width, words, slot, D));
    if (raw != RT_INVALID_ID)
    {
      float lo[D], hi[D];
#pragma unroll
      for (int d = 0; d < D; ++d)
      {
        lo[d] = float_at(block, s.width, words, slot, d);
        hi[d] = leaf ? lo[d] : float_at(block, s.width, words, slot, D + 1 + d);
      }
      quant_slot<D>(lo, hi, grid->lo, grid->scale, a);
      if (PACKED_LEAVES && leaf)
      {
        // a point says the same thing in both halves of its record: keep only the q-words
        // G | q(p_2j) | q(p_2j+1) << 16 (the first D/2 of quant_slot's words, which is what `a` holds
        // already); the traversal derives the n-words (rt_traverse_box.cuh), so they are zeroed
#pragma unroll
        for (int j = D / 2; j < D; ++j)
          a[j] = 0u;
      }
      link = raw; // leaf entry: the id
      if (!leaf)
      {
        // child link of the float image (byte offset / 16) -> breadth-first node index -> link of
        // the quantised image, where every node has the same stride
        const uint64_t offset = (uint64_t)raw * 16u;
        const uint32_t child = offset < s.leaf_offset
                                   ? (uint32_t)(offset / s.internal_stride)
                                   : s.num_internal + (uint32_t)((offset - s.leaf_offset) / s.leaf_stride);
        link = child * (s.qstride / 16u);
      }
    }
  }
  unsigned char* out = qblocks + (uint64_t)node * s.qstride;
#pragma unroll
  for (int j = 0; j < D; ++j)
    *reinterpret_cast<uint32_t*>(out + block_word_offset(s.width, D + 1, slot, j)) = a[j];
  *reinterpret_cast<uint32_t*>(out + block_word_offset(s.width, D + 1, slot, D)) = link;
}

template <int D>
void launch_build(const Shape& s, const Grid* grid, unsigned char* qblocks, cudaStream_t stream)
{
  const uint64_t threads = ((uint64_t)s.num_nodes + 1) * s.width;
  quant_build_kernel<D><<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(s, grid, qblocks);
}
} // namespace

// the (dim, width) pairs box_traverse_kernel<..., QUANT> is instantiated for (rt_traverse_box.cuh)
bool quant_supported(uint32_t scalar, uint32_t dim, uint32_t width, uint32_t key_kind)
{
  if (scalar != RT_F32 || key_kind != RT_KEY_POINT)
    return false;
  return (dim == 2 && width == 8) || (dim == 3 && width == 8) || (dim == 8 && width == 16);
}

size_t quant_image_bytes(uint32_t num_nodes, uint32_t dim, uint32_t width)
{
  return ((size_t)num_nodes + 1) * rt_block_bytes(width, dim + 1) + sizeof(Grid) + 256;
}

cudaError_t build_quant_image(TreeDev& tree, const rt_tree_desc& desc, uint32_t num_internal, unsigned char* qblocks,
                              cudaStream_t stream)
{
  Shape s;
  s.blocks = tree.blocks;
  s.leaf_offset = desc.leaf_offset;
  s.internal_stride = desc.internal_stride;
  s.leaf_stride = desc.leaf_stride;
  s.internal_words = rt_record_words(desc.dim, desc.scalar, 1);
  s.leaf_words = rt_record_words(desc.dim, desc.scalar, 0);
  s.width = desc.width;
  s.dim = desc.dim;
  s.num_nodes = desc.num_nodes;
  s.num_internal = num_internal;
  s.qstride = rt_block_bytes(desc.width, desc.dim + 1);
  const size_t image = ((size_t)desc.num_nodes + 1) * s.qstride;
  Grid* d_grid = reinterpret_cast<Grid*>(qblocks + (image + 255) / 256 * 256);
  quant_grid_kernel<<<1, 32, 0, stream>>>(s, d_grid);
  switch (desc.dim)
  {
  case 2: launch_build<2>(s, d_grid, qblocks, stream); break;
  case 3: launch_build<3>(s, d_grid, qblocks, stream); break;
  case 8: launch_build<8>(s, d_grid, qblocks, stream); break;
  default: return cudaErrorInvalidValue;
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess)
    return err;
  Grid h;
  if ((err = cudaMemcpyAsync(&h, d_grid, sizeof h, cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
    return err;
  if ((err = cudaStreamSynchronize(stream)) != cudaSuccess)
    return err;
  for (uint32_t d = 0; d < 8; ++d)
  {
    tree.qlo[d] = d < desc.dim ? h.lo[d] : 0.0f;
    tree.qscale[d] = d < desc.dim ? h.scale[d] : 0.0f;
  }
  tree.qblocks = qblocks;
  tree.leaf_blocks = tree.blocks + desc.leaf_offset;
  tree.qleaf_link_begin = num_internal * (s.qstride / 16u);
  tree.qnull_link = desc.num_nodes * (s.qstride / 16u);
  return cudaSuccess;
}
} // namespace rtb
