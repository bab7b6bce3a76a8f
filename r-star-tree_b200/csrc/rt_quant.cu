// rt_quant.cu -- builds the quantised image (rt_quant.cuh) from the uploaded float blocks, on the
// device, with the same quantiser the traversal applies to the queries.
#include "rt_quant.cuh"

namespace rtb
{
namespace
{
struct Grid
{
  float lo[3], scale[3];
};

// one warp: the grid spans the bounds of the root's children (= of all keys)
__global__ void quant_grid_kernel(const unsigned char* blocks, bool root_is_leaf, uint64_t leaf_offset, Grid* grid)
{
  const int lane = threadIdx.x;
  float lo[3], hi[3];
  bool valid = false;
  if (lane < 8)
  {
    const unsigned char* blk = blocks + (root_is_leaf ? leaf_offset : 0);
    const uint4 r0 = *reinterpret_cast<const uint4*>(blk + lane * 16);
    valid = r0.w != RT_INVALID_ID;
    uint4 r1 = r0;
    if (!root_is_leaf)
      r1 = *reinterpret_cast<const uint4*>(blk + 128 + lane * 16);
    lo[0] = __uint_as_float(r0.x), lo[1] = __uint_as_float(r0.y), lo[2] = __uint_as_float(r0.z);
    hi[0] = __uint_as_float(r1.x), hi[1] = __uint_as_float(r1.y), hi[2] = __uint_as_float(r1.z);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d)
  {
    float l = valid ? lo[d] : __uint_as_float(0x7F800000u);
    float h = valid ? hi[d] : __uint_as_float(0xFF800000u);
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

// one thread per (node, slot)
__global__ void __launch_bounds__(256) quant_build_kernel(const unsigned char* blocks, uint32_t num_nodes,
                                                          uint32_t num_internal, uint64_t leaf_offset,
                                                          const Grid* grid, unsigned char* qblocks)
{
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t node = (uint32_t)(tid >> 3), slot = (uint32_t)(tid & 7);
  if (node > num_nodes) // node == num_nodes: the null block
    return;
  uint4 out = make_uint4(QUANT_GUARD, QUANT_GUARD, QUANT_GUARD, RT_INVALID_ID); // an empty slot
  if (node < num_nodes)
  {
    const bool leaf = node >= num_internal;
    const unsigned char* blk = leaf ? blocks + leaf_offset + (uint64_t)(node - num_internal) * 128
                                    : blocks + (uint64_t)node * 256;
    const uint4 r0 = *reinterpret_cast<const uint4*>(blk + slot * 16);
    if (r0.w != RT_INVALID_ID)
    {
      uint4 r1 = r0;
      if (!leaf)
        r1 = *reinterpret_cast<const uint4*>(blk + 128 + slot * 16);
      const float lo[3] = { __uint_as_float(r0.x), __uint_as_float(r0.y), __uint_as_float(r0.z) };
      const float hi[3] = { __uint_as_float(r1.x), __uint_as_float(r1.y), __uint_as_float(r1.z) };
      const float glo[3] = { grid->lo[0], grid->lo[1], grid->lo[2] };
      const float gscale[3] = { grid->scale[0], grid->scale[1], grid->scale[2] };
      uint32_t a[3];
      quant_slot(lo, hi, glo, gscale, a);
      uint32_t link = r0.w; // leaf entry: the id
      if (!leaf)
      {
        // child link of the float image (byte offset / 16) -> link of the quantised image
        const uint32_t leaf_link = (uint32_t)(leaf_offset / 16);
        link = r0.w < leaf_link ? r0.w / 2                                   // internal: 256 -> 128 bytes
                                : num_internal * (QUANT_BLOCK / 16) + (r0.w - leaf_link); // leaf: same stride
      }
      out = make_uint4(a[0], a[1], a[2], link);
    }
  }
  *reinterpret_cast<uint4*>(qblocks + (uint64_t)node * QUANT_BLOCK + slot * 16) = out;
}
} // namespace

size_t quant_image_bytes(uint32_t num_nodes)
{
  return ((size_t)num_nodes + 1) * QUANT_BLOCK + sizeof(Grid);
}

cudaError_t build_quant_image(TreeDev& tree, uint32_t num_internal, uint64_t leaf_offset, unsigned char* qblocks,
                              cudaStream_t stream)
{
  Grid* d_grid = reinterpret_cast<Grid*>(qblocks + ((size_t)tree.num_nodes + 1) * QUANT_BLOCK);
  quant_grid_kernel<<<1, 32, 0, stream>>>(tree.blocks, num_internal == 0, leaf_offset, d_grid);
  const uint64_t threads = ((uint64_t)tree.num_nodes + 1) * 8;
  quant_build_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(tree.blocks, tree.num_nodes, num_internal,
                                                                         leaf_offset, d_grid, qblocks);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess)
    return err;
  Grid h;
  if ((err = cudaMemcpyAsync(&h, d_grid, sizeof h, cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
    return err;
  if ((err = cudaStreamSynchronize(stream)) != cudaSuccess)
    return err;
  for (int d = 0; d < 3; ++d)
  {
    tree.qlo[d] = h.lo[d];
    tree.qscale[d] = h.scale[d];
  }
  tree.qblocks = qblocks;
  tree.leaf_blocks = tree.blocks + leaf_offset;
  tree.qleaf_link_begin = num_internal * (QUANT_BLOCK / 16);
  tree.qnull_link = tree.num_nodes * (QUANT_BLOCK / 16);
  return cudaSuccess;
}
} // namespace rtb
