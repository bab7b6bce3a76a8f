// rt_api.cu -- the C ABI of include/rtree_b200.h, part 1: the handle.  rt_upload copies the tree
// image to the device (and builds the quantised image), rt_free releases everything; options, stats
// and the small accessors.  The query entry points live in rt_api_boxes.cu / rt_api_rays.cu /
// rt_api_knn.cu; no query-path arithmetic lives in any of them.
#include "rt_handle.cuh"

namespace
{
thread_local std::string g_last_error = "";
}

namespace rtb
{
int fail(int code, const std::string& what)
{
  g_last_error = what;
  return code;
}
} // namespace rtb

using namespace rtb;

namespace
{
int check_desc(const rt_tree_desc* d)
{
  if (d == nullptr)
    return fail(RT_E_INVALID, "rt_upload: desc is NULL");
  if (d->abi_version != RT_ABI_VERSION)
    return fail(RT_E_INVALID, "rt_upload: abi_version mismatch");
  if (d->dim < 1 || d->dim > 8)
    return fail(RT_E_INVALID, "rt_upload: dim must be 1..8");
  if (d->scalar > RT_I32)
    return fail(RT_E_INVALID, "rt_upload: unknown scalar type");
  if (d->width != 8 && d->width != 16)
    return fail(RT_E_UNSUPPORTED, "rt_upload: block width must be 8 or 16");
  if (d->key_kind > RT_KEY_BOX)
    return fail(RT_E_INVALID, "rt_upload: unknown key kind");
  if (d->num_levels != d->leaf_level + 1 || d->num_levels > 64)
    return fail(RT_E_INVALID, "rt_upload: num_levels must be leaf_level + 1 (<= 64)");
  if (d->level_node_begin == nullptr || d->blocks == nullptr)
    return fail(RT_E_INVALID, "rt_upload: NULL buffer in desc");
  if (d->num_nodes == 0 || d->level_node_begin[0] != 0 || d->level_node_begin[d->num_levels] != d->num_nodes)
    return fail(RT_E_INVALID, "rt_upload: level_node_begin inconsistent with num_nodes");
  for (uint32_t l = 0; l < d->num_levels; ++l)
  {
    if (d->level_node_begin[l] >= d->level_node_begin[l + 1])
      return fail(RT_E_INVALID, "rt_upload: empty level in level_node_begin");
  }
  if (d->level_node_begin[1] != 1)
    return fail(RT_E_INVALID, "rt_upload: level 0 must hold exactly the root");
  const uint32_t internal = rt_block_bytes(d->width, rt_record_words(d->dim, d->scalar, 1));
  const uint32_t leaf = rt_block_bytes(d->width, rt_record_words(d->dim, d->scalar, d->key_kind == RT_KEY_BOX));
  if (d->internal_stride != internal || d->leaf_stride != leaf)
    return fail(RT_E_INVALID, "rt_upload: block strides do not match rt_block_bytes() for this tree");
  if (d->leaf_offset % 128)
    return fail(RT_E_INVALID, "rt_upload: leaf_offset must be a multiple of 128");
  const uint32_t leaf_begin = d->level_node_begin[d->leaf_level];
  const uint64_t need_internal = (uint64_t)leaf_begin * d->internal_stride;
  const uint64_t need = d->leaf_offset + (uint64_t)(d->num_nodes - leaf_begin) * d->leaf_stride;
  if (d->leaf_offset < need_internal || d->blocks_bytes < need)
    return fail(RT_E_INVALID, "rt_upload: blocks_bytes smaller than the layout needs");
  if ((d->blocks_bytes + 4096) / 16 >= RT_INVALID_ID)
    return fail(RT_E_INVALID, "rt_upload: tree image larger than 64 GB");
  if (d->num_values >= RT_INVALID_ID)
    return fail(RT_E_INVALID, "rt_upload: too many values");
  return RT_OK;
}
} // namespace

extern "C"
{

const char* rt_last_error(void) { return g_last_error.c_str(); }
uint32_t rt_abi_version(void) { return RT_ABI_VERSION; }

int rt_device_count(void)
{
  int n = 0;
  RT_CUDA(cudaGetDeviceCount(&n));
  return n;
}

int rt_upload(const rt_tree_desc* desc, int device, rt_tree** out)
{
  if (out == nullptr)
    return fail(RT_E_INVALID, "rt_upload: out is NULL");
  *out = nullptr;
  int rc = check_desc(desc);
  if (rc != RT_OK)
    return rc;
  if (rtb::find_box_kernel(desc->scalar, desc->dim, desc->width, desc->key_kind, RT_INTERSECTS) == nullptr)
    return fail(RT_E_UNSUPPORTED, "rt_upload: no kernel instantiated for this (scalar, dim, width)");
  int count = 0;
  RT_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count)
    return fail(RT_E_CUDA, "rt_upload: no such CUDA device");
  RT_CUDA(cudaSetDevice(device));

  rt_tree* t = new (std::nothrow) rt_tree();
  if (t == nullptr)
    return fail(RT_E_NOMEM, "rt_upload: out of host memory");
  t->device = device;
  t->desc = *desc;
  t->levels.assign(desc->level_node_begin, desc->level_node_begin + desc->num_levels + 1);
  t->desc.level_node_begin = t->levels.data();
  std::memset(&t->stats, 0, sizeof t->stats);
  if (const char* batch = std::getenv("RTB_PIPELINE_BATCH"))
  {
    const uint64_t b = std::strtoull(batch, nullptr, 10);
    t->pipeline_batch = (b != 0 && b < MIN_PIPELINE_BATCH) ? MIN_PIPELINE_BATCH : b;
  }

  auto cleanup = [&](int code, const std::string& what)
  {
    rt_free(t);
    return fail(code, what);
  };
  // the image is followed by one "null" block whose slots are all empty (every word 0xFFFFFFFF:
  // link = RT_INVALID_ID); idle lane groups of the traversal kernels read it instead of branching
  const uint64_t null_offset = (desc->blocks_bytes + 127) / 128 * 128;
  const uint64_t null_bytes = desc->internal_stride > desc->leaf_stride ? desc->internal_stride : desc->leaf_stride;
  cudaError_t e = cudaMalloc((void**)&t->d_blocks, null_offset + null_bytes);
  if (e != cudaSuccess)
    return cleanup(RT_E_NOMEM, std::string("rt_upload: cudaMalloc(blocks): ") + cudaGetErrorString(e));
  e = cudaStreamCreateWithFlags(&t->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess)
    return cleanup(RT_E_CUDA, std::string("rt_upload: stream: ") + cudaGetErrorString(e));
  t->stream = t->own_stream;
  for (Slot& s : t->slot)
  {
    if ((e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess
        || (e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming)) != cudaSuccess)
      return cleanup(RT_E_CUDA, std::string("rt_upload: slot stream: ") + cudaGetErrorString(e));
  }
  if ((e = cudaEventCreate(&t->ev_fork)) != cudaSuccess || (e = cudaEventCreate(&t->ev_join)) != cudaSuccess)
    return cleanup(RT_E_CUDA, std::string("rt_upload: event: ") + cudaGetErrorString(e));
  e = cudaMemcpyAsync(t->d_blocks, desc->blocks, desc->blocks_bytes,
                      desc->blocks_memory == RT_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                      t->stream);
  if (e == cudaSuccess)
    e = cudaMemsetAsync(t->d_blocks + null_offset, 0xFF, null_bytes, t->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(t->stream);
  if (e != cudaSuccess)
    return cleanup(RT_E_CUDA, std::string("rt_upload: copy of blocks: ") + cudaGetErrorString(e));

  t->desc.blocks = t->d_blocks;
  t->desc.blocks_memory = RT_MEM_DEVICE;
  t->dev.blocks = t->d_blocks;
  t->dev.leaf_link_begin = (uint32_t)(desc->leaf_offset / 16);
  t->dev.null_link = (uint32_t)(null_offset / 16);
  t->dev.num_nodes = desc->num_nodes;
  t->dev.num_levels = desc->num_levels;
  t->dev.stage_links = 0;
  if (rtb::quant_supported(desc->scalar, desc->dim, desc->width, desc->key_kind)
      && std::getenv("RTB_NO_QUANT") == nullptr)
  {
    e = cudaMalloc((void**)&t->d_qblocks, rtb::quant_image_bytes(desc->num_nodes, desc->dim, desc->width));
    if (e != cudaSuccess)
      return cleanup(RT_E_NOMEM, std::string("rt_upload: cudaMalloc(quantised image): ") + cudaGetErrorString(e));
    e = rtb::build_quant_image(t->dev, t->desc, desc->level_node_begin[desc->leaf_level], t->d_qblocks, t->stream);
    if (e != cudaSuccess)
      return cleanup(RT_E_CUDA, std::string("rt_upload: quantised image: ") + cudaGetErrorString(e));
    // levels 0..2 (at most 128 nodes): what the RTB_STAGE_TOP experiment keeps in shared memory
    const uint32_t top = desc->level_node_begin[desc->num_levels < 3 ? desc->num_levels : 3];
    t->dev.stage_links = (top < 128 ? top : 128) * (rt_block_bytes(desc->width, desc->dim + 1) / 16u);
  }
  *out = t;
  return RT_OK;
}

void rt_free(rt_tree* t)
{
  if (t == nullptr)
    return;
  cudaSetDevice(t->device);
  if (t->own_stream)
    cudaStreamSynchronize(t->own_stream);
  for (Slot& s : t->slot)
  {
    if (s.stream)
      cudaStreamSynchronize(s.stream);
    s.release();
  }
  DevBuf* dev[] = { &t->d_ray_t, &t->d_ray_id, &t->d_ray_any, &t->d_knn_ids, &t->d_knn_d2 };
  for (DevBuf* b : dev)
    b->release();
  HostBuf* host[] = { &t->h_offsets, &t->h_counts, &t->h_ids,  &t->h_ray_t,    &t->h_ray_id, &t->h_ray_any, &t->h_batch_counters,
                      &t->h_knn_ids, &t->h_knn_d2 };
  for (HostBuf* b : host)
    b->release();
  if (t->d_blocks)
    cudaFree(t->d_blocks);
  if (t->d_qblocks)
    cudaFree(t->d_qblocks);
  for (cudaEvent_t e : t->events)
    cudaEventDestroy(e);
  if (t->ev_fork)
    cudaEventDestroy(t->ev_fork);
  if (t->ev_join)
    cudaEventDestroy(t->ev_join);
  if (t->own_stream)
    cudaStreamDestroy(t->own_stream);
  cudaGetLastError();
  delete t;
}

int rt_set_stream(rt_tree* t, void* cuda_stream)
{
  if (t == nullptr)
    return fail(RT_E_INVALID, "rt_set_stream: tree is NULL");
  t->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : t->own_stream;
  return RT_OK;
}

int rt_set_option(rt_tree* t, uint32_t option, uint64_t value)
{
  if (t == nullptr)
    return fail(RT_E_INVALID, "rt_set_option: tree is NULL");
  switch (option)
  {
  case RT_OPT_PIPELINE_BATCH:
    if (value != 0 && value < MIN_PIPELINE_BATCH)
      return fail(RT_E_INVALID, "rt_set_option: RT_OPT_PIPELINE_BATCH must be 0 (never pipeline) or at least 4096");
    t->pipeline_batch = value;
    return RT_OK;
  case RT_OPT_QUERY_CHUNK: t->query_chunk = (uint32_t)value; return RT_OK;
  case RT_OPT_WORK_RANGES: t->work_ranges = (uint32_t)value; return RT_OK;
  case RT_OPT_TRAVERSE_GRID: t->traverse_grid = (int)value; return RT_OK;
  case RT_OPT_QUANTISED: t->use_quant = value != 0; return RT_OK;
  default: return fail(RT_E_INVALID, "rt_set_option: unknown option");
  }
}

int rt_get_stats(const rt_tree* t, rt_stats* out)
{
  if (t == nullptr || out == nullptr)
    return fail(RT_E_INVALID, "rt_get_stats: NULL argument");
  *out = t->stats;
  return RT_OK;
}

int rt_tree_device_blocks(const rt_tree* t, void** device_ptr, uint64_t* bytes)
{
  if (t == nullptr || device_ptr == nullptr || bytes == nullptr)
    return fail(RT_E_INVALID, "rt_tree_device_blocks: NULL argument");
  *device_ptr = t->d_blocks;
  *bytes = t->desc.blocks_bytes;
  return RT_OK;
}

int rt_tree_get_desc(const rt_tree* t, rt_tree_desc* out)
{
  if (t == nullptr || out == nullptr)
    return fail(RT_E_INVALID, "rt_tree_get_desc: NULL argument");
  *out = t->desc;
  return RT_OK;
}

} // extern "C"

namespace rtb
{
box_launch_fn find_box_kernel(uint32_t scalar, uint32_t dim, uint32_t width, uint32_t key_kind, uint32_t mode)
{
  switch (scalar)
  {
  case RT_F32: return find_box_kernel_f32(dim, width, key_kind, mode);
  case RT_F64: return find_box_kernel_f64(dim, width, key_kind, mode);
  case RT_I32: return find_box_kernel_i32(dim, width, key_kind, mode);
  default: return nullptr;
  }
}
} // namespace rtb
