// rt_api_knn.cu -- rt_query_knn and rt_refit: buffers and copies around their kernels.
#include "rt_handle.cuh"

using namespace rtb;

namespace
{
int run_knn_call(rt_tree* t, const float* points, uint64_t n, uint32_t k, uint32_t flags, rt_knn_result* out);
int run_refit_call(rt_tree* t, const void* keys, uint64_t n_keys, uint32_t flags);
} // namespace

extern "C"
{

int rt_query_knn(rt_tree* t, const float* points, uint64_t n, uint32_t k, uint32_t flags, rt_knn_result* out)
{
  const int rc = run_knn_call(t, points, n, k, flags, out);
  // argument errors come back before anything was issued; anything else may have left work in flight
  if (rc != RT_OK && rc != RT_E_INVALID && rc != RT_E_UNSUPPORTED && t != nullptr)
    return abandon_call(t, rc);
  return rc;
}

int rt_refit(rt_tree* t, const void* keys, uint64_t n_keys, uint32_t flags)
{
  const int rc = run_refit_call(t, keys, n_keys, flags);
  if (rc != RT_OK && rc != RT_E_INVALID && rc != RT_E_UNSUPPORTED && t != nullptr)
    return abandon_call(t, rc);
  return rc;
}

} // extern "C"

namespace
{
int run_knn_call(rt_tree* t, const float* points, uint64_t n, uint32_t k, uint32_t flags, rt_knn_result* out)
{
  if (t == nullptr || out == nullptr)
    return fail(RT_E_INVALID, "rt_query_knn: NULL argument");
  if (points == nullptr && n > 0)
    return fail(RT_E_INVALID, "rt_query_knn: points is NULL");
  if (k < 1 || k > 32)
    return fail(RT_E_INVALID, "rt_query_knn: k must be 1..32");
  if (!rtb::knn_supported(t->desc.scalar, t->desc.dim, t->desc.width))
    return fail(RT_E_UNSUPPORTED, "rt_query_knn: kNN queries need a float32 tree (dim 1-4 or 8 with 8-wide blocks, "
                                  "dim 2, 3 or 8 with 16-wide blocks)");
  if (n >= 0xFFFFFFFFull)
    return fail(RT_E_INVALID, "rt_query_knn: at most 2^32-2 queries per call");
  int rc = use_device(t);
  if (rc != RT_OK)
    return rc;

  std::memset(out, 0, sizeof *out);
  std::memset(&t->stats, 0, sizeof t->stats);
  cudaStream_t st = t->stream;
  Slot& w = t->slot[0];
  cudaEvent_t* ev = nullptr;
  if ((rc = batch_events(t, 0, &ev)) != RT_OK)
    return rc;
  const uint32_t dim = t->desc.dim;
  const size_t point_bytes = (size_t)n * dim * sizeof(float);
  const size_t out_elems = (size_t)n * k;
  const bool on_device = (flags & RT_QUERIES_ON_DEVICE) != 0;
  const bool keep_on_device = (flags & RT_RESULT_ON_DEVICE) != 0;
  const bool collect = (flags & RT_COLLECT_VISITS) != 0;
  const bool reorder = (flags & RT_NO_REORDER) == 0 && n >= 4096;
  uint32_t launches = 0;

  RT_CUDA(w.d_counters.reserve(sizeof(Counters)));
  RT_CUDA(w.d_cursors.reserve(rtb::MAX_WORK_RANGES * sizeof(unsigned int)));
  RT_CUDA(w.h_counters.reserve(sizeof(Counters)));
  if (!on_device)
    RT_CUDA(w.d_queries.reserve(point_bytes ? point_bytes : 16));
  if (reorder)
  {
    RT_CUDA(w.d_order.reserve((size_t)n * sizeof(uint32_t)));
    RT_CUDA(w.d_scratch.reserve(rtb::reorder_scratch_bytes(n)));
  }
  RT_CUDA(t->d_knn_ids.reserve((out_elems + 1) * sizeof(uint32_t)));
  RT_CUDA(t->d_knn_d2.reserve((out_elems + 1) * sizeof(float)));

  RT_CUDA(cudaEventRecord(ev[EV_BEGIN], st));
  const float* d_points = points;
  if (!on_device)
  {
    if (point_bytes)
      RT_CUDA(cudaMemcpyAsync(w.d_queries.p, points, point_bytes, cudaMemcpyHostToDevice, st));
    d_points = w.d_queries.as<float>();
  }
  RT_CUDA(cudaMemsetAsync(w.d_counters.p, 0, sizeof(Counters), st));
  RT_CUDA(cudaEventRecord(ev[EV_H2D], st));
  const uint32_t* d_order = nullptr;
  if (reorder)
  {
    // the coherence pass of the box queries, with the point as its own centre
    RT_CUDA(rtb::launch_reorder(d_points, RT_F32, dim, dim, n, w.d_order.as<uint32_t>(), w.d_scratch.p, st,
                                &launches, 0u));
    d_order = w.d_order.as<uint32_t>();
  }
  RT_CUDA(cudaEventRecord(ev[EV_REORDER], st));

  Counters* dc = w.d_counters.as<Counters>();
  rtb::KnnLaunch L;
  std::memset(&L, 0, sizeof L);
  L.tree = t->dev;
  L.points = d_points;
  L.order = d_order;
  L.n = n;
  L.k = k;
  L.collect = collect ? 1u : 0u;
  L.ids_out = t->d_knn_ids.as<uint32_t>();
  L.d2_out = t->d_knn_d2.as<float>();
  L.work.cursors = w.d_cursors.as<unsigned int>();
  L.work.chunk = t->query_chunk;
  L.work.ranges = t->work_ranges;
  L.visit_stats = dc->visits;
  L.grid = t->traverse_grid;
  L.stream = st;
  if (n > 0)
  {
    RT_CUDA(rtb::launch_knn_kernel(L, dim, t->desc.width, t->desc.key_kind));
    ++launches;
  }
  RT_CUDA(cudaEventRecord(ev[EV_TRAVERSE], st));
  RT_CUDA(cudaMemcpyAsync(w.h_counters.p, dc, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  RT_CUDA(cudaEventRecord(ev[EV_FINISH], st));

  out->n_queries = n;
  out->k = k;
  if (keep_on_device)
  {
    out->ids = t->d_knn_ids.as<uint32_t>();
    out->dist2 = t->d_knn_d2.as<float>();
    out->memory = RT_MEM_DEVICE;
  }
  else
  {
    RT_CUDA(t->h_knn_ids.reserve((out_elems + 1) * sizeof(uint32_t)));
    RT_CUDA(t->h_knn_d2.reserve((out_elems + 1) * sizeof(float)));
    if (out_elems)
    {
      RT_CUDA(cudaMemcpyAsync(t->h_knn_ids.p, t->d_knn_ids.p, out_elems * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      RT_CUDA(cudaMemcpyAsync(t->h_knn_d2.p, t->d_knn_d2.p, out_elems * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    out->ids = t->h_knn_ids.as<uint32_t>();
    out->dist2 = t->h_knn_d2.as<float>();
    out->memory = RT_MEM_HOST;
  }
  RT_CUDA(cudaEventRecord(ev[EV_D2H], st));
  RT_CUDA(cudaStreamSynchronize(st));

  const Counters hc = *w.h_counters.as<Counters>();
  rt_stats& s = t->stats;
  s.total_ms = elapsed(ev[EV_BEGIN], ev[EV_D2H]);
  s.batches = 1;
  s.h2d_ms = elapsed(ev[EV_BEGIN], ev[EV_H2D]);
  s.reorder_ms = elapsed(ev[EV_H2D], ev[EV_REORDER]);
  s.traverse_ms = elapsed(ev[EV_REORDER], ev[EV_TRAVERSE]);
  s.finish_ms = elapsed(ev[EV_TRAVERSE], ev[EV_FINISH]);
  s.d2h_ms = elapsed(ev[EV_FINISH], ev[EV_D2H]);
  s.kernel_launches = launches;
  if (collect)
  {
    s.visits_internal = hc.visits[0];
    s.visits_leaf = hc.visits[1];
    const uint64_t M = t->desc.max_entries, D = dim;
    const uint64_t b_int = M * (2 * D * 4 + 4);
    const uint64_t b_leaf = t->desc.key_kind == RT_KEY_BOX ? b_int : M * (D * 4 + 4);
    s.algorithmic_bytes = hc.visits[0] * b_int + hc.visits[1] * b_leaf + n * (D * 4 + 8ull * k);
  }
  return RT_OK;
}

int run_refit_call(rt_tree* t, const void* keys, uint64_t n_keys, uint32_t flags)
{
  if (t == nullptr)
    return fail(RT_E_INVALID, "rt_refit: tree is NULL");
  if (n_keys != t->desc.num_values)
    return fail(RT_E_INVALID, "rt_refit: n_keys must equal the tree's num_values");
  if (keys == nullptr && n_keys > 0)
    return fail(RT_E_INVALID, "rt_refit: keys is NULL");
  int rc = use_device(t);
  if (rc != RT_OK)
    return rc;
  std::memset(&t->stats, 0, sizeof t->stats);
  cudaStream_t st = t->stream;
  Slot& w = t->slot[0];
  cudaEvent_t* ev = nullptr;
  if ((rc = batch_events(t, 0, &ev)) != RT_OK)
    return rc;
  const bool on_device = (flags & RT_QUERIES_ON_DEVICE) != 0;
  const size_t key_bytes = (size_t)n_keys * t->desc.dim * (t->desc.key_kind == RT_KEY_BOX ? 2 : 1)
                           * (t->desc.scalar == RT_F64 ? 8u : 4u);
  const uint32_t n_leaves = t->desc.num_nodes - t->levels[t->desc.leaf_level];
  uint32_t launches = 0;
  RT_CUDA(w.d_scratch.reserve(rtb::refit_scratch_bytes(n_leaves)));
  if (!on_device)
    RT_CUDA(w.d_queries.reserve(key_bytes ? key_bytes : 16));
  RT_CUDA(cudaEventRecord(ev[EV_BEGIN], st));
  const void* d_keys = keys;
  if (!on_device)
  {
    if (key_bytes)
      RT_CUDA(cudaMemcpyAsync(w.d_queries.p, keys, key_bytes, cudaMemcpyHostToDevice, st));
    d_keys = w.d_queries.p;
  }
  RT_CUDA(cudaEventRecord(ev[EV_H2D], st));
  RT_CUDA(rtb::launch_refit(t->d_blocks, t->desc, t->levels.data(), d_keys, w.d_scratch.p, st, &launches));
  if (t->d_qblocks)
  {
    // the quantised image follows the float blocks (the grid is re-derived from the new root bounds)
    RT_CUDA(rtb::build_quant_image(t->dev, t->desc, t->levels[t->desc.leaf_level], t->d_qblocks, st));
    launches += 2;
  }
  RT_CUDA(cudaEventRecord(ev[EV_TRAVERSE], st));
  RT_CUDA(cudaStreamSynchronize(st));
  t->stats.total_ms = elapsed(ev[EV_BEGIN], ev[EV_TRAVERSE]);
  t->stats.h2d_ms = elapsed(ev[EV_BEGIN], ev[EV_H2D]);
  t->stats.traverse_ms = elapsed(ev[EV_H2D], ev[EV_TRAVERSE]);
  t->stats.kernel_launches = launches;
  t->stats.batches = 1;
  return RT_OK;
}

} // namespace
