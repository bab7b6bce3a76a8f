// rt_api_rays.cu -- rt_query_rays: buffers, batches and copies around the ray traversal kernels.
//
// closest-hit / any-hit: one kernel per batch.  all-hits: count traversal -> scan -> [64-byte D2H:
// total] -> write traversal -> sort of the long lists (the CSR contract of rt_query_boxes).
// A call whose result stays on the device, or a small one, is one batch on the handle's stream.  A
// large call with a HOST result is cut into batches of about 2 x RT_OPT_PIPELINE_BATCH rays that
// rotate through the NSLOT working sets on their own streams, so that the H2D copy of the next batch
// and the D2H copy of the previous one run under the traversal of the current one (at 28 bytes per
// ray the call is otherwise bound by the H2D copy alone).
#include "rt_handle.cuh"

using namespace rtb;

namespace
{
struct RayCall
{
  const float* rays;
  uint64_t n;
  uint32_t mode;
  bool on_device, keep_on_device, count_only, collect, sorted;
};

// what a visit-counting ray call reports (a ray is 28 bytes in; 8 out for closest-hit,
// 1 for any-hit, offset + ids for all-hits)
void ray_visit_stats(rt_tree* t, const RayCall& c, const Counters& hc, uint64_t total)
{
  rt_stats& s = t->stats;
  s.visits_internal += hc.visits[0];
  s.visits_leaf += hc.visits[1];
  s.traverse_steps += hc.visits[3];
  const uint64_t M = t->desc.max_entries;
  const uint64_t b_int = M * (2 * 3 * 4 + 4);
  const uint64_t b_leaf = t->desc.key_kind == RT_KEY_BOX ? b_int : M * (3 * 4 + 4);
  const uint64_t out_bytes = c.mode == RT_RAY_ALL_HITS ? 8 : (c.mode == RT_RAY_CLOSEST ? 8 : 1);
  s.algorithmic_bytes = s.visits_internal * b_int + s.visits_leaf * b_leaf + c.n * (28 + out_bytes) + 4 * total;
}

void fill_launch(rt_tree* t, Slot& s, const RayCall& c, const float* d_rays, uint64_t m, uint64_t first,
                 cudaStream_t st, bool pipelined, rtb::RayLaunch& L)
{
  Counters* dc = s.d_counters.as<Counters>();
  std::memset(&L, 0, sizeof L);
  L.tree = t->dev;
  L.width = t->desc.width;
  L.key_kind = t->desc.key_kind;
  L.rays = d_rays;
  L.order = nullptr;
  L.n = m;
  L.mode = c.mode;
  L.pass = c.collect ? rtb::PASS_COUNT_STATS : rtb::PASS_COUNT;
  L.counts = s.d_counts.as<uint32_t>();
  L.offsets = s.d_offsets.as<unsigned long long>();
  L.big_list = s.d_big.as<uint32_t>();
  L.big_count = &dc->big;
  L.max_list = &dc->max_list;
  L.t_out = t->d_ray_t.as<float>() + first;
  L.id_out = t->d_ray_id.as<uint32_t>() + first;
  L.any_out = t->d_ray_any.as<uint8_t>() + first;
  L.work.cursors = s.d_cursors.as<unsigned int>();
  L.work.chunk = t->query_chunk;
  L.work.ranges = t->work_ranges;
  L.grid = t->traverse_grid ? t->traverse_grid : (pipelined ? -1 : 0);
  L.visit_stats = dc->visits;
  L.sorted = c.sorted ? 1u : 0u;
  L.stream = st;
}

// ---- closest / any, host result, many rays: one kernel per batch --------------------------------
int run_rays_pipelined(rt_tree* t, const RayCall& c, uint64_t per_batch, rt_ray_result* out)
{
  int rc = RT_OK;
  const uint64_t n = c.n;
  cudaStream_t st = t->stream;
  const uint64_t n_batches = (n + per_batch - 1) / per_batch;
  uint32_t launches = 0;
  if (c.mode == RT_RAY_CLOSEST)
  {
    RT_CUDA(t->d_ray_t.reserve((size_t)(n + 1) * sizeof(float)));
    RT_CUDA(t->d_ray_id.reserve((size_t)(n + 1) * sizeof(uint32_t)));
    RT_CUDA(t->h_ray_t.reserve((size_t)(n + 1) * sizeof(float)));
    RT_CUDA(t->h_ray_id.reserve((size_t)(n + 1) * sizeof(uint32_t)));
  }
  else
  {
    RT_CUDA(t->d_ray_any.reserve((size_t)n + 16));
    RT_CUDA(t->h_ray_any.reserve((size_t)n + 16));
  }
  RT_CUDA(t->h_batch_counters.reserve((size_t)n_batches * sizeof(Counters)));
  cudaEvent_t* unused = nullptr;
  if ((rc = batch_events(t, (size_t)n_batches - 1, &unused)) != RT_OK)
    return rc;
  RT_CUDA(cudaEventRecord(t->ev_fork, st));
  for (Slot& s : t->slot)
    RT_CUDA(cudaStreamWaitEvent(s.stream, t->ev_fork, 0));
  for (uint64_t k = 0; k < n_batches; ++k)
  {
    Slot& s = t->slot[k % NSLOT];
    cudaStream_t bs = s.stream;
    cudaEvent_t* bev = t->events.data() + (size_t)k * EV_COUNT;
    const uint64_t first = k * per_batch;
    const uint64_t m = (k + 1 == n_batches) ? n - first : per_batch;
    RT_CUDA(s.d_counters.reserve(sizeof(Counters)));
    RT_CUDA(s.d_cursors.reserve(rtb::MAX_WORK_RANGES * sizeof(unsigned int)));
    if (!c.on_device)
      RT_CUDA(s.d_queries.reserve((size_t)per_batch * 7 * sizeof(float)));
    RT_CUDA(cudaEventRecord(bev[EV_BEGIN], bs));
    const float* d_rays = c.rays + first * 7;
    if (!c.on_device)
    {
      RT_CUDA(cudaMemcpyAsync(s.d_queries.p, c.rays + first * 7, (size_t)m * 7 * sizeof(float), cudaMemcpyHostToDevice,
                              bs));
      d_rays = s.d_queries.as<float>();
    }
    RT_CUDA(cudaMemsetAsync(s.d_counters.p, 0, sizeof(Counters), bs));
    RT_CUDA(cudaEventRecord(bev[EV_H2D], bs));
    rtb::RayLaunch L;
    fill_launch(t, s, c, d_rays, m, first, bs, true, L);
    RT_CUDA(rtb::launch_ray_kernel(L));
    ++launches;
    RT_CUDA(cudaEventRecord(bev[EV_TRAVERSE], bs));
    RT_CUDA(rtb::launch_publish_counters(s.d_counters.p, sizeof(Counters), 0, nullptr,
                                         t->h_batch_counters.as<Counters>() + k, bs, &launches));
    if (c.mode == RT_RAY_CLOSEST)
    {
      RT_CUDA(cudaMemcpyAsync(t->h_ray_t.as<float>() + first, L.t_out, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost,
                              bs));
      RT_CUDA(cudaMemcpyAsync(t->h_ray_id.as<uint32_t>() + first, L.id_out, (size_t)m * sizeof(uint32_t),
                              cudaMemcpyDeviceToHost, bs));
    }
    else
    {
      RT_CUDA(cudaMemcpyAsync(t->h_ray_any.as<uint8_t>() + first, L.any_out, (size_t)m, cudaMemcpyDeviceToHost, bs));
    }
    RT_CUDA(cudaEventRecord(bev[EV_D2H], bs));
  }
  for (Slot& s : t->slot)
  {
    RT_CUDA(cudaEventRecord(s.done, s.stream));
    RT_CUDA(cudaStreamWaitEvent(st, s.done, 0));
  }
  RT_CUDA(cudaEventRecord(t->ev_join, st));
  RT_CUDA(cudaStreamSynchronize(st));

  out->n_rays = n;
  out->memory = RT_MEM_HOST;
  if (c.mode == RT_RAY_CLOSEST)
  {
    out->t = t->h_ray_t.as<float>();
    out->id = t->h_ray_id.as<uint32_t>();
  }
  else
  {
    out->any = t->h_ray_any.as<uint8_t>();
  }
  rt_stats& rs = t->stats;
  rs.total_ms = elapsed(t->ev_fork, t->ev_join);
  rs.batches = (uint32_t)n_batches;
  rs.kernel_launches = launches;
  Counters sum;
  std::memset(&sum, 0, sizeof sum);
  for (uint64_t k = 0; k < n_batches; ++k)
  {
    cudaEvent_t* bev = t->events.data() + (size_t)k * EV_COUNT;
    rs.h2d_ms += elapsed(bev[EV_BEGIN], bev[EV_H2D]);
    rs.traverse_ms += elapsed(bev[EV_H2D], bev[EV_TRAVERSE]);
    rs.d2h_ms += elapsed(bev[EV_TRAVERSE], bev[EV_D2H]);
    const Counters& hc = t->h_batch_counters.as<Counters>()[k];
    out->n_hit += hc.visits[2];
    sum.visits[0] += hc.visits[0], sum.visits[1] += hc.visits[1], sum.visits[3] += hc.visits[3];
  }
  if (c.collect)
    ray_visit_stats(t, c, sum, 0);
  return RT_OK;
}

// ---- all-hits, host result, many rays: batches in two phases, the host one phase behind ----------
struct RayBatch
{
  Slot* slot;
  cudaEvent_t* ev;
  uint64_t first, n;
  rtb::RayLaunch L;
  Counters hc;
  uint32_t launches;
};

int run_all_hits_pipelined(rt_tree* t, const RayCall& c, uint64_t per_batch, rt_ray_result* out)
{
  int rc = RT_OK;
  const uint64_t n = c.n;
  cudaStream_t st = t->stream;
  const uint64_t n_batches = (n + per_batch - 1) / per_batch;
  std::vector<RayBatch> batches((size_t)n_batches);
  cudaEvent_t* unused = nullptr;
  if ((rc = batch_events(t, (size_t)n_batches - 1, &unused)) != RT_OK)
    return rc;
  RT_CUDA(t->h_offsets.reserve((size_t)(n + 1) * sizeof(unsigned long long)));
  if (!c.count_only && t->ids_wanted)
    RT_CUDA(t->h_ids.reserve((size_t)(t->ids_wanted + t->ids_wanted / 32 + 1024) * sizeof(uint32_t)));
  RT_CUDA(cudaEventRecord(t->ev_fork, st));
  for (Slot& s : t->slot)
    RT_CUDA(cudaStreamWaitEvent(s.stream, t->ev_fork, 0));

  unsigned long long total = 0;
  uint64_t done_rays = 0;
  const uint64_t lag = NSLOT - 1;
  for (uint64_t k = 0; k < n_batches + lag; ++k)
  {
    if (k < n_batches)
    {
      // phase 1: H2D -> count traversal -> scan -> counters on their way to the host
      RayBatch& b = batches[(size_t)k];
      b.slot = &t->slot[k % NSLOT];
      b.ev = t->events.data() + (size_t)k * EV_COUNT;
      b.first = k * per_batch;
      b.n = (k + 1 == n_batches) ? n - b.first : per_batch;
      b.launches = 0;
      Slot& s = *b.slot;
      cudaStream_t bs = s.stream;
      RT_CUDA(s.d_counters.reserve(sizeof(Counters)));
      RT_CUDA(s.d_cursors.reserve(rtb::MAX_WORK_RANGES * sizeof(unsigned int)));
      RT_CUDA(s.h_counters.reserve(sizeof(Counters)));
      RT_CUDA(s.d_counts.reserve((size_t)(per_batch + 4) * sizeof(uint32_t)));
      RT_CUDA(s.d_offsets.reserve((size_t)(per_batch + 1) * sizeof(unsigned long long)));
      RT_CUDA(s.d_scratch.reserve(rtb::scan_scratch_bytes(per_batch)));
      RT_CUDA(s.d_big.reserve((size_t)(per_batch + 1) * sizeof(uint32_t)));
      if (!c.on_device)
        RT_CUDA(s.d_queries.reserve((size_t)per_batch * 7 * sizeof(float)));
      RT_CUDA(cudaEventRecord(b.ev[EV_BEGIN], bs));
      const float* d_rays = c.rays + b.first * 7;
      if (!c.on_device)
      {
        RT_CUDA(cudaMemcpyAsync(s.d_queries.p, d_rays, (size_t)b.n * 7 * sizeof(float), cudaMemcpyHostToDevice, bs));
        d_rays = s.d_queries.as<float>();
      }
      RT_CUDA(cudaMemsetAsync(s.d_counters.p, 0, sizeof(Counters), bs));
      RT_CUDA(cudaEventRecord(b.ev[EV_H2D], bs));
      fill_launch(t, s, c, d_rays, b.n, 0, bs, true, b.L);
      RT_CUDA(rtb::launch_ray_kernel(b.L));
      ++b.launches;
      RT_CUDA(cudaEventRecord(b.ev[EV_TRAVERSE], bs));
      Counters* dc = s.d_counters.as<Counters>();
      RT_CUDA(rtb::launch_exclusive_scan(s.d_counts.as<uint32_t>(), b.n, s.d_offsets.as<unsigned long long>(),
                                         s.d_scratch.p, bs, &b.launches));
      RT_CUDA(rtb::launch_publish_counters(dc, sizeof(Counters), offsetof(Counters, total),
                                           s.d_offsets.as<unsigned long long>() + b.n, s.h_counters.p, bs, &b.launches));
      RT_CUDA(cudaEventRecord(b.ev[EV_COUNTED], bs));
    }
    if (k < lag)
      continue;
    // phase 2 of the batch issued `lag` rounds ago: write traversal -> sort -> D2H
    RayBatch& b = batches[(size_t)(k - lag)];
    Slot& s = *b.slot;
    cudaStream_t bs = s.stream;
    RT_CUDA(cudaEventSynchronize(b.ev[EV_COUNTED]));
    b.hc = *s.h_counters.as<Counters>();
    const unsigned long long batch_total = b.hc.total;
    const bool last = k - lag + 1 == n_batches;
    RT_CUDA(cudaEventRecord(b.ev[EV_WRITE], bs));
    if (!c.count_only)
    {
      const unsigned long long need = total + batch_total;
      const double seen = (double)(done_rays + b.n);
      const unsigned long long hint = (unsigned long long)((double)need / (seen > 0 ? seen : 1.0) * (double)n * 1.03) + 4096;
      if ((rc = grow_host_ids(t, need ? need : 1, total, hint)) != RT_OK)
        return rc;
      RT_CUDA(s.d_ids.reserve((size_t)(batch_total ? batch_total : 1) * sizeof(uint32_t)));
      if (b.n > 0 && batch_total > 0)
      {
        b.L.pass = rtb::PASS_WRITE;
        b.L.ids = s.d_ids.as<uint32_t>();
        RT_CUDA(rtb::launch_ray_kernel(b.L));
        ++b.launches;
        if (c.sorted)
        {
          void* long_scratch = nullptr;
          if (b.hc.max_list > 8192)
          {
            RT_CUDA(s.d_long.reserve(rtb::long_sort_scratch_bytes(batch_total)));
            long_scratch = s.d_long.p;
          }
          RT_CUDA(rtb::launch_sort_segments(s.d_big.as<uint32_t>(), &s.d_counters.as<Counters>()->big, (uint32_t)b.n,
                                            s.d_offsets.as<unsigned long long>(), s.d_ids.as<uint32_t>(), bs,
                                            &b.launches, b.hc.max_list, long_scratch, batch_total));
        }
      }
    }
    RT_CUDA(cudaEventRecord(b.ev[EV_FINISH], bs));
    if (total)
      RT_CUDA(rtb::launch_add_base(s.d_offsets.as<unsigned long long>(), b.n + 1, total, bs, &b.launches));
    const size_t n_off = (size_t)b.n + (last ? 1 : 0);
    if (n_off)
      RT_CUDA(cudaMemcpyAsync(t->h_offsets.as<unsigned long long>() + b.first, s.d_offsets.p,
                              n_off * sizeof(unsigned long long), cudaMemcpyDeviceToHost, bs));
    if (!c.count_only && batch_total)
      RT_CUDA(cudaMemcpyAsync(t->h_ids.as<uint32_t>() + total, s.d_ids.p, (size_t)batch_total * sizeof(uint32_t),
                              cudaMemcpyDeviceToHost, bs));
    RT_CUDA(cudaEventRecord(b.ev[EV_D2H], bs));
    total += batch_total;
    done_rays += b.n;
  }
  for (Slot& s : t->slot)
  {
    RT_CUDA(cudaEventRecord(s.done, s.stream));
    RT_CUDA(cudaStreamWaitEvent(st, s.done, 0));
  }
  t->ids_wanted = total;
  RT_CUDA(cudaEventRecord(t->ev_join, st));
  RT_CUDA(cudaStreamSynchronize(st));

  out->n_rays = n;
  out->n_hit = total;
  out->memory = RT_MEM_HOST;
  out->hits.n_queries = n;
  out->hits.total = total;
  out->hits.memory = RT_MEM_HOST;
  out->hits.offsets = t->h_offsets.as<uint64_t>();
  out->hits.ids = c.count_only ? nullptr : t->h_ids.as<uint32_t>();
  rt_stats& rs = t->stats;
  rs.total_ms = elapsed(t->ev_fork, t->ev_join);
  rs.batches = (uint32_t)n_batches;
  Counters sum;
  std::memset(&sum, 0, sizeof sum);
  for (const RayBatch& b : batches)
  {
    rs.h2d_ms += elapsed(b.ev[EV_BEGIN], b.ev[EV_H2D]);
    rs.traverse_ms += elapsed(b.ev[EV_H2D], b.ev[EV_TRAVERSE]);
    rs.finish_ms += elapsed(b.ev[EV_TRAVERSE], b.ev[EV_COUNTED]) + elapsed(b.ev[EV_WRITE], b.ev[EV_FINISH]);
    rs.d2h_ms += elapsed(b.ev[EV_FINISH], b.ev[EV_D2H]);
    rs.kernel_launches += b.launches;
    sum.visits[0] += b.hc.visits[0], sum.visits[1] += b.hc.visits[1], sum.visits[3] += b.hc.visits[3];
  }
  if (c.collect)
    ray_visit_stats(t, c, sum, total);
  return RT_OK;
}

// ---- one batch on the caller's stream -------------------------------------------------------------
int run_rays_single(rt_tree* t, const RayCall& c, rt_ray_result* out)
{
  int rc = RT_OK;
  const uint64_t n = c.n;
  const uint32_t mode = c.mode;
  cudaStream_t st = t->stream;
  Slot& w = t->slot[0];
  cudaEvent_t* ev = nullptr;
  if ((rc = batch_events(t, 0, &ev)) != RT_OK)
    return rc;
  RT_CUDA(w.d_counters.reserve(sizeof(Counters)));
  RT_CUDA(w.d_cursors.reserve(rtb::MAX_WORK_RANGES * sizeof(unsigned int)));
  RT_CUDA(w.h_counters.reserve(sizeof(Counters)));
  const size_t ray_bytes = (size_t)n * 7 * sizeof(float);
  uint32_t launches = 0;
  if (!c.on_device)
    RT_CUDA(w.d_queries.reserve(ray_bytes ? ray_bytes : 16));
  if (mode == RT_RAY_ALL_HITS)
  {
    RT_CUDA(w.d_counts.reserve((size_t)(n + 4) * sizeof(uint32_t)));
    RT_CUDA(w.d_offsets.reserve((size_t)(n + 1) * sizeof(unsigned long long)));
    RT_CUDA(w.d_scratch.reserve(rtb::scan_scratch_bytes(n)));
    RT_CUDA(w.d_big.reserve((size_t)(n + 1) * sizeof(uint32_t)));
  }
  else if (mode == RT_RAY_CLOSEST)
  {
    RT_CUDA(t->d_ray_t.reserve((size_t)(n + 1) * sizeof(float)));
    RT_CUDA(t->d_ray_id.reserve((size_t)(n + 1) * sizeof(uint32_t)));
  }
  else
  {
    RT_CUDA(t->d_ray_any.reserve((size_t)n + 16));
  }

  RT_CUDA(cudaEventRecord(ev[EV_BEGIN], st));
  const float* d_rays = c.rays;
  if (!c.on_device)
  {
    if (ray_bytes)
      RT_CUDA(cudaMemcpyAsync(w.d_queries.p, c.rays, ray_bytes, cudaMemcpyHostToDevice, st));
    d_rays = w.d_queries.as<float>();
  }
  RT_CUDA(cudaMemsetAsync(w.d_counters.p, 0, sizeof(Counters), st));
  RT_CUDA(cudaEventRecord(ev[EV_H2D], st));
  RT_CUDA(cudaEventRecord(ev[EV_REORDER], st));

  Counters* dc = w.d_counters.as<Counters>();
  rtb::RayLaunch L;
  fill_launch(t, w, c, d_rays, n, 0, st, false, L);
  if (n > 0)
  {
    RT_CUDA(rtb::launch_ray_kernel(L));
    ++launches;
  }
  RT_CUDA(cudaEventRecord(ev[EV_TRAVERSE], st));

  Counters hc;
  std::memset(&hc, 0, sizeof hc);
  unsigned long long total = 0;
  if (mode == RT_RAY_ALL_HITS)
  {
    RT_CUDA(rtb::launch_exclusive_scan(w.d_counts.as<uint32_t>(), n, w.d_offsets.as<unsigned long long>(),
                                       w.d_scratch.p, st, &launches));
    RT_CUDA(cudaMemcpyAsync(&dc->total, w.d_offsets.as<unsigned long long>() + n, sizeof(unsigned long long),
                            cudaMemcpyDeviceToDevice, st));
  }
  RT_CUDA(cudaMemcpyAsync(w.h_counters.p, dc, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  RT_CUDA(cudaStreamSynchronize(st));
  hc = *w.h_counters.as<Counters>();
  total = hc.total;

  if (mode == RT_RAY_ALL_HITS && !c.count_only)
  {
    RT_CUDA(w.d_ids.reserve((size_t)(total ? total : 1) * sizeof(uint32_t)));
    if (n > 0 && total > 0)
    {
      L.pass = rtb::PASS_WRITE;
      L.ids = w.d_ids.as<uint32_t>();
      RT_CUDA(rtb::launch_ray_kernel(L));
      ++launches;
      if (c.sorted)
      {
        void* long_scratch = nullptr;
        if (hc.max_list > 8192)
        {
          RT_CUDA(w.d_long.reserve(rtb::long_sort_scratch_bytes(total)));
          long_scratch = w.d_long.p;
        }
        RT_CUDA(rtb::launch_sort_segments(w.d_big.as<uint32_t>(), &dc->big, (uint32_t)n,
                                          w.d_offsets.as<unsigned long long>(), w.d_ids.as<uint32_t>(), st,
                                          &launches, hc.max_list, long_scratch, total));
      }
    }
  }
  RT_CUDA(cudaEventRecord(ev[EV_FINISH], st));

  out->n_rays = n;
  out->memory = c.keep_on_device ? RT_MEM_DEVICE : RT_MEM_HOST;
  if (mode == RT_RAY_ALL_HITS)
  {
    out->n_hit = total;
    out->hits.n_queries = n;
    out->hits.total = total;
    out->hits.memory = out->memory;
    if (c.keep_on_device)
    {
      out->hits.offsets = w.d_offsets.as<uint64_t>();
      out->hits.counts = w.d_counts.as<uint32_t>();
      out->hits.ids = c.count_only ? nullptr : w.d_ids.as<uint32_t>();
    }
    else
    {
      RT_CUDA(t->h_offsets.reserve((size_t)(n + 1) * sizeof(unsigned long long)));
      RT_CUDA(cudaMemcpyAsync(t->h_offsets.p, w.d_offsets.p, (size_t)(n + 1) * sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, st));
      if (!c.count_only)
      {
        RT_CUDA(t->h_ids.reserve((size_t)(total ? total : 1) * sizeof(uint32_t)));
        if (total)
          RT_CUDA(cudaMemcpyAsync(t->h_ids.p, w.d_ids.p, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                  st));
      }
      out->hits.offsets = t->h_offsets.as<uint64_t>();
      out->hits.ids = c.count_only ? nullptr : t->h_ids.as<uint32_t>();
    }
  }
  else if (mode == RT_RAY_CLOSEST)
  {
    out->n_hit = hc.visits[2];
    if (c.keep_on_device)
    {
      out->t = t->d_ray_t.as<float>();
      out->id = t->d_ray_id.as<uint32_t>();
    }
    else
    {
      RT_CUDA(t->h_ray_t.reserve((size_t)(n + 1) * sizeof(float)));
      RT_CUDA(t->h_ray_id.reserve((size_t)(n + 1) * sizeof(uint32_t)));
      if (n)
      {
        RT_CUDA(cudaMemcpyAsync(t->h_ray_t.p, t->d_ray_t.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
        RT_CUDA(cudaMemcpyAsync(t->h_ray_id.p, t->d_ray_id.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      }
      out->t = t->h_ray_t.as<float>();
      out->id = t->h_ray_id.as<uint32_t>();
    }
  }
  else
  {
    out->n_hit = hc.visits[2];
    if (c.keep_on_device)
    {
      out->any = t->d_ray_any.as<uint8_t>();
    }
    else
    {
      RT_CUDA(t->h_ray_any.reserve((size_t)n + 16));
      if (n)
        RT_CUDA(cudaMemcpyAsync(t->h_ray_any.p, t->d_ray_any.p, (size_t)n, cudaMemcpyDeviceToHost, st));
      out->any = t->h_ray_any.as<uint8_t>();
    }
  }
  RT_CUDA(cudaEventRecord(ev[EV_D2H], st));
  RT_CUDA(cudaStreamSynchronize(st));

  rt_stats& s = t->stats;
  s.total_ms = elapsed(ev[EV_BEGIN], ev[EV_D2H]);
  s.batches = 1;
  s.h2d_ms = elapsed(ev[EV_BEGIN], ev[EV_H2D]);
  s.traverse_ms = elapsed(ev[EV_REORDER], ev[EV_TRAVERSE]);
  s.finish_ms = elapsed(ev[EV_TRAVERSE], ev[EV_FINISH]);
  s.d2h_ms = elapsed(ev[EV_FINISH], ev[EV_D2H]);
  s.kernel_launches = launches;
  if (c.collect)
    ray_visit_stats(t, c, hc, mode == RT_RAY_ALL_HITS ? total : 0);
  return RT_OK;
}
} // namespace

extern "C"
{

int rt_query_rays(rt_tree* t, const float* rays, uint64_t n, uint32_t mode, uint32_t flags, rt_ray_result* out)
{
  if (t == nullptr || out == nullptr)
    return fail(RT_E_INVALID, "rt_query_rays: NULL argument");
  if (rays == nullptr && n > 0)
    return fail(RT_E_INVALID, "rt_query_rays: rays is NULL");
  if (mode > RT_RAY_ANY)
    return fail(RT_E_INVALID, "rt_query_rays: unknown mode");
  if (t->desc.dim != 3 || t->desc.scalar != RT_F32)
    return fail(RT_E_UNSUPPORTED, "rt_query_rays: ray queries need a 3-D float32 tree");
  if (n >= 0xFFFFFFFFull)
    return fail(RT_E_INVALID, "rt_query_rays: at most 2^32-2 rays per call");
  int rc = use_device(t);
  if (rc != RT_OK)
    return rc;

  std::memset(out, 0, sizeof *out);
  std::memset(&t->stats, 0, sizeof t->stats);
  RayCall c;
  c.rays = rays;
  c.n = n;
  c.mode = mode;
  c.on_device = (flags & RT_QUERIES_ON_DEVICE) != 0;
  c.keep_on_device = (flags & RT_RESULT_ON_DEVICE) != 0;
  c.count_only = (flags & RT_COUNT_ONLY) != 0;
  c.collect = (flags & RT_COLLECT_VISITS) != 0;
  c.sorted = (flags & RT_UNSORTED) == 0;

  // a host result and at least two batches of about 2 x RT_OPT_PIPELINE_BATCH rays: pipelined
  uint64_t B = t->pipeline_batch;
  if (B != 0 && B < MIN_PIPELINE_BATCH)
    B = MIN_PIPELINE_BATCH;
  while (B != 0 && n / (2 * B) > MAX_BATCHES)
    B *= 2;
  if (!c.keep_on_device && B > 0 && n / (2 * B) >= 2)
  {
    const uint64_t wanted = n / (2 * B);
    const uint64_t per_batch = ((n + wanted - 1) / wanted + 15) / 16 * 16; // keeps every copy 16-byte aligned
    rc = mode == RT_RAY_ALL_HITS ? run_all_hits_pipelined(t, c, per_batch, out) : run_rays_pipelined(t, c, per_batch, out);
  }
  else
  {
    rc = run_rays_single(t, c, out);
  }
  if (rc != RT_OK)
  {
    std::memset(out, 0, sizeof *out);
    return abandon_call(t, rc);
  }
  return RT_OK;
}

} // extern "C"
