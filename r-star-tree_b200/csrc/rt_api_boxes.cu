// rt_api_boxes.cu -- rt_query_boxes: how a call is cut into batches and the two-phase pipeline of a
// batch.  No query-path arithmetic lives here.
//
// Pipeline (default flags), per batch of queries:
//   H2D queries -> reorder (Morton cells; quantised trees: the queries themselves are re-laid in that
//   order, on the tree's grid) -> traversal PASS_STASH (counts + sorted short lists into the stash)
//   -> scan (CSR offsets) -> [64-byte D2H: total, #deferred] -> gather stash -> PASS_WRITE over the
//   deferred queries only -> sort their long lists -> D2H offsets (or counts) + ids.
// With RT_COLLECT_VISITS / RT_COUNT_ONLY the plain two-pass form (PASS_COUNT[_STATS], scan,
// PASS_WRITE over all queries) runs instead.
// A call whose result stays on the device is ONE batch on the handle's stream.  A large call with a
// host result is cut into batches that rotate through NSLOT working sets on their own streams
// (forked from / joined to the handle's stream), so the PCIe copies of neighbouring batches run
// under the traversal of the current one.
#include "rt_handle.cuh"

using namespace rtb;

namespace
{
uint64_t algorithmic_bytes_boxes(const rt_tree* t, uint64_t n, uint64_t total, uint64_t v_int, uint64_t v_leaf)
{
  // V_int*B_int + V_leaf*B_leaf + query bytes + 8 (CSR offset) + 4*hits
  const uint64_t s = t->desc.scalar == RT_F64 ? 8 : 4;
  const uint64_t M = t->desc.max_entries, D = t->desc.dim;
  const uint64_t b_int = M * (2 * D * s + 4);
  const uint64_t b_leaf = t->desc.key_kind == RT_KEY_BOX ? b_int : M * (D * s + 4);
  return v_int * b_int + v_leaf * b_leaf + n * (2 * D * s + 8) + 4 * total;
}

// ---- one rt_query_boxes call = one or more batches, each issued in two phases ----------------
struct BoxCall
{
  rtb::box_launch_fn launch;
  size_t query_stride; // bytes per query
  bool on_device, keep_on_device, count_only, collect, sorted, reorder, two_pass, pipelined, any_hit, counts32, strict;
};

struct Batch
{
  Slot* slot = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t* ev = nullptr;
  const void* queries = nullptr; // this batch's queries, host or device memory
  uint64_t first = 0, n = 0;
  rtb::BoxLaunch L;
  const uint32_t* d_order = nullptr;
  Counters hc;
  uint32_t launches = 0, deferred = 0;
};

// phase 1: H2D -> reorder -> first traversal -> scan -> counters on their way to the host
int box_issue_count(rt_tree* t, const BoxCall& c, Batch& b)
{
  Slot& s = *b.slot;
  const uint64_t n = b.n;
  cudaStream_t st = b.st;
  const bool reorder = c.reorder && n >= 4096;

  RT_CUDA(s.d_counters.reserve(sizeof(Counters)));
  RT_CUDA(s.d_cursors.reserve(rtb::MAX_WORK_RANGES * sizeof(unsigned int)));
  RT_CUDA(s.h_counters.reserve(sizeof(Counters)));
  RT_CUDA(s.d_counts.reserve((size_t)(n + 4) * sizeof(uint32_t)));
  RT_CUDA(s.d_offsets.reserve((size_t)(n + 1) * sizeof(unsigned long long)));
  size_t scratch = rtb::scan_scratch_bytes(n);
  if (reorder)
    scratch = scratch > rtb::reorder_scratch_bytes(n) ? scratch : rtb::reorder_scratch_bytes(n);
  RT_CUDA(s.d_scratch.reserve(scratch));
  RT_CUDA(s.d_big.reserve((size_t)(n + 1) * sizeof(uint32_t)));
  const size_t query_bytes = (size_t)n * c.query_stride;
  if (!c.on_device)
    RT_CUDA(s.d_queries.reserve(query_bytes ? query_bytes : 16));
  // quantised trees of dim <= 3 take their queries packed: in processing order, on the tree's grid
  const bool packed = t->use_quant && !c.strict && t->dev.qblocks != nullptr && t->desc.dim <= 3;
  if (packed)
    RT_CUDA(s.d_packed.reserve((size_t)(n ? n : 1) * sizeof(uint4)));
  else if (reorder)
    RT_CUDA(s.d_order.reserve((size_t)n * sizeof(uint32_t)));
  unsigned long long stash_capacity = 0;
  if (!c.two_pass)
  {
    RT_CUDA(s.d_stash_pos.reserve((size_t)(n + 1) * sizeof(uint32_t)));
    RT_CUDA(s.d_deferred.reserve((size_t)(n + 1) * sizeof(uint32_t)));
    // pool: what the previous batch needed (+25%), else ~12 ids per query; never more than a
    // quarter of the free memory, never more than 32-bit positions can address
    // (a batch whose long lists were deferred used little of the pool: also allow for the hits seen,
    // scaled to this batch, plus the quarter slab a warp may leave unused)
    unsigned long long want = s.stash_wanted ? s.stash_wanted + s.stash_wanted / 4 : n * 12ull;
    if (s.last_n)
    {
      const unsigned long long by_hits = (unsigned long long)((double)s.last_total / (double)s.last_n * (double)n * 1.5);
      want = want > by_hits ? want : by_hits;
    }
    want += 4ull * s.stash_slab * 1024;
    if (want > 0xFFFFF000ull)
      want = 0xFFFFF000ull;
    const unsigned long long have = s.d_stash.cap / sizeof(uint32_t);
    if (want > have)
    {
      // growing: cap the pool at a quarter of the free memory (cudaMemGetInfo costs
      // milliseconds, so it is only asked when the pool really has to grow)
      size_t free_b = 0, total_b = 0;
      RT_CUDA(cudaMemGetInfo(&free_b, &total_b));
      const unsigned long long limit = (unsigned long long)(free_b / 4 / sizeof(uint32_t)) + have;
      if (want > limit)
        want = limit;
    }
    if (want > have)
    {
      cudaError_t e = s.d_stash.reserve((size_t)want * sizeof(uint32_t));
      if (e != cudaSuccess)
        cudaGetLastError(); // keep whatever we have; overflowing queries are deferred
    }
    stash_capacity = s.d_stash.cap / sizeof(uint32_t);
    if (stash_capacity > 0xFFFFF000ull)
      stash_capacity = 0xFFFFF000ull;
    // test knob: pretend the pool is this small so the exhaustion path can be exercised
    if (const char* limit = std::getenv("RTB_STASH_LIMIT_IDS"))
    {
      const unsigned long long cap = std::strtoull(limit, nullptr, 10);
      if (cap < stash_capacity)
        stash_capacity = cap;
    }
  }

  // ---- H2D ---------------------------------------------------------------------------
  RT_CUDA(cudaEventRecord(b.ev[EV_BEGIN], st));
  const void* d_q = b.queries;
  if (!c.on_device)
  {
    if (query_bytes)
      RT_CUDA(cudaMemcpyAsync(s.d_queries.p, b.queries, query_bytes, cudaMemcpyHostToDevice, st));
    d_q = s.d_queries.p;
  }
  RT_CUDA(cudaMemsetAsync(s.d_counters.p, 0, sizeof(Counters), st));
  RT_CUDA(cudaEventRecord(b.ev[EV_H2D], st));

  // ---- coherence pass ----------------------------------------------------------------
  b.d_order = nullptr;
  if (packed)
  {
    if (reorder)
      RT_CUDA(rtb::launch_reorder(d_q, t->desc.scalar, t->desc.dim, 2 * t->desc.dim, n, nullptr, s.d_scratch.p, st,
                                  &b.launches, 0xFFFFFFFFu, s.d_packed.as<uint4>(), &t->dev));
    else
      RT_CUDA(rtb::launch_pack_queries(d_q, t->desc.dim, n, s.d_packed.as<uint4>(), t->dev, st, &b.launches));
  }
  else if (reorder)
  {
    RT_CUDA(rtb::launch_reorder(d_q, t->desc.scalar, t->desc.dim, 2 * t->desc.dim, n, s.d_order.as<uint32_t>(),
                                s.d_scratch.p, st, &b.launches));
    b.d_order = s.d_order.as<uint32_t>();
  }
  RT_CUDA(cudaEventRecord(b.ev[EV_REORDER], st));

  // ---- first traversal -----------------------------------------------------------------
  Counters* dc = s.d_counters.as<Counters>();
  rtb::BoxLaunch& L = b.L;
  std::memset(&L, 0, sizeof L);
  L.tree = t->dev;
  if (!t->use_quant || c.strict) // the strict filter is not conservative for the quantised grid: float blocks
    L.tree.qblocks = nullptr;
  L.strict = c.strict ? 1u : 0u;
  L.queries = d_q;
  L.packed = packed ? s.d_packed.as<uint4>() : nullptr;
  L.order = b.d_order;
  L.n = n;
  L.counts = s.d_counts.as<uint32_t>();
  L.offsets = s.d_offsets.as<unsigned long long>();
  L.stash = s.d_stash.as<uint32_t>();
  L.stash_pos = s.d_stash_pos.as<uint32_t>();
  L.stash_capacity = stash_capacity;
  L.stash_slab = s.stash_slab;
  L.stash_cursor = &dc->stash_cursor;
  L.deferred_list = s.d_deferred.as<uint32_t>();
  L.deferred_count = &dc->deferred;
  L.big_list = s.d_big.as<uint32_t>();
  L.big_count = &dc->big;
  L.max_list = &dc->max_list;
  L.work.cursors = s.d_cursors.as<unsigned int>();
  L.work.chunk = t->query_chunk;
  L.work.ranges = t->work_ranges;
  L.grid = t->traverse_grid ? t->traverse_grid : (c.pipelined ? -1 : 0);
  L.visit_stats = dc->visits;
  L.confirm_stats = &dc->lookups;
  L.sorted = c.sorted ? 1u : 0u;
  L.stop_at_first = c.any_hit ? 1u : 0u;
  L.stream = st;
  L.pass = c.two_pass ? (c.collect ? rtb::PASS_COUNT_STATS : rtb::PASS_COUNT) : rtb::PASS_STASH;
  if (n > 0)
  {
    RT_CUDA(c.launch(L));
    ++b.launches;
  }
  RT_CUDA(cudaEventRecord(b.ev[EV_TRAVERSE], st));

  // ---- CSR offsets, then the 64-byte read-back that sizes the id buffer -----------------
  RT_CUDA(rtb::launch_exclusive_scan(s.d_counts.as<uint32_t>(), n, s.d_offsets.as<unsigned long long>(),
                                     s.d_scratch.p, st, &b.launches));
  RT_CUDA(rtb::launch_publish_counters(dc, sizeof(Counters), offsetof(Counters, total),
                                       s.d_offsets.as<unsigned long long>() + n, s.h_counters.p, st, &b.launches));
  RT_CUDA(cudaEventRecord(b.ev[EV_COUNTED], st));
  return RT_OK;
}

// between the phases: wait for the batch's counters
int box_wait_counted(Batch& b)
{
  RT_CUDA(cudaEventSynchronize(b.ev[EV_COUNTED]));
  b.hc = *b.slot->h_counters.as<Counters>();
  return RT_OK;
}

// phase 2: gather the stash -> second traversal of the deferred queries -> sort long lists ->
// D2H.  `base` = hits of the batches before this one; a host result goes to
// h_offsets[first ..] (rebased) and h_ids[base ..].
int box_issue_write(rt_tree* t, const BoxCall& c, Batch& b, unsigned long long base, bool last_batch)
{
  Slot& s = *b.slot;
  const uint64_t n = b.n;
  cudaStream_t st = b.st;
  Counters* dc = s.d_counters.as<Counters>();
  const unsigned long long total = b.hc.total;
  if (!c.two_pass)
  {
    s.stash_wanted = b.hc.stash_cursor;
    // slab size for the next batch on this slot: room for lists 16 times the average, so that the
    // long ones spill into the stash instead of costing a second traversal
    const unsigned long long average = n ? total / n : 0;
    uint32_t slab = rtb::STASH_SLAB;
    while (slab < rtb::STASH_SLAB_MAX && slab < 16 * average)
      slab *= 2;
    s.stash_slab = slab;
    s.last_total = total;
    s.last_n = n;
  }
  rtb::BoxLaunch& L = b.L;

  RT_CUDA(cudaEventRecord(b.ev[EV_WRITE], st));
  if (!c.count_only)
  {
    RT_CUDA(s.d_ids.reserve((size_t)(total ? total : 1) * sizeof(uint32_t)));
    bool second = false;
    if (c.two_pass)
    {
      second = n > 0 && total > 0;
      L.order = b.d_order;
      L.n = n;
      b.deferred = (uint32_t)n;
    }
    else
    {
      RT_CUDA(rtb::launch_gather_stash(s.d_stash.as<uint32_t>(), s.d_stash_pos.as<uint32_t>(),
                                       s.d_offsets.as<unsigned long long>(), n, s.d_ids.as<uint32_t>(), st,
                                       &b.launches));
      b.deferred = b.hc.deferred;
      second = b.deferred > 0;
      L.order = s.d_deferred.as<uint32_t>();
      L.n = b.deferred;
    }
    if (second)
    {
      L.pass = rtb::PASS_WRITE;
      L.ids = s.d_ids.as<uint32_t>();
      RT_CUDA(c.launch(L));
      ++b.launches;
    }
    // long lists -- spilled into the stash by the first traversal, or written by the second -- are sorted here
    // (lists beyond 8192 ids together, by a global-memory network: needs a small table)
    if (c.sorted && (second || b.hc.big > 0))
    {
      void* long_scratch = nullptr;
      if (b.hc.max_list > 8192)
      {
        RT_CUDA(s.d_long.reserve(rtb::long_sort_scratch_bytes(total)));
        long_scratch = s.d_long.p;
      }
      RT_CUDA(rtb::launch_sort_segments(s.d_big.as<uint32_t>(), &dc->big, (uint32_t)n,
                                        s.d_offsets.as<unsigned long long>(), s.d_ids.as<uint32_t>(), st,
                                        &b.launches, b.hc.max_list, long_scratch, total));
    }
  }
  RT_CUDA(cudaEventRecord(b.ev[EV_FINISH], st));

  if (!c.keep_on_device)
  {
    if (c.counts32)
    {
      // 32-bit hits per query instead of 64-bit prefix sums: half the bytes, and nothing to rebase
      if (n)
        RT_CUDA(cudaMemcpyAsync(t->h_counts.as<uint32_t>() + b.first, s.d_counts.p, (size_t)n * sizeof(uint32_t),
                                cudaMemcpyDeviceToHost, st));
    }
    else
    {
      if (base)
        RT_CUDA(rtb::launch_add_base(s.d_offsets.as<unsigned long long>(), n + 1, base, st, &b.launches));
      const size_t n_off = (size_t)n + (last_batch ? 1 : 0);
      if (n_off)
        RT_CUDA(cudaMemcpyAsync(t->h_offsets.as<unsigned long long>() + b.first, s.d_offsets.p,
                                n_off * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    }
    if (!c.count_only && total)
      RT_CUDA(cudaMemcpyAsync(t->h_ids.as<uint32_t>() + base, s.d_ids.p, (size_t)total * sizeof(uint32_t),
                              cudaMemcpyDeviceToHost, st));
  }
  RT_CUDA(cudaEventRecord(b.ev[EV_D2H], st));
  return RT_OK;
}

// How a call is cut into batches.  B = RT_OPT_PIPELINE_BATCH.  Fewer than 2 B queries (or a device
// result): one batch.  Up to 6 B: equal batches of about B.  Beyond: a ramp B/2, B, 2 B up to batches
// of about 4 B and down again -- big batches in the middle because a launch over more queries sorts
// more coherently (2 M-query launches traverse 13 % slower per query than 16 M ones, 4 M ones 5 %),
// small ones at both ends because the first H2D copy and the last D2H copy cannot hide behind any
// traversal.  B is never taken below MIN_PIPELINE_BATCH and is doubled until the call has at most
// MAX_BATCHES batches (each batch owns EV_COUNT events).
std::vector<uint64_t> plan_batches(uint64_t n, uint64_t pipeline_batch, bool single)
{
  std::vector<uint64_t> sizes;
  uint64_t B = pipeline_batch;
  if (B != 0)
  {
    if (B < MIN_PIPELINE_BATCH)
      B = MIN_PIPELINE_BATCH;
    while (n / B > 4 * MAX_BATCHES)
      B *= 2;
  }
  if (single || B == 0 || n / B < 2)
  {
    sizes.push_back(n);
  }
  else if (n / B < 6)
  {
    const uint64_t count = n / B;
    const uint64_t per = ((n + count - 1) / count + 3) / 4 * 4;
    for (uint64_t done = 0; done < n; done += per)
      sizes.push_back(n - done < per ? n - done : per);
  }
  else
  {
    // ramp B/2, B, 2 B, [4 B ...], 2 B, B, B/2; what does not fill a 4 B batch is split evenly
    const uint64_t half = (B / 2 + 3) / 4 * 4; // >= MIN_PIPELINE_BATCH / 2, never 0
    const uint64_t ramp[3] = { half, 2 * half, 4 * half };
    const uint64_t big = 8 * half;
    uint64_t middle = n;
    size_t steps = 0;
    while (steps < 3 && middle >= 2 * ramp[steps] + (steps + 1 < 3 ? 2 * ramp[steps + 1] : big))
      middle -= 2 * ramp[steps++];
    for (size_t k = 0; k < steps; ++k)
      sizes.push_back(ramp[k]);
    const uint64_t bigs = (middle + big - 1) / big; // batches of at most 4 B
    const uint64_t per = ((middle + bigs - 1) / bigs + 3) / 4 * 4;
    for (uint64_t done = 0; done < middle; done += per)
      sizes.push_back(middle - done < per ? middle - done : per);
    for (size_t k = steps; k-- > 0;)
      sizes.push_back(ramp[k]);
  }
  return sizes;
}

// RTB_BATCH_PLAN="a,b,*K,c,d" (a tuning tool, like RTB_TRACE_BATCHES): batches of a, b queries, then what is
// left over in K equal batches, then c, d.  Ignored unless it is well-formed and adds up to the call
std::vector<uint64_t> plan_from_env(uint64_t n, const std::vector<uint64_t>& fallback)
{
  const char* e = std::getenv("RTB_BATCH_PLAN");
  if (e == nullptr || fallback.size() < 2)
    return fallback;
  std::vector<uint64_t> sizes;
  uint64_t fixed = 0, split = 0;
  size_t split_at = 0;
  while (*e != '\0')
  {
    const bool star = *e == '*';
    char* end = nullptr;
    const unsigned long long v = std::strtoull(star ? e + 1 : e, &end, 10);
    if (end == (star ? e + 1 : e) || v == 0 || (star && split != 0))
      return fallback;
    if (star)
      split = v, split_at = sizes.size();
    else if (v < MIN_PIPELINE_BATCH || v % 4 != 0)
      return fallback;
    else
      sizes.push_back(v), fixed += v;
    e = *end == ',' ? end + 1 : end;
    if (*end != ',' && *end != '\0')
      return fallback;
  }
  if (fixed > n || (split == 0 && fixed != n) || (split != 0 && (n - fixed) / split < MIN_PIPELINE_BATCH))
    return fallback;
  if (split != 0)
  {
    const uint64_t rest = n - fixed, per = ((rest + split - 1) / split + 3) / 4 * 4;
    std::vector<uint64_t> middle;
    for (uint64_t done = 0; done < rest; done += per)
      middle.push_back(rest - done < per ? rest - done : per);
    sizes.insert(sizes.begin() + (long)split_at, middle.begin(), middle.end());
  }
  return sizes.size() <= 4 * MAX_BATCHES ? sizes : fallback;
}

// everything of a call that touches the device; on failure the caller quiets the handle (abandon_call)
int run_box_call(rt_tree* t, BoxCall& c, const void* boxes, uint64_t n, rt_hits* out)
{
  int rc = RT_OK;
  const std::vector<uint64_t> sizes = plan_from_env(n, plan_batches(n, t->pipeline_batch, c.keep_on_device));
  const uint64_t n_batches = sizes.size();
  const bool pipelined = n_batches > 1;
  c.pipelined = pipelined;
  std::vector<Batch> batches((size_t)n_batches);
  uint64_t next_first = 0;
  cudaEvent_t* unused = nullptr;
  if ((rc = batch_events(t, (size_t)n_batches - 1, &unused)) != RT_OK)
    return rc;
  for (uint64_t k = 0; k < n_batches; ++k)
  {
    Batch& b = batches[(size_t)k];
    b.slot = &t->slot[pipelined ? k % NSLOT : 0];
    b.st = pipelined ? b.slot->stream : t->stream;
    b.first = next_first;
    b.n = sizes[(size_t)k];
    next_first += b.n;
    b.queries = static_cast<const unsigned char*>(boxes) + (size_t)b.first * c.query_stride;
    b.ev = t->events.data() + (size_t)k * EV_COUNT;
  }
  if (!c.keep_on_device)
  {
    if (c.counts32)
      RT_CUDA(t->h_counts.reserve((size_t)(n + 1) * sizeof(uint32_t)));
    else
      RT_CUDA(t->h_offsets.reserve((size_t)(n + 1) * sizeof(unsigned long long)));
    if (pipelined && !c.count_only && t->ids_wanted)
      RT_CUDA(t->h_ids.reserve((size_t)(t->ids_wanted + t->ids_wanted / 32 + 1024) * sizeof(uint32_t)));
  }

  // ---- issue ---------------------------------------------------------------------------------
  cudaStream_t st = t->stream;
  RT_CUDA(cudaEventRecord(t->ev_fork, st));
  if (pipelined)
  {
    for (Slot& s : t->slot)
      RT_CUDA(cudaStreamWaitEvent(s.stream, t->ev_fork, 0));
  }
  unsigned long long total = 0;
  uint64_t done_queries = 0;
  const uint64_t lag = pipelined ? NSLOT - 1 : 0;
  for (uint64_t k = 0; k < n_batches + lag; ++k)
  {
    if (k < n_batches && (rc = box_issue_count(t, c, batches[(size_t)k])) != RT_OK)
      return rc;
    if (k < lag)
      continue;
    Batch& b = batches[(size_t)(k - lag)];
    if ((rc = box_wait_counted(b)) != RT_OK)
      return rc;
    if (!c.keep_on_device && !c.count_only)
    {
      // size the pinned id buffer: exactly for a single batch, by extrapolation while batches
      // are still to come (the first call of a size pays for the growth, later ones do not)
      const unsigned long long need = total + b.hc.total;
      unsigned long long hint = need;
      if (pipelined)
      {
        const double seen = (double)(done_queries + b.n);
        hint = (unsigned long long)((double)need / (seen > 0 ? seen : 1.0) * (double)n * 1.03) + 4096;
      }
      if ((rc = grow_host_ids(t, need ? need : 1, total, hint)) != RT_OK)
        return rc;
    }
    if ((rc = box_issue_write(t, c, b, total, k - lag + 1 == n_batches)) != RT_OK)
      return rc;
    total += b.hc.total;
    done_queries += b.n;
  }
  if (pipelined)
  {
    for (Slot& s : t->slot)
    {
      RT_CUDA(cudaEventRecord(s.done, s.stream));
      RT_CUDA(cudaStreamWaitEvent(st, s.done, 0));
    }
    t->ids_wanted = total;
  }
  RT_CUDA(cudaEventRecord(t->ev_join, st));
  RT_CUDA(cudaStreamSynchronize(st));

  // ---- result ----------------------------------------------------------------------------------
  out->n_queries = n;
  out->total = total;
  if (c.keep_on_device)
  {
    Slot& s = t->slot[0];
    out->offsets = s.d_offsets.as<uint64_t>();
    out->counts = s.d_counts.as<uint32_t>();
    out->ids = c.count_only ? nullptr : s.d_ids.as<uint32_t>();
    out->memory = RT_MEM_DEVICE;
  }
  else
  {
    out->offsets = c.counts32 ? nullptr : t->h_offsets.as<uint64_t>();
    out->counts = c.counts32 ? t->h_counts.as<uint32_t>() : nullptr;
    out->ids = c.count_only ? nullptr : t->h_ids.as<uint32_t>();
    out->memory = RT_MEM_HOST;
  }

  // ---- stats: per-phase times are sums over the batches (they overlap when pipelined) ------------
  rt_stats& s = t->stats;
  s.total_ms = elapsed(t->ev_fork, t->ev_join);
  s.batches = (uint32_t)n_batches;
  uint64_t v_int = 0, v_leaf = 0;
  for (const Batch& b : batches)
  {
    s.h2d_ms += elapsed(b.ev[EV_BEGIN], b.ev[EV_H2D]);
    s.reorder_ms += elapsed(b.ev[EV_H2D], b.ev[EV_REORDER]);
    s.traverse_ms += elapsed(b.ev[EV_REORDER], b.ev[EV_TRAVERSE]);
    s.finish_ms += elapsed(b.ev[EV_TRAVERSE], b.ev[EV_COUNTED]) + elapsed(b.ev[EV_WRITE], b.ev[EV_FINISH]);
    s.d2h_ms += elapsed(b.ev[EV_FINISH], b.ev[EV_D2H]);
    s.kernel_launches += b.launches;
    s.deferred_queries += c.count_only ? 0 : b.deferred;
    v_int += b.hc.visits[0];
    v_leaf += b.hc.visits[1];
    s.traverse_steps += b.hc.visits[3];
    s.leaf_lookups += b.hc.lookups;
  }
  // RTB_TRACE_BATCHES=1: where every batch's phases lie on the call's timeline (ms after the fork),
  // to stderr -- the tool the batch ramp of plan_batches() was tuned with
  if (std::getenv("RTB_TRACE_BATCHES") != nullptr)
  {
    static const char* const names[EV_COUNT] = { "begin", "h2d", "reorder", "traverse", "counted", "write", "finish", "d2h" };
    for (size_t k = 0; k < batches.size(); ++k)
    {
      std::fprintf(stderr, "[rtb] batch %2zu n %8llu:", k, (unsigned long long)batches[k].n);
      for (int e = 0; e < EV_COUNT; ++e)
        std::fprintf(stderr, " %s %.2f", names[e], elapsed(t->ev_fork, batches[k].ev[e]));
      std::fprintf(stderr, "\n");
    }
    std::fprintf(stderr, "[rtb] call: %.2f ms, %zu batches\n", s.total_ms, batches.size());
  }
  if (c.collect)
  {
    s.visits_internal = v_int;
    s.visits_leaf = v_leaf;
    s.algorithmic_bytes = algorithmic_bytes_boxes(t, n, total, v_int, v_leaf);
  }
  else
  {
    s.traverse_steps = 0;
    s.leaf_lookups = 0;
  }
  return RT_OK;
}
} // namespace

extern "C"
{

int64_t rt_plan_batches(uint64_t n, uint64_t pipeline_batch, uint64_t* sizes, uint64_t capacity)
{
  if (n >= 0xFFFFFFFFull)
    return fail(RT_E_INVALID, "rt_plan_batches: at most 2^32-2 queries per call");
  const std::vector<uint64_t> plan = plan_batches(n, pipeline_batch, false);
  for (size_t k = 0; k < plan.size() && sizes != nullptr && k < capacity; ++k)
    sizes[k] = plan[k];
  return (int64_t)plan.size();
}

int rt_query_boxes(rt_tree* t, const void* boxes, uint64_t n, uint32_t mode, uint32_t flags, rt_hits* out)
{
  if (t == nullptr || out == nullptr)
    return fail(RT_E_INVALID, "rt_query_boxes: NULL argument");
  if (boxes == nullptr && n > 0)
    return fail(RT_E_INVALID, "rt_query_boxes: boxes is NULL");
  if (mode > RT_CONTAINS)
    return fail(RT_E_INVALID, "rt_query_boxes: unknown mode");
  if (t->desc.key_kind == RT_KEY_BOX && mode == RT_POINT_IN_BOX)
    return fail(RT_E_INVALID, "rt_query_boxes: RT_POINT_IN_BOX needs a point-keyed tree "
                              "(use RT_INTERSECTS or RT_CONTAINS)");
  if (n >= 0xFFFFFFFFull)
    return fail(RT_E_INVALID, "rt_query_boxes: at most 2^32-2 queries per call");
  const bool any_hit = (flags & RT_ANY_HIT) != 0;
  if (any_hit && (flags & RT_COLLECT_VISITS) != 0)
    return fail(RT_E_INVALID, "rt_query_boxes: RT_ANY_HIT cannot be combined with RT_COLLECT_VISITS");
  BoxCall c;
  c.launch = rtb::find_box_kernel(t->desc.scalar, t->desc.dim, t->desc.width, t->desc.key_kind, mode);
  if (c.launch == nullptr)
    return fail(RT_E_UNSUPPORTED, "rt_query_boxes: no kernel for this tree");
  int rc = use_device(t);
  if (rc != RT_OK)
    return rc;

  std::memset(out, 0, sizeof *out);
  std::memset(&t->stats, 0, sizeof t->stats);
  c.query_stride = (size_t)2 * t->desc.dim * (t->desc.scalar == RT_F64 ? 8u : 4u);
  c.on_device = (flags & RT_QUERIES_ON_DEVICE) != 0;
  c.keep_on_device = (flags & RT_RESULT_ON_DEVICE) != 0;
  c.count_only = (flags & RT_COUNT_ONLY) != 0 || any_hit;
  c.any_hit = any_hit;
  c.collect = (flags & RT_COLLECT_VISITS) != 0;
  c.sorted = (flags & RT_UNSORTED) == 0;
  c.reorder = (flags & RT_NO_REORDER) == 0;
  c.two_pass = c.collect || c.count_only;
  c.counts32 = (flags & RT_RESULT_COUNTS32) != 0 && !c.keep_on_device;
  c.strict = (flags & RT_STRICT_OVERLAP) != 0;
  c.pipelined = false;
  if ((rc = run_box_call(t, c, boxes, n, out)) != RT_OK)
  {
    std::memset(out, 0, sizeof *out);
    return abandon_call(t, rc);
  }
  return RT_OK;
}

} // extern "C"
