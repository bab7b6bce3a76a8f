// rt_api.cu -- the C ABI of include/rtree_b200.h: device buffers, streams, events and the
// per-call pipeline around the traversal kernels.  No query-path arithmetic lives here.
//
// rt_query_boxes pipeline (default flags), per batch of queries:
//   H2D queries -> reorder (Morton cells) -> traversal PASS_STASH (counts + sorted short
//   lists into the stash) -> scan (CSR offsets) -> [64-byte D2H: total, #deferred] ->
//   gather stash -> PASS_WRITE over the deferred queries only -> sort their long lists ->
//   D2H offsets + ids.
// With RT_COLLECT_VISITS / RT_COUNT_ONLY the plain two-pass form (PASS_COUNT[_STATS],
// scan, PASS_WRITE over all queries) runs instead.
// A call whose result stays on the device is ONE batch on the handle's stream.  A large call
// with a host result is cut into batches that rotate through NSLOT working sets on their own
// streams (forked from / joined to the handle's stream), so the PCIe copies of neighbouring
// batches run under the traversal of the current one.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "rt_common.cuh"

namespace
{
thread_local std::string g_last_error = "";

int fail(int code, const std::string& what)
{
  g_last_error = what;
  return code;
}

#define RT_CUDA(expr)                                                                         \
  do                                                                                          \
  {                                                                                           \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
    {                                                                                         \
      cudaGetLastError();                                                                     \
      return fail(e_ == cudaErrorMemoryAllocation ? RT_E_NOMEM : RT_E_CUDA,                   \
                  std::string(#expr) + ": " + cudaGetErrorString(e_));                        \
    }                                                                                         \
  } while (0)

// grow-only device / pinned-host buffers owned by a handle
struct DevBuf
{
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes)
  {
    if (bytes <= cap)
      return cudaSuccess;
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess)
    {
      cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      if (e != cudaSuccess)
        return e;
      cap = bytes;
      return cudaSuccess;
    }
    cap = want;
    return cudaSuccess;
  }
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const
  {
    return static_cast<T*>(p);
  }
};
struct HostBuf
{
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes)
  {
    if (bytes <= cap)
      return cudaSuccess;
    if (p)
      cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess)
      return e;
    cap = want;
    return cudaSuccess;
  }
  void release()
  {
    if (p)
      cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const
  {
    return static_cast<T*>(p);
  }
};

// device-side counters of one batch (one 64-byte block, zeroed per batch)
struct Counters
{
  unsigned long long unused;
  unsigned long long stash_cursor;
  unsigned long long visits[4];   // internal visits, leaf visits, rays with a hit, loop steps
  unsigned long long total;       // filled by copying offsets[n]
  uint32_t deferred;
  uint32_t big;

};
static_assert(sizeof(Counters) == 64, "Counters layout");

enum
{
  EV_BEGIN = 0,
  EV_H2D,
  EV_REORDER,
  EV_TRAVERSE,
  EV_COUNTED, // scan done, counters on their way to the host
  EV_WRITE,   // phase 2 begins
  EV_FINISH,
  EV_D2H,
  EV_COUNT
};

constexpr int NSLOT = 3;
constexpr uint64_t DEFAULT_PIPELINE_BATCH = 1u << 20; // base batch size of a pipelined call (queries; rays: twice)

// The working set of one batch of queries.  A call that leaves its result on the device, or a
// small one, is a single batch on slot 0 and the caller's stream.  A large call with a host
// result is cut into batches that rotate through the slots, each slot on its own stream, so
// that the H2D copy of batch k+1 and the D2H copy of batch k-1 run under the traversal of
// batch k (PCIe is full duplex and both copy engines are otherwise idle).
struct Slot
{
  DevBuf d_queries, d_order, d_counts, d_offsets, d_ids, d_stash, d_stash_pos, d_deferred, d_big, d_scratch,
      d_counters, d_cursors;
  HostBuf h_counters;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  unsigned long long stash_wanted = 0; // pool size the last batch on this slot would have needed
  uint32_t stash_slab = rtb::STASH_SLAB; // slab size for the next batch (grows with the lists seen)
  unsigned long long last_total = 0, last_n = 0; // hits / queries of the last batch on this slot

  void release()
  {
    DevBuf* dev[] = { &d_queries, &d_order, &d_counts, &d_offsets, &d_ids, &d_stash,
                      &d_stash_pos, &d_deferred, &d_big, &d_scratch, &d_counters, &d_cursors };
    for (DevBuf* b : dev)
      b->release();
    h_counters.release();
    if (done)
      cudaEventDestroy(done);
    if (stream)
      cudaStreamDestroy(stream);
    done = nullptr;
    stream = nullptr;
  }
};
} // namespace

struct rt_tree
{
  int device = 0;
  rt_tree_desc desc;
  std::vector<uint32_t> levels;
  rtb::TreeDev dev;
  unsigned char* d_blocks = nullptr;
  unsigned char* d_qblocks = nullptr; // quantised image (3-D float32 point trees, W = 8), else NULL
  bool use_quant = true;              // RT_OPT_QUANTISED
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<cudaEvent_t> events; // EV_COUNT per batch of a call, grow-only

  Slot slot[NSLOT];
  DevBuf d_ray_t, d_ray_id, d_ray_any, d_knn_ids, d_knn_d2;
  HostBuf h_offsets, h_ids, h_ray_t, h_ray_id, h_ray_any, h_batch_counters, h_knn_ids, h_knn_d2;
  uint64_t pipeline_batch = DEFAULT_PIPELINE_BATCH;
  uint32_t query_chunk = 0, work_ranges = 0; // 0 = the kernels' defaults
  int traverse_grid = 0;                     // 0 = fill the GPU
  unsigned long long ids_wanted = 0; // hit total of the last pipelined call (sizes h_ids up front)
  rt_stats stats;
};

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

int use_device(const rt_tree* t)
{
  RT_CUDA(cudaSetDevice(t->device));
  return RT_OK;
}

float elapsed(cudaEvent_t a, cudaEvent_t b)
{
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess)
  {
    cudaGetLastError();
    return 0.f;
  }
  return ms;
}

// the EV_COUNT events of batch `index` of the running call
int batch_events(rt_tree* t, size_t index, cudaEvent_t** out)
{
  const size_t need = (index + 1) * EV_COUNT;
  while (t->events.size() < need)
  {
    cudaEvent_t e = nullptr;
    RT_CUDA(cudaEventCreate(&e));
    t->events.push_back(e);
  }
  *out = t->events.data() + index * EV_COUNT;
  return RT_OK;
}

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
  bool on_device, keep_on_device, count_only, collect, sorted, reorder, two_pass, pipelined, any_hit;
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
  if (reorder)
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
  if (reorder)
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
  if (!t->use_quant)
    L.tree.qblocks = nullptr;
  L.queries = d_q;
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
  L.work.cursors = s.d_cursors.as<unsigned int>();
  L.work.chunk = t->query_chunk;
  L.work.ranges = t->work_ranges;
  L.grid = t->traverse_grid ? t->traverse_grid : (c.pipelined ? -1 : 0);
  L.visit_stats = dc->visits;
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
  RT_CUDA(cudaMemcpyAsync(&dc->total, s.d_offsets.as<unsigned long long>() + n, sizeof(unsigned long long),
                          cudaMemcpyDeviceToDevice, st));
  RT_CUDA(cudaMemcpyAsync(s.h_counters.p, dc, sizeof(Counters), cudaMemcpyDeviceToHost, st));
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
    if (c.sorted && (second || b.hc.big > 0))
      RT_CUDA(rtb::launch_sort_segments(s.d_big.as<uint32_t>(), &dc->big, (uint32_t)n,
                                        s.d_offsets.as<unsigned long long>(), s.d_ids.as<uint32_t>(), st,
                                        &b.launches));
  }
  RT_CUDA(cudaEventRecord(b.ev[EV_FINISH], st));

  if (!c.keep_on_device)
  {
    if (base)
      RT_CUDA(rtb::launch_add_base(s.d_offsets.as<unsigned long long>(), n + 1, base, st, &b.launches));
    const size_t n_off = (size_t)n + (last_batch ? 1 : 0);
    if (n_off)
      RT_CUDA(cudaMemcpyAsync(t->h_offsets.as<unsigned long long>() + b.first, s.d_offsets.p,
                              n_off * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    if (!c.count_only && total)
      RT_CUDA(cudaMemcpyAsync(t->h_ids.as<uint32_t>() + base, s.d_ids.p, (size_t)total * sizeof(uint32_t),
                              cudaMemcpyDeviceToHost, st));
  }
  RT_CUDA(cudaEventRecord(b.ev[EV_D2H], st));
  return RT_OK;
}

// Make the pinned id buffer hold `need` ids, keeping the first `keep` (copies into it may be in
// flight on the slot streams).  `hint` is the expected final size.
int grow_host_ids(rt_tree* t, unsigned long long need, unsigned long long keep, unsigned long long hint)
{
  if ((size_t)need * sizeof(uint32_t) <= t->h_ids.cap)
    return RT_OK;
  for (Slot& s : t->slot)
    RT_CUDA(cudaStreamSynchronize(s.stream));
  HostBuf bigger;
  const unsigned long long want = need > hint ? need : hint;
  RT_CUDA(bigger.reserve((size_t)want * sizeof(uint32_t)));
  if (keep)
    std::memcpy(bigger.p, t->h_ids.p, (size_t)keep * sizeof(uint32_t));
  t->h_ids.release();
  t->h_ids = bigger;
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
    t->pipeline_batch = std::strtoull(batch, nullptr, 10);

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
  if (rtb::quant_supported(desc->scalar, desc->dim, desc->width, desc->key_kind)
      && std::getenv("RTB_NO_QUANT") == nullptr)
  {
    e = cudaMalloc((void**)&t->d_qblocks, rtb::quant_image_bytes(desc->num_nodes, desc->dim, desc->width));
    if (e != cudaSuccess)
      return cleanup(RT_E_NOMEM, std::string("rt_upload: cudaMalloc(quantised image): ") + cudaGetErrorString(e));
    e = rtb::build_quant_image(t->dev, t->desc, desc->level_node_begin[desc->leaf_level], t->d_qblocks, t->stream);
    if (e != cudaSuccess)
      return cleanup(RT_E_CUDA, std::string("rt_upload: quantised image: ") + cudaGetErrorString(e));
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
  HostBuf* host[] = { &t->h_offsets, &t->h_ids,  &t->h_ray_t,    &t->h_ray_id, &t->h_ray_any, &t->h_batch_counters,
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
  case RT_OPT_PIPELINE_BATCH: t->pipeline_batch = value; return RT_OK;
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
  const bool any_hit = (flags & RT_ANY_HIT) != 0;
  if (any_hit && (flags & RT_COLLECT_VISITS) != 0)
    return fail(RT_E_INVALID, "rt_query_boxes: RT_ANY_HIT cannot be combined with RT_COLLECT_VISITS");
  c.count_only = (flags & RT_COUNT_ONLY) != 0 || any_hit;
  c.any_hit = any_hit;
  c.collect = (flags & RT_COLLECT_VISITS) != 0;
  c.sorted = (flags & RT_UNSORTED) == 0;
  c.reorder = (flags & RT_NO_REORDER) == 0;
  c.two_pass = c.collect || c.count_only;

  // ---- how the call is cut into batches ------------------------------------------------------
  // B = RT_OPT_PIPELINE_BATCH.  Fewer than 2 B queries (or a device result): one batch.  Up to 6 B:
  // equal batches of about B.  Beyond: a ramp B/2, B, 2 B up to batches of about 4 B and down again
  // -- big batches in the middle because a launch over more queries sorts more coherently (2 M-query
  // launches traverse 13 % slower per query than 16 M ones, 4 M ones 5 %), small ones at both ends
  // because the first H2D copy and the last D2H copy cannot hide behind any traversal.
  std::vector<uint64_t> sizes;
  const uint64_t B = t->pipeline_batch;
  if (c.keep_on_device || B == 0 || n / B < 2)
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
    const uint64_t half = (B / 2 + 3) / 4 * 4;
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
  const uint64_t n_batches = sizes.size();
  const bool pipelined = n_batches > 1;
  c.pipelined = pipelined;
  std::vector<Batch> batches((size_t)n_batches);
  uint64_t next_first = 0;
  for (uint64_t k = 0; k < n_batches; ++k)
  {
    Batch& b = batches[(size_t)k];
    b.slot = &t->slot[pipelined ? k % NSLOT : 0];
    b.st = pipelined ? b.slot->stream : t->stream;
    b.first = next_first;
    b.n = sizes[(size_t)k];
    next_first += b.n;
    b.queries = static_cast<const unsigned char*>(boxes) + (size_t)b.first * c.query_stride;
    if ((rc = batch_events(t, (size_t)k, &b.ev)) != RT_OK)
      return rc;
  }
  for (uint64_t k = 0; k < n_batches; ++k) // the vector of events may have moved while growing
    batches[(size_t)k].ev = t->events.data() + (size_t)k * EV_COUNT;
  if (!c.keep_on_device)
  {
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
    out->ids = c.count_only ? nullptr : s.d_ids.as<uint32_t>();
    out->memory = RT_MEM_DEVICE;
  }
  else
  {
    out->offsets = t->h_offsets.as<uint64_t>();
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
  }
  return RT_OK;
}

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
  cudaStream_t st = t->stream;
  Slot& w = t->slot[0]; // rays run as one batch on the caller's stream
  cudaEvent_t* ev = nullptr;
  if ((rc = batch_events(t, 0, &ev)) != RT_OK)
    return rc;
  RT_CUDA(w.d_counters.reserve(sizeof(Counters)));
  RT_CUDA(w.d_cursors.reserve(rtb::MAX_WORK_RANGES * sizeof(unsigned int)));
  RT_CUDA(w.h_counters.reserve(sizeof(Counters)));
  const size_t ray_bytes = (size_t)n * 7 * sizeof(float);
  const bool on_device = (flags & RT_QUERIES_ON_DEVICE) != 0;
  const bool keep_on_device = (flags & RT_RESULT_ON_DEVICE) != 0;
  const bool count_only = (flags & RT_COUNT_ONLY) != 0;
  const bool collect = (flags & RT_COLLECT_VISITS) != 0;
  const bool sorted = (flags & RT_UNSORTED) == 0;
  uint32_t launches = 0;

  // ---- closest / any with a host result and many rays: pipelined batches ------------------------
  // (one kernel per batch; the H2D copy of the next batch and the D2H copy of the previous one
  // run under it -- at 28 bytes per ray the call is otherwise bound by the H2D copy alone)
  if (mode != RT_RAY_ALL_HITS && !keep_on_device && t->pipeline_batch > 0 && n / (2 * t->pipeline_batch) >= 2)
  {
    const uint64_t wanted = n / (2 * t->pipeline_batch);
    const uint64_t per_batch = ((n + wanted - 1) / wanted + 15) / 16 * 16; // keeps every copy 16-byte aligned
    const uint64_t n_batches = (n + per_batch - 1) / per_batch;
    if (mode == RT_RAY_CLOSEST)
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
      if (!on_device)
        RT_CUDA(s.d_queries.reserve((size_t)per_batch * 7 * sizeof(float)));
      RT_CUDA(cudaEventRecord(bev[EV_BEGIN], bs));
      const float* d_rays = rays + first * 7;
      if (!on_device)
      {
        RT_CUDA(cudaMemcpyAsync(s.d_queries.p, rays + first * 7, (size_t)m * 7 * sizeof(float),
                                cudaMemcpyHostToDevice, bs));
        d_rays = s.d_queries.as<float>();
      }
      RT_CUDA(cudaMemsetAsync(s.d_counters.p, 0, sizeof(Counters), bs));
      RT_CUDA(cudaEventRecord(bev[EV_H2D], bs));
      Counters* dc = s.d_counters.as<Counters>();
      rtb::RayLaunch L;
      std::memset(&L, 0, sizeof L);
      L.tree = t->dev;
      L.width = t->desc.width;
      L.key_kind = t->desc.key_kind;
      L.rays = d_rays;
      L.n = m;
      L.mode = mode;
      L.t_out = t->d_ray_t.as<float>() + first;
      L.id_out = t->d_ray_id.as<uint32_t>() + first;
      L.any_out = t->d_ray_any.as<uint8_t>() + first;
      L.work.cursors = s.d_cursors.as<unsigned int>();
      L.work.chunk = t->query_chunk;
      L.work.ranges = t->work_ranges;
      L.grid = t->traverse_grid ? t->traverse_grid : -1;
      L.visit_stats = dc->visits;
      L.sorted = 1u;
      L.stream = bs;
      RT_CUDA(rtb::launch_ray_kernel(L));
      ++launches;
      RT_CUDA(cudaEventRecord(bev[EV_TRAVERSE], bs));
      RT_CUDA(cudaMemcpyAsync(t->h_batch_counters.as<Counters>() + k, dc, sizeof(Counters), cudaMemcpyDeviceToHost, bs));
      if (mode == RT_RAY_CLOSEST)
      {
        RT_CUDA(cudaMemcpyAsync(t->h_ray_t.as<float>() + first, L.t_out, (size_t)m * sizeof(float),
                                cudaMemcpyDeviceToHost, bs));
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
    if (mode == RT_RAY_CLOSEST)
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
    for (uint64_t k = 0; k < n_batches; ++k)
    {
      cudaEvent_t* bev = t->events.data() + (size_t)k * EV_COUNT;
      rs.h2d_ms += elapsed(bev[EV_BEGIN], bev[EV_H2D]);
      rs.traverse_ms += elapsed(bev[EV_H2D], bev[EV_TRAVERSE]);
      rs.d2h_ms += elapsed(bev[EV_TRAVERSE], bev[EV_D2H]);
      out->n_hit += t->h_batch_counters.as<Counters>()[k].visits[2];
    }
    return RT_OK;
  }

  if (!on_device)
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
  const float* d_rays = rays;
  if (!on_device)
  {
    if (ray_bytes)
      RT_CUDA(cudaMemcpyAsync(w.d_queries.p, rays, ray_bytes, cudaMemcpyHostToDevice, st));
    d_rays = w.d_queries.as<float>();
  }
  RT_CUDA(cudaMemsetAsync(w.d_counters.p, 0, sizeof(Counters), st));
  RT_CUDA(cudaEventRecord(ev[EV_H2D], st));
  RT_CUDA(cudaEventRecord(ev[EV_REORDER], st));

  Counters* dc = w.d_counters.as<Counters>();
  rtb::RayLaunch L;
  std::memset(&L, 0, sizeof L);
  L.tree = t->dev;
  L.width = t->desc.width;
  L.key_kind = t->desc.key_kind;
  L.rays = d_rays;
  L.order = nullptr;
  L.n = n;
  L.mode = mode;
  L.pass = collect ? rtb::PASS_COUNT_STATS : rtb::PASS_COUNT;
  L.counts = w.d_counts.as<uint32_t>();
  L.offsets = w.d_offsets.as<unsigned long long>();
  L.big_list = w.d_big.as<uint32_t>();
  L.big_count = &dc->big;
  L.t_out = t->d_ray_t.as<float>();
  L.id_out = t->d_ray_id.as<uint32_t>();
  L.any_out = t->d_ray_any.as<uint8_t>();
  L.work.cursors = w.d_cursors.as<unsigned int>();
  L.work.chunk = t->query_chunk;
  L.work.ranges = t->work_ranges;
  L.grid = t->traverse_grid;
  L.visit_stats = dc->visits;
  L.sorted = sorted ? 1u : 0u;
  L.stream = st;
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

  if (mode == RT_RAY_ALL_HITS && !count_only)
  {
    RT_CUDA(w.d_ids.reserve((size_t)(total ? total : 1) * sizeof(uint32_t)));
    if (n > 0 && total > 0)
    {
      L.pass = rtb::PASS_WRITE;
      L.ids = w.d_ids.as<uint32_t>();
      RT_CUDA(rtb::launch_ray_kernel(L));
      ++launches;
      if (sorted)
        RT_CUDA(rtb::launch_sort_segments(w.d_big.as<uint32_t>(), &dc->big, (uint32_t)n,
                                          w.d_offsets.as<unsigned long long>(), w.d_ids.as<uint32_t>(), st,
                                          &launches));
    }
  }
  RT_CUDA(cudaEventRecord(ev[EV_FINISH], st));

  out->n_rays = n;
  out->memory = keep_on_device ? RT_MEM_DEVICE : RT_MEM_HOST;
  if (mode == RT_RAY_ALL_HITS)
  {
    out->n_hit = total;
    out->hits.n_queries = n;
    out->hits.total = total;
    out->hits.memory = out->memory;
    if (keep_on_device)
    {
      out->hits.offsets = w.d_offsets.as<uint64_t>();
      out->hits.ids = count_only ? nullptr : w.d_ids.as<uint32_t>();
    }
    else
    {
      RT_CUDA(t->h_offsets.reserve((size_t)(n + 1) * sizeof(unsigned long long)));
      RT_CUDA(cudaMemcpyAsync(t->h_offsets.p, w.d_offsets.p, (size_t)(n + 1) * sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, st));
      if (!count_only)
      {
        RT_CUDA(t->h_ids.reserve((size_t)(total ? total : 1) * sizeof(uint32_t)));
        if (total)
          RT_CUDA(cudaMemcpyAsync(t->h_ids.p, w.d_ids.p, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                  st));
      }
      out->hits.offsets = t->h_offsets.as<uint64_t>();
      out->hits.ids = count_only ? nullptr : t->h_ids.as<uint32_t>();
    }
  }
  else if (mode == RT_RAY_CLOSEST)
  {
    out->n_hit = hc.visits[2];
    if (keep_on_device)
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
    if (keep_on_device)
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
  if (collect && mode == RT_RAY_ALL_HITS)
  {
    s.visits_internal = hc.visits[0];
    s.visits_leaf = hc.visits[1];
    s.traverse_steps = hc.visits[3];
    const uint64_t M = t->desc.max_entries;
    const uint64_t b_int = M * (2 * 3 * 4 + 4);
    const uint64_t b_leaf = t->desc.key_kind == RT_KEY_BOX ? b_int : M * (3 * 4 + 4);
    s.algorithmic_bytes = hc.visits[0] * b_int + hc.visits[1] * b_leaf + n * (28 + 8) + 4 * total;
  }
  return RT_OK;
}

int rt_query_knn(rt_tree* t, const float* points, uint64_t n, uint32_t k, uint32_t flags, rt_knn_result* out)
{
  if (t == nullptr || out == nullptr)
    return fail(RT_E_INVALID, "rt_query_knn: NULL argument");
  if (points == nullptr && n > 0)
    return fail(RT_E_INVALID, "rt_query_knn: points is NULL");
  if (k < 1 || k > 32)
    return fail(RT_E_INVALID, "rt_query_knn: k must be 1..32");
  if (!rtb::knn_supported(t->desc.scalar, t->desc.dim, t->desc.width))
    return fail(RT_E_UNSUPPORTED, "rt_query_knn: kNN queries need a float32 tree with dim 2 or 3 and 8-wide blocks");
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

int rt_refit(rt_tree* t, const void* keys, uint64_t n_keys, uint32_t flags)
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
