// rt_handle.cuh -- what the translation units behind the C ABI share (internal header): the handle
// (rt_tree), its grow-only buffers, the per-batch working sets and the error plumbing.
//
//   rt_api.cu        rt_upload / rt_free / options / stats
//   rt_api_boxes.cu  rt_query_boxes: the batch planner and the two-phase pipeline
//   rt_api_rays.cu   rt_query_rays
//   rt_api_knn.cu    rt_query_knn, rt_refit
#pragma once

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "rt_common.cuh"

namespace rtb
{
// records the text rt_last_error() returns (thread-local) and hands the code back
int fail(int code, const std::string& what);
} // namespace rtb

#define RT_CUDA(expr)                                                                         \
  do                                                                                          \
  {                                                                                           \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
    {                                                                                         \
      cudaGetLastError();                                                                     \
      return rtb::fail(e_ == cudaErrorMemoryAllocation ? RT_E_NOMEM : RT_E_CUDA,              \
                       std::string(#expr) + ": " + cudaGetErrorString(e_));                   \
    }                                                                                         \
  } while (0)

namespace rtb
{
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
    // (mapped: the counter blocks are written by a store from the SM, launch_publish_counters)
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocMapped);
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
  unsigned long long lookups;     // leaf candidates looked up on the float records (quantised kernels, COUNT_STATS)
  unsigned long long stash_cursor;
  unsigned long long visits[4];   // internal visits, leaf visits, rays with a hit, loop steps
  unsigned long long total;       // filled by copying offsets[n]
  uint32_t deferred;
  uint32_t big;
  uint32_t max_list;              // longest list that goes to the segment sort
  uint32_t pad;
};
static_assert(sizeof(Counters) == 72, "Counters layout");

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
constexpr uint64_t MIN_PIPELINE_BATCH = 4096;         // smaller values of RT_OPT_PIPELINE_BATCH are refused
constexpr uint64_t MAX_BATCHES = 1024;                // a call is never cut finer than about this

// The working set of one batch of queries.  A call that leaves its result on the device, or a
// small one, is a single batch on slot 0 and the caller's stream.  A large call with a host
// result is cut into batches that rotate through the slots, each slot on its own stream, so
// that the H2D copy of batch k+1 and the D2H copy of batch k-1 run under the traversal of
// batch k (PCIe is full duplex and both copy engines are otherwise idle).
struct Slot
{
  DevBuf d_queries, d_packed, d_order, d_counts, d_offsets, d_ids, d_stash, d_stash_pos, d_deferred, d_big, d_scratch,
      d_counters, d_cursors, d_long;
  HostBuf h_counters;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  unsigned long long stash_wanted = 0; // pool size the last batch on this slot would have needed
  uint32_t stash_slab = rtb::STASH_SLAB; // slab size for the next batch (grows with the lists seen)
  unsigned long long last_total = 0, last_n = 0; // hits / queries of the last batch on this slot

  void release()
  {
    DevBuf* dev[] = { &d_queries, &d_packed, &d_order, &d_counts, &d_offsets, &d_ids, &d_stash,
                      &d_stash_pos, &d_deferred, &d_big, &d_scratch, &d_counters, &d_cursors, &d_long };
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
} // namespace rtb

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

  rtb::Slot slot[rtb::NSLOT];
  rtb::DevBuf d_ray_t, d_ray_id, d_ray_any, d_knn_ids, d_knn_d2;
  rtb::HostBuf h_offsets, h_counts, h_ids, h_ray_t, h_ray_id, h_ray_any, h_batch_counters, h_knn_ids, h_knn_d2;
  uint64_t pipeline_batch = rtb::DEFAULT_PIPELINE_BATCH;
  uint32_t query_chunk = 0, work_ranges = 0; // 0 = the kernels' defaults
  int traverse_grid = 0;                     // 0 = fill the GPU
  unsigned long long ids_wanted = 0; // hit total of the last pipelined call (sizes h_ids up front)
  rt_stats stats;
};

namespace rtb
{
inline int use_device(const rt_tree* t)
{
  RT_CUDA(cudaSetDevice(t->device));
  return RT_OK;
}

inline float elapsed(cudaEvent_t a, cudaEvent_t b)
{
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess)
  {
    cudaGetLastError();
    return 0.f;
  }
  return ms;
}

// the EV_COUNT events of batch `index` of the running call (the vector may move while it grows:
// take pointers into it only after the last call)
inline int batch_events(rt_tree* t, size_t index, cudaEvent_t** out)
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

// Make the pinned id buffer hold `need` ids, keeping the first `keep` (copies into it may be in
// flight on the slot streams).  `hint` is the expected final size.
inline int grow_host_ids(rt_tree* t, unsigned long long need, unsigned long long keep, unsigned long long hint)
{
  if ((size_t)need * sizeof(uint32_t) <= t->h_ids.cap)
    return RT_OK;
  // test knob: the pinned id buffer may not grow beyond this many ids (exercises the failure path of
  // a pipelined call: the call returns RT_E_NOMEM and the handle must stay usable)
  if (const char* cap = std::getenv("RTB_HOST_IDS_LIMIT"))
  {
    const unsigned long long limit = std::strtoull(cap, nullptr, 10);
    if (need > limit)
      return fail(RT_E_NOMEM, "rt_query_boxes: pinned id buffer limited by RTB_HOST_IDS_LIMIT");
    if (hint > limit)
      hint = limit;
  }
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

// After a failure in the middle of a call: wait for everything the call has put on the handle's
// streams, so that no batch is still writing into the handle's buffers and the next call starts
// from a quiet handle (the slot streams of a pipelined call are forked from the caller's stream and
// would otherwise never be joined).  Returns `code`.
inline int abandon_call(rt_tree* t, int code)
{
  const std::string what = std::string(rt_last_error());
  for (Slot& s : t->slot)
  {
    if (s.stream)
      cudaStreamSynchronize(s.stream);
  }
  if (t->stream)
    cudaStreamSynchronize(t->stream);
  cudaGetLastError();
  return fail(code, what);
}
} // namespace rtb
