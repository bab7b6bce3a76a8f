// rt_passes.cu -- the small passes around the traversal kernels (hand-written, sm_100a):
//   * 64-bit exclusive scan of the per-query hit counts  -> CSR offsets
//   * gather of the stashed (already sorted) hit lists to their CSR position
//   * ascending sort of the long hit lists the traversal could not sort in a warp
//   * query reordering by Morton cell (coherence pass)
// All are bandwidth-trivial next to the traversal (a few bytes per query).
#include "rt_common.cuh"
#include "rt_quant.cuh"

#include <cstdlib>

namespace rtb
{

// SMs of the current device (148 on a B200); grids of the grid-stride passes are multiples of it
static unsigned sm_count()
{
  int device = 0, sms = 0;
  if (cudaGetDevice(&device) != cudaSuccess
      || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0)
  {
    cudaGetLastError();
    return 148u;
  }
  return (unsigned)sms;
}

// ================================ work distribution =================================
cudaError_t plan_work(const void* kernel, size_t smem, uint64_t n, int forced_grid, cudaStream_t stream,
                      WorkRanges& work, int& grid)
{
  cudaError_t err = cudaSuccess;
  grid = forced_grid;
  if (grid <= 0)
  {
    // forced_grid < 0: leave that many CTA slots per SM free (pipelined calls: the small kernels
    // of the neighbouring batches must not queue behind a GPU-filling persistent kernel)
    const int spare = -forced_grid;
    int device = 0, sms = 0, per_sm = 0;
    if ((err = cudaGetDevice(&device)) != cudaSuccess)
      return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess)
      return err;
    if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TRAVERSE_THREADS, smem))
        != cudaSuccess)
      return err;
    per_sm = per_sm - spare < 1 ? 1 : per_sm - spare;
    grid = sms * per_sm; // persistent: every resident warp pulls chunks until none is left
  }
  const uint32_t chunk = work.chunk >= 1 && work.chunk <= 1024 ? work.chunk : QUERY_CHUNK;
  const unsigned long long chunks = (n + chunk - 1) / chunk;
  const unsigned long long needed = (chunks + TRAVERSE_THREADS / 32 - 1) / (TRAVERSE_THREADS / 32);
  if ((unsigned long long)grid > needed)
    grid = (int)(needed > 0 ? needed : 1);
  // default: one range per SM, warps start on the range of the SM they run on; an explicit
  // count cuts by CTA index instead (1 = a single global cursor)
  work.by_sm = work.ranges == 0 ? 1u : 0u;
  uint32_t ranges = work.ranges;
  if (ranges == 0)
  {
    int device = 0, sms = 0;
    if ((err = cudaGetDevice(&device)) != cudaSuccess)
      return err;
    if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess)
      return err;
    ranges = (uint32_t)sms;
  }
  if (ranges > (uint32_t)grid && !work.by_sm)
    ranges = (uint32_t)grid;
  if (ranges > MAX_WORK_RANGES)
    ranges = MAX_WORK_RANGES;
  const unsigned long long per = (n + ranges - 1) / ranges;
  work.ranges = ranges;
  work.chunk = chunk;
  work.per_range = (uint32_t)((per + chunk - 1) / chunk * chunk);
  if (work.per_range == 0)
    work.per_range = chunk;
  return cudaMemsetAsync(work.cursors, 0, (size_t)ranges * sizeof(unsigned int), stream);
}

// ================================ exclusive scan ====================================
namespace
{
// 256-thread CTAs on purpose: in a pipelined call these passes run in the one CTA slot per SM that the
// persistent traversal kernel of a neighbouring batch leaves free (8 warps); a 1024-thread CTA could not
// start before that traversal had drained
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned long long warp_inclusive(unsigned long long v, int lane)
{
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    const unsigned long long up = __shfl_up_sync(0xFFFFFFFFu, v, d);
    if (lane >= d)
      v += up;
  }
  return v;
}

// inclusive scan of one value per thread across the block; returns the block total via *total
__device__ __forceinline__ unsigned long long block_inclusive(unsigned long long v, unsigned long long* warp_sums,
                                                              unsigned long long* total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_inclusive(v, lane);
  if (lane == 31)
    warp_sums[warp] = v;
  __syncthreads();
  if (warp == 0)
  {
    unsigned long long s = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0ull;
    s = warp_inclusive(s, lane);
    warp_sums[lane] = s;
  }
  __syncthreads();
  if (warp > 0)
    v += warp_sums[warp - 1];
  *total = warp_sums[(blockDim.x >> 5) - 1];
  return v;
}

__device__ __forceinline__ void load_tile(const uint32_t* counts, uint64_t n, uint64_t base, uint32_t (&x)[SCAN_ITEMS])
{
  if (base + SCAN_ITEMS <= n)
  {
    const uint4 v = *reinterpret_cast<const uint4*>(counts + base);
    x[0] = v.x, x[1] = v.y, x[2] = v.z, x[3] = v.w;
  }
  else
  {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
      x[k] = (base + k < n) ? counts[base + k] : 0u;
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const uint32_t* counts, uint64_t n,
                                                               unsigned long long* tile_sums)
{
  __shared__ unsigned long long warp_sums[32];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t x[SCAN_ITEMS];
  load_tile(counts, n, base, x);
  unsigned long long total;
  block_inclusive((unsigned long long)x[0] + x[1] + x[2] + x[3], warp_sums, &total);
  if (threadIdx.x == 0)
    tile_sums[blockIdx.x] = total;
}

// one block: exclusive scan of the tile sums in place, carrying across chunks
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_offsets(unsigned long long* tile_sums, uint64_t tiles)
{
  __shared__ unsigned long long warp_sums[32];
  unsigned long long carry = 0;
  for (uint64_t base = 0; base < tiles; base += SCAN_THREADS)
  {
    const uint64_t i = base + threadIdx.x;
    const unsigned long long v = (i < tiles) ? tile_sums[i] : 0ull;
    unsigned long long total;
    const unsigned long long inc = block_inclusive(v, warp_sums, &total);
    if (i < tiles)
      tile_sums[i] = carry + inc - v;
    carry += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_write_offsets(const uint32_t* counts, uint64_t n,
                                                                   const unsigned long long* tile_offsets,
                                                                   unsigned long long* offsets)
{
  __shared__ unsigned long long warp_sums[32];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t x[SCAN_ITEMS];
  load_tile(counts, n, base, x);
  const unsigned long long mine = (unsigned long long)x[0] + x[1] + x[2] + x[3];
  unsigned long long total;
  const unsigned long long inc = block_inclusive(mine, warp_sums, &total);
  unsigned long long run = tile_offsets[blockIdx.x] + inc - mine;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
  {
    if (base + k < n)
      offsets[base + k] = run;
    run += x[k];
    if (base + k + 1 == n)
      offsets[n] = run;
  }
}

__global__ void scan_empty(unsigned long long* offsets) { offsets[0] = 0; }
} // namespace

size_t scan_scratch_bytes(uint64_t n)
{
  const uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  return (size_t)(tiles + 1) * sizeof(unsigned long long);
}

cudaError_t launch_exclusive_scan(const uint32_t* counts, uint64_t n, unsigned long long* offsets,
                                  void* scratch, cudaStream_t stream, uint32_t* launches)
{
  if (n == 0)
  {
    scan_empty<<<1, 1, 0, stream>>>(offsets);
    *launches += 1;
    return cudaGetLastError();
  }
  const uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  unsigned long long* tile_sums = static_cast<unsigned long long*>(scratch);
  scan_tile_sums<<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(counts, n, tile_sums);
  scan_tile_offsets<<<1, SCAN_THREADS, 0, stream>>>(tile_sums, tiles);
  scan_write_offsets<<<(unsigned)tiles, SCAN_THREADS, 0, stream>>>(counts, n, tile_sums, offsets);
  *launches += 3;
  return cudaGetLastError();
}

// ================================ counters to the host ==============================
namespace
{
// dst (mapped pinned host memory) = the batch's counter block, with `total` taken from offsets[n].  A
// store from the SM instead of a 64-byte cudaMemcpy: the copy engine serves copies in order, and a tiny
// read-back queued behind the previous batch's 100 MB of hit ids held the whole pipeline back.
__global__ void publish_counters_kernel(const unsigned long long* __restrict__ counters, uint32_t words,
                                        uint32_t total_word, const unsigned long long* __restrict__ total,
                                        unsigned long long* dst)
{
  const uint32_t i = threadIdx.x;
  if (i < words)
    dst[i] = (i == total_word && total != nullptr) ? *total : counters[i];
  __threadfence_system();
}
} // namespace

cudaError_t launch_publish_counters(const void* d_counters, uint32_t bytes, uint32_t total_offset_bytes,
                                    const unsigned long long* d_total, void* h_mapped, cudaStream_t stream,
                                    uint32_t* launches)
{
  void* d_view = nullptr;
  cudaError_t err = cudaHostGetDevicePointer(&d_view, h_mapped, 0);
  if (err != cudaSuccess)
    return err;
  if (bytes % 8 != 0 || bytes > 256)
    return cudaErrorInvalidValue;
  publish_counters_kernel<<<1, 32, 0, stream>>>(static_cast<const unsigned long long*>(d_counters), bytes / 8,
                                                total_offset_bytes / 8, d_total,
                                                static_cast<unsigned long long*>(d_view));
  *launches += 1;
  return cudaGetLastError();
}

// ================================ stash gather ======================================
namespace
{
// 8 lanes per query: the common list (about 10 ids) moves in two coalesced steps
__global__ void __launch_bounds__(256) gather_stash_kernel(const uint32_t* __restrict__ stash,
                                                           const uint32_t* __restrict__ stash_pos,
                                                           const unsigned long long* __restrict__ offsets,
                                                           uint64_t n, uint32_t* __restrict__ ids)
{
  const uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const uint32_t sub = threadIdx.x & 7;
  if (q >= n)
    return;
  const uint32_t from = stash_pos[q];
  if (from == NO_STASH)
    return;
  const unsigned long long begin = offsets[q];
  const uint32_t count = (uint32_t)(offsets[q + 1] - begin);
  for (uint32_t k = sub; k < count; k += 8)
    ids[begin + k] = stash[(unsigned long long)from + k];
}
} // namespace

cudaError_t launch_gather_stash(const uint32_t* stash, const uint32_t* stash_pos,
                                const unsigned long long* offsets, uint64_t n, uint32_t* ids,
                                cudaStream_t stream, uint32_t* launches)
{
  if (n == 0)
    return cudaSuccess;
  const uint64_t threads = n * 8;
  const uint64_t blocks = (threads + 255) / 256;
  gather_stash_kernel<<<(unsigned)blocks, 256, 0, stream>>>(stash, stash_pos, offsets, n, ids);
  *launches += 1;
  return cudaGetLastError();
}

// ================================ offset rebasing ===================================
namespace
{
__global__ void __launch_bounds__(256) add_base_kernel(unsigned long long* offsets, uint64_t n, unsigned long long base)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    offsets[i] += base;
}
} // namespace

cudaError_t launch_add_base(unsigned long long* offsets, uint64_t n, unsigned long long base,
                            cudaStream_t stream, uint32_t* launches)
{
  if (n == 0 || base == 0)
    return cudaSuccess;
  add_base_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(offsets, n, base);
  *launches += 1;
  return cudaGetLastError();
}

// ================================ segment sort ======================================
namespace
{
constexpr int SORT_THREADS = 512;
constexpr uint32_t SORT_SMEM_IDS = 8192;

// Bitonic network with every comparison ascending (mirror partner on the first step of a
// merge), so indices >= len act as +infinity padding and never need to exist.
__device__ void bitonic_ascending(uint32_t* a, uint64_t len)
{
  uint64_t padded = 1;
  while (padded < len)
    padded <<= 1;
  for (uint64_t k = 2; k <= padded; k <<= 1)
  {
    for (uint64_t j = k >> 1; j > 0; j >>= 1)
    {
      for (uint64_t i = threadIdx.x; i < len; i += blockDim.x)
      {
        const uint64_t partner = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
        if (partner > i && partner < len)
        {
          const uint32_t x = a[i], y = a[partner];
          if (x > y)
          {
            a[i] = y;
            a[partner] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(SORT_THREADS) sort_segments_kernel(const uint32_t* __restrict__ big_list,
                                                                     const uint32_t* __restrict__ big_count,
                                                                     const unsigned long long* __restrict__ offsets,
                                                                     uint32_t* ids, uint32_t skip_long)
{
  __shared__ uint32_t buf[SORT_SMEM_IDS];
  const uint32_t segments = *big_count;
  for (uint32_t s = blockIdx.x; s < segments; s += gridDim.x)
  {
    const uint32_t q = big_list[s];
    const unsigned long long begin = offsets[q];
    const uint64_t len = offsets[q + 1] - begin;
    uint32_t* seg = ids + begin;
    if (len <= SORT_SMEM_IDS)
    {
      for (uint64_t i = threadIdx.x; i < len; i += blockDim.x)
        buf[i] = seg[i];
      __syncthreads();
      bitonic_ascending(buf, len);
      for (uint64_t i = threadIdx.x; i < len; i += blockDim.x)
        seg[i] = buf[i];
      __syncthreads();
    }
    else if (!skip_long)
    {
      __syncthreads();
      bitonic_ascending(seg, len); // in global memory; __syncthreads orders the block's accesses
      __syncthreads();
    }
  }
}

// ---- lists longer than SORT_SMEM_IDS: one bitonic network over ALL of them at once ------------
// LongTable = { count, work, then per long list: begin (u64), work_begin (u64), len (u32) } in scratch.
struct LongSeg
{
  unsigned long long begin;      // first id of the list in ids[]
  unsigned long long work_begin; // exclusive prefix sum of the lengths of the long lists before it
  unsigned long long len;
};
struct LongHeader
{
  unsigned int count;            // long lists found
  unsigned int pad;
  unsigned long long work;       // sum of their lengths
};

__global__ void long_init_kernel(LongHeader* h)
{
  h->count = 0;
  h->pad = 0;
  h->work = 0;
}

__global__ void __launch_bounds__(256) long_collect_kernel(const uint32_t* __restrict__ big_list,
                                                           const uint32_t* __restrict__ big_count,
                                                           const unsigned long long* __restrict__ offsets,
                                                           LongHeader* h, LongSeg* table, uint32_t capacity)
{
  const uint32_t segments = *big_count;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < segments; s += gridDim.x * blockDim.x)
  {
    const uint32_t q = big_list[s];
    const unsigned long long begin = offsets[q];
    const unsigned long long len = offsets[q + 1] - begin;
    if (len > SORT_SMEM_IDS)
    {
      const unsigned int at = atomicAdd(&h->count, 1u);
      if (at < capacity)
      {
        table[at].begin = begin;
        table[at].len = len;
      }
    }
  }
}

// one CTA: work_begin = exclusive scan of len over the table (a few thousand entries at most)
__global__ void __launch_bounds__(SCAN_THREADS) long_prefix_kernel(LongHeader* h, LongSeg* table, uint32_t capacity)
{
  __shared__ unsigned long long warp_sums[32];
  const unsigned int count = h->count < capacity ? h->count : capacity;
  unsigned long long carry = 0;
  for (unsigned int base = 0; base < count; base += SCAN_THREADS)
  {
    const unsigned int i = base + threadIdx.x;
    const unsigned long long v = i < count ? table[i].len : 0ull;
    unsigned long long total;
    const unsigned long long inc = block_inclusive(v, warp_sums, &total);
    if (i < count)
      table[i].work_begin = carry + inc - v;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0)
  {
    h->count = count;
    h->work = carry;
  }
}

// one step (k, j) of the network of bitonic_ascending() over every element of every long list
__global__ void __launch_bounds__(256) long_step_kernel(const LongHeader* __restrict__ h,
                                                        const LongSeg* __restrict__ table, uint32_t* ids,
                                                        unsigned long long k, unsigned long long j)
{
  const unsigned int count = h->count;
  const unsigned long long work = h->work;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < work;
       e += (unsigned long long)gridDim.x * blockDim.x)
  {
    // the list this element belongs to: the last one whose work_begin <= e
    unsigned int lo = 0, hi = count;
    while (hi - lo > 1)
    {
      const unsigned int mid = (lo + hi) >> 1;
      if (table[mid].work_begin <= e)
        lo = mid;
      else
        hi = mid;
    }
    const unsigned long long len = table[lo].len;
    const unsigned long long i = e - table[lo].work_begin;
    const unsigned long long partner = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
    if (partner > i && partner < len)
    {
      uint32_t* seg = ids + table[lo].begin;
      const uint32_t x = seg[i], y = seg[partner];
      if (x > y)
      {
        seg[i] = y;
        seg[partner] = x;
      }
    }
  }
}
} // namespace

// long lists hold more than SORT_SMEM_IDS ids each, so a batch of `total_ids` has at most this many
static uint32_t long_capacity(unsigned long long total_ids)
{
  return (uint32_t)(total_ids / SORT_SMEM_IDS + 1);
}

size_t long_sort_scratch_bytes(unsigned long long total_ids)
{
  return sizeof(LongHeader) + (size_t)long_capacity(total_ids) * sizeof(LongSeg) + 64;
}

cudaError_t launch_sort_segments(const uint32_t* big_list, const uint32_t* big_count_dev,
                                 uint32_t max_segments, const unsigned long long* offsets,
                                 uint32_t* ids, cudaStream_t stream, uint32_t* launches, uint32_t longest,
                                 void* long_scratch, unsigned long long total_ids)
{
  if (max_segments == 0)
    return cudaSuccess;
  const bool long_path = longest > SORT_SMEM_IDS && long_scratch != nullptr;
  const unsigned wave = sm_count() * 4u;
  const unsigned grid = max_segments < wave ? max_segments : wave;
  sort_segments_kernel<<<grid, SORT_THREADS, 0, stream>>>(big_list, big_count_dev, offsets, ids, long_path ? 1u : 0u);
  *launches += 1;
  if (!long_path)
    return cudaGetLastError();

  LongHeader* h = static_cast<LongHeader*>(long_scratch);
  LongSeg* table = reinterpret_cast<LongSeg*>(h + 1);
  const uint32_t capacity = long_capacity(total_ids);
  long_init_kernel<<<1, 1, 0, stream>>>(h);
  const unsigned cgrid = (max_segments + 255) / 256 < wave ? (max_segments + 255) / 256 : wave;
  long_collect_kernel<<<cgrid, 256, 0, stream>>>(big_list, big_count_dev, offsets, h, table, capacity);
  long_prefix_kernel<<<1, SCAN_THREADS, 0, stream>>>(h, table, capacity);
  *launches += 3;
  // the network of bitonic_ascending(): every comparison ascending, mirror partner on the first step
  // of a merge, indices beyond a list's length act as +infinity.  `longest` bounds every list, so a
  // list shorter than the padded size of the longest one simply sees no-op steps at the end.
  unsigned long long padded = 1;
  while (padded < longest)
    padded <<= 1;
  const unsigned sgrid = sm_count() * 8u;
  for (unsigned long long k = 2; k <= padded; k <<= 1)
  {
    for (unsigned long long j = k >> 1; j > 0; j >>= 1)
    {
      long_step_kernel<<<sgrid, 256, 0, stream>>>(h, table, ids, k, j);
      *launches += 1;
    }
  }
  return cudaGetLastError();
}

// ================================ query reordering ==================================
// Counting sort of the queries by the Morton cell of their box centre.  Which query a warp
// works on next then shares most of its root-to-leaf path with the previous one, so node
// blocks are found in L1/L2 instead of HBM.  Results are written by ORIGINAL query index,
// so the order inside a cell (decided by atomics) never shows in the output.
namespace
{
constexpr uint32_t CELL_BITS_PER_AXIS = 7;                       // 128 cells per axis
constexpr uint32_t CELL_COUNT = 1u << (3 * CELL_BITS_PER_AXIS);  // 2 M cells

// `hi` = offset of the second corner in the query record (0 for a point: the centre is the point)
template <typename S>
__device__ __forceinline__ float centre_of(const S* q, uint32_t hi, uint32_t axis)
{
  return 0.5f * ((float)q[axis] + (float)q[hi + axis]);
}

__device__ __forceinline__ uint32_t spread3(uint32_t v)
{
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// order-preserving float <-> uint mapping for atomicMin/Max
__device__ __forceinline__ uint32_t float_key(float f)
{
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k)
{
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

struct ReorderScratch
{
  uint32_t ext_min[3];
  uint32_t ext_max[3];
};

__global__ void reorder_init(ReorderScratch* s, uint32_t* cell_count)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3)
  {
    s->ext_min[i] = 0xFFFFFFFFu;
    s->ext_max[i] = 0u;
  }
  for (uint32_t c = i; c <= CELL_COUNT; c += gridDim.x * blockDim.x)
    cell_count[c] = 0;
}

template <typename S>
__global__ void __launch_bounds__(256) reorder_extent(const S* __restrict__ q, uint32_t dim, uint32_t hi_at,
                                                      uint32_t stride, uint64_t n, ReorderScratch* s)
{
  const uint32_t axes = dim < 3 ? dim : 3;
  float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
  {
    for (uint32_t a = 0; a < axes; ++a)
    {
      const float c = centre_of(q + i * stride, hi_at, a);
      if (c == c && fabsf(c) != INFINITY)
      {
        lo[a] = fminf(lo[a], c);
        hi[a] = fmaxf(hi[a], c);
      }
    }
  }
  for (uint32_t a = 0; a < axes; ++a)
  {
    for (int d = 16; d > 0; d >>= 1)
    {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], d));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], d));
    }
    if ((threadIdx.x & 31) == 0 && lo[a] <= hi[a])
    {
      atomicMin(&s->ext_min[a], float_key(lo[a]));
      atomicMax(&s->ext_max[a], float_key(hi[a]));
    }
  }
}

template <typename S>
__device__ __forceinline__ uint32_t cell_of(const S* q, uint32_t dim, uint32_t hi, const ReorderScratch* s)
{
  const uint32_t axes = dim < 3 ? dim : 3;
  uint32_t code = 0;
  for (uint32_t a = 0; a < axes; ++a)
  {
    const float lo = key_float(s->ext_min[a]), vhi = key_float(s->ext_max[a]);
    const float c = centre_of(q, hi, a);
    float t = (vhi > lo) ? (c - lo) / (vhi - lo) : 0.0f;
    t = (t == t) ? fminf(fmaxf(t, 0.0f), 1.0f) : 0.0f;
    const uint32_t cell = min((uint32_t)(t * (float)(1u << CELL_BITS_PER_AXIS)), (1u << CELL_BITS_PER_AXIS) - 1u);
    code |= spread3(cell) << a;
  }
  return code;
}

template <typename S>
__global__ void __launch_bounds__(256) reorder_histogram(const S* __restrict__ q, uint32_t dim, uint32_t hi,
                                                         uint32_t stride, uint64_t n, const ReorderScratch* s,
                                                         uint32_t* cell_count, uint32_t* cell_of_query)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
  {
    const uint32_t c = cell_of(q + i * stride, dim, hi, s);
    cell_of_query[i] = c;
    atomicAdd(&cell_count[c], 1u);
  }
}

__global__ void __launch_bounds__(256) reorder_scatter(const uint32_t* __restrict__ cell_of_query, uint64_t n,
                                                       const unsigned long long* __restrict__ cell_begin,
                                                       uint32_t* cell_fill, uint32_t* order)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
  {
    const uint32_t c = cell_of_query[i];
    const uint32_t k = atomicAdd(&cell_fill[c], 1u);
    order[cell_begin[c] + k] = (uint32_t)i;
  }
}

// the quantisation grid of a tree, by value
struct PackGrid
{
  float lo[3], scale[3];
};

// query i of the caller's array as a packed record: its box on the tree's grid + its index
template <int D>
__device__ __forceinline__ uint4 pack_query(const float* __restrict__ q, uint64_t i, const PackGrid& g)
{
  float lo[D], hi[D];
#pragma unroll
  for (int d = 0; d < D; ++d)
  {
    lo[d] = q[i * (2 * D) + d];
    hi[d] = q[i * (2 * D) + D + d];
  }
  uint32_t b[D];
  quant_query<D>(lo, hi, g.lo, g.scale, b);
  return D == 3 ? make_uint4(b[0], b[1], b[2], (uint32_t)i) : make_uint4(b[0], b[1], (uint32_t)i, 0u);
}

// the scatter of the coherence pass for packed launches: the query itself, quantised, goes to its
// place in the processing order (BoxLaunch::packed)
template <int D>
__global__ void __launch_bounds__(256) reorder_scatter_packed(const float* __restrict__ q,
                                                              const uint32_t* __restrict__ cell_of_query, uint64_t n,
                                                              const unsigned long long* __restrict__ cell_begin,
                                                              uint32_t* cell_fill, uint4* packed, PackGrid g)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
  {
    const uint32_t c = cell_of_query[i];
    const uint32_t k = atomicAdd(&cell_fill[c], 1u);
    packed[cell_begin[c] + k] = pack_query<D>(q, i, g);
  }
}

template <int D>
__global__ void __launch_bounds__(256) pack_queries_kernel(const float* __restrict__ q, uint64_t n, uint4* packed,
                                                           PackGrid g)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    packed[i] = pack_query<D>(q, i, g);
}

static PackGrid pack_grid(const TreeDev& tree)
{
  PackGrid g;
  for (int d = 0; d < 3; ++d)
  {
    g.lo[d] = tree.qlo[d];
    g.scale[d] = tree.qscale[d];
  }
  return g;
}

__global__ void zero_u32(uint32_t* p, uint32_t n)
{
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    p[i] = 0;
}
} // namespace

// scratch layout: ReorderScratch | cell_count[CELL_COUNT+1] u32 | cell_begin[CELL_COUNT+2] u64 |
//                 scan scratch | cell_of_query[n] u32
static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

size_t reorder_scratch_bytes(uint64_t n)
{
  return align256(sizeof(ReorderScratch)) + align256((CELL_COUNT + 1) * sizeof(uint32_t))
         + align256((CELL_COUNT + 2) * sizeof(unsigned long long)) + align256(scan_scratch_bytes(CELL_COUNT + 1))
         + align256(n * sizeof(uint32_t));
}

cudaError_t launch_pack_queries(const void* queries, uint32_t dim, uint64_t n, uint4* packed, const TreeDev& tree,
                                cudaStream_t stream, uint32_t* launches)
{
  if (n == 0)
    return cudaSuccess;
  const uint64_t want = (n + 255) / 256;
  const unsigned cap = sm_count() * 8u;
  const unsigned grid = want < cap ? (unsigned)want : cap;
  const float* q = static_cast<const float*>(queries);
  if (dim == 3)
    pack_queries_kernel<3><<<grid, 256, 0, stream>>>(q, n, packed, pack_grid(tree));
  else if (dim == 2)
    pack_queries_kernel<2><<<grid, 256, 0, stream>>>(q, n, packed, pack_grid(tree));
  else
    return cudaErrorInvalidValue;
  *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_reorder(const void* queries, uint32_t scalar, uint32_t dim, uint32_t scalars_per_query,
                           uint64_t n, uint32_t* order, void* scratch, cudaStream_t stream,
                           uint32_t* launches, uint32_t hi_offset, uint4* packed, const TreeDev* tree)
{
  if (n == 0)
    return cudaSuccess;
  if (packed != nullptr && (tree == nullptr || scalar != RT_F32 || (dim != 2 && dim != 3) || scalars_per_query != 2 * dim))
    return cudaErrorInvalidValue;
  const uint32_t hi = hi_offset == 0xFFFFFFFFu ? dim : hi_offset;
  unsigned char* base = static_cast<unsigned char*>(scratch);
  ReorderScratch* rs = reinterpret_cast<ReorderScratch*>(base);
  base += align256(sizeof(ReorderScratch));
  uint32_t* cell_count = reinterpret_cast<uint32_t*>(base);
  base += align256((CELL_COUNT + 1) * sizeof(uint32_t));
  unsigned long long* cell_begin = reinterpret_cast<unsigned long long*>(base);
  base += align256((CELL_COUNT + 2) * sizeof(unsigned long long));
  void* scan_scratch = base;
  base += align256(scan_scratch_bytes(CELL_COUNT + 1));
  uint32_t* cell_of_query = reinterpret_cast<uint32_t*>(base);

  const unsigned grid = sm_count() * 8u;
  reorder_init<<<grid, 256, 0, stream>>>(rs, cell_count);
  *launches += 1;
  switch (scalar)
  {
  case RT_F32:
    reorder_extent<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(queries), dim, hi, scalars_per_query, n, rs);
    reorder_histogram<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(queries), dim, hi, scalars_per_query, n,
                                                       rs, cell_count, cell_of_query);
    break;
  case RT_F64:
    reorder_extent<double><<<grid, 256, 0, stream>>>(static_cast<const double*>(queries), dim, hi, scalars_per_query, n, rs);
    reorder_histogram<double><<<grid, 256, 0, stream>>>(static_cast<const double*>(queries), dim, hi, scalars_per_query,
                                                        n, rs, cell_count, cell_of_query);
    break;
  case RT_I32:
    reorder_extent<int32_t><<<grid, 256, 0, stream>>>(static_cast<const int32_t*>(queries), dim, hi, scalars_per_query, n, rs);
    reorder_histogram<int32_t><<<grid, 256, 0, stream>>>(static_cast<const int32_t*>(queries), dim, hi, scalars_per_query,
                                                         n, rs, cell_count, cell_of_query);
    break;
  default: return cudaErrorInvalidValue;
  }
  *launches += 2;
  cudaError_t err = launch_exclusive_scan(cell_count, CELL_COUNT + 1, cell_begin, scan_scratch, stream, launches);
  if (err != cudaSuccess)
    return err;
  zero_u32<<<grid, 256, 0, stream>>>(cell_count, CELL_COUNT + 1);
  if (packed == nullptr)
    reorder_scatter<<<grid, 256, 0, stream>>>(cell_of_query, n, cell_begin, cell_count, order);
  else if (dim == 3)
    reorder_scatter_packed<3><<<grid, 256, 0, stream>>>(static_cast<const float*>(queries), cell_of_query, n, cell_begin,
                                                        cell_count, packed, pack_grid(*tree));
  else
    reorder_scatter_packed<2><<<grid, 256, 0, stream>>>(static_cast<const float*>(queries), cell_of_query, n, cell_begin,
                                                        cell_count, packed, pack_grid(*tree));
  *launches += 2;
  return cudaGetLastError();
}

} // namespace rtb
