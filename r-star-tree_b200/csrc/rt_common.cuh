// rt_common.cuh -- shared declarations of the sm_100a query engine (internal header).
//
// Division of labour inside librtree_b200.so:
//   rt_api*.cu         the C ABI over rt_handle.cuh (handle, buffers, streams / events): upload / free /
//                      options (rt_api.cu), the box pipeline (rt_api_boxes.cu), rays, kNN + refit
//   rt_box_*.cu        instantiations of the box traversal kernel (rt_traverse_box.cuh)
//   rt_ray.cu, rt_knn.cu   the ray / kNN traversal kernels (rt_traverse_ray.cuh, rt_traverse_knn.cuh)
//   rt_quant.cu        builds the quantised image; rt_refit.cu refits bounds after keys moved
//   rt_passes.cu       hand-written scan, gather, segment sort and query reordering / packing
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include <rtree_b200.h>

namespace rtb
{

// The uploaded tree as kernels see it (layout: rt_tree_desc in include/rtree_b200.h).
// Nodes are named by the byte offset of their block divided by 16 ("link"): the root is
// link 0 and a link at or above leaf_link_begin names a leaf block.
struct TreeDev
{
  const unsigned char* blocks; // internal blocks, then (at leaf_offset) leaf blocks
  uint32_t leaf_link_begin;    // leaf_offset / 16
  uint32_t null_link;          // a block past the leaves whose slots are all empty (appended by rt_upload)
  // Quantised image (rt_quant.cuh; float32 point-keyed trees of the instantiated (dim, width) pairs,
  // else NULL): one block of W records of dim + 1 words per node -- internal nodes first, then from
  // qleaf_link_begin the leaves in the order of `blocks` -- holding conservative 15-bit bounds + the link.
  const unsigned char* qblocks;
  const unsigned char* leaf_blocks; // blocks + leaf_offset: the exact records behind the leaf slots
  uint32_t qleaf_link_begin;
  uint32_t qnull_link;
  float qlo[8], qscale[8];          // the grid: q(v) = clamp(floor((v - qlo) * qscale) + 1, 1, 32767)
  uint32_t stage_links;             // links below this name the nodes of levels 0..2 (RTB_STAGE_TOP experiment)
  uint32_t num_nodes;
  uint32_t num_levels;
};

// traversal passes
enum : uint32_t
{
  PASS_COUNT = 0,       // counts[q] only
  PASS_COUNT_STATS = 1, // counts[q] + node-visit totals (RT_COLLECT_VISITS)
  PASS_STASH = 2,       // counts[q]; sorted hit lists of <= HIT_BUF hits go to the stash
  PASS_WRITE = 3        // hits go to ids[offsets[q] ...), small lists sorted in the warp
};

constexpr uint32_t HIT_BUF = 32;        // per-warp hit buffer (one id per lane)
constexpr uint32_t STASH_SLAB = 512;    // ids a warp reserves from the stash per atomic (smallest slab)
constexpr uint32_t STASH_SLAB_MAX = 16384;
constexpr uint32_t NO_STASH = 0xFFFFFFFFu;
constexpr int TRAVERSE_THREADS = 256;   // 8 warps per CTA
constexpr uint32_t QUERY_CHUNK = 4;     // queries a warp claims per atomic (default)
constexpr uint32_t MAX_WORK_RANGES = 1u << 16;

// Work distribution of the persistent traversal kernels.  The n queries, in processing order
// (Morton order after the coherence pass), are cut into `ranges` contiguous ranges of `per_range`
// queries.  By default there is one range per SM: every warp resident on SM s starts on range s
// and claims `chunk` consecutive queries at a time from that range's cursor, so the 64 warps of an
// SM walk one short window of neighbouring queries and find each other's nodes in L1.  A warp
// whose range is used up steals chunks from the following ranges.  cursors[] is zeroed by the
// launcher.
struct WorkRanges
{
  unsigned int* cursors; // [ranges], queries handed out per range
  uint32_t ranges;
  uint32_t per_range;    // a multiple of chunk
  uint32_t chunk;
  uint32_t by_sm;        // start range = %smid % ranges (else blockIdx.x % ranges)
};

// dynamic shared memory of a traversal CTA: per warp, the hit buffer, G = 32/W words holding the
// null node and the stack (at most 32 entries per level are ever live)
inline size_t traverse_smem_bytes(uint32_t num_levels, uint32_t width)
{
  return (size_t)(TRAVERSE_THREADS / 32) * (32u / width + 32u * num_levels + HIT_BUF) * sizeof(uint32_t);
}

struct BoxLaunch
{
  TreeDev tree;
  const void* queries;      // [n][2][dim] scalars (device)
  // Quantised trees of dim <= 3 ("packed" launches): the queries as the coherence pass (or
  // launch_pack_queries) left them -- in processing order, already on the tree's grid, 16 bytes each:
  // { b_0 .. b_{dim-1}, original index } (dim 2: one unused word).  A query is then one 128-bit load and
  // the float box is only looked at when a leaf candidate sits in a boundary cell of the query.
  const uint4* packed;
  // nullable: i-th processed query is order[i] -- an original query index, or for packed launches a
  // position in `packed` (only the deferred second pass of a packed launch has one)
  const uint32_t* order;
  uint64_t n;               // number of queries to process (length of order if given)
  uint32_t pass;
  // outputs / inputs by pass
  uint32_t* counts;                 // [n_total] hits per query (COUNT*, STASH)
  const unsigned long long* offsets; // [n_total+1] (WRITE)
  uint32_t* ids;                    // [total]     (WRITE)
  uint32_t* stash;                  // stash pool  (STASH)
  uint32_t* stash_pos;              // [n_total] position in the pool or NO_STASH (STASH)
  unsigned long long stash_capacity;
  uint32_t stash_slab;              // ids per slab: a list of up to this many ids can be stashed
  unsigned long long* stash_cursor; // device counter
  uint32_t* deferred_list;          // queries that need PASS_WRITE (STASH)
  uint32_t* deferred_count;
  uint32_t* big_list;               // queries whose list must be sorted afterwards (STASH spills, WRITE)
  uint32_t* big_count;
  uint32_t* max_list;               // longest list that will go to the segment sort (device counter, atomicMax)
  WorkRanges work;                  // cursors = MAX_WORK_RANGES entries; the rest is set by the launcher
  unsigned long long* visit_stats;  // [0] internal / [1] leaf visits, [3] loop steps (COUNT_STATS)
  unsigned long long* confirm_stats; // [0] leaf candidates looked up on the float records (COUNT_STATS)
  uint32_t sorted;                  // 0: keep traversal order
  uint32_t stop_at_first;           // PASS_COUNT: abort a query at its first hit (RT_ANY_HIT): counts are 0 / 1
  uint32_t strict;                  // RT_STRICT_OVERLAP: the reference's strict is_overlap as the filter (float blocks)
  int grid;                         // 0: fill the GPU; > 0: exactly; < 0: fill, less -grid CTAs per SM
  cudaStream_t stream;
};

typedef cudaError_t (*box_launch_fn)(const BoxLaunch&);

// rt_box_*.cu: nullptr when the combination is not instantiated
box_launch_fn find_box_kernel(uint32_t scalar, uint32_t dim, uint32_t width, uint32_t key_kind,
                              uint32_t mode);
box_launch_fn find_box_kernel_f32(uint32_t dim, uint32_t width, uint32_t key_kind, uint32_t mode);
box_launch_fn find_box_kernel_f64(uint32_t dim, uint32_t width, uint32_t key_kind, uint32_t mode);
box_launch_fn find_box_kernel_i32(uint32_t dim, uint32_t width, uint32_t key_kind, uint32_t mode);

struct RayLaunch
{
  TreeDev tree;
  uint32_t width;
  uint32_t key_kind;
  const float* rays;     // [n][7]
  const uint32_t* order;
  uint64_t n;
  uint32_t mode;         // RT_RAY_*
  uint32_t pass;         // ALL_HITS: PASS_COUNT / PASS_COUNT_STATS / PASS_WRITE; closest / any: PASS_COUNT_STATS
                         // also counts the node visits, anything else does not
  uint32_t* counts;
  const unsigned long long* offsets;
  uint32_t* ids;
  uint32_t* big_list;
  uint32_t* big_count;
  uint32_t* max_list; // longest list that will go to the segment sort
  float* t_out;
  uint32_t* id_out;
  uint8_t* any_out;
  WorkRanges work;
  unsigned long long* visit_stats;
  uint32_t sorted;
  int grid;
  cudaStream_t stream;
};
cudaError_t launch_ray_kernel(const RayLaunch&); // rt_ray.cu; cudaErrorInvalidValue if unsupported

struct KnnLaunch
{
  TreeDev tree;
  const float* points;   // [n][dim]
  const uint32_t* order; // nullable: i-th processed query is order[i]
  uint64_t n;
  uint32_t k;            // 1 .. 32
  uint32_t collect;      // also count node visits
  uint32_t* ids_out;     // [n][k]
  float* d2_out;         // [n][k]
  WorkRanges work;
  unsigned long long* visit_stats;
  int grid;
  cudaStream_t stream;
};
bool knn_supported(uint32_t scalar, uint32_t dim, uint32_t width); // rt_knn.cu
cudaError_t launch_knn_kernel(const KnnLaunch&, uint32_t dim, uint32_t width, uint32_t key_kind);

// grid of a persistent traversal launch (occupancy x SMs, clamped to the work) and the split of
// the n queries into ranges; zeroes the cursors on `stream`.  On entry work.chunk / work.ranges
// hold the caller's wishes (0 = default: QUERY_CHUNK, one range per SM).
cudaError_t plan_work(const void* kernel, size_t smem, uint64_t n, int forced_grid, cudaStream_t stream,
                      WorkRanges& work, int& grid);

// rt_quant.cu: the quantised image of a float32 point-keyed tree.  quant_supported(): is there a
// traversal kernel for it; `qblocks` must hold quant_image_bytes() bytes; build_quant_image() fills
// tree.qblocks / leaf_blocks / qleaf_link_begin / qnull_link / qlo / qscale (synchronises the stream
// once to read the grid back).
bool quant_supported(uint32_t scalar, uint32_t dim, uint32_t width, uint32_t key_kind);
size_t quant_image_bytes(uint32_t num_nodes, uint32_t dim, uint32_t width);
cudaError_t build_quant_image(TreeDev& tree, const rt_tree_desc& desc, uint32_t num_internal, unsigned char* qblocks,
                              cudaStream_t stream);

// rt_refit.cu: new keys (flatten data order) into the leaf slots, then all bounds bottom-up
size_t refit_scratch_bytes(uint32_t n_leaves);
cudaError_t launch_refit(unsigned char* blocks, const rt_tree_desc& d, const uint32_t* levels, const void* keys,
                         void* scratch, cudaStream_t stream, uint32_t* launches);

// ---- rt_passes.cu -----------------------------------------------------------------
// offsets[0..n] = exclusive prefix sum of counts[0..n) (64-bit); scratch >= scan_scratch_bytes(n)
size_t scan_scratch_bytes(uint64_t n);
cudaError_t launch_exclusive_scan(const uint32_t* counts, uint64_t n, unsigned long long* offsets,
                                  void* scratch, cudaStream_t stream, uint32_t* launches);
// the batch's counter block to mapped pinned host memory by a store from the SM (not through the copy
// engine, whose queue may hold 100 MB of an earlier batch's result); the 8-byte word at
// total_offset_bytes is taken from *d_total (offsets[n]) when d_total is not NULL
cudaError_t launch_publish_counters(const void* d_counters, uint32_t bytes, uint32_t total_offset_bytes,
                                    const unsigned long long* d_total, void* h_mapped, cudaStream_t stream,
                                    uint32_t* launches);
// offsets[0..n) += base (rebases the CSR offsets of one batch of a pipelined call)
cudaError_t launch_add_base(unsigned long long* offsets, uint64_t n, unsigned long long base,
                            cudaStream_t stream, uint32_t* launches);
// ids[offsets[q] + k] = stash[stash_pos[q] + k] for every stashed query
cudaError_t launch_gather_stash(const uint32_t* stash, const uint32_t* stash_pos,
                                const unsigned long long* offsets, uint64_t n, uint32_t* ids,
                                cudaStream_t stream, uint32_t* launches);
// sort ids[offsets[q] .. offsets[q+1]) ascending for the queries in big_list.  Lists of up to 8192 ids
// are sorted by one CTA each in shared memory.  `longest` = an upper bound of the longest list (the
// counter the traversal keeps); when it exceeds 8192 the longer lists are sorted TOGETHER by a bitonic
// network in global memory, one launch per network step over all their elements, so that one huge list
// (a query covering much of the tree) gets the whole GPU instead of a single CTA.  `long_scratch` must
// then hold long_sort_scratch_bytes(total ids of the batch) bytes.
size_t long_sort_scratch_bytes(unsigned long long total_ids);
cudaError_t launch_sort_segments(const uint32_t* big_list, const uint32_t* big_count_dev,
                                 uint32_t max_segments, const unsigned long long* offsets,
                                 uint32_t* ids, cudaStream_t stream, uint32_t* launches,
                                 uint32_t longest = 0, void* long_scratch = nullptr, unsigned long long total_ids = 0);
// coherence pass: order[] = query indices grouped by the Morton cell of the box centre
// (first 3 axes), cells sized from the batch's own extent; scratch >= reorder_scratch_bytes()
size_t reorder_scratch_bytes(uint64_t n);
// hi_offset: where the second corner of a query starts (dim for boxes; 0 for points, whose centre is
// the point itself)
// packed != nullptr (float32 boxes, dim 2 or 3): instead of order[], write the queries themselves in
// processing order, quantised on `tree`'s grid (BoxLaunch::packed)
cudaError_t launch_reorder(const void* queries, uint32_t scalar, uint32_t dim, uint32_t floats_per_query,
                           uint64_t n, uint32_t* order, void* scratch, cudaStream_t stream,
                           uint32_t* launches, uint32_t hi_offset = 0xFFFFFFFFu, uint4* packed = nullptr,
                           const TreeDev* tree = nullptr);
// packed[i] = query i on the tree's grid, in input order (RT_NO_REORDER and small batches)
cudaError_t launch_pack_queries(const void* queries, uint32_t dim, uint64_t n, uint4* packed, const TreeDev& tree,
                                cudaStream_t stream, uint32_t* launches);

} // namespace rtb
