/* rtree_b200.h -- the C ABI of the B200 batched query engine (librtree_b200.so).
 *
 * This is the drop-in boundary for the read-only query path of ehwan/R-Star-Tree.
 * The reference has no FFI layer -- its "GPU support" is RTree::flatten()
 * (reference include/RTree/rtree.hpp:966-981) returning host vectors the user is told
 * to upload themselves (reference README.md:265-274).  These entry points are what a
 * binding for that path attaches to:
 *
 *   rt_upload        <- consumes what RTree::flatten_device() emits, i.e. the
 *                       flatten_result_t of reference rtree.hpp:852-870 re-laid out
 *   rt_query_boxes   <- replaces a host loop of RTree::search(filter, functor)
 *                       (reference rtree.hpp:1141-1146, hot loop :984-1019) with the
 *                       canonical closed-overlap filter (reference
 *                       example/sample/sample.cpp:48-54, test/rtree.cpp:403) and the
 *                       leaf test geometry_traits<aabb>::is_inside (reference
 *                       include/RTree/aabb.hpp:206-210; test/rtree.cpp:390)
 *   rt_query_rays    <- the same search() driven by a stateful slab-test filter
 *                       (legal per reference rtree.hpp:1145); all-hits, closest-hit,
 *                       any-hit (functor returning true, reference rtree.hpp:994-997)
 *   rt_query_knn     <- the same search() with a distance-pruning filter and a k-best functor
 *   rt_free          <- releases the device copy
 *
 * Plain C, plain pointers and sizes; no exceptions cross the boundary.  Every call
 * returns RT_OK (0) or a negative RT_E_* code; rt_last_error() gives the text.
 * There is NO CPU fallback: without a CUDA device every call fails with RT_E_CUDA.
 *
 * Threading: calls on one handle are serialised by the caller; different handles
 * (different GPUs) may be driven from different host threads.
 */
#ifndef RTREE_B200_H
#define RTREE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_ABI_VERSION 2u

/* error codes */
#define RT_OK 0
#define RT_E_INVALID (-1)     /* bad argument / descriptor */
#define RT_E_UNSUPPORTED (-2) /* (scalar, dim, width, key kind, mode) not instantiated */
#define RT_E_CUDA (-3)        /* CUDA runtime error, or no device */
#define RT_E_NOMEM (-4)       /* host or device allocation failed */
#define RT_E_INTERNAL (-5)

/* rt_tree_desc.scalar */
#define RT_F32 0u
#define RT_F64 1u
#define RT_I32 2u

/* rt_tree_desc.key_kind */
#define RT_KEY_POINT 0u /* leaf records hold one corner: p_0 .. p_{D-1} id             */
#define RT_KEY_BOX 1u   /* leaf records look like internal ones: lo/hi pairs, id      */

/* memory spaces */
#define RT_MEM_HOST 0u
#define RT_MEM_DEVICE 1u

/* rt_query_boxes mode = the leaf predicate (the node filter is always the closed
 * overlap  !(hi < qmin || lo > qmax)  per axis)                                     */
#define RT_POINT_IN_BOX 0u /* point keys:  !(p < qmin || p > qmax) per axis           */
#define RT_INTERSECTS 1u   /* box keys:    closed overlap with the key box            */
#define RT_CONTAINS 2u     /* box keys:    !(klo < qmin || khi > qmax) per axis       */

/* rt_query_rays mode */
#define RT_RAY_ALL_HITS 0u /* CSR of key ids whose AABB the ray hits                  */
#define RT_RAY_CLOSEST 1u  /* (t, id): least entry distance, ties to the smaller id   */
#define RT_RAY_ANY 2u      /* one byte per ray                                        */

/* flags (OR-ed) */
#define RT_QUERIES_ON_DEVICE 0x1u /* the query pointer is device memory on the tree's GPU */
#define RT_RESULT_ON_DEVICE 0x2u  /* leave the result in device memory (no D2H copy)      */
#define RT_UNSORTED 0x4u          /* per-query ids in traversal order (still deterministic)*/
#define RT_COUNT_ONLY 0x8u        /* fill offsets / total only, no ids                    */
#define RT_COLLECT_VISITS 0x10u   /* also count node visits (rt_stats), slower; counted on the
                                     image the call traverses (see RT_OPT_QUANTISED)         */
#define RT_NO_REORDER 0x20u       /* process queries in input order (no coherence pass)   */
#define RT_ANY_HIT 0x40u          /* rt_query_boxes: stop each query at its first hit -- the
                                     functor-returns-true abort of reference rtree.hpp:994-997.
                                     Implies RT_COUNT_ONLY; offsets[q+1] - offsets[q] is 0 or 1
                                     ("does the box hit anything"; which hit came first depends
                                     on the traversal order and is not reported)            */

#define RT_RESULT_COUNTS32 0x80u   /* rt_query_boxes with a HOST result: deliver rt_hits.counts (32-bit hits per
                                     query) instead of rt_hits.offsets (64-bit prefix sums, which the ids make
                                     redundant): 4 instead of 8 bytes per query cross PCIe.  rt_hits.offsets
                                     is then NULL; ids are laid out exactly as without the flag            */

#define RT_STRICT_OVERLAP 0x100u   /* rt_query_boxes: filter with the reference's N-D geometry_traits<aabb>::
                                     is_overlap(aabb, aabb) (include/RTree/aabb.hpp:225-236), whose inequalities
                                     are STRICT -- boxes that only touch do not overlap -- instead of the closed
                                     overlap of the reference's sample / test / README.  Applies to the node
                                     filter and to box keys under RT_INTERSECTS; point keys keep
                                     is_overlap(aabb, point) = is_inside (closed), RT_CONTAINS its containment
                                     test.  As in the reference this filter is NOT conservative for keys on a
                                     face of the query box: such a key is found only if the bounds above it
                                     overlap the query strictly, i.e. the hit set depends on the tree -- the
                                     engine returns exactly what the reference's search() returns on the same
                                     tree with these functors.  1-D trees: no effect (their is_overlap is closed,
                                     aabb.hpp:147-159).  Runs on the float blocks.                          */

/* An id slot that holds no entry.  Ids / value counts must stay below it.            */
#define RT_INVALID_ID 0xFFFFFFFFu

typedef struct rt_tree rt_tree; /* opaque; owns device memory on ONE GPU */

/* What RTree::flatten_device() fills (r-star-tree_b200/include/RTree/device_layout.hpp).
 *
 * Nodes are renumbered in breadth-first (level) order: node 0 is the root, level l
 * occupies [level_node_begin[l], level_node_begin[l+1]), leaves are the last level.
 * `blocks` holds one fixed-width block per node, W = width slots each; internal blocks
 * first (node i at i * internal_stride), then from leaf_offset the leaf blocks (j-th leaf
 * at leaf_offset + j * leaf_stride).
 *
 * A slot's RECORD is a run of 32-bit words (a 64-bit scalar takes two):
 *     internal node, or leaf of a box-keyed tree :  lo_0 .. lo_{D-1}  link  hi_0 .. hi_{D-1}
 *     leaf of a point-keyed tree                 :  p_0 .. p_{D-1}  link
 *   (with 64-bit scalars one unused word follows the link of a two-corner record, so that
 *   the hi corner stays 8-byte aligned).  The one-corner record is a prefix of the
 *   two-corner one: a lane reads "lo corner + link" the same way in both, and only
 *   internal nodes fetch the hi corner.
 *   link = for a child node: byte offset of the child's block in `blocks`, divided by 16
 *          (so the root is 0 and "is a leaf" is  link >= leaf_offset / 16);
 *          for a leaf entry: the 32-bit id reported for a hit;
 *          RT_INVALID_ID for an empty slot (whose bound is an inverted box).
 *
 * A block stores the W records as a structure of arrays OF 16-BYTE VECTORS, so that lane c
 * of a warp fetches "its" child with one coalesced 128-bit load per row:
 *     row k (k < words/4)   W x 16 bytes at k*W*16: words 4k..4k+3 of slot 0, of slot 1, ...
 *     tail (words % 4 = r)  at (words/4)*W*16: W elements of 4 (r=1), 8 (r=2) or 16 (r=3,
 *                           last word unused) bytes holding the remaining words of each slot
 * see rt_record_words() / rt_block_bytes() below.  3-D float32, W = 8: an internal block is
 * two 128-byte rows (lo_x lo_y lo_z link | hi_x hi_y hi_z -, 256 B), a point leaf exactly
 * one (x y z id per slot, 128 B).                                                         */
typedef struct rt_tree_desc
{
  uint32_t abi_version; /* RT_ABI_VERSION */
  uint32_t dim;         /* 1..8 */
  uint32_t scalar;      /* RT_F32 | RT_F64 | RT_I32 */
  uint32_t width;       /* W: MAX_ENTRIES rounded up to 8 or 16 */
  uint32_t max_entries; /* Config::MAX_ENTRIES of the tree that was flattened */
  uint32_t key_kind;    /* RT_KEY_POINT | RT_KEY_BOX */
  uint32_t leaf_level;  /* flatten_result_t::leaf_level */
  uint32_t num_levels;  /* leaf_level + 1 */
  uint32_t num_nodes;
  uint32_t num_values;  /* number of leaf entries */
  uint32_t internal_stride; /* == rt_block_bytes(width, rt_record_words(dim, scalar, 1)) */
  uint32_t leaf_stride;     /* == rt_block_bytes(width, rt_record_words(dim, scalar, key is box)) */
  uint64_t leaf_offset;     /* byte offset of the first leaf block, a multiple of 128 */
  uint64_t blocks_bytes;
  const void* blocks;
  const uint32_t* level_node_begin; /* [num_levels + 1], host memory */
  uint32_t blocks_memory;           /* RT_MEM_HOST | RT_MEM_DEVICE (same GPU) */
  uint32_t reserved;
} rt_tree_desc;

/* 32-bit words in a slot record: corners + link (+ alignment word, see above) */
static inline uint32_t rt_record_words(uint32_t dim, uint32_t scalar, int two_corners)
{
  const uint32_t sw = scalar == RT_F64 ? 2u : 1u;
  return two_corners ? 2u * dim * sw + sw : dim * sw + 1u;
}
/* bytes of one block of `width` slots with `words`-word records */
static inline uint32_t rt_block_bytes(uint32_t width, uint32_t words)
{
  static const uint32_t tail_bytes[4] = { 0u, 4u, 8u, 16u };
  return width * (16u * (words / 4u) + tail_bytes[words % 4u]);
}

/* CSR hit lists.  The buffers belong to the tree handle and stay valid until the next
 * query on that handle or rt_free().  ids[offsets[q] .. offsets[q+1]) are the hits of
 * query q, ascending unless RT_UNSORTED.                                             */
typedef struct rt_hits
{
  uint64_t n_queries;
  uint64_t total;
  const uint64_t* offsets; /* [n_queries + 1]; NULL with RT_RESULT_COUNTS32 */
  const uint32_t* ids;     /* [total]; NULL with RT_COUNT_ONLY */
  uint32_t memory;         /* RT_MEM_HOST (pinned) | RT_MEM_DEVICE */
  uint32_t reserved;
  const uint32_t* counts;  /* [n_queries] hits per query: with RT_RESULT_COUNTS32, and always for a
                              single-batch device result; else NULL */
} rt_hits;

typedef struct rt_ray_result
{
  uint64_t n_rays;
  uint64_t n_hit;   /* rays with at least one hit (closest / any), total hits (all) */
  rt_hits hits;     /* RT_RAY_ALL_HITS */
  const float* t;   /* RT_RAY_CLOSEST: entry distance, +inf on a miss */
  const uint32_t* id; /* RT_RAY_CLOSEST: key id, RT_INVALID_ID on a miss */
  const uint8_t* any; /* RT_RAY_ANY */
  uint32_t memory;
  uint32_t reserved;
} rt_ray_result;

/* Timings of the LAST query call on the handle (CUDA events on the handle's stream)
 * and, with RT_COLLECT_VISITS, the node-visit totals that define the algorithmic
 * bytes of the batch.  total_ms is the whole call; the per-phase times are sums over the
 * batches of the call, which overlap one another when the call was pipelined.          */
typedef struct rt_stats
{
  float total_ms;    /* first H2D to last D2H */
  float h2d_ms;
  float reorder_ms;  /* coherence pass over the queries */
  float traverse_ms; /* the traversal kernel(s) */
  float finish_ms;   /* scan + compaction + deferred queries */
  float d2h_ms;
  uint32_t kernel_launches; /* kernels of this library launched by the call */
  uint32_t deferred_queries; /* queries answered by the second traversal */
  uint64_t visits_internal;
  uint64_t visits_leaf;
  uint64_t algorithmic_bytes; /* SURVEY.md 8d formula on the visit totals */
  uint64_t traverse_steps;    /* warp loop iterations; visits / (steps * 32/W) = lane-group use */
  uint32_t batches;           /* batches the call was cut into (> 1: pipelined, see RT_OPT_PIPELINE_BATCH) */
  uint32_t reserved;
  uint64_t leaf_lookups;      /* RT_COLLECT_VISITS on a quantised image: leaf candidates that had to be looked
                                 up on the float records (those in a boundary cell of their query) */
} rt_stats;

/* Copy the tree image to `device` (cudaSetDevice(device) is called by every entry
 * point).  *out receives the handle.                                                  */
int rt_upload(const rt_tree_desc* desc, int device, rt_tree** out);

/* boxes: [n][2][dim] scalars of the tree's type, min corner then max corner.          */
int rt_query_boxes(rt_tree* tree, const void* boxes, uint64_t n, uint32_t mode,
                   uint32_t flags, rt_hits* out);

/* rays: [n][7] floats = origin[3], 1/direction[3] (computed by the caller so that
 * every consumer sees the same bits), tmax; tmin is 0.  3-D f32 trees only.            */
int rt_query_rays(rt_tree* tree, const float* rays, uint64_t n, uint32_t mode,
                  uint32_t flags, rt_ray_result* out);

/* k nearest neighbours: RTree::search() driven by a pruning filter and a k-best
 * functor, defined in oracle/knn.h.  ids[q*k .. q*k+k) / dist2[...] are the k keys nearest to point
 * q, ascending by (dist2, id); dist2 is the squared Euclidean distance to the key (to the key's box
 * for box keys), float32; when the tree holds fewer than k keys the tail is (RT_INVALID_ID, +inf).  */
typedef struct rt_knn_result
{
  uint64_t n_queries;
  uint32_t k;
  uint32_t memory;       /* RT_MEM_HOST (pinned) | RT_MEM_DEVICE */
  const uint32_t* ids;   /* [n_queries][k] */
  const float* dist2;    /* [n_queries][k] */
} rt_knn_result;

/* points: [n][dim] float32.  float32 trees with dim 2 or 3 and 8-wide blocks; 1 <= k <= 32.
 * Flags: RT_QUERIES_ON_DEVICE, RT_RESULT_ON_DEVICE, RT_NO_REORDER, RT_COLLECT_VISITS.      */
int rt_query_knn(rt_tree* tree, const float* points, uint64_t n, uint32_t k, uint32_t flags,
                 rt_knn_result* out);

/* Moving objects: replace every key and refit every bound on the device, keeping
 * the topology -- the batch form of changing keys through the iterators and calling
 * RTree::rebound(iterator) for each (reference README.md:342-347, include/RTree/rtree.hpp:479-500).
 * keys: [n_keys] keys of the tree's kind ([dim] scalars for points, [2][dim] for boxes) in
 * flatten_result_t::data order, i.e. the iteration order of RTree::begin()..end(); n_keys must be
 * the tree's num_values.  Afterwards the device image equals flatten_device() of the refitted host
 * tree byte for byte.  Like rebound() it does not rebalance: query cost degrades as keys drift.
 * Flags: RT_QUERIES_ON_DEVICE (keys is device memory).                                       */
int rt_refit(rt_tree* tree, const void* keys, uint64_t n_keys, uint32_t flags);

/* Release the handle and every buffer it handed out.  NULL is a no-op.                */
void rt_free(rt_tree* tree);

/* ---- auxiliaries -------------------------------------------------------------- */
/* Launch on this CUDA stream (a cudaStream_t) instead of the handle's own.          */
int rt_set_stream(rt_tree* tree, void* cuda_stream);
int rt_get_stats(const rt_tree* tree, rt_stats* out);
/* Tuning knobs.
 * RT_OPT_PIPELINE_BATCH: a query call with a HOST result and at least two batches' worth of
 *   queries is cut into batches of about this many queries (long calls: half-size batches at
 *   both ends, double-size ones in the middle); the copies of neighbouring batches (H2D of the
 *   next, D2H of the previous) run under the traversal of the current one.  The result is the
 *   same CSR as an unpipelined call.  Ray calls use batches of twice this size.  0 = never
 *   pipeline.  Default 1 Mi
 *   (environment override at rt_upload: RTB_PIPELINE_BATCH).                              */
#define RT_OPT_PIPELINE_BATCH 1u
/* RT_OPT_QUERY_CHUNK: queries a warp claims at a time (default 4).
 * RT_OPT_WORK_RANGES: 0 (default) = the queries of a launch are cut into one contiguous range per
 *   SM, the warps of an SM working through "their" range first and stealing from the others
 *   afterwards; N > 0 = N ranges, CTA b starting on range b mod N (1 = a single global cursor).
 *   Neither changes any result.                                                            */
#define RT_OPT_QUERY_CHUNK 2u
#define RT_OPT_WORK_RANGES 3u
/* RT_OPT_TRAVERSE_GRID: CTAs of the persistent traversal kernels; 0 (default) = as many as are
 *   resident at once (SMs x occupancy; one CTA slot per SM is left free in pipelined calls).    */
#define RT_OPT_TRAVERSE_GRID 4u
/* RT_OPT_QUANTISED: 1 (default) = box queries on float32 point-keyed trees (2-D / 3-D with 8-wide
 *   blocks, 8-D with 16-wide blocks) traverse the quantised image rt_upload builds for them (slot
 *   records with conservative 15-bit bounds; a leaf candidate strictly inside the query on the grid is a
 *   hit, one in a boundary cell is confirmed on its float record, so results are identical);
 *   0 = always the float blocks (the reference's exact filter: what SURVEY.md 8d's visit counts
 *   are defined on).  Keys with NaN coordinates are outside the contract on either image.        */
#define RT_OPT_QUANTISED 5u
int rt_set_option(rt_tree* tree, uint32_t option, uint64_t value);
/* How rt_query_boxes would cut a host-result call of n queries with RT_OPT_PIPELINE_BATCH =
 * pipeline_batch: writes at most `capacity` batch sizes to sizes[] (NULL: none) and returns the
 * number of batches, or RT_E_INVALID.  Pure host arithmetic (no device needed); values below the
 * minimum of 4096 are taken as 4096 here, as the RTB_PIPELINE_BATCH environment override is.     */
int64_t rt_plan_batches(uint64_t n, uint64_t pipeline_batch, uint64_t* sizes, uint64_t capacity);
/* Device copy of the blocks, e.g. to broadcast the tree to peer GPUs.               */
int rt_tree_device_blocks(const rt_tree* tree, void** device_ptr, uint64_t* bytes);
int rt_tree_get_desc(const rt_tree* tree, rt_tree_desc* out);
const char* rt_last_error(void);
uint32_t rt_abi_version(void);
/* Number of CUDA devices visible, or RT_E_CUDA.                                      */
int rt_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RTREE_B200_H */
