/* oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See oracle.c. */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_F32 = 0, ORACLE_F64 = 1, ORACLE_I32 = 2 };
enum { ORACLE_MODE_POINT_IN_BOX = 0, ORACLE_MODE_INTERSECTS = 1, ORACLE_MODE_CONTAINS = 2 };
/* OR-ed into the mode: filter with the reference's own is_overlap(aabb, aabb) (strict for N-D,
 * include/RTree/aabb.hpp:225-236) instead of the closed overlap; see RT_STRICT_OVERLAP */
enum { ORACLE_MODE_STRICT = 0x100 };
enum { ORACLE_RAY_ALL = 0, ORACLE_RAY_CLOSEST = 1, ORACLE_RAY_ANY = 2 };

/* The four vectors of flatten_result_t (include/RTree/rtree.hpp:852-870) as arrays. */
typedef struct oracle_flat
{
  uint32_t leaf_level;
  uint32_t root;
  uint32_t dim;
  uint32_t scalar;        /* ORACLE_F32 | ORACLE_F64 | ORACLE_I32 */
  uint64_t n_nodes;
  const uint32_t* nodes;  /* [n_nodes][3] = offset, size, parent (flatten_node_t) */
  const void* bounds;     /* children_bound: [n_children][2][dim], min then max */
  const uint32_t* children;
  const uint32_t* data;   /* mapped values as 32-bit ids */
} oracle_flat;

/* offsets: [n+1] caller-allocated; *ids: malloc'ed here, release with oracle_free.
 * Per query the ids are sorted ascending.  visits (nullable): [3] = internal node
 * visits, leaf node visits, entry tests, summed over the batch.  Returns total hits. */
int64_t oracle_query_boxes(const oracle_flat* f, const void* queries, uint64_t n,
                           int mode, uint64_t* offsets, uint32_t** ids,
                           uint64_t* visits);
int64_t oracle_brute_boxes(const void* keys, uint64_t n_keys, int key_is_box,
                           const uint32_t* key_ids, uint32_t dim, uint32_t scalar,
                           const void* queries, uint64_t n, int mode,
                           uint64_t* offsets, uint32_t** ids);
/* rays: [n][7] = o[3], inv_d[3], tmax */
int64_t oracle_query_rays(const oracle_flat* f, const float* rays, uint64_t n,
                          int mode, uint64_t* offsets, uint32_t** ids,
                          float* t_out, uint32_t* id_out, uint8_t* any_out,
                          uint64_t* visits);
int64_t oracle_brute_rays(const float* key_boxes, uint64_t n_keys,
                          const uint32_t* key_ids, const float* rays, uint64_t n,
                          int mode, uint64_t* offsets, uint32_t** ids,
                          float* t_out, uint32_t* id_out, uint8_t* any_out);
/* kNN (oracle/knn.h), float32 trees: points [n][dim]; ids_out / d2_out: [n][k] ascending by (d2, id),
 * padded with (0xFFFFFFFF, +inf).  visits (nullable): [2] internal / leaf node visits of the
 * depth-first traversal in child order.  Returns 0, or -1 for an unsupported tree. */
int64_t oracle_query_knn(const oracle_flat* f, int key_is_box, const float* points, uint64_t n, uint32_t k,
                         uint32_t* ids_out, float* d2_out, uint64_t* visits);
int64_t oracle_brute_knn(const float* keys, uint64_t n_keys, int key_is_box, const uint32_t* key_ids,
                         uint32_t dim, const float* points, uint64_t n, uint32_t k, uint32_t* ids_out,
                         float* d2_out);
void oracle_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
