/* oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the reference's read-only query
 * path, walking the buffers RTree::flatten() emits
 * (include/RTree/rtree.hpp:841-870, :927-962):
 *
 *   - box queries : RTree::search_recursive (include/RTree/rtree.hpp:984-1019) with
 *                   the canonical functors of SURVEY.md 8a, in the flat-form shape
 *                   of search_flatten (test/rtree.cpp:376-409)       oracle_box.inc
 *   - ray queries : the same traversal driven by the slab test of oracle/slab.h
 *                   (all-hits / closest-hit with non-strict pruning / any-hit abort,
 *                   include/RTree/rtree.hpp:994-997, :1004-1015)
 *   - brute force : O(N*Q) over the raw keys, topology independent
 *
 * PARITY PINNING: this restatement is pinned against the reference itself run in
 * this repository's container -- oracle/_ref/libref_oracle.so, built from
 * /root/reference/include by oracle/Makefile -- by tests/test_oracle.py, including a
 * replay of the reference's one known-answer test on the path (test/rtree.cpp:
 * 410-467, "Flatten") and of example/sample/sample.cpp:33-64, and against the
 * committed golden vectors under tests/golden/ (generated from the reference by
 * tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load this library.  The product never does.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "slab.h"
#include "knn.h"
#include <math.h>

#define ORACLE_MAX_DIM 16

/* ---- growable id vector ---------------------------------------------------- */
typedef struct
{
  uint32_t* p;
  uint64_t n, cap;
} idvec;
static void idvec_init(idvec* v)
{
  v->cap = 1024;
  v->n = 0;
  v->p = (uint32_t*)malloc(sizeof(uint32_t) * v->cap);
}
static void idvec_push(idvec* v, uint32_t x)
{
  if (v->n == v->cap)
  {
    v->cap *= 2;
    v->p = (uint32_t*)realloc(v->p, sizeof(uint32_t) * v->cap);
  }
  v->p[v->n++] = x;
}
static int cmp_u32(const void* a, const void* b)
{
  const uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
  return (x > y) - (x < y);
}

#define S float
#define SFX f32
#include "oracle_box.inc"
#undef S
#undef SFX

#define S double
#define SFX f64
#include "oracle_box.inc"
#undef S
#undef SFX

#define S int32_t
#define SFX i32
#include "oracle_box.inc"
#undef S
#undef SFX

int64_t oracle_query_boxes(const oracle_flat* f, const void* queries, uint64_t n,
                           int mode, uint64_t* offsets, uint32_t** ids,
                           uint64_t* visits)
{
  if (f->dim == 0 || f->dim > ORACLE_MAX_DIM)
    return -1;
  switch (f->scalar)
  {
  case ORACLE_F32: return query_boxes_f32(f, queries, n, mode, offsets, ids, visits);
  case ORACLE_F64: return query_boxes_f64(f, queries, n, mode, offsets, ids, visits);
  case ORACLE_I32: return query_boxes_i32(f, queries, n, mode, offsets, ids, visits);
  default: return -1;
  }
}

int64_t oracle_brute_boxes(const void* keys, uint64_t n_keys, int key_is_box,
                           const uint32_t* key_ids, uint32_t dim, uint32_t scalar,
                           const void* queries, uint64_t n, int mode,
                           uint64_t* offsets, uint32_t** ids)
{
  if (dim == 0 || dim > ORACLE_MAX_DIM)
    return -1;
  switch (scalar)
  {
  case ORACLE_F32:
    return brute_boxes_f32(keys, n_keys, key_is_box, key_ids, dim, queries, n, mode, offsets, ids);
  case ORACLE_F64:
    return brute_boxes_f64(keys, n_keys, key_is_box, key_ids, dim, queries, n, mode, offsets, ids);
  case ORACLE_I32:
    return brute_boxes_i32(keys, n_keys, key_is_box, key_ids, dim, queries, n, mode, offsets, ids);
  default: return -1;
  }
}

/* ---- rays (f32, 3-D) ----------------------------------------------------------- */
typedef struct
{
  const oracle_flat* f;
  const float* ray; /* o[3], inv[3], tmax */
  int mode;
  idvec* out;
  float t_best;
  uint32_t id_best;
  int found;
  uint64_t v_int, v_leaf, tests;
} rayctx;

/* returns 1 to abort the whole search (include/RTree/rtree.hpp:994-997,:1006-1013) */
static int walk_ray(rayctx* c, uint32_t node, uint32_t level)
{
  const oracle_flat* f = c->f;
  const uint32_t off = f->nodes[3 * (size_t)node + 0];
  const uint32_t size = f->nodes[3 * (size_t)node + 1];
  const float* bounds = (const float*)f->bounds;
  const float* r = c->ray;
  uint32_t i;
  if (level == f->leaf_level)
  {
    c->v_leaf++;
    for (i = 0; i < size; ++i)
    {
      const float* b = bounds + (size_t)(off + i) * 6;
      float t;
      c->tests++;
      if (!rt_slab_hit(b, b + 3, r, r + 3, r[6], &t))
        continue;
      {
        const uint32_t id = f->data[f->children[off + i]];
        if (c->mode == ORACLE_RAY_ALL)
        {
          idvec_push(c->out, id);
        }
        else if (c->mode == ORACLE_RAY_CLOSEST)
        {
          if (!c->found || t < c->t_best || (t == c->t_best && id < c->id_best))
          {
            c->found = 1;
            c->t_best = t;
            c->id_best = id;
          }
        }
        else
        {
          c->found = 1;
          return 1;
        }
      }
    }
  }
  else
  {
    c->v_int++;
    for (i = 0; i < size; ++i)
    {
      const float* b = bounds + (size_t)(off + i) * 6;
      float t;
      c->tests++;
      if (!rt_slab_hit(b, b + 3, r, r + 3, r[6], &t))
        continue;
      if (c->mode == ORACLE_RAY_CLOSEST && !(t <= c->t_best))
        continue;
      if (walk_ray(c, f->children[off + i], level + 1))
        return 1;
    }
  }
  return 0;
}

int64_t oracle_query_rays(const oracle_flat* f, const float* rays, uint64_t n,
                          int mode, uint64_t* offsets, uint32_t** ids,
                          float* t_out, uint32_t* id_out, uint8_t* any_out,
                          uint64_t* visits)
{
  idvec all;
  rayctx c;
  uint64_t i;
  int64_t total = 0;
  if (f->dim != 3 || f->scalar != ORACLE_F32)
    return -1;
  idvec_init(&all);
  c.f = f;
  c.mode = mode;
  c.out = &all;
  c.v_int = c.v_leaf = c.tests = 0;
  for (i = 0; i < n; ++i)
  {
    const uint64_t begin = all.n;
    c.ray = rays + i * 7;
    c.t_best = c.ray[6];
    c.id_best = 0xFFFFFFFFu;
    c.found = 0;
    if (offsets)
      offsets[i] = begin;
    if (f->n_nodes > 0)
      walk_ray(&c, f->root, 0);
    if (mode == ORACLE_RAY_ALL)
    {
      qsort(all.p + begin, (size_t)(all.n - begin), sizeof(uint32_t), cmp_u32);
    }
    else if (mode == ORACLE_RAY_CLOSEST)
    {
      if (t_out)
        t_out[i] = c.found ? c.t_best : RT_SLAB_MISS_T;
      if (id_out)
        id_out[i] = c.id_best;
      total += c.found;
    }
    else
    {
      if (any_out)
        any_out[i] = (uint8_t)c.found;
      total += c.found;
    }
  }
  if (offsets)
    offsets[n] = all.n;
  if (visits)
  {
    visits[0] = c.v_int;
    visits[1] = c.v_leaf;
    visits[2] = c.tests;
  }
  if (mode == ORACLE_RAY_ALL)
  {
    if (ids)
      *ids = all.p;
    else
      free(all.p);
    return (int64_t)all.n;
  }
  free(all.p);
  return total;
}

/* O(N*Q) brute force for rays over raw key boxes [n_keys][2][3]. */
int64_t oracle_brute_rays(const float* key_boxes, uint64_t n_keys,
                          const uint32_t* key_ids, const float* rays, uint64_t n,
                          int mode, uint64_t* offsets, uint32_t** ids,
                          float* t_out, uint32_t* id_out, uint8_t* any_out)
{
  idvec all;
  uint64_t i, k;
  int64_t total = 0;
  idvec_init(&all);
  for (i = 0; i < n; ++i)
  {
    const float* r = rays + i * 7;
    const uint64_t begin = all.n;
    float t_best = r[6];
    uint32_t id_best = 0xFFFFFFFFu;
    int found = 0;
    if (offsets)
      offsets[i] = begin;
    for (k = 0; k < n_keys; ++k)
    {
      const float* b = key_boxes + k * 6;
      const uint32_t id = key_ids ? key_ids[k] : (uint32_t)k;
      float t;
      if (!rt_slab_hit(b, b + 3, r, r + 3, r[6], &t))
        continue;
      if (mode == ORACLE_RAY_ALL)
        idvec_push(&all, id);
      if (!found || t < t_best || (t == t_best && id < id_best))
      {
        found = 1;
        t_best = t;
        id_best = id;
      }
    }
    if (mode == ORACLE_RAY_ALL)
      qsort(all.p + begin, (size_t)(all.n - begin), sizeof(uint32_t), cmp_u32);
    if (mode == ORACLE_RAY_CLOSEST)
    {
      if (t_out)
        t_out[i] = found ? t_best : RT_SLAB_MISS_T;
      if (id_out)
        id_out[i] = id_best;
    }
    if (mode == ORACLE_RAY_ANY && any_out)
      any_out[i] = (uint8_t)found;
    total += found;
  }
  if (offsets)
    offsets[n] = all.n;
  if (mode == ORACLE_RAY_ALL)
  {
    if (ids)
      *ids = all.p;
    else
      free(all.p);
    return (int64_t)all.n;
  }
  free(all.p);
  return total;
}

void oracle_free(void* p) { free(p); }

/* ---- kNN (f32) ------------------------------------------------------------------ */
typedef struct
{
  float d2;
  uint32_t id;
} knn_pair;

typedef struct
{
  const oracle_flat* f;
  int key_is_box;
  const float* q;
  uint32_t k, have;
  knn_pair* best; /* [k], unordered; worst tracked below */
  uint32_t worst; /* index of the worst pair once have == k */
  uint64_t v_int, v_leaf;
} knnctx;

static void knn_find_worst(knnctx* c)
{
  uint32_t i, w = 0;
  for (i = 1; i < c->k; ++i)
    if (rt_knn_less(c->best[w].d2, c->best[w].id, c->best[i].d2, c->best[i].id))
      w = i;
  c->worst = w;
}
static void knn_offer(knnctx* c, float d2, uint32_t id)
{
  if (c->have < c->k)
  {
    c->best[c->have].d2 = d2;
    c->best[c->have].id = id;
    if (++c->have == c->k)
      knn_find_worst(c);
    return;
  }
  if (rt_knn_less(d2, id, c->best[c->worst].d2, c->best[c->worst].id))
  {
    c->best[c->worst].d2 = d2;
    c->best[c->worst].id = id;
    knn_find_worst(c);
  }
}
static float knn_bound(const knnctx* c)
{
  return c->have < c->k ? INFINITY : c->best[c->worst].d2;
}
static void walk_knn(knnctx* c, uint32_t node, uint32_t level)
{
  const oracle_flat* f = c->f;
  const uint32_t off = f->nodes[3 * node], size = f->nodes[3 * node + 1], dim = f->dim;
  const float* b = (const float*)f->bounds;
  uint32_t i;
  if (level == f->leaf_level)
  {
    c->v_leaf++;
    for (i = 0; i < size; ++i)
    {
      const float* lo = b + (uint64_t)(off + i) * 2 * dim;
      knn_offer(c, rt_knn_dist2(lo, lo + dim, c->q, (int)dim), f->data[f->children[off + i]]);
    }
    return;
  }
  c->v_int++;
  for (i = 0; i < size; ++i)
  {
    const float* lo = b + (uint64_t)(off + i) * 2 * dim;
    if (rt_knn_dist2(lo, lo + dim, c->q, (int)dim) <= knn_bound(c))
      walk_knn(c, f->children[off + i], level + 1);
  }
}
static int cmp_knn(const void* a, const void* b)
{
  const knn_pair* x = (const knn_pair*)a;
  const knn_pair* y = (const knn_pair*)b;
  if (x->d2 < y->d2)
    return -1;
  if (x->d2 > y->d2)
    return 1;
  return (x->id > y->id) - (x->id < y->id);
}
static void knn_emit(knn_pair* best, uint32_t have, uint32_t k, uint32_t* ids_out, float* d2_out)
{
  uint32_t i;
  qsort(best, have, sizeof(knn_pair), cmp_knn);
  for (i = 0; i < k; ++i)
  {
    ids_out[i] = i < have ? best[i].id : 0xFFFFFFFFu;
    d2_out[i] = i < have ? best[i].d2 : INFINITY;
  }
}

int64_t oracle_query_knn(const oracle_flat* f, int key_is_box, const float* points, uint64_t n, uint32_t k,
                         uint32_t* ids_out, float* d2_out, uint64_t* visits)
{
  knnctx c;
  uint64_t i;
  if (f->scalar != ORACLE_F32 || f->dim == 0 || f->dim > ORACLE_MAX_DIM || k == 0)
    return -1;
  memset(&c, 0, sizeof c);
  c.f = f;
  c.key_is_box = key_is_box;
  c.k = k;
  c.best = (knn_pair*)malloc(sizeof(knn_pair) * k);
  for (i = 0; i < n; ++i)
  {
    c.q = points + i * f->dim;
    c.have = 0;
    if (f->n_nodes > 0)
      walk_knn(&c, f->root, 0);
    knn_emit(c.best, c.have, k, ids_out + i * k, d2_out + i * k);
  }
  if (visits)
  {
    visits[0] = c.v_int;
    visits[1] = c.v_leaf;
  }
  free(c.best);
  return 0;
}

int64_t oracle_brute_knn(const float* keys, uint64_t n_keys, int key_is_box, const uint32_t* key_ids,
                         uint32_t dim, const float* points, uint64_t n, uint32_t k, uint32_t* ids_out,
                         float* d2_out)
{
  knnctx c;
  uint64_t i, j;
  const uint64_t stride = (uint64_t)dim * (key_is_box ? 2 : 1);
  if (dim == 0 || dim > ORACLE_MAX_DIM || k == 0)
    return -1;
  memset(&c, 0, sizeof c);
  c.k = k;
  c.best = (knn_pair*)malloc(sizeof(knn_pair) * k);
  for (i = 0; i < n; ++i)
  {
    c.have = 0;
    for (j = 0; j < n_keys; ++j)
    {
      const float* lo = keys + j * stride;
      const float* hi = key_is_box ? lo + dim : lo;
      knn_offer(&c, rt_knn_dist2(lo, hi, points + i * dim, (int)dim), key_ids ? key_ids[j] : (uint32_t)j);
    }
    knn_emit(c.best, c.have, k, ids_out + i * k, d2_out + i * k);
  }
  free(c.best);
  return 0;
}
