"""Generate the golden vectors under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container (where /root/reference exists and `make -C oracle` has produced
oracle/_ref/libref_oracle.so, the unmodified reference behind a C shim):

    python tests/golden/make_golden.py

Each .npz holds (a) the inputs, (b) the reference's own flatten() output, and (c) the hit
lists of the reference's own RTree::search() with the canonical functors, sorted per query.
The fixtures are small (a few hundred kB) and committed; tests replay them against the
C restatement (CPU) and against the CUDA engine (GPU).
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracles as O  # noqa: E402

W = importlib.import_module("r-star-tree_b200").workloads


def save(name, flat, **arrays):
    path = os.path.join(HERE, name + ".npz")
    flat.save(path, **arrays)
    print("%-28s %7.1f kB" % (name + ".npz", os.path.getsize(path) / 1e3))


def boxes_case(name, kind, keys, queries, mode):
    t = O.RefTree(kind).insert(keys)
    flat = t.flatten()
    off, ids = t.search_boxes(queries, mode)
    save(name, flat, kind=kind, keys=keys, queries=queries, mode=mode, offsets=off, ids=ids)


def main():
    rng = np.random.default_rng(7)

    # 1. the reference's one known-answer test on the path: test/rtree.cpp:410-467 ("Flatten")
    a = rng.integers(1, 3000, size=(200, 2)).astype(np.int32)
    a.sort(axis=1)
    boxes_case("flatten_test_1d_i32", 0, np.arange(3000, dtype=np.int32), a.reshape(-1, 2, 1),
               O.MODE_POINT_IN_BOX)

    # 2. example/sample/sample.cpp:33-64: 50 points 0..49, rebalance(), range [10, 20]
    t = O.RefTree(1).insert(np.arange(50, dtype=np.float64))
    t.rebalance()
    q = np.array([[[10.0], [20.0]]])
    off, ids = t.search_boxes(q)
    save("sample_1d_f64", t.flatten(), kind=1, keys=np.arange(50, dtype=np.float64), queries=q,
         mode=0, offsets=off, ids=ids)

    # 3. BASELINE config 1: visualize_2d points, 10k boxes (2,000 kept in the fixture)
    boxes_case("config1_2d_f64", 2, W.config1_points(1000, seed=1), W.config1_queries(2000, seed=2),
               O.MODE_POINT_IN_BOX)

    # 4. a cut of BASELINE config 2: 3-D uniform points, small cubes
    n = 20000
    boxes_case("config2_cut_3d_f32", 4, W.uniform_points(n, seed=42), W.cube_queries(3000, n, seed=43),
               O.MODE_POINT_IN_BOX)

    # 5. a cut of BASELINE config 4: 8-D mixture, MAX_ENTRIES 16, QuadraticSplit
    pts = W.mixture_points(6000, seed=45)
    boxes_case("config4_cut_8d_f32", 6, pts, W.mixture_queries(1000, pts, 0.035, seed=46),
               O.MODE_POINT_IN_BOX)

    # 6. box keys (triangle AABBs of the procedural teapot): INTERSECTS and CONTAINS
    tris = W.teapot_triangles(0.5)
    boxes = W.triangle_aabbs(tris)
    c = rng.uniform(-3.0, 3.5, (1500, 3)).astype(np.float32)
    h = rng.uniform(0.05, 0.8, (1500, 3)).astype(np.float32)
    qb = np.stack([c - h, c + h], axis=1)
    boxes_case("teapot_boxes_intersects", 5, boxes, qb, O.MODE_INTERSECTS)
    boxes_case("teapot_boxes_contains", 5, boxes, qb, O.MODE_CONTAINS)

    # 7. BASELINE config 3: rays against the teapot AABBs, all three modes
    t = O.RefTree(5).insert(boxes)
    rays = W.random_rays(3000, boxes, seed=44, tmax=1e30)
    off, ids = t.search_rays(rays, O.RAY_ALL)
    tt, tid = t.search_rays(rays, O.RAY_CLOSEST)
    anyhit = t.search_rays(rays, O.RAY_ANY)
    save("teapot_rays", t.flatten(), kind=5, keys=boxes, rays=rays, offsets=off, ids=ids, closest_t=tt,
         closest_id=tid, anyhit=anyhit)

    # 8. 2-D integer lattice (the Flatten test lifted to 2-D): exact known answer
    g = np.stack(np.meshgrid(np.arange(40), np.arange(40), indexing="ij"), axis=2).reshape(-1, 2).astype(np.int32)
    lo = rng.integers(0, 40, size=(300, 2)).astype(np.int32)
    hi = lo + rng.integers(0, 12, size=(300, 2)).astype(np.int32)
    boxes_case("lattice_2d_i32", 8, g, np.stack([lo, hi], axis=1), O.MODE_POINT_IN_BOX)


def knn_cases():
    """kNN (oracle/knn.h) through the reference's search(): 3-D points with duplicate keys (ties), 2-D
    points, 3-D box keys; k = 1, 8, 32; query points include exact data points."""
    rng = np.random.default_rng(17)
    pts3 = W.uniform_points(20000, seed=48)
    pts3[:300] = pts3[300:600]
    pts2 = rng.random((8000, 2)).astype(np.float32)
    boxes = W.triangle_aabbs(W.teapot_triangles(0.5))
    for name, kind, keys, lo, hi in (("knn_points_3d", 4, pts3, 0.0, 1.0), ("knn_points_2d", 3, pts2, -0.2, 1.2),
                                     ("knn_boxes_3d", 5, boxes, -3.5, 3.5)):
        t = O.RefTree(kind).insert(keys)
        dim = keys.shape[-1]
        q = rng.uniform(lo, hi, (1500, dim)).astype(np.float32)
        if kind != 5:
            q[:200] = keys[:200]
        extra = {}
        for k in (1, 8, 32):
            ids, d2 = t.search_knn(q, k)
            extra["ids_k%d" % k] = ids
            extra["d2_k%d" % k] = d2
        save(name, t.flatten(), kind=kind, keys=keys, points=q, **extra)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "knn":
        knn_cases()
    else:
        main()
        knn_cases()
