"""CPU tests that PIN the oracle (oracle/oracle.c) before anything trusts it.

1. against the committed golden vectors, which were produced by the unmodified reference
   (tests/golden/make_golden.py) -- runs everywhere, needs only oracle/liboracle.so;
2. against closed-form known answers the reference's own test holds for the path
   (test/rtree.cpp:410-467 "Flatten", example/sample/sample.cpp:33-64);
3. against the reference itself run live (oracle/_ref/libref_oracle.so) on fresh seeds, and
   against O(N*Q) brute force.
"""
import glob
import os

import numpy as np
import pytest

import oracles as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/libref_oracle.so not built")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, O.Flat.load(z)


def csr_same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_oracle_library_is_built():
    assert O.have_oracle(), "run `make -C oracle` (or __graft_entry__.build())"


BOX_FIXTURES = ["flatten_test_1d_i32", "sample_1d_f64", "config1_2d_f64", "config2_cut_3d_f32",
                "config4_cut_8d_f32", "teapot_boxes_intersects", "teapot_boxes_contains", "lattice_2d_i32"]


KNN_FIXTURES = ["knn_points_3d", "knn_points_2d", "knn_boxes_3d"]


def test_all_fixtures_present():
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    assert names == sorted(BOX_FIXTURES + ["teapot_rays"] + KNN_FIXTURES)


@pytest.mark.parametrize("name", BOX_FIXTURES)
def test_restatement_matches_golden_boxes(name):
    z, flat = load(name)
    off, ids = O.oracle_query_boxes(flat, z["queries"], int(z["mode"]))
    assert np.array_equal(off, z["offsets"])
    assert np.array_equal(ids, z["ids"])


@pytest.mark.parametrize("name", BOX_FIXTURES)
def test_brute_force_matches_golden_boxes(name):
    z, flat = load(name)
    off, ids = O.oracle_brute_boxes(z["keys"], flat.key_is_box, None, flat.dim, flat.scalar, z["queries"],
                                    int(z["mode"]))
    assert np.array_equal(off, z["offsets"])
    assert np.array_equal(ids, z["ids"])


def test_flatten_known_answer():
    """reference test/rtree.cpp:450-467: a closed range [min,max] over points 0..2999 returns
    exactly min..max."""
    z, flat = load("flatten_test_1d_i32")
    assert flat.root == 0 and len(flat.data) == 3000
    assert sorted(flat.data.tolist()) == list(range(3000))          # :422-433
    sizes = flat.nodes[1:, 1]
    assert sizes.min() >= 4 and sizes.max() <= 8                      # :435-445
    q = z["queries"].reshape(-1, 2)
    off, ids = O.oracle_query_boxes(flat, z["queries"], O.MODE_CONTAINS)
    for i, (lo, hi) in enumerate(q):
        assert np.array_equal(ids[off[i]:off[i + 1]], np.arange(lo, hi + 1))


def test_sample_known_answer():
    """reference example/sample/sample.cpp:33-64: 50 points, range [10,20] -> values 10..20."""
    z, flat = load("sample_1d_f64")
    off, ids = O.oracle_query_boxes(flat, z["queries"])
    assert ids.tolist() == list(range(10, 21))


def test_lattice_known_answer():
    z, flat = load("lattice_2d_i32")
    q = z["queries"]
    off, ids = O.oracle_query_boxes(flat, q)
    for i in range(len(q)):
        (x0, y0), (x1, y1) = q[i]
        xs = np.arange(max(x0, 0), min(x1, 39) + 1)
        ys = np.arange(max(y0, 0), min(y1, 39) + 1)
        want = (xs[:, None] * 40 + ys[None, :]).reshape(-1)
        assert np.array_equal(ids[off[i]:off[i + 1]], np.sort(want))


def test_restatement_matches_golden_rays():
    z, flat = load("teapot_rays")
    off, ids = O.oracle_query_rays(flat, z["rays"], O.RAY_ALL)
    assert np.array_equal(off, z["offsets"]) and np.array_equal(ids, z["ids"])
    t, tid = O.oracle_query_rays(flat, z["rays"], O.RAY_CLOSEST)
    assert np.array_equal(t.view(np.uint32), z["closest_t"].view(np.uint32))
    assert np.array_equal(tid, z["closest_id"])
    assert np.array_equal(O.oracle_query_rays(flat, z["rays"], O.RAY_ANY), z["anyhit"])
    # topology-independent answers
    boff, bids = O.oracle_brute_rays(z["keys"], None, z["rays"], O.RAY_ALL)
    assert np.array_equal(boff, off) and np.array_equal(bids, ids)
    bt, bid = O.oracle_brute_rays(z["keys"], None, z["rays"], O.RAY_CLOSEST)
    assert np.array_equal(bt.view(np.uint32), t.view(np.uint32)) and np.array_equal(bid, tid)
    assert z["anyhit"].sum() > 100  # the workload does hit the mesh


def test_closest_is_min_of_all_hits():
    z, flat = load("teapot_rays")
    t, tid = O.oracle_query_rays(flat, z["rays"], O.RAY_CLOSEST)
    anyhit = O.oracle_query_rays(flat, z["rays"], O.RAY_ANY)
    off, _ = O.oracle_query_rays(flat, z["rays"], O.RAY_ALL)
    counts = np.diff(off.astype(np.int64))
    assert np.array_equal(anyhit.astype(bool), counts > 0)
    assert np.array_equal(tid != 0xFFFFFFFF, counts > 0)
    assert np.all(np.isinf(t[counts == 0]))


# ---- live against the reference -------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("kind,n,nq", [(0, 700, 300), (2, 1500, 500), (3, 3000, 500), (4, 5000, 800),
                                        (6, 3000, 300), (8, 900, 300), (9, 4000, 500), (11, 3000, 400),
                                        (12, 3000, 400), (13, 2000, 300)])
def test_restatement_matches_reference_search_points(kind, n, nq):
    scalar, dim, is_box, _ = O.KINDS[kind]
    rng = np.random.default_rng(100 + kind)
    if scalar == O.I32:
        keys = rng.integers(0, 60, (n, dim)).astype(np.int32)
        lo = rng.integers(-5, 60, (nq, dim))
        q = np.stack([lo, lo + rng.integers(0, 15, (nq, dim))], axis=1).astype(np.int32)
    else:
        keys = rng.random((n, dim))
        c = rng.random((nq, dim))
        h = 0.5 * (10.0 / n) ** (1.0 / dim) * rng.uniform(0.5, 2.0, (nq, 1))
        q = np.stack([c - h, c + h], axis=1)
    tree = O.RefTree(kind).insert(keys)
    flat = tree.flatten()
    ro, ri = tree.search_boxes(q)
    oo, oi = O.oracle_query_boxes(flat, q)
    bo, bi = O.oracle_brute_boxes(np.ascontiguousarray(keys, tree.dtype), False, None, dim, scalar, q)
    assert np.array_equal(ro, oo) and np.array_equal(ri, oi)
    assert np.array_equal(ro, bo) and np.array_equal(ri, bi)
    assert ri.size > 0


@needs_ref
@pytest.mark.parametrize("mode", [O.MODE_INTERSECTS, O.MODE_CONTAINS])
def test_restatement_matches_reference_search_box_keys(mode):
    rng = np.random.default_rng(5)
    c = rng.random((4000, 3)).astype(np.float32)
    h = rng.uniform(0.0, 0.03, (4000, 3)).astype(np.float32)
    keys = np.stack([c - h, c + h], axis=1)
    qc = rng.random((600, 3)).astype(np.float32)
    qh = rng.uniform(0.01, 0.2, (600, 3)).astype(np.float32)
    q = np.stack([qc - qh, qc + qh], axis=1)
    tree = O.RefTree(5).insert(keys)
    ro, ri = tree.search_boxes(q, mode)
    oo, oi = O.oracle_query_boxes(tree.flatten(), q, mode)
    bo, bi = O.oracle_brute_boxes(keys, True, None, 3, O.F32, q, mode)
    assert np.array_equal(ro, oo) and np.array_equal(ri, oi)
    assert np.array_equal(ro, bo) and np.array_equal(ri, bi)
    assert ri.size > 0


@needs_ref
@pytest.mark.parametrize("kind", [8, 13, 4, 5, 7, 2, 6, 0])
def test_strict_overlap_restatement_matches_the_reference_traits(kind):
    """MODE_STRICT: search() filtered by the reference's OWN geometry_traits<aabb>::is_overlap
    (include/RTree/aabb.hpp:225-236; strict for N-D boxes, closed in 1-D) instead of the closed overlap
    of its sample / test.  Lattice coordinates make touching boxes and keys on query faces the common
    case; the restatement walks the reference's flatten() and must deliver the same lists (this filter
    is not conservative for keys on a face, so the hit set depends on the tree: same tree, same answer)."""
    scalar, dim, is_box, _ = O.KINDS[kind]
    dt = O.SCALAR_DTYPES[scalar]
    rng = np.random.default_rng(500 + kind)
    n, nq = 4000, 3000
    side = {1: 3000, 2: 60, 3: 16, 8: 3}[dim]
    c = rng.integers(0, side, (n, dim))
    if is_box:
        keys = np.stack([c, c + rng.integers(0, 3, (n, dim))], axis=1).astype(dt)
    else:
        keys = c.astype(dt)
    lo = rng.integers(-1, side, (nq, dim))
    q = np.stack([lo, lo + rng.integers(0, 4, (nq, dim))], axis=1).astype(dt)
    ref = O.RefTree(kind).insert(keys)
    flat = ref.flatten()
    base = O.MODE_INTERSECTS if is_box else O.MODE_POINT_IN_BOX
    strict = ref.search_boxes(q, base | O.MODE_STRICT)
    closed = ref.search_boxes(q, base)
    assert csr_same(O.oracle_query_boxes(flat, q, base | O.MODE_STRICT), strict)
    # never more than the closed filter finds; in 1-D the reference's is_overlap IS closed
    assert np.all(np.diff(strict[0].astype(np.int64)) <= np.diff(closed[0].astype(np.int64)))
    if dim == 1:
        assert csr_same(strict, closed)
    else:
        assert strict[1].size < closed[1].size
    if is_box:
        assert csr_same(O.oracle_query_boxes(flat, q, O.MODE_CONTAINS | O.MODE_STRICT),
                        ref.search_boxes(q, O.MODE_CONTAINS | O.MODE_STRICT))


@needs_ref
def test_nan_and_degenerate_queries_match_reference():
    """N-D predicates are written as "not greater" per axis (reference aabb.hpp:249-259), so a
    NaN bound simply does not constrain that side; inverted and zero-size boxes behave too."""
    rng = np.random.default_rng(9)
    pts = rng.random((3000, 3)).astype(np.float32)
    tree = O.RefTree(4).insert(pts)
    q = np.zeros((6, 2, 3), np.float32)
    q[0] = [[0.2, 0.2, 0.2], [0.4, np.nan, 0.4]]
    q[1] = [[np.nan, 0.1, 0.1], [0.3, 0.3, 0.3]]
    q[2] = [[0.6, 0.6, 0.6], [0.5, 0.5, 0.5]]          # inverted: no hits
    q[3] = [pts[17], pts[17]]                          # zero-size box on a point
    q[4] = [[-np.inf] * 3, [np.inf] * 3]               # everything
    q[5] = [[np.nan] * 3, [np.nan] * 3]                # everything (nothing constrains)
    ro, ri = tree.search_boxes(q)
    oo, oi = O.oracle_query_boxes(tree.flatten(), q)
    assert np.array_equal(ro, oo) and np.array_equal(ri, oi)
    counts = np.diff(ro.astype(np.int64))
    assert counts[2] == 0 and counts[3] >= 1 and counts[4] == 3000 and counts[5] == 3000
    assert counts[0] > 0 and counts[1] > 0


@needs_ref
def test_empty_and_root_leaf_trees():
    """reference rtree.hpp:990 / appendix B4: a root that is a leaf hands every value to the
    functor; an empty tree flattens to one node of size 0."""
    empty = O.RefTree(4)
    flat = empty.flatten()
    assert flat.leaf_level == 0 and len(flat.nodes) == 1 and flat.nodes[0, 1] == 0
    q = np.array([[[0, 0, 0], [1, 1, 1]]], np.float32)
    off, ids = O.oracle_query_boxes(flat, q)
    assert off.tolist() == [0, 0] and ids.size == 0
    small = O.RefTree(4).insert(np.array([[0.1, 0.1, 0.1], [0.5, 0.5, 0.5], [0.9, 0.9, 0.9]], np.float32))
    flat = small.flatten()
    assert flat.leaf_level == 0
    ro, ri = small.search_boxes(np.array([[[0.4, 0.4, 0.4], [1, 1, 1]]], np.float32))
    oo, oi = O.oracle_query_boxes(flat, np.array([[[0.4, 0.4, 0.4], [1, 1, 1]]], np.float32))
    assert ri.tolist() == [1, 2] and oi.tolist() == [1, 2] and np.array_equal(ro, oo)


@needs_ref
def test_rays_match_reference_search_and_brute_force():
    rng = np.random.default_rng(11)
    c = rng.uniform(-1, 1, (2500, 3)).astype(np.float32)
    h = rng.uniform(0.0, 0.08, (2500, 3)).astype(np.float32)
    keys = np.stack([c - h, c + h], axis=1)
    tree = O.RefTree(5).insert(keys)
    flat = tree.flatten()
    o = rng.uniform(-3, 3, (1500, 3)).astype(np.float32)
    d = rng.normal(size=(1500, 3)).astype(np.float32)
    d[:100, 0] = 0.0                                   # axis-parallel rays: inv = +-inf, NaN paths
    d[100:200, 1] = -0.0
    with np.errstate(divide="ignore"):
        inv = (np.float32(1.0) / d).astype(np.float32)
    rays = np.concatenate([o, inv, np.full((1500, 1), 1e30, np.float32)], axis=1)
    rays[200:300, 6] = 2.0                             # short rays
    for mode in (O.RAY_ALL, O.RAY_CLOSEST, O.RAY_ANY):
        ref = tree.search_rays(rays, mode)
        mine = O.oracle_query_rays(flat, rays, mode)
        brute = O.oracle_brute_rays(keys, None, rays, mode)
        if mode == O.RAY_ANY:
            assert np.array_equal(ref, mine) and np.array_equal(ref, brute)
        elif mode == O.RAY_CLOSEST:
            assert np.array_equal(ref[0].view(np.uint32), mine[0].view(np.uint32))
            assert np.array_equal(ref[0].view(np.uint32), brute[0].view(np.uint32))
            assert np.array_equal(ref[1], mine[1]) and np.array_equal(ref[1], brute[1])
        else:
            assert np.array_equal(ref[0], mine[0]) and np.array_equal(ref[1], mine[1])
            assert np.array_equal(ref[0], brute[0]) and np.array_equal(ref[1], brute[1])
            assert ref[1].size > 300


@needs_ref
def test_visit_counters_define_algorithmic_bytes():
    """bytes(query) = V_int*B_int + V_leaf*B_leaf + ...; the counters the oracle
    reports are what the engine's RT_COLLECT_VISITS totals are checked against on the GPU."""
    n = 30000
    rng = np.random.default_rng(3)
    pts = rng.random((n, 3)).astype(np.float32)
    c = rng.random((500, 3)).astype(np.float32)
    q = np.stack([c - 0.03, c + 0.03], axis=1).astype(np.float32)
    flat = O.RefTree(4).insert(pts).flatten()
    _, ids, visits = O.oracle_query_boxes(flat, q, want_visits=True)
    v_int, v_leaf, tests = [int(x) for x in visits]
    assert v_int >= 500 * flat.leaf_level and v_leaf > 0 and tests >= v_int + v_leaf
    assert ids.size > 0


@pytest.mark.parametrize("name", KNN_FIXTURES)
def test_knn_restatement_and_brute_force_match_golden(name):
    """kNN has no reference code: the golden vectors come from the reference's search() driven by the
    functors of oracle/knn.h; the flat-form restatement and the O(N*Q) brute force must agree with
    them bit for bit (ids and float32 distances), ties by the smaller id"""
    z, flat = load(name)
    q = z["points"]
    for k in (1, 8, 32):
        ids, d2 = O.oracle_query_knn(flat, q, k)
        assert np.array_equal(ids, z["ids_k%d" % k])
        assert np.array_equal(d2.view(np.uint32), z["d2_k%d" % k].view(np.uint32))
        sub = slice(0, 300)
        bi, bd = O.oracle_brute_knn(z["keys"], flat.key_is_box, None, flat.dim, q[sub], k)
        assert np.array_equal(bi, z["ids_k%d" % k][sub])
        assert np.array_equal(bd.view(np.uint32), z["d2_k%d" % k][sub].view(np.uint32))
    assert np.all(np.diff(z["d2_k32"], axis=1) >= 0)


@pytest.mark.skipif(not O.have_ref(), reason="reference oracle not built")
def test_knn_through_reference_search_live():
    rng = np.random.default_rng(33)
    pts = rng.random((7000, 3)).astype(np.float32)
    q = rng.random((400, 3)).astype(np.float32)
    ref = O.RefTree(4).insert(pts)
    flat = ref.flatten()
    for k in (3, 17):
        a, b = ref.search_knn(q, k), O.oracle_query_knn(flat, q, k)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    few = O.RefTree(4).insert(pts[:5])
    ids, d2 = few.search_knn(q[:10], 8)
    assert np.all(ids[:, 5:] == 0xFFFFFFFF) and np.all(np.isinf(d2[:, 5:])) and np.all(ids[:, :5] < 5)
