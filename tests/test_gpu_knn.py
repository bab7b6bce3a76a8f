"""GPU parity tests for rt_query_knn through the C ABI.  kNN has no reference code: the definition is
oracle/knn.h (the reference's search() driven by a pruning filter and a k-best functor), the golden
vectors were produced by running exactly that through the reference, and dist2 is specified operation
for operation, so the bar is BIT-EXACT: same ids in the same order, same float32 distances."""
import os

import numpy as np
import pytest

import oracles as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def same(got, want):
    return np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32))


@pytest.mark.parametrize("name", ["knn_points_3d", "knn_points_2d", "knn_boxes_3d"])
def test_golden_knn(rtb, name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    flat = O.Flat.load(z)
    image = rtb.DeviceImage.from_flat(int(z["kind"]), flat.leaf_level, flat.nodes, flat.bounds, flat.children,
                                      flat.data)
    with image.upload(0) as tree:
        for k in (1, 8, 32):
            want = (z["ids_k%d" % k], z["d2_k%d" % k])
            assert same(tree.query_knn(z["points"], k), want)
            assert tree.stats()["kernel_launches"] > 0
            assert same(tree.query_knn(z["points"], k, rtb.RT_NO_REORDER), want)
        # visit counters are reported; pruning makes them order dependent, so only sanity-check them
        tree.query_knn(z["points"], 8, rtb.RT_COLLECT_VISITS)
        s = tree.stats()
        assert s["visits_leaf"] >= len(z["points"]) and s["visits_internal"] >= len(z["points"])


def test_knn_vs_reference_search_and_brute_force(rtb):
    rng = np.random.default_rng(71)
    n = 60000
    pts = rng.random((n, 3)).astype(np.float32)
    pts[:2000] = pts[rng.integers(2000, n, 2000)]            # duplicates: ties broken by id
    q = rng.uniform(-0.1, 1.1, (9000, 3)).astype(np.float32)  # > 4096: takes the coherence pass
    q[:500] = pts[:500]
    ref = O.RefTree(4).insert(pts)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        for k in (1, 5, 16, 32):
            assert same(tree.query_knn(q, k), ref.search_knn(q, k))
        got = tree.query_knn(q[:300], 10)
        assert same(got, O.oracle_brute_knn(pts, False, None, 3, q[:300], 10))
        assert np.all(got[1][:300, 0][:300] >= 0) and np.all(got[1][:300, 0] == 0.0)   # the data points themselves


@pytest.mark.parametrize("kind", [6, 10, 11, 3, 7, 5, 9])
def test_knn_on_every_float32_tree_kind(rtb, kind):
    """every float32 tree kind the engine accepts answers kNN -- 8-D with 16-wide blocks (BASELINE
    config 4's kind, QuadraticSplit), 1-D, 4-D, 2-D points, 2-D / 3-D box keys, QuadraticSplit 3-D --
    bit-exact against the reference's search() with the oracle/knn.h functors and against brute force"""
    scalar, dim, is_box, _ = O.KINDS[kind]
    assert scalar == O.F32
    rng = np.random.default_rng(400 + kind)
    n, nq = 30000, 6000
    c = rng.random((n, dim)).astype(np.float32)
    keys = np.stack([c, c + rng.uniform(0, 0.02, (n, dim)).astype(np.float32)], axis=1) if is_box else c
    keys = np.ascontiguousarray(keys, np.float32)
    q = rng.uniform(-0.1, 1.1, (nq, dim)).astype(np.float32)
    q[:200] = c[:200]
    ref = O.RefTree(kind).insert(keys)
    with rtb.RTree(kind).insert(keys).flatten_device().upload(0) as tree:
        for k in (1, 8, 32):
            assert same(tree.query_knn(q, k), ref.search_knn(q, k))
        assert same(tree.query_knn(q[:200], 7), O.oracle_brute_knn(keys, is_box, None, dim, q[:200], 7))


def test_knn_edge_cases(rtb):
    rng = np.random.default_rng(5)
    pts = rng.random((5, 3)).astype(np.float32)               # root is a leaf, fewer keys than k
    q = rng.random((40, 3)).astype(np.float32)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        ids, d2 = tree.query_knn(q, 8)
        assert same((ids, d2), O.oracle_brute_knn(pts, False, None, 3, q, 8))
        assert np.all(ids[:, 5:] == 0xFFFFFFFF) and np.all(np.isinf(d2[:, 5:]))
        ids, d2 = tree.query_knn(np.zeros((0, 3), np.float32), 4)                      # empty batch
        assert ids.shape == (0, 4)
        for bad_k in (0, 33):
            with pytest.raises(rtb.RtError) as err:
                tree.query_knn(q, bad_k)
            assert err.value.code == rtb.capi.RT_E_INVALID
    same_pts = np.tile(np.float32([[0.5, 0.25, 0.75]]), (3000, 1))                    # every key identical
    with rtb.RTree(4).insert(same_pts).flatten_device().upload(0) as tree:
        ids, d2 = tree.query_knn(q[:5], 32)
        assert np.array_equal(ids, np.tile(np.arange(32, dtype=np.uint32), (5, 1)))     # ties: the smallest ids
    with rtb.RTree(2).insert(rng.random((100, 2))).flatten_device().upload(0) as tree:   # float64 tree
        with pytest.raises(rtb.RtError) as err:
            tree.query_knn(np.zeros((1, 2), np.float32), 1)
        assert err.value.code == rtb.capi.RT_E_UNSUPPORTED


def test_knn_device_resident_and_scale_properties(rtb):
    """1 M queries against 2 M points: device-resident I/O; the k-th distance is non-decreasing in k and
    the k nearest of a data point start with that point"""
    import torch
    n, nq = 2_000_000, 1_000_000
    pts = rtb.workloads.uniform_points(n, seed=61)
    q = rtb.workloads.uniform_points(nq, seed=62)
    q[:1000] = pts[:1000]
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        dq = torch.from_numpy(q).cuda()
        flags = rtb.RT_QUERIES_ON_DEVICE | rtb.RT_RESULT_ON_DEVICE
        out = tree.query_knn_raw(dq.data_ptr(), nq, 8, flags)
        ids = torch.as_tensor(rtb.capi.DeviceBuffer(out.ids, nq * 8 * 4), device="cuda").view(torch.int32).view(nq, 8)
        d2 = torch.as_tensor(rtb.capi.DeviceBuffer(out.dist2, nq * 8 * 4), device="cuda").view(torch.float32).view(nq, 8)
        ids, d2 = ids.cpu().numpy().view(np.uint32), d2.cpu().numpy()
        host = tree.query_knn(q[:20000], 8)
        assert np.array_equal(host[0], ids[:20000]) and np.array_equal(host[1], d2[:20000])
    assert np.all(np.diff(d2, axis=1) >= 0) and np.all(ids < n)
    assert np.all(d2[:1000, 0] == 0.0)
    sample = np.random.default_rng(2).integers(0, nq, 300)
    assert same((ids[sample], d2[sample]), O.oracle_brute_knn(pts, False, None, 3, q[sample], 8))
