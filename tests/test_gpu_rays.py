"""GPU parity tests for rt_query_rays through the C ABI.  The slab test is only sub / mul /
compare-select with the reciprocal direction precomputed by the caller, so host and device
round identically and the bar is BIT-EXACT (tighter than BASELINE's 1-ulp / 1e-6 allowance):
hit sets equal, closest t equal as bit patterns, closest id equal."""
import os

import numpy as np
import pytest

import oracles as O
from helpers import csr_equal

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def check_all_modes(rtb, tree, rays, want_all, want_closest, want_any):
    assert csr_equal(tree.query_rays(rays, rtb.RT_RAY_ALL_HITS), want_all)
    assert csr_equal(tree.query_rays(rays, rtb.RT_RAY_ALL_HITS, rtb.RT_COLLECT_VISITS), want_all)
    t, tid = tree.query_rays(rays, rtb.RT_RAY_CLOSEST)
    assert np.array_equal(t.view(np.uint32), want_closest[0].view(np.uint32))
    assert np.array_equal(tid, want_closest[1])
    assert np.array_equal(tree.query_rays(rays, rtb.RT_RAY_ANY), want_any)


def test_golden_teapot_rays(rtb):
    z = np.load(os.path.join(GOLDEN, "teapot_rays.npz"))
    flat = O.Flat.load(z)
    image = rtb.DeviceImage.from_flat(5, flat.leaf_level, flat.nodes, flat.bounds, flat.children, flat.data)
    with image.upload(0) as tree:
        check_all_modes(rtb, tree, z["rays"], (z["offsets"], z["ids"]), (z["closest_t"], z["closest_id"]),
                        z["anyhit"])
        s_rays = tree.query_rays(z["rays"], rtb.RT_RAY_ALL_HITS, rtb.RT_COLLECT_VISITS)
        stats = tree.stats()
        _, _, visits = O.oracle_query_rays(flat, z["rays"], O.RAY_ALL, want_visits=True)
        assert (stats["visits_internal"], stats["visits_leaf"]) == (int(visits[0]), int(visits[1]))
        assert s_rays[1].size == z["ids"].size


def test_rays_vs_reference_search_with_degenerate_rays(rtb):
    rng = np.random.default_rng(11)
    c = rng.uniform(-1, 1, (30000, 3)).astype(np.float32)
    h = rng.uniform(0.0, 0.03, (30000, 3)).astype(np.float32)
    keys = np.stack([c - h, c + h], axis=1)
    nr = 20000
    o = rng.uniform(-3, 3, (nr, 3)).astype(np.float32)
    d = rng.normal(size=(nr, 3)).astype(np.float32)
    d[:500, 0] = 0.0          # axis-parallel: inv = inf, (lo - o) * inf, 0 * inf = NaN paths
    d[500:1000, 1] = -0.0
    d[1000:1200, :2] = 0.0
    o[1200:1400] = keys[:200, 0]   # origin exactly on a box corner
    with np.errstate(divide="ignore"):
        inv = (np.float32(1.0) / d).astype(np.float32)
    rays = np.concatenate([o, inv, np.full((nr, 1), 1e30, np.float32)], axis=1)
    rays[1400:2000, 6] = 1.5       # short rays
    ref = O.RefTree(5).insert(keys)
    want_all = ref.search_rays(rays, O.RAY_ALL)
    assert want_all[1].size > 10000
    with rtb.RTree(5).insert(keys).flatten_device().upload(0) as tree:
        check_all_modes(rtb, tree, rays, want_all, ref.search_rays(rays, O.RAY_CLOSEST),
                        ref.search_rays(rays, O.RAY_ANY))


def test_rays_on_point_keys_and_wide_blocks(rtb):
    """point-keyed leaves (one corner stored) and 16-wide... (kind 4: W=8 points)"""
    pts = rtb.workloads.uniform_points(20000, seed=3)
    rng = np.random.default_rng(4)
    target = pts[rng.integers(0, 20000, 5000)]
    o = rng.uniform(-1, 2, (5000, 3)).astype(np.float32)
    d = (target - o).astype(np.float32)      # aimed at data points: some exact hits
    with np.errstate(divide="ignore"):
        inv = (np.float32(1.0) / d).astype(np.float32)
    rays = np.concatenate([o, inv, np.full((5000, 1), 1e30, np.float32)], axis=1)
    ref = O.RefTree(4).insert(pts)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        check_all_modes(rtb, tree, rays, ref.search_rays(rays, O.RAY_ALL), ref.search_rays(rays, O.RAY_CLOSEST),
                        ref.search_rays(rays, O.RAY_ANY))


def test_ray_properties_at_scale(rtb):
    """BASELINE config 3 shape (teapot AABBs, random-direction rays) at 2 M rays: closest-hit is
    the minimum over all-hits, any-hit is its indicator, results are reproducible."""
    boxes = rtb.workloads.triangle_aabbs(rtb.workloads.teapot_triangles(2.0))
    rays = rtb.workloads.random_rays(2_000_000, boxes, seed=44, tmax=1e30)
    with rtb.RTree(5).insert(boxes).flatten_device().upload(0) as tree:
        off, ids = tree.query_rays(rays, rtb.RT_RAY_ALL_HITS)
        t, tid = tree.query_rays(rays, rtb.RT_RAY_CLOSEST)
        anyhit = tree.query_rays(rays, rtb.RT_RAY_ANY)
        t2, tid2 = tree.query_rays(rays, rtb.RT_RAY_CLOSEST)
    counts = np.diff(off.astype(np.int64))
    assert np.array_equal(anyhit.astype(bool), counts > 0)
    assert np.array_equal(tid != 0xFFFFFFFF, counts > 0)
    assert np.array_equal(t.view(np.uint32), t2.view(np.uint32)) and np.array_equal(tid, tid2)
    assert 0.15 < anyhit.mean() < 0.9
    # the closest id is one of the ray's all-hits ids, checked exactly on a sample with the oracle
    sample = np.random.default_rng(1).integers(0, len(rays), 20000)
    bt, bid = O.oracle_brute_rays(boxes, None, rays[sample], O.RAY_CLOSEST)
    assert np.array_equal(bt.view(np.uint32), t[sample].view(np.uint32)) and np.array_equal(bid, tid[sample])


def test_rays_need_a_3d_f32_tree(rtb):
    pts = rtb.workloads.config1_points(500)
    with rtb.RTree(2).insert(pts).flatten_device().upload(0) as tree:
        with pytest.raises(rtb.RtError) as err:
            tree.query_rays(np.zeros((1, 7), np.float32), rtb.RT_RAY_ANY)
        assert err.value.code == rtb.capi.RT_E_UNSUPPORTED


def test_pipelined_ray_batches_equal_single_batch(rtb):
    """ray calls with a host result are cut into batches whose copies overlap the traversal of their
    neighbours (closest / any: one kernel per batch; all-hits: the two-phase CSR pipeline); results,
    hit counts and visit counters equal the unpipelined call and the reference's search()"""
    boxes = rtb.workloads.triangle_aabbs(rtb.workloads.teapot_triangles(1.0))
    rays = rtb.workloads.random_rays(50_003, boxes, seed=45, tmax=1e30)
    ref = O.RefTree(5).insert(boxes)
    want_t, want_id = ref.search_rays(rays, O.RAY_CLOSEST)
    want_any = ref.search_rays(rays, O.RAY_ANY)
    want_all = ref.search_rays(rays, O.RAY_ALL)
    with rtb.RTree(5).insert(boxes).flatten_device().upload(0) as tree:
        visits = {}
        for batch in (0, 4096, 5001, 10000):
            tree.set_option(rtb.RT_OPT_PIPELINE_BATCH, batch)
            for _ in range(2):
                r = np.ascontiguousarray(rays)
                out = tree.query_rays_raw(r.ctypes.data, len(r), rtb.RT_RAY_CLOSEST, 0)
                assert tree.stats()["batches"] == (max(1, len(r) // (2 * batch)) if batch else 1)
                assert out.n_hit == int((want_id != 0xFFFFFFFF).sum())
                t, tid = tree.query_rays(rays, rtb.RT_RAY_CLOSEST)
                assert np.array_equal(t.view(np.uint32), want_t.view(np.uint32)) and np.array_equal(tid, want_id)
                assert np.array_equal(tree.query_rays(rays, rtb.RT_RAY_ANY), want_any)
                assert csr_equal(tree.query_rays(rays, rtb.RT_RAY_ALL_HITS), want_all)
                assert tree.stats()["batches"] == (max(1, len(r) // (2 * batch)) if batch else 1)
                off, none = tree.query_rays(rays, rtb.RT_RAY_ALL_HITS, rtb.RT_COUNT_ONLY)
                assert np.array_equal(off, want_all[0]) and none.size == 0
            # the visit counters exist for every mode and do not depend on how the call was cut
            for mode in (rtb.RT_RAY_ALL_HITS, rtb.RT_RAY_CLOSEST, rtb.RT_RAY_ANY):
                tree.query_rays(rays, mode, rtb.RT_COLLECT_VISITS)
                s = tree.stats()
                seen = (s["visits_internal"], s["visits_leaf"], s["algorithmic_bytes"])
                assert seen[0] >= len(rays) and seen[2] > 0
                assert visits.setdefault(mode, seen) == seen
        # pruning and the first-hit abort visit fewer nodes than collecting every hit
        assert visits[rtb.RT_RAY_ANY][0] <= visits[rtb.RT_RAY_CLOSEST][0] <= visits[rtb.RT_RAY_ALL_HITS][0]
