"""GPU parity tests for rt_query_boxes -- every call goes through the C ABI
(include/rtree_b200.h) into the sm_100a kernels and is compared bit-exactly with the oracle:
golden vectors made by the reference, the reference run live (oracle/_ref), the C restatement
and brute force.  Integer / index work: the bar is exact equality."""
import os

import numpy as np
import pytest

import oracles as O
from helpers import csr_equal

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BOX_FIXTURES = ["flatten_test_1d_i32", "sample_1d_f64", "config1_2d_f64", "config2_cut_3d_f32",
                "config4_cut_8d_f32", "teapot_boxes_intersects", "teapot_boxes_contains", "lattice_2d_i32"]


def upload_flat(rtb, kind, flat):
    image = rtb.DeviceImage.from_flat(kind, flat.leaf_level, flat.nodes, flat.bounds, flat.children, flat.data)
    return image.upload(0)


def sort_segments(off, ids):
    out = ids.copy()
    for i in range(len(off) - 1):
        out[off[i]:off[i + 1]].sort()
    return out


@pytest.mark.parametrize("name", BOX_FIXTURES)
def test_golden_vectors_all_pipelines(rtb, name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    flat = O.Flat.load(z)
    kind, mode, q = int(z["kind"]), int(z["mode"]), z["queries"]
    want = (z["offsets"], z["ids"])
    with upload_flat(rtb, kind, flat) as tree:
        assert csr_equal(tree.query_boxes(q, mode), want)                                   # stash pipeline
        assert tree.stats()["kernel_launches"] > 0
        assert csr_equal(tree.query_boxes(q, mode, rtb.RT_NO_REORDER), want)
        # two-pass pipeline with visit counters: on the quantised image (where the tree has one) the
        # conservative filter may visit a few nodes more than the reference's exact filter ...
        off, ids = tree.query_boxes(q, mode, rtb.RT_COLLECT_VISITS)
        assert csr_equal((off, ids), want)
        sq = tree.stats()
        _, _, visits = O.oracle_query_boxes(flat, q, mode, want_visits=True)
        exact = (int(visits[0]), int(visits[1]))
        assert sq["visits_internal"] >= exact[0] and sq["visits_leaf"] >= exact[1]
        assert sq["visits_internal"] + sq["visits_leaf"] <= 1.05 * (exact[0] + exact[1]) + 16
        assert sq["leaf_lookups"] <= 16 * sq["visits_leaf"]
        # ... and on the float blocks exactly the nodes the reference's traversal visits
        tree.set_option(rtb.RT_OPT_QUANTISED, 0)
        off, ids = tree.query_boxes(q, mode, rtb.RT_COLLECT_VISITS)
        tree.set_option(rtb.RT_OPT_QUANTISED, 1)
        assert csr_equal((off, ids), want)
        s = tree.stats()
        assert (s["visits_internal"], s["visits_leaf"]) == exact
        assert s["algorithmic_bytes"] > 0 and s["leaf_lookups"] == 0
        off, ids = tree.query_boxes(q, mode, rtb.RT_UNSORTED)
        assert np.array_equal(off, want[0]) and np.array_equal(sort_segments(off, ids), want[1])
        off, ids = tree.query_boxes(q, mode, rtb.RT_COUNT_ONLY)
        assert np.array_equal(off, want[0]) and ids.size == 0


@pytest.mark.parametrize("kind", sorted(O.KINDS))
def test_product_tree_vs_reference_search(rtb, kind):
    """tree built by the product's host library, queried on the GPU, against the reference's own
    RTree::search() on the reference's own tree (hit sets are topology independent)"""
    scalar, dim, is_box, _ = O.KINDS[kind]
    rng = np.random.default_rng(200 + kind)
    n, nq = 20000, 6000
    if scalar == O.I32:
        keys = rng.integers(0, 300, (n, dim)).astype(np.int32)
        lo = rng.integers(-5, 300, (nq, dim))
        q = np.stack([lo, lo + rng.integers(0, 9, (nq, dim))], axis=1).astype(np.int32)
    else:
        dt = O.SCALAR_DTYPES[scalar]
        c = rng.random((n, dim))
        if is_box:
            h = rng.uniform(0, 0.01, (n, dim))
            keys = np.stack([c - h, c + h], axis=1).astype(dt)
        else:
            keys = c.astype(dt)
        qc = rng.random((nq, dim))
        qh = 0.5 * (10.0 / n) ** (1.0 / dim) * rng.uniform(0.3, 3.0, (nq, 1))
        q = np.stack([qc - qh, qc + qh], axis=1).astype(dt)
    modes = [rtb.RT_INTERSECTS, rtb.RT_CONTAINS] if is_box else [rtb.RT_POINT_IN_BOX]
    ref = O.RefTree(kind).insert(keys)
    with rtb.RTree(kind).insert(keys).flatten_device().upload(0) as tree:
        for mode in modes:
            want = ref.search_boxes(q, mode)
            assert want[1].size > 0
            assert csr_equal(tree.query_boxes(q, mode), want)
            assert csr_equal(tree.query_boxes(q, mode, rtb.RT_COLLECT_VISITS), want)


def test_point_tree_accepts_every_mode_box_tree_rejects_point_mode(rtb):
    pts = rtb.workloads.uniform_points(5000, seed=1)
    q = rtb.workloads.cube_queries(500, 5000, seed=2)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        a = tree.query_boxes(q, rtb.RT_POINT_IN_BOX)
        assert csr_equal(a, tree.query_boxes(q, rtb.RT_INTERSECTS))
        assert csr_equal(a, tree.query_boxes(q, rtb.RT_CONTAINS))
    boxes = np.stack([pts, pts + 0.01], axis=1)
    with rtb.RTree(5).insert(boxes).flatten_device().upload(0) as tree:
        with pytest.raises(rtb.RtError) as err:
            tree.query_boxes(q, rtb.RT_POINT_IN_BOX)
        assert err.value.code == rtb.capi.RT_E_INVALID


def test_edge_cases_nan_inverted_everything(rtb):
    rng = np.random.default_rng(9)
    pts = rng.random((3000, 3)).astype(np.float32)
    q = np.zeros((7, 2, 3), np.float32)
    q[0] = [[0.2, 0.2, 0.2], [0.4, np.nan, 0.4]]
    q[1] = [[np.nan, 0.1, 0.1], [0.3, 0.3, 0.3]]
    q[2] = [[0.6, 0.6, 0.6], [0.5, 0.5, 0.5]]
    q[3] = [pts[17], pts[17]]
    q[4] = [[-np.inf] * 3, [np.inf] * 3]
    q[5] = [[np.nan] * 3, [np.nan] * 3]
    q[6] = [[2, 2, 2], [3, 3, 3]]
    want = O.RefTree(4).insert(pts).search_boxes(q)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        assert csr_equal(tree.query_boxes(q), want)
        assert csr_equal(tree.query_boxes(q, flags=rtb.RT_COLLECT_VISITS), want)
        # a large batch takes the reorder path even with NaN / inf centres in it
        big = np.concatenate([q, rtb.workloads.cube_queries(9000, 3000, seed=5)])
        want_big = O.RefTree(4).insert(pts).search_boxes(big)
        assert csr_equal(tree.query_boxes(big), want_big)


def test_empty_tree_root_leaf_tree_and_empty_batch(rtb):
    q = rtb.workloads.cube_queries(10, 10, seed=2)
    with rtb.RTree(4).flatten_device().upload(0) as tree:                   # appendix B4
        off, ids = tree.query_boxes(q)
        assert off.tolist() == [0] * 11 and ids.size == 0
    pts = np.array([[0.1, 0.1, 0.1], [0.5, 0.5, 0.5], [0.9, 0.9, 0.9]], np.float32)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:      # root is a leaf
        off, ids = tree.query_boxes(np.array([[[0.4, 0.4, 0.4], [1, 1, 1]], [[0, 0, 0], [0.2, 0.2, 0.2]]], np.float32))
        assert off.tolist() == [0, 2, 3] and ids.tolist() == [1, 2, 0]
        off, ids = tree.query_boxes(np.zeros((0, 2, 3), np.float32))       # n == 0
        assert off.tolist() == [0] and ids.size == 0
        off, ids = tree.query_boxes(np.array([[[0, 0, 0], [1, 1, 1]]], np.float32))
        assert ids.tolist() == [0, 1, 2]


def test_long_hit_lists_take_the_deferred_and_sort_paths(rtb):
    """lists beyond the 32-id hit buffer spill into the stash slab (here ~200 ids) or, beyond the slab
    (9000 ids, the whole tree), take the second traversal; 33..8192 ids are sorted in shared memory,
    longer lists in global memory; the whole-tree query returns every id exactly once"""
    n = 40000
    pts = rtb.workloads.uniform_points(n, seed=3)
    ref = O.RefTree(4).insert(pts)
    q = np.concatenate([rtb.workloads.cube_queries(300, n, seed=4, hits_per_query=200.0),
                        rtb.workloads.cube_queries(40, n, seed=5, hits_per_query=9000.0),
                        np.array([[[-1, -1, -1], [2, 2, 2]]], np.float32),
                        rtb.workloads.cube_queries(300, n, seed=6, hits_per_query=5.0)])
    want = ref.search_boxes(q)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        got = tree.query_boxes(q)
        assert csr_equal(got, want)
        s = tree.stats()
        assert 41 <= s["deferred_queries"] < 341          # the 9000-id lists and the whole tree; not the ~200-id ones
        whole = got[1][got[0][340]:got[0][341]]
        assert np.array_equal(whole, np.arange(n))
        assert csr_equal(tree.query_boxes(q, flags=rtb.RT_COLLECT_VISITS), want)


def test_huge_hit_lists_are_sorted_by_the_whole_gpu(rtb):
    """queries that cover much of a 500 k-point tree (lists of 500 k, ~250 k, ~60 k and ~10 k ids, next
    to ordinary ones): lists beyond 8192 ids are sorted together by a global-memory network, one launch
    per step over all their elements, instead of one thread block per list -- still the reference's
    sorted lists, in single-batch, pipelined and two-pass calls, and in well under a second"""
    n = 500_000
    pts = rtb.workloads.uniform_points(n, seed=31)
    ref = O.RefTree(4).insert(pts)
    small = rtb.workloads.cube_queries(9000, n, seed=32)
    big = np.array([[[-1, -1, -1], [2, 2, 2]],
                    [[0, 0, 0], [0.5, 1, 1]],
                    [[0.2, 0.2, 0.2], [0.7, 0.7, 0.7]],
                    [[0.6, 0.1, 0.3], [0.9, 0.4, 0.52]]], np.float32)
    q = np.concatenate([small[:3000], big[:2], small[3000:6000], big[2:], small[6000:]])
    want = ref.search_boxes(q)
    lengths = np.diff(want[0].astype(np.int64))
    assert lengths.max() == n and (lengths > 8192).sum() == 4
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        for batch, flags in ((0, 0), (4096, 0), (0, rtb.RT_COLLECT_VISITS)):
            tree.set_option(rtb.RT_OPT_PIPELINE_BATCH, batch)
            got = tree.query_boxes(q, flags=flags)
            assert csr_equal(got, want)
            assert tree.stats()["finish_ms"] < 500.0
            assert tree.stats()["kernel_launches"] > 150       # the steps of the network are launches of their own


def test_stash_exhaustion_falls_back_to_the_second_traversal(rtb, monkeypatch):
    n = 30000
    pts = rtb.workloads.uniform_points(n, seed=7)
    q = rtb.workloads.cube_queries(20000, n, seed=8)
    want = O.RefTree(4).insert(pts).search_boxes(q)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        monkeypatch.setenv("RTB_STASH_LIMIT_IDS", "4096")
        got = tree.query_boxes(q)
        assert tree.stats()["deferred_queries"] > 1000
        assert csr_equal(got, want)
        monkeypatch.setenv("RTB_STASH_LIMIT_IDS", "0")
        got = tree.query_boxes(q)
        assert tree.stats()["deferred_queries"] >= int((np.diff(want[0].astype(np.int64)) > 0).sum())
        assert csr_equal(got, want)
        monkeypatch.delenv("RTB_STASH_LIMIT_IDS")
        assert csr_equal(tree.query_boxes(q), want)
        assert tree.stats()["deferred_queries"] == 0


def test_device_resident_queries_and_results(rtb):
    """RT_QUERIES_ON_DEVICE | RT_RESULT_ON_DEVICE with torch owning the memory and the stream"""
    import ctypes as C
    import torch
    n = 20000
    pts = rtb.workloads.uniform_points(n, seed=1)
    q = rtb.workloads.cube_queries(30000, n, seed=2)
    want = O.RefTree(4).insert(pts).search_boxes(q)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        stream = torch.cuda.Stream()
        tree.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            dq = torch.from_numpy(q).cuda(non_blocking=False)
            stream.synchronize()
            out = tree.query_boxes_raw(dq.data_ptr(), len(q), rtb.RT_POINT_IN_BOX,
                                       rtb.RT_QUERIES_ON_DEVICE | rtb.RT_RESULT_ON_DEVICE)
            assert out.memory == rtb.capi.RT_MEM_DEVICE and out.total == len(want[1])
            # wrap the engine-owned device buffers without copying (__cuda_array_interface__)
            off = torch.as_tensor(rtb.capi.DeviceBuffer(out.offsets, (len(q) + 1) * 8), device="cuda").view(torch.int64)
            ids = torch.as_tensor(rtb.capi.DeviceBuffer(out.ids, out.total * 4), device="cuda").view(torch.int32)
        assert np.array_equal(off.cpu().numpy().astype(np.uint64), want[0])
        assert np.array_equal(ids.cpu().numpy().view(np.uint32), want[1])
        tree.set_stream(0)


def test_query_sharding_emulated_on_one_gpu(rtb):
    """R contiguous shards answered separately concatenate to the 1-shard answer"""
    n, nq = 50000, 40001
    pts = rtb.workloads.uniform_points(n, seed=11)
    q = rtb.workloads.cube_queries(nq, n, seed=12)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        full = tree.query_boxes(q)
        for world in (2, 4, 8):
            parts = []
            for r in range(world):
                b, e = rtb.sharding.shard_range(nq, world, r)
                parts.append(tree.query_boxes(q[b:e]))
            assert csr_equal(rtb.sharding.concat_csr(parts), full)
        # a second handle (replica) on the same device gives the same answer
        image = rtb.RTree(4).insert(pts).flatten_device()
        with image.upload(0) as replica:
            assert csr_equal(replica.query_boxes(q), full)
            blocks = replica.device_blocks()
            assert blocks.nbytes == image.desc.blocks_bytes


def test_config2_scale_properties(rtb):
    """BASELINE config 2 at a few million points / queries, checked through size-independent
    properties: both pipelines agree, per-query lists are strictly ascending, every reported
    point lies in its box, the count equals a brute-force count on a sample, and the result
    is reproducible."""
    n, nq = 2_000_000, 3_000_000
    pts = rtb.workloads.uniform_points(n, seed=42)
    q = rtb.workloads.cube_queries(nq, n, seed=43)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        off, ids = tree.query_boxes(q)
        assert tree.stats()["deferred_queries"] == 0
        off2, ids2 = tree.query_boxes(q, flags=rtb.RT_COLLECT_VISITS)
        assert np.array_equal(off, off2) and np.array_equal(ids, ids2)
        off3, ids3 = tree.query_boxes(q)
        assert np.array_equal(off, off3) and np.array_equal(ids, ids3)
    counts = np.diff(off.astype(np.int64))
    assert 9.0 < counts.mean() < 11.0
    owner = np.repeat(np.arange(nq), counts)
    p = pts[ids]
    assert np.all((p >= q[owner, 0]) & (p <= q[owner, 1]))                 # every hit is inside
    same_owner = owner[1:] == owner[:-1]
    assert np.all(ids[1:][same_owner] > ids[:-1][same_owner])             # ascending, no duplicates
    sample = np.random.default_rng(0).integers(0, nq, 300)
    for i in sample:                                                       # nothing is missing
        inside = np.all((pts >= q[i, 0]) & (pts <= q[i, 1]), axis=1)
        assert np.array_equal(np.nonzero(inside)[0].astype(np.uint32), ids[off[i]:off[i + 1]])


@pytest.mark.parametrize("flags", [0, "visits", "unsorted", "count_only", "no_reorder"])
def test_pipelined_call_equals_single_batch(rtb, flags):
    """a host-result call cut into batches that rotate through the working sets (H2D / traversal /
    D2H of neighbouring batches overlapped) returns the very CSR of the unpipelined call, for every
    pipeline flavour, ragged last batch and a batch count that is not a multiple of the slots"""
    flag = {0: 0, "visits": rtb.RT_COLLECT_VISITS, "unsorted": rtb.RT_UNSORTED,
            "count_only": rtb.RT_COUNT_ONLY, "no_reorder": rtb.RT_NO_REORDER}[flags]
    n, nq = 60000, 47003
    pts = rtb.workloads.uniform_points(n, seed=11)
    q = rtb.workloads.cube_queries(nq, n, seed=12)
    q[1000:1040, 1] += 0.2  # a run of long hit lists (deferred + segment sort) inside one batch
    ref = O.RefTree(4).insert(pts).search_boxes(q)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        tree.set_option(rtb.RT_OPT_PIPELINE_BATCH, 0)
        single = tree.query_boxes(q, flags=flag)
        assert tree.stats()["batches"] == 1
        single_visits = (tree.stats()["visits_internal"], tree.stats()["visits_leaf"])
        for batch in (5000, 6500, 20000):
            tree.set_option(rtb.RT_OPT_PIPELINE_BATCH, batch)
            for _ in range(2):  # the second call reuses every buffer the first one grew
                got = tree.query_boxes(q, flags=flag)
                s = tree.stats()
                assert s["batches"] >= 2 and (s["batches"] == nq // batch or nq // batch >= 6)
                assert np.array_equal(got[0], single[0]) and np.array_equal(got[1], single[1])
                if flags == "visits":
                    assert (s["visits_internal"], s["visits_leaf"]) == single_visits
        if flags == "unsorted":
            assert np.array_equal(single[0], ref[0]) and np.array_equal(sort_segments(*single), ref[1])
        elif flags == "count_only":
            assert np.array_equal(single[0], ref[0])
        else:
            assert csr_equal(single, ref)


def test_counts32_result_and_failed_call_leaves_the_handle_usable(rtb, monkeypatch):
    """RT_RESULT_COUNTS32 delivers 32-bit hits per query (whose prefix sums are the offsets) with the
    same ids, single batch and pipelined.  A call that fails in the middle of its pipeline (here: the
    pinned id buffer is not allowed to grow) reports RT_E_NOMEM, and the next call on the handle is
    answered as if nothing had happened."""
    n, nq = 60000, 47003
    pts = rtb.workloads.uniform_points(n, seed=21)
    q = rtb.workloads.cube_queries(nq, n, seed=22)
    q[2000:2030, 1] += 0.2
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        tree.set_option(rtb.RT_OPT_PIPELINE_BATCH, 0)
        off, ids = tree.query_boxes(q)
        for batch in (0, 5000):
            tree.set_option(rtb.RT_OPT_PIPELINE_BATCH, batch)
            counts, ids32 = tree.query_boxes(q, flags=rtb.RT_RESULT_COUNTS32)
            assert counts.dtype == np.uint32 and np.array_equal(np.cumsum(counts, dtype=np.uint64), off[1:])
            assert np.array_equal(ids32, ids)
            counts, none = tree.query_boxes(q, flags=rtb.RT_RESULT_COUNTS32 | rtb.RT_COUNT_ONLY)
            assert np.array_equal(np.cumsum(counts, dtype=np.uint64), off[1:]) and none.size == 0
        # values the planner cannot work with are refused, not divided by
        for bad in (1, 7, 4095):
            with pytest.raises(rtb.RtError) as e:
                tree.set_option(rtb.RT_OPT_PIPELINE_BATCH, bad)
            assert e.value.code == rtb.capi.RT_E_INVALID
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:  # a fresh handle: no pinned ids yet
        tree.set_option(rtb.RT_OPT_PIPELINE_BATCH, 5000)
        monkeypatch.setenv("RTB_HOST_IDS_LIMIT", str(len(ids) // 3))
        with pytest.raises(rtb.RtError) as e:
            tree.query_boxes(q)
        assert e.value.code == rtb.capi.RT_E_NOMEM
        monkeypatch.delenv("RTB_HOST_IDS_LIMIT")
        for _ in range(2):
            assert csr_equal(tree.query_boxes(q), (off, ids))
        tree.set_option(rtb.RT_OPT_PIPELINE_BATCH, 0)
        assert csr_equal(tree.query_boxes(q), (off, ids))


def _boundary_queries(pts, rng, n):
    """query boxes whose faces pass exactly through data points (closed intervals: those points are hits)"""
    a = pts[rng.integers(0, len(pts), n)]
    b = pts[rng.integers(0, len(pts), n)]
    return np.stack([np.minimum(a, b), np.maximum(a, b)], axis=1).astype(np.float32)


QUANT_CASES = ["uniform", "flat_axis", "huge_range", "tiny_range", "negative_denormal", "single_leaf",
               "single_leaf_extreme", "duplicates"]
QUANT_KINDS = {4: QUANT_CASES, 3: ["uniform", "flat_axis", "tiny_range", "single_leaf_extreme"],
               6: ["uniform", "flat_axis", "tiny_range", "single_leaf", "duplicates"]}


@pytest.mark.parametrize("kind,case", [(k, c) for k, cases in QUANT_KINDS.items() for c in cases])
def test_quantised_image_gives_the_exact_hit_sets(rtb, kind, case):
    """float32 point trees (3-D and 2-D with 8-wide blocks, 8-D with 16-wide ones) are traversed on a
    quantised image (conservative 15-bit bounds, leaf candidates confirmed on the float records): hit
    lists must equal the float-block path and the reference's search() bit for bit, also where
    quantisation is at its worst"""
    _, dim, _, _ = O.KINDS[kind]
    rng = np.random.default_rng(700 + QUANT_CASES.index(case) + 50 * kind)
    n = 40000 if dim <= 3 else 25000
    pts = rng.random((n, dim)).astype(np.float32)
    if dim == 8:
        pts = (rng.random((64, dim))[rng.integers(0, 64, n)] + rng.normal(0, 0.02, (n, dim))).astype(np.float32)
    if case == "flat_axis":
        pts[:, 1] = np.float32(0.25)                       # zero extent on one axis
    elif case == "huge_range":
        # (the reference's insert segfaults once bound areas overflow float32, i.e. beyond ~7e12 per
        # axis -- stay below; an overflowing extent is exercised by single_leaf_extreme)
        pts = (pts * np.float32(2e12) - np.float32(1e12)).astype(np.float32)
    elif case == "tiny_range":
        pts = (np.float32(1000.0) + pts * np.float32(1e-3)).astype(np.float32)   # few ulps between keys
    elif case == "negative_denormal":
        pts = (pts - np.float32(0.5)).astype(np.float32)
        pts[:200] *= np.float32(1e-42)                     # denormals around zero
    elif case == "single_leaf":
        pts = pts[:6]
    elif case == "single_leaf_extreme":
        pts = pts[:6]
        pts[0] = np.where(np.arange(dim) % 2 == 0, 3e38, -3e38)   # root extent overflows to +inf: scale 0
        pts[1] = -pts[0]
    elif case == "duplicates":
        pts = pts[rng.integers(0, 50, n)]                  # 50 distinct points, long equal runs
    with np.errstate(over="ignore", invalid="ignore"):
        span = pts.max(axis=0) - pts.min(axis=0)
        c = pts[rng.integers(0, len(pts), 3000)]
        width = 0.02 if dim <= 3 else 0.12
        h = (np.where(np.isfinite(span), span, np.float32(1e30)) * np.float32(width)).astype(np.float32)
        q = np.concatenate([
            np.stack([c - h, c + h], axis=1).astype(np.float32),
            _boundary_queries(pts, rng, 3000),
            np.stack([c, c], axis=1),                                     # zero-size boxes on data points
            np.stack([c + h, c - h], axis=1).astype(np.float32)[:200],    # inverted
        ])
    special = np.zeros((8, 2, dim), np.float32)
    special[0] = [[-np.inf] * dim, [np.inf] * dim]
    special[1] = [[np.nan] * dim, [np.nan] * dim]
    special[2] = [pts.min(axis=0), pts.max(axis=0)]                       # exactly the root box
    special[3] = [pts.max(axis=0), [np.inf] * dim]                        # touches the root box in one corner
    special[4] = [[-np.inf] * dim, pts.min(axis=0)]
    special[5] = [[-np.inf] * dim, [np.inf] * dim]
    special[5][0, 0], special[5][1, 1] = np.nan, np.nan
    special[5][:, dim - 1] = pts[0, dim - 1]                              # NaN / inf bounds + a zero-width axis
    special[6] = [pts.max(axis=0) + np.float32(1.0), pts.max(axis=0) + np.float32(2.0)]   # outside
    special[7] = [[-3e38] * dim, [3e38] * dim]
    q = np.concatenate([q, special])
    want = O.RefTree(kind).insert(pts).search_boxes(q)
    assert want[1].size > 0
    with rtb.RTree(kind).insert(pts).flatten_device().upload(0) as tree:
        for flags in (0, rtb.RT_NO_REORDER, rtb.RT_UNSORTED, rtb.RT_COUNT_ONLY):
            got = {}
            for quantised in (1, 0):
                tree.set_option(rtb.RT_OPT_QUANTISED, quantised)
                got[quantised] = tree.query_boxes(q, flags=flags)
            assert np.array_equal(got[1][0], got[0][0]) and np.array_equal(got[1][0], want[0])
            if flags == rtb.RT_UNSORTED:
                assert np.array_equal(sort_segments(*got[1]), want[1]) and np.array_equal(sort_segments(*got[0]), want[1])
            elif flags != rtb.RT_COUNT_ONLY:
                assert np.array_equal(got[1][1], want[1]) and np.array_equal(got[0][1], want[1])


def test_reloaded_image_answers_like_the_original(rtb, tmp_path):
    """an image saved with save_device_tree and loaded back uploads and answers identically"""
    pts = rtb.workloads.uniform_points(30000, seed=21)
    q = rtb.workloads.cube_queries(5000, 30000, seed=22)
    image = rtb.RTree(4).insert(pts).flatten_device()
    path = str(tmp_path / "tree.rtb")
    image.save(path)
    with image.upload(0) as a, rtb.DeviceImage.load(path).upload(0) as b:
        assert csr_equal(a.query_boxes(q), b.query_boxes(q))


def test_heavy_hit_lists_at_scale(rtb):
    """every list far beyond the 32-id hit buffer (~300 hits per query, 100 k queries, 30 M hits): they
    spill into the warps' stash slabs (or, when a slab is too small, are written by a second traversal),
    the segment sort orders them, pipelined or not; a sample is checked against the reference's
    search(), the whole batch through sortedness and the count-only totals"""
    n, nq = 1_000_000, 100_000
    pts = rtb.workloads.uniform_points(n, seed=31)
    q = rtb.workloads.cube_queries(nq, n, seed=32, hits_per_query=300.0)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:
        tree.set_option(rtb.RT_OPT_PIPELINE_BATCH, 0)
        off, ids = tree.query_boxes(q)
        first = tree.stats()["deferred_queries"]
        again = tree.query_boxes(q)                       # slab size now follows the lists seen
        assert np.array_equal(again[0], off) and np.array_equal(again[1], ids)
        assert tree.stats()["deferred_queries"] <= first and tree.stats()["deferred_queries"] < 0.01 * nq
        tree.set_option(rtb.RT_OPT_PIPELINE_BATCH, 10_000)
        off2, ids2 = tree.query_boxes(q)
        assert tree.stats()["batches"] > 2
        assert np.array_equal(off, off2) and np.array_equal(ids, ids2)
        coff, _ = tree.query_boxes(q, flags=rtb.RT_COUNT_ONLY)
        assert np.array_equal(coff, off)
    counts = np.diff(off.astype(np.int64))
    assert 250 < counts.mean() < 350 and ids.size == int(off[-1])
    # ascending inside every segment: a descent can only sit on a segment boundary
    descents = np.nonzero(np.diff(ids.astype(np.int64)) < 0)[0] + 1
    assert np.all(np.isin(descents, off[1:-1].astype(np.int64)))
    sample = np.random.default_rng(3).integers(0, nq, 400)
    want = O.RefTree(4).insert(pts).search_boxes(q[sample])
    for k, qi in enumerate(sample):
        assert np.array_equal(ids[off[qi]:off[qi + 1]], want[1][want[0][k]:want[0][k + 1]])


@pytest.mark.parametrize("kind", [4, 5, 2, 6])
def test_any_hit_is_the_abort_semantics(rtb, kind):
    """RT_ANY_HIT stops a query at its first hit (functor returns true, reference rtree.hpp:994-997):
    the per-query boolean equals "the reference's search() delivered at least one value", on the
    quantised and the float path, pipelined or not"""
    scalar, dim, is_box, _ = O.KINDS[kind]
    dt = O.SCALAR_DTYPES[scalar]
    rng = np.random.default_rng(300 + kind)
    n, nq = 30000, 12000
    c = rng.random((n, dim))
    keys = np.stack([c, c + rng.uniform(0, 0.01, (n, dim))], axis=1).astype(dt) if is_box else c.astype(dt)
    qc = rng.uniform(-0.3, 1.3, (nq, dim))                       # a good share of the boxes hit nothing
    qh = 0.5 * (2.0 / n) ** (1.0 / dim) * rng.uniform(0.2, 2.0, (nq, 1))
    q = np.stack([qc - qh, qc + qh], axis=1).astype(dt)
    mode = rtb.RT_INTERSECTS if is_box else rtb.RT_POINT_IN_BOX
    want = np.diff(O.RefTree(kind).insert(keys).search_boxes(q, mode)[0].astype(np.int64)) > 0
    assert 0.1 < want.mean() < 0.9
    with rtb.RTree(kind).insert(keys).flatten_device().upload(0) as tree:
        for quantised in (1, 0):
            tree.set_option(rtb.RT_OPT_QUANTISED, quantised)
            for batch in (0, 4096):
                tree.set_option(rtb.RT_OPT_PIPELINE_BATCH, batch)
                off, ids = tree.query_boxes(q, mode, rtb.RT_ANY_HIT)
                assert tree.stats()["batches"] == (2 if batch else 1)
                assert ids.size == 0 and np.array_equal(np.diff(off.astype(np.int64)), want.astype(np.int64))
        with pytest.raises(rtb.RtError):
            tree.query_boxes(q, mode, rtb.RT_ANY_HIT | rtb.RT_COLLECT_VISITS)


@pytest.mark.parametrize("kind", [8, 13, 4, 5, 7, 2, 6, 0])
def test_strict_overlap_flag_equals_the_reference_traits(rtb, kind):
    """RT_STRICT_OVERLAP: the node filter (and the key test of box keys) is the reference's own
    geometry_traits<aabb>::is_overlap(aabb, aabb) (include/RTree/aabb.hpp:225-236) -- strict
    inequalities, so touching boxes do not overlap; 1-D trees: closed, as in the reference.  On lattice
    coordinates, against the reference's search() with those traits on the same tree, for int32 /
    float32 / float64, point and box keys, 8-wide and 16-wide blocks, pipelined and not."""
    scalar, dim, is_box, _ = O.KINDS[kind]
    dt = O.SCALAR_DTYPES[scalar]
    rng = np.random.default_rng(500 + kind)
    n, nq = 20000, 9000
    side = {1: 15000, 2: 130, 3: 26, 8: 4}[dim]
    c = rng.integers(0, side, (n, dim))
    keys = np.stack([c, c + rng.integers(0, 3, (n, dim))], axis=1).astype(dt) if is_box else c.astype(dt)
    lo = rng.integers(-1, side, (nq, dim))
    q = np.stack([lo, lo + rng.integers(0, 4, (nq, dim))], axis=1).astype(dt)
    ref = O.RefTree(kind).insert(keys)
    flat = ref.flatten()
    image = rtb.DeviceImage.from_flat(kind, flat.leaf_level, flat.nodes, flat.bounds, flat.children, flat.data)
    modes = [(rtb.RT_INTERSECTS, O.MODE_INTERSECTS), (rtb.RT_CONTAINS, O.MODE_CONTAINS)] if is_box \
        else [(rtb.RT_POINT_IN_BOX, O.MODE_POINT_IN_BOX)]
    with image.upload(0) as tree:
        for mode, omode in modes:
            want = ref.search_boxes(q, omode | O.MODE_STRICT)
            closed = ref.search_boxes(q, omode)
            for batch in (0, 4096):
                tree.set_option(rtb.RT_OPT_PIPELINE_BATCH, batch)
                assert csr_equal(tree.query_boxes(q, mode, rtb.RT_STRICT_OVERLAP), want)
                assert csr_equal(tree.query_boxes(q, mode), closed)        # and the default is untouched
            off, none = tree.query_boxes(q, mode, rtb.RT_STRICT_OVERLAP | rtb.RT_COUNT_ONLY)
            assert np.array_equal(off, want[0])
            if dim > 1:
                assert want[1].size < closed[1].size


@pytest.mark.parametrize("kind,dim", [(4, 3), (3, 2)])
def test_float_lattice_known_answer(rtb, kind, dim):
    """the reference's Flatten known-answer test (test/rtree.cpp:410-467: a closed range over integer
    points returns exactly min..max) lifted to float32 lattices: every query face passes exactly
    through lattice points, the hit count is the product of the closed index ranges -- on the
    quantised image and on the float blocks"""
    side = 24 if dim == 3 else 90
    axes = np.meshgrid(*[np.arange(side)] * dim, indexing="ij")
    pts = np.stack([a.ravel() for a in axes], axis=1).astype(np.float32) * np.float32(0.125)
    rng = np.random.default_rng(77)
    lo = rng.integers(-2, side, (4000, dim))
    hi = lo + rng.integers(0, 7, (4000, dim))
    q = np.stack([lo, hi], axis=1).astype(np.float32) * np.float32(0.125)
    want_counts = np.prod(np.clip(np.minimum(hi, side - 1) - np.maximum(lo, 0) + 1, 0, None), axis=1)
    with rtb.RTree(kind).insert(pts).flatten_device().upload(0) as tree:
        for quantised in (1, 0):
            tree.set_option(rtb.RT_OPT_QUANTISED, quantised)
            off, ids = tree.query_boxes(q)
            assert np.array_equal(np.diff(off.astype(np.int64)), want_counts)
            k = int(np.argmax(want_counts))
            got = pts[ids[off[k]:off[k + 1]]]
            assert np.all(got >= q[k, 0]) and np.all(got <= q[k, 1])
