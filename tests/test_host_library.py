"""CPU tests of the header-only host library (r-star-tree_b200/include/RTree.hpp) through the
host bridge: the build side produces valid trees, flatten() is field-for-field the reference's
format, search() returns the reference's hit sets, and flatten_device() lays the tree out as
include/rtree_b200.h documents."""
import numpy as np
import pytest

import oracles as O
from helpers import decode_image, csr_equal, node_link, record_words, block_bytes

needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/libref_oracle.so not built")


def make_keys(kind, n, seed):
    scalar, dim, is_box, _ = O.KINDS[kind]
    rng = np.random.default_rng(seed)
    if scalar == O.I32:
        return rng.integers(0, 200, (n, dim)).astype(np.int32)
    if is_box:
        c = rng.random((n, dim))
        h = rng.uniform(0.0, 0.02, (n, dim))
        return np.stack([c - h, c + h], axis=1).astype(O.SCALAR_DTYPES[scalar])
    return rng.random((n, dim)).astype(O.SCALAR_DTYPES[scalar])


def make_queries(kind, nq, n, seed):
    scalar, dim, _, _ = O.KINDS[kind]
    rng = np.random.default_rng(seed)
    if scalar == O.I32:
        lo = rng.integers(-5, 200, (nq, dim))
        return np.stack([lo, lo + rng.integers(0, 40, (nq, dim))], axis=1).astype(np.int32)
    c = rng.random((nq, dim))
    h = 0.5 * (12.0 / n) ** (1.0 / dim) * rng.uniform(0.5, 2.0, (nq, 1))
    return np.stack([c - h, c + h], axis=1).astype(O.SCALAR_DTYPES[scalar])


@pytest.mark.parametrize("kind", sorted(O.KINDS))
def test_insert_keeps_the_structural_invariants(rtb, kind):
    """the invariants reference test/rtree.cpp:32-95 asserts after every insert"""
    keys = make_keys(kind, 1200, kind)
    tree = rtb.RTree(kind)
    for start in range(0, 1200, 150):
        tree.insert(keys[start:start + 150], first_id=start)
        assert tree.check_invariants() == 0
        assert tree.size() == start + 150
    flat = tree.flatten()
    assert sorted(flat["data"].tolist()) == list(range(1200))
    assert flat["root"] == 0 and flat["nodes"][0, 2] == 0
    sizes = flat["nodes"][1:, 1]
    m = 8 if O.KINDS[kind][3] == 16 else 4
    assert sizes.min() >= m and sizes.max() <= O.KINDS[kind][3]


@needs_ref
@pytest.mark.parametrize("kind", sorted(O.KINDS))
def test_flatten_is_bit_identical_to_the_reference(rtb, kind):
    """same inputs -> the same tree and the same four flatten vectors as the reference"""
    keys = make_keys(kind, 3000, 50 + kind)
    mine = rtb.RTree(kind).insert(keys).flatten()
    ref = O.RefTree(kind).insert(keys).flatten()
    assert mine["leaf_level"] == ref.leaf_level
    assert np.array_equal(mine["nodes"], ref.nodes)
    assert np.array_equal(mine["children"], ref.children)
    assert np.array_equal(mine["data"], ref.data)
    assert np.array_equal(mine["children_bound"].view(np.uint8), ref.bounds.view(np.uint8))


@needs_ref
@pytest.mark.parametrize("kind", [0, 2, 4, 5, 6, 8])
def test_host_search_matches_reference_search(rtb, kind):
    keys = make_keys(kind, 4000, 70 + kind)
    q = make_queries(kind, 400, 4000, 80 + kind)
    mode = O.MODE_INTERSECTS if O.KINDS[kind][2] else O.MODE_POINT_IN_BOX
    mine = rtb.RTree(kind).insert(keys).search_boxes(q, mode)
    ref = O.RefTree(kind).insert(keys).search_boxes(q, mode)
    assert csr_equal(mine, ref) and ref[1].size > 0


def test_erase_clear_rebalance(rtb):
    """reference test/rtree.cpp:98-259 (Erase / Clear): invariants hold while deleting"""
    keys = make_keys(4, 2000, 5)
    tree = rtb.RTree(4).insert(keys)
    rng = np.random.default_rng(6)
    order = rng.permutation(2000).astype(np.uint32)
    alive = set(range(2000))
    for start in range(0, 2000, 250):
        batch = order[start:start + 250]
        assert tree.erase_ids(batch) == len(batch)
        alive -= set(batch.tolist())
        assert tree.check_invariants() == 0
        assert tree.size() == len(alive)
        assert sorted(tree.flatten()["data"].tolist()) == sorted(alive)
    assert tree.size() == 0 and tree.leaf_level() == 0
    tree.insert(keys[:100])
    tree.rebalance()
    assert tree.size() == 100 and tree.check_invariants() == 0
    tree.clear()
    assert tree.size() == 0 and tree.leaf_level() == 0
    flat = tree.flatten()
    assert len(flat["nodes"]) == 1 and flat["nodes"][0, 1] == 0


def _same_flat(mine, ref):
    return (mine["leaf_level"] == ref.leaf_level and np.array_equal(mine["nodes"], ref.nodes)
            and np.array_equal(mine["children"], ref.children) and np.array_equal(mine["data"], ref.data)
            and np.array_equal(mine["children_bound"].view(np.uint8), ref.bounds.view(np.uint8)))


@needs_ref
@pytest.mark.parametrize("kind", sorted(O.KINDS))
def test_erase_gives_the_reference_tree(rtb, kind):
    """erase(iterator) against the reference's own (include/RTree/rtree.hpp:372-457): after every batch of
    deletions -- under-full nodes dissolved, their content re-inserted, the root collapsed -- flatten() is
    bit-identical, down to the empty tree; then the same for inserts into the thinned tree and rebalance()"""
    n = 900
    keys = make_keys(kind, n, 200 + kind)
    mine, ref = rtb.RTree(kind).insert(keys), O.RefTree(kind).insert(keys)
    order = np.random.default_rng(kind).permutation(n).astype(np.uint32)
    for start in range(0, 600, 150):
        batch = order[start:start + 150]
        assert mine.erase_ids(batch) == ref.erase_ids(batch) == len(batch)
        assert _same_flat(mine.flatten(), ref.flatten())
    # values go back into the thinned tree (ids beyond the first batch's)
    more = make_keys(kind, 200, 300 + kind)
    mine.insert(more, first_id=n)
    ref.insert(more, first_id=n)
    assert _same_flat(mine.flatten(), ref.flatten())
    mine.rebalance()
    ref.rebalance()
    assert _same_flat(mine.flatten(), ref.flatten())
    # an id that is not there is not an error, and the rest goes down to the empty tree
    assert mine.erase_ids(order[:5]) == ref.erase_ids(order[:5]) == 0
    rest = np.concatenate([order[600:], np.arange(n, n + 200, dtype=np.uint32)])
    assert mine.erase_ids(rest) == ref.erase_ids(rest) == len(rest)
    assert mine.size() == ref.size() == 0
    assert _same_flat(mine.flatten(), ref.flatten())


@needs_ref
@pytest.mark.parametrize("kind", [2, 3, 4, 5, 6, 8])
def test_degenerate_key_sets_give_the_reference_tree(rtb, kind):
    """ties everywhere -- all keys equal, a 3-valued lattice, collinear keys: zero-area bounds, equal
    enlargements, equal sort keys in the split -- and the tree is still the reference's bit for bit, before and
    after erasing a third of it"""
    scalar, dim, is_box, _ = O.KINDS[kind]
    dt = O.SCALAR_DTYPES[scalar]
    rng = np.random.default_rng(900 + kind)
    n = 1200
    line = np.concatenate([np.linspace(0, 1, n)[:, None], np.zeros((n, dim - 1))], axis=1)
    for name, pts, extent in (("all equal", np.repeat(rng.random((1, dim)), n, axis=0), 0.0),
                              ("lattice", rng.integers(0, 3, (n, dim)) / 2.0, 0.25),
                              ("collinear", line, 0.0)):
        pts = pts.astype(dt)
        keys = np.stack([pts, pts + dt(extent)], axis=1) if is_box else pts
        mine, ref = rtb.RTree(kind).insert(keys), O.RefTree(kind).insert(keys)
        assert mine.check_invariants() == 0, name
        assert _same_flat(mine.flatten(), ref.flatten()), name
        gone = rng.permutation(n).astype(np.uint32)[:400]
        assert mine.erase_ids(gone) == ref.erase_ids(gone) == 400
        assert _same_flat(mine.flatten(), ref.flatten()), name


@pytest.mark.parametrize("kind", [0, 2, 4, 5, 6])
def test_flatten_device_layout(rtb, kind):
    """level-ordered SoA blocks: every node's entries are those of flatten(), in slot order,
    children renumbered breadth-first, empty slots inverted + RT_INVALID_ID"""
    capi = rtb.capi
    keys = make_keys(kind, 2500, 90 + kind)
    tree = rtb.RTree(kind).insert(keys)
    flat = tree.flatten()
    image = tree.flatten_device(ids_from_mapped=True)
    d = image.desc
    scalar, dim, is_box, max_entries = O.KINDS[kind]
    assert (d.dim, d.scalar, d.max_entries) == (dim, scalar, max_entries)
    assert d.width == (8 if max_entries <= 8 else 16)
    assert d.key_kind == (capi.RT_KEY_BOX if is_box else capi.RT_KEY_POINT)
    assert d.leaf_level == flat["leaf_level"] and d.num_levels == d.leaf_level + 1
    assert d.num_nodes == len(flat["nodes"]) and d.num_values == 2500
    s = np.dtype(capi.SCALAR_DTYPES[scalar]).itemsize
    assert d.internal_stride == block_bytes(d.width, record_words(dim, s, True))
    assert d.leaf_stride == block_bytes(d.width, record_words(dim, s, is_box))
    if (kind, d.width) == (4, 8):
        assert (d.internal_stride, d.leaf_stride) == (256, 128)   # 3-D f32: 2 rows / 1 row of 128 B
    assert d.internal_stride % 16 == 0 and d.leaf_stride % 16 == 0 and d.leaf_offset % 128 == 0
    assert d.blocks % 256 == 0
    levels = image.level_node_begin
    assert levels[0] == 0 and levels[1] == 1 and levels[-1] == d.num_nodes
    to_dfs = image.bfs_to_dfs
    assert sorted(to_dfs.tolist()) == list(range(d.num_nodes)) and to_dfs[0] == 0
    nodes = decode_image(image, capi)
    next_child = 1
    for b, (is_leaf, lo, hi, link, valid) in enumerate(nodes):
        off, size, _ = flat["nodes"][to_dfs[b]]
        assert valid[:size].all() and not valid[size:].any()
        bound = flat["children_bound"][off:off + size]
        assert np.array_equal(lo, bound[:, 0, :]) and np.array_equal(hi, bound[:, 1, :])
        if is_leaf:
            assert np.array_equal(link, flat["data"][flat["children"][off:off + size]])
        else:
            kids = np.arange(next_child, next_child + size)
            assert np.array_equal(link, [node_link(image, k) for k in kids])
            assert np.array_equal(to_dfs[kids], flat["children"][off:off + size])
            next_child += size
    # data-index ids instead of mapped values
    image2 = tree.flatten_device(ids_from_mapped=False)
    leaves = [n for n in decode_image(image2, capi) if n[0]]
    assert np.array_equal(np.concatenate([n[3] for n in leaves]), np.arange(2500))


@needs_ref
def test_flatten_device_accepts_the_reference_flatten(rtb):
    """flatten_device() is duck-typed on the reference's flatten_result_t field names"""
    keys = make_keys(4, 3000, 3)
    ref = O.RefTree(4).insert(keys).flatten()
    a = rtb.DeviceImage.from_flat(4, ref.leaf_level, ref.nodes, ref.bounds, ref.children, ref.data)
    b = rtb.RTree(4).insert(keys).flatten_device()
    assert np.array_equal(a.blocks(), b.blocks())
    assert np.array_equal(a.level_node_begin, b.level_node_begin)


def test_empty_tree_image(rtb):
    image = rtb.RTree(4).flatten_device()
    d = image.desc
    assert d.num_nodes == 1 and d.leaf_level == 0 and d.num_values == 0
    (is_leaf, lo, hi, link, valid), = decode_image(image, rtb.capi)
    assert is_leaf and link.size == 0 and not valid.any()


REF_INCLUDE = "/root/reference/include"


@pytest.mark.skipif(not __import__("os").path.isdir(REF_INCLUDE), reason="reference sources not present")
def test_device_layout_compiles_against_reference_headers(rtb, tmp_path):
    """the reference's RTree + our device_layout.hpp, nothing else of
    ours, produce the same image as our host library."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "with_reference.cpp"
    src.write_text(r'''
#include <RTree.hpp>                  // the REFERENCE's umbrella header
#include <RTree/device_layout.hpp>    // ours
#include <cstdio>
#include <random>
int main()
{
  using point = eh::rtree::point_t<float, 3>;
  using box = eh::rtree::aabb_t<point>;
  using tree_t = eh::rtree::RTree<box, point, unsigned>;
  tree_t tree;
  std::mt19937 rng(5);
  std::uniform_real_distribution<float> u(0.f, 1.f);
  for (unsigned i = 0; i < 5000; ++i)
  {
    float x = u(rng), y = u(rng), z = u(rng);
    tree.insert({ point(x, y, z), i });
    std::printf("%a %a %a\n", x, y, z);
  }
  auto image = eh::rtree::flatten_device<tree_t::MAX_ENTRIES, false>(tree.flatten(), eh::rtree::id_source::mapped_value);
  rt_tree_desc d = image.desc();
  unsigned long long h = 1469598103934665603ull;
  for (unsigned long long i = 0; i < d.blocks_bytes; ++i) h = (h ^ image.blocks.data()[i]) * 1099511628211ull;
  std::fprintf(stderr, "%u %u %llu %llu\n", d.num_nodes, d.leaf_level, (unsigned long long)d.blocks_bytes, h);
  return 0;
}
''')
    exe = tmp_path / "with_reference"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", REF_INCLUDE, "-I",
                    os.path.join(root, "r-star-tree_b200", "include"), "-I", os.path.join(root, "include"),
                    "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True)
    pts = np.array([[float.fromhex(v) for v in line.split()] for line in out.stdout.split("\n") if line], np.float32)
    num_nodes, leaf_level, nbytes, fnv = [int(v) for v in out.stderr.split()]
    image = rtb.RTree(4).insert(pts).flatten_device(ids_from_mapped=True)
    assert (image.desc.num_nodes, image.desc.leaf_level, image.desc.blocks_bytes) == (num_nodes, leaf_level, nbytes)
    h = 1469598103934665603
    for byte in image.blocks().tobytes():
        h = ((h ^ byte) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert h == fnv


def test_device_image_file_round_trip_and_corruption(rtb, tmp_path):
    """save_device_tree / load_device_tree (RTree/device_layout.hpp): the reloaded image is the saved
    one byte for byte; truncated, corrupted or foreign files are refused"""
    rng = np.random.default_rng(5)
    for kind, keys in ((4, rng.random((5000, 3)).astype(np.float32)),
                       (2, rng.random((700, 2))),
                       (5, np.sort(rng.random((900, 2, 3)).astype(np.float32), axis=1)),
                       (4, rng.random((3, 3)).astype(np.float32))):
        image = rtb.RTree(kind).insert(keys).flatten_device()
        path = str(tmp_path / ("tree%d_%d.rtb" % (kind, len(keys))))
        image.save(path)
        back = rtb.DeviceImage.load(path)
        for name, _ in rtb.capi.RtTreeDesc._fields_:
            if name not in ("blocks", "level_node_begin", "reserved"):
                assert getattr(back.desc, name) == getattr(image.desc, name), name
        assert np.array_equal(back.level_node_begin, image.level_node_begin)
        assert np.array_equal(back.bfs_to_dfs, image.bfs_to_dfs)
        assert np.array_equal(back.blocks(), image.blocks())
    raw = bytearray(open(path, "rb").read())
    cases = {"truncated": raw[:len(raw) - 7], "trailing": raw + b"x", "magic": b"NOTANIMG" + raw[8:],
             "header_only": raw[:128], "empty": b""}
    flipped = bytearray(raw)
    flipped[-1] ^= 0x40                      # one bit in the blocks
    cases["bitflip"] = flipped
    version = bytearray(raw)
    version[8] = 99
    cases["version"] = version
    for name, data in cases.items():
        bad = str(tmp_path / (name + ".rtb"))
        with open(bad, "wb") as f:
            f.write(data)
        with pytest.raises(RuntimeError):
            rtb.DeviceImage.load(bad)
    with pytest.raises(RuntimeError):
        rtb.DeviceImage.load(str(tmp_path / "does_not_exist.rtb"))


@pytest.mark.parametrize("kind", [4, 5, 2])
def test_set_keys_rebound_matches_the_reference(rtb, kind):
    """keys changed in place through the iterators + rebound() for every value: the host library's
    tree flattens to the same arrays as the reference's after the same mutation"""
    scalar, dim, is_box, _ = O.KINDS[kind]
    dt = O.SCALAR_DTYPES[scalar]
    rng = np.random.default_rng(40 + kind)
    n = 6000
    c = rng.random((n, dim))
    keys = np.stack([c, c + rng.uniform(0, 0.02, (n, dim))], axis=1).astype(dt) if is_box else c.astype(dt)
    mine, ref = rtb.RTree(kind).insert(keys), O.RefTree(kind).insert(keys)
    order = mine.flatten()["data"]
    moved = (keys + rng.normal(0, 0.05, keys.shape)).astype(dt)
    if is_box:
        moved = np.sort(moved, axis=1)
    mine.set_keys(moved[order])
    ref.set_keys(moved[order])
    a, b = mine.flatten(), ref.flatten()
    assert np.array_equal(a["nodes"], b.nodes) and np.array_equal(a["children"], b.children)
    assert np.array_equal(a["data"], b.data)
    assert np.array_equal(np.asarray(a["children_bound"]).view(np.uint8), b.bounds.view(np.uint8))
