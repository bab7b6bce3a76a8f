"""GPU parity tests for rt_refit: new keys into the leaf slots and every bound recomputed bottom-up on
the device.  The checker is the reference itself: the same keys changed in place through its iterators,
RTree::rebound(iterator) called for every value (reference README.md:342-347, rtree.hpp:479-500), then
flatten().  min / max only, so the bar is byte equality of the whole image, and bit-exact queries."""
import numpy as np
import pytest

import oracles as O
from helpers import csr_equal

pytestmark = pytest.mark.gpu


def device_blocks(tree, nbytes):
    import torch
    return torch.as_tensor(tree.device_blocks(), device="cuda").cpu().numpy()[:nbytes]


@pytest.mark.parametrize("kind", [4, 5, 2, 8, 6, 12])
def test_refit_equals_reference_rebound(rtb, kind):
    scalar, dim, is_box, _ = O.KINDS[kind]
    dt = O.SCALAR_DTYPES[scalar]
    rng = np.random.default_rng(90 + kind)
    n = 20000
    if scalar == O.I32:
        keys = rng.integers(0, 500, (n, dim)).astype(np.int32)
        moved = (keys + rng.integers(-20, 21, keys.shape)).astype(np.int32)
    else:
        c = rng.random((n, dim))
        keys = np.stack([c, c + rng.uniform(0, 0.02, (n, dim))], axis=1).astype(dt) if is_box else c.astype(dt)
        moved = (keys + rng.normal(0, 0.03, keys.shape)).astype(dt)
        if is_box:
            moved = np.sort(moved, axis=1)
    ref = O.RefTree(kind).insert(keys)
    flat = ref.flatten()
    image = rtb.DeviceImage.from_flat(kind, flat.leaf_level, flat.nodes, flat.bounds, flat.children, flat.data)
    in_data_order = moved[flat.data]
    ref.set_keys(in_data_order)
    flat2 = ref.flatten()
    assert np.array_equal(flat2.nodes, flat.nodes) and np.array_equal(flat2.data, flat.data)   # topology kept
    want = rtb.DeviceImage.from_flat(kind, flat2.leaf_level, flat2.nodes, flat2.bounds, flat2.children, flat2.data)
    nq = 3000
    if scalar == O.I32:
        lo = rng.integers(-30, 520, (nq, dim))
        q = np.stack([lo, lo + rng.integers(0, 12, (nq, dim))], axis=1).astype(np.int32)
    else:
        qc = rng.uniform(-0.1, 1.1, (nq, dim))
        qh = 0.5 * (10.0 / n) ** (1.0 / dim) * rng.uniform(0.5, 3.0, (nq, 1))
        q = np.stack([qc - qh, qc + qh], axis=1).astype(dt)
    mode = rtb.RT_INTERSECTS if is_box else rtb.RT_POINT_IN_BOX
    with image.upload(0) as tree:
        before = tree.query_boxes(q, mode)
        tree.refit(in_data_order)
        assert tree.stats()["kernel_launches"] >= flat.leaf_level + 2
        got = device_blocks(tree, want.desc.blocks_bytes)
        assert np.array_equal(got, want.blocks())
        after = tree.query_boxes(q, mode)
        assert csr_equal(after, ref.search_boxes(q, mode))
        assert not csr_equal(after, before)
        tree.refit(keys[flat.data])                      # and back: the original image and answers
        assert np.array_equal(device_blocks(tree, image.desc.blocks_bytes), image.blocks())
        assert csr_equal(tree.query_boxes(q, mode), before)


def test_refit_argument_checks_and_root_leaf(rtb):
    pts = rtb.workloads.uniform_points(6, seed=3)
    with rtb.RTree(4).insert(pts).flatten_device().upload(0) as tree:      # the root is the only leaf
        moved = (pts + np.float32(2.0)).astype(np.float32)
        tree.refit(moved)
        q = np.array([[[1.9, 1.9, 1.9], [3.1, 3.1, 3.1]], [[0, 0, 0], [1, 1, 1]]], np.float32)
        off, ids = tree.query_boxes(q)
        assert list(off) == [0, 6, 6] and sorted(ids) == list(range(6))
        with pytest.raises(rtb.RtError) as err:
            tree.refit(moved[:5])
        assert err.value.code == rtb.capi.RT_E_INVALID
