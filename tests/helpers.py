"""Shared helpers for the tests: decoding a device image back into per-node entries."""
import numpy as np


def decode_image(image, capi):
    """Decode RTree::flatten_device() output (rt_tree_desc + blocks) into, per node in
    breadth-first order: (is_leaf, lo[n][dim], hi[n][dim], link[n]) with empty slots removed."""
    d = image.desc
    dt = np.dtype(capi.SCALAR_DTYPES[d.scalar])
    W, D = d.width, d.dim
    raw = image.blocks()
    levels = image.level_node_begin
    leaf_begin = int(levels[d.leaf_level])
    out = []
    for b in range(d.num_nodes):
        is_leaf = b >= leaf_begin
        if is_leaf:
            start = d.leaf_offset + (b - leaf_begin) * d.leaf_stride
        else:
            start = b * d.internal_stride
        one_corner = is_leaf and d.key_kind == capi.RT_KEY_POINT
        rows = D if one_corner else 2 * D
        vals = np.frombuffer(raw[start:start + rows * W * dt.itemsize].tobytes(), dt).reshape(rows, W)
        link = np.frombuffer(raw[start + rows * W * dt.itemsize:start + rows * W * dt.itemsize + 4 * W].tobytes(),
                             np.uint32)
        valid = link != capi.RT_INVALID_ID
        lo = vals[:D, valid].T
        hi = lo if one_corner else vals[D:, valid].T
        out.append((is_leaf, lo, hi, link[valid], valid))
    return out


def csr_equal(a, b):
    return (np.array_equal(np.asarray(a[0], np.uint64), np.asarray(b[0], np.uint64))
            and np.array_equal(np.asarray(a[1], np.uint32), np.asarray(b[1], np.uint32)))
