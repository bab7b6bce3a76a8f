"""Shared helpers for the tests: decoding a device image back into per-node entries."""
import numpy as np


TAIL_BYTES = (0, 4, 8, 16)


def record_words(dim, scalar_bytes, two_corners):
    sw = scalar_bytes // 4
    return 2 * dim * sw + sw if two_corners else dim * sw + 1


def block_bytes(width, words):
    return width * (16 * (words // 4) + TAIL_BYTES[words % 4])


def block_records(raw, start, width, words):
    """uint32 [width][words]: the slot records of the block at byte offset `start`
    (rows of 16-byte vectors + tail row, include/rtree_b200.h)."""
    full = words // 4
    rec = np.zeros((width, words), np.uint32)
    for k in range(full):
        row = np.frombuffer(raw[start + k * width * 16:start + (k + 1) * width * 16].tobytes(), np.uint32)
        rec[:, 4 * k:4 * k + 4] = row.reshape(width, 4)
    r = words % 4
    if r:
        eb = TAIL_BYTES[r]
        t0 = start + full * width * 16
        tail = np.frombuffer(raw[t0:t0 + width * eb].tobytes(), np.uint32).reshape(width, eb // 4)
        rec[:, 4 * full:] = tail[:, :r]
    return rec


def decode_image(image, capi):
    """Decode RTree::flatten_device() output (rt_tree_desc + blocks) into, per node in
    breadth-first order: (is_leaf, lo[n][dim], hi[n][dim], link[n], valid[W]) with empty slots
    removed from lo / hi / link."""
    d = image.desc
    dt = np.dtype(capi.SCALAR_DTYPES[d.scalar])
    W, D, sb = d.width, d.dim, dt.itemsize
    raw = image.blocks()
    levels = image.level_node_begin
    leaf_begin = int(levels[d.leaf_level])
    out = []
    for b in range(d.num_nodes):
        is_leaf = b >= leaf_begin
        start = d.leaf_offset + (b - leaf_begin) * d.leaf_stride if is_leaf else b * d.internal_stride
        one_corner = is_leaf and d.key_kind == capi.RT_KEY_POINT
        words = record_words(D, sb, not one_corner)
        rec = block_records(raw, start, W, words)
        sw = sb // 4
        link = rec[:, D * sw]                          # record: lo corner, link (+pad), hi corner
        valid = link != capi.RT_INVALID_ID
        lo = np.ascontiguousarray(rec[:, :D * sw]).view(dt)[valid]
        if one_corner:
            hi = lo
        else:
            hi = np.ascontiguousarray(rec[:, (D + 1) * sw:(2 * D + 1) * sw]).view(dt)[valid]
        out.append((is_leaf, lo, hi, link[valid], valid))
    return out


def node_link(image, b):
    """the link (block byte offset / 16) that names breadth-first node b"""
    d = image.desc
    leaf_begin = int(image.level_node_begin[d.leaf_level])
    off = d.leaf_offset + (b - leaf_begin) * d.leaf_stride if b >= leaf_begin else b * d.internal_stride
    return off // 16


def csr_equal(a, b):
    return (np.array_equal(np.asarray(a[0], np.uint64), np.asarray(b[0], np.uint64))
            and np.array_equal(np.asarray(a[1], np.uint32), np.asarray(b[1], np.uint32)))
