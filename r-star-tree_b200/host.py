"""Handle-based Python mirror of the host library (include/RTree.hpp) through
csrc/librtree_b200_host.so (host_bridge.cpp).

``RTree(kind)`` is one instantiation of ``eh::rtree::RTree<Geometry, Key, uint32_t, Config>``;
it builds on the CPU with the library's own insert()/split/reinsert, and
``flatten_device()`` gives the image ``DeviceImage.upload()`` hands to ``rt_upload``.
"""
import ctypes as C
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "csrc", "librtree_b200_host.so")

F32, F64, I32 = capi.RT_F32, capi.RT_F64, capi.RT_I32

# kind -> (scalar, dim, key_is_box, max_entries); see host_bridge.cpp
KINDS = {
    0: (I32, 1, False, 8),
    1: (F64, 1, False, 8),
    2: (F64, 2, False, 8),
    3: (F32, 2, False, 8),
    4: (F32, 3, False, 8),
    5: (F32, 3, True, 8),
    6: (F32, 8, False, 16),
    7: (F32, 2, True, 8),
    8: (I32, 2, False, 8),
    9: (F32, 3, False, 8),
    10: (F32, 1, False, 8),
    11: (F32, 4, False, 8),
    12: (F64, 3, False, 8),
    13: (I32, 3, False, 8),
}

_lib = None


def have_library():
    return os.path.exists(HOST_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError("host bridge %s is missing: run __graft_entry__.build()" % HOST_LIB_PATH)
        L = C.CDLL(HOST_LIB_PATH)
        vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
        L.rtb_last_error.restype = C.c_char_p
        L.rtb_tree_create.restype = vp
        L.rtb_tree_create.argtypes = [C.c_int]
        L.rtb_tree_destroy.argtypes = [vp]
        L.rtb_tree_insert.argtypes = [vp, vp, u64, u32]
        L.rtb_tree_erase_ids.restype = C.c_int64
        L.rtb_tree_erase_ids.argtypes = [vp, vp, u64]
        L.rtb_tree_clear.argtypes = [vp]
        L.rtb_tree_set_keys.restype = C.c_int64
        L.rtb_tree_set_keys.argtypes = [vp, vp, u64]
        L.rtb_tree_rebalance.argtypes = [vp]
        L.rtb_tree_leaf_level.restype = u32
        L.rtb_tree_leaf_level.argtypes = [vp]
        L.rtb_tree_size.restype = u64
        L.rtb_tree_size.argtypes = [vp]
        L.rtb_tree_check_invariants.argtypes = [vp]
        L.rtb_tree_flatten.argtypes = [vp, C.POINTER(u64)]
        L.rtb_tree_flat_copy.argtypes = [vp, vp, vp, vp, vp]
        L.rtb_tree_search_boxes.restype = C.c_int64
        L.rtb_tree_search_boxes.argtypes = [vp, vp, u64, C.c_int, vp, C.POINTER(C.POINTER(u32))]
        L.rtb_free.argtypes = [vp]
        L.rtb_flatten_device.restype = vp
        L.rtb_flatten_device.argtypes = [vp, C.c_int]
        L.rtb_flatten_device_from_arrays.restype = vp
        L.rtb_flatten_device_from_arrays.argtypes = [C.c_int, u32, vp, u64, vp, vp, u64, vp, u64, C.c_int]
        L.rtb_device_tree_desc.argtypes = [vp, C.POINTER(capi.RtTreeDesc)]
        L.rtb_device_tree_bfs_to_dfs.restype = C.POINTER(u32)
        L.rtb_device_tree_bfs_to_dfs.argtypes = [vp]
        L.rtb_device_tree_free.argtypes = [vp]
        L.rtb_device_tree_save.restype = C.c_int
        L.rtb_device_tree_save.argtypes = [vp, C.c_char_p]
        L.rtb_device_tree_load.restype = vp
        L.rtb_device_tree_load.argtypes = [C.c_char_p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _err():
    return lib().rtb_last_error().decode("utf-8", "replace")


class DeviceImage:
    """What RTree::flatten_device() returns: rt_tree_desc + the host block buffer."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("flatten_device failed: " + _err())
        self._h = handle
        self.desc = capi.RtTreeDesc()
        lib().rtb_device_tree_desc(self._h, C.byref(self.desc))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rtb_device_tree_free(self._h)
            self._h = None

    @property
    def level_node_begin(self):
        d = self.desc
        return np.ctypeslib.as_array(d.level_node_begin, shape=(d.num_levels + 1,)).copy()

    @property
    def bfs_to_dfs(self):
        return np.ctypeslib.as_array(lib().rtb_device_tree_bfs_to_dfs(self._h),
                                     shape=(self.desc.num_nodes,)).copy()

    def blocks(self):
        """The block buffer as a uint8 numpy view (borrowed from this object)."""
        d = self.desc
        buf = (C.c_uint8 * d.blocks_bytes).from_address(d.blocks)
        return np.frombuffer(buf, dtype=np.uint8)

    def upload(self, device=0):
        """rt_upload(): copy the image to a GPU, returns a capi.DeviceTree."""
        return capi.DeviceTree(self.desc, device, keepalive=self)

    def save(self, path):
        """save_device_tree(): write the image to one file (header + tables + blocks, checksummed)."""
        if lib().rtb_device_tree_save(self._h, os.fsencode(path)) != 0:
            raise RuntimeError("save_device_tree failed: " + _err())

    @staticmethod
    def load(path):
        """load_device_tree(): read an image back; raises if the file is truncated or corrupted."""
        h = lib().rtb_device_tree_load(os.fsencode(path))
        if not h:
            raise RuntimeError("load_device_tree failed: " + _err())
        return DeviceImage(h)

    @staticmethod
    def from_flat(kind, leaf_level, nodes, bounds, children, data, ids_from_mapped=True):
        """flatten_device() applied to a flatten_result_t given as arrays (for instance the
        output of the reference's own flatten())."""
        scalar, dim, _, _ = KINDS[kind]
        nodes = np.ascontiguousarray(nodes, np.uint32).reshape(-1, 3)
        bounds = np.ascontiguousarray(bounds, capi.SCALAR_DTYPES[scalar]).reshape(-1, 2, dim)
        children = np.ascontiguousarray(children, np.uint32)
        data = np.ascontiguousarray(data, np.uint32)
        h = lib().rtb_flatten_device_from_arrays(kind, int(leaf_level), _ptr(nodes), len(nodes),
                                                 _ptr(bounds), _ptr(children), len(children),
                                                 _ptr(data), len(data), int(ids_from_mapped))
        return DeviceImage(h)


class RTree:
    """eh::rtree::RTree<...> (one of KINDS) built on the CPU by the host library."""

    def __init__(self, kind):
        self.kind = kind
        self.scalar, self.dim, self.key_is_box, self.max_entries = KINDS[kind]
        self.dtype = capi.SCALAR_DTYPES[self.scalar]
        self._h = lib().rtb_tree_create(kind)
        if not self._h:
            raise ValueError("unknown tree kind %r" % (kind,))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rtb_tree_destroy(self._h)
            self._h = None

    def insert(self, keys, first_id=0):
        """insert keys[i] with mapped value first_id + i (points [n][dim], boxes [n][2][dim])."""
        keys = np.ascontiguousarray(keys, self.dtype)
        per = self.dim * (2 if self.key_is_box else 1)
        assert keys.size % per == 0
        if lib().rtb_tree_insert(self._h, _ptr(keys), keys.size // per, first_id) != 0:
            raise RuntimeError(_err())
        return self

    def set_keys(self, keys):
        """change every key in place (iteration order = flatten data order) and rebound() every
        value: bounds refitted, topology unchanged -- the host-side form of rt_refit()."""
        keys = np.ascontiguousarray(keys, self.dtype)
        per = self.dim * (2 if self.key_is_box else 1)
        if lib().rtb_tree_set_keys(self._h, _ptr(keys), keys.size // per) != keys.size // per:
            raise RuntimeError("set_keys: key count differs from the tree's size " + _err())
        return self

    def erase_ids(self, ids):
        ids = np.ascontiguousarray(ids, np.uint32)
        return lib().rtb_tree_erase_ids(self._h, _ptr(ids), len(ids))

    def clear(self):
        lib().rtb_tree_clear(self._h)

    def rebalance(self):
        lib().rtb_tree_rebalance(self._h)

    def leaf_level(self):
        return lib().rtb_tree_leaf_level(self._h)

    def size(self):
        return lib().rtb_tree_size(self._h)

    def check_invariants(self):
        return lib().rtb_tree_check_invariants(self._h)

    def flatten(self):
        """RTree::flatten() -> dict of numpy arrays named like flatten_result_t."""
        sizes = (C.c_uint64 * 5)()
        lib().rtb_tree_flatten(self._h, sizes)
        n_nodes, n_children, n_data, leaf_level, root = [int(x) for x in sizes]
        nodes = np.zeros((n_nodes, 3), np.uint32)
        bounds = np.zeros((n_children, 2, self.dim), self.dtype)
        children = np.zeros(n_children, np.uint32)
        data = np.zeros(n_data, np.uint32)
        lib().rtb_tree_flat_copy(self._h, _ptr(nodes), _ptr(bounds), _ptr(children), _ptr(data))
        return {"leaf_level": leaf_level, "root": root, "nodes": nodes, "children_bound": bounds,
                "children": children, "data": data}

    def search_boxes(self, queries, mode=capi.RT_POINT_IN_BOX):
        """The host library's CPU RTree::search() per query (host-API parity, not the GPU path)."""
        q = np.ascontiguousarray(queries, self.dtype).reshape(-1, 2, self.dim)
        offsets = np.zeros(len(q) + 1, np.uint64)
        pp = C.POINTER(C.c_uint32)()
        total = lib().rtb_tree_search_boxes(self._h, _ptr(q), len(q), mode, _ptr(offsets), C.byref(pp))
        if total < 0:
            raise RuntimeError(_err())
        ids = np.ctypeslib.as_array(pp, shape=(total,)).copy() if total else np.zeros(0, np.uint32)
        lib().rtb_free(C.cast(pp, C.c_void_p))
        return offsets, ids

    def flatten_device(self, ids_from_mapped=True):
        """RTree::flatten_device(): hits report the mapped value (default) or the data index."""
        return DeviceImage(lib().rtb_flatten_device(self._h, int(ids_from_mapped)))
