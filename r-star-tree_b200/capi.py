"""ctypes bindings of the C ABI declared in include/rtree_b200.h (librtree_b200.so).

Nothing here computes anything: every query goes to the CUDA library.  If the library
is missing the import of this module still works (so CPU-only tooling can inspect the
ABI), but the first call raises ``RtError`` -- there is no fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RTB_LIB names another build of the same library (the experiment variants of csrc/Makefile)
LIB_PATH = os.environ.get("RTB_LIB") or os.path.join(_HERE, "csrc", "librtree_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "rtree_b200.h")

RT_ABI_VERSION = 2
RT_OK, RT_E_INVALID, RT_E_UNSUPPORTED, RT_E_CUDA, RT_E_NOMEM, RT_E_INTERNAL = 0, -1, -2, -3, -4, -5
RT_F32, RT_F64, RT_I32 = 0, 1, 2
RT_KEY_POINT, RT_KEY_BOX = 0, 1
RT_MEM_HOST, RT_MEM_DEVICE = 0, 1
RT_POINT_IN_BOX, RT_INTERSECTS, RT_CONTAINS = 0, 1, 2
RT_RAY_ALL_HITS, RT_RAY_CLOSEST, RT_RAY_ANY = 0, 1, 2
RT_QUERIES_ON_DEVICE, RT_RESULT_ON_DEVICE, RT_UNSORTED = 0x1, 0x2, 0x4
RT_COUNT_ONLY, RT_COLLECT_VISITS, RT_NO_REORDER, RT_ANY_HIT = 0x8, 0x10, 0x20, 0x40
RT_RESULT_COUNTS32 = 0x80
RT_STRICT_OVERLAP = 0x100
RT_INVALID_ID = 0xFFFFFFFF
RT_OPT_PIPELINE_BATCH, RT_OPT_QUERY_CHUNK, RT_OPT_WORK_RANGES, RT_OPT_TRAVERSE_GRID = 1, 2, 3, 4
RT_OPT_QUANTISED = 5

SCALAR_DTYPES = {RT_F32: np.float32, RT_F64: np.float64, RT_I32: np.int32}

# every symbol include/rtree_b200.h declares
EXPORTS = ["rt_upload", "rt_query_boxes", "rt_query_rays", "rt_query_knn", "rt_refit", "rt_free", "rt_set_stream",
           "rt_set_option", "rt_plan_batches", "rt_get_stats", "rt_tree_device_blocks", "rt_tree_get_desc", "rt_last_error",
           "rt_abi_version", "rt_device_count"]


class RtError(RuntimeError):
    def __init__(self, code, what):
        RuntimeError.__init__(self, "rtree_b200 error %d: %s" % (code, what))
        self.code = code


class RtTreeDesc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("dim", C.c_uint32), ("scalar", C.c_uint32),
                ("width", C.c_uint32), ("max_entries", C.c_uint32), ("key_kind", C.c_uint32),
                ("leaf_level", C.c_uint32), ("num_levels", C.c_uint32), ("num_nodes", C.c_uint32),
                ("num_values", C.c_uint32), ("internal_stride", C.c_uint32),
                ("leaf_stride", C.c_uint32), ("leaf_offset", C.c_uint64),
                ("blocks_bytes", C.c_uint64), ("blocks", C.c_void_p),
                ("level_node_begin", C.POINTER(C.c_uint32)), ("blocks_memory", C.c_uint32),
                ("reserved", C.c_uint32)]


class RtHits(C.Structure):
    _fields_ = [("n_queries", C.c_uint64), ("total", C.c_uint64), ("offsets", C.c_void_p),
                ("ids", C.c_void_p), ("memory", C.c_uint32), ("reserved", C.c_uint32),
                ("counts", C.c_void_p)]


class RtRayResult(C.Structure):
    _fields_ = [("n_rays", C.c_uint64), ("n_hit", C.c_uint64), ("hits", RtHits),
                ("t", C.c_void_p), ("id", C.c_void_p), ("any", C.c_void_p),
                ("memory", C.c_uint32), ("reserved", C.c_uint32)]


class RtKnnResult(C.Structure):
    _fields_ = [("n_queries", C.c_uint64), ("k", C.c_uint32), ("memory", C.c_uint32),
                ("ids", C.c_void_p), ("dist2", C.c_void_p)]


class RtStats(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("h2d_ms", C.c_float), ("reorder_ms", C.c_float),
                ("traverse_ms", C.c_float), ("finish_ms", C.c_float), ("d2h_ms", C.c_float),
                ("kernel_launches", C.c_uint32), ("deferred_queries", C.c_uint32),
                ("visits_internal", C.c_uint64), ("visits_leaf", C.c_uint64),
                ("algorithmic_bytes", C.c_uint64), ("traverse_steps", C.c_uint64),
                ("batches", C.c_uint32), ("reserved", C.c_uint32), ("leaf_lookups", C.c_uint64)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


_lib = None


def have_library():
    return os.path.exists(LIB_PATH)


def lib():
    """Load librtree_b200.so; raises loudly when it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RtError(RT_E_CUDA, "CUDA library %s is missing: run __graft_entry__.build() "
                                     "(make -C r-star-tree_b200/csrc); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.rt_upload.restype = C.c_int
        L.rt_upload.argtypes = [C.POINTER(RtTreeDesc), C.c_int, C.POINTER(C.c_void_p)]
        L.rt_query_boxes.restype = C.c_int
        L.rt_query_boxes.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                     C.POINTER(RtHits)]
        L.rt_query_rays.restype = C.c_int
        L.rt_query_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                    C.POINTER(RtRayResult)]
        L.rt_free.restype = None
        L.rt_free.argtypes = [C.c_void_p]
        L.rt_query_knn.restype = C.c_int
        L.rt_query_knn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32,
                                   C.POINTER(RtKnnResult)]
        L.rt_refit.restype = C.c_int
        L.rt_refit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
        L.rt_set_stream.restype = C.c_int
        L.rt_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.rt_set_option.restype = C.c_int
        L.rt_set_option.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64]
        L.rt_plan_batches.restype = C.c_int64
        L.rt_plan_batches.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint64]
        L.rt_get_stats.restype = C.c_int
        L.rt_get_stats.argtypes = [C.c_void_p, C.POINTER(RtStats)]
        L.rt_tree_device_blocks.restype = C.c_int
        L.rt_tree_device_blocks.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.rt_tree_get_desc.restype = C.c_int
        L.rt_tree_get_desc.argtypes = [C.c_void_p, C.POINTER(RtTreeDesc)]
        L.rt_last_error.restype = C.c_char_p
        L.rt_abi_version.restype = C.c_uint32
        L.rt_device_count.restype = C.c_int
        _lib = L
    return _lib


def _check(code):
    if code != RT_OK:
        raise RtError(code, lib().rt_last_error().decode("utf-8", "replace"))


def _host_array(ptr, count, dtype):
    """Copy `count` elements of engine-owned host memory into a fresh numpy array."""
    if count == 0 or not ptr:
        return np.zeros(0, dtype)
    buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


def plan_batches(n, pipeline_batch):
    """the batch sizes rt_query_boxes cuts a host-result call of n queries into (host arithmetic only)"""
    count = lib().rt_plan_batches(n, pipeline_batch, None, 0)
    if count < 0:
        _check(int(count))
    sizes = (C.c_uint64 * max(1, count))()
    lib().rt_plan_batches(n, pipeline_batch, sizes, count)
    return [int(x) for x in sizes[:count]]


class DeviceBuffer:
    """A raw device pointer exposed through __cuda_array_interface__ (so torch.as_tensor can
    wrap it without a copy).  Does not own the memory."""

    def __init__(self, ptr, nbytes, owner=None):
        self.ptr, self.nbytes, self._owner = ptr, nbytes, owner
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 2, "strides": None}


class DeviceTree:
    """An uploaded tree on one GPU: rt_upload() ... rt_free()."""

    def __init__(self, desc, device=0, keepalive=None):
        self._h = C.c_void_p()
        self.device = device
        _check(lib().rt_upload(C.byref(desc), device, C.byref(self._h)))
        d = RtTreeDesc()
        _check(lib().rt_tree_get_desc(self._h, C.byref(d)))
        self.dim, self.scalar, self.width, self.key_kind = d.dim, d.scalar, d.width, d.key_kind
        self.leaf_level, self.num_nodes, self.num_values = d.leaf_level, d.num_nodes, d.num_values
        self.internal_stride, self.leaf_stride = d.internal_stride, d.leaf_stride
        self.blocks_bytes = d.blocks_bytes
        self.dtype = SCALAR_DTYPES[d.scalar]
        del keepalive  # the upload copied everything it needs

    def close(self):
        if getattr(self, "_h", None):
            lib().rt_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream):
        _check(lib().rt_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def set_option(self, option, value):
        _check(lib().rt_set_option(self._h, C.c_uint32(option), C.c_uint64(int(value))))

    def stats(self):
        s = RtStats()
        _check(lib().rt_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def device_blocks(self):
        p, n = C.c_void_p(), C.c_uint64()
        _check(lib().rt_tree_device_blocks(self._h, C.byref(p), C.byref(n)))
        return DeviceBuffer(p.value, n.value, self)

    # ---- queries ----------------------------------------------------------------
    def query_boxes_raw(self, boxes_ptr, n, mode=RT_POINT_IN_BOX, flags=0):
        """Thin call: returns the RtHits struct (engine-owned buffers)."""
        out = RtHits()
        _check(lib().rt_query_boxes(self._h, C.c_void_p(boxes_ptr), n, mode, flags, C.byref(out)))
        return out

    def query_boxes(self, boxes, mode=RT_POINT_IN_BOX, flags=0):
        """boxes: host array [n][2][dim] -> (offsets[n+1] uint64, ids uint32) numpy copies
        (with RT_RESULT_COUNTS32: (counts[n] uint32, ids))."""
        q = np.ascontiguousarray(boxes, self.dtype).reshape(-1, 2, self.dim)
        flags &= ~(RT_QUERIES_ON_DEVICE | RT_RESULT_ON_DEVICE)
        out = self.query_boxes_raw(q.ctypes.data, len(q), mode, flags)
        ids = _host_array(out.ids, 0 if flags & (RT_COUNT_ONLY | RT_ANY_HIT) else out.total, np.uint32)
        if flags & RT_RESULT_COUNTS32:  # 32-bit hits per query instead of the 64-bit offsets
            return _host_array(out.counts, out.n_queries, np.uint32), ids
        offsets = _host_array(out.offsets, out.n_queries + 1, np.uint64)
        return offsets, ids

    def refit(self, keys):
        """rt_refit(): new keys (flatten data order) + all bounds recomputed on the device."""
        per = self.dim * (2 if self.key_kind == RT_KEY_BOX else 1)
        k = np.ascontiguousarray(keys, self.dtype).reshape(-1, per)
        _check(lib().rt_refit(self._h, C.c_void_p(k.ctypes.data), len(k), 0))

    def query_knn_raw(self, points_ptr, n, k, flags=0):
        out = RtKnnResult()
        _check(lib().rt_query_knn(self._h, C.c_void_p(points_ptr), n, k, flags, C.byref(out)))
        return out

    def query_knn(self, points, k, flags=0):
        """points: host array [n][dim] float32 -> (ids[n][k] uint32, dist2[n][k] float32)."""
        p = np.ascontiguousarray(points, np.float32).reshape(-1, self.dim)
        flags &= ~(RT_QUERIES_ON_DEVICE | RT_RESULT_ON_DEVICE)
        out = self.query_knn_raw(p.ctypes.data, len(p), k, flags)
        return (_host_array(out.ids, len(p) * k, np.uint32).reshape(len(p), k),
                _host_array(out.dist2, len(p) * k, np.float32).reshape(len(p), k))

    def query_rays_raw(self, rays_ptr, n, mode, flags=0):
        out = RtRayResult()
        _check(lib().rt_query_rays(self._h, C.c_void_p(rays_ptr), n, mode, flags, C.byref(out)))
        return out

    def query_rays(self, rays, mode, flags=0):
        """rays: host array [n][7] float32.
        ALL_HITS -> (offsets, ids); CLOSEST -> (t, id); ANY -> uint8[n]."""
        r = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
        flags &= ~(RT_QUERIES_ON_DEVICE | RT_RESULT_ON_DEVICE)
        out = self.query_rays_raw(r.ctypes.data, len(r), mode, flags)
        if mode == RT_RAY_ALL_HITS:
            return (_host_array(out.hits.offsets, out.n_rays + 1, np.uint64),
                    _host_array(out.hits.ids, 0 if flags & RT_COUNT_ONLY else out.hits.total, np.uint32))
        if mode == RT_RAY_CLOSEST:
            return _host_array(out.t, out.n_rays, np.float32), _host_array(out.id, out.n_rays, np.uint32)
        return _host_array(out.any, out.n_rays, np.uint8)
