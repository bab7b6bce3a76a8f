"""ctypes bindings for the two CPU checkers under oracle/ (TEST INFRASTRUCTURE).

* ``RefTree``   -- oracle/_ref/libref_oracle.so : the unmodified reference library
  (its own insert / search / flatten) behind oracle/ref_harness.cpp.
* ``oracle_*``  -- oracle/liboracle.so : the plain-C restatement (oracle/oracle.c).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl
reference`` leg import this module.  The product never does.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so")

F32, F64, I32 = 0, 1, 2
SCALAR_DTYPES = {F32: np.float32, F64: np.float64, I32: np.int32}
MODE_POINT_IN_BOX, MODE_INTERSECTS, MODE_CONTAINS = 0, 1, 2
MODE_STRICT = 0x100  # OR-ed in: the reference's own (strict) is_overlap as the filter, see RT_STRICT_OVERLAP
RAY_ALL, RAY_CLOSEST, RAY_ANY = 0, 1, 2

# kind -> (scalar, dim, key_is_box, max_entries); must match make_tree() in
# oracle/ref_harness.cpp and the product's host bridge.
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

_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_u32pp = C.POINTER(_u32p)


def have_ref():
    return os.path.exists(REF_SO)


def have_oracle():
    return os.path.exists(ORACLE_SO)


_ref_lib = None
_oracle_lib = None


def ref_lib():
    global _ref_lib
    if _ref_lib is None:
        lib = C.CDLL(REF_SO)
        lib.ref_tree_create.restype = C.c_void_p
        lib.ref_tree_create.argtypes = [C.c_int]
        lib.ref_tree_destroy.argtypes = [C.c_void_p]
        lib.ref_tree_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
        lib.ref_tree_leaf_level.restype = C.c_uint32
        lib.ref_tree_leaf_level.argtypes = [C.c_void_p]
        lib.ref_tree_size.restype = C.c_uint64
        lib.ref_tree_size.argtypes = [C.c_void_p]
        lib.ref_tree_rebalance.argtypes = [C.c_void_p]
        lib.ref_tree_flatten.argtypes = [C.c_void_p, _u64p]
        lib.ref_tree_flat_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_search_boxes.restype = C.c_int64
        lib.ref_search_boxes.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int,
                                         C.c_void_p, _u32pp, C.POINTER(C.c_double)]
        lib.ref_search_rays.restype = C.c_int64
        lib.ref_search_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int,
                                        C.c_void_p, _u32pp, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_double)]
        lib.ref_tree_set_keys.restype = C.c_int64
        lib.ref_tree_set_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        if hasattr(lib, "ref_tree_erase_ids"):  # (a prebuilt _ref from before the symbol existed still serves the rest)
            lib.ref_tree_erase_ids.restype = C.c_uint64
            lib.ref_tree_erase_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        lib.ref_search_knn.restype = C.c_int64
        lib.ref_search_knn.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int,
                                       C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        lib.ref_free.argtypes = [C.c_void_p]
        lib.ref_max_threads.restype = C.c_int
        _ref_lib = lib
    return _ref_lib


class OracleFlat(C.Structure):
    _fields_ = [("leaf_level", C.c_uint32), ("root", C.c_uint32), ("dim", C.c_uint32),
                ("scalar", C.c_uint32), ("n_nodes", C.c_uint64), ("nodes", C.c_void_p),
                ("bounds", C.c_void_p), ("children", C.c_void_p), ("data", C.c_void_p)]


def oracle_lib():
    global _oracle_lib
    if _oracle_lib is None:
        lib = C.CDLL(ORACLE_SO)
        lib.oracle_query_boxes.restype = C.c_int64
        lib.oracle_query_boxes.argtypes = [C.POINTER(OracleFlat), C.c_void_p, C.c_uint64, C.c_int,
                                           C.c_void_p, _u32pp, C.c_void_p]
        lib.oracle_brute_boxes.restype = C.c_int64
        lib.oracle_brute_boxes.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint32,
                                           C.c_uint32, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p,
                                           _u32pp]
        lib.oracle_query_rays.restype = C.c_int64
        lib.oracle_query_rays.argtypes = [C.POINTER(OracleFlat), C.c_void_p, C.c_uint64, C.c_int,
                                          C.c_void_p, _u32pp, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
        lib.oracle_brute_rays.restype = C.c_int64
        lib.oracle_brute_rays.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                          C.c_int, C.c_void_p, _u32pp, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
        lib.oracle_query_knn.restype = C.c_int64
        lib.oracle_query_knn.argtypes = [C.POINTER(OracleFlat), C.c_int, C.c_void_p, C.c_uint64, C.c_uint32,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_brute_knn.restype = C.c_int64
        lib.oracle_brute_knn.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p,
                                         C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        lib.oracle_free.argtypes = [C.c_void_p]
        _oracle_lib = lib
    return _oracle_lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _take_ids(free_fn, pp, total):
    if total > 0:
        ids = np.ctypeslib.as_array(pp, shape=(total,)).copy()
    else:
        ids = np.zeros(0, np.uint32)
    free_fn(C.cast(pp, C.c_void_p))
    return ids


class Flat:
    """flatten_result_t (include/RTree/rtree.hpp:852-870 of the reference) as numpy arrays."""

    def __init__(self, scalar, dim, key_is_box, max_entries, leaf_level, root, nodes, bounds,
                 children, data):
        self.scalar, self.dim, self.key_is_box, self.max_entries = scalar, dim, key_is_box, max_entries
        self.leaf_level, self.root = int(leaf_level), int(root)
        self.nodes = np.ascontiguousarray(nodes, np.uint32).reshape(-1, 3)
        self.bounds = np.ascontiguousarray(bounds, SCALAR_DTYPES[scalar]).reshape(-1, 2, dim)
        self.children = np.ascontiguousarray(children, np.uint32)
        self.data = np.ascontiguousarray(data, np.uint32)

    def c_struct(self):
        return OracleFlat(self.leaf_level, self.root, self.dim, self.scalar, len(self.nodes),
                          _ptr(self.nodes), _ptr(self.bounds), _ptr(self.children), _ptr(self.data))

    def save(self, path, **extra):
        np.savez_compressed(path, scalar=self.scalar, dim=self.dim, key_is_box=self.key_is_box,
                            max_entries=self.max_entries, leaf_level=self.leaf_level,
                            root=self.root, nodes=self.nodes, bounds=self.bounds,
                            children=self.children, data=self.data, **extra)

    @staticmethod
    def load(z):
        return Flat(int(z["scalar"]), int(z["dim"]), bool(z["key_is_box"]), int(z["max_entries"]),
                    int(z["leaf_level"]), int(z["root"]), z["nodes"], z["bounds"], z["children"],
                    z["data"])


class RefTree:
    """The reference's RTree (built by its own insert()), one of the KINDS instantiations."""

    def __init__(self, kind):
        self.kind = kind
        self.scalar, self.dim, self.key_is_box, self.max_entries = KINDS[kind]
        self.dtype = SCALAR_DTYPES[self.scalar]
        self.lib = ref_lib()
        self.h = self.lib.ref_tree_create(kind)
        if not self.h:
            raise ValueError("unknown tree kind %r" % kind)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_tree_destroy(self.h)
            self.h = None

    def insert(self, keys, first_id=0):
        keys = np.ascontiguousarray(keys, self.dtype)
        per = self.dim * (2 if self.key_is_box else 1)
        assert keys.size % per == 0
        self.lib.ref_tree_insert(self.h, _ptr(keys), keys.size // per, first_id)
        return self

    def set_keys(self, keys):
        """mutate every key in place (iteration order = flatten data order), then rebound() every
        value (reference README.md:342-347): bounds refitted, topology unchanged"""
        keys = np.ascontiguousarray(keys, self.dtype)
        per = self.dim * (2 if self.key_is_box else 1)
        got = self.lib.ref_tree_set_keys(self.h, _ptr(keys), keys.size // per)
        assert got == keys.size // per, "set_keys: key count differs from the tree's size"
        return self

    def erase_ids(self, ids):
        """the reference's erase(iterator) (include/RTree/rtree.hpp:372-457) on the value holding each id,
        in the order given; returns how many were found"""
        ids = np.ascontiguousarray(ids, np.uint32)
        return int(self.lib.ref_tree_erase_ids(self.h, _ptr(ids), ids.size))

    def leaf_level(self):
        return self.lib.ref_tree_leaf_level(self.h)

    def size(self):
        return self.lib.ref_tree_size(self.h)

    def rebalance(self):
        self.lib.ref_tree_rebalance(self.h)

    def flatten(self):
        sizes = (C.c_uint64 * 5)()
        self.lib.ref_tree_flatten(self.h, sizes)
        n_nodes, n_children, n_data, leaf_level, root = [int(x) for x in sizes]
        nodes = np.zeros((n_nodes, 3), np.uint32)
        bounds = np.zeros((n_children, 2, self.dim), self.dtype)
        children = np.zeros(n_children, np.uint32)
        data = np.zeros(n_data, np.uint32)
        self.lib.ref_tree_flat_copy(self.h, _ptr(nodes), _ptr(bounds), _ptr(children), _ptr(data))
        return Flat(self.scalar, self.dim, self.key_is_box, self.max_entries, leaf_level, root,
                    nodes, bounds, children, data)

    def search_boxes(self, queries, mode=MODE_POINT_IN_BOX, threads=0):
        """reference RTree::search() per query -> (offsets[n+1], ids sorted per query)."""
        q = np.ascontiguousarray(queries, self.dtype).reshape(-1, 2, self.dim)
        n = len(q)
        offsets = np.zeros(n + 1, np.uint64)
        pp = _u32p()
        total = self.lib.ref_search_boxes(self.h, _ptr(q), n, mode, threads, _ptr(offsets),
                                          C.byref(pp), None)
        return offsets, _take_ids(self.lib.ref_free, pp, total)

    def time_search_boxes(self, queries, mode=MODE_POINT_IN_BOX, threads=0):
        """count-only timed loop (the CPU baseline): returns (seconds, total_hits)."""
        q = np.ascontiguousarray(queries, self.dtype).reshape(-1, 2, self.dim)
        sec = C.c_double(0)
        total = self.lib.ref_search_boxes(self.h, _ptr(q), len(q), mode, threads, None, None,
                                          C.byref(sec))
        return sec.value, int(total)

    def search_rays(self, rays, mode, threads=0, timed_only=False):
        r = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
        n = len(r)
        sec = C.c_double(0)
        if mode == RAY_ALL:
            if timed_only:
                total = self.lib.ref_search_rays(self.h, _ptr(r), n, mode, threads, None, None,
                                                 None, None, None, C.byref(sec))
                return sec.value, int(total)
            offsets = np.zeros(n + 1, np.uint64)
            pp = _u32p()
            total = self.lib.ref_search_rays(self.h, _ptr(r), n, mode, threads, _ptr(offsets),
                                             C.byref(pp), None, None, None, C.byref(sec))
            return offsets, _take_ids(self.lib.ref_free, pp, total)
        if mode == RAY_CLOSEST:
            t = np.zeros(n, np.float32)
            ids = np.zeros(n, np.uint32)
            total = self.lib.ref_search_rays(self.h, _ptr(r), n, mode, threads, None, None,
                                             _ptr(t), _ptr(ids), None, C.byref(sec))
            return (sec.value, int(total)) if timed_only else (t, ids)
        anyhit = np.zeros(n, np.uint8)
        total = self.lib.ref_search_rays(self.h, _ptr(r), n, mode, threads, None, None, None, None,
                                         _ptr(anyhit), C.byref(sec))
        return (sec.value, int(total)) if timed_only else anyhit


    def search_knn(self, points, k, threads=0, timed_only=False):
        """kNN through the reference's search() (oracle/knn.h) -> (ids[n][k], d2[n][k])."""
        p = np.ascontiguousarray(points, np.float32).reshape(-1, self.dim)
        ids = np.zeros((len(p), k), np.uint32)
        d2 = np.zeros((len(p), k), np.float32)
        sec = C.c_double(0)
        rc = self.lib.ref_search_knn(self.h, _ptr(p), len(p), k, threads, _ptr(ids), _ptr(d2), C.byref(sec))
        assert rc == 0, "kNN needs a float32 tree"
        return sec.value if timed_only else (ids, d2)


def ref_max_threads():
    return ref_lib().ref_max_threads()


def oracle_query_knn(flat, points, k, want_visits=False):
    lib = oracle_lib()
    p = np.ascontiguousarray(points, np.float32).reshape(-1, flat.dim)
    ids = np.zeros((len(p), k), np.uint32)
    d2 = np.zeros((len(p), k), np.float32)
    visits = np.zeros(2, np.uint64)
    cs = flat.c_struct()
    rc = lib.oracle_query_knn(C.byref(cs), int(flat.key_is_box), _ptr(p), len(p), k, _ptr(ids), _ptr(d2),
                              _ptr(visits))
    assert rc == 0
    return (ids, d2, visits) if want_visits else (ids, d2)


def oracle_brute_knn(keys, key_is_box, key_ids, dim, points, k):
    lib = oracle_lib()
    keys = np.ascontiguousarray(keys, np.float32)
    n_keys = keys.size // (dim * (2 if key_is_box else 1))
    key_ids = None if key_ids is None else np.ascontiguousarray(key_ids, np.uint32)
    p = np.ascontiguousarray(points, np.float32).reshape(-1, dim)
    ids = np.zeros((len(p), k), np.uint32)
    d2 = np.zeros((len(p), k), np.float32)
    rc = lib.oracle_brute_knn(_ptr(keys), n_keys, int(key_is_box), _ptr(key_ids), dim, _ptr(p), len(p), k,
                              _ptr(ids), _ptr(d2))
    assert rc == 0
    return ids, d2


def oracle_query_boxes(flat, queries, mode=MODE_POINT_IN_BOX, want_visits=False):
    lib = oracle_lib()
    q = np.ascontiguousarray(queries, SCALAR_DTYPES[flat.scalar]).reshape(-1, 2, flat.dim)
    n = len(q)
    offsets = np.zeros(n + 1, np.uint64)
    visits = np.zeros(3, np.uint64)
    pp = _u32p()
    cs = flat.c_struct()
    total = lib.oracle_query_boxes(C.byref(cs), _ptr(q), n, mode, _ptr(offsets), C.byref(pp),
                                   _ptr(visits))
    assert total >= 0
    ids = _take_ids(lib.oracle_free, pp, total)
    return (offsets, ids, visits) if want_visits else (offsets, ids)


def oracle_brute_boxes(keys, key_is_box, key_ids, dim, scalar, queries, mode=MODE_POINT_IN_BOX):
    lib = oracle_lib()
    dt = SCALAR_DTYPES[scalar]
    keys = np.ascontiguousarray(keys, dt)
    n_keys = keys.size // (dim * (2 if key_is_box else 1))
    key_ids = None if key_ids is None else np.ascontiguousarray(key_ids, np.uint32)
    q = np.ascontiguousarray(queries, dt).reshape(-1, 2, dim)
    offsets = np.zeros(len(q) + 1, np.uint64)
    pp = _u32p()
    total = lib.oracle_brute_boxes(_ptr(keys), n_keys, int(key_is_box), _ptr(key_ids), dim, scalar,
                                   _ptr(q), len(q), mode, _ptr(offsets), C.byref(pp))
    assert total >= 0
    return offsets, _take_ids(lib.oracle_free, pp, total)


def oracle_query_rays(flat, rays, mode, want_visits=False):
    lib = oracle_lib()
    r = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
    n = len(r)
    visits = np.zeros(3, np.uint64)
    cs = flat.c_struct()
    if mode == RAY_ALL:
        offsets = np.zeros(n + 1, np.uint64)
        pp = _u32p()
        total = lib.oracle_query_rays(C.byref(cs), _ptr(r), n, mode, _ptr(offsets), C.byref(pp),
                                      None, None, None, _ptr(visits))
        assert total >= 0
        out = (offsets, _take_ids(lib.oracle_free, pp, total))
    elif mode == RAY_CLOSEST:
        t = np.zeros(n, np.float32)
        ids = np.zeros(n, np.uint32)
        total = lib.oracle_query_rays(C.byref(cs), _ptr(r), n, mode, None, None, _ptr(t), _ptr(ids),
                                      None, _ptr(visits))
        assert total >= 0
        out = (t, ids)
    else:
        anyhit = np.zeros(n, np.uint8)
        total = lib.oracle_query_rays(C.byref(cs), _ptr(r), n, mode, None, None, None, None,
                                      _ptr(anyhit), _ptr(visits))
        assert total >= 0
        out = (anyhit,)
    return out + (visits,) if want_visits else (out if len(out) > 1 else out[0])


def oracle_brute_rays(key_boxes, key_ids, rays, mode):
    lib = oracle_lib()
    kb = np.ascontiguousarray(key_boxes, np.float32).reshape(-1, 2, 3)
    key_ids = None if key_ids is None else np.ascontiguousarray(key_ids, np.uint32)
    r = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
    n = len(r)
    if mode == RAY_ALL:
        offsets = np.zeros(n + 1, np.uint64)
        pp = _u32p()
        total = lib.oracle_brute_rays(_ptr(kb), len(kb), _ptr(key_ids), _ptr(r), n, mode,
                                      _ptr(offsets), C.byref(pp), None, None, None)
        return offsets, _take_ids(lib.oracle_free, pp, total)
    if mode == RAY_CLOSEST:
        t = np.zeros(n, np.float32)
        ids = np.zeros(n, np.uint32)
        lib.oracle_brute_rays(_ptr(kb), len(kb), _ptr(key_ids), _ptr(r), n, mode, None, None,
                              _ptr(t), _ptr(ids), None)
        return t, ids
    anyhit = np.zeros(n, np.uint8)
    lib.oracle_brute_rays(_ptr(kb), len(kb), _ptr(key_ids), _ptr(r), n, mode, None, None, None,
                          None, _ptr(anyhit))
    return anyhit
