"""rtree-b200: B200-native batched query engine for the read-only query path of
ehwan/R-Star-Tree (box range queries and ray queries over the tree RTree::flatten() emits).

The product is C++/CUDA:

* ``include/rtree_b200.h`` + ``csrc/librtree_b200.so``   -- the C ABI (rt_upload /
  rt_query_boxes / rt_query_rays / rt_free) over hand-written sm_100a kernels;
* ``include/RTree.hpp``                                   -- the header-only host library
  (the reference's RTree API plus flatten_device()).

This Python package is only the thin driver the tests and bench.py use: ctypes
bindings (``capi``), a handle-based mirror of the host API (``host``), the synthetic
workloads of BASELINE.json (``workloads``) and the multi-GPU query partition
(``sharding``).  The directory name has dashes, so import it with
``importlib.import_module("r-star-tree_b200")``.

There is no CPU fallback: if the CUDA library is missing, or no GPU is present, the
query calls raise.
"""
from . import capi, host, sharding, workloads  # noqa: F401
from .capi import (DeviceTree, RtError, RT_POINT_IN_BOX, RT_INTERSECTS, RT_CONTAINS,  # noqa: F401
                   RT_RAY_ALL_HITS, RT_RAY_CLOSEST, RT_RAY_ANY, RT_QUERIES_ON_DEVICE,
                   RT_RESULT_ON_DEVICE, RT_UNSORTED, RT_COUNT_ONLY, RT_COLLECT_VISITS,
                   RT_NO_REORDER, RT_ANY_HIT, RT_RESULT_COUNTS32, RT_STRICT_OVERLAP, RT_INVALID_ID, RT_OPT_PIPELINE_BATCH, RT_OPT_QUERY_CHUNK,
                   RT_OPT_WORK_RANGES, RT_OPT_QUANTISED)
from .host import RTree, DeviceImage, KINDS  # noqa: F401

__all__ = ["capi", "host", "sharding", "workloads", "DeviceTree", "RtError", "RTree",
           "DeviceImage", "KINDS"]
