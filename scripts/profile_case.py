"""The command profiled with ncu: BASELINE config 2 (or a cut: PROFILE_N / PROFILE_Q), device-resident
queries, a few calls of rt_query_boxes.  Prints per-call stats so the plain run documents itself."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
rtb = importlib.import_module("r-star-tree_b200")
W, capi = rtb.workloads, rtb.capi
N = int(os.environ.get("PROFILE_N", 10_000_000)); Q = int(os.environ.get("PROFILE_Q", 16_000_000))
CALLS = int(os.environ.get("PROFILE_CALLS", 3))
t0 = time.time()
tree = rtb.RTree(4).insert(W.uniform_points(N, seed=42)).flatten_device().upload(0)
print("built + uploaded N=%d in %.1f s" % (N, time.time() - t0), flush=True)
dq = torch.from_numpy(W.cube_queries(Q, N, seed=43)).cuda()
flags = capi.RT_QUERIES_ON_DEVICE | capi.RT_RESULT_ON_DEVICE
extra = int(os.environ.get("PROFILE_FLAGS", 0))
for i in range(CALLS):
    out = tree.query_boxes_raw(dq.data_ptr(), Q, capi.RT_POINT_IN_BOX, flags | extra)
    s = tree.stats()
    print("call %d: total hits %d, reorder %.2f ms, traverse %.2f ms, finish %.2f ms, launches %d, deferred %d"
          % (i, out.total, s["reorder_ms"], s["traverse_ms"], s["finish_ms"], s["kernel_launches"], s["deferred_queries"]), flush=True)
